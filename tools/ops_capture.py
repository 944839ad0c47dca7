"""Launch every standalone kernel of the path once at its bench size (for `ncu --set full`):
LayerNorm, RMSNorm, patchify, RoPE + KV append, VQ argmin, SiLU-gate GEMM epilogue, embedding, the two tcgen05
attention kernels, and one cached decode forward of a 2-layer 13B-shaped LLaMA (GEMV, decode attention).
usage: ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r01_ops python tools/ops_capture.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L, synth
import bench

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
def r(*s, std=1.0):
    return (torch.randn(*s, device=dev, generator=g) * std).half()

# ViT LayerNorm: 256 images x 257 tokens x 1408
x = r(65792, 1408); w = r(1408); b = r(1408)
L.layernorm(x, w, b, 1e-6)
# LLaMA RMSNorm: 2048 x 4096
xr = r(2048, 4096); wr = r(4096)
L.rmsnorm(xr, wr, 1e-6)
# patchify: 256 images
img = r(256, 3, 224, 224)
L.patchify(img)
# RoPE + KV append: S=2048, H=32, D=128
qkv = r(2048, 3 * 4096)
kc = torch.zeros(1, 32, 2048, 128, device=dev, dtype=torch.float16); vc = torch.zeros_like(kc)
L.rope_kv_append(qkv, None, 1, 2048, 32, 128, 0, kc, vc)
# VQ argmin: 256 images x 32 queries against the 8192 x 32 codebook
z = r(8192, 32, std=0.26); cb = r(8192, 32, std=0.28)
L.vq_argmin(z, cb, L.VQ_FP16)
# SiLU-gate GEMM (LLaMA-7B gate/up, interleaved): M=2048, N=2*11008, K=4096
a = r(2048, 4096, std=0.5); wgu = r(2 * 11008, 4096, std=0.02)
L.gemm(a, wgu, mode=1, ctas=2)
# embedding gather: 2048 ids from a 40194 x 4096 table
tab = r(40194, 4096, std=0.02); ids = torch.randint(0, 40194, (2048,), device=dev)
L.embedding(tab, ids)
# tcgen05 attention kernels: ViT (B=64 -> 1024 items) and LLaMA causal S=2048
qkvv = r(64 * 257, 3 * 16 * 88).view(64, 257, 3, 16, 88)
q, k, v = (qkvv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
L.attention(q, k, v, 88 ** -0.5, False)
ql = r(1, 2048, 32, 128).permute(0, 2, 1, 3); kl = r(1, 32, 2048, 128); vl = r(1, 32, 2048, 128)
L.attention(ql, kl, vl, 128 ** -0.5, True)
torch.cuda.synchronize()
# cached decode forward, 2 layers of the 13B shape (GEMV with fused RMSNorm, decode attention, PDL chain)
model = bench.random_llama(dev, 0, 5120, 2, 40, 13824, 40194, 512, 2)
pids = synth.prompt_ids(1, 256, 4, seed=99).to(dev)
o = model.forward(input_ids=pids, use_cache=True, last_logits_only=True)
nxt = o.logits[:, -1].float().argmax(-1)[:, None]
model.forward(input_ids=nxt, past_key_values=o.past_key_values, use_cache=True, last_logits_only=True)
torch.cuda.synchronize()
print("ops_capture done")
