"""A/B of the ragged-tail n-tile (option gemm_tail) on the ViT / LLaMA shapes with N % 256 != 0 (CUDA events, best of 6)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L

def timeit(fn, iters=6):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)

for name, M, N, K, res, bn in (("proj", 65792, 1408, 1408, True, 0), ("fc2", 65792, 1408, 6144, True, 0),
                               ("qkv_bn192", 65792, 4224, 1408, False, 192), ("qkv_bn256", 65792, 4224, 1408, False, 256),
                               ("lm_head", 2048, 40194, 4096, False, 0)):
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
    bias = torch.randn(N, device="cuda", dtype=torch.float16)
    ld = (N + 7) // 8 * 8
    out = torch.empty(M, ld, device="cuda", dtype=torch.float16)[:, :N]
    r = out if res else None
    row = {"shape": name, "M": M, "N": N, "K": K}
    for tail in (0, 1):
        L.set_option("gemm_tail", tail)
        t = timeit(lambda: L.gemm(a, w, bias=bias, residual=r, out=out, ctas=2, bn=bn))
        row[f"tail{tail}_ms"] = round(t, 4)
        row[f"tail{tail}_tflops"] = round(2.0 * M * N * K / t / 1e9, 1)
    print(json.dumps(row), flush=True)
L.set_option("gemm_tail", 1)
