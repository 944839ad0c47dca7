"""Run the ViT attention shape a few times (for ncu): python tools/one_attn.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, N, D = 16, 257, 88
qkv = torch.randn(B * N, 3 * H * D, device="cuda", dtype=torch.float16)
v4 = qkv.view(B, N, 3, H, D)
q, k, v = (v4[:, :, i].permute(0, 2, 1, 3) for i in range(3))
for _ in range(3):
    o = L.attention(q, k, v, D ** -0.5, False)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    o = L.attention(q, k, v, D ** -0.5, False)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print(f"B={B} attention {ms:.3f} ms  {4*B*H*N*N*D/ms/1e9:.1f} TFLOP/s")
