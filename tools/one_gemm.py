"""Run one GEMM shape a few times (for ncu): python tools/one_gemm.py M N K act residual ctas [bn]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L
M, N, K, act, res, ctas = [int(v) for v in sys.argv[1:7]]
bn = int(sys.argv[7]) if len(sys.argv) > 7 else 0
a = torch.randn(M, K, device="cuda", dtype=torch.float16)
w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
bias = torch.randn(N, device="cuda", dtype=torch.float16)
out = torch.randn(M, N, device="cuda", dtype=torch.float16)
for _ in range(4):
    L.gemm(a, w, bias=bias, act=act, residual=out if res else None, out=out, ctas=ctas, bn=bn)
torch.cuda.synchronize()
print("ok")
