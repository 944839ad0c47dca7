"""Sweep tile-N / stage depth for the narrow-N ViT GEMMs (proj, fc2)."""
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
for name, M, N, K in (("proj", 65792, 1408, 1408), ("fc2", 65792, 1408, 6144), ("qkv", 65792, 4224, 1408)):
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
    bias = torch.randn(N, device="cuda", dtype=torch.float16)
    out = torch.randn(M, N, device="cuda", dtype=torch.float16)
    res = out if name != "qkv" else None
    for bn in ((128, 176, 256) if name != "qkv" else (192, 256)):
        for ctas in (1, 2):
            for ksub in (1, 2):
                L.set_option("gemm_ksub", ksub)
                try:
                    t = timeit(lambda: L.gemm(a, w, bias=bias, residual=res, out=out, ctas=ctas, bn=bn))
                    print(json.dumps({"shape": name, "bn": bn, "ctas": ctas, "ksub": ksub, "ms": round(t, 4), "tflops": round(2.0 * M * N * K / t / 1e9, 1)}), flush=True)
                except Exception as ex:
                    print(name, bn, ctas, ksub, "ERR", str(ex)[:100])
L.set_option("gemm_ksub", 0)
