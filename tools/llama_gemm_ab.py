"""A/B the tile width for the LLaMA-7B prefill shapes (M=2048)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L
def timeit(fn, iters=20):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]
M = 2048
for name, N, K in (("qkv", 12288, 4096), ("o", 4096, 4096), ("down", 4096, 11008), ("lm_head", 40194, 4096)):
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for bn in (0, 256, 192, 128):
        try:
            t = timeit(lambda: L.gemm(a, w, out=out, ctas=2, bn=bn))
            print(json.dumps({"shape": name, "bn": bn, "ms": round(t, 4), "tflops": round(2.0 * M * N * K / t / 1e9, 1)}), flush=True)
        except Exception as ex:
            print(name, bn, "ERR", str(ex)[:80])
