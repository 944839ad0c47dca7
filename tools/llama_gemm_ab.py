"""A/B of the tile plan for the LLaMA prefill GEMM shapes: 7B at M = 2048 (config #3) and 13B at M = 256 (the prompt of
config #5).  `auto0` = fixed heuristics (option gemm_sched 0), `auto1` = load-balance model (seedb200_gemm_plan);
explicit widths run in rotated round-robin order (`rr`) and in balanced-tail order (`bt`).
Writes gpurun_out/llama_gemm_ab.json."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


SHAPES = [("7b_qkv", 2048, 12288, 4096), ("7b_o", 2048, 4096, 4096), ("7b_down", 2048, 4096, 11008),
          ("13b_qkv_m256", 256, 15360, 5120), ("13b_o_m256", 256, 5120, 5120), ("13b_down_m256", 256, 5120, 13824)]
if len(sys.argv) > 1:
    SHAPES = [s for s in SHAPES if any(a in s[0] for a in sys.argv[1:])]
rows = []
for name, M, N, K in SHAPES:
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    variants = [("auto0", 0, 0, 2), ("auto1", 1, 0, 2)]
    for bn in (256, 224, 192, 128):
        variants += [(f"bn{bn}_rr", 0, bn, 2), (f"bn{bn}_bt", 2, bn, 2)]
    if M <= 256:
        variants += [(f"bn{bn}_c1", 0, bn, 1) for bn in (256, 128, 64)]
    for tag, sched, bn, ctas in variants:
        L.set_option("gemm_sched", sched)
        try:
            t = timeit(lambda: L.gemm(a, w, out=out, ctas=ctas, bn=bn))
            r = {"shape": name, "M": M, "N": N, "K": K, "variant": tag, "ms": round(t, 4),
                 "tflops": round(2.0 * M * N * K / t / 1e9, 1), "gbs": round(2.0 * (N * K + M * K + M * N) / t / 1e6, 1)}
        except Exception as ex:
            r = {"shape": name, "variant": tag, "error": str(ex)[:100]}
        finally:
            L.set_option("gemm_sched", 1)
        rows.append(r)
        print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/llama_gemm_ab.json", "w"), indent=1)
