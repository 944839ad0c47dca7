"""Time the tcgen05 GEMM on the shapes of the encode / prefill path (CUDA events, L2-cold operands are
larger than L2 at B=256) next to torch.matmul (cuBLAS) as the yardstick.  Not part of bench.py."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L

SHAPES = [  # name, M, N, K, act, residual
    ("vit_qkv", 65792, 4224, 1408, 0, False),
    ("vit_proj", 65792, 1408, 1408, 0, True),
    ("vit_fc1", 65792, 6144, 1408, 1, False),
    ("vit_fc2", 65792, 1408, 6144, 0, True),
    ("llama_qkv", 2048, 12288, 4096, 0, False),
    ("llama_down", 2048, 4096, 11008, 0, True),
    ("square8k", 8192, 8192, 8192, 0, False),
]

def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts), sorted(ts)[len(ts)//2]

def main():
    only = sys.argv[1:] 
    for name, M, N, K, act, res in SHAPES:
        if only and name not in only: continue
        a = torch.randn(M, K, device="cuda", dtype=torch.float16)
        w = torch.randn(N, K, device="cuda", dtype=torch.float16) * K ** -0.5
        bias = torch.randn(N, device="cuda", dtype=torch.float16)
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        r = out if res else None
        flops = 2.0 * M * N * K
        row = {"shape": name, "M": M, "N": N, "K": K}
        tb, tm = timeit(lambda: torch.matmul(a, w.t(), out=out))
        row["cublas_tflops"] = round(flops / tb / 1e9, 1)
        for ctas in (1, 2):
            for bn in ([0] if name != "square8k" else [256, 128]):
                try:
                    tb, tm = timeit(lambda: L.gemm(a, w, bias=bias, act=act, residual=r, out=out, ctas=ctas, bn=bn))
                    row[f"sb_ctas{ctas}_bn{bn}_tflops"] = round(flops / tb / 1e9, 1)
                    row[f"sb_ctas{ctas}_bn{bn}_ms"] = round(tb, 3)
                except Exception as ex:
                    row[f"sb_ctas{ctas}_bn{bn}_err"] = str(ex)[:200]
        print(json.dumps(row), flush=True)

if __name__ == "__main__":
    main()
