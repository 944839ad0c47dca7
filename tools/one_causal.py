"""Time the LLaMA-7B prefill attention shape (B=1, H=32, S, D=128, causal): tcgen05 kernel with TMA / cp.async loaders,
mma.sync kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
H, D = 32, 128
q = torch.randn(1, S, H, D, device="cuda", dtype=torch.float16).permute(0, 2, 1, 3)
k = torch.randn(1, H, S, D, device="cuda", dtype=torch.float16)
v = torch.randn(1, H, S, D, device="cuda", dtype=torch.float16)
def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
outs = {}
rows = []
for tc, tma in ((1, 1), (1, 0), (0, 0)):
    L.set_option("causal_attention_tc", tc)
    L.set_option("causal_attention_tma", tma)
    o = L.attention(q, k, v, D ** -0.5, True); torch.cuda.synchronize()
    outs[(tc, tma)] = o.float()
    t = timeit(lambda: L.attention(q, k, v, D ** -0.5, True))
    rows.append({"tc": tc, "tma": tma, "S": S, "ms": round(t, 4), "tflops_causal": round(2.0 * H * S * S * D * 2 / 2 / t / 1e9, 1)})
    print(json.dumps(rows[-1]), flush=True)
L.set_option("causal_attention_tc", 1); L.set_option("causal_attention_tma", 1)
print("rel diff tc vs mma:", ((outs[(1, 1)] - outs[(0, 0)]).norm() / outs[(0, 0)].norm()).item(),
      "tma == cp.async:", bool(torch.equal(outs[(1, 1)], outs[(1, 0)])))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/one_causal_{S}.json", "w"), indent=1)
