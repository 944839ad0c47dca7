"""(historical: the gemv_occupancy / gemv_deep options it sweeps were removed after this A/B; the fused-attention rows still run)
A/B of the decode-step options on the 13B cached decode forward (eager launches through the C handle, programmatic
dependent launches on): fused RoPE+append+attention, GEMV CTAs-per-SM cap, deep-load GEMV for the short-N projections.
Writes gpurun_out/decode_ab.json.  `python tools/decode_ab.py [7b]`"""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seed_b200 import synth, lib as L

small = len(sys.argv) > 1 and sys.argv[1] == "7b"
dims = (4096, 32, 32, 11008, 40194) if small else (5120, 40, 40, 13824, 40194)
dev = torch.device("cuda", 0)
model = bench.random_llama(dev, 0, *dims, 512, 2)
ids = synth.prompt_ids(1, 256, 4, seed=99).to(dev)
o = model.forward(input_ids=ids, use_cache=True, last_logits_only=True)
nxt = o.logits[:, -1].float().argmax(-1)[:, None]
torch.cuda.synchronize()
llm = model._llm
N = 48
h, nl, _, ffn, V = dims
bytes_per_token = 2.0 * (nl * (4 * h * h + 3 * h * ffn) + V * h) + 2.0 * nl * 2 * 300 * h   # weights + ~300 cached keys


def run(tag, **opts):
    for k, v in opts.items():
        try:
            L.set_option(k, v)
        except RuntimeError:
            if v != 0:
                print(json.dumps({"tag": tag, "skipped": f"option {k} no longer exists"}), flush=True)
                return {"tag": tag, "skipped": k}
    for i in range(4):
        llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=256 + i, last_only=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(N):
        llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=260 + i, last_only=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / N
    r = dict(opts, tag=tag, ms_per_token=round(ms, 4), gbs=round(bytes_per_token / ms / 1e6, 1))
    print(json.dumps(r), flush=True)
    return r


rows = []
base = dict(decode_fused_attention=0, gemv_occupancy=0, gemv_deep=0)
rows.append(run("r02 default", **base))
rows.append(run("fused attention", **dict(base, decode_fused_attention=1)))
for occ in (3, 2):
    rows.append(run(f"fused + occupancy {occ}", **dict(base, decode_fused_attention=1, gemv_occupancy=occ)))
for deep in (2560, 3000, 8000):
    rows.append(run(f"fused + deep<={deep}", **dict(base, decode_fused_attention=1, gemv_deep=deep)))
for occ, deep in ((3, 2560), (2, 2560), (3, 8000)):
    rows.append(run(f"fused + occupancy {occ} + deep<={deep}", **dict(base, decode_fused_attention=1, gemv_occupancy=occ, gemv_deep=deep)))
rows.append(run("fused attention (repeat)", **dict(base, decode_fused_attention=1)))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/decode_ab.json", "w"), indent=1)
