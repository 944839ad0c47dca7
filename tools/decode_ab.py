"""A/B of decode-step options on the 13B cached decode forward (eager launches through the C handle, programmatic
dependent launches on).  `python tools/decode_ab.py key=v0,v1,... [key=...]` times every listed value of one option at a
time against the defaults, twice and interleaved (clocks drift over a run), e.g.
`python tools/decode_ab.py gemv_no_allocate=0,1 decode_fused_attention=0,1`.  Writes gpurun_out/decode_ab.json.
(profiles/r02_decode_ab.json was written by an earlier form of this tool that also swept options since removed.)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seed_b200 import synth, lib as L

sweeps = [a.split("=") for a in sys.argv[1:] if "=" in a] or [["decode_fused_attention", "0,1"], ["gemv_no_allocate", "0,1"]]
dims = (5120, 40, 40, 13824, 40194)
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


def run(key, value):
    L.set_option(key, value)
    for i in range(4):
        llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=256 + i, last_only=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(N):
        llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=260 + i, last_only=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / N
    r = {"option": key, "value": value, "ms_per_token": round(ms, 4), "gbs": round(bytes_per_token / ms / 1e6, 1)}
    print(json.dumps(r), flush=True)
    return r


rows = []
for key, vals in sweeps:
    vals = [int(v) for v in vals.split(",")]
    default = vals[-1]
    for rep in range(2):
        for v in vals:
            rows.append(run(key, v))
    L.set_option(key, default)               # list the default last
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/decode_ab.json", "w"), indent=1)
