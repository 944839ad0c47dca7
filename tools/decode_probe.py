"""Where does a cached decode forward spend its time?  C-level loop vs Python mirror, 13B by default."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seed_b200 import synth
small = len(sys.argv) > 1 and sys.argv[1] == "7b"
dims = (4096, 32, 32, 11008, 40194) if small else (5120, 40, 40, 13824, 40194)
dev = torch.device("cuda", 0)
model = bench.random_llama(dev, 0, *dims, 512, 2)
ids = synth.prompt_ids(1, 256, 4, seed=99).to(dev)
o = model.forward(input_ids=ids, use_cache=True, last_logits_only=True)
nxt = o.logits[:, -1].float().argmax(-1)[:, None]
past = o.past_key_values
torch.cuda.synchronize()
def ev():
    return torch.cuda.Event(enable_timing=True)
N = 32
# (a) python mirror
e0, e1 = ev(), ev(); t0 = time.perf_counter(); e0.record()
for _ in range(N):
    o = model.forward(input_ids=nxt, past_key_values=past, use_cache=True, last_logits_only=True); past = o.past_key_values
e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
print(json.dumps({"path": "python forward", "gpu_ms_per_token": e0.elapsed_time(e1) / N, "host_enqueue_ms_per_token": 1e3 * t_host / N}))
# (b) C handle only, A/B over the decode options
from seed_b200 import lib as L
llm = model._llm
pl = 256 + N
for pdl, per_sm in ((1, 0), (1, 1), (1, 2), (1, 4), (0, 0)):
    L.set_option("decode_pdl", pdl); L.set_option("gemv_ksplit", per_sm)
    for i in range(4):
        llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=pl + i, last_only=True)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    for i in range(N):
        llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=pl + i, last_only=True)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"path": "C handle forward", "pdl": pdl, "gemv_ksplit": per_sm, "gpu_ms_per_token": round(e0.elapsed_time(e1) / N, 4)}), flush=True)
L.set_option("decode_pdl", 1); L.set_option("gemv_ksplit", 0)
# (c) host cost of one forward with an empty queue
torch.cuda.synchronize()
t0 = time.perf_counter()
llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=300, last_only=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(json.dumps({"path": "single forward, empty queue", "host_enqueue_ms": 1e3 * (t1 - t0), "until_done_ms": 1e3 * (t2 - t0)}))
# (d) real timeline from CUPTI (torch.profiler): kernel durations and the gaps between consecutive kernels
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        llm.forward(input_ids=nxt, inputs_embeds=None, position_ids=None, past_len=310 + i, last_only=True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
import collections, re
agg = collections.OrderedDict(); gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    gaps.append(b.time_range.start - a.time_range.end)
for e in evs:
    n = re.sub(r"\(.*$", "", e.name.replace("void ", "").replace("sb::", ""))
    d = agg.setdefault(n, [0, 0.0]); d[0] += 1; d[1] += e.time_range.end - e.time_range.start
tot = sum(v[1] for v in agg.values())
span = evs[-1].time_range.end - evs[0].time_range.start
print(json.dumps({"kernels": len(evs), "sum_kernel_us": tot, "span_us": span, "gap_us_mean": sum(gaps) / len(gaps),
                  "gap_us_max": max(gaps)}))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n={n:4d} avg={us / n:8.2f} us total={us:9.1f}")
