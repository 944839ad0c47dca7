"""Per-role clock64 timeline of block 0 of a tcgen05 ViT attention kernel, plus the kernel time at the bench shape.

    python tools/attn_timeline.py [variant]        variant 1 (default): attention_tc.cu, 2: attention_tc2.cu
"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, N, D = 16, 257, 88
lib = L.load()


def make(B):
    qkv = torch.randn(B * N, 3 * H * D, device="cuda", dtype=torch.float16)
    v4 = qkv.view(B, N, 3, H, D)
    return [v4[:, :, i].permute(0, 2, 1, 3) for i in range(3)]


# kernel time at the bench shape (B = 256: 4096 items over 148 persistent CTAs), both variants
q, k, v = make(256)
for var, tma in ((1, 1), (1, 0), (2, 0)):
    L.set_option("vit_attention_tc", var)
    L.set_option("vit_attention_tma", tma)
    for _ in range(3):
        L.attention(q, k, v, D ** -0.5, False)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        s.record(); L.attention(q, k, v, D ** -0.5, False); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    print(f"variant {var} (Q/K by {'TMA' if (tma or var == 2) else 'cp.async'}): B=256 kernel time min {ts[0]:.4f} ms  median {ts[len(ts)//2]:.4f} ms "
          f"({4.0 * 256 * H * N * N * D / ts[0] / 1e9:.0f} TFLOP/s)")

L.set_option("vit_attention_tc", variant)
L.set_option("vit_attention_tma", 1)
q, k, v = make(64)
dbg = torch.zeros(16 * 64 * 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    L.attention(q, k, v, D ** -0.5, False)
lib.seedb200_debug_set_attn_timeline.argtypes = [C.c_void_p]
lib.seedb200_debug_set_attn_timeline(dbg.data_ptr())
L.attention(q, k, v, D ** -0.5, False)
torch.cuda.synchronize()
lib.seedb200_debug_set_attn_timeline(None)
L.set_option("vit_attention_tc", 1)
t = dbg.cpu().view(16, 64, 8)
if variant == 2:
    names = {0: "softmax w0", 4: "softmax w4", 8: "mma", 9: "row256", 10: "loader 0", 13: "loader 3"}
    ev = {8: ["-", "S0 issue", "S1 issue", "-", "-", "PV0 issue", "PV1 issue"],
          0: ["start", "q,k ok", "dots done", "S ok", "max done", "P done", "O ok", "end"],
          9: ["start", "q,k ok", "key256", "scores ok", "p written", "row stored"],
          10: ["start", "Q0 landed", "V landed", "K landed", "Q1 landed"]}
    ev[4] = ev[0]; ev[13] = ev[10]
    t0 = int(t[0, 0, 0])
else:
    names = {0: "softmax w0", 4: "softmax w4", 8: "mma", 9: "row256", 10: "ld Q0", 11: "ld K", 12: "ld Q1", 13: "ld V",
             14: "helper w8", 15: "helper w9"}
    ev = {8: ["start", "S0 issue", "S1 issue", "wait v", "v ok", "PV0 issue", "PV1 issue", "end"],
          0: ["start", "-", "-", "S ok", "max done", "P done", "O ok", "end"],
          10: ["start", "empty ok", "-", "issued/full"],
          9: ["start", "q,k ok", "key256", "scores ok", "p written", "share done"],
          14: ["start", "q,k ok", "dots done", "p,v ok", "share done"]}
    ev[4] = ev[0]; ev[11] = ev[12] = ev[13] = ev[10]; ev[15] = ev[14]
    t0 = int(t[8, 0, 0])
print(f"--- timeline of block 0, variant {variant} (cycles since the first stamp)")
for item in range(0, 6):
    print(f"--- item {item}")
    for slot in sorted(names):
        row = [int(x) - t0 for x in t[slot, item, :len(ev[slot])]]
        print(f"{names[slot]:11s} " + "  ".join(f"{e}={x}" for e, x in zip(ev[slot], row) if e != "-"))
