"""Dump the per-role clock64 timeline of block 0 of the tcgen05 ViT attention kernel (debug)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L
B = 64
H, N, D = 16, 257, 88
qkv = torch.randn(B * N, 3 * H * D, device="cuda", dtype=torch.float16)
v4 = qkv.view(B, N, 3, H, D)
q, k, v = (v4[:, :, i].permute(0, 2, 1, 3) for i in range(3))
dbg = torch.zeros(16 * 64 * 8, dtype=torch.int64, device="cuda")
lib = L.load()
for _ in range(2):
    L.attention(q, k, v, D ** -0.5, False)
lib.seedb200_debug_set_attn_timeline.argtypes = [C.c_void_p]
lib.seedb200_debug_set_attn_timeline(dbg.data_ptr())
L.attention(q, k, v, D ** -0.5, False)
torch.cuda.synchronize()
lib.seedb200_debug_set_attn_timeline(None)
t = dbg.cpu().view(16, 64, 8)
t0 = int(t[8, 0, 0])
names = {0: "softmax w0", 4: "softmax w4", 8: "mma", 9: "row256", 10: "ld Q0", 11: "ld K", 12: "ld Q1", 13: "ld V"}
ev = {8: ["start", "S0 issue", "S1 issue", "wait v", "v ok", "PV0 issue", "PV1 issue", "end"],
      0: ["start", "q,k ok", "s256 done", "S ok", "max done", "P done", "O ok", "end"],
      10: ["start", "empty ok", "issued", "full"]}
ev[9] = ["start", "q,k ok", "key256", "scores ok", "p written"]; ev[4] = ev[0]; ev[11] = ev[12] = ev[13] = ev[10]
for item in range(0, 6):
    print(f"--- item {item}")
    for slot in (8, 9, 0, 4, 10, 11, 12, 13):
        row = [int(x) - t0 for x in t[slot, item, :len(ev[slot])]]
        print(f"{names[slot]:11s} " + "  ".join(f"{e}={x}" for e, x in zip(ev[slot], row)))
