"""The four ViT-g GEMM shapes exactly as the encode step launches them (M = 65 792 = 256 images x 257 tokens):
qkv and fc1 with the LayerNorm-folded epilogue reading the residual stream itself, proj and fc2 with bias + in-place
residual, BN = 256 CTA-pair tiles + narrow tail tile.  Warm-up round, then one launch per shape (for `ncu --set full`):

    ncu --set full --clock-control none -k regex:gemm_tcgen05 --launch-skip 4 -c 4 -f -o gpurun_out/r02_gemm python tools/vit_gemm_capture.py
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L

M, D, FF = 65792, 1408, 6144
g = torch.Generator(device="cuda").manual_seed(1)
def r(*s, std=1.0):
    return (torch.randn(*s, device="cuda", generator=g) * std).half()

x = r(M, D)
gamma, beta = (1.0 + 0.1 * torch.randn(D, device="cuda")).half(), r(D, std=0.02)
w_qkv, b_qkv = r(3 * D, D, std=D ** -0.5), r(3 * D, std=0.02)
w_fc1, b_fc1 = r(FF, D, std=D ** -0.5), r(FF, std=0.02)
w_proj, b_proj = r(D, D, std=D ** -0.5), r(D, std=0.02)
w_fc2, b_fc2 = r(D, FF, std=FF ** -0.5), r(D, std=0.02)
qkv_f = L.ln_fold_weights(w_qkv, gamma, beta, b_qkv)
fc1_f = L.ln_fold_weights(w_fc1, gamma, beta, b_fc1)
stats = L.row_stats(x, 1e-6)
qkv = torch.empty(M, 3 * D, device="cuda", dtype=torch.float16)
hid = torch.empty(M, FF, device="cuda", dtype=torch.float16)
att = r(M, D)

def run():
    L.gemm(x, qkv_f[0], out=qkv, ctas=2, ln=(stats, qkv_f[1], qkv_f[2]))                       # norm1 + qkv
    L.gemm(att, w_proj, bias=b_proj, residual=x, out=x, ctas=2)                                # proj + residual
    L.gemm(x, fc1_f[0], act=L.ACT_GELU, out=hid, ctas=2, ln=(stats, fc1_f[1], fc1_f[2]))       # norm2 + fc1 + GELU
    L.gemm(hid, w_fc2, bias=b_fc2, residual=x, out=x, ctas=2)                                  # fc2 + residual

run(); torch.cuda.synchronize()
run(); torch.cuda.synchronize()
# event timing of the same four launches (not under ncu: compare)
names = ["qkv_4224x1408_lnfold", "proj_1408x1408", "fc1_6144x1408_lnfold_gelu", "fc2_1408x6144"]
flops = [2.0 * M * 3 * D * D, 2.0 * M * D * D, 2.0 * M * FF * D, 2.0 * M * D * FF]
if "--time" in sys.argv:
    fns = [lambda: L.gemm(x, qkv_f[0], out=qkv, ctas=2, ln=(stats, qkv_f[1], qkv_f[2])),
           lambda: L.gemm(att, w_proj, bias=b_proj, residual=x, out=x, ctas=2),
           lambda: L.gemm(x, fc1_f[0], act=L.ACT_GELU, out=hid, ctas=2, ln=(stats, fc1_f[1], fc1_f[2])),
           lambda: L.gemm(hid, w_fc2, bias=b_fc2, residual=x, out=x, ctas=2)]
    for nm, fl, fn in zip(names, flops, fns):
        ts = []
        for _ in range(6):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        print(json.dumps({"shape": nm, "ms": round(min(ts), 4), "tflops": round(fl / min(ts) / 1e9, 1)}))
print("vit_gemm_capture done")
if "--ab" in sys.argv:
    # LayerNorm-folded epilogue vs plain bias epilogue on the same shapes, interleaved (same clocks)
    ln = L.layernorm(x, gamma, beta, 1e-6)
    pairs = [("qkv", lambda: L.gemm(x, qkv_f[0], out=qkv, ctas=2, ln=(stats, qkv_f[1], qkv_f[2])),
              lambda: L.gemm(ln, w_qkv, bias=b_qkv, out=qkv, ctas=2), flops[0]),
             ("fc1", lambda: L.gemm(x, fc1_f[0], act=L.ACT_GELU, out=hid, ctas=2, ln=(stats, fc1_f[1], fc1_f[2])),
              lambda: L.gemm(ln, w_fc1, bias=b_fc1, act=L.ACT_GELU, out=hid, ctas=2), flops[2])]
    for nm, f_ln, f_bias, fl in pairs:
        t = {"ln": [], "bias": []}
        for _ in range(8):
            for key, fn in (("ln", f_ln), ("bias", f_bias)):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); fn(); e.record(); torch.cuda.synchronize(); t[key].append(s.elapsed_time(e))
        print(json.dumps({"shape": nm, "lnfold_ms": round(min(t["ln"]), 4), "bias_ms": round(min(t["bias"]), 4),
                          "lnfold_tflops": round(fl / min(t["ln"]) / 1e9, 1), "bias_tflops": round(fl / min(t["bias"]) / 1e9, 1)}))

if "--ksub" in sys.argv:
    # pipeline-stage depth per shape (option gemm_ksub: 0 = heuristic, 1 = 64-deep stages, 2 = 128-deep), interleaved
    fns = [lambda: L.gemm(x, qkv_f[0], out=qkv, ctas=2, ln=(stats, qkv_f[1], qkv_f[2])),
           lambda: L.gemm(att, w_proj, bias=b_proj, residual=x, out=x, ctas=2),
           lambda: L.gemm(x, fc1_f[0], act=L.ACT_GELU, out=hid, ctas=2, ln=(stats, fc1_f[1], fc1_f[2])),
           lambda: L.gemm(hid, w_fc2, bias=b_fc2, residual=x, out=x, ctas=2)]
    rows = []
    for nm, fl, fn in zip(names, flops, fns):
        t = {0: [], 1: [], 2: []}
        for _ in range(8):
            for ks in (0, 1, 2):
                L.set_option("gemm_ksub", ks)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); fn(); e.record(); torch.cuda.synchronize(); t[ks].append(s.elapsed_time(e))
        L.set_option("gemm_ksub", 0)
        rows.append({"shape": nm, **{f"ksub{ks}_ms": round(min(v), 4) for ks, v in t.items()},
                     **{f"ksub{ks}_tflops": round(fl / min(v) / 1e9, 1) for ks, v in t.items()}})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/vit_gemm_ksub.json", "w"), indent=1)

if "--act" in sys.argv:
    # is the fc1 GEMM bound by its GELU epilogue?  same shape, LN-folded, with and without the activation, interleaved
    f_gelu = lambda: L.gemm(x, fc1_f[0], act=L.ACT_GELU, out=hid, ctas=2, ln=(stats, fc1_f[1], fc1_f[2]))
    f_none = lambda: L.gemm(x, fc1_f[0], act=L.ACT_NONE, out=hid, ctas=2, ln=(stats, fc1_f[1], fc1_f[2]))
    f_relu = lambda: L.gemm(x, fc1_f[0], act=L.ACT_RELU, out=hid, ctas=2, ln=(stats, fc1_f[1], fc1_f[2]))
    t = {"gelu": [], "none": [], "relu": []}
    for _ in range(10):
        for key, fn in (("gelu", f_gelu), ("none", f_none), ("relu", f_relu)):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); t[key].append(s.elapsed_time(e))
    row = {"shape": "fc1_6144x1408_lnfold", **{k + "_ms": round(min(v), 4) for k, v in t.items()},
           **{k + "_tflops": round(flops[2] / min(v) / 1e9, 1) for k, v in t.items()}}
    print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(row, open("gpurun_out/vit_gemm_act.json", "w"), indent=1)

if "--bn" in sys.argv:
    # tile width per shape with the current pipeline (the tuning table went stale once: re-measure), interleaved
    def mk(bn):
        return [lambda: L.gemm(x, qkv_f[0], out=qkv, ctas=2, bn=bn, ln=(stats, qkv_f[1], qkv_f[2])),
                lambda: L.gemm(att, w_proj, bias=b_proj, residual=x, out=x, ctas=2, bn=bn),
                lambda: L.gemm(x, fc1_f[0], act=L.ACT_GELU, out=hid, ctas=2, bn=bn, ln=(stats, fc1_f[1], fc1_f[2])),
                lambda: L.gemm(hid, w_fc2, bias=b_fc2, residual=x, out=x, ctas=2, bn=bn)]
    widths = (0, 256, 192, 128)
    fns = {bn: mk(bn) for bn in widths}
    rows = []
    for i, (nm, fl) in enumerate(zip(names, flops)):
        t = {bn: [] for bn in widths}
        for _ in range(8):
            for bn in widths:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); fns[bn][i](); e.record(); torch.cuda.synchronize(); t[bn].append(s.elapsed_time(e))
        rows.append({"shape": nm, **{f"bn{k}_ms": round(min(v), 4) for k, v in t.items()},
                     **{f"bn{k}_tflops": round(fl / min(v) / 1e9, 1) for k, v in t.items()}})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/vit_gemm_bn.json", "w"), indent=1)
