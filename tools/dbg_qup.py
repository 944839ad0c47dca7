import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import synth, lib as L
from models.seed_qformer.qformer_quantizer import Blip2QformerQuantizer
vd, ql, dd = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctas = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sd = synth.encoder_state_dict(vd, ql, dd)
m = Blip2QformerQuantizer(sd, device="cuda", max_batch=4, gemm_ctas=ctas, vq_mode=1)
x = synth.images(2).cuda()
for rep in range(3):
    ids, qup = m.get_codebook_indices(x)
    torch.cuda.synchronize()
    cb = sd["quantize.embedding.weight"].cuda()
    quant = cb[ids.reshape(-1)]
    h = torch.tanh(quant @ sd["decode_task_layer.0.weight"].cuda().t() + sd["decode_task_layer.0.bias"].cuda())
    ref = h @ sd["decode_task_layer.2.weight"].cuda().t() + sd["decode_task_layer.2.bias"].cuda()
    d = (qup.float().reshape(-1, 768) - ref)
    per_row = d.norm(dim=1) / ref.norm(dim=1)
    print(f"depth {vd}/{ql}/{dd} ctas {ctas} rep {rep}: rel {d.norm()/ref.norm():.4f} bad rows {(per_row > 0.01).nonzero().flatten().tolist()[:20]} cols of row0 bad: {(d[0].abs() > 0.05).nonzero().flatten().tolist()[:12]}")
