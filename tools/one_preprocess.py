"""The two preprocessing kernels once at the bench shape (256 images of 480 x 640 -> 224 x 224 bicubic), for
`ncu --set full -k regex:resize_ python tools/one_preprocess.py`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seed_b200 import lib as L
u8 = torch.randint(0, 256, (256, 480, 640, 3), dtype=torch.uint8, device="cuda")
plan = L.Preprocess(480, 640, 224, "bicubic", max_batch=256)
plan(u8); torch.cuda.synchronize()
plan(u8); torch.cuda.synchronize()
print("one_preprocess done")
