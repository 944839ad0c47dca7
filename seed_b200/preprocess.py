"""GPU image preprocessing with the reference's exact arithmetic.

`GpuClipTransform` is the device-side twin of the two CPU pipelines the reference uses in front of the tokenizer:

* `models.transforms.get_transform('clip', keep_ratio=False, 224)` (models/transforms.py:4-19) -- PIL bilinear;
* `ImageTokenizer.processor` (models/seed_llama_tokenizer.py:50-56) -- PIL bicubic (`interpolation=3`).

It takes PIL images / uint8 HWC arrays, uploads the raw bytes, and runs `seedb200_preprocess_run`
(seed_b200/csrc/preprocess.cu): Pillow's 8-bit two-pass fixed-point resize, ToTensor, Normalize and the fp16 cast
of `ImageTokenizer.encode`, bit for bit.  Plans (weight tables + intermediate buffer) are cached per image size.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple, Union

import numpy as np
import torch

from . import lib as L


class GpuClipTransform:
    def __init__(self, image_size: int = 224, interpolation: Union[str, int] = "bilinear", device="cuda",
                 max_batch: int = 256):
        if interpolation not in L.Preprocess.FILTERS:
            raise ValueError("interpolation must be 'bilinear' (PIL 2) or 'bicubic' (PIL 3)")
        self.image_size, self.filter, self.max_batch = image_size, interpolation, max_batch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuClipTransform needs a CUDA device (use models.transforms.get_transform on the CPU)")
        self._plans: Dict[Tuple[int, int], L.Preprocess] = {}

    def _plan(self, h: int, w: int) -> L.Preprocess:
        p = self._plans.get((h, w))
        if p is None:
            p = self._plans[(h, w)] = L.Preprocess(h, w, self.image_size, self.filter, self.max_batch)
        return p

    @staticmethod
    def _to_u8(img) -> torch.Tensor:
        if isinstance(img, torch.Tensor):
            t = img
        else:
            if hasattr(img, "mode") and img.mode != "RGB":     # the reference scripts do .convert('RGB') themselves
                raise ValueError(f"expected an RGB image, got mode {img.mode}")
            t = torch.from_numpy(np.array(img, dtype=np.uint8, copy=True))
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError(f"expected uint8 [H,W,3], got {t.dtype} {tuple(t.shape)}")
        return t

    def __call__(self, images) -> torch.Tensor:
        """one image -> [3,S,S]; a sequence of images -> [n,3,S,S] (sizes may differ: grouped per size)."""
        single = not isinstance(images, (list, tuple))
        imgs: List[torch.Tensor] = [self._to_u8(i) for i in ([images] if single else images)]
        out = torch.empty((len(imgs), 3, self.image_size, self.image_size), dtype=torch.float16, device=self.device)
        groups: Dict[Tuple[int, int], List[int]] = {}
        for i, t in enumerate(imgs):
            groups.setdefault((int(t.shape[0]), int(t.shape[1])), []).append(i)
        for (h, w), idx in groups.items():
            batch = torch.stack([imgs[i] for i in idx])
            if not batch.is_cuda:
                batch = batch.pin_memory().to(self.device, non_blocking=True)
            res = self._plan(h, w)(batch)
            out[torch.as_tensor(idx, device=self.device)] = res
        return out[0] if single else out
