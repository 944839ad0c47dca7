"""GPU image preprocessing with the reference's exact arithmetic.

`GpuClipTransform` is the device-side twin of the two CPU pipelines the reference uses in front of the tokenizer:

* `models.transforms.get_transform('clip', keep_ratio=False, 224)` (models/transforms.py:4-19) -- PIL bilinear;
* `ImageTokenizer.processor` (models/seed_llama_tokenizer.py:50-56) -- PIL bicubic (`interpolation=3`).

It takes PIL images / uint8 HWC arrays, uploads the raw bytes, and runs `seedb200_preprocess_run`
(seed_b200/csrc/preprocess.cu): Pillow's 8-bit two-pass fixed-point resize, ToTensor, Normalize and the fp16 cast
of `ImageTokenizer.encode`, bit for bit.  Plans (weight tables + intermediate buffer) live in a bounded LRU keyed by
source size; `keep_ratio=True` evaluates Resize(S) -> CenterCrop(S) (the reference default) as a windowed resample.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import numpy as np
import torch

from . import lib as L


def keep_ratio_geometry(h: int, w: int, size: int):
    """torchvision's arithmetic for Resize(size) -> CenterCrop(size) (models/transforms.py:6-9, keep_ratio=True):
    the shorter side becomes `size`, the other int(size * long / short) (transforms.functional
    ._compute_resized_output_size); the crop origin is int(round((dim - size) / 2.0)) with Python's round
    (half to even) as in functional.center_crop.  -> ((resize_h, resize_w), (crop_top, crop_left))"""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    rw, rh = (new_short, new_long) if w <= h else (new_long, new_short)
    top = int(round((rh - size) / 2.0))
    left = int(round((rw - size) / 2.0))
    return (rh, rw), (top, left)


class GpuClipTransform:
    """PIL images / uint8 HWC arrays -> fp16 [n,3,S,S] on the device.  One plan (weight tables + an intermediate
    sized for `max_batch` images) per distinct source size, kept in a small LRU: a dataset with heterogeneous image
    sizes (the reference extractor's use case) must not accumulate one plan per size it has ever seen."""

    def __init__(self, image_size: int = 224, interpolation: Union[str, int] = "bilinear", device="cuda",
                 max_batch: int = 16, keep_ratio: bool = False, max_plans: int = 8):
        if interpolation not in L.Preprocess.FILTERS:
            raise ValueError("interpolation must be 'bilinear' (PIL 2) or 'bicubic' (PIL 3)")
        self.image_size, self.filter, self.max_batch = image_size, interpolation, max_batch
        self.keep_ratio, self.max_plans = keep_ratio, max(1, int(max_plans))
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuClipTransform needs a CUDA device (use models.transforms.get_transform on the CPU)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._plans: "OrderedDict[Tuple[int, int], L.Preprocess]" = OrderedDict()

    def _plan(self, h: int, w: int) -> L.Preprocess:
        p = self._plans.get((h, w))
        if p is not None:
            self._plans.move_to_end((h, w))
            return p
        while len(self._plans) >= self.max_plans:          # evict the least recently used plan and free its buffers
            _, old = self._plans.popitem(last=False)
            torch.cuda.current_stream(self.device).synchronize()    # its last run() may still be in flight
            old.close()
        if self.keep_ratio:
            resize, crop = keep_ratio_geometry(h, w, self.image_size)
            if min(h, w) < 1:
                raise ValueError("empty image")
        else:
            resize, crop = None, (0, 0)
        p = self._plans[(h, w)] = L.Preprocess(h, w, self.image_size, self.filter, self.max_batch, resize=resize,
                                               crop=crop, device=self.device)
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
