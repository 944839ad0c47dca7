"""`models.transforms.get_transform` -- the hydra target of configs/transform/clip_transform.yaml
(reference models/transforms.py:4-19).  Same contract: a callable PIL image -> float32 [3, S, S] tensor, resized
(optionally aspect preserving + centre crop) and normalised with the CLIP statistics.  `get_gpu_transform` is the
device-side twin (both `keep_ratio` settings, bit-identical output, seed_b200/csrc/preprocess.cu)."""
from torchvision import transforms as _T

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _geometry(image_size: int, keep_ratio: bool):
    """keep_ratio: shorter side to `image_size`, then a centre crop; otherwise a plain (anisotropic) resize"""
    if keep_ratio:
        return [_T.Resize(image_size), _T.CenterCrop(image_size)]
    return [_T.Resize((image_size, image_size))]


def _clip(image_size: int, keep_ratio: bool):
    return _T.Compose(_geometry(image_size, keep_ratio) + [_T.ToTensor(), _T.Normalize(mean=CLIP_MEAN, std=CLIP_STD)])


_BUILDERS = {"clip": _clip}


def get_transform(type="clip", keep_ratio=True, image_size=224):
    builder = _BUILDERS.get(type)
    if builder is None:
        raise NotImplementedError
    return builder(image_size, keep_ratio)


def get_gpu_transform(type="clip", keep_ratio=True, image_size=224, device="cuda", **kwargs):
    """PIL image(s) / uint8 HWC arrays -> fp16 [n, 3, S, S] on `device`, equal bit for bit to
    `get_transform(type, keep_ratio, image_size)(img).half()` (Pillow bilinear resize [+ centre crop] + ToTensor +
    Normalize); same defaults as get_transform (reference models/transforms.py:4: keep_ratio=True)."""
    if type != "clip":
        raise NotImplementedError
    from seed_b200.preprocess import GpuClipTransform

    return GpuClipTransform(image_size, "bilinear", device=device, keep_ratio=keep_ratio, **kwargs)
