"""Host-side mirror of models/seed_llama_tokenizer.py: ImageTokenizer and SeedLlamaTokenizer.

Signatures, attribute names, assertions and return types follow the reference (file:line cited per method) so
that scripts/seed_tokenizer_inference.py, scripts/seed_llama_inference_*.py and
MultiModalLLM/src/tools/extract_image_ids_to_torchdata_parallel.py keep working; the tensor work runs in
libseedb200 (seed_b200/qformer_quantizer.py).  Added on top of the reference surface (all optional):
`encode_image_sharded` (rank-sharded encode + one NCCL all-gather of ids) and
`image_ids_to_tokens` (direct id arithmetic instead of the '<img_%05d>' string round trip, section 8f).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from .qformer_quantizer import Blip2QformerQuantizer

WEIGHTS_NAME = "seed_quantizer.pt"
DIFFUSION_NAME = "diffusion_model"

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _make_processor(image_size: int):
    """seed_llama_tokenizer.py:50-56: Resize((s,s), bicubic) -> ToTensor -> CLIP Normalize (torchvision, CPU)."""
    from torchvision import transforms

    return transforms.Compose([
        transforms.Resize((image_size, image_size), interpolation=3),
        transforms.ToTensor(),
        transforms.Normalize(mean=CLIP_MEAN, std=CLIP_STD),
    ])


class ImageTokenizer(nn.Module):
    """seed_llama_tokenizer.py:24-113."""

    def __init__(self, model_path, diffusion_model_path=None, load_diffusion=False, image_size=224, device="cuda",
                 fp16=True, **kwargs):
        super().__init__()
        if not fp16:
            raise ValueError("seed_b200 implements the reference's fp16 mode only (fp16=True)")
        if isinstance(model_path, dict):      # in-memory state dict (tests / synthetic weights)
            model = Blip2QformerQuantizer(model_path, device=device, vit_precision="fp16", **kwargs)
        else:
            model = Blip2QformerQuantizer.from_pretrained(pretrained_model_path=model_path, device=device,
                                                          vit_precision="fp16", **kwargs)
        self.diffusion_model = None
        if diffusion_model_path is not None and load_diffusion:
            # OUT OF SCOPE for the kernels: the unCLIP UNet/VAE stays the reference's diffusers
            # pipeline when that package is installed; this mirror only feeds it the 1024-d embedding.
            try:
                from diffusers import StableUnCLIPImg2ImgPipeline  # type: ignore
            except Exception as e:  # pragma: no cover - diffusers is not in this image
                raise RuntimeError("load_diffusion=True needs the `diffusers` package (not installed here)") from e
            self.diffusion_model = StableUnCLIPImg2ImgPipeline.from_pretrained(
                diffusion_model_path, torch_dtype=torch.float16).to(device)
        self.processor = _make_processor(image_size)          # the reference's CPU pipeline (attribute kept)
        # its device-side twin: same bytes in, bit-identical fp16 tensor out (seed_b200/csrc/preprocess.cu)
        from .preprocess import GpuClipTransform

        self.gpu_processor = GpuClipTransform(image_size, "bicubic", device=device)
        # fixed latents / noise for the diffusion decoder (seed_llama_tokenizer.py:61-65)
        self.latents = torch.randn(torch.Size([1, 4, 96, 96]), generator=None, device=device, dtype=torch.float16)
        self.noise = torch.randn(torch.Size([1, 1024]), generator=None, device=device, dtype=torch.float16)
        self.model = model
        self.device = device
        self.fp16 = fp16

    def __len__(self):
        return self.model.n_embed

    def encode(self, image_torch):
        """Convert a batch of img to code (seed_llama_tokenizer.py:75-90): [b,c,h,w] (or [c,h,w]) -> LongTensor [b,32]."""
        if len(image_torch.shape) == 3:
            image_torch = image_torch.unsqueeze(0)
        img = image_torch
        if self.fp16:
            img = img.half()
        with torch.no_grad():
            id = self.model.encode_ids(img)
        return id.view(img.shape[0], -1)

    def decode_embeds(self, indices):
        """ids -> [B,1024] fp16 embedding fed to the unCLIP decoder (the in-scope part of decode)."""
        return self.model.get_codebook_entry(indices)

    def decode(self, indices, negative_indices=None, guidance_scale=10, num_inference_steps=20):
        """seed_llama_tokenizer.py:92-113."""
        image_embeds = self.model.get_codebook_entry(indices)
        if negative_indices is not None:
            assert indices.shape == negative_indices.shape, "Negative indices must have the same shape with indices"
            negative_image_embeds = self.model.get_codebook_entry(negative_indices)
        else:
            negative_image_embeds = None
        if self.diffusion_model is None:
            raise RuntimeError("decode() needs the Stable-unCLIP pipeline (load_diffusion=True with `diffusers` "
                               "installed); use decode_embeds() for the 1024-d image embedding")
        image = self.diffusion_model(
            image_embeds=image_embeds,
            negative_image_embeds=negative_image_embeds,
            guidance_scale=guidance_scale,
            noise_level=0,
            num_inference_steps=num_inference_steps,
            latents=self.latents,
        ).images
        return image


class SeedImageTokenMixin:
    """The image half of SeedLlamaTokenizer (seed_llama_tokenizer.py:144-213), independent of the text vocab."""

    def _init_image_side(self, device="cuda", fp16=True, load_diffusion=False, encoder_url=None, diffusion_path=None,
                         image_tokenizer_kwargs: Optional[Dict[str, Any]] = None):
        self.device = device
        self.fp16 = fp16
        self.load_diffusion = load_diffusion
        self.encoder_url = encoder_url
        self.diffusion_path = diffusion_path
        self._image_tokenizer_kwargs = dict(image_tokenizer_kwargs or {})

    def _model_path(self):
        if self.encoder_url is not None:
            return self.encoder_url
        assert hasattr(self, "name_or_path") and os.path.exists(self.name_or_path)
        return os.path.join(self.name_or_path, WEIGHTS_NAME)

    def load_image_tokenizer(self):
        if not hasattr(self, "_image_tokenizer"):
            self._image_tokenizer = ImageTokenizer(model_path=self._model_path(),
                                                   diffusion_model_path=self.diffusion_path,
                                                   load_diffusion=self.load_diffusion, device=self.device,
                                                   fp16=self.fp16, **self._image_tokenizer_kwargs)

    @property
    def image_tokenizer(self):
        self.load_image_tokenizer()
        return self._image_tokenizer

    @property
    def num_image_tokens(self):
        return 8192  # seed_llama_tokenizer.py:181-183

    def to(self, device):
        self.device = device
        if hasattr(self, "_image_tokenizer"):
            self._image_tokenizer.to(device=device)

    def encode_image(self, image_path=None, image_pil=None, image_torch=None, image_size: int = 224):
        """seed_llama_tokenizer.py:185-202: exactly one of the three inputs."""
        assert (image_path is None) + (image_pil is None) + (image_torch is None) == 2
        if image_path is not None:
            from PIL import Image

            image_pil = Image.open(image_path).convert("RGB")
        if image_pil is not None:
            # same arithmetic as `self.image_tokenizer.processor(image_pil)` (Pillow bicubic resize, ToTensor,
            # Normalize, then the .half() of encode), computed by seedb200_preprocess_run on the raw bytes
            image_torch = self.image_tokenizer.gpu_processor(image_pil)
        return self.image_tokenizer.encode(image_torch)

    def decode_image(self, indices, negative_indices=None, guidance_scale=10):
        """seed_llama_tokenizer.py:204-213."""
        indices = indices.to(self.device)
        if negative_indices is not None:
            negative_indices = negative_indices.to(self.device)
        return self.image_tokenizer.decode(indices, negative_indices=negative_indices, guidance_scale=guidance_scale)

    # ---- additions (not in the reference) ------------------------------------------------------
    def encode_image_sharded(self, image_torch, group=None):
        """Data-parallel encode: every rank tokenizes its own [B_local,3,224,224] shard and ONE NCCL all-gather
        returns the ids of all ranks, rank-major, on every rank (replaces the per-rank tar shards
        of MultiModalLLM/src/tools/extract_image_ids_to_torchdata_parallel.py:109-127).  int32 on the wire."""
        import torch.distributed as dist

        ids = self.image_tokenizer.encode(image_torch)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return ids
        return all_gather_ids(ids, group)

    @staticmethod
    def image_ids_to_tokens(ids, image_id_shift: int = 32000, boi: Optional[int] = None, eoi: Optional[int] = None,
                            out: Optional[torch.Tensor] = None):
        """[B,32] codebook ids -> [B,34] LLaMA token ids `<img> <img_xxxxx>*32 </img>` by arithmetic
        (scripts/seed_llama_inference_8B.py:16-23,60,98-100 build the same ids through a string round trip).
        CUDA ids go through seedb200_image_ids_to_tokens (the span can land inside a prompt buffer via `out`);
        host ids are host-side bookkeeping like the reference's string formatting.  The defaults assume the
        added-token order of the released checkpoints (`<img_00000>` = 32000 ... `<img>` = 40192, `</img>` = 40193);
        `image_token_ids()` on a tokenizer instance resolves the three ids through the vocabulary instead."""
        boi = image_id_shift + 8192 if boi is None else boi
        eoi = image_id_shift + 8193 if eoi is None else eoi
        if ids.is_cuda:
            from . import lib as L

            return L.image_ids_to_tokens(ids, image_id_shift, boi, eoi, out=out)
        B = ids.shape[0]
        res = torch.empty((B, 34), dtype=torch.int64, device=ids.device) if out is None else out
        res[:, 0] = boi
        res[:, 1:33] = ids + image_id_shift
        res[:, 33] = eoi
        return res

    def image_token_ids(self):
        """(image_id_shift, boi, eoi) looked up in the text vocabulary the way the reference scripts do
        (`tokenizer(BOI_TOKEN).input_ids[0]`, scripts/seed_llama_inference_8B.py:42-43), cached; falls back to the
        released layout only when this object carries no vocabulary (bare SeedImageTokenMixin)."""
        cached = getattr(self, "_image_token_ids", None)
        if cached is not None:
            return cached
        shift, boi, eoi = 32000, 40192, 40193
        conv = getattr(self, "convert_tokens_to_ids", None)
        if callable(conv):
            unk = getattr(self, "unk_token_id", None)
            got = [conv(t) for t in ("<img_00000>", "<img>", "</img>")]
            if any(g is None or g == unk for g in got):
                raise KeyError("the tokenizer vocabulary lacks '<img_00000>', '<img>' or '</img>' "
                               "(added tokens of the SEED-LLaMA checkpoints)")
            shift, boi, eoi = (int(g) for g in got)
            last = conv("<img_%05d>" % (self.num_image_tokens - 1))
            if last != shift + self.num_image_tokens - 1:
                raise ValueError("image tokens are not contiguous in the vocabulary: id arithmetic does not apply")
        self._image_token_ids = (shift, boi, eoi)
        return self._image_token_ids

    def encode_image_tokens(self, image_torch, out: Optional[torch.Tensor] = None):
        """images -> [B,34] LLaMA token ids without leaving the device (encode fused with the id arithmetic)."""
        shift, boi, eoi = self.image_token_ids()
        tok = self.image_tokenizer
        if len(image_torch.shape) == 3:
            image_torch = image_torch.unsqueeze(0)
        return tok.model.encode_tokens(image_torch, shift, boi, eoi, out=out)


def all_gather_ids(ids: torch.Tensor, group=None) -> torch.Tensor:
    """the single collective on the path: all-gather [B_local,32] ids (int32 on the wire) -> [world*B_local,32] int64."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    local = ids.to(torch.int32).contiguous()
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out.to(torch.int64)


try:  # transformers' LlamaTokenizer needs a sentencepiece vocab that only real checkpoints ship
    from transformers import LlamaTokenizer as _LlamaTokenizer
except Exception:  # pragma: no cover
    _LlamaTokenizer = object


class SeedLlamaTokenizer(SeedImageTokenMixin, _LlamaTokenizer):
    """seed_llama_tokenizer.py:116-213: LlamaTokenizer + image tokenizer."""

    def __init__(self, vocab_file=None, unk_token="<unk>", bos_token="<s>", eos_token="</s>", pad_token=None,
                 sp_model_kwargs: Optional[Dict[str, Any]] = None, add_bos_token=True, add_eos_token=False,
                 clean_up_tokenization_spaces=False, device="cuda", fp16=True, load_diffusion=False,
                 encoder_url=None, diffusion_path=None, **kwargs):
        super().__init__(vocab_file=vocab_file, unk_token=unk_token, bos_token=bos_token, eos_token=eos_token,
                         pad_token=pad_token, sp_model_kwargs=sp_model_kwargs, add_bos_token=add_bos_token,
                         add_eos_token=add_eos_token, clean_up_tokenization_spaces=clean_up_tokenization_spaces,
                         **kwargs)
        self._init_image_side(device, fp16, load_diffusion, encoder_url, diffusion_path)
        self.pad_token = self.unk_token
        self.load_image_tokenizer()
