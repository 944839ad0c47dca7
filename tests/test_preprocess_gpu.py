"""GPU parity of seedb200_preprocess_run (seed_b200/csrc/preprocess.cu) against the reference's own CPU
pipeline -- torchvision + Pillow, executed here on the same bytes -- BIT-EXACT on the fp16 output:
  models/transforms.py:4-19             get_transform('clip', keep_ratio=False, 224)      (PIL bilinear)
  models/seed_llama_tokenizer.py:50-56  Resize((224,224), interpolation=3) -> ... `processor` (PIL bicubic)
followed by the .half() of ImageTokenizer.encode (:84-85)."""
import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms

pytestmark = pytest.mark.gpu

MEAN, STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def reference_pipeline(interp):
    kw = {} if interp == "bilinear" else {"interpolation": 3}
    return transforms.Compose([transforms.Resize((224, 224), **kw), transforms.ToTensor(), transforms.Normalize(MEAN, STD)])


def rand_image(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if h > 4:
        img[: h // 3] = (img[: h // 3] // 128) * 255          # saturated blocks: exercises clip8 on bicubic overshoot
    return img


SIZES = [(224, 224), (480, 640), (1000, 800), (64, 48), (225, 223), (1, 1), (7, 1000), (333, 500), (1536, 2048),
         (100, 224), (449, 448), (3000, 17), (501, 5), (500, 5), (5, 3000), (2300, 1700)]


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("h,w", SIZES)
def test_preprocess_bit_exact_vs_torchvision_pillow(h, w, interp):
    from seed_b200.preprocess import GpuClipTransform

    img = rand_image(h, w, seed=h * 7 + w)
    ref = reference_pipeline(interp)(Image.fromarray(img, "RGB")).half()
    out = GpuClipTransform(224, interp)(Image.fromarray(img, "RGB"))
    torch.cuda.synchronize()
    assert out.dtype == torch.float16 and tuple(out.shape) == (3, 224, 224)
    assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16)), \
        f"{(out.cpu().float() - ref.float()).abs().max().item()} max abs diff"


def test_preprocess_batch_mixed_sizes_and_device_input():
    from seed_b200.preprocess import GpuClipTransform

    t = GpuClipTransform(224, "bicubic", max_batch=3)
    imgs = [rand_image(300, 400, 1), rand_image(300, 400, 2), rand_image(128, 96, 3), rand_image(300, 400, 4),
            rand_image(300, 400, 5), rand_image(300, 400, 6)]          # 5 of one size: crosses the plan's max_batch
    out = t([Image.fromarray(i, "RGB") for i in imgs])
    ref = torch.stack([reference_pipeline("bicubic")(Image.fromarray(i, "RGB")).half() for i in imgs])
    assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))
    # uint8 tensors that already live on the device
    dev = torch.from_numpy(np.stack([imgs[0], imgs[1]])).cuda()
    out2 = t._plan(300, 400)(dev)
    assert torch.equal(out2.cpu().view(torch.int16), ref[:2].view(torch.int16))
    with pytest.raises(ValueError):
        t(Image.fromarray(imgs[0][:, :, 0], "L"))
    with pytest.raises(ValueError):
        t._plan(300, 400)(torch.zeros(1, 10, 10, 3, dtype=torch.uint8, device="cuda"))


def test_encode_image_pil_equals_reference_processor_path():
    """SeedLlamaTokenizer.encode_image(image_pil=...) (GPU preprocessing) gives the ids of the reference flow
    `processor(image_pil)` (CPU) -> encode(image_torch)."""
    from models.seed_llama_tokenizer import ImageTokenizer
    from oracle import synth

    sd = synth.encoder_state_dict(2, 2, 0)
    tok = ImageTokenizer(sd, device="cuda", fp16=True, max_batch=4)
    pil = Image.fromarray(rand_image(375, 500, 11), "RGB")
    ids_gpu_pre = tok.encode(tok.gpu_processor(pil))
    ids_cpu_pre = tok.encode(tok.processor(pil).to("cuda"))
    assert torch.equal(ids_gpu_pre, ids_cpu_pre)
