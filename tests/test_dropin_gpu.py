"""Script-level drop-in proof: the reference's entry scripts, run against checkpoints on disk through
the same dotted `_target_` paths their hydra configs name -- with this repo's `models/` package answering.

What is built under tmp_path (no network: real checkpoints cannot be fetched, so tiny synthetic ones are written in
the formats the reference loads):
  * `seed_quantizer.pt`      torch.save of a depth-1 tokenizer state dict (qformer_quantizer.py:366-374 torch.load path)
  * `tokenizer.model`        a sentencepiece BPE model trained on the spot (+ tokenizer_config.json), the 8192 + 2
                             image tokens added the way the released SEED-LLaMA tokenizers carry them
  * `llama/`                 HF LLaMA directory: config.json + model.safetensors (model_tools.py:5-18 from_pretrained)

The body of each test follows the script it cites line by line; `instantiate` is the 10-line hydra.utils.instantiate
stand-in (hydra is not installed here): resolve `_target_`, call it with the remaining keys + overrides.
"""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R, synth

pytestmark = pytest.mark.gpu

BOI_TOKEN, EOI_TOKEN, IMG_TOKEN = "<img>", "</img>", "<img_{:05d}>"        # scripts/seed_llama_inference_8B.py:16-18
NUM_IMG_CODES = 8192


def instantiate(cfg: dict, **overrides):
    """hydra.utils.instantiate for the flat configs under configs/{tokenizer,transform,llm}/*.yaml"""
    cfg = dict(cfg)
    cfg.update(overrides)
    target = cfg.pop("_target_")
    obj = importlib.import_module(target.split(".")[0])
    parts = target.split(".")
    for i in range(1, len(parts)):
        try:
            obj = getattr(obj, parts[i])
        except AttributeError:
            obj = importlib.import_module(".".join(parts[: i + 1]))
    return obj(**cfg)


@pytest.fixture(scope="module")
def checkpoints(tmp_path_factory):
    import sentencepiece as spm
    from safetensors.torch import save_file

    root = tmp_path_factory.mktemp("seed_ckpt")
    tok_dir = root / "seed-tokenizer-2"
    tok_dir.mkdir()
    # ---- image tokenizer weights: reference names, depth 1/1/1 ----
    enc_sd = synth.encoder_state_dict(1, 1, 1, seed=77)
    torch.save({k: v.half() for k, v in enc_sd.items()}, tok_dir / "seed_quantizer.pt")
    # ---- text vocabulary ----
    corpus = root / "corpus.txt"
    corpus.write_text("\n".join(["USER: what is this animal ? ASSISTANT: a cat on the green grass .",
                                 "can you generate an image of a dog on the green grass ?",
                                 "the quick brown fox jumps over the lazy dog"] * 40))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tok_dir / "tokenizer"), vocab_size=64,
                                   model_type="bpe", unk_id=0, bos_id=1, eos_id=2, pad_id=-1, character_coverage=1.0,
                                   minloglevel=2)
    json.dump({"tokenizer_class": "SeedLlamaTokenizer", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>",
               "add_bos_token": True, "add_eos_token": False, "model_max_length": 2048},
              open(tok_dir / "tokenizer_config.json", "w"))
    # ---- LLaMA: HF directory ----
    text_vocab = 64
    vocab = text_vocab + NUM_IMG_CODES + 2
    h, nl, nh, ffn = 512, 2, 4, 1408
    llm_dir = root / "seed_llama_tiny"
    llm_dir.mkdir()
    llm_sd = synth.llama_state_dict(h, nl, ffn, vocab, seed=78)
    save_file({k: v.half().contiguous() for k, v in llm_sd.items()}, str(llm_dir / "model.safetensors"))
    json.dump({"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": h, "intermediate_size": ffn,
               "num_hidden_layers": nl, "num_attention_heads": nh, "num_key_value_heads": nh, "vocab_size": vocab,
               "rms_norm_eps": 1e-6, "max_position_embeddings": 512, "bos_token_id": 1, "eos_token_id": 2,
               "torch_dtype": "float16"}, open(llm_dir / "config.json", "w"))
    return {"tok_dir": str(tok_dir), "llm_dir": str(llm_dir), "enc_sd": enc_sd, "llm_sd": llm_sd, "text_vocab": text_vocab,
            "vocab": vocab, "dims": (h, nl, nh, ffn)}


def _tokenizer(ck, **overrides):
    # configs/tokenizer/seed_llama_tokenizer_hf.yaml (encoder_url: a local file instead of the https URL; the
    # reference resolves `name_or_path/seed_quantizer.pt` when encoder_url is None, seed_llama_tokenizer.py:147-151)
    cfg = {"_target_": "models.seed_llama_tokenizer.SeedLlamaTokenizer.from_pretrained",
           "pretrained_model_name_or_path": ck["tok_dir"], "fp16": True, "load_diffusion": False, "encoder_url": None,
           "diffusion_path": None}
    tok = instantiate(cfg, **overrides)
    # the released tokenizers carry the image vocabulary as added tokens: <img_00000> .. <img_08191>, <img>, </img>
    tok.add_tokens([IMG_TOKEN.format(i) for i in range(NUM_IMG_CODES)] + [BOI_TOKEN, EOI_TOKEN], special_tokens=False)
    return tok


def test_seed_tokenizer_inference_script(checkpoints):
    """scripts/seed_tokenizer_inference.py:20-29 up to the ids (the unCLIP image decode needs diffusers)."""
    from PIL import Image

    ck = checkpoints
    device = "cuda"
    tokenizer = _tokenizer(ck, device=device)                                       # :20-21
    transform = instantiate({"_target_": "models.transforms.get_transform", "type": "clip", "image_size": 224,
                             "keep_ratio": False})                                  # :23-24, clip_transform.yaml
    rng = np.random.default_rng(5)
    image = Image.fromarray(rng.integers(0, 256, (300, 400, 3), dtype=np.uint8), "RGB")   # stands in for images/cat.jpg
    image_tensor = transform(image).to(device)                                      # :28
    image_ids = tokenizer.encode_image(image_torch=image_tensor)                    # :29
    assert image_ids.dtype == torch.int64 and tuple(image_ids.shape) == (1, 32) and image_ids.is_cuda
    assert tokenizer.num_image_tokens == 8192 and len(tokenizer.image_tokenizer) == 8192
    # same ids as the dict-constructed model (no file round trip) and as the CPU oracle above the margin
    from models.seed_qformer.qformer_quantizer import Blip2QformerQuantizer

    direct = Blip2QformerQuantizer(ck["enc_sd"], device=device, max_batch=2)
    assert torch.equal(direct.encode_ids(image_tensor[None]), image_ids)
    with torch.no_grad():
        ref = R.encode(image_tensor[None].half().float().cpu(), ck["enc_sd"], 1, 1)
    neq = image_ids.cpu() != ref["ids"]
    assert not bool((neq.reshape(-1) & (ref["margin"] > 0.02)).any())
    # PIL and path inputs run the bicubic `processor` (seed_llama_tokenizer.py:50-56,195-200) on the device,
    # bit-identical to the CPU pipeline the reference applies
    via_pil = tokenizer.encode_image(image_pil=image)
    cpu_proc = tokenizer.image_tokenizer.processor(image).to(device)
    assert torch.equal(via_pil, tokenizer.encode_image(image_torch=cpu_proc))
    # decode side up to the 1024-d embedding (the unCLIP pipeline is out of scope)
    emb = tokenizer.image_tokenizer.decode_embeds(image_ids)
    assert tuple(emb.shape) == (1, 1024)
    with pytest.raises(RuntimeError, match="unCLIP"):
        tokenizer.decode_image(image_ids)


def test_seed_llama_inference_script(checkpoints):
    """scripts/seed_llama_inference_8B.py:26-38,66-100: tokenizer + transform + model from their configs, image ->
    ids -> '<img>...</img>' string -> tokenizer -> model.generate -> decode."""
    from PIL import Image

    ck = checkpoints
    device = "cuda"
    tokenizer = _tokenizer(ck, device=device)                                                             # :68-70
    transform = instantiate({"_target_": "models.transforms.get_transform", "type": "clip", "image_size": 224,
                             "keep_ratio": False})                                                        # :72-74
    model = instantiate({"_target_": "models.model_tools.get_pretrained_llama_causal_model",
                         "pretrained_model_name_or_path": ck["llm_dir"], "torch_dtype": "fp16",
                         "low_cpu_mem_usage": True}, torch_dtype=torch.float16)                          # :76-77
    model = model.eval().to(device)                                                                       # :78
    generation_config = {"temperature": 1.0, "num_beams": 1, "max_new_tokens": 24, "top_p": 0.5, "do_sample": True}

    rng = np.random.default_rng(6)
    image = Image.fromarray(rng.integers(0, 256, (256, 320, 3), dtype=np.uint8), "RGB")
    image_tensor = transform(image).to(device)                                                            # :94-96
    img_ids = tokenizer.encode_image(image_torch=image_tensor)                                            # :97
    img_ids_np = img_ids.view(-1).cpu().numpy()                                                           # :98
    img_tokens = BOI_TOKEN + "".join([IMG_TOKEN.format(item) for item in img_ids_np]) + EOI_TOKEN         # :99
    question = "what is this animal ?"
    input_tokens = tokenizer.bos_token + "USER:" + " " + img_tokens + question + "\n" + "ASSISTANT:"      # :103
    input_ids = tokenizer(input_tokens, add_special_tokens=False, return_tensors="pt").input_ids.to(device)   # :28-29
    # (1) the string round trip lands on the ids the device arithmetic produces
    shift, boi, eoi = tokenizer.image_token_ids()
    assert (shift, boi, eoi) == (ck["text_vocab"], ck["text_vocab"] + 8192, ck["text_vocab"] + 8193)
    span = tokenizer.encode_image_tokens(image_tensor)                               # [1,34], never leaves the GPU
    pos = int((input_ids[0] == boi).nonzero()[0])
    assert torch.equal(input_ids[0, pos:pos + 34], span[0])
    assert torch.equal(span[0, 1:33] - shift, img_ids[0])
    # (2) generate through the reference call pattern; same tokens as the dict-constructed model
    generate_ids = model.generate(input_ids=input_ids, seed=11, **generation_config)                      # :31-34
    assert generate_ids.shape[0] == 1 and torch.equal(generate_ids[:, :input_ids.shape[1]], input_ids)
    new_ids = generate_ids[0][input_ids.shape[1]:]                                                        # :35
    assert 1 <= new_ids.numel() <= 24 and int(new_ids.max()) < ck["vocab"]
    from transformers.models.llama.configuration_llama import LlamaConfig
    from models.llama_xformer import LlamaForCausalLM

    h, nl, nh, ffn = ck["dims"]
    cfg = LlamaConfig(vocab_size=ck["vocab"], hidden_size=h, intermediate_size=ffn, num_hidden_layers=nl,
                      num_attention_heads=nh, num_key_value_heads=nh, rms_norm_eps=1e-6, max_position_embeddings=512,
                      eos_token_id=2)
    direct = LlamaForCausalLM(cfg, ck["llm_sd"], device=device, max_batch=1, max_seq=512)
    same = direct.generate(input_ids=input_ids, seed=11, **generation_config)
    assert torch.equal(same, generate_ids)
    # logits of the loaded checkpoint vs the CPU oracle on the loaded weights
    with torch.no_grad():
        ref_logits, _, _ = R.llama_forward(ck["llm_sd"], input_ids.cpu(), nh, nl)
    out = model(input_ids=input_ids)
    err = ((out.logits.float().cpu() - ref_logits).norm() / ref_logits.norm()).item()
    assert err <= 1e-2, err
    # (3) decode_image_text (:40-63): text ids decode; an image span in the output maps back to codebook ids
    text = tokenizer.decode(new_ids[new_ids < shift], skip_special_tokens=True)
    assert isinstance(text, str)
    fake = torch.cat([torch.tensor([boi], device=device), img_ids[0] + shift, torch.tensor([eoi], device=device)])
    boi_list = torch.where(fake == tokenizer(BOI_TOKEN, add_special_tokens=False).input_ids[0])[0]        # :42
    eoi_list = torch.where(fake == tokenizer(EOI_TOKEN, add_special_tokens=False).input_ids[0])[0]        # :43
    back = (fake[boi_list[0] + 1:eoi_list[0]] - shift).reshape(1, -1)                                     # :60
    assert torch.equal(back, img_ids)


def test_loader_errors_and_dtype_strings(checkpoints, tmp_path):
    from models.model_tools import get_pretrained_llama_causal_model

    ck = checkpoints
    with pytest.raises(RuntimeError, match="local checkpoint"):
        get_pretrained_llama_causal_model(pretrained_model_name_or_path=str(tmp_path / "missing"), torch_dtype="fp16")
    with pytest.raises(ValueError, match="fp16"):
        get_pretrained_llama_causal_model(pretrained_model_name_or_path=ck["llm_dir"], torch_dtype="bf16")
    m = get_pretrained_llama_causal_model(pretrained_model_name_or_path=ck["llm_dir"], torch_dtype="float16",
                                          low_cpu_mem_usage=True)
    assert m.config.vocab_size == ck["vocab"]
    # strict=False semantics of the tokenizer loader: unknown keys ignored, a missing hot-path key is an error
    from models.seed_qformer.qformer_quantizer import Blip2QformerQuantizer

    sd = {k: v.half() for k, v in ck["enc_sd"].items()}
    sd["some.unrelated.key"] = torch.zeros(3)
    p = tmp_path / "extra.pt"
    torch.save(sd, p)
    Blip2QformerQuantizer.from_pretrained(pretrained_model_path=str(p), device="cuda", max_batch=1)
    del sd["ln_vision.weight"]
    p2 = tmp_path / "broken.pt"
    torch.save(sd, p2)
    with pytest.raises(RuntimeError, match="ln_vision.weight"):
        Blip2QformerQuantizer.from_pretrained(pretrained_model_path=str(p2), device="cuda", max_batch=1)
