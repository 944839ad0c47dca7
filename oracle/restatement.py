"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch, fp32) of the reference algorithm on the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
It travels to the GPU box (which has no /root/reference) and is pinned against the real reference code by
tests/test_oracle.py: here in the build container against the live reference through oracle/ref_shim.py,
and everywhere against tests/golden/*.pt, which were generated from the reference by oracle/make_golden.py.

Each function follows the reference file:line it cites and takes the reference's own state-dict.  The
reference's CPU-runnable configuration is fp32 (`fp16=False`; its fp16 mode only runs under CUDA autocast,
SURVEY.md section 8c), so this is the fp32 oracle against which activation tolerances are stated.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------------
# EVA ViT-g/14  (models/seed_qformer/eva_vit.py)
# --------------------------------------------------------------------------------------------------
def vit_block(x: torch.Tensor, sd: SD, p: str, heads: int = 16) -> torch.Tensor:
    """Block.forward (eva_vit.py:199-206) with gamma_1 None: x + attn(norm1(x)); x + mlp(norm2(x))."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    # Attention.forward (eva_vit.py:129-159): bias = (q_bias, 0, v_bias); q scaled before q k^T
    qkv_bias = torch.cat((sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.v_bias"]), sd[p + "attn.v_bias"]))
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], qkv_bias)
    qkv = qkv.reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (C // heads) ** -0.5
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    a = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    x = x + F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    # Mlp.forward (eva_vit.py:59-66): fc1 -> GELU(erf) -> fc2
    h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                 sd[p + "mlp.fc2.bias"])
    return x + h


def vit_forward_features(image: torch.Tensor, sd: SD, depth: int) -> torch.Tensor:
    """VisionTransformer.forward_features (eva_vit.py:369-385) + PatchEmbed.forward (:223-230)."""
    pre = "visual_encoder."
    x = F.conv2d(image, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"], stride=14)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[pre + "cls_token"].expand(x.shape[0], -1, -1)
    x = torch.cat((cls, x), dim=1) + sd[pre + "pos_embed"]
    for i in range(depth):
        x = vit_block(x, sd, f"{pre}blocks.{i}.")
    return x


# --------------------------------------------------------------------------------------------------
# causal Q-Former  (models/seed_qformer/qformer_causual.py)
# --------------------------------------------------------------------------------------------------
def _bert_attention(hidden, kv_src, sd: SD, p: str, mask: Optional[torch.Tensor], heads: int = 12):
    """BertSelfAttention.forward (:148-241) + BertSelfOutput.forward (:251-255)."""
    B, N, C = hidden.shape

    def split(t):
        return t.view(t.shape[0], t.shape[1], heads, C // heads).permute(0, 2, 1, 3)

    k = split(F.linear(kv_src, sd[p + "self.key.weight"], sd[p + "self.key.bias"]))
    v = split(F.linear(kv_src, sd[p + "self.value.weight"], sd[p + "self.value.bias"]))
    q = split(F.linear(hidden, sd[p + "self.query.weight"], sd[p + "self.query.bias"]))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(C // heads)
    if mask is not None:
        scores = scores + mask
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(B, N, C)
    out = F.linear(ctx, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return F.layer_norm(out + hidden, (C,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)


def qformer_forward(image_embeds: torch.Tensor, sd: SD, layers: int) -> torch.Tensor:
    """BertModel.forward with query_embeds only (:769-931): 32 queries => is_casual (:813-816), additive mask
    (1 - tril) * -10000 (:712-714,765-767); BertLayer.forward (:359-444) self-attn -> cross-attn on even
    layers (encoder mask all ones => 0) -> feed_forward_chunk_query (:441-444)."""
    B = image_embeds.shape[0]
    pre = "Qformer.bert."
    q = sd["query_tokens"].expand(B, -1, -1)
    h = F.layer_norm(q, (768,), sd[pre + "embeddings.LayerNorm.weight"], sd[pre + "embeddings.LayerNorm.bias"], 1e-12)
    n = h.shape[1]
    ids = torch.arange(n, device=h.device)
    causal = (ids[None, :] <= ids[:, None]).to(h.dtype)
    mask = ((1.0 - causal) * -10000.0)[None, None]
    for l in range(layers):
        p = f"{pre}encoder.layer.{l}."
        h = _bert_attention(h, h, sd, p + "attention.", mask)
        if l % 2 == 0:
            h = _bert_attention(h, image_embeds, sd, p + "crossattention.", None)
        inter = F.gelu(F.linear(h, sd[p + "intermediate_query.dense.weight"], sd[p + "intermediate_query.dense.bias"]))
        out = F.linear(inter, sd[p + "output_query.dense.weight"], sd[p + "output_query.dense.bias"])
        h = F.layer_norm(out + h, (768,), sd[p + "output_query.LayerNorm.weight"], sd[p + "output_query.LayerNorm.bias"],
                         1e-12)
    return h


# --------------------------------------------------------------------------------------------------
# Blip2QformerQuantizer  (models/seed_qformer/qformer_quantizer.py)
# --------------------------------------------------------------------------------------------------
def vq_forward(z: torch.Tensor, codebook: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """VectorQuantizer2.forward (:83-123), eval path: the distance expression verbatim, argmin, margins."""
    zf = z.reshape(-1, z.shape[-1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) - 2 * torch.einsum(
        "bd,dn->bn", zf, codebook.t())
    ids = torch.argmin(d, dim=1)
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    return ids, (top2[:, 1] - top2[:, 0])


def encode(image: torch.Tensor, sd: SD, vit_depth: int, qformer_layers: int) -> Dict[str, torch.Tensor]:
    """get_codebook_indices (:288-307) in the reference's CPU mode (no autocast, fp32)."""
    vit = vit_forward_features(image, sd, vit_depth)
    image_embeds = F.layer_norm(vit, (1408,), sd["ln_vision.weight"], sd["ln_vision.bias"], 1e-5)   # blip2.py:179-184
    qout = qformer_forward(image_embeds, sd, qformer_layers)
    # encode_task_layer (:219-223): Linear - Tanh - Linear
    z = F.linear(torch.tanh(F.linear(qout, sd["encode_task_layer.0.weight"], sd["encode_task_layer.0.bias"])),
                 sd["encode_task_layer.2.weight"], sd["encode_task_layer.2.bias"])
    ids, margin = vq_forward(z, sd["quantize.embedding.weight"])
    quant = F.embedding(ids, sd["quantize.embedding.weight"]).view(z.shape)
    up = F.linear(torch.tanh(F.linear(quant, sd["decode_task_layer.0.weight"], sd["decode_task_layer.0.bias"])),
                  sd["decode_task_layer.2.weight"], sd["decode_task_layer.2.bias"])
    B = image.shape[0]
    return {"ids": ids.view(B, -1), "margin": margin.view(B, -1), "z": z, "vit": vit, "image_embeds": image_embeds,
            "qformer": qout, "query_output_up": up}


def _timm_block(x: torch.Tensor, sd: SD, p: str, heads: int = 12) -> torch.Tensor:
    """vit.Block.forward (vit.py:147-150) / vit.Attention.forward (:84-104): scale applied after q k^T."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, N, 3, heads, C // heads)
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    attn = ((q @ k.transpose(-2, -1)) * (C // heads) ** -0.5).softmax(dim=-1)
    a = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                 sd[p + "mlp.fc2.bias"])
    return x + h


def detokenize(ids: torch.Tensor, sd: SD, detok_depth: int) -> torch.Tensor:
    """get_codebook_entry (:309-338), use_qformer_image False branch: -> [B,1024]."""
    quant = F.embedding(ids, sd["quantize.embedding.weight"])
    up = F.linear(torch.tanh(F.linear(quant, sd["decode_task_layer.0.weight"], sd["decode_task_layer.0.bias"])),
                  sd["decode_task_layer.2.weight"], sd["decode_task_layer.2.bias"])
    x = up + sd["pos_embed_image"].repeat(up.shape[0], 1, 1)
    for i in range(detok_depth):
        x = _timm_block(x, sd, f"blocks_image.{i}.")
    r = F.linear(F.relu(F.linear(F.relu(F.linear(x, sd["image_down.0.weight"])), sd["image_down.2.weight"])),
                 sd["image_down.4.weight"])
    r = r.reshape(r.shape[0], -1)
    return F.linear(r, sd["distill_image_proj.weight"], sd["distill_image_proj.bias"])


# --------------------------------------------------------------------------------------------------
# LLaMA  (models/llama_xformer.py)
# --------------------------------------------------------------------------------------------------
def _rms(x, w, eps):
    """LlamaRMSNorm.forward (:105-113)."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def _rope(q, k, position_ids, base=10000.0):
    """LlamaRotaryEmbedding (:116-150) + apply_rotary_pos_emb / rotate_half (:152-168)."""
    D = q.shape[-1]
    inv_freq = 1.0 / (base ** (torch.arange(0, D, 2).float() / D))
    n = int(position_ids.max().item()) + 1
    freqs = torch.einsum("i,j->ij", torch.arange(n, dtype=inv_freq.dtype), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos = emb.cos()[position_ids].unsqueeze(1).to(q.dtype)
    sin = emb.sin()[position_ids].unsqueeze(1).to(q.dtype)

    def rot(x):
        return torch.cat((-x[..., D // 2:], x[..., : D // 2]), dim=-1)

    return q * cos + rot(q) * sin, k * cos + rot(k) * sin


def llama_forward(sd: SD, input_ids: torch.Tensor, heads: int, layers: int, eps: float = 1e-6,
                  position_ids: Optional[torch.Tensor] = None,
                  past: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None):
    """LlamaForCausalLM.forward (:661-743) / LlamaModel.forward (:496-627) / LlamaDecoderLayer (:280-332) /
    LlamaAttention (:212-263).  Attention = xformers memory_efficient_attention with LowerTriangularMask when
    q_len > 1 and no mask when q_len == 1 (:240-256; the padding mask is ignored by the reference).  For
    q_len > 1 with a non-empty past the causal mask is bottom-right aligned (each new token sees the whole
    past); the reference's behaviour there depends on the xformers version and is not used by its scripts.
    Returns (logits [B,S,V], hidden [B,S,h], new_past)."""
    B, S = input_ids.shape
    past_len = 0 if past is None else past[0][0].shape[2]
    if position_ids is None:
        position_ids = torch.arange(past_len, past_len + S).unsqueeze(0).expand(B, S)
    x = F.embedding(input_ids, sd["model.embed_tokens.weight"])
    h = x.shape[-1]
    D = h // heads
    new_past = []
    for l in range(layers):
        p = f"model.layers.{l}."
        r = x
        n = _rms(x, sd[p + "input_layernorm.weight"], eps)
        q = F.linear(n, sd[p + "self_attn.q_proj.weight"]).view(B, S, heads, D).transpose(1, 2)
        k = F.linear(n, sd[p + "self_attn.k_proj.weight"]).view(B, S, heads, D).transpose(1, 2)
        v = F.linear(n, sd[p + "self_attn.v_proj.weight"]).view(B, S, heads, D).transpose(1, 2)
        q, k = _rope(q, k, position_ids)
        if past is not None:
            k = torch.cat([past[l][0], k], dim=2)
            v = torch.cat([past[l][1], v], dim=2)
        new_past.append((k, v))
        scores = (q @ k.transpose(-1, -2)) / math.sqrt(D)
        if S > 1:
            nk = k.shape[2]
            i = torch.arange(S)[:, None]
            j = torch.arange(nk)[None, :]
            scores = scores.masked_fill(j > i + (nk - S), float("-inf"))
        a = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(B, S, h)
        x = r + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        r = x
        n = _rms(x, sd[p + "post_attention_layernorm.weight"], eps)
        # LlamaMLP.forward (:185-186)
        m = F.linear(F.silu(F.linear(n, sd[p + "mlp.gate_proj.weight"])) * F.linear(n, sd[p + "mlp.up_proj.weight"]),
                     sd[p + "mlp.down_proj.weight"])
        x = r + m
    hidden = _rms(x, sd["model.norm.weight"], eps)
    logits = F.linear(hidden, sd["lm_head.weight"])
    return logits, hidden, new_past
