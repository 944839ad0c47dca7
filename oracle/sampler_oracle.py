"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the next-token selection the reference obtains from HF
`GenerationMixin` at its call sites (scripts/seed_llama_inference_8B.py:26-38: temperature=1.0, top_p=0.5,
do_sample=True, num_beams=1; gradio_demo/seed_llama_flask.py:172).  Only tests/ may import this.

The arithmetic lives in an un-vendored dependency (transformers==4.30.2, requirements.txt:8), so its published
algorithm is restated here and pinned against the `transformers` installed in this image
(tests/test_oracle.py::test_sampler_oracle_matches_transformers_warpers):

  TemperatureLogitsWarper   scores / temperature
  TopPLogitsWarper          sort ascending, cumulative softmax, remove tokens with cumsum <= 1 - top_p,
                            always keep the last (most probable) one.  Tokens whose score EQUALS the least kept score
                            straddle the boundary in HF by sort order (implementation-defined: torch.sort is not stable
                            on CUDA); here -- and in the kernel -- all of them are kept (`ties="all"`), which is the
                            threshold form of the same rule and is identical to HF on tie-free scores
  sample                    softmax over the kept scores, one multinomial draw

The multinomial draw itself is RNG specific; seedb200's kernel (seed_b200/csrc/sampler.cu) defines it as the inverse
CDF over the kept tokens in index order evaluated at u * K, u = one Philox4x32-10 uniform keyed by (seed; offset +
step, row).  `sample_ref` restates exactly that in float64 and reports how far the draw is from the nearest CDF
boundary, so a test can tell a real mismatch from an fp32 summation-order flip.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Philox4x32 with 10 rounds (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)."""
    c = [int(x) & MASK for x in counter]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def philox_uniform(seed: int, offset: int, row: int) -> np.float32:
    """uniform in (0, 1]: counter (offset_lo, offset_hi, row, 0), key = seed, first word, (x + 0.5) / 2^32 in fp32."""
    c = philox4x32_10((offset & MASK, (offset >> 32) & MASK, row, 0), (seed & MASK, (seed >> 32) & MASK))
    return np.float32(np.float32(c[0]) * np.float32(2.3283064365386963e-10) + np.float32(1.1641532182693481e-10))


def warp(logits: np.ndarray, temperature: float, top_p: float, ties: str = "all"):
    """-> (probs over the full vocabulary after both warpers (0 outside the nucleus), kept mask, boundary margin).
    The margin is |cumsum - (1 - top_p)| of the token closest to the nucleus boundary: a kernel that sums in another
    order may legitimately differ on a token whose margin is ~1e-6."""
    x = logits.astype(np.float64) / float(temperature)
    p = np.exp(x - x.max())
    p /= p.sum()
    keep = np.ones_like(p, dtype=bool)
    margin = 1.0
    if top_p < 1.0:
        order = np.argsort(x, kind="stable")                 # ascending, like torch.sort(descending=False)
        cs = np.cumsum(p[order])
        remove_sorted = cs <= (1.0 - top_p)
        remove_sorted[-1:] = False                           # min_tokens_to_keep = 1
        keep[order] = ~remove_sorted
        if ties == "all":                                    # every token tied with the least kept score stays
            keep = x >= x[keep].min()
        margin = float(np.min(np.abs(cs - (1.0 - top_p))))
    q = np.where(keep, p, 0.0)
    return q / q.sum(), keep, margin


def sample_ref(logits: np.ndarray, do_sample: bool, temperature: float = 1.0, top_p: float = 1.0, seed: int = 0,
               offset: int = 0, step: int = 0, row: int = 0):
    """-> (token, draw_margin, nucleus_margin).  Greedy: argmax, ties to the lowest id (torch.argmax)."""
    if not do_sample:
        return int(np.argmax(logits.astype(np.float32))), 1.0, 1.0
    q, keep, nmargin = warp(logits, temperature, top_p)
    cdf = np.cumsum(q)
    u = float(philox_uniform(seed, offset + step, row))
    tok = int(np.searchsorted(cdf, u, side="left"))          # first index with cdf >= u
    tok = min(tok, len(q) - 1)
    while not keep[tok]:                                     # (cdf is flat over removed tokens: step to a kept one)
        tok += 1
    edges = np.concatenate([[0.0], cdf])
    dmargin = float(np.min(np.abs(edges - u)))
    return tok, dmargin, nmargin
