/*
 * vq_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * CPU restatement, in plain C, of VectorQuantizer2.forward's nearest-neighbour search
 * (/root/reference/models/seed_qformer/qformer_quantizer.py:94-98):
 *
 *     d = torch.sum(z ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * einsum('bd,dn->bn', z, E^T)
 *     ids = torch.argmin(d, dim=1)
 *
 * The arithmetic is pinned operation by operation so that the CUDA kernel (seed_b200/csrc/vq.cu) can be
 * checked BIT-EXACTLY against it:
 *   - every dot product is one binary32 fma chain over d = 0..dim-1 in index order;
 *   - squared norms are binary32 sums in index order of the binary32 squares;
 *   - mode 0 ("fp16", the reference's fp16 GPU mode, configs/tokenizer/seed_llama_tokenizer_hf.yaml:3):
 *     each tensor torch materialises in half precision is rounded to binary16 where torch rounds it:
 *     z**2 and E**2 elementwise, the two sums, the einsum result, A+B, 2*C and the final subtraction;
 *   - mode 1 ("fp32", the reference's CPU / fp16=False mode): d = (A + B) - 2*C in binary32, in that order;
 *   - argmin keeps the first (lowest) index among equal minima, like torch.argmin.
 * tests/test_oracle.py pins this file against torch running the reference expression itself (golden
 * vectors under tests/golden/) -- exact agreement wherever the top-2 margin is not a rounding tie.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o libvq_oracle.so vq_oracle.c -lm  (seed_b200/build.py)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float h2f(uint16_t bits) {
  _Float16 h;
  memcpy(&h, &bits, sizeof(h));
  return (float)h;
}
static float r16(float x) { return (float)(_Float16)x; } /* round-to-nearest-even to binary16 */

/* z: [n, dim] binary16 bit patterns; codebook: [n_codes, dim]; ids: [n] int64; margin (nullable): [n]
 * d(second best) - d(best) as seen by this arithmetic. Returns 0, or -1 on bad arguments. */
int vq_oracle_argmin(const uint16_t* z, const uint16_t* codebook, int n, int n_codes, int dim, int mode,
                     int64_t* ids, float* margin) {
  if (!z || !codebook || !ids || n <= 0 || n_codes <= 0 || dim <= 0 || dim > 256 || (mode != 0 && mode != 1))
    return -1;
  float zr[256], er[256];
  for (int r = 0; r < n; ++r) {
    float A = 0.0f;
    for (int d = 0; d < dim; ++d) {
      zr[d] = h2f(z[(size_t)r * dim + d]);
      float sq = zr[d] * zr[d];
      if (mode == 0) sq = r16(sq);
      A = A + sq;
    }
    if (mode == 0) A = r16(A);
    float best = INFINITY, second = INFINITY;
    int64_t best_i = 0;
    int have = 0;
    for (int c = 0; c < n_codes; ++c) {
      float B = 0.0f, C = 0.0f;
      for (int d = 0; d < dim; ++d) {
        er[d] = h2f(codebook[(size_t)c * dim + d]);
        float sq = er[d] * er[d];
        if (mode == 0) sq = r16(sq);
        B = B + sq;
        C = fmaf(zr[d], er[d], C);
      }
      float dist;
      if (mode == 0) {
        B = r16(B);
        const float C16 = r16(C);
        const float t = r16(A + B);
        dist = r16(t - r16(2.0f * C16));
      } else {
        const float t = A + B;
        dist = t - 2.0f * C;
      }
      if (dist < best) {
        second = best;
        best = dist;
        best_i = c;
        have = 1;
      } else if (dist < second) {
        second = dist;
      }
    }
    ids[r] = have ? best_i : 0;
    if (margin) margin[r] = second - best;
  }
  return 0;
}
