// rowwise.cu -- the HBM-bound kernels of the path: LayerNorm, RMSNorm, patch unfold, RoPE + KV-cache
// append, embedding gather, row broadcast.  All are single-pass, 16-byte vectorised and keep one row in
// registers, so the algorithmic traffic (read x once, write y once) is also the DRAM traffic.
#include "common.cuh"

namespace sb {

// ----------------------------------------------------------------------------
// LayerNorm / RMSNorm
//   LayerNorm: eva_vit.py:201-202 (norm1/norm2, eps 1e-6, fp32 under autocast), blip2.py:179-184
//   (ln_vision, eps 1e-5), qformer_causual.py:96,254,336 (eps 1e-12), vit.py:147-150 (eps 1e-6).
//   RMSNorm: llama_xformer.py:105-113 -- variance and x*rsqrt in fp32, ROUND to fp16, then * weight.
// TPR threads cooperate on one row; each holds VPT 8-half vectors.
// ----------------------------------------------------------------------------
// STATS: only the per-row (mean, rstd) pair is written (float2 per row, through `y`): the LayerNorm-folded GEMM
// (seedb200_gemm_desc.ln_stats) applies the normalisation in its epilogue and the normalised tensor never exists.
template <int TPR, int VPT, bool RMS, bool STATS = false>
__global__ void __launch_bounds__(256)
norm_kernel(const __half* __restrict__ x, long long ldx, const __half* __restrict__ w,
            const __half* __restrict__ bvec, __half* __restrict__ y, long long ldy, int rows, int cols,
            float eps) {
  constexpr int ROWS_PER_BLOCK = 256 / TPR;
  __shared__ float red[2][8];
  const int tid = threadIdx.x;
  const int sub = tid % TPR;
  const int row = blockIdx.x * ROWS_PER_BLOCK + tid / TPR;
  const bool active = row < rows;
  const int nvec = cols / 8;

  float v[VPT][8];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = sub + i * TPR;
    if (active && vi < nvec) {
      const uint4 raw = *reinterpret_cast<const uint4*>(x + (long long)row * ldx + vi * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[i][2 * j] = f.x; v[i][2 * j + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += RMS ? v[i][j] * v[i][j] : v[i][j];
  }

  auto row_reduce = [&](float val, int slot) -> float {
    val = warp_sum(val);
    if constexpr (TPR > 32) {
      if ((tid & 31) == 0) red[slot][tid >> 5] = val;
      __syncthreads();
      float t = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[slot][i];
      return t;
    } else {
      return val;
    }
  };

  const float inv_n = 1.0f / (float)cols;
  float mean = 0.0f, rstd;
  if constexpr (RMS) {
    const float ss = row_reduce(sum, 0);
    rstd = rsqrtf(ss * inv_n + eps);
  } else {
    mean = row_reduce(sum, 0) * inv_n;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = sub + i * TPR;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float dlt = v[i][j] - mean; sq += dlt * dlt; }
      }
    }
    const float var = row_reduce(sq, 1) * inv_n;
    rstd = rsqrtf(var + eps);
  }
  if constexpr (STATS) {
    if (active && sub == 0) reinterpret_cast<float2*>(y)[row] = make_float2(mean, rstd);
    return;
  }

#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = sub + i * TPR;
    if (active && vi < nvec) {
      const uint4 wraw = __ldg(reinterpret_cast<const uint4*>(w + vi * 8));
      const __half* wh = reinterpret_cast<const __half*>(&wraw);
      uint4 outv;
      __half* oh = reinterpret_cast<__half*>(&outv);
      if constexpr (RMS) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __half n16 = __float2half_rn(v[i][j] * rstd);
          oh[j] = __float2half_rn(__half2float(n16) * __half2float(wh[j]));
        }
      } else {
        const uint4 braw = __ldg(reinterpret_cast<const uint4*>(bvec + vi * 8));
        const __half* bh = reinterpret_cast<const __half*>(&braw);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          oh[j] = __float2half_rn((v[i][j] - mean) * rstd * __half2float(wh[j]) + __half2float(bh[j]));
      }
      *reinterpret_cast<uint4*>(y + (long long)row * ldy + vi * 8) = outv;
    }
  }
}

template <bool RMS>
static int launch_norm(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy, int rows,
                       int cols, float eps, cudaStream_t stream) {
  SB_REQUIRE(rows > 0 && cols > 0, "norm: empty input");
  SB_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "norm: cols/ld must be multiples of 8 (cols=%d)", cols);
  SB_REQUIRE(x && w && y && (RMS || b), "norm: null operand");
  const __half* xp = static_cast<const __half*>(x);
  const __half* wp = static_cast<const __half*>(w);
  const __half* bp = static_cast<const __half*>(b);
  __half* yp = static_cast<__half*>(y);
  const int nvec = cols / 8;
#define SB_NORM_LAUNCH(TPR_, VPT_)                                                              \
  {                                                                                             \
    const int rpb = 256 / TPR_;                                                                 \
    norm_kernel<TPR_, VPT_, RMS><<<(rows + rpb - 1) / rpb, 256, 0, stream>>>(xp, ldx, wp, bp, yp, ldy, rows, \
                                                                              cols, eps);       \
    SB_LAUNCH_CHECK();                                                                          \
    return 0;                                                                                   \
  }
  if (nvec <= 32 * 1) SB_NORM_LAUNCH(32, 1)
  if (nvec <= 32 * 2) SB_NORM_LAUNCH(32, 2)
  if (nvec <= 32 * 3) SB_NORM_LAUNCH(32, 3)      // 768
  if (nvec <= 32 * 6) SB_NORM_LAUNCH(32, 6)      // 1408
  if (nvec <= 256 * 2) SB_NORM_LAUNCH(256, 2)    // 4096
  if (nvec <= 256 * 3) SB_NORM_LAUNCH(256, 3)    // 5120
  if (nvec <= 256 * 8) SB_NORM_LAUNCH(256, 8)
#undef SB_NORM_LAUNCH
  set_error("norm: cols=%d too large", cols);
  return SEEDB200_ERR_UNSUPPORTED;
}

// (mean, rstd) from the 64-column (sum, sum of squares) groups a GEMM epilogue left behind (GemmParams::row_moments):
// thread per row, groups added in index order (deterministic), the cancellation-prone E[x^2] - mean^2 in fp64
__global__ void stats_from_moments_kernel(const float2* __restrict__ mom, int rows, int groups, float inv_cols, float eps,
                                          float2* __restrict__ stats) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float2* mp = mom + (long long)r * groups;
  float s = 0.0f, q = 0.0f;
  for (int g = 0; g < groups; ++g) {
    const float2 v = mp[g];
    s += v.x;
    q += v.y;
  }
  const double mean = (double)s * (double)inv_cols;
  double var = (double)q * (double)inv_cols - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[r] = make_float2((float)mean, rsqrtf((float)var + eps));
}

int row_stats_from_moments(const void* moments, int rows, int cols, float eps, void* stats, cudaStream_t stream) {
  SB_REQUIRE(moments && stats && rows > 0 && cols > 0 && cols % 64 == 0, "row_stats_from_moments: bad arguments (cols=%d)", cols);
  stats_from_moments_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(static_cast<const float2*>(moments), rows, cols / 64,
                                                                    1.0f / (float)cols, eps, static_cast<float2*>(stats));
  SB_LAUNCH_CHECK();
  return 0;
}

int row_stats(const void* x, int64_t ldx, int rows, int cols, float eps, void* stats, cudaStream_t stream) {
  SB_REQUIRE(x && stats && rows > 0 && cols > 0, "row_stats: bad arguments");
  SB_REQUIRE(cols % 8 == 0 && ldx % 8 == 0, "row_stats: cols/ld must be multiples of 8 (cols=%d)", cols);
  const __half* xp = static_cast<const __half*>(x);
  __half* yp = static_cast<__half*>(stats);
  const int nvec = cols / 8;
#define SB_STATS_LAUNCH(TPR_, VPT_)                                                              \
  {                                                                                              \
    const int rpb = 256 / TPR_;                                                                  \
    norm_kernel<TPR_, VPT_, false, true><<<(rows + rpb - 1) / rpb, 256, 0, stream>>>(xp, ldx, nullptr, nullptr, yp, 0, \
                                                                                      rows, cols, eps);              \
    SB_LAUNCH_CHECK();                                                                           \
    return 0;                                                                                    \
  }
  if (nvec <= 32 * 1) SB_STATS_LAUNCH(32, 1)
  if (nvec <= 32 * 2) SB_STATS_LAUNCH(32, 2)
  if (nvec <= 32 * 3) SB_STATS_LAUNCH(32, 3)
  if (nvec <= 32 * 6) SB_STATS_LAUNCH(32, 6)      // 1408
  if (nvec <= 256 * 2) SB_STATS_LAUNCH(256, 2)
  if (nvec <= 256 * 8) SB_STATS_LAUNCH(256, 8)
#undef SB_STATS_LAUNCH
  set_error("row_stats: cols=%d too large", cols);
  return SEEDB200_ERR_UNSUPPORTED;
}

// W' = fp16(W * gamma), c[n] = sum_k W'[n,k] (of the ROUNDED values), b'[n] = sum_k W[n,k] beta[k] + bias[n].
// One warp per output row; run once per weight at handle creation.
__global__ void __launch_bounds__(256)
ln_fold_weights_kernel(const __half* __restrict__ W, long long ldw, const __half* __restrict__ gamma,
                       const __half* __restrict__ beta, const __half* __restrict__ bias, int N, int K,
                       __half* __restrict__ Wo, float* __restrict__ c, float* __restrict__ b) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.0f, bs = 0.0f;
  for (int k = lane; k < K; k += 32) {
    const float w = __half2float(W[(long long)n * ldw + k]);
    const __half wf = __float2half_rn(w * __half2float(gamma[k]));
    Wo[(long long)n * K + k] = wf;
    cs += __half2float(wf);
    bs = fmaf(w, __half2float(beta[k]), bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    c[n] = cs;
    b[n] = bs + (bias != nullptr ? __half2float(bias[n]) : 0.0f);
  }
}

int ln_fold_weights(const void* W, int64_t ldw, const void* gamma, const void* beta, const void* bias, int N, int K,
                    void* W_out, void* c_out, void* b_out, cudaStream_t stream) {
  SB_REQUIRE(W && gamma && beta && W_out && c_out && b_out && N > 0 && K > 0, "ln_fold_weights: bad arguments");
  ln_fold_weights_kernel<<<(N + 7) / 8, 256, 0, stream>>>(
      static_cast<const __half*>(W), ldw, static_cast<const __half*>(gamma), static_cast<const __half*>(beta),
      static_cast<const __half*>(bias), N, K, static_cast<__half*>(W_out), static_cast<float*>(c_out),
      static_cast<float*>(b_out));
  SB_LAUNCH_CHECK();
  return 0;
}

int layernorm(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy, int rows, int cols,
              float eps, cudaStream_t stream) {
  return launch_norm<false>(x, ldx, w, b, y, ldy, rows, cols, eps, stream);
}
int rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int cols, float eps,
            cudaStream_t stream) {
  return launch_norm<true>(x, ldx, w, nullptr, y, ldy, rows, cols, eps, stream);
}

// ----------------------------------------------------------------------------
// Patch unfold (eva_vit.py:222,229: Conv2d(3,1408,k=14,s=14) == GEMM over unfolded patches).
// One CTA per (image, patch row): the 3 x 14 image rows (42 x 448 bytes) are read with 16-byte loads into shared
// memory, then the 16 GEMM rows of kpad halves are assembled from there and written with 16-byte stores
// (HBM-bound: 301 KB in + 303 KB out per image; the first version moved single halves and ran at 21 % of the
// copy bandwidth).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_kernel(const __half* __restrict__ img, __half* __restrict__ cols, int kpad) {
  __shared__ __align__(16) __half tile[42 * 224];        // [c*14 + dy][x]
  const int b = blockIdx.x >> 4, py = blockIdx.x & 15;
  const __half* src = img + (long long)b * 3 * 224 * 224;
  for (int i = threadIdx.x; i < 42 * 28; i += 256) {     // 28 16-byte vectors per image row
    const int r = i / 28, v = i - r * 28;
    const int c = r / 14, dy = r - c * 14;
    reinterpret_cast<uint4*>(tile)[i] =
        __ldg(reinterpret_cast<const uint4*>(src + (long long)(c * 224 + py * 14 + dy) * 224) + v);
  }
  __syncthreads();
  __half* dst = cols + ((long long)b * 256 + py * 16) * kpad;
  const int vpr = kpad / 8;                              // 16-byte vectors per GEMM row
  for (int i = threadIdx.x; i < 16 * vpr; i += 256) {
    const int px = i / vpr, v = i - px * vpr;
    uint4 o;
    __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = v * 8 + j;
      __half val = __float2half(0.0f);
      if (col < 588) {
        const int r = col / 14, dx = col - r * 14;       // r = c*14 + dy
        val = tile[r * 224 + px * 14 + dx];
      }
      oh[j] = val;
    }
    *reinterpret_cast<uint4*>(dst + (long long)px * kpad + v * 8) = o;
  }
}

int patchify(const void* images, int B, void* cols, int kpad, cudaStream_t stream) {
  SB_REQUIRE(images && cols && B > 0, "patchify: bad arguments");
  SB_REQUIRE(kpad >= 588 && kpad % 8 == 0, "patchify: kpad=%d must be >= 588 and a multiple of 8", kpad);
  patchify_kernel<<<B * 16, 256, 0, stream>>>(static_cast<const __half*>(images), static_cast<__half*>(cols), kpad);
  SB_LAUNCH_CHECK();
  return 0;
}

// dst[g * group_stride + r, :] = src[r, :] for r < src_rows (cls row of every image, query tokens of
// every image: eva_vit.py:372-373 cls_token.expand, qformer_quantizer.py:293 query_tokens.expand)
__global__ void __launch_bounds__(256)
broadcast_rows_kernel(const __half* __restrict__ src, int src_rows, int cols, __half* __restrict__ dst,
                      long long ldd, long long group_stride_rows, int groups) {
  const int nvec = cols / 8;
  const long long total = (long long)groups * src_rows * nvec;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int vi = (int)(i % nvec);
    const long long t = i / nvec;
    const int r = (int)(t % src_rows);
    const long long g = t / src_rows;
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(src + (long long)r * cols + vi * 8));
    *reinterpret_cast<uint4*>(dst + (g * group_stride_rows + r) * ldd + vi * 8) = val;
  }
}

int broadcast_rows(const void* src, int src_rows, int cols, void* dst, int64_t ldd, int64_t group_stride_rows,
                   int groups, cudaStream_t stream) {
  SB_REQUIRE(cols % 8 == 0 && ldd % 8 == 0, "broadcast_rows: cols/ld must be multiples of 8");
  const long long total = (long long)groups * src_rows * (cols / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  broadcast_rows_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __half*>(src), src_rows, cols,
                                                    static_cast<__half*>(dst), ldd, group_stride_rows, groups);
  SB_LAUNCH_CHECK();
  return 0;
}

// ----------------------------------------------------------------------------
// Embedding gather: out[i,:] = table[ids[i],:] (llama_xformer.py:544, qformer_quantizer.py:133).
// Ids outside [0, n_rows) produce a zero row (the reference raises an IndexError on the host side;
// the Python mirror validates ids where the reference would have failed).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embedding_kernel(const __half* __restrict__ table, long long ld, const long long* __restrict__ ids, int n, int cols,
                 __half* __restrict__ out, long long ldo, long long n_rows) {
  const int nvec = cols / 8;
  const long long total = (long long)n * nvec;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int vi = (int)(i % nvec);
    const long long r = i / nvec;
    const long long id = ids[r];
    uint4 val = make_uint4(0, 0, 0, 0);
    if (id >= 0 && id < n_rows) val = __ldg(reinterpret_cast<const uint4*>(table + id * ld + vi * 8));
    *reinterpret_cast<uint4*>(out + r * ldo + vi * 8) = val;
  }
}

int embedding(const void* table, int64_t ld, const int64_t* ids, int n, int cols, void* out, int64_t ldo,
              int64_t n_rows, cudaStream_t stream) {
  SB_REQUIRE(table && ids && out && n > 0, "embedding: bad arguments");
  SB_REQUIRE(cols % 8 == 0 && ld % 8 == 0 && ldo % 8 == 0, "embedding: cols/ld must be multiples of 8");
  const long long total = (long long)n * (cols / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  embedding_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __half*>(table), ld,
                                               reinterpret_cast<const long long*>(ids), n, cols,
                                               static_cast<__half*>(out), ldo, n_rows);
  SB_LAUNCH_CHECK();
  return 0;
}

// ----------------------------------------------------------------------------
// RoPE + KV-cache append (llama_xformer.py:152-161 apply_rotary_pos_emb with the rotate-half
// convention, fp16 products and sum; :234-239 cache growth by torch.cat replaced by an in-place append).
// cos/sin tables [max_pos, D/2] hold fp16(cos(fp32(pos * inv_freq))) like LlamaRotaryEmbedding.
// One thread handles 8 consecutive dims of the low half and the matching 8 of the high half.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rope_kv_kernel(const __half* __restrict__ qkv, const long long* __restrict__ positions, int S, int H, int D,
               int past_len, int max_seq, int max_pos, const __half* __restrict__ cos_t,
               const __half* __restrict__ sin_t, __half* __restrict__ q_out, __half* __restrict__ k_cache,
               __half* __restrict__ v_cache, long long total, const int* __restrict__ dyn) {
  const int half_d = D / 2;
  const int vec_per_head = half_d / 8;
  const long long HD = (long long)H * D;
  pdl_trigger();
  pdl_wait();
  if (dyn != nullptr) past_len = dyn[0];     // graph-replayed decode step: the cache length lives in device memory
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int vi = (int)(i % vec_per_head);
    long long t = i / vec_per_head;
    const int h = (int)(t % H);
    t /= H;                                  // token index in [0, B*S)
    const int b = (int)(t / S), s = (int)(t % S);
    long long pos = positions ? positions[t] : (long long)(past_len + s);
    if (pos < 0) pos = 0;
    if (pos >= max_pos) pos = max_pos - 1;
    const uint4 craw = __ldg(reinterpret_cast<const uint4*>(cos_t + pos * half_d + vi * 8));
    const uint4 sraw = __ldg(reinterpret_cast<const uint4*>(sin_t + pos * half_d + vi * 8));
    const __half* ch = reinterpret_cast<const __half*>(&craw);
    const __half* sh = reinterpret_cast<const __half*>(&sraw);
    const __half* row = qkv + t * 3 * HD + (long long)h * D;
    int crow = past_len + s;
    if (crow >= max_seq) crow = max_seq - 1;           // only reachable through dyn (host-checked otherwise)
    const long long cache_row = (((long long)b * H + h) * max_seq + crow) * D;
#pragma unroll
    for (int which = 0; which < 2; ++which) {          // 0 = q, 1 = k
      const uint4 lo_raw = *reinterpret_cast<const uint4*>(row + which * HD + vi * 8);
      const uint4 hi_raw = *reinterpret_cast<const uint4*>(row + which * HD + half_d + vi * 8);
      const __half* lo = reinterpret_cast<const __half*>(&lo_raw);
      const __half* hi = reinterpret_cast<const __half*>(&hi_raw);
      uint4 olo_raw, ohi_raw;
      __half* olo = reinterpret_cast<__half*>(&olo_raw);
      __half* ohi = reinterpret_cast<__half*>(&ohi_raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // q_embed = q*cos + rotate_half(q)*sin, every op rounded to fp16 as torch does on fp16 tensors
        const float c = __half2float(ch[j]), sn = __half2float(sh[j]);
        const float xl = __half2float(lo[j]), xh = __half2float(hi[j]);
        const float a_lo = __half2float(__float2half_rn(xl * c));
        const float b_lo = __half2float(__float2half_rn(-xh * sn));
        const float a_hi = __half2float(__float2half_rn(xh * c));
        const float b_hi = __half2float(__float2half_rn(xl * sn));
        olo[j] = __float2half_rn(a_lo + b_lo);
        ohi[j] = __float2half_rn(a_hi + b_hi);
      }
      __half* dst = (which == 0) ? (q_out + t * HD + (long long)h * D) : (k_cache + cache_row);
      *reinterpret_cast<uint4*>(dst + vi * 8) = olo_raw;
      *reinterpret_cast<uint4*>(dst + half_d + vi * 8) = ohi_raw;
    }
    const uint4 v_lo = *reinterpret_cast<const uint4*>(row + 2 * HD + vi * 8);
    const uint4 v_hi = *reinterpret_cast<const uint4*>(row + 2 * HD + half_d + vi * 8);
    *reinterpret_cast<uint4*>(v_cache + cache_row + vi * 8) = v_lo;
    *reinterpret_cast<uint4*>(v_cache + cache_row + half_d + vi * 8) = v_hi;
  }
}

int rope_kv_append_tables(const void* qkv, const int64_t* positions, int B, int S, int H, int D, int past_len,
                          int max_seq, int max_pos, const void* cos_t, const void* sin_t, void* q_out,
                          void* k_cache, void* v_cache, cudaStream_t stream, const int* dyn) {
  SB_REQUIRE(qkv && q_out && k_cache && v_cache && cos_t && sin_t, "rope_kv_append: null operand");
  SB_REQUIRE(D % 16 == 0, "rope_kv_append: head_dim %d must be a multiple of 16", D);
  SB_REQUIRE(past_len + S <= max_seq, "rope_kv_append: past_len %d + S %d exceeds max_seq %d", past_len, S, max_seq);
  const long long total = (long long)B * S * H * (D / 16);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  SB_CHECK_CUDA(launch_chain(rope_kv_kernel, dim3(blocks), dim3(256), 0, stream, static_cast<const __half*>(qkv),
                             reinterpret_cast<const long long*>(positions), S, H, D, past_len, max_seq, max_pos,
                             static_cast<const __half*>(cos_t), static_cast<const __half*>(sin_t),
                             static_cast<__half*>(q_out), static_cast<__half*>(k_cache), static_cast<__half*>(v_cache),
                             total, dyn));
  SB_LAUNCH_CHECK();
  return 0;
}

// cos/sin tables exactly as LlamaRotaryEmbedding builds them (llama_xformer.py:118-135):
// inv_freq = 1 / base^(2i/D) in fp32, freqs = pos * inv_freq in fp32, cos/sin in fp32, cast to fp16.
__global__ void rope_table_kernel(__half* cos_t, __half* sin_t, int max_pos, int half_d, float base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * half_d) return;
  const int pos = i / half_d, j = i - pos * half_d;
  const float expo = (float)(2 * j) / (float)(2 * half_d);
  const float inv_freq = 1.0f / powf(base, expo);
  const float f = (float)pos * inv_freq;
  cos_t[i] = __float2half_rn(cosf(f));
  sin_t[i] = __float2half_rn(sinf(f));
}

int build_rope_tables(void* cos_t, void* sin_t, int max_pos, int D, float base, cudaStream_t stream) {
  const int n = max_pos * (D / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<__half*>(cos_t), static_cast<__half*>(sin_t),
                                                         max_pos, D / 2, base);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sb

extern "C" {

int seedb200_layernorm(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy, int rows,
                       int cols, float eps, void* stream) {
  return sb::layernorm(x, ldx, w, b, y, ldy, rows, cols, eps, static_cast<cudaStream_t>(stream));
}
int seedb200_row_stats_from_moments(const void* moments, int rows, int cols, float eps, void* stats_out, void* stream) {
  return sb::row_stats_from_moments(moments, rows, cols, eps, stats_out, static_cast<cudaStream_t>(stream));
}
int seedb200_row_stats(const void* x, int64_t ldx, int rows, int cols, float eps, void* stats_out, void* stream) {
  return sb::row_stats(x, ldx, rows, cols, eps, stats_out, static_cast<cudaStream_t>(stream));
}
int seedb200_ln_fold_weights(const void* W, int64_t ldw, const void* gamma, const void* beta, const void* bias, int N,
                             int K, void* W_out, void* c_out, void* b_out, void* stream) {
  return sb::ln_fold_weights(W, ldw, gamma, beta, bias, N, K, W_out, c_out, b_out, static_cast<cudaStream_t>(stream));
}
int seedb200_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int cols, float eps,
                     void* stream) {
  return sb::rmsnorm(x, ldx, w, y, ldy, rows, cols, eps, static_cast<cudaStream_t>(stream));
}
int seedb200_patchify(const void* images, int B, void* cols, int kpad, void* stream) {
  return sb::patchify(images, B, cols, kpad, static_cast<cudaStream_t>(stream));
}
int seedb200_embedding(const void* table, int64_t ld, const int64_t* ids, int n, int cols, void* out, int64_t ldo,
                       int64_t n_rows, void* stream) {
  return sb::embedding(table, ld, ids, n, cols, out, ldo, n_rows, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
