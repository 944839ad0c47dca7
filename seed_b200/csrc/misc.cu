// misc.cu -- small-M kernels of the LLaMA decode step (HBM-bound: every weight byte is read once per
// token, llama_xformer.py:745-776 generation loop) and a row-wise add used at handle creation.
//   gemv              y[m,:] = x[m,:] . W^T for m <= 4 rows: the batch-1 decode form of every nn.Linear in
//                     LlamaDecoderLayer (llama_xformer.py:186,223-225,258) and lm_head (:718)
//   decode_attention  one query token against the KV cache (llama_xformer.py:240-256 with attn_bias=None),
//                     split over the key axis so that B*H*splits CTAs cover the SMs
#include "common.cuh"
#include "ops.h"

namespace sb {

__global__ void add_rows_kernel(const __half* a, const __half* b, __half* out, int rows, int cols, int b_rows) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  out[i] = __float2half_rn(__half2float(a[i]) + __half2float(b[(long long)(r % b_rows) * cols + c]));
}

int add_rows(const void* a, const void* b, void* out, int rows, int cols, int b_rows, cudaStream_t stream) {
  const long long n = (long long)rows * cols;
  add_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(static_cast<const __half*>(a),
                                                                  static_cast<const __half*>(b),
                                                                  static_cast<__half*>(out), rows, cols, b_rows);
  SB_LAUNCH_CHECK();
  return 0;
}

// ----------------------------------------------------------------------------
// GEMV: M <= GEMV_MAXM activation rows staged in shared memory; each warp owns output columns and streams
// the matching weight rows with 16-byte loads (2 rows x 4 vectors in flight per lane).
// mode 1: W rows are [128 gate | 128 up] blocks and out[m,j] = silu(gate_j) * up_j (same rounding points
// as the GEMM epilogue, llama_xformer.py:186).
// ----------------------------------------------------------------------------
constexpr int GEMV_MAXM = 4;

__device__ __forceinline__ void dot8(const uint4& w, const uint4& x, float& acc) {
  const __half2* wh = reinterpret_cast<const __half2*>(&w);
  const __half2* xh = reinterpret_cast<const __half2*>(&x);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = __half22float2(wh[j]), b = __half22float2(xh[j]);
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
  }
}

template <int M, int MODE>
__global__ void __launch_bounds__(256)
gemv_kernel(const __half* __restrict__ x, const __half* __restrict__ W, long long ldw, __half* __restrict__ out,
            const __half* __restrict__ residual, int N, int K) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* xs = reinterpret_cast<uint4*>(smem_raw);     // [M][K/8]
  const int nvec = K / 8;
  for (int i = threadIdx.x; i < M * nvec; i += 256) xs[i] = reinterpret_cast<const uint4*>(x)[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_out = (MODE == 1) ? N / 2 : N;
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  for (int j = gw; j < n_out; j += nw) {
    long long r0, r1;
    if (MODE == 1) { r0 = (long long)(j / 128) * 256 + (j % 128); r1 = r0 + 128; }
    else { r0 = j; r1 = j; }
    const uint4* w0 = reinterpret_cast<const uint4*>(W + r0 * ldw);
    const uint4* w1 = reinterpret_cast<const uint4*>(W + r1 * ldw);
    float a0[M], a1[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { a0[m] = 0.0f; a1[m] = 0.0f; }
    for (int v = lane; v < nvec; v += 128) {
      uint4 wa[4], wb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int vi = v + u * 32;
        if (vi < nvec) {
          wa[u] = __ldg(w0 + vi);
          if (MODE == 1) wb[u] = __ldg(w1 + vi);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int vi = v + u * 32;
        if (vi < nvec) {
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const uint4 xv = xs[m * nvec + vi];
            dot8(wa[u], xv, a0[m]);
            if (MODE == 1) dot8(wb[u], xv, a1[m]);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
      a0[m] = warp_sum(a0[m]);
      if (MODE == 1) a1[m] = warp_sum(a1[m]);
    }
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
        __half h;
        if (MODE == 1) {
          const float g = __half2float(__float2half_rn(a0[m]));
          const float u = __half2float(__float2half_rn(a1[m]));
          const float s = __half2float(__float2half_rn(g / (1.0f + __expf(-g))));
          h = __float2half_rn(s * u);
        } else {
          h = __float2half_rn(a0[m]);
          if (residual != nullptr)
            h = __float2half_rn(__half2float(h) + __half2float(residual[(long long)m * n_out + j]));
        }
        out[(long long)m * n_out + j] = h;
      }
    }
  }
}

int gemv(const void* x, const void* W, int64_t ldw, void* out, const void* residual, int M, int N, int K, int mode,
         cudaStream_t stream) {
  SB_REQUIRE(M >= 1 && M <= GEMV_MAXM, "gemv: M=%d outside [1,%d]", M, GEMV_MAXM);
  SB_REQUIRE(K % 8 == 0 && ldw % 8 == 0, "gemv: K and ldw must be multiples of 8");
  SB_REQUIRE(mode == 0 || (mode == 1 && N % 256 == 0 && residual == nullptr), "gemv: bad mode/shape");
  const size_t smem = (size_t)M * K * 2;
  SB_REQUIRE(smem <= 200 * 1024, "gemv: activation rows do not fit shared memory (M=%d K=%d)", M, K);
  const int n_out = mode == 1 ? N / 2 : N;
  int blocks = (n_out + 7) / 8;
  const int cap = num_sms() * 4;
  if (blocks > cap) blocks = cap;
  const __half* xp = static_cast<const __half*>(x);
  const __half* wp = static_cast<const __half*>(W);
  const __half* rp = static_cast<const __half*>(residual);
  __half* op = static_cast<__half*>(out);
#define SB_GEMV(M_, MD_)                                                                                   \
  if (M == M_ && mode == MD_) {                                                                            \
    auto kern = gemv_kernel<M_, MD_>;                                                                      \
    if (smem > 48 * 1024) SB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<blocks, 256, smem, stream>>>(xp, wp, ldw, op, rp, N, K);                                        \
    SB_LAUNCH_CHECK();                                                                                     \
    return 0;                                                                                              \
  }
  SB_GEMV(1, 0) SB_GEMV(2, 0) SB_GEMV(3, 0) SB_GEMV(4, 0)
  SB_GEMV(1, 1) SB_GEMV(2, 1) SB_GEMV(3, 1) SB_GEMV(4, 1)
#undef SB_GEMV
  set_error("gemv: unsupported configuration");
  return SEEDB200_ERR_UNSUPPORTED;
}

// ----------------------------------------------------------------------------
// Decode attention: q [B,H,D] (one token), caches [B,H,max_seq,D], D = 128.
// Kernel 1: CTA (split, h, b) -> 4 warps walk keys split*chunk .. ; lane owns 4 dims; partial (m, l, o).
// Kernel 2: merge the splits.
// ----------------------------------------------------------------------------
constexpr int DA_D = 128;
constexpr int DA_MAX_SPLITS = 32;

__global__ void __launch_bounds__(128)
decode_attn_partial(const __half* __restrict__ q, const __half* __restrict__ kc, const __half* __restrict__ vc,
                    float* __restrict__ ws, int H, int kv_len, int max_seq, int chunk, float scale_log2) {
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z, nsplit = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float s_m[4], s_l[4], s_o[4][DA_D];
  const __half* qp = q + ((long long)b * H + h) * DA_D + lane * 4;
  const float2 q01 = __half22float2(*reinterpret_cast<const __half2*>(qp));
  const float2 q23 = __half22float2(*reinterpret_cast<const __half2*>(qp + 2));
  const long long base = ((long long)b * H + h) * max_seq * DA_D;
  const int k0 = split * chunk, k1 = min(kv_len, k0 + chunk);
  float m = -INFINITY, l = 0.0f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
  for (int key = k0 + warp; key < k1; key += 4) {
    const uint2 kraw = __ldg(reinterpret_cast<const uint2*>(kc + base + (long long)key * DA_D + lane * 4));
    const uint2 vraw = __ldg(reinterpret_cast<const uint2*>(vc + base + (long long)key * DA_D + lane * 4));
    const float2 ka = __half22float2(*reinterpret_cast<const __half2*>(&kraw.x));
    const float2 kb = __half22float2(*reinterpret_cast<const __half2*>(&kraw.y));
    float s = q01.x * ka.x + q01.y * ka.y + q23.x * kb.x + q23.y * kb.y;
    s = warp_sum(s) * scale_log2;
    const float m_new = fmaxf(m, s);
    const float corr = exp2f(m - m_new);
    const float p = exp2f(s - m_new);
    const float2 va = __half22float2(*reinterpret_cast<const __half2*>(&vraw.x));
    const float2 vb = __half22float2(*reinterpret_cast<const __half2*>(&vraw.y));
    l = l * corr + p;
    o0 = o0 * corr + p * va.x; o1 = o1 * corr + p * va.y;
    o2 = o2 * corr + p * vb.x; o3 = o3 * corr + p * vb.y;
    m = m_new;
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  s_o[warp][lane * 4 + 0] = o0; s_o[warp][lane * 4 + 1] = o1;
  s_o[warp][lane * 4 + 2] = o2; s_o[warp][lane * 4 + 3] = o3;
  __syncthreads();
  if (warp == 0) {
    float mm = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    const float mu = (mm == -INFINITY) ? 0.0f : mm;
    float ll = 0.0f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float c = exp2f(s_m[w] - mu);
      ll += s_l[w] * c;
      a0 += s_o[w][lane * 4 + 0] * c; a1 += s_o[w][lane * 4 + 1] * c;
      a2 += s_o[w][lane * 4 + 2] * c; a3 += s_o[w][lane * 4 + 3] * c;
    }
    float* dst = ws + (((long long)b * H + h) * nsplit + split) * (DA_D + 2);
    if (lane == 0) { dst[0] = mm; dst[1] = ll; }
    dst[2 + lane * 4 + 0] = a0; dst[2 + lane * 4 + 1] = a1;
    dst[2 + lane * 4 + 2] = a2; dst[2 + lane * 4 + 3] = a3;
  }
}

__global__ void __launch_bounds__(DA_D)
decode_attn_merge(const float* __restrict__ ws, __half* __restrict__ out, int nsplit) {
  const long long bh = blockIdx.x;
  const int d = threadIdx.x;
  const float* src = ws + bh * nsplit * (DA_D + 2);
  float mm = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mm = fmaxf(mm, src[s * (DA_D + 2)]);
  const float mu = (mm == -INFINITY) ? 0.0f : mm;
  float ll = 0.0f, acc = 0.0f;
  for (int s = 0; s < nsplit; ++s) {
    const float c = exp2f(src[s * (DA_D + 2)] - mu);
    ll += src[s * (DA_D + 2) + 1] * c;
    acc += src[s * (DA_D + 2) + 2 + d] * c;
  }
  out[bh * DA_D + d] = __float2half_rn(ll > 0.0f ? acc / ll : 0.0f);
}

int decode_attention(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H, int D,
                     int kv_len, int max_seq, float scale, void* workspace, cudaStream_t stream) {
  SB_REQUIRE(D == DA_D, "decode_attention: head_dim %d unsupported (LLaMA uses 128)", D);
  SB_REQUIRE(kv_len >= 1 && kv_len <= max_seq, "decode_attention: kv_len %d outside [1,%d]", kv_len, max_seq);
  int nsplit = (kv_len + 255) / 256;
  if (nsplit > DA_MAX_SPLITS) nsplit = DA_MAX_SPLITS;
  const int chunk = (kv_len + nsplit - 1) / nsplit;
  dim3 grid(nsplit, H, B);
  decode_attn_partial<<<grid, 128, 0, stream>>>(static_cast<const __half*>(q), static_cast<const __half*>(k_cache),
                                                static_cast<const __half*>(v_cache), static_cast<float*>(workspace),
                                                H, kv_len, max_seq, chunk, scale * 1.4426950408889634f);
  SB_LAUNCH_CHECK();
  decode_attn_merge<<<B * H, DA_D, 0, stream>>>(static_cast<const float*>(workspace), static_cast<__half*>(out), nsplit);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sb
