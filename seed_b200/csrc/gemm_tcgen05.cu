// gemm_tcgen05.cu -- out = epilogue(A[M,K] . W[N,K]^T), fp16 operands, fp32 accumulate.
//
// This is the kernel that carries >97% of the encode FLOPs (SURVEY.md 2.4 rows E1,E4,E6,E7,
// E10-E15) and all LLaMA linears (rows L4,L8,L9,L10).  It replaces the cuBLAS calls behind
// torch.nn.functional.linear in the reference (eva_vit.py:133-135,157,60-65;
// qformer_causual.py:176-181,251-255,320-337; llama_xformer.py:186,223-225,258,718).
//
// Design (B200 / sm_100a):
//   * persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer
//     (one elected thread), warp 2 = TMEM allocator, warps 4..11 = epilogue;
//   * A and W tiles are fetched by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
//     multi-stage shared-memory ring guarded by full/empty mbarriers;
//   * accumulators live in TMEM, double buffered (2 x BN columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1;
//   * CTAS == 2: a CTA pair (cluster of 2) runs one 256 x BN tcgen05.mma.cta_group::2 tile;
//     each CTA loads its own 128 rows of A and half of the W tile, halving the shared-memory
//     and L2 traffic per FLOP;
//   * epilogue: tcgen05.ld -> registers -> bias / activation / residual with the reference's
//     fp16 rounding points -> 16-byte global stores.  SiLU-gate mode reads the gate and up
//     halves of the same accumulator tile (llama_xformer.py:186).
#include <stdio.h>

#include "common.cuh"

namespace sb {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 384;      // 4 control warps + 8 epilogue warps
constexpr int GEMM_EPI_WARP0 = 4;
constexpr int GEMM_EPI_THREADS = 256;

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles;      // m_tiles counts CTAS*128-row tiles
  const __half* bias;
  const __half* residual;
  long long ldr;
  __half* out;
  long long ldo;
  int act;
  int row_group, row_stride, row_offset;
  int res_mod, res_offset;
  const float2* ln_stats;    // LayerNorm folded into this GEMM (see seedb200_gemm_desc.ln_stats): per-row (mean, rstd),
  const float* ln_c;         // per-column c[n] = sum_k W'[n,k] and b'[n] = sum_k W[n,k] beta[k] + bias[n]:
  const float* ln_b;         //   out = rstd * (acc - mean * c) + b'
  int tile_shift;            // round r of the persistent schedule hands unit u the tile r*units + (u + r*tile_shift) % units:
                             // with a cheap tail column the plain round robin (shift 0) gives some units all the cheap
                             // tiles and others none whenever units % n_tiles shares a factor with n_tiles
  int sched;                 // 0: round robin (above).  1: balanced tail -- the full-width tiles go round robin over the units
                             // first, then the units that got one full tile fewer take the narrow tail tiles; with few tiles
                             // per unit (M = 2048 prefill: 128 tiles of 256 x 256 on 74 CTA pairs = 2 rounds, the second
                             // 27 % empty) a narrower tile + this order lands every unit within ~1 % of the mean
  float2* row_moments;       // staged epilogue only: (sum, sum of squares) per 64-column group of the stored output rows
  int out_tma;               // 1: the epilogue parks 32 x 32 output boxes in shared memory and TMA-stores them (tmap_o)
  int res_tma;               // 1: ... and the residual boxes arrive by TMA as well (tmap_r), two boxes ahead
  int tail_w;                // > 0: the last n-tile is only tail_w (< BN) columns wide -- loaded through the tail tensor
                             // map, multiplied with a narrower UMMA and read out chunk-limited, so a ragged N (1408 =
                             // 5.5 x 256, 40194 = 157 x 256 + 2) costs its columns, not a whole tile
};

// Persistent schedule: the tile (mt * n_tiles + nt) of unit `unit`'s round-th iteration, >= m_tiles * n_tiles when the
// unit is done.  sched 0: round robin with a per-round rotation (GemmParams::tile_shift).  sched 1: balanced tail.
__host__ __device__ inline int sched_tile(int sched, int round, int unit, int units, int m_tiles, int n_tiles, int tile_shift) {
  const int total_tiles = m_tiles * n_tiles;
  if (sched == 0) {
    const int t = round * units + (unit + round * tile_shift) % units;
    return t < total_tiles ? t : total_tiles;
  }
  const int ncol_full = n_tiles - 1;
  const int F = m_tiles * ncol_full;               // full-width tiles; the m_tiles tiles of the last column come last
  const int q = F / units, r = F % units;
  const int nf = q + (unit < r ? 1 : 0);
  if (round < nf) {
    const int f = round * units + unit;
    return (f / ncol_full) * n_tiles + (f % ncol_full);
  }
  if (r != 0 && unit < r) return total_tiles;      // already has one full tile more than the others
  const int S = (r == 0) ? units : units - r;
  const int t = (unit - (r == 0 ? 0 : r)) + (round - nf) * S;
  return t < m_tiles ? t * n_tiles + ncol_full : total_tiles;
}

// KSUB: 64-wide K sub-blocks per pipeline stage.  KSUB = 2 halves the per-stage fixed cost in the MMA issuer
// (one mbarrier wait + one tcgen05.commit per 8 MMAs instead of per 4), which is what bounded the
// short-N tiles (ncu: tensor pipe 64% active with neither operands nor the epilogue late).
template <int BN, int CTAS, int KSUB = 2>
struct GemmCfg {
  static constexpr int LOAD_N = BN / CTAS;
  static constexpr int A_SUB = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;      // one 64-wide sub-block of A
  static constexpr int B_SUB = LOAD_N * GEMM_BLOCK_K * 2;
  static constexpr int A_BYTES = KSUB * A_SUB;
  static constexpr int B_BYTES = KSUB * B_SUB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // Epilogue staging: each of the 8 epilogue warps owns one 32-row x 32-column fp16 box (2 KB, 64-byte rows, SWIZZLE_64B)
  // that it fills with 16-byte shared-memory stores and hands to a TMA store -- a direct 16-byte global store per thread
  // touches 32 different lines per instruction, and at 4096 such wavefronts per 128 x 256 tile (8192 with the residual
  // loads) the SM's load/store unit, not the tensor pipe, set the tile time of the K = 1408 shapes (ncu: LSU data-pipe
  // 41-50 % busy over the whole kernel, tensor pipe 58-74 %).  CTA pairs with 64-deep stages also have room for two
  // residual boxes per warp, loaded by TMA two boxes ahead.
  static constexpr bool STAGED_OUT = (BN % 64 == 0);
  static constexpr bool STAGED_RES = STAGED_OUT && CTAS == 2 && KSUB == 1;
  static constexpr int OUT_STAGE_BYTES = STAGED_OUT ? 8 * 2048 : 0;
  static constexpr int RES_STAGE_BYTES = STAGED_RES ? 2 * 8 * 2048 : 0;
  static constexpr int STAGES_RAW = (220 * 1024 - OUT_STAGE_BYTES - RES_STAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_STRIDE = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16 + 2 * 256 * 2 + 2 * 2 * 256 * 4 + 16 * 8;   // barriers, tmem slot, bias stage, LN-fold c / b' stages, residual-box barriers
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + OUT_STAGE_BYTES + RES_STAGE_BYTES + BAR_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  // keep one CTA per SM (TMEM is allocated per CTA): request more than half of the SM's smem
  static constexpr int SMEM_REQUEST = SMEM_BYTES < 120 * 1024 ? 120 * 1024 : SMEM_BYTES;
  static_assert(B_SUB % 1024 == 0, "W stage must keep 1024-byte alignment for SWIZZLE_128B");
  static_assert(BN % 16 == 0 && BN <= 256, "invalid UMMA N");
  static_assert(STAGES >= 2, "pipeline too shallow");
};

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float gelu_erf(float x) {
  // 0.5 x (1 + erf(x / sqrt 2)) = 0.5 (x + |x| erf(|x| / sqrt 2)); erf by Abramowitz-Stegun 7.1.26
  // (|err| < 1.5e-7, far below fp16 resolution); ~15 issue slots per element, 2 of them MUFU
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(ax * ax * (-0.5f * 1.4426950408889634f));   // exp(-(|x|/sqrt2)^2)
  const float erf_abs = fmaf(-p, e, 1.0f);
  return 0.5f * fmaf(ax, erf_abs, x);
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case SEEDB200_ACT_GELU: return gelu_erf(x);
    case SEEDB200_ACT_TANH: return tanhf(x);
    case SEEDB200_ACT_RELU: return fmaxf(x, 0.0f);
    default: return x;
  }
}

// One 16-column chunk of one output row: bias -> fp16 -> act -> fp16 -> (+residual) -> fp16 -> store.
// `bias_s` points at this chunk's 16 bias values in shared memory (nullptr: no bias); `res` holds the
// chunk's 16 residual values when `res_loaded` (prefetched one chunk ahead by the caller).
template <int MODE>
__device__ __forceinline__ void epilogue_store16(const GemmParams& p, const uint32_t (&acc)[16],
                                                 const uint32_t (&acc2)[16], const __half* bias_s,
                                                 const __half* res_row, __half* out_row, bool res_loaded,
                                                 const uint4& res0, const uint4& res1, int n0, int n_limit,
                                                 const float* lnc_s = nullptr, const float* lnb_s = nullptr,
                                                 float ln_mean = 0.0f, float ln_rstd = 1.0f) {
  if (n0 >= n_limit) return;
  __half h[16];
  if constexpr (MODE == 1) {
    // SiLU-gate: silu(fp16(gate)) rounded to fp16, times fp16(up), rounded (llama_xformer.py:186)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float g = __half2float(__float2half_rn(__uint_as_float(acc[j])));
      const float u = __half2float(__float2half_rn(__uint_as_float(acc2[j])));
      const float s = __half2float(__float2half_rn(g * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * g))));
      h[j] = __float2half_rn(s * u);
    }
  } else {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
    if (lnc_s != nullptr) {
      // LayerNorm folded into the GEMM: acc = sum_k W'[n,k] x[m,k] with W' = W diag(gamma), so
      // W LN(x) + bias = rstd * (acc - mean * c[n]) + b'[n]   (fp32, one rounding to fp16 below)
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(lnc_s + j);
        const float4 b4 = *reinterpret_cast<const float4*>(lnb_s + j);
        v[j + 0] = fmaf(ln_rstd, fmaf(-ln_mean, c4.x, v[j + 0]), b4.x);
        v[j + 1] = fmaf(ln_rstd, fmaf(-ln_mean, c4.y, v[j + 1]), b4.y);
        v[j + 2] = fmaf(ln_rstd, fmaf(-ln_mean, c4.z, v[j + 2]), b4.z);
        v[j + 3] = fmaf(ln_rstd, fmaf(-ln_mean, c4.w, v[j + 3]), b4.w);
      }
    }
    if (bias_s != nullptr) {
      const uint4 b0 = *reinterpret_cast<const uint4*>(bias_s);
      const uint4 b1 = *reinterpret_cast<const uint4*>(bias_s + 8);
      const __half2* bh0 = reinterpret_cast<const __half2*>(&b0);
      const __half2* bh1 = reinterpret_cast<const __half2*>(&b1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f0 = __half22float2(bh0[j]);
        const float2 f1 = __half22float2(bh1[j]);
        v[2 * j] += f0.x; v[2 * j + 1] += f0.y;
        v[8 + 2 * j] += f1.x; v[8 + 2 * j + 1] += f1.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(v[j]);
    if (p.act == SEEDB200_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(gelu_erf(__half2float(h[j])));
    } else if (p.act != SEEDB200_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(apply_act(__half2float(h[j]), p.act));
    }
  }

  if (res_row != nullptr) {
    if (res_loaded) {
      const __half* rh0 = reinterpret_cast<const __half*>(&res0);
      const __half* rh1 = reinterpret_cast<const __half*>(&res1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        h[j] = __float2half_rn(__half2float(h[j]) + __half2float(rh0[j]));
        h[8 + j] = __float2half_rn(__half2float(h[8 + j]) + __half2float(rh1[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + j < n_limit) h[j] = __float2half_rn(__half2float(h[j]) + __half2float(res_row[n0 + j]));
    }
  }

  __half* op = out_row + n0;
  if ((n0 + 16 <= n_limit) && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
    uint4 o0, o1;
    __half* oh0 = reinterpret_cast<__half*>(&o0);
    __half* oh1 = reinterpret_cast<__half*>(&o1);
#pragma unroll
    for (int j = 0; j < 8; ++j) { oh0[j] = h[j]; oh1[j] = h[8 + j]; }
    *reinterpret_cast<uint4*>(op) = o0;
    *reinterpret_cast<uint4*>(op + 8) = o1;
  } else if ((n0 + 16 <= n_limit) && ((reinterpret_cast<uintptr_t>(op) & 3) == 0)) {
    // rows that are only 4-byte aligned (lm_head: V = 40194 columns): half2 stores instead of 16 scalar ones
#pragma unroll
    for (int j = 0; j < 8; ++j) reinterpret_cast<__half2*>(op)[j] = __halves2half2(h[2 * j], h[2 * j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < n_limit) op[j] = h[j];
  }
}

// The arithmetic of epilogue_store16 for MODE 0 without the store: 16 accumulator columns of one row -> 16 fp16
// values (two 16-byte words), same rounding points.  `res0/res1` are the row's 16 residual values when has_res.
__device__ __forceinline__ void epilogue_compute16(const GemmParams& p, const uint32_t* acc, const __half* bias_s,
                                                   const float* lnc_s, const float* lnb_s, float ln_mean, float ln_rstd,
                                                   bool has_res, const uint4& res0, const uint4& res1, uint4& o0, uint4& o1) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  if (lnc_s != nullptr) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 c4 = *reinterpret_cast<const float4*>(lnc_s + j);
      const float4 b4 = *reinterpret_cast<const float4*>(lnb_s + j);
      v[j + 0] = fmaf(ln_rstd, fmaf(-ln_mean, c4.x, v[j + 0]), b4.x);
      v[j + 1] = fmaf(ln_rstd, fmaf(-ln_mean, c4.y, v[j + 1]), b4.y);
      v[j + 2] = fmaf(ln_rstd, fmaf(-ln_mean, c4.z, v[j + 2]), b4.z);
      v[j + 3] = fmaf(ln_rstd, fmaf(-ln_mean, c4.w, v[j + 3]), b4.w);
    }
  }
  if (bias_s != nullptr) {
    const uint4 b0 = *reinterpret_cast<const uint4*>(bias_s);
    const uint4 b1 = *reinterpret_cast<const uint4*>(bias_s + 8);
    const __half2* bh0 = reinterpret_cast<const __half2*>(&b0);
    const __half2* bh1 = reinterpret_cast<const __half2*>(&b1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f0 = __half22float2(bh0[j]);
      const float2 f1 = __half22float2(bh1[j]);
      v[2 * j] += f0.x; v[2 * j + 1] += f0.y;
      v[8 + 2 * j] += f1.x; v[8 + 2 * j + 1] += f1.y;
    }
  }
  __half h[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(v[j]);
  if (p.act == SEEDB200_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(gelu_erf(__half2float(h[j])));
  } else if (p.act != SEEDB200_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(apply_act(__half2float(h[j]), p.act));
  }
  if (has_res) {
    const __half* rh0 = reinterpret_cast<const __half*>(&res0);
    const __half* rh1 = reinterpret_cast<const __half*>(&res1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = __float2half_rn(__half2float(h[j]) + __half2float(rh0[j]));
      h[8 + j] = __float2half_rn(__half2float(h[8 + j]) + __half2float(rh1[j]));
    }
  }
  __half* oh0 = reinterpret_cast<__half*>(&o0);
  __half* oh1 = reinterpret_cast<__half*>(&o1);
#pragma unroll
  for (int j = 0; j < 8; ++j) { oh0[j] = h[j]; oh1[j] = h[8 + j]; }
}

__device__ __forceinline__ void gm_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld that names the registers it completes, so their uses cannot be scheduled above it while the next
// box's load is already in flight
__device__ __forceinline__ void gm_tmem_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

template <int BN, int CTAS, int MODE, int KSUB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_bt, const __grid_constant__ CUtensorMap tmap_o,
                    const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  using Cfg = GemmCfg<BN, CTAS, KSUB>;
  constexpr int STAGE_K = GEMM_BLOCK_K * KSUB;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned stage bases (same offset in both CTAs of a pair)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * Cfg::A_BYTES;
  const uint32_t ostage_base = smem_base + STAGES * Cfg::STAGE_BYTES;          // [8 warps][2 KB] output boxes
  const uint32_t rstage_base = ostage_base + Cfg::OUT_STAGE_BYTES;             // [8 warps][2][2 KB] residual boxes
  const uint32_t bar_base = rstage_base + Cfg::RES_STAGE_BYTES;
  const uint32_t full_bar = bar_base;                   // [STAGES]
  const uint32_t empty_bar = bar_base + STAGES * 8;     // [STAGES]
  const uint32_t tfull_bar = bar_base + 2 * STAGES * 8; // [2]
  const uint32_t tempty_bar = tfull_bar + 16;           // [2]
  const uint32_t tmem_slot = tempty_bar + 16;           // uint32
  const uint32_t bias_off = tmem_slot + 16;             // [2][256] halves
  const uint32_t lnc_off = bias_off + 2 * 256 * 2;      // [2][256] floats
  const uint32_t lnb_off = lnc_off + 2 * 256 * 4;       // [2][256] floats
  const uint32_t res_bar = lnb_off + 2 * 256 * 4;       // [8 warps][2] residual boxes landed
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  if constexpr (CTAS == 2) cluster_sync_all();  // both CTAs resident before the paired TMEM allocation

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tail_w > 0) tma_prefetch_desc(&tmap_bt);
    if (p.out_tma) tma_prefetch_desc(&tmap_o);
    if (p.res_tma) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, CTAS);   // one arrive per CTA's producer (leader's copy is the one waited on)
      mbar_init(empty_bar + 8 * s, 1);     // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, CTAS * (GEMM_EPI_THREADS / 32));   // one arrive per epilogue warp
    }
    for (int i = 0; i < 16; ++i) mbar_init(res_bar + 8 * i, 1);
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc<CTAS>(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int num_kb = (p.K + STAGE_K - 1) / STAGE_K;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int unit = (CTAS == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int units = (CTAS == 2) ? (gridDim.x >> 1) : gridDim.x;
  auto tile_of = [&](int round) -> int {
    return sched_tile(p.sched, round, unit, units, p.m_tiles, p.n_tiles, p.tile_shift);
  };
  const int tile0 = tile_of(0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int round = 0;; ++round) {
        const int tile = tile_of(round);
        if (tile >= total_tiles) break;
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        const int m_idx = (mt * CTAS + (int)cta_rank) * GEMM_BLOCK_M;
        const bool tail_tile = p.tail_w > 0 && nt == p.n_tiles - 1;
        const int load_n = tail_tile ? p.tail_w / CTAS : Cfg::LOAD_N;       // W rows this CTA loads per k sub-block
        const CUtensorMap* tb = tail_tile ? &tmap_bt : &tmap_b;
        const uint32_t stage_tx = (uint32_t)(Cfg::A_BYTES + KSUB * load_n * GEMM_BLOCK_K * 2);
        const int n_idx = nt * BN + (int)cta_rank * load_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_relaxed(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_a + stage * Cfg::A_BYTES;
          const uint32_t sb_ = smem_b + stage * Cfg::B_BYTES;
          if constexpr (CTAS == 1) {
            mbar_arrive_expect_tx(full_bar + 8 * stage, stage_tx);
#pragma unroll
            for (int ks = 0; ks < KSUB; ++ks) {
              tma_load_2d(sa + ks * Cfg::A_SUB, &tmap_a, full_bar + 8 * stage, kb * STAGE_K + ks * GEMM_BLOCK_K, m_idx);
              tma_load_2d(sb_ + ks * Cfg::B_SUB, tb, full_bar + 8 * stage, kb * STAGE_K + ks * GEMM_BLOCK_K, n_idx);
            }
          } else {
            const uint32_t lead_bar = mapa_shared(full_bar + 8 * stage, 0);
            if (leader) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * stage_tx);
#pragma unroll
            for (int ks = 0; ks < KSUB; ++ks) {
              tma_load_2d_2cta(sa + ks * Cfg::A_SUB, &tmap_a, lead_bar, kb * STAGE_K + ks * GEMM_BLOCK_K, m_idx);
              tma_load_2d_2cta(sb_ + ks * Cfg::B_SUB, tb, lead_bar, kb * STAGE_K + ks * GEMM_BLOCK_K, n_idx);
            }
            if (!leader) mbar_arrive_cluster(lead_bar);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc_full = make_idesc_f16(GEMM_BLOCK_M * CTAS, BN);
      const uint32_t idesc_tail = make_idesc_f16(GEMM_BLOCK_M * CTAS, p.tail_w > 0 ? p.tail_w : BN);
      int stage = 0; uint32_t phase = 0; int iter = 0;
      for (int round = 0;; ++round, ++iter) {
        const int tile = tile_of(round);
        if (tile >= total_tiles) break;
        const int as = iter & 1;
        const uint32_t idesc = (p.tail_w > 0 && (tile % p.n_tiles) == p.n_tiles - 1) ? idesc_tail : idesc_full;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(tempty_bar + 8 * as, aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < KSUB; ++ks) {
            const uint64_t adesc = make_smem_desc_sw128(smem_a + stage * Cfg::A_BYTES + ks * Cfg::A_SUB);
            const uint64_t bdesc = make_smem_desc_sw128(smem_b + stage * Cfg::B_BYTES + ks * Cfg::B_SUB);
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              // advance 16 halves = 32 bytes inside the 128-byte swizzle atom: +2 in 16-byte units
              umma_f16<CTAS>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | ks | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit<CTAS>(empty_bar + 8 * stage);            // frees the smem slot (both CTAs)
          if (kb == num_kb - 1) umma_commit<CTAS>(tfull_bar + 8 * as);  // accumulator ready
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (CTAS == 2) {
        // drain: make sure every remote tmem-empty arrive has landed before teardown
        for (int back = 1; back <= 2 && back <= iter; ++back) {
          const int it = iter - back;
          mbar_wait(tempty_bar + 8 * (it & 1), (it >> 1) & 1);
        }
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ===================== epilogue =====================
    const int ew = warp - GEMM_EPI_WARP0;      // 0..7
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half_id = ew >> 2;               // which half of the column chunks
    const int etid = threadIdx.x - GEMM_EPI_WARP0 * 32;   // 0..255
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t lead_tempty0 = (CTAS == 2) ? mapa_shared(tempty_bar, 0) : tempty_bar;
    __half* bias_smem = reinterpret_cast<__half*>(smem_gen + (bias_off - smem_base));
    constexpr int NCHUNK = (MODE == 1) ? (BN / 32) : (BN / 16);   // 16-column output chunks per tile
    const int n_limit = (MODE == 1) ? p.N / 2 : p.N;
    float* lnc_smem = reinterpret_cast<float*>(smem_gen + (lnc_off - smem_base));
    float* lnb_smem = reinterpret_cast<float*>(smem_gen + (lnb_off - smem_base));
    const bool has_ln = (MODE == 0) && (p.ln_stats != nullptr);
    const bool has_bias = (MODE == 0) && (p.bias != nullptr) && !has_ln;
    const bool has_cols = has_bias || has_ln;          // some per-column vector is staged one tile ahead
    if (has_cols && etid < BN && tile0 < total_tiles) {
      const int n = (tile0 % p.n_tiles) * BN + etid;
      if (has_ln) {
        lnc_smem[etid] = (n < p.N) ? p.ln_c[n] : 0.0f;
        lnb_smem[etid] = (n < p.N) ? p.ln_b[n] : 0.0f;
      } else {
        bias_smem[etid] = (n < p.N) ? p.bias[n] : __float2half(0.0f);
      }
    }
    int iter = 0;
    uint32_t res_phase = 0;                    // bit s: parity to wait for on this warp's residual slot s
    for (int round = 0;; ++round, ++iter) {
      const int tile = tile_of(round);
      if (tile >= total_tiles) break;
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      // the two warps of a TMEM lane quarter split the tile's 16-column chunks; the ragged last tile has fewer
      const int nchunk = (p.tail_w > 0 && nt == p.n_tiles - 1) ? p.tail_w / 16 : NCHUNK;
      const int ch0 = (nchunk + 1) / 2;
      const int c_begin = half_id == 0 ? 0 : ch0;
      const int c_end = half_id == 0 ? ch0 : nchunk;
      const int n_tile0 = (MODE == 1) ? nt * (BN / 2) : nt * BN;
      const int m = (mt * CTAS + (int)cta_rank) * GEMM_BLOCK_M + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      // per-row pointers (output row remap / periodic residual rows, see seedb200_gemm_desc)
      long long orow = m;
      if (p.row_group > 0) orow = (long long)(m / p.row_group) * p.row_stride + (m % p.row_group) + p.row_offset;
      long long rrow = orow;
      if (p.res_mod > 0) rrow = (m % p.res_mod) + p.res_offset;
      __half* out_row = p.out + orow * p.ldo;
      const __half* res_row = (p.residual != nullptr && row_ok) ? p.residual + rrow * p.ldr : nullptr;
      const bool res_vec = (res_row != nullptr) && ((reinterpret_cast<uintptr_t>(res_row) & 15) == 0);

      // stage this tile's bias in shared memory and prefetch the first residual chunk while the MMAs of
      // the tile are still running (both are global-memory latencies that used to sit in the chunk loop)
      __half bias_next = __float2half(0.0f);
      float lnc_next = 0.0f, lnb_next = 0.0f;
      float2 ln_st = make_float2(0.0f, 1.0f);
      if (has_ln && row_ok) ln_st = p.ln_stats[m];       // (mean, rstd) of this thread's row
      if (has_cols) {
        // this tile's vectors were staged one tile ago (or in the prologue); the barrier publishes them
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int next = tile_of(round + 1);
        if (next < total_tiles && etid < BN) {     // issue the load now, consume it after the chunk loop
          const int n = (next % p.n_tiles) * BN + etid;
          if (n < p.N) {
            if (has_ln) { lnc_next = p.ln_c[n]; lnb_next = p.ln_b[n]; }
            else bias_next = p.bias[n];
          }
        }
      }
      if constexpr (MODE == 0 && Cfg::STAGED_OUT) {
        if (p.out_tma) {
          // ---------- staged read-out: 32 x 32 boxes through shared memory, TMA store (and TMA residual) ----------
          const int nb = (c_end - c_begin) >> 1;           // boxes of this warp (the host guarantees even chunk counts)
          const uint32_t obuf = ostage_base + ew * 2048;
          uint8_t* obuf_g = smem_gen + (obuf - smem_base) + lane * 64;
          const int sw = (lane >> 1) & 3;                  // SWIZZLE_64B: 16-byte chunk index ^ bits 1..2 of the row
          const int m0 = (mt * CTAS + (int)cta_rank) * GEMM_BLOCK_M + quarter * 32;
          const bool use_res = p.residual != nullptr;
          const bool res_tma = Cfg::STAGED_RES && p.res_tma != 0;
          auto res_issue = [&](int b) {                    // one thread: residual box b of this tile -> slot b & 1
            const uint32_t slot = (uint32_t)(ew * 2 + (b & 1));
            mbar_arrive_expect_tx(res_bar + 8 * slot, 2048u);
            tma_load_2d(rstage_base + slot * 2048, &tmap_r, res_bar + 8 * slot, n_tile0 + (c_begin + 2 * b) * 16, m0);
          };
          if (res_tma && use_res && lane == 0) {           // two boxes ahead, before the accumulator is even ready
            if (nb > 0) res_issue(0);
            if (nb > 1) res_issue(1);
          }
          mbar_wait_relaxed(tfull_bar + 8 * as, aphase);
          tc_fence_after();
          const uint32_t t_acc = tmem_base + as * Cfg::ACC_STRIDE + lane_addr;
          if (nb == 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CTAS == 2) mbar_arrive_cluster(lead_tempty0 + 8 * as);
              else mbar_arrive(tempty_bar + 8 * as);
            }
          }
          uint32_t ra[32], rb[32];
          float mom_s = 0.0f, mom_q = 0.0f;               // moments of this thread's row over this warp's columns
          auto box = [&](int b, uint32_t(&cur)[32], uint32_t(&nxt)[32]) {
            const int cc = c_begin + 2 * b;
            gm_tmem_ld_wait32(cur);
            if (b + 1 < nb) {
              gm_tmem_ld32(t_acc + (cc + 2) * 16, nxt);
            } else {
              // every TMEM read of this accumulator stage is done: hand it back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (CTAS == 2) mbar_arrive_cluster(lead_tempty0 + 8 * as);
                else mbar_arrive(tempty_bar + 8 * as);
              }
            }
            uint4 rs[4];
            rs[0] = rs[1] = rs[2] = rs[3] = make_uint4(0, 0, 0, 0);
            bool has_res = false;
            if (use_res) {
              if (res_tma) {
                const uint32_t slot = (uint32_t)(ew * 2 + (b & 1));
                mbar_wait_relaxed(res_bar + 8 * slot, (res_phase >> (b & 1)) & 1u);
                res_phase ^= 1u << (b & 1);
                const uint8_t* rp = smem_gen + (rstage_base + slot * 2048 - smem_base) + lane * 64;
#pragma unroll
                for (int q = 0; q < 4; ++q) rs[q] = *reinterpret_cast<const uint4*>(rp + ((q ^ sw) << 4));
                __syncwarp();
                if (lane == 0 && b + 2 < nb) res_issue(b + 2);
                has_res = true;
              } else if (row_ok && n_tile0 + cc * 16 < p.N) {      // (N % 32 == 0: a box is inside or outside)
                const __half* rr = res_row + n_tile0 + cc * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q) rs[q] = *reinterpret_cast<const uint4*>(rr + 8 * q);
                has_res = true;
              }
            }
            uint4 o[4];
            epilogue_compute16(p, &cur[0], has_bias ? bias_smem + as * 256 + cc * 16 : nullptr,
                               has_ln ? lnc_smem + as * 256 + cc * 16 : nullptr, lnb_smem + as * 256 + cc * 16, ln_st.x,
                               ln_st.y, has_res, rs[0], rs[1], o[0], o[1]);
            epilogue_compute16(p, &cur[16], has_bias ? bias_smem + as * 256 + (cc + 1) * 16 : nullptr,
                               has_ln ? lnc_smem + as * 256 + (cc + 1) * 16 : nullptr, lnb_smem + as * 256 + (cc + 1) * 16,
                               ln_st.x, ln_st.y, has_res, rs[2], rs[3], o[2], o[3]);
            if (p.row_moments != nullptr) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&o[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __half22float2(h2[j]);
                  mom_s += f.x + f.y;
                  mom_q = fmaf(f.x, f.x, mom_q);
                  mom_q = fmaf(f.y, f.y, mom_q);
                }
              }
            }
            // the previous box has left the staging buffer (its store only has to have READ shared memory)
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(obuf_g + ((q ^ sw) << 4)) = o[q];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_o, obuf, n_tile0 + cc * 16, m0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          };
          if (nb > 0) gm_tmem_ld32(t_acc + c_begin * 16, ra);
          for (int b = 0; b < nb; b += 2) {
            box(b, ra, rb);
            if (b + 1 < nb) box(b + 1, rb, ra);
          }
          if (p.row_moments != nullptr && row_ok && nb > 0) {
            const int groups = p.N >> 6;                   // 64-column groups per row
            const int g0 = (n_tile0 + c_begin * 16) >> 6, gn = (nb * 32 + 63) >> 6;
            float2* mp = p.row_moments + (long long)m * groups + g0;
            mp[0] = make_float2(mom_s, mom_q);
            for (int g = 1; g < gn && g0 + g < groups; ++g) mp[g] = make_float2(0.0f, 0.0f);
          }
          if (has_cols && etid < BN) {
            if (has_ln) { lnc_smem[(as ^ 1) * 256 + etid] = lnc_next; lnb_smem[(as ^ 1) * 256 + etid] = lnb_next; }
            else bias_smem[(as ^ 1) * 256 + etid] = bias_next;
          }
          continue;
        }
      }
      uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
      bool rn_ok = false;
      {
        const int n0 = n_tile0 + c_begin * 16;
        rn_ok = res_vec && (n0 + 16 <= n_limit);
        if (rn_ok) {
          rn0 = *reinterpret_cast<const uint4*>(res_row + n0);
          rn1 = *reinterpret_cast<const uint4*>(res_row + n0 + 8);
        }
      }

      mbar_wait_relaxed(tfull_bar + 8 * as, aphase);     // a whole tile of MMAs away: park, do not spin
      tc_fence_after();
      const uint32_t t_acc = tmem_base + as * Cfg::ACC_STRIDE + lane_addr;
      if (c_begin >= c_end) {       // (one-chunk tail tile: this warp has nothing to read, but still releases the stage)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTAS == 2) mbar_arrive_cluster(lead_tempty0 + 8 * as);
          else mbar_arrive(tempty_bar + 8 * as);
        }
      }
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r0[16], r1[16];
        tmem_ld16(t_acc + c * 16, r0);
        if constexpr (MODE == 1) tmem_ld16(t_acc + BN / 2 + c * 16, r1);
        const uint4 rc0 = rn0, rc1 = rn1;
        const bool rc_ok = rn_ok;
        if (c + 1 < c_end) {   // prefetch the next chunk's residual
          const int n1 = n_tile0 + (c + 1) * 16;
          rn_ok = res_vec && (n1 + 16 <= n_limit);
          if (rn_ok) {
            rn0 = *reinterpret_cast<const uint4*>(res_row + n1);
            rn1 = *reinterpret_cast<const uint4*>(res_row + n1 + 8);
          }
        }
        tmem_ld_wait();
        if (c == c_end - 1) {
          // all TMEM reads of this accumulator stage are done (tcgen05.wait::ld is warp-collective):
          // one lane per warp hands the stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTAS == 2) mbar_arrive_cluster(lead_tempty0 + 8 * as);
            else mbar_arrive(tempty_bar + 8 * as);
          }
        }
        if (row_ok)
          epilogue_store16<MODE>(p, r0, r1, has_bias ? bias_smem + as * 256 + c * 16 : nullptr, res_row, out_row,
                                 rc_ok, rc0, rc1, n_tile0 + c * 16, n_limit,
                                 has_ln ? lnc_smem + as * 256 + c * 16 : nullptr, lnb_smem + as * 256 + c * 16, ln_st.x,
                                 ln_st.y);
      }
      // stage the next tile's vectors (other accumulator stage: nobody reads it until the next barrier)
      if (has_cols && etid < BN) {
        if (has_ln) { lnc_smem[(as ^ 1) * 256 + etid] = lnc_next; lnb_smem[(as ^ 1) * 256 + etid] = lnb_next; }
        else bias_smem[(as ^ 1) * 256 + etid] = bias_next;
      }
    }
    if (p.out_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CTAS>(tmem_base, Cfg::TMEM_COLS);
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
int get_option(const char* key);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 [rows, cols] row-major with leading dimension ld (elements); box = box_rows x 64 columns, 128B swizzle
static int make_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return SEEDB200_ERR_CUDA;
  }
  SB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "gemm: operand pointer %p not 16-byte aligned", ptr);
  SB_REQUIRE((ld * 2) % 16 == 0, "gemm: leading dimension %lld (elements) is not a multiple of 8", (long long)ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)GEMM_BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box_rows=%d)", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_rows);
    return SEEDB200_ERR_CUDA;
  }
  return 0;
}

// fp16 [rows, cols] row-major, 32 x 32 boxes with 64-byte rows (SWIZZLE_64B): the epilogue's output / residual boxes
static int make_box_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return SEEDB200_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (epilogue box) failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
    return SEEDB200_ERR_CUDA;
  }
  return 0;
}

// Host side of the persistent schedule: tile grid, ragged tail width, units and the round-robin rotation.
struct TileSchedule { int m_tiles, n_tiles, tail_w, units, tile_shift, sched; };

static TileSchedule make_schedule(int M, int N, int bn, int ctas, int mode, int sched, int sms, bool tail_opt) {
  TileSchedule t;
  t.m_tiles = (M + GEMM_BLOCK_M * ctas - 1) / (GEMM_BLOCK_M * ctas);
  t.n_tiles = (N + bn - 1) / bn;
  // ragged N: the last n-tile is loaded / multiplied / read out at its own width (rounded up to what UMMA and the
  // 8-row core-matrix groups allow: 16 columns per CTA of the pair)
  t.tail_w = 0;
  if (mode == 0 && N % bn != 0 && tail_opt) {
    const int q = 16 * ctas;
    t.tail_w = (N % bn + q - 1) / q * q;
    if (t.tail_w >= bn) t.tail_w = 0;
  }
  const int tiles = t.m_tiles * t.n_tiles;
  t.units = sms / ctas;
  if (t.units > tiles) t.units = tiles;
  if (t.units < 1) t.units = 1;
  t.sched = (sched == 1 && t.n_tiles >= 2) ? 1 : 0;
  t.tile_shift = 0;
  if (t.sched == 0 && t.tail_w > 0 && t.n_tiles > 1) {
    // advance of a unit's n-tile index per round = (units + shift) mod n_tiles: make it coprime with n_tiles so that
    // every unit meets the cheap tail column once every n_tiles rounds
    auto gcd = [](int a, int b) { while (b) { const int x = a % b; a = b; b = x; } return a; };
    for (int sft = 0; sft < t.n_tiles; ++sft)
      if (gcd((t.units + sft) % t.n_tiles, t.n_tiles) == 1) { t.tile_shift = sft; break; }
  }
  return t;
}

template <int BN, int CTAS, int MODE, int KSUB>
static int launch_gemm(const seedb200_gemm_desc& d, cudaStream_t stream, int sched) {
  using Cfg = GemmCfg<BN, CTAS, KSUB>;
  static bool attr_set_dev[SB_MAX_DEVICES] = {};   // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_dev[cur_device()];
  auto kern = gemm_tcgen05_kernel<BN, CTAS, MODE, KSUB>;
  if (!attr_set) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_REQUEST));
    attr_set = true;
  }
  CUtensorMap ta, tb, tbt;
  SB_PROPAGATE(make_tmap(&ta, d.A, d.M, d.K, d.lda, GEMM_BLOCK_M));
  SB_PROPAGATE(make_tmap(&tb, d.W, d.N, d.K, d.ldw, Cfg::LOAD_N));
  const TileSchedule ts = make_schedule(d.M, d.N, BN, CTAS, MODE, sched, num_sms(), get_option("gemm_tail") != 0);
  const int tail_w = ts.tail_w;
  if (tail_w > 0) SB_PROPAGATE(make_tmap(&tbt, d.W, d.N, d.K, d.ldw, tail_w / CTAS));
  else tbt = tb;

  // staged read-out (see GemmCfg): plain row-major output, every epilogue warp owns whole 32-column boxes
  CUtensorMap to, tr;
  to = tb; tr = tb;
  int out_tma = 0, res_tma = 0;
  if (MODE == 0 && Cfg::STAGED_OUT && d.row_group == 0 && d.res_mod == 0 && get_option("gemm_out_tma") != 0 &&
      (tail_w == 0 || tail_w % 64 == 0) && d.ldo % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 &&
      (d.residual == nullptr ||
       (d.N % 32 == 0 && d.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(d.residual) & 15) == 0))) {
    out_tma = 1;
    SB_PROPAGATE(make_box_tmap(&to, d.out, d.M, d.N, d.ldo));
    if (Cfg::STAGED_RES && d.residual != nullptr) {
      res_tma = 1;
      SB_PROPAGATE(make_box_tmap(&tr, d.residual, d.M, d.N, d.ldr));
    }
  }

  GemmParams p;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.m_tiles = ts.m_tiles;
  p.n_tiles = ts.n_tiles;
  p.bias = static_cast<const __half*>(d.bias);
  p.residual = static_cast<const __half*>(d.residual);
  p.ldr = d.ldr;
  p.out = static_cast<__half*>(d.out);
  p.ldo = d.ldo;
  p.act = d.act;
  p.row_group = d.row_group; p.row_stride = d.row_stride; p.row_offset = d.row_offset;
  p.res_mod = d.res_mod; p.res_offset = d.res_offset;
  p.tail_w = tail_w;
  p.out_tma = out_tma; p.res_tma = res_tma;
  p.row_moments = static_cast<float2*>(d.row_moments);
  if (d.row_moments != nullptr) {
    SB_REQUIRE(MODE == 0 && d.N % 64 == 0, "gemm: row_moments needs mode 0 and N %% 64 == 0 (N=%d)", d.N);
    if (!out_tma) {
      set_error("gemm: row_moments needs the staged epilogue (plain row-major output, 64-column-divisible tiles); "
                "this call (N=%d bn=%d) takes the direct one", d.N, BN);
      return SEEDB200_ERR_UNSUPPORTED;
    }
  }
  p.tile_shift = ts.tile_shift;
  p.sched = ts.sched;
  p.ln_stats = static_cast<const float2*>(d.ln_stats);
  p.ln_c = static_cast<const float*>(d.ln_c);
  p.ln_b = static_cast<const float*>(d.ln_b);

  const int units = ts.units;

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * CTAS);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_REQUEST;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  profile_mark_begin(0, stream);
  SB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tbt, to, tr, p));
  profile_mark_end(0, stream, 2.0 * (double)d.M * (double)d.N * (double)d.K);
  count_launch();
  return 0;
}

static int pick_bn(int N, int mode, int ctas) {
  if (mode == 1) return 256;
  if (N % 256 == 0) return 256;
  // CTA pairs: 256-wide tiles plus a narrow tail tile (the ragged last column costs its own width, and the persistent
  // schedule spreads the cheap tiles over the units) beat the exactly dividing narrower tilings: N=1408 proj 1143 vs
  // 967 (BN=176) TFLOP/s, N=4224 qkv 1355-1370 vs 1328 (BN=192)
  if (ctas == 2 && N >= 1024) return 256;
  if (N % 192 == 0) return 192;
  if (N % 176 == 0) return 176;
  if (N % 128 == 0) return 128;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
}

int get_option(const char* key);

// Largest number of output columns any unit of the persistent schedule works through.
static int plan_max_cols(int M, int N, int bn, int ctas, int sched, int sms) {
  const TileSchedule t = make_schedule(M, N, bn, ctas, 0, sched, sms, true);
  const int tiles = t.m_tiles * t.n_tiles, w_last = t.tail_w > 0 ? t.tail_w : bn;
  int worst = 0;
  for (int u = 0; u < t.units; ++u) {
    int cols = 0;
    for (int round = 0;; ++round) {
      const int tile = sched_tile(t.sched, round, u, t.units, t.m_tiles, t.n_tiles, t.tile_shift);
      if (tile >= tiles) break;
      cols += (tile % t.n_tiles == t.n_tiles - 1) ? w_last : bn;
    }
    if (cols > worst) worst = cols;
  }
  return worst;
}

// One row of tiles (M <= 256 on a CTA pair: the 256-token prompt of config #5 on the 13B shapes): with 256-wide tiles
// N = 5120 keeps 20 CTA pairs busy; 128-wide tiles double that and won on the GPU (tools/llama_gemm_ab.py, r02:
// o_proj 0.048 -> 0.042 ms, down_proj 0.087 -> 0.070 ms), while N = 15360 (60 pairs busy) stays on 256.
// Narrower tiles for the M = 2048 prefill shapes (128 tiles of 256 x 256 on 74 pairs = two rounds, the second 27 %
// empty) were measured and lost: a 224-wide tile costs the tensor core as much as a 256-wide one, 192 / 128-wide
// tiles lose more per tile than the better balance returns (profiles/r02_summary.md).
static void plan_single_row(const seedb200_gemm_desc& d, int sms, int& bn, int& ctas, int& sched) {
  struct Cand { int bn; double eff; };
  static const Cand cands[] = {{256, 1.00}, {128, 0.80}};
  double best = 1e30;
  for (const Cand& c : cands) {
    const double cost = plan_max_cols(d.M, d.N, c.bn, ctas, 0, sms) / c.eff;
    // busy units matter as much as the busiest one's columns: the shapes are half HBM-bound (every W byte once)
    const int tiles = (d.N + c.bn - 1) / c.bn, units = sms / ctas;
    const double busy = tiles >= units ? 1.0 : (double)tiles / units;
    const double score = cost / (0.5 + 0.5 * busy);
    if (score < best) { best = score; bn = c.bn; }
  }
  sched = 0;
}

struct GemmPlan { int bn, ctas, sched, ksub; };
static int choose_plan(const seedb200_gemm_desc& d, int sms, GemmPlan& plan) {
  SB_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "gemm: non-positive shape M=%d N=%d K=%d", d.M, d.N, d.K);
  SB_REQUIRE(d.mode == 0 || d.mode == 1, "gemm: unknown mode %d", d.mode);
  SB_REQUIRE(d.K % 8 == 0, "gemm: K=%d must be a multiple of 8 (16-byte TMA rows)", d.K);
  if (d.ln_stats != nullptr) {
    SB_REQUIRE(d.mode == 0 && d.ln_c != nullptr && d.ln_b != nullptr && d.bias == nullptr && d.row_group == 0,
               "gemm: LayerNorm-folded mode takes ln_stats + ln_c + ln_b, no bias (it is inside ln_b), mode 0, no row remap");
  }
  if (d.mode == 1) {
    SB_REQUIRE(d.N % 256 == 0, "gemm: SiLU-gate mode needs N %% 256 == 0 (got %d)", d.N);
    SB_REQUIRE(d.bias == nullptr && d.act == 0, "gemm: SiLU-gate mode takes no bias/activation");
  }
  int ctas = d.ctas > 0 ? d.ctas : 1;
  SB_REQUIRE(ctas == 1 || ctas == 2, "gemm: ctas must be 1 or 2 (got %d)", ctas);
  if (ctas == 2 && d.M <= GEMM_BLOCK_M) ctas = 1;                 // pairs only pay off on big tiles
  int bn = d.bn > 0 ? d.bn : pick_bn(d.N, d.mode, ctas);
  // the epilogue moments need whole 64-column groups per warp: 176-wide tiles (N = 1408 on single CTAs) do not have them
  if (d.row_moments != nullptr && d.bn == 0 && bn % 64 != 0) bn = d.N % 128 == 0 ? 128 : 256;
  if (ctas == 2 && bn < 64) ctas = 1;
  int sched = 0;
  if (d.bn == 0 && d.mode == 0 && d.N >= 1024 && get_option("gemm_sched") != 0 && d.ln_stats == nullptr &&
      d.row_moments == nullptr && d.row_group == 0 && d.res_mod == 0 && get_option("gemm_tail") != 0) {
    if (d.M <= GEMM_BLOCK_M * ctas && d.N % 128 == 0) plan_single_row(d, sms, bn, ctas, sched);
  } else if (d.bn != 0 && get_option("gemm_sched") == 2) {
    sched = 1;                                           // A/B runs with an explicit tile width
  }

  // 128-deep stages pay off for CTA pairs on long reductions or exactly tiled wide outputs (measured on B200:
  // qkv/fc1/fc2 +5..10%); they lose on single CTAs (-5%).
  // option gemm_ksub: 0 = this heuristic, 1 / 2 = force.
  // Re-measured after the staged epilogue (tools/vit_gemm_capture.py --ksub, r02): the ragged 256-wide tilings gain too
  // (qkv 4224x1408: 0.607 -> 0.530 ms, proj 1408x1408: 0.226 -> 0.217 ms), so every wide CTA-pair tiling takes them.
  int ksub = (d.K > 64 && ctas == 2 && (bn >= 192 || d.K >= 4096)) ? 2 : 1;
  if (get_option("gemm_ksub") == 1) ksub = 1;
  if (get_option("gemm_ksub") == 2 && d.K > 64) ksub = 2;
  plan.bn = bn; plan.ctas = ctas; plan.sched = sched; plan.ksub = ksub;
  return 0;
}

int gemm(const seedb200_gemm_desc& d, cudaStream_t stream) {
  GemmPlan plan;
  SB_PROPAGATE(choose_plan(d, num_sms(), plan));
  SB_REQUIRE(d.A && d.W && d.out, "gemm: null operand");
  const int bn = plan.bn, ctas = plan.ctas, sched = plan.sched, ksub = plan.ksub;
#define SB_GEMM_CASE(BN_, CT_, MD_)                                                   \
  if (bn == BN_ && ctas == CT_ && d.mode == MD_) {                                    \
    if (ksub == 2) return launch_gemm<BN_, CT_, MD_, 2>(d, stream, sched);            \
    return launch_gemm<BN_, CT_, MD_, 1>(d, stream, sched);                           \
  }
  SB_GEMM_CASE(256, 1, 0) SB_GEMM_CASE(256, 2, 0)
  SB_GEMM_CASE(192, 1, 0) SB_GEMM_CASE(192, 2, 0)
  SB_GEMM_CASE(176, 1, 0) SB_GEMM_CASE(176, 2, 0)
  SB_GEMM_CASE(128, 1, 0) SB_GEMM_CASE(128, 2, 0)
  SB_GEMM_CASE(64, 1, 0)  SB_GEMM_CASE(64, 2, 0)
  SB_GEMM_CASE(32, 1, 0)
  SB_GEMM_CASE(256, 1, 1) SB_GEMM_CASE(256, 2, 1)
#undef SB_GEMM_CASE
  set_error("gemm: unsupported tile configuration bn=%d ctas=%d mode=%d", bn, ctas, d.mode);
  return SEEDB200_ERR_UNSUPPORTED;
}

}  // namespace sb

extern "C" int seedb200_gemm_plan(const seedb200_gemm_desc* d, int sms, int32_t* out9) {
  if (d == nullptr || out9 == nullptr || sms <= 0) {
    sb::set_error("seedb200_gemm_plan: null argument or sms <= 0");
    return SEEDB200_ERR_INVALID;
  }
  sb::GemmPlan plan;
  SB_PROPAGATE(sb::choose_plan(*d, sms, plan));
  const sb::TileSchedule t = sb::make_schedule(d->M, d->N, plan.bn, plan.ctas, d->mode, plan.sched, sms,
                                               sb::get_option("gemm_tail") != 0);
  const int32_t v[9] = {plan.bn, plan.ctas, t.sched, plan.ksub, t.m_tiles, t.n_tiles, t.units, t.tile_shift, t.tail_w};
  for (int i = 0; i < 9; ++i) out9[i] = v[i];
  return 0;
}

extern "C" int seedb200_gemm_schedule_tile(int sched, int round, int unit, int units, int m_tiles, int n_tiles,
                                           int tile_shift) {
  return sb::sched_tile(sched, round, unit, units, m_tiles, n_tiles, tile_shift);
}

extern "C" int seedb200_gemm(const seedb200_gemm_desc* d, void* stream) {
  if (d == nullptr) {
    sb::set_error("seedb200_gemm: null descriptor");
    return SEEDB200_ERR_INVALID;
  }
  return sb::gemm(*d, static_cast<cudaStream_t>(stream));
}
