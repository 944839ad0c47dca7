// attention_causal_tc.cu -- causal prefill attention of llama_xformer (head_dim 128) on the 5th-gen tensor cores.
//
// Replaces llama_xformer.py:240-256 (xops.memory_efficient_attention with LowerTriangularMask over the K/V that
// were just appended to the cache) for q_len >= 128.  The mma.sync kernel in attention.cu stays the path for
// short chunks and other head sizes.
//
// One persistent CTA per SM walks work items = (batch, head, pair of 128-row query tiles), longest items first
// and dealt to the CTAs in snake order so that the causal triangle balances.  Roles (416 threads):
//   warps 0-3 / 4-7   softmax + correction + epilogue of query tile 0 / 1 (one query row per thread)
//   warps 8-11        loaders (Q once per item, then K and V tiles of 128 keys, double buffered): one TMA issuing
//                     thread per operand when the layout allows it (below), else four cp.async warps
//   warp 12           tcgen05.mma issuer: S_u = Q_u K_j^T (SS form), O_u += P_u V_j (TS form: P stays in TMEM)
// TMEM (512 columns): per tile u  S [256u, 256u+128) fp32, P aliased in place (fp16 x2 / column, 64 columns),
// O [256u+128, 256u+256) fp32.  The two tiles ping-pong: while the softmax warps of one tile work, the tensor
// core runs the other tile's MMAs.  Online softmax with lazy rescaling: O and the running sum are only rescaled
// when some row's maximum grew by more than 2^8 (the stale maximum is used consistently for P and the sum, so
// the result is exact up to rounding).
//
// Operands sit in shared memory in the canonical non-swizzled UMMA layout (8 x 16-byte core matrices):
//   byte(r, c) = (r / 8) * 2048 + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2        (r = token, c = head dim)
// read K-major for Q and K (LBO 128, SBO 2048) and MN-major for V (SBO 128, LBO 2048), so V needs no transpose.
//
// TMA writes that image directly: a 4-D tensor map (8 elements = one 16-byte chunk | token rows | 16 chunks of a head |
// heads / batches) whose 8 x 8 x 16 x 1 box IS one 8-row group (2 KB); a 128-row tile is 16 such boxes on one
// mbarrier, rows past the end of the sequence are zero-filled by the out-of-bounds rule.  The cp.async loaders (2048
// 16-byte copies per tile, ~78 cycles of the SM's load/store unit per 512-byte warp instruction = ~10 k cycles per
// K + V block against ~2 k cycles of MMAs) were what bounded the kernel.
#include <cuda.h>

#include "common.cuh"
#include "ops.h"

namespace sb {

constexpr int CA_D = 128, CA_T = 128;
constexpr int CA_G = (CA_D / 8) * 128;            // 2048 bytes per 8-row group
constexpr int CA_TILE_BYTES = (CA_T / 8) * CA_G;  // 32 KB
constexpr int CA_SMEM = 6 * CA_TILE_BYTES + 256 + 128;
constexpr int CA_THREADS = 416;
constexpr int CA_TMEM_COLS = 512;
constexpr float CA_RESCALE_LOG2 = 8.0f;

struct CausalAttnParams {
  const __half* q; const __half* k; const __half* v; __half* o;
  long long q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts;
  int batch_heads, heads, nq, nk, past, pairs, items;
  float scale_log2;
  int use_tma;               // 1: Q / K / V tiles by TMA (tensor maps passed beside this struct)
  int q_span, k_span, v_span;   // tensor-map form per operand: 1 = chunks span the heads (coords: chunk = h*16, outer = b),
                                // 0 = heads in the outer dimension (coords: chunk = 0, outer = b*heads + h)
};

struct CaItem { int b, h, pair, n0, n1; };

__device__ __forceinline__ CaItem ca_item(const CausalAttnParams& p, int i) {
  CaItem it;
  it.pair = p.pairs - 1 - i / p.batch_heads;        // longest (last) pairs first
  const int bh = i - (i / p.batch_heads) * p.batch_heads;
  it.b = bh / p.heads; it.h = bh - it.b * p.heads;
  const int r0 = it.pair * 2 * CA_T, r1 = r0 + CA_T;
  it.n0 = (p.past + min(r0 + CA_T - 1, p.nq - 1)) / CA_T + 1;      // r0 < nq always
  it.n1 = r1 < p.nq ? (p.past + min(r1 + CA_T - 1, p.nq - 1)) / CA_T + 1 : 0;
  return it;
}
// snake deal: round r hands item r*G + c to CTA c (even rounds) or CTA G-1-c (odd rounds)
__device__ __forceinline__ int ca_next(int round, int cta, int grid) {
  return round * grid + ((round & 1) ? grid - 1 - cta : cta);
}

__device__ __forceinline__ uint64_t ca_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__device__ __forceinline__ void ca_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ca_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void ca_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void ca_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void ca_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ca_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ca_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ca_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t ca_pack2(float a, float b, float& sum) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  sum += f.x + f.y;
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(CA_THREADS, 1)
causal_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                           const __grid_constant__ CUtensorMap tm_v, const CausalAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sK = sQ + 2 * CA_TILE_BYTES, sV = sK + 2 * CA_TILE_BYTES;
  const uint32_t bars = sV + 2 * CA_TILE_BYTES;
  const uint32_t bar_s = bars, bar_p = bars + 16, bar_o = bars + 32;            // [2] each: per query tile
  const uint32_t q_full = bars + 48, q_empty = bars + 64;                       // [2] each
  const uint32_t k_full = bars + 80, k_empty = bars + 96, v_full = bars + 112, v_empty = bars + 128;   // [2] stages
  const uint32_t tmem_slot = bars + 144;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (tid == 0) {
    for (int u = 0; u < 2; ++u) {
      mbar_init(bar_s + 8 * u, 1); mbar_init(bar_p + 8 * u, 4); mbar_init(bar_o + 8 * u, 1);
      const uint32_t producers = p.use_tma ? 1u : 2u;     // one expect-tx arrive, or one arrive per cp.async warp
      mbar_init(q_full + 8 * u, producers); mbar_init(q_empty + 8 * u, 1);
      mbar_init(k_full + 8 * u, producers); mbar_init(k_empty + 8 * u, 1);
      mbar_init(v_full + 8 * u, producers); mbar_init(v_empty + 8 * u, 1);
    }
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<1>(tmem_slot, CA_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  constexpr uint32_t IDESC_S = make_idesc_f16(128, 128);
  constexpr uint32_t IDESC_O = make_idesc_f16(128, 128) | (1u << 16);       // B (= V) is MN-major

  if (warp >= 8 && warp < 12 && p.use_tma) {
    // ======================= loaders (TMA) =======================
    // warp 8: Q tile 0 then the K tiles; warp 10: Q tile 1 then the V tiles; one thread each
    const int isv = (warp - 8) >> 1;
    if (((warp - 8) & 1) == 0 && lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(isv ? &tm_v : &tm_k);
      const CUtensorMap* tm = isv ? &tm_v : &tm_k;
      const int span = isv ? p.v_span : p.k_span;
      auto load_tile = [&](const CUtensorMap* map, int sp, const CaItem& it, int row0, uint32_t dst, uint32_t bar) {
        const int c2 = sp ? it.h * (CA_D / 8) : 0, c3 = sp ? it.b : it.b * p.heads + it.h;
        mbar_arrive_expect_tx(bar, CA_TILE_BYTES);
#pragma unroll 4
        for (int g = 0; g < CA_T / 8; ++g) tma_load_4d(dst + g * CA_G, map, bar, 0, row0 + 8 * g, c2, c3);
      };
      uint32_t kc = 0, qc = 0;
      for (int round = 0;; ++round) {
        const int i = ca_next(round, cta, G);
        if (i >= p.items) break;
        const CaItem it = ca_item(p, i);
        const int nu = isv ? it.n1 : it.n0, nmax = max(it.n0, it.n1);
        if (nu > 0) {
          mbar_wait_relaxed(q_empty + 8 * isv, (qc & 1) ^ 1);
          load_tile(&tm_q, p.q_span, it, it.pair * 2 * CA_T + isv * CA_T, sQ + isv * CA_TILE_BYTES, q_full + 8 * isv);
          ++qc;
        }
        const uint32_t dst0 = isv ? sV : sK, full = isv ? v_full : k_full, empty = isv ? v_empty : k_empty;
        for (int j = 0; j < nmax; ++j, ++kc) {
          const uint32_t s = kc & 1;
          mbar_wait_relaxed(empty + 8 * s, ((kc >> 1) & 1) ^ 1);
          load_tile(tm, span, it, j * CA_T, dst0 + s * CA_TILE_BYTES, full + 8 * s);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 8 && warp < 12) {
    // ======================= loaders (cp.async: layouts the tensor maps cannot describe) =======================
    // warps 8,9: Q tile 0 then the K tiles (rows 0..63 / 64..127 of each); warps 10,11: Q tile 1 then the V tiles.
    const int lw = warp - 8, half = lw & 1, isv = lw >> 1;
    const int r8 = lane & 7, cq = lane >> 3;
    auto load_half = [&](const __half* src, long long ts, int row0, int nrows, uint32_t dst) {
      // 64 rows x 16 chunks of this warp's half tile; rows >= nrows are zero-filled
      const int rbase = half * 64;
#pragma unroll 2
      for (int g = 0; g < 8; ++g) {
        const int r = rbase + g * 8 + r8;
        const bool ok = row0 + r < nrows;
        const __half* rp = src + (long long)(ok ? row0 + r : 0) * ts + cq * 8;
        const uint32_t dp = dst + (uint32_t)(r >> 3) * CA_G + cq * 128 + r8 * 16;
        const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) ca_cp_async16(dp + c4 * 4 * 128, rp + c4 * 32, nb);
      }
    };
    uint32_t kc = 0, qc = 0;          // tiles loaded so far (this warp's operand), Q tiles loaded (this warp's tile)
    for (int round = 0;; ++round) {
      const int i = ca_next(round, cta, G);
      if (i >= p.items) break;
      const CaItem it = ca_item(p, i);
      const int nu = isv ? it.n1 : it.n0, nmax = max(it.n0, it.n1);
      if (nu > 0) {
        mbar_wait_relaxed(q_empty + 8 * isv, (qc & 1) ^ 1);
        load_half(p.q + it.b * p.q_bs + it.h * p.q_hs, p.q_ts, it.pair * 2 * CA_T + isv * CA_T, p.nq, sQ + isv * CA_TILE_BYTES);
        asm volatile("cp.async.wait_all;" ::: "memory");
        ca_fence_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(q_full + 8 * isv);
        ++qc;
      }
      const __half* src = isv ? p.v + it.b * p.v_bs + it.h * p.v_hs : p.k + it.b * p.k_bs + it.h * p.k_hs;
      const long long ts = isv ? p.v_ts : p.k_ts;
      const uint32_t dst0 = isv ? sV : sK, full = isv ? v_full : k_full, empty = isv ? v_empty : k_empty;
      for (int j = 0; j < nmax; ++j, ++kc) {
        const uint32_t s = kc & 1;
        mbar_wait_relaxed(empty + 8 * s, ((kc >> 1) & 1) ^ 1);
        load_half(src, ts, j * CA_T, p.nk, dst0 + s * CA_TILE_BYTES);
        asm volatile("cp.async.wait_all;" ::: "memory");
        ca_fence_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * s);
      }
    }
  } else if (warp == 12) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      uint32_t kc = 0, qc[2] = {0, 0}, pc[2] = {0, 0};
      auto issue_s = [&](int u, uint32_t ks) {
        const uint32_t qa = sQ + u * CA_TILE_BYTES, ka = sK + ks * CA_TILE_BYTES;
#pragma unroll
        for (int j = 0; j < CA_D / 16; ++j)
          umma_f16<1>(tmem + u * 256, ca_desc(qa + j * 256, 128, CA_G), ca_desc(ka + j * 256, 128, CA_G), IDESC_S, j > 0);
        umma_commit<1>(bar_s + 8 * u);
      };
      auto issue_pv = [&](int u, uint32_t vs, bool acc) {
        const uint32_t tb = tmem + u * 256, va = sV + vs * CA_TILE_BYTES;
#pragma unroll
        for (int j = 0; j < CA_T / 16; ++j)
          ca_umma_ts(tb + 128, tb + j * 8, ca_desc(va + j * 2 * CA_G, CA_G, 128), IDESC_O, (acc || j > 0) ? 1u : 0u);
        umma_commit<1>(bar_o + 8 * u);
      };
      for (int round = 0;; ++round) {
        const int i = ca_next(round, cta, G);
        if (i >= p.items) break;
        const CaItem it = ca_item(p, i);
        const int n[2] = {it.n0, it.n1};
        const int nmax = max(it.n0, it.n1);
        for (int u = 0; u < 2; ++u)
          if (n[u] > 0) { mbar_wait(q_full + 8 * u, qc[u] & 1); ++qc[u]; }
        mbar_wait(k_full + 8 * (kc & 1), (kc >> 1) & 1);
        tc_fence_after();
        for (int u = 0; u < 2; ++u)
          if (n[u] > 0) issue_s(u, kc & 1);
        umma_commit<1>(k_empty + 8 * (kc & 1));
        for (int j = 0; j < nmax; ++j) {
          const uint32_t t = kc + j;
          mbar_wait(v_full + 8 * (t & 1), (t >> 1) & 1);
          if (j + 1 < nmax) mbar_wait(k_full + 8 * ((t + 1) & 1), ((t + 1) >> 1) & 1);
          for (int u = 0; u < 2; ++u) {
            if (j >= n[u]) continue;
            mbar_wait(bar_p + 8 * u, pc[u] & 1); ++pc[u];
            tc_fence_after();
            issue_pv(u, t & 1, j > 0);
            if (j + 1 < n[u]) issue_s(u, (t + 1) & 1);
            else umma_commit<1>(q_empty + 8 * u);
          }
          umma_commit<1>(v_empty + 8 * (t & 1));
          if (j + 1 < nmax) umma_commit<1>(k_empty + 8 * ((t + 1) & 1));
        }
        kc += nmax;
      }
    }
    __syncwarp();
  } else {
    // ======================= softmax / correction / epilogue =======================
    const int quarter = warp & 3, u = warp >> 2;
    const int rl = quarter * 32 + lane;
    const uint32_t tS = tmem + u * 256 + ((uint32_t)(quarter * 32) << 16), tO = tS + 128;
    uint32_t sc = 0;                  // S tiles consumed == PV tiles issued for this query tile
    for (int round = 0;; ++round) {
      const int i = ca_next(round, cta, G);
      if (i >= p.items) break;
      const CaItem it = ca_item(p, i);
      const int nu = u == 0 ? it.n0 : it.n1;
      if (nu == 0) continue;
      const int row = it.pair * 2 * CA_T + u * CA_T + rl;
      const int limit = p.past + row;                         // last visible key of this row
      const int warp_limit = p.past + row - lane;             // ... of the warp's first row
      float m = -INFINITY, l = 0.0f;
      for (int j = 0; j < nu; ++j, ++sc) {
        mbar_wait_relaxed(bar_s + 8 * u, sc & 1);
        tc_fence_after();
        const int kb = j * CA_T;
        const bool masked = kb + CA_T - 1 > warp_limit;       // warp-uniform
        // pass 1: row maximum of the visible scores
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          ca_ld32(tS + c * 32, r);
          tmem_ld_wait();
          if (masked) {
#pragma unroll
            for (int e = 0; e < 32; ++e) if (kb + c * 32 + e <= limit) mx = fmaxf(mx, __uint_as_float(r[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
          }
        }
        const float m_new = fmaxf(m, mx * p.scale_log2);      // scale > 0
        if (j == 0) {
          m = m_new;                                          // key 0 is visible to every row: finite
        } else if (__any_sync(0xffffffffu, m_new > m + CA_RESCALE_LOG2)) {
          // correction: O *= 2^(m - m_new).  PV(j-1) retired before S(j) (same in-order pipe); the wait makes it formal.
          mbar_wait(bar_o + 8 * u, (sc - 1) & 1);
          tc_fence_after();
          const float alpha = ca_ex2(m - m_new);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            ca_ld32(tO + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
            ca_st32(tO + c * 32, r);
          }
          l *= alpha;
          m = m_new;
        }
        // pass 2: P = 2^(s * scale * log2e - m), fp16, in place over the S columns already consumed
        float sum = 0.0f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          ca_ld32(tS + c * 32, r);
          tmem_ld_wait();
          uint32_t pk[16];
          if (masked) {
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              const int key = kb + c * 32 + 2 * g;
              const float a = key <= limit ? ca_ex2(fmaf(__uint_as_float(r[2 * g]), p.scale_log2, -m)) : 0.0f;
              const float b = key + 1 <= limit ? ca_ex2(fmaf(__uint_as_float(r[2 * g + 1]), p.scale_log2, -m)) : 0.0f;
              pk[g] = ca_pack2(a, b, sum);
            }
          } else {
#pragma unroll
            for (int g = 0; g < 16; ++g)
              pk[g] = ca_pack2(ca_ex2(fmaf(__uint_as_float(r[2 * g]), p.scale_log2, -m)),
                               ca_ex2(fmaf(__uint_as_float(r[2 * g + 1]), p.scale_log2, -m)), sum);
          }
          ca_st16(tS + c * 16, pk);
        }
        l += sum;
        ca_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p + 8 * u);
      }
      // epilogue: O / l -> fp16 -> global
      mbar_wait_relaxed(bar_o + 8 * u, (sc - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      __half* og = p.o + it.b * p.o_bs + it.h * p.o_hs + (long long)row * p.o_ts;
      float unused = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        ca_ld32(tO + c * 32, r);
        tmem_ld_wait();
        if (row < p.nq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = ca_pack2(__uint_as_float(r[8 * g + 0]) * inv, __uint_as_float(r[8 * g + 1]) * inv, unused);
            o.y = ca_pack2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv, unused);
            o.z = ca_pack2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv, unused);
            o.w = ca_pack2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv, unused);
            *reinterpret_cast<uint4*>(og + c * 32 + g * 8) = o;
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<1>(tmem, CA_TMEM_COLS);
}

bool causal_attention_tc_applicable(const seedb200_attn_desc& d) {
  return d.head_dim == CA_D && d.causal != 0 && d.nq >= CA_T && d.nk >= d.nq && d.o_hs % 8 == 0 && d.o_ts % 8 == 0 &&
         d.o_bs % 8 == 0 && (reinterpret_cast<uintptr_t>(d.o) & 15) == 0 && d.scale > 0.0f;
}

int get_option(const char* key);

typedef CUresult (*EncodeTiledFnC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 4-D map of one operand [batch][head][rows][128] with element strides (bs, hs, ts): (8 elements | rows | chunks | outer).
// span = 1: the heads of a token are contiguous (hs == 128) -> the chunk dimension runs over all heads, outer = batch;
// span = 0: batch and head merge into one outer dimension (bs == heads * hs, or a single batch).
// Returns 1 when the layout fits (map written), 0 when it does not (caller keeps cp.async), < 0 on a driver error.
static int make_causal_tmap(CUtensorMap* tm, int* span, const void* base, long long bs, long long hs, long long ts,
                            int batch, int heads, int rows) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ts % 8 != 0 || hs % 8 != 0 || bs % 8 != 0 || ts < CA_D) return 0;
  cuuint64_t gdim[4], gstr[3];
  gdim[0] = 8; gdim[1] = (cuuint64_t)rows; gstr[0] = (cuuint64_t)ts * 2; gstr[1] = 16;
  if (hs == CA_D && ts >= (long long)heads * CA_D) {
    *span = 1;
    gdim[2] = (cuuint64_t)heads * (CA_D / 8); gdim[3] = (cuuint64_t)batch; gstr[2] = (cuuint64_t)(batch > 1 ? bs : ts) * 2;
  } else if (batch == 1 || bs == (long long)heads * hs) {
    *span = 0;
    gdim[2] = CA_D / 8; gdim[3] = (cuuint64_t)batch * heads; gstr[2] = (cuuint64_t)hs * 2;
  } else {
    return 0;
  }
  if (gstr[0] >= (1ull << 40) || gstr[2] >= (1ull << 40)) return 0;
  static EncodeTiledFnC fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnC>(ptr);
  }
  if (fn == nullptr) return 0;
  cuuint32_t box[4] = {8, 8, CA_D / 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("causal_attention: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return -1;
  }
  return 1;
}

int causal_attention_tc(const seedb200_attn_desc& d, cudaStream_t stream) {
  static bool attr_set_dev[SB_MAX_DEVICES] = {};   // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(causal_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CA_SMEM));
    attr_set = true;
  }
  CausalAttnParams p;
  p.q = static_cast<const __half*>(d.q); p.k = static_cast<const __half*>(d.k);
  p.v = static_cast<const __half*>(d.v); p.o = static_cast<__half*>(d.o);
  p.q_bs = d.q_bs; p.q_hs = d.q_hs; p.q_ts = d.q_ts;
  p.k_bs = d.k_bs; p.k_hs = d.k_hs; p.k_ts = d.k_ts;
  p.v_bs = d.v_bs; p.v_hs = d.v_hs; p.v_ts = d.v_ts;
  p.o_bs = d.o_bs; p.o_hs = d.o_hs; p.o_ts = d.o_ts;
  p.batch_heads = d.batch * d.heads; p.heads = d.heads; p.nq = d.nq; p.nk = d.nk; p.past = d.nk - d.nq;
  p.pairs = (d.nq + 2 * CA_T - 1) / (2 * CA_T);
  p.items = p.pairs * p.batch_heads;
  p.scale_log2 = d.scale * 1.4426950408889634f;
  // TMA loaders when all three operands fit a 4-D tensor map (the LLaMA layouts do: q rows of the projection buffer,
  // K / V rows of the [B,H,max_seq,128] caches); option causal_attention_tma = 0 keeps the cp.async warps
  CUtensorMap tq, tk, tv;
  memset(&tq, 0, sizeof(tq)); memset(&tk, 0, sizeof(tk)); memset(&tv, 0, sizeof(tv));
  p.use_tma = 0; p.q_span = p.k_span = p.v_span = 0;
  if (get_option("causal_attention_tma") != 0) {
    const int rq = make_causal_tmap(&tq, &p.q_span, d.q, d.q_bs, d.q_hs, d.q_ts, d.batch, d.heads, d.nq);
    const int rk = rq == 1 ? make_causal_tmap(&tk, &p.k_span, d.k, d.k_bs, d.k_hs, d.k_ts, d.batch, d.heads, d.nk) : 0;
    const int rv = rk == 1 ? make_causal_tmap(&tv, &p.v_span, d.v, d.v_bs, d.v_hs, d.v_ts, d.batch, d.heads, d.nk) : 0;
    if (rq < 0 || rk < 0 || rv < 0) return SEEDB200_ERR_CUDA;
    p.use_tma = (rq == 1 && rk == 1 && rv == 1) ? 1 : 0;
  }
  int grid = num_sms();
  if (grid > p.items) grid = p.items;
  profile_mark_begin(1, stream);
  causal_attention_tc_kernel<<<grid, CA_THREADS, CA_SMEM, stream>>>(tq, tk, tv, p);
  profile_mark_end(1, stream, 4.0 * (double)d.batch * d.heads * (double)d.nq * d.nk * d.head_dim * 0.5);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sb
