// attention_tc.cu -- ViT-g/14 attention (257 x 257 tokens, 16 heads x 88) on the 5th-gen tensor cores.
//
// Replaces eva_vit.py:139-156 (`attn = softmax(q*scale @ k^T); x = attn @ v`) for the shape that is 23% of the
// encode step on the mma.sync kernel (attention.cu) although it is only 2.8% of its FLOPs.
//
// One persistent CTA per SM walks (image, head) items.  Per item:
//   * all 8 compute warps stage Q, K (K-major, head_dim zero-padded 88 -> 96) and V^T (transposed on the fly
//     so that keys become the contraction dimension) into shared memory in the canonical NON-swizzled UMMA
//     layout (8 x 16-byte core matrices; LBO = 128 B along K, SBO between 8-row groups).  No swizzle is what
//     lets 88/257 be padded freely with plain stores;
//   * for each 128-row query tile: one thread issues S = Q K^T as tcgen05.mma 128x256x16 + 128x16x16 into TMEM
//     (272 fp32 columns), the compute warps (one row per thread, two warps per TMEM lane quarter splitting
//     the columns) run an exact two-pass fp32 softmax straight out of TMEM (tcgen05.ld), write P (fp16, as
//     the reference rounds it under autocast) back to shared memory in the same canonical layout, the
//     issuer runs O = P V (128x96x16 x 17) into 96 more TMEM columns, and the compute warps normalise and
//     store O.
// Synchronisation: tcgen05.commit -> mbarrier for "S ready" / "O ready", mbarrier arrives (one per compute
// warp) for "P written" / "TMEM free", fence.proxy.async between generic-proxy smem writes and UMMA reads.
#include <string.h>

#include "common.cuh"
#include "attention_tc_common.cuh"

namespace sb {

// ---- geometry of this kernel (attention_tc2.cu keeps the constants of attention_tc_common.cuh) ----
constexpr int V1_THREADS = 480;                    // 8 softmax warps, 3 helper warps, V loader, MMA issuer, row-256 warp, TMA producer
constexpr int V1_V_BYTES = 34 * VA_G;              // keys 0..271, ONE buffer (the second one became the O staging area)
constexpr int V1_DATA_BYTES = VA_K_BYTES + VA_Q0_BYTES + VA_Q1_BYTES + V1_V_BYTES;    // zero-initialised operand buffers
constexpr int V1_STAGE_ROW = VA_D * 2;             // 176 bytes: one row of O, dense (the TMA store's box is 88 x 1 x 32)
constexpr int V1_STAGE_WARP = 32 * V1_STAGE_ROW;   // 5632 bytes per softmax warp
constexpr int V1_STAGE_BYTES = 8 * V1_STAGE_WARP;
constexpr int V1_MISC_BYTES = 2 * VA_CLS_LD * 4 + 2 * VA_CLS_LD * 2 + 2 * 256 * 4 + 256;
constexpr int V1_SMEM = V1_DATA_BYTES + V1_STAGE_BYTES + V1_MISC_BYTES + 1024;
// TMEM map of one tile pipeline u (base = 256 u); everything aliases the 256 fp32 columns of S:
//   S      keys 0..255                         [0, 256)
//   P      keys 0..255 (fp16 x2 / column)      [0, 128)    chunk c of S (32 columns) -> [16 c, 16 c + 16), behind the reads
//   O      dims 0..95                          [128, 224)  one N = 96 MMA per 16 keys; dim 88 = ones-column of V = row sum
//   P      key 256 (+ 15 zero keys)            [224, 232)
constexpr int V1_O_COL = 128, V1_P256_COL = 224;

__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// TMA = true: Q and K arrive by TMA (3-D tensor maps over the packed qkv buffer: 88 elements per head slot with OOB zero
// fill up to 96, 48 head slots, rows) in the K-major swizzled layouts the GEMM uses -- a 128-byte-swizzled block of head
// dims 0..63 and a 64-byte-swizzled block of dims 64..95 -- instead of 16-byte cp.async copies into the no-swizzle
// core-matrix layout: the cp.async path cost ~20 LSU cycles per 512 bytes (tools/attn_timeline.py: 5-8 k cycles per
// operand).  V (an MN-major operand) stays on the cp.async / no-swizzle path.
//
// Who does what (profiles/r02_attention.md: the softmax warps' own serial work per item WAS the kernel's period, so
// everything that is not the 256 x 256 softmax of a tile moved off them):
//   warps 0-7   softmax of the two 128-row tiles (thread per row): max, exp2, P back to TMEM, O read-out through a
//               per-warp staging area + one TMA store of 32 rows
//   warps 8-11  the 257th token on the legacy tensor pipe (mma.sync): scores of key 256 for every query row, scores of
//               query 256 for every key, and their share of row 256's P.V; with TMA = false they first load
//               Q0 / K / Q1 / V with cp.async
//   warp 12     tcgen05.mma issuer
//   warp 13     softmax of query row 256 + its share of that row's P.V
//   warp 14     TMA producer for Q0 / K / Q1 / V (one thread).  V keeps the no-swizzle core-matrix image (it is the
//               MN-major operand): a 4-D tensor map (8 elements | rows | 11 chunks | head slots) with an 8 x 8 x 11 box
//               writes one 8-key group per copy, 33 copies per item -- the cp.async version kept the SM's load/store
//               unit busy for ~5 k cycles per item, which is what the helper warps' ldmatrix queued behind
template <bool TMA>
__global__ void __launch_bounds__(V1_THREADS, 1)
vit_attention_tc_kernel(const VitAttnParams p, const __grid_constant__ CUtensorMap tm_a64,
                        const __grid_constant__ CUtensorMap tm_a32, const __grid_constant__ CUtensorMap tm_r64,
                        const __grid_constant__ CUtensorMap tm_r32, const __grid_constant__ CUtensorMap tm_o,
                        const __grid_constant__ CUtensorMap tm_v8, const __grid_constant__ CUtensorMap tm_v1) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK0 = base, sQ0 = sK0 + VA_K_BYTES, sQ1 = sQ0 + VA_Q0_BYTES, sV0 = sQ1 + VA_Q1_BYTES;
  const uint32_t sStage = sV0 + V1_V_BYTES;
  const uint32_t misc = sStage + V1_STAGE_BYTES;
  float* s_clsb = reinterpret_cast<float*>(gen + (misc - base));    // [2][VA_CLS_LD]: scores of query row 256 (+ its sum at [VA_KP])
  __half* s_clsh = reinterpret_cast<__half*>(s_clsb + 2 * VA_CLS_LD);   // [2][VA_CLS_LD]: its probabilities, the fp16 A operand
  float* s_s256 = s_clsb + 3 * VA_CLS_LD;                           // [2][256]: score of key 256 for every query row
  const uint32_t bars = misc + 2 * VA_CLS_LD * 4 + 2 * VA_CLS_LD * 2 + 2 * 256 * 4;
  const uint32_t bar_s = bars, bar_p = bars + 16, bar_o = bars + 32, bar_free = bars + 48;      // [2] each: per tile pipeline
  const uint32_t q_full = bars + 64 /*[2]*/, q_empty = bars + 80 /*[2]*/;
  const uint32_t k_full = bars + 96, k_empty = bars + 104, v_full = bars + 112, v_empty = bars + 120;
  const uint32_t cls_bar = bars + 128;           // helper warps -> row-256 warp: scores of query 256 are in s_clsb
  const uint32_t cls_p = bars + 136;             // row-256 warp -> helper warps: probabilities of row 256 are in s_clsh
  const uint32_t s256_full = bars + 144;         // helper warps -> softmax warps: scores of key 256 are in s_s256
  const uint32_t tmem_slot = bars + 160;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  uint8_t* gQ0 = gen + (sQ0 - base);
  uint8_t* gQ1 = gen + (sQ1 - base);
  const uint8_t* gK = gen + (sK0 - base);
  // 16-byte chunk c (8 head dims) of row r of a Q / K buffer with `rows8` 8-row groups: the ldmatrix readers of the
  // 257th token see either the no-swizzle core-matrix image or the two swizzled blocks (Swizzle<3,4,3> on 128-byte
  // rows for dims 0..63, Swizzle<2,4,3> on 64-byte rows for dims 64..95)
  auto qk_chunk = [&](const uint8_t* buf, int rows8, int r, int c) -> const uint4* {
    if constexpr (TMA) {
      if (c < 8) return reinterpret_cast<const uint4*>(buf + r * 128 + ((c ^ (r & 7)) << 4));
      return reinterpret_cast<const uint4*>(buf + rows8 * 1024 + r * 64 + (((c - 8) ^ ((r >> 1) & 3)) << 4));
    } else {
      return reinterpret_cast<const uint4*>(buf + (uint32_t)(r >> 3) * VA_G + c * 128 + (r & 7) * 16);
    }
  };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    if constexpr (TMA) {
      tma_prefetch_desc(&tm_a64); tma_prefetch_desc(&tm_a32); tma_prefetch_desc(&tm_r64); tma_prefetch_desc(&tm_r32);
      tma_prefetch_desc(&tm_v8); tma_prefetch_desc(&tm_v1);
    }
    if (p.o_tma) tma_prefetch_desc(&tm_o);
    for (int u = 0; u < 2; ++u) {
      mbar_init(bar_s + 8 * u, 1); mbar_init(bar_p + 8 * u, 4); mbar_init(bar_o + 8 * u, 1); mbar_init(bar_free + 8 * u, 4);
      mbar_init(q_full + 8 * u, 1);
    }
    mbar_init(q_empty, 5);                 // S(0) retired + 4 helper warps (key-256 scores of rows 0..127)
    mbar_init(q_empty + 8, 6);             // S(1) retired + 4 helper warps + row-256 warp (query row 256)
    mbar_init(k_full, 1);
    mbar_init(k_empty, 6);                 // S(1) retired + 4 helper warps + row-256 warp (key row 256)
    mbar_init(v_full, 1);
    mbar_init(v_empty, 6);                 // P.V(1) retired + 4 helper warps + row-256 warp (their shares of row 256's P.V)
    mbar_init(cls_bar, 4);
    mbar_init(cls_p, 1);
    mbar_init(s256_full, 4);
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<1>(tmem_slot, VA_TMEM_COLS);
  // zero every operand buffer once: the padding (head_dim 88..95, rows/keys 257..271) is never written again --
  // except the ones-column of V: head dim 88 (first element of the padding chunk) of keys 0..256 is 1.0, so that column
  // 88 of O = P V is the row sum of the fp16-rounded probabilities, accumulated in fp32 by the tensor core
  for (uint32_t off = tid * 16; off < (uint32_t)V1_DATA_BYTES; off += V1_THREADS * 16) {
    uint32_t first = 0;
    if (off >= sV0 - base) {
      const uint32_t rel = off - (sV0 - base), within = rel % VA_G;
      if ((within >> 7) == 11 && (rel / VA_G) * 8 + ((within & 127) >> 4) < (uint32_t)VA_N) first = 0x3C00u;
    }
    *reinterpret_cast<uint4*>(gen + off) = make_uint4(first, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  constexpr uint32_t IDESC_S256 = make_idesc_f16(128, 256);
  constexpr uint32_t IDESC_O = make_idesc_f16(128, VA_DP) | (1u << 16);   // B (= V) is MN-major

  // O[256, 8 nt .. ] = P[256, :] V for `cnt` (<= 3) groups of 8 head dims on the legacy tensor pipe (mma.sync m16n8k16,
  // only row 0 of A is populated): per group nine ldmatrix.x4.trans of the no-swizzle V image (four 8-key x 8-dim core
  // matrices = two k-steps each) and 17 MMAs; the groups are independent accumulator chains.
  auto row256_pv = [&](int nt0, int cnt, const __half* ph, float inv256, __half* orow) {
    float c[3][4];
#pragma unroll
    for (int q = 0; q < 3; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.0f;
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(ph);
    const uint32_t va = sV0 + nt0 * 128 + (lane & 7) * 16;
    const int t4 = lane & 3;
    const bool row0 = lane < 4;
#pragma unroll
    for (int kp = 0; kp < 9; ++kp) {
      const int grp = kp < 8 ? 4 * kp + (lane >> 3) : 32 + ((lane >> 3) & 1);    // keys 256..271 are groups 32, 33
      uint32_t a0 = 0, a2 = 0, a4 = 0, a6 = 0;
      if (row0) {
        a0 = pw[16 * kp + t4]; a2 = pw[16 * kp + 4 + t4];
        if (kp < 8) { a4 = pw[16 * kp + 8 + t4]; a6 = pw[16 * kp + 12 + t4]; }
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (q < cnt) {
          uint32_t bm[4];
          va_ldsm_x4_t(bm, va + grp * VA_G + q * 128);
          va_mma16816(c[q], a0, 0u, a2, 0u, bm[0], bm[1]);
          if (kp < 8) va_mma16816(c[q], a4, 0u, a6, 0u, bm[2], bm[3]);
        }
      }
    }
    if (row0) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (q < cnt)
          *reinterpret_cast<__half2*>(orow + (nt0 + q) * 8 + 2 * t4) = __floats2half2_rn(c[q][0] * inv256, c[q][1] * inv256);
    }
  };

  // Q0 / K / Q1 / V for item number n (which = 0 / 1 / 2 / 3): wait until the buffer's previous contents are consumed,
  // then bulk-tensor copies from one thread (TMA) or 16-byte cp.async copies from the whole warp
  auto load_op = [&](int which, int item, uint32_t n) {
    const int b = item / p.heads, h = item - b * p.heads;
    const int row0 = which == 2 ? 128 : 0;
    const uint32_t dst = which == 0 ? sQ0 : (which == 1 ? sK0 : (which == 2 ? sQ1 : sV0));
    const uint32_t full = which == 0 ? q_full : (which == 1 ? k_full : (which == 2 ? q_full + 8 : v_full));
    const uint32_t empty = which == 0 ? q_empty : (which == 1 ? k_empty : (which == 2 ? q_empty + 8 : v_empty));
    VA_STAMP(10 + which, 0);
    mbar_wait_relaxed(empty, (n & 1) ^ 1);  // passes immediately the first time
    VA_STAMP(10 + which, 1);
    if constexpr (TMA) {
      if (lane == 0) {
        const int grow = b * VA_N + row0;
        if (which == 3) {
          // 32 groups of 8 keys (8 x 8 x 11 box -> 1408 bytes at the group's 1536-byte slot: the 12th chunk, which
          // holds the ones-column, is never touched) + key 256 through a 5-D view whose second dimension has extent 1
          // under a box of 8: rows 257..263 of the last group are out of bounds there and arrive as zeros, and the
          // chunk stride stays 128 bytes (a box of one row would pack its 11 chunks 16 bytes apart)
          mbar_arrive_expect_tx(full, (uint32_t)(33 * 8 * VA_D * 2));
          for (int g = 0; g < 32; ++g) tma_load_4d(dst + g * VA_G, &tm_v8, full, 0, grow + 8 * g, 0, 32 + h);
          tma_load_5d(dst + 32 * VA_G, &tm_v1, full, 0, 0, 0, grow + 256, 32 + h);
        } else {
          const int slot = (which == 1 ? 16 : 0) + h;          // q heads 0..15, k heads 16..31 (v: 32..47)
          const int rows8 = which == 0 ? 16 : (which == 1 ? 33 : 17);
          const uint32_t b0 = dst, b1 = dst + rows8 * 1024;
          const int big = which == 1 ? 2 : 1;                  // 128-row boxes
          const bool last_row = which != 0;                    // K and the second Q buffer also hold token 256
          mbar_arrive_expect_tx(full, (uint32_t)(big * 128 * 192 + (last_row ? 192 : 0)));
          for (int i = 0; i < big; ++i) {
            tma_load_3d(b0 + i * 128 * 128, &tm_a64, full, 0, slot, grow + i * 128);
            tma_load_3d(b1 + i * 128 * 64, &tm_a32, full, 64, slot, grow + i * 128);
          }
          if (last_row) {
            const int lr = big * 128;                          // local row of token 256 (start of its own 8-row group)
            tma_load_3d(b0 + lr * 128, &tm_r64, full, 0, slot, grow + lr);
            tma_load_3d(b1 + lr * 64, &tm_r32, full, 64, slot, grow + lr);
          }
        }
      }
      VA_STAMP(10 + which, 3);
    } else {
      // One instruction moves 8 rows x 4 chunks (512 B): the 8 rows fill one 128-byte core-matrix column each, so the
      // shared-memory side needs the minimum 4 wavefronts, and the addresses are pure adds (no divisions).
      const __half* src = which == 1 ? p.k + b * p.k_bs + h * p.k_hs
                                     : (which == 3 ? p.v + b * p.v_bs + h * p.v_hs : p.q + b * p.q_bs + h * p.q_hs);
      const long long ts = which == 1 ? p.k_ts : (which == 3 ? p.v_ts : p.q_ts);
      const int rows = which == 0 ? 128 : (which == 2 ? VA_N - 128 : VA_N);
      const int r8 = lane & 7, cq = lane >> 3;
      const int groups = (rows + 7) >> 3;
      const __half* rp = src + (long long)(row0 + r8) * ts + cq * 8;
      uint32_t dp = dst + cq * 128 + r8 * 16;
      for (int g = 0; g < groups; ++g) {
        if (g * 8 + r8 < rows) {
          cp_async16_tc(dp, rp);                                  // chunks 0..3
          cp_async16_tc(dp + 4 * 128, rp + 32);                   // chunks 4..7
          if (cq < 3) cp_async16_tc(dp + 8 * 128, rp + 64);       // chunks 8..10 (chunk 11: zero padding / ones-column)
        }
        rp += 8 * ts;
        dp += VA_G;
      }
      cp_async_wait_all_tc();
      fence_proxy_async_smem();             // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full);
      VA_STAMP(10 + which, 3);
    }
  };

  if (warp >= 8 && warp < 12) {
    // ======================= helper warps: the 257th token on the legacy tensor pipe =======================
    // 32 jobs per item, each a 16 x 96 by 96 x 1 product (six m16n8k16 with the vector in column 0 of B):
    //   jobs 0..15   s[r] = q_r . k_256 for query rows 16 j .. 16 j + 15   -> s_s256 (read by the softmax thread of row r)
    //   jobs 16..31  t[k] = q_256 . k_k for keys 16 (j - 16) ..            -> s_clsb (read by the row-256 warp)
    // dealt round-robin to the four warps, four jobs (= four independent accumulator chains) in flight.
    const int hw = warp - 8;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const uint32_t pq = n & 1;
      if constexpr (!TMA) load_op(hw, item, n);
      if (hw < 2) VA_STAMP(14 + hw, 0);
      mbar_wait_relaxed(q_full, pq);
      mbar_wait_relaxed(q_full + 8, pq);
      mbar_wait_relaxed(k_full, pq);
      if (hw < 2) VA_STAMP(14 + hw, 1);
      uint32_t bk[VA_DP / 16][2], bq[VA_DP / 16][2];
#pragma unroll
      for (int ks = 0; ks < VA_DP / 16; ++ks) {
        bk[ks][0] = bk[ks][1] = bq[ks][0] = bq[ks][1] = 0u;
        if (lane < 4) {
          bk[ks][0] = reinterpret_cast<const uint32_t*>(qk_chunk(gK, 33, 256, 2 * ks))[lane];
          bk[ks][1] = reinterpret_cast<const uint32_t*>(qk_chunk(gK, 33, 256, 2 * ks + 1))[lane];
          bq[ks][0] = reinterpret_cast<const uint32_t*>(qk_chunk(gQ1, 17, 128, 2 * ks))[lane];
          bq[ks][1] = reinterpret_cast<const uint32_t*>(qk_chunk(gQ1, 17, 128, 2 * ks + 1))[lane];
        }
      }
      float* sdst = s_s256 + pq * 256;
      float* tdst = s_clsb + pq * VA_CLS_LD;
      const int ar = lane & 15, ac = lane >> 4;   // ldmatrix: matrices 0/1 = rows 0..15 of chunk 2 ks, 2/3 = chunk 2 ks + 1
      // The legacy HMMA issues once per ~32 cycles per scheduler but takes ~110 cycles to return: four independent
      // jobs in flight, with the A fragments of the next k-step loaded (ldmatrix) before the current MMAs are issued.
#pragma unroll 1
      for (int j0 = hw; j0 < 32; j0 += 16) {
        float acc[4][4];
        int jj[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0f;
          jj[q] = j0 + 4 * q;
        }
        auto a_addr = [&](int q, int ks) -> uint32_t {
          const int j = jj[q], mt = j & 15;
          const int c = 2 * ks + ac;
          if (j >= 16) return smem_u32(qk_chunk(gK, 33, 16 * mt + ar, c));
          if (mt < 8) return smem_u32(qk_chunk(gQ0, 16, 16 * mt + ar, c));
          return smem_u32(qk_chunk(gQ1, 17, 16 * (mt - 8) + ar, c));
        };
        uint32_t af[2][4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) va_ldsm_x4(af[0][q], a_addr(q, 0));
#pragma unroll
        for (int ks = 0; ks < VA_DP / 16; ++ks) {
          if (ks + 1 < VA_DP / 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) va_ldsm_x4(af[(ks + 1) & 1][q], a_addr(q, ks + 1));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool tj = jj[q] >= 16;
            va_mma16816(acc[q], af[ks & 1][q][0], af[ks & 1][q][1], af[ks & 1][q][2], af[ks & 1][q][3],
                        tj ? bq[ks][0] : bk[ks][0], tj ? bq[ks][1] : bk[ks][1]);
          }
        }
        // column 0 of the accumulators: lanes 0, 4, 8, ... hold rows g and g + 8 of the job's 16
        if ((lane & 3) == 0) {
          const int g = lane >> 2;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float* dst = (jj[q] >= 16 ? tdst : sdst) + 16 * (jj[q] & 15);
            dst[g] = acc[q][0];
            dst[g + 8] = acc[q][2];
          }
        }
        if (j0 < 16) {
          // first round = this warp's four s-jobs: the softmax warps wait for these, the row-256 warp can wait longer
          __syncwarp();
          if (lane == 0) mbar_arrive(s256_full);
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(q_empty); mbar_arrive(q_empty + 8); mbar_arrive(k_empty);
        mbar_arrive(cls_bar);
      }
      if (hw < 2) VA_STAMP(14 + hw, 2);
      // this warp's share of row 256's P.V: 16 of the head dims 0..63
      mbar_wait_relaxed(cls_p, pq);
      mbar_wait_relaxed(v_full, pq);
      if (hw < 2) VA_STAMP(14 + hw, 3);
      {
        const int b = item / p.heads, h = item - b * p.heads;
        row256_pv(2 * hw, 2, s_clsh + pq * VA_CLS_LD, 1.0f / s_clsb[pq * VA_CLS_LD + VA_KP],
                  p.o + b * p.o_bs + h * p.o_hs + (long long)(VA_N - 1) * p.o_ts);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(v_empty);
      if (hw < 2) VA_STAMP(14 + hw, 4);
    }
  } else if (warp == 14) {
    // ======================= TMA producer for Q0 / K / Q1 =======================
    if constexpr (TMA) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        load_op(1, item, n);
        load_op(0, item, n);
        load_op(2, item, n);
        load_op(3, item, n);
      }
    }
  } else if (warp == 12) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        const uint32_t pn = n & 1;
        auto issue_s = [&](int u) {             // S_u = Q_u K^T (keys 0..255) into the tile's 256 TMEM columns
          const uint32_t qa = u == 0 ? sQ0 : sQ1;
          if constexpr (TMA) {
            const uint32_t qa1 = qa + (u == 0 ? 16 : 17) * 1024, ka1 = sK0 + 33 * 1024;    // the 64-byte-swizzled blocks
#pragma unroll
            for (int j = 0; j < 4; ++j)              // head dims 0..63: +32 bytes per 16-element k-step inside the atom
              umma_f16<1>(tmem + u * VA_TILE_COLS, make_smem_desc_sw128(qa) + 2 * j, make_smem_desc_sw128(sK0) + 2 * j,
                          IDESC_S256, j > 0);
#pragma unroll
            for (int j = 0; j < 2; ++j)              // head dims 64..95 (88..95 are the TMA's zero fill)
              umma_f16<1>(tmem + u * VA_TILE_COLS, make_smem_desc_sw64(qa1) + 2 * j, make_smem_desc_sw64(ka1) + 2 * j,
                          IDESC_S256, 1u);
          } else {
#pragma unroll
            for (int j = 0; j < VA_DP / 16; ++j)
              umma_f16<1>(tmem + u * VA_TILE_COLS, make_desc_nosw(qa + j * 256, 128, VA_G),
                          make_desc_nosw(sK0 + j * 256, 128, VA_G), IDESC_S256, j > 0);
          }
          umma_commit<1>(bar_s + 8 * u);
          umma_commit<1>(q_empty + 8 * u);      // the Q rows may be overwritten once S has retired (and the helper
        };                                      // warps have read them for the scores of key 256)
        auto issue_pv = [&](int u) {            // O_u = P_u V: A = P in TMEM, B = V (MN-major), 17 steps of 16 keys
          const uint32_t tb = tmem + u * VA_TILE_COLS;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            umma_f16_ts(tb + V1_O_COL, tb + j * 8, make_desc_nosw(sV0 + j * 2 * VA_G, VA_G, 128), IDESC_O, j > 0);
          // keys 256..271: P of key 256 sits in its own 8 columns, V rows 257.. are zero
          umma_f16_ts(tb + V1_O_COL, tb + V1_P256_COL, make_desc_nosw(sV0 + 32 * VA_G, VA_G, 128), IDESC_O, 1u);
          umma_commit<1>(bar_o + 8 * u);
        };
        VA_STAMP(8, 0);
        mbar_wait(k_full, pn);
        for (int u = 0; u < 2; ++u) {
          mbar_wait(q_full + 8 * u, pn);
          if (n > 0) mbar_wait(bar_free + 8 * u, pn ^ 1);    // O of this pipeline's previous tile has been read out
          tc_fence_after();
          VA_STAMP(8, 1 + u);
          issue_s(u);
        }
        umma_commit<1>(k_empty);
        VA_STAMP(8, 3);
        mbar_wait(v_full, pn);
        VA_STAMP(8, 4);
        for (int u = 0; u < 2; ++u) {
          mbar_wait(bar_p + 8 * u, pn);          // P_u in TMEM, S_u fully read
          tc_fence_after();
          VA_STAMP(8, 5 + u);
          issue_pv(u);
        }
        umma_commit<1>(v_empty);
        VA_STAMP(8, 7);
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ======================= query row 256 (the 257th token): its softmax + its share of P.V =======================
    // 1 row x 257 keys x 88 dims: a third 128-row MMA tile would be 99% padding.  The helper warps leave the 256
    // scores q_256 . k_key in s_clsb, this warp adds key 256 and runs the softmax over the 257 scores; the P.V
    // product of the row is split by head dims over the helper warps and this one (dims 64..87).
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const uint32_t pq = n & 1;
      float* cls = s_clsb + pq * VA_CLS_LD;
      __half* clsh = s_clsh + pq * VA_CLS_LD;
      const int b = item / p.heads, h = item - b * p.heads;
      VA_STAMP(9, 0);
      mbar_wait_relaxed(q_full + 8, pq);
      mbar_wait_relaxed(k_full, pq);
      VA_STAMP(9, 1);
      // score of key 256: lanes 0..10 take one 8-dim chunk each
      float part = 0.0f;
      if (lane < VA_D / 8) {
        const uint4 qa = *qk_chunk(gQ1, 17, 128, lane);   // query row 256 = local row 128 of the second Q buffer
        const uint4 ka = *qk_chunk(gK, 33, 256, lane);    // key row 256
        const __half2* q2 = reinterpret_cast<const __half2*>(&qa);
        const __half2* k2 = reinterpret_cast<const __half2*>(&ka);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 qf = __half22float2(q2[j]), kf = __half22float2(k2[j]);
          part = fmaf(qf.x, kf.x, part);
          part = fmaf(qf.y, kf.y, part);
        }
      }
      part = warp_sum(part);
      __syncwarp();
      if (lane == 0) { mbar_arrive(k_empty); mbar_arrive(q_empty + 8); }
      VA_STAMP(9, 2);
      mbar_wait_relaxed(cls_bar, pq);                     // the 256 distributed scores are in cls[0..255]
      VA_STAMP(9, 3);
      float sc[9];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int key = lane + 32 * i;
        float v = -INFINITY;
        if (key < VA_N - 1) v = cls[key] * p.scale_log2;
        else if (key == VA_N - 1) v = part * p.scale_log2;
        sc[i] = v;
        mx = fmaxf(mx, v);
      }
      mx = warp_max(mx);
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int key = lane + 32 * i;
        const __half ph = __float2half_rn(ex2f(sc[i] - mx));               // P rounded to fp16 like the tile path
        sum += __half2float(ph);
        if (key < VA_KP) clsh[key] = ph;                   // 0 for keys 257..271
      }
      sum = warp_sum(sum);
      if (lane == 0) cls[VA_KP] = sum;
      __syncwarp();
      if (lane == 0) mbar_arrive(cls_p);                  // release: the probabilities are visible to the waiters
      VA_STAMP(9, 4);
      mbar_wait_relaxed(v_full, pq);
      row256_pv(8, 3, clsh, 1.0f / sum, p.o + b * p.o_bs + h * p.o_hs + (long long)(VA_N - 1) * p.o_ts);
      __syncwarp();
      if (lane == 0) mbar_arrive(v_empty);
      VA_STAMP(9, 5);
    }
  } else if (warp < 8) {
    // ======================= softmax + read-out: warps 0-3 own tile 0 (rows 0..127), warps 4-7 tile 1 =======================
    const int quarter = warp & 3, u = warp >> 2;
    const int rl = quarter * 32 + lane;               // row inside the tile: one full row per thread
    const int row = u * 128 + rl;
    const uint32_t trow = tmem + u * VA_TILE_COLS + ((uint32_t)(quarter * 32) << 16);
    const uint32_t stage = sStage + warp * V1_STAGE_WARP;
    uint8_t* gstage = gen + (stage - base) + lane * V1_STAGE_ROW;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.heads, h = item - b * p.heads;
      const uint32_t pn = n & 1;
      VA_STAMP(warp, 0);
      mbar_wait_relaxed(bar_s + 8 * u, pn);
      tc_fence_after();
      VA_STAMP(warp, 3);
      // Both passes walk the row in 32-column chunks with the NEXT chunk's tcgen05.ld already in flight while the
      // current one is consumed (two register buffers): a thread only ever has one round trip to tensor memory
      // exposed, not one per chunk.
      uint32_t r0[32], r1[32];
      // pass 1: row maximum
      float mx = -INFINITY;
      auto chunk_max = [&](const uint32_t(&cur)[32]) {
        float m0 = __uint_as_float(cur[0]), m1 = __uint_as_float(cur[1]);
#pragma unroll
        for (int j = 2; j < 32; j += 2) {
          m0 = fmaxf(m0, __uint_as_float(cur[j]));
          m1 = fmaxf(m1, __uint_as_float(cur[j + 1]));
        }
        mx = fmaxf(mx, fmaxf(m0, m1));
      };
      tmem_ld32(trow, r0);
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {                     // chunk pairs: buffer roles are compile-time inside the body
        tmem_ld_wait32(r0);
        tmem_ld32(trow + (c + 1) * 32, r1);
        chunk_max(r0);
        tmem_ld_wait32(r1);
        tmem_ld32(trow + ((c + 2) & 7) * 32, r0);          // after chunk 7 this is chunk 0 again: pass 2's first load
        chunk_max(r1);
      }
      // the score of key 256 for this row comes from the helper warps (computed while the previous item was in flight)
      mbar_wait_relaxed(s256_full, pn);
      const float s256 = s_s256[pn * 256 + row];
      mx = fmaxf(mx, s256);
      VA_STAMP(warp, 4);
      const float m = mx * p.scale_log2;                 // scale > 0
      // pass 2: P = exp2(s*scale*log2e - m), rounded to fp16 (as the reference does under autocast), written in
      // place: chunk c of S (32 columns) becomes the 16 packed columns [16 c, 16 c + 16), inside chunks already loaded
      auto chunk_exp = [&](const uint32_t(&cur)[32], uint32_t dst) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                   // two halves of 16 columns: bounds the live temporaries
          uint32_t pk[8];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            pk[g] = pack2n(ex2f(fmaf(__uint_as_float(cur[hf * 16 + 2 * g]), p.scale_log2, -m)),
                           ex2f(fmaf(__uint_as_float(cur[hf * 16 + 2 * g + 1]), p.scale_log2, -m)));
          tmem_st8(dst + hf * 8, pk);
        }
      };
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait32(r0);
        tmem_ld32(trow + (c + 1) * 32, r1);
        chunk_exp(r0, trow + c * 16);
        tmem_ld_wait32(r1);
        if (c + 2 < 8) tmem_ld32(trow + (c + 2) * 32, r0);
        chunk_exp(r1, trow + (c + 1) * 16);
      }
      {
        // key 256: its probability as the A operand of a 17th P.V step (the 15 keys after it are zero rows of V)
        uint32_t pk[8] = {pack2n(ex2f(fmaf(s256, p.scale_log2, -m)), 0.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(trow + V1_P256_COL, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * u);
      VA_STAMP(warp, 5);

      mbar_wait_relaxed(bar_o + 8 * u, pn);
      tc_fence_after();
      VA_STAMP(warp, 6);
      {
        // O read-out.  dims 64..95 first: dim 88 holds the row sum (ones-column of V), all 257 keys included.  With a
        // TMA store each thread parks its row (176 bytes) in the warp's staging area -- at a 176-byte pitch the
        // 16-byte stores of a quarter warp hit 8 distinct bank groups -- and one thread stores the 32 rows; the direct
        // path (any other output layout) costs 32 partial sectors per store instruction.
        const bool otma = p.o_tma != 0;
        __half* og = p.o + b * p.o_bs + h * p.o_hs + (long long)row * p.o_ts;
        auto store8 = [&](const uint32_t* src, int chunk, float inv) {
          uint4 o;
          o.x = pack2n(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          o.y = pack2n(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          o.z = pack2n(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          o.w = pack2n(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          if (otma) *reinterpret_cast<uint4*>(gstage + chunk * 16) = o;
          else *reinterpret_cast<uint4*>(og + chunk * 8) = o;
        };
        uint32_t c2[32], c0[32];
        tmem_ld32(trow + V1_O_COL + 64, c2);               // dims 64..95
        tmem_ld32(trow + V1_O_COL, c0);                    // dims 0..31
        if (otma && n > 0) {                               // the previous item's TMA store has read the staging area
          if (lane == 0) bulk_wait_group_read0();
          __syncwarp();
        }
        tmem_ld_wait32(c2);
        tmem_ld_wait32(c0);
        const float inv = 1.0f / __uint_as_float(c2[24]);
#pragma unroll
        for (int g = 0; g < 4; ++g) store8(&c0[g * 8], g, inv);
        uint32_t c1[32];
        tmem_ld32(trow + V1_O_COL + 32, c1);               // dims 32..63
#pragma unroll
        for (int g = 0; g < 3; ++g) store8(&c2[g * 8], 8 + g, inv);    // dims 64..87 (chunk 11 is padding)
        tmem_ld_wait32(c1);
        // every column of O is in registers: hand the tile's TMEM columns back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 8 * u);
#pragma unroll
        for (int g = 0; g < 4; ++g) store8(&c1[g * 8], 4 + g, inv);
        if (otma) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tm_o, stage, 0, h, b * VA_N + u * 128 + quarter * 32);
            bulk_commit_group();
          }
        }
      }
      VA_STAMP(warp, 7);
    }
    if (p.o_tma != 0 && lane == 0) bulk_wait_group0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<1>(tmem, VA_TMEM_COLS);
}

long long get_option64(const char* key);

bool vit_attention_tc_applicable(const seedb200_attn_desc& d) {
  return d.head_dim == VA_D && d.nq == VA_N && d.nk == VA_N && d.causal == 0 && d.o_hs % 8 == 0 && d.o_ts % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(d.o) & 15) == 0;
}

typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D map over the packed projection buffer: (88 elements of a head | 48 head slots = q, k, v x 16 | token rows);
// the same shape with 16 slots describes the [B*257, 16, 88] attention output for the TMA store
int make_qkv_tmap(CUtensorMap* tm, const void* base, long long rows, long long pitch_elems, int box_elems,
                  int box_rows, CUtensorMapSwizzle swz, int slots) {
  static EncodeTiledFnA fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnA>(ptr);
  }
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SEEDB200_ERR_CUDA;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)VA_D, (cuuint64_t)slots, (cuuint64_t)rows};
  cuuint64_t gstr[2] = {(cuuint64_t)VA_D * 2, (cuuint64_t)pitch_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_elems, 1, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("vit_attention: cuTensorMapEncodeTiled failed with CUresult %d (box %d x %d)", (int)r, box_elems, box_rows);
    return SEEDB200_ERR_CUDA;
  }
  return 0;
}

// Maps for V in the no-swizzle core-matrix image.  4-D: (8 elements = one 16-byte chunk | token rows | 11 chunks of a
// head | 48 head slots); a box of 8 x 8 x 11 x 1 lands as 11 core-matrix columns of 8 rows x 16 bytes.  single_row: the
// 5-D view (8 | 1 | 11 | rows | 48) with the same 8 x 8 x 11 box -- indices 1..7 of the extent-1 dimension are out of
// bounds and zero-filled, which loads ONE token row into row 0 of a group and clears the other seven.
int make_v_tmap(CUtensorMap* tm, const void* base, long long rows, long long pitch_elems, bool single_row) {
  static EncodeTiledFnA fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnA>(ptr);
  }
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SEEDB200_ERR_CUDA;
  }
  const cuuint64_t pitch_b = (cuuint64_t)pitch_elems * 2;
  cuuint64_t gdim4[4] = {8, (cuuint64_t)rows, (cuuint64_t)(VA_D / 8), 48};
  cuuint64_t gstr4[3] = {pitch_b, 16, (cuuint64_t)VA_D * 2};
  cuuint32_t box4[4] = {8, 8, (cuuint32_t)(VA_D / 8), 1};
  cuuint64_t gdim5[5] = {8, 1, (cuuint64_t)(VA_D / 8), (cuuint64_t)rows, 48};
  cuuint64_t gstr5[4] = {pitch_b, 16, pitch_b, (cuuint64_t)VA_D * 2};
  cuuint32_t box5[5] = {8, 8, (cuuint32_t)(VA_D / 8), 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = single_row
                         ? fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), gdim5, gstr5, box5, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                         : fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim4, gstr4, box4, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("vit_attention: cuTensorMapEncodeTiled (V%s) failed with CUresult %d", single_row ? ", last row" : "", (int)r);
    return SEEDB200_ERR_CUDA;
  }
  return 0;
}

int get_option(const char* key);

// TMA needs q, k, v to be the [B*257, 3, 16, 88] views of ONE projection buffer (what the fused qkv GEMM writes,
// eva_vit.py:133-138); any other strided layout takes the cp.async loaders
bool vit_attention_packed_qkv(const seedb200_attn_desc& d) {
  const __half* qp = static_cast<const __half*>(d.q);
  return d.heads == 16 && d.q_hs == VA_D && d.k_hs == VA_D && d.q_ts == d.k_ts && d.q_ts % 8 == 0 &&
         d.q_ts >= 48 * VA_D && d.q_bs == (int64_t)VA_N * d.q_ts && d.k_bs == d.q_bs &&
         static_cast<const __half*>(d.k) == qp + 16 * VA_D && (reinterpret_cast<uintptr_t>(qp) & 15) == 0 &&
         d.v_hs == VA_D && d.v_ts == d.q_ts && d.v_bs == d.q_bs && static_cast<const __half*>(d.v) == qp + 32 * VA_D;
}

int vit_attention_tc(const seedb200_attn_desc& d, cudaStream_t stream) {
  const __half* qp = static_cast<const __half*>(d.q);
  const bool tma = vit_attention_packed_qkv(d) && get_option("vit_attention_tma") != 0;
  auto kern = tma ? vit_attention_tc_kernel<true> : vit_attention_tc_kernel<false>;
  static bool attr_set_dev[SB_MAX_DEVICES][2] = {};   // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_dev[cur_device()][tma ? 1 : 0];
  if (!attr_set) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, V1_SMEM));
    attr_set = true;
  }
  VitAttnParams p;
  p.q = static_cast<const __half*>(d.q); p.k = static_cast<const __half*>(d.k);
  p.v = static_cast<const __half*>(d.v); p.o = static_cast<__half*>(d.o);
  p.q_bs = d.q_bs; p.q_hs = d.q_hs; p.q_ts = d.q_ts;
  p.k_bs = d.k_bs; p.k_hs = d.k_hs; p.k_ts = d.k_ts;
  p.v_bs = d.v_bs; p.v_hs = d.v_hs; p.v_ts = d.v_ts;
  p.o_bs = d.o_bs; p.o_hs = d.o_hs; p.o_ts = d.o_ts;
  p.items = d.batch * d.heads; p.heads = d.heads;
  p.scale_log2 = d.scale * 1.4426950408889634f;
  p.dbg = reinterpret_cast<long long*>(static_cast<uintptr_t>(get_option64("vit_attention_dbg_ptr")));
  CUtensorMap ta64, ta32, tr64, tr32, to, tv8, tv1;
  memset(&ta64, 0, sizeof(ta64)); ta32 = ta64; tr64 = ta64; tr32 = ta64; to = ta64; tv8 = ta64; tv1 = ta64;
  const long long rows = (long long)d.batch * VA_N;
  if (tma) {
    SB_PROPAGATE(make_qkv_tmap(&ta64, qp, rows, d.q_ts, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 48));
    SB_PROPAGATE(make_qkv_tmap(&ta32, qp, rows, d.q_ts, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 48));
    SB_PROPAGATE(make_qkv_tmap(&tr64, qp, rows, d.q_ts, 64, 1, CU_TENSOR_MAP_SWIZZLE_128B, 48));
    SB_PROPAGATE(make_qkv_tmap(&tr32, qp, rows, d.q_ts, 32, 1, CU_TENSOR_MAP_SWIZZLE_64B, 48));
    SB_PROPAGATE(make_v_tmap(&tv8, qp, rows, d.q_ts, false));
    SB_PROPAGATE(make_v_tmap(&tv1, qp, rows, d.q_ts, true));
  }
  // the output goes out by TMA when it is the [B*257, heads, 88] view of a row-major buffer (what the encoder passes)
  p.o_tma = (d.o_hs == VA_D && d.o_bs == (int64_t)VA_N * d.o_ts && d.o_ts % 8 == 0 && d.o_ts >= (int64_t)d.heads * VA_D &&
             get_option("vit_attention_tma") != 0) ? 1 : 0;
  if (p.o_tma) SB_PROPAGATE(make_qkv_tmap(&to, d.o, rows, d.o_ts, VA_D, 32, CU_TENSOR_MAP_SWIZZLE_NONE, d.heads));
  int grid = num_sms();
  if (grid > p.items) grid = p.items;
  profile_mark_begin(1, stream);
  kern<<<grid, V1_THREADS, V1_SMEM, stream>>>(p, ta64, ta32, tr64, tr32, to, tv8, tv1);
  profile_mark_end(1, stream, 4.0 * (double)d.batch * d.heads * (double)d.nq * d.nk * d.head_dim);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sb
