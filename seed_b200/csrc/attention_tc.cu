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

// TMA = true: Q and K arrive by TMA (3-D tensor maps over the packed qkv buffer: 88 elements per head slot with OOB zero
// fill up to 96, 48 head slots, rows) in the K-major swizzled layouts the GEMM uses -- a 128-byte-swizzled block of head
// dims 0..63 and a 64-byte-swizzled block of dims 64..95 -- instead of 16-byte cp.async copies into the no-swizzle
// core-matrix layout: the cp.async path cost ~20 LSU cycles per 512 bytes (tools/attn_timeline.py: 5-8 k cycles per
// operand), and the SM's one load/store unit was what bounded the kernel's period (profiles/r02_attention.md).
// V (an MN-major operand) stays on the cp.async / no-swizzle path.
template <bool TMA>
__global__ void __launch_bounds__(VA_THREADS, 1)
vit_attention_tc_kernel(const VitAttnParams p, const __grid_constant__ CUtensorMap tm_a64,
                        const __grid_constant__ CUtensorMap tm_a32, const __grid_constant__ CUtensorMap tm_r64,
                        const __grid_constant__ CUtensorMap tm_r32) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = TMA ? ((smem_u32(smem_raw) + 1023u) & ~1023u) : ((smem_u32(smem_raw) + 127u) & ~127u);
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK0 = base, sQ0 = sK0 + VA_K_BYTES, sQ1 = sQ0 + VA_Q0_BYTES, sV0 = sQ1 + VA_Q1_BYTES;
  const uint32_t misc = sV0 + 2 * VA_V_BYTES;
  float* s_clsb = reinterpret_cast<float*>(gen + (misc - base));    // [2][VA_CLS_LD]: scores, then probabilities of query row 256
  __half* s_clsh = reinterpret_cast<__half*>(s_clsb + 2 * VA_CLS_LD);   // [2][VA_CLS_LD]: probabilities of row 256 as the fp16 A operand
  float* s_sx = s_clsb + 2 * VA_CLS_LD + VA_CLS_LD;                 // [8][32]: per-warp hand-over of the key-256 scores
  const uint32_t bars = misc + 2 * VA_CLS_LD * 4 + 8 * VA_PART_LD * 4;
  const uint32_t bar_s = bars, bar_p = bars + 16, bar_o = bars + 32, bar_free = bars + 48;      // [2] each: per tile pipeline
  const uint32_t q_full = bars + 64 /*[2]*/, q_empty = bars + 80 /*[2]*/;
  const uint32_t k_full = bars + 96, k_empty = bars + 104, v_full = bars + 112 /*[2]*/, v_empty = bars + 128 /*[2]*/;
  const uint32_t cls_bar = bars + 144;           // 8 softmax warps -> row-256 warp: scores of query 256 are in s_cls
  const uint32_t cls_p = bars + 152;             // row-256 warp -> softmax warps: probabilities of row 256 are in s_cls
  const uint32_t tmem_slot = bars + 160;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  uint8_t* gQ0 = gen + (sQ0 - base);
  uint8_t* gQ1 = gen + (sQ1 - base);
  uint8_t* gV0 = gen + (sV0 - base);
  const uint8_t* gK = gen + (sK0 - base);
  // 16-byte chunk c (8 head dims) of row r of a Q / K buffer with `rows8` 8-row groups: the CUDA-core readers of the
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
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(bar_s + 8 * u, 1); mbar_init(bar_p + 8 * u, 4); mbar_init(bar_o + 8 * u, 1); mbar_init(bar_free + 8 * u, 4);
      mbar_init(q_full + 8 * u, 1);
      mbar_init(v_full + 8 * u, 1);
      mbar_init(v_empty + 8 * u, 10);      // P.V(1) retired + 8 softmax warps + row-256 warp (their shares of row 256's P.V)
    }
    mbar_init(q_empty, 5);                 // S(0) retired + 4 softmax warps of tile 0 (their q rows, for key 256)
    mbar_init(q_empty + 8, 10);            // S(1) retired + all 8 softmax warps (tile-1 rows, query row 256) + row-256 warp
    mbar_init(cls_bar, 8);
    mbar_init(cls_p, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 10);                // S(1) retired + 8 softmax warps + row-256 warp (key row 256)
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<1>(tmem_slot, VA_TMEM_COLS);
  // zero every operand buffer once: the padding (head_dim 88..95, rows/keys 257..271) is never written again
  // except the ones-column of V: head dim 88 (first element of the padding chunk) of keys 0..256 is 1.0, so that column
  // 88 of O = P V is the row sum of the fp16-rounded probabilities, accumulated in fp32 by the tensor core
  for (uint32_t off = tid * 16; off < (uint32_t)VA_DATA_BYTES; off += VA_THREADS * 16) {
    uint32_t first = 0;
    if (off >= sV0 - base) {
      const uint32_t rel = (off - (sV0 - base)) % VA_V_BYTES, within = rel % VA_G;
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
  constexpr uint32_t IDESC_O = make_idesc_f16(128, 48) | (1u << 16);      // half of the head dim; B (= V) is MN-major
  constexpr int CH = VA_D / 8;            // 11 16-byte chunks per row

  // O[256, 8 nt .. 8 nt + 7] = P[256, :] V on the legacy tensor pipe (mma.sync m16n8k16, only row 0 of A is populated):
  // nine ldmatrix.x4.trans of the no-swizzle V image (four 8-key x 8-dim core matrices = two k-steps each) and 17 MMAs
  // per 8 head dims.  The eight softmax warps take dims 0..63, the row-256 warp dims 64..87.
  auto row256_pv = [&](int nt, const __half* ph, const uint8_t* gV, float inv256, __half* orow) {
    float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(ph);
    const uint32_t va = smem_u32(gV) + nt * 128 + (lane & 7) * 16;
    const int t4 = lane & 3;
    const bool row0 = lane < 4;
#pragma unroll
    for (int kp = 0; kp < 9; ++kp) {
      const int grp = kp < 8 ? 4 * kp + (lane >> 3) : 32 + ((lane >> 3) & 1);    // keys 256..271 are groups 32, 33
      uint32_t bm[4];
      va_ldsm_x4_t(bm, va + grp * VA_G);
      uint32_t a0 = 0, a2 = 0;
      if (row0) { a0 = pw[16 * kp + t4]; a2 = pw[16 * kp + 4 + t4]; }
      va_mma16816(c, a0, 0u, a2, 0u, bm[0], bm[1]);
      if (kp < 8) {
        uint32_t a4 = 0, a6 = 0;
        if (row0) { a4 = pw[16 * kp + 8 + t4]; a6 = pw[16 * kp + 12 + t4]; }
        va_mma16816(c, a4, 0u, a6, 0u, bm[2], bm[3]);
      }
    }
    if (row0) *reinterpret_cast<__half2*>(orow + nt * 8 + 2 * t4) = __floats2half2_rn(c[0] * inv256, c[1] * inv256);
  };

  if (warp >= 8 && warp < 12) {
    // ======================= loaders: one warp per operand buffer, cp.async 16-byte copies =======================
    // One instruction moves 8 rows x 4 chunks (512 B): the 8 rows fill one 128-byte core-matrix column each, so the
    // shared-memory side needs the minimum 4 wavefronts, and the addresses are pure adds (no divisions).
    const int which = warp - 8;             // 0: Q rows 0..127, 1: K, 2: Q rows 128..256, 3: V
    const int r8 = lane & 7, cq = lane >> 3;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.heads, h = item - b * p.heads;
      const __half* src; long long ts; int row0, rows; uint32_t dst, full, empty, par;
      if (which == 0)      { src = p.q + b * p.q_bs + h * p.q_hs; ts = p.q_ts; row0 = 0;   rows = 128; dst = sQ0; full = q_full; empty = q_empty; par = n & 1; }
      else if (which == 1) { src = p.k + b * p.k_bs + h * p.k_hs; ts = p.k_ts; row0 = 0;   rows = VA_N; dst = sK0; full = k_full; empty = k_empty; par = n & 1; }
      else if (which == 2) { src = p.q + b * p.q_bs + h * p.q_hs; ts = p.q_ts; row0 = 128; rows = VA_N - 128; dst = sQ1; full = q_full + 8; empty = q_empty + 8; par = n & 1; }
      else                 { src = p.v + b * p.v_bs + h * p.v_hs; ts = p.v_ts; row0 = 0;   rows = VA_N; dst = sV0 + (n & 1) * VA_V_BYTES; full = v_full + 8 * (n & 1); empty = v_empty + 8 * (n & 1); par = (n >> 1) & 1; }
      VA_STAMP(10 + which, 0);
      mbar_wait_relaxed(empty, par ^ 1);    // previous contents consumed (passes immediately the first time)
      VA_STAMP(10 + which, 1);
      if constexpr (TMA) {
        if (which < 3) {
          // one thread, a handful of bulk-tensor copies: 128-row boxes of the two swizzled blocks (+ row 256 alone)
          if (lane == 0) {
            const int slot = (which == 1 ? 16 : 0) + h;          // q heads 0..15, k heads 16..31 (v: 32..47)
            const int grow = b * VA_N + row0;
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
          VA_STAMP(10 + which, 3);
          continue;
        }
      }
      const int groups = (rows + 7) >> 3;
      const __half* rp = src + (long long)(row0 + r8) * ts + cq * 8;
      uint32_t dp = dst + cq * 128 + r8 * 16;
      for (int g = 0; g < groups; ++g) {
        if (g * 8 + r8 < rows) {
          cp_async16_tc(dp, rp);                                  // chunks 0..3
          cp_async16_tc(dp + 4 * 128, rp + 32);                   // chunks 4..7
          if (cq < 3) cp_async16_tc(dp + 8 * 128, rp + 64);       // chunks 8..10 (chunk 11 is the zero padding)
        }
        rp += 8 * ts;
        dp += VA_G;
      }
      VA_STAMP(10 + which, 2);
      cp_async_wait_all_tc();
      fence_proxy_async_smem();             // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full);
      VA_STAMP(10 + which, 3);
    }
  } else if (warp == 12) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        const uint32_t vb = n & 1, pn = n & 1;
        const uint32_t sV = sV0 + vb * VA_V_BYTES;
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
          umma_commit<1>(q_empty + 8 * u);      // the Q rows may be overwritten once S has retired (and the softmax
        };                                      // warps have read their rows for key 256)
        auto issue_pv = [&](int u) {            // O_u = P_u V: A = P in TMEM, B = V (MN-major), two 48-wide halves of d
          const uint32_t tb = tmem + u * VA_TILE_COLS;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t pa = tb + (j < 8 ? j * 8 : 128 + (j - 8) * 8);
            umma_f16_ts(tb + VA_OLO_COL, pa, make_desc_nosw(sV + j * 2 * VA_G, VA_G, 128), IDESC_O, j > 0);
            umma_f16_ts(tb + VA_OHI_COL, pa, make_desc_nosw(sV + j * 2 * VA_G + 6 * 128, VA_G, 128), IDESC_O, j > 0);
          }
          // keys 256..271: P of key 256 sits in its own 8 columns, V rows 257.. are zero
          umma_f16_ts(tb + VA_OLO_COL, tb + VA_P256_COL, make_desc_nosw(sV + 32 * VA_G, VA_G, 128), IDESC_O, 1u);
          umma_f16_ts(tb + VA_OHI_COL, tb + VA_P256_COL, make_desc_nosw(sV + 32 * VA_G + 6 * 128, VA_G, 128), IDESC_O, 1u);
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
        mbar_wait(v_full + 8 * vb, (n >> 1) & 1);
        VA_STAMP(8, 4);
        for (int u = 0; u < 2; ++u) {
          mbar_wait(bar_p + 8 * u, pn);          // P_u in TMEM, S_u fully read
          tc_fence_after();
          VA_STAMP(8, 5 + u);
          issue_pv(u);
        }
        umma_commit<1>(v_empty + 8 * vb);
        VA_STAMP(8, 7);
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ======================= query row 256 (the 257th token): its softmax =======================
    // 1 row x 257 keys x 88 dims: a third 128-row MMA tile would be 99% padding.  The 256 softmax threads each
    // contribute the score of "their" key (thread <-> key), this warp adds key 256 and runs the softmax over the 257
    // scores; the P.V product of the row is spread over the 256 softmax threads again (it was 12 k cycles per item on
    // this one warp -- the whole kernel's period).  s_cls is double buffered by item parity.
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      float* cls = s_clsb + (n & 1) * VA_CLS_LD;
      __half* clsh = s_clsh + (n & 1) * VA_CLS_LD;
      const int b = item / p.heads, h = item - b * p.heads;
      VA_STAMP(9, 0);
      mbar_wait_relaxed(q_full + 8, n & 1);
      mbar_wait_relaxed(k_full, n & 1);
      VA_STAMP(9, 1);
      // score of key 256: lanes 0..10 take one 8-dim chunk each
      float part = 0.0f;
      if (lane < CH) {
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
      mbar_wait_relaxed(cls_bar, n & 1);                  // the 256 distributed scores are in cls[0..255]
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
      __syncwarp();
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
      // this warp's share of the row's P.V: head dims 64..87
      {
        const uint32_t vb = n & 1;
        mbar_wait_relaxed(v_full + 8 * vb, (n >> 1) & 1);
        __half* orow = p.o + b * p.o_bs + h * p.o_hs + (long long)(VA_N - 1) * p.o_ts;
        const float inv256 = 1.0f / sum;
#pragma unroll 1
        for (int nt = 8; nt < 11; ++nt) row256_pv(nt, clsh, gV0 + vb * VA_V_BYTES, inv256, orow);
        __syncwarp();
        if (lane == 0) mbar_arrive(v_empty + 8 * vb);
      }
      VA_STAMP(9, 5);
    }
  } else {
    // ======================= softmax + epilogue: warps 0-3 own tile 0 (rows 0..127), warps 4-7 tile 1 =======================
    const int quarter = warp & 3, u = warp >> 2;
    const int rl = quarter * 32 + lane;               // row inside the tile: one full row (256 + 1 keys) per thread
    const int row = u * 128 + rl;
    const uint32_t trow = tmem + u * VA_TILE_COLS + ((uint32_t)(quarter * 32) << 16);
    // ---- the 257th token on the CUDA cores: s256 = q_row . k_256 (key 256 for this thread's row) and
    //      t = q_256 . k_key (this thread's key for query row 256, left in s_cls for the row-256 warp).  The dot products
    //      of item n+1 are computed while the P.V MMA of item n runs (its Q and K are already in shared memory: both S
    //      MMAs of item n retired long ago and the TMA refills the buffers in ~1 k cycles), so they are off the
    //      S -> softmax -> P.V -> read-out chain; only item 0's are exposed.
    auto dots = [&](uint32_t nn) -> float {
      const uint32_t pq = nn & 1;
      mbar_wait_relaxed(q_full + 8 * u, pq);
      if (u == 0) mbar_wait_relaxed(q_full + 8, pq);                // query row 256 lives in the second Q buffer
      mbar_wait_relaxed(k_full, pq);
      // both as 32 x 96 by 96 x 1 products on the legacy tensor pipe: A = this warp's 32 query rows (resp. key rows)
      // through ldmatrix from the swizzled / core-matrix image, B = key 256 (resp. query 256) in column 0 of the n = 8
      float sc[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
      float tc[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
      const uint8_t* qbuf = u == 0 ? gQ0 : gQ1;
      const int qg = u == 0 ? 16 : 17;
      const int ar = quarter * 32 + (lane & 15);          // ldmatrix row of this lane inside m-tile 0 (matrices 0/1: rows 0..15)
      const int ac = lane >> 4;                           // matrices 2/3: the k-step's second 8-dim chunk
#pragma unroll
      for (int ks = 0; ks < VA_DP / 16; ++ks) {
        uint32_t bk0 = 0, bk1 = 0, bq0 = 0, bq1 = 0;
        if (lane < 4) {
          bk0 = reinterpret_cast<const uint32_t*>(qk_chunk(gK, 33, 256, 2 * ks))[lane];
          bk1 = reinterpret_cast<const uint32_t*>(qk_chunk(gK, 33, 256, 2 * ks + 1))[lane];
          bq0 = reinterpret_cast<const uint32_t*>(qk_chunk(gQ1, 17, 128, 2 * ks))[lane];
          bq1 = reinterpret_cast<const uint32_t*>(qk_chunk(gQ1, 17, 128, 2 * ks + 1))[lane];
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t a[4];
          va_ldsm_x4(a, smem_u32(qk_chunk(qbuf, qg, ar + 16 * mt, 2 * ks + ac)));
          va_mma16816(sc[mt], a[0], a[1], a[2], a[3], bk0, bk1);
          va_ldsm_x4(a, smem_u32(qk_chunk(gK, 33, u * 128 + ar + 16 * mt, 2 * ks + ac)));
          va_mma16816(tc[mt], a[0], a[1], a[2], a[3], bq0, bq1);
        }
      }
      // column 0 of the accumulators: lanes 0, 4, 8, ... hold rows g and g + 8 of each m-tile
      if ((lane & 3) == 0) {
        const int g = lane >> 2;
        float* tdst = s_clsb + (nn & 1) * VA_CLS_LD + u * 128 + quarter * 32;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          s_sx[warp * 32 + mt * 16 + g] = sc[mt][0];
          s_sx[warp * 32 + mt * 16 + g + 8] = sc[mt][2];
          tdst[mt * 16 + g] = tc[mt][0];
          tdst[mt * 16 + g + 8] = tc[mt][2];
        }
      }
      __syncwarp();
      const float s256 = s_sx[warp * 32 + lane];
      __syncwarp();
      if (lane == 0) { mbar_arrive(q_empty + 8 * u); if (u == 0) mbar_arrive(q_empty + 8); mbar_arrive(k_empty); mbar_arrive(cls_bar); }
      return s256;
    };
    uint32_t n = 0;
    float s256 = 0.0f;
    if ((int)blockIdx.x < p.items) s256 = dots(0);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.heads, h = item - b * p.heads;
      const uint32_t vb = n & 1, pn = n & 1;
      VA_STAMP(warp, 0);
      mbar_wait_relaxed(bar_s + 8 * u, pn);
      tc_fence_after();
      VA_STAMP(warp, 3);
      // Both passes walk the row in 32-column chunks with the NEXT chunk's tcgen05.ld already in flight while the
      // current one is consumed (two register buffers, loops fully unrolled): a thread only ever has one round trip to
      // tensor memory exposed, not one per chunk (14 serialized ld+wait round trips per item were most of the period).
      uint32_t r0[32], r1[32];
      // pass 1: row maximum
      float mx = s256;
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
      VA_STAMP(warp, 4);
      const float m = mx * p.scale_log2;                 // scale > 0
      // pass 2: P = exp2(s*scale*log2e - m), rounded to fp16 (as the reference does under autocast), written in
      // place: chunk c of S (32 columns) becomes 16 packed columns that lie inside chunks already loaded
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
        const uint32_t dst = trow + (c < 4 ? c * 16 : 128 + (c - 4) * 16);
        tmem_ld_wait32(r0);
        tmem_ld32(trow + (c + 1) * 32, r1);
        chunk_exp(r0, dst);
        tmem_ld_wait32(r1);
        if (c + 2 < 8) tmem_ld32(trow + (c + 2) * 32, r0);
        chunk_exp(r1, dst + 16);
      }
      {
        // key 256: its probability as the A operand of a 17th P.V step (the 15 keys after it are zero rows of V)
        uint32_t pk[8] = {pack2n(ex2f(fmaf(s256, p.scale_log2, -m)), 0.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(trow + VA_P256_COL, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * u);
      VA_STAMP(warp, 5);

      // ---- query row 256: this warp's share of its P.V product (8 of the 88 head dims), in the shadow of the tile's
      //      P.V MMA ----
      {
        mbar_wait_relaxed(cls_p, pn);
        mbar_wait_relaxed(v_full + 8 * vb, (n >> 1) & 1);
        const float* pc = s_clsb + (n & 1) * VA_CLS_LD;
        row256_pv(warp, s_clsh + (n & 1) * VA_CLS_LD, gV0 + vb * VA_V_BYTES, 1.0f / pc[VA_KP],
                  p.o + b * p.o_bs + h * p.o_hs + (long long)(VA_N - 1) * p.o_ts);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(v_empty + 8 * vb);         // this warp has read its share of V
      VA_STAMP(warp, 1);
      // next item's 257th-token dot products, in the same shadow
      float s256_next = 0.0f;
      if (item + (int)gridDim.x < p.items) s256_next = dots(n + 1);
      VA_STAMP(warp, 2);
      mbar_wait_relaxed(bar_o + 8 * u, pn);
      tc_fence_after();
      VA_STAMP(warp, 6);
      __half* og = p.o + b * p.o_bs + h * p.o_hs + (long long)row * p.o_ts;
      {
        // dims 80..95 first: column 88 holds the row sum (ones-column of V), all 257 keys included.  Three loads in
        // flight at most (64 registers), each one overlapped with the stores of the previous.
        auto store8 = [&](const uint32_t* src, int chunk, float inv) {
          uint4 o;
          o.x = pack2n(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          o.y = pack2n(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          o.z = pack2n(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          o.w = pack2n(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          *reinterpret_cast<uint4*>(og + chunk * 8) = o;
        };
        uint32_t a1[16], b0[32];
        tmem_ld16(trow + VA_OHI_COL + 32, a1);             // dims 80..95
        tmem_ld32(trow + VA_OLO_COL, b0);                  // dims 0..31
        tmem_ld_wait16(a1);
        tmem_ld_wait32(b0);
        const float inv = 1.0f / __uint_as_float(a1[8]);
        uint32_t b1[16];
        tmem_ld16(trow + VA_OLO_COL + 32, b1);             // dims 32..47
#pragma unroll
        for (int g = 0; g < 4; ++g) store8(&b0[g * 8], g, inv);
        store8(&a1[0], 10, inv);                           // dims 80..87 (chunk 11 is padding)
        tmem_ld_wait16(b1);
        uint32_t a0[32];
        tmem_ld32(trow + VA_OHI_COL, a0);                  // dims 48..79
        store8(&b1[0], 4, inv);
        store8(&b1[8], 5, inv);
        tmem_ld_wait32(a0);
        // every column of O is in registers: hand the tile's TMEM columns back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 8 * u);
#pragma unroll
        for (int g = 0; g < 4; ++g) store8(&a0[g * 8], 6 + g, inv);
      }
      VA_STAMP(warp, 7);
      s256 = s256_next;
    }
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

// 3-D map over the packed projection buffer: (88 elements of a head | 48 head slots = q, k, v x 16 | token rows)
int make_qkv_tmap(CUtensorMap* tm, const void* base, long long rows, long long pitch_elems, int box_elems,
                         int box_rows, CUtensorMapSwizzle swz) {
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
  cuuint64_t gdim[3] = {(cuuint64_t)VA_D, 48, (cuuint64_t)rows};
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

int get_option(const char* key);

// TMA needs q, k, v to be the [B*257, 3, 16, 88] views of ONE projection buffer (what the fused qkv GEMM writes,
// eva_vit.py:133-138); any other strided layout takes the cp.async loaders
bool vit_attention_packed_qkv(const seedb200_attn_desc& d) {
  const __half* qp = static_cast<const __half*>(d.q);
  return d.heads == 16 && d.q_hs == VA_D && d.k_hs == VA_D && d.q_ts == d.k_ts && d.q_ts % 8 == 0 &&
         d.q_ts >= 48 * VA_D && d.q_bs == (int64_t)VA_N * d.q_ts && d.k_bs == d.q_bs &&
         static_cast<const __half*>(d.k) == qp + 16 * VA_D && (reinterpret_cast<uintptr_t>(qp) & 15) == 0;
}

int vit_attention_tc(const seedb200_attn_desc& d, cudaStream_t stream) {
  const __half* qp = static_cast<const __half*>(d.q);
  const bool tma = vit_attention_packed_qkv(d) && get_option("vit_attention_tma") != 0;
  auto kern = tma ? vit_attention_tc_kernel<true> : vit_attention_tc_kernel<false>;
  static bool attr_set_dev[SB_MAX_DEVICES][2] = {};   // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_dev[cur_device()][tma ? 1 : 0];
  if (!attr_set) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, VA_SMEM));
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
  CUtensorMap ta64, ta32, tr64, tr32;
  memset(&ta64, 0, sizeof(ta64)); ta32 = ta64; tr64 = ta64; tr32 = ta64;
  if (tma) {
    const long long rows = (long long)d.batch * VA_N;
    SB_PROPAGATE(make_qkv_tmap(&ta64, qp, rows, d.q_ts, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B));
    SB_PROPAGATE(make_qkv_tmap(&ta32, qp, rows, d.q_ts, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B));
    SB_PROPAGATE(make_qkv_tmap(&tr64, qp, rows, d.q_ts, 64, 1, CU_TENSOR_MAP_SWIZZLE_128B));
    SB_PROPAGATE(make_qkv_tmap(&tr32, qp, rows, d.q_ts, 32, 1, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  int grid = num_sms();
  if (grid > p.items) grid = p.items;
  profile_mark_begin(1, stream);
  kern<<<grid, VA_THREADS, VA_SMEM, stream>>>(p, ta64, ta32, tr64, tr32);
  profile_mark_end(1, stream, 4.0 * (double)d.batch * d.heads * (double)d.nq * d.nk * d.head_dim);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sb
