// attention_tc2.cu -- ViT-g/14 attention (257 x 257 tokens, 16 heads x 88) on tcgen05, staggered tile pipelines.
//
// Same arithmetic and the same tensor-memory layout as attention_tc.cu (eva_vit.py:139-156: S = Q K^T in fp32, exact
// two-pass softmax, P rounded to fp16, O = P V with P kept in tensor memory), different schedule.  The per-block
// timeline of attention_tc.cu (tools/attn_timeline.py, profiles/r02_attention_timeline.md) showed a 15.6 k-cycle period
// per (image, head) item against ~3.5 k cycles of MMA and ~4.1 k cycles of MUFU work: both 128-row query tiles were
// released by the same event (K landed), ran their softmax passes in lock step and fought for the one MUFU pipe, while
// everything that is not an exponential (the CUDA-core dot products of the 257th token, the max pass, the row-256
// P.V share, the O read-out) waited its turn in the same warps.  Here the two tiles are HALF A PERIOD APART:
//   * the MMA warp is a two-stream state machine; S of tile 1 is issued only after P.V of tile 0 of the same item, so
//     tile 1 runs its exponentials while tile 0 reads out / prepares the next item, and vice versa;
//   * Q and K arrive by TMA (swizzled K-major blocks, as in attention_tc.cu), so K -- single buffered, freed only when
//     S of tile 1 has retired -- is back ~1-3 k cycles later instead of ~8 k with cp.async;
//   * the softmax row sums come out of the P.V MMA itself (V's padding column 88 holds 1.0), which removes the
//     unpack + add per probability from the MUFU-bound pass;
//   * the 257th query row: scores by the 256 softmax threads, softmax by warp 13, P.V spread over the softmax threads
//     in the shadow of their tile's P.V MMA, final 8-way sum and store by warp 13 (no CTA-wide barrier anywhere).
#include "attention_tc_common.cuh"

namespace sb {

// Q and K live in the swizzled K-major blocks the TMA writes (see attention_tc.cu, TMA = true): per buffer a
// 128-byte-swizzled block of head dims 0..63 (1024 B per 8-row group) followed by a 64-byte-swizzled block of dims
// 64..95 (512 B per group).  Buffer order K, Q1, Q0, V keeps every swizzled block on its 1024 / 512-byte boundary.
constexpr int V2_K_G = 33, V2_Q0_G = 17, V2_Q1_G = 16;   // 8-row groups: keys 0..256 | rows 0..127 + row 256 | rows 128..255
constexpr int V2_K_BYTES = 34 * VA_G;               // 33 groups used; 34 keeps the next buffer 1024-byte aligned
constexpr int V2_Q1_BYTES = V2_Q1_G * VA_G;
constexpr int V2_Q0_BYTES = V2_Q0_G * VA_G;
constexpr int V2_V_BYTES = 33 * VA_G;               // keys 0..263, x2 buffers, no-swizzle core-matrix image (cp.async)
constexpr int V2_DATA_BYTES = V2_K_BYTES + V2_Q0_BYTES + V2_Q1_BYTES + 2 * V2_V_BYTES;
constexpr int V2_MISC_BYTES = 2 * VA_CLS_LD * 4 + 2 * 8 * VA_PART_LD * 4 + 256;
constexpr int V2_SMEM = V2_DATA_BYTES + V2_MISC_BYTES + 1024;

__device__ __forceinline__ void cp_async_commit_tc() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t pack2_nosum(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(VA_THREADS, 1)
vit_attention_tc2_kernel(const VitAttnParams p, const __grid_constant__ CUtensorMap tm_a64,
                         const __grid_constant__ CUtensorMap tm_a32, const __grid_constant__ CUtensorMap tm_r64,
                         const __grid_constant__ CUtensorMap tm_r32) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK0 = base, sQ1 = sK0 + V2_K_BYTES, sQ0 = sQ1 + V2_Q1_BYTES, sV0 = sQ0 + V2_Q0_BYTES;
  const uint32_t misc = sV0 + 2 * V2_V_BYTES;
  float* s_clsb = reinterpret_cast<float*>(gen + (misc - base));    // [2][VA_CLS_LD]: scores, then probabilities of row 256
  float* s_part = s_clsb + 2 * VA_CLS_LD;                           // [2][8][VA_PART_LD]: per-warp partial P.V of row 256
  const uint32_t bars = misc + 2 * VA_CLS_LD * 4 + 2 * 8 * VA_PART_LD * 4;
  const uint32_t bar_s = bars, bar_p = bars + 16, bar_o = bars + 32, bar_free = bars + 48;      // [2] each: per tile
  const uint32_t q_full = bars + 64 /*[2]*/, q_empty = bars + 80 /*[2]*/;
  const uint32_t k_full = bars + 96, k_empty = bars + 104, v_full = bars + 112 /*[2]*/, v_empty = bars + 128 /*[2]*/;
  const uint32_t cls_bar = bars + 144;           // 8 softmax warps -> warp 13: scores of query 256 are in s_cls
  const uint32_t cls_p = bars + 152;             // warp 13 -> softmax warps: probabilities of row 256 are in s_cls
  const uint32_t part_bar = bars + 160;          // 8 softmax warps -> warp 13: partial P.V of row 256 are in s_part
  const uint32_t tmem_slot = bars + 168;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  uint8_t* gQ0 = gen + (sQ0 - base);
  uint8_t* gQ1 = gen + (sQ1 - base);
  uint8_t* gV0 = gen + (sV0 - base);
  const uint8_t* gK = gen + (sK0 - base);
  // 16-byte chunk c (8 head dims) of row r of a swizzled Q / K buffer with `rows8` 8-row groups
  auto qk_chunk = [&](const uint8_t* buf, int rows8, int r, int c) -> const uint4* {
    if (c < 8) return reinterpret_cast<const uint4*>(buf + r * 128 + ((c ^ (r & 7)) << 4));
    return reinterpret_cast<const uint4*>(buf + rows8 * 1024 + r * 64 + (((c - 8) ^ ((r >> 1) & 3)) << 4));
  };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tma_prefetch_desc(&tm_a64); tma_prefetch_desc(&tm_a32); tma_prefetch_desc(&tm_r64); tma_prefetch_desc(&tm_r32);
    for (int u = 0; u < 2; ++u) {
      mbar_init(bar_s + 8 * u, 1); mbar_init(bar_p + 8 * u, 4); mbar_init(bar_o + 8 * u, 1); mbar_init(bar_free + 8 * u, 4);
      mbar_init(q_full + 8 * u, 1);        // the TMA's expect_tx arrive
      mbar_init(v_full + 8 * u, 1);
      mbar_init(v_empty + 8 * u, 9);       // P.V(1) retired + 8 softmax warps (row-256 share, value row 256)
    }
    mbar_init(q_empty, 10);                // S(0) retired + 8 softmax warps (tile-0 rows / query row 256) + warp 13
    mbar_init(q_empty + 8, 5);             // S(1) retired + 4 softmax warps of tile 1
 mbar_init(k_full, 1);
    mbar_init(k_empty, 10);                // S(1) retired + 8 softmax warps + warp 13 (key row 256)
    mbar_init(cls_bar, 8);
    mbar_init(cls_p, 1);
    mbar_init(part_bar, 8);
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<1>(tmem_slot, VA_TMEM_COLS);
  // zero every operand buffer once: the padding (head_dim 88..95, rows 257..263 of the last key group, rows 1..7 of
  // the query-256 group) is never written again ...
  for (uint32_t off = tid * 16; off < (uint32_t)V2_DATA_BYTES; off += VA_THREADS * 16)
    *reinterpret_cast<uint4*>(gen + off) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // ... except column 88 of V (first padding column), which holds 1.0 for keys 0..255: the P.V MMA then leaves the
  // row sum of the fp16 probabilities in column 88 of O (fp32), for free
  for (int i = tid; i < 2 * 256; i += VA_THREADS) {
    const int vb = i >> 8, k = i & 255;
    *reinterpret_cast<__half*>(gV0 + vb * V2_V_BYTES + (k >> 3) * VA_G + 11 * 128 + (k & 7) * 16) = __float2half(1.0f);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  constexpr uint32_t IDESC_S256 = make_idesc_f16(128, 256);
  constexpr uint32_t IDESC_O = make_idesc_f16(128, 48) | (1u << 16);      // half of the head dim; B (= V) is MN-major
  constexpr int CH = VA_D / 8;            // 11 16-byte chunks per row

  if (warp >= 8 && warp < 12) {
    // ======================= loaders: warp 8 = Q0 (+ query row 256), 9 = K, 10 = Q1 by TMA; warp 11 = V by cp.async ===========
    const int which = warp - 8;
    const int r8 = lane & 7, cq = lane >> 3;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.heads, h = item - b * p.heads;
      const uint32_t vb = n & 1, par = n & 1;
      VA_STAMP(10 + which, 0);
      if (which < 3) {
        const uint32_t empty = which == 0 ? q_empty : (which == 1 ? k_empty : q_empty + 8);
        const uint32_t full = which == 0 ? q_full : (which == 1 ? k_full : q_full + 8);
        mbar_wait_relaxed(empty, par ^ 1);
        VA_STAMP(10 + which, 1);
        if (lane == 0) {
          const int slot = (which == 1 ? 16 : 0) + h;            // q heads 0..15, k heads 16..31
          const int grow = b * VA_N;
          if (which == 0) {                                      // rows 0..127, then token 256 at local row 128
            const uint32_t b0 = sQ0, b1 = sQ0 + V2_Q0_G * 1024;
            mbar_arrive_expect_tx(full, 128 * 192 + 192);
            tma_load_3d(b0, &tm_a64, full, 0, slot, grow);
            tma_load_3d(b1, &tm_a32, full, 64, slot, grow);
            tma_load_3d(b0 + 128 * 128, &tm_r64, full, 0, slot, grow + 256);
            tma_load_3d(b1 + 128 * 64, &tm_r32, full, 64, slot, grow + 256);
          } else if (which == 1) {                               // keys 0..255 in two boxes, key 256 alone
            const uint32_t b0 = sK0, b1 = sK0 + V2_K_G * 1024;
            mbar_arrive_expect_tx(full, 256 * 192 + 192);
            for (int i = 0; i < 2; ++i) {
              tma_load_3d(b0 + i * 128 * 128, &tm_a64, full, 0, slot, grow + i * 128);
              tma_load_3d(b1 + i * 128 * 64, &tm_a32, full, 64, slot, grow + i * 128);
            }
            tma_load_3d(b0 + 256 * 128, &tm_r64, full, 0, slot, grow + 256);
            tma_load_3d(b1 + 256 * 64, &tm_r32, full, 64, slot, grow + 256);
          } else {                                               // rows 128..255
            const uint32_t b0 = sQ1, b1 = sQ1 + V2_Q1_G * 1024;
            mbar_arrive_expect_tx(full, 128 * 192);
            tma_load_3d(b0, &tm_a64, full, 0, slot, grow + 128);
            tma_load_3d(b1, &tm_a32, full, 64, slot, grow + 128);
          }
        }
        VA_STAMP(10 + which, 3);
        continue;
      }
      // V: no-swizzle core-matrix image (MN-major B operand of the P.V MMA), 16-byte cp.async copies
      mbar_wait_relaxed(v_empty + 8 * vb, ((n >> 1) & 1) ^ 1);
      VA_STAMP(10 + which, 1);
      const __half* rp = p.v + b * p.v_bs + h * p.v_hs + (long long)r8 * p.v_ts + cq * 8;
      uint32_t dp = sV0 + vb * V2_V_BYTES + cq * 128 + r8 * 16;
      for (int g = 0; g < 33; ++g) {
        if (g * 8 + r8 < VA_N) {
          cp_async16_tc(dp, rp);                                  // chunks 0..3
          cp_async16_tc(dp + 4 * 128, rp + 32);                   // chunks 4..7
          if (cq < 3) cp_async16_tc(dp + 8 * 128, rp + 64);       // chunks 8..10 (chunk 11: the ones column / padding)
        }
        rp += 8 * p.v_ts;
        dp += VA_G;
      }
      cp_async_wait_all_tc();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(v_full + 8 * vb);
      VA_STAMP(10 + which, 3);
    }
  } else if (warp == 12) {
    // ======================= MMA issuer: two streams (tile 0 / tile 1), whichever is ready goes =======================
    if (lane == 0) {
      int n_items = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) ++n_items;
      auto issue_s = [&](int u) {             // S_u = Q_u K^T (keys 0..255) into the tile's 256 TMEM columns
        const uint32_t qa = u == 0 ? sQ0 : sQ1;
        const uint32_t qa1 = qa + (u == 0 ? V2_Q0_G : V2_Q1_G) * 1024, ka1 = sK0 + V2_K_G * 1024;
#pragma unroll
        for (int j = 0; j < 4; ++j)              // head dims 0..63: +32 bytes per 16-element k-step inside the swizzle atom
          umma_f16<1>(tmem + u * VA_TILE_COLS, make_smem_desc_sw128(qa) + 2 * j, make_smem_desc_sw128(sK0) + 2 * j,
                      IDESC_S256, j > 0);
#pragma unroll
        for (int j = 0; j < 2; ++j)              // head dims 64..95 (88..95: the TMA's zero fill)
          umma_f16<1>(tmem + u * VA_TILE_COLS, make_smem_desc_sw64(qa1) + 2 * j, make_smem_desc_sw64(ka1) + 2 * j,
                      IDESC_S256, 1u);
        umma_commit<1>(bar_s + 8 * u);
        umma_commit<1>(q_empty + 8 * u);      // + the softmax warps' own arrivals: the Q rows may be overwritten
      };
      auto issue_pv = [&](int u, uint32_t vb) {   // O_u = P_u V: A = P in TMEM, B = V (MN-major), two 48-wide halves of d
        const uint32_t tb = tmem + u * VA_TILE_COLS;
        const uint32_t sV = sV0 + vb * V2_V_BYTES;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t pa = tb + (j < 8 ? j * 8 : 128 + (j - 8) * 8);
          umma_f16_ts(tb + VA_OLO_COL, pa, make_desc_nosw(sV + j * 2 * VA_G, VA_G, 128), IDESC_O, j > 0);
          umma_f16_ts(tb + VA_OHI_COL, pa, make_desc_nosw(sV + j * 2 * VA_G + 6 * 128, VA_G, 128), IDESC_O, j > 0);
        }
        umma_commit<1>(bar_o + 8 * u);
      };
      // stream state: step 2n = S(n), step 2n+1 = P.V(n)
      int st0 = 0, st1 = 0;
      const int end = 2 * n_items;
      const long long t_start = clock64();
      while (st0 < end || st1 < end) {
        bool progress = false;
        if (st0 < end) {
          const int n = st0 >> 1;
          const uint32_t par = n & 1;
          if ((st0 & 1) == 0) {
            if (mbar_test_wait(q_full, par) && mbar_test_wait(k_full, par) && (n == 0 || mbar_test_wait(bar_free, par ^ 1))) {
              tc_fence_after();
              VA_STAMP(8, 1);
              issue_s(0);
              ++st0; progress = true;
            }
          } else {
            if (mbar_test_wait(bar_p, par) && mbar_test_wait(v_full + 8 * (n & 1), (n >> 1) & 1)) {
              tc_fence_after();
              VA_STAMP(8, 5);
              issue_pv(0, n & 1);
              ++st0; progress = true;
            }
          }
        }
        if (st1 < end) {
          const int n = st1 >> 1;
          const uint32_t par = n & 1;
          if ((st1 & 1) == 0) {
            // the stagger: S of tile 1 only after P.V of tile 0 of the same item has been issued (st0 >= 2n + 2)
            if (st0 >= 2 * n + 2 && mbar_test_wait(q_full + 8, par) && (n == 0 || mbar_test_wait(bar_free + 8, par ^ 1))) {
              tc_fence_after();
              VA_STAMP(8, 2);
              issue_s(1);
              umma_commit<1>(k_empty);        // both S MMAs of the item have been issued: K may go once they retire
              ++st1; progress = true;
            }
          } else {
            if (mbar_test_wait(bar_p + 8, par)) {   // (V landed: P.V(0) of this item already waited for it)
              tc_fence_after();
              VA_STAMP(8, 6);
              issue_pv(1, n & 1);
              umma_commit<1>(v_empty + 8 * (n & 1));
              ++st1; progress = true;
            }
          }
        }
        if (!progress) {
          __nanosleep(20);
          if (clock64() - t_start > SB_MBAR_TIMEOUT_CYCLES) {
            printf("seedb200: vit_attention_tc2 MMA stream stalled block %d st0 %d st1 %d\n", blockIdx.x, st0, st1);
            __trap();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ======================= query row 256 (the 257th token): its softmax, and the final sum of its P.V =======================
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.heads, h = item - b * p.heads;
      float* cls = s_clsb + (n & 1) * VA_CLS_LD;
      VA_STAMP(9, 0);
      mbar_wait_relaxed(q_full, n & 1);
      mbar_wait_relaxed(k_full, n & 1);
      VA_STAMP(9, 1);
      // score of key 256: lanes 0..10 take one 8-dim chunk each
      float part = 0.0f;
      if (lane < CH) {
        const uint4 qa = *qk_chunk(gQ0, V2_Q0_G, 128, lane);    // query row 256 = local row 128 of the first Q buffer
        const uint4 ka = *qk_chunk(gK, V2_K_G, 256, lane);      // key row 256
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
      if (lane == 0) { mbar_arrive(k_empty); mbar_arrive(q_empty); }
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
        const float pr = __half2float(__float2half_rn(ex2f(sc[i] - mx)));   // P rounded to fp16 like the tile path
        sum += pr;
        if (key < VA_KP) cls[key] = pr;                    // 0 for keys 257..271
      }
      sum = warp_sum(sum);
      __syncwarp();
      if (lane == 0) mbar_arrive(cls_p);                  // release: the probabilities are visible to the waiters
      VA_STAMP(9, 4);
      // final sum of the 8 per-warp partial products, normalise, store (writers of the next use of this s_part
      // buffer wait for cls_p two items on, i.e. behind this read in program order)
      mbar_wait_relaxed(part_bar, n & 1);
      const float* pp = s_part + (n & 1) * 8 * VA_PART_LD;
      const float inv = 1.0f / sum;
      __half* og = p.o + b * p.o_bs + h * p.o_hs + (long long)(VA_N - 1) * p.o_ts;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int d = lane + 32 * i;
        if (d < VA_D) {
          float acc = 0.0f;
#pragma unroll
          for (int w = 0; w < 8; ++w) acc += pp[w * VA_PART_LD + d];
          og[d] = __float2half_rn(acc * inv);
        }
      }
      VA_STAMP(9, 5);
    }
  } else {
    // ======================= softmax + epilogue: warps 0-3 own tile 0 (rows 0..127), warps 4-7 tile 1 =======================
    const int quarter = warp & 3, u = warp >> 2;
    const int rl = quarter * 32 + lane;               // row inside the tile: one full row (256 + 1 keys) per thread
    const int row = u * 128 + rl;
    const uint32_t trow = tmem + u * VA_TILE_COLS + ((uint32_t)(quarter * 32) << 16);
    // ---- the 257th token on the CUDA cores: s256 = q_row . k_256 and t = q_256 . k_key (left in s_cls for warp 13).
    //      The dot products of item n+1 run in the shadow of item n's P.V MMA (needs only Q and K in shared memory).
    auto dots = [&](uint32_t nn) -> float {
      const uint32_t pq = nn & 1;
      mbar_wait_relaxed(q_full, pq);                                 // tile-0 rows and query row 256
      if (u == 1) mbar_wait_relaxed(q_full + 8, pq);
      mbar_wait_relaxed(k_full, pq);
      float s256 = 0.0f, t256 = 0.0f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint4 qa = *qk_chunk(u == 0 ? gQ0 : gQ1, u == 0 ? V2_Q0_G : V2_Q1_G, rl, c);   // this thread's query row
        const uint4 ka = *qk_chunk(gK, V2_K_G, 256, c);                                       // key 256
        const uint4 qb = *qk_chunk(gQ0, V2_Q0_G, 128, c);                                     // query 256
        const uint4 kb_ = *qk_chunk(gK, V2_K_G, row, c);                                      // key index == row index
        const __half2* q2 = reinterpret_cast<const __half2*>(&qa);
        const __half2* k2 = reinterpret_cast<const __half2*>(&ka);
        const __half2* q3 = reinterpret_cast<const __half2*>(&qb);
        const __half2* k3 = reinterpret_cast<const __half2*>(&kb_);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 qf = __half22float2(q2[j]), kf = __half22float2(k2[j]);
          const float2 qg = __half22float2(q3[j]), kg = __half22float2(k3[j]);
          s256 = fmaf(qf.x, kf.x, s256);
          s256 = fmaf(qf.y, kf.y, s256);
          t256 = fmaf(qg.x, kg.x, t256);
          t256 = fmaf(qg.y, kg.y, t256);
        }
      }
      s_clsb[(nn & 1) * VA_CLS_LD + row] = t256;
      __syncwarp();
      if (lane == 0) { mbar_arrive(q_empty); if (u == 1) mbar_arrive(q_empty + 8); mbar_arrive(k_empty); mbar_arrive(cls_bar); }
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
      // Both passes walk the row in 32-column chunks with the NEXT chunk's tcgen05.ld already in flight.
      uint32_t r0[32], r1[32];
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
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait32(r0);
        tmem_ld32(trow + (c + 1) * 32, r1);
        chunk_max(r0);
        tmem_ld_wait32(r1);
        tmem_ld32(trow + ((c + 2) & 7) * 32, r0);          // after chunk 7 this is chunk 0 again: pass 2's first load
        chunk_max(r1);
      }
      VA_STAMP(warp, 4);
      const float m = mx * p.scale_log2;                 // scale > 0
      // pass 2: P = exp2(s*scale*log2e - m) rounded to fp16, written in place behind the chunks already loaded.
      // (the row sum is not accumulated here: it is column 88 of O)
      auto chunk_exp = [&](const uint32_t(&cur)[32], uint32_t dst) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t pk[8];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            pk[g] = pack2_nosum(ex2f(fmaf(__uint_as_float(cur[hf * 16 + 2 * g]), p.scale_log2, -m)),
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
      const float p256 = __half2float(__float2half_rn(ex2f(fmaf(s256, p.scale_log2, -m))));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * u);
      VA_STAMP(warp, 5);

      // ---- query row 256: this warp's share of its P.V product, in the shadow of the tile's P.V MMA ----
      {
        mbar_wait_relaxed(cls_p, pn);
        mbar_wait_relaxed(v_full + 8 * vb, (n >> 1) & 1);
        const float* pc = s_clsb + (n & 1) * VA_CLS_LD;
        const uint8_t* gV = gV0 + vb * V2_V_BYTES;
        const int k8 = lane & 7, c4 = lane >> 3;
        float a[3][8];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) a[q][j] = 0.0f;
#pragma unroll 1
        for (int kg = warp; kg < 33; kg += 8) {            // keys 257..263 of the last group are zero rows, p = 0
          const float pk = pc[kg * 8 + k8];
          const uint8_t* vrow = gV + (uint32_t)kg * VA_G + k8 * 16 + c4 * 128;
#pragma unroll
          for (int q = 0; q < 3; ++q) {                    // chunk q*4 + c4 (chunk 11: the ones column, dropped below)
            const uint4 vv = *reinterpret_cast<const uint4*>(vrow + q * 512);
            const __half2* v2 = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 vf = __half22float2(v2[j]);
              a[q][2 * j] = fmaf(pk, vf.x, a[q][2 * j]);
              a[q][2 * j + 1] = fmaf(pk, vf.y, a[q][2 * j + 1]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v = a[q][j];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            a[q][j] = v;
          }
        if (k8 == 0) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q * 4 + c4 < CH) {
              float* dst = s_part + ((n & 1) * 8 + warp) * VA_PART_LD + (q * 4 + c4) * 8;
              *reinterpret_cast<float4*>(dst) = make_float4(a[q][0], a[q][1], a[q][2], a[q][3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(a[q][4], a[q][5], a[q][6], a[q][7]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(part_bar);
      }
      VA_STAMP(warp, 1);
      // next item's 257th-token dot products: now if its Q and K have already landed (tile 1: always; tile 0: K is
      // only freed when S of tile 1 of THIS item has retired), otherwise after the read-out below
      const bool has_next = item + (int)gridDim.x < p.items;
      float s256_next = 0.0f;
      bool dots_done = false;
      if (has_next) {
        const uint32_t pq = (n + 1) & 1;
        const bool ready = mbar_test_wait(k_full, pq) && mbar_test_wait(q_full, pq) && (u == 0 || mbar_test_wait(q_full + 8, pq));
        if (__all_sync(0xffffffffu, ready)) { s256_next = dots(n + 1); dots_done = true; }
      }
      VA_STAMP(warp, 2);

      mbar_wait_relaxed(bar_o + 8 * u, pn);
      tc_fence_after();
      VA_STAMP(warp, 6);
      uint32_t rsum;
      tmem_ld1(trow + VA_OHI_COL + (VA_D - 48), rsum);     // column 88 of O = sum of this row's fp16 probabilities
      tmem_ld_wait();
      const float inv = 1.0f / (__uint_as_float(rsum) + p256);
      const float w256 = p256 * inv;
      __half* og = p.o + b * p.o_bs + h * p.o_hs + (long long)row * p.o_ts;
      const uint8_t* v256 = gV0 + vb * V2_V_BYTES + 32 * VA_G;   // V row 256 = first row of group 32
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {                    // dims 0..47, then 48..95 (48..87 exist)
        uint32_t o0[32], o1[16];
        tmem_ld32(trow + (hh == 0 ? VA_OLO_COL : VA_OHI_COL), o0);
        tmem_ld16(trow + (hh == 0 ? VA_OLO_COL : VA_OHI_COL) + 32, o1);
        tmem_ld_wait();
        if (hh == 1) {
          // both halves of O are in registers / stored: hand the tile's TMEM columns back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_free + 8 * u);
        }
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          if (hh == 0 || g < 5) {                          // chunk 11 (dims 88..95) is padding
            const uint4 vv = *reinterpret_cast<const uint4*>(v256 + (hh * 6 + g) * 128);
            const __half2* v2 = reinterpret_cast<const __half2*>(&vv);
            float o8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o8[j] = __uint_as_float(g < 4 ? o0[g * 8 + j] : o1[(g - 4) * 8 + j]) * inv;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 vf = __half22float2(v2[j]);
              o8[2 * j] = fmaf(w256, vf.x, o8[2 * j]);
              o8[2 * j + 1] = fmaf(w256, vf.y, o8[2 * j + 1]);
            }
            uint4 o;
            o.x = pack2_nosum(o8[0], o8[1]); o.y = pack2_nosum(o8[2], o8[3]);
            o.z = pack2_nosum(o8[4], o8[5]); o.w = pack2_nosum(o8[6], o8[7]);
            *reinterpret_cast<uint4*>(og + (hh * 6 + g) * 8) = o;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(v_empty + 8 * vb);         // row-256 share and V row 256 have been read
      VA_STAMP(warp, 7);
      if (has_next && !dots_done) s256_next = dots(n + 1);
      s256 = s256_next;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<1>(tmem, VA_TMEM_COLS);
}

long long get_option64(const char* key);

int make_qkv_tmap(CUtensorMap* tm, const void* base, long long rows, long long pitch_elems, int box_elems, int box_rows,
                  CUtensorMapSwizzle swz, int slots);
bool vit_attention_packed_qkv(const seedb200_attn_desc& d);
int vit_attention_tc(const seedb200_attn_desc& d, cudaStream_t stream);

int vit_attention_tc2(const seedb200_attn_desc& d, cudaStream_t stream) {
  if (!vit_attention_packed_qkv(d)) return vit_attention_tc(d, stream);   // this variant takes Q / K by TMA only
  static bool attr_set_dev[SB_MAX_DEVICES] = {};   // cudaFuncSetAttribute is per device
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(vit_attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM));
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
  const long long rows = (long long)d.batch * VA_N;
  SB_PROPAGATE(make_qkv_tmap(&ta64, d.q, rows, d.q_ts, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 48));
  SB_PROPAGATE(make_qkv_tmap(&ta32, d.q, rows, d.q_ts, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 48));
  SB_PROPAGATE(make_qkv_tmap(&tr64, d.q, rows, d.q_ts, 64, 1, CU_TENSOR_MAP_SWIZZLE_128B, 48));
  SB_PROPAGATE(make_qkv_tmap(&tr32, d.q, rows, d.q_ts, 32, 1, CU_TENSOR_MAP_SWIZZLE_64B, 48));
  int grid = num_sms();
  if (grid > p.items) grid = p.items;
  profile_mark_begin(1, stream);
  vit_attention_tc2_kernel<<<grid, VA_THREADS, V2_SMEM, stream>>>(p, ta64, ta32, tr64, tr32);
  profile_mark_end(1, stream, 4.0 * (double)d.batch * d.heads * (double)d.nq * d.nk * d.head_dim);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sb
