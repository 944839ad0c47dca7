// attention_tc_common.cuh -- constants and inline-PTX helpers shared by the tcgen05 ViT attention kernels
// (attention_tc.cu: lock-step tile pipelines; attention_tc2.cu: staggered tile pipelines).
#pragma once

#include "common.cuh"
#include "ops.h"

namespace sb {
constexpr int VA_D = 88, VA_DP = 96, VA_N = 257, VA_KP = 272;
constexpr int VA_G = (VA_DP / 8) * 128;            // 1536: bytes of one 8-row group (12 chunks of 8 halves)
constexpr int VA_K_BYTES = 34 * VA_G;              // keys 0..271 (the N=256 MMA reads groups 0..31; row 256 sits in group 32)
constexpr int VA_Q0_BYTES = 16 * VA_G;             // query rows 0..127
constexpr int VA_Q1_BYTES = 17 * VA_G;             // query rows 128..255 and the group holding row 256
constexpr int VA_V_BYTES = 34 * VA_G;              // keys 0..271, x2 buffers (P.V walks 17 steps of 16 keys; rows 257.. stay zero)
constexpr int VA_DATA_BYTES = VA_K_BYTES + VA_Q0_BYTES + VA_Q1_BYTES + 2 * VA_V_BYTES;
constexpr int VA_CLS_LD = 288;                     // floats per row-256 probability buffer: 272 keys + the row's sum
constexpr int VA_PART_LD = 96;                     // floats per warp of row-256 P.V partials (88 dims, padded)
constexpr int VA_MISC_BYTES = 2 * VA_CLS_LD * 4 + 8 * VA_PART_LD * 4 + 256;   // row-256 buffers, partials / fp16 probabilities + s256 exchange, barriers
constexpr int VA_SMEM = VA_DATA_BYTES + VA_MISC_BYTES + 1024;   // slack: 1024-byte alignment of the swizzled blocks
constexpr int VA_THREADS = 448;                       // 8 softmax warps, 4 loader warps, MMA warp, row-256 warp
// (warp ids matter: the SM's arbiter favours high warp ids, so the two latency-critical single warps come last)
constexpr int VA_TMEM_COLS = 512;
// TMEM map of one tile pipeline u (base = 256 * u); everything aliases the 256 fp32 columns of S:
//   S      keys 0..255                      [0, 256)
//   P      keys 0..127 (fp16 x2 / column)   [0, 64)     written in place behind the S chunks already consumed
//          keys 128..255                    [128, 192)
//   O      dims 0..47                       [64, 112)   written by P.V after every S column has been read
//          dims 48..95                      [192, 240)  dim 88 is the ones-column of V: O[:, 88] = sum of the row's P
//   P      key 256 (+ 15 zero keys)         [112, 120)  written after the last S chunk has been consumed
constexpr int VA_TILE_COLS = 256;
constexpr int VA_OLO_COL = 64, VA_OHI_COL = 192, VA_P256_COL = 112;

struct VitAttnParams {
  const __half* q; const __half* k; const __half* v; __half* o;
  long long q_bs, q_hs, q_ts, k_bs, k_hs, k_ts, v_bs, v_hs, v_ts, o_bs, o_hs, o_ts;
  int items, heads;
  float scale_log2;
  int o_tma;         // attention_tc.cu: the output is written by TMA stores (tensor map passed next to the params)
  long long* dbg;    // optional timeline of block 0: [16 slots][64 items][8 events] clock64 stamps (tools/attn_timeline.py)
};

// No-swizzle canonical layouts (8 x 16-byte core matrices, 128 B each).  The shared-memory image used here is
//   byte(r, c) = (r / 8) * VA_G + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2      (r = token, c = head dim)
// * as a K-major operand (Q, K: contraction over c): LBO = 128 (next core matrix along K), SBO = VA_G;
// * as an MN-major B operand (V: contraction over r = keys, N = c): SBO = 128 (next 8 of N), LBO = VA_G (next
//   8 keys) -- the same bytes, so V needs no transposition.
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__device__ __forceinline__ void cp_async16_tc(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_tc() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
// tcgen05.wait::ld that also names the 32 destination registers of the chunk it completes: uses of r[] cannot be
// scheduled above the wait, which matters once the NEXT chunk's load is in flight while this one is consumed
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (P, fp16 packed two per 32-bit column) stays in tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float a, float b, float& sum) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  sum += f.x + f.y;
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ uint32_t pack2n(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// legacy-pipe tensor ops for the 257th token (a 1 x 257 row / column of the score matrix is matrix-vector work: the
// CUDA-core version cost ~1100 instructions per thread and item, profiles/r02_attention.md)
__device__ __forceinline__ void va_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void va_ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void va_mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                            uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

#define VA_STAMP(slot, ev)                                                                            \
  do {                                                                                                \
    if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && n < 64) p.dbg[((slot) * 64 + n) * 8 + (ev)] = clock64(); \
  } while (0)

}  // namespace sb
