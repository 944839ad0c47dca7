// vq.cu -- VectorQuantizer2.forward nearest-neighbour search (qformer_quantizer.py:94-98):
//     d = sum(z^2, dim=1, keepdim) + sum(e^2, dim=1) - 2 * einsum('bd,dn->bn', z, e^T);  ids = argmin(d, dim=1)
//
// Integer output => the arithmetic is pinned exactly (oracle/vq_oracle.c restates it in C and the parity
// tests are bit-exact against that):
//   * dot(z, e_j): ONE fp32 fma chain over d = 0..31 in index order (products of two fp16 values are exact
//     in fp32, so only the accumulation order matters -- tensor cores would make it implementation
//     defined, which is why this kernel stays on the CUDA cores; it is 0.003% of the encode FLOPs);
//   * |z|^2, |e_j|^2: fp32 sums in index order; in FP16 mode each square is first rounded to fp16
//     (torch materialises z**2 as an fp16 tensor) and the sum is rounded to fp16;
//   * FP16 mode (the reference's fp16 GPU path): C16 = fp16(dot); t = fp16(A16 + B16_j);
//     d = fp16(t - fp16(2*C16));  FP32 mode: d = (A + B_j) - 2*C in fp32, evaluated in that order;
//   * argmin: smallest d, ties -> lowest index (torch.argmin), NaN never produced by finite inputs.
//
// Mapping: a CTA owns 32 z rows (lane <-> row, the row lives in 32 fp32 registers).  The codebook is walked in
// tiles of 256 codes that the whole CTA converts to fp32 ONCE into shared memory (the conversion and |e_j|^2 are
// the same for every row: doing them per lane was 3/4 of the instructions of the first version, 610 us at
// 8192 x 8192); warp w then scores codes 32w..32w+31 of the tile for its 32 rows with broadcast 16-byte shared
// loads, four independent fma chains in flight.  Partial (d, id) pairs are merged lexicographically.
#include "common.cuh"

namespace sb {

constexpr int VQ_DIM = 32;
constexpr int VQ_WARPS = 8;
constexpr int VQ_TILE = 256;
constexpr int VQ_LD = 36;          // padded fp32 row (16-byte aligned, 4-way instead of 32-way conflicts column-wise)

__device__ __forceinline__ float round16(float x) { return __half2float(__float2half_rn(x)); }

// (d, id) as one 64-bit key whose unsigned order is the lexicographic order (smaller d first, then the lower index):
// the float is mapped to an order-preserving unsigned; NaN never wins (it never does in the strict `<` scan either)
__device__ __forceinline__ unsigned long long vq_key(float d, int id) {
  unsigned u = __float_as_uint(d);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  if (d != d) u = 0xffffffffu;
  return ((unsigned long long)u << 32) | (unsigned)id;
}

// SPLIT: blockIdx.y owns the code range [y * codes_per_split, ...) and the per-row winners of the splits meet in
// ids[] through a 64-bit atomicMin on the (d, id) key (ids[] preset to all ones, low word extracted by vq_finish_kernel).
// Few rows (a single image = 32 rows) otherwise leave one CTA walking all 8192 codes: 167 us for 64 rows.
template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(VQ_WARPS * 32)
vq_argmin_kernel(const __half* __restrict__ z, const __half* __restrict__ codebook, int n, int n_codes,
                 long long* __restrict__ ids, int codes_per_split) {
  __shared__ __align__(16) float s_e[VQ_TILE * VQ_LD];
  __shared__ float s_b[VQ_TILE];
  __shared__ float s_d[VQ_WARPS][32];
  __shared__ int s_i[VQ_WARPS][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = blockIdx.x * 32 + lane;
  const int rrow = row < n ? row : n - 1;

  float zr[VQ_DIM];
  {
    const uint4* zp = reinterpret_cast<const uint4*>(z + (long long)rrow * VQ_DIM);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 raw = zp[q];
      const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) zr[q * 8 + j] = __half2float(h[j]);
    }
  }
  float A = 0.0f;
#pragma unroll
  for (int d = 0; d < VQ_DIM; ++d) {
    float sq = __fmul_rn(zr[d], zr[d]);          // intrinsics: never contracted into an fma
    if (MODE == SEEDB200_VQ_FP16) sq = round16(sq);
    A = __fadd_rn(A, sq);
  }
  if (MODE == SEEDB200_VQ_FP16) A = round16(A);

  float best = INFINITY;
  int best_i = 0x7fffffff;
  const int code_begin = SPLIT ? blockIdx.y * codes_per_split : 0;
  const int code_end = SPLIT ? min(n_codes, code_begin + codes_per_split) : n_codes;
  for (int tile0 = code_begin; tile0 < code_end; tile0 += VQ_TILE) {
    __syncthreads();                               // the previous tile has been consumed
    // stage: 256 codes x 64 bytes = 1024 16-byte vectors, coalesced; fp16 -> fp32
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = tid + k * 256, code = idx >> 2, q = idx & 3;
      uint4 raw = make_uint4(0, 0, 0, 0);
      if (tile0 + code < n_codes) raw = __ldg(reinterpret_cast<const uint4*>(codebook + (long long)(tile0 + code) * VQ_DIM) + q);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
      const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
      float* dst = s_e + code * VQ_LD + q * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(f0.x, f0.y, f1.x, f1.y);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    __syncthreads();
    {   // |e|^2 of code `tid` of the tile, index order
      const float* e = s_e + tid * VQ_LD;
      float Bn = 0.0f;
#pragma unroll
      for (int d = 0; d < VQ_DIM; ++d) {
        float sq = __fmul_rn(e[d], e[d]);
        if (MODE == SEEDB200_VQ_FP16) sq = round16(sq);
        Bn = __fadd_rn(Bn, sq);
      }
      if (MODE == SEEDB200_VQ_FP16) Bn = round16(Bn);
      s_b[tid] = Bn;
    }
    __syncthreads();
    const int cbase = warp * 32;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const int c = tile0 + cbase + j;
      const float* e = s_e + (cbase + j) * VQ_LD;
      float C = 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(e + q * 4);       // broadcast
        C = __fmaf_rn(zr[q * 4 + 0], v.x, C);
        C = __fmaf_rn(zr[q * 4 + 1], v.y, C);
        C = __fmaf_rn(zr[q * 4 + 2], v.z, C);
        C = __fmaf_rn(zr[q * 4 + 3], v.w, C);
      }
      const float Bn = s_b[cbase + j];
      float dist;
      if (MODE == SEEDB200_VQ_FP16) {
        const float C16 = round16(C);
        const float t = round16(__fadd_rn(A, Bn));
        dist = round16(__fsub_rn(t, round16(__fmul_rn(2.0f, C16))));
      } else {
        const float t = __fadd_rn(A, Bn);
        dist = __fsub_rn(t, __fmul_rn(2.0f, C));
      }
      if (c < code_end && dist < best) { best = dist; best_i = c; }   // ascending c per warp: strict < keeps the lowest index
    }
  }
  s_d[warp][lane] = best;
  s_i[warp][lane] = best_i;
  __syncthreads();
  if (warp == 0 && row < n) {
    float bd = s_d[0][lane];
    int bi = s_i[0][lane];
#pragma unroll
    for (int w = 1; w < VQ_WARPS; ++w) {
      const float dd = s_d[w][lane];
      const int ii = s_i[w][lane];
      if (dd < bd || (dd == bd && ii < bi)) { bd = dd; bi = ii; }   // ties -> lowest index (torch.argmin)
    }
    if constexpr (SPLIT) {
      if (bi != 0x7fffffff) atomicMin(reinterpret_cast<unsigned long long*>(ids) + row, vq_key(bd, bi));
    } else {
      if (bi == 0x7fffffff) bi = 0;                  // all distances NaN/inf: torch.argmin returns 0 for all-inf
      ids[row] = (long long)bi;
    }
  }
}

__global__ void vq_finish_kernel(long long* __restrict__ ids, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = reinterpret_cast<unsigned long long*>(ids)[i];
  ids[i] = (k == ~0ull) ? 0LL : (long long)(k & 0xffffffffull);     // no finite / inf distance at all: 0 like the scan
}

int vq_argmin(const void* z, const void* codebook, int n, int n_codes, int dim, int mode, int64_t* ids,
              cudaStream_t stream) {
  SB_REQUIRE(z && codebook && ids, "vq_argmin: null operand");
  SB_REQUIRE(dim == VQ_DIM, "vq_argmin: dim=%d unsupported (codebook_embed_dim is 32)", dim);
  SB_REQUIRE(n > 0 && n_codes > 0, "vq_argmin: empty problem");
  SB_REQUIRE(mode == SEEDB200_VQ_FP16 || mode == SEEDB200_VQ_FP32, "vq_argmin: unknown mode %d", mode);
  const int blocks = (n + 31) / 32;
  const int tiles = (n_codes + VQ_TILE - 1) / VQ_TILE;
  const __half* zp = static_cast<const __half*>(z);
  const __half* cp = static_cast<const __half*>(codebook);
  long long* ip = reinterpret_cast<long long*>(ids);
  // few row blocks: split the codebook over blockIdx.y so that about one CTA per SM is busy
  int splits = 1;
  if (blocks * 2 <= num_sms() && tiles > 1) {
    splits = num_sms() / blocks;
    if (splits > tiles) splits = tiles;
  }
  if (splits > 1) {
    const int tiles_per_split = (tiles + splits - 1) / splits;
    splits = (tiles + tiles_per_split - 1) / tiles_per_split;
    const int cps = tiles_per_split * VQ_TILE;
    SB_CHECK_CUDA(cudaMemsetAsync(ip, 0xff, (size_t)n * sizeof(long long), stream));
    if (mode == SEEDB200_VQ_FP16)
      vq_argmin_kernel<SEEDB200_VQ_FP16, true><<<dim3(blocks, splits), VQ_WARPS * 32, 0, stream>>>(zp, cp, n, n_codes, ip, cps);
    else
      vq_argmin_kernel<SEEDB200_VQ_FP32, true><<<dim3(blocks, splits), VQ_WARPS * 32, 0, stream>>>(zp, cp, n, n_codes, ip, cps);
    SB_LAUNCH_CHECK();
    vq_finish_kernel<<<(n + 255) / 256, 256, 0, stream>>>(ip, n);
    SB_LAUNCH_CHECK();
    return 0;
  }
  if (mode == SEEDB200_VQ_FP16)
    vq_argmin_kernel<SEEDB200_VQ_FP16, false><<<blocks, VQ_WARPS * 32, 0, stream>>>(zp, cp, n, n_codes, ip, n_codes);
  else
    vq_argmin_kernel<SEEDB200_VQ_FP32, false><<<blocks, VQ_WARPS * 32, 0, stream>>>(zp, cp, n, n_codes, ip, n_codes);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sb

extern "C" int seedb200_vq_argmin(const void* z, const void* codebook, int n, int n_codes, int dim, int mode,
                                  int64_t* ids, void* stream) {
  return sb::vq_argmin(z, codebook, n, n_codes, dim, mode, ids, static_cast<cudaStream_t>(stream));
}
