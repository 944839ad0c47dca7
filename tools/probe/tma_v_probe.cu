// probe: the 4-D TMA box (8 elems | rows | 11 chunks | slots) that writes V's no-swizzle core-matrix image, issued the
// way attention_tc.cu does (32 groups of 8 rows + one single row per item)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/tma_v_probe.bin tools/probe/tma_v_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
constexpr int G = 1536, NH = 34 * G / 2;
__global__ void probe(const __grid_constant__ CUtensorMap tm8, const __grid_constant__ CUtensorMap tm1, uint16_t* out, int grow, int slot) {
  extern __shared__ __align__(1024) uint16_t buf[];
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < NH; i += blockDim.x) buf[i] = 0xFFFF;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(33 * 1408));
    for (int g = 0; g < 32; ++g)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(d + g * G), "l"(&tm8), "r"(b), "r"(0), "r"(grow + 8 * g), "r"(0), "r"(slot) : "memory");
    // the 257th row alone: a 5-D view whose second dimension has extent 1 under a box of 8 -> rows 1..7 of the group are
    // out of bounds and zero-filled, so the chunk stride stays 128 bytes
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(d + 32 * G), "l"(&tm1), "r"(b), "r"(0), "r"(0), "r"(0), "r"(grow + 256), "r"(slot) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(0));
  __syncthreads();
  for (int i = threadIdx.x; i < NH; i += blockDim.x) out[i] = buf[i];
}
int main() {
  const int B = 3, rows = B * 257, pitch = 4224;
  std::vector<uint16_t> h((size_t)rows * pitch);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < pitch; ++c) h[(size_t)r * pitch + c] = (uint16_t)(((c / 88) * 1013 + r * 89 + (c % 88)) & 0xFFFF);
  uint16_t *g, *o; cudaMalloc(&g, h.size() * 2); cudaMalloc(&o, NH * 2);
  cudaMemcpy(g, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cuInit(0);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * G);
  for (int cfg = 0; cfg < 4; ++cfg) {
    CUtensorMap tm8, tm1;
    cuuint64_t gdim[4] = {8, (cuuint64_t)rows, 11, 48}; cuuint64_t gstr[3] = {pitch * 2, 16, 176};
    cuuint32_t box8[4] = {8, 8, 11, 1}, box1[4] = {8, 1, 11, 1}, es[4] = {1, 1, 1, 1};
    CUtensorMapDataType dt = (cfg & 1) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    CUtensorMapL2promotion pr = (cfg & 2) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    CUresult r1 = cuTensorMapEncodeTiled(&tm8, dt, 4, g, gdim, gstr, box8, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t gdim5[5] = {8, 1, 11, (cuuint64_t)rows, 48}; cuuint64_t gstr5[4] = {pitch * 2, 16, pitch * 2, 176};
    cuuint32_t box5[5] = {8, 8, 11, 1, 1}, es5[5] = {1, 1, 1, 1, 1};
    (void)box1;
    CUresult r2 = cuTensorMapEncodeTiled(&tm1, dt, 5, g, gdim5, gstr5, box5, es5, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int b = 1, slot = 32 + 7, grow = b * 257;
    probe<<<1, 256, 34 * G>>>(tm8, tm1, o, grow, slot);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> res(NH); cudaMemcpy(res.data(), o, NH * 2, cudaMemcpyDeviceToHost);
    int bad = 0, untouched_bad = 0;
    for (int i = 0; i < NH; ++i) {
      int grp = i / (G / 2), w = i % (G / 2), chunk = w / 64, r8 = (w % 64) / 8, e8 = w % 8, row = grp * 8 + r8;
      bool written = chunk < 11 && row < 257;
      uint16_t want = written ? (uint16_t)((slot * 1013 + (grow + row) * 89 + chunk * 8 + e8) & 0xFFFF)
                              : ((grp == 32 && chunk < 11) ? 0 : 0xFFFF);      // rows 257..263: zero-filled by the 5-D box
      if (res[i] != want) { if (written) { if (bad++ < 6) printf("  idx %d grp %d chunk %d r8 %d e %d: got %u want %u\n", i, grp, chunk, r8, e8, res[i], want); } else untouched_bad++; }
    }
    printf("cfg %d (dtype %s, promo %s): encode %d %d sync %d bad %d untouched-bad %d\n", cfg, (cfg & 1) ? "f16" : "u16", (cfg & 2) ? "128B" : "none", (int)r1, (int)r2, (int)e, bad, untouched_bad);
  }
  return 0;
}
