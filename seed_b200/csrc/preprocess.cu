// preprocess.cu -- the image preprocessing in front of the tokenizer, on the GPU and bit-exact with the reference's
// CPU pipeline:
//   models/transforms.py:4-19            Resize((224,224)) [PIL BILINEAR] -> ToTensor -> Normalize(CLIP mean/std)
//   models/seed_llama_tokenizer.py:50-56 Resize((224,224), interpolation=3) [PIL BICUBIC] -> ToTensor -> Normalize
// followed by the `.half()` of ImageTokenizer.encode (:84-85).
//
// torchvision resizes a PIL image with Pillow's ImagingResample (8-bit path): per output coordinate a window of
// filter weights computed in double, normalised, converted to 22-bit fixed point; a horizontal pass and a vertical
// pass, each accumulating in int32 from 1 << 21 and storing clip8(acc >> 22) -- the intermediate image is 8-bit, so
// the two passes cannot be merged algebraically.  The same integer arithmetic runs here:
//   * host (create): the two weight tables, exactly as Pillow builds them (double), stored transposed
//     [tap][output coordinate] so that neighbouring threads read neighbouring words;
//   * resize_h_kernel: one CTA per (image, source row the vertical pass needs): the row (W x 3 bytes) is staged in
//     shared memory with 16-byte loads, thread xx accumulates its taps for the three channels, the 672-byte
//     output row leaves through shared memory as 16-byte stores;
//   * resize_v_norm_kernel: one CTA per (image, output row): vertical taps over the 8-bit intermediate (L2
//     resident), then ToTensor ((float)v / 255), Normalize ((x - mean) / std, fp32, IEEE division) and the fp16
//     rounding, written planar [3, S, S].
// HBM traffic: the source bytes once + 3*S*S*2 bytes out per image; the intermediate (rows x 672 B) stays in L2.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

struct seedb200_preprocess {
  int in_h, in_w, out, filter, max_batch;
  int ksh, ksv, y_first, tmp_rows;
  int2* bh; int32_t* kh;     // horizontal bounds [out], weights [ksh][out]
  int2* bv; int32_t* kv;     // vertical bounds [out] (ymin relative to y_first), weights [ksv][out]
  uint8_t* tmp;              // [max_batch][tmp_rows][out][3]
};

namespace sb {

constexpr int PP_PRECISION_BITS = 32 - 8 - 2;

static double pp_bilinear(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}
static double pp_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for the full-image box
static int pp_coeffs(int in_size, int out_size, int filter, std::vector<int2>& bounds, std::vector<int32_t>& kk_t) {
  double (*f)(double) = filter == 3 ? pp_bicubic : pp_bilinear;
  const double fsupport = filter == 3 ? 2.0 : 1.0;
  const double scale = (double)((float)in_size - 0.0f) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = fsupport * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  std::vector<double> k(ksize);
  bounds.resize(out_size);
  kk_t.assign((size_t)ksize * out_size, 0);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = f((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x];
      kk_t[(size_t)x * out_size + xx] = v < 0 ? (int)(-0.5 + v * (1 << PP_PRECISION_BITS)) : (int)(0.5 + v * (1 << PP_PRECISION_BITS));
    }
    bounds[xx] = make_int2(xmin, xmax);
  }
  return ksize;
}

__device__ __forceinline__ uint8_t pp_clip8(int v) {
  v >>= PP_PRECISION_BITS;
  return (uint8_t)min(255, max(0, v));
}

__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, long long image_stride, int in_w, int out, int y_first, int tmp_rows,
                const int2* __restrict__ bounds, const int32_t* __restrict__ kk_t, uint8_t* __restrict__ tmp) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const int y = blockIdx.x, img = blockIdx.y;
  const uint8_t* row = src + (long long)img * image_stride + (long long)(y + y_first) * in_w * 3;
  const int nbytes = in_w * 3;
  const int mis = (int)(reinterpret_cast<uintptr_t>(row) & 15);   // keep shared offsets congruent to the global address
  uint8_t* srow = pp_smem + mis;
  uint8_t* sout = pp_smem + ((16 + nbytes + 15) & ~15);
  {
    const int head = min(nbytes, (16 - mis) & 15);
    for (int i = threadIdx.x; i < head; i += 256) srow[i] = row[i];
    const int body = (nbytes - head) >> 4;
    const uint4* g = reinterpret_cast<const uint4*>(row + head);
    uint4* s = reinterpret_cast<uint4*>(srow + head);
    for (int i = threadIdx.x; i < body; i += 256) s[i] = __ldg(g + i);
    for (int i = head + body * 16 + threadIdx.x; i < nbytes; i += 256) srow[i] = row[i];
  }
  __syncthreads();
  for (int xx = threadIdx.x; xx < out; xx += 256) {
    const int2 b = bounds[xx];
    int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    const uint8_t* p = srow + b.x * 3;
    for (int x = 0; x < b.y; ++x) {
      const int k = __ldg(kk_t + (long long)x * out + xx);
      s0 += p[3 * x + 0] * k;
      s1 += p[3 * x + 1] * k;
      s2 += p[3 * x + 2] * k;
    }
    sout[xx * 3 + 0] = pp_clip8(s0);
    sout[xx * 3 + 1] = pp_clip8(s1);
    sout[xx * 3 + 2] = pp_clip8(s2);
  }
  __syncthreads();
  uint8_t* dst = tmp + ((long long)img * tmp_rows + y) * out * 3;
  const int obytes = out * 3;
  if ((obytes & 15) == 0) {
    for (int i = threadIdx.x; i < (obytes >> 4); i += 256)
      reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(sout)[i];
  } else {
    for (int i = threadIdx.x; i < obytes; i += 256) dst[i] = sout[i];
  }
}

__global__ void __launch_bounds__(256)
resize_v_norm_kernel(const uint8_t* __restrict__ tmp, int out, int tmp_rows, const int2* __restrict__ bounds,
                     const int32_t* __restrict__ kk_t, __half* __restrict__ dst) {
  const int yy = blockIdx.x, img = blockIdx.y;
  const int2 b = bounds[yy];
  const uint8_t* base = tmp + ((long long)img * tmp_rows + b.x) * out * 3;
  // torchvision: ToTensor -> float32 / 255; Normalize -> (x - mean) / std with float32 mean/std (transforms.py:16)
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  for (int xx = threadIdx.x; xx < out; xx += 256) {
    int s[3] = {1 << (PP_PRECISION_BITS - 1), 1 << (PP_PRECISION_BITS - 1), 1 << (PP_PRECISION_BITS - 1)};
    for (int y = 0; y < b.y; ++y) {
      const int k = __ldg(kk_t + (long long)y * out + yy);
      const uint8_t* p = base + ((long long)y * out + xx) * 3;
      s[0] += p[0] * k; s[1] += p[1] * k; s[2] += p[2] * k;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn((float)pp_clip8(s[c]), 255.0f);
      const float nrm = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
      dst[(((long long)img * 3 + c) * out + yy) * out + xx] = __float2half_rn(nrm);
    }
  }
}

}  // namespace sb

extern "C" {

int seedb200_preprocess_create(int in_h, int in_w, int out_size, int filter, int max_batch, seedb200_preprocess** out) {
  using namespace sb;
  if (!out) { set_error("preprocess_create: null output"); return SEEDB200_ERR_INVALID; }
  *out = nullptr;
  SB_REQUIRE(in_h > 0 && in_w > 0 && out_size > 0 && max_batch > 0, "preprocess_create: non-positive size");
  SB_REQUIRE(filter == 2 || filter == 3, "preprocess_create: filter %d (2 = PIL BILINEAR, 3 = PIL BICUBIC)", filter);
  SB_REQUIRE((long long)in_w * 3 + out_size * 3 + 64 <= 200 * 1024, "preprocess_create: image width %d too large", in_w);
  std::vector<int2> bh, bv;
  std::vector<int32_t> kh, kv;
  seedb200_preprocess* p = new seedb200_preprocess();
  memset(p, 0, sizeof(*p));
  p->in_h = in_h; p->in_w = in_w; p->out = out_size; p->filter = filter; p->max_batch = max_batch;
  p->ksh = pp_coeffs(in_w, out_size, filter, bh, kh);
  p->ksv = pp_coeffs(in_h, out_size, filter, bv, kv);
  p->y_first = bv[0].x;
  p->tmp_rows = bv[out_size - 1].x + bv[out_size - 1].y - p->y_first;
  for (auto& b : bv) b.x -= p->y_first;
  auto fail = [&](const char* what) {
    set_error("preprocess_create: %s failed", what);
    seedb200_preprocess_destroy(p);
    return SEEDB200_ERR_CUDA;
  };
  if (cudaMalloc(&p->bh, sizeof(int2) * out_size) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMalloc(&p->bv, sizeof(int2) * out_size) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMalloc(&p->kh, sizeof(int32_t) * kh.size()) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMalloc(&p->kv, sizeof(int32_t) * kv.size()) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMalloc(&p->tmp, (size_t)max_batch * p->tmp_rows * out_size * 3) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMemcpy(p->bh, bh.data(), sizeof(int2) * out_size, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->bv, bv.data(), sizeof(int2) * out_size, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->kh, kh.data(), sizeof(int32_t) * kh.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->kv, kv.data(), sizeof(int32_t) * kv.size(), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail("cudaMemcpy");
  *out = p;
  return 0;
}

void seedb200_preprocess_destroy(seedb200_preprocess* p) {
  if (!p) return;
  cudaFree(p->bh); cudaFree(p->bv); cudaFree(p->kh); cudaFree(p->kv); cudaFree(p->tmp);
  delete p;
}

int seedb200_preprocess_run(seedb200_preprocess* p, const void* images_u8, int n, void* out_f16, void* stream) {
  using namespace sb;
  SB_REQUIRE(p && images_u8 && out_f16, "preprocess_run: null argument");
  SB_REQUIRE(n > 0 && n <= p->max_batch, "preprocess_run: batch %d outside [1,%d]", n, p->max_batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)((16 + p->in_w * 3 + 15) & ~15) + (size_t)((p->out * 3 + 15) & ~15) + 16;
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  resize_h_kernel<<<dim3(p->tmp_rows, n), 256, smem, st>>>(static_cast<const uint8_t*>(images_u8),
                                                          (long long)p->in_h * p->in_w * 3, p->in_w, p->out, p->y_first,
                                                          p->tmp_rows, p->bh, p->kh, p->tmp);
  SB_LAUNCH_CHECK();
  resize_v_norm_kernel<<<dim3(p->out, n), 256, 0, st>>>(p->tmp, p->out, p->tmp_rows, p->bv, p->kv,
                                                        static_cast<__half*>(out_f16));
  SB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
