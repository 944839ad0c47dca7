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
  int vertical_first;        // Pillow's Image.resize: images more than 100x taller than wide that shrink vertically
  int2* bh; int32_t* kh;     // horizontal bounds [out], weights [ksh][out]
  int2* bv; int32_t* kv;     // vertical bounds [out] (ymin relative to y_first), weights [ksv][out]
  uint8_t* tmp;              // [max_batch][tmp_rows][out][3]
  int device;
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

// torchvision: ToTensor -> float32 / 255; Normalize -> (x - mean) / std with float32 mean/std (transforms.py:15-16);
// then the .half() of ImageTokenizer.encode.  IEEE fp32 division and subtraction, round-to-nearest-even to fp16.
__device__ __forceinline__ __half pp_normalize(uint8_t v, int c) {
  const float mean = c == 0 ? 0.48145466f : (c == 1 ? 0.4578275f : 0.40821073f);
  const float stdv = c == 0 ? 0.26862954f : (c == 1 ? 0.26130258f : 0.27577711f);
  const float x = __fdiv_rn((float)v, 255.0f);
  return __float2half_rn(__fdiv_rn(__fsub_rn(x, mean), stdv));
}

constexpr int PP_ROWS = 16;      // max source rows per CTA of the horizontal pass (fewer, fatter CTAs: the
                                 // block scheduler was the limit at 4 rows / 1 row per CTA)
constexpr int PP_VROWS = 8;      // output rows per CTA of the vertical pass
constexpr int PP_MAX_TAPS = 16;  // taps on the register/word path; wider windows (bicubic downscale > 3.5x) finish on a byte loop

// Horizontal pass: rows [row0, row0 + PP_ROWS) of image blockIdx.y.  NORM = false: 8-bit intermediate
// dst8 [img][row][out][3]; NORM = true (vertical-first order): normalised fp16 planar dstf [img][3][rows][out].
template <bool NORM>
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, long long image_stride, int in_w, int out, int row_first, int rows,
                int rpc, const int2* __restrict__ bounds, const int32_t* __restrict__ kk_t, uint8_t* __restrict__ dst8,
                __half* __restrict__ dstf) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const int img = blockIdx.y, row0 = blockIdx.x * rpc;
  const int nrows = min(rpc, rows - row0);
  const int nbytes = in_w * 3;
  const int rstride = (nbytes + 16 + 15) & ~15;          // per-row shared-memory slot (16 bytes of alignment slack)
  const int obytes = out * 3;
  const int ostride = (obytes + 15) & ~15;
  uint8_t* sout = pp_smem + rpc * rstride;
  for (int r = 0; r < nrows; ++r) {
    const uint8_t* row = src + (long long)img * image_stride + (long long)(row0 + r + row_first) * nbytes;
    const int mis = (int)(reinterpret_cast<uintptr_t>(row) & 15);   // shared offset congruent to the global address
    uint8_t* srow = pp_smem + r * rstride + mis;
    const int head = min(nbytes, (16 - mis) & 15);
    for (int i = threadIdx.x; i < head; i += 256) srow[i] = row[i];
    const int body = (nbytes - head) >> 4;
    const uint4* g = reinterpret_cast<const uint4*>(row + head);
    uint4* s4 = reinterpret_cast<uint4*>(srow + head);
    for (int i = threadIdx.x; i < body; i += 256) s4[i] = __ldg(g + i);
    for (int i = head + body * 16 + threadIdx.x; i < nbytes; i += 256) srow[i] = row[i];
  }
  __syncthreads();
  for (int xx = threadIdx.x; xx < out; xx += 256) {
    const int2 b = bounds[xx];
    int kreg[PP_MAX_TAPS];
#pragma unroll
    for (int x = 0; x < PP_MAX_TAPS; ++x) kreg[x] = x < b.y ? __ldg(kk_t + (long long)x * out + xx) : 0;
    for (int r = 0; r < nrows; ++r) {
      const uint8_t* row = src + (long long)img * image_stride + (long long)(row0 + r + row_first) * nbytes;
      const uint8_t* p = pp_smem + r * rstride + (int)(reinterpret_cast<uintptr_t>(row) & 15) + b.x * 3;
      int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
      {
        // the window's 3 * taps bytes as aligned 32-bit shared loads + funnel shifts (one byte load per tap and
        // channel made the shared-memory pipe the bottleneck: 39 loads -> 13 for a 13-tap bicubic window)
        constexpr int NW = (3 * PP_MAX_TAPS + 3) / 4;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(3));
        const uint32_t sh = (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p)) & 3u) * 8u;
        uint32_t w[NW + 1];
#pragma unroll
        for (int k = 0; k <= NW; ++k) w[k] = (4 * k < 3 * b.y + 4) ? wp[k] : 0u;
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] = __funnelshift_r(w[k], w[k + 1], sh);
#pragma unroll
        for (int x = 0; x < PP_MAX_TAPS; ++x) {
          const int i0 = 3 * x, i1 = 3 * x + 1, i2 = 3 * x + 2;
          s0 += (int)((w[i0 >> 2] >> ((i0 & 3) * 8)) & 0xffu) * kreg[x];
          s1 += (int)((w[i1 >> 2] >> ((i1 & 3) * 8)) & 0xffu) * kreg[x];
          s2 += (int)((w[i2 >> 2] >> ((i2 & 3) * 8)) & 0xffu) * kreg[x];
        }
      }
      for (int x = PP_MAX_TAPS; x < b.y; ++x) {
        const int k = __ldg(kk_t + (long long)x * out + xx);
        s0 += p[3 * x + 0] * k;
        s1 += p[3 * x + 1] * k;
        s2 += p[3 * x + 2] * k;
      }
      if (NORM) {
        const long long o = ((long long)img * 3 * rows + (row0 + r)) * out + xx;
        dstf[o] = pp_normalize(pp_clip8(s0), 0);
        dstf[o + (long long)rows * out] = pp_normalize(pp_clip8(s1), 1);
        dstf[o + 2LL * rows * out] = pp_normalize(pp_clip8(s2), 2);
      } else {
        uint8_t* so = sout + r * ostride + xx * 3;
        so[0] = pp_clip8(s0); so[1] = pp_clip8(s1); so[2] = pp_clip8(s2);
      }
    }
  }
  if (!NORM) {
    __syncthreads();
    for (int r = 0; r < nrows; ++r) {
      uint8_t* dst = dst8 + ((long long)img * rows + row0 + r) * obytes;
      const uint8_t* so = sout + r * ostride;
      if ((obytes & 15) == 0) {
        for (int i = threadIdx.x; i < (obytes >> 4); i += 256)
          reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(so)[i];
      } else {
        for (int i = threadIdx.x; i < obytes; i += 256) dst[i] = so[i];
      }
    }
  }
}

// Vertical pass over rows of `width` pixels: output row blockIdx.x of image blockIdx.y.  NORM = true (horizontal-
// first order, width == out): normalised fp16 planar; NORM = false: 8-bit intermediate [img][out_rows][width][3].
template <bool NORM>
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ src, long long image_stride, int width, int out_rows,
                const int2* __restrict__ bounds, const int32_t* __restrict__ kk_t, uint8_t* __restrict__ dst8,
                __half* __restrict__ dstf) {
  const int img = blockIdx.y;
  const int rowbytes = width * 3;
  for (int yy = blockIdx.x * PP_VROWS; yy < min(out_rows, (blockIdx.x + 1) * PP_VROWS); ++yy) {
  const int2 b = bounds[yy];
  const uint8_t* base = src + (long long)img * image_stride + (long long)b.x * width * 3;
  if ((rowbytes & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 3) == 0) {
    // one thread per 32-bit word of the row: the vertical weights are the same for every byte of a row
    for (int j = threadIdx.x; j < (rowbytes >> 2); j += 256) {
      int s[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] = 1 << (PP_PRECISION_BITS - 1);
#pragma unroll 4
      for (int y = 0; y < b.y; ++y) {     // unrolled: the loads of 4 taps are in flight together (L2 latency bound)
        const int k = __ldg(kk_t + (long long)y * out_rows + yy);
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)y * rowbytes + 4 * j));
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] += (int)((w >> (8 * i)) & 0xffu) * k;
      }
      if (NORM) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int bi = 4 * j + i, xx = bi / 3, c = bi - 3 * xx;
          dstf[(((long long)img * 3 + c) * out_rows + yy) * width + xx] = pp_normalize(pp_clip8(s[i]), c);
        }
      } else {
        const uint32_t o = (uint32_t)pp_clip8(s[0]) | ((uint32_t)pp_clip8(s[1]) << 8) | ((uint32_t)pp_clip8(s[2]) << 16) |
                           ((uint32_t)pp_clip8(s[3]) << 24);
        *reinterpret_cast<uint32_t*>(dst8 + ((long long)img * out_rows + yy) * rowbytes + 4 * j) = o;
      }
    }
    continue;
  }
  for (int xx = threadIdx.x; xx < width; xx += 256) {
    int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
#pragma unroll 4
    for (int y = 0; y < b.y; ++y) {
      const int k = __ldg(kk_t + (long long)y * out_rows + yy);
      const uint8_t* p = base + ((long long)y * width + xx) * 3;
      s0 += p[0] * k; s1 += p[1] * k; s2 += p[2] * k;
    }
    if (NORM) {
      const long long o = ((long long)img * 3 * out_rows + yy) * width + xx;
      dstf[o] = pp_normalize(pp_clip8(s0), 0);
      dstf[o + (long long)out_rows * width] = pp_normalize(pp_clip8(s1), 1);
      dstf[o + 2LL * out_rows * width] = pp_normalize(pp_clip8(s2), 2);
    } else {
      uint8_t* o = dst8 + (((long long)img * out_rows + yy) * width + xx) * 3;
      o[0] = pp_clip8(s0); o[1] = pp_clip8(s1); o[2] = pp_clip8(s2);
    }
  }
  }
}

}  // namespace sb

extern "C" {

int seedb200_preprocess_create(int in_h, int in_w, int out_size, int filter, int max_batch, seedb200_preprocess** out) {
  return seedb200_preprocess_create_ex(in_h, in_w, out_size, out_size, 0, 0, out_size, filter, max_batch, out);
}

int seedb200_preprocess_create_ex(int in_h, int in_w, int resize_h, int resize_w, int crop_top, int crop_left,
                                  int out_size, int filter, int max_batch, seedb200_preprocess** out) {
  using namespace sb;
  if (!out) { set_error("preprocess_create: null output"); return SEEDB200_ERR_INVALID; }
  *out = nullptr;
  SB_REQUIRE(in_h > 0 && in_w > 0 && out_size > 0 && max_batch > 0, "preprocess_create: non-positive size");
  SB_REQUIRE(resize_h >= out_size && resize_w >= out_size && crop_top >= 0 && crop_left >= 0 &&
                 crop_top + out_size <= resize_h && crop_left + out_size <= resize_w,
             "preprocess_create: crop window %d+%d x %d+%d outside the %dx%d resize", crop_top, out_size, crop_left,
             out_size, resize_h, resize_w);
  SB_REQUIRE(filter == 2 || filter == 3, "preprocess_create: filter %d (2 = PIL BILINEAR, 3 = PIL BICUBIC)", filter);
  SB_REQUIRE((long long)(in_w * 3 + 48 + out_size * 3) + 64 <= 200 * 1024, "preprocess_create: image width %d too large", in_w);
  std::vector<int2> bh, bv;
  std::vector<int32_t> kh, kv;
  seedb200_preprocess* p = new seedb200_preprocess();
  memset(p, 0, sizeof(*p));
  p->in_h = in_h; p->in_w = in_w; p->out = out_size; p->filter = filter; p->max_batch = max_batch;
  p->device = cur_device();
  // coefficients of the full resize_w x resize_h resample, then only the crop window's columns / rows are kept: a
  // centre crop of a separable resample is the same resample evaluated on fewer output coordinates
  // (models/transforms.py:6-9 keep_ratio=True: Resize(S) -> CenterCrop(S))
  auto window = [&](int in_size, int full, int first, std::vector<int2>& b, std::vector<int32_t>& k) {
    std::vector<int2> fb; std::vector<int32_t> fk;
    const int ks = pp_coeffs(in_size, full, filter, fb, fk);
    b.assign(fb.begin() + first, fb.begin() + first + out_size);
    k.assign((size_t)ks * out_size, 0);
    for (int t = 0; t < ks; ++t)
      for (int i = 0; i < out_size; ++i) k[(size_t)t * out_size + i] = fk[(size_t)t * full + first + i];
    return ks;
  };
  p->ksh = window(in_w, resize_w, crop_left, bh, kh);
  p->ksv = window(in_h, resize_h, crop_top, bv, kv);
  // PIL/Image.py resize(): "if self.size[1] > self.size[0] * 100 and size[1] < self.size[1]" -> vertical pass first
  p->vertical_first = ((long long)in_h > (long long)in_w * 100 && resize_h < in_h) ? 1 : 0;
  if (p->vertical_first) {
    p->y_first = 0;
    p->tmp_rows = out_size;                                  // intermediate [out][in_w][3]
  } else {
    p->y_first = bv[0].x;
    p->tmp_rows = bv[out_size - 1].x + bv[out_size - 1].y - p->y_first;   // intermediate [tmp_rows][out][3]
    for (auto& b : bv) b.x -= p->y_first;
  }
  auto fail = [&](const char* what) {
    set_error("preprocess_create: %s failed", what);
    seedb200_preprocess_destroy(p);
    return SEEDB200_ERR_CUDA;
  };
  if (cudaMalloc(&p->bh, sizeof(int2) * out_size) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMalloc(&p->bv, sizeof(int2) * out_size) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMalloc(&p->kh, sizeof(int32_t) * kh.size()) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMalloc(&p->kv, sizeof(int32_t) * kv.size()) != cudaSuccess) return fail("cudaMalloc");
  const size_t tmp_bytes = (size_t)max_batch * p->tmp_rows * (p->vertical_first ? in_w : out_size) * 3;
  if (cudaMalloc(&p->tmp, tmp_bytes) != cudaSuccess) return fail("cudaMalloc");
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
  DeviceGuard guard(p->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* src = static_cast<const uint8_t*>(images_u8);
  __half* dst = static_cast<__half*>(out_f16);
  const size_t row_slot = (((size_t)p->in_w * 3 + 16 + 15) & ~(size_t)15) + (((size_t)p->out * 3 + 15) & ~(size_t)15);
  int rpc = (int)((200 * 1024 - 64) / row_slot);            // source rows per CTA that fit shared memory
  if (rpc > PP_ROWS) rpc = PP_ROWS;
  if (rpc < 1) rpc = 1;
  const size_t smem = (size_t)rpc * row_slot + 64;          // + word over-read slack
  static size_t attr_smem_dev[SB_MAX_DEVICES][2] = {};     // per device: cudaFuncSetAttribute is
  size_t* attr_smem = attr_smem_dev[cur_device()];
  if (attr_smem[0] == 0) attr_smem[0] = attr_smem[1] = 48 * 1024;
  if (smem > attr_smem[p->vertical_first]) {
    if (p->vertical_first)
      SB_CHECK_CUDA(cudaFuncSetAttribute(resize_h_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      SB_CHECK_CUDA(cudaFuncSetAttribute(resize_h_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[p->vertical_first] = smem;
  }
  const long long in_stride = (long long)p->in_h * p->in_w * 3;
  if (!p->vertical_first) {
    resize_h_kernel<false><<<dim3((p->tmp_rows + rpc - 1) / rpc, n), 256, smem, st>>>(
        src, in_stride, p->in_w, p->out, p->y_first, p->tmp_rows, rpc, p->bh, p->kh, p->tmp, nullptr);
    SB_LAUNCH_CHECK();
    resize_v_kernel<true><<<dim3((p->out + PP_VROWS - 1) / PP_VROWS, n), 256, 0, st>>>(p->tmp, (long long)p->tmp_rows * p->out * 3, p->out, p->out,
                                                           p->bv, p->kv, nullptr, dst);
    SB_LAUNCH_CHECK();
  } else {
    resize_v_kernel<false><<<dim3((p->out + PP_VROWS - 1) / PP_VROWS, n), 256, 0, st>>>(src, in_stride, p->in_w, p->out, p->bv, p->kv, p->tmp, nullptr);
    SB_LAUNCH_CHECK();
    resize_h_kernel<true><<<dim3((p->out + rpc - 1) / rpc, n), 256, smem, st>>>(
        p->tmp, (long long)p->out * p->in_w * 3, p->in_w, p->out, 0, p->out, rpc, p->bh, p->kh, nullptr, dst);
    SB_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
