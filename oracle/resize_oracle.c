/* resize_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * CPU restatement of the image resize the reference runs before tokenisation:
 *   models/transforms.py:4-19          transforms.Resize((224, 224))                  -> PIL BILINEAR
 *   models/seed_llama_tokenizer.py:50-56  transforms.Resize((224, 224), interpolation=3) -> PIL BICUBIC
 * torchvision hands a PIL image to PIL.Image.resize, i.e. Pillow's ImagingResample (src/libImaging/Resample.c;
 * Pillow is a dependency of the reference through torchvision, requirements.txt, and is NOT vendored under
 * /root/reference).  Its published algorithm for 8-bit images, restated here:
 *   - per output coordinate xx: center = (xx + 0.5) * scale, support = filter_support * max(scale, 1),
 *     xmin = (int)(center - support + 0.5) clamped to 0, xmax = (int)(center + support + 0.5) clamped to the input
 *     size, weights w = filter((x + xmin - center + 0.5) / max(scale, 1)) in double, normalised by their sum;
 *   - weights are converted to fixed point with 22 fractional bits, rounding half away from zero;
 *   - horizontal pass over the rows the vertical pass needs, then vertical pass; each pass accumulates in int32
 *     starting from 1 << 21 and stores clip8(acc >> 22) -- the intermediate image is 8-bit.
 *   - pass order: horizontal first, EXCEPT (Pillow >= 11, PIL/Image.py `resize`: "if self.size[1] > self.size[0] * 100
 *     and size[1] < self.size[1]") for images more than 100 times taller than wide that shrink vertically, which
 *     are resized vertically first.  Older Pillow releases (the ones contemporary with the reference's
 *     requirements.txt) always ran horizontal first; the two orders differ by up to 19 grey levels on such images.
 *     The oracle follows the Pillow that is installed next to it, because that is what it is pinned against.
 * Pinned by tests/test_oracle.py::test_resize_oracle_matches_pillow against the installed Pillow itself
 * (random images, up- and down-scaling, both filters): bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

static double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

/* filter: 2 = PIL BILINEAR, 3 = PIL BICUBIC.  Returns ksize; bounds[2*out], kk[out*ksize] are malloc'ed. */
static int precompute(int in_size, int out_size, int filter, int** bounds_out, int32_t** kk_out) {
  double (*f)(double) = filter == 3 ? bicubic_filter : bilinear_filter;
  const double fsupport = filter == 3 ? 2.0 : 1.0;
  double scale = (double)((float)in_size - 0.0f) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = fsupport * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  double* pre = (double*)calloc((size_t)out_size * ksize, sizeof(double));
  int* bounds = (int*)malloc(sizeof(int) * 2 * out_size);
  int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * (size_t)out_size * ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double* k = pre + (size_t)xx * ksize;
    int x;
    for (x = 0; x < xmax; ++x) {
      const double w = f((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  for (size_t i = 0; i < (size_t)out_size * ksize; ++i) {
    if (pre[i] < 0) kk[i] = (int)(-0.5 + pre[i] * (1 << PRECISION_BITS));
    else kk[i] = (int)(0.5 + pre[i] * (1 << PRECISION_BITS));
  }
  free(pre);
  *bounds_out = bounds;
  *kk_out = kk;
  return ksize;
}

static uint8_t clip8(int32_t v) {
  v >>= PRECISION_BITS;           /* arithmetic shift, as Pillow's lookup index */
  if (v < 0) return 0;
  if (v > 255) return 255;
  return (uint8_t)v;
}

/* exported for the tests of the product's coefficient tables */
int resize_oracle_coeffs(int in_size, int out_size, int filter, int* bounds, int32_t* kk, int kk_capacity) {
  int* b; int32_t* k;
  const int ksize = precompute(in_size, out_size, filter, &b, &k);
  if ((long long)ksize * out_size > kk_capacity) { free(b); free(k); return -ksize; }
  memcpy(bounds, b, sizeof(int) * 2 * out_size);
  memcpy(kk, k, sizeof(int32_t) * (size_t)ksize * out_size);
  free(b); free(k);
  return ksize;
}

static int resize_two_pass(const uint8_t* src, int H, int W, int out_h, int out_w, int filter, uint8_t* dst) {
  int *bh, *bv; int32_t *kh, *kv;
  const int ksh = precompute(W, out_w, filter, &bh, &kh);
  const int ksv = precompute(H, out_h, filter, &bv, &kv);
  const int y_first = bv[0];
  const int y_last = bv[2 * out_h - 2] + bv[2 * out_h - 1];
  const int th = y_last - y_first;
  uint8_t* tmp = (uint8_t*)malloc((size_t)th * out_w * 3);
  for (int y = 0; y < th; ++y) {
    const uint8_t* line = src + (size_t)(y + y_first) * W * 3;
    for (int xx = 0; xx < out_w; ++xx) {
      const int xmin = bh[2 * xx], xmax = bh[2 * xx + 1];
      const int32_t* k = kh + (size_t)xx * ksh;
      int32_t s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
      for (int x = 0; x < xmax; ++x) {
        s0 += line[(x + xmin) * 3 + 0] * k[x];
        s1 += line[(x + xmin) * 3 + 1] * k[x];
        s2 += line[(x + xmin) * 3 + 2] * k[x];
      }
      uint8_t* o = tmp + ((size_t)y * out_w + xx) * 3;
      o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
    }
  }
  for (int yy = 0; yy < out_h; ++yy) {
    const int ymin = bv[2 * yy] - y_first, ymax = bv[2 * yy + 1];
    const int32_t* k = kv + (size_t)yy * ksv;
    for (int xx = 0; xx < out_w; ++xx) {
      int32_t s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
      for (int y = 0; y < ymax; ++y) {
        const uint8_t* p = tmp + ((size_t)(y + ymin) * out_w + xx) * 3;
        s0 += p[0] * k[y]; s1 += p[1] * k[y]; s2 += p[2] * k[y];
      }
      uint8_t* o = dst + ((size_t)yy * out_w + xx) * 3;
      o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
    }
  }
  free(tmp); free(bh); free(bv); free(kh); free(kv);
  return 0;
}

/* src [H, W, 3] uint8 (RGB, interleaved) -> dst [out_h, out_w, 3] uint8 */
int resize_oracle_u8(const uint8_t* src, int H, int W, int out_h, int out_w, int filter, uint8_t* dst) {
  if (!src || !dst || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || (filter != 2 && filter != 3)) return 1;
  if ((long long)H > (long long)W * 100 && out_h < H) {
    /* PIL/Image.py: vertical resize to (W, out_h) first, then the horizontal one */
    uint8_t* mid = (uint8_t*)malloc((size_t)out_h * W * 3);
    resize_two_pass(src, H, W, out_h, W, filter, mid);      /* width unchanged: identity horizontal weights */
    const int rc = resize_two_pass(mid, out_h, W, out_h, out_w, filter, dst);
    free(mid);
    return rc;
  }
  return resize_two_pass(src, H, W, out_h, out_w, filter, dst);
}
