// GPU-side replacement for the HF `ViTImageProcessor` call of the reference's collator
// (datasets/collator.py:11-22: `self.processor(pixel_values, return_tensors="pt")` on a list of PIL images):
// resize to S x S with PIL's BILINEAR resampling, rescale by 1/255, normalise per channel.
//
// PIL resamples 8-bit images in fixed point, horizontally then vertically, rounding to uint8 after EACH pass
// (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
// Vertical_8bpc).  That arithmetic is restated exactly: the coefficient tables are built on the host in double
// precision as Pillow builds them (odevit_pil_bilinear_tables), the kernels do the integer accumulation, so the
// resized uint8 image is bit-identical to Pillow's and `pixel_values` differs from the processor's by fp32 rounding
// of the normalisation only (tests/test_preprocess.py).  Uploading uint8 pixels instead of processed fp32 tensors cuts
// the host-to-device bytes per image from 602 KB to 3 KB (CIFAR) and takes the resize off the data-loader workers.
#include <cmath>

#include "internal.h"

namespace odevit {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Resample.c

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[b, y, xo, c] = clip8(round(sum_x img[b, y, xmin + x, c] * k[xo][x]))        (one thread per (b, y, xo))
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ tmp,
                                                         const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                         int ksize, long long rows, int W, int So) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * So) return;
  const int xo = (int)(i % So);
  const long long row = i / So;
  const int xmin = bounds[2 * xo], n = bounds[2 * xo + 1];
  const int32_t* k = kk + (long long)xo * ksize;
  const uint8_t* src = img + (row * W + xmin) * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int w = k[x];
    s0 += src[3 * x] * w; s1 += src[3 * x + 1] * w; s2 += src[3 * x + 2] * w;
  }
  uint8_t* dst = tmp + i * 3;
  dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

// out[b, c, yo, xo] = (clip8(round(sum_y tmp[b, ymin + y, xo, c] * k[yo][y])) * rescale - mean_c) / std_c
__global__ void __launch_bounds__(256) resample_v_norm_kernel(const uint8_t* __restrict__ tmp, float* __restrict__ out,
                                                              uint8_t* __restrict__ out_u8,
                                                              const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                              int ksize, int B, int H, int So_h, int So_w, float rescale,
                                                              float m0, float m1, float m2, float d0, float d1, float d2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)So_h * So_w;
  if (i >= (long long)B * per) return;
  const int xo = (int)(i % So_w);
  const int yo = (int)((i / So_w) % So_h);
  const long long b = i / per;
  const int ymin = bounds[2 * yo], n = bounds[2 * yo + 1];
  const int32_t* k = kk + (long long)yo * ksize;
  const uint8_t* src = tmp + ((b * H + ymin) * So_w + xo) * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < n; ++y) {
    const int w = k[y];
    const uint8_t* p = src + (long long)y * So_w * 3;
    s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
  }
  const uint8_t u0 = clip8(s0), u1 = clip8(s1), u2 = clip8(s2);
  if (out_u8) {
    uint8_t* q = out_u8 + i * 3;
    q[0] = u0; q[1] = u1; q[2] = u2;
  }
  float* o = out + b * 3 * per + (long long)yo * So_w + xo;
  o[0] = __fdiv_rn(__fsub_rn(__fmul_rn((float)u0, rescale), m0), d0);
  o[per] = __fdiv_rn(__fsub_rn(__fmul_rn((float)u1, rescale), m1), d1);
  o[2 * per] = __fdiv_rn(__fsub_rn(__fmul_rn((float)u2, rescale), m2), d2);
}

}  // namespace

int preprocess_u8(const uint8_t* images, int B, int H, int W, int So_h, int So_w, const int32_t* bounds_h,
                  const int32_t* kk_h, int ksize_h, const int32_t* bounds_v, const int32_t* kk_v, int ksize_v,
                  float rescale, const float* mean, const float* stdv, uint8_t* tmp, float* out, uint8_t* out_u8,
                  cudaStream_t s) {
  ProfScope prof(KC_OTHER, s);
  {
    const long long n = (long long)B * H * So_w;
    resample_h_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(images, tmp, bounds_h, kk_h, ksize_h, (long long)B * H, W, So_w);
    ODV_LAUNCH_CHECK();
  }
  {
    const long long n = (long long)B * So_h * So_w;
    resample_v_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tmp, out, out_u8, bounds_v, kk_v, ksize_v, B, H, So_h, So_w,
                                                                      rescale, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
    ODV_LAUNCH_CHECK();
  }
  return 0;
}

// Pillow's coefficient tables for BILINEAR (support 1.0) resampling of `in_size` samples to `out_size`, full box
// (Resample.c: precompute_coeffs + normalize_coeffs_8bpc), in HOST memory.  Plain C double arithmetic, as Pillow's.
int pil_bilinear_ksize(int in_size, int out_size) {
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

void pil_bilinear_tables(int in_size, int out_size, int32_t* bounds, int32_t* kk) {
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double kd[64];
    volatile double w_tmp;   // (keeps each product / difference a separately rounded double, as the x86-64 build does)
    int x = 0;
    for (; x < xmax && x < 64; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w_tmp = (a < 1.0) ? 1.0 - a : 0.0;
      kd[x] = w_tmp;
      ww += w_tmp;
    }
    for (x = 0; x < xmax && x < 64; ++x)
      if (ww != 0.0) kd[x] /= ww;
    for (x = 0; x < ksize; ++x) {
      const double v = (x < xmax && x < 64) ? kd[x] : 0.0;
      kk[(long long)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PRECISION_BITS)) : (int)(0.5 + v * (1 << PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

}  // namespace odevit
