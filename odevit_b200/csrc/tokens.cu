// The two ends of ViTNeuralODE.forward around the solve:
//   * token assembly, ode_transformer_gpt.py:148-182 -- the patch projection (Conv2d with kernel = stride = patch, as a
//     GEMM on im2col rows) lands DIRECTLY in the token tensor x0 [B, N, D] at its row offset with the bias and the
//     positional rows added in the GEMM's epilogue (EPI_TOKENS); the cls / distillation / register rows (parameters
//     broadcast over the batch) are written by one small kernel.  No torch.cat, no separate bias / pos_embed passes.
//   * head + loss, :588-589 and :626 -- logits = head(final[:, 0]) and CrossEntropy(label_smoothing) in one launch per
//     direction: a CTA per image forms the C logits from the CLS row, the log-sum-exp and the smoothed loss; the
//     backward forms d logits (from the loss and from an optional cotangent on the logits themselves), d cls, and the
//     weight / bias gradients.
#include "internal.h"

namespace odevit {

namespace {

__global__ void __launch_bounds__(256) special_rows_kernel(const float* __restrict__ rows, const int* __restrict__ index,
                                                           int n_special, int B, int N, int D, float* __restrict__ x0) {
  // one float4 per thread: (b, j, d4)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int D4 = D >> 2;
  if (i >= (long long)B * n_special * D4) return;
  const int d4 = (int)(i % D4);
  const int j = (int)((i / D4) % n_special);
  const long long b = i / ((long long)D4 * n_special);
  reinterpret_cast<float4*>(x0 + (b * N + index[j]) * D)[d4] = reinterpret_cast<const float4*>(rows + (long long)j * D)[d4];
}

// 1024 threads per CTA: the kernels are chains of dependent loads over a [classes, dim] weight that sits in L2 -- 32
// warps per image (forward: ~3 dot products per warp instead of 13) cut them from ~40 us to a few
constexpr int HEAD_THREADS = 1024;

__device__ __forceinline__ float warp_sum(float x) {
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// one CTA per image: logits[c] = <W[c], cls> + bias[c] (a warp per class, lanes over D), then lse and the loss
__global__ void __launch_bounds__(HEAD_THREADS, 1) head_ce_fwd_kernel(const float* __restrict__ x, long long x_stride, const float* __restrict__ W,
                                                                   const float* __restrict__ bias, const long long* __restrict__ labels,
                                                                   int C, int D, float eps, float* __restrict__ logits,
                                                                   float* __restrict__ loss, float* __restrict__ lse_out) {
  extern __shared__ float sh[];     // [D] cls row, [C] logits
  float* cls = sh;
  float* z = sh + D;
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int d = t; d < D; d += HEAD_THREADS) cls[d] = x[b * x_stride + d];
  __syncthreads();
  for (int c = warp; c < C; c += HEAD_THREADS / 32) {
    const float* w = W + (long long)c * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(w[d], cls[d], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + (bias ? bias[c] : 0.f);
      z[c] = v;
      logits[(long long)b * C + c] = v;
    }
  }
  __syncthreads();
  if (warp == 0 && loss) {
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f, sz = 0.f;
    for (int c = lane; c < C; c += 32) { se += expf(z[c] - mx); sz += z[c]; }
    se = warp_sum(se); sz = warp_sum(sz);
    if (lane == 0) {
      const float lse = mx + logf(se);
      const long long y = labels[b];
      // F.cross_entropy(label_smoothing = eps): (1 - eps) * (lse - z_y) + eps * (lse - mean_c z_c)
      loss[b] = (1.f - eps) * (lse - z[y]) + eps * (lse - sz / C);
      lse_out[b] = lse;
    }
  }
}

// d z[b, c] = g_logits[b, c] + g_loss / B * (softmax - (1 - eps) onehot - eps / C);  d cls_b = sum_c d z W[c]
__global__ void __launch_bounds__(HEAD_THREADS, 1) head_ce_bwd_rows_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                                                                        const long long* __restrict__ labels, const float* __restrict__ W,
                                                                        const float* __restrict__ g_logits, const float* __restrict__ g_loss,
                                                                        int B, int C, int D, float eps, float* __restrict__ dz,
                                                                        float* __restrict__ g_x, long long gx_stride) {
  extern __shared__ float sh[];     // [C] d z
  const int b = blockIdx.x, t = threadIdx.x;
  const float gl = g_loss ? (*g_loss) / B : 0.f;
  for (int c = t; c < C; c += HEAD_THREADS) {
    float v = g_logits ? g_logits[(long long)b * C + c] : 0.f;
    if (g_loss) {
      const float p = expf(logits[(long long)b * C + c] - lse[b]);
      v += gl * (p - (c == labels[b] ? 1.f - eps : 0.f) - eps / C);
    }
    sh[c] = v;
    dz[(long long)b * C + c] = v;
  }
  __syncthreads();
  if (g_x) {
    for (int d = t; d < D; d += HEAD_THREADS) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      int c = 0;
      for (; c + 4 <= C; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(sh[c + u], W[(long long)(c + u) * D + d], acc[u]);
      }
      for (; c < C; ++c) acc[0] = fmaf(sh[c], W[(long long)c * D + d], acc[0]);
      g_x[b * gx_stride + d] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
  }
}

// g_W[c, :] += sum_b dz[b, c] cls_b,  g_bias[c] += sum_b dz[b, c]     (a CTA per class, threads over D, images in order)
__global__ void __launch_bounds__(HEAD_THREADS, 1) head_ce_bwd_w_kernel(const float* __restrict__ dz, const float* __restrict__ x, long long x_stride,
                                                                     int B, int C, int D, float* __restrict__ g_W, float* __restrict__ g_bias) {
  const int c = blockIdx.x, t = threadIdx.x;
  for (int d = t; d < D; d += HEAD_THREADS) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int b = 0;
    for (; b + 4 <= B; b += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(dz[(long long)(b + u) * C + c], x[(b + u) * x_stride + d], acc[u]);
    }
    for (; b < B; ++b) acc[0] = fmaf(dz[(long long)b * C + c], x[b * x_stride + d], acc[0]);
    g_W[(long long)c * D + d] += (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  if (t == 0 && g_bias) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dz[(long long)b * C + c];
    g_bias[c] += acc;
  }
}

}  // namespace

int tokens_special_rows(const float* rows, const int* index, int n_special, int B, int N, int D, float* x0, cudaStream_t s) {
  if (n_special <= 0) return 0;
  if (D % 4) return set_error(ODEVIT_ERR_UNSUPPORTED, "token assembly needs D %% 4 == 0");
  ProfScope prof(KC_OTHER, s);
  const long long n = (long long)B * n_special * (D / 4);
  special_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rows, index, n_special, B, N, D, x0);
  ODV_LAUNCH_CHECK();
  return 0;
}

int head_ce_fwd(const float* x, long long x_stride, const float* W, const float* bias, const long long* labels, int B, int C, int D,
                float eps, float* logits, float* loss, float* lse, cudaStream_t s) {
  ProfScope prof(KC_OTHER, s);
  head_ce_fwd_kernel<<<B, HEAD_THREADS, (size_t)(D + C) * 4, s>>>(x, x_stride, W, bias, labels, C, D, eps, logits, loss, lse);
  ODV_LAUNCH_CHECK();
  return 0;
}

int head_ce_bwd(const float* x, long long x_stride, const float* W, const long long* labels, const float* logits, const float* lse,
                const float* g_logits, const float* g_loss, int B, int C, int D, float eps, float* dz, float* g_x,
                long long gx_stride, float* g_W, float* g_bias, cudaStream_t s) {
  ProfScope prof(KC_OTHER, s);
  head_ce_bwd_rows_kernel<<<B, HEAD_THREADS, (size_t)C * 4, s>>>(logits, lse, labels, W, g_logits, g_loss, B, C, D, eps, dz, g_x, gx_stride);
  ODV_LAUNCH_CHECK();
  if (g_W) {
    head_ce_bwd_w_kernel<<<C, HEAD_THREADS, 0, s>>>(dz, x, x_stride, B, C, D, g_W, g_bias);
    ODV_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace odevit
