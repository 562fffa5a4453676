// Fused GEMM epilogues shared by the FFMA and the tcgen05 GEMM kernels (element-wise form).
//
//   EPI_FWD1  packed in-proj + fc1:  q|k|v store, exact-erf GELU      (ode_transformer_gpt.py:193-200, :228)
//   EPI_RK    out_proj + fc2 + `*scaler` + Runge-Kutta stage combine   (:277, :320 + torchdiffeq step)
//   EPI_BWD3  d[O|h] with GELU' applied to the fc1 half
//   EPI_ACCUM weight-gradient accumulation
#pragma once
#include "internal.h"

namespace odevit {

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ void store_elem(void* p, long long idx, int type, float v) {
  if (type == DT_F32) reinterpret_cast<float*>(p)[idx] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
}
__device__ __forceinline__ float load_elem_rw(const void* p, long long idx, int type) {
  if (type == DT_F32) return reinterpret_cast<const float*>(p)[idx];
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
}

template <int EPI, bool ATOMIC>
__device__ __forceinline__ void epi_apply(const Epi& e, int m, int n, float acc, int zo, int zi) {
  if constexpr (EPI == EPI_STORE) {
    float v = e.alpha * acc;
    if (e.bias) v += e.bias[n];
    const long long idx = (long long)zo * e.out_bo + (long long)zi * e.out_bi + (long long)m * e.ld_out + n;
    store_elem(e.out, idx, e.out_type, v);
  } else if constexpr (EPI == EPI_FWD1) {
    float v = acc;
    if (e.bias) v += e.bias[n];
    if (n < e.split) {
      store_elem(e.out, (long long)m * e.ld_out + n, e.out_type, v);
    } else {
      const int c = n - e.split;
      if (e.out3) store_elem(e.out3, (long long)m * e.ld_out3 + c, e.aux_type, v);
      store_elem(e.out2, (long long)m * e.ld_out2 + c, e.aux_type, gelu_erf(v));
    }
  } else if constexpr (EPI == EPI_RK) {
    float v = acc;
    if (e.bias) v += e.bias[n];
    v *= e.alpha;
    const long long idx = (long long)m * e.ld_out + n;
    if (e.k_store) e.k_store[idx] = v;
    float r = e.c_new * v;
    if (e.y) r = fmaf(e.y_coef, e.y[idx], r);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (e.kin[i]) r = fmaf(e.c_k[i], e.kin[i][idx], r);
    reinterpret_cast<float*>(e.out)[idx] = r;
  } else if constexpr (EPI == EPI_BWD3) {
    if (n < e.split) {
      store_elem(e.out, (long long)m * e.ld_out + n, e.out_type, acc);
    } else {
      const int c = n - e.split;
      const float hp = load_elem_rw(e.aux, (long long)m * e.ld_aux + c, e.aux_type);
      store_elem(e.out2, (long long)m * e.ld_out2 + c, e.aux_type, acc * gelu_erf_grad(hp));
    }
  } else if constexpr (EPI == EPI_ACCUM) {
    float* o = reinterpret_cast<float*>(e.out) + (long long)m * e.ld_out + n;
    if constexpr (ATOMIC) atomicAdd(o, e.alpha * acc);
    else *o += e.alpha * acc;
  }
}

}  // namespace odevit
