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
    for (int i = 0; i < Epi::kMaxTerms; ++i)
      if (e.kin[i]) r = fmaf(e.c_k[i], e.kin[i][idx], r);
    if (e.out) reinterpret_cast<float*>(e.out)[idx] = r;
    if (e.out2) store_elem(e.out2, idx, e.aux_type, e.out2_scale * r);
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

// ---- vectorised form for the tcgen05 kernels: one thread holds 16 consecutive columns
// [n, n+16) of row m (n % 16 == 0, all 16 in range; row pointers 16-byte aligned). -------------
__device__ __forceinline__ void store16(void* p, long long idx, int type, const float* v) {
  if (type == DT_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + idx);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + idx);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
__device__ __forceinline__ void load16(const void* p, long long idx, int type, float* v) {
  if (type == DT_F32) {
    const float4* in = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + idx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = in[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
    const uint4* in = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + idx);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 t = in[h];
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        v[8 * h + 2 * i] = __low2float(b);
        v[8 * h + 2 * i + 1] = __high2float(b);
      }
    }
  }
}

template <int EPI, bool ATOMIC>
__device__ __forceinline__ void epi_chunk16(const Epi& e, int m, int n, float* v) {
  if constexpr (EPI == EPI_STORE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = e.alpha * v[j] + (e.bias ? e.bias[n + j] : 0.f);
    store16(e.out, (long long)m * e.ld_out + n, e.out_type, v);
  } else if constexpr (EPI == EPI_FWD1) {
    if (e.bias) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += e.bias[n + j];
    }
    if (n < e.split) {
      store16(e.out, (long long)m * e.ld_out + n, e.out_type, v);
    } else {
      const int c = n - e.split;
      if (e.out3) store16(e.out3, (long long)m * e.ld_out3 + c, e.aux_type, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
      store16(e.out2, (long long)m * e.ld_out2 + c, e.aux_type, v);
    }
  } else if constexpr (EPI == EPI_RK) {
    const long long idx = (long long)m * e.ld_out + n;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = e.alpha * (v[j] + (e.bias ? e.bias[n + j] : 0.f));
    if (e.k_store) store16(e.k_store, idx, DT_F32, v);
    float r[16], t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = e.c_new * v[j];
    if (e.y) {
      load16(e.y, idx, DT_F32, t);
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = fmaf(e.y_coef, t[j], r[j]);
    }
#pragma unroll
    for (int i = 0; i < Epi::kMaxTerms; ++i) {
      if (e.kin[i]) {
        load16(e.kin[i], idx, DT_F32, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = fmaf(e.c_k[i], t[j], r[j]);
      }
    }
    if (e.out) store16(e.out, idx, DT_F32, r);
    if (e.out2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] *= e.out2_scale;
      store16(e.out2, idx, e.aux_type, r);
    }
  } else if constexpr (EPI == EPI_BWD3) {
    if (n < e.split) {
      store16(e.out, (long long)m * e.ld_out + n, e.out_type, v);
    } else {
      const int c = n - e.split;
      float hp[16];
      load16(e.aux, (long long)m * e.ld_aux + c, e.aux_type, hp);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= gelu_erf_grad(hp[j]);
      store16(e.out2, (long long)m * e.ld_out2 + c, e.aux_type, v);
    }
  } else if constexpr (EPI == EPI_ACCUM) {
    float* o = reinterpret_cast<float*>(e.out) + (long long)m * e.ld_out + n;
    if constexpr (ATOMIC) {
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(o + j, e.alpha * v[j]);
    } else {
      float t[16];
      load16(o, 0, DT_F32, t);
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = fmaf(e.alpha, v[j], t[j]);
      store16(o, 0, DT_F32, t);
    }
  }
}

}  // namespace odevit
