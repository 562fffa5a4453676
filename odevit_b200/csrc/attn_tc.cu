// Fused softmax attention on tcgen05 tensor cores (bf16 mode, head dim 64, N <= 256 tokens).
//
//   per (image b, head h):  S = q k^T  ->  P = softmax_keys(S)  ->  O = P v
//
// replaces the bmm / softmax / bmm / head split+merge copies of nn.MultiheadAttention's explicit
// path (models/ode_transformer_gpt.py:226-232).  q already carries the 1/sqrt(d) (folded into the
// in-proj weight rows, rows.cu::fold_w1_kernel).
//
// One CTA (128 threads) per (b, h, 128-query tile); two CTAs are resident per SM (<= 81 KB of
// shared memory and 256 tensor-memory columns each) so that the TMA loads / MMAs of one overlap
// the softmax of the other.
//   thread 0      TMA: Q tile [128 x 64], K and V [NP x 64] of this (b, h) out of the packed
//                 [B, N, 3D] qkv buffer (3-D tensor map: rows past N are zero-filled);
//                 MMA1 (SS): S[128 x NP] = Q K^T into TMEM columns [0, NP)
//   all 4 warps   thread = query row: max and sum over its S row straight from TMEM, then the
//                 un-normalised exp() is packed to bf16 and written back IN PLACE over the S
//                 columns [0, NP/2) (tcgen05.st) -- P never touches shared or global memory
//   thread 0      MMA2 (TS): O[128 x 64] = P (TMEM) * V (smem, MN-major) into columns [128, 192)
//   all 4 warps   O row * 1/sum -> bf16 -> columns [h*64, h*64+64) of the [O | h] buffer
// When the caller can observe P (block.attentions, attention_trajectory, the JaSMin window) the
// EXPORT variant also writes the normalised fp32 P[b, h, :, :].
#include <cuda.h>

#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace odevit {

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1_elems,
                      uint64_t ld2_elems, uint32_t b0, uint32_t b1, uint32_t b2);

namespace {

constexpr int HD = 64;        // head dim
constexpr int BMQ = 128;      // query rows per CTA
constexpr int O_COL = 128;    // TMEM column of the O accumulator
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnArgs {
  int B, N, H, NP, tiles_m;
  int D;                 // embed dim (column offsets of k, v inside the packed row)
  void* oh;              // [B*N, ld_oh] bf16; O goes to columns [h*64, h*64+64)
  long long ld_oh;
  float* p_out;          // [B, H, N, N] fp32 or null
  float* lse_out;        // [B, H, N] fp32 or null: log2-domain log-sum-exp of each row (for the VJP)
  Drop drop;             // attention-map dropout (element (b*H+h)*N + i, j); thresh 0 = off
  // JaSMin statistic of this evaluation's map without exporting it (ode_transformer_gpt.py:419-456): per (b, h)
  // the max over query rows of log(g_1 / (g_k + 1e-12) + 1e-12), atomically maxed into jas_out[b*H + h]
  float* jas_out;        // [B, H] fp32 (pre-set to -inf) or null
  int jas_k;             // 0..3
};

// running 4 largest values (descending), branch-free insertion
__device__ __forceinline__ void top4_insert(float (&t)[4], float v) {
  float a;
  a = fmaxf(t[0], v); v = fminf(t[0], v); t[0] = a;
  a = fmaxf(t[1], v); v = fminf(t[1], v); t[1] = a;
  a = fmaxf(t[2], v); v = fminf(t[2], v); t[2] = a;
  t[3] = fmaxf(t[3], v);
}
// float atomic max through the ordered-integer views (the target starts at -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Softmax of this thread's query row straight from tensor memory (S at t_row, n_chunks 16-column chunks; 0 for a
// warp without valid rows), un-normalised exp packed to bf16 IN PLACE over the consumed S columns; EXPORT also
// writes the normalised fp32 row.  Returns 1 / row sum (what the O epilogue needs; 1 in EXPORT mode's P).
template <bool EXPORT, bool DROP, bool JAS = false>
__device__ __forceinline__ float attn_softmax_row(const AttnArgs& a, uint32_t t_row, int n_chunks, int b, int h, int row,
                                                  float* stage = nullptr) {
  constexpr float LOG2E = 1.4426950408889634f;
  // Tensor-memory loads are issued several chunks at a time and waited for once, and the row max / row sum run
  // as four independent chains: with one row per thread and two warps per scheduler, a load / wait round trip
  // per 16 columns and one 208-long dependent FMNMX / FADD chain per pass left the passes latency-bound.
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float tt[JAS ? 4 : 1][4];     // JAS: the 4 largest logits of each of the four chains
  if constexpr (JAS) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
      for (int i = 0; i < 4; ++i) tt[q4][i] = -INFINITY;
  }
  // (the chunk that straddles N is peeled off as a tail, as in attn_softmax_row_fast: the others carry no column tests;
  //  loads may run past NP inside the slot's 256 columns)
  const int last = n_chunks - 1;            // the chunk that holds column N - 1
  const int nv = a.N - last * 16;           // its key columns, 1..16
  for (int c0 = 0; c0 < last; c0 += 4) {
    float v[4][16];
#pragma unroll
    for (int u = 0; u < 4; ++u) ptx::tmem_ld16(t_row + (c0 + u) * 16, v[u]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c0 + u < last) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if constexpr (JAS) top4_insert(tt[j & 3], v[u][j]);
          else m4[j & 3] = fmaxf(m4[j & 3], v[u][j]);
        }
      }
  }
  if (n_chunks > 0) {
    float v[16];
    ptx::tmem_ld16(t_row + last * 16, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x = (j < nv) ? v[j] : -INFINITY;
      if constexpr (JAS) top4_insert(tt[j & 3], x);
      else m4[j & 3] = fmaxf(m4[j & 3], x);
    }
  }
  if constexpr (JAS) {
#pragma unroll
    for (int q4 = 1; q4 < 4; ++q4)
#pragma unroll
      for (int i = 0; i < 4; ++i) top4_insert(tt[0], tt[q4][i]);
    m4[0] = tt[0][0];
  }
  const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  const float mxs = mx * LOG2E;
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
  float inv = 1.f;
  if constexpr (EXPORT) {
    for (int c0 = 0; c0 < last; c0 += 4) {
      float v[4][16];
#pragma unroll
      for (int u = 0; u < 4; ++u) ptx::tmem_ld16(t_row + (c0 + u) * 16, v[u]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + u < last) {
#pragma unroll
          for (int j = 0; j < 16; ++j) s4[j & 3] += ex2_fast(fmaf(v[u][j], LOG2E, -mxs));
        }
    }
    if (n_chunks > 0) {
      float v[16];
      ptx::tmem_ld16(t_row + last * 16, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e = ex2_fast(fmaf(v[j], LOG2E, -mxs));
        s4[j & 3] += (j < nv) ? e : 0.f;
      }
    }
    inv = 1.f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
  }
  const uint32_t drow = (uint32_t)(((long long)b * a.H + h) * a.N + row);   // dropout coordinate of this row
  // EXPORT: the normalised map leaves through a per-warp staging tile [32 rows][32 columns] in shared memory, so
  // that a store instruction writes 128 contiguous bytes of ONE row (a thread owns a row: storing from registers
  // scatters every instruction over 32 rows, 8x the memory transactions)
  const int lane = threadIdx.x & 31;
  const int row0 = row - lane;                     // first row of this warp
  float* p_base = nullptr;
  if constexpr (EXPORT) {
    if (a.p_out) p_base = a.p_out + ((long long)b * a.H + h) * a.N * a.N;
  }
  for (int c0 = 0; c0 < n_chunks; c0 += 2) {
    float v[2][16];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (c0 + u < n_chunks) ptx::tmem_ld16(t_row + (c0 + u) * 16, v[u]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int c = c0 + u;
      if (c < n_chunks) {
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[u][j] = ex2_fast(fmaf(v[u][j], LOG2E, -mxs));
        if (c == last) {   // (a uniform branch: only the chunk that straddles N pays for column tests)
#pragma unroll
          for (int j = 0; j < 16; ++j) v[u][j] = (j < nv) ? v[u][j] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float e = v[u][j];
          if constexpr (EXPORT) {
            e *= inv;
            if constexpr (DROP) e *= drop_factor(a.drop, drow, c * 16 + j);   // the exported map is post-dropout
            stage[lane * 33 + u * 16 + j] = e;
          } else {
            s4[j & 3] += e;
            if constexpr (DROP) e *= drop_factor(a.drop, drow, c * 16 + j);
          }
          v[u][j] = e;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(v[u][2 * j], v[u][2 * j + 1]);
          packed[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        // P chunk c (16 keys = 8 packed columns) lands on columns this thread has already consumed
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t_row + c * 8),
                     "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3]), "r"(packed[4]), "r"(packed[5]),
                     "r"(packed[6]), "r"(packed[7])
                     : "memory");
      }
    }
    if constexpr (EXPORT) {
      __syncwarp();
      const int col = c0 * 16 + lane;
      if (p_base && col < a.N && (c0 + (lane >> 4)) < n_chunks) {
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr)
          if (row0 + rr < a.N) p_base[(long long)(row0 + rr) * a.N + col] = stage[rr * 33 + lane];
      }
      __syncwarp();
    }
  }
  const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  if constexpr (!EXPORT) inv = 1.f / sum;
  if (a.lse_out && row < a.N) a.lse_out[((long long)b * a.H + h) * a.N + row] = mxs + log2f(sum);
  if constexpr (JAS) {
    // x_(j): j-th largest probability, clamped like the reference's clamp(P, 1e-12, 1); its renormalisation by
    // (sum of the clamped row + 1e-12) differs from 1 by < N * 1e-12 and is dropped
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = fmaxf(ex2_fast(fmaf(tt[0][i], LOG2E, -mxs)) * inv, 1e-12f);
    const float g1 = x[0] * (1.f - x[0] + x[1]);
    float val;
    if (a.jas_k == 0) {
      val = logf(g1 + 1e-12f);
    } else {
      const float xk = (a.jas_k == 1) ? x[0] : (a.jas_k == 2) ? x[1] : x[2];
      const float xk1 = (a.jas_k == 1) ? x[1] : (a.jas_k == 2) ? x[2] : x[3];
      val = logf(g1 / (xk * (1.f - xk + xk1) + 1e-12f) + 1e-12f);
    }
    if (!(row < a.N) || n_chunks == 0) val = -INFINITY;
#pragma unroll
    for (int o = 16; o; o >>= 1) val = fmaxf(val, __shfl_xor_sync(0xffffffffu, val, o));
    if ((threadIdx.x & 31) == 0 && val > -INFINITY) atomic_max_float(a.jas_out + (long long)b * a.H + h, val);
  }
  return inv;
}

// Non-exporting form of attn_softmax_row -- the hot loop of both fused forwards.  Same arithmetic in the same order
// (bitwise equal results), but the instruction stream of a full 16-key chunk is exactly FFMA, MUFU.EX2, FADD per
// element plus a pack per pair: the chunk that straddles N is peeled off as a tail with its column predicates against
// compile-time column numbers, so the other chunks carry none (the generic loop spent 3 of its 7 issue slots per
// element on them, and two warps per scheduler have no latency to hide that behind).  Loads go four chunks per
// wait and may run past NP: they stay inside the slot's 256 columns and those registers are never used.
// (tools/ubench/tmem_ld.cu: tensor memory delivers ~600-900 B/clk/SM to 4-8 warps and a load round trip is ~55
// cycles -- the passes are bound by MUFU, 8 cycles per warp-wide ex2, not by the read port.)
template <bool DROP, bool JAS>
__device__ __forceinline__ float attn_softmax_row_fast(const AttnArgs& a, uint32_t t_row, int n_chunks, int b, int h, int row) {
  constexpr float LOG2E = 1.4426950408889634f;
  const int last = n_chunks - 1;            // the chunk that holds column N - 1
  const int nv = a.N - last * 16;           // its key columns, 1..16
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float tt[JAS ? 4 : 1][4];
  if constexpr (JAS) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
      for (int i = 0; i < 4; ++i) tt[q4][i] = -INFINITY;
  }
  // ---- pass 1: row max (JAS: the four largest logits) ----
  for (int c0 = 0; c0 < last; c0 += 4) {
    float v[4][16];
#pragma unroll
    for (int u = 0; u < 4; ++u) ptx::tmem_ld16(t_row + (c0 + u) * 16, v[u]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c0 + u < last) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if constexpr (JAS) top4_insert(tt[j & 3], v[u][j]);
          else m4[j & 3] = fmaxf(m4[j & 3], v[u][j]);
        }
      }
  }
  if (n_chunks > 0) {
    float v[16];
    ptx::tmem_ld16(t_row + last * 16, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x = (j < nv) ? v[j] : -INFINITY;
      if constexpr (JAS) top4_insert(tt[j & 3], x);
      else m4[j & 3] = fmaxf(m4[j & 3], x);
    }
  }
  if constexpr (JAS) {
#pragma unroll
    for (int q4 = 1; q4 < 4; ++q4)
#pragma unroll
      for (int i = 0; i < 4; ++i) top4_insert(tt[0], tt[q4][i]);
    m4[0] = tt[0][0];
  }
  const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  const float mxs = mx * LOG2E;
  // ---- pass 2: un-normalised exp, row sum, bf16 P in place (P chunk c = 8 packed columns at 8 c: columns this
  // thread has already loaded) ----
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t drow = (uint32_t)(((long long)b * a.H + h) * a.N + row);   // dropout coordinate of this row
  auto pack_store = [&](int c, float (&e)[16]) {
    uint32_t packed[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 hh = __floats2bfloat162_rn(e[2 * j], e[2 * j + 1]);
      packed[j] = *reinterpret_cast<uint32_t*>(&hh);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t_row + c * 8),
                 "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3]), "r"(packed[4]), "r"(packed[5]),
                 "r"(packed[6]), "r"(packed[7])
                 : "memory");
  };
  for (int c0 = 0; c0 < last; c0 += 4) {
    float v[4][16];
#pragma unroll
    for (int u = 0; u < 4; ++u) ptx::tmem_ld16(t_row + (c0 + u) * 16, v[u]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c0 + u < last) {
        const int c = c0 + u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float e = ex2_fast(fmaf(v[u][j], LOG2E, -mxs));
          s4[j & 3] += e;
          if constexpr (DROP) e *= drop_factor(a.drop, drow, c * 16 + j);
          v[u][j] = e;
        }
        pack_store(c, v[u]);
      }
  }
  if (n_chunks > 0) {
    float v[16];
    ptx::tmem_ld16(t_row + last * 16, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float e = ex2_fast(fmaf(v[j], LOG2E, -mxs));
      e = (j < nv) ? e : 0.f;
      s4[j & 3] += e;
      if constexpr (DROP) e *= drop_factor(a.drop, drow, last * 16 + j);
      v[j] = e;
    }
    pack_store(last, v);
  }
  const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  const float inv = 1.f / sum;
  if (a.lse_out && row < a.N) a.lse_out[((long long)b * a.H + h) * a.N + row] = mxs + log2f(sum);
  if constexpr (JAS) {
    // x_(j): j-th largest probability, clamped like the reference's clamp(P, 1e-12, 1); its renormalisation by
    // (sum of the clamped row + 1e-12) differs from 1 by < N * 1e-12 and is dropped
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = fmaxf(ex2_fast(fmaf(tt[0][i], LOG2E, -mxs)) * inv, 1e-12f);
    const float g1 = x[0] * (1.f - x[0] + x[1]);
    float val;
    if (a.jas_k == 0) {
      val = logf(g1 + 1e-12f);
    } else {
      const float xk = (a.jas_k == 1) ? x[0] : (a.jas_k == 2) ? x[1] : x[2];
      const float xk1 = (a.jas_k == 1) ? x[1] : (a.jas_k == 2) ? x[2] : x[3];
      val = logf(g1 / (xk * (1.f - xk + xk1) + 1e-12f) + 1e-12f);
    }
    if (!(row < a.N) || n_chunks == 0) val = -INFINITY;
#pragma unroll
    for (int o = 16; o; o >>= 1) val = fmaxf(val, __shfl_xor_sync(0xffffffffu, val, o));
    if ((threadIdx.x & 31) == 0 && val > -INFINITY) atomic_max_float(a.jas_out + (long long)b * a.H + h, val);
  }
  return inv;
}

// O row: tensor memory (columns t_o ..+64) -> registers, one wait for the four loads
__device__ __forceinline__ void attn_load_o_row(uint32_t t_o, float (&v)[HD]) {
#pragma unroll
  for (int c = 0; c < HD / 16; ++c) ptx::tmem_ld16(t_o + c * 16, &v[c * 16]);
  ptx::tmem_ld_wait();
}
// O row (tensor memory, columns t_o ..+64) * inv -> bf16 -> columns [h*64, h*64+64) of the [O | h] buffer
template <bool EXPORT>
__device__ __forceinline__ void attn_store_o_row(const AttnArgs& a, uint32_t t_o, int b, int h, int row, float inv) {
  __nv_bfloat16* o_ptr = reinterpret_cast<__nv_bfloat16*>(a.oh) + ((long long)b * a.N + row) * a.ld_oh + h * HD;
#pragma unroll
  for (int c = 0; c < HD / 16; ++c) {
    float v[16];
    ptx::tmem_ld16(t_o + c * 16, v);
    ptx::tmem_ld_wait();
    if (row < a.N) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s0 = EXPORT ? v[2 * j] : v[2 * j] * inv;
        const float s1 = EXPORT ? v[2 * j + 1] : v[2 * j + 1] * inv;
        __nv_bfloat162 hh = __floats2bfloat162_rn(s0, s1);
        w[j] = *reinterpret_cast<uint32_t*>(&hh);
      }
      uint4* o = reinterpret_cast<uint4*>(o_ptr + c * 16);
      o[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}

template <bool EXPORT, bool DROP, bool JAS = false>
__global__ void __launch_bounds__(128, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_bytes = a.NP * HD * 2;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + BMQ * HD * 2;
  uint8_t* sV = sK + kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kv_bytes);
  uint64_t* bar_qk = bars;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x;
  const int mt = unit % a.tiles_m;
  const int h = (unit / a.tiles_m) % a.H;
  const int b = unit / (a.tiles_m * a.H);

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmKV);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar_qk, BMQ * HD * 2 + kv_bytes);
    ptx::tma_load_3d(sQ, &tmQ, bar_qk, h * HD, mt * BMQ, b);
    ptx::tma_load_3d(sK, &tmKV, bar_qk, a.D + h * HD, 0, b);
    ptx::mbar_expect_tx(bar_v, kv_bytes);
    ptx::tma_load_3d(sV, &tmKV, bar_v, 2 * a.D + h * HD, 0, b);
    // ---- MMA1: S = Q K^T (both K-major, 128-byte swizzle) ----
    ptx::mbar_wait(bar_qk, 0);
    ptx::tc_fence_after();
    const uint32_t idesc1 = ptx::idesc_bf16(BMQ, a.NP, 0, 0);
    const uint32_t q_addr = ptx::smem_u32(sQ), k_addr = ptx::smem_u32(sK);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) {
      const uint64_t dq = ptx::smem_desc_sw128(q_addr + k * 32, 16, 1024);
      const uint64_t dk = ptx::smem_desc_sw128(k_addr + k * 32, 16, 1024);
      ptx::mma_bf16_ss(tmem, dq, dk, idesc1, k > 0 ? 1u : 0u);
    }
    ptx::mma_commit(bar_s);
  }

  // ---- softmax over this thread's row, straight from TMEM ----
  ptx::mbar_wait(bar_s, 0);
  ptx::tc_fence_after();
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  // a warp whose 32 query rows all lie past N (the last tile of an image) skips the passes: its P rows stay
  // whatever S left there and only feed O rows that are never stored
  const int n_chunks = (mt * BMQ + warp * 32 < a.N) ? a.NP / 16 : 0;
  const int row = mt * BMQ + warp * 32 + lane;
  float* stage = EXPORT ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 64) + warp * (32 * 33) : nullptr;
  float inv;
  if constexpr (EXPORT) inv = attn_softmax_row<EXPORT, DROP, JAS>(a, t_row, n_chunks, b, h, row, stage);
  else inv = attn_softmax_row_fast<DROP, JAS>(a, t_row, n_chunks, b, h, row);
  ptx::tmem_st_wait();
  ptx::tc_fence_before();
  __syncthreads();

  // ---- MMA2: O = P V   (A = P from TMEM, B = V from smem, MN-major: [key][d]) ----
  if (threadIdx.x == 0) {
    ptx::tc_fence_after();
    ptx::mbar_wait(bar_v, 0);
    ptx::tc_fence_after();
    const uint32_t idesc2 = ptx::idesc_bf16(BMQ, HD, 0, 1);
    const uint32_t v_addr = ptx::smem_u32(sV);
    for (int k = 0; k < a.NP / 16; ++k) {
      const uint64_t dv = ptx::smem_desc_sw128(v_addr + k * 2048, 8192, 1024);
      ptx::mma_bf16_ts(tmem + O_COL, tmem + k * 8, dv, idesc2, k > 0 ? 1u : 0u);
    }
    ptx::mma_commit(bar_o);
  }

  // ---- epilogue: O row / sum -> bf16 ----
  ptx::mbar_wait(bar_o, 0);
  ptx::tc_fence_after();
  attn_store_o_row<EXPORT>(a, t_row + O_COL, b, h, row, inv);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// Persistent ping-pong variant of the forward: one CTA per SM walks the (query tile, image, head) units
// c, c+G, c+2G, ...; two softmax warpgroups own one tensor-memory slot each (unit parity) and alternate, a
// loader thread streams Q | K of the next unit of each slot and a ring of three V tiles, each slot has its own MMA-issuing warp (S, then P.V of the slot's units: the two
// chains only meet in the tensor pipe), so the tensor core, the TMA engine and the two warpgroups overlap.
// What a CTA-per-unit launch serialises for every unit (launch, TMEM allocation, barrier setup, the TMA round
// trip) is paid once per SM here.
//   warps 8, 11    MMA issue for slot 0 / slot 1 (predicated on the elected lane)
//   warp 10        TMA loads of Q / K / V (lane 0)
//   warp 9         TMA stores of the O tiles: a warpgroup parks its scaled bf16 O tile [128 x 64] in its slot's
//                  staging tile, 128-byte swizzled, and goes on; this warp's lane 0 issues the tensor store (rows
//                  past N are clipped by the map) and hands the tile back once it has been read.
//                  (A thread storing its own row from registers scatters every instruction over 32 lines: that
//                  epilogue cost a warpgroup 1.6 k cycles per unit, a fifth of its time.)
//   warps 0-3      slot 0: units 0, 2, 4, ... of this CTA      (thread = query row, quadrant = warp & 3)
//   warps 4-7      slot 1: units 1, 3, 5, ...
constexpr int PP_THREADS = 384;
constexpr int PP_V_STAGES = 3;

template <bool EXPORT, bool DROP, bool JAS = false>
__global__ void __launch_bounds__(PP_THREADS, 1)
attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // shared memory: Q|K of the unit in each slot (dead once S is in tensor memory: reloaded a whole softmax ahead of
  // the slot's next unit), a ring of three V tiles (dead after P.V), one O staging tile per slot
  const int kv_bytes = a.NP * HD * 2;
  const int qk_bytes = BMQ * HD * 2 + kv_bytes;
  uint8_t* sm_qk = smem;                                  // [2][Q 128 x 64 | K NP x 64]
  uint8_t* sm_v = sm_qk + 2 * qk_bytes;                   // [PP_V_STAGES][NP x 64]
  uint8_t* sm_o = sm_v + PP_V_STAGES * kv_bytes;          // [2][128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_o + 2 * BMQ * HD * 2);
  uint64_t* bar_qk = bars;                         // [slot] Q and K landed
  uint64_t* bar_s = bars + 2;                      // [slot] S is in tensor memory (and Q | K may be overwritten)
  uint64_t* bar_p = bars + 4;                      // [slot] P is packed (128 arrivals)
  uint64_t* bar_o = bars + 6;                      // [slot] O is in tensor memory
  uint64_t* bar_epi = bars + 8;                    // [slot] the epilogue has read O (128 arrivals)
  uint64_t* bar_ost = bars + 10;                   // [slot] the O tile is staged in shared memory (128 arrivals)
  uint64_t* bar_ofree = bars + 12;                 // [slot] the staged O tile has been read by its tensor store
  uint64_t* bar_v = bars + 14;                     // [v stage] V landed
  uint64_t* bar_vfree = bar_v + PP_V_STAGES;       // [v stage] P.V of the stage's unit has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_vfree + PP_V_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef ATTN_FWD_TRACE
  const long long ft_entry = clock64();
  unsigned long long gt_entry;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_entry));
#endif
  const int items = a.B * a.H;
  const int n_units = items * a.tiles_m;
  const int n_mine = (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmKV);
    ptx::prefetch_tensormap(&tmO);
    for (int i = 0; i < PP_V_STAGES; ++i) { ptx::mbar_init(bar_v + i, 1); ptx::mbar_init(bar_vfree + i, 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_qk + i, 1); ptx::mbar_init(bar_s + i, 1); ptx::mbar_init(bar_p + i, 128); ptx::mbar_init(bar_o + i, 1);
      ptx::mbar_init(bar_epi + i, 128); ptx::mbar_init(bar_ost + i, 128); ptx::mbar_init(bar_ofree + i, 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 8) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // unit i of this CTA: all first tiles come before all second tiles, so every CTA gets both kinds
  auto decode = [&](int i, int& b, int& h, int& mt) {
    const int idx = (int)blockIdx.x + i * (int)gridDim.x;
    mt = idx / items;
    const int item = idx - mt * items;
    h = item % a.H;
    b = item / a.H;
  };

  if (warp == 8 || warp == 11) {
    // one issuing warp per slot: its units' S and P.V in their natural order, never waiting for the other warpgroup
    const int sl = warp == 8 ? 0 : 1;
    const bool leader = ptx::elect_one();
    const uint32_t lead = leader ? 1u : 0u;
    const uint32_t idesc1 = ptx::idesc_bf16(BMQ, a.NP, 0, 0);
    const uint32_t idesc2 = ptx::idesc_bf16(BMQ, HD, 0, 1);
    const uint64_t DESC0 = ptx::smem_desc_sw128(0, 16, 1024);
    const uint32_t smem0 = ptx::smem_u32(smem);
    const uint32_t t = tmem + sl * 256;
#ifdef ATTN_FWD_TRACE
    long long mt_[4] = {0, 0, 0, 0}, mt0 = 0;
#define MT0() mt0 = clock64()
#define MTA(i) mt_[i] += clock64() - mt0
#else
#define MT0() do { } while (0)
#define MTA(i) do { } while (0)
#endif
    for (int u = sl; u < n_mine; u += 2) {
      const int vs = u % PP_V_STAGES;
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      MT0();
      ptx::mbar_wait(bar_qk + sl, ph);
      MTA(0);
      MT0();
      if (u >= 2) ptx::mbar_wait(bar_epi + sl, ph ^ 1u);   // the slot's previous O has been read
      MTA(1);
      ptx::tc_fence_after();
      const uint64_t dq = DESC0 + ((smem0 + sl * qk_bytes) >> 4), dk = dq + ((BMQ * HD * 2) >> 4);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) ptx::mma_ss_pred(t, dq + 2 * k, dk + 2 * k, idesc1, k > 0 ? 1u : 0u, lead);
      ptx::commit_pred(ptx::smem_u32(bar_s + sl), lead);
      MT0();
      ptx::mbar_wait(bar_v + vs, (uint32_t)((u / PP_V_STAGES) & 1));
      MTA(2);
      MT0();
      ptx::mbar_wait(bar_p + sl, ph);
      MTA(3);
      ptx::tc_fence_after();
      const uint32_t v_addr = ptx::smem_u32(sm_v) + vs * kv_bytes;
      for (int k = 0; k < a.NP / 16; ++k)
        ptx::mma_ts_pred(t + O_COL, t + k * 8, ptx::smem_desc_sw128(v_addr + k * 2048, 8192, 1024), idesc2, k > 0 ? 1u : 0u, lead);
      ptx::commit_pred(ptx::smem_u32(bar_o + sl), lead);
      ptx::commit_pred(ptx::smem_u32(bar_vfree + vs), lead);
    }
#ifdef ATTN_FWD_TRACE
    if (blockIdx.x == 0 && lane == 0)
      printf("FT issuer %d | wait qk %lld | wait epi %lld | wait v %lld | wait p %lld\n", sl, mt_[0], mt_[1], mt_[2], mt_[3]);
#endif
  } else if (warp == 10) {
    // loads: Q | K of unit i as soon as S of the slot's previous unit is in tensor memory, V of unit i once P.V of
    // unit i - 3 has retired.  (Their own thread: waiting for a buffer inside an MMA warp kept the next S from
    // being issued; freeing Q | K | V together, after P.V, left a load ~1.5 us to land and S waited for it.)
    if (lane == 0) {
      for (int i = 0; i < n_mine; ++i) {
        const int sl = i & 1, vs = i % PP_V_STAGES;
        int b, h, mt;
        decode(i, b, h, mt);
        if (i >= 2) ptx::mbar_wait(bar_s + sl, (uint32_t)(((i >> 1) - 1) & 1));
        uint8_t* q = sm_qk + sl * qk_bytes;
        ptx::mbar_expect_tx(bar_qk + sl, qk_bytes);
        ptx::tma_load_3d(q, &tmQ, bar_qk + sl, h * HD, mt * BMQ, b);
        ptx::tma_load_3d(q + BMQ * HD * 2, &tmKV, bar_qk + sl, a.D + h * HD, 0, b);
        if (i >= PP_V_STAGES) ptx::mbar_wait(bar_vfree + vs, (uint32_t)((i / PP_V_STAGES - 1) & 1));
        ptx::mbar_expect_tx(bar_v + vs, kv_bytes);
        ptx::tma_load_3d(sm_v + vs * kv_bytes, &tmKV, bar_v + vs, 2 * a.D + h * HD, 0, b);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      for (int u = 0; u < n_mine; ++u) {
        int b, h, mt;
        decode(u, b, h, mt);
        ptx::mbar_wait(bar_ost + (u & 1), (uint32_t)((u >> 1) & 1));
        ptx::tma_store_3d(&tmO, sm_o + (u & 1) * (BMQ * HD * 2), h * HD, mt * BMQ, b);
        ptx::bulk_commit_group();
        ptx::bulk_wait_group_read0();
        ptx::mbar_arrive(bar_ofree + (u & 1));
      }
      ptx::bulk_wait_group0();
    }
  } else {
    const int wg = warp >> 2, quarter = warp & 3;
    const uint32_t t_row = tmem + wg * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
#ifdef ATTN_FWD_TRACE   // make trace TRACE=-DATTN_FWD_TRACE: where a softmax warp's cycles go
    long long ft[4] = {0, 0, 0, 0}, ft0 = 0, ft_begin = clock64();
#define FT0() ft0 = clock64()
#define FTA(i) ft[i] += clock64() - ft0
#else
#define FT0() do { } while (0)
#define FTA(i) do { } while (0)
#endif
    for (int u = wg; u < n_mine; u += 2) {
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      int b, h, mt;
      decode(u, b, h, mt);
      const int row = mt * BMQ + quarter * 32 + lane;
      const int n_chunks = (mt * BMQ + quarter * 32 < a.N) ? a.NP / 16 : 0;
      FT0();
      ptx::mbar_wait(bar_s + wg, ph);
      FTA(0);
      ptx::tc_fence_after();
      FT0();
      static_assert(!EXPORT, "the ping-pong forward does not export the map");
      const float inv = attn_softmax_row_fast<DROP, JAS>(a, t_row, n_chunks, b, h, row);
      ptx::tmem_st_wait();
      FTA(1);
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_p + wg);
      FT0();
      ptx::mbar_wait(bar_o + wg, ph);
      FTA(2);
      ptx::tc_fence_after();
      FT0();
      // O is in registers after one round trip: the slot goes back to the MMA warp (S of this warpgroup's next unit)
      // BEFORE the scaled row is converted and stored
      float o[HD];
      attn_load_o_row(t_row + O_COL, o);
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_epi + wg);
      {   // scaled bf16 row -> the slot's staging tile, swizzled as the tensor map expects: 16-byte chunk c of tile
          // row r sits at chunk c ^ (r & 7)
        const int r = quarter * 32 + lane;
        uint8_t* dst = sm_o + wg * (BMQ * HD * 2) + r * 128;
        if (u >= 2) ptx::mbar_wait(bar_ofree + wg, ph ^ 1u);   // the tile of this slot's previous unit has left
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 hh = __floats2bfloat162_rn(o[c * 8 + 2 * j] * inv, o[c * 8 + 2 * j + 1] * inv);
            w[j] = *reinterpret_cast<uint32_t*>(&hh);
          }
          *reinterpret_cast<uint4*>(dst + ((c ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        ptx::fence_async_shared();
        ptx::mbar_arrive(bar_ost + wg);
      }
      FTA(3);
    }
#ifdef ATTN_FWD_TRACE
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 2 || warp == 4))
      printf("FT warp %d units %d total %lld | wait S %lld | softmax %lld | wait O %lld | store O %lld\n", warp, (n_mine - wg + 1) / 2,
             clock64() - ft_begin, ft[0], ft[1], ft[2], ft[3]);
#endif
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
#ifdef ATTN_FWD_TRACE
  if ((blockIdx.x == 0 || blockIdx.x == 147) && threadIdx.x == 0) {
    unsigned long long gt_exit;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_exit));
    printf("FT block %d entry->exit %lld cycles, %llu ns (entry at %llu)\n", blockIdx.x, clock64() - ft_entry, gt_exit - gt_entry, gt_entry);
  }
#endif
}

// ================================================================================================
// Fused attention VJP (bf16 mode, head dim 64, N <= 256).  Per (image b, head h), with the
// log-sum-exp `lse2` of every row saved by the forward and delta_i = sum_d dO_id O_id:
//     S = q k^T,  P = exp2(S*log2e - lse2),  dP = dO v^T,  dS = P o (dP - delta)
//     dq = dS k,  dk = dS^T q,  dv = P^T dO
// PERSISTENT kernel: one CTA per SM walks the (b, h) items c, c+G, c+2G, ...; everything is in the
// TRANSPOSED orientation (thread = key row), so that P^T and dS^T are A operands that never leave
// tensor memory.  An iteration is one (128-key chunk kc, 128-query tile qt) pair, its S^T / dP^T tiles split into
// two HALVES of 64 query columns with their own "ready" / "done" barriers:
//   warp 16             every MMA (all lanes walk the schedule; tcgen05 instructions predicated on the elected lane).
//        per iteration, for half hf = 0, 1:
//          A(hf):  S^T[hf] = K_kc Q_(qt,hf)^T,  dP^T[hf] = V_kc dO_(qt,hf)^T       (SS)   128 x 64 fp32 each
//          C(hf):  dV_kc += P^T[hf] dO_(qt,hf),  dK_kc += dS^T[hf] Q_(qt,hf)        (TS: A = bf16 in TMEM)
//        and once both halves are done   dQ_qt += dS K_kc   (A = the dS^T tile in shared memory, MN-major).
//        Issue order  C(i,0) A(i+1,0) | C(i,1) A(i+1,1) dQ(i)  (at a key chunk's last query tile: C, C, dQ, A, A -- the
//        compute warps go to the accumulator read-out first).  dK, dV, dQ_0, dQ_1 accumulate in TMEM [256,512).
//   warp 17, lane 0     TMA loads: the operand tiles {K, Q, V, dO} x 128 rows live in a ring of three 64 KB sets over the
//        sequence of half-item loads (full / free barriers; "free" is a tcgen05.commit of the MMA warp).
//   warp 17, lane 1     TMA stores: dq | dk | dv leave as [128 x 64] tiles staged in the dS^T tile (see acc_stage32).
//        (Issuing a 16 KB TMA box costs its thread several hundred cycles: neither the MMA warp nor a compute warp.)
//   warps 0-15          four warpgroups split the query columns of a tile (32 each: warpgroups 0, 1 = half 0); thread =
//          key row: P^T, dS^T of its 128 x 32 block from S^T, dP^T (per-query lse / delta broadcast from shared memory),
//          packed to bf16 and written back IN PLACE over columns the same thread has already read (warpgroup w parks its
//          packed columns at [32 w, 32 w + 16) of the tile: no cross-thread hazard, no barrier); dS^T also goes to the
//          128-byte-swizzled shared-memory tile.  Key rows past N need no predicate: their K / V rows are TMA zero
//          fill, so whatever they produce multiplies zeros or lands in accumulator rows the stores clip.  Read-outs:
//          dK | dV per key chunk and dQ at the item's end, 32 x 32 per warp, signalled to the MMA warp as soon as the
//          values are in registers.
//   warps 18-19         lse and delta = <dO_i, O_i> of the NEXT item into shared memory: the only global
//          loads that are not TMA, kept off the compute warps.
// Measured (B200, 64 x 12 items of 207 tokens): 95 us (round 1: 8 compute warps, one barrier per phase, TMA and MMA
// issued by one thread, register -> global epilogue) -> 72 us.  What bounds it now (clock trace, -DATTN_TRACE): per
// iteration ~1.2 k cycles of P^T / dS^T formation (issue- and MUFU-bound with all 16 warps busy), ~1.3-1.9 k cycles of
// MMAs (64-column MMAs cost ~48 cycles, 128-column ones ~68: tools/ubench/mma_rate.cu) and ~1 k cycles of
// hand-offs, only partly overlapped because tensor memory holds a single S^T / dP^T tile next to the accumulators.
struct AttnBwdArgs {
  int B, N, H, D, R;
  int n_t;                      // 128-row tiles per item: 1 or 2 (keys and queries alike)
  int items;                    // B * H
  const float* lse2;            // [B,H,N]
  const __nv_bfloat16* dO;      // [B*N, D]
  const __nv_bfloat16* O;       // [B*N, ld_o]
  long long ld_o;
  void* dz;                     // [B*N, R] bf16: dq | dk | dv at columns h*64, D + h*64, 2D + h*64
  Drop drop;                    // the forward's attention-map dropout, regenerated here
  // cotangent of the EXPORTED map (the `attentions` output of the last evaluation): dS = P o (dP + gp - delta),
  // delta = <dO, O> + dext with dext[b,h,i] = sum_j P_ij gp_ij formed beforehand (rows.cu::rowdot_rows)
  const float* gp;              // [B,H,N,N] fp32 or null
  const float* dext;            // [B,H,N] fp32 or null
  // bias gradient of the q rows of the in-projection: column sums of dq over every token, accumulated (atomics)
  // into dq_colsum[h*64 + c] when not null -- the rows are in registers at the item's end anyway
  float* dq_colsum;
};

constexpr int BWD_TMEM_COLS = 512;
constexpr int T_S0 = 0, T_DP0 = 128, T_DK = 256, T_DV = 320, T_DQ = 384;  // dQ_qt at T_DQ + 64 qt
constexpr int BWD_COMPUTE_WARPS = 16;
constexpr int BWD_MMA_WARP = 16;
constexpr int BWD_TMA_WARP = 17;
constexpr int BWD_LOADER_WARP = 18;
constexpr int BWD_LOADER_THREADS = 64;   // 2 warps: 20 warps in all leave 96 registers per thread
constexpr int BWD_THREADS = (BWD_LOADER_WARP * 32) + BWD_LOADER_THREADS;
constexpr int SLOT = 16384;               // one [128 x 64] bf16 tile
constexpr int SET = 4 * SLOT;             // {K, Q, V, dO}
enum { SL_K = 0, SL_Q = 1, SL_V = 2, SL_DO = 3 };

__device__ __forceinline__ void store_bf16x8_sw128(uint8_t* tile, int r, int col, const uint32_t* w) {
  // tile: atoms [128 rows x 64 columns] of 16 KB; 16-byte chunk index XOR (row % 8)
  const int atom = col >> 6, chunk = (col & 63) >> 3;
  uint4* dst = reinterpret_cast<uint4*>(tile + atom * 16384 + r * 128 + ((chunk ^ (r & 7)) << 4));
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// Accumulator read-out: this warp's 32 rows x 32 fp32 columns -> 16 packed bf16 pairs per thread (thread = row).
__device__ __forceinline__ void acc_load_pack32(uint32_t t_addr, uint32_t (&w)[16]) {
  float v[32];
  ptx::tmem_ld32(t_addr, v);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&hh);
  }
}
// The same read-out, plus the sums over this warp's 32 rows of each of the 32 columns: a butterfly over the lanes in
// which every step halves the number of columns a lane carries (16 + 8 + 4 + 2 + 1 shuffles); lane j ends up with the
// sum of column j.
__device__ __forceinline__ float acc_load_pack32_colsum(uint32_t t_addr, uint32_t (&w)[16], int lane, bool row_live) {
  float v[32];
  ptx::tmem_ld32(t_addr, v);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&hh);
  }
  // rows past N are not zeros (their dS columns are whatever the shared tile held: the tensor store clips them)
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = row_live ? v[i] : 0.f;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = up ? v[i + o] : v[i];
      const float send = up ? v[i] : v[i + o];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
// ... -> the 16-byte slots of the dS^T tile that THIS thread writes in the main loop (row `trow`, atom `hf`, chunks
// sh*4 .. sh*4+3 under the 128-byte swizzle).  With warpgroups (0, 1) holding the column halves of one accumulator and
// (2, 3) of the other, atom 0 and atom 1 become two [128 rows x 64 columns] tiles in exactly the layout a
// SWIZZLE_128B tensor map describes: one TMA store per tile takes them to global memory (rows past N are clipped by
// the map), asynchronously -- the threads do not wait for the store traffic, which arrives in bursts because the
// persistent CTAs run in phase.
__device__ __forceinline__ void acc_stage32(const uint32_t (&w)[16], uint8_t* atom, int trow, int sh) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(atom + trow * 128 + (((sh * 4 + c) ^ (trow & 7)) << 4)) =
        make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

#ifdef ATTN_TRACE
__device__ uint32_t tr[3][256];
__device__ uint8_t tr_id[3][256];
__device__ int tr_n[3];
// (the event counter lives in a register of the one tracing thread per role: a trace point is two plain stores)
#define TR(role, id) do { if (tr_on && ((role) < 20) && tr_i < 256) { tr[role][tr_i] = (uint32_t)clock(); tr_id[role][tr_i] = (id); ++tr_i; } } while (0)
#else
#define TR(role, id) do { } while (0)
#endif

template <bool DROP, bool GP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ AttnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sSets = smem;                       // 3 sets x 4 slots x 16 KB
  uint8_t* sDS = sSets + 3 * SET;              // 32 KB: dS^T tile [128 keys][128 queries] (two 64-query atoms)
  float* sLse = reinterpret_cast<float*>(sDS + 32768);   // [256]
  float* sDelta = sLse + 256;                             // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* bar_full = bars;      // [3] operand set landed
  uint64_t* bar_sdp = bars + 3;   // [2] S^T / dP^T of a half (64 query columns) are in TMEM
  uint64_t* bar_pds = bars + 5;   // [2] P^T / dS^T of a half written (TMEM + shared memory), 256 arrivals
  uint64_t* bar_c = bars + 7;     // every MMA of an iteration (dV, dK, dQ) has retired
  uint64_t* bar_epi = bars + 8;   // accumulators of a key chunk (and dQ at the item's end) read out
  uint64_t* bar_aux = bars + 9;   // lse / delta of an item are in shared memory
  uint64_t* bar_item = bars + 10; // the compute warps no longer read the item's lse / delta
  uint64_t* bar_kv = bars + 11;   // dV / dK of a key chunk are complete (its last query tile's C MMAs have retired)
  uint64_t* bar_free = bars + 12; // [3] every MMA that reads an operand set has retired
  uint64_t* bar_staged = bars + 15;      // accumulator tiles are staged in the dS^T tile (512 arrivals)
  uint64_t* bar_stage_free = bars + 16;  // their TMA stores have read the tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef ATTN_TRACE
  const bool tr_on = (blockIdx.x == 0) && (lane == 0) && (warp == 0 || warp == 8 || warp == BWD_MMA_WARP);
  int tr_i = 0;
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { ptx::mbar_init(bar_full + i, 1); ptx::mbar_init(bar_free + i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_sdp + i, 1); ptx::mbar_init(bar_pds + i, BWD_COMPUTE_WARPS * 32); }
    ptx::mbar_init(bar_c, 1);
    ptx::mbar_init(bar_kv, 1);
    ptx::mbar_init(bar_staged, BWD_COMPUTE_WARPS * 32);
    ptx::mbar_init(bar_stage_free, 1);
    ptx::mbar_init(bar_epi, BWD_COMPUTE_WARPS * 32);
    ptx::mbar_init(bar_aux, BWD_LOADER_THREADS);
    ptx::mbar_init(bar_item, BWD_COMPUTE_WARPS * 32);
    ptx::fence_barrier_init();
  }
  if (warp == BWD_MMA_WARP) ptx::tmem_alloc(tmem_slot, BWD_TMEM_COLS);
  if (warp < BWD_COMPUTE_WARPS) {
    // the dS^T tile feeds dQ with rows / columns the compute warps never write (keys, queries past N): those meet
    // zero operand rows, but must be finite
    uint4* z = reinterpret_cast<uint4*>(sDS) + threadIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_async_shared();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int NP = (a.N + 15) / 16 * 16;
  const int nt = a.n_t;
  const int n_iter = nt * nt;
  const int n_mine = (a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == BWD_TMA_WARP) {
    // =========================== operand loader: TMA only ===========================
    // The operand sets form a ring of three over the sequence of half-item loads j = nt * item index + t (rows
    // [128 t, 128 t + 128) of the item): load j goes to set j % 3 once the MMAs that read load j - 3 have retired
    // (bar_free, committed by the MMA warp).  (Issuing a 64 KB set costs the issuing thread ~1-2 k cycles: kept off the
    // warp that feeds the tensor core.)
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmQKV);
      ptx::prefetch_tensormap(&tmDO);
      const int n_loads = n_mine * nt;
      for (int j = 0; j < n_loads; ++j) {
        const int set = j % 3, idx = j / nt, t = j - idx * nt;
        const int item = (int)blockIdx.x + idx * (int)gridDim.x;
        const int hh = item % a.H, bb = item / a.H;
        if (j >= 3) ptx::mbar_wait(bar_free + set, (uint32_t)((j / 3 - 1) & 1));
        uint8_t* base = sSets + set * SET;
        ptx::mbar_expect_tx(bar_full + set, SET);
        ptx::tma_load_3d(base + SL_K * SLOT, &tmQKV, bar_full + set, a.D + hh * HD, t * 128, bb);
        ptx::tma_load_3d(base + SL_Q * SLOT, &tmQKV, bar_full + set, hh * HD, t * 128, bb);
        ptx::tma_load_3d(base + SL_V * SLOT, &tmQKV, bar_full + set, 2 * a.D + hh * HD, t * 128, bb);
        ptx::tma_load_3d(base + SL_DO * SLOT, &tmDO, bar_full + set, hh * HD, t * 128, bb);
      }
    } else if (lane == 1) {
      // ---- accumulator tiles out: per item, one group per key chunk ([dK | dV] staged in the two atoms of the dS^T
      //      tile) and a last one with [dQ_0 | dQ_1].  (A second thread of this warp: issuing a TMA store costs its
      //      thread several hundred cycles per 16 KB tile, which none of the compute warps can spare.) ----
      ptx::prefetch_tensormap(&tmDZ);
      uint32_t n_grp = 0;
      for (int idx = 0; idx < n_mine; ++idx) {
        const int item = (int)blockIdx.x + idx * (int)gridDim.x;
        const int hh = item % a.H, bb = item / a.H;
        for (int grp = 0; grp <= nt; ++grp, ++n_grp) {
          ptx::mbar_wait(bar_staged, n_grp & 1);
          if (grp < nt) {
            ptx::tma_store_3d(&tmDZ, sDS, a.D + hh * HD, grp * 128, bb);               // dK tile = atom 0
            ptx::tma_store_3d(&tmDZ, sDS + 16384, 2 * a.D + hh * HD, grp * 128, bb);   // dV tile = atom 1
          } else {
            ptx::tma_store_3d(&tmDZ, sDS, hh * HD, 0, bb);
            if (nt > 1) ptx::tma_store_3d(&tmDZ, sDS + 16384, hh * HD, 128, bb);
          }
          ptx::bulk_commit_group();
          ptx::bulk_wait_group_read0();
          ptx::mbar_arrive(bar_stage_free);
        }
      }
      ptx::bulk_wait_group0();   // the last tiles have reached global memory
    }
  } else if (warp == BWD_MMA_WARP) {
    // =========================== producer: every MMA ===========================
    // The whole warp walks the schedule (uniform control flow keeps descriptors in uniform registers);
    // only the elected lane issues MMA / commit.  The schedule is flattened over (item, kc, qt).
    const bool leader = ptx::elect_one();
    uint32_t n_epi = 0;      // accumulator read-outs the producer has waited for
    uint32_t n_epi_due = 0;  // read-outs the compute warps will have signalled before the next first-of-chunk C
    struct Iter {
      uint64_t dk_k, dq_k, dv_k, ddo_k;   // K-major descriptors (S^T, dP^T)
      uint64_t ddo_mn, dq_mn, dk_mn;      // MN-major descriptors (dV, dK, dQ)
      uint32_t id_s[2], acc_q, acc_k;
      int w[2];                           // query columns of each half (multiples of 16, 0 = empty half)
      int nq, nk, qt, it;                 // 16-query / 16-key steps of the tile / chunk
    };
    // describes iteration `it` of item index `idx` (rows [0,128) in set (nt idx) % 3, rows [128,256) in the next one of
    // the ring); waits for the operand sets the iteration touches first
    auto make_iter = [&](int idx, int it) {
      Iter r;
      const int kc = (nt == 2) ? (it >> 1) : 0, qt = (nt == 2) ? (it & 1) : 0;
      const int j0 = nt * idx, j1 = j0 + 1;
      const int s0 = j0 % 3, s1 = j1 % 3;
      if (it == 0) ptx::mbar_wait(bar_full + s0, (uint32_t)((j0 / 3) & 1));
      else if (it == 1) ptx::mbar_wait(bar_full + s1, (uint32_t)((j1 / 3) & 1));
      const int qw = min(128, NP - qt * 128), cw = min(128, NP - kc * 128);  // multiples of 16
      const uint32_t kset = ptx::smem_u32(sSets + (kc ? s1 : s0) * SET), qset = ptx::smem_u32(sSets + (qt ? s1 : s0) * SET);
      const uint32_t k_addr = kset + SL_K * SLOT, v_addr = kset + SL_V * SLOT;
      const uint32_t q_addr = qset + SL_Q * SLOT, do_addr = qset + SL_DO * SLOT;
      r.dk_k = ptx::smem_desc_sw128(k_addr, 16, 1024);
      r.dq_k = ptx::smem_desc_sw128(q_addr, 16, 1024);
      r.dv_k = ptx::smem_desc_sw128(v_addr, 16, 1024);
      r.ddo_k = ptx::smem_desc_sw128(do_addr, 16, 1024);
      r.ddo_mn = ptx::smem_desc_sw128(do_addr, 8192, 1024);
      r.dq_mn = ptx::smem_desc_sw128(q_addr, 8192, 1024);
      r.dk_mn = ptx::smem_desc_sw128(k_addr, 8192, 1024);
      r.w[0] = min(64, qw); r.w[1] = max(0, qw - 64);
      r.id_s[0] = ptx::idesc_bf16(128, r.w[0], 0, 0);
      r.id_s[1] = ptx::idesc_bf16(128, max(16, r.w[1]), 0, 0);
      r.nq = qw / 16; r.nk = cw / 16;
      r.acc_q = qt > 0 ? 1u : 0u; r.acc_k = kc > 0 ? 1u : 0u;
      r.qt = qt; r.it = it;
      return r;
    };
    // Every tcgen05 instruction below is PREDICATED on the elected lane instead of sitting in an `if (leader)` branch:
    // inside a divergent region the compiler has to form each descriptor in vector registers and move it to uniform
    // ones (R2UR) right in front of every MMA -- ~13 dependent instructions per MMA pair, more than the MMA itself
    // takes -- while a converged warp keeps the whole schedule in the uniform datapath.
    const uint32_t lead = leader ? 1u : 0u;
    const uint32_t a_sdp = ptx::smem_u32(bar_sdp), a_c = ptx::smem_u32(bar_c), a_kv = ptx::smem_u32(bar_kv);
    // S^T[hf] = K Q_hf^T, dP^T[hf] = V dO_hf^T for the 64 query columns of half hf (query rows hf*64.. of the Q / dO
    // tiles: +8192 bytes = 512 descriptor units).  A 64-wide MMA costs ~48 cycles against ~68 for 128 columns
    // (tools/ubench/mma_rate.cu), but the halves give the compute warps two independent buffers to alternate on.
    auto issue_a = [&](const Iter& r, int hf) {
      if (r.w[hf] > 0) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::mma_ss_pred(tmem + T_S0 + hf * 64, r.dk_k + 2 * k, r.dq_k + 512 * hf + 2 * k, r.id_s[hf], k > 0 ? 1u : 0u, lead);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::mma_ss_pred(tmem + T_DP0 + hf * 64, r.dv_k + 2 * k, r.ddo_k + 512 * hf + 2 * k, r.id_s[hf], k > 0 ? 1u : 0u, lead);
      }
      ptx::commit_pred(a_sdp + 8 * hf, lead);
    };
    const uint32_t id_ts = ptx::idesc_bf16(128, HD, 0, 1);   // A from TMEM, B MN-major
    const uint32_t id_dq = ptx::idesc_bf16(128, HD, 1, 1);   // A, B MN-major
    const uint64_t dds_mn = ptx::smem_desc_sw128(ptx::smem_u32(sDS), 16384, 1024);
    // dV += P^T dO, dK += dS^T Q over the half's queries, 16 per MMA; the packed columns of k-step ks (written by
    // warpgroup ks & 3) sit at [16 ks, 16 ks + 8) of the tile
    auto issue_c = [&](const Iter& r, int hf) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int ks = hf * 4 + k4;
        if (ks < r.nq) {
          const uint32_t acc = ks > 0 ? 1u : r.acc_q;
          const uint32_t a_col = 16 * ks;
          ptx::mma_ts_pred(tmem + T_DV, tmem + T_S0 + a_col, r.ddo_mn + 128 * ks, id_ts, acc, lead);
          ptx::mma_ts_pred(tmem + T_DK, tmem + T_DP0 + a_col, r.dq_mn + 128 * ks, id_ts, acc, lead);
        }
      }
    };
    auto issue_dq = [&](const Iter& r) {   // contraction over the chunk's keys
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        if (ks < r.nk)
          ptx::mma_ss_pred(tmem + T_DQ + r.qt * 64, dds_mn + 128 * ks, r.dk_mn + 128 * ks, id_dq, ks > 0 ? 1u : r.acc_k, lead);
      ptx::commit_pred(a_c, lead);
    };
    const int total = n_mine * n_iter;
    if (n_mine > 0) {
      Iter cur = make_iter(0, 0);
      issue_a(cur, 0);
      issue_a(cur, 1);
      int idx = 0;
      for (int gi = 0; gi < total; ++gi) {
        // ---- the next iteration's descriptors, before the waits ----
        const bool has_next = gi + 1 < total;
        const bool chunk_end = (cur.qt == nt - 1);
        const int cur_it = cur.it, cur_idx = idx;
        Iter nxt = cur;
        if (has_next) {
          if (cur.it == n_iter - 1) { ++idx; nxt = make_iter(idx, 0); }
          else nxt = make_iter(idx, cur.it + 1);
        }
        // ---- half 0: P^T / dS^T ready -> dV, dK partial sums; S^T / dP^T of the next iteration's half 0 ----
        TR(2, 1);
        ptx::mbar_wait(bar_pds + 0, gi & 1);
        ptx::tc_fence_after();
        TR(2, 3);
        if (cur.qt == 0 && n_epi < n_epi_due) {  // the accumulators of the previous chunk have been read out
          ptx::mbar_wait(bar_epi, n_epi & 1);
          ++n_epi;
          ptx::tc_fence_after();
        }
        issue_c(cur, 0);
        // (at a key chunk's last query tile every compute warp goes to the read-out first: dK / dV and dQ are finished
        //  before the next S^T / dP^T are queued)
        if (has_next && !chunk_end) issue_a(nxt, 0);
        TR(2, 4);
        // ---- half 1 ----
        ptx::mbar_wait(bar_pds + 1, gi & 1);
        ptx::tc_fence_after();
        issue_c(cur, 1);
        if (chunk_end) {
          ptx::commit_pred(a_kv, lead);
          ++n_epi_due;  // the compute warps read this chunk's accumulators out next
          issue_dq(cur);
          if (has_next) { issue_a(nxt, 0); issue_a(nxt, 1); }
        } else {
          if (has_next) issue_a(nxt, 1);
          issue_dq(cur);
        }
        TR(2, 6);
        // ---- operand sets this iteration was the last to read go back to the TMA warp ----
        if (nt == 1) {
          ptx::commit_pred(ptx::smem_u32(bar_free + cur_idx % 3), lead);
        } else if (cur_it >= 2) {
          ptx::commit_pred(ptx::smem_u32(bar_free + (2 * cur_idx + cur_it - 2) % 3), lead);   // it 2: rows [0,128); it 3: rows [128,256)
        }
        TR(2, 5);
        cur = nxt;
      }
    }
  } else if (warp >= BWD_LOADER_WARP) {
    // =========================== lse / delta loader (64 threads) ===========================
    // The values of item idx are formed in registers while item idx-1 is being computed and dropped
    // into the (single) shared-memory buffer the moment that item is finished.
    // (Item 0 is the exception: nothing hides these 64 threads' four dependent rounds of row loads at the start of the
    //  kernel -- ~13 k cycles before the first exponential -- so the 512 idle compute threads form it, one round.)
    const int tl = threadIdx.x - BWD_LOADER_WARP * 32;
    if (n_mine > 0) ptx::mbar_arrive(bar_aux);
    for (int idx = 1; idx < n_mine; ++idx) {
      const int item = (int)blockIdx.x + idx * (int)gridDim.x;
      const int hh = item % a.H, bb = item / a.H;
      float lse[4], dl[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int q = tl + r * BWD_LOADER_THREADS;
        lse[r] = INFINITY;
        dl[r] = 0.f;
        if (q < a.N) {
          lse[r] = a.lse2[(long long)item * a.N + q];
          const uint4* pd = reinterpret_cast<const uint4*>(a.dO + ((long long)bb * a.N + q) * a.D + hh * HD);
          const uint4* po = reinterpret_cast<const uint4*>(a.O + ((long long)bb * a.N + q) * a.ld_o + hh * HD);
          uint4 x[8], y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { x[i] = pd[i]; y[i] = po[i]; }
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t xs[4] = {x[i].x, x[i].y, x[i].z, x[i].w}, ys[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162*>(&xs[j]);
              const __nv_bfloat162 yb = *reinterpret_cast<const __nv_bfloat162*>(&ys[j]);
              acc = fmaf(__low2float(xb), __low2float(yb), acc);
              acc = fmaf(__high2float(xb), __high2float(yb), acc);
            }
          }
          dl[r] = acc;
          if constexpr (GP) dl[r] += a.dext[(long long)item * a.N + q];
        }
      }
      TR(20, 20);
      ptx::mbar_wait(bar_item, (idx - 1) & 1);  // the previous item no longer reads the buffer
      TR(20, 21);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (tl + r * BWD_LOADER_THREADS < 256) {
          sLse[tl + r * BWD_LOADER_THREADS] = lse[r];
          sDelta[tl + r * BWD_LOADER_THREADS] = dl[r];
        }
      }
      ptx::mbar_arrive(bar_aux);
      TR(20, 22);
    }
  } else {
    // =========================== compute warps (thread = key row) ===========================
    // Every warpgroup works on BOTH halves of an iteration, one after the other: in half hf it owns the 16 query
    // columns [64 hf + 16 wg, +16).  The halves are then separated in TIME -- while the sixteen warps exponentiate
    // half 1 the tensor core runs dV / dK of half 0 and S^T / dP^T of the next iteration's half 0 -- instead of two
    // groups of eight warps working on both halves at once with the tensor core idle, then waiting together.
    const int wg = warp >> 2, quarter = warp & 3;
    [[maybe_unused]] const int trr = warp == 0 ? 0 : 1;   // trace role (ATTN_TRACE builds: warps 0 and 8 record)
    const int eh = wg >> 1, sh = wg & 1;   // read-outs: accumulator / atom of the dS^T tile and 32-column half of it
    const int trow = quarter * 32 + lane;  // row inside a 128-row tile / chunk
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t g = 0;  // global iteration counter (phases of bar_sdp / bar_pds / bar_c)
    uint32_t n_chunk = 0;   // key chunks finished (phase of bar_kv)
    bool store_pending = false;
    uint32_t n_grp = 0;     // store groups whose completion has been waited for (phase of bar_stage_free)
    if (n_mine > 0) {
      // lse / delta of the CTA's first item: two threads per query row, 32 elements of <dO, O> each
      const int item = (int)blockIdx.x, hh = item % a.H, bb = item / a.H;
      const int q = threadIdx.x >> 1, part = threadIdx.x & 1;
      float acc = 0.f;
      if (q < a.N) {
        const uint4* pd = reinterpret_cast<const uint4*>(a.dO + ((long long)bb * a.N + q) * a.D + hh * HD) + part * 4;
        const uint4* po = reinterpret_cast<const uint4*>(a.O + ((long long)bb * a.N + q) * a.ld_o + hh * HD) + part * 4;
        uint4 x[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { x[i] = pd[i]; y[i] = po[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t xs[4] = {x[i].x, x[i].y, x[i].z, x[i].w}, ys[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162*>(&xs[j]);
            const __nv_bfloat162 yb = *reinterpret_cast<const __nv_bfloat162*>(&ys[j]);
            acc = fmaf(__low2float(xb), __low2float(yb), acc);
            acc = fmaf(__high2float(xb), __high2float(yb), acc);
          }
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (part == 0) {
        float l = INFINITY, d = 0.f;
        if (q < a.N) {
          l = a.lse2[(long long)item * a.N + q];
          d = acc;
          if constexpr (GP) d += a.dext[(long long)item * a.N + q];
        }
        sLse[q] = l;
        sDelta[q] = d;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(BWD_COMPUTE_WARPS * 32) : "memory");
    }
    for (int idx = 0; idx < n_mine; ++idx) {
      const int item = (int)blockIdx.x + idx * (int)gridDim.x;
      const uint32_t drow0 = (uint32_t)((long long)item * a.N);   // dropout row coordinate of query 0
      TR(trr, 14);
      ptx::mbar_wait(bar_aux, idx & 1);
      TR(trr, 15);
      for (int it = 0; it < n_iter; ++it, ++g) {
        const int kc = it / nt, qt = it - kc * nt;
        const int qw = min(128, NP - qt * 128);
        const bool rows_live = kc * 128 + quarter * 32 < a.N;   // a warp whose 32 key rows all lie past N has nothing
                                                                // to do: its S^T / dP^T rows are exact zeros
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int col0 = hf * 64 + wg * 16;                   // first query column (inside the tile) of this chunk
          const bool live = rows_live && col0 < qw;
          const uint32_t t_s = t_lane + T_S0 + col0, t_dp = t_lane + T_DP0 + col0;
          TR(trr, 0 + 32 * hf);
          ptx::mbar_wait(bar_sdp + hf, g & 1);
          ptx::tc_fence_after();
          TR(trr, 2 + 32 * hf);
          // load, form P^T / dS^T, pack, park both in tensor memory over columns this very thread has consumed
          uint32_t pDS[8];
          if (live) {
            float sv[16], dp[16];
            ptx::tmem_ld16(t_s, sv);
            ptx::tmem_ld16(t_dp, dp);
            if constexpr (GP) {
              // + the cotangent of the exported map, read transposed: lanes = consecutive keys of query row q
              const int key = kc * 128 + trow;
              float gq[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int q = qt * 128 + col0 + j;
                gq[j] = (key < a.N && q < a.N) ? __ldg(a.gp + ((long long)item * a.N + q) * a.N + key) : 0.f;
              }
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) dp[j] += gq[j];
            } else {
              ptx::tmem_ld_wait();
            }
            const float4* l4p = reinterpret_cast<const float4*>(sLse + qt * 128 + col0);
            const float4* d4p = reinterpret_cast<const float4*>(sDelta + qt * 128 + col0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 l4 = l4p[j4], d4 = d4p[j4];
              const float lsq[4] = {l4.x, l4.y, l4.z, l4.w}, dlq[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = j4 * 4 + jj;
                const float p = ex2_fast(fmaf(sv[j], LOG2E, -lsq[jj]));
                if constexpr (DROP) {
                  // O = drop(P) V:  dV needs drop(P)^T, dS = P o (drop'(dP) - delta)
                  const float f = drop_factor(a.drop, drow0 + qt * 128 + col0 + j, kc * 128 + trow);
                  sv[j] = p * f;
                  dp[j] = p * (dp[j] * f - dlq[jj]);
                } else {
                  sv[j] = p;
                  dp[j] = p * (dp[j] - dlq[jj]);
                }
              }
            }
            uint32_t pP[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 hp = __floats2bfloat162_rn(sv[2 * j], sv[2 * j + 1]);
              __nv_bfloat162 hd = __floats2bfloat162_rn(dp[2 * j], dp[2 * j + 1]);
              pP[j] = *reinterpret_cast<uint32_t*>(&hp);
              pDS[j] = *reinterpret_cast<uint32_t*>(&hd);
            }
            tmem_st8(t_s, pP);
            tmem_st8(t_dp, pDS);
          }
          TR(trr, 4 + 32 * hf);
          if (hf == 0) {
            // the dS^T tile is single-buffered: the dQ MMA of the previous iteration must have retired
            if (g > 0) ptx::mbar_wait(bar_c, (g - 1) & 1);
            if (store_pending) {   // the accumulator tiles staged in the dS^T tile have been read by their TMA stores
              ptx::mbar_wait(bar_stage_free, n_grp & 1);
              ++n_grp;
              store_pending = false;
            }
          }
          TR(trr, 6 + 32 * hf);
          if (live) {
            store_bf16x8_sw128(sDS, trow, col0, pDS);
            store_bf16x8_sw128(sDS, trow, col0 + 8, pDS + 4);
          }
          ptx::tmem_st_wait();
          ptx::fence_async_shared();
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar_pds + hf);
          TR(trr, 8 + 32 * hf);
        }
        if (it == n_iter - 1) ptx::mbar_arrive(bar_item);   // lse / delta of this item have been read for the last time
        if (qt == nt - 1) {
          // ---- accumulators of this key chunk, 32 rows x 32 columns per warp (thread = key row): warpgroups 0, 1 ->
          //      the two column halves of dK, warpgroups 2, 3 -> of dV.  The producer may overwrite them as soon as
          //      they are in registers (bar_epi); packing, staging and the global stores follow. ----
          ptx::mbar_wait(bar_kv, n_chunk & 1);
          ++n_chunk;
          ptx::tc_fence_after();
          TR(trr, 10);
          const bool item_end = (kc == nt - 1);
          const int key0 = kc * 128 + quarter * 32;  // first of this warp's 32 key rows
          uint32_t wkv[16];
          if (key0 < a.N) acc_load_pack32(t_lane + (wg < 2 ? T_DK : T_DV) + (wg & 1) * 32, wkv);
          if (!item_end) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_epi);
          }
          TR(trr, 11);
          // the staging slots belong to the dS^T tile the dQ product of this iteration reads
          ptx::mbar_wait(bar_c, g & 1);
          TR(trr, 13);
          uint8_t* atom = sDS + eh * 16384;
          if (key0 < a.N) acc_stage32(wkv, atom, trow, sh);
          ptx::fence_async_shared();
          ptx::mbar_arrive(bar_staged);      // -> the store thread (TMA warp) takes [dK | dV] out
          store_pending = true;
          if (item_end) {
            // ---- dQ of query tile wg >> 1, column half wg & 1 (thread = query row): every MMA of the item has retired;
            //      the tiles go through the same two atoms once the dK / dV stores have read them
            const int tq = wg >> 1;
            const bool have = (tq < nt) && (tq * 128 + quarter * 32 < a.N);
            ptx::tc_fence_after();
            float dq_cs = 0.f;
            if (have) {
              if (a.dq_colsum) dq_cs = acc_load_pack32_colsum(t_lane + T_DQ + tq * 64 + (wg & 1) * 32, wkv, lane, tq * 128 + trow < a.N);
              else acc_load_pack32(t_lane + T_DQ + tq * 64 + (wg & 1) * 32, wkv);
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_epi);
            if (have && a.dq_colsum) atomicAdd(a.dq_colsum + (item % a.H) * HD + (wg & 1) * 32 + lane, dq_cs);
            TR(trr, 16);
            ptx::mbar_wait(bar_stage_free, n_grp & 1);
            TR(trr, 17);
            ++n_grp;
            if (have) acc_stage32(wkv, atom, trow, sh);
            ptx::fence_async_shared();
            ptx::mbar_arrive(bar_staged);
          }
          TR(trr, 12);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
#ifdef ATTN_TRACE
  if (tr_on) tr_n[warp == 0 ? 0 : warp == 8 ? 1 : 2] = tr_i;
  __threadfence();
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int w = 0; w < 3; ++w)
      for (int i = 0; i < tr_n[w]; ++i) printf("TR %d %d %u\n", w, (int)tr_id[w][i], tr[w][i]);
  }
#endif
  if (warp == BWD_MMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, BWD_TMEM_COLS);
  }
}

}  // namespace

namespace {
__global__ void fill_f32_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
}  // namespace

bool attn_fwd_tc_jasmin_supports(int N, int k) { return k >= 0 && k <= 3 && N >= 4; }

bool attn_fwd_tc_supports(int N, int D, int H, int act_type, long long ld_oh) {
  return act_type == DT_BF16 && D % H == 0 && D / H == HD && N >= 1 && N <= 256 && D % 8 == 0 && ld_oh % 8 == 0;
}

// qkv: [B, N, 3D] bf16 (q | k | v per row); oh: [B*N, ld_oh] bf16; p_out: [B,H,N,N] fp32 or null.
int attn_fwd_tc(const void* qkv, void* oh, long long ld_oh, float* p_out, float* lse_out, int B, int N, int H, int D,
                Drop drop, cudaStream_t s, float* jas_out, int jas_k) {
  if (!attn_fwd_tc_supports(N, D, H, DT_BF16, ld_oh))
    return set_error(ODEVIT_ERR_UNSUPPORTED, "attn_fwd_tc: unsupported shape N=%d D=%d H=%d", N, D, H);
  ProfScope prof(p_out ? KC_FUSED_ATTN_EXPORT : KC_FUSED_ATTN, s);
  AttnArgs a;
  a.B = B; a.N = N; a.H = H; a.D = D;
  a.NP = (N + 15) / 16 * 16;
  a.tiles_m = (N + BMQ - 1) / BMQ;
  a.oh = oh; a.ld_oh = ld_oh; a.p_out = p_out; a.lse_out = lse_out;
  a.drop = drop;
  a.jas_out = jas_out; a.jas_k = jas_k;
  if (jas_out) {
    if (p_out || drop.thresh || !attn_fwd_tc_jasmin_supports(N, jas_k))
      return set_error(ODEVIT_ERR_UNSUPPORTED, "attn_fwd_tc: in-kernel JaSMin needs k <= 3, N >= 4, no export, no dropout");
    fill_f32_kernel<<<(B * H + 255) / 256, 256, 0, s>>>(jas_out, -INFINITY, B * H);
    ODV_LAUNCH_CHECK();
  }
  CUtensorMap tq, tkv, to;
  ODV_TRY(make_tmap_3d_bf16(&tq, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, BMQ, 1));
  ODV_TRY(make_tmap_3d_bf16(&tkv, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, a.NP, 1));
  ODV_TRY(make_tmap_3d_bf16(&to, oh, D, N, B, ld_oh, (uint64_t)N * ld_oh, HD, BMQ, 1));   // O tiles of the ping-pong kernel
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  const int n_units = B * H * a.tiles_m;
  const char* env = getenv("ODEVIT_ATTN_PERSIST");
  if (!(env && env[0] == '0') && n_units >= 2 * sms && !p_out) {
    // enough units for every SM to pipeline: the persistent ping-pong kernel (the exporting variant is bound
    // by its row stores and measured slower there: it stays CTA-per-unit)
    const int smem_pp = 2 * (BMQ * HD * 2 + a.NP * HD * 2) + PP_V_STAGES * a.NP * HD * 2 + 2 * BMQ * HD * 2 + 1024 + 256;
    static bool configured_pp = false;
    if (!configured_pp) {
      const int max_smem = 227 * 1024;
      ODV_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      ODV_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      ODV_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      configured_pp = true;
    }
    const int grid_pp = n_units < sms ? n_units : sms;
    if (jas_out) attn_fwd_pp_kernel<false, false, true><<<grid_pp, PP_THREADS, smem_pp, s>>>(tq, tkv, to, a);
    else if (drop.thresh) attn_fwd_pp_kernel<false, true><<<grid_pp, PP_THREADS, smem_pp, s>>>(tq, tkv, to, a);
    else attn_fwd_pp_kernel<false, false><<<grid_pp, PP_THREADS, smem_pp, s>>>(tq, tkv, to, a);
    ODV_LAUNCH_CHECK();
    return 0;
  }
  const int smem = BMQ * HD * 2 + 2 * a.NP * HD * 2 + 1024 + 64;
  const int grid = B * H * a.tiles_m;
  static bool configured = false;
  if (!configured) {
    const int max_smem = BMQ * HD * 2 + 2 * 256 * HD * 2 + 1024 + 64 + 4 * 32 * 33 * 4;
    ODV_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    ODV_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    ODV_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    ODV_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    ODV_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  const int smem_x = smem + 4 * 32 * 33 * 4;   // + the export staging tiles
  if (jas_out) {
    attn_fwd_tc_kernel<false, false, true><<<grid, 128, smem, s>>>(tq, tkv, a);
  } else if (drop.thresh) {
    if (p_out) attn_fwd_tc_kernel<true, true><<<grid, 128, smem_x, s>>>(tq, tkv, a);
    else attn_fwd_tc_kernel<false, true><<<grid, 128, smem, s>>>(tq, tkv, a);
  } else {
    if (p_out) attn_fwd_tc_kernel<true, false><<<grid, 128, smem_x, s>>>(tq, tkv, a);
    else attn_fwd_tc_kernel<false, false><<<grid, 128, smem, s>>>(tq, tkv, a);
  }
  ODV_LAUNCH_CHECK();
  return 0;
}

size_t attn_bwd_tc_scratch_floats(int B, int N, int H) {
  (void)B; (void)N; (void)H;
  return 0;  // dQ accumulates in tensor memory across key chunks
}

// qkv [B,N,3D] bf16; dO [B*N, D] bf16; O = oh[:, 0:D] (ld_oh); lse2 [B,H,N] fp32; dz [B*N, R] bf16
// receives dq | dk | dv.  (`delta`, `dq_scratch` are unused: delta is formed in the kernel.)
int attn_bwd_tc(const void* qkv, const void* dO, const void* oh, long long ld_oh, const float* lse2, float* delta,
                void* dz, int R, float* dq_scratch, int B, int N, int H, int D, Drop drop, cudaStream_t s, const float* gp,
                const float* dext, float* dq_colsum) {
  (void)delta; (void)dq_scratch;
  if ((gp != nullptr) != (dext != nullptr) || (gp && drop.thresh))
    return set_error(ODEVIT_ERR_INVALID_ARG, "attn_bwd_tc: map cotangent needs its row dots and no dropout");
  if (!attn_fwd_tc_supports(N, D, H, DT_BF16, ld_oh) || R % 8)
    return set_error(ODEVIT_ERR_UNSUPPORTED, "attn_bwd_tc: unsupported shape N=%d D=%d H=%d", N, D, H);
  ProfScope prof(KC_FUSED_ATTN_BWD, s);
  AttnBwdArgs a;
  a.B = B; a.N = N; a.H = H; a.D = D; a.R = R;
  a.n_t = (N + 127) / 128;
  a.items = B * H;
  a.lse2 = lse2; a.dz = dz;
  a.drop = drop;
  a.gp = gp; a.dext = dext;
  a.dq_colsum = dq_colsum;
  a.dO = reinterpret_cast<const __nv_bfloat16*>(dO);
  a.O = reinterpret_cast<const __nv_bfloat16*>(oh);
  a.ld_o = ld_oh;
  CUtensorMap tqkv, tdo;
  ODV_TRY(make_tmap_3d_bf16(&tqkv, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, 128, 1));
  ODV_TRY(make_tmap_3d_bf16(&tdo, dO, D, N, B, D, (uint64_t)N * D, HD, 128, 1));
  CUtensorMap tdz;   // dq | dk | dv tiles leave through TMA stores: [B][N][R], boxes of 128 rows x 64 columns
  ODV_TRY(make_tmap_3d_bf16(&tdz, dz, R, N, B, R, (uint64_t)N * R, HD, 128, 1));
  const int smem = 3 * SET + 32768 + 2048 + 256;   // operand sets, dS^T tile, lse | delta, 17 mbarriers + the TMEM slot
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    ODV_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ODV_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ODV_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev = 0;
    ODV_CUDA(cudaGetDevice(&dev));
    ODV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const int grid = a.items < sms ? a.items : sms;
  if (gp) attn_bwd_tc_kernel<false, true><<<grid, BWD_THREADS, smem, s>>>(tqkv, tdo, tdz, a);
  else if (drop.thresh) attn_bwd_tc_kernel<true, false><<<grid, BWD_THREADS, smem, s>>>(tqkv, tdo, tdz, a);
  else attn_bwd_tc_kernel<false, false><<<grid, BWD_THREADS, smem, s>>>(tqkv, tdo, tdz, a);
  ODV_LAUNCH_CHECK();
  return 0;
}

}  // namespace odevit
