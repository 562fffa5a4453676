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
  for (int c0 = 0; c0 < n_chunks; c0 += 4) {
    float v[4][16];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c0 + u < n_chunks) ptx::tmem_ld16(t_row + (c0 + u) * 16, v[u]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c0 + u < n_chunks) {
        if constexpr (JAS) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            top4_insert(tt[j & 3], ((c0 + u) * 16 + j < a.N) ? v[u][j] : -INFINITY);
        } else if ((c0 + u + 1) * 16 <= a.N) {   // every column of the chunk is a key: no per-element predicate
#pragma unroll
          for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], v[u][j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if ((c0 + u) * 16 + j < a.N) m4[j & 3] = fmaxf(m4[j & 3], v[u][j]);
        }
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
    for (int c0 = 0; c0 < n_chunks; c0 += 4) {
      float v[4][16];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + u < n_chunks) ptx::tmem_ld16(t_row + (c0 + u) * 16, v[u]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + u < n_chunks) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if ((c0 + u) * 16 + j < a.N) s4[j & 3] += ex2_fast(fmaf(v[u][j], LOG2E, -mxs));
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
        const bool full = (c + 1) * 16 <= a.N;   // every column of the chunk is a key
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float e = ex2_fast(fmaf(v[u][j], LOG2E, -mxs));
          if (!full) e = (c * 16 + j < a.N) ? e : 0.f;
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
  const float inv = attn_softmax_row<EXPORT, DROP, JAS>(a, t_row, n_chunks, b, h, row, stage);
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
// producer warp streams Q / K / V of the units through a ring of shared-memory stages (three when they fit:
// loads run a full unit ahead) and issues both MMAs of every unit, S of unit u+1 BEFORE P.V of unit u, so
// the tensor core, the TMA engine and the two warpgroups overlap.  What a CTA-per-unit launch serialises for
// every unit (launch, TMEM allocation, barrier setup, the TMA round trip) is paid once per SM here.
//   warp 8         TMA + MMA issue (predicated on the elected lane)
//   warps 0-3      slot 0: units 0, 2, 4, ... of this CTA      (thread = query row, quadrant = warp & 3)
//   warps 4-7      slot 1: units 1, 3, 5, ...
constexpr int PP_THREADS = 288;
constexpr int PP_MAX_STAGES = 3;

template <bool EXPORT, bool DROP, bool JAS = false>
__global__ void __launch_bounds__(PP_THREADS, 1)
attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ AttnArgs a, int n_stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_bytes = a.NP * HD * 2;
  const int stage_bytes = BMQ * HD * 2 + 2 * kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * stage_bytes);
  uint64_t* bar_qk = bars;                         // [stage] Q and K landed
  uint64_t* bar_v = bars + PP_MAX_STAGES;          // [stage] V landed
  uint64_t* bar_free = bars + 2 * PP_MAX_STAGES;   // [stage] P.V of the stage's unit has retired
  uint64_t* bar_s = bars + 3 * PP_MAX_STAGES;      // [slot] S is in tensor memory
  uint64_t* bar_p = bar_s + 2;                     // [slot] P is packed (128 arrivals)
  uint64_t* bar_o = bar_s + 4;                     // [slot] O is in tensor memory
  uint64_t* bar_epi = bar_s + 6;                   // [slot] the epilogue has read O (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = a.B * a.H;
  const int n_units = items * a.tiles_m;
  const int n_mine = (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmKV);
    for (int i = 0; i < PP_MAX_STAGES; ++i) { ptx::mbar_init(bar_qk + i, 1); ptx::mbar_init(bar_v + i, 1); ptx::mbar_init(bar_free + i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_s + i, 1); ptx::mbar_init(bar_p + i, 128); ptx::mbar_init(bar_o + i, 1); ptx::mbar_init(bar_epi + i, 128); }
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

  if (warp == 8) {
    const bool leader = ptx::elect_one();
    const uint32_t lead = leader ? 1u : 0u;
    const uint32_t idesc1 = ptx::idesc_bf16(BMQ, a.NP, 0, 0);
    const uint32_t idesc2 = ptx::idesc_bf16(BMQ, HD, 0, 1);
    const uint64_t DESC0 = ptx::smem_desc_sw128(0, 16, 1024);
    const uint32_t smem0 = ptx::smem_u32(smem);
    int loaded = 0;
    auto issue_load = [&](int i) {
      const int st = i % n_stages;
      if (i >= n_stages) ptx::mbar_wait(bar_free + st, (uint32_t)((i / n_stages - 1) & 1));
      if (leader) {
        int b, h, mt;
        decode(i, b, h, mt);
        uint8_t* q = smem + st * stage_bytes;
        ptx::mbar_expect_tx(bar_qk + st, BMQ * HD * 2 + kv_bytes);
        ptx::tma_load_3d(q, &tmQ, bar_qk + st, h * HD, mt * BMQ, b);
        ptx::tma_load_3d(q + BMQ * HD * 2, &tmKV, bar_qk + st, a.D + h * HD, 0, b);
        ptx::mbar_expect_tx(bar_v + st, kv_bytes);
        ptx::tma_load_3d(q + BMQ * HD * 2 + kv_bytes, &tmKV, bar_v + st, 2 * a.D + h * HD, 0, b);
      }
      __syncwarp();
    };
    // P.V of unit v, then the refill of the oldest free stage
    auto finish = [&](int v) {
      const int sv = v & 1, stv = v % n_stages;
      ptx::mbar_wait(bar_p + sv, (uint32_t)((v >> 1) & 1));
      ptx::mbar_wait(bar_v + stv, (uint32_t)((v / n_stages) & 1));
      ptx::tc_fence_after();
      const uint32_t v_addr = smem0 + stv * stage_bytes + BMQ * HD * 2 + kv_bytes;
      const uint32_t t = tmem + sv * 256;
      for (int k = 0; k < a.NP / 16; ++k)
        ptx::mma_ts_pred(t + O_COL, t + k * 8, ptx::smem_desc_sw128(v_addr + k * 2048, 8192, 1024), idesc2, k > 0 ? 1u : 0u, lead);
      ptx::commit_pred(ptx::smem_u32(bar_o + sv), lead);
      ptx::commit_pred(ptx::smem_u32(bar_free + stv), lead);
      if (loaded < n_mine) { issue_load(loaded); ++loaded; }
    };
    for (; loaded < n_stages && loaded < n_mine; ++loaded) issue_load(loaded);
    for (int u = 0; u < n_mine; ++u) {
      const int sl = u & 1, st = u % n_stages;
      ptx::mbar_wait(bar_qk + st, (uint32_t)((u / n_stages) & 1));
      if (u >= 2) ptx::mbar_wait(bar_epi + sl, (uint32_t)(((u >> 1) - 1) & 1));   // the slot's previous O has been read
      ptx::tc_fence_after();
      const uint64_t dq = DESC0 + ((smem0 + st * stage_bytes) >> 4), dk = dq + ((BMQ * HD * 2) >> 4);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) ptx::mma_ss_pred(tmem + sl * 256, dq + 2 * k, dk + 2 * k, idesc1, k > 0 ? 1u : 0u, lead);
      ptx::commit_pred(ptx::smem_u32(bar_s + sl), lead);
      if (u >= 1) finish(u - 1);
    }
    if (n_mine > 0) finish(n_mine - 1);
  } else {
    const int wg = warp >> 2, quarter = warp & 3;
    const uint32_t t_row = tmem + wg * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    for (int u = wg; u < n_mine; u += 2) {
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      int b, h, mt;
      decode(u, b, h, mt);
      const int row = mt * BMQ + quarter * 32 + lane;
      const int n_chunks = (mt * BMQ + quarter * 32 < a.N) ? a.NP / 16 : 0;
      ptx::mbar_wait(bar_s + wg, ph);
      ptx::tc_fence_after();
      const float inv = attn_softmax_row<EXPORT, DROP, JAS>(a, t_row, n_chunks, b, h, row);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_p + wg);
      ptx::mbar_wait(bar_o + wg, ph);
      ptx::tc_fence_after();
      attn_store_o_row<EXPORT>(a, t_row + O_COL, b, h, row, inv);
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_epi + wg);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ================================================================================================
// Fused attention VJP (bf16 mode, head dim 64, N <= 256).  Per (image b, head h), with the
// log-sum-exp `lse2` of every row saved by the forward and delta_i = sum_d dO_id O_id:
//     S = q k^T,  P = exp2(S*log2e - lse2),  dP = dO v^T,  dS = P o (dP - delta)
//     dq = dS k,  dk = dS^T q,  dv = P^T dO
// PERSISTENT kernel: one CTA per SM walks the (b, h) items c, c+G, c+2G, ...; everything is in the
// TRANSPOSED orientation (thread = key row), so that P^T and dS^T are A operands that never leave
// tensor memory, and the operand tiles of the NEXT item stream in while this one is computed:
//   warp 8 (one lane)  TMA + all MMAs.  Operand tiles live in three rotating 64 KB sets
//        {K, Q, V, dO} x 128 rows: item i keeps rows [0,128) in set s0 and rows [128,256) in s1; the
//        third set receives rows [0,128) of item i+1 at the start of item i, and s0 is refilled with
//        rows [128,256) of item i+1 as soon as its last reader has retired.
//        per (key chunk kc, query tile qt):
//          S^T  = K_kc Q_qt^T,  dP^T = V_kc dO_qt^T                 (SS)  -> TMEM [0,128), [128,256)
//          dV_kc += P^T dO_qt,  dK_kc += dS^T Q_qt                  (TS: A = bf16 P^T / dS^T in TMEM)
//          dQ_qt += dS K_kc     (A = dS^T tile in shared memory read MN-major, B = K chunk MN-major)
//        dK, dV, dQ_0, dQ_1 accumulate in TMEM [256,512): no partial sum ever goes through memory.
//   warps 0-7          two warpgroups split the query columns of a tile; thread = key row:
//          P^T, dS^T from S^T, dP^T (per-query lse / delta broadcast from shared memory), packed to
//          bf16 in registers; dS^T also goes to a 128-byte-swizzled shared-memory tile; after a
//          barrier among the 256 threads the packed values overwrite S^T / dP^T in place
//          (tcgen05.st).  Epilogues: dK | dV per key chunk (one warpgroup each), dQ at the item's end.
//   warps 9-12         lse and delta = <dO_i, O_i> of the NEXT item into shared memory (double
//          buffered): the only global loads that are not TMA, kept off the compute warps.
// The cotangent of an exported P (`attentions`, last evaluation only) is not handled here; that
// single evaluation takes the CUDA-core path (api.cu::attention_vjp).
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
};

constexpr int BWD_TMEM_COLS = 512;
constexpr int T_ST = 0, T_DPT = 128, T_DK = 256, T_DV = 320, T_DQ = 384;  // T_DQ + 64*qt
constexpr int BWD_THREADS = 13 * 32;
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
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// This warp's 32 accumulator rows x 64 fp32 columns -> bf16 -> global rows of `stride` elements.
// Thread = row out of TMEM; a 4 KB staging tile (16-byte chunks XOR-swizzled by row) turns the
// row-per-thread layout into 128 contiguous bytes per 8 lanes, so every store instruction covers
// 4 full lines instead of 32 partial ones.
__device__ __forceinline__ void store_rows_bf16(__nv_bfloat16* g_row0, long long stride, uint32_t t_addr,
                                                uint8_t* stage, int lane, int rows_valid) {
  float v[64];
#pragma unroll
  for (int c = 0; c < HD / 16; ++c) ptx::tmem_ld16(t_addr + c * 16, v + c * 16);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 hh = __floats2bfloat162_rn(v[c * 8 + 2 * j], v[c * 8 + 2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&hh);
    }
    *reinterpret_cast<uint4*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncwarp();
  const int c = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 128 + ((c ^ (row & 7)) << 4));
    if (row < rows_valid) *reinterpret_cast<uint4*>(g_row0 + row * stride + c * 8) = val;
  }
  __syncwarp();
}

#ifdef ATTN_TRACE
__device__ uint32_t tr[3][100];
__device__ uint8_t tr_id[3][100];
__device__ int tr_n[3];
#define TRC(slot) ((slot) >= 100 ? 2 : ((slot) & 1))
#define TR(slot) do { if (blockIdx.x == 0 && tr_on && tr_n[TRC(slot)] < 100) { const int _c = TRC(slot); const int _i = tr_n[_c]++; tr[_c][_i] = (uint32_t)clock(); tr_id[_c][_i] = (slot); } } while (0)
#else
#define TR(slot) do { } while (0)
#endif

template <bool DROP, bool GP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ AttnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sSets = smem;                       // 3 sets x 4 slots x 16 KB
  uint8_t* sDS = sSets + 3 * SET;              // 32 KB: dS^T tile
  float* sLse = reinterpret_cast<float*>(sDS + 32768);   // [256]
  float* sDelta = sLse + 256;                             // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* bar_full = bars;      // [3] operand set landed
  uint64_t* bar_sdp = bars + 3;   // S^T / dP^T of an iteration are in TMEM
  uint64_t* bar_pds = bars + 4;   // P^T / dS^T written (TMEM + shared memory)
  uint64_t* bar_c = bars + 5;     // the dV / dK / dQ MMAs of an iteration have retired
  uint64_t* bar_epi = bars + 6;   // accumulators of a key chunk (and dQ at the item's end) read out
  uint64_t* bar_aux = bars + 7;   // lse / delta of an item are in shared memory
  uint64_t* bar_item = bars + 8;  // an item is finished by the compute warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
#ifdef ATTN_TRACE
  if (threadIdx.x == 0 && blockIdx.x == 0) { tr_n[0] = 0; tr_n[1] = 0; tr_n[2] = 0; }
#endif

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef ATTN_TRACE
  const bool tr_on = (threadIdx.x == 0) || (warp == 8 && lane == 0) || (warp == 9 && lane == 0);
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) ptx::mbar_init(bar_full + i, 1);
    ptx::mbar_init(bar_sdp, 1);
    ptx::mbar_init(bar_pds, 256);
    ptx::mbar_init(bar_c, 1);
    ptx::mbar_init(bar_epi, 256);
    ptx::mbar_init(bar_aux, 128);
    ptx::mbar_init(bar_item, 256);
    ptx::fence_barrier_init();
  }
  if (warp == 8) ptx::tmem_alloc(tmem_slot, BWD_TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int NP = (a.N + 15) / 16 * 16;
  const int nt = a.n_t;
  const int n_iter = nt * nt;
  const int n_mine = (a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 8) {
    // =========================== producer: TMA + every MMA ===========================
    // The whole warp walks the schedule (uniform control flow keeps descriptors in uniform registers);
    // only the elected lane issues TMA / MMA / commit.  The schedule is flattened over (item, kc, qt):
    // the descriptors of iteration gi+1 are formed BEFORE waiting for the compute warps of iteration
    // gi, so that S^T / dP^T of gi+1 are issued back to back with the accumulating MMAs of gi.
    const bool leader = ptx::elect_one();
    if (leader) {
      ptx::prefetch_tensormap(&tmQKV);
      ptx::prefetch_tensormap(&tmDO);
    }
    uint32_t ph_full[3] = {0, 0, 0};
    uint32_t n_epi = 0;      // accumulator read-outs the producer has waited for
    uint32_t n_epi_due = 0;  // read-outs the compute warps will have signalled before the next first-of-chunk C
    // rows [128*t, 128*t+128) of item `item` -> set `set`
    auto load_set = [&](int item, int t, int set) {
      const int hh = item % a.H, bb = item / a.H;
      uint8_t* base = sSets + set * SET;
      if (leader) {
        ptx::mbar_expect_tx(bar_full + set, SET);
        ptx::tma_load_3d(base + SL_K * SLOT, &tmQKV, bar_full + set, a.D + hh * HD, t * 128, bb);
        ptx::tma_load_3d(base + SL_Q * SLOT, &tmQKV, bar_full + set, hh * HD, t * 128, bb);
        ptx::tma_load_3d(base + SL_V * SLOT, &tmQKV, bar_full + set, 2 * a.D + hh * HD, t * 128, bb);
        ptx::tma_load_3d(base + SL_DO * SLOT, &tmDO, bar_full + set, hh * HD, t * 128, bb);
      }
      __syncwarp();
    };
    struct Iter {
      uint64_t dk_k, dq_k, dv_k, ddo_k;   // K-major descriptors (S^T, dP^T)
      uint64_t ddo_mn, dq_mn, dk_mn;      // MN-major descriptors (dV, dK, dQ)
      uint32_t id_s, acc_q, acc_k;
      int nq, nk, qt, it;
    };
    // nt == 2: an item holds two sets (rows [0,128) in s0, rows [128,256) in s1), the third one (sf)
    //          receives rows [0,128) of the next item during iteration 0 and s0 is refilled with its rows
    //          [128,256) during iteration 3;  (s0, s1, sf) <- (sf, s0, s1) at every item boundary.
    // nt == 1: an item holds one set; item i lives in set i % 3 and is loaded two items ahead.
    int s0 = 0, s1 = 1, sf = 2;
    bool t1_ready = (nt == 1);
    // describes iteration `it` of the item whose sets are (s0, s1); waits for the operand sets it touches first
    auto make_iter = [&](int it) {
      Iter r;
      const int kc = (nt == 2) ? (it >> 1) : 0, qt = (nt == 2) ? (it & 1) : 0;
      if (it == 0) {
        ptx::mbar_wait(bar_full + s0, ph_full[s0]);
        ph_full[s0] ^= 1;
        t1_ready = (nt == 1);
      } else if (!t1_ready) {
        ptx::mbar_wait(bar_full + s1, ph_full[s1]);
        ph_full[s1] ^= 1;
        t1_ready = true;
      }
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
      r.id_s = ptx::idesc_bf16(128, qw, 0, 0);
      r.nq = qw / 16; r.nk = cw / 16;
      r.acc_q = qt > 0 ? 1u : 0u; r.acc_k = kc > 0 ? 1u : 0u;
      r.qt = qt; r.it = it;
      return r;
    };
    auto issue_a = [&](const Iter& r) {  // S^T = K Q^T, dP^T = V dO^T
      if (leader) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) ptx::mma_bf16_ss(tmem + T_ST, r.dk_k + 2 * k, r.dq_k + 2 * k, r.id_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) ptx::mma_bf16_ss(tmem + T_DPT, r.dv_k + 2 * k, r.ddo_k + 2 * k, r.id_s, k > 0 ? 1u : 0u);
        ptx::mma_commit(bar_sdp);
      }
      __syncwarp();
    };
    const uint32_t id_ts = ptx::idesc_bf16(128, HD, 0, 1);   // A from TMEM, B MN-major
    const uint32_t id_dq = ptx::idesc_bf16(128, HD, 1, 1);   // A, B MN-major
    const uint64_t dds_mn = ptx::smem_desc_sw128(ptx::smem_u32(sDS), 16384, 1024);
    const int total = n_mine * n_iter;
    if (n_mine > 0) {
      load_set(blockIdx.x, 0, s0);
      if (nt > 1) load_set(blockIdx.x, 1, s1);
      else if (n_mine > 1) load_set((int)blockIdx.x + (int)gridDim.x, 0, 1);
      TR(1);
      Iter cur = make_iter(0);
      issue_a(cur);
      int idx = 0;
      for (int gi = 0; gi < total; ++gi) {
        // ---- operand prefetch, into sets whose last reader has retired ----
        if (nt == 1) {
          if (idx + 2 < n_mine) {
            if (gi > 0) ptx::mbar_wait(bar_c, (gi - 1) & 1);
            load_set((int)blockIdx.x + (idx + 2) * (int)gridDim.x, 0, (idx + 2) % 3);   // the set of item idx-1
          }
        } else if (idx + 1 < n_mine && (cur.it == 0 || cur.it == 3)) {
          const int next_item = (int)blockIdx.x + (idx + 1) * (int)gridDim.x;
          if (gi > 0) ptx::mbar_wait(bar_c, (gi - 1) & 1);   // every accumulating MMA issued so far has retired
          if (cur.it == 0) load_set(next_item, 0, sf);        // sf: rows [128,256) of the previous item
          else load_set(next_item, 1, s0);                    // s0: last read by iteration 2
        }
        // ---- the next iteration's descriptors, before the wait ----
        const bool has_next = gi + 1 < total;
        Iter nxt = cur;
        if (has_next) {
          if (cur.it == n_iter - 1) {
            ++idx;
            if (nt == 1) { s0 = idx % 3; }
            else { const int t = s0; s0 = sf; sf = s1; s1 = t; }   // (s0, s1, sf) <- (sf, s0, s1)
            nxt = make_iter(0);
          } else {
            nxt = make_iter(cur.it + 1);
          }
        }
        TR(3);
        // ---- wait for P^T / dS^T (TMEM) and dS^T (shared memory) of iteration gi ----
        ptx::mbar_wait(bar_pds, gi & 1);
        ptx::tc_fence_after();
        if (cur.qt == 0 && n_epi < n_epi_due) {  // the accumulators of the previous chunk have been read out
          ptx::mbar_wait(bar_epi, n_epi & 1);
          ++n_epi;
          ptx::tc_fence_after();
        }
        TR(5);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // contraction over the tile's queries (16 rows = 2048 B = 128 units)
            if (ks < cur.nq) {
              ptx::mma_bf16_ts(tmem + T_DV, tmem + T_ST + ks * 8, cur.ddo_mn + 128 * ks, id_ts, ks > 0 ? 1u : cur.acc_q);
              ptx::mma_bf16_ts(tmem + T_DK, tmem + T_DPT + ks * 8, cur.dq_mn + 128 * ks, id_ts, ks > 0 ? 1u : cur.acc_q);
            }
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)    // contraction over the chunk's keys
            if (ks < cur.nk)
              ptx::mma_bf16_ss(tmem + T_DQ + cur.qt * 64, dds_mn + 128 * ks, cur.dk_mn + 128 * ks, id_dq,
                               ks > 0 ? 1u : cur.acc_k);
          ptx::mma_commit(bar_c);
        }
        __syncwarp();
        if (cur.qt == nt - 1) ++n_epi_due;  // the compute warps read this chunk's accumulators out next
        if (has_next) issue_a(nxt);
        TR(7);
        cur = nxt;
      }
    }
  } else if (warp >= 9) {
    // =========================== lse / delta loader (128 threads) ===========================
    // The values of item idx are formed in registers while item idx-1 is being computed and dropped
    // into the (single) shared-memory buffer the moment that item is finished.
    const int tl = threadIdx.x - 9 * 32;
    for (int idx = 0; idx < n_mine; ++idx) {
      const int item = (int)blockIdx.x + idx * (int)gridDim.x;
      const int hh = item % a.H, bb = item / a.H;
      TR(100);
      float lse[2], dl[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int q = tl + r * 128;
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
      TR(102);
      if (idx >= 1) ptx::mbar_wait(bar_item, (idx - 1) & 1);  // the previous item no longer reads the buffer
      TR(104);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        sLse[tl + r * 128] = lse[r];
        sDelta[tl + r * 128] = dl[r];
      }
      ptx::mbar_arrive(bar_aux);
    }
  } else {
    // =========================== compute warps (thread = key row) ===========================
    const int wg = warp >> 2, quarter = warp & 3;
    const int trow = quarter * 32 + lane;  // row inside a 128-row tile / chunk
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(a.dz);
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t g = 0;  // global iteration counter (phases of bar_sdp / bar_c)
    for (int idx = 0; idx < n_mine; ++idx) {
      const int item = (int)blockIdx.x + idx * (int)gridDim.x;
      const int h = item % a.H, b = item / a.H;
      const float* lse_s = sLse;
      const float* dl_s = sDelta;
      const uint32_t drow0 = (uint32_t)((long long)item * a.N);   // dropout row coordinate of query 0
      TR(22);
      ptx::mbar_wait(bar_aux, idx & 1);
      TR(24);
      for (int it = 0; it < n_iter; ++it, ++g) {
        const int kc = it / nt, qt = it - kc * nt;
        const int qw = min(128, NP - qt * 128);
        const int nch = qw / 16;
        const int c_lo = wg ? (nch + 1) / 2 : 0, c_hi = wg ? nch : (nch + 1) / 2;   // this warpgroup's chunks
        const bool key_ok = (kc * 128 + trow) < a.N;
        TR(0);
        ptx::mbar_wait(bar_sdp, g & 1);
        ptx::tc_fence_after();
        TR(2);
        uint32_t pP[4][8], pDS[4][8];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = c_lo + ci;
          if (c < c_hi) {
            float sv[16], dp[16];
            ptx::tmem_ld16(t_lane + T_ST + c * 16, sv);
            ptx::tmem_ld16(t_lane + T_DPT + c * 16, dp);
            // per-query lse / delta of the chunk: 16-byte broadcast loads (the row offset is a multiple of 16 floats)
            float lsq[16], dlq[16];
            float gq[GP ? 16 : 1];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 l4 = reinterpret_cast<const float4*>(lse_s + qt * 128 + c * 16)[j4];
              const float4 d4 = reinterpret_cast<const float4*>(dl_s + qt * 128 + c * 16)[j4];
              lsq[4 * j4] = l4.x; lsq[4 * j4 + 1] = l4.y; lsq[4 * j4 + 2] = l4.z; lsq[4 * j4 + 3] = l4.w;
              dlq[4 * j4] = d4.x; dlq[4 * j4 + 1] = d4.y; dlq[4 * j4 + 2] = d4.z; dlq[4 * j4 + 3] = d4.w;
            }
            if constexpr (GP) {
              // + the cotangent of the exported map, read transposed: lanes = consecutive keys of query row q
              const int key = kc * 128 + trow;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int q = qt * 128 + c * 16 + j;
                gq[j] = (key_ok && q < a.N) ? __ldg(a.gp + ((long long)item * a.N + q) * a.N + key) : 0.f;
              }
            }
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int q = qt * 128 + c * 16 + j;
              if constexpr (GP) dp[j] += gq[j];
              const float p = key_ok ? ex2_fast(fmaf(sv[j], LOG2E, -lsq[j])) : 0.f;
              if constexpr (DROP) {
                // O = drop(P) V:  dV needs drop(P)^T, dS = P o (drop'(dP) - delta)
                const float f = drop_factor(a.drop, drow0 + q, kc * 128 + trow);
                sv[j] = p * f;
                dp[j] = p * (dp[j] * f - dlq[j]);
              } else {
                sv[j] = p;
                dp[j] = p * (dp[j] - dlq[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 hp = __floats2bfloat162_rn(sv[2 * j], sv[2 * j + 1]);
              __nv_bfloat162 hd = __floats2bfloat162_rn(dp[2 * j], dp[2 * j + 1]);
              pP[ci][j] = *reinterpret_cast<uint32_t*>(&hp);
              pDS[ci][j] = *reinterpret_cast<uint32_t*>(&hd);
            }
          }
        }
        // the dS^T tile is single-buffered: the dQ MMA of the previous iteration must have retired
        TR(4);
        if (g > 0) ptx::mbar_wait(bar_c, (g - 1) & 1);
        TR(6);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = c_lo + ci;
          if (c < c_hi) {
            store_bf16x8_sw128(sDS, trow, c * 16, pDS[ci]);
            store_bf16x8_sw128(sDS, trow, c * 16 + 8, pDS[ci] + 4);
          }
        }
        ptx::tc_fence_before();
        TR(8);
        compute_bar_sync();  // every thread has read its S^T / dP^T columns: they may be overwritten
        ptx::tc_fence_after();
        TR(10);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = c_lo + ci;
          if (c < c_hi) {
            tmem_st8(t_lane + T_ST + c * 8, pP[ci]);
            tmem_st8(t_lane + T_DPT + c * 8, pDS[ci]);
          }
        }
        ptx::tmem_st_wait();
        TR(12);
        ptx::fence_async_shared();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_pds);
        TR(14);
        if (qt == nt - 1) {
          // ---- dK | dV of this key chunk: warpgroup 0 -> dK, warpgroup 1 -> dV (thread = key row) ----
          ptx::mbar_wait(bar_c, g & 1);
          ptx::tc_fence_after();
          TR(16);
          // (the dS^T tile is idle here -- every MMA that reads it has retired -- and serves as staging)
          uint8_t* stage = sDS + warp * 4096;
          const int key0 = kc * 128 + quarter * 32;  // first of this warp's 32 key rows
          store_rows_bf16(dz + ((long long)b * a.N + key0) * a.R + (wg + 1) * a.D + h * HD, a.R,
                          t_lane + (wg ? T_DV : T_DK), stage, lane, a.N - key0);
          if (kc == nt - 1 && wg < nt) {
            // ---- dQ: warpgroup qt reads tile qt (thread = query row); every MMA of the item has retired
            const int q0 = wg * 128 + quarter * 32;
            store_rows_bf16(dz + ((long long)b * a.N + q0) * a.R + h * HD, a.R, t_lane + T_DQ + wg * 64, stage, lane,
                            a.N - q0);
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar_epi);
          compute_bar_sync();  // staging regions are dS^T rows of other warps in the next iteration
          TR(18);
        }
      }
      TR(20);
      ptx::mbar_arrive(bar_item);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
#ifdef ATTN_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int w = 0; w < 3; ++w)
      for (int i = 0; i < tr_n[w]; ++i) printf("TR %d %d %u\n", w, (int)tr_id[w][i], tr[w][i]);
  }
#endif
  if (warp == 8) {
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
  CUtensorMap tq, tkv;
  ODV_TRY(make_tmap_3d_bf16(&tq, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, BMQ, 1));
  ODV_TRY(make_tmap_3d_bf16(&tkv, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, a.NP, 1));
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
    const int stage_bytes = BMQ * HD * 2 + 2 * a.NP * HD * 2;
    const int n_stages = (PP_MAX_STAGES * stage_bytes + 2048 <= 227 * 1024) ? PP_MAX_STAGES : 2;
    const int smem_pp = n_stages * stage_bytes + 1024 + 256;
    static bool configured_pp = false;
    if (!configured_pp) {
      const int max_smem = 227 * 1024;
      ODV_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      ODV_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      ODV_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      configured_pp = true;
    }
    const int grid_pp = n_units < sms ? n_units : sms;
    if (jas_out) attn_fwd_pp_kernel<false, false, true><<<grid_pp, PP_THREADS, smem_pp, s>>>(tq, tkv, a, n_stages);
    else if (drop.thresh) attn_fwd_pp_kernel<false, true><<<grid_pp, PP_THREADS, smem_pp, s>>>(tq, tkv, a, n_stages);
    else attn_fwd_pp_kernel<false, false><<<grid_pp, PP_THREADS, smem_pp, s>>>(tq, tkv, a, n_stages);
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
                const float* dext) {
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
  a.dO = reinterpret_cast<const __nv_bfloat16*>(dO);
  a.O = reinterpret_cast<const __nv_bfloat16*>(oh);
  a.ld_o = ld_oh;
  CUtensorMap tqkv, tdo;
  ODV_TRY(make_tmap_3d_bf16(&tqkv, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, 128, 1));
  ODV_TRY(make_tmap_3d_bf16(&tdo, dO, D, N, B, D, (uint64_t)N * D, HD, 128, 1));
  const int smem = 3 * SET + 32768 + 2048 + 128;
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
  if (gp) attn_bwd_tc_kernel<false, true><<<grid, BWD_THREADS, smem, s>>>(tqkv, tdo, a);
  else if (drop.thresh) attn_bwd_tc_kernel<true, false><<<grid, BWD_THREADS, smem, s>>>(tqkv, tdo, a);
  else attn_bwd_tc_kernel<false, false><<<grid, BWD_THREADS, smem, s>>>(tqkv, tdo, a);
  ODV_LAUNCH_CHECK();
  return 0;
}

}  // namespace odevit
