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

struct AttnArgs {
  int B, N, H, NP, tiles_m;
  int D;                 // embed dim (column offsets of k, v inside the packed row)
  void* oh;              // [B*N, ld_oh] bf16; O goes to columns [h*64, h*64+64)
  long long ld_oh;
  float* p_out;          // [B, H, N, N] fp32 or null
  float* lse_out;        // [B, H, N] fp32 or null: log2-domain log-sum-exp of each row (for the VJP)
};

template <bool EXPORT>
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
  const int n_chunks = a.NP / 16;
  constexpr float LOG2E = 1.4426950408889634f;
  float mx = -INFINITY;
  for (int c = 0; c < n_chunks; ++c) {
    float v[16];
    ptx::tmem_ld16(t_row + c * 16, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c * 16 + j < a.N) mx = fmaxf(mx, v[j]);
  }
  const float mxs = mx * LOG2E;
  float sum = 0.f;
  float inv = 1.f;
  if constexpr (EXPORT) {
    for (int c = 0; c < n_chunks; ++c) {
      float v[16];
      ptx::tmem_ld16(t_row + c * 16, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c * 16 + j < a.N) sum += exp2f(fmaf(v[j], LOG2E, -mxs));
    }
    inv = 1.f / sum;
  }
  const int row = mt * BMQ + warp * 32 + lane;
  float* p_row = nullptr;
  if constexpr (EXPORT) {
    if (a.p_out && row < a.N) p_row = a.p_out + (((long long)b * a.H + h) * a.N + row) * a.N;
  }
  for (int c = 0; c < n_chunks; ++c) {
    float v[16];
    ptx::tmem_ld16(t_row + c * 16, v);
    ptx::tmem_ld_wait();
    uint32_t packed[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float e = (c * 16 + j < a.N) ? exp2f(fmaf(v[j], LOG2E, -mxs)) : 0.f;
      if constexpr (EXPORT) {
        e *= inv;
        if (p_row && c * 16 + j < a.N) p_row[c * 16 + j] = e;
      } else {
        sum += e;
      }
      v[j] = e;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      packed[j] = *reinterpret_cast<uint32_t*>(&hh);
    }
    // P chunk c (16 keys = 8 packed columns) lands on columns this thread has already consumed
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t_row + c * 8),
                 "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3]), "r"(packed[4]), "r"(packed[5]),
                 "r"(packed[6]), "r"(packed[7])
                 : "memory");
  }
  if constexpr (!EXPORT) inv = 1.f / sum;
  if (a.lse_out && row < a.N) a.lse_out[((long long)b * a.H + h) * a.N + row] = mxs + log2f(sum);
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
    for (int k = 0; k < n_chunks; ++k) {
      const uint64_t dv = ptx::smem_desc_sw128(v_addr + k * 2048, 8192, 1024);
      ptx::mma_bf16_ts(tmem + O_COL, tmem + k * 8, dv, idesc2, k > 0 ? 1u : 0u);
    }
    ptx::mma_commit(bar_o);
  }

  // ---- epilogue: O row / sum -> bf16 ----
  ptx::mbar_wait(bar_o, 0);
  ptx::tc_fence_after();
  {
    __nv_bfloat16* o_ptr = reinterpret_cast<__nv_bfloat16*>(a.oh) + ((long long)b * a.N + row) * a.ld_oh + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      float v[16];
      ptx::tmem_ld16(t_row + O_COL + c * 16, v);
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
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, TMEM_COLS);
  }
}


// ================================================================================================
// Fused attention VJP (bf16 mode, head dim 64, N <= 256).  Per (image b, head h), with the
// log-sum-exp `lse2` of every row saved by the forward recomputation and
// delta_i = sum_d dO_id O_id  (== sum_j P_ij dP_ij):
//     S = q k^T,  P = exp2(S*log2e - lse2),  dP = dO v^T,  dS = P o (dP - delta)
//     dq = dS k,  dk = dS^T q,  dv = P^T dO
// One CTA (128 threads, thread = query row) per (b, h).  Keys are processed in chunks of 128
// (outer loop), queries in tiles of 128 (inner loop):
//   thread 0   TMA (Q and dO tiles once; K / V chunk per outer iteration), and all MMAs:
//              S and dP (SS) into TMEM [0,128) / [128,256);
//   all        S, dP rows from TMEM -> P, dS as bf16 into 128-byte-swizzled shared-memory tiles
//   thread 0   dQ_tile  = dS K      (A = dS tile K-major,  B = K chunk MN-major)   TMEM [384,448)
//              dK_chunk += dS^T Q   (A = dS tile MN-major, B = Q tile MN-major)    TMEM [256,320)
//              dV_chunk += P^T dO   (A = P tile MN-major,  B = dO tile MN-major)   TMEM [320,384)
//   all        dQ epilogue (thread = query row); with two key chunks the first chunk's partial
//              goes through an fp32 scratch row that the same thread re-reads on the second;
//              dK / dV epilogue (thread = key row) after the inner loop.
// The cotangent of an exported P (`attentions`, last evaluation only) is not handled here; that
// single evaluation takes the CUDA-core path (api.cu::eval_vjp).
struct AttnBwdArgs {
  int B, N, H, D, R;
  int n_kc, n_qt;
  const float* lse2;    // [B,H,N]
  const float* delta;   // [B,H,N]
  void* dz;             // [B*N, R] bf16: dq | dk | dv at columns h*64, D + h*64, 2D + h*64
  float* dq_scratch;    // [B*H, 256, 64] fp32 (used when n_kc == 2)
};

constexpr int BWD_TMEM_COLS = 512;
constexpr int T_S = 0, T_DP = 128, T_DK = 256, T_DV = 320, T_DQ = 384;

__device__ __forceinline__ void store_bf16x8_sw128(uint8_t* tile, int r, int key, const float* v) {
  // tile: atoms [128 rows x 64 keys] of 16 KB; 16-byte chunk index XOR (row % 8)
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&hh);
  }
  const int atom = key >> 6, chunk = (key & 63) >> 3;
  uint4* dst = reinterpret_cast<uint4*>(tile + atom * 16384 + r * 128 + ((chunk ^ (r & 7)) << 4));
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(128, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // 2 x 16 KB
  uint8_t* sDO = sQ + 2 * 16384;      // 2 x 16 KB
  uint8_t* sK = sDO + 2 * 16384;      // 16 KB
  uint8_t* sV = sK + 16384;           // 16 KB
  uint8_t* sP = sV + 16384;           // 32 KB (2 atoms)
  uint8_t* sDS = sP + 32768;          // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 32768);
  uint64_t* bar_qdo = bars;
  uint64_t* bar_kv = bars + 1;
  uint64_t* bar_sdp = bars + 2;
  uint64_t* bar_mma2 = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5;
  const int t = threadIdx.x;
  const int h = blockIdx.x % a.H;
  const int b = blockIdx.x / a.H;

  if (t == 0) {
    ptx::prefetch_tensormap(&tmQKV);
    ptx::prefetch_tensormap(&tmDO);
    ptx::mbar_init(bar_qdo, 1);
    ptx::mbar_init(bar_kv, 1);
    ptx::mbar_init(bar_sdp, 1);
    ptx::mbar_init(bar_mma2, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, BWD_TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);

  if (t == 0) {
    ptx::mbar_expect_tx(bar_qdo, a.n_qt * 2 * 16384);
    for (int qt = 0; qt < a.n_qt; ++qt) {
      ptx::tma_load_3d(sQ + qt * 16384, &tmQKV, bar_qdo, h * HD, qt * 128, b);
      ptx::tma_load_3d(sDO + qt * 16384, &tmDO, bar_qdo, h * HD, qt * 128, b);
    }
  }
  constexpr float LOG2E = 1.4426950408889634f;
  const int NP = (a.N + 15) / 16 * 16;
  uint32_t ph_sdp = 0, ph_mma2 = 0;
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(a.dz);

  for (int kc = 0; kc < a.n_kc; ++kc) {
    const int cw = min(128, NP - kc * 128);  // chunk width (multiple of 16)
    if (t == 0) {
      ptx::mbar_expect_tx(bar_kv, 2 * 16384);
      ptx::tma_load_3d(sK, &tmQKV, bar_kv, a.D + h * HD, kc * 128, b);
      ptx::tma_load_3d(sV, &tmQKV, bar_kv, 2 * a.D + h * HD, kc * 128, b);
    }
    for (int qt = 0; qt < a.n_qt; ++qt) {
      if (t == 0) {
        if (kc == 0 && qt == 0) ptx::mbar_wait(bar_qdo, 0);
        if (qt == 0) ptx::mbar_wait(bar_kv, kc & 1);
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::idesc_bf16(128, cw, 0, 0);
        const uint32_t q_addr = ptx::smem_u32(sQ + qt * 16384), do_addr = ptx::smem_u32(sDO + qt * 16384);
        const uint32_t k_addr = ptx::smem_u32(sK), v_addr = ptx::smem_u32(sV);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::mma_bf16_ss(tmem + T_S, ptx::smem_desc_sw128(q_addr + k * 32, 16, 1024),
                           ptx::smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::mma_bf16_ss(tmem + T_DP, ptx::smem_desc_sw128(do_addr + k * 32, 16, 1024),
                           ptx::smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
        ptx::mma_commit(bar_sdp);
      }
      // ---- this thread's query row: P and dS of the chunk -> shared memory ----
      const int qrow = qt * 128 + t;
      float lse = INFINITY, dlt = 0.f;
      if (qrow < a.N) {
        const long long si = ((long long)b * a.H + h) * a.N + qrow;
        lse = a.lse2[si];
        dlt = a.delta[si];
      }
      ptx::mbar_wait(bar_sdp, ph_sdp);
      ph_sdp ^= 1;
      ptx::tc_fence_after();
      for (int c = 0; c < cw / 16; ++c) {
        float sv[16], dp[16];
        ptx::tmem_ld16(t_lane + T_S + c * 16, sv);
        ptx::tmem_ld16(t_lane + T_DP + c * 16, dp);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int key = kc * 128 + c * 16 + j;
          const float p = (key < a.N) ? exp2f(fmaf(sv[j], LOG2E, -lse)) : 0.f;
          sv[j] = p;
          dp[j] = p * (dp[j] - dlt);
        }
        store_bf16x8_sw128(sP, t, c * 16, sv);
        store_bf16x8_sw128(sP, t, c * 16 + 8, sv + 8);
        store_bf16x8_sw128(sDS, t, c * 16, dp);
        store_bf16x8_sw128(sDS, t, c * 16 + 8, dp + 8);
      }
      ptx::fence_async_shared();
      ptx::tc_fence_before();
      __syncthreads();
      if (t == 0) {
        ptx::tc_fence_after();
        const uint32_t ds_addr = ptx::smem_u32(sDS), p_addr = ptx::smem_u32(sP);
        const uint32_t q_addr = ptx::smem_u32(sQ + qt * 16384), do_addr = ptx::smem_u32(sDO + qt * 16384);
        const uint32_t k_addr = ptx::smem_u32(sK);
        // dQ = dS K : contraction over the chunk's keys
        const uint32_t id_q = ptx::idesc_bf16(128, HD, 0, 1);
        for (int ks = 0; ks < cw / 16; ++ks)
          ptx::mma_bf16_ss(tmem + T_DQ,
                           ptx::smem_desc_sw128(ds_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                           ptx::smem_desc_sw128(k_addr + ks * 2048, 8192, 1024), id_q, ks > 0 ? 1u : 0u);
        // dK += dS^T Q, dV += P^T dO : contraction over the tile's 128 query rows
        const uint32_t id_kv = ptx::idesc_bf16(128, HD, 1, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t acc = (qt > 0 || ks > 0) ? 1u : 0u;
          ptx::mma_bf16_ss(tmem + T_DK, ptx::smem_desc_sw128(ds_addr + ks * 2048, 16384, 1024),
                           ptx::smem_desc_sw128(q_addr + ks * 2048, 8192, 1024), id_kv, acc);
          ptx::mma_bf16_ss(tmem + T_DV, ptx::smem_desc_sw128(p_addr + ks * 2048, 16384, 1024),
                           ptx::smem_desc_sw128(do_addr + ks * 2048, 8192, 1024), id_kv, acc);
        }
        ptx::mma_commit(bar_mma2);
      }
      ptx::mbar_wait(bar_mma2, ph_mma2);
      ph_mma2 ^= 1;
      ptx::tc_fence_after();
      // ---- dQ epilogue (thread = query row) ----
      {
        float* scr = a.dq_scratch ? a.dq_scratch + ((long long)blockIdx.x * 256 + qrow) * HD : nullptr;
        const bool last = (kc == a.n_kc - 1);
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
          float v[16];
          ptx::tmem_ld16(t_lane + T_DQ + c * 16, v);
          ptx::tmem_ld_wait();
          if (qrow < a.N) {
            if (kc > 0) {
              const float4* sp = reinterpret_cast<const float4*>(scr + c * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 x = sp[i];
                v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
              }
            }
            if (last) {
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&hh);
              }
              uint4* o = reinterpret_cast<uint4*>(dz + ((long long)b * a.N + qrow) * a.R + h * HD + c * 16);
              o[0] = make_uint4(w[0], w[1], w[2], w[3]);
              o[1] = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
              float4* sp = reinterpret_cast<float4*>(scr + c * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) sp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncthreads();
    }
    // ---- dK / dV epilogue of this key chunk (thread = key row) ----
    {
      const int key = kc * 128 + t;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* dst = dz + ((long long)b * a.N + key) * a.R + (which + 1) * a.D + h * HD;
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
          float v[16];
          ptx::tmem_ld16(t_lane + (which ? T_DV : T_DK) + c * 16, v);
          ptx::tmem_ld_wait();
          if (key < a.N) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              w[j] = *reinterpret_cast<uint32_t*>(&hh);
            }
            uint4* o = reinterpret_cast<uint4*>(dst + c * 16);
            o[0] = make_uint4(w[0], w[1], w[2], w[3]);
            o[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, BWD_TMEM_COLS);
  }
}

// delta[b,h,i] = sum_d dO[i, h*64+d] * O[i, h*64+d]    (one warp per token row, all heads)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, long long ld_do,
                                                         const __nv_bfloat16* __restrict__ O, long long ld_o,
                                                         float* __restrict__ delta, int B, int N, int H) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * N) return;
  const int b = row / N, i = row - b * N;
  for (int h = 0; h < H; ++h) {
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(dO + (long long)row * ld_do + h * HD + lane * 2);
    const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(O + (long long)row * ld_o + h * HD + lane * 2);
    float s = __low2float(x) * __low2float(y) + __high2float(x) * __high2float(y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) delta[((long long)b * H + h) * N + i] = s;
  }
}

}  // namespace

bool attn_fwd_tc_supports(int N, int D, int H, int act_type, long long ld_oh) {
  return act_type == DT_BF16 && D % H == 0 && D / H == HD && N >= 1 && N <= 256 && D % 8 == 0 && ld_oh % 8 == 0;
}

// qkv: [B, N, 3D] bf16 (q | k | v per row); oh: [B*N, ld_oh] bf16; p_out: [B,H,N,N] fp32 or null.
int attn_fwd_tc(const void* qkv, void* oh, long long ld_oh, float* p_out, float* lse_out, int B, int N, int H, int D,
                cudaStream_t s) {
  if (!attn_fwd_tc_supports(N, D, H, DT_BF16, ld_oh))
    return set_error(ODEVIT_ERR_UNSUPPORTED, "attn_fwd_tc: unsupported shape N=%d D=%d H=%d", N, D, H);
  ProfScope prof(KC_FUSED_ATTN, s);
  AttnArgs a;
  a.B = B; a.N = N; a.H = H; a.D = D;
  a.NP = (N + 15) / 16 * 16;
  a.tiles_m = (N + BMQ - 1) / BMQ;
  a.oh = oh; a.ld_oh = ld_oh; a.p_out = p_out; a.lse_out = lse_out;
  CUtensorMap tq, tkv;
  ODV_TRY(make_tmap_3d_bf16(&tq, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, BMQ, 1));
  ODV_TRY(make_tmap_3d_bf16(&tkv, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, a.NP, 1));
  const int smem = BMQ * HD * 2 + 2 * a.NP * HD * 2 + 1024 + 64;
  const int grid = B * H * a.tiles_m;
  static bool configured = false;
  if (!configured) {
    const int max_smem = BMQ * HD * 2 + 2 * 256 * HD * 2 + 1024 + 64;
    ODV_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    ODV_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  if (p_out) attn_fwd_tc_kernel<true><<<grid, 128, smem, s>>>(tq, tkv, a);
  else attn_fwd_tc_kernel<false><<<grid, 128, smem, s>>>(tq, tkv, a);
  ODV_LAUNCH_CHECK();
  return 0;
}

size_t attn_bwd_tc_scratch_floats(int B, int N, int H) { return N > 128 ? (size_t)B * H * 256 * HD : 0; }

// qkv [B,N,3D] bf16; dO [B*N, D] bf16; O = oh[:, 0:D] (ld_oh); lse2, delta [B,H,N] fp32 (delta is
// written here); dz [B*N, R] bf16 receives dq | dk | dv.
int attn_bwd_tc(const void* qkv, const void* dO, const void* oh, long long ld_oh, const float* lse2, float* delta,
                void* dz, int R, float* dq_scratch, int B, int N, int H, int D, cudaStream_t s) {
  if (!attn_fwd_tc_supports(N, D, H, DT_BF16, ld_oh) || R % 8)
    return set_error(ODEVIT_ERR_UNSUPPORTED, "attn_bwd_tc: unsupported shape N=%d D=%d H=%d", N, D, H);
  ProfScope prof(KC_FUSED_ATTN_BWD, s);
  attn_delta_kernel<<<(B * N + 7) / 8, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dO), D,
                                                     reinterpret_cast<const __nv_bfloat16*>(oh), ld_oh, delta, B, N, H);
  ODV_LAUNCH_CHECK();
  AttnBwdArgs a;
  a.B = B; a.N = N; a.H = H; a.D = D; a.R = R;
  a.n_kc = (N + 127) / 128; a.n_qt = (N + 127) / 128;
  a.lse2 = lse2; a.delta = delta; a.dz = dz; a.dq_scratch = dq_scratch;
  if (a.n_kc > 1 && !dq_scratch) return set_error(ODEVIT_ERR_WORKSPACE, "attn_bwd_tc: scratch missing");
  CUtensorMap tqkv, tdo;
  ODV_TRY(make_tmap_3d_bf16(&tqkv, qkv, 3 * D, N, B, 3 * D, (uint64_t)N * 3 * D, HD, 128, 1));
  ODV_TRY(make_tmap_3d_bf16(&tdo, dO, D, N, B, D, (uint64_t)N * D, HD, 128, 1));
  const int smem = 10 * 16384 + 1024 + 64;
  static bool configured = false;
  if (!configured) {
    ODV_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attn_bwd_tc_kernel<<<B * H, 128, smem, s>>>(tqkv, tdo, a);
  ODV_LAUNCH_CHECK();
  return 0;
}

}  // namespace odevit
