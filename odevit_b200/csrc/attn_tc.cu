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

}  // namespace

bool attn_fwd_tc_supports(int N, int D, int H, int act_type, long long ld_oh) {
  return act_type == DT_BF16 && D % H == 0 && D / H == HD && N >= 1 && N <= 256 && D % 8 == 0 && ld_oh % 8 == 0;
}

// qkv: [B, N, 3D] bf16 (q | k | v per row); oh: [B*N, ld_oh] bf16; p_out: [B,H,N,N] fp32 or null.
int attn_fwd_tc(const void* qkv, void* oh, long long ld_oh, float* p_out, int B, int N, int H, int D, cudaStream_t s) {
  if (!attn_fwd_tc_supports(N, D, H, DT_BF16, ld_oh))
    return set_error(ODEVIT_ERR_UNSUPPORTED, "attn_fwd_tc: unsupported shape N=%d D=%d H=%d", N, D, H);
  ProfScope prof(KC_FUSED_ATTN, s);
  AttnArgs a;
  a.B = B; a.N = N; a.H = H; a.D = D;
  a.NP = (N + 15) / 16 * 16;
  a.tiles_m = (N + BMQ - 1) / BMQ;
  a.oh = oh; a.ld_oh = ld_oh; a.p_out = p_out;
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

}  // namespace odevit
