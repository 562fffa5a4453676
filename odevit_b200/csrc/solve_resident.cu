// The on-chip-state solver for small-token shapes (CIFAR: N ~ 65-69 tokens, D = 64..256, head dim 64):
// ONE persistent CTA per image runs EVERY solver step of odeint(odefunc, x0, t, method)
// (ode_transformer_gpt.py:571-578 with the parallel field :274-277, :317-330), inference only.
//
// What stays on the chip for the whole solve of an image:
//   shared memory   the ODE state y [N, D] fp32 (row-padded), the centred stage input xc as the bf16
//                   A-operand tile, per-head q/k/v/O tiles, the GELU(fc1) chunk tile, a 4-stage ring of
//                   weight tiles
//   tensor memory   R0 [0,192): q|k|v accumulator of a head -> S -> P (bf16 in place) -> O; then the fc1
//                   chunk accumulator;  R1 [192,192+D): the field output accumulator, summed over all heads'
//                   out-projections and all fc2 chunks; it doubles as scratch for the next stage input
// HBM / L2 sees only: the weights (streamed by TMA, 2.25 * (3D+hid) * D... bytes per evaluation, L2 hits
// after the first CTA), x0 once, each trajectory row once (if requested), the last attention map (if
// requested), and -- for multi-stage methods -- the k_l stage vectors in a per-CTA L2-resident scratch.
//
//   warp 4 (elected lane)  TMA weight tiles + every tcgen05.mma
//   warps 0-3              thread = token row: bias / softmax / GELU / Runge-Kutta update / centring
// Per evaluation and head h:  [q|k|v]_h = xc W^T (3 MMAs N=64) -> tiles -> S = q k^T -> softmax -> O = P v (TS)
// -> OUT += O Wo_h^T;  per 128-wide hidden chunk c:  Hc = xc W1_c^T -> GELU -> OUT += Hc W2_c^T;  then
// k = scaler (OUT + b2), the stage combine of the Butcher tableau on the resident state, and the next xc.
#include <cuda.h>

#include <cstdlib>

#include "epilogue.cuh"
#include "host.h"
#include "ptx.cuh"

namespace odevit {

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer);

namespace {

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

constexpr int RES_THREADS = 160;
constexpr int NST = 4;          // weight-ring stages
constexpr int T_R0 = 0, T_O = 128, T_R1 = 192;
constexpr int kMaxGrid = 129;   // grid points carried in the kernel parameters

struct ResArgs {
  int B, N, RA, NK, D, H, hid;
  int n_grid, S;
  float scaler;
  float dt[kMaxGrid - 1];
  float a[4][4], b[4];
  const float* x0;       // [B, N, D]
  float* states;         // [T, B, N, D] or null (row 0 written by the host wrapper)
  float* final_state;    // [B, N, D] or null
  float* p_last;         // [B, H, N, N] or null
  const float* b1cat;    // [3D + hid]
  const float* b2;       // [D]
  float* kbuf;           // [grid][3][D][128] fp32 scratch (stage vectors of multi-stage methods)
};

struct Smem {
  int xc_atom;           // bytes of one 64-column atom of RA rows
  int off_xc, off_q, off_k, off_v, off_o, off_ring, off_y, off_bars, total;
  int stage_bytes;
};
__host__ __device__ inline Smem smem_layout(int RA, int N, int D) {
  Smem s;
  s.xc_atom = RA * 128;
  s.stage_bytes = D * 128;
  int o = 0;
  s.off_xc = o; o += (D / 64) * s.xc_atom;
  s.off_q = o; o += s.xc_atom;      // q, k: also the two atoms of the GELU(fc1) chunk tile
  s.off_k = o; o += s.xc_atom;
  s.off_v = o; o += s.xc_atom;
  s.off_o = o; o += s.xc_atom;
  o = (o + 1023) & ~1023;
  s.off_ring = o; o += NST * s.stage_bytes;
  s.off_y = o; o += N * (D + 1) * 4;
  o = (o + 15) & ~15;
  s.off_bars = o; o += 128;
  s.total = o;
  return s;
}

// bf16 row of a K-major SWIZZLE_128B tile: 8 consecutive columns [col, col+8) of row r
__device__ __forceinline__ void st_tile8(uint8_t* atom0, int atom_bytes, int r, int col, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&hh);
  }
  const int atom = col >> 6, chunk = (col & 63) >> 3;
  *reinterpret_cast<uint4*>(atom0 + atom * atom_bytes + r * 128 + ((chunk ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float* v) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
  ptx::tmem_st16(taddr, r);
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(RES_THREADS, 1)
solve_resident_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                      const __grid_constant__ ResArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Smem L = smem_layout(a.RA, a.N, a.D);
  uint8_t* sXC = smem + L.off_xc;
  uint8_t* sQ = smem + L.off_q;
  uint8_t* sK = smem + L.off_k;
  uint8_t* sV = smem + L.off_v;
  uint8_t* sO = smem + L.off_o;
  uint8_t* sRing = smem + L.off_ring;
  float* sY = reinterpret_cast<float*>(smem + L.off_y);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bars);
  uint64_t* bar_full = bars;          // [NST] weight tile landed
  uint64_t* bar_free = bars + NST;    // [NST] the MMAs that read the stage have retired
  uint64_t* bar_mma = bars + 2 * NST;       // an MMA group the compute warps wait for has retired
  uint64_t* bar_cmp = bars + 2 * NST + 1;   // a compute phase is finished (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2);

  const int warp = threadIdx.x >> 5;
  const int D = a.D, H = a.H, N = a.N, RA = a.RA, NK = a.NK;
  const int n_chunks = a.hid / 128;
  const int KD = D / 16;                       // k-steps over the model dimension
  const int tiles_per_eval = 4 * H + 4 * n_chunks;
  const int n_evals = (a.n_grid - 1) * a.S;
  const int n_mine = (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(bar_full + i, 1); ptx::mbar_init(bar_free + i, 1); }
    ptx::mbar_init(bar_mma, 1);
    ptx::mbar_init(bar_cmp, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 4) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ======================= producer: weight tiles + every MMA =======================
    const bool leader = ptx::elect_one();
    if (leader) { ptx::prefetch_tensormap(&tmW1); ptx::prefetch_tensormap(&tmW2); }
    long long loaded = 0, consumed = 0;       // weight tiles issued to TMA / handed to the tensor core
    const long long total_tiles = (long long)n_mine * n_evals * tiles_per_eval;
    uint32_t ph_cmp = 0;
    // tile t of an evaluation:  head h: 4h+{0,1,2} = Wq_h, Wk_h, Wv_h rows (A-type: 64 rows x D), 4h+3 = Wo k-atom h
    // (B-type: D rows x 64);  chunk c: base+4c+{0,1} = fc1 rows (A-type), base+4c+{2,3} = fc2 k-atoms (B-type)
    auto issue_load = [&](long long gt) {
      const int t = (int)(gt % tiles_per_eval);
      const int st = (int)(gt % NST);
      if (gt >= NST) ptx::mbar_wait(bar_free + st, (uint32_t)((gt / NST - 1) & 1));
      uint8_t* dst = sRing + st * L.stage_bytes;
      if (leader) {
        ptx::mbar_expect_tx(bar_full + st, L.stage_bytes);
        if (t < 4 * H) {
          const int h = t >> 2, m = t & 3;
          if (m < 3) {
            for (int ka = 0; ka < D / 64; ++ka) ptx::tma_load_2d(dst + ka * 8192, &tmW1, bar_full + st, ka * 64, m * D + h * 64);
          } else {
            ptx::tma_load_2d(dst, &tmW2, bar_full + st, h * 64, 0);
          }
        } else {
          const int u = t - 4 * H, c = u >> 2, m = u & 3;
          if (m < 2) {
            for (int ka = 0; ka < D / 64; ++ka)
              ptx::tma_load_2d(dst + ka * 8192, &tmW1, bar_full + st, ka * 64, 3 * D + c * 128 + m * 64);
          } else {
            ptx::tma_load_2d(dst, &tmW2, bar_full + st, D + c * 128 + (m - 2) * 64, 0);
          }
        }
      }
      __syncwarp();
    };
    auto top_up = [&]() {
      while (loaded < total_tiles && loaded < consumed + NST) { issue_load(loaded); ++loaded; }
    };
    // waits for tile `consumed`, returns its shared-memory address; release() after its MMAs were issued
    auto acquire = [&]() -> uint32_t {
      top_up();
      const int st = (int)(consumed % NST);
      ptx::mbar_wait(bar_full + st, (uint32_t)((consumed / NST) & 1));
      ptx::tc_fence_after();
      return ptx::smem_u32(sRing + st * L.stage_bytes);
    };
    auto release = [&]() {
      const int st = (int)(consumed % NST);
      if (leader) ptx::mma_commit(bar_free + st);
      __syncwarp();
      ++consumed;
    };
    auto wait_cmp = [&]() {
      ptx::mbar_wait(bar_cmp, ph_cmp);
      ph_cmp ^= 1;
      ptx::tc_fence_after();
    };
    auto commit_mma = [&]() {
      if (leader) ptx::mma_commit(bar_mma);
      __syncwarp();
    };
    const uint32_t xc_addr = ptx::smem_u32(sXC), q_addr = ptx::smem_u32(sQ), k_addr = ptx::smem_u32(sK);
    const uint32_t v_addr = ptx::smem_u32(sV), o_addr = ptx::smem_u32(sO);
    const uint32_t id_n64 = ptx::idesc_bf16(128, 64, 0, 0);
    const uint32_t id_s = ptx::idesc_bf16(128, NK, 0, 0);
    const uint32_t id_pv = ptx::idesc_bf16(128, 64, 0, 1);
    const uint32_t id_out = ptx::idesc_bf16(128, D, 0, 0);
    // D[128 x 64] (+)= xc[128 x D] * tile[64 x D]^T
    auto mma_xc_tile = [&](uint32_t d_tmem, uint32_t tile) {
      if (leader) {
        for (int kk = 0; kk < KD; ++kk) {
          const uint64_t da = ptx::smem_desc_sw128(xc_addr + (kk >> 2) * L.xc_atom + (kk & 3) * 32, 16, 1024);
          const uint64_t db = ptx::smem_desc_sw128(tile + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
          ptx::mma_bf16_ss(d_tmem, da, db, id_n64, kk > 0 ? 1u : 0u);
        }
      }
      __syncwarp();
    };

    for (int idx = 0; idx < n_mine; ++idx) {
      for (int ev = 0; ev < n_evals; ++ev) {
        wait_cmp();                                   // xc of this evaluation is in shared memory, R1 has been read
        for (int h = 0; h < H; ++h) {
          for (int m = 0; m < 3; ++m) {               // q | k | v of head h
            const uint32_t tile = acquire();
            mma_xc_tile(tmem + T_R0 + m * 64, tile);
            release();
          }
          commit_mma();
          wait_cmp();                                 // q, k, v tiles written
          if (leader) {
            for (int kk = 0; kk < 4; ++kk)
              ptx::mma_bf16_ss(tmem + T_R0, ptx::smem_desc_sw128(q_addr + kk * 32, 16, 1024),
                               ptx::smem_desc_sw128(k_addr + kk * 32, 16, 1024), id_s, kk > 0 ? 1u : 0u);
          }
          __syncwarp();
          commit_mma();
          wait_cmp();                                 // P (bf16) is in tensor memory
          if (leader) {
            for (int ks = 0; ks < NK / 16; ++ks)
              ptx::mma_bf16_ts(tmem + T_O, tmem + T_R0 + ks * 8, ptx::smem_desc_sw128(v_addr + ks * 2048, 8192, 1024),
                               id_pv, ks > 0 ? 1u : 0u);
          }
          __syncwarp();
          commit_mma();
          wait_cmp();                                 // O tile written
          {
            const uint32_t tile = acquire();          // Wo k-atom h: [D rows x 64]
            if (leader) {
              for (int kk = 0; kk < 4; ++kk)
                ptx::mma_bf16_ss(tmem + T_R1, ptx::smem_desc_sw128(o_addr + kk * 32, 16, 1024),
                                 ptx::smem_desc_sw128(tile + kk * 32, 16, 1024), id_out, (h > 0 || kk > 0) ? 1u : 0u);
            }
            __syncwarp();
            release();
          }
        }
        for (int c = 0; c < n_chunks; ++c) {
          for (int m = 0; m < 2; ++m) {               // fc1 rows [c*128 + m*64, +64)
            const uint32_t tile = acquire();
            mma_xc_tile(tmem + T_R0 + m * 64, tile);
            release();
          }
          commit_mma();
          wait_cmp();                                 // GELU(fc1) chunk tile written (over the q / k tiles)
          for (int m = 0; m < 2; ++m) {               // fc2 k-atoms
            const uint32_t tile = acquire();
            if (leader) {
              for (int kk = 0; kk < 4; ++kk)
                ptx::mma_bf16_ss(tmem + T_R1, ptx::smem_desc_sw128(q_addr + m * L.xc_atom + kk * 32, 16, 1024),
                                 ptx::smem_desc_sw128(tile + kk * 32, 16, 1024), id_out, 1u);
            }
            __syncwarp();
            release();
          }
        }
        commit_mma();                                 // the field output of this evaluation is complete in R1
      }
      wait_cmp();                                     // last stage combine of the image done
    }
  } else {
    // ======================= compute warps: thread = token row =======================
    const int r = threadIdx.x;                         // 0..127
    const bool row_ok = r < N;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t ph_mma = 0;
    auto wait_mma = [&]() {
      ptx::mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      ptx::tc_fence_after();
    };
    auto done = [&]() {   // shared-memory writes -> async proxy, tensor-memory traffic ordered, then signal
      ptx::fence_async_shared();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_cmp);
    };
    float* yrow = sY + (size_t)(row_ok ? r : 0) * (D + 1);
    float* kb = a.kbuf + (size_t)blockIdx.x * 3 * D * 128;
    constexpr float LOG2E = 1.4426950408889634f;

    // writes the centred bf16 row of `u` (held in R1 as fp32) into the xc tile; `sum` = row sum of u
    auto centre_from_r1 = [&](float sum) {
      const float mean = sum / (float)D;
      for (int c = 0; c < D / 16; ++c) {
        float v[16];
        ptx::tmem_ld16(t_lane + T_R1 + c * 16, v);
        ptx::tmem_ld_wait();
        if (r < RA) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = row_ok ? v[j] - mean : 0.f;
          st_tile8(sXC, L.xc_atom, r, c * 16, v);
          st_tile8(sXC, L.xc_atom, r, c * 16 + 8, v + 8);
        }
      }
    };

    for (int idx = 0; idx < n_mine; ++idx) {
      const int img = (int)blockIdx.x + idx * (int)gridDim.x;
      // ---- load x0 -> resident state, first xc ----
      {
        float sum = 0.f;
        const float* src = a.x0 + ((size_t)img * N + (row_ok ? r : 0)) * D;
        for (int c = 0; c < D / 16; ++c) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 f = row_ok ? *reinterpret_cast<const float4*>(src + c * 16 + j * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum += v[j];
            if (row_ok) yrow[c * 16 + j] = v[j];
          }
          tmem_st16f(t_lane + T_R1 + c * 16, v);
        }
        ptx::tmem_st_wait();
        centre_from_r1(sum);
        done();
      }
      for (int ev = 0; ev < n_evals; ++ev) {
        const int step = ev / a.S, st = ev - step * a.S;
        const bool last_eval = (ev == n_evals - 1);
        for (int h = 0; h < H; ++h) {
          // ---- q | k | v accumulators -> bias -> bf16 tiles ----
          wait_mma();
          for (int m = 0; m < 3; ++m) {
            uint8_t* dst = (m == 0) ? sQ : (m == 1) ? sK : sV;
            const float* bias = a.b1cat + m * D + h * 64;
            for (int c = 0; c < 4; ++c) {
              float v[16];
              ptx::tmem_ld16(t_lane + T_R0 + m * 64 + c * 16, v);
              ptx::tmem_ld_wait();
              if (r < RA) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = row_ok ? v[j] + __ldg(bias + c * 16 + j) : 0.f;
                st_tile8(dst, L.xc_atom, r, c * 16, v);
                st_tile8(dst, L.xc_atom, r, c * 16 + 8, v + 8);
              }
            }
          }
          done();
          // ---- softmax over the keys of this row, P packed in place ----
          wait_mma();
          {
            float v[8][16];
            const int nch = NK / 16;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < nch) ptx::tmem_ld16(t_lane + T_R0 + c * 16, v[c]);
            ptx::tmem_ld_wait();
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < nch) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (c * 16 + j >= N) v[c][j] = -INFINITY;
                  mx = fmaxf(mx, v[c][j]);
                }
              }
            const float mxs = mx * LOG2E;
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < nch) {
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  v[c][j] = ex2f(fmaf(v[c][j], LOG2E, -mxs));
                  sum += v[c][j];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  __nv_bfloat162 hh = __floats2bfloat162_rn(v[c][2 * j], v[c][2 * j + 1]);
                  packed[j] = *reinterpret_cast<uint32_t*>(&hh);
                }
                tmem_st8u(t_lane + T_R0 + c * 8, packed);
              }
            const float inv = 1.f / sum;
            // the O epilogue needs 1/sum: park it in the state row's padding column
            if (row_ok) yrow[D] = inv;
            if (last_eval && a.p_last && row_ok) {
              float* p_row = a.p_last + (((size_t)img * H + h) * N + r) * N;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                if (c < nch) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c * 16 + j < N) p_row[c * 16 + j] = v[c][j] * inv;
                }
            }
            ptx::tmem_st_wait();
          }
          done();
          // ---- O row * 1/sum -> bf16 tile ----
          wait_mma();
          {
            const float inv = row_ok ? yrow[D] : 0.f;
            for (int c = 0; c < 4; ++c) {
              float v[16];
              ptx::tmem_ld16(t_lane + T_O + c * 16, v);
              ptx::tmem_ld_wait();
              if (r < RA) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = row_ok ? v[j] * inv : 0.f;
                st_tile8(sO, L.xc_atom, r, c * 16, v);
                st_tile8(sO, L.xc_atom, r, c * 16 + 8, v + 8);
              }
            }
          }
          done();
        }
        for (int ch = 0; ch < n_chunks; ++ch) {
          // ---- fc1 chunk -> bias -> GELU -> bf16 tile (two atoms over the q / k tiles) ----
          wait_mma();
          const float* bias = a.b1cat + 3 * D + ch * 128;
          for (int c = 0; c < 8; ++c) {
            float v[16];
            ptx::tmem_ld16(t_lane + T_R0 + c * 16, v);
            ptx::tmem_ld_wait();
            if (r < RA) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = row_ok ? gelu_fast(v[j] + __ldg(bias + c * 16 + j)) : 0.f;
              st_tile8(sQ, L.xc_atom, r, c * 16, v);
              st_tile8(sQ, L.xc_atom, r, c * 16 + 8, v + 8);
            }
          }
          done();
        }
        // ---- k = scaler (OUT + b2); stage combine on the resident state; next stage input -> xc ----
        wait_mma();
        {
          const float dt = a.dt[step];
          const bool last_stage = (st == a.S - 1);
          const float* coef = last_stage ? a.b : a.a[st + 1];
          float sum = 0.f;
          float* srow = (last_stage && a.states && row_ok) ? a.states + (((size_t)(step + 1) * a.B + img) * N + r) * D : nullptr;
          float* frow = (last_eval && a.final_state && row_ok) ? a.final_state + ((size_t)img * N + r) * D : nullptr;
          for (int c = 0; c < D / 16; ++c) {
            float v[16];
            ptx::tmem_ld16(t_lane + T_R1 + c * 16, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = c * 16 + j;
              const float kv = a.scaler * (v[j] + __ldg(a.b2 + col));
              if (!last_stage) kb[((size_t)st * D + col) * 128 + r] = kv;   // a later stage combines it
              float acc = coef[st] * kv;
              for (int l = 0; l < st; ++l)
                if (coef[l] != 0.f) acc = fmaf(coef[l], kb[((size_t)l * D + col) * 128 + r], acc);
              const float yv = row_ok ? yrow[col] : 0.f;
              const float u = fmaf(dt, acc, yv);
              if (last_stage && row_ok) yrow[col] = u;
              v[j] = u;
              sum += u;
            }
            if (srow) {
#pragma unroll
              for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(srow + c * 16)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (frow) {
#pragma unroll
              for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(frow + c * 16)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            tmem_st16f(t_lane + T_R1 + c * 16, v);
          }
          ptx::tmem_st_wait();
          centre_from_r1(sum);
        }
        done();
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

size_t solve_resident_scratch_floats(const Plan& p) {
  const int sms = num_sms();
  const int grid = p.B < sms ? p.B : sms;
  return (size_t)grid * 3 * p.D * 128;
}

bool solve_resident_shape_ok(const Plan& p) {
  if (p.variant != ODEVIT_FIELD_PARALLEL || p.precision != ODEVIT_BF16) return false;
  if (p.d != 64 || p.D % 64 || p.D > 256 || p.hid % 128 || p.N > 128) return false;
  const int RA = (p.N + 15) / 16 * 16;
  return smem_layout(RA, p.N, p.D).total <= 227 * 1024;
}

bool solve_resident_supports(const Plan& p, int n_grid, bool wants_p_traj, bool has_tape) {
  const char* env = getenv("ODEVIT_RESIDENT");
  if (env && env[0] == '0') return false;
  if (wants_p_traj || has_tape || p.split_out || p.p_attn > 0.f) return false;
  if (n_grid < 2 || n_grid > kMaxGrid) return false;
  return solve_resident_shape_ok(p);
}

int solve_resident(const Plan& p, const WeightBufs& wb, int S, const float (*ta)[4], const float* tbv, const float* x0,
                   const float* t_host, int n_grid, float* states, float* final_state, float* p_last, float* kbuf,
                   cudaStream_t s) {
  ProfScope prof(KC_RESIDENT, s);
  ResArgs a;
  a.B = p.B; a.N = p.N; a.D = p.D; a.H = p.H; a.hid = p.hid;
  a.RA = (p.N + 15) / 16 * 16;
  a.NK = a.RA;
  a.n_grid = n_grid; a.S = S; a.scaler = p.scaler;
  for (int j = 0; j + 1 < n_grid; ++j) a.dt[j] = t_host[j + 1] - t_host[j];
  for (int i = 0; i < 4; ++i) {
    a.b[i] = tbv[i];
    for (int j = 0; j < 4; ++j) a.a[i][j] = ta[i][j];
  }
  a.x0 = x0; a.states = states; a.final_state = final_state; a.p_last = p_last;
  a.b1cat = wb.b1cat; a.b2 = wb.b2; a.kbuf = kbuf;
  const int R = 3 * p.D + p.hid, K2 = p.D + p.hid;
  CUtensorMap t1, t2;
  ODV_TRY(make_tmap_2d_bf16(&t1, wb.w1cat, p.D, R, p.D, 64, 64));
  ODV_TRY(make_tmap_2d_bf16(&t2, wb.w2cat, K2, p.D, K2, 64, p.D));
  const Smem L = smem_layout(a.RA, a.N, a.D);
  static int configured_bytes = 0;
  if (L.total > configured_bytes) {
    ODV_CUDA(cudaFuncSetAttribute(solve_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    configured_bytes = L.total;
  }
  const int sms = num_sms();
  const int grid = p.B < sms ? p.B : sms;
  solve_resident_kernel<<<grid, RES_THREADS, L.total, s>>>(t1, t2, a);
  ODV_LAUNCH_CHECK();
  return 0;
}

}  // namespace odevit
