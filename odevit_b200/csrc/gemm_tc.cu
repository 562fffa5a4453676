// tcgen05 / TMA GEMM with fused epilogues -- the dense contractions of the bf16 mode.
//
//   C[m,n] = sum_k A(m,k) * B(n,k)       bf16 operands, fp32 accumulation in tensor memory
//
// replaces aten::mm of models/ode_transformer_gpt.py:193-200 (fc1/fc2) and :228 (packed in-proj,
// out_proj), and their autograd counterparts, with the CenterNorm affine / GELU / `*scaler` /
// Runge-Kutta stage combine fused into the epilogues (epilogue.cuh).
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor 128-byte-swizzled boxes of A and B into a
//            STAGES-deep shared-memory ring, completion on `full` mbarriers
//   warp 1   allocates tensor memory; one elected thread issues tcgen05.mma (UMMA 128 x BN x 16,
//            A and B straight from shared memory through matrix descriptors) and releases ring
//            slots / publishes accumulators with tcgen05.commit
//   warps 2-5 epilogue: tcgen05.ld the fp32 accumulator (thread = row, 16 columns at a time),
//            apply the fused epilogue, vectorised global stores.  Two accumulator buffers in TMEM
//            let the epilogue of tile i overlap the main loop of tile i+1.
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows], i.e. the
// transposed products of the weight-gradient GEMMs); out-of-range rows / K are zero-filled by TMA.
#include <cuda.h>

#include <mutex>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace odevit {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int NUM_THREADS = 192;

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 128) ? 6 : 4;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator buffers (256 or 512: powers of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct DevArgs {
  int M, N, K;
  int tiles_m, tiles_n, split_k, kb_per_split, num_kb;
  Epi epi;
};

template <int BN, int EPI, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ DevArgs g) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* acc_full = bars + 2 * C::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int i = 0; i < C::STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = g.tiles_m * g.tiles_n * g.split_k;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % g.tiles_n;
        const int mt = (tile / g.tiles_n) % g.tiles_m;
        const int ks = tile / (g.tiles_n * g.tiles_m);
        const int kb0 = ks * g.kb_per_split;
        const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          ptx::mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          if constexpr (!A_MN) {
            ptx::tma_load_2d(sa, &tmA, &full[stage], kb * BK, mt * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              ptx::tma_load_2d(sa + i * (64 * BK * 2), &tmA, &full[stage], mt * BM + i * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            ptx::tma_load_2d(sb, &tmB, &full[stage], kb * BK, nt * BN);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              ptx::tma_load_2d(sb + i * (64 * BK * 2), &tmB, &full[stage], nt * BN + i * 64, kb * BK);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major: 8-row groups 1024 B apart, one UMMA_K (16 elements) = 32 B along the row.
      // MN-major: 64-wide MN blocks (BK*128 B apart), 8-k-row groups 1024 B apart, UMMA_K = 2 groups.
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, A_SBO = 1024, A_KSTEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16, B_SBO = 1024, B_KSTEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int ks = tile / (g.tiles_n * g.tiles_m);
        const int kb0 = ks * g.kb_per_split;
        const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = ptx::smem_desc_sw128(sa + k * A_KSTEP, A_LBO, A_SBO);
            const uint64_t db = ptx::smem_desc_sw128(sb + k * B_KSTEP, B_LBO, B_SBO);
            ptx::mma_bf16_ss(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::mma_commit(&empty[stage]);  // ring slot reusable once these MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&acc_full[acc]);  // accumulator complete
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % g.tiles_n;
      const int mt = (tile / g.tiles_n) % g.tiles_m;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      ptx::mbar_wait(&acc_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = mt * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 2
      for (int c = 0; c < BN / 16; ++c) {
        float v[16];
        ptx::tmem_ld16(taddr + c * 16, v);
        ptx::tmem_ld_wait();
        const int n = nt * BN + c * 16;
        if (row < g.M && n < g.N) {
          if (g.split_k > 1) epi_chunk16<EPI, true>(g.epi, row, n, v);
          else epi_chunk16<EPI, false>(g.epi, row, n, v);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int EPI, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const DevArgs& d, int grid, cudaStream_t s) {
  auto kern = gemm_tc_kernel<BN, EPI, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    ODV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    configured = true;
  }
  kern<<<grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, s>>>(ta, tb, d);
  ODV_LAUNCH_CHECK();
  return 0;
}

template <int BN, int EPI>
int launch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const DevArgs& d, int grid,
                 cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, EPI, false, false>(ta, tb, d, grid, s);
  if (a_mn && b_mn) return launch<BN, EPI, true, true>(ta, tb, d, grid, s);
  return set_error(ODEVIT_ERR_UNSUPPORTED, "gemm_tc: mixed operand majors are not instantiated");
}

}  // namespace

// dims: inner (contiguous) extent, outer extent, outer stride in elements; box = {box_inner, box_outer}
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeFn fn = encode_fn();
  if (!fn) return set_error(ODEVIT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(ODEVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1_elems,
                      uint64_t ld2_elems, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeFn fn = encode_fn();
  if (!fn) return set_error(ODEVIT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld1_elems * 2, ld2_elems * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(ODEVIT_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return 0;
}

bool gemm_tc_supports(const GemmArgs& g) {
  if (g.a_type != DT_BF16 || g.b_type != DT_BF16) return false;
  if (g.batch_outer * g.batch_inner != 1) return false;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if (g.N % 16) return false;
  const bool a_k = (g.a_cs == 1), a_mn = (g.a_rs == 1 && g.a_cs != 1);
  const bool b_k = (g.b_cs == 1), b_mn = (g.b_rs == 1 && g.b_cs != 1);
  if (!(a_k || a_mn) || !(b_k || b_mn)) return false;
  if ((a_k && !a_mn) != (b_k && !b_mn)) return false;  // both K-major or both MN-major
  const long long lda = a_mn ? g.a_cs : g.a_rs, ldb = b_mn ? g.b_cs : g.b_rs;
  if (lda % 8 || ldb % 8) return false;  // TMA: 16-byte global strides
  if ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.B)) & 15) return false;
  if (a_mn ? (g.M % 8) : (g.K % 8)) return false;
  if (b_mn ? (g.N % 8) : (g.K % 8)) return false;
  switch (g.epi_mode) {
    case EPI_STORE: case EPI_FWD1: case EPI_RK: case EPI_BWD3: case EPI_ACCUM: break;
    default: return false;
  }
  if (g.epi_mode == EPI_FWD1 || g.epi_mode == EPI_BWD3) {
    if (g.epi.split % 16 || g.epi.ld_out2 % 8 || (g.epi.out3 && g.epi.ld_out3 % 8)) return false;
  }
  if (g.epi.ld_out % (g.epi.out_type == DT_F32 || g.epi_mode == EPI_RK || g.epi_mode == EPI_ACCUM ? 4 : 8)) return false;
  return true;
}

int gemm_tc(const GemmArgs& g, cudaStream_t s) {
  if (!gemm_tc_supports(g)) return set_error(ODEVIT_ERR_UNSUPPORTED, "gemm_tc: unsupported problem");
  ProfScope prof(g.kclass, s);
  const bool mn = (g.a_cs != 1);
  constexpr int BN = 128;
  DevArgs d;
  d.M = g.M; d.N = g.N; d.K = g.K;
  d.tiles_m = (g.M + BM - 1) / BM;
  d.tiles_n = (g.N + BN - 1) / BN;
  d.num_kb = (g.K + BK - 1) / BK;
  d.split_k = 1;
  const int sms = num_sms();
  if (g.epi_mode == EPI_ACCUM) {
    // weight-gradient shapes: few output tiles, very long K -> split K across the idle SMs
    const int tiles = d.tiles_m * d.tiles_n;
    int sk = sms / tiles;
    if (sk > d.num_kb / 8) sk = d.num_kb / 8;
    if (sk > 1) d.split_k = sk;
  }
  d.kb_per_split = (d.num_kb + d.split_k - 1) / d.split_k;
  d.split_k = (d.num_kb + d.kb_per_split - 1) / d.kb_per_split;
  d.epi = g.epi;
  CUtensorMap ta, tb;
  if (!mn) {
    ODV_TRY(make_tmap_2d_bf16(&ta, g.A, g.K, g.M, g.a_rs, BK, BM));
    ODV_TRY(make_tmap_2d_bf16(&tb, g.B, g.K, g.N, g.b_rs, BK, BN));
  } else {
    ODV_TRY(make_tmap_2d_bf16(&ta, g.A, g.M, g.K, g.a_cs, 64, BK));
    ODV_TRY(make_tmap_2d_bf16(&tb, g.B, g.N, g.K, g.b_cs, 64, BK));
  }
  const int total = d.tiles_m * d.tiles_n * d.split_k;
  const int grid = total < sms ? total : sms;
  switch (g.epi_mode) {
    case EPI_STORE: return launch_major<BN, EPI_STORE>(mn, mn, ta, tb, d, grid, s);
    case EPI_FWD1: return launch_major<BN, EPI_FWD1>(mn, mn, ta, tb, d, grid, s);
    case EPI_RK: return launch_major<BN, EPI_RK>(mn, mn, ta, tb, d, grid, s);
    case EPI_BWD3: return launch_major<BN, EPI_BWD3>(mn, mn, ta, tb, d, grid, s);
    case EPI_ACCUM: return launch_major<BN, EPI_ACCUM>(mn, mn, ta, tb, d, grid, s);
    default: return set_error(ODEVIT_ERR_INVALID_ARG, "gemm_tc: bad epilogue %d", g.epi_mode);
  }
}

}  // namespace odevit
