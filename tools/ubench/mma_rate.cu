// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16) for the shapes / operand sources the attention
// kernels use.  One CTA per SM (all 148 run, CTA 0 reports), 64 back-to-back MMAs per case, clock() around
// issue .. commit .. mbarrier wait.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../odevit_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace odevit;

__global__ void __launch_bounds__(128, 1) k(int* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(&slot, 512);
  ptx::fence_async_shared();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t base = ptx::smem_u32(smem);
    uint32_t ph = 0;
    int res[16];
    int nres = 0;
    auto run = [&](auto issue) {
      // warm
      issue(); ptx::mma_commit(&bar); ptx::mbar_wait(&bar, ph); ph ^= 1;
      const long long t0 = clock64();
      issue();
      ptx::mma_commit(&bar);
      ptx::mbar_wait(&bar, ph); ph ^= 1;
      res[nres++] = (int)(clock64() - t0);
    };
    const uint64_t dA_k = ptx::smem_desc_sw128(base, 16, 1024);             // K-major A tile [128 x 64]
    const uint64_t dB_k = ptx::smem_desc_sw128(base + 32768, 16, 1024);     // K-major B tile [<=256 x 64]
    const uint64_t dA_mn = ptx::smem_desc_sw128(base + 65536, 16384, 1024); // MN-major A [K rows][128] (2 atoms)
    const uint64_t dB_mn = ptx::smem_desc_sw128(base + 98304, 8192, 1024);  // MN-major B [K rows][64]
    // 0: SS K-major N=128 (S^T tile), 64 MMAs accumulating into one tile
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ss(tmem, dA_k + 2 * (i & 3), dB_k + 2 * (i & 3), ptx::idesc_bf16(128, 128, 0, 0), 1u); });
    // 1: SS K-major N=64
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ss(tmem, dA_k + 2 * (i & 3), dB_k + 2 * (i & 3), ptx::idesc_bf16(128, 64, 0, 0), 1u); });
    // 2: SS K-major N=256
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ss(tmem, dA_k + 2 * (i & 3), dB_k + 2 * (i & 3), ptx::idesc_bf16(128, 256, 0, 0), 1u); });
    // 3: TS (A in TMEM) N=64, B MN-major -- dV / dK, P.V
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ts(tmem + 256, tmem + (i & 7) * 8, dB_mn + 128 * (i & 7), ptx::idesc_bf16(128, 64, 0, 1), 1u); });
    // 4: TS N=64, two accumulators interleaved (dV, dK)
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ts(tmem + 256 + (i & 1) * 64, tmem + (i & 7) * 8, dB_mn + 128 * (i & 7), ptx::idesc_bf16(128, 64, 0, 1), 1u); });
    // 5: SS, A MN-major (M=128: two atoms), B MN-major, N=64 -- dQ
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ss(tmem + 384, dA_mn + 128 * (i & 7), dB_mn + 128 * (i & 7), ptx::idesc_bf16(128, 64, 1, 1), 1u); });
    // 6: TS N=128, B MN-major (two atoms) -- hypothetical wider value tile
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ts(tmem + 256, tmem + (i & 7) * 8, ptx::smem_desc_sw128(base + 98304, 16384, 1024) + 128 * (i & 7), ptx::idesc_bf16(128, 128, 0, 1), 1u); });
    // 7: TS N=64, B K-major
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ts(tmem + 256, tmem + (i & 7) * 8, dB_k + 2 * (i & 3), ptx::idesc_bf16(128, 64, 0, 0), 1u); });
    // 8: SS K-major N=208 (the forward's S tile at N=207)
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ss(tmem, dA_k + 2 * (i & 3), dB_k + 2 * (i & 3), ptx::idesc_bf16(128, 208, 0, 0), 1u); });
    // 9: SS K-major N=16
    run([&] { for (int i = 0; i < 64; ++i) ptx::mma_bf16_ss(tmem, dA_k + 2 * (i & 3), dB_k + 2 * (i & 3), ptx::idesc_bf16(128, 16, 0, 0), 1u); });
    if (blockIdx.x == 0) for (int i = 0; i < nres; ++i) out[i] = res[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  int* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 200 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  int h[16]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const char* names[] = {"SS K-major N=128", "SS K-major N=64", "SS K-major N=256", "TS N=64 B MN-major", "TS N=64 x2 accumulators",
                         "SS A,B MN-major N=64 (dQ)", "TS N=128 B MN-major", "TS N=64 B K-major", "SS K-major N=208", "SS K-major N=16"};
  printf("status %s\n", cudaGetErrorString(e));
  for (int i = 0; i < 10; ++i) printf("%-32s %6d cycles / 64 MMAs = %.1f per MMA\n", names[i], h[i], h[i] / 64.0);
  return 0;
}
