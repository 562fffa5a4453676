// How many bytes per clock can TMA deliver into one SM's shared memory from L2, and does multicast raise it?
// Decides whether the tcgen05 GEMM (operand-supply-bound at ~46 B/clk/SM with unicast 32 KB stages) gains from
// cluster multicast.  Every CTA runs a 4-stage ring of 32 KB stages (two 16 KB boxes of 128 rows x 64 bf16):
//   mode 0  unicast, every CTA streams its OWN rows                 (what the GEMM does for A)
//   mode 1  unicast, all CTAs of a cluster stream the SAME rows     (L2-side dedup?)
//   mode 2  multicast: each CTA loads 1/CS of every box and multicasts it to the whole cluster
// The source matrix is L2-resident (32 MB).  Reports bytes landed per SM per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_feed tma_feed.cu -lcuda && ./tma_feed
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int STAGES = 4, BOX = 16384, STAGE = 2 * BOX;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) { while (!mbar_try(b, ph)) {} }
__device__ __forceinline__ uint32_t mapa(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void arrive_cluster(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

// tm: box 64 x 128 rows; tm_part: box 64 x (128 / CS) rows
__global__ void __launch_bounds__(128, 1)
feed(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_part, int mode, int cs, int iters,
     int rows, int kblocks, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const uint32_t rank = cs > 1 ? cta_rank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], mode == 2 ? cs : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (cs > 1) cluster_sync(); else __syncthreads();
  const int cluster_id = blockIdx.x / cs;
  // rows streamed by this CTA: own 128-row panel (mode 0) or the cluster's panel (modes 1, 2)
  const int panel = (mode == 0 ? (int)blockIdx.x : cluster_id) * 128 % rows;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {          // producer
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      const int kb = it % kblocks;
      mbar_wait(&empty[st], ph ^ 1);
      uint8_t* dst = ring + st * STAGE;
      mbar_expect(&full[st], STAGE);
      if (mode != 2) {
        tma_2d(dst, &tm, &full[st], kb * 64, panel);
        tma_2d(dst + BOX, &tm, &full[st], kb * 64, (panel + rows / 2) % rows);
      } else {
        const int part = 128 / cs;   // rows of each box this CTA fetches for everybody
        const uint16_t mask = (uint16_t)((1u << cs) - 1);
        tma_2d_mc(dst + rank * part * 128, &tm_part, &full[st], kb * 64, panel + rank * part, mask);
        tma_2d_mc(dst + BOX + rank * part * 128, &tm_part, &full[st], kb * 64, (panel + rows / 2) % rows + rank * part, mask);
      }
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer: hands every stage straight back (to every producer that writes it)
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[st], ph);
      if (mode == 2) { for (int r = 0; r < cs; ++r) arrive_cluster(mapa(&empty[st], r)); }
      else arrive_cluster(mapa(&empty[st], rank));
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (cs > 1) cluster_sync();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 4096, K = 4096;   // 32 MB of bf16: L2-resident
  void* d; CK(cudaMalloc(&d, (size_t)rows * K * 2)); CK(cudaMemset(d, 0, (size_t)rows * K * 2));
  long long* cyc; CK(cudaMalloc(&cyc, 256 * sizeof(long long)));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  auto make = [&](int box_rows) {
    CUtensorMap m; cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  const int smem = STAGES * STAGE + 1024;
  CK(cudaFuncSetAttribute(feed, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int cs : {1, 2, 4, 8}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (cs == 1 && mode != 0) continue;
      CUtensorMap tm = make(128), tmp = make(128 / cs);
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = a; cfg.numAttrs = 1;
      cfg.gridDim = dim3(cs * 64);
      int nclusters = 0; CK(cudaOccupancyMaxActiveClusters(&nclusters, feed, &cfg));
      if (nclusters * cs > 148) nclusters = 148 / cs;
      cfg.gridDim = dim3(nclusters * cs);
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, feed, tm, tmp, mode, cs, iters, rows, K / 64, cyc));
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nclusters * cs, cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < nclusters * cs; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes_per_sm = (double)iters * STAGE;
      printf("cluster %d mode %d (%s): %3d CTAs  %.1f us  %.1f B/clk/SM landed  (%.2f TB/s landed chip-wide, %.2f GHz)\n", cs, mode,
             mode == 0 ? "unicast own rows" : mode == 1 ? "unicast same rows" : "multicast", nclusters * cs, ms * 1e3,
             bytes_per_sm / (double)mx, bytes_per_sm * nclusters * cs / (ms * 1e-3) / 1e12, (double)mx / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
