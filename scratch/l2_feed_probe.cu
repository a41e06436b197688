// l2_feed_probe.cu — how fast can the SMs of a B200 pull GEMM operand tiles out of L2 with TMA, and what does sharing the B tile
// between two CTA pairs of a 4-CTA cluster (TMA multicast) buy? No math: a producer thread per CTA streams the tiles of a
// 50176 x 3072 x 768 GEMM walk (A 128 x 64 + B 128 x 64 bf16 per step = 32 KB per CTA) through a 5-stage ring; a consumer thread
// frees each stage as soon as it is full.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/l2_feed_probe scratch/l2_feed_probe.cu && scratch/l2_feed_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int STAGE_BYTES = 32768, KSTEPS = 12, NBLK = 12;

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bar_arrive_remote(uint64_t* b, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(s32(b)), "r"(cta) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 2000000000LL) { printf("probe: barrier timeout block %d\n", blockIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(s32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(s32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// MODE 0: every CTA loads its own A and B halves (what a CTA pair does today); MODE 1: clusters of 4 = two pairs on neighbouring row
// blocks and the same column block: every CTA loads its A half and HALF of its B half, multicast to the CTA of the other pair that
// needs the same columns.
template <int MODE, int STAGES>
__global__ void __launch_bounds__(128, 1) feed(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                               const __grid_constant__ CUtensorMap tmBq, int steps, int m_blocks, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const uint32_t rank = MODE ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { bar_init(&full[s], 1); bar_init(&empty[s], MODE ? 2 : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE) cluster_sync();
  const long long t0 = clock64();
  // row block (128 rows) of this CTA: pairs own 256 rows; in MODE 1 a cluster owns 512
  const int pair = blockIdx.x >> 1, half = blockIdx.x & 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < steps; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      bar_wait(&empty[s], ph ^ 1);
      const int tile = i / KSTEPS, k = (i % KSTEPS) * 64;
      const int nb = tile % NBLK;
      const int mb = (pair + (tile / NBLK) * (gridDim.x >> 1)) % m_blocks;
      uint8_t* st = smem + s * STAGE_BYTES;
      bar_expect(&full[s], STAGE_BYTES);
      tma2d(st, &tmA, &full[s], k, mb * 256 + half * 128);
      if (MODE == 0) {
        tma2d(st + 16384, &tmB, &full[s], k, nb * 256 + half * 128);
      } else {
        const uint32_t q = rank >> 1;                                  // which 64 columns of the shared B half this CTA fetches
        const uint16_t mask = static_cast<uint16_t>((1u << (rank & 1)) | (1u << ((rank & 1) + 2)));
        tma2d_mc(st + 16384 + q * 8192, &tmBq, &full[s], k, nb * 256 + (rank & 1) * 128 + q * 64, mask);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < steps; ++i) {
      const int s = i % STAGES;
      bar_wait(&full[s], (i / STAGES) & 1);
      if (MODE == 0) bar_arrive(&empty[s]);
      else { bar_arrive_remote(&empty[s], rank); bar_arrive_remote(&empty[s], rank ^ 2); }   // both CTAs that write this stage
    }
  }
  __syncthreads();
  if (MODE) cluster_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

static CUtensorMap make_map(void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  using EncFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncFn enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &q));
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
  return m;
}

template <int STAGES>
void run_all(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBq, int steps, int M, long long* cyc, cudaEvent_t e0, cudaEvent_t e1) {
  const int smem = STAGES * STAGE_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(feed<0, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(feed<1, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  printf("-- %d stages of 32 KB in flight per CTA\n", STAGES);
  for (int mode = 0; mode < 2; ++mode) {
    for (int grid : {148, 132}) {
      if (mode == 1 && grid == 148) continue;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {mode ? 4u : 2u, 1, 1};
      cfg.attrs = at; cfg.numAttrs = 1;
      float best = 1e9f; long long c = 0;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) CK(cudaLaunchKernelEx(&cfg, feed<0, STAGES>, tmA, tmB, tmBq, steps, M / 256, cyc));
        else CK(cudaLaunchKernelEx(&cfg, feed<1, STAGES>, tmA, tmB, tmBq, steps, M / 256, cyc));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost)); }
      }
      const double filled = (double)grid * steps * STAGE_BYTES;            // bytes that landed in shared memory
      const double pulled = mode == 0 ? filled : filled * 0.75;            // bytes that left L2
      printf("%-28s grid %3d: %7.3f ms, %7.1f cycles per 32 KB stage per CTA = %5.1f B/clk/SM into smem; %6.2f TB/s into smem, %6.2f TB/s out of L2; "
             "a 128x256x64 MMA step takes 512 cycles -> feed allows %4.0f %% tensor duty on %d SMs (x%d/148 = %4.0f %% of the chip)\n",
             mode ? "clusters of 4, B multicast" : "pairs, no multicast", grid, best, (double)c / steps, STAGE_BYTES / ((double)c / steps),
             filled / best / 1e9, pulled / best / 1e9, 100.0 * 512.0 / ((double)c / steps), grid, grid,
             100.0 * 512.0 / ((double)c / steps) * grid / 148.0);
    }
  }
}

int main() {
  const int M = 50176, N = 3072, K = 768, steps = KSTEPS * NBLK * 6;
  void *A, *B; long long* cyc;
  CK(cudaMalloc(&A, (size_t)M * K * 2)); CK(cudaMalloc(&B, (size_t)N * K * 2)); CK(cudaMalloc(&cyc, 8));
  CK(cudaMemset(A, 0, (size_t)M * K * 2)); CK(cudaMemset(B, 0, (size_t)N * K * 2));
  CUtensorMap tmA = make_map(A, M, K, 128), tmB = make_map(B, N, K, 128), tmBq = make_map(B, N, K, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  run_all<2>(tmA, tmB, tmBq, steps, M, cyc, e0, e1);
  run_all<3>(tmA, tmB, tmBq, steps, M, cyc, e0, e1);
  run_all<5>(tmA, tmB, tmBq, steps, M, cyc, e0, e1);
  run_all<6>(tmA, tmB, tmBq, steps, M, cyc, e0, e1);
  return 0;
}

