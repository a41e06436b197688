// poll_probe.cu — do warps that POLL an mbarrier slow down the exponentials of the working warps on the same sub-partition?
// One CTA: 2 warps per sub-partition run the softmax pattern (FFMA2 -> 2 x MUFU.EX2 -> FADD2 + pack); 0 / 1 / 2 / 3 more warps per
// sub-partition wait on an mbarrier phase that only completes when the workers are done, with (a) a try_wait loop, (b) try_wait
// with a 10 ms suspend hint, (c) try_wait + __nanosleep(256) back-off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/poll_probe scratch/poll_probe.cu && scratch/poll_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t f2fp(float lo, float hi) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}

constexpr int N = 64, REP = 64, WORK_WARPS = 8;

template <int POLL>   // 0 try_wait loop, 1 suspend hint, 2 nanosleep back-off
__global__ void __launch_bounds__(640, 1) probe(float* out, long long* cyc, float seed, int pollers_per_smsp) {
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(WORK_WARPS));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= WORK_WARPS + 4 * pollers_per_smsp) return;
  if (warp >= WORK_WARPS) {   // pollers
    if (POLL == 0) { while (!try_wait(&bar, 0)) {} }
    else if (POLL == 1) { while (!try_wait_hint(&bar, 0, 10000000u)) {} }
    else { while (!try_wait(&bar, 0)) __nanosleep(256); }
    return;
  }
  float s[N];
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = seed + 0.001f * i + threadIdx.x * 1e-6f;
  uint64_t acc0 = pack2(0.f, 0.f), acc1 = acc0;
  uint32_t pk = 0;
  const uint64_t sc = pack2(0.999f, 0.999f), nb = pack2(-0.01f, -0.01f);
  const long long t0 = clock64();
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      float x0, x1;
      unpack2(ffma2(pack2(s[i], s[i + 1]), sc, nb), x0, x1);
      const float p0 = ex2(x0), p1 = ex2(x1);
      if (i & 2) acc1 = fadd2(acc1, pack2(p0, p1)); else acc0 = fadd2(acc0, pack2(p0, p1));
      pk ^= f2fp(p0, p1);
      s[i] = __uint_as_float(__float_as_uint(x0) ^ (pk & 1u)); s[i + 1] = __uint_as_float(__float_as_uint(x1) ^ (pk & 1u));
    }
  }
  const long long t1 = clock64();
  float a, b; unpack2(fadd2(acc0, acc1), a, b);
  float sum = a + b;
#pragma unroll
  for (int i = 0; i < N; ++i) sum += s[i];
  out[threadIdx.x] = sum;
  if (lane == 0) {
    cyc[warp] = t1 - t0;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024); cudaMalloc(&cyc, 64 * 8);
  const char* names[3] = {"try_wait loop", "try_wait + 10 ms suspend hint", "try_wait + nanosleep(256)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int pollers = 0; pollers <= 3; ++pollers) {
      if (mode == 0) probe<0><<<1, 640>>>(out, cyc, 0.5f, pollers);
      if (mode == 1) probe<1><<<1, 640>>>(out, cyc, 0.5f, pollers);
      if (mode == 2) probe<2><<<1, 640>>>(out, cyc, 0.5f, pollers);
      cudaError_t e = cudaDeviceSynchronize();
      long long c[8]; cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
      printf("%-30s %d polling warp(s) per sub-partition: %6.2f cycles per element per sub-partition (2 working warps)  [%s]\n", names[mode], pollers,
             (double)c[0] / (REP * N) / 2, cudaGetErrorString(e));
    }
  return 0;
}
