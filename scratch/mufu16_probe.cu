// mufu16_probe.cu — is MUFU.EX2.F16 / .BF16 issued faster than the fp32 MUFU.EX2 on sm_100a? (cycles per warp instruction per sub-partition)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int N = 32, REP = 256;
template <int MODE>
__global__ void probe(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = seed + i * 0x00010001u + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (MODE == 0) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(v[i]))); v[i] = __float_as_uint(y) & 0x3effffffu; }
      if (MODE == 1) { unsigned short h; asm volatile("ex2.approx.f16 %0, %1;" : "=h"(h) : "h"((unsigned short)v[i])); v[i] = h & 0x3bffu; }
      if (MODE == 2) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(v[i])); v[i] = y & 0x3bff3bffu; }
      if (MODE == 3) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(v[i])); v[i] = y & 0x3eff3effu; }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int per, uint32_t* out, long long* cyc) {
  for (int w = 1; w <= 4; w *= 2) {
    probe<MODE><<<1, 128 * w>>>(out, cyc, 0x3c003c00u);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %d warp(s)/SMSP: %6.2f cycles per exponential per sub-partition\n", name, w, (double)c / (REP * N * per) / w);
  }
}
int main() {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  run<0>("ex2.approx.ftz.f32", 1, out, cyc);
  run<1>("ex2.approx.f16", 1, out, cyc);
  run<2>("ex2.approx.f16x2 (2 MUFU)", 2, out, cyc);
  run<3>("ex2.approx.ftz.bf16x2 (2 MUFU)", 2, out, cyc);
  return 0;
}
