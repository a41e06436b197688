// xu_probe.cu — what bounds the softmax inner loop on sm_100a? Cycles per warp-instruction of MUFU.EX2 alone, with the bf16
// pack (F2FP vs PRMT), and in the softmax pattern (FFMA2 -> 2 x MUFU -> FADD2 + pack), for 1 / 2 / 4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/xu_probe scratch/xu_probe.cu && scratch/xu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp(float lo, float hi) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
__device__ __forceinline__ uint32_t prmt(float lo, float hi) { uint32_t d; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(d) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi))); return d; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int N = 64;      // elements per thread per repetition (as one softmax step of 64 keys)
constexpr int REP = 64;

// MODE 7: MUFU + FADD2 + F2FP (no FFMA2); 8: FFMA2 + MUFU + F2FP (no FADD2); 9: softmax pattern in two phases (all FFMA2 + MUFU of
// the 64 elements first, the row sum started from the LAST pair so that no FADD2 can be hoisted, then a branch on the last
// exponential, then the packs); 10: as 3 but the sum chain starts from the last pair only (no branch)
// MODE 0: MUFU only (N independent); 1: MUFU + F2FP per pair; 2: MUFU + PRMT per pair;
// 3: softmax pattern with F2FP; 4: softmax pattern with PRMT; 5: F2FP only; 6: MUFU, dependent chain (latency)
template <int MODE>
__global__ void probe(float* out, long long* cycles, float seed) {
  float s[N];
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = seed + 0.001f * i + threadIdx.x * 1e-6f;
  uint64_t acc0 = pack2(0.f, 0.f), acc1 = acc0;
  uint32_t pk = 0;
  const uint64_t sc = pack2(0.999f, 0.999f), nb = pack2(-0.01f, -0.01f);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < REP; ++r) {
    if (MODE == 6) {
      float x = s[0];
#pragma unroll
      for (int i = 0; i < N; ++i) x = ex2(x * 0.001f);
      s[0] = x;
    } else if (MODE == 9 || MODE == 10) {
      float e[N];
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        float x0, x1;
        unpack2(ffma2(pack2(s[i], s[i + 1]), sc, nb), x0, x1);
        e[i] = ex2(x0); e[i + 1] = ex2(x1);
      }
      if (MODE == 9) { if (e[N - 1] < 0.f) { out[threadIdx.x] = e[N - 1]; return; } }   // never taken; ends the basic block
      uint64_t a0 = pack2(e[N - 2], e[N - 1]), a1 = pack2(e[N - 4], e[N - 3]);
#pragma unroll
      for (int i = 0; i < N - 4; i += 2) {
        if (i & 2) a1 = fadd2(a1, pack2(e[i], e[i + 1])); else a0 = fadd2(a0, pack2(e[i], e[i + 1]));
      }
      acc0 = fadd2(acc0, a0); acc1 = fadd2(acc1, a1);
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        pk ^= f2fp(e[i], e[i + 1]);
        s[i] = __uint_as_float(__float_as_uint(s[i]) ^ (pk & 1u)); s[i + 1] = __uint_as_float(__float_as_uint(s[i + 1]) ^ (pk & 1u));
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        float x0 = s[i], x1 = s[i + 1];
        if (MODE == 3 || MODE == 4 || MODE == 8) unpack2(ffma2(pack2(x0, x1), sc, nb), x0, x1);
        float p0 = x0, p1 = x1;
        if (MODE != 5) { p0 = ex2(x0); p1 = ex2(x1); }
        if (MODE == 3 || MODE == 4 || MODE == 7) {
          if (i & 2) acc1 = fadd2(acc1, pack2(p0, p1)); else acc0 = fadd2(acc0, pack2(p0, p1));
        }
        if (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 7 || MODE == 8) pk ^= f2fp(p0, p1);
        if (MODE == 2 || MODE == 4) pk ^= prmt(p0, p1);
        if (MODE == 0) { s[i] = p0 * 0.5f; s[i + 1] = p1 * 0.5f; }
        else { s[i] = __uint_as_float(__float_as_uint(x0) ^ (pk & 1u)); s[i + 1] = __uint_as_float(__float_as_uint(x1) ^ (pk & 1u)); }   // keep every repetition live
      }
    }
  }
  const long long t1 = clock64();
  float a, b; unpack2(fadd2(acc0, acc1), a, b);
  float sum = a + b + __uint_as_float(pk & 0x3f800000u);
#pragma unroll
  for (int i = 0; i < N; ++i) sum += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, int n_mufu_per_elem_x2, float* out, long long* cyc) {
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    probe<MODE><<<1, 128 * warps_per_smsp>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_elem = (double)c / (REP * N);                 // cycles per element of one warp
    printf("%-44s %d warp(s)/SMSP: %7.2f cycles per element per warp, %6.2f per element per sub-partition\n", name, warps_per_smsp,
           per_elem, per_elem / warps_per_smsp);
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024); cudaMalloc(&cyc, 8);
  run<0>("MUFU.EX2 only", 2, out, cyc);
  run<5>("F2FP (cvt.rn.bf16x2.f32) only, per pair/2", 0, out, cyc);
  run<1>("MUFU.EX2 + F2FP pack", 2, out, cyc);
  run<2>("MUFU.EX2 + PRMT pack", 2, out, cyc);
  run<3>("softmax pattern: FFMA2, 2 MUFU, FADD2, F2FP", 2, out, cyc);
  run<4>("softmax pattern: FFMA2, 2 MUFU, FADD2, PRMT", 2, out, cyc);
  run<6>("MUFU.EX2 dependent chain (latency)", 2, out, cyc);
  run<7>("2 MUFU, FADD2, F2FP (no FFMA2)", 2, out, cyc);
  run<8>("FFMA2, 2 MUFU, F2FP (no FADD2)", 2, out, cyc);
  run<9>("two-phase softmax pattern, branch between", 2, out, cyc);
  run<10>("two-phase softmax pattern, sum from the last pair", 2, out, cyc);
  return 0;
}
