// softmax_probe.cu — the exponential phase of one softmax step as vf_attention.cu runs it (64 scores per thread in registers ->
// FFMA2 scale/shift -> exp2 -> FADD2 row sum + bf16 pack -> 16-byte stores), with every POLY-th pair evaluated on the FMA pipe
// (round-to-nearest split + degree-3 polynomial in packed f32x2 + exponent insertion) instead of MUFU.EX2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -maxrregcount=104 -o scratch/softmax_probe scratch/softmax_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp(float lo, float hi) { uint32_t d; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// 2^x for a pair, FMA pipe only. x clamped at -126 (smaller results flush to ~1e-38 instead of 0: irrelevant for a softmax term).
__device__ __forceinline__ void exp2_poly2(uint64_t X, float& p0, float& p1) {
  const float MAGIC = 12582912.f;   // 1.5 * 2^23: x + MAGIC rounds x to the nearest integer, kept in the low mantissa bits
  float x0, x1;
  unpack2(X, x0, x1);
  x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
  X = pack2(x0, x1);
  const uint64_t J = fadd2(X, pack2(MAGIC, MAGIC));
  const uint64_t N = fadd2(J, pack2(-MAGIC, -MAGIC));
  const uint64_t Fr = ffma2(N, pack2(-1.f, -1.f), X);                    // f = x - n in [-0.5, 0.5]
  uint64_t P = ffma2(Fr, pack2(0.05508868f, 0.05508868f), pack2(0.24260405f, 0.24260405f));
  P = ffma2(P, Fr, pack2(0.69327624f, 0.69327624f));
  P = ffma2(P, Fr, pack2(0.99992894f, 0.99992894f));
  float j0, j1, q0, q1;
  unpack2(J, j0, j1);
  unpack2(P, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(j0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(j1) << 23));
}

constexpr int KT = 64;
constexpr int REP = 128;

// POLY = 0: every pair on the MUFU; POLY = n: pair e of a 16-pair chunk goes to the FMA pipe when e % n == n - 1
template <int POLY>
__global__ void __maxnreg__(104) probe(float* out, long long* cycles, const float* in, float scale) {
  extern __shared__ uint4 sink[];
  float s[KT];
#pragma unroll
  for (int i = 0; i < KT; ++i) s[i] = in[i * 32 + (threadIdx.x & 31)];
  float l = 0.f;
  uint4* my = sink + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < REP; ++r) {
    const float mb = 0.25f + 1e-3f * r;
    const uint64_t sc2 = pack2(scale, scale), nb2 = pack2(-mb, -mb);
    uint64_t acc0 = pack2(0.f, 0.f), acc1 = acc0;
#pragma unroll
    for (int c = 0; c < KT / 32; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const uint64_t X = ffma2(pack2(s[c * 32 + 2 * e], s[c * 32 + 2 * e + 1]), sc2, nb2);
        float p0, p1;
        if (POLY > 0 && e % (POLY > 0 ? POLY : 1) == POLY - 1) exp2_poly2(X, p0, p1);
        else { float x0, x1; unpack2(X, x0, x1); p0 = ex2(x0); p1 = ex2(x1); }
        if (e & 1) acc1 = fadd2(acc1, pack2(p0, p1)); else acc0 = fadd2(acc0, pack2(p0, p1));
        pk[e] = f2fp(p0, p1);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) my[(c * 4 + q) * blockDim.x] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
    float a0, a1;
    unpack2(fadd2(acc0, acc1), a0, a1);
    l += a0 + a1;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int POLY>
void run(const char* name, float* out, long long* cyc, const float* in) {
  cudaFuncSetAttribute(probe<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 512 * 16);
  for (int w = 1; w <= 4; w *= 2) {
    probe<POLY><<<1, 128 * w, 8 * 128 * w * 16>>>(out, cyc, in, 0.18f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_elem = (double)c / (REP * KT);
    printf("%-34s %d warp(s)/SMSP: %6.2f cycles per element per warp, %6.2f per element per sub-partition\n", name, w, per_elem, per_elem / w);
  }
}

__global__ void check(float* err) {   // accuracy of the polynomial path against exp2f on (-126, 4)
  float worst = 0.f;
  for (int i = threadIdx.x; i < 1 << 20; i += blockDim.x) {
    const float x = -126.f + 130.f * (i / float(1 << 20));
    float p0, p1;
    exp2_poly2(pack2(x, x - 0.37f), p0, p1);
    worst = fmaxf(worst, fabsf(p0 / exp2f(x) - 1.f));
    worst = fmaxf(worst, fabsf(p1 / exp2f(x - 0.37f < -126.f ? -126.f : x - 0.37f) - 1.f));
  }
  atomicMax(reinterpret_cast<int*>(err), __float_as_int(worst));
}

int main() {
  float *out, *in, *err; long long* cyc;
  cudaMalloc(&out, 4 * 1024); cudaMalloc(&cyc, 8); cudaMalloc(&in, 64 * 32 * 4); cudaMalloc(&err, 4);
  float h[64 * 32];
  for (int i = 0; i < 64 * 32; ++i) h[i] = -20.f * ((i * 2654435761u) >> 8 & 0xffff) / 65536.f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMemset(err, 0, 4);
  check<<<1, 256>>>(err);
  float e; cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost);
  printf("polynomial exp2: max relative error %.3e\n", e);
  run<0>("all pairs on MUFU.EX2", out, cyc, in);
  run<8>("every 8th pair on the FMA pipe", out, cyc, in);
  run<4>("every 4th pair on the FMA pipe", out, cyc, in);
  run<3>("every 3rd pair on the FMA pipe", out, cyc, in);
  run<2>("every 2nd pair on the FMA pipe", out, cyc, in);
  return 0;
}
