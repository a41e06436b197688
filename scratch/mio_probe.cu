// mio_probe.cu — do MUFU results share a return path with shared-memory / tensor-memory loads?
// One CTA, 8 warps (2 per sub-partition): warps 0-3 run independent MUFU.EX2, warps 4-7 run (a) nothing, (b) LDS.128
// streams, (c) tcgen05.ld 32x32b.x32 streams. Prints cycles per MUFU warp-instruction for the MUFU warps in each case and
// the bytes per cycle the other warps moved.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mio_probe scratch/mio_probe.cu && scratch/mio_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

constexpr int N = 64, REP = 128;

template <int OTHER>   // 0 idle, 1 LDS.128, 2 tcgen05.ld x32
__global__ void __launch_bounds__(256, 1) probe(float* out, long long* cyc, float seed) {
  __shared__ __align__(16) float buf[8192];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += 256) buf[i] = seed * i;
  if (OTHER == 2 && warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = OTHER == 2 ? tmem_slot : 0;
  float s[N];
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = seed + 0.001f * i + threadIdx.x * 1e-6f;
  float sink = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 4) {
    for (int r = 0; r < REP; ++r) {
#pragma unroll
      for (int i = 0; i < N; ++i) s[i] = ex2(s[i]) * 0.5f;
    }
  } else if (OTHER == 1) {
    for (int r = 0; r < REP; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {     // 16 x 128-bit loads per repetition and lane: 8 KB per warp
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(buf) + ((i * 32 + lane) * 16) % 32768));
        sink += v.x + v.y + v.z + v.w;
      }
    }
  } else if (OTHER == 2) {
    const uint32_t taddr = tbase + ((static_cast<uint32_t>((warp & 3) * 32)) << 16);
    for (int r = 0; r < REP; ++r) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {      // 2 x (32 lanes x 32 columns x 4 B) = 8 KB per warp per repetition
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr + i * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 32; ++k) sink += __uint_as_float(v[k] & 0x3f800000u);
      }
    }
  }
  const long long t1 = clock64();
  float sum = sink;
#pragma unroll
  for (int i = 0; i < N; ++i) sum += s[i];
  out[threadIdx.x] = sum;
  if (lane == 0) cyc[warp] = t1 - t0;
  __syncthreads();
  if (OTHER == 2 && warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int OTHER>
void run(const char* name, float* out, long long* cyc) {
  probe<OTHER><<<1, 256>>>(out, cyc, 0.5f);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[8];
  cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
  printf("%-34s MUFU warps: %6.2f cycles per MUFU.EX2 (alone: 8.0)", name, (double)c[0] / (REP * N));
  if (OTHER) printf("; other warps: %7.1f cycles per 8 KB per warp = %5.1f B/clk per SM for the four of them", (double)c[4] / REP, 4 * 8192.0 * REP / c[4]);
  printf("  [%s]\n", cudaGetErrorString(e));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 256); cudaMalloc(&cyc, 64);
  run<0>("other warps idle", out, cyc);
  run<1>("other warps: LDS.128 stream", out, cyc);
  run<2>("other warps: tcgen05.ld x32 stream", out, cyc);
  return 0;
}
