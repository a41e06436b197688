// Probe: how does TMA lay out a box whose inner dimension (32 B) is narrower than the swizzle span?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t* out, int out_elems, int expect_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint16_t* s = reinterpret_cast<uint16_t*>(smem);
  for (int i = threadIdx.x; i < out_elems; i += blockDim.x) s[i] = 0xFFFF;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(expect_bytes));
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(d), "l"((uint64_t)&tm), "r"(b), "r"(0), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0; long long t0 = clock64();
    while (!ok && clock64() - t0 < 200000000LL) {
      asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0,1,0,P; }" : "=r"(ok) : "r"(b));
    }
    if (!ok) printf("TIMEOUT (expect_tx %d never completed)\n", expect_bytes);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < out_elems; i += blockDim.x) out[i] = s[i];
}
int main() {
  // global tensor: [rows=64][py=4][16 px] as a 3-D map: dim0 px(16, stride 1), dim1 py(4, stride 1024 elems), dim2 tok(64, stride 16 elems)
  const int W = 1024;
  std::vector<uint16_t> h(4 * W);
  for (int py = 0; py < 4; ++py) for (int x = 0; x < W; ++x) h[py * W + x] = (uint16_t)((py << 12) | x);  // value encodes (py, x): tok = x/16, px = x%16
  uint16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* sym; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  auto enc = (EncodeTiledFn)sym;
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {16, 4, 64}; cuuint64_t str[2] = {W * 2, 32}; cuuint32_t box[3] = {16, 4, 16}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d encode -> %d\n", mode, (int)r);
    const int out_elems = 8192;  // 16 KB window
    uint16_t* o; cudaMalloc(&o, out_elems * 2);
    probe<<<1, 128, out_elems * 2>>>(tm, o, out_elems, 16 * 4 * 16 * 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync -> %s\n", cudaGetErrorString(e));
    std::vector<uint16_t> ho(out_elems); cudaMemcpy(ho.data(), o, out_elems * 2, cudaMemcpyDeviceToHost);
    int last = -1; for (int i = 0; i < out_elems; ++i) if (ho[i] != 0xFFFF) last = i;
    printf("last written element index %d (bytes %d)\n", last, (last + 1) * 2);
    // print 16-byte chunks of the first 16 smem rows of 128 B: each chunk as (tok, py, px0)
    for (int row = 0; row < 18; ++row) {
      printf("smem+%4d:", row * 128);
      for (int c = 0; c < 8; ++c) { uint16_t v = ho[row * 64 + c * 8]; if (v == 0xFFFF) printf("  ----  "); else printf(" t%02dy%dx%02d", (v & 0xFFF) / 16, v >> 12, (v & 0xFFF) % 16); }
      printf("\n");
    }
  }
  return 0;
}
