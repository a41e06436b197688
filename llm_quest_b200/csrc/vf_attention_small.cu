// vf_attention_small.cu — bidirectional attention for head dims the tcgen05 kernel is not built for.
//
// vf_attention_fwd (vf_attention.cu) is a head_dim-64 tensor-core kernel. The reference's TINY_VIT_CONFIG
// (config.py:175-186: emb 256, 8 heads -> head_dim 32, 4x4 patches of 32x32 images -> S = 65) is a 65 x 65 x 32
// problem per head: far below one MMA tile, so it runs on the CUDA cores instead — same token-major layout in and out,
// same semantics (softmax(q k^T * scale) v, vit_attention.py:74-82), fp32 scores, probabilities and accumulation.
//
//   grid (B*H, ceil(S / 32)), 8 warps per CTA; K and V of the (sample, head) staged once in shared memory (rows padded
//   by 16 B: conflict-free 128-bit reads with lane = key); a warp owns one query row at a time:
//     scores: lane = key (k, k+32, ...), 128-bit dot products against the query held in shared memory;
//     softmax: warp max / sum over the lanes' register-resident scores;
//     context: probabilities parked in shared memory, lane = output dims (d, d+32, ...).
// Limits (checked on the host): head_dim % 8 == 0, head_dim <= 128, S <= 2048, (2*S*(head_dim+8) + 8*(S+hd)*2) * 2 B of smem.
#include "vf_common.cuh"

#include <math.h>

namespace vf {

constexpr int AS_WARPS = 8;
constexpr int AS_ROWS = 32;          // query rows per CTA
constexpr int AS_MAX_KPL = 64;       // keys per lane (S <= 2048)

__global__ void __launch_bounds__(AS_WARPS * 32)
attention_small_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int S, int H, int hd, float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int ldk = hd + 8;                                   // padded row, elements
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sV = sK + (size_t)S * ldk;
  float* sQ = reinterpret_cast<float*>(sV + (size_t)S * ldk);   // [AS_WARPS][hd]
  float* sP = sQ + AS_WARPS * hd;                               // [AS_WARPS][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const long long ldq = 3ll * H * hd;
  const __nv_bfloat16* base = qkv + (long long)b * S * ldq + h * hd;
  pdl_wait();
  pdl_launch_dependents();

  const int vec_per_row = hd / 8;
  for (int i = threadIdx.x; i < S * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, c = (i % vec_per_row) * 8;
    *reinterpret_cast<uint4*>(sK + (size_t)r * ldk + c) = *reinterpret_cast<const uint4*>(base + (long long)r * ldq + (long long)H * hd + c);
    *reinterpret_cast<uint4*>(sV + (size_t)r * ldk + c) = *reinterpret_cast<const uint4*>(base + (long long)r * ldq + 2ll * H * hd + c);
  }
  __syncthreads();

  float* q = sQ + warp * hd;
  float* pw = sP + (size_t)warp * S;
  for (int r = blockIdx.y * AS_ROWS + warp; r < S && r < (blockIdx.y + 1) * AS_ROWS; r += AS_WARPS) {
    for (int d = lane; d < hd; d += 32) q[d] = __bfloat162float(base[(long long)r * ldq + d]);
    __syncwarp();
    float sc[AS_MAX_KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < AS_MAX_KPL; ++i) {
      const int k = lane + 32 * i;
      sc[i] = -INFINITY;
      if (k < S) {
        float acc = 0.f;
        const __nv_bfloat16* kr = sK + (size_t)k * ldk;
        for (int c = 0; c < hd; c += 8) {
          const uint4 t = *reinterpret_cast<const uint4*>(kr + c);
          acc = fmaf(q[c + 0], bf16_lo(t.x), acc); acc = fmaf(q[c + 1], bf16_hi(t.x), acc);
          acc = fmaf(q[c + 2], bf16_lo(t.y), acc); acc = fmaf(q[c + 3], bf16_hi(t.y), acc);
          acc = fmaf(q[c + 4], bf16_lo(t.z), acc); acc = fmaf(q[c + 5], bf16_hi(t.z), acc);
          acc = fmaf(q[c + 6], bf16_lo(t.w), acc); acc = fmaf(q[c + 7], bf16_hi(t.w), acc);
        }
        sc[i] = acc * scale_log2;
        mx = fmaxf(mx, sc[i]);
      }
      if (32 * (i + 1) >= S) break;
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < AS_MAX_KPL; ++i) {
      const int k = lane + 32 * i;
      if (k < S) {
        const float p = exp2f(sc[i] - mx);
        sum += p;
        pw[k] = p;
      }
      if (32 * (i + 1) >= S) break;
    }
    const float inv = 1.0f / warp_sum(sum);
    __syncwarp();
    __nv_bfloat16* o = out + ((long long)b * S + r) * ((long long)H * hd) + h * hd;
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int k = 0; k < S; ++k) acc = fmaf(pw[k], __bfloat162float(sV[(size_t)k * ldk + d]), acc);
      o[d] = __float2bfloat16_rn(acc * inv);
    }
    __syncwarp();
  }
}

}  // namespace vf

using namespace vf;

// Called by vf_attention_fwd_hd for head dims other than 64.
int vf_attention_small_launch(const void* qkv, void* out, int B, int S, int H, int hd, float scale, cudaStream_t stream) {
  VF_REQUIRE(hd % 8 == 0 && hd >= 8 && hd <= 128, VF_ERR_ARG, "vf_attention_fwd_hd: head_dim %d unsupported (multiple of 8 up to 128; 64 = tensor-core kernel)", hd);
  VF_REQUIRE(S <= 32 * AS_MAX_KPL, VF_ERR_ARG, "vf_attention_fwd_hd: S=%d too long for the CUDA-core kernel (head_dim %d, S <= %d)", S, hd, 32 * AS_MAX_KPL);
  const size_t smem = (size_t)2 * S * (hd + 8) * 2 + (size_t)AS_WARPS * (hd + S) * 4;
  VF_REQUIRE(smem <= 200 * 1024, VF_ERR_ARG, "vf_attention_fwd_hd: S=%d x head_dim %d needs %zu B of shared memory (limit 200 KB)", S, hd, smem);
  static std::atomic<uint64_t> configured{0};
  if (int e = ensure_dynamic_smem(attention_small_kernel, 200 * 1024, configured)) return e;
  dim3 grid(B * H, (S + AS_ROWS - 1) / AS_ROWS);
  VF_CUDA(launch_pdl(attention_small_kernel, grid, dim3(AS_WARPS * 32), smem, stream, 1, static_cast<const __nv_bfloat16*>(qkv),
                     static_cast<__nv_bfloat16*>(out), S, H, hd, scale * 1.4426950408889634f));
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
