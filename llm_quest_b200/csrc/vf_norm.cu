// vf_norm.cu — LayerNorm kernels (HBM-bound): one warp per row, 128-bit loads, fp32 statistics.
//
// variant 0: nn.LayerNorm semantics      (x-mean)/sqrt(var+eps)*w+b
//            reference: llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:213-214,229,234,406,422
// variant 1: Part-1 LayerNorm semantics   (x-mean)/(std+eps)*w+b   (eps added to the std)
//            reference: llm_quest/multimodal/vision_transformer/vit_transformer_block.py:21-31
// merge > 1 fuses ViTMergeAdapter's view/permute/contiguous (vision_model.py:425-427) into the
// store: the normalised row of token (f, r, c) lands in merged row ((f*(nh/m)+r/m)*(nw/m)+c/m) at
// feature slot (r%m)*m + c%m, so the 2x2 gather costs no extra pass over HBM.
//
// Algorithmic traffic: rows*D*(in_bytes + out_bytes); the row lives in registers between the two
// statistics passes, so HBM sees each element once in and once out.
#include "vf_common.cuh"

namespace vf {

constexpr int LN_WARPS = 8;  // rows per CTA

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
  }
};

// NV = D / 128 (each lane owns NV groups of 4 consecutive elements, group g at column (g*32+lane)*4)
template <typename TIn, typename TOut, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_kernel(const TIn* __restrict__ x, long long ldx, const float* __restrict__ w,
                 const float* __restrict__ b, TOut* __restrict__ out, long long rows, float eps,
                 int variant, int merge, int nh, int nw, float* __restrict__ mean_out) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  pdl_wait();
  pdl_launch_dependents();
  if (row >= rows) return;

  float v[NV][4];
  const TIn* xr = x + row * ldx;
#pragma unroll
  for (int g = 0; g < NV; ++g) Vec4<TIn>::load(xr + (g * 32 + lane) * 4, v[g]);

  float s = 0.f;
#pragma unroll
  for (int g = 0; g < NV; ++g) s += (v[g][0] + v[g][1]) + (v[g][2] + v[g][3]);
  const float mean = warp_sum(s) * (1.0f / D);
  if (mean_out != nullptr && lane == 0) mean_out[row] = mean;   // the folded-LayerNorm producers' first row shift
  float q = 0.f;
#pragma unroll
  for (int g = 0; g < NV; ++g) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d = v[g][e] - mean;
      q = fmaf(d, d, q);
    }
  }
  const float var = warp_sum(q) * (1.0f / D);
  const float rstd = variant == 0 ? rsqrtf(var + eps) : 1.0f / (sqrtf(var) + eps);

  long long orow = row;
  int slot = 0;
  if (merge > 1) {
    const int n = nh * nw;
    const long long fidx = row / n;            // global frame index (sample-major)
    const int sp = static_cast<int>(row - fidx * n);
    const int r = sp / nw, c = sp - r * nw;
    orow = (fidx * (nh / merge) + r / merge) * (nw / merge) + c / merge;
    slot = (r % merge) * merge + (c % merge);
  }
  TOut* o = out + (merge > 1 ? (orow * (merge * merge) + slot) * D : orow * static_cast<long long>(D));
#pragma unroll
  for (int g = 0; g < NV; ++g) {
    const int col = (g * 32 + lane) * 4;
    const float4 w4 = *reinterpret_cast<const float4*>(w + col);
    const float4 b4 = *reinterpret_cast<const float4*>(b + col);
    float y[4];
    y[0] = (v[g][0] - mean) * rstd * w4.x + b4.x;
    y[1] = (v[g][1] - mean) * rstd * w4.y + b4.y;
    y[2] = (v[g][2] - mean) * rstd * w4.z + b4.z;
    y[3] = (v[g][3] - mean) * rstd * w4.w + b4.w;
    Vec4<TOut>::store(o + col, y);
  }
}

template <typename TIn, typename TOut>
static int launch_ln(const void* x, long long ldx, const float* w, const float* b, void* out,
                     long long rows, int D, float eps, int variant, int merge, int nh, int nw,
                     float* mean_out, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((rows + LN_WARPS - 1) / LN_WARPS);
  const TIn* xi = static_cast<const TIn*>(x);
  TOut* o = static_cast<TOut*>(out);
#define VF_LN_CASE(NV)                                                                            \
  case NV:                                                                                        \
    VF_CUDA(launch_pdl(layernorm_kernel<TIn, TOut, NV>, dim3(grid), dim3(LN_WARPS * 32), 0, s, 1, xi, \
                       static_cast<long long>(ldx), w, b, o, static_cast<long long>(rows), eps, variant, merge, nh, nw, mean_out)); \
    break;
  switch (D / 128) {
    VF_LN_CASE(1) VF_LN_CASE(2) VF_LN_CASE(3) VF_LN_CASE(4) VF_LN_CASE(6) VF_LN_CASE(8)
    VF_LN_CASE(10) VF_LN_CASE(12) VF_LN_CASE(16)
    default:
      set_last_error("vf_layernorm: D=%d unsupported (need D/128 in {1,2,3,4,6,8,10,12,16})", D);
      return VF_ERR_ARG;
  }
#undef VF_LN_CASE
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

__global__ void vit_cls_pos_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                                   float* __restrict__ out, long long rows_per_sample, int D) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    out[static_cast<long long>(b) * rows_per_sample * D + d] = cls[d] + pos[d];
}

// One block = 32 rows x 8 groups of partials: thread (g, r) adds the partials j = g, g+8, ... of row r (coalesced 256-byte
// rows of float2), the 8 group sums meet in shared memory and warp 0 adds them in fixed order.
__global__ void __launch_bounds__(256) ln_row_stats_kernel(const float2* __restrict__ part, int parts, long long ld,
                                                           long long rows, float inv_d, float eps, int variant,
                                                           float2* __restrict__ out, float* __restrict__ shift) {
  __shared__ float2 acc[8][32];
  pdl_wait();
  pdl_launch_dependents();
  const int g = threadIdx.x >> 5, l = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 32 + l;
  float s = 0.f, q = 0.f;
  if (r < rows) {
#pragma unroll 4
    for (int j = g; j < parts; j += 8) {
      const float2 t = __ldg(part + (long long)j * ld + r);
      s += t.x;
      q += t.y;
    }
  }
  acc[g][l] = make_float2(s, q);
  __syncthreads();
  if (g == 0 && r < rows) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      s += acc[k][l].x;
      q += acc[k][l].y;
    }
    const float mu = s * inv_d;   // mean of the SHIFTED row: the consumer GEMM multiplies the shifted bf16 copy
    const float var = fmaxf(fmaf(-mu, mu, q * inv_d), 0.f);
    out[r] = make_float2(mu, variant == 0 ? rsqrtf(var + eps) : 1.0f / (sqrtf(var) + eps));
    if (shift != nullptr) shift[r] += mu;   // the row's true mean: what the next producer subtracts before rounding
  }
}

}  // namespace vf

using namespace vf;

extern "C" int vf_ln_row_stats(const void* partials, int32_t parts, int64_t ld, int64_t rows, int32_t D, float eps,
                               int32_t variant, void* out, float* shift, void* stream) {
  VF_REQUIRE(partials && out && parts > 0 && rows > 0 && D > 0 && ld >= rows && (variant == 0 || variant == 1), VF_ERR_ARG,
             "vf_ln_row_stats: bad arguments");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(partials) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, VF_ERR_ALIGN,
             "vf_ln_row_stats: pointers must be 8-byte aligned");
  const unsigned grid = static_cast<unsigned>((rows + 31) / 32);
  VF_CUDA(launch_pdl(ln_row_stats_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                     static_cast<const float2*>(partials), (int)parts, (long long)ld, (long long)rows,
                     1.0f / static_cast<float>(D), eps, (int)variant, static_cast<float2*>(out), shift));
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_layernorm(const void* x, int32_t in_dtype, int64_t ldx, const float* w,
                            const float* b, void* out, int32_t out_dtype, int64_t rows, int32_t D,
                            float eps, int32_t variant, int32_t merge, int32_t nh, int32_t nw,
                            float* mean_out, void* stream) {
  VF_REQUIRE(x && w && b && out, VF_ERR_ARG, "vf_layernorm: null pointer");
  VF_REQUIRE(rows > 0 && D > 0 && D % 128 == 0, VF_ERR_ARG, "vf_layernorm: rows=%lld D=%d (D must be a multiple of 128)",
             (long long)rows, D);
  VF_REQUIRE(variant == 0 || variant == 1, VF_ERR_ARG, "vf_layernorm: variant must be 0 or 1");
  VF_REQUIRE(ldx >= D && ldx % 4 == 0, VF_ERR_ALIGN, "vf_layernorm: ldx must be >= D and a multiple of 4");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0,
             VF_ERR_ALIGN, "vf_layernorm: pointers must be 16-byte aligned");
  if (merge > 1) {
    VF_REQUIRE(nh > 0 && nw > 0 && nh % merge == 0 && nw % merge == 0 && rows % ((int64_t)nh * nw) == 0,
               VF_ERR_ARG, "vf_layernorm: merge=%d needs nh,nw divisible by merge and rows %% (nh*nw) == 0", merge);
  } else {
    merge = 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in_dtype == 0 && out_dtype == 1)
    return launch_ln<float, __nv_bfloat16>(x, ldx, w, b, out, rows, D, eps, variant, merge, nh, nw, mean_out, s);
  if (in_dtype == 0 && out_dtype == 0)
    return launch_ln<float, float>(x, ldx, w, b, out, rows, D, eps, variant, merge, nh, nw, mean_out, s);
  if (in_dtype == 1 && out_dtype == 1)
    return launch_ln<__nv_bfloat16, __nv_bfloat16>(x, ldx, w, b, out, rows, D, eps, variant, merge, nh, nw, mean_out, s);
  if (in_dtype == 1 && out_dtype == 0)
    return launch_ln<__nv_bfloat16, float>(x, ldx, w, b, out, rows, D, eps, variant, merge, nh, nw, mean_out, s);
  set_last_error("vf_layernorm: dtype codes must be 0 (fp32) or 1 (bf16)");
  return VF_ERR_ARG;
}

extern "C" int vf_vit_cls_pos(const float* cls, const float* pos, float* out, int32_t B,
                              int64_t rows_per_sample, int32_t D, void* stream) {
  VF_REQUIRE(cls && pos && out && B > 0 && D > 0, VF_ERR_ARG, "vf_vit_cls_pos: bad arguments");
  vit_cls_pos_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(cls, pos, out, rows_per_sample, D);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
