// vf_elementwise.cu — the small HBM-bound pieces of the module surface that have no GEMM to ride on.
//
//   vf_gelu            stand-alone GELU (erf or tanh form): drop-in for GELU.forward of the Part-1 ViT
//                      (llm_quest/multimodal/vision_transformer/vit_transformer_block.py:43-44) and nn.GELU modules called
//                      on their own; on the path proper the activation lives in a GEMM epilogue.
//   vf_rmsnorm_zc      ZeroCenteredRMSNorm.forward (llm_quest/qwen/qwen3_next/qwen3_next_attention.py:41-46), fp32 inside.
//   vf_embed_pos_concat   Part-2 text half of the fusion: get_embeddings (token + position embedding,
//                      llm_quest/multimodal/vlm_engine.py:5-20) written straight into rows row_off.. of the fused
//                      [b, n_vision + seq, D] buffer (the torch.cat of vlm_engine.py:114 / vlm_generation.py:66).
//   vf_im2col_patches  non-overlapping P x P patches -> bf16 rows [B*n, C*P*P] for patch sizes the TMA gather GEMM is not
//                      built for (TINY_VIT_CONFIG: 4x4 patches, config.py:175-186).
//   vf_fill_rows_f32   out[b, r, :] = src[r, :] (+ add0[:] on row 0): class-token / position-embedding rows of the
//                      Part-1 ViT for every row of every sample (vit_model.py:86-87,145).
//
// All kernels: one warp per row (or grid-stride over 16-byte vectors), 128-bit accesses, fp32 math.
#include "vf_common.cuh"

#include <math.h>

namespace vf {

// ------------------------------------------------------------------------------------------------ GELU
template <typename T>
struct V8;   // 16 bytes of T as floats
template <>
struct V8<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct V8<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
    v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
};

__device__ __forceinline__ float gelu_exact(float x, int tanh_form) {
  if (tanh_form) {
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
  }
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752f));
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_kernel(const T* __restrict__ x, T* __restrict__ out, long long n, int tanh_form) {
  constexpr int N = V8<T>::N;
  const long long nvec = n / N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[N];
    V8<T>::load(x + i * N, v);
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = gelu_exact(v[e], tanh_form);
    V8<T>::store(out + i * N, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < n - nvec * N) {   // ragged tail, scalar
    const long long i = nvec * N + threadIdx.x;
    if constexpr (std::is_same<T, float>::value) out[i] = gelu_exact(x[i], tanh_form);
    else out[i] = __float2bfloat16_rn(gelu_exact(__bfloat162float(x[i]), tanh_form));
  }
}

// ------------------------------------------------------------------------------------------------ zero-centred RMSNorm
// one warp per row; D % 8 == 0 for bf16 (D % 4 for fp32); (x * rsqrt(mean(x^2) + eps)) * w with w = 1 + scale given in fp32
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_zc_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                         T* __restrict__ out, long long ldo, long long rows, int D, float eps) {
  constexpr int N = V8<T>::N;
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ldx;
  float ss = 0.f;
  for (int c = lane * N; c < D; c += 32 * N) {
    float v[N];
    V8<T>::load(xr + c, v);
#pragma unroll
    for (int e = 0; e < N; ++e) ss = fmaf(v[e], v[e], ss);
  }
  const float r = rsqrtf(warp_sum(ss) / static_cast<float>(D) + eps);
  T* orow = out + row * ldo;
  for (int c = lane * N; c < D; c += 32 * N) {   // second read hits L1/L2: the row was just touched by this warp
    float v[N];
    V8<T>::load(xr + c, v);
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = (v[e] * r) * __ldg(w + c + e);
    V8<T>::store(orow + c, v);
  }
}

// ------------------------------------------------------------------------------------------------ token + position embedding rows
template <typename T>
__global__ void __launch_bounds__(256) embed_pos_concat_kernel(const long long* __restrict__ ids, const T* __restrict__ tok,
                                                               long long vocab, const T* __restrict__ pos, float* __restrict__ out,
                                                               int b, int seq, int D, long long rows_per_sample, long long row_off) {
  constexpr int N = V8<T>::N;
  const int lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (r >= (long long)b * seq) return;
  const int t = static_cast<int>(r % seq);
  const long long bi = r / seq;
  long long id = ids[r];
  if (id < 0 || id >= vocab) id = 0;   // nn.Embedding would raise; the host mirror checks the range when asked to
  const T* tr = tok + id * D;
  const T* pr = pos + (long long)t * D;
  float* o = out + (bi * rows_per_sample + row_off + t) * D;
  for (int c = lane * N; c < D; c += 32 * N) {
    float a[N], p2[N];
    V8<T>::load(tr + c, a);
    V8<T>::load(pr + c, p2);
#pragma unroll
    for (int e = 0; e < N; ++e) a[e] += p2[e];
    if constexpr (N == 4) {
      *reinterpret_cast<float4*>(o + c) = make_float4(a[0], a[1], a[2], a[3]);
    } else {
      // the sum of two bf16 values is rounded to bf16 by the reference (bf16 + bf16), then widened
      *reinterpret_cast<float4*>(o + c) = make_float4(bf16_lo(pack_bf16(a[0], 0.f)), bf16_lo(pack_bf16(a[1], 0.f)),
                                                      bf16_lo(pack_bf16(a[2], 0.f)), bf16_lo(pack_bf16(a[3], 0.f)));
      *reinterpret_cast<float4*>(o + c + 4) = make_float4(bf16_lo(pack_bf16(a[4], 0.f)), bf16_lo(pack_bf16(a[5], 0.f)),
                                                          bf16_lo(pack_bf16(a[6], 0.f)), bf16_lo(pack_bf16(a[7], 0.f)));
    }
  }
}

// ------------------------------------------------------------------------------------------------ im2col for small patches
// pixels [B, C, H, W] (fp32 or bf16) -> rows [B * nh * nw, C*P*P] bf16, column order (c, py, px) = the conv weight flattened
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int C, int H,
                                                     int W, int P, long long ld_out) {
  const int nh = H / P, nw = W / P, K = C * P * P;
  const long long total = (long long)B * nh * nw * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const long long row = i / K;
    const int pw = static_cast<int>(row % nw);
    const int ph = static_cast<int>((row / nw) % nh);
    const long long bi = row / ((long long)nw * nh);
    const int px = k % P, py = (k / P) % P, c = k / (P * P);
    const T v = x[((bi * C + c) * H + ph * P + py) * W + pw * P + px];
    float f;
    if constexpr (std::is_same<T, float>::value) f = v;
    else f = __bfloat162float(v);
    out[row * ld_out + k] = __float2bfloat16_rn(f);
  }
}

__global__ void __launch_bounds__(256) fill_rows_kernel(const float* __restrict__ src, const float* __restrict__ add0,
                                                        float* __restrict__ out, long long rows_per_sample, int D, int B) {
  const long long total = (long long)B * rows_per_sample * (D / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % (D / 4)) * 4;
    const long long r = (i / (D / 4)) % rows_per_sample;
    float4 v = *reinterpret_cast<const float4*>(src + r * D + c);
    if (r == 0 && add0) {
      const float4 a = *reinterpret_cast<const float4*>(add0 + c);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(out + (i / (D / 4)) * D + c) = v;
  }
}

static unsigned grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = 148ll * 16;
  return static_cast<unsigned>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace vf

using namespace vf;

extern "C" int vf_gelu(const void* x, void* out, int32_t dtype, int64_t n, int32_t tanh_form, void* stream) {
  VF_REQUIRE(x && out && n > 0, VF_ERR_ARG, "vf_gelu: bad arguments");
  VF_REQUIRE(dtype == 0 || dtype == 1, VF_ERR_ARG, "vf_gelu: dtype codes must be 0 (fp32) or 1 (bf16)");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, VF_ERR_ALIGN,
             "vf_gelu: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == 0)
    gelu_kernel<float><<<grid_for(n / 4, 256), 256, 0, s>>>(static_cast<const float*>(x), static_cast<float*>(out), n, tanh_form);
  else
    gelu_kernel<__nv_bfloat16><<<grid_for(n / 8, 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x),
                                                                     static_cast<__nv_bfloat16*>(out), n, tanh_form);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_rmsnorm_zc(const void* x, int64_t ldx, const float* one_plus_scale, void* out, int64_t ldo, int32_t dtype,
                             int64_t rows, int32_t D, float eps, void* stream) {
  VF_REQUIRE(x && out && one_plus_scale && rows > 0 && D > 0, VF_ERR_ARG, "vf_rmsnorm_zc: bad arguments");
  VF_REQUIRE(dtype == 0 || dtype == 1, VF_ERR_ARG, "vf_rmsnorm_zc: dtype codes must be 0 (fp32) or 1 (bf16)");
  const int vec = dtype == 0 ? 4 : 8;
  VF_REQUIRE(D % vec == 0 && ldx % vec == 0 && ldo % vec == 0 && ldx >= D && ldo >= D, VF_ERR_ALIGN,
             "vf_rmsnorm_zc: D and the row pitches must be multiples of %d elements", vec);
  VF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, VF_ERR_ALIGN,
             "vf_rmsnorm_zc: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (dtype == 0)
    rmsnorm_zc_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), ldx, one_plus_scale, static_cast<float*>(out), ldo, rows, D, eps);
  else
    rmsnorm_zc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, one_plus_scale,
                                                           static_cast<__nv_bfloat16*>(out), ldo, rows, D, eps);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_embed_pos_concat(const int64_t* input_ids, const void* tok_table, int64_t vocab, const void* pos_table,
                                   int64_t n_pos, int32_t table_dtype, float* out, int32_t b, int32_t seq, int32_t D,
                                   int64_t rows_per_sample, int64_t row_off, void* stream) {
  VF_REQUIRE(input_ids && tok_table && pos_table && out && b > 0 && seq > 0 && D > 0, VF_ERR_ARG, "vf_embed_pos_concat: bad arguments");
  VF_REQUIRE(table_dtype == 0 || table_dtype == 1, VF_ERR_ARG, "vf_embed_pos_concat: dtype codes must be 0 (fp32) or 1 (bf16)");
  VF_REQUIRE(seq <= n_pos, VF_ERR_ARG, "vf_embed_pos_concat: seq %d exceeds the %lld position embeddings", seq, (long long)n_pos);
  VF_REQUIRE(row_off >= 0 && row_off + seq <= rows_per_sample, VF_ERR_ARG, "vf_embed_pos_concat: rows [row_off, row_off + seq) leave the sample");
  VF_REQUIRE(D % 8 == 0 && (reinterpret_cast<uintptr_t>(tok_table) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos_table) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             VF_ERR_ALIGN, "vf_embed_pos_concat: D %% 8 == 0 and 16-byte aligned pointers");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(((long long)b * seq + 7) / 8);
  if (table_dtype == 0)
    embed_pos_concat_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const long long*>(input_ids), static_cast<const float*>(tok_table), vocab,
                                                        static_cast<const float*>(pos_table), out, b, seq, D, rows_per_sample, row_off);
  else
    embed_pos_concat_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const long long*>(input_ids),
                                                                 static_cast<const __nv_bfloat16*>(tok_table), vocab,
                                                                 static_cast<const __nv_bfloat16*>(pos_table), out, b, seq, D,
                                                                 rows_per_sample, row_off);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_im2col_patches(const void* pixels, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W, int32_t P, void* out,
                                 int64_t ld_out, void* stream) {
  VF_REQUIRE(pixels && out && B > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0, VF_ERR_ARG, "vf_im2col_patches: bad arguments");
  VF_REQUIRE(dtype == 0 || dtype == 1, VF_ERR_ARG, "vf_im2col_patches: dtype codes must be 0 (fp32) or 1 (bf16)");
  VF_REQUIRE(ld_out >= (int64_t)C * P * P, VF_ERR_ARG, "vf_im2col_patches: ld_out smaller than C*P*P");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long total = (long long)B * (H / P) * (W / P) * C * P * P;
  if (dtype == 0)
    im2col_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(static_cast<const float*>(pixels), static_cast<__nv_bfloat16*>(out), B, C, H, W, P, ld_out);
  else
    im2col_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(pixels),
                                                                      static_cast<__nv_bfloat16*>(out), B, C, H, W, P, ld_out);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_fill_rows_f32(const float* src, const float* add_row0, float* out, int32_t B, int64_t rows_per_sample, int32_t D,
                                void* stream) {
  VF_REQUIRE(src && out && B > 0 && rows_per_sample > 0 && D > 0 && D % 4 == 0, VF_ERR_ARG, "vf_fill_rows_f32: bad arguments");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(add_row0) & 15) == 0,
             VF_ERR_ALIGN, "vf_fill_rows_f32: pointers must be 16-byte aligned");
  const long long total = (long long)B * rows_per_sample * (D / 4);
  fill_rows_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, add_row0, out, rows_per_sample, D, B);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
