// vf_preprocess.cu — uint8 image -> normalised [B, C, T, H, W] pixel tensor (HBM-bound, one pass).
//
// Replaces the host-side pre-processing that feeds PatchEmbedding3D in the reference:
//   to_tensor (HWC uint8 -> CHW float / 255), normalize ((x - mean) / std), duplicate the frame along the
//   temporal axis and permute to (B, C, T, H, W)
//   (llm_quest/qwen/qwen3_5/qwen3_5_generate_multimodal.py:40-46; dataset.py:336-351 for the ViT datasets).
// The resize before it is PIL's antialiased bilinear filter and stays on the host.
//
// fp32 arithmetic in the reference's order (divide by 255, subtract, divide: three IEEE-rounded steps), so an
// fp32 output is bit-identical to torchvision's; a bf16 output is that value rounded once.
// One thread converts 4 consecutive pixels of a row: 12 input bytes, and per channel and temporal copy one
// 8-byte (bf16) or 16-byte (fp32) store. Algorithmic traffic per pixel: 3 B in, C*T*sizeof(out) out.
#include "vf_common.cuh"

namespace vf {

template <typename OutT>
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ img, OutT* __restrict__ out, long long n_quads, int H, int W, int T,
                     float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // quad of 4 pixels
  if (q >= n_quads) return;
  const int wq = W >> 2;
  const long long row = q / wq;                 // b*H + y
  const int x0 = static_cast<int>(q - row * wq) * 4;
  const long long b = row / H;
  const int y = static_cast<int>(row - b * H);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(img + (row * W + x0) * 3);   // 12 bytes, 4-byte aligned
  const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
  const uint8_t px[12] = {
      (uint8_t)(w0), (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
      (uint8_t)(w1), (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
      (uint8_t)(w2), (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  const long long plane = static_cast<long long>(H) * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      v[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(px[3 * i + c]), 255.0f), mean[c]), sd[c]);
    OutT* dst = out + ((b * 3 + c) * T) * plane + static_cast<long long>(y) * W + x0;
    for (int t = 0; t < T; ++t) {
      if constexpr (sizeof(OutT) == 4)
        *reinterpret_cast<float4*>(dst + t * plane) = make_float4(v[0], v[1], v[2], v[3]);
      else
        *reinterpret_cast<uint2*>(dst + t * plane) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
  }
}

}  // namespace vf

using namespace vf;

extern "C" int vf_preprocess_u8(const uint8_t* img, int32_t B, int32_t H, int32_t W, int32_t T, const float* mean3,
                                const float* std3, void* out, int32_t out_dtype, void* stream) {
  VF_REQUIRE(img && out && mean3 && std3, VF_ERR_ARG, "vf_preprocess_u8: null pointer");
  VF_REQUIRE(B > 0 && H > 0 && W > 0 && T > 0, VF_ERR_ARG, "vf_preprocess_u8: bad shape B=%d H=%d W=%d T=%d", B, H, W, T);
  VF_REQUIRE((W & 3) == 0, VF_ERR_ALIGN, "vf_preprocess_u8: image width must be a multiple of 4 pixels");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(img) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, VF_ERR_ALIGN,
             "vf_preprocess_u8: image must be 4-byte aligned, output 16-byte aligned");
  VF_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, VF_ERR_ARG, "vf_preprocess_u8: zero std");
  const long long n_quads = static_cast<long long>(B) * H * (W / 4);
  const unsigned grid = static_cast<unsigned>((n_quads + 255) / 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_dtype == 0)
    preprocess_u8_kernel<float><<<grid, 256, 0, s>>>(img, reinterpret_cast<float*>(out), n_quads, H, W, T, mean3[0],
                                                    mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  else if (out_dtype == 1)
    preprocess_u8_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(img, reinterpret_cast<__nv_bfloat16*>(out), n_quads, H, W, T,
                                                            mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  else {
    set_last_error("vf_preprocess_u8: out_dtype must be 0 (fp32) or 1 (bf16)");
    return VF_ERR_ARG;
  }
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
