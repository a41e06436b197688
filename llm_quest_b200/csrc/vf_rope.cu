// vf_rope.cu — rotary-embedding kernels (HBM-bound, 128-bit accesses).
//
//  * rope_apply_kernel: rotate-half RoPE with a [rows, rot] cos/sin table, positions = s or
//    position_ids[b, s]. Drop-in for VisionRoPE.apply / RoPE.apply
//    (reference llm_quest/common/rope.py:180-243, 485-500).
//  * mrope_apply_kernel: MRoPE-I. Half-dim slot j takes its angle from position axis
//        H if j % 3 == 1 and j < 3*sec_h,  W if j % 3 == 2 and j < 3*sec_w,  else T
//    (reference rope.py:283-294 builds exactly this by strided overwrites), gathers cos/sin at that
//    axis' position id, rotates the first `rot` dims and passes the rest through (rope.py:331-358).
//    Optionally fuses the zero-centred RMSNorm that precedes it in the text model
//    (qwen3_next_attention.py:41-46, call site qwen3_5_text_model.py:227-233).
//
// For fp32 tensors the arithmetic order mirrors the reference (two rounded products, one rounded
// sum) so results are bit-identical; bf16 tensors are computed in fp32 and rounded once.
// Algorithmic traffic: one read + one write of x (cos/sin rows are L2/L1 resident).
#include "vf_common.cuh"

namespace vf {

template <typename T>
struct Ld4;
template <>
struct Ld4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Ld4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
  }
};

// One thread = one unit of 4 elements: either a rotation pair-quad (i..i+3 and i+half..i+half+3) or
// a pass-through quad. units_per_row = rot/8 + (hd-rot)/4.
template <typename T>
__global__ void __launch_bounds__(256)
rope_apply_kernel(const T* __restrict__ x, T* __restrict__ out, long long n_rows, int H, int S, int hd,
                  const float* __restrict__ cos, const float* __restrict__ sin, int rot,
                  const long long* __restrict__ position_ids) {
  const int half = rot >> 1;
  const int pair_units = half >> 2;
  const int units = pair_units + ((hd - rot) >> 2);
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= n_rows * units) return;
  const long long row = gid / units;          // (b*H + h)*S + s
  const int u = static_cast<int>(gid - row * units);
  const T* xr = x + row * hd;
  T* orow = out + row * hd;
  if (u >= pair_units) {
    const int c = rot + (u - pair_units) * 4;
    float v[4];
    Ld4<T>::ld(xr + c, v);
    Ld4<T>::st(orow + c, v);
    return;
  }
  const int s = static_cast<int>(row % S);
  const long long b = row / (static_cast<long long>(S) * H);
  const long long pos = position_ids ? position_ids[b * S + s] : s;
  const int i = u * 4;
  float x1[4], x2[4], c1[4], c2[4], s1[4], s2[4], o1[4], o2[4];
  Ld4<T>::ld(xr + i, x1);
  Ld4<T>::ld(xr + i + half, x2);
  Ld4<float>::ld(cos + pos * rot + i, c1);
  Ld4<float>::ld(cos + pos * rot + i + half, c2);
  Ld4<float>::ld(sin + pos * rot + i, s1);
  Ld4<float>::ld(sin + pos * rot + i + half, s2);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    // cos*x + sin*rotate_half(x), each product rounded before the sum (as the reference does)
    o1[e] = __fsub_rn(__fmul_rn(c1[e], x1[e]), __fmul_rn(s1[e], x2[e]));
    o2[e] = __fadd_rn(__fmul_rn(c2[e], x2[e]), __fmul_rn(s2[e], x1[e]));
  }
  Ld4<T>::st(orow + i, o1);
  Ld4<T>::st(orow + i + half, o2);
}

// EPL consecutive elements of one lane kept in their storage format (so that eight rows in flight cost 32
// registers at bf16 / head_dim 256), moved with the widest access the size allows.
template <typename T, int EPL>
struct LaneRaw {
  static constexpr int NU = EPL * sizeof(T) / 8;   // 8-byte units: 1, 2 or 4
  uint2 u[NU];
  __device__ __forceinline__ void load(const T* p) {
    if constexpr (NU == 1) {
      u[0] = *reinterpret_cast<const uint2*>(p);
    } else {
#pragma unroll
      for (int q = 0; q < NU / 2; ++q) {
        const uint4 t = reinterpret_cast<const uint4*>(p)[q];
        u[2 * q] = make_uint2(t.x, t.y);
        u[2 * q + 1] = make_uint2(t.z, t.w);
      }
    }
  }
  __device__ __forceinline__ void store(T* p) const {
    if constexpr (NU == 1) {
      *reinterpret_cast<uint2*>(p) = u[0];
    } else {
#pragma unroll
      for (int q = 0; q < NU / 2; ++q)
        reinterpret_cast<uint4*>(p)[q] = make_uint4(u[2 * q].x, u[2 * q].y, u[2 * q + 1].x, u[2 * q + 1].y);
    }
  }
  __device__ __forceinline__ void unpack(float (&v)[EPL]) const {
    if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int q = 0; q < NU; ++q) {
        v[4 * q] = bf16_lo(u[q].x); v[4 * q + 1] = bf16_hi(u[q].x);
        v[4 * q + 2] = bf16_lo(u[q].y); v[4 * q + 3] = bf16_hi(u[q].y);
      }
    } else {
#pragma unroll
      for (int q = 0; q < NU; ++q) { v[2 * q] = __uint_as_float(u[q].x); v[2 * q + 1] = __uint_as_float(u[q].y); }
    }
  }
  __device__ __forceinline__ void pack(const float (&v)[EPL]) {
    if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int q = 0; q < NU; ++q) u[q] = make_uint2(pack_bf16(v[4 * q], v[4 * q + 1]), pack_bf16(v[4 * q + 2], v[4 * q + 3]));
    } else {
#pragma unroll
      for (int q = 0; q < NU; ++q) u[q] = make_uint2(__float_as_uint(v[2 * q]), __float_as_uint(v[2 * q + 1]));
    }
  }
};

// One warp per TOKEN (b, s): the position ids and the gathered cos/sin coefficients depend on the token
// only, so they are fetched once and reused for every head; the heads' rows (512 B each at head_dim 256
// bf16) are loaded eight at a time before any arithmetic, which is what keeps enough bytes in flight
// (one row per warp with the id -> table -> x dependency chain reached only 22 % of the copy bandwidth).
// Lane l owns elements [l*EPL, (l+1)*EPL) of a row, EPL = hd/32 (4 or 8).
template <typename T, int EPL>
__global__ void __launch_bounds__(256)
mrope_apply_kernel(const T* x, T* out, int B, int H, int S, long long xs_b, long long xs_h, long long xs_s,
                   long long os_b, long long os_h, long long os_s, const float* __restrict__ cos, const float* __restrict__ sin, int rot,
                   const long long* __restrict__ position_ids, int sec_h, int sec_w,
                   const float* __restrict__ norm_w, float norm_eps) {
  constexpr int HD = EPL * 32;
  constexpr int HC = sizeof(T) == 2 ? 8 : 4;   // heads (rows) in flight per warp
  const int lane = threadIdx.x & 31;
  const long long tok = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= static_cast<long long>(B) * S) return;
  const int s = static_cast<int>(tok % S);
  const long long b = tok / S;

  const int half = rot >> 1;
  const int lanes_half = half / EPL;  // lanes holding the first half of the rotated block
  const bool rotates = lane < 2 * lanes_half;
  // out = c*v + sg*partner with sg = -sin for the first half and +sin for the second: the same two rounded
  // products and one rounded sum as the reference's cos*x + sin*rotate_half(x)
  float c[EPL], sg[EPL], w[EPL];
  if (rotates) {
    const bool first = lane < lanes_half;
    const long long pt = position_ids[(0LL * B + b) * S + s];
    const long long ph = position_ids[(1LL * B + b) * S + s];
    const long long pw = position_ids[(2LL * B + b) * S + s];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int j = (lane * EPL + e) % half;  // half-dim slot
      const int r3 = j % 3;
      long long pos = pt;
      if (r3 == 1 && j < 3 * sec_h) pos = ph;
      if (r3 == 2 && j < 3 * sec_w) pos = pw;
      c[e] = __ldg(cos + pos * rot + j);
      const float sn = __ldg(sin + pos * rot + j);
      sg[e] = first ? -sn : sn;
    }
  }
  if (norm_w) {
#pragma unroll
    for (int q = 0; q < EPL / 4; ++q) Ld4<float>::ld(norm_w + lane * EPL + q * 4, *reinterpret_cast<float(*)[4]>(&w[q * 4]));
  }

  // element strides (batch, head, token) of x and out: [B,H,S,hd] contiguous, or head columns of a token-major
  // projection output; x == out (in place) is fine, a row is read completely before it is written
  const long long xbase = b * xs_b + s * xs_s + lane * EPL;
  const long long obase = b * os_b + s * os_s + lane * EPL;
  for (int h0 = 0; h0 < H; h0 += HC) {
    LaneRaw<T, EPL> raw[HC];
#pragma unroll
    for (int i = 0; i < HC; ++i)
      if (h0 + i < H) raw[i].load(x + xbase + (h0 + i) * xs_h);
#pragma unroll
    for (int i = 0; i < HC; ++i) {
      if (h0 + i >= H) break;
      float v[EPL];
      raw[i].unpack(v);
      if (norm_w) {
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) ss = fmaf(v[e], v[e], ss);
        const float rms = rsqrtf(warp_sum(ss) * (1.0f / HD) + norm_eps);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          float y = v[e] * rms * w[e];
          if (sizeof(T) == 2) y = __bfloat162float(__float2bfloat16_rn(y));  // the norm returns x.dtype
          v[e] = y;
        }
      }
      // partner element (j <-> j+half) lives lanes_half lanes away
      float partner[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) partner[e] = __shfl_xor_sync(0xffffffffu, v[e], lanes_half);
      if (rotates) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          if (sizeof(T) == 4) v[e] = __fadd_rn(__fmul_rn(c[e], v[e]), __fmul_rn(sg[e], partner[e]));
          else v[e] = c[e] * v[e] + sg[e] * partner[e];
        }
      }
      raw[i].pack(v);
      raw[i].store(out + obase + (h0 + i) * os_h);
    }
  }
}

}  // namespace vf

using namespace vf;

extern "C" int vf_rope_apply(const void* x, void* out, int32_t dtype, int32_t B, int32_t H, int32_t S,
                             int32_t hd, const float* cos, const float* sin, int32_t rot,
                             int64_t table_rows, const int64_t* position_ids, void* stream) {
  VF_REQUIRE(x && out && cos && sin, VF_ERR_ARG, "vf_rope_apply: null pointer");
  VF_REQUIRE(B > 0 && H > 0 && S > 0 && hd > 0, VF_ERR_ARG, "vf_rope_apply: bad shape");
  VF_REQUIRE(rot > 0 && rot <= hd && rot % 8 == 0 && hd % 4 == 0, VF_ERR_ARG,
             "vf_rope_apply: need rot %% 8 == 0, rot <= hd, hd %% 4 == 0 (rot=%d hd=%d)", rot, hd);
  VF_REQUIRE(position_ids || table_rows >= S, VF_ERR_ARG,
             "vf_rope_apply: cos/sin table has %lld rows but seq_len is %d", (long long)table_rows, S);
  VF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(cos) & 15) == 0 && (reinterpret_cast<uintptr_t>(sin) & 15) == 0,
             VF_ERR_ALIGN, "vf_rope_apply: pointers must be 16-byte aligned");
  const long long rows = (long long)B * H * S;
  const int units = rot / 8 + (hd - rot) / 4;
  const long long total = rows * units;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long* pid = reinterpret_cast<const long long*>(position_ids);
  if (dtype == 0)
    rope_apply_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)out, rows, H, S, hd, cos, sin, rot, pid);
  else if (dtype == 1)
    rope_apply_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows,
                                                         H, S, hd, cos, sin, rot, pid);
  else {
    set_last_error("vf_rope_apply: dtype must be 0 (fp32) or 1 (bf16)");
    return VF_ERR_ARG;
  }
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_mrope_apply_strided(const void* x, void* out, int32_t dtype, int32_t B, int32_t H, int32_t S,
                                      int32_t hd, const int64_t* x_strides, const int64_t* out_strides,
                                      const float* cos, const float* sin, int32_t rot, int64_t table_rows,
                                      const int64_t* position_ids, int32_t sec_t, int32_t sec_h, int32_t sec_w,
                                      const float* norm_weight, float norm_eps, void* stream) {
  VF_REQUIRE(x && out && cos && sin && position_ids && x_strides && out_strides, VF_ERR_ARG, "vf_mrope_apply: null pointer");
  VF_REQUIRE(B > 0 && H > 0 && S > 0, VF_ERR_ARG, "vf_mrope_apply: bad shape");
  VF_REQUIRE(hd == 128 || hd == 256, VF_ERR_ARG, "vf_mrope_apply: head_dim %d unsupported (128 or 256)", hd);
  VF_REQUIRE(rot > 0 && rot <= hd && (rot / 2) % (hd / 32) == 0, VF_ERR_ARG,
             "vf_mrope_apply: rotation dim %d incompatible with head_dim %d", rot, hd);
  VF_REQUIRE(sec_t + sec_h + sec_w == rot / 2, VF_ERR_ARG,
             "vf_mrope_apply: mrope_section must sum to rot/2 (%d+%d+%d != %d)", sec_t, sec_h, sec_w, rot / 2);
  (void)table_rows;
  VF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(norm_weight) & 15) == 0,
             VF_ERR_ALIGN, "vf_mrope_apply: pointers must be 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    VF_REQUIRE(x_strides[i] % 8 == 0 && out_strides[i] % 8 == 0, VF_ERR_ALIGN,
               "vf_mrope_apply: strides must be multiples of 8 elements");
  const long long tokens = (long long)B * S;
  const unsigned grid = (unsigned)((tokens + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long* pid = reinterpret_cast<const long long*>(position_ids);
#define VF_MROPE(T, EPL)                                                                                          \
  mrope_apply_kernel<T, EPL><<<grid, 256, 0, s>>>((const T*)x, (T*)out, B, H, S, x_strides[0], x_strides[1], x_strides[2], \
                                                  out_strides[0], out_strides[1], out_strides[2], cos, sin, rot, pid,     \
                                                  sec_h, sec_w, norm_weight, norm_eps)
  if (dtype == 0 && hd == 256) VF_MROPE(float, 8);
  else if (dtype == 0 && hd == 128) VF_MROPE(float, 4);
  else if (dtype == 1 && hd == 256) VF_MROPE(__nv_bfloat16, 8);
  else if (dtype == 1 && hd == 128) VF_MROPE(__nv_bfloat16, 4);
  else {
    set_last_error("vf_mrope_apply: dtype must be 0 (fp32) or 1 (bf16)");
    return VF_ERR_ARG;
  }
#undef VF_MROPE
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_mrope_apply(const void* x, void* out, int32_t dtype, int32_t B, int32_t H, int32_t S,
                              int32_t hd, const float* cos, const float* sin, int32_t rot,
                              int64_t table_rows, const int64_t* position_ids, int32_t sec_t,
                              int32_t sec_h, int32_t sec_w, const float* norm_weight, float norm_eps,
                              void* stream) {
  const int64_t st[3] = {(int64_t)H * S * hd, (int64_t)S * hd, hd};   // [B, H, S, hd] contiguous
  return vf_mrope_apply_strided(x, out, dtype, B, H, S, hd, st, st, cos, sin, rot, table_rows, position_ids, sec_t, sec_h,
                                sec_w, norm_weight, norm_eps, stream);
}
