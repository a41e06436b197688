// vf_fuse.cu — the integer / byte-moving half of early fusion (all HBM-bound, bit-exact).
//
//  * mrope_position_ids_kernel — Qwen3_5VLM.compute_3d_position_ids
//      (reference llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py:85-176), one CTA per sample.
//  * fuse_scan_kernel         — exclusive scan of the image mask in flat (b, seq) order: the rank the
//      reference's masked_scatter assigns to every placeholder (vlm_model.py:206-211).
//  * embed_gather_scatter_kernel — emb_dict(input_ids) and the masked scatter in ONE pass: a
//      placeholder row is filled from vision row `rank`, any other row from the embedding table;
//      placeholder rows never touch the table (vlm_model.py:198,209).
//  * casts fp32 <-> bf16.
#include "vf_common.cuh"

namespace vf {

constexpr int MAX_FEEDS = 240;  // feeds travel in the kernel-parameter block (no H2D copy, no sync)

struct Feeds {
  int n;
  int t[MAX_FEEDS];
  int hm[MAX_FEEDS];  // h / merge
  int wm[MAX_FEEDS];  // w / merge
};

// Block-wide inclusive scan of one int per thread (blockDim.x == 1024). Returns the inclusive
// prefix; *total receives the block sum. `wsum` is a 32-entry shared scratch.
__device__ __forceinline__ long long block_scan_incl(long long v, long long* wsum, long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  __syncthreads();  // protect wsum from the previous use
  if (lane == 31) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    long long w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  if (warp > 0) v += wsum[warp - 1];
  *total = wsum[31];
  return v;
}

__global__ void __launch_bounds__(1024)
mrope_position_ids_kernel(const long long* __restrict__ ids, const uint8_t* __restrict__ mask,
                          long long image_token_id, const Feeds feeds, int b_total, int seq,
                          long long* __restrict__ out) {
  __shared__ long long wsum[32];
  __shared__ int s_used_feeds;        // how many feeds this sample consumes (reference: break at :148)
  __shared__ long long s_used_tokens; // placeholders covered by those feeds
  const int b = blockIdx.x;
  const long long* row_ids = ids + static_cast<long long>(b) * seq;
  const uint8_t* row_mask = mask ? mask + static_cast<long long>(b) * seq : nullptr;
  const long long plane = static_cast<long long>(b_total) * seq;
  long long* o_t = out + static_cast<long long>(b) * seq;
  long long* o_h = o_t + plane;
  long long* o_w = o_h + plane;

  // pass 1: number of placeholders in this sample
  long long cnt = 0;
  for (int i = threadIdx.x; i < seq; i += blockDim.x)
    cnt += row_mask ? (row_mask[i] != 0) : (row_ids[i] == image_token_id);
  long long n_img;
  block_scan_incl(cnt, wsum, &n_img);
  if (threadIdx.x == 0) {
    int used = 0;
    long long pos = 0;
    if (n_img > 0) {
      for (int f = 0; f < feeds.n; ++f) {
        const long long nt = static_cast<long long>(feeds.t[f]) * feeds.hm[f] * feeds.wm[f];
        if (pos + nt > n_img) break;
        pos += nt;
        ++used;
      }
    }
    s_used_feeds = used;
    s_used_tokens = pos;
  }
  __syncthreads();
  const int used_feeds = s_used_feeds;
  const long long used_tokens = s_used_tokens;

  // pass 2: chunks of 1024 tokens, carrying the placeholder rank and the position across chunks
  long long rank_carry = 0, pos_carry = 0;
  for (int base = 0; base < seq; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const bool in = i < seq;
    const bool is_img = in && (row_mask ? (row_mask[i] != 0) : (row_ids[i] == image_token_id));
    long long chunk_imgs;
    const long long rank = rank_carry + block_scan_incl(is_img ? 1 : 0, wsum, &chunk_imgs) - (is_img ? 1 : 0);

    long long inc = in ? (is_img ? 0 : 1) : 0;
    long long lt = 0, lh = 0, lw = 0;
    if (is_img && rank < used_tokens) {
      // which feed does this placeholder belong to?
      long long start = 0;
      int f = 0;
      for (; f < used_feeds; ++f) {
        const long long nt = static_cast<long long>(feeds.t[f]) * feeds.hm[f] * feeds.wm[f];
        if (rank < start + nt) break;
        start += nt;
      }
      const long long local = rank - start;
      const long long hw = static_cast<long long>(feeds.hm[f]) * feeds.wm[f];
      const long long nt = hw * feeds.t[f];
      lt = local / hw;
      const long long flat = local % hw;
      lh = flat / feeds.wm[f];
      lw = flat % feeds.wm[f];
      if (local == nt - 1) inc = max(feeds.t[f], max(feeds.hm[f], feeds.wm[f]));
    }
    long long chunk_inc;
    const long long incl = block_scan_incl(inc, wsum, &chunk_inc);
    const long long g = pos_carry + incl - inc;  // exclusive cumsum (reference :171)
    if (in) {
      o_t[i] = g + lt;
      o_h[i] = g + lh;
      o_w[i] = g + lw;
    }
    rank_carry += chunk_imgs;
    pos_carry += chunk_inc;
  }
}

// ------------------------------------------------------------------------------------------------
// flat exclusive scan of the placeholder mask: three tiny kernels (count / scan counts / write)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;

__device__ __forceinline__ bool is_placeholder(const long long* ids, const uint8_t* mask, long long tok,
                                               long long i) {
  return mask ? (mask[i] != 0) : (ids[i] == tok);
}

__global__ void __launch_bounds__(SCAN_BLOCK)
fuse_count_kernel(const long long* __restrict__ ids, const uint8_t* __restrict__ mask, long long tok,
                  long long n, int* __restrict__ block_counts) {
  __shared__ int wcnt[32];
  const long long i = static_cast<long long>(blockIdx.x) * SCAN_BLOCK + threadIdx.x;
  const bool f = i < n && is_placeholder(ids, mask, tok, i);
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wcnt[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = v;
  }
}

// single CTA: exclusive scan of block_counts in place; total -> n_placeholders
__global__ void __launch_bounds__(1024)
fuse_scan_counts_kernel(int* __restrict__ block_counts, int n_blocks, int* __restrict__ n_placeholders) {
  __shared__ long long wsum[32];
  long long carry = 0;
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < n_blocks ? block_counts[i] : 0;
    long long total;
    const long long incl = block_scan_incl(v, wsum, &total);
    if (i < n_blocks) block_counts[i] = static_cast<int>(carry + incl - v);
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && n_placeholders) *n_placeholders = static_cast<int>(carry);
}

__global__ void __launch_bounds__(SCAN_BLOCK)
fuse_rank_kernel(const long long* __restrict__ ids, const uint8_t* __restrict__ mask, long long tok,
                 long long n, const int* __restrict__ block_offsets, int* __restrict__ row_map,
                 int* __restrict__ inv_map, long long inv_cap) {
  __shared__ int wcnt[32];
  const long long i = static_cast<long long>(blockIdx.x) * SCAN_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool f = i < n && is_placeholder(ids, mask, tok, i);
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  if (lane == 0) wcnt[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int v = wcnt[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    wcnt[lane] = v - wcnt[lane];  // exclusive warp offsets
  }
  __syncthreads();
  if (i < n) {
    const int r = block_offsets[blockIdx.x] + wcnt[warp] + __popc(bal & ((1u << lane) - 1u));
    row_map[i] = f ? r : -1;
    if (f && inv_map && r < inv_cap) inv_map[r] = static_cast<int>(i);
  }
}

// one warp per output row; D*2 bytes moved with 16-byte accesses
template <typename TV>
__global__ void __launch_bounds__(256)
embed_gather_scatter_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                            long long vocab, int D, const TV* __restrict__ vision, long long n_vis,
                            const int* __restrict__ row_map, __nv_bfloat16* __restrict__ out,
                            long long n_tokens, int skip_vision) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_tokens) return;
  const int lane = threadIdx.x & 31;
  const int r = row_map ? row_map[row] : -1;
  uint4* dst = reinterpret_cast<uint4*>(out + row * D);
  if (r >= 0 && r < n_vis) {
    if (skip_vision) return;  // a GEMM epilogue (VF_EPI_SCATTER_BF16) writes these rows
    if (sizeof(TV) == 2) {
      const uint4* src = reinterpret_cast<const uint4*>(vision + static_cast<long long>(r) * D);
      for (int c = lane; c < D / 8; c += 32) dst[c] = __ldg(src + c);
    } else {
      const float4* src = reinterpret_cast<const float4*>(vision + static_cast<long long>(r) * D);
      for (int c = lane; c < D / 8; c += 32) {
        const float4 a = __ldg(src + 2 * c), b = __ldg(src + 2 * c + 1);
        dst[c] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
      }
    }
  } else {
    long long id = ids[row];
    if (id < 0 || id >= vocab) {  // nn.Embedding would raise; never read out of bounds
      for (int c = lane; c < D / 8; c += 32) dst[c] = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
      return;
    }
    const uint4* src = reinterpret_cast<const uint4*>(table + id * D);
    for (int c = lane; c < D / 8; c += 32) dst[c] = __ldg(src + c);
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x + i + 4));
      *reinterpret_cast<uint4*>(out + i) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    } else {
      for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(x[j]);
    }
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + i));
      *reinterpret_cast<float4*>(out + i) = make_float4(bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y));
      *reinterpret_cast<float4*>(out + i + 4) = make_float4(bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w));
    } else {
      for (long long j = i; j < n; ++j) out[j] = __bfloat162float(x[j]);
    }
  }
}

}  // namespace vf

using namespace vf;

extern "C" int vf_mrope_position_ids(const int64_t* input_ids, const uint8_t* image_mask,
                                     int64_t image_token_id, const int64_t* feeds_host, int32_t n_feeds,
                                     int32_t merge, int32_t b, int32_t seq, int64_t* out, void* stream) {
  VF_REQUIRE(input_ids && out, VF_ERR_ARG, "vf_mrope_position_ids: null pointer");
  VF_REQUIRE(b > 0 && seq > 0, VF_ERR_ARG, "vf_mrope_position_ids: bad shape b=%d seq=%d", b, seq);
  VF_REQUIRE(n_feeds >= 0 && n_feeds <= MAX_FEEDS, VF_ERR_ARG,
             "vf_mrope_position_ids: n_feeds=%d out of range [0,%d]", n_feeds, MAX_FEEDS);
  VF_REQUIRE(n_feeds == 0 || (feeds_host && merge > 0), VF_ERR_ARG, "vf_mrope_position_ids: feeds/merge missing");
  Feeds f{};
  f.n = n_feeds;
  for (int i = 0; i < n_feeds; ++i) {
    const int64_t t = feeds_host[3 * i], h = feeds_host[3 * i + 1], w = feeds_host[3 * i + 2];
    VF_REQUIRE(t > 0 && h / merge > 0 && w / merge > 0 && t < (1 << 30) && h < (1 << 30) && w < (1 << 30),
               VF_ERR_ARG, "vf_mrope_position_ids: feed %d has an empty or oversized (t,h,w)", i);
    f.t[i] = (int)t;
    f.hm[i] = (int)(h / merge);
    f.wm[i] = (int)(w / merge);
  }
  // n_feeds == 0: no feed is ever consumed, every placeholder keeps increment 0 / offset 0. The
  // host mirror handles the reference's "feeds_3d_shape is None" branch (plain arange) by passing
  // an all-zero mask.
  mrope_position_ids_kernel<<<b, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(input_ids), image_mask, image_token_id, f, b, seq,
      reinterpret_cast<long long*>(out));
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_fuse_scan(const int64_t* input_ids, const uint8_t* image_mask, int64_t image_token_id,
                            int64_t n_tokens, int32_t* row_map, int32_t* n_placeholders, int32_t* inv_map,
                            int64_t inv_cap, int32_t* scratch, void* stream) {
  VF_REQUIRE(input_ids && row_map && scratch, VF_ERR_ARG, "vf_fuse_scan: null pointer");
  VF_REQUIRE(n_tokens > 0 && n_tokens < (1ll << 31), VF_ERR_ARG, "vf_fuse_scan: bad n_tokens");
  const int n_blocks = (int)((n_tokens + SCAN_BLOCK - 1) / SCAN_BLOCK);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long* ids = reinterpret_cast<const long long*>(input_ids);
  fuse_count_kernel<<<n_blocks, SCAN_BLOCK, 0, s>>>(ids, image_mask, image_token_id, n_tokens, scratch);
  count_launch();
  fuse_scan_counts_kernel<<<1, 1024, 0, s>>>(scratch, n_blocks, n_placeholders);
  count_launch();
  if (inv_map && inv_cap > 0) VF_CUDA(cudaMemsetAsync(inv_map, 0xFF, inv_cap * sizeof(int32_t), s));  // all -1
  fuse_rank_kernel<<<n_blocks, SCAN_BLOCK, 0, s>>>(ids, image_mask, image_token_id, n_tokens, scratch, row_map,
                                                   inv_map, inv_cap);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_embed_gather_scatter(const int64_t* input_ids, const void* table, int64_t vocab, int32_t D,
                                       const void* vision, int32_t vis_dtype, int64_t n_vis,
                                       const int32_t* row_map, void* out, int64_t n_tokens,
                                       int32_t skip_vision, void* stream) {
  VF_REQUIRE(input_ids && table && out, VF_ERR_ARG, "vf_embed_gather_scatter: null pointer");
  VF_REQUIRE(n_tokens > 0 && D > 0 && D % 8 == 0, VF_ERR_ARG, "vf_embed_gather_scatter: D must be a multiple of 8");
  VF_REQUIRE(n_vis == 0 || skip_vision || vision, VF_ERR_ARG, "vf_embed_gather_scatter: vision rows missing");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(table) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(vision) & 15) == 0,
             VF_ERR_ALIGN, "vf_embed_gather_scatter: pointers must be 16-byte aligned");
  const unsigned grid = (unsigned)((n_tokens + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long* ids = reinterpret_cast<const long long*>(input_ids);
  if (vis_dtype == 1)
    embed_gather_scatter_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(
        ids, (const __nv_bfloat16*)table, vocab, D, (const __nv_bfloat16*)vision, n_vis, row_map,
        (__nv_bfloat16*)out, n_tokens, skip_vision);
  else if (vis_dtype == 0)
    embed_gather_scatter_kernel<float><<<grid, 256, 0, s>>>(ids, (const __nv_bfloat16*)table, vocab, D,
                                                           (const float*)vision, n_vis, row_map,
                                                           (__nv_bfloat16*)out, n_tokens, skip_vision);
  else {
    set_last_error("vf_embed_gather_scatter: vis_dtype must be 0 (fp32) or 1 (bf16)");
    return VF_ERR_ARG;
  }
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_cast_f32_to_bf16(const float* x, void* out, int64_t n, void* stream) {
  VF_REQUIRE(x && out && n > 0, VF_ERR_ARG, "vf_cast_f32_to_bf16: bad arguments");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             VF_ERR_ALIGN, "vf_cast_f32_to_bf16: pointers must be 16-byte aligned");
  long long blocks = (n / 8 + 255) / 256;
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(out), n);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

extern "C" int vf_cast_bf16_to_f32(const void* x, float* out, int64_t n, void* stream) {
  VF_REQUIRE(x && out && n > 0, VF_ERR_ARG, "vf_cast_bf16_to_f32: bad arguments");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             VF_ERR_ALIGN, "vf_cast_bf16_to_f32: pointers must be 16-byte aligned");
  long long blocks = (n / 8 + 255) / 256;
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_bf16_f32_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), out, n);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
