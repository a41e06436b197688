// vf_attention.cu — fused bidirectional attention for sm_100a (head_dim 64, bf16, fp32 softmax).
//
// Replaces F.scaled_dot_product_attention and the head-major transposes around it
// (llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:169-190, vit_attention.py:62-87 in the reference).
// Q, K, V are read IN PLACE from the token-major [B*S, 3*H*64] buffer the QKV GEMM writes (TMA boxes
// of 128 tokens x 64 columns at column offsets h*64, H*64+h*64, 2*H*64+h*64), and the context is
// written token-major [B*S, H*64] — no head-major copy exists anywhere.
//
// One persistent CTA per SM; a work item is (sample b, head h, block of 256 queries) = two 128-row
// query tiles that share every K/V tile:
//   warp 0        TMA loader: Q0,Q1 once per item; K and V tiles through two 3-stage rings
//   warp 1        tcgen05.mma issuer:  S_t = Q_t K_j^T  (SS, M=128,N=128,K=64)
//                                      O_t += P_t V_j   (TS: P_t bf16 in TMEM; V MN-major smem)
//   warps 4-7     softmax warpgroup for tile 0   } one thread per query row (tcgen05.ld 32x32b):
//   warps 8-11    softmax warpgroup for tile 1   } rowmax / exp2 / rowsum / P->TMEM / lazy O rescale
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384); P_t aliases the first 64 columns of S_t.
// The MMA warp issues S_t(j+1) right behind PV_t(j), so one tile's softmax overlaps the other
// tile's MMAs. tcgen05.mma instructions of one thread retire in order, which is what makes the
// S/P aliasing and the in-place O rescale race-free (see comments at the barriers).
#include "vf_common.cuh"

#include <math.h>

namespace vf {

constexpr int ATT_THREADS = 384;
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 tokens x 64 dims bf16

struct AttnParams {
  int B, S, H;
  int n_qblk;      // ceil(S / 256)
  int n_kt;        // ceil(S / 128)
  int n_items;     // B * H * n_qblk
  float scale_log2;
  __nv_bfloat16* out;
};

struct AttnSmem {
  static constexpr int Q_OFF = 0;                             // 2 tiles
  static constexpr int K_OFF = 2 * TILE_BYTES;                // KV_STAGES tiles
  static constexpr int V_OFF = K_OFF + KV_STAGES * TILE_BYTES;
  static constexpr int BAR_OFF = V_OFF + KV_STAGES * TILE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

__device__ __forceinline__ void att_wait(uint64_t* bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("vf_attention: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQKV) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;                  // [KV_STAGES]
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;       // [2]
  uint64_t* p_full = s_full + 2;                // [2]
  uint64_t* o_full = p_full + 2;                // [2]
  uint64_t* o_empty = o_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int HD3 = 3 * p.H * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);   // one arrival per softmax warp
      mbar_init(&o_full[t], 1);
      mbar_init(&o_empty[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // loader / MMA / idle warpgroup: give registers back to the softmax warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader
    if (lane == 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, qph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int qb = item % p.n_qblk;
        const int h = (item / p.n_qblk) % p.H;
        const int b = item / (p.n_qblk * p.H);
        const int row0 = b * p.S;
        att_wait(q_empty, qph ^ 1);
        mbar_expect_tx(q_full, 2 * TILE_BYTES);
        tma_load_2d(smem + AttnSmem::Q_OFF, &tmQKV, q_full, h * 64, row0 + qb * 256);
        tma_load_2d(smem + AttnSmem::Q_OFF + TILE_BYTES, &tmQKV, q_full, h * 64, row0 + qb * 256 + 128);
        qph ^= 1;
        for (int j = 0; j < p.n_kt; ++j) {
          att_wait(&k_empty[ks], kph ^ 1);
          mbar_expect_tx(&k_full[ks], TILE_BYTES);
          tma_load_2d(smem + AttnSmem::K_OFF + ks * TILE_BYTES, &tmQKV, &k_full[ks], p.H * 64 + h * 64,
                      row0 + j * 128);
          if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
          att_wait(&v_empty[vs], vph ^ 1);
          mbar_expect_tx(&v_full[vs], TILE_BYTES);
          tma_load_2d(smem + AttnSmem::V_OFF + vs * TILE_BYTES, &tmQKV, &v_full[vs], 2 * p.H * 64 + h * 64,
                      row0 + j * 128);
          if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
        }
      }
    }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      // The whole warp walks the (warp-uniform) schedule so that addresses and descriptors live in
      // uniform registers; one elected lane issues the tcgen05 instructions.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // Q K^T : both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // P V   : V is MN-major
      const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + AttnSmem::Q_OFF));
      const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + AttnSmem::K_OFF));
      const uint64_t v_desc = umma_desc_sw128(smem_u32(smem + AttnSmem::V_OFF));
      constexpr uint64_t TILE_DESC = TILE_BYTES >> 4;   // one 16 KB tile further, in descriptor units
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, qph = 0;
      uint32_t pph0 = 0, pph1 = 0, oeph0 = 0, oeph1 = 0;

#define VF_ISSUE_S(T, KSTAGE)                                                                        \
  do {                                                                                               \
    if (elect_one()) {                                                                               \
      const uint64_t a_ = q_desc + (T) * TILE_DESC;                                                  \
      const uint64_t b_ = k_desc + (KSTAGE) * TILE_DESC;                                             \
      _Pragma("unroll") for (int k_ = 0; k_ < 4; ++k_)                                               \
          umma_ss(tmem_base + (T) * 128, a_ + 2 * k_, b_ + 2 * k_, idesc_s, k_ != 0);                \
      umma_commit(&s_full[T]);                                                                       \
    }                                                                                                \
    __syncwarp();                                                                                    \
  } while (0)
#define VF_ISSUE_PV(T, VSTAGE, ACC)                                                                  \
  do {                                                                                               \
    if (elect_one()) {                                                                               \
      const uint64_t b_ = v_desc + (VSTAGE) * TILE_DESC;                                             \
      _Pragma("unroll") for (int k_ = 0; k_ < 8; ++k_) /* 16 keys: 8 TMEM cols of bf16x2, 16 V rows */ \
          umma_ts(tmem_base + 256 + (T) * 64, tmem_base + (T) * 128 + k_ * 8, b_ + k_ * (2048 >> 4),  \
                  idesc_o, (ACC) || k_ != 0);                                                        \
    }                                                                                                \
    __syncwarp();                                                                                    \
  } while (0)
#define VF_COMMIT(BAR)                                                                               \
  do {                                                                                               \
    if (elect_one()) umma_commit(BAR);                                                               \
    __syncwarp();                                                                                    \
  } while (0)

      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int qb = item % p.n_qblk;
        const bool t1 = qb * 256 + 128 < p.S;  // second query tile has at least one valid row
        att_wait(q_full, qph);
        // O_t / S_t of the previous item must have been drained by the softmax warpgroups
        att_wait(&o_empty[0], oeph0 ^ 1); oeph0 ^= 1;
        if (t1) { att_wait(&o_empty[1], oeph1 ^ 1); oeph1 ^= 1; }

        // prologue: S_t(0)
        att_wait(&k_full[ks], kph);
        tc_fence_after();
        VF_ISSUE_S(0, ks);
        if (t1) VF_ISSUE_S(1, ks);
        VF_COMMIT(&k_empty[ks]);
        if (p.n_kt == 1) VF_COMMIT(q_empty);
        if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }

        for (int j = 0; j < p.n_kt; ++j) {
          const bool more = (j + 1 < p.n_kt);
          att_wait(&v_full[vs], vph);
          att_wait(&p_full[0], pph0); pph0 ^= 1;
          tc_fence_after();
          VF_ISSUE_PV(0, vs, j > 0);
          if (more) {
            att_wait(&k_full[ks], kph);
            tc_fence_after();
            VF_ISSUE_S(0, ks);
          }
          if (t1) {
            att_wait(&p_full[1], pph1); pph1 ^= 1;
            tc_fence_after();
            VF_ISSUE_PV(1, vs, j > 0);
            if (more) VF_ISSUE_S(1, ks);
          }
          VF_COMMIT(&v_empty[vs]);
          if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
          if (more) {
            VF_COMMIT(&k_empty[ks]);
            if (j + 2 == p.n_kt) VF_COMMIT(q_empty);  // last S MMAs of this item are in flight
            if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
          }
        }
        VF_COMMIT(&o_full[0]);
        if (t1) VF_COMMIT(&o_full[1]);
        qph ^= 1;
      }
#undef VF_ISSUE_S
#undef VF_ISSUE_PV
#undef VF_COMMIT
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int t = (warp - 4) >> 2;        // query tile 0 / 1
    const int quarter = warp & 3;         // TMEM lane quarter
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_sel + t * 128;
    const uint32_t o_addr = tmem_base + lane_sel + 256 + t * 64;
    const int r_local = quarter * 32 + lane;
    uint32_t sph = 0, oph = 0;

    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int qb = item % p.n_qblk;
      const int h = (item / p.n_qblk) % p.H;
      const int b = item / (p.n_qblk * p.H);
      if (t == 1 && !(qb * 256 + 128 < p.S)) continue;  // whole tile out of range (uniform per warp)
      const int q_in_sample = qb * 256 + t * 128 + r_local;

      float m = -INFINITY;   // running (possibly stale) row max, raw score units
      float l = 0.f;         // running row sum
      for (int j = 0; j < p.n_kt; ++j) {
        att_wait(&s_full[t], sph); sph ^= 1;
        tc_fence_after();
        // All MMAs issued before S_t(j) — in particular PV_t(j-1) — have retired: O_t is stable
        // until this warpgroup arrives on p_full[t].
        float s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(s_addr + c * 32, reinterpret_cast<uint32_t*>(s) + c * 32);
        tmem_ld_wait();
        const int valid = p.S - j * 128;  // keys [0, valid) of this tile exist
        if (valid < 128) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= valid) s[c] = -INFINITY;
        }
        float mx = s[0];
#pragma unroll
        for (int c = 1; c < 128; ++c) mx = fmaxf(mx, s[c]);

        if (j == 0) {
          m = mx;
        } else {
          const bool grow = (mx - m) * p.scale_log2 > 8.0f;  // lazy rescale threshold: 2^8 headroom
          if (__any_sync(0xffffffffu, grow)) {
            const float f = grow ? fast_exp2((m - mx) * p.scale_log2) : 1.0f;
            if (grow) { m = mx; l *= f; }
            uint32_t o[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              tmem_ld_x32(o_addr + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
              tmem_st_x32(o_addr + c * 32, o);
            }
          }
        }
        const float mb = m * p.scale_log2;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = fast_exp2(fmaf(s[c * 32 + 2 * e], p.scale_log2, -mb));
            const float p1 = fast_exp2(fmaf(s[c * 32 + 2 * e + 1], p.scale_log2, -mb));
            sum += p0 + p1;
            pk[e] = pack_bf16(p0, p1);
          }
          tmem_st_x16(s_addr + c * 16, pk);   // P_t (bf16 pairs) over the first 64 columns of S_t
        }
        l += sum;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }

      // ---- item epilogue: O_t / l -> global
      att_wait(&o_full[t], oph); oph ^= 1;
      tc_fence_after();
      const float inv = 1.0f / l;
      const bool row_ok = q_in_sample < p.S;
      __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.S + q_in_sample) * (p.H * 64) + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld_x32(o_addr + c * 32, o);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            reinterpret_cast<uint4*>(orow + c * 32)[q] = v;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[t]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  (void)HD3;
}

}  // namespace vf

using namespace vf;

extern "C" int vf_attention_fwd(const void* qkv, void* out, int32_t B, int32_t S, int32_t H,
                                float scale, void* stream) {
  VF_REQUIRE(qkv && out, VF_ERR_ARG, "vf_attention_fwd: null pointer");
  VF_REQUIRE(B > 0 && S > 0 && H > 0, VF_ERR_ARG, "vf_attention_fwd: bad shape B=%d S=%d H=%d", B, S, H);
  VF_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             VF_ERR_ALIGN, "vf_attention_fwd: pointers must be 16-byte aligned");
  VF_REQUIRE((long long)B * S < (1ll << 31), VF_ERR_ARG, "vf_attention_fwd: B*S too large");

  AttnParams p{};
  p.B = B; p.S = S; p.H = H;
  p.n_qblk = (S + 255) / 256;
  p.n_kt = (S + 127) / 128;
  p.n_items = B * H * p.n_qblk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);

  CUtensorMap tm;
  uint64_t dims[2] = {(uint64_t)3 * H * 64, (uint64_t)B * S};
  uint64_t strides[1] = {(uint64_t)3 * H * 64 * 2};
  uint32_t box[2] = {64, 128};
  int e = encode_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box,
                      CU_TENSOR_MAP_SWIZZLE_128B);
  if (e) return e;

  static bool configured = false;
  if (!configured) {
    VF_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AttnSmem::TOTAL));
    configured = true;
  }
  const int sms = device_sm_count();
  VF_REQUIRE(sms > 0, VF_ERR_NO_DEVICE, "no CUDA device");
  const int grid = p.n_items < sms ? p.n_items : sms;
  attention_kernel<<<grid, ATT_THREADS, AttnSmem::TOTAL, static_cast<cudaStream_t>(stream)>>>(p, tm);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
