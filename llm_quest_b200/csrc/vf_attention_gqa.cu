// vf_attention_gqa.cu — causal grouped-query attention, head_dim 256, for sm_100a (bf16, fp32 softmax).
//
// The attention core of the first consumer of the fused embeddings and MRoPE-I position ids, the text model's
// MRoPEGatedAttention in prefill (reference llm_quest/qwen/qwen3_5/qwen3_5_text_model.py:194-267):
//     F.scaled_dot_product_attention(q, k, v, causal mask, enable_gqa=True)   8 query heads / 2 kv heads x 256
//     ctx = ctx * sigmoid(gate)                                               (:262)
// q, k, v are read token-major ([B*S, heads*256], the layout the projection GEMMs write) through TMA boxes at
// the head's column offset; the context is written token-major, already multiplied by sigmoid(gate) when a gate
// pointer is given (the gate columns live next to the query columns in the w_queries_gate output, :235-238).
//
// At head_dim 256 the balance is the opposite of the vision kernel (vf_attention.cu): a 128 x 64 score tile costs
// 1024 tensor cycles (QK^T over K=256 + PV over N=256) against 512 MUFU cycles, so ONE chain per CTA is enough if
// the tensor core never waits for the softmax:
//   TMEM (all 512 columns): S double-buffered at [0,64) / [64,128), P(j) overwrites S(j) in its buffer; Q (copied
//                           from shared memory once per item, tcgen05.cp) at [128,256); O at [256,512).
//   warps 0..3  softmax, one thread per query row (lean step: packed f32x2 math, first-tile max as reference,
//               step redone with a fresh max only when a row sum runs past 2^60 — see vf_attention.cu);
//   warp 4      TMA loader: Q (4 swizzle atoms of 64 dims) once per item, K and V tiles through 2-stage rings;
//   warp 5      issuer of S(j+1) = Q K_{j+1}^T: runs one step ahead, under the softmax of step j (waits only for the PV
//               that last read the target buffer); warp 6: issuer of O += P(j) V_j as soon as P(j) is stored.
// Measured and not kept (round 1, B=32 S=2832 with gate, before Q moved to TMEM): eight softmax warps, two per lane
// quarter each owning 32 of a tile's 64 key columns (1329 us vs 1230), and additionally four S buffers + a 3-stage K
// ring so that S can run three steps ahead (1352 us): the step was bound neither by the softmax nor by the S/PV
// dependency but by the shared-memory reads of the N = 64 QK^T MMAs (see the S issuer).
// Work item = (sample, query head, 128-row query tile), causal: key tiles 0 .. 2*tile+1 only; items are walked in
// decreasing cost (last query tiles first), boustrophedon over the CTAs.
#include "vf_common.cuh"

#include <math.h>

namespace vf {

constexpr int GD = 256;                      // head dim
constexpr int GKT = 64;                      // keys per tile
constexpr int G_ATOMS = GD / 64;             // 128-byte swizzle atoms per row
constexpr int G_THREADS = 224;
constexpr int G_STAGES = 2;
constexpr int GQ_ATOM_BYTES = 128 * 128;     // 128 rows x 64 dims
constexpr int GKV_ATOM_BYTES = GKT * 128;    // 64 keys x 64 dims
constexpr int GQ_BYTES = G_ATOMS * GQ_ATOM_BYTES;     // 64 KB
constexpr int GKV_BYTES = G_ATOMS * GKV_ATOM_BYTES;   // 32 KB
constexpr int GT_S = 0, GT_Q = 128, GT_O = 256;   // TMEM columns: S/P 2 x 64, Q 128 (bf16 pairs), O 256

struct GqaParams {
  int B, S, Hq, Hkv;
  int n_qt, n_items, n_bh;
  int causal;
  float scale_log2;
  __nv_bfloat16* out;
  long long ldo;
  const __nv_bfloat16* gate;   // optional: out *= sigmoid(gate[row, gate_col0 + h*gate_head_stride + d])
  long long ldg;
  int gate_col0, gate_head_stride;
  int q_col0, q_head_stride;   // column of head h in the q tensor map = q_col0 + h*q_head_stride
};

struct GqaSmem {
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = GQ_BYTES;
  static constexpr int V_OFF = K_OFF + G_STAGES * GKV_BYTES;
  static constexpr int E_OFF = V_OFF + G_STAGES * GKV_BYTES;   // epilogue staging: per softmax warp a gate tile and an
  static constexpr int BAR_OFF = E_OFF + 4 * 8192;              // output tile of 32 rows x 64 columns bf16 (4 KB each)
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

// MN-major SW128 operand spanning several 64-element atoms along N: atoms `lbo_bytes` apart, 8-row (K) groups 1024 B
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct GqaItem {
  int b, h, kvh, qt, n_kt;
};
__device__ __forceinline__ GqaItem gqa_decode(const GqaParams& p, int item) {
  GqaItem it;
  const int r = item / p.n_bh;
  const int bh = item - r * p.n_bh;
  it.qt = p.n_qt - 1 - r;                      // expensive (late) query tiles first
  it.b = bh / p.Hq;
  it.h = bh - it.b * p.Hq;
  it.kvh = it.h / (p.Hq / p.Hkv);
  const int all = (p.S + GKT - 1) / GKT;
  const int need = 2 * it.qt + 2;              // key tiles that reach the diagonal of this query tile
  it.n_kt = (p.causal && need < all) ? need : all;
  return it;
}

struct GqaIter {   // same boustrophedon walk as vf_attention.cu
  int r, n;
  __device__ explicit GqaIter(int n_items) : r(0), n(n_items) {}
  __device__ __forceinline__ bool next(int& item) {
    const int G = gridDim.x;
    while (r * G < n) {
      const int k = (r & 1) ? G - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
      const int idx = r * G + k;
      ++r;
      if (idx < n) { item = idx; return true; }
    }
    return false;
  }
};

__global__ void __launch_bounds__(G_THREADS, 1)
attention_gqa_kernel(const GqaParams p, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmG) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GqaSmem::BAR_OFF);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* q_empty = bars + 1;            // [1]
  uint64_t* k_full = bars + 2;             // [G_STAGES]
  uint64_t* k_empty = k_full + G_STAGES;
  uint64_t* v_full = k_empty + G_STAGES;
  uint64_t* v_empty = v_full + G_STAGES;
  uint64_t* s_full = v_empty + G_STAGES;   // [2] per S buffer
  uint64_t* p_full = s_full + 2;           // [2]
  uint64_t* pv_done = p_full + 2;          // [2] per S/P buffer: PV that read P from it has retired
  uint64_t* o_full = pv_done + 2;          // [1]
  uint64_t* o_empty = o_full + 1;          // [1]
  uint64_t* g_full = o_empty + 1;          // [4] per softmax warp: its gate tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_full + 4);
  const char* const WHO = "vf_attention_gqa";

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    for (int i = 0; i < 4; ++i) mbar_init(&g_full[i], 1);
    tma_prefetch_desc(&tmO);
    if (p.gate) tma_prefetch_desc(&tmG);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // -------------------------------------------------------------------- TMA loader
    if (lane == 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, qph = 0;
      int item;
      for (GqaIter it(p.n_items); it.next(item);) {
        const GqaItem w = gqa_decode(p, item);
        const int row0 = w.b * p.S;
        mbar_wait_or_trap(q_empty, qph ^ 1, WHO);
        mbar_expect_tx(q_full, GQ_BYTES);
#pragma unroll
        for (int a = 0; a < G_ATOMS; ++a)
          tma_load_2d(smem + GqaSmem::Q_OFF + a * GQ_ATOM_BYTES, &tmQ, q_full, p.q_col0 + w.h * p.q_head_stride + a * 64,
                      row0 + w.qt * 128);
        qph ^= 1;
        for (int j = 0; j < w.n_kt; ++j) {
          mbar_wait_or_trap(&k_empty[ks], kph ^ 1, WHO);
          mbar_expect_tx(&k_full[ks], GKV_BYTES);
#pragma unroll
          for (int a = 0; a < G_ATOMS; ++a)
            tma_load_2d(smem + GqaSmem::K_OFF + ks * GKV_BYTES + a * GKV_ATOM_BYTES, &tmK, &k_full[ks],
                        w.kvh * GD + a * 64, row0 + j * GKT);
          if (++ks == G_STAGES) { ks = 0; kph ^= 1; }
          mbar_wait_or_trap(&v_empty[vs], vph ^ 1, WHO);
          mbar_expect_tx(&v_full[vs], GKV_BYTES);
#pragma unroll
          for (int a = 0; a < G_ATOMS; ++a)
            tma_load_2d(smem + GqaSmem::V_OFF + vs * GKV_BYTES + a * GKV_ATOM_BYTES, &tmV, &v_full[vs],
                        w.kvh * GD + a * 64, row0 + j * GKT);
          if (++vs == G_STAGES) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // -------------------------------------------------------------------- issuer of S(j) = Q K_j^T
    // Its own warp: issuing the 16 MMAs of one S takes longer than the four of a PV, and in one thread it would sit
    // between P(j) becoming ready and PV(j) being issued. S(j) may overwrite its buffer once the PV that read P(j-2)
    // from it has retired (pv_done of that buffer) — the two issuers are separate threads, so program order does not
    // give that any more.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, GKT, 0, 0);
    const uint32_t q_addr = smem_u32(smem + GqaSmem::Q_OFF);
    const uint32_t k_addr = smem_u32(smem + GqaSmem::K_OFF);
    int ks = 0;
    uint32_t kph = 0, qph = 0;
    unsigned g = 0;   // global key-step counter: S buffer = g & 1, barrier parity = (g >> 1) & 1
    int item;
    for (GqaIter it(p.n_items); it.next(item);) {
      const GqaItem w = gqa_decode(p, item);
      mbar_wait_or_trap(q_full, qph, WHO); qph ^= 1;
      tc_fence_after();
      if (elect_one()) {
        // Q: shared memory -> tensor memory once per item (16 copies of 128 rows x 16 dims). As an operand read from
        // TMEM it costs the QK^T MMAs no shared-memory bandwidth: with Q in smem an N = 64 MMA read 6 KB per 32 cycles
        // of math (A 128x16 + B 64x16), above the 128 B/clk an SM delivers. tcgen05.cp and tcgen05.mma of one thread
        // execute in issue order, so the copies wait for the previous item's last S and this item's S wait for them.
#pragma unroll
        for (int kk = 0; kk < GD / 16; ++kk) {
          const int a = kk >> 2, i = kk & 3;
          asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_base + GT_Q + kk * 8),
                       "l"(umma_desc_sw128(q_addr + a * GQ_ATOM_BYTES) + 2 * i)
                       : "memory");
        }
        umma_commit(q_empty);   // the Q tile in shared memory is free once the copies have retired
      }
      __syncwarp();
      for (int j = 0; j < w.n_kt; ++j, ++g) {
        mbar_wait_or_trap(&k_full[ks], kph, WHO);
        if (g >= 2) mbar_wait_or_trap(&pv_done[g & 1], ((g >> 1) - 1) & 1, WHO);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + GT_S + (g & 1) * 64;
#pragma unroll
          for (int kk = 0; kk < GD / 16; ++kk) {
            const int a = kk >> 2, i = kk & 3;
            umma_ts(d, tmem_base + GT_Q + kk * 8,
                    umma_desc_sw128(k_addr + ks * GKV_BYTES + a * GKV_ATOM_BYTES) + 2 * i, idesc_s, kk != 0);
          }
          umma_commit(&s_full[g & 1]);
          umma_commit(&k_empty[ks]);
        }
        __syncwarp();
        if (++ks == G_STAGES) { ks = 0; kph ^= 1; }
      }
    }
  } else if (warp == 6) {
    // -------------------------------------------------------------------- issuer of O += P(j) V_j
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, GD, 0, 1);   // N = 256: V is MN-major, four 64-dim atoms (LBO)
    const uint32_t v_addr = smem_u32(smem + GqaSmem::V_OFF);
    int vs = 0;
    uint32_t vph = 0, oeph = 0;
    unsigned g = 0;
    int item;
    for (GqaIter it(p.n_items); it.next(item);) {
      const GqaItem w = gqa_decode(p, item);
      for (int j = 0; j < w.n_kt; ++j, ++g) {
        mbar_wait_or_trap(&v_full[vs], vph, WHO);
        mbar_wait_or_trap(&p_full[g & 1], (g >> 1) & 1, WHO);
        if (j == 0) { mbar_wait_or_trap(o_empty, oeph ^ 1, WHO); oeph ^= 1; }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_t = tmem_base + GT_S + (g & 1) * 64;
#pragma unroll
          for (int k_ = 0; k_ < GKT / 16; ++k_) {   // 16 keys per MMA: 8 TMEM columns of bf16 pairs / 16 V rows
            umma_ts(tmem_base + GT_O, a_t + k_ * 8,
                    umma_desc_sw128_mn(v_addr + vs * GKV_BYTES, GKV_ATOM_BYTES) + k_ * (2048 >> 4), idesc_o,
                    j > 0 || k_ != 0);
          }
          umma_commit(&v_empty[vs]);
          umma_commit(&pv_done[g & 1]);
          if (j + 1 == w.n_kt) umma_commit(o_full);
        }
        __syncwarp();
        if (++vs == G_STAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warps (one thread per query row)
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t o_addr = tmem_base + lane_sel + GT_O;
    unsigned g = 0;
    uint32_t oph = 0, gph = 0;
    uint8_t* gslot = smem + GqaSmem::E_OFF + warp * 8192;     // gate tile in, [32 rows][64 columns] bf16, 128-byte swizzle
    uint8_t* oslot = gslot + 4096;                            // output tile out, same shape
    const uint32_t grow = smem_u32(gslot) + lane * 128, orow = smem_u32(oslot) + lane * 128;
    int item;
    for (GqaIter it(p.n_items); it.next(item);) {
      const GqaItem w = gqa_decode(p, item);
      const int q_pos = w.qt * 128 + warp * 32 + lane;        // position of my row inside the sample
      float m = -INFINITY, l = 0.f;
      // The epilogue moves 32 rows x 64 columns at a time through shared memory: gate tiles come in and output tiles go out as
      // TMA boxes (full 128-byte lines; a row per thread straight from / to global memory costs 32 lines per warp instruction,
      // 48 such instructions per row). The first gate tile is requested now, the others while their predecessor is in use.
      const bool warp_ok = w.qt * 128 + warp * 32 < p.S;       // at least one row of this warp exists
      const int g_col = p.gate_col0 + w.h * p.gate_head_stride;
      if (p.gate && warp_ok && lane == 0) {
        mbar_expect_tx(&g_full[warp], 4096);
        tma_load_3d(gslot, &tmG, &g_full[warp], g_col, w.qt * 128 + warp * 32, w.b);
      }
      for (int j = 0; j < w.n_kt; ++j, ++g) {
        const uint32_t s_addr = tmem_base + lane_sel + GT_S + (g & 1) * 64;
        mbar_wait_or_trap(&s_full[g & 1], (g >> 1) & 1, WHO);
        tc_fence_after();
        float s[GKT];
        tmem_ld_x32(s_addr, reinterpret_cast<uint32_t*>(s));
        tmem_ld_x32(s_addr + 32, reinterpret_cast<uint32_t*>(s) + 32);
        tmem_ld_wait();
        // keys [0, limit] of this tile are visible to my row: sequence end and, if causal, the diagonal
        int limit = p.S - 1 - j * GKT;
        if (p.causal && q_pos - j * GKT < limit) limit = q_pos - j * GKT;
        if (__any_sync(0xffffffffu, limit < GKT - 1)) {
#pragma unroll
          for (int c = 0; c < GKT; ++c)
            if (c > limit) s[c] = -INFINITY;
        }
        auto row_max = [&]() {
          float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
          for (int c = 4; c < GKT; c += 2) {
            mx0 = fmaxf(mx0, s[c]);
            mx1 = fmaxf(mx1, s[c + 1]);
          }
          return fmaxf(mx0, mx1);
        };
        auto exp_store = [&](float mb) {
          const uint64_t sc2 = pack2(p.scale_log2, p.scale_log2), nb2 = pack2(-mb, -mb);
          uint64_t acc0 = pack2(0.f, 0.f), acc1 = acc0;
#pragma unroll
          for (int c = 0; c < GKT / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float x0, x1;
              unpack2(ffma2(pack2(s[c * 32 + 2 * e], s[c * 32 + 2 * e + 1]), sc2, nb2), x0, x1);
              const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
              if (e & 1) acc1 = fadd2(acc1, pack2(p0, p1));
              else acc0 = fadd2(acc0, pack2(p0, p1));
              pk[e] = pack_bf16(p0, p1);
            }
            tmem_st_x16(s_addr + c * 16, pk);   // P(j) over the first 32 columns of its S buffer
          }
          float a0, a1;
          unpack2(fadd2(acc0, acc1), a0, a1);
          return a0 + a1;
        };
        if (j == 0) m = row_max();               // key 0 is visible to every row, so m is finite
        float ssum = exp_store(m * p.scale_log2);
        if (j > 0 && __any_sync(0xffffffffu, !(ssum <= 0x1p60f))) {   // runaway exponent: redo with a fresh max
          mbar_wait_or_trap(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, WHO);   // O is stable once PV(previous step) retired
          tc_fence_after();
          const float mx = row_max();
          const bool grow = mx > m;
          const float f = grow ? fast_exp2((m - mx) * p.scale_log2) : 1.0f;
          if (grow) { m = mx; l *= f; }
#pragma unroll
          for (int c = 0; c < GD / 16; ++c) {
            uint32_t o[16];
            tmem_ld_x16(o_addr + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st_x16(o_addr + c * 16, o);
          }
          ssum = exp_store(m * p.scale_log2);
        }
        l += ssum;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1]);
      }
      // ---- item epilogue: O / l (* sigmoid(gate)) -> global, 64 columns at a time
      mbar_wait_or_trap(o_full, oph, WHO); oph ^= 1;
      tc_fence_after();
      const float inv = 1.0f / l;
      if (warp_ok) {
#pragma unroll 1
        for (int c = 0; c < GD / 64; ++c) {
          uint32_t o[64];
          tmem_ld_x32(o_addr + c * 64, o);
          tmem_ld_x32(o_addr + c * 64 + 32, o + 32);
          tmem_ld_wait();
          uint32_t pk[32];
          if (p.gate) {
            mbar_wait_or_trap(&g_full[warp], gph, WHO); gph ^= 1;
            uint32_t gq[32];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(gq[4 * k]), "=r"(gq[4 * k + 1]), "=r"(gq[4 * k + 2]), "=r"(gq[4 * k + 3])
                           : "r"(grow + ((k ^ (lane & 7)) << 4)));
            __syncwarp();                        // every lane has its gate values: the tile may be overwritten
            if (lane == 0 && c + 1 < GD / 64) {
              mbar_expect_tx(&g_full[warp], 4096);
              tma_load_3d(gslot, &tmG, &g_full[warp], g_col + (c + 1) * 64, w.qt * 128 + warp * 32, w.b);
            }
#pragma unroll
            for (int e = 0; e < 32; ++e) {       // sigmoid(x) = 0.5 tanh(x/2) + 0.5: one MUFU op per element
              const float v0 = __uint_as_float(o[2 * e]) * inv * fmaf(0.5f, fast_tanh(0.5f * bf16_lo(gq[e])), 0.5f);
              const float v1 = __uint_as_float(o[2 * e + 1]) * inv * fmaf(0.5f, fast_tanh(0.5f * bf16_hi(gq[e])), 0.5f);
              pk[e] = pack_bf16(v0, v1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) pk[e] = pack_bf16(__uint_as_float(o[2 * e]) * inv, __uint_as_float(o[2 * e + 1]) * inv);
          }
          if (lane == 0) tma_store_wait_read<0>();   // the previous store out of the output tile has left shared memory
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + ((k ^ (lane & 7)) << 4)), "r"(pk[4 * k]),
                         "r"(pk[4 * k + 1]), "r"(pk[4 * k + 2]), "r"(pk[4 * k + 3])
                         : "memory");
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmO, oslot, w.h * GD + c * 64, w.qt * 128 + warp * 32, w.b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  if (warp < 4 && lane == 0) tma_store_wait<0>();   // shared memory must outlive the last output stores
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vf

using namespace vf;

extern "C" int vf_attention_gqa_fwd(const void* q, int64_t ldq, int32_t q_col0, int32_t q_head_stride, const void* k,
                                    int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo, const void* gate,
                                    int64_t ldg, int32_t gate_col0, int32_t gate_head_stride, int32_t B, int32_t S,
                                    int32_t Hq, int32_t Hkv, int32_t head_dim, float scale, int32_t causal,
                                    void* stream) {
  VF_REQUIRE(q && k && v && out, VF_ERR_ARG, "vf_attention_gqa_fwd: null pointer");
  VF_REQUIRE(B > 0 && S > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, VF_ERR_ARG,
             "vf_attention_gqa_fwd: bad shape B=%d S=%d Hq=%d Hkv=%d", B, S, Hq, Hkv);
  VF_REQUIRE(head_dim == GD, VF_ERR_ARG, "vf_attention_gqa_fwd: head_dim %d unsupported (256)", head_dim);
  VF_REQUIRE((long long)B * S < (1ll << 31), VF_ERR_ARG, "vf_attention_gqa_fwd: B*S too large");
  VF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && (!gate || ldg % 8 == 0) && q_col0 % 8 == 0 &&
                 q_head_stride % 8 == 0 && gate_col0 % 8 == 0 && gate_head_stride % 8 == 0,
             VF_ERR_ALIGN, "vf_attention_gqa_fwd: row pitches and column offsets must be multiples of 8 elements");
  VF_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
               reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gate)) & 15) == 0,
             VF_ERR_ALIGN, "vf_attention_gqa_fwd: pointers must be 16-byte aligned");
  VF_REQUIRE(q_col0 + (long long)(Hq - 1) * q_head_stride + GD <= ldq && (long long)Hkv * GD <= ldk &&
                 (long long)Hkv * GD <= ldv && (long long)Hq * GD <= ldo,
             VF_ERR_ARG, "vf_attention_gqa_fwd: heads do not fit the row pitch");

  GqaParams p{};
  p.B = B; p.S = S; p.Hq = Hq; p.Hkv = Hkv;
  p.n_qt = (S + 127) / 128;
  p.n_bh = B * Hq;
  p.n_items = p.n_bh * p.n_qt;
  p.causal = causal ? 1 : 0;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.gate = reinterpret_cast<const __nv_bfloat16*>(gate);
  p.ldg = ldg; p.gate_col0 = gate_col0; p.gate_head_stride = gate_head_stride;
  p.q_col0 = q_col0; p.q_head_stride = q_head_stride;

  CUtensorMap tmQ, tmK, tmV;
  const uint64_t rows = (uint64_t)B * S;
  {
    uint64_t dims[2] = {(uint64_t)ldq, rows};
    uint64_t strides[1] = {(uint64_t)ldq * 2};
    uint32_t box[2] = {64, 128};
    int e = encode_tmap(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)ldk, rows};
    uint64_t strides[1] = {(uint64_t)ldk * 2};
    uint32_t box[2] = {64, GKT};
    int e = encode_tmap(&tmK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, k, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)ldv, rows};
    uint64_t strides[1] = {(uint64_t)ldv * 2};
    uint32_t box[2] = {64, GKT};
    int e = encode_tmap(&tmV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, v, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  CUtensorMap tmO, tmG;
  {   // 3-D (column, row in sample, sample): boxes that run past the end of a sample are clipped / zero-filled
    uint64_t dims[3] = {(uint64_t)Hq * GD, (uint64_t)S, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)ldo * 2, (uint64_t)S * ldo * 2};
    uint32_t box[3] = {64, 32, 1};
    int e = encode_tmap(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
    tmG = tmO;
    if (gate) {
      VF_REQUIRE(gate_col0 + (long long)(Hq - 1) * gate_head_stride + GD <= ldg, VF_ERR_ARG,
                 "vf_attention_gqa_fwd: gate heads do not fit the row pitch");
      uint64_t gd[3] = {(uint64_t)gate_col0 + (uint64_t)(Hq - 1) * gate_head_stride + GD, (uint64_t)S, (uint64_t)B};
      uint64_t gs[2] = {(uint64_t)ldg * 2, (uint64_t)S * ldg * 2};
      e = encode_tmap(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, gate, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (e) return e;
    }
  }
  const int sms = device_sm_count();
  VF_REQUIRE(sms > 0, VF_ERR_NO_DEVICE, "no CUDA device");
  const int grid = p.n_items < sms ? p.n_items : sms;
  static std::atomic<uint64_t> configured{0};
  if (int e2 = ensure_dynamic_smem(attention_gqa_kernel, GqaSmem::TOTAL, configured)) return e2;
  attention_gqa_kernel<<<grid, G_THREADS, GqaSmem::TOTAL, static_cast<cudaStream_t>(stream)>>>(p, tmQ, tmK, tmV, tmO, tmG);
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
