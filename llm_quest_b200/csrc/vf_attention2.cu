// vf_attention2.cu — fused bidirectional attention for sm_100a, head_dim 64 (bf16 in/out, fp32 softmax): second design.
//
// Replaces F.scaled_dot_product_attention and the head-major transposes around it
// (llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:169-190, vit_attention.py:62-87 in the reference). Q, K, V are read IN
// PLACE from the token-major [B*S, 3*H*64] buffer the QKV GEMM writes (TMA boxes at column offsets h*64, H*64+h*64,
// 2*H*64+h*64); the context is written token-major [B*S, H*64].
//
// At head_dim 64 the exponentials bound this kernel, not the tensor pipe: a 128 x 128 score tile costs 512 tensor cycles
// (QK^T + PV) but 1024 MUFU cycles (16 ex2/clk/SM). The design therefore keeps the MUFU fed and everything else off the
// softmax warps' critical path:
//
//   one persistent CTA per SM, 12 warps; a work item is (sample b, head h, block of 256 queries) = TWO 128-row query
//   tiles ("chains") that share every 128-key K/V tile:
//     warps 0..3 / 4..7   softmax warpgroups of chain 0 / 1, one thread per query row: S -> registers (4 x tcgen05.ld
//                         x32), row max, packed-f32x2 scale, exp2, row sum, P (bf16) -> TMEM; O is only rescaled when the
//                         row max grows by more than 2^8 (otherwise the stale max stays the reference: bf16 P and fp32
//                         O / l keep full relative precision at any common scale); final O / l store
//     warp 8              TMA loader + item scheduler: Q0, Q1 once per item (double-buffered), K and V tiles through
//                         two 3-stage rings
//     warps 9, 10         tcgen05.mma issuers, one per chain:  S_t = Q_t K_j^T (SS, M=128, N=128, K=64)
//                                                              O_t += P_t V_j  (TS: P_t bf16 in TMEM; V MN-major smem)
//   TMEM (512 columns): S_t at [128t, 128t+128), O_t at [256+64t, +64), P_t at [384+64t, +64). P has its OWN columns, so
//   the issuer starts S_t(j+1) = Q_t K_{j+1}^T as soon as the softmax warps have S_t(j) in registers — the QK^T of the
//   next step runs under the exponentials of this one and a chain's step is (tcgen05.ld, exponentials, tcgen05.st), with
//   no MMA round trip in it. The two chains ALTERNATE on the MUFU (ping-pong, enforced by a token: chain t starts the
//   exponentials of a step when chain 1-t has finished its own): left alone they fall into lock step — both in the
//   exponentials (each at half rate), then both in tcgen05.ld / row max / tcgen05.st with the MUFU idle (measured:
//   4800 cycles per step pair, 43 % MUFU) — with the token one chain's loads, maxima and stores run under the other's
//   1024 MUFU cycles.
//   Ragged edges: the last key step only computes ceil(valid/32) 32-key chunks (QK^T with a smaller N, fewer PV slices,
//   fewer exponentials); softmax warps whose 32 rows lie beyond the sequence only keep the barrier protocol going.
//   Scheduling: items are handed out by an atomic counter (the loader fetches one item ahead), ordered (b, h)-major with
//   the query block fastest: the CTAs of the grid work on the same few (b, h) at any time, so K/V tiles are read from
//   DRAM once and from L2 afterwards (the static cost-sorted walk of the first design re-read them per query block:
//   1.77x the algorithmic DRAM bytes at S = 784), and no static assignment has to guess the cost of ragged blocks.
#include "vf_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace vf {
namespace a2 {

constexpr int KT = 128;                      // keys per K/V tile
constexpr int NCH = 2;                       // query tiles (chains) per work item
constexpr int THREADS = 384;
constexpr int KV_STAGES = 3;
constexpr int LOADER_WARP = 8;
constexpr int MMA_WARP = 9;                  // warps 9, 10: chain 0, 1
constexpr int Q_TILE_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int KV_TILE_BYTES = KT * 64 * 2;   // 16 KB
constexpr float TAU = 8.0f;                  // log2 units: P may reach 2^8 before O is rescaled
constexpr int SCHED_SLOTS = 64;

struct Params {
  int B, S, H;
  int n_qblk;      // ceil(S / 256)
  int n_kt;        // ceil(S / KT)
  int n_items;     // B * H * n_qblk
  float scale_log2;
  __nv_bfloat16* out;
  int* sched;      // [2]: next item, CTAs done — both zero before and after every launch
  unsigned long long* trace;   // diagnosis build only: [4 rows][trace_n steps][8] clock64 stamps of block 0
  int trace_first, trace_n;
};

struct Smem {
  static constexpr int Q_OFF = 0;                               // 2 buffers x NCH tiles (the next item's Q loads early)
  static constexpr int K_OFF = 2 * NCH * Q_TILE_BYTES;
  static constexpr int V_OFF = K_OFF + KV_STAGES * KV_TILE_BYTES;
  static constexpr int BAR_OFF = V_OFF + KV_STAGES * KV_TILE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

__device__ int g_sched[SCHED_SLOTS][2];   // zero-initialised per device at module load; every launch leaves its slot zero

// Bounded wait: a protocol bug becomes a trap (CUDA error) instead of a hung GPU. No printf in here: a call inside the
// softmax warps' loop makes ptxas spill every live register around it (caller-saved ABI).
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {     // try_wait suspends the thread for a time slice by itself
    if (++spins > (1u << 24)) __trap();
  }
}

// Order-pinned primitives of the exponential phase (volatile: ptxas keeps them in source order). One softmax warp per
// SM sub-partition holds the MUFU at a time, so there is no second warp to hide the MUFU's latency: the consumers of an
// exponential (row sum, bf16 pack) are issued EXP_DIST pairs behind it by hand. ptxas on its own placed them one pair
// behind, and the lone warp ran at one MUFU per ~15 cycles instead of the pipe's 8.
__device__ __forceinline__ float ex2_pinned(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t fadd2_pinned(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16_pinned(float lo, float hi) {
  uint32_t d;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
constexpr int EXP_DIST = 8;   // pairs between an exponential and its consumers

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// TRACE build: block 0 records clock64 stamps per key step — rows 0/1: softmax warp 0 of chain 0/1, rows 2/3: the issuers
template <bool TRACE>
__device__ __forceinline__ void stamp(const Params& p, int row, uint32_t step, int k) {
  if constexpr (TRACE) {
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
      const int i = static_cast<int>(step) - p.trace_first;
      if (i >= 0 && i < p.trace_n) p.trace[(static_cast<long long>(row) * p.trace_n + i) * 8 + k] = clock64();
    }
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
attention2_kernel(const Params p, const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BAR_OFF);
  uint64_t* q_full = bars + 0;                  // [2]
  uint64_t* q_empty = bars + 2;                 // [2]
  uint64_t* k_full = bars + 4;                  // [KV_STAGES]
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;       // [NCH]  S_t(j) complete (tcgen05.commit)
  uint64_t* s_free = s_full + NCH;              // [NCH]  S_t(j) is in the softmax warps' registers
  uint64_t* p_full = s_free + NCH;              // [NCH]  P_t(j) stored
  uint64_t* pv_done = p_full + NCH;             // [NCH]  O_t += P_t(j) V_j complete (tcgen05.commit)
  uint64_t* turn = pv_done + NCH;               // [NCH]  chain t may run its exponentials (ping-pong token, see below)
  uint64_t* sch_full = turn + NCH;              // [2]
  uint64_t* sch_empty = sch_full + 2;           // [2]
  int* sch_item = reinterpret_cast<int*>(sch_empty + 2);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sch_item + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == LOADER_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], NCH);          // both issuers
      mbar_init(&sch_full[i], 1);
      mbar_init(&sch_empty[i], NCH + NCH * 4);   // issuers + softmax warps
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], NCH);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], NCH);
    }
    for (int t = 0; t < NCH; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);             // one arrival per softmax warp
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
      mbar_init(&turn[t], 4);               // the four softmax warps of the OTHER chain
    }
    for (int i = 0; i < 4; ++i) mbar_arrive(&turn[0]);   // chain 0 goes first
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // the QKV projection's output is complete and visible from here on
  pdl_launch_dependents();

  const int n_bh = p.B * p.H;
  (void)n_bh;
  // item -> (sample, head, query block): (b, h)-major, query block fastest (K/V of a (b, h) stay hot in L2)
  auto decode = [&](int item, int& b, int& h, int& qb) {
    const int bh = item / p.n_qblk;
    qb = item - bh * p.n_qblk;
    b = bh / p.H;
    h = bh - b * p.H;
  };

  if (warp >= NCH * 4) {
    // loader / MMA / idle warps hand registers to the softmax warpgroups: the CTA owns 384 x 168 registers at launch;
    // 128 x 88 + 256 x 208 fits inside that pool
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == LOADER_WARP) {
      // ---------------------------------------------------------------- scheduler + TMA loader
      if (lane == 0) {
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int n = 0;; ++n) {
          const int slot = n & 1;
          wait(&sch_empty[slot], ((n >> 1) & 1) ^ 1);
          int item = atomicAdd(&p.sched[0], 1);
          if (item >= p.n_items) item = -1;
          sch_item[slot] = item;
          mbar_arrive(&sch_full[slot]);
          if (item < 0) break;
          int b, h, qb;
          decode(item, b, h, qb);
          const int row0 = b * p.S;
          const int qbuf = n & 1;
          wait(&q_empty[qbuf], ((n >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qbuf], NCH * Q_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < NCH; ++t)
            tma_load_2d(smem + Smem::Q_OFF + (qbuf * NCH + t) * Q_TILE_BYTES, &tmQ, &q_full[qbuf], h * 64,
                        row0 + qb * (128 * NCH) + t * 128);
          for (int j = 0; j < p.n_kt; ++j) {
            wait(&k_empty[ks], kph ^ 1);
            mbar_expect_tx(&k_full[ks], KV_TILE_BYTES);
            tma_load_2d(smem + Smem::K_OFF + ks * KV_TILE_BYTES, &tmKV, &k_full[ks], p.H * 64 + h * 64, row0 + j * KT);
            if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
            wait(&v_empty[vs], vph ^ 1);
            mbar_expect_tx(&v_full[vs], KV_TILE_BYTES);
            tma_load_2d(smem + Smem::V_OFF + vs * KV_TILE_BYTES, &tmKV, &v_full[vs], 2 * p.H * 64 + h * 64, row0 + j * KT);
            if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
          }
        }
        // the last CTA to run out of items leaves the counters at zero for the next launch
        __threadfence();
        if (atomicAdd(&p.sched[1], 1) == static_cast<int>(gridDim.x) - 1) {
          p.sched[0] = 0;
          p.sched[1] = 0;
        }
      }
    } else if (warp == MMA_WARP || warp == MMA_WARP + 1) {
      // ---------------------------------------------------------------- MMA issuer of chain t
      const int t = warp - MMA_WARP;
      // The whole warp walks the (warp-uniform) schedule so that addresses and descriptors live in uniform registers;
      // one elected lane issues the tcgen05 instructions.
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // P V : V is MN-major
      const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + Smem::Q_OFF));
      const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + Smem::K_OFF));
      const uint64_t v_desc = umma_desc_sw128(smem_u32(smem + Smem::V_OFF));
      constexpr uint64_t QT_DESC = Q_TILE_BYTES >> 4;
      constexpr uint64_t KVT_DESC = KV_TILE_BYTES >> 4;
      const uint32_t s_tm = tmem_base + t * 128, o_tm = tmem_base + 256 + t * 64, p_tm = tmem_base + 384 + t * 64;

      auto issue_s = [&](int kstage, int qbuf, int nkeys) {       // S_t = Q_t K^T over the first nkeys (multiple of 32) keys
        if (elect_one()) {
          const uint32_t idesc_s = umma_idesc_bf16(128, nkeys, 0, 0);   // both operands K-major
          const uint64_t a_ = q_desc + (qbuf * NCH + t) * QT_DESC;
          const uint64_t b_ = k_desc + kstage * KVT_DESC;
#pragma unroll
          for (int k_ = 0; k_ < 4; ++k_)      // head_dim 64 = 4 x K16
            umma_ss(s_tm, a_ + 2 * k_, b_ + 2 * k_, idesc_s, k_ != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int vstage, bool accumulate, int nkeys) {
        if (elect_one()) {
          const uint64_t bv_ = v_desc + vstage * KVT_DESC;
          for (int k_ = 0; k_ < nkeys / 16; ++k_)  // 16 keys per MMA: 8 TMEM columns of bf16x2 / 16 V rows
            umma_ts(o_tm, p_tm + k_ * 8, bv_ + k_ * (2048 >> 4), idesc_o, accumulate || k_ != 0);
          umma_commit(&pv_done[t]);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      auto keys_of = [&](int j) {                                  // keys of step j that exist, rounded up to 32
        const int valid = p.S - j * KT;
        return valid >= KT ? KT : ((valid + 31) & ~31);
      };

      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t g = 0;            // steps issued by this chain so far (phase bookkeeping of s_free / p_full)
      for (int n = 0;; ++n) {
        const int slot = n & 1;
        wait(&sch_full[slot], (n >> 1) & 1);
        const int item = sch_item[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sch_empty[slot]);
        if (item < 0) break;
        int b, h, qb;
        decode(item, b, h, qb);
        const bool mine = qb * (128 * NCH) + t * 128 < p.S;      // tiles beyond the last query only recycle K/V slots
        const int qbuf = n & 1;
        wait(&q_full[qbuf], (n >> 1) & 1);
        // ---- S(0)
        wait(&k_full[ks], kph);
        if (mine) {
          if (g > 0) wait(&s_free[t], (g - 1) & 1);
          tc_fence_after();
          issue_s(ks, qbuf, keys_of(0));
        }
        commit(&k_empty[ks]);
        if (p.n_kt == 1) commit(&q_empty[qbuf]);
        if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
        for (int j = 0; j < p.n_kt; ++j) {
          if (j + 1 < p.n_kt) {
            // ---- S(j+1): as soon as the softmax warps hold S(j) in registers — it runs under their exponentials
            wait(&k_full[ks], kph);
            if (mine) {
              stamp<TRACE>(p, 2 + t, g, 0);
              wait(&s_free[t], g & 1);
              stamp<TRACE>(p, 2 + t, g, 1);
              tc_fence_after();
              issue_s(ks, qbuf, keys_of(j + 1));
              stamp<TRACE>(p, 2 + t, g, 2);
            }
            commit(&k_empty[ks]);
            if (j + 2 == p.n_kt) commit(&q_empty[qbuf]);
            if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
          }
          // ---- O += P(j) V(j)
          wait(&v_full[vs], vph);
          if (mine) {
            stamp<TRACE>(p, 2 + t, g, 3);
            wait(&p_full[t], g & 1);
            stamp<TRACE>(p, 2 + t, g, 4);
            tc_fence_after();
            issue_pv(vs, j > 0, keys_of(j));
            stamp<TRACE>(p, 2 + t, g, 5);
            ++g;
          }
          commit(&v_empty[vs]);
          if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int t = warp >> 2;              // chain 0..1
    const int quarter = warp & 3;         // TMEM lane quarter (== SM sub-partition)
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_sel + t * 128;
    const uint32_t o_addr = tmem_base + lane_sel + 256 + t * 64;
    const uint32_t p_addr = tmem_base + lane_sel + 384 + t * 64;
    const int r_local = quarter * 32 + lane;
    uint32_t g = 0;                       // steps done by this chain (same count as its issuer)
    uint32_t tk = 0;                      // token steps (every key step of every item, also of tiles this chain skips)

    for (int n = 0;; ++n) {
      const int slot = n & 1;
      wait(&sch_full[slot], (n >> 1) & 1);
      const int item = sch_item[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sch_empty[slot]);
      if (item < 0) break;
      int b, h, qb;
      decode(item, b, h, qb);
      const int tile_row0 = qb * (128 * NCH) + t * 128;
      if (tile_row0 >= p.S) {                              // whole tile out of range (uniform per chain): only pass the token on
        for (int j = 0; j < p.n_kt; ++j, ++tk) {
          wait(&turn[t], tk & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&turn[t ^ 1]);
        }
        continue;
      }
      const bool active = tile_row0 + quarter * 32 < p.S;  // this warp owns at least one real query row
      const int q_in_sample = tile_row0 + r_local;

      float m = -INFINITY;   // reference max of the row (possibly stale), raw score units
      float l = 0.f;         // running row sum
      for (int j = 0; j < p.n_kt; ++j, ++g, ++tk) {
        if (quarter == 0) stamp<TRACE>(p, t, g, 0);
        wait(&s_full[t], g & 1);
        if (quarter == 0) stamp<TRACE>(p, t, g, 1);
        tc_fence_after();
        const int valid = p.S - j * KT;                    // keys [0, valid) of this tile exist
        const int nch = valid >= KT ? 4 : ((valid + 31) >> 5);   // 32-key chunks that were computed
        float s[KT];
        if (active) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < nch) tmem_ld_x32(s_addr + c * 32, reinterpret_cast<uint32_t*>(s) + c * 32);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);            // S_t(j+1) may overwrite the columns now
        if (quarter == 0) stamp<TRACE>(p, t, g, 2);
        if (active) {
          if (valid < KT) {
#pragma unroll
            for (int c = 0; c < KT; ++c)
              if (c >= valid) s[c] = -INFINITY;
          }
          // row max of the step (3-input max: 2 elements per issued instruction); the full-tile path carries no guards so
          // that the compiler software-pipelines all 128 elements
          float mx = -INFINITY;
          auto chunk_max = [&](int c) {
            float m0 = fmax3(s[c * 32], s[c * 32 + 1], s[c * 32 + 2]), m1 = fmax3(s[c * 32 + 3], s[c * 32 + 4], s[c * 32 + 5]);
#pragma unroll
            for (int e = 6; e < 30; e += 4) {
              m0 = fmax3(m0, s[c * 32 + e], s[c * 32 + e + 1]);
              m1 = fmax3(m1, s[c * 32 + e + 2], s[c * 32 + e + 3]);
            }
            return fmax3(m0, s[c * 32 + 30], fmaxf(s[c * 32 + 31], m1));
          };
          if (nch == 4) {
            mx = fmax3(chunk_max(0), chunk_max(1), fmaxf(chunk_max(2), chunk_max(3)));
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              if (c < nch) mx = fmaxf(mx, chunk_max(c));
          }
          if (quarter == 0) stamp<TRACE>(p, t, g, 3);
          // the stale max stays the reference unless the new one is more than 2^TAU above it
          const bool grow = mx * p.scale_log2 > m * p.scale_log2 + TAU;   // (first step: m = -inf)
          bool pv_waited = j == 0;
          if (j > 0 && __any_sync(0xffffffffu, grow)) {
            wait(&pv_done[t], (g - 1) & 1);                // PV(j-1) has updated O
            pv_waited = true;
            tc_fence_after();
            const float f = grow ? fast_exp2((m - mx) * p.scale_log2) : 1.0f;
            l *= f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld_x16(o_addr + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
              tmem_st_x16(o_addr + c * 16, o);
            }
          }
          if (grow) m = mx;
          // P = exp2(s*c - m*c) as bf16 pairs; row sum in packed f32x2. All exponentials first (a score register dies as
          // its pair is packed), the stores afterwards: by then PV(j-1), which still reads P(j-1), has long retired and
          // the wait for it costs nothing.
          wait(&turn[t], tk & 1);                          // the other chain has finished its exponentials
          if (quarter == 0) stamp<TRACE>(p, t, g, 4);
          const float mb = m * p.scale_log2;
          const uint64_t sc2 = pack2(p.scale_log2, p.scale_log2), nb2 = pack2(-mb, -mb);
          uint64_t acc0 = pack2(0.f, 0.f), acc1 = acc0;
          if (nch == 4) {
            // software pipeline over the 64 key pairs of the step: pair i is exponentiated while pair i - EXP_DIST is
            // summed and packed; the 16 packed words of a 32-key chunk are stored as soon as the chunk is complete
            uint32_t pk[16];
            float e0[EXP_DIST], e1[EXP_DIST];                // exponentials in flight
            auto issue = [&](int i) {
              float x0, x1;
              unpack2(ffma2(pack2(s[2 * i], s[2 * i + 1]), sc2, nb2), x0, x1);
              e0[i % EXP_DIST] = ex2_pinned(x0);
              e1[i % EXP_DIST] = ex2_pinned(x1);
            };
            auto consume = [&](int k) {
              const float p0 = e0[k % EXP_DIST], p1 = e1[k % EXP_DIST];
              if (k & 1) acc1 = fadd2_pinned(acc1, pack2(p0, p1));
              else acc0 = fadd2_pinned(acc0, pack2(p0, p1));
              pk[k & 15] = pack_bf16_pinned(p0, p1);
            };
#pragma unroll
            for (int i = 0; i < EXP_DIST; ++i) issue(i);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { issue(16 * c + 8 + i); consume(16 * c + i); }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (c < 3) issue(16 * c + 16 + i);
                consume(16 * c + 8 + i);
              }
              if (c == 0 && !pv_waited) wait(&pv_done[t], (g - 1) & 1);   // PV(j-1), which still read P(j-1), has long retired
              if (c == 0 && quarter == 0) stamp<TRACE>(p, t, g, 5);
              tmem_st_x16(p_addr + c * 16, pk);
            }
          } else {
            // ragged last key step: chunk by chunk, no pipeline
            if (!pv_waited) wait(&pv_done[t], (g - 1) & 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              if (c < nch) {
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
                tmem_st_x16(p_addr + c * 16, pk);
              }
            }
          }
          float a0, a1;
          unpack2(fadd2(acc0, acc1), a0, a1);
          l += a0 + a1;
          __syncwarp();
          if (lane == 0) mbar_arrive(&turn[t ^ 1]);        // the MUFU is the other chain's now
          if (quarter == 0) stamp<TRACE>(p, t, g, 6);
          tmem_st_wait();
          if (quarter == 0) stamp<TRACE>(p, t, g, 7);
        } else {
          if (j > 0) wait(&pv_done[t], (g - 1) & 1);       // keep the phase bookkeeping of the chain in step
          wait(&turn[t], tk & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&turn[t ^ 1]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }

      // ---- item epilogue: O_t / l -> global
      wait(&pv_done[t], (g - 1) & 1);
      tc_fence_after();
      if (active) {
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
      }
      tc_fence_before();   // the next item's first PV (accumulate = 0) is ordered behind these loads by p_full
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace a2
}  // namespace vf

using namespace vf;

static unsigned long long* g_trace_buf = nullptr;
static int g_trace_first = 0, g_trace_n = 0;

extern "C" int vf_attention_set_trace(void* buf, int32_t first_step, int32_t n_steps) {
  g_trace_buf = reinterpret_cast<unsigned long long*>(buf);
  g_trace_first = first_step;
  g_trace_n = buf ? n_steps : 0;
  return VF_OK;
}

int vf_attention2_launch(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, float scale, void* stream) {
  a2::Params p{};
  p.trace = g_trace_buf; p.trace_first = g_trace_first; p.trace_n = g_trace_n;
  p.B = B; p.S = S; p.H = H;
  p.n_qblk = (S + 128 * a2::NCH - 1) / (128 * a2::NCH);
  p.n_kt = (S + a2::KT - 1) / a2::KT;
  p.n_items = B * H * p.n_qblk;
  const int sms = device_sm_count();
  VF_REQUIRE(sms > 0, VF_ERR_NO_DEVICE, "no CUDA device");
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  // one counter pair per launch in flight (round robin over 64 slots; a captured graph keeps the slot it was captured with)
  static std::atomic<unsigned> launch_no{0};
  static std::atomic<int*> sched_of_device[64];
  int dev = 0;
  VF_CUDA(cudaGetDevice(&dev));
  int* sched_base = sched_of_device[dev & 63].load(std::memory_order_acquire);
  if (sched_base == nullptr) {
    VF_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&sched_base), a2::g_sched));
    sched_of_device[dev & 63].store(sched_base, std::memory_order_release);
  }
  p.sched = sched_base + 2 * (launch_no.fetch_add(1, std::memory_order_relaxed) % a2::SCHED_SLOTS);
  CUtensorMap tmQ, tmKV;
  uint64_t dims[2] = {(uint64_t)3 * H * 64, (uint64_t)B * S};
  uint64_t strides[1] = {(uint64_t)3 * H * 64 * 2};
  uint32_t boxq[2] = {64, 128};
  uint32_t boxkv[2] = {64, a2::KT};
  int e = encode_tmap(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, boxq, CU_TENSOR_MAP_SWIZZLE_128B);
  if (e) return e;
  e = encode_tmap(&tmKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, boxkv, CU_TENSOR_MAP_SWIZZLE_128B);
  if (e) return e;
  using kern_t = void (*)(const a2::Params, const CUtensorMap, const CUtensorMap);
  const int tr = g_trace_buf != nullptr ? 1 : 0;          // the trace build only runs while a trace buffer is set
  const kern_t kern = tr ? static_cast<kern_t>(a2::attention2_kernel<true>) : static_cast<kern_t>(a2::attention2_kernel<false>);
  static std::atomic<uint64_t> configured[2];
  if (int e2 = ensure_dynamic_smem(kern, a2::Smem::TOTAL, configured[tr])) return e2;
  const int grid = p.n_items < sms ? p.n_items : sms;
  VF_CUDA(launch_pdl(kern, dim3(grid), dim3(a2::THREADS), a2::Smem::TOTAL, static_cast<cudaStream_t>(stream), 1, p, tmQ, tmKV));
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}
