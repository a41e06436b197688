// vf_attention.cu — fused bidirectional attention for sm_100a (head_dim 64, bf16, fp32 softmax).
//
// Replaces F.scaled_dot_product_attention and the head-major transposes around it
// (llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:169-190, vit_attention.py:62-87 in the reference).
// Q, K, V are read IN PLACE from the token-major [B*S, 3*H*64] buffer the QKV GEMM writes (TMA boxes
// at column offsets h*64, H*64+h*64, 2*H*64+h*64), and the context is written token-major
// [B*S, H*64] — no head-major copy exists anywhere.
//
// At head_dim 64 the tensor pipe is NOT what bounds this kernel: one 128x64 score tile costs 256 MMA cycles (QK^T +
// PV) but 512 MUFU cycles (16 ex2/clk/SM) and about as many issue cycles of the softmax warps (the FMA-pipe
// polynomial exponential made it slower: issue slots are the scarcer of the two). The design goal is therefore to keep
// every SM sub-partition busy with softmax work, which needs many independent chains:
//
//   one persistent CTA per SM; a work item is (sample b, head h, block of 512 queries) = FOUR
//   128-row query tiles ("chains") that share every 64-key K/V tile:
//     warps 0..15   four softmax warpgroups (one per query tile), one thread per query row: first-tile row max,
//                   packed-f32x2 scale/exp2/row sum, P->TMEM, redo of a step with a fresh max + O rescale only when
//                   its row sum runs past 2^60, final O/l store
//     warp 16       TMA loader: Q0..Q3 once per item (double-buffered); K and V tiles through two 4-stage rings
//     warps 17,18   tcgen05.mma issuers (tiles {0,1} and {2,3}), each walking its two tiles as independent chains:
//                                        S_t = Q_t K_j^T  (SS, M=128, N=64,  K=64)
//                                        O_t += P_t V_j   (TS: P_t bf16 in TMEM; V MN-major smem)
//                   Measured: one issuer needs ~380 cycles to issue the 8 small MMAs of a tile-step,
//                   more than the 256 tensor cycles they take, so a single issuer starves all four
//                   softmax chains; two issuers halve that. They carry the HIGHEST warp ids of their
//                   sub-partitions because the warp scheduler favours high warp ids.
//   TMEM (512 cols): S_t at [64t, 64t+64), P_t aliases the first 32 columns of S_t (64 bf16),
//                    O_t at [256+64t, 256+64t+64).
//   Each softmax warp sits on one SM sub-partition together with the three warps that own the same
//   lane quarter of the other tiles, so while one waits on its MMAs the others keep the sub-partition fed.
//   A walker issues S_t(j+1) right behind PV_t(j); tcgen05.mma instructions of one thread retire
//   in order, which is what makes the S/P aliasing and the in-place O rescale race-free.
//   S <= 256 (pair mode): an item is a PAIR of (sample, head); chains 0,1 serve the first, 2,3 the second, each
//   member with its own two-stage K/V ring.
//
//   Tried and measured slower on B200 (round 1): THREE chains with P in its own TMEM columns so
//   that S_t(j+1) is computed while S_t(j) is exponentiated (one issuer warp per tile and product, exponentials
//   kept in registers until PV_t(j-1) retires): 2.96 ms vs 2.23 ms at S=6272 — a chain's step is bound by the
//   softmax warp's own ~2500-cycle serial path and by the issue cost of the small MMAs (~60 cycles per
//   tcgen05.mma, ~100-200 per barrier operation), not by the QK^T round trip, so four chains beat three.
//   Also: P in shared memory with S_t(j+1) issued as soon as S_t(j) is in registers (3.57 ms vs 3.28 ms per 12
//   layers); more variants are listed above the kernel.
#include "vf_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace vf {

constexpr int NQ = 4;                       // query tiles per work item
constexpr int KT = 64;                      // keys per K/V tile
constexpr int ATT_THREADS = 128 + NQ * 128;
constexpr int KV_STAGES = 4;
constexpr int LOADER_WARP = NQ * 4;      // warp 16 (sub-partition 0)
constexpr int MMA_WARP = NQ * 4 + 1;     // warps 17, 18 (sub-partitions 1, 2): two tiles each
constexpr int Q_TILE_BYTES = 128 * 64 * 2;  // 16 KB
constexpr int KV_TILE_BYTES = KT * 64 * 2;  // 8 KB

// FLAGS bit 1 = trace build: block 0 records clock64 stamps per chain and key step into AttnParams::trace
// (vf_attention_set_trace); diagnosis only, never the default.
constexpr int ATT_TRACE = 2;
// The softmax step is "lean": packed f32x2 arithmetic (FFMA2 / FADD2: the softmax warps are bound by issue
// slots, not only by the MUFU) and NO per-step row max: the max of the first key tile stays the reference
// and a step is redone with a fresh max only when its row sum shows that an exponent ran away (> 2^60).
// bf16 P and fp32 O/l keep full relative precision at any common scale, so the result is unchanged.
constexpr int TRACE_STAMPS = 8;

struct AttnParams {
  int B, S, H;
  int n_qblk;      // ceil(S / (128*NQ))
  int n_kt;        // ceil(S / KT)
  int n_items;     // B * H * n_qblk
  int n_bh;        // B * H
  int pair;        // S <= 256 (at most two query tiles): an item is a PAIR of (sample, head) — chains 0,1 run the
                   // first, chains 2,3 the second, each with its own 2-stage K/V ring (all four chains stay busy)
  float scale_log2;
  unsigned long long* trace;   // [20 rows][trace_n steps][TRACE_STAMPS] clock64 stamps of block 0, or nullptr
  int trace_first, trace_n;
  __nv_bfloat16* out;
};

struct AttnSmem {
  static constexpr int Q_OFF = 0;                               // 2 buffers x NQ tiles (next item's Q loads early)
  static constexpr int K_OFF = 2 * NQ * Q_TILE_BYTES;
  static constexpr int V_OFF = K_OFF + KV_STAGES * KV_TILE_BYTES;
  static constexpr int O_OFF = V_OFF + KV_STAGES * KV_TILE_BYTES;   // one 32-row x 32-column bf16 staging slot per softmax warp
  static constexpr int BAR_OFF = O_OFF + NQ * 4 * 2048;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

__device__ __forceinline__ void att_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;   // fast path: no watchdog bookkeeping
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("vf_attention: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

template <int FLAGS>
__device__ __forceinline__ void att_trace(const AttnParams& p, int row, int step, int k) {
  if constexpr (FLAGS & ATT_TRACE) {
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
      const int i = step - p.trace_first;
      if (i >= 0 && i < p.trace_n) p.trace[(static_cast<long long>(row) * p.trace_n + i) * TRACE_STAMPS + k] = clock64();
    }
  }
}

// Work items are ordered by decreasing cost: first every (b,h)'s full 4-tile block(s), the ragged last
// block of each (b,h) at the end, so the static round-robin tail is short.
// Boustrophedon ("snake") walk of the cost-sorted item list: round r hands items r*G .. r*G+G-1 to the
// CTAs in ascending order when r is even and descending when r is odd, so the CTAs that received the
// extra expensive item of a partial round get the cheap end of the next one (max load 38 instead of
// 39 tile-units at S=784, B*H=768, G=148; the mean is 36.3).
struct ItemIter {
  int r, n;
  __device__ explicit ItemIter(int n_items) : r(0), n(n_items) {}
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

// (sample, head)-major with the query block fastest: the CTAs of the grid work on neighbouring items at any time, i.e. on
// the query blocks of the same few (sample, head) pairs, so a pair's K/V tiles come from DRAM once and from L2 for its
// other query blocks. (Round 1 walked all first blocks, then all second blocks, ...: every query block re-read K/V from
// DRAM — 1.77x the algorithmic bytes at S = 784, ncu.) The snake walk below keeps the load balanced: the grid size is
// even, so consecutive rounds hand a CTA items of alternating parity (full / ragged query block at S = 784).
__device__ __forceinline__ void decode_item(const AttnParams& p, int item, int& b, int& h, int& qb) {
  const int bh = item / p.n_qblk;
  qb = item - bh * p.n_qblk;
  b = bh / p.H;
  h = bh - b * p.H;
}

// Pair mode: (sample, head) index of member m (0 or 1) of an item, or -1 when the last item has no second member.
__device__ __forceinline__ int pair_bh(const AttnParams& p, int item, int m) {
  const int bh = 2 * item + m;
  return bh < p.n_bh ? bh : -1;
}

// The two MMA warps walk their two query tiles each as INDEPENDENT chains (any-order issue: whichever tile's P
// is ready is served first) instead of tile 0 then tile 1 of every key step: the in-order walk couples the
// softmax chains (a chain that runs ahead has to wait for its pair). Measured slower and removed (round 1):
// a FIFO token that serialises the exp phases per sub-partition (hand-off latency eats the gain: 2675 vs
// 2459 us at S=6272), a one-time 250-800 cycle stagger of the four chains (2620-2930 vs 2340 us), polling with
// mbarrier.test_wait instead of the suspending try_wait (2522 vs 2328 us), skipping the exponentials of masked
// 16-key chunks and running the ragged last query tile on one lane quarter only (no change at S=784: a key
// step there is bound by the chain's serial latency, not by the amount of exponentials), and evaluating every 4th /
// 3rd / 2nd pair of exponentials on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, packed f32x2:
// 2414 / 2501 / 2813 us against 2296 at S=6272 — the softmax warps are short of issue slots, not of MUFU cycles), and
// ex2.approx.ftz.bf16x2 for both exponentials of a pair (ptxas splits it into two MUFU.EX2.BF16: no packed MUFU on
// sm_100a; 2479 us and twice the error).

template <int FLAGS>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                 const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmO) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::BAR_OFF);
  uint64_t* q_full = bars + 0;                  // [2]
  uint64_t* q_empty = bars + 2;                 // [2]
  uint64_t* k_full = bars + 4;                  // [KV_STAGES]
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;       // [NQ]
  uint64_t* p_full = s_full + NQ;               // [NQ]
  uint64_t* o_full = p_full + NQ;               // [NQ]
  uint64_t* o_empty = o_full + NQ;              // [NQ]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + NQ);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == LOADER_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], NQ);   // every tile walker
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], p.pair ? 2 : NQ);   // the walkers that consume the stage
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], p.pair ? 2 : NQ);
    }
    for (int t = 0; t < NQ; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);   // one arrival per softmax warp
      mbar_init(&o_full[t], 1);
      mbar_init(&o_empty[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // the QKV projection's output is complete and visible from here on
  pdl_launch_dependents();

  if (warp >= NQ * 4) {
    // loader / MMA / idle warps: hand registers to the softmax warpgroups. Budget: the CTA owns
    // 640 x 96 registers at launch; 128 x 56 + 512 x 104 fits inside that pool (setmaxnreg.inc can
    // only draw from what the CTA already holds — asking for more deadlocks the warpgroup).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == LOADER_WARP) {
      // ---------------------------------------------------------------- TMA loader
      if (lane == 0) {
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0, qph = 0;   // qph: bit i = phase of q buffer i
        int qbuf = 0;
        int item;
        if (p.pair) {
          // ring m (stages 2m, 2m+1) belongs to member m; ks/vs/kph/vph hold one 2-bit-wide lane per ring
          int kst[2] = {0, 0}, vst[2] = {0, 0};
          uint32_t kp[2] = {0, 0}, vp[2] = {0, 0};
          for (ItemIter it(p.n_items); it.next(item);) {
            att_wait(&q_empty[qbuf], ((qph >> qbuf) & 1) ^ 1);
            mbar_expect_tx(&q_full[qbuf], NQ * Q_TILE_BYTES);
            for (int t = 0; t < NQ; ++t) {
              int bh = pair_bh(p, item, t >> 1);
              if (bh < 0) bh = 2 * item;          // no second member: the tiles are loaded but never used
              tma_load_2d(smem + AttnSmem::Q_OFF + (qbuf * NQ + t) * Q_TILE_BYTES, &tmQ, &q_full[qbuf], (bh % p.H) * 64,
                          (bh / p.H) * p.S + (t & 1) * 128);
            }
            qph ^= 1u << qbuf;
            qbuf ^= 1;
            for (int j = 0; j < p.n_kt; ++j) {
              for (int m = 0; m < 2; ++m) {
                const int bh = pair_bh(p, item, m);
                if (bh < 0) continue;
                const int row0 = (bh / p.H) * p.S, h = bh % p.H;
                const int sk = 2 * m + kst[m], sv = 2 * m + vst[m];
                att_wait(&k_empty[sk], kp[m] ^ 1);
                mbar_expect_tx(&k_full[sk], KV_TILE_BYTES);
                tma_load_2d(smem + AttnSmem::K_OFF + sk * KV_TILE_BYTES, &tmKV, &k_full[sk], p.H * 64 + h * 64, row0 + j * KT);
                if (++kst[m] == 2) { kst[m] = 0; kp[m] ^= 1; }
                att_wait(&v_empty[sv], vp[m] ^ 1);
                mbar_expect_tx(&v_full[sv], KV_TILE_BYTES);
                tma_load_2d(smem + AttnSmem::V_OFF + sv * KV_TILE_BYTES, &tmKV, &v_full[sv], 2 * p.H * 64 + h * 64, row0 + j * KT);
                if (++vst[m] == 2) { vst[m] = 0; vp[m] ^= 1; }
              }
            }
          }
        } else
        for (ItemIter it(p.n_items); it.next(item);) {
          int b, h, qb;
          decode_item(p, item, b, h, qb);
          const int row0 = b * p.S;
          att_wait(&q_empty[qbuf], ((qph >> qbuf) & 1) ^ 1);
          mbar_expect_tx(&q_full[qbuf], NQ * Q_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < NQ; ++t)
            tma_load_2d(smem + AttnSmem::Q_OFF + (qbuf * NQ + t) * Q_TILE_BYTES, &tmQ, &q_full[qbuf], h * 64,
                        row0 + qb * (128 * NQ) + t * 128);
          qph ^= 1u << qbuf;
          qbuf ^= 1;
          for (int j = 0; j < p.n_kt; ++j) {
            att_wait(&k_empty[ks], kph ^ 1);
            mbar_expect_tx(&k_full[ks], KV_TILE_BYTES);
            tma_load_2d(smem + AttnSmem::K_OFF + ks * KV_TILE_BYTES, &tmKV, &k_full[ks], p.H * 64 + h * 64,
                        row0 + j * KT);
            if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
            att_wait(&v_empty[vs], vph ^ 1);
            mbar_expect_tx(&v_full[vs], KV_TILE_BYTES);
            tma_load_2d(smem + AttnSmem::V_OFF + vs * KV_TILE_BYTES, &tmKV, &v_full[vs], 2 * p.H * 64 + h * 64,
                        row0 + j * KT);
            if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
          }
        }
      }
    } else if (warp == MMA_WARP || warp == MMA_WARP + 1) {
      // ---------------------------------------------------------------- MMA issuers
      const int t_lo = (warp - MMA_WARP) * 2;   // this issuer owns query tiles t_lo, t_lo+1
      // The whole warp walks the (warp-uniform) schedule so that addresses and descriptors live in
      // uniform registers; one elected lane issues the tcgen05 instructions.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, KT, 0, 0);   // Q K^T : both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // P V   : V is MN-major
      const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + AttnSmem::Q_OFF));
      const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + AttnSmem::K_OFF));
      const uint64_t v_desc = umma_desc_sw128(smem_u32(smem + AttnSmem::V_OFF));
      constexpr uint64_t QT_DESC = Q_TILE_BYTES >> 4;
      constexpr uint64_t KVT_DESC = KV_TILE_BYTES >> 4;

      auto issue_s_q = [&](int t, int kstage, int qb_) {
        if (elect_one()) {
          const uint64_t a_ = q_desc + (qb_ * NQ + t) * QT_DESC;
          const uint64_t b_ = k_desc + kstage * KVT_DESC;
#pragma unroll
          for (int k_ = 0; k_ < 4; ++k_)      // head_dim 64 = 4 x K16
            umma_ss(tmem_base + t * 64, a_ + 2 * k_, b_ + 2 * k_, idesc_s, k_ != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto issue_pv_s_q = [&](int t, int vstage, bool accumulate, bool with_s, int kstage, int qb_) {
        if (elect_one()) {
          const uint64_t bv_ = v_desc + vstage * KVT_DESC;
#pragma unroll
          for (int k_ = 0; k_ < KT / 16; ++k_)  // 16 keys per MMA: 8 TMEM columns of bf16x2 / 16 V rows
            umma_ts(tmem_base + 256 + t * 64, tmem_base + t * 64 + k_ * 8, bv_ + k_ * (2048 >> 4), idesc_o,
                    accumulate || k_ != 0);
          if (with_s) {
            const uint64_t a_ = q_desc + (qb_ * NQ + t) * QT_DESC;
            const uint64_t bk_ = k_desc + kstage * KVT_DESC;
#pragma unroll
            for (int k_ = 0; k_ < 4; ++k_)
              umma_ss(tmem_base + t * 64, a_ + 2 * k_, bk_ + 2 * k_, idesc_s, k_ != 0);
            umma_commit(&s_full[t]);
          }
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };

      {
        // ---- two independent tile walkers, served in whatever order their barriers complete
        struct Walk {
          ItemIter it;
          int t, nt, j, qbuf, ks, vs;          // j = -1: S(t,0) of the current item is the next action
          int base, depth;                     // K/V ring of this walker: stages [base, base + depth)
          bool member;                         // pair mode: the walker's (sample, head) exists in this item
          uint32_t qph, kph, vph, pph, oeph;   // qph: bit i = phase of q buffer i
          bool done;
          int gstep;
          __device__ Walk(int n_items) : it(n_items), gstep(0) {}
        };
        auto next_item = [&](Walk& w) {
          int item;
          if (!w.it.next(item)) { w.done = true; return; }
          if (p.pair) {
            w.member = pair_bh(p, item, w.t >> 1) >= 0;
            w.nt = (p.S + 127) / 128;          // tiles per member (1 or 2); the walker's tile inside it is t & 1
          } else {
            int b_, h_, qb_;
            decode_item(p, item, b_, h_, qb_);
            int nt = (p.S - qb_ * (128 * NQ) + 127) / 128;
            w.nt = nt > NQ ? NQ : nt;
            w.member = true;
          }
          w.j = -1;
        };
        // one action of walker w if everything it needs has arrived; never blocks
        auto advance = [&](Walk& w) -> bool {
          const int t = w.t;
          // tiles beyond the item's last query tile only recycle K/V slots
          const bool mine = (p.pair ? (t & 1) : t) < w.nt;
          if (!w.member) {                     // pair mode, last item without a second member: only release the Q buffer
            if (!mbar_try_wait(&q_full[w.qbuf], (w.qph >> w.qbuf) & 1)) return false;
            commit(&q_empty[w.qbuf]);
            w.qph ^= 1u << w.qbuf;
            w.qbuf ^= 1;
            next_item(w);
            return true;
          }
          if (w.j < 0) {
            if (!mbar_try_wait(&q_full[w.qbuf], (w.qph >> w.qbuf) & 1)) return false;
            if (!mbar_try_wait(&k_full[w.ks], w.kph)) return false;
            tc_fence_after();
            att_trace<FLAGS>(p, 16 + t, w.gstep, 2);
            if (mine) issue_s_q(t, w.ks, w.qbuf);
            att_trace<FLAGS>(p, 16 + t, w.gstep, 3);
            commit(&k_empty[w.ks]);
            if (p.n_kt == 1) commit(&q_empty[w.qbuf]);
            if (++w.ks == w.base + w.depth) { w.ks = w.base; w.kph ^= 1; }
            w.j = 0;
            return true;
          }
          const bool more = (w.j + 1 < p.n_kt);
          if (!mbar_try_wait(&v_full[w.vs], w.vph)) return false;
          if (more && !mbar_try_wait(&k_full[w.ks], w.kph)) return false;
          if (mine) {
            if (!mbar_try_wait(&p_full[t], w.pph)) return false;
            if (w.j == 0 && !mbar_try_wait(&o_empty[t], w.oeph ^ 1)) return false;
            w.pph ^= 1;
            if (w.j == 0) w.oeph ^= 1;
            tc_fence_after();
            att_trace<FLAGS>(p, 16 + t, w.gstep, 0);
            issue_pv_s_q(t, w.vs, w.j > 0, more, w.ks, w.qbuf);
            att_trace<FLAGS>(p, 16 + t, w.gstep, 1);
            ++w.gstep;
          }
          commit(&v_empty[w.vs]);
          if (++w.vs == w.base + w.depth) { w.vs = w.base; w.vph ^= 1; }
          if (more) {
            commit(&k_empty[w.ks]);
            if (w.j + 2 == p.n_kt) commit(&q_empty[w.qbuf]);
            if (++w.ks == w.base + w.depth) { w.ks = w.base; w.kph ^= 1; }
            ++w.j;
          } else {
            if (mine) commit(&o_full[t]);
            w.qph ^= 1u << w.qbuf;
            w.qbuf ^= 1;
            next_item(w);
          }
          return true;
        };
        Walk w0(p.n_items), w1(p.n_items);
        Walk* ws[2] = {&w0, &w1};
        for (int i = 0; i < 2; ++i) {
          Walk& w = *ws[i];
          w.t = t_lo + i; w.qbuf = 0;
          w.base = p.pair ? 2 * (w.t >> 1) : 0;
          w.depth = p.pair ? 2 : KV_STAGES;
          w.ks = w.base; w.vs = w.base;
          w.qph = 0; w.kph = 0; w.vph = 0; w.pph = 0; w.oeph = 0; w.done = false;
          next_item(w);
        }
        long long last = clock64();
        while (!(w0.done && w1.done)) {
          bool prog = false;
          if (!w0.done) prog |= advance(w0);
          if (!w1.done) prog |= advance(w1);
          if (prog) last = clock64();
          else if (clock64() - last > 4000000000LL) {
            printf("vf_attention: MMA walker stalled (block %d warp %d)\n", blockIdx.x, warp);
            __trap();
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int t = warp >> 2;              // query tile 0..3
    const int quarter = warp & 3;         // TMEM lane quarter (== SM sub-partition)
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_sel + t * 64;
    const uint32_t o_addr = tmem_base + lane_sel + 256 + t * 64;
    uint32_t sph = 0, oph = 0;
    int gstep = 0;   // key steps done by this warp (trace index)

    int item;
    for (ItemIter it(p.n_items); it.next(item);) {
      int b, h, tile_row0;
      if (p.pair) {   // chains 0,1 -> member 0, chains 2,3 -> member 1; tile inside the member = t & 1
        const int bh = pair_bh(p, item, t >> 1);
        tile_row0 = (t & 1) * 128;
        if (bh < 0 || tile_row0 >= p.S) continue;
        b = bh / p.H;
        h = bh - b * p.H;
      } else {
        int qb;
        decode_item(p, item, b, h, qb);
        tile_row0 = qb * (128 * NQ) + t * 128;
        if (tile_row0 >= p.S) continue;                // whole tile out of range (uniform per warp)
      }

      float m = -INFINITY;   // running (possibly stale) row max, raw score units
      float l = 0.f;         // running row sum
      for (int j = 0; j < p.n_kt; ++j) {
        att_trace<FLAGS>(p, warp, gstep, 0);
        att_wait(&s_full[t], sph); sph ^= 1;
        tc_fence_after();
        att_trace<FLAGS>(p, warp, gstep, 1);
        // All MMAs issued before S_t(j) — in particular PV_t(j-1) — have retired: O_t is stable
        // until this warpgroup arrives on p_full[t].
        float s[KT];
        tmem_ld_x32(s_addr, reinterpret_cast<uint32_t*>(s));
        tmem_ld_x32(s_addr + 32, reinterpret_cast<uint32_t*>(s) + 32);
        tmem_ld_wait();
        att_trace<FLAGS>(p, warp, gstep, 2);
        const int valid = p.S - j * KT;  // keys [0, valid) of this tile exist
        if (valid < KT) {
#pragma unroll
          for (int c = 0; c < KT; ++c)
            if (c >= valid) s[c] = -INFINITY;
        }
        auto row_max = [&]() {
          float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
          for (int c = 4; c < KT; c += 2) {
            mx0 = fmaxf(mx0, s[c]);
            mx1 = fmaxf(mx1, s[c + 1]);
          }
          return fmaxf(mx0, mx1);
        };
        auto rescale_o = [&](float f) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t o[16];
            tmem_ld_x16(o_addr + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st_x16(o_addr + c * 16, o);
          }
        };
        // P_t = exp2(s*c - m*c) as bf16 pairs over the first 32 columns of S_t; returns the row sum
        auto exp_store = [&](float mb) {
          const uint64_t sc2 = pack2(p.scale_log2, p.scale_log2), nb2 = pack2(-mb, -mb);
          uint64_t acc0 = pack2(0.f, 0.f), acc1 = acc0;
#pragma unroll
          for (int c = 0; c < KT / 32; ++c) {
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
            tmem_st_x16(s_addr + c * 16, pk);
          }
          float a0, a1;
          unpack2(fadd2(acc0, acc1), a0, a1);
          return a0 + a1;
        };
        if (j == 0) m = row_max();
        att_trace<FLAGS>(p, warp, gstep, 3);
        float ssum = exp_store(m * p.scale_log2);
        // runaway exponent (sum beyond 2^60, inf or NaN) in any row of the warp: redo the step with a fresh max
        if (j > 0 && __any_sync(0xffffffffu, !(ssum <= 0x1p60f))) {
          const float mx = row_max();
          const bool grow = mx > m;
          const float f = grow ? fast_exp2((m - mx) * p.scale_log2) : 1.0f;
          if (grow) { m = mx; l *= f; }
          rescale_o(f);
          ssum = exp_store(m * p.scale_log2);
        }
        l += ssum;
        att_trace<FLAGS>(p, warp, gstep, 4);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        att_trace<FLAGS>(p, warp, gstep, 5);
        ++gstep;
      }

      // ---- item epilogue: O_t / l -> global. Every thread owns one row of 64 bf16 (128 bytes, rows 1536 bytes apart): written
      // straight from registers that is 8 x 32 scattered 16-byte stores per warp, and the 16 warps of a CTA finish their items
      // together — ~4000 cycles of load/store-unit time per item (a third of a key step per step at S = 784). Instead each warp
      // stages its 32 rows x 32 columns in shared memory (64-byte swizzle) and one TMA store writes them as full lines; the
      // output map is 3-D (column, row in sample, sample) so that rows past the end of the sample are clipped, not written
      // into the next sample.
      att_wait(&o_full[t], oph); oph ^= 1;
      tc_fence_after();
      att_trace<FLAGS>(p, warp, gstep - 1, 6);
      const float inv = 1.0f / l;
      const bool warp_ok = tile_row0 + quarter * 32 < p.S;   // at least one row of this warp exists
      uint8_t* stage = smem + AttnSmem::O_OFF + warp * 2048;
      const uint32_t srow = smem_u32(stage) + lane * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld_x32(o_addr + c * 32, o);
        tmem_ld_wait();
        if (warp_ok) {
          if (lane == 0) tma_store_wait_read<0>();   // the previous store out of this slot has left shared memory
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t w0 = pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            const uint32_t w1 = pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            const uint32_t w2 = pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            const uint32_t w3 = pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((q ^ ((lane >> 1) & 3)) << 4)), "r"(w0), "r"(w1),
                         "r"(w2), "r"(w3)
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmO, stage, h * 64 + c * 32, tile_row0 + quarter * 32, b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[t]);
      att_trace<FLAGS>(p, warp, gstep - 1, 7);
    }
  }

  if (warp < NQ * 4 && lane == 0) tma_store_wait<0>();   // shared memory must outlive the last output stores
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


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

extern "C" int vf_attention_fwd(const void* qkv, void* out, int32_t B, int32_t S, int32_t H,
                                float scale, void* stream) {
  VF_REQUIRE(qkv && out, VF_ERR_ARG, "vf_attention_fwd: null pointer");
  VF_REQUIRE(B > 0 && S > 0 && H > 0, VF_ERR_ARG, "vf_attention_fwd: bad shape B=%d S=%d H=%d", B, S, H);
  VF_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             VF_ERR_ALIGN, "vf_attention_fwd: pointers must be 16-byte aligned");
  VF_REQUIRE((long long)B * S < (1ll << 31), VF_ERR_ARG, "vf_attention_fwd: B*S too large");

  AttnParams p{};
  p.B = B; p.S = S; p.H = H;
  p.n_qblk = (S + 128 * NQ - 1) / (128 * NQ);
  p.n_kt = (S + KT - 1) / KT;
  p.n_bh = B * H;
  p.n_items = p.n_bh * p.n_qblk;
  const int sms = device_sm_count();
  VF_REQUIRE(sms > 0, VF_ERR_NO_DEVICE, "no CUDA device");
  // at most two query tiles per (sample, head): two of them share an item — once there are more (sample, head)
  // pairs than SMs (below that, pairing would only leave SMs idle)
  p.pair = (S <= 256 && p.n_bh > sms) ? 1 : 0;
  if (p.pair) p.n_items = (p.n_bh + 1) / 2;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  CUtensorMap tmQ, tmKV, tmO;
  uint64_t dims[2] = {(uint64_t)3 * H * 64, (uint64_t)B * S};
  uint64_t strides[1] = {(uint64_t)3 * H * 64 * 2};
  uint32_t boxq[2] = {64, 128};
  uint32_t boxkv[2] = {64, KT};
  int e = encode_tmap(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, boxq, CU_TENSOR_MAP_SWIZZLE_128B);
  if (e) return e;
  e = encode_tmap(&tmKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, boxkv, CU_TENSOR_MAP_SWIZZLE_128B);
  if (e) return e;

  {
    uint64_t od[3] = {(uint64_t)H * 64, (uint64_t)S, (uint64_t)B};
    uint64_t os[2] = {(uint64_t)H * 64 * 2, (uint64_t)S * H * 64 * 2};
    uint32_t ob[3] = {32, 32, 1};
    e = encode_tmap(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, od, os, ob, CU_TENSOR_MAP_SWIZZLE_64B);
    if (e) return e;
  }
  // VF_ATTN_FLAGS=2 selects the trace build (see vf_attention_set_trace)
  static int flags = -1;
  if (flags < 0) {
    const char* e_ = getenv("VF_ATTN_FLAGS");
    flags = (e_ && (atoi(e_) & ATT_TRACE)) ? 1 : 0;
  }
  p.trace = g_trace_buf;
  p.trace_first = g_trace_first;
  p.trace_n = g_trace_n;
  using kern_t = void (*)(const AttnParams, const CUtensorMap, const CUtensorMap, const CUtensorMap);
  static const kern_t kerns[2] = {attention_kernel<0>, attention_kernel<ATT_TRACE>};
  static std::atomic<uint64_t> configured[2];
  if (int e2 = ensure_dynamic_smem(kerns[flags], AttnSmem::TOTAL, configured[flags])) return e2;
  const int grid = p.n_items < sms ? p.n_items : sms;
  VF_CUDA(launch_pdl(kerns[flags], dim3(grid), dim3(ATT_THREADS), AttnSmem::TOTAL, static_cast<cudaStream_t>(stream), 1, p,
                     tmQ, tmKV, tmO));
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

int vf_attention_small_launch(const void* qkv, void* out, int B, int S, int H, int hd, float scale, cudaStream_t stream);

extern "C" int vf_attention_fwd_hd(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, int32_t head_dim, float scale,
                                   void* stream) {
  if (head_dim == 64) return vf_attention_fwd(qkv, out, B, S, H, scale, stream);
  VF_REQUIRE(qkv && out, VF_ERR_ARG, "vf_attention_fwd_hd: null pointer");
  VF_REQUIRE(B > 0 && S > 0 && H > 0 && head_dim > 0, VF_ERR_ARG, "vf_attention_fwd_hd: bad shape B=%d S=%d H=%d hd=%d", B, S, H, head_dim);
  VF_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, VF_ERR_ALIGN,
             "vf_attention_fwd_hd: pointers must be 16-byte aligned");
  return vf_attention_small_launch(qkv, out, B, S, H, head_dim, scale, static_cast<cudaStream_t>(stream));
}
