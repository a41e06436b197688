// vf_gemm.cu — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   out = epilogue(A[M,K] · W[N,K]^T)      A, W bf16 (K-major), fp32 accumulation in TMEM.
//
// Replaces every nn.Linear / nn.ConvNd(k=s) on the reference's vision path (see include/vfuse.h
// for the call-site list). Design:
//   * persistent: one CTA per SM — or one CTA PAIR (cta_group::2, a 256 x 256 tile, each CTA stages its own 128 A rows
//     and half of the W tile) for every 256-wide problem — static round-robin over the output tiles, n fastest so that
//     concurrently running CTAs share the same A rows in L2;
//   * warp 0 / lane 0: TMA producer, ring of {A 128x64, W (BN/CG)x64} bf16 stages (six for pairs, four otherwise),
//     128B swizzle;
//   * warp 1 / lane 0: tcgen05.mma issuer (M = 128 * CG, N = BN, K = 16 per instruction), accumulators double-buffered
//     in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1;
//   * warps 2..9: epilogue. tcgen05.ld 32x32b gives every thread one accumulator ROW. Three epilogue families:
//       - staged (RoPE, scatter, row remap, patch mode, small problems): 32x32 fp32 blocks are transposed through a
//         private XOR-swizzled smem block so that all global I/O is coalesced;
//       - bf16 rows + TMA store (bias / GELU, + folded LayerNorm): the math runs on the row, the 32x32 bf16 block is
//         written to a 64B-swizzled slot and stored by TMA — a fifth of the LSU traffic, no transposition;
//       - fp32 residual through TMA (short-K residual GEMMs): the residual block is TMA-loaded into a 128B-swizzled
//         slot, the accumulator row is added in place, the slot is TMA-stored (template flag RING, four stages);
//     plus the folded LayerNorm (producer: bf16 copy + partial row sums; consumer: rstd * (acc - mean * colsum) + b')
//     and the fused all-gather (peer / multicast stores);
//   * patch-embedding mode gathers the A tile with ONE 5-D TMA box per stage straight from the
//     [B,C,T,H,W] pixel tensor (im2col-free): tile rows are a 16 x 8 rectangle of patches. A patch
//     row is only 32 B wide, and TMA pads narrower-than-span rows under the 128 B swizzle (measured:
//     scratch/tma_probe.cu), so this operand uses the 32 B swizzle instead: the box lands as
//     [ph 16][py 4][pw 8][16 px], i.e. for every K=16 slice (one pixel row py) a dense K-major
//     SW32 tile whose 8-row groups (one patch row ph) are 1024 B apart.
#include "vf_common.cuh"
#include "../../include/vfuse.h"

#include <math.h>
#include <stdlib.h>
#include <atomic>
#include <type_traits>

namespace vf {

constexpr int BM = 128;
constexpr int BK = 64;
// warp0 TMA, warp1 MMA, then eight epilogue warps (two per TMEM lane quarter, each taking half of the
// tile's columns — the epilogue is latency-bound, so the extra warps are what hides it).
template <int EPI>
struct EpiCfg {
  static constexpr int WARPS = 8;
  static constexpr int THREADS = 64 + 32 * WARPS;
};

struct GemmParams {
  int M, N, K;
  int num_m_blk, num_n_blk, num_kb;
  const float* bias;
  void* out;
  long long ldo;
  const float* res;
  long long ldr;
  int grp_rows;
  long long grp_stride;
  long long row_off;
  const float* rope_cos;
  const float* rope_sin;
  int rope_period;
  int rope_cols;
  const int* dst_rows;
  int vec_ok;          // out (and res) rows are 16-byte aligned: 128-bit stores allowed
  int res_tma;         // residual epilogue through TMA (identity row map, N % 32 == 0): see the RING path
  int out_tma;         // bf16 epilogues (bias / GELU): math on accumulator rows, 32x32 bf16 blocks TMA-stored
  int n_peers;         // > 0: fused all-gather, every element goes to peer[0..n_peers) (NVLink peer memory) instead of out
  char* peer[8];
  // patch-embedding mode
  int patch;
  int PW, PH;          // tile rectangle in patches (PW*PH == 128)
  int nw, nh;          // patches per row / column of one frame
  int n_pwb, n_phb;    // tile rectangles per frame
  int Tp, T, C, tp, P; // frames after temporal merge, raw frames, channels, temporal patch, patch
  const float* pos;
  long long ld_pos;
  int peer_mc;         // peer[0] is an NVSwitch multicast (multimem) address: stores go through multimem.st
  // LayerNorm folded into the neighbouring GEMMs (see vf_epilogue in vfuse.h)
  __nv_bfloat16* ln_xb;      // producer: bf16 copy of the fp32 output rows, minus the row's shift
  long long ln_ldxb;
  float2* ln_stat_out;       // producer: [N/32][ln_stat_ld] partial (sum, sum of squares) of the shifted rows
  long long ln_stat_ld;
  const float* ln_shift;     // producer: [M] per-row shift (an estimate of the row mean) or nullptr
  const float2* ln_row_stats;  // consumer: [M] (mean of the shifted row, rstd)
  const float* ln_colsum;      // consumer: [N]
  // consumer, small problems: the producer's partial sums instead of finished statistics (no vf_ln_row_stats launch)
  const float2* ln_part_in;    // [K/32][ln_stat_ld]
  float* ln_shift_rw;          // [M] or nullptr: advanced to the rows' means by the first column tile of every row block
  float ln_eps;
  int ln_variant;
  long long* dbg;              // diagnosis (vf_gemm_set_debug): per CTA {issuer loop cycles, cycles waiting for operands,
                               //   cycles waiting for a free accumulator, tiles}
};

// CG = 1: one CTA per 128 x BN tile. CG = 2: a CTA pair (cta_group::2) per 256 x BN tile — each CTA stages
// its own 128 A rows and HALF of the W tile, so a k-block costs 32 KB of L2->SM traffic per SM instead of
// 48 KB (the 1-CTA kernel was capped by exactly that traffic, ~70 % tensor-pipe active) and six stages fit.
// RING (the fp32 residual epilogue of the CTA-pair kernel): four stages, and instead of the 32 KB of staging blocks
// every epilogue warp owns 12 KB: three 32x32 fp32 slots (TMA-loaded residual block, updated in place, TMA-stored), or
// two of them plus one 32x32 bf16 slot when the folded-LayerNorm copy of the rows is written as well.
template <int BN, int CG, bool RING = false>
struct SmemLayout {
  static constexpr int STAGES = RING ? 4 : (CG == 2 ? 6 : 4);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFF = BAR_OFF + 1024;   // TMA-swizzled epilogue slots: swizzle phase = address bits, keep them tile-relative
  static constexpr int RING_WARP_BYTES = 3 * 4096;   // three fp32 slots, or two + the bf16 slot (folded-LN producer)
  static constexpr int EPI_BYTES = RING ? 8 * RING_WARP_BYTES : 4 * 8192;   // else: 32x32 fp32 staging blocks
  static constexpr int TOTAL = EPI_OFF + EPI_BYTES + 1024;  // + alignment slack
};

struct EpiMaps {   // tensor maps of the TMA residual epilogue (32x32 boxes): residual in, fp32 out, bf16 copy out
  CUtensorMap res, out, xb;
};


__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
  // Bounded wait: a protocol bug becomes a trap (CUDA error) instead of a hung GPU.
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("vf_gemm: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(k0 (x + k1 x^3))) with the constants folded: 3 FMA-pipe ops, 1 MUFU, 2 more
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float u = x * fmaf(x * x, k01, k0);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  // x * Phi(x) with Phi from the Abramowitz-Stegun 7.1.26 form of erfc: q = 0.5 * poly(t) * exp(-x^2/2), t = 1/(1 + p|x|/sqrt2),
  // Phi = q for x < 0 and 1 - q otherwise. |error| < 5e-7 on GELU (three orders below the bf16 rounding of the output),
  // branch-free, 2 MUFU + 10 FMA-pipe ops; erff() costs ~30 instructions with a divergent branch and made the erf-GELU
  // epilogue issue-bound (777 vs 1288 TF for the tanh-GELU GEMM of the same shape).
  const float ax = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float q = 0.5f * poly * t * fast_exp2(-ax * ax * 1.4426950408889634f);
  return x * (x < 0.f ? q : 1.0f - q);
}

template <int EPI, int BN, bool PATCH, int CG, bool RING = false>
__global__ void __launch_bounds__(EpiCfg<EPI>::THREADS, 1)
gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmA,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ EpiMaps em) {
  static_assert(CG == 1 || (CG == 2 && BN == 256), "CTA pairs: 256-wide tiles");
  static_assert(!RING || (EPI == VF_EPI_BIAS_RES_F32 && !PATCH && CG == 2), "the TMA ring belongs to the CTA-pair residual kernel");
  using L = SmemLayout<BN, CG, RING>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  [[maybe_unused]] uint64_t* rfull_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF + 256);   // RING: [8 warps][3 slots]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tile walk: a "tile" is BM*CG rows x BN columns; CTA `rank` of a pair owns its rows [rank*BM, +BM)
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int num_tiles = ((p.num_m_blk + CG - 1) / CG) * p.num_n_blk;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EpiCfg<EPI>::WARPS * CG);  // one arrival per epilogue warp (of both CTAs)
    }
    if constexpr (RING) {
      tma_prefetch_desc(&em.res);
      for (int i = 0; i < 24; ++i) mbar_init(&rfull_bar[i], 1);
    }
    if (p.res_tma || p.out_tma) tma_prefetch_desc(&em.out);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair<2 * BN>(tmem_slot);
    else tmem_alloc<2 * BN>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // barrier inits + TMEM of both CTAs visible to the pair
  else __syncthreads();
  tc_fence_after();
  pdl_wait();                 // everything above overlapped the previous kernel's tail; its results are visible now
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int n_blk = tile % p.num_n_blk;
        const int m_blk = (tile / p.num_n_blk) * CG + rank;
        int pc2 = 0, pc3 = 0, pimg = 0;  // patch mode: pw0, ph0, (b*C)*T + t'*tp
        if (PATCH) {
          int t = m_blk;
          const int pwb = t % p.n_pwb; t /= p.n_pwb;
          const int phb = t % p.n_phb; t /= p.n_phb;
          const int tpr = t % p.Tp;
          const int b = t / p.Tp;
          pc2 = pwb * p.PW;
          pc3 = phb * p.PH;
          pimg = b * p.C * p.T + tpr * p.tp;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          wait_or_trap(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
          uint8_t* b_dst = a_dst + L::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES * CG);   // the pair's bytes land on the leader
          if (PATCH) {
            // K index = ((c*tp + dt)*P + py)*P + px ; one 64-wide k-block = (64/P) pixel rows.
            const int rows_per_kb = BK / p.P;
            const int kb_per_plane = p.P / rows_per_kb;       // k-blocks per (c,dt) plane
            const int plane = kb / kb_per_plane;              // c*tp + dt
            const int py0 = (kb % kb_per_plane) * rows_per_kb;
            const int c = plane / p.tp, dt = plane % p.tp;
            if (CG == 2) tma_load_5d_pair(a_dst, &tmA, &full_bar[stage], 0, pc2, py0, pc3, pimg + c * p.T + dt);
            else tma_load_5d(a_dst, &tmA, &full_bar[stage], 0, pc2, py0, pc3, pimg + c * p.T + dt);
          } else if (CG == 2) {
            tma_load_2d_pair(a_dst, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
            tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          }
          if (CG == 2) tma_load_2d_pair(b_dst, &tmB, &full_bar[stage], kb * BK, n_blk * BN + rank * (BN / 2));
          else tma_load_2d(b_dst, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long t_loop = 0, t_ops = 0, t_acc = 0, n_tiles = 0;
      if (p.dbg) t_loop = clock64();
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        if (p.dbg) { const long long c0 = clock64(); wait_or_trap(&tempty_bar[acc], acc_phase ^ 1); t_acc += clock64() - c0; ++n_tiles; }
        else wait_or_trap(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (p.dbg) { const long long c0 = clock64(); wait_or_trap(&full_bar[stage], phase); t_ops += clock64() - c0; }
          else wait_or_trap(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(a_addr);
          const uint64_t b_desc = umma_desc_sw128(a_addr + L::A_BYTES);
          if (PATCH) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {  // slice k = pixel row py0+k: [16 ph][8 pw][32 B], SBO 1024 B
              if (CG == 2) umma_ss_pair(d_tmem, umma_desc_sw32(a_addr + k * 256, 1024), b_desc + 2 * k, idesc, (kb | k) != 0);
              else umma_ss(d_tmem, umma_desc_sw32(a_addr + k * 256, 1024), b_desc + 2 * k, idesc, (kb | k) != 0);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in the (addr>>4) field
              if (CG == 2) umma_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              else umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
          }
          // smem slot is free (in both CTAs of a pair) once these MMAs retire
          if (CG == 2) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (each CTA of a pair drains its own 128 rows)
        if (CG == 2) umma_commit_pair(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (p.dbg) {
        long long* d = p.dbg + 4ll * blockIdx.x;
        d[0] = clock64() - t_loop; d[1] = t_ops; d[2] = t_acc; d[3] = n_tiles;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // tcgen05.ld hands every thread one accumulator ROW; global memory wants warps to touch whole
    // 128-byte lines. Each warp therefore transposes 32x32 fp32 blocks through a private, XOR-swizzled
    // (conflict-free) shared-memory buffer and does all epilogue math + I/O in the coalesced mapping:
    // lane l owns columns 4*(l%8)..+3 of rows (l/8)+4i, i = 0..7.
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    // staging blocks are addressed in the shared window explicitly (ld/st.shared), never through
    // generic pointers
    constexpr int EW = EpiCfg<EPI>::WARPS;
    constexpr int COLS_PER_WARP = BN / (EW / 4);      // columns of the tile this warp drains
    const int chalf = (warp - 2) >> 2;                // 0, or 0/1 with eight warps
    const uint32_t stA = smem_u32(smem + L::EPI_OFF + (warp - 2) * (32768 / EW));
    const int cl = lane & 7;       // 16-byte column group inside a 32-column block
    const int rl = lane >> 3;      // row offset inside a group of 4 rows
    int acc = 0;
    uint32_t acc_phase = 0;

    // write this thread's accumulator row (32 fp32) into the staging block, chunk j -> slot j ^ (row & 7)
    auto stage_row = [&](uint32_t st, const uint32_t* r) {
      const uint32_t row = st + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((j ^ (lane & 7)) << 4)), "r"(r[4 * j]),
                     "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                     : "memory");
    };
    auto read_staged = [&](uint32_t st, int rr) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(st + rr * 128 + ((cl ^ (rr & 7)) << 4))
                   : "memory");
      return v;
    };

    // 8 bytes of the slot read_staged(rr) returns: the upper half for rows 8..15 and 24..31, so that the per-row read-back
    // (lane = row) meets 2-way instead of 4-way bank conflicts
    auto write_staged2 = [&](uint32_t st, int rr, float a, float b) {
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(st + rr * 128 + ((cl ^ (rr & 7)) << 4) + (rr & 8)), "f"(a),
                   "f"(b)
                   : "memory");
    };

    // Folded LayerNorm, consumer side: lane L fetches (mean', rstd) of tile row L. Normally they are final when the launch
    // starts (vf_ln_row_stats between the producing and the consuming GEMM). For SMALL problems (ln_part_in) the launch
    // in between is dropped: every epilogue warp adds up the K/32 partial sums of its 32 rows itself — in exactly
    // vf_ln_row_stats' order (eight interleaved groups, then the groups in order), so a row's statistics do not depend on
    // which of the two routes its batch size selects — and the first column tile of a row block advances the row shift.
    [[maybe_unused]] auto ln_fetch = [&](int row_first, int n_blk, float& mu, float& rs) {
      mu = 0.f; rs = 0.f;
      const int row = row_first + lane;
      if (row >= p.M) return;
      if (p.ln_part_in == nullptr) {
        const float2 t = __ldg(p.ln_row_stats + row);
        mu = t.x; rs = t.y;
        return;
      }
      const int parts = p.K >> 5;
      float gs[8], gq[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
      for (int j0 = 0; j0 < parts; j0 += 8) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (j0 + g < parts) {
            const float2 t = __ldcg(p.ln_part_in + (long long)(j0 + g) * p.ln_stat_ld + row);
            gs[g] += t.x; gq[g] += t.y;
          }
        }
      }
      float s_ = gs[0], q_ = gq[0];
#pragma unroll
      for (int g = 1; g < 8; ++g) { s_ += gs[g]; q_ += gq[g]; }
      const float inv_d = 1.0f / static_cast<float>(p.K);
      mu = s_ * inv_d;
      const float var = fmaxf(fmaf(-mu, mu, q_ * inv_d), 0.f);
      rs = p.ln_variant == 0 ? rsqrtf(var + p.ln_eps) : 1.0f / (sqrtf(var) + p.ln_eps);
      if (p.ln_shift_rw != nullptr && n_blk == 0 && chalf == 0) p.ln_shift_rw[row] += mu;
    };

    // Residual epilogue: the fp32 residual slab of a tile (128 KB) is pulled into L2 one tile ahead, while
    // the tensor pipe is still busy with the current one, so the epilogue's loads are L2 hits instead of
    // serialised DRAM round trips (measured: the proj GEMM, K=768, was bound by exactly that latency).
    auto prefetch_res = [&](int tile) {
      if constexpr (EPI == VF_EPI_BIAS_RES_F32 && !PATCH) {
        if (tile < num_tiles && p.grp_rows == 0) {
          const int m = ((tile / p.num_n_blk) * CG + rank) * BM + quarter * 32 + lane;
          const int c0 = (tile % p.num_n_blk) * BN + chalf * COLS_PER_WARP;
          if (m < p.M) {
            const char* rp = reinterpret_cast<const char*>(p.res + (long long)m * p.ldr + c0);
#pragma unroll
            for (int i = 0; i < COLS_PER_WARP * 4 / 128; ++i)
              if (c0 + i * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + i * 128));
          }
        }
      }
    };
    prefetch_res(tile0);

    // ------------------------------------------------------------------ residual epilogue through TMA (RING)
    // The staged epilogue moves every fp32 element through the SM's load/store unit four times (stage, read back,
    // residual load, store): 512 KB per 128x256 tile, more LSU time than the K=768 main loop takes — the proj GEMM
    // was bound by exactly that, not by DRAM (a residual that always hits L1 cost the same +25 us). Here TMA brings
    // the 32x32 residual block into a swizzled slot, every thread adds its accumulator ROW in place (thread = row, so
    // the folded-LayerNorm row sums are thread-local: no shuffles, no second pass) and TMA stores the slot: 256 KB of
    // LSU traffic per tile, no transposition, no global load/store instructions at all.
    bool ring_done = false;
    if constexpr (RING) {
      if (p.res_tma) {
        ring_done = true;
        const int ew = warp - 2;
        uint8_t* ring = smem + L::EPI_OFF + ew * L::RING_WARP_BYTES;
        const uint32_t ring_u = smem_u32(ring);
        uint64_t* rf = rfull_bar + ew * 3;
        const bool ln_out = p.ln_xb != nullptr;
        const int R = ln_out ? 2 : 3;                        // fp32 slots (the third one is the bf16 slot with ln_out)
        // load cursor: block (nl_tile, nl_c) goes into slot nl_slot next (lane 0)
        int nl_tile = tile0, nl_c = 0, nl_slot = 0;
        auto issue_load = [&]() {
          if (nl_tile >= num_tiles) return;
          const int col = (nl_tile % p.num_n_blk) * BN + chalf * COLS_PER_WARP + nl_c * 32;
          const int row = ((nl_tile / p.num_n_blk) * CG + rank) * BM + quarter * 32;
          mbar_expect_tx(&rf[nl_slot], 4096);
          tma_load_2d(ring + nl_slot * 4096, &em.res, &rf[nl_slot], col, row);
          if (++nl_c == COLS_PER_WARP / 32) { nl_c = 0; nl_tile += tile_step; }
          if (++nl_slot == R) nl_slot = 0;
        };
        int slot = 0;
        uint32_t ph = 0;   // parity of rf[slot] for the block being processed
        if (lane == 0)
          for (int i = 0; i < R - 1; ++i) issue_load();      // R-1 blocks in flight ahead of the one being processed
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
          const int col0 = (tile % p.num_n_blk) * BN + chalf * COLS_PER_WARP;
          const int row0 = ((tile / p.num_n_blk) * CG + rank) * BM + quarter * 32;
          prefetch_res(tile + tile_step);
          // folded-LayerNorm producer: the bf16 copy and the partial sums are taken of x - shift[row] (thread = row)
          const float sh = (ln_out && p.ln_shift && row0 + lane < p.M) ? __ldg(p.ln_shift + row0 + lane) : 0.f;
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + chalf * COLS_PER_WARP;
#pragma unroll 1
          for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
            const int col = col0 + c * 32;
            if (lane == 0) {
              // every earlier store has left shared memory: the slot of the previous block (and the bf16 slot) may be
              // rewritten, so the load that runs R-1 blocks ahead goes into it
              tma_store_wait_read<0>();
              issue_load();
            }
            __syncwarp();
            if (c == 0) {
              wait_or_trap(&tfull_bar[acc], acc_phase);
              tc_fence_after();
            }
            uint32_t r[32];
            tmem_ld_x32(t_row + c * 32, r);
            // the 32 bias values of the block (uniform addresses: one broadcast wavefront per load), requested before
            // the waits; loads, adds and stores below are batched so that nothing waits on a single round trip
            float4 bv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              bv[j] = (p.bias && col < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col + 4 * j))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);   // N % 32 == 0: a block is all in or all out
            wait_or_trap(&rf[slot], ph);
            const uint32_t rowb = ring_u + slot * 4096 + lane * 128;
            const uint32_t xrow = ring_u + 8192 + lane * 64;
            float4 x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(x[j].x), "=f"(x[j].y), "=f"(x[j].z), "=f"(x[j].w)
                           : "r"(rowb + ((j ^ (lane & 7)) << 4))
                           : "memory");
            tmem_ld_wait();
            if (c == COLS_PER_WARP / 32 - 1) {   // all tcgen05.ld of this warp for the tile are complete
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
            }
            float s_ = 0.f, q_ = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              x[j].x += __uint_as_float(r[4 * j]) + bv[j].x;
              x[j].y += __uint_as_float(r[4 * j + 1]) + bv[j].y;
              x[j].z += __uint_as_float(r[4 * j + 2]) + bv[j].z;
              x[j].w += __uint_as_float(r[4 * j + 3]) + bv[j].w;
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((j ^ (lane & 7)) << 4)), "f"(x[j].x),
                           "f"(x[j].y), "f"(x[j].z), "f"(x[j].w)
                           : "memory");
            }
            if (ln_out) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                // same association as the staged epilogue (4-column shares added in column order): a row's statistics
                // must not depend on which of the two paths its batch size selects
                x[j].x -= sh; x[j].y -= sh; x[j].z -= sh; x[j].w -= sh;
                s_ += (x[j].x + x[j].y) + (x[j].z + x[j].w);
                q_ += fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, fmaf(x[j].z, x[j].z, x[j].w * x[j].w)));
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)   // 64-byte rows, 64 B swizzle: chunk ^= (row >> 1) & 3
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xrow + ((k ^ ((lane >> 1) & 3)) << 4)),
                             "r"(pack_bf16(x[2 * k].x, x[2 * k].y)), "r"(pack_bf16(x[2 * k].z, x[2 * k].w)),
                             "r"(pack_bf16(x[2 * k + 1].x, x[2 * k + 1].y)), "r"(pack_bf16(x[2 * k + 1].z, x[2 * k + 1].w))
                             : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&em.out, ring + slot * 4096, col, row0);
              if (ln_out) tma_store_2d(&em.xb, ring + 8192, col, row0);
              tma_store_commit();
            }
            if (ln_out && col < p.N && row0 + lane < p.M)   // (a block right of N is zero-filled by TMA and clipped on store)
              p.ln_stat_out[(long long)(col >> 5) * p.ln_stat_ld + row0 + lane] = make_float2(s_, q_);
            if (++slot == R) { slot = 0; ph ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait<0>();
      }
    }

    // ------------------------------------------------------------------ bf16 epilogues through a TMA store
    // Same idea for bias / GELU (+ folded LayerNorm) outputs: the math runs on the accumulator ROW a thread gets from
    // tcgen05.ld (bias and colsum are uniform broadcast loads, mean / rstd of the row are two scalars of the thread),
    // the 32x32 bf16 block goes into a 64B-swizzled 2 KB slot with four 16-byte stores per thread and TMA writes it
    // out: 64 KB of LSU traffic per 128x256 tile instead of 320 KB (stage fp32, read back, store), no transposition.
    if constexpr (EPI == VF_EPI_BIAS_BF16 || EPI == VF_EPI_GELU_TANH_BF16 || EPI == VF_EPI_GELU_ERF_BF16) {
      if (p.out_tma) {
        ring_done = true;
        const uint32_t slots = smem_u32(smem + L::EPI_OFF + (warp - 2) * 4096);   // two 2 KB slots per warp
        const bool ln_fold = p.ln_row_stats != nullptr || p.ln_part_in != nullptr;
        int sl = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
          const int col0 = (tile % p.num_n_blk) * BN + chalf * COLS_PER_WARP;
          const int row0 = ((tile / p.num_n_blk) * CG + rank) * BM + quarter * 32;
          float mu = 0.f, rs = 1.f;
          if (ln_fold) ln_fetch(row0, tile % p.num_n_blk, mu, rs);
          wait_or_trap(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + chalf * COLS_PER_WARP;
#pragma unroll 1
          for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
            const int col = col0 + c * 32;
            const bool live = col < p.N;          // N % 32 == 0: a block is entirely inside or outside (uniform)
            uint32_t r[32];
            tmem_ld_x32(t_row + c * 32, r);
            float4 bv[8], cs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              bv[j] = (live && p.bias) ? __ldg(reinterpret_cast<const float4*>(p.bias + col + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
              cs[j] = (live && ln_fold) ? __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (lane == 0) tma_store_wait_read<1>();   // the store that used this slot two blocks ago has left shared memory
            __syncwarp();
            tmem_ld_wait();
            if (c == COLS_PER_WARP / 32 - 1) {         // all tcgen05.ld of this warp for the tile are complete
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);
                else mbar_arrive(&tempty_bar[acc]);
              }
            }
            if (live) {
              const uint32_t srow = slots + sl * 2048 + lane * 64;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint32_t w[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int j = 2 * k + h;
                  float v[4] = {__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3])};
                  const float bb[4] = {bv[j].x, bv[j].y, bv[j].z, bv[j].w};
                  const float cc4[4] = {cs[j].x, cs[j].y, cs[j].z, cs[j].w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    // identical expressions to the staged epilogue: both paths must give the same bits
                    v[e] = ln_fold ? fmaf(rs, fmaf(-mu, cc4[e], v[e]), bb[e]) : v[e] + bb[e];
                    if constexpr (EPI == VF_EPI_GELU_TANH_BF16) v[e] = gelu_tanh_f(v[e]);
                    if constexpr (EPI == VF_EPI_GELU_ERF_BF16) v[e] = gelu_erf_f(v[e]);
                  }
                  w[2 * h] = pack_bf16(v[0], v[1]);
                  w[2 * h + 1] = pack_bf16(v[2], v[3]);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((k ^ ((lane >> 1) & 3)) << 4)), "r"(w[0]),
                             "r"(w[1]), "r"(w[2]), "r"(w[3])
                             : "memory");
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&em.out, reinterpret_cast<const void*>(smem + L::EPI_OFF + (warp - 2) * 4096 + sl * 2048), col, row0);
                tma_store_commit();
              }
              sl ^= 1;
            }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait<0>();
      }
    }

    for (int tile = ring_done ? num_tiles : tile0; tile < num_tiles; tile += tile_step) {
      const int n_blk = tile % p.num_n_blk;
      const int m_blk = (tile / p.num_n_blk) * CG + rank;
      const int col0 = n_blk * BN;
      prefetch_res(tile + tile_step);

      // ---- output row (or -1) of the 8 rows this lane touches: rows r0, r0+4, ..., r0+28 of the tile.
      // One integer division per tile at most; the other rows follow incrementally.
      long long orow[8];
      int aux[8];  // patch: spatial index (pos-embed row); rope: cos/sin table row
      [[maybe_unused]] long long lane_orow = -1;   // output row of tile row quarter*32 + lane (folded-LN producer)
      {
        const int r0 = quarter * 32 + rl;
        if constexpr (PATCH) {
          int t = m_blk;
          const int pwb = t % p.n_pwb; t /= p.n_pwb;
          const int phb = t % p.n_phb; t /= p.n_phb;
          const int tpr = t % p.Tp;
          const int b = t / p.Tp;
          const long long base = (long long)b * p.grp_stride + p.row_off + (long long)tpr * p.nh * p.nw;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r_local = r0 + 4 * i;          // tile rows are a 16 x 8 rectangle of patches
            const int ph = phb * 16 + (r_local >> 3);
            const int pw = pwb * 8 + (r_local & 7);
            const bool ok = ph < p.nh && pw < p.nw && m_blk < p.num_m_blk;   // (pair mode: the last pair may lack its second row block)
            aux[i] = ok ? ph * p.nw + pw : 0;
            orow[i] = ok ? base + aux[i] : -1;
          }
        } else {
          const int m0 = m_blk * BM + r0;
          if (m_blk * BM + quarter * 32 + lane < p.M) lane_orow = m_blk * BM + quarter * 32 + lane;   // identity map
          int q = 0, rem = 0;                         // remap / rope: running quotient and remainder
          if (EPI == VF_EPI_QKV_ROPE_BF16) rem = m0 % p.rope_period;
          else if (p.grp_rows > 0) { q = m0 / p.grp_rows; rem = m0 - q * p.grp_rows; }
          const int period = EPI == VF_EPI_QKV_ROPE_BF16 ? p.rope_period : p.grp_rows;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m0 + 4 * i;
            long long o = m;
            if (EPI == VF_EPI_SCATTER_BF16) o = m < p.M ? p.dst_rows[m] : -1;
            else if (EPI != VF_EPI_QKV_ROPE_BF16 && p.grp_rows > 0) o = (long long)q * p.grp_stride + rem + p.row_off;
            orow[i] = m < p.M ? o : -1;
            aux[i] = rem;
            if (period > 0) {
              rem += 4;
              while (rem >= period) { rem -= period; ++q; }
            }
          }
        }
      }

      // Residual epilogue: the residual values of a 32-column block are requested one block ahead of their use
      // (block 0 before the accumulator is even complete), into the registers the previous block just freed.
      [[maybe_unused]] float4 ex[8];
      [[maybe_unused]] const bool fast_res = p.vec_ok && ((p.N & 3) == 0);
      [[maybe_unused]] auto load_res = [&](int c) {
        if constexpr (EPI == VF_EPI_BIAS_RES_F32 && !PATCH) {
          const int cc = col0 + chalf * COLS_PER_WARP + c * 32 + cl * 4;
          if (fast_res && cc < p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const long long o = orow[i] < 0 ? 0 : orow[i];
              ex[i] = *reinterpret_cast<const float4*>(p.res + o * p.ldr + cc);
            }
          }
        }
      };
      load_res(0);
      [[maybe_unused]] float shv[8];
      if constexpr (EPI == VF_EPI_BIAS_RES_F32 && !PATCH) {
#pragma unroll
        for (int i = 0; i < 8; ++i) shv[i] = (p.ln_xb && p.ln_shift && orow[i] >= 0) ? __ldg(p.ln_shift + orow[i]) : 0.f;
      }

      // LayerNorm folded into this GEMM (consumer side): lane L fetches (mean, rstd) of tile row quarter*32 + L while the
      // tensor pipe is still busy with the tile; the coalesced mapping below gets the values of its rows by shuffle.
      [[maybe_unused]] float ln_mu = 0.f, ln_rs = 0.f;
      [[maybe_unused]] const bool ln_in = p.ln_row_stats != nullptr || p.ln_part_in != nullptr;
      if constexpr (EPI == VF_EPI_BIAS_BF16 || EPI == VF_EPI_GELU_TANH_BF16 || EPI == VF_EPI_GELU_ERF_BF16 || EPI == VF_EPI_QKV_ROPE_BF16) {
        if (ln_in) ln_fetch(m_blk * BM + quarter * 32, n_blk, ln_mu, ln_rs);
      }

      wait_or_trap(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + chalf * COLS_PER_WARP;

      if constexpr (EPI == VF_EPI_QKV_ROPE_BF16) {
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
        // cos/sin depend on (row, pair column) only: fetched once per tile, reused by every head
        float4 cv[8], sv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          cv[i] = __ldg(reinterpret_cast<const float4*>(p.rope_cos + (long long)aux[i] * 32) + cl);
          sv[i] = __ldg(reinterpret_cast<const float4*>(p.rope_sin + (long long)aux[i] * 32) + cl);
        }
#pragma unroll 1
        for (int hh = 0; hh < COLS_PER_WARP / 64; ++hh) {
          const int hc = col0 + chalf * COLS_PER_WARP + hh * 64;   // first column of this head
          // The rotation pairs column i with i+32 of the same row. Both halves go through ONE staging
          // block: the first half is parked in registers (coalesced mapping) while the second is staged.
          uint32_t r[32];
          tmem_ld_x32(t_row + hh * 64, r);
          tmem_ld_wait();
          __syncwarp();                          // previous head fully consumed by all lanes
          stage_row(stA, r);
          tmem_ld_x32(t_row + hh * 64 + 32, r);  // second half in flight while the first is read back
          __syncwarp();
          float4 x1s[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x1s[i] = read_staged(stA, rl + 4 * i);
          tmem_ld_wait();
          __syncwarp();
          stage_row(stA, r);
          __syncwarp();
          if (hc >= p.N) continue;
          const bool rot = hc < p.rope_cols;
          float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f), b2 = b1, cs1 = b1, cs2 = b1;
          if (p.bias) {
            b1 = __ldg(reinterpret_cast<const float4*>(p.bias + hc) + cl);
            b2 = __ldg(reinterpret_cast<const float4*>(p.bias + hc + 32) + cl);
          }
          if (ln_in) {
            cs1 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + hc) + cl);
            cs2 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + hc + 32) + cl);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = rl + 4 * i;
            float4 x1 = x1s[i], x2 = read_staged(stA, rr);
            if (ln_in) {   // rstd * (acc - mean * colsum) + bias'
              const float mu = __shfl_sync(0xffffffffu, ln_mu, rr), rs = __shfl_sync(0xffffffffu, ln_rs, rr);
              x1.x = fmaf(rs, fmaf(-mu, cs1.x, x1.x), b1.x); x1.y = fmaf(rs, fmaf(-mu, cs1.y, x1.y), b1.y);
              x1.z = fmaf(rs, fmaf(-mu, cs1.z, x1.z), b1.z); x1.w = fmaf(rs, fmaf(-mu, cs1.w, x1.w), b1.w);
              x2.x = fmaf(rs, fmaf(-mu, cs2.x, x2.x), b2.x); x2.y = fmaf(rs, fmaf(-mu, cs2.y, x2.y), b2.y);
              x2.z = fmaf(rs, fmaf(-mu, cs2.z, x2.z), b2.z); x2.w = fmaf(rs, fmaf(-mu, cs2.w, x2.w), b2.w);
            } else {
              x1.x += b1.x; x1.y += b1.y; x1.z += b1.z; x1.w += b1.w;
              x2.x += b2.x; x2.y += b2.y; x2.z += b2.z; x2.w += b2.w;
            }
            const float4 c = rot ? cv[i] : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 sn = rot ? sv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 y1, y2;
            y1.x = x1.x * c.x - x2.x * sn.x; y2.x = x2.x * c.x + x1.x * sn.x;
            y1.y = x1.y * c.y - x2.y * sn.y; y2.y = x2.y * c.y + x1.y * sn.y;
            y1.z = x1.z * c.z - x2.z * sn.z; y2.z = x2.z * c.z + x1.z * sn.z;
            y1.w = x1.w * c.w - x2.w * sn.w; y2.w = x2.w * c.w + x1.w * sn.w;
            if (orow[i] >= 0) {
              __nv_bfloat16* o = outp + orow[i] * p.ldo + hc + cl * 4;
              *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(y1.x, y1.y), pack_bf16(y1.z, y1.w));
              *reinterpret_cast<uint2*>(o + 32) = make_uint2(pack_bf16(y2.x, y2.y), pack_bf16(y2.z, y2.w));
            }
          }
        }
      } else {
        constexpr bool OUT_F32 = (EPI == VF_EPI_BIAS_F32 || EPI == VF_EPI_BIAS_RES_F32);
        constexpr int ESZ = OUT_F32 ? 4 : 2;
        // Uniform fast path: 16-byte aligned rows and N % 4 == 0, so a lane's 4 columns are either all
        // inside the matrix or all outside and every access is one vector instruction.
        const bool fast = p.vec_ok && ((p.N & 3) == 0);
        const bool rows_all_valid = !PATCH && EPI != VF_EPI_SCATTER_BF16 && (m_blk + 1) * BM <= p.M;
        // per-row base pointers, once per tile (rows that are masked out point at row 0 and never store)
        char* obase[8];
        const char* rbase[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long o = orow[i] < 0 ? 0 : orow[i];
          obase[i] = reinterpret_cast<char*>(p.out) + o * p.ldo * ESZ;
          rbase[i] = EPI == VF_EPI_BIAS_RES_F32 ? reinterpret_cast<const char*>(p.res) + o * p.ldr * 4 : nullptr;
        }
        uint32_t r[32];
        tmem_ld_x32(t_row, r);
#pragma unroll 1
        for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
          tmem_ld_wait();
          __syncwarp();                          // previous block fully consumed by all lanes
          stage_row(stA, r);
          if (c + 1 < COLS_PER_WARP / 32) tmem_ld_x32(t_row + (c + 1) * 32, r);  // overlaps the math below
          __syncwarp();
          const int cc = col0 + chalf * COLS_PER_WARP + c * 32 + cl * 4;   // first of this lane's 4 columns
          if (cc >= p.N) continue;
          if (fast) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + cc));
            [[maybe_unused]] float4 csv = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (EPI == VF_EPI_BIAS_BF16 || EPI == VF_EPI_GELU_TANH_BF16 || EPI == VF_EPI_GELU_ERF_BF16)
              if (ln_in) csv = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + cc));
            [[maybe_unused]] const bool ln_out = p.ln_xb != nullptr;
            // residual / pos-embed values of all 8 rows are fetched up front: they may alias `out`, so
            // the compiler cannot hoist them across the stores by itself
            if constexpr (PATCH) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                ex[i] = p.pos ? __ldg(reinterpret_cast<const float4*>(p.pos + (long long)aux[i] * p.ld_pos + cc))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // rows_all_valid (uniform): no per-row branch at all, so the 8 rows interleave freely
            auto do_rows = [&](auto guarded) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 x = read_staged(stA, rl + 4 * i);
                float v[4] = {x.x + bv.x, x.y + bv.y, x.z + bv.z, x.w + bv.w};
                if constexpr (EPI == VF_EPI_BIAS_BF16 || EPI == VF_EPI_GELU_TANH_BF16 || EPI == VF_EPI_GELU_ERF_BF16) {
                  if (ln_in) {   // rstd * (acc - mean * colsum) + bias'
                    const float mu = __shfl_sync(0xffffffffu, ln_mu, rl + 4 * i);
                    const float rs = __shfl_sync(0xffffffffu, ln_rs, rl + 4 * i);
                    v[0] = fmaf(rs, fmaf(-mu, csv.x, x.x), bv.x); v[1] = fmaf(rs, fmaf(-mu, csv.y, x.y), bv.y);
                    v[2] = fmaf(rs, fmaf(-mu, csv.z, x.z), bv.z); v[3] = fmaf(rs, fmaf(-mu, csv.w, x.w), bv.w);
                  }
                }
                if constexpr (EPI == VF_EPI_BIAS_RES_F32 || PATCH) {
                  v[0] += ex[i].x; v[1] += ex[i].y; v[2] += ex[i].z; v[3] += ex[i].w;
                }
                if constexpr (EPI == VF_EPI_BIAS_RES_F32 && !PATCH) {
                  if (ln_out) {
                    // producer side of the folded LayerNorm: bf16 copy of the row minus its shift (the next GEMM's A
                    // operand), and this lane's share of the shifted row's (sum, sum of squares) parked in the staging slot
                    // it has just read
                    const float u[4] = {v[0] - shv[i], v[1] - shv[i], v[2] - shv[i], v[3] - shv[i]};
                    write_staged2(stA, rl + 4 * i, (u[0] + u[1]) + (u[2] + u[3]),
                                  fmaf(u[0], u[0], fmaf(u[1], u[1], fmaf(u[2], u[2], u[3] * u[3]))));
                    if (!decltype(guarded)::value || orow[i] >= 0)
                      *reinterpret_cast<uint2*>(p.ln_xb + orow[i] * p.ln_ldxb + cc) =
                          make_uint2(pack_bf16(u[0], u[1]), pack_bf16(u[2], u[3]));
                  }
                }
                if constexpr (EPI == VF_EPI_GELU_TANH_BF16) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[e] = gelu_tanh_f(v[e]);
                }
                if constexpr (EPI == VF_EPI_GELU_ERF_BF16) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[e] = gelu_erf_f(v[e]);
                }
                if (!decltype(guarded)::value || orow[i] >= 0) {
                  if constexpr (EPI == VF_EPI_BIAS_BF16 || EPI == VF_EPI_BIAS_F32) {
                    if (p.n_peers > 0) {   // fused all-gather: the same element to every GPU's copy of the gathered buffer
                      const long long off = (obase[i] - reinterpret_cast<char*>(p.out)) + (long long)cc * ESZ;
                      if (p.peer_mc) {
                        // NVSwitch multicast mapping: only multimem.* may touch it; the switch replicates the store
                        // into every GPU's copy of the gathered buffer
                        if constexpr (OUT_F32)
                          asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.peer[0] + off), "f"(v[0]),
                                       "f"(v[1]), "f"(v[2]), "f"(v[3])
                                       : "memory");
                        else
                          asm volatile("multimem.st.weak.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(p.peer[0] + off),
                                       "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3]))
                                       : "memory");
                        continue;
                      }
                      for (int r = 0; r < p.n_peers; ++r) {
                        if constexpr (OUT_F32)
                          *reinterpret_cast<float4*>(p.peer[r] + off) = make_float4(v[0], v[1], v[2], v[3]);
                        else
                          *reinterpret_cast<uint2*>(p.peer[r] + off) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
                      }
                      continue;
                    }
                  }
                  if constexpr (OUT_F32)
                    *reinterpret_cast<float4*>(obase[i] + cc * 4) = make_float4(v[0], v[1], v[2], v[3]);
                  else
                    *reinterpret_cast<uint2*>(obase[i] + cc * 2) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
                }
              }
            };
            if (rows_all_valid) do_rows(std::false_type{});
            else do_rows(std::true_type{});
            if constexpr (EPI == VF_EPI_BIAS_RES_F32 && !PATCH) {
              if (ln_out) {
                // lane L adds up the 8 shares of tile row quarter*32 + L (no shuffles, no extra live registers) and the
                // warp stores the 32 partial sums of this 32-column block with one coalesced 256-byte store
                __syncwarp();
                float s_ = 0.f, q_ = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  float a_, b_;
                  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                               : "=f"(a_), "=f"(b_)
                               : "r"(stA + lane * 128 + ((k ^ (lane & 7)) << 4) + (lane & 8))
                               : "memory");
                  s_ += a_; q_ += b_;
                }
                if (lane_orow >= 0) p.ln_stat_out[(long long)(cc >> 5) * p.ln_stat_ld + lane_orow] = make_float2(s_, q_);
              }
            }
            if (c + 1 < COLS_PER_WARP / 32) load_res(c + 1);
          } else {
            // generic path (unaligned rows or N % 4 != 0, e.g. a 10-class head): scalar, per-element guards
#pragma unroll
            for (int i = 0; i < 8; ++i) {   // unrolled so that the per-row arrays stay in registers
              if (orow[i] < 0) continue;
              const float4 x = read_staged(stA, rl + 4 * i);
              const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (cc + e >= p.N) continue;
                float v = xv[e] + (p.bias ? p.bias[cc + e] : 0.f);
                if constexpr (EPI == VF_EPI_BIAS_RES_F32) v += reinterpret_cast<const float*>(rbase[i])[cc + e];
                if constexpr (PATCH) { if (p.pos) v += p.pos[(long long)aux[i] * p.ld_pos + cc + e]; }
                if constexpr (EPI == VF_EPI_GELU_TANH_BF16) v = gelu_tanh_f(v);
                if constexpr (EPI == VF_EPI_GELU_ERF_BF16) v = gelu_erf_f(v);
                if constexpr (OUT_F32) reinterpret_cast<float*>(obase[i])[cc + e] = v;
                else reinterpret_cast<__nv_bfloat16*>(obase[i])[cc + e] = __float2bfloat16_rn(v);
              }
            }
          }
        }
      }
      // release the accumulator buffer (all tcgen05.ld of this warp have completed)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);   // the leader's MMA warp reuses the buffer pair-wide
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the peer may still be reading this CTA's smem / TMEM
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair<2 * BN>(tmem_base);
    else tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int EPI, int BN, bool PATCH = false, int CG = 1, bool RING = false>
static int launch_gemm(const GemmParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB,
                       cudaStream_t stream, const EpiMaps* em_in = nullptr) {
  using L = SmemLayout<BN, CG, RING>;
  static const EpiMaps no_maps{};
  const EpiMaps& em = em_in ? *em_in : no_maps;
  auto kfn = gemm_kernel<EPI, BN, PATCH, CG, RING>;
  static std::atomic<uint64_t> configured{0};
  if (int e = ensure_dynamic_smem(kfn, L::TOTAL, configured)) return e;
  const int sms = device_sm_count();
  VF_REQUIRE(sms > 0, VF_ERR_NO_DEVICE, "no CUDA device");
  const int tiles = ((p.num_m_blk + CG - 1) / CG) * p.num_n_blk;
  const int slots = sms / CG;                       // persistent: one CTA (pair) per SM (pair)
  const int grid = (tiles < slots ? tiles : slots) * CG;
  VF_CUDA(launch_pdl(kfn, dim3(grid), dim3(EpiCfg<EPI>::THREADS), L::TOTAL, stream, CG, p, tmA, tmB, em));
  count_launch();
  VF_CUDA(cudaGetLastError());
  return VF_OK;
}

template <int BN, int CG>
static int dispatch_epi(int mode, const GemmParams& p, const CUtensorMap& a, const CUtensorMap& b,
                        cudaStream_t s, const EpiMaps* em = nullptr) {
  switch (mode) {
    case VF_EPI_BIAS_BF16: return launch_gemm<VF_EPI_BIAS_BF16, BN, false, CG>(p, a, b, s, em);
    case VF_EPI_BIAS_F32: return launch_gemm<VF_EPI_BIAS_F32, BN, false, CG>(p, a, b, s);
    case VF_EPI_BIAS_RES_F32:
      if constexpr (CG == 2) {
        if (p.res_tma) return launch_gemm<VF_EPI_BIAS_RES_F32, BN, false, CG, true>(p, a, b, s, em);
      }
      return launch_gemm<VF_EPI_BIAS_RES_F32, BN, false, CG>(p, a, b, s, em);
    case VF_EPI_GELU_TANH_BF16: return launch_gemm<VF_EPI_GELU_TANH_BF16, BN, false, CG>(p, a, b, s, em);
    case VF_EPI_GELU_ERF_BF16: return launch_gemm<VF_EPI_GELU_ERF_BF16, BN, false, CG>(p, a, b, s, em);
    case VF_EPI_QKV_ROPE_BF16: return launch_gemm<VF_EPI_QKV_ROPE_BF16, BN, false, CG>(p, a, b, s);
    case VF_EPI_SCATTER_BF16: return launch_gemm<VF_EPI_SCATTER_BF16, BN, false, CG>(p, a, b, s);
    default: break;
  }
  set_last_error("vf_gemm_bf16: unknown epilogue mode %d", mode);
  return VF_ERR_ARG;
}

static int pick_bn(int M, int N) {
  // 256-wide tiles halve the number of A re-reads and MMA issue overhead; use 128 when N is small
  // or when 256-wide tiles would leave most SMs idle.
  if (N <= 128) return 128;
  const long long tiles256 = (long long)((M + BM - 1) / BM) * ((N + 255) / 256);
  if (tiles256 < 100 && N > 128) return 128;
  return 256;
}

}  // namespace vf

using namespace vf;

static long long* g_gemm_dbg = nullptr;
// Diagnosis: every following GEMM launch writes, per CTA, {cycles of the MMA issuer's loop, cycles it waited for operands
// (full barriers), cycles it waited for a free accumulator (epilogue), tiles} into buf (int64 [grid][4]); NULL = off.
extern "C" int vf_gemm_set_debug(void* buf) {
  g_gemm_dbg = static_cast<long long*>(buf);
  return VF_OK;
}

extern "C" int vf_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int32_t M,
                            int32_t N, int32_t K, const vf_epilogue* ep, void* stream) {
  VF_REQUIRE(A && W && ep && ep->out, VF_ERR_ARG, "vf_gemm_bf16: null pointer");
  VF_REQUIRE(M > 0 && N > 0 && K > 0, VF_ERR_ARG, "vf_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  VF_REQUIRE(lda >= K && ldw >= K, VF_ERR_ARG, "vf_gemm_bf16: lda/ldw smaller than K");
  VF_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, VF_ERR_ALIGN,
             "vf_gemm_bf16: row pitches must be multiples of 8 elements (16 B) for TMA");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
             VF_ERR_ALIGN, "vf_gemm_bf16: A/W must be 16-byte aligned");
  const bool out_f32 = ep->mode == VF_EPI_BIAS_F32 || ep->mode == VF_EPI_BIAS_RES_F32;
  VF_REQUIRE(ep->ldo >= N, VF_ERR_ARG, "vf_gemm_bf16: ldo smaller than N");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(ep->bias) & 15) == 0, VF_ERR_ALIGN, "vf_gemm_bf16: bias must be 16-byte aligned");
  // 128-bit stores need 16-byte aligned rows; otherwise (e.g. a 10-class head) fall back to scalar stores
  bool vec_ok = (reinterpret_cast<uintptr_t>(ep->out) & 15) == 0 && (ep->ldo % (out_f32 ? 4 : 8)) == 0;
  if (ep->mode == VF_EPI_BIAS_RES_F32) {
    VF_REQUIRE(ep->res, VF_ERR_ARG, "vf_gemm_bf16: residual pointer missing");
    vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(ep->res) & 15) == 0 && (ep->ldr % 4) == 0;
  }
  if (ep->mode == VF_EPI_QKV_ROPE_BF16)
    VF_REQUIRE(vec_ok, VF_ERR_ALIGN, "vf_gemm_bf16: the QKV+RoPE epilogue needs 16-byte aligned output rows");
  if (ep->mode == VF_EPI_QKV_ROPE_BF16)
    VF_REQUIRE(ep->rope_cos && ep->rope_sin && ep->rope_period > 0 && (ep->rope_cols % 64) == 0 &&
                   (N % 64) == 0,
               VF_ERR_ARG, "vf_gemm_bf16: rope epilogue needs cos/sin, period>0, N and rope_cols %% 64 == 0");
  if (ep->mode == VF_EPI_SCATTER_BF16)
    VF_REQUIRE(ep->dst_rows, VF_ERR_ARG, "vf_gemm_bf16: scatter epilogue needs dst_rows");

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  const int bn = pick_bn(M, N);
  p.num_m_blk = (M + BM - 1) / BM;
  p.num_n_blk = (N + bn - 1) / bn;
  p.num_kb = (K + BK - 1) / BK;
  p.bias = ep->bias;
  p.out = ep->out; p.ldo = ep->ldo;
  p.res = ep->res; p.ldr = ep->ldr;
  p.grp_rows = ep->grp_rows; p.grp_stride = ep->grp_stride; p.row_off = ep->row_off;
  p.rope_cos = ep->rope_cos; p.rope_sin = ep->rope_sin;
  p.rope_period = ep->rope_period; p.rope_cols = ep->rope_cols;
  p.dst_rows = ep->dst_rows;
  p.vec_ok = vec_ok ? 1 : 0;
  p.dbg = g_gemm_dbg;
  p.n_peers = 0;
  if (ep->n_peers > 0) {
    VF_REQUIRE(ep->n_peers <= 8, VF_ERR_ARG, "vf_gemm_bf16: at most 8 peer buffers");
    VF_REQUIRE(ep->mode == VF_EPI_BIAS_BF16 || ep->mode == VF_EPI_BIAS_F32, VF_ERR_ARG,
               "vf_gemm_bf16: the fused all-gather needs the bias_bf16 or bias_f32 epilogue");
    VF_REQUIRE(vec_ok && (N & 3) == 0, VF_ERR_ALIGN, "vf_gemm_bf16: the fused all-gather needs aligned rows and N %% 4 == 0");
    for (int i = 0; i < ep->n_peers; ++i) {
      VF_REQUIRE(ep->peer_out[i] && (reinterpret_cast<uintptr_t>(ep->peer_out[i]) & 15) == 0, VF_ERR_ALIGN,
                 "vf_gemm_bf16: peer buffer %d is null or not 16-byte aligned", i);
      p.peer[i] = reinterpret_cast<char*>(ep->peer_out[i]);
    }
    p.n_peers = ep->n_peers;
    p.peer_mc = ep->peer_multicast ? 1 : 0;
    VF_REQUIRE(!p.peer_mc || p.n_peers == 1, VF_ERR_ARG, "vf_gemm_bf16: a multicast destination is ONE address (n_peers == 1)");
  }

  if (ep->ln_xb_out || ep->ln_stat_out) {
    VF_REQUIRE(ep->mode == VF_EPI_BIAS_RES_F32 && ep->grp_rows <= 0, VF_ERR_ARG,
               "vf_gemm_bf16: ln_xb_out / ln_stat_out need the bias_res_f32 epilogue with the identity row map");
    VF_REQUIRE(ep->ln_xb_out && ep->ln_stat_out && ep->ln_stat_ld >= M, VF_ERR_ARG,
               "vf_gemm_bf16: ln_xb_out, ln_stat_out and ln_stat_ld >= M go together");
    VF_REQUIRE(vec_ok && (N % 32) == 0 && ep->ln_ldxb >= N && (ep->ln_ldxb % 4) == 0 &&
                   (reinterpret_cast<uintptr_t>(ep->ln_xb_out) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(ep->ln_stat_out) & 7) == 0,
               VF_ERR_ALIGN, "vf_gemm_bf16: folded LayerNorm (producer) needs aligned rows and N %% 32 == 0");
    p.ln_xb = static_cast<__nv_bfloat16*>(ep->ln_xb_out);
    p.ln_ldxb = ep->ln_ldxb;
    p.ln_stat_out = static_cast<float2*>(ep->ln_stat_out);
    p.ln_stat_ld = ep->ln_stat_ld;
    p.ln_shift = ep->ln_shift;
    VF_REQUIRE((reinterpret_cast<uintptr_t>(ep->ln_shift) & 3) == 0, VF_ERR_ALIGN, "vf_gemm_bf16: ln_shift must be 4-byte aligned");
  }
  if (ep->ln_row_stats || ep->ln_part_in) {
    VF_REQUIRE(!(ep->ln_row_stats && ep->ln_part_in), VF_ERR_ARG, "vf_gemm_bf16: give ln_row_stats OR ln_part_in, not both");
    VF_REQUIRE((ep->mode == VF_EPI_BIAS_BF16 && ep->n_peers == 0) || ep->mode == VF_EPI_GELU_TANH_BF16 ||
                   ep->mode == VF_EPI_GELU_ERF_BF16 || ep->mode == VF_EPI_QKV_ROPE_BF16,
               VF_ERR_ARG, "vf_gemm_bf16: the folded LayerNorm (consumer) needs a bf16 bias / GELU / QKV+RoPE epilogue");
    VF_REQUIRE(ep->ln_colsum, VF_ERR_ARG, "vf_gemm_bf16: folded LayerNorm (consumer) needs ln_colsum");
    // N % 32: every 32-column block of a tile is entirely inside or outside the matrix, so no lane of an epilogue warp
    // leaves the block loop while the others still exchange (mean, rstd) by shuffle
    VF_REQUIRE(vec_ok && (N % 32) == 0 && (reinterpret_cast<uintptr_t>(ep->ln_colsum) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(ep->ln_row_stats) & 7) == 0,
               VF_ERR_ALIGN, "vf_gemm_bf16: folded LayerNorm (consumer) needs aligned rows and N %% 32 == 0");
    p.ln_row_stats = static_cast<const float2*>(ep->ln_row_stats);
    p.ln_colsum = ep->ln_colsum;
    if (ep->ln_part_in) {
      VF_REQUIRE((K % 32) == 0 && ep->ln_stat_ld >= M && (reinterpret_cast<uintptr_t>(ep->ln_part_in) & 7) == 0 &&
                     (ep->ln_variant == 0 || ep->ln_variant == 1),
                 VF_ERR_ARG, "vf_gemm_bf16: ln_part_in needs K %% 32 == 0, ln_stat_ld >= M and ln_variant 0 / 1");
      p.ln_part_in = static_cast<const float2*>(ep->ln_part_in);
      p.ln_stat_ld = ep->ln_stat_ld;
      p.ln_shift_rw = ep->ln_shift_update;
      p.ln_eps = ep->ln_eps;
      p.ln_variant = ep->ln_variant;
    }
  }

  // CTA pairs for every 256-wide problem with at least one full pair of row blocks
  const int cg = (bn == 256 && p.num_m_blk >= 2) ? 2 : 1;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    int e = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)(bn / cg)};
    int e = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // The TMA residual ring takes two of the six main-loop stages: it pays where the epilogue bounds the kernel (short K:
  // proj, 96 -> 81 us) and costs where the main loop does (K = 3072, cold operands: lin2 240 vs 211 us in the step).
  constexpr int kResTmaMaxK = 1024;
  EpiMaps em{};
  if (cg == 2 && ep->mode == VF_EPI_BIAS_RES_F32 && ep->grp_rows <= 0 && vec_ok && (N % 32) == 0 && K <= kResTmaMaxK &&
      (!p.ln_xb || ((p.ln_ldxb % 8) == 0 && (reinterpret_cast<uintptr_t>(p.ln_xb) & 15) == 0))) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint32_t box[2] = {32, 32};
    uint64_t sr[1] = {(uint64_t)ep->ldr * 4}, so[1] = {(uint64_t)ep->ldo * 4}, sx[1] = {(uint64_t)p.ln_ldxb * 2};
    int e = encode_tmap(&em.res, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ep->res, dims, sr, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
    e = encode_tmap(&em.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ep->out, dims, so, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
    if (p.ln_xb) {
      e = encode_tmap(&em.xb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.ln_xb, dims, sx, box, CU_TENSOR_MAP_SWIZZLE_64B);
      if (e) return e;
    }
    p.res_tma = 1;
  }
  if ((ep->mode == VF_EPI_BIAS_BF16 || ep->mode == VF_EPI_GELU_TANH_BF16 || ep->mode == VF_EPI_GELU_ERF_BF16) &&
      ep->grp_rows <= 0 && p.n_peers == 0 && vec_ok && (N % 32) == 0) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint32_t box[2] = {32, 32};
    uint64_t so[1] = {(uint64_t)ep->ldo * 2};
    int e = encode_tmap(&em.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ep->out, dims, so, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (e) return e;
    p.out_tma = 1;
  }
  if (cg == 2) return dispatch_epi<256, 2>(ep->mode, p, tmA, tmB, s, &em);
  return bn == 256 ? dispatch_epi<256, 1>(ep->mode, p, tmA, tmB, s, &em)
                   : dispatch_epi<128, 1>(ep->mode, p, tmA, tmB, s, &em);
}

extern "C" int vf_patch_embed(const void* pixels, int32_t B, int32_t C, int32_t T, int32_t H,
                              int32_t W, int32_t P, int32_t tp, const void* weight,
                              const float* bias, const float* pos, int64_t ld_pos, int32_t N,
                              float* out, int64_t ldo, int64_t out_rows_per_sample,
                              int64_t out_row_off, void* stream) {
  VF_REQUIRE(pixels && weight && out, VF_ERR_ARG, "vf_patch_embed: null pointer");
  VF_REQUIRE(B > 0 && C > 0 && T > 0 && H > 0 && W > 0 && N > 0, VF_ERR_ARG, "vf_patch_embed: bad shape");
  VF_REQUIRE(P == 16, VF_ERR_ARG,
             "vf_patch_embed: patch size %d unsupported (the TMA gather is built for 16x16 patches)", P);
  VF_REQUIRE(tp >= 1 && T % tp == 0 && H % P == 0 && W % P == 0, VF_ERR_ARG,
             "vf_patch_embed: T/H/W not divisible by the patch shape");
  VF_REQUIRE((W % 8) == 0, VF_ERR_ALIGN, "vf_patch_embed: image width must be a multiple of 8 pixels");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(pixels) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ldo % 4) == 0,
             VF_ERR_ALIGN, "vf_patch_embed: pointers must be 16-byte aligned");
  VF_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0 &&
                 (ld_pos % 4) == 0 && (N % 4) == 0,
             VF_ERR_ALIGN, "vf_patch_embed: bias/pos must be 16-byte aligned, ld_pos and N multiples of 4");
  const int K = C * tp * P * P;
  VF_REQUIRE(K % BK == 0, VF_ERR_ARG, "vf_patch_embed: C*tp*P*P must be a multiple of 64");
  const int nh = H / P, nw = W / P, Tp = T / tp;

  // tile rectangle: 16 patch rows x 8 patch columns (see the SW32 layout note at the top)
  const int best_pw = 8;
  GemmParams p{};
  p.patch = 1;
  p.PW = best_pw; p.PH = BM / best_pw;
  p.nw = nw; p.nh = nh;
  p.n_pwb = (nw + p.PW - 1) / p.PW;
  p.n_phb = (nh + p.PH - 1) / p.PH;
  p.Tp = Tp; p.T = T; p.C = C; p.tp = tp; p.P = P;
  p.pos = pos; p.ld_pos = ld_pos;
  p.M = B * Tp * nh * nw; p.N = N; p.K = K;
  const int bn = N <= 128 ? 128 : 256;
  p.num_m_blk = B * Tp * p.n_phb * p.n_pwb;
  // CTA pairs (two row blocks = two patch rectangles per 256 x 256 tile, each CTA stages half of the W tile): a single CTA
  // pulls 48 KB of operands per 512 tensor cycles through L2 -> SM (96 B/clk, above what the fabric delivers per SM: 610-670 TF),
  // a pair 32 KB
  const int cg = (bn == 256 && p.num_m_blk >= 2) ? 2 : 1;
  p.num_n_blk = (N + bn - 1) / bn;
  p.num_kb = K / BK;
  p.bias = bias;
  p.out = out; p.ldo = ldo;
  p.grp_stride = out_rows_per_sample; p.row_off = out_row_off;
  p.vec_ok = 1;

  CUtensorMap tmA, tmB;
  {
    // dims (fastest first): px[P], pw[nw], py[P], ph[nh], image plane [B*C*T]
    uint64_t dims[5] = {(uint64_t)P, (uint64_t)nw, (uint64_t)P, (uint64_t)nh, (uint64_t)B * C * T};
    uint64_t strides[4] = {(uint64_t)P * 2, (uint64_t)W * 2, (uint64_t)W * P * 2, (uint64_t)H * W * 2};
    uint32_t box[5] = {(uint32_t)P, (uint32_t)p.PW, (uint32_t)(BK / P), (uint32_t)p.PH, 1};
    int e = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, pixels, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_32B);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {BK, (uint32_t)(bn / cg)};
    int e = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, weight, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
    if (e) return e;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cg == 2) return launch_gemm<VF_EPI_BIAS_F32, 256, true, 2>(p, tmA, tmB, s);
  return bn == 256 ? launch_gemm<VF_EPI_BIAS_F32, 256, true>(p, tmA, tmB, s)
                   : launch_gemm<VF_EPI_BIAS_F32, 128, true>(p, tmA, tmB, s);
}
