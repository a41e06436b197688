/* vfuse.h — C ABI of libvfuse.so: the B200 (sm_100a) vision-encode-and-fuse kernels.
 *
 * The reference (casinca/LLM-quest) is pure Python/PyTorch and has NO FFI of its own for this path
 * (SURVEY.md §8b); its boundary is a set of nn.Module.forward methods. Every entry point below
 * therefore cites the reference *Python* call site whose ATen ops it replaces (paths relative to
 * the reference root). The Python drop-in modules in llm_quest_b200/ bind these with ctypes
 * (llm_quest_b200/_lib.py); INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - nothing allocates, nothing synchronises: work is enqueued on `stream` (a cudaStream_t passed
 *     as void*; NULL = legacy default stream);
 *   - return value 0 = ok, negative = error (VF_ERR_*); vf_last_error() gives the message
 *     (thread-local);
 *   - bf16 = raw IEEE bfloat16 bits (uint16_t), row-major, innermost dimension contiguous;
 *   - there is NO CPU fallback: on a machine without an sm_100 GPU every compute entry point
 *     returns VF_ERR_NO_DEVICE / VF_ERR_CUDA.
 */
#ifndef VFUSE_H_
#define VFUSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VF_VERSION 200 /* 0.2.0: one folded-LayerNorm protocol (row shift), multimem all-gather stores, head_dim != 64 attention */

enum {
  VF_OK = 0,
  VF_ERR_ARG = -1,
  VF_ERR_CUDA = -2,
  VF_ERR_NO_DEVICE = -3,
  VF_ERR_ALIGN = -4
};

int vf_version(void);
const char* vf_last_error(void);
/* number of kernels this library has launched since load (or since the last reset); bench.py
 * reports it as "gpu_launches". */
int64_t vf_launch_count(void);
void vf_launch_count_reset(void);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction: out = epilogue(A[M,K] · W[N,K]^T)  — tcgen05/TMEM GEMM, TMA-fed.
 * A, W bf16; fp32 accumulate. Replaces every nn.Linear on the path:
 *   qkv / proj        llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:150-151,168,190
 *   ffn lin1 / lin2   llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:120-125
 *   merger lin1/lin2  llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:407-409,429
 *   Part-1 q/k/v/out  llm_quest/multimodal/vision_transformer/vit_attention.py:58-60,89
 *   Part-1 ffn, head  llm_quest/multimodal/vision_transformer/vit_transformer_block.py:58-67,
 *                     vit_model.py:158-159
 *   ViTAdapter        llm_quest/multimodal/vision_transformer/vit_engine.py:44-59
 * ------------------------------------------------------------------------------------------- */
typedef enum {
  VF_EPI_BIAS_BF16 = 0,      /* out_bf16 = acc + bias                                         */
  VF_EPI_BIAS_F32 = 1,       /* out_f32  = acc + bias                                         */
  VF_EPI_BIAS_RES_F32 = 2,   /* out_f32  = acc + bias + res_f32 (res may alias out)           */
  VF_EPI_GELU_TANH_BF16 = 3, /* out_bf16 = gelu_tanh(acc + bias)   (vision_model.py:122)      */
  VF_EPI_GELU_ERF_BF16 = 4,  /* out_bf16 = gelu_erf(acc + bias)    (vision_model.py:408)      */
  VF_EPI_QKV_ROPE_BF16 = 5,  /* out_bf16 = rope2d(acc + bias) on cols < rope_cols, head=64    */
  VF_EPI_SCATTER_BF16 = 6    /* out_bf16[dst_rows[m]] = acc + bias (early-fusion scatter)     */
} vf_epilogue_mode;

typedef struct {
  int32_t mode;            /* vf_epilogue_mode */
  const float* bias;       /* [N] fp32 or NULL */
  void* out;               /* bf16 or fp32 per mode */
  int64_t ldo;             /* out row pitch in elements */
  const float* res;        /* [.., ldr] fp32 residual (VF_EPI_BIAS_RES_F32) */
  int64_t ldr;
  /* output row remap: out_row = (m / grp_rows) * grp_stride + (m % grp_rows) + row_off.
   * grp_rows <= 0 means identity. Used to write adapter rows into a wider fused buffer
   * (multimodal/vlm_engine.py:114 torch.cat) */
  int32_t grp_rows;
  int64_t grp_stride;
  int64_t row_off;
  /* VF_EPI_QKV_ROPE_BF16: cos/sin tables [rope_period, 32] fp32 (the first half of the reference's
   * duplicated [n,64] tables, common/rope.py:477-480); row m uses table row m % rope_period. */
  const float* rope_cos;
  const float* rope_sin;
  int32_t rope_period;
  int32_t rope_cols;       /* columns [0, rope_cols) are rotated (q and k); multiple of 64 */
  /* VF_EPI_SCATTER_BF16 */
  const int32_t* dst_rows; /* [M] destination row or -1 */
  /* Fused all-gather (VF_EPI_BIAS_BF16 / VF_EPI_BIAS_F32, n_peers > 0): every output element is stored to the same
   * (row, column) of EACH of the n_peers buffers instead of `out` — peer-mapped device memory of the other GPUs of
   * the box (NVLink P2P) plus this GPU's own copy, each pointer already offset to this rank's first row. Replaces
   * "GEMM, then ncclAllGather of its output" for the sample-sharded path. `out` is ignored. */
  int32_t n_peers;         /* 0 = store to `out` only */
  void* peer_out[8];
  /* peer_out[0] is an NVSwitch MULTICAST mapping of the gathered buffer (n_peers must be 1): the epilogue writes it
   * with multimem.st — the only instruction family that may touch such an address — and the switch replicates every
   * store into all GPUs' copies. */
  int32_t peer_multicast;
  /* LayerNorm folded into the two GEMMs around it (pre-LN block, qwen3_5_vision_model.py:195-238):
   *   LN(x) @ W^T + b  =  rstd * (x' @ (gamma.W)^T - mean' * colsum) + (b + W beta),   colsum[n] = sum_k bf16(gamma_k W[n,k]),
   * for ANY per-row shift s with x' = x - s and mean' = mean(x') (LayerNorm is shift-invariant). The bf16 rounding sits
   * on x' instead of LN(x), so its error grows by sqrt(1 + (mean'/sigma)^2): with s = the row's mean at the previous
   * LayerNorm point (the residual stream moves slowly) mean' stays far below sigma even for rows whose mean is 50 sigma.
   * PRODUCER side (VF_EPI_BIAS_RES_F32, identity row map): besides the fp32 row x the epilogue writes bf16(x - ln_shift[row])
   * to ln_xb_out (the next GEMM's A operand) and, per 32-column block j, the partial row sums
   * ln_stat_out[j * ln_stat_ld + row] = (sum x', sum x'^2) as float2 — N/32 partials per row, no atomics, fixed order.
   * vf_ln_row_stats() turns the partials into (mean', rstd) per row and advances the shift to the row's true mean.
   * CONSUMER side (VF_EPI_BIAS_BF16, VF_EPI_GELU_*_BF16, VF_EPI_QKV_ROPE_BF16; N % 32 == 0): ln_row_stats holds (mean', rstd) of every row
   * of A; the epilogue applies the identity above with ln_colsum [N] fp32. `bias` must already hold b + W beta and W
   * must already be gamma-scaled (host side, once per weight). */
  void* ln_xb_out;           /* bf16 [rows, ln_ldxb] or NULL */
  int64_t ln_ldxb;
  void* ln_stat_out;         /* float2 [N/32][ln_stat_ld] or NULL (required with ln_xb_out) */
  int64_t ln_stat_ld;        /* rows per partial plane (>= M) */
  const float* ln_shift;     /* fp32 [M] row shifts or NULL (= 0) */
  const void* ln_row_stats;  /* float2 [M] (mean', rstd) or NULL */
  const float* ln_colsum;    /* [N] fp32 (required with ln_row_stats / ln_part_in) */
  /* CONSUMER, small problems: instead of ln_row_stats give the producer's partial sums (ln_part_in = its ln_stat_out,
   * K/32 planes of ln_stat_ld rows) and the launch of vf_ln_row_stats in between is dropped: every epilogue warp adds up
   * the partials of its 32 rows itself, in vf_ln_row_stats' summation order (same bits), eps = ln_eps, ln_variant as in
   * vf_ln_row_stats; the first column tile of every row block advances ln_shift_update[row] by mean' (may be NULL). Every
   * column tile re-reads the partials, so this only pays while the launch costs more than K/32 x 8 bytes per row and
   * column tile (M up to a few thousand rows). */
  const void* ln_part_in;    /* float2 [K/32][ln_stat_ld] or NULL */
  float* ln_shift_update;    /* fp32 [M] or NULL */
  float ln_eps;
  int32_t ln_variant;
} vf_epilogue;

int vf_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int32_t M, int32_t N,
                 int32_t K, const vf_epilogue* ep, void* stream);
/* Diagnosis only: every following vf_gemm_bf16 launch writes, per CTA, {cycles of the MMA issuer's loop, cycles it waited
 * for operands, cycles it waited for a free accumulator, tiles} into buf (device int64 [grid][4]); NULL switches it off. */
int vf_gemm_set_debug(void* buf);

/* ---------------------------------------------------------------------------------------------
 * Patch embedding as an im2col-free GEMM: the A operand is gathered by 5-D TMA boxes straight from
 * the pixel tensor. Replaces
 *   nn.Conv3d(k=s=(tp,P,P)) + flatten(2).transpose(1,2) + pos-embed add
 *       llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:79-86,105-107,353-358
 *   nn.Conv2d(k=s=P) + flatten/transpose (+ cls/pos add done by vf_vit_cls_pos)
 *       llm_quest/multimodal/vision_transformer/vit_model.py:48-55,77-87,145   (T = tp = 1)
 * pixels: bf16 [B, C, T, H, W]; weight: bf16 [N, C*tp*P*P] (the conv weight flattened, K order
 * c,dt,py,px); out: fp32, token (b, t', ph, pw) is written to row
 *   b*out_rows_per_sample + out_row_off + (t'*nh + ph)*nw + pw,   pitch ldo elements,
 * and gets  + bias[N] + pos[(ph*nw + pw) * ld_pos + :]  (pos may be NULL).
 * ------------------------------------------------------------------------------------------- */
int vf_patch_embed(const void* pixels, int32_t B, int32_t C, int32_t T, int32_t H, int32_t W,
                   int32_t P, int32_t tp, const void* weight, const float* bias, const float* pos,
                   int64_t ld_pos, int32_t N, float* out, int64_t ldo, int64_t out_rows_per_sample,
                   int64_t out_row_off, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused bidirectional attention, head_dim 64, bf16 in/out, fp32 softmax (tcgen05 + TMEM).
 * Replaces F.scaled_dot_product_attention(q,k,v) + the transposes around it
 *   llm_quest/qwen/qwen3_5/qwen3_5_vision_model.py:169-190
 *   llm_quest/multimodal/vision_transformer/vit_attention.py:62-87
 * qkv: bf16 [B*S, 3*H*64] token-major, columns [q heads | k heads | v heads] (the layout
 * nn.Linear(d, 3d) produces, vision_model.py:168-173); out: bf16 [B*S, H*64] token-major.
 * Attention is over the S tokens of one sample; scale = softmax scale (1/sqrt(64)).
 * ------------------------------------------------------------------------------------------- */
int vf_attention_fwd(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, float scale,
                     void* stream);
/* Same contract for any head_dim that is a multiple of 8 up to 128 (qkv [B*S, 3*H*head_dim], out [B*S, H*head_dim]):
 * head_dim 64 is the tensor-core kernel above; other head dims (TINY_VIT_CONFIG, config.py:175-186: head_dim 32, S = 65)
 * run a CUDA-core kernel with K/V of a (sample, head) staged in shared memory (S <= 2048, 200 KB of shared memory). */
int vf_attention_fwd_hd(const void* qkv, void* out, int32_t B, int32_t S, int32_t H, int32_t head_dim, float scale,
                        void* stream);
/* Diagnosis only: with a trace build selected (VF_ATTN_FLAGS bit 1) block 0 of every following
 * vf_attention_fwd writes clock64 stamps into buf — uint64 [20 chains][n_steps][8] — for the key steps
 * [first_step, first_step + n_steps) of each softmax chain / MMA walker. buf = NULL switches it off. */
int vf_attention_set_trace(void* buf, int32_t first_step, int32_t n_steps);

/* ---------------------------------------------------------------------------------------------
 * Causal grouped-query attention, head_dim 256, bf16 in/out, fp32 softmax (tcgen05 + TMEM): the attention core
 * of the first consumer of the fused embeddings and position ids. Replaces
 *   F.scaled_dot_product_attention(q, k, v, attn_mask=causal, enable_gqa=True) and `ctx * sigmoid(gate)`
 *   llm_quest/qwen/qwen3_5/qwen3_5_text_model.py:246-262 (prefill: no KV cache, no padding mask)
 * q: bf16 token-major [B*S, ldq], head h at columns q_col0 + h*q_head_stride .. +256; k, v: [B*S, ldk/ldv], kv head g
 * at columns 256g; query head h uses kv head h / (Hq/Hkv). out: bf16 [B*S, ldo], head h at columns 256h.
 * gate (optional, may be NULL): bf16 [B*S, ldg]; out is multiplied by sigmoid(gate[row, gate_col0 +
 * h*gate_head_stride + d]) before the single bf16 rounding. causal != 0: key t' <= query t inside a sample.
 * ------------------------------------------------------------------------------------------- */
int vf_attention_gqa_fwd(const void* q, int64_t ldq, int32_t q_col0, int32_t q_head_stride, const void* k,
                         int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo, const void* gate,
                         int64_t ldg, int32_t gate_col0, int32_t gate_head_stride, int32_t B, int32_t S, int32_t Hq,
                         int32_t Hkv, int32_t head_dim, float scale, int32_t causal, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last dim, fp32 or bf16 in, bf16 or fp32 out, fp32 statistics.
 *   variant 0: (x-mean)/sqrt(var+eps)*w+b   nn.LayerNorm    vision_model.py:213-214,229,234,406
 *   variant 1: (x-mean)/(std+eps)*w+b       Part-1 LayerNorm vit_transformer_block.py:21-31
 * merge > 1 additionally applies the 2x2 (merge x merge) spatial-merge gather of ViTMergeAdapter
 * (vision_model.py:425-427): input token (f, r, c) of a sample with nh x nw patches per frame is
 * written to row ((f*(nh/m) + r/m)*(nw/m) + c/m) of the sample, feature slot (r%m)*m + c%m.
 * in_dtype / out_dtype: 0 = fp32, 1 = bf16.
 * ------------------------------------------------------------------------------------------- */
int vf_layernorm(const void* x, int32_t in_dtype, int64_t ldx, const float* w, const float* b,
                 void* out, int32_t out_dtype, int64_t rows, int32_t D, float eps, int32_t variant,
                 int32_t merge, int32_t nh, int32_t nw, float* mean_out, void* stream);
/* mean_out (fp32 [rows] or NULL): the row means, i.e. the first row shift of a chain of folded LayerNorms
 * (vf_epilogue.ln_shift). */

/* (mean', rstd) per row from the partial sums a folded-LayerNorm producer epilogue left (vf_epilogue.ln_stat_out):
 * partials float2 [parts][ld] -> out float2 [rows], mean' = sum / D, var = max(sumsq / D - mean'^2, 0),
 * rstd = rsqrt(var + eps) (variant 0, nn.LayerNorm) or 1 / (sqrt(var) + eps) (variant 1, Part-1 LayerNorm).
 * shift (fp32 [rows] or NULL): shift[row] += mean' — the producer subtracted shift[row], so this is the row's true mean,
 * which the NEXT producer subtracts. Summation order is fixed (part 0, 1, ...): results do not depend on the batch a
 * row sits in. */
int vf_ln_row_stats(const void* partials, int32_t parts, int64_t ld, int64_t rows, int32_t D, float eps,
                    int32_t variant, void* out, float* shift, void* stream);

/* Part-1 class-token rows: out[b*S + 0, :] = cls[:] + pos[0, :]   (vit_model.py:86-87,145) */
int vf_vit_cls_pos(const float* cls, const float* pos, float* out, int32_t B, int64_t rows_per_sample,
                   int32_t D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rotate-half RoPE, standalone (drop-in for VisionRoPE.apply / RoPE.apply,
 * llm_quest/common/rope.py:180-243,485-500). x: [B, H, S, hd] (dtype 0 fp32 / 1 bf16), contiguous.
 * cos/sin: fp32 [>=S, rot] (duplicated halves as the reference builds them); position_ids: int64
 * [B, S] or NULL (then position = s). rot <= hd; columns >= rot pass through.
 * ------------------------------------------------------------------------------------------- */
int vf_rope_apply(const void* x, void* out, int32_t dtype, int32_t B, int32_t H, int32_t S,
                  int32_t hd, const float* cos, const float* sin, int32_t rot, int64_t table_rows,
                  const int64_t* position_ids, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MRoPE-I apply with optional fused zero-centred RMSNorm over hd.
 * Replaces RoPE.apply_mrope + interleave_mrope_coeffs (common/rope.py:246-358) and, when
 * norm_weight != NULL (fp32 [hd], already 1+scale), the ZeroCenteredRMSNorm before it
 * (qwen/qwen3_next/qwen3_next_attention.py:41-46; call site qwen3_5_text_model.py:227-233).
 * x/out: [B, H, S, hd]; cos/sin: fp32 [table_rows, rot]; position_ids: int64 [3, B, S];
 * sections: the three mrope_section ints (T, H, W).
 * ------------------------------------------------------------------------------------------- */
int vf_mrope_apply(const void* x, void* out, int32_t dtype, int32_t B, int32_t H, int32_t S,
                   int32_t hd, const float* cos, const float* sin, int32_t rot, int64_t table_rows,
                   const int64_t* position_ids, int32_t sec_t, int32_t sec_h, int32_t sec_w,
                   const float* norm_weight, float norm_eps, void* stream);
/* Same, with explicit element strides {batch, head, token} for x and out (multiples of 8): lets q / k be
 * normalised and rotated IN PLACE inside the token-major output of the w_queries_gate / w_keys projections
 * (qwen3_5_text_model.py:227-233 without the transposes). x == out is allowed. */
int vf_mrope_apply_strided(const void* x, void* out, int32_t dtype, int32_t B, int32_t H, int32_t S, int32_t hd,
                           const int64_t* x_strides, const int64_t* out_strides, const float* cos, const float* sin,
                           int32_t rot, int64_t table_rows, const int64_t* position_ids, int32_t sec_t, int32_t sec_h,
                           int32_t sec_w, const float* norm_weight, float norm_eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MRoPE 3-D position ids (bit-exact integer work). Replaces Qwen3_5VLM.compute_3d_position_ids
 * (llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py:85-176).
 * input_ids int64 [b, seq]; image_mask uint8/bool [b, seq] or NULL (then ids == image_token_id);
 * feeds_host: HOST int64 [n_feeds, 3] (t, h, w) — the reference keeps it on the CPU too (:83);
 * out int64 [3, b, seq]. n_feeds == 0 gives the text-only arange.
 * ------------------------------------------------------------------------------------------- */
int vf_mrope_position_ids(const int64_t* input_ids, const uint8_t* image_mask, int64_t image_token_id,
                          const int64_t* feeds_host, int32_t n_feeds, int32_t merge, int32_t b,
                          int32_t seq, int64_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Early fusion: embedding gather + masked scatter of vision rows in one pass. Replaces
 *   emb_dict(input_ids) ... masked_scatter(image_mask, vision_embeds.to(dtype))
 *   llm_quest/qwen/qwen3_5/qwen3_5_vlm_model.py:198-211
 * table bf16 [vocab, D]; vision [n_vis, D] (vis_dtype 0 fp32 / 1 bf16); out bf16 [b*seq, D].
 * The j-th placeholder in flat (b, seq) order receives vision row j. row_map (int32 [b*seq],
 * optional) receives j for placeholder rows and -1 elsewhere; n_placeholders (int32[1], optional)
 * the count. scratch: int32 [b*seq + 1024] workspace. If there are more placeholders than n_vis
 * rows the extra rows are left as the table row and *n_placeholders still reports the count (the
 * host mirror raises like masked_scatter does).
 * vf_fuse_scan produces row_map / n_placeholders and, when inv_map != NULL, the inverse map
 * inv_map[j] = flat token row of the j-th placeholder (-1 for j >= #placeholders; int32 [inv_cap]),
 * which is what VF_EPI_SCATTER_BF16 consumes as dst_rows.
 * ------------------------------------------------------------------------------------------- */
int vf_fuse_scan(const int64_t* input_ids, const uint8_t* image_mask, int64_t image_token_id,
                 int64_t n_tokens, int32_t* row_map, int32_t* n_placeholders, int32_t* inv_map,
                 int64_t inv_cap, int32_t* scratch, void* stream);
int vf_embed_gather_scatter(const int64_t* input_ids, const void* table, int64_t vocab, int32_t D,
                            const void* vision, int32_t vis_dtype, int64_t n_vis,
                            const int32_t* row_map, void* out, int64_t n_tokens, int32_t skip_vision,
                            void* stream);

/* dtype casts used at the module seam (fp32 pixels -> bf16, vision_model.py:210 .to(dtype)) */
int vf_cast_f32_to_bf16(const float* x, void* out, int64_t n, void* stream);
int vf_cast_bf16_to_f32(const void* x, float* out, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * uint8 image -> normalised pixel tensor for PatchEmbedding3D. Replaces, after the
 * host-side resize, to_tensor + normalize + temporal duplication + permute of
 *   llm_quest/qwen/qwen3_5/qwen3_5_generate_multimodal.py:40-46 (and dataset.py:336-351).
 * img: uint8 [B, H, W, 3] (HWC, device); mean3 / std3: HOST float[3]; out: [B, 3, T, H, W] fp32 (out_dtype 0,
 * bit-identical to torchvision) or bf16 (1); every temporal slot holds the same frame. W % 4 == 0.
 * ------------------------------------------------------------------------------------------- */
int vf_preprocess_u8(const uint8_t* img, int32_t B, int32_t H, int32_t W, int32_t T, const float* mean3,
                     const float* std3, void* out, int32_t out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone pieces of the module surface (on the path proper they are GEMM epilogues or fused elsewhere).
 * dtype: 0 = fp32, 1 = bf16 (in and out).
 * ------------------------------------------------------------------------------------------- */
/* GELU.forward of the Part-1 ViT (x * 0.5 * (1 + erf(x / sqrt 2)), vit_transformer_block.py:43-44; tanh_form = 0) and
 * nn.GELU(approximate="tanh") (qwen3_5_vision_model.py:122; tanh_form = 1), erff / tanhf accuracy. */
int vf_gelu(const void* x, void* out, int32_t dtype, int64_t n, int32_t tanh_form, void* stream);
/* ZeroCenteredRMSNorm.forward (qwen/qwen3_next/qwen3_next_attention.py:41-46): fp32 inside,
 * out = ((x * rsqrt(mean(x^2) + eps)) * one_plus_scale) cast back to dtype; one_plus_scale fp32 [D] = 1 + scale. */
int vf_rmsnorm_zc(const void* x, int64_t ldx, const float* one_plus_scale, void* out, int64_t ldo, int32_t dtype,
                  int64_t rows, int32_t D, float eps, void* stream);
/* Part-2 text half of the early fusion: get_embeddings (multimodal/vlm_engine.py:5-20) + torch.cat([vision, text], 1)
 * (vlm_engine.py:114, vlm_generation.py:66). out fp32 [b, rows_per_sample, D]:
 *   out[bi, row_off + t, :] = tok_table[input_ids[bi, t], :] + pos_table[t, :]      (sum rounded to the tables' dtype)
 * tables fp32 or bf16 (table_dtype), D % 8 == 0, seq <= n_pos. The vision rows [0, row_off) are written by the
 * adapter's GEMM through its row remap (vf_epilogue.grp_rows). */
int vf_embed_pos_concat(const int64_t* input_ids, const void* tok_table, int64_t vocab, const void* pos_table,
                        int64_t n_pos, int32_t table_dtype, float* out, int32_t b, int32_t seq, int32_t D,
                        int64_t rows_per_sample, int64_t row_off, void* stream);
/* Patch sizes other than 16 (TINY_VIT_CONFIG: 4 x 4): pixels [B, C, H, W] (dtype) -> bf16 rows [B*nh*nw, ld_out],
 * columns (c, py, px) = the flattened conv weight; the patch embedding is then vf_gemm_bf16 on those rows
 * (vit_model.py:77-83). */
int vf_im2col_patches(const void* pixels, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W, int32_t P, void* out,
                      int64_t ld_out, void* stream);
/* out[b, r, :] = src[r, :] (+ add_row0[:] when r == 0) for every sample: the position-embedding (+ class token) rows
 * the small-patch path accumulates its patch GEMM onto (vit_model.py:86-87,145). */
int vf_fill_rows_f32(const float* src, const float* add_row0, float* out, int32_t B, int64_t rows_per_sample, int32_t D,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VFUSE_H_ */
