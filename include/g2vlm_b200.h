/*
 * g2vlm_b200 — C ABI of the B200-native (sm_100a) kernels behind G2VLM's reconstruction forward
 * pass (`G2VLM.recon`, reference modeling/g2vlm/g2vlm.py:1240-1303).
 *
 * The reference has NO FFI / native layer: its "operators" are PyTorch
 * library calls. Each entry point below therefore names the reference *Python call site* it
 * replaces (file:line under /root/reference). Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless the field
 *     comment says HOST; the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it;
 *   - return value: 0 = ok, otherwise a G2VLM_ERR_* code; g2vlm_last_error() returns the message of
 *     the last failure on the calling thread; no exceptions cross the boundary;
 *   - bf16 tensors are row-major with a leading dimension given in ELEMENTS.
 */
#ifndef G2VLM_B200_H_
#define G2VLM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2VLM_OK 0
#define G2VLM_ERR_INVALID 1 /* bad argument / unsupported shape */
#define G2VLM_ERR_CUDA 2    /* a CUDA runtime / driver call failed */

#define G2VLM_ABI_VERSION 4 /* 2: gemm out_col_group/out_col_stride, attention item_causal + out_head_cols;
                               3: gemm FORCE_PAIR/FORCE_SINGLE flags, attention lse_out + max_ctas, g2vlm_attention_merge;
                               4: g2vlm_ply_pack filter_nonfinite; decode_step fused_ws (one-kernel step),
                                  g2vlm_und_decode_workspace_bytes, g2vlm_und_prefill, g2vlm_sp_kv_exchange */

/* Version of this ABI (G2VLM_ABI_VERSION of the built library). */
int g2vlm_abi_version(void);
/* Message of the last error raised on this thread ("" if none). */
const char* g2vlm_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Grouped (token-type-routed) bf16 GEMM on tcgen05 tensor cores:  C[rows of group g] = A · B_g^T
 *
 * Replaces every `nn.Linear` executed under `torch.amp.autocast(bf16)` on the path:
 *   MoT routed q/k/v/o_proj[_moe_geo]      modeling/g2vlm/qwen2vl.py:584-594, 653-658
 *   MoT routed Qwen2MLP gate/up/down       modeling/qwen2vl/modeling_qwen2_vl.py:519-521,
 *                                          modeling/g2vlm/qwen2vl.py:894-909
 *   DINO query/key/value/dense/fc1/fc2     modeling/g2vlm/dinov2_model.py:45-47, 74-78, 174-178
 *   DINO patch projection (im2col GEMM)    modeling/dinov2_with_registers/modeling_dinov2_with_registers.py:71
 *   dino2llm                               modeling/g2vlm/g2vlm.py:1001
 *   Pi3 qkv/proj/fc1/fc2/linear_out        modeling/pi3/models/layers/attention.py:360-379,
 *                                          modeling/pi3/models/layers/transformer_head.py:55
 *   fp32 head Linears (split-bf16 x3)      modeling/pi3/models/layers/transformer_head.py:76,
 *                                          modeling/pi3/models/layers/camera_head.py:25-61
 *
 * A is [a_rows, K] bf16 (tokens, already permuted so that each expert's rows are contiguous);
 * B is [n_groups * N, K] bf16 = the experts' nn.Linear weights stacked along rows. Group g owns A
 * rows [group_row0[g], group_row0[g] + group_rows[g]) and weight rows [g*N, (g+1)*N). Output rows
 * are the A rows (same index). fp32 accumulation in TMEM; the epilogue applies the reference's
 * rounding points (autocast returns bf16 from every Linear). Calls with <= 8 rows in a single group (the
 * decode steps of generate_text) are HBM-bound and run a warp-per-column GEMV kernel with the same epilogues.
 * ---------------------------------------------------------------------------------------------- */
#define G2VLM_EPI_STORE_BF16 0  /* out_bf16 = bf16(acc + bias) [; gelu] */
#define G2VLM_EPI_SWIGLU_BF16 1 /* B rows interleave gate/up in blocks of 128; out[:, N/2] */
#define G2VLM_EPI_RESID_F32 2   /* out_f32 += [bf16](scale * bf16(acc + bias)) */
#define G2VLM_EPI_STORE_F32 3   /* out_f32 = [resid +] [relu] [scale *] [gelu] [bf16] (acc + bias) [+ out_f32] */

#define G2VLM_GEMM_GELU 1u              /* STORE_BF16: exact-erf GELU on the bf16-rounded value; STORE_F32 (fp32 mode): on the fp32 value */
#define G2VLM_GEMM_ROUND_AFTER_SCALE 2u /* RESID_F32: round scale*x to bf16 (MoT ls1/ls2) */
#define G2VLM_GEMM_RELU 4u              /* STORE_F32 */
#define G2VLM_GEMM_ACCUMULATE 8u        /* STORE_F32: out += value (second pass of split-bf16) */
#define G2VLM_GEMM_ROUND_BF16 16u       /* STORE_F32: round (acc+bias) to bf16 before storing */
#define G2VLM_GEMM_QUICK_GELU 32u       /* STORE_BF16: x*sigmoid(1.702x) on the bf16-rounded value (Qwen2-VL ViT MLP) */
#define G2VLM_GEMM_ROUND_SUM 64u        /* RESID_F32: round the updated stream value to bf16 (bf16 residual stream) */
#define G2VLM_GEMM_FORCE_PAIR 128u      /* run the CTA-pair kernel (256x256 tiles) whatever the problem size */
#define G2VLM_GEMM_FORCE_SINGLE 256u    /* run the 1-CTA kernel (128x256 tiles) whatever the problem size */
#define G2VLM_GEMM_NO_TMA_OUT 512u      /* RESID_F32: SM-side read-modify-write instead of the TMA reduce-add (A/B timing) */

typedef struct g2vlm_gemm_args {
  const void* A; /* bf16 [a_rows, K], leading dimension lda */
  int64_t lda;
  int64_t a_rows;
  const void* B; /* bf16 [n_groups * N, K], leading dimension ldb */
  int64_t ldb;
  int32_t N;
  int32_t K;
  int32_t n_groups; /* 1 or 2 */
  int32_t group_row0[2];
  int32_t group_rows[2];
  int32_t epilogue; /* G2VLM_EPI_* */
  uint32_t flags;   /* G2VLM_GEMM_* */
  void* out;        /* bf16 or fp32 depending on the epilogue */
  int64_t ldo;
  const float* bias;       /* fp32 [n_groups, N] or NULL */
  const float* scale;      /* fp32 [N] LayerScale / lambda, or NULL */
  uint32_t scale_groups;   /* bit g set: group g rows are multiplied by `scale` */
  const float* residual;   /* STORE_F32: fp32 [rows, ldr] added to the result, or NULL */
  int64_t ldr;
  /* STORE_BF16 only: output column c is written at (c / out_col_group) * out_col_stride + c % out_col_group
   * (both multiples of 32, stride >= group; 0 = plain).  Lets a Linear whose outputs are heads of 96 write
   * straight into the 128-wide head slots the attention kernel reads, without multiplying zero weight rows
   * (Pi3 decoders: modeling/pi3/models/layers/attention.py:353-357). The pad columns are not written. */
  int32_t out_col_group;
  int32_t out_col_stride;
  /* ABI v3, STORE_F32 only (fp32 mode): 0, or the number of 64-element k-blocks after which the accumulator leaves
   * the tensor core; the chunks are then summed in fp32 round-to-nearest by the epilogue (bias with the first chunk,
   * gelu / scale / relu / residual with the last).  The tensor core accumulates with truncation: a split-form K of
   * tens of thousands loses ~1e-5 relative in one go, 4 k-blocks (16 MMA steps) per chunk keep it below 1e-6. */
  int32_t k_chunk_blocks;
} g2vlm_gemm_args;

int g2vlm_gemm_bf16(const g2vlm_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Varlen / GQA attention on tcgen05 tensor cores:  out = softmax(scale * Q K^T [+ causal mask]) V
 *
 * Replaces
 *   flash_attn_varlen_func (MoT shared attention)   modeling/g2vlm/qwen2vl.py:643-652
 *   flash_attn_varlen_func (DINO, segments of P)    modeling/g2vlm/dinov2_model.py:49-58
 *   SDPA self-attention per view (Pi3 decoders)     modeling/pi3/models/layers/attention.py:370-372
 *   SDPA cross-attention to view 0                  modeling/pi3/models/layers/attention.py:255-259
 *
 * q is [q_rows, num_q_heads*head_dim] bf16 (leading dim ldq), k / v are [kv_rows,
 * num_kv_heads*head_dim] (ldk / ldv), out is [q_rows, num_q_heads*head_dim] (ldo). The segment
 * structure (cu_seqlens_q / cu_seqlens_k of flash-attn; per-view batches of SDPA) is given as a
 * DEVICE table of work items, 8 int32 each:
 *   {q_tile_begin, q_seg_begin, q_seg_end, k_begin, k_end, item_causal, 0, 0}
 * one item per <= 256 consecutive query rows [q_tile_begin, min(q_tile_begin+256, q_seg_end)) of a
 * segment whose queries are rows [q_seg_begin, q_seg_end) and whose keys are rows [k_begin, k_end).
 * Query rows covered by no item are NOT written (flash-attn leaves them uninitialised, SURVEY.md
 * quirk Q1; the caller defines them, the host mirror zero-fills). causal = bottom-right aligned
 * mask as in flash-attn, for the whole launch (args.causal) or for one item (item_causal != 0: lets
 * the causal prompt rows ride in the same launch as the non-causal geo step). K/V rows in [k_end, round_up(k_end,128)) that lie inside kv_rows must hold
 * finite values. head_dim in {64, 128} (the 96-wide Pi3 heads sit zero-padded in 128-wide q/k/v slots; see
 * out_head_cols for the compact output).
 * ---------------------------------------------------------------------------------------------- */
typedef struct g2vlm_attn_args {
  const void* q;
  int64_t ldq;
  int64_t q_rows;
  const void* k;
  int64_t ldk;
  const void* v;
  int64_t ldv;
  int64_t kv_rows;
  void* out;
  int64_t ldo;
  int32_t num_q_heads;
  int32_t num_kv_heads;
  int32_t head_dim;
  int32_t causal;
  float softmax_scale;
  int32_t n_items;
  const int32_t* work_items; /* DEVICE int32 [n_items][8] */
  /* 0, or the number of columns of each head that are written to `out`, heads packed at that stride
   * (multiple of 8, <= head_dim): a 96-wide head computed in a 128-wide slot leaves as [q_rows, heads*96]. */
  int32_t out_head_cols;
  /* ABI v3.  lse_out: NULL, or fp32 [q_rows, num_q_heads] receiving ln(sum_k exp(scale * q.k)) of every covered row
   * (-inf for a row without visible keys) so that partial results over disjoint key sets can be merged
   * (g2vlm_attention_merge) — the view-sharded path runs the local keys while the remote K/V are in flight.
   * max_ctas: 0, or an upper bound on the persistent grid (leaves SMs to a concurrent communication kernel). */
  float* lse_out;
  int32_t max_ctas;
} g2vlm_attn_args;

int g2vlm_attention(const g2vlm_attn_args* args, void* stream);

/* fp32 MODE (north_star "fp32 mode <= 1e-4"; ground truth = oracle/restate.py mode="fp32"): the same call with
 * fp32 q / k / v / out (leading dimensions in floats, multiples of 4), head_dim in {16, 32, 64, 96, 128} (no padding of
 * the 96-wide Pi3 heads), FP32-pipe dot products and exp2f softmax.  lse_out / max_ctas / out_head_cols unsupported. */
int g2vlm_attention_f32(const g2vlm_attn_args* args, void* stream);

/* Merge of two attention partials over disjoint key sets (same queries): out = (w_a o_a + w_b o_b) with
 * w_x = exp(lse_x) / (exp(lse_a) + exp(lse_b)), fp32 math on the bf16 partials; rows whose two lse are -inf get
 * zeros.  o_a / o_b / out: bf16 [rows, heads*head_cols] (row strides ld*, in elements; out may alias o_a);
 * lse_a / lse_b: fp32 [rows, heads].  No reference counterpart (the reference has one device): this is the
 * online-softmax combination flash-attn performs internally across key blocks (g2vlm/qwen2vl.py:643-652). */
int g2vlm_attention_merge(const void* o_a, int64_t lda, const float* lse_a, const void* o_b, int64_t ldb,
                          const float* lse_b, void* out, int64_t ldo, int64_t rows, int32_t heads,
                          int32_t head_cols, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound kernels (vectorised, coalesced, warp-shuffle reductions). "Routed" kernels take the
 * expert-permuted row order: rows [0, n_first) use the *_a weights, rows [n_first, rows) the *_b
 * weights (geo expert first, und expert second).
 * ---------------------------------------------------------------------------------------------- */

/* dst[i, :] = src[idx[i], :] (gather) or dst[idx[i], :] = src[i, :] (scatter); idx == NULL is the
 * identity. Rows are `row_bytes` long (multiple of 16; pointers / pitches 16-byte aligned).
 * Replaces the index_select / index_put_ routing copies of modeling/g2vlm/qwen2vl.py:584-606,
 * 657-658, 863-864, 1326-1328, the embedding lookup g2vlm.py:984 and the KV merge :621-638. */
int g2vlm_gather_rows(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes,
                      const int64_t* idx, int64_t n_rows, int64_t row_bytes, int32_t scatter,
                      void* stream);

/* Qwen2RMSNorm (modeling/qwen2vl/modeling_qwen2_vl.py:496-501), routed (g2vlm/qwen2vl.py:862-865,
 * 897-898, 1326-1328): out = w * (x * rsqrt(mean(x^2) + eps)), x fp32 [rows, dim].
 * out_bf16 bit 0: out is bf16 (the `.to(torch.bfloat16)` that follows the norm), else fp32;
 * bit 1 (training forward, bf16 module): the normalised value is rounded to bf16 before the weight multiply
 * (`self.weight * hidden_states.to(input_dtype)` with a bf16 input, :501). */
int g2vlm_rmsnorm_routed(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t out_bf16,
                         const float* w_a, const float* w_b, int64_t rows, int64_t n_first,
                         int32_t dim, float eps, void* stream);

/* nn.LayerNorm(eps, affine) fp32 -> bf16 or fp32 (DINO norm1/norm2/layernorm
 * g2vlm/dinov2_model.py:203-210,285; Pi3 norm1/2/3/norm_y pi3/models/layers/block.py:281-382).
 * If seg_in > 0 rows are grouped in segments of seg_in rows of which the first seg_skip are
 * dropped from the output (final DINO norm + `[:, 1+num_register_tokens:]`, dinov2_model.py:351-354);
 * output rows are then compacted. */
int g2vlm_layernorm(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t out_bf16,
                    const float* w, const float* b, int64_t rows, int32_t dim, float eps,
                    int32_t seg_in, int32_t seg_skip, void* stream);

/* M-RoPE angle table (Qwen2VLRotaryEmbedding.forward modeling_qwen2_vl.py:141-166 + section gather
 * of apply_multimodal_rotary_pos_emb :223-227): position_ids int64 [3, rows] (row stride ld_pos),
 * inv_freq fp32 [half], sections s0+s1+s2 = half -> cos, sin fp32 [rows, half]. */
int g2vlm_mrope_table(const int64_t* position_ids, int64_t ld_pos, const float* inv_freq,
                      float* cos_out, float* sin_out, int64_t rows, int32_t half, int32_t s0,
                      int32_t s1, void* stream);

/* Per-head RMSNorm + M-RoPE, in place on a fused bf16 QKV buffer [rows, ld] whose columns are
 * [q heads | k heads | v heads] x head_dim (=128): routed q_norm/k_norm weights, fp32 math, bf16
 * result (g2vlm/qwen2vl.py:600-619; rotate_half modeling_qwen2_vl.py:170-174, 229-231).
 * round_normed = 1 reproduces the `und` branch where the norm runs on a bf16 tensor (:575-576);
 * round_normed = 2 is the bf16 MODULE of the training forward (:493-512): additionally the weight multiply, cos/sin
 * (`cos.to(x.dtype)`) and both rotation products are rounded to bf16, as bf16 tensor arithmetic does. */
int g2vlm_qknorm_mrope(void* qkv, int64_t ld, int64_t rows, int64_t n_first, int32_t n_q_heads,
                       int32_t n_kv_heads, int32_t head_dim, const float* qw_a, const float* kw_a,
                       const float* qw_b, const float* kw_b, const float* cos_tab,
                       const float* sin_tab, float eps, int32_t round_normed, void* stream);

/* DINO patch embedding input: images fp32 [n, 3, H, W] -> bf16 patches [n*gh*gw, k_pad] in
 * (channel, py, px) order = nn.Conv2d(3, D, 14, 14) as a GEMM
 * (dinov2_with_registers/modeling_dinov2_with_registers.py:62,71); columns >= 3*p*p are zero.
 * mean3 / std3: HOST pointers to 3 floats each, or both NULL. When given, the ImageNet normalisation
 * (x - mean[c]) / std[c] of g2vlm.py:950 is applied on the fly (IEEE fp32 sub + div, bit-identical to
 * the host-side torchvision op), so only the raw [0,1] images have to cross PCIe. */
int g2vlm_im2col_patches(const float* images, void* out, int32_t n, int32_t H, int32_t W,
                         int32_t patch, int32_t k_pad, const float* mean3, const float* std3,
                         void* stream);

/* Dinov2WithRegistersEmbeddings.forward (:147-171): rows per image = [cls+pos0, registers,
 * patch_i + pos_{1+i}]; patch_emb bf16 [n*P, dim], pos fp32 [1+P, dim] (already resampled if the
 * grid differs from the table), out fp32 [n*(1+n_reg+P), dim]. */
int g2vlm_dino_embed(const void* patch_emb, int64_t ld_patch, const float* cls, const float* reg,
                     const float* pos, float* out, int32_t n, int32_t P, int32_t n_reg, int32_t dim,
                     void* stream);

/* RoPE2D (pi3/models/layers/pos_embed.py:112-159) in place on the q and k parts of a fused bf16
 * buffer [rows, ld]: n_heads_total heads of `head_stride` columns starting at column 0, of which the
 * first head_dim columns are real (96 padded to 128). Token position: n = row % tokens_per_view,
 * (y, x) = (n / grid_w, n % grid_w). cos/sin tables fp32 [n_pos, head_dim/4] are built by the host
 * exactly as the reference's cache (:118-128; bf16-rounded angles in bf16 mode, quirk Q2).
 * bf16_ops != 0: every elementwise product / sum is rounded to bf16 as in the reference. */
int g2vlm_rope2d(void* buf, int64_t ld, int64_t rows, int32_t n_heads_total, int32_t head_stride,
                 int32_t head_dim, int32_t tokens_per_view, int32_t grid_w, const float* cos_tab,
                 const float* sin_tab, int32_t bf16_ops, void* stream);

/* apply_rotary_pos_emb_vision (modeling/qwen2vl/modeling_qwen2_vl.py:236-246) in place on the q and k heads of a
 * fused bf16 buffer [rows, ld]: n_heads_total heads of `head_stride` columns from column 0, the first head_dim
 * real; rotate_half over the whole head (pair d, d + head_dim/2); cos / sin fp32 [rows, head_dim/2] per token
 * (built by the host from the 2-D patch positions like rot_pos_emb :1024-1048); fp32 math, one bf16 rounding. */
int g2vlm_rope_vision(void* buf, int64_t ld, int64_t rows, int32_t n_heads_total, int32_t head_stride,
                      int32_t head_dim, const float* cos_tab, const float* sin_tab, void* stream);

/* Pi3LinearPts3d pixel-shuffle (transformer_head.py:77-81) fused with the recon epilogue
 * (g2vlm.py:1203-1205, 1226; homogenize_points pi3/utils/geometry.py:108-113).
 * feat fp32 [n*gh*gw, 3*p*p]. mode 0: out0 = pixel-shuffled (n,H,W,3) (global_points);
 * mode 1: out0 = local_points = (x*e^z, y*e^z, e^z), out1 = points = R*local + t with poses fp32
 * [n,4,4]; mode 2: feat fp32 [n*gh*gw, p*p], out0 = single-channel pixel shuffle (n,H,W,1) (conf_head,
 * g2vlm.py:1208-1212). */
int g2vlm_points_epilogue(const float* feat, int64_t ld_feat, const float* poses, float* out0,
                          float* out1, int32_t n, int32_t H, int32_t W, int32_t patch, int32_t mode,
                          void* stream);

/* out[v, :] = mean over the tokens of view v (AdaptiveAvgPool2d(1), camera_head.py:55). */
int g2vlm_mean_pool(const float* x, int64_t ldx, float* out, int32_t n_views, int32_t tokens,
                    int32_t dim, void* stream);

/* fp32 [rows, k] -> bf16 [rows, 3k] = [hi | hi | lo] (split-bf16 operand for fp32-accurate GEMMs on
 * the bf16 tensor cores; weights are stored as [hi | lo | hi]). */
int g2vlm_split3_f32(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t k,
                     void* stream);

/* Elementwise dtype cast fp32 [rows, cols] -> bf16 (the implicit autocast cast in front of a Linear
 * whose input is the fp32 residual stream, e.g. Pi3 linear_out, transformer_head.py:55). */
int g2vlm_cast_f32_to_bf16(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t rows,
                           int32_t cols, void* stream);

/* fc_t / fc_rot + svd_orthogonalize + 4x4 assembly (camera_head.py:59-93), fp32:
 * feat fp32 [n, dim]; w_t [3, dim], b_t [3], w_r [9, dim], b_r [9] -> poses fp32 [n, 4, 4]. */
int g2vlm_camera_pose(const float* feat, int64_t ldf, const float* w_t, const float* b_t,
                      const float* w_r, const float* b_r, float* poses, int32_t n, int32_t dim,
                      void* stream);

/* Decode-shaped attention (ONE query token, head_dim 128, GQA): flash-decoding split over the keys + merge.
 * Replaces flash_attn_varlen_func at g2vlm/qwen2vl.py:643-652 for the decode steps of generate_text
 * (g2vlm.py:1100-1113; with one query row the bottom-right causal mask hides nothing).
 * q bf16 [num_q_heads*128]; k / v bf16 [kv_len, num_kv_heads*128] (ldk / ldv); out bf16 [num_q_heads*128];
 * workspace fp32, at least ceil(kv_len/160) * num_q_heads * 130 floats.
 * kv_len_dev (DEVICE int32, or NULL): if given, the number of keys is *kv_len_dev + kv_len_extra, read on
 * the device, and kv_len is only an upper bound that sizes the split grid — one captured launch (CUDA graph)
 * then serves every step of a growing cache. */
int g2vlm_attention_decode(const void* q, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           int64_t kv_len, const int32_t* kv_len_dev, int32_t kv_len_extra, void* out,
                           int32_t num_q_heads, int32_t num_kv_heads, int32_t head_dim,
                           float softmax_scale, float* workspace, int64_t workspace_floats, void* stream);

/* In-place append of a step's rows to a cache buffer: dst[base + i, :] = src[i, :], base = *len_dev (DEVICE
 * int32) if len_dev != NULL else static_row. Replaces the per-step re-allocation + index scatter of the whole
 * cache in the reference (g2vlm/qwen2vl.py:621-638). row_bytes and pitches multiples of 16. */
int g2vlm_kv_append(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes,
                    const int32_t* len_dev, int64_t static_row, int64_t rows, int64_t row_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32-mode elementwise kernels (no bf16 rounding anywhere; every Linear of that mode is a split-bf16 GEMM
 * [hi|hi|lo] x [hi|lo|hi] through g2vlm_split3_f32 + g2vlm_gemm_bf16 with the STORE_F32 epilogue).
 * ---------------------------------------------------------------------------------------------- */
/* g2vlm_im2col_patches with an fp32 output [n*gh*gw, k_pad]. */
int g2vlm_im2col_patches_f32(const float* images, float* out, int32_t n, int32_t H, int32_t W, int32_t patch,
                             int32_t k_pad, const float* mean3, const float* std3, void* stream);
/* g2vlm_dino_embed with an fp32 patch embedding [n*P, dim] (ld_patch in floats). */
int g2vlm_dino_embed_f32(const float* patch_emb, int64_t ld_patch, const float* cls, const float* reg,
                         const float* pos, float* out, int32_t n, int32_t P, int32_t n_reg, int32_t dim,
                         void* stream);
/* g2vlm_qknorm_mrope on an fp32 fused QKV buffer (head_dim 128): per-head RMSNorm with the routed weights, then
 * x*cos + rotate_half(x)*sin, all fp32 (g2vlm/qwen2vl.py:600-619 without the bf16 casts). */
int g2vlm_qknorm_mrope_f32(float* qkv, int64_t ld, int64_t rows, int64_t n_first, int32_t n_q_heads,
                           int32_t n_kv_heads, int32_t head_dim, const float* qw_a, const float* kw_a,
                           const float* qw_b, const float* kw_b, const float* cos_tab, const float* sin_tab,
                           float eps, void* stream);
/* g2vlm_rope2d on an fp32 buffer; tables fp32 [n_pos, head_dim/4] of EXACT fp32 angles (pos_embed.py:120-128). */
int g2vlm_rope2d_f32(float* buf, int64_t ld, int64_t rows, int32_t n_heads_total, int32_t head_stride,
                     int32_t head_dim, int32_t tokens_per_view, int32_t grid_w, const float* cos_tab,
                     const float* sin_tab, void* stream);
/* fp32 [rows, k] -> bf16 [rows, 6k] = [m | l | h | m | h | h] with x = h + m + l exactly (three bf16 pieces hold all
 * 24 significand bits); weights are stored as [m | h | l | h | m | h], so ONE bf16 GEMM over 6k sums the six products
 * mm + lh + hl + mh + hm + hh, smallest first (dropped terms <= 2^-24 relative): fp32-accurate Linear layers on the
 * tensor cores when combined with g2vlm_gemm_args.k_chunk_blocks. */
int g2vlm_split6_f32(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t k, void* stream);
/* Qwen2MLP gate (modeling_qwen2_vl.py:519-521) in fp32: out[r, c] = silu(gu[r, c]) * gu[r, inter + c]. */
int g2vlm_swiglu_f32(const float* gate_up, int64_t ld_gu, float* out, int64_t ldo, int64_t rows, int32_t inter,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * One greedy decode step of generate_text (g2vlm.py:1086-1131), `und` expert, batch 1, enqueued as ONE call:
 * embed(cur_token) -> L x [RMSNorm -> qkv GEMV -> q/k-norm + M-RoPE -> append K|V at row *cache_len ->
 * split-K attention over *cache_len + 1 keys -> o_proj GEMV (+residual) -> RMSNorm -> gate/up GEMV + SwiGLU ->
 * down GEMV (+residual)] -> final RMSNorm -> lm_head GEMV -> argmax -> cur_token; position += 1; cache_len += 1.
 * Everything a step depends on (token, position, cache length) lives on the DEVICE, so the host issues the
 * same ~260 launches per token from native code (no per-kernel Python / FFI overhead; graph-capturable).
 * All pointers are device pointers except `layers` and `kv` (HOST arrays of num_layers entries).
 * ---------------------------------------------------------------------------------------------- */
typedef struct g2vlm_und_layer_weights {
  const void* wqkv;  const float* bqkv;  /* bf16 [(nq+2nkv)*hd, H], fp32 bias (bf16-rounded values) */
  const void* wo;                        /* bf16 [H, nq*hd] */
  const void* wgu;                       /* bf16 [2I, H], gate/up interleaved in blocks of 128 rows */
  const void* wdown;                     /* bf16 [H, I] */
  const float* input_norm;  const float* post_norm;  const float* q_norm;  const float* k_norm;
} g2vlm_und_layer_weights;

typedef struct g2vlm_decode_step_args {
  int32_t num_layers, hidden, intermediate, n_q_heads, n_kv_heads, head_dim, vocab;
  float rms_eps;
  int32_t mrope_s0, mrope_s1;
  const g2vlm_und_layer_weights* layers; /* HOST array [num_layers] */
  void* const* kv;                       /* HOST array [num_layers] of device bf16 [kv_capacity, 2*nkv*hd] (K|V) */
  int64_t kv_capacity;                   /* rows available in every cache buffer */
  int64_t kv_bound;                      /* upper bound of *cache_len + 1 during this generation (split sizing) */
  const float* embed;                    /* fp32 [vocab_rows, H] */
  const float* final_norm;               /* fp32 [H] */
  const void* lm_head;                   /* bf16 [vocab, H] */
  const float* inv_freq;                 /* fp32 [head_dim/2] */
  int64_t* cur_token;                    /* in: token to process, out: next token */
  int64_t* position;                     /* int64 [3] (t,h,w rope ids), incremented */
  int32_t* cache_len;                    /* int32 [1], incremented */
  /* workspaces */
  float* x;        /* fp32 [H] residual stream */
  void* h;         /* bf16 [max(H, I)] */
  void* qkv;       /* bf16 [(nq+2nkv)*hd] */
  void* attn;      /* bf16 [nq*hd] */
  void* act;       /* bf16 [I] */
  float* y;        /* fp32 [H] */
  float* cos_sin;  /* fp32 [head_dim] (cos | sin) */
  float* attn_ws;  int64_t attn_ws_floats;
  void* logits;    /* bf16 [round_up(vocab, 8)] */
  /* ABI v4.  Non-null: the step runs as ONE persistent cooperative kernel (one CTA per SM, grid barriers between the
   * phases of a layer, L2 prefetch of the next phases' weights across them; csrc/decode_fused.cu) instead of ~280
   * launches.  Device buffer of >= g2vlm_und_decode_workspace_bytes() bytes, 16-byte aligned, ZEROED ONCE by the caller
   * when it is allocated (it carries the barrier counter from step to step) and owned by ONE generation at a time. */
  void* fused_ws;  int64_t fused_ws_bytes;
  /* ABI v4.  Non-zero: the step leaves *cur_token alone (the caller samples the next token from `logits`: the do_sample tail of
   * generate_text, g2vlm.py:1122-1124); position and cache length advance as usual. */
  int32_t keep_token;
} g2vlm_decode_step_args;

int g2vlm_und_decode_step(const g2vlm_decode_step_args* args, void* stream);
/* Size of `fused_ws` for this head geometry on the current device. */
int64_t g2vlm_und_decode_workspace_bytes(int32_t n_q_heads, int32_t n_kv_heads);

/* Qwen2VLModel.forward_inference(mode="und") for T >= 2 rows on top of an append-style cache, enqueued as ONE call: text
 * prefill (causal; g2vlm.py:701-733), the ViT step (non-causal; g2vlm.py:735-790) — the und branches of
 * g2vlm/qwen2vl.py:570-576, 859-860, 891-893, 1322-1323.  M-RoPE table of the rows' positions, then per layer RMSNorm -> qkv
 * GEMM (+bias) -> q/k-norm on the bf16 tensor + M-RoPE -> append K|V at rows [cache_len, cache_len + rows) -> attention over
 * cache_len + rows keys -> o_proj (+residual) -> RMSNorm -> gate/up + SwiGLU -> down (+residual); final RMSNorm -> y.
 * The same per-op entry points in the same order as the host mirror issues them one by one (~9 x num_layers FFI calls).
 * `layers` and `kv` are HOST arrays like in g2vlm_decode_step_args; everything else is device memory. */
typedef struct g2vlm_und_prefill_args {
  int32_t num_layers, hidden, intermediate, n_q_heads, n_kv_heads, head_dim;
  float rms_eps;
  int32_t mrope_s0, mrope_s1;
  const g2vlm_und_layer_weights* layers;
  void* const* kv;
  int64_t kv_capacity;
  int64_t cache_len;                 /* rows already in the cache */
  int32_t rows, causal;
  const float* final_norm;           /* fp32 [H] */
  const float* inv_freq;             /* fp32 [head_dim/2] */
  const int64_t* position_ids;       /* int64 [3, rows] */
  const int32_t* work;  int32_t n_items;   /* attention work table (g2vlm_attention) over q rows [0, rows), keys [0, cache_len + rows) */
  float* x;                          /* fp32 [rows, H]: in = the embedded rows, updated in place (residual stream) */
  float* y;                          /* fp32 [rows, H]: out = `norm`-ed hidden states */
  void* h;  void* qkv;  void* attn;  void* act;   /* bf16 workspaces [rows, H], [rows, (nq+2nkv)*hd], [rows, nq*hd], [rows, I] */
  float* cos;  float* sin;           /* fp32 workspaces [rows, head_dim/2] */
} g2vlm_und_prefill_args;

int g2vlm_und_prefill(const g2vlm_und_prefill_args* args, void* stream);
/* The same entry point under the name SURVEY.md §8(b) gives it. */
int g2vlm_mot_prefill_und(const g2vlm_und_prefill_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * View-sharded scene (one long scene split by view over the GPUs of a node), for a host that owns a raw NCCL
 * communicator (the Python mirror uses torch.distributed: symmetric memory + copy engines, or the same NCCL exchange).
 * The K|V exchange of ONE MoT layer (the step the reference's dataflow g2vlm/qwen2vl.py:621-652 needs once its keys live on
 * several GPUs): this rank's rows kv_send bf16 [rank_rows[rank], kvw] go to every peer, and peer j's rows land in kv_remote
 * bf16 [sum of the other ranks' rows, kvw] at row offset sum(rank_rows[i] for i < j, i != rank) — ONE NCCL group of
 * point-to-point sends / receives enqueued on `stream` (a side stream: the caller runs g2vlm_attention over its local keys
 * with `lse_out` and a `max_ctas` bound meanwhile, then over kv_remote, and combines the two with g2vlm_attention_merge).
 * nccl_comm: an ncclComm_t of `world` ranks created by the caller with the NCCL library of the process (resolved with
 * dlopen("libnccl.so.2"): the library is not a link-time dependency).  rank_rows: HOST int64 [world].
 * ---------------------------------------------------------------------------------------------- */
int g2vlm_sp_kv_exchange(void* nccl_comm, int32_t rank, int32_t world, const int64_t* rank_rows, int64_t kvw,
                         const void* kv_send, void* kv_remote, void* stream);

/* Greedy token selection of generate_text (g2vlm.py:1122-1126, `torch.argmax(pred_logits, dim=-1)`):
 * logits bf16 [rows, vocab] (leading dim ld) -> int64 [rows]; ties resolve to the lowest index. */
int g2vlm_argmax_bf16(const void* logits, int64_t ld, int64_t rows, int32_t vocab, int64_t* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stack-level entry points: an opaque per-model context and ONE call per stage of G2VLM.recon.
 * The host mirror (`G2VLMFast`) packs the reference's state_dict tensors into the layouts below and registers them;
 * every stage call then enqueues its whole kernel sequence from native code through the per-op entry points above
 * (same kernels, order and rounding points as the per-op path: bit-identical results).  Ownership: the context owns no
 * device memory — weights and the workspace belong to the caller.  Threading: a context may be used by one host thread
 * at a time; distinct contexts are independent.  After g2vlm_recon_plan the stage calls only launch kernels (no
 * allocation, no host<->device copy, no synchronisation): CUDA-graph capturable.
 * ---------------------------------------------------------------------------------------------- */
typedef struct g2vlm_dims {
  /* Qwen2-VL MoT language model (modeling/g2vlm/qwen2vl.py:50-234) */
  int32_t hidden, layers, heads, kv_heads, intermediate;
  float rms_eps;
  int32_t mrope_s0, mrope_s1; /* mrope_section[0..1] (16, 24; hard-coded in the reference, modeling_qwen2_vl.py:562-566) */
  /* DINOv2-with-registers encoder */
  int32_t dino_hidden, dino_layers, dino_heads, dino_mlp_ratio, dino_patch, dino_registers;
  float dino_ln_eps;
  /* Pi3 decoders / heads (g2vlm.py:162-226) */
  int32_t dec_depth, dec_heads, dec_mlp_ratio, point_dim, camera_dim;
  int32_t train_conf; /* 1: conf_decoder + conf_head exist (train_conf_pi3) */
} g2vlm_dims;

typedef struct g2vlm_ctx g2vlm_ctx;

#define G2VLM_DTYPE_F32 0
#define G2VLM_DTYPE_BF16 1

int g2vlm_ctx_create(const g2vlm_dims* dims, g2vlm_ctx** ctx);
int g2vlm_ctx_destroy(g2vlm_ctx* ctx);

/* Registers ONE packed weight tensor (device pointer, 16-byte aligned, row-major [rows, cols]) under its slot name.
 * Slots (geo expert first wherever two experts are stacked; "bias" tensors hold bf16-rounded values as fp32):
 *   embed f32 [vocab_rows,H] | inv_freq f32 [64] | norm_geo, norm_und f32 [H] | dino2llm.w bf16 [H,D] | dino2llm.b f32 [H]
 *   mot.{i}.wqkv bf16 [2*(nq+2nkv)*128, H]   mot.{i}.bqkv f32   mot.{i}.wo bf16 [2H, nq*128]
 *   mot.{i}.wgu bf16 [2*2I, H] (gate/up interleaved in blocks of 128 rows)   mot.{i}.wdown bf16 [2H, I]
 *   mot.{i}.{input_layernorm,post_attention_layernorm,q_norm,k_norm}_{geo,und} f32   mot.{i}.ls1, ls2 f32 [H]
 *   dino.wpatch bf16 [D, kpad] | dino.bpatch | dino.cls f32 [D] | dino.reg f32 [n_reg, D] | dino.lnw, dino.lnb
 *   dino.{i}.wqkv bf16 [3*heads*hp, D] (heads zero-padded to hp = 64 | 128)  bqkv  wdense bf16 [D, heads*hp]  bdense
 *   dino.{i}.wfc1, bfc1, wfc2, bfc2, norm1w, norm1b, norm2w, norm2b, ls1, ls2
 *   dec.{point_decoder|camera_decoder|global_points_decoder|conf_decoder}.{b}.{wqkv,bqkv,wproj,bproj,wfc1,bfc1,wfc2,
 *        bfc2,norm1w,norm1b,norm2w,norm2b} (+ norm3w, norm3b, norm_yw, norm_yb, wcq, bcq, wckv, bckv, wcproj, bcproj for the
 *        cross-attention decoder)   dec.{name}.wout, bout
 *   head.{point_head|global_point_head|conf_head}.whi, wlo bf16 [3*14*14 | 14*14, point_dim] (w = hi + lo), .b f32
 *   cam.r{0,1}{1,2,3}w, cam.m{0,2}w bf16 [C, 3C] = [hi|lo|hi]; cam.*b f32; cam.fc_tw f32 [3,C], fc_tb, fc_rotw [9,C], fc_rotb
 * A missing slot makes the stage call that needs it fail with G2VLM_ERR_INVALID and a message naming it. */
int g2vlm_load_weights(g2vlm_ctx* ctx, const char* name, const void* ptr, int32_t dtype, int64_t rows, int64_t cols);

/* Bytes of caller-owned workspace the three stages need for one scene geometry (-1: bad arguments). */
int64_t g2vlm_workspace_bytes(const g2vlm_ctx* ctx, int32_t n_views, int32_t H, int32_t W, int32_t n_prompt);

/* Once per geometry (NOT capturable: host->device copies): writes the attention work tables into the workspace and
 * clears the regions that must be zero where no kernel writes.  dino_seqlens: HOST int32 [n_views], the caller's
 * cu_seqlens lengths (the reference passes patch counts, quirk Q1 g2vlm.py:988-990).  n_prompt: rows of the und
 * prompt prefill fused into the geo step (K0 = 7 in recon; 0 = none). */
int g2vlm_recon_plan(g2vlm_ctx* ctx, int32_t n_views, int32_t H, int32_t W, int32_t n_prompt,
                     const int32_t* dino_seqlens, void* workspace, int64_t workspace_bytes, void* stream);

/* Where a named intermediate of the planned geometry lives inside the workspace (debugging / stage-by-stage parity):
 * e.g. "dino.tokens" bf16 [n*P, D], "mot.x" fp32 [T+n_prompt, H] (expert-permuted rows), "rec.point_hidden" bf16,
 * "rec.camera_hidden" fp32, "rec.global_hidden" bf16, "dec.x" fp32 (residual stream of the last decoder run). */
int g2vlm_workspace_region(const g2vlm_ctx* ctx, const char* name, int64_t* offset, int64_t* bytes);

/* Dinov2WithRegistersModel.forward (g2vlm/dinov2_model.py:301-356): images fp32 [n,3,H,W] (normalize != 0: raw [0,1]
 * views, ImageNet-normalised on the fly) -> bf16 tokens [n*P, D] inside the workspace (*tokens_out).  pos_embed: fp32
 * [1+P, D], the position table already resampled to this grid (interpolate_pos_encoding, :93-145). */
int g2vlm_dino_forward(g2vlm_ctx* ctx, const float* images, int32_t n_views, int32_t H, int32_t W, int32_t normalize,
                       const float* pos_embed, void* workspace, void** tokens_out, void* stream);

/* forward_cache_update_dino after the encoder (g2vlm.py:984-1039) = dino2llm + scatter + Qwen2VLModel.forward_inference
 * (mode="geo", g2vlm/qwen2vl.py:1267-1337) with the und prompt prefill (g2vlm.py:701-733) riding along as n_prompt extra
 * rows.  All index tensors are DEVICE int64 built by the host exactly as the reference does (prepare_dino_images_pi3):
 * packed_text_ids / packed_text_indexes [2n], packed_dino_token_indexes [n*P], packed_position_ids [3, T] contiguous,
 * prompt_ids [n_prompt], prompt_position_ids [3, n_prompt] contiguous.  last_hidden: fp32 [T, H] (caller buffer).
 * attention_events: NULL, or HOST array of 2*layers cudaEvent_t recorded before / after every shared-attention launch
 * (bench.py's roofline timing). */
int g2vlm_mot_forward_geo(g2vlm_ctx* ctx, const void* dino_tokens, const int64_t* packed_text_ids,
                          const int64_t* packed_text_indexes, const int64_t* packed_dino_token_indexes,
                          const int64_t* packed_position_ids, const int64_t* prompt_ids,
                          const int64_t* prompt_position_ids, void* workspace, float* last_hidden,
                          void* const* attention_events, void* stream);

/* G2VLM.reconstruct (g2vlm.py:1143-1238): 3 (4 with conf) Pi3 decoders, camera head, point heads, epilogue.
 * rope_cos / rope_sin: fp32 [max(gh,gw), head_dim/4] RoPE2D tables built like the reference's cache (pos_embed.py:118-128).
 * Outputs (caller buffers): points, local_points, global_points fp32 [n,H,W,3]; camera_poses fp32 [n,4,4]; conf fp32
 * [n,H,W,1] or NULL. */
int g2vlm_recon_heads(g2vlm_ctx* ctx, const float* last_hidden, const int64_t* packed_dino_token_indexes,
                      const float* rope_cos, const float* rope_sin, void* workspace, float* points, float* local_points,
                      float* global_points, float* camera_poses, float* conf, void* stream);

/* Kernels launched so far by the stage drivers of this process (bench.py's gpu_launches). */
int64_t g2vlm_driver_launches(void);

/* ------------------------------------------------------------------------------------------------
 * Output side of the path ("next" row f3): device-side replacement of the numpy stage of
 * save_ply_visualization (g2vlm_utils.py:84-149): drop points with a NaN/Inf coordinate (:126-143; filter_nonfinite = 0
 * keeps every point, the reference's filter_nan=False),
 * keep the original order, and pack binary-little-endian PLY vertex records
 *   double x, y, z ; uchar red, green, blue          (27 bytes, the layout Open3D writes)
 * points fp32 [n_views*H*W, 3] (= pred["points"]), images fp32 [n_views, 3, H, W] in [0,1]
 * (= pred["images"]; colour of point (v,y,x) is images[v,:,y,x], :121). out must hold n*27 bytes;
 * block_counts is an int32 workspace of ceil(n/1024)+1 entries; *n_valid (device int64) receives the
 * number of records written. colour byte = round(clamp(c, 0, 1) * 255), Open3D's utility::ColorToUint8.
 * ---------------------------------------------------------------------------------------------- */
int g2vlm_ply_pack(const float* points, const float* images, int32_t n_views, int32_t H, int32_t W, void* out,
                   int32_t* block_counts, int64_t* n_valid, int32_t filter_nonfinite, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input side of the path ("next" row f3): device-side replacement of the per-view
 *   img_pil.resize((TARGET_W, TARGET_H), Image.Resampling.LANCZOS) ; ToTensor()
 * of load_images (data/transforms_vggt.py:437-441), bit-exact with Pillow's 8-bit two-pass resampler
 * (horizontal pass into an 8-bit intermediate, then vertical; per output sample
 *   clip8((2^21 + sum_k coef[k] * pixel[k]) >> 22),  coefficients = normalised Lanczos-3 weights in 22-bit fixed
 * point).  The coefficient tables depend only on (input size, output size) and are built on the host
 * (g2vlm_b200.host_prep.lanczos_tables restates Pillow's precompute_coeffs / normalize_coeffs_8bpc).
 * src uint8 [H, W, 3] interleaved RGB (row pitch src_pitch bytes);
 * hbounds int32 [out_w][2] = {first input column, taps}, hcoef int32 [out_w][hk]; vbounds/vcoef likewise for rows;
 * pass NULL tables for an axis whose size does not change (Pillow skips that pass).
 * tmp uint8 [H, out_w, 3] workspace (unused when hbounds is NULL);
 * out_u8 uint8 [out_h, out_w, 3] and/or out_f32 fp32 [3, out_h, out_w] = value / 255 (ToTensor); either may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int g2vlm_resize_lanczos_u8(const uint8_t* src, int32_t H, int32_t W, int64_t src_pitch, const int32_t* hbounds,
                            const int32_t* hcoef, int32_t hk, const int32_t* vbounds, const int32_t* vcoef,
                            int32_t vk, uint8_t* tmp, int32_t out_h, int32_t out_w, uint8_t* out_u8,
                            float* out_f32, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* G2VLM_B200_H_ */
