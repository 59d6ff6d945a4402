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

#define G2VLM_ABI_VERSION 1

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
 * rounding points (autocast returns bf16 from every Linear).
 * ---------------------------------------------------------------------------------------------- */
#define G2VLM_EPI_STORE_BF16 0  /* out_bf16 = bf16(acc + bias) [; gelu] */
#define G2VLM_EPI_SWIGLU_BF16 1 /* B rows interleave gate/up in blocks of 128; out[:, N/2] */
#define G2VLM_EPI_RESID_F32 2   /* out_f32 += [bf16](scale * bf16(acc + bias)) */
#define G2VLM_EPI_STORE_F32 3   /* out_f32 = [resid +] [relu] [bf16] (acc + bias) [+ out_f32] */

#define G2VLM_GEMM_GELU 1u              /* STORE_BF16: exact-erf GELU on the bf16-rounded value */
#define G2VLM_GEMM_ROUND_AFTER_SCALE 2u /* RESID_F32: round scale*x to bf16 (MoT ls1/ls2) */
#define G2VLM_GEMM_RELU 4u              /* STORE_F32 */
#define G2VLM_GEMM_ACCUMULATE 8u        /* STORE_F32: out += value (second pass of split-bf16) */
#define G2VLM_GEMM_ROUND_BF16 16u       /* STORE_F32: round (acc+bias) to bf16 before storing */

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
 *   {q_tile_begin, q_seg_begin, q_seg_end, k_begin, k_end, 0, 0, 0}
 * one item per <= 256 consecutive query rows [q_tile_begin, min(q_tile_begin+256, q_seg_end)) of a
 * segment whose queries are rows [q_seg_begin, q_seg_end) and whose keys are rows [k_begin, k_end).
 * Query rows covered by no item are NOT written (flash-attn leaves them uninitialised, SURVEY.md
 * quirk Q1; the caller defines them, the host mirror zero-fills). causal = bottom-right aligned
 * mask as in flash-attn. K/V rows in [k_end, round_up(k_end,128)) that lie inside kv_rows must hold
 * finite values. head_dim in {64, 128} (the 96-wide Pi3 heads are zero-padded to 128 by the caller).
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
} g2vlm_attn_args;

int g2vlm_attention(const g2vlm_attn_args* args, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* G2VLM_B200_H_ */
