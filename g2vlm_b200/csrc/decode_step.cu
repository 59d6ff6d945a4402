// g2vlm_b200 — native driver of one greedy decode step. A decode step is ~260 tiny
// bandwidth-bound launches; issued one by one from Python they cost 15 us of interpreter + FFI each (4 ms per
// token, GPU 25 % busy). This translation unit issues the whole step from C++ through the same C-ABI entry
// points the host mirror uses, so the step is bounded by the GPU again.
#include "common.cuh"

namespace g2 {

__global__ void advance_step_kernel(long long* position, int* cache_len) {
  if (threadIdx.x < 3) position[threadIdx.x] += 1;
  if (threadIdx.x == 3) cache_len[0] += 1;
}

}  // namespace g2

#define G2_TRY(expr)             \
  do {                           \
    int _rc = (expr);            \
    if (_rc != G2VLM_OK) return _rc; \
  } while (0)

extern "C" int g2vlm_und_decode_step(const g2vlm_decode_step_args* a, void* stream) {
  using namespace g2;
  G2_REQUIRE(a && a->layers && a->kv, "decode_step: null args");
  G2_REQUIRE(a->head_dim == 128 && a->num_layers > 0, "decode_step: head_dim must be 128");
  G2_REQUIRE(a->kv_bound > 0 && a->kv_bound <= a->kv_capacity, "decode_step: kv_bound exceeds the cache capacity");
  if (a->fused_ws != nullptr) return launch_decode_fused(a, (cudaStream_t)stream);
  const int H = a->hidden, I = a->intermediate, nq = a->n_q_heads, nkv = a->n_kv_heads, hd = a->head_dim;
  const int qkv_w = (nq + 2 * nkv) * hd, kvw = 2 * nkv * hd;
  const float scale = static_cast<float>(1.0 / sqrt(static_cast<double>(hd)));  // same rounding as the host mirror (1/math.sqrt in double, then float)
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(a->qkv);

  // embedding row of the current token (device index) and the M-RoPE angles of the current position
  G2_TRY(g2vlm_gather_rows(a->embed, (int64_t)H * 4, a->x, (int64_t)H * 4, a->cur_token, 1, (int64_t)H * 4, 0, stream));
  G2_TRY(g2vlm_mrope_table(a->position, 1, a->inv_freq, a->cos_sin, a->cos_sin + hd / 2, 1, hd / 2, a->mrope_s0,
                           a->mrope_s1, stream));

  g2vlm_gemm_args g;
  auto gemv = [&](const void* x, int K, const void* w, int N, int epilogue, uint32_t flags, void* out,
                  const float* bias) -> int {
    memset(&g, 0, sizeof(g));
    g.A = x; g.lda = K; g.a_rows = 1;
    g.B = w; g.ldb = K; g.N = N; g.K = K;
    g.n_groups = 1; g.group_row0[0] = 0; g.group_rows[0] = 1;
    g.epilogue = epilogue; g.flags = flags;
    g.out = out; g.ldo = (epilogue == G2VLM_EPI_SWIGLU_BF16) ? N / 2 : N;
    g.bias = bias;
    return g2vlm_gemm_bf16(&g, stream);
  };

  for (int l = 0; l < a->num_layers; ++l) {
    const g2vlm_und_layer_weights& w = a->layers[l];
    void* kvbuf = a->kv[l];
    G2_TRY(g2vlm_rmsnorm_routed(a->x, H, a->h, H, 1, w.input_norm, w.input_norm, 1, 0, H, a->rms_eps, stream));
    G2_TRY(gemv(a->h, H, w.wqkv, qkv_w, G2VLM_EPI_STORE_BF16, 0, a->qkv, w.bqkv));
    G2_TRY(g2vlm_qknorm_mrope(a->qkv, qkv_w, 1, 0, nq, nkv, hd, w.q_norm, w.k_norm, w.q_norm, w.k_norm, a->cos_sin,
                              a->cos_sin + hd / 2, a->rms_eps, 1, stream));
    G2_TRY(g2vlm_kv_append(qkv + nq * hd, (int64_t)qkv_w * 2, kvbuf, (int64_t)kvw * 2, a->cache_len, 0, 1,
                           (int64_t)kvw * 2, stream));
    const __nv_bfloat16* kb = reinterpret_cast<const __nv_bfloat16*>(kvbuf);
    G2_TRY(g2vlm_attention_decode(a->qkv, kb, kvw, kb + nkv * hd, kvw, a->kv_bound, a->cache_len, 1, a->attn, nq, nkv,
                                  hd, scale, a->attn_ws, a->attn_ws_floats, stream));
    G2_TRY(gemv(a->attn, nq * hd, w.wo, H, G2VLM_EPI_RESID_F32, 0, a->x, nullptr));
    G2_TRY(g2vlm_rmsnorm_routed(a->x, H, a->h, H, 1, w.post_norm, w.post_norm, 1, 0, H, a->rms_eps, stream));
    G2_TRY(gemv(a->h, H, w.wgu, 2 * I, G2VLM_EPI_SWIGLU_BF16, 0, a->act, nullptr));
    G2_TRY(gemv(a->act, I, w.wdown, H, G2VLM_EPI_RESID_F32, 0, a->x, nullptr));
  }
  G2_TRY(g2vlm_rmsnorm_routed(a->x, H, a->y, H, 0, a->final_norm, a->final_norm, 1, 0, H, a->rms_eps, stream));
  G2_TRY(g2vlm_cast_f32_to_bf16(a->y, H, a->h, H, 1, H, stream));
  const int vpad = (a->vocab + 7) / 8 * 8;
  memset(&g, 0, sizeof(g));
  g.A = a->h; g.lda = H; g.a_rows = 1; g.B = a->lm_head; g.ldb = H; g.N = a->vocab; g.K = H; g.n_groups = 1;
  g.group_rows[0] = 1; g.epilogue = G2VLM_EPI_STORE_BF16; g.out = a->logits; g.ldo = vpad;
  G2_TRY(g2vlm_gemm_bf16(&g, stream));
  if (!a->keep_token) G2_TRY(g2vlm_argmax_bf16(a->logits, vpad, 1, a->vocab, a->cur_token, stream));
  advance_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(a->position), a->cache_len);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

// Text prefill / ViT step of the und expert, T >= 2 rows: the per-op sequence of G2VLMFast._und_forward issued natively.
extern "C" int g2vlm_und_prefill(const g2vlm_und_prefill_args* a, void* stream) {
  using namespace g2;
  G2_REQUIRE(a && a->layers && a->kv && a->x && a->y && a->h && a->qkv && a->attn && a->act && a->cos && a->sin &&
                 a->position_ids && a->work, "und_prefill: null argument");
  G2_REQUIRE(a->head_dim == 128 && a->num_layers > 0 && a->rows >= 2, "und_prefill: head_dim must be 128, rows >= 2");
  G2_REQUIRE(a->cache_len >= 0 && a->cache_len + a->rows <= a->kv_capacity, "und_prefill: the rows do not fit the cache");
  const int H = a->hidden, I = a->intermediate, nq = a->n_q_heads, nkv = a->n_kv_heads, hd = a->head_dim, T = a->rows;
  const int qkv_w = (nq + 2 * nkv) * hd, kvw = 2 * nkv * hd;
  const float scale = static_cast<float>(1.0 / sqrt(static_cast<double>(hd)));
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(a->qkv);
  G2_TRY(g2vlm_mrope_table(a->position_ids, T, a->inv_freq, a->cos, a->sin, T, hd / 2, a->mrope_s0, a->mrope_s1, stream));
  g2vlm_gemm_args g;
  auto gemm = [&](const void* x, int K, const void* w, int N, int epilogue, void* out, const float* bias) -> int {
    memset(&g, 0, sizeof(g));
    g.A = x; g.lda = K; g.a_rows = T;
    g.B = w; g.ldb = K; g.N = N; g.K = K;
    g.n_groups = 1; g.group_row0[0] = 0; g.group_rows[0] = T;
    g.epilogue = epilogue;
    g.flags = epilogue == G2VLM_EPI_RESID_F32 ? G2VLM_GEMM_ROUND_AFTER_SCALE : 0;
    g.out = out; g.ldo = (epilogue == G2VLM_EPI_SWIGLU_BF16) ? N / 2 : N;
    g.bias = bias;
    return g2vlm_gemm_bf16(&g, stream);
  };
  for (int l = 0; l < a->num_layers; ++l) {
    const g2vlm_und_layer_weights& w = a->layers[l];
    __nv_bfloat16* kvbuf = reinterpret_cast<__nv_bfloat16*>(a->kv[l]);
    G2_TRY(g2vlm_rmsnorm_routed(a->x, H, a->h, H, 1, w.input_norm, w.input_norm, T, 0, H, a->rms_eps, stream));
    G2_TRY(gemm(a->h, H, w.wqkv, qkv_w, G2VLM_EPI_STORE_BF16, a->qkv, w.bqkv));
    G2_TRY(g2vlm_qknorm_mrope(a->qkv, qkv_w, T, 0, nq, nkv, hd, w.q_norm, w.k_norm, w.q_norm, w.k_norm, a->cos, a->sin,
                              a->rms_eps, 1, stream));
    G2_TRY(g2vlm_kv_append(qkv + nq * hd, (int64_t)qkv_w * 2, kvbuf, (int64_t)kvw * 2, nullptr, a->cache_len, T,
                           (int64_t)kvw * 2, stream));
    g2vlm_attn_args at;
    memset(&at, 0, sizeof(at));
    at.q = a->qkv; at.ldq = qkv_w; at.q_rows = T;
    at.k = kvbuf; at.ldk = kvw; at.v = kvbuf + nkv * hd; at.ldv = kvw; at.kv_rows = a->cache_len + T;
    at.out = a->attn; at.ldo = nq * hd;
    at.num_q_heads = nq; at.num_kv_heads = nkv; at.head_dim = hd;
    at.causal = a->causal; at.softmax_scale = scale;
    at.n_items = a->n_items; at.work_items = a->work;
    G2_TRY(g2vlm_attention(&at, stream));
    G2_TRY(gemm(a->attn, nq * hd, w.wo, H, G2VLM_EPI_RESID_F32, a->x, nullptr));
    G2_TRY(g2vlm_rmsnorm_routed(a->x, H, a->h, H, 1, w.post_norm, w.post_norm, T, 0, H, a->rms_eps, stream));
    G2_TRY(gemm(a->h, H, w.wgu, 2 * I, G2VLM_EPI_SWIGLU_BF16, a->act, nullptr));
    G2_TRY(gemm(a->act, I, w.wdown, H, G2VLM_EPI_RESID_F32, a->x, nullptr));
  }
  G2_TRY(g2vlm_rmsnorm_routed(a->x, H, a->y, H, 0, a->final_norm, a->final_norm, T, 0, H, a->rms_eps, stream));
  return G2VLM_OK;
}

extern "C" int g2vlm_mot_prefill_und(const g2vlm_und_prefill_args* a, void* stream) { return g2vlm_und_prefill(a, stream); }
