// g2vlm_b200 — stack-level entry points of the recon path: an opaque per-model context plus ONE call
// per stage, so that a host in any language drives the whole forward pass with three calls instead of ~630:
//
//   g2vlm_ctx_create / g2vlm_ctx_destroy     model dimensions, the table of packed weights, the plan of one geometry
//   g2vlm_load_weights                       registers one packed weight tensor (layouts: include/g2vlm_b200.h)
//   g2vlm_workspace_bytes / g2vlm_recon_plan caller-owned workspace: size, and the per-geometry tables written into it
//   g2vlm_dino_forward                       Dinov2WithRegistersModel.forward            g2vlm/dinov2_model.py:301-356
//   g2vlm_mot_forward_geo                    dino2llm + scatter + the 28-layer MoT stack g2vlm.py:984-1039, g2vlm/qwen2vl.py:1267-1337
//                                            with the 7-token und prefill (g2vlm.py:701-733) riding along as extra rows
//   g2vlm_recon_heads                        G2VLM.reconstruct                            g2vlm.py:1143-1238
//
// The drivers issue the same per-op entry points the host mirror uses (same kernels, same order, same rounding points:
// results are bit-identical to the per-op path, tests/test_native_stack_gpu.py).  After g2vlm_recon_plan they launch
// kernels only — no allocation, no host<->device copy, no synchronisation — so a stage is CUDA-graph capturable.
// The reference has no native layer at all; nothing here is derived from its code.
#include <atomic>
#include <string>
#include <unordered_map>
#include <vector>

#include <math.h>
#include <string.h>

#include "common.cuh"

namespace g2 {

std::atomic<long long> g_driver_launches{0};

struct Slot {
  const void* ptr;
  int dtype;
  long long rows, cols;
};

struct Region {
  long long off = 0, bytes = 0;
};

// Workspace layout of one geometry: every buffer of the three stages has its own region (no aliasing, so the regions
// that must stay zero where no kernel writes — DINO attention rows of no segment, pad columns of the 128-wide Pi3 head
// slots — are cleared once by g2vlm_recon_plan).
struct Layout {
  int n_views = 0, H = 0, W = 0, Kp = 0;
  int gh = 0, gw = 0, P = 0, S = 0, T = 0, R = 0, n_geo = 0, n_und = 0;
  int dino_hp = 0, dec_hp = 0, dec_compact = 0;
  long long total = 0;
  std::unordered_map<std::string, Region> r;
  int n_dino_items = 0, n_mot_items = 0, n_dec_items = 0, n_cross_items = 0;
};

}  // namespace g2

struct g2vlm_ctx {
  g2vlm_dims d;
  int device;
  std::unordered_map<std::string, g2::Slot> w;
  g2::Layout plan;
  bool planned = false;
};

namespace g2 {

static int pad_dim(int hd) { return hd <= 64 ? 64 : 128; }

static Region& add(Layout& L, const char* name, long long bytes) {
  Region& reg = L.r[name];
  reg.off = L.total;
  reg.bytes = bytes;
  L.total += (bytes + 255) / 256 * 256;
  return reg;
}

static int make_layout(const g2vlm_dims& d, int n_views, int H, int W, int Kp, Layout* out) {
  G2_REQUIRE(n_views > 0 && H > 0 && W > 0 && Kp >= 0, "plan: bad geometry");
  G2_REQUIRE(d.dino_patch > 0 && H % d.dino_patch == 0 && W % d.dino_patch == 0, "plan: H, W must be multiples of the patch");
  Layout L;
  L.n_views = n_views; L.H = H; L.W = W; L.Kp = Kp;
  L.gh = H / d.dino_patch; L.gw = W / d.dino_patch;
  L.P = L.gh * L.gw;
  L.S = L.P + 1 + d.dino_registers;
  L.T = n_views * (L.P + 2);
  L.R = L.T + Kp;
  L.n_geo = n_views * L.P;
  L.n_und = 2 * n_views;
  const int dino_hd = d.dino_hidden / d.dino_heads, dec_hd = d.hidden / d.dec_heads;
  L.dino_hp = pad_dim(dino_hd);
  L.dec_hp = pad_dim(dec_hd);
  L.dec_compact = (dec_hd != L.dec_hp && dec_hd % 32 == 0) ? 1 : 0;
  const long long N = n_views, P = L.P, S = L.S, T = L.T, R = L.R, D = d.dino_hidden, Hd = d.hidden, I = d.intermediate;
  const long long kpad = (3LL * d.dino_patch * d.dino_patch + 63) / 64 * 64;
  const int hd = d.hidden / d.heads;
  const long long qkvw = (long long)(d.heads + 2 * d.kv_heads) * hd;
  // tables
  add(L, "tab.dino_work", (N + (N * P) / 256 + 8) * 32);
  add(L, "tab.mot_work", (R / 256 + 8) * 32);
  add(L, "tab.dec_work", (N * (P / 256 + 2)) * 32);
  add(L, "tab.cross_work", (N * (P / 256 + 2)) * 32);
  add(L, "tab.perm", T * 8);
  // DINO
  add(L, "dino.patches", N * P * kpad * 2);
  add(L, "dino.emb", N * P * D * 2);
  add(L, "dino.x", N * S * D * 4);
  add(L, "dino.h", N * S * D * 2);
  add(L, "dino.qkv", N * S * 3 * d.dino_heads * L.dino_hp * 2);
  add(L, "dino.attn", N * S * d.dino_heads * L.dino_hp * 2);          // zero where no segment covers a row (Q1)
  add(L, "dino.mid", N * S * D * d.dino_mlp_ratio * 2);
  add(L, "dino.tokens", N * P * D * 2);
  // MoT
  add(L, "mot.geo_emb", N * P * Hd * 4);
  add(L, "mot.packed", T * Hd * 4);
  add(L, "mot.txt", (long long)L.n_und * Hd * 4);
  add(L, "mot.x", R * Hd * 4);
  add(L, "mot.cos_p", T * (hd / 2) * 4);
  add(L, "mot.sin_p", T * (hd / 2) * 4);
  add(L, "mot.cos", R * (hd / 2) * 4);
  add(L, "mot.sin", R * (hd / 2) * 4);
  add(L, "mot.qkv", R * qkvw * 2);
  add(L, "mot.attn", R * d.heads * hd * 2);
  add(L, "mot.act", R * I * 2);
  add(L, "mot.h", R * Hd * 2);
  add(L, "mot.y", T * Hd * 4);
  // heads
  const long long rows = N * P, dh = d.dec_heads, hp = L.dec_hp, ehd = dec_hd;
  add(L, "rec.hidden", rows * Hd * 4);
  add(L, "dec.x", rows * Hd * 4);
  add(L, "dec.h", rows * Hd * 2);
  add(L, "dec.qkv", rows * 3 * dh * hp * 2);                           // pad columns of the head slots stay zero
  add(L, "dec.attn", rows * dh * (L.dec_compact ? ehd : hp) * 2);
  add(L, "dec.mid", rows * Hd * d.dec_mlp_ratio * 2);
  add(L, "dec.yh", P * Hd * 2);
  add(L, "dec.kvc", P * 2 * dh * hp * 2);                              // zero pad columns
  add(L, "dec.qc", rows * dh * hp * 2);                                // zero pad columns
  add(L, "rec.point_hidden", rows * d.point_dim * 2);
  add(L, "rec.camera_hidden", rows * d.camera_dim * 4);
  add(L, "rec.global_hidden", rows * d.point_dim * 2);
  add(L, "rec.conf_hidden", d.train_conf ? rows * d.point_dim * 2 : 0);
  add(L, "cam.t1", rows * d.camera_dim * 4);
  add(L, "cam.t2", rows * d.camera_dim * 4);
  add(L, "cam.f2", rows * d.camera_dim * 4);
  add(L, "cam.f3", rows * d.camera_dim * 4);
  add(L, "cam.split", rows * 3 * d.camera_dim * 2);
  add(L, "cam.pooled", N * d.camera_dim * 4);
  add(L, "cam.m1", N * d.camera_dim * 4);
  add(L, "cam.m2", N * d.camera_dim * 4);
  add(L, "rec.feat_pts", rows * 3 * d.dino_patch * d.dino_patch * 4);
  *out = L;
  return G2VLM_OK;
}

struct Ws {
  uint8_t* base;
  const Layout* L;
  template <typename T>
  T* p(const char* name) const { return reinterpret_cast<T*>(base + L->r.at(name).off); }
};

// `rows x 256`-row work items of a segment table (flash-attn cu_seqlens), as ops.attention_work_table builds them
static void push_items(std::vector<int32_t>& v, int qb, int qe, int kb, int ke, int causal) {
  for (int t0 = qb; t0 < qe; t0 += 256) {
    const int32_t it[8] = {t0, qb, qe, kb, ke, causal, 0, 0};
    v.insert(v.end(), it, it + 8);
  }
}

static const Slot* find(const g2vlm_ctx* c, const std::string& name) {
  auto it = c->w.find(name);
  return it == c->w.end() ? nullptr : &it->second;
}

#define G2_TRY(expr)                          \
  do {                                        \
    int _rc = (expr);                         \
    if (_rc != G2VLM_OK) return _rc;          \
    g2::g_driver_launches.fetch_add(1, std::memory_order_relaxed); \
  } while (0)

// weight lookup that fails the call with a message naming the missing slot
#define G2_W(var, type, name_expr)                                                     \
  const type* var = nullptr;                                                           \
  {                                                                                    \
    const std::string _n = (name_expr);                                                \
    const g2::Slot* _s = g2::find(ctx, _n);                                            \
    if (_s == nullptr) {                                                               \
      g2::set_last_error(__FILE__, __LINE__, ("weight not loaded: " + _n).c_str());    \
      return G2VLM_ERR_INVALID;                                                        \
    }                                                                                  \
    var = reinterpret_cast<const type*>(_s->ptr);                                      \
  }

struct Gemm {
  g2vlm_gemm_args a;
  Gemm(const void* A, long long lda, long long rows, const void* B, int N, int K, int epilogue, void* out, long long ldo) {
    memset(&a, 0, sizeof(a));
    a.A = A; a.lda = lda; a.a_rows = rows; a.B = B; a.ldb = K; a.N = N; a.K = K;
    a.n_groups = 1; a.group_row0[0] = 0; a.group_rows[0] = (int32_t)rows;
    a.epilogue = epilogue; a.out = out; a.ldo = ldo;
  }
  Gemm& groups(int n_first, int n_second) {
    a.n_groups = 2;
    a.group_row0[0] = 0; a.group_rows[0] = n_first;
    a.group_row0[1] = n_first; a.group_rows[1] = n_second;
    return *this;
  }
  Gemm& bias(const float* b) { a.bias = b; return *this; }
  Gemm& flags(uint32_t f) { a.flags = f; return *this; }
  Gemm& scale(const float* s, uint32_t g) { a.scale = s; a.scale_groups = g; return *this; }
  Gemm& residual(const float* r, long long ldr) { a.residual = r; a.ldr = ldr; return *this; }
  Gemm& regroup(int group, int stride) { a.out_col_group = group; a.out_col_stride = stride; return *this; }
  int run(void* stream) const { return g2vlm_gemm_bf16(&a, stream); }
};

static int attention(const void* q, long long ldq, long long q_rows, const void* k, long long ldk, const void* v, long long ldv,
                     long long kv_rows, void* out, long long ldo, int hq, int hk, int hd, float scale, const int32_t* work,
                     int n_items, int out_head_cols, void* stream) {
  g2vlm_attn_args a;
  memset(&a, 0, sizeof(a));
  a.q = q; a.ldq = ldq; a.q_rows = q_rows; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.kv_rows = kv_rows;
  a.out = out; a.ldo = ldo; a.num_q_heads = hq; a.num_kv_heads = hk; a.head_dim = hd; a.softmax_scale = scale;
  a.n_items = n_items; a.work_items = work; a.out_head_cols = out_head_cols;
  return g2vlm_attention(&a, stream);
}

static std::string key(const char* fmt, int i, const char* leaf) {
  char buf[160];
  snprintf(buf, sizeof(buf), fmt, i, leaf);
  return buf;
}

}  // namespace g2

using namespace g2;
typedef __nv_bfloat16 bf16;

extern "C" int64_t g2vlm_driver_launches(void) { return g_driver_launches.load(std::memory_order_relaxed); }

extern "C" int g2vlm_ctx_create(const g2vlm_dims* dims, g2vlm_ctx** ctx) {
  G2_REQUIRE(dims != nullptr && ctx != nullptr, "ctx_create: null argument");
  G2_REQUIRE(dims->hidden > 0 && dims->heads > 0 && dims->kv_heads > 0 && dims->hidden / dims->heads == 128,
             "ctx_create: LLM head_dim must be 128 (mrope_section is hard-coded to [16,24,24])");
  G2_REQUIRE(dims->dino_hidden > 0 && dims->dino_heads > 0 && dims->dino_hidden / dims->dino_heads <= 128 &&
                 dims->dec_heads > 0 && dims->hidden / dims->dec_heads <= 128, "ctx_create: head dims above 128 are not supported");
  g2vlm_ctx* c = new (std::nothrow) g2vlm_ctx();
  G2_REQUIRE(c != nullptr, "ctx_create: out of host memory");
  c->d = *dims;
  G2_CUDA_OK(cudaGetDevice(&c->device));
  *ctx = c;
  return G2VLM_OK;
}

extern "C" int g2vlm_ctx_destroy(g2vlm_ctx* ctx) {
  delete ctx;   // the context owns no device memory: weights and workspace belong to the caller
  return G2VLM_OK;
}

extern "C" int g2vlm_load_weights(g2vlm_ctx* ctx, const char* name, const void* ptr, int32_t dtype, int64_t rows,
                                  int64_t cols) {
  G2_REQUIRE(ctx && name && ptr, "load_weights: null argument");
  G2_REQUIRE(dtype == G2VLM_DTYPE_F32 || dtype == G2VLM_DTYPE_BF16, "load_weights: dtype must be G2VLM_DTYPE_F32 / _BF16");
  G2_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "load_weights: tensors must be 16-byte aligned");
  ctx->w[name] = Slot{ptr, dtype, rows, cols};
  return G2VLM_OK;
}

extern "C" int64_t g2vlm_workspace_bytes(const g2vlm_ctx* ctx, int32_t n_views, int32_t H, int32_t W, int32_t n_prompt) {
  if (ctx == nullptr) return -1;
  Layout L;
  if (make_layout(ctx->d, n_views, H, W, n_prompt, &L) != G2VLM_OK) return -1;
  return L.total;
}

extern "C" int g2vlm_recon_plan(g2vlm_ctx* ctx, int32_t n_views, int32_t H, int32_t W, int32_t n_prompt,
                                const int32_t* dino_seqlens, void* workspace, int64_t workspace_bytes, void* stream) {
  G2_REQUIRE(ctx && workspace && dino_seqlens, "plan: null argument");
  Layout L;
  if (int rc = make_layout(ctx->d, n_views, H, W, n_prompt, &L)) return rc;
  G2_REQUIRE(workspace_bytes >= L.total, "plan: workspace smaller than g2vlm_workspace_bytes");
  G2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "plan: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  // ---- attention work tables (one item per <= 256 query rows of a segment) ----
  std::vector<int32_t> dino, mot, dec, cross;
  int cu = 0;
  for (int v = 0; v < n_views; ++v) {                                   // caller's cu_seqlens (quirk Q1: patch counts)
    G2_REQUIRE(dino_seqlens[v] >= 0, "plan: negative dino_seqlens entry");
    push_items(dino, cu, cu + dino_seqlens[v], cu, cu + dino_seqlens[v], 0);
    cu += dino_seqlens[v];
  }
  G2_REQUIRE(cu <= n_views * L.S, "plan: dino_seqlens exceed the number of DINO rows");
  push_items(mot, 0, L.T, 0, L.R, 0);                                   // geo-step rows see every key incl. the prompt's
  if (L.Kp > 0) push_items(mot, L.T, L.R, L.T, L.R, 1);                 // the prompt rows: causal among themselves
  for (int v = 0; v < n_views; ++v) {
    push_items(dec, v * L.P, (v + 1) * L.P, v * L.P, (v + 1) * L.P, 0); // per-view self-attention of the Pi3 decoders
    for (int t0 = v * L.P; t0 < (v + 1) * L.P; t0 += 256) {             // every view attends to view 0's tokens (context)
      const int32_t it[8] = {t0, v * L.P, (v + 1) * L.P, 0, L.P, 0, 0, 0};
      cross.insert(cross.end(), it, it + 8);
    }
  }
  struct { const char* name; std::vector<int32_t>* v; int* n; } tabs[4] = {
      {"tab.dino_work", &dino, &L.n_dino_items}, {"tab.mot_work", &mot, &L.n_mot_items},
      {"tab.dec_work", &dec, &L.n_dec_items}, {"tab.cross_work", &cross, &L.n_cross_items}};
  for (auto& t : tabs) {
    *t.n = (int)(t.v->size() / 8);
    G2_REQUIRE((long long)t.v->size() * 4 <= L.r.at(t.name).bytes, "plan: internal table overflow");
    if (!t.v->empty())   // pageable source: the runtime stages it before returning, the vector may die afterwards
      G2_CUDA_OK(cudaMemcpyAsync(base + L.r.at(t.name).off, t.v->data(), t.v->size() * 4, cudaMemcpyHostToDevice, st));
  }
  for (const char* z : {"dino.attn", "dec.qkv", "dec.kvc", "dec.qc"})
    G2_CUDA_OK(cudaMemsetAsync(base + L.r.at(z).off, 0, (size_t)L.r.at(z).bytes, st));
  ctx->plan = L;
  ctx->planned = true;
  return G2VLM_OK;
}

extern "C" int g2vlm_workspace_region(const g2vlm_ctx* ctx, const char* name, int64_t* offset, int64_t* bytes) {
  G2_REQUIRE(ctx && ctx->planned && name && offset && bytes, "workspace_region: needs a planned context");
  auto it = ctx->plan.r.find(name);
  G2_REQUIRE(it != ctx->plan.r.end(), "workspace_region: unknown region");
  *offset = it->second.off;
  *bytes = it->second.bytes;
  return G2VLM_OK;
}

#define G2_PLAN_CHECK(nv, hh, ww)                                                                              \
  G2_REQUIRE(ctx && ctx->planned, "g2vlm_recon_plan must run before the stage calls");                         \
  G2_REQUIRE(ctx->plan.n_views == (nv) && ctx->plan.H == (hh) && ctx->plan.W == (ww), "geometry differs from the plan")

extern "C" int g2vlm_dino_forward(g2vlm_ctx* ctx, const float* images, int32_t n_views, int32_t H, int32_t W,
                                  int32_t normalize, const float* pos_embed, void* workspace, void** tokens_out,
                                  void* stream) {
  G2_PLAN_CHECK(n_views, H, W);
  G2_REQUIRE(images && pos_embed && workspace, "dino_forward: null argument");
  const g2vlm_dims& d = ctx->d;
  const Layout& L = ctx->plan;
  const Ws ws{reinterpret_cast<uint8_t*>(workspace), &L};
  const int N = n_views, P = L.P, S = L.S, D = d.dino_hidden, nh = d.dino_heads, hp = L.dino_hp;
  const long long rows = (long long)N * S;
  const int kpad = (3 * d.dino_patch * d.dino_patch + 63) / 64 * 64;
  static const float mean3[3] = {0.485f, 0.456f, 0.406f}, std3[3] = {0.229f, 0.224f, 0.225f};   // g2vlm.py:33-34, 950
  bf16* patches = ws.p<bf16>("dino.patches");
  bf16* emb = ws.p<bf16>("dino.emb");
  float* x = ws.p<float>("dino.x");
  bf16* h = ws.p<bf16>("dino.h");
  bf16* qkv = ws.p<bf16>("dino.qkv");
  bf16* attn = ws.p<bf16>("dino.attn");
  bf16* mid = ws.p<bf16>("dino.mid");
  bf16* tokens = ws.p<bf16>("dino.tokens");
  const int32_t* work = ws.p<int32_t>("tab.dino_work");
  G2_W(wpatch, bf16, "dino.wpatch");
  G2_W(bpatch, float, "dino.bpatch");
  G2_W(cls, float, "dino.cls");
  G2_W(reg, float, "dino.reg");
  G2_TRY(g2vlm_im2col_patches(images, patches, N, H, W, d.dino_patch, kpad, normalize ? mean3 : nullptr,
                              normalize ? std3 : nullptr, stream));
  G2_TRY(Gemm(patches, kpad, (long long)N * P, wpatch, D, kpad, G2VLM_EPI_STORE_BF16, emb, D).bias(bpatch).run(stream));
  G2_TRY(g2vlm_dino_embed(emb, D, cls, reg, pos_embed, x, N, P, d.dino_registers, D, stream));
  const float scale = (float)(1.0 / sqrt((double)(D / nh)));
  const int qw = nh * hp;
  for (int l = 0; l < d.dino_layers; ++l) {
    G2_W(n1w, float, key("dino.%d.%s", l, "norm1w")); G2_W(n1b, float, key("dino.%d.%s", l, "norm1b"));
    G2_W(n2w, float, key("dino.%d.%s", l, "norm2w")); G2_W(n2b, float, key("dino.%d.%s", l, "norm2b"));
    G2_W(wqkv, bf16, key("dino.%d.%s", l, "wqkv")); G2_W(bqkv, float, key("dino.%d.%s", l, "bqkv"));
    G2_W(wdense, bf16, key("dino.%d.%s", l, "wdense")); G2_W(bdense, float, key("dino.%d.%s", l, "bdense"));
    G2_W(wfc1, bf16, key("dino.%d.%s", l, "wfc1")); G2_W(bfc1, float, key("dino.%d.%s", l, "bfc1"));
    G2_W(wfc2, bf16, key("dino.%d.%s", l, "wfc2")); G2_W(bfc2, float, key("dino.%d.%s", l, "bfc2"));
    G2_W(ls1, float, key("dino.%d.%s", l, "ls1")); G2_W(ls2, float, key("dino.%d.%s", l, "ls2"));
    G2_TRY(g2vlm_layernorm(x, D, h, D, 1, n1w, n1b, rows, D, d.dino_ln_eps, 0, 0, stream));
    G2_TRY(Gemm(h, D, rows, wqkv, 3 * qw, D, G2VLM_EPI_STORE_BF16, qkv, 3 * qw).bias(bqkv).run(stream));
    G2_TRY(attention(qkv, 3 * qw, rows, qkv + qw, 3 * qw, qkv + 2 * qw, 3 * qw, rows, attn, qw, nh, nh, hp, scale, work,
                     L.n_dino_items, 0, stream));
    G2_TRY(Gemm(attn, qw, rows, wdense, D, qw, G2VLM_EPI_RESID_F32, x, D).bias(bdense).scale(ls1, 1).run(stream));
    G2_TRY(g2vlm_layernorm(x, D, h, D, 1, n2w, n2b, rows, D, d.dino_ln_eps, 0, 0, stream));
    G2_TRY(Gemm(h, D, rows, wfc1, D * d.dino_mlp_ratio, D, G2VLM_EPI_STORE_BF16, mid, D * d.dino_mlp_ratio)
               .bias(bfc1).flags(G2VLM_GEMM_GELU).run(stream));
    G2_TRY(Gemm(mid, D * d.dino_mlp_ratio, rows, wfc2, D, D * d.dino_mlp_ratio, G2VLM_EPI_RESID_F32, x, D)
               .bias(bfc2).scale(ls2, 1).run(stream));
  }
  G2_W(lnw, float, "dino.lnw");
  G2_W(lnb, float, "dino.lnb");
  G2_TRY(g2vlm_layernorm(x, D, tokens, D, 1, lnw, lnb, rows, D, d.dino_ln_eps, S, 1 + d.dino_registers, stream));
  if (tokens_out) *tokens_out = tokens;
  return G2VLM_OK;
}

extern "C" int g2vlm_mot_forward_geo(g2vlm_ctx* ctx, const void* dino_tokens, const int64_t* packed_text_ids,
                                     const int64_t* packed_text_indexes, const int64_t* packed_dino_token_indexes,
                                     const int64_t* packed_position_ids, const int64_t* prompt_ids,
                                     const int64_t* prompt_position_ids, void* workspace, float* last_hidden,
                                     void* const* attention_events, void* stream) {
  G2_REQUIRE(ctx && ctx->planned, "g2vlm_recon_plan must run before the stage calls");
  G2_REQUIRE(dino_tokens && packed_text_ids && packed_text_indexes && packed_dino_token_indexes && packed_position_ids &&
                 workspace && last_hidden, "mot_forward_geo: null argument");
  const g2vlm_dims& d = ctx->d;
  const Layout& L = ctx->plan;
  G2_REQUIRE(L.Kp == 0 || (prompt_ids && prompt_position_ids), "mot_forward_geo: the plan has prompt rows but no prompt was given");
  const Ws ws{reinterpret_cast<uint8_t*>(workspace), &L};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int H = d.hidden, I = d.intermediate, nq = d.heads, nkv = d.kv_heads, hd = H / nq, half = hd / 2;
  const int T = L.T, R = L.R, Kp = L.Kp, n_geo = L.n_geo, n_und = L.n_und, D = d.dino_hidden;
  const int qkvw = (nq + 2 * nkv) * hd;
  float* geo_emb = ws.p<float>("mot.geo_emb");
  float* packed = ws.p<float>("mot.packed");
  float* txt = ws.p<float>("mot.txt");
  float* x = ws.p<float>("mot.x");
  float* cos_p = ws.p<float>("mot.cos_p");
  float* sin_p = ws.p<float>("mot.sin_p");
  float* cosb = ws.p<float>("mot.cos");
  float* sinb = ws.p<float>("mot.sin");
  bf16* qkv = ws.p<bf16>("mot.qkv");
  bf16* attn = ws.p<bf16>("mot.attn");
  bf16* act = ws.p<bf16>("mot.act");
  bf16* hb = ws.p<bf16>("mot.h");
  float* y = ws.p<float>("mot.y");
  int64_t* perm = ws.p<int64_t>("tab.perm");
  const int32_t* work = ws.p<int32_t>("tab.mot_work");
  G2_W(embed, float, "embed");
  G2_W(w_d2l, bf16, "dino2llm.w");
  G2_W(b_d2l, float, "dino2llm.b");
  G2_W(inv_freq, float, "inv_freq");
  G2_W(norm_geo, float, "norm_geo");
  G2_W(norm_und, float, "norm_und");
  // g2vlm.py:984-1010: embed the <start>/<end> ids, dino2llm on the patch tokens, scatter both into the packed sequence
  G2_TRY(Gemm(dino_tokens, D, n_geo, w_d2l, H, D, G2VLM_EPI_STORE_F32, geo_emb, H).bias(b_d2l).flags(G2VLM_GEMM_ROUND_BF16).run(stream));
  G2_TRY(g2vlm_gather_rows(embed, (int64_t)H * 4, txt, (int64_t)H * 4, packed_text_ids, n_und, (int64_t)H * 4, 0, stream));
  G2_TRY(g2vlm_gather_rows(txt, (int64_t)H * 4, packed, (int64_t)H * 4, packed_text_indexes, n_und, (int64_t)H * 4, 1, stream));
  G2_TRY(g2vlm_gather_rows(geo_emb, (int64_t)H * 4, packed, (int64_t)H * 4, packed_dino_token_indexes, n_geo, (int64_t)H * 4, 1, stream));
  // expert-permuted row order [geo rows | und rows | prompt rows]: internal row i <- packed row perm[i]
  G2_CUDA_OK(cudaMemcpyAsync(perm, packed_dino_token_indexes, (size_t)n_geo * 8, cudaMemcpyDeviceToDevice, st));
  G2_CUDA_OK(cudaMemcpyAsync(perm + n_geo, packed_text_indexes, (size_t)n_und * 8, cudaMemcpyDeviceToDevice, st));
  G2_TRY(g2vlm_gather_rows(packed, (int64_t)H * 4, x, (int64_t)H * 4, perm, T, (int64_t)H * 4, 0, stream));
  G2_TRY(g2vlm_mrope_table(packed_position_ids, T, inv_freq, cos_p, sin_p, T, half, d.mrope_s0, d.mrope_s1, stream));
  G2_TRY(g2vlm_gather_rows(cos_p, (int64_t)half * 4, cosb, (int64_t)half * 4, perm, T, (int64_t)half * 4, 0, stream));
  G2_TRY(g2vlm_gather_rows(sin_p, (int64_t)half * 4, sinb, (int64_t)half * 4, perm, T, (int64_t)half * 4, 0, stream));
  if (Kp > 0) {
    G2_TRY(g2vlm_gather_rows(embed, (int64_t)H * 4, x + (long long)T * H, (int64_t)H * 4, prompt_ids, Kp, (int64_t)H * 4, 0, stream));
    G2_TRY(g2vlm_mrope_table(prompt_position_ids, Kp, inv_freq, cosb + (long long)T * half, sinb + (long long)T * half, Kp,
                             half, d.mrope_s0, d.mrope_s1, stream));
  }
  const float scale = (float)(1.0 / sqrt((double)hd));
  const int n_second = R - n_geo;
  for (int l = 0; l < d.layers; ++l) {
    G2_W(in_g, float, key("mot.%d.%s", l, "input_layernorm_geo")); G2_W(in_u, float, key("mot.%d.%s", l, "input_layernorm_und"));
    G2_W(po_g, float, key("mot.%d.%s", l, "post_attention_layernorm_geo")); G2_W(po_u, float, key("mot.%d.%s", l, "post_attention_layernorm_und"));
    G2_W(qn_g, float, key("mot.%d.%s", l, "q_norm_geo")); G2_W(qn_u, float, key("mot.%d.%s", l, "q_norm_und"));
    G2_W(kn_g, float, key("mot.%d.%s", l, "k_norm_geo")); G2_W(kn_u, float, key("mot.%d.%s", l, "k_norm_und"));
    G2_W(wqkv, bf16, key("mot.%d.%s", l, "wqkv")); G2_W(bqkv, float, key("mot.%d.%s", l, "bqkv"));
    G2_W(wo, bf16, key("mot.%d.%s", l, "wo")); G2_W(wgu, bf16, key("mot.%d.%s", l, "wgu")); G2_W(wdown, bf16, key("mot.%d.%s", l, "wdown"));
    G2_W(ls1, float, key("mot.%d.%s", l, "ls1")); G2_W(ls2, float, key("mot.%d.%s", l, "ls2"));
    G2_TRY(g2vlm_rmsnorm_routed(x, H, hb, H, 1, in_g, in_u, R, n_geo, H, d.rms_eps, stream));
    G2_TRY(Gemm(hb, H, R, wqkv, qkvw, H, G2VLM_EPI_STORE_BF16, qkv, qkvw).groups(n_geo, n_second).bias(bqkv).run(stream));
    G2_TRY(g2vlm_qknorm_mrope(qkv, qkvw, T, n_geo, nq, nkv, hd, qn_g, kn_g, qn_u, kn_u, cosb, sinb, d.rms_eps, 0, stream));
    if (Kp > 0)   // the und prefill normalises a bf16 tensor: normalised value rounded before the weight multiply
      G2_TRY(g2vlm_qknorm_mrope(qkv + (long long)T * qkvw, qkvw, Kp, 0, nq, nkv, hd, qn_g, kn_g, qn_u, kn_u,
                                cosb + (long long)T * half, sinb + (long long)T * half, d.rms_eps, 1, stream));
    if (attention_events) G2_CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(attention_events[2 * l]), st));
    G2_TRY(attention(qkv, qkvw, R, qkv + nq * hd, qkvw, qkv + (nq + nkv) * hd, qkvw, R, attn, nq * hd, nq, nkv, hd, scale,
                     work, L.n_mot_items, 0, stream));
    if (attention_events) G2_CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(attention_events[2 * l + 1]), st));
    G2_TRY(Gemm(attn, nq * hd, R, wo, H, nq * hd, G2VLM_EPI_RESID_F32, x, H).groups(n_geo, n_second).scale(ls1, 1)
               .flags(G2VLM_GEMM_ROUND_AFTER_SCALE).run(stream));
    G2_TRY(g2vlm_rmsnorm_routed(x, H, hb, H, 1, po_g, po_u, R, n_geo, H, d.rms_eps, stream));
    G2_TRY(Gemm(hb, H, R, wgu, 2 * I, H, G2VLM_EPI_SWIGLU_BF16, act, I).groups(n_geo, n_second).run(stream));
    G2_TRY(Gemm(act, I, R, wdown, H, I, G2VLM_EPI_RESID_F32, x, H).groups(n_geo, n_second).scale(ls2, 1)
               .flags(G2VLM_GEMM_ROUND_AFTER_SCALE).run(stream));
  }
  G2_TRY(g2vlm_rmsnorm_routed(x, H, y, H, 0, norm_geo, norm_und, T, n_geo, H, d.rms_eps, stream));
  G2_TRY(g2vlm_gather_rows(y, (int64_t)H * 4, last_hidden, (int64_t)H * 4, perm, T, (int64_t)H * 4, 1, stream));
  return G2VLM_OK;
}

// ---- Pi3 decoders + heads ---------------------------------------------------------------------------------------
static int run_decoder(g2vlm_ctx* ctx, const Ws& ws, const char* name, bool cross, const float* hidden, const float* context,
                       const float* cos_tab, const float* sin_tab, void* out, bool out_f32, int out_dim, void* stream) {
  const g2vlm_dims& d = ctx->d;
  const Layout& L = *ws.L;
  const int N = L.n_views, P = L.P, H = d.hidden, dh = d.dec_heads, hp = L.dec_hp, ehd = H / dh, gw = L.gw;
  const long long rows = (long long)N * P;
  const bool cmp = L.dec_compact != 0;
  const int grp = cmp ? ehd : 0, strd = cmp ? hp : 0, ohc = cmp ? ehd : 0;
  const int aw = dh * (cmp ? ehd : hp);               // attention output width
  const int F = H * d.dec_mlp_ratio;
  float* x = ws.p<float>("dec.x");
  bf16* h = ws.p<bf16>("dec.h");
  bf16* qkv = ws.p<bf16>("dec.qkv");
  bf16* attn = ws.p<bf16>("dec.attn");
  bf16* mid = ws.p<bf16>("dec.mid");
  bf16* yh = ws.p<bf16>("dec.yh");
  bf16* kvc = ws.p<bf16>("dec.kvc");
  bf16* qc = ws.p<bf16>("dec.qc");
  const int32_t* work = ws.p<int32_t>("tab.dec_work");
  const int32_t* cwork = ws.p<int32_t>("tab.cross_work");
  const float scale = (float)(1.0 / sqrt((double)ehd));
  const int qw = dh * hp;
  const std::string pre = std::string("dec.") + name + ".";
  G2_TRY(g2vlm_gather_rows(hidden, (int64_t)H * 4, x, (int64_t)H * 4, nullptr, rows, (int64_t)H * 4, 0, stream));
  for (int b = 0; b < d.dec_depth; ++b) {
    const std::string bp = pre + std::to_string(b) + ".";
    G2_W(n1w, float, bp + "norm1w"); G2_W(n1b, float, bp + "norm1b");
    G2_W(n2w, float, bp + "norm2w"); G2_W(n2b, float, bp + "norm2b");
    G2_W(wqkv, bf16, bp + "wqkv"); G2_W(bqkv, float, bp + "bqkv");
    G2_W(wproj, bf16, bp + "wproj"); G2_W(bproj, float, bp + "bproj");
    G2_W(wfc1, bf16, bp + "wfc1"); G2_W(bfc1, float, bp + "bfc1");
    G2_W(wfc2, bf16, bp + "wfc2"); G2_W(bfc2, float, bp + "bfc2");
    const int n_qkv = 3 * dh * (cmp ? ehd : hp);       // unpadded Linear when the heads are regrouped into slots
    G2_TRY(g2vlm_layernorm(x, H, h, H, 1, n1w, n1b, rows, H, 1e-6f, 0, 0, stream));
    G2_TRY(Gemm(h, H, rows, wqkv, n_qkv, H, G2VLM_EPI_STORE_BF16, qkv, 3 * qw).bias(bqkv).regroup(grp, strd).run(stream));
    G2_TRY(g2vlm_rope2d(qkv, 3 * qw, rows, 2 * dh, hp, ehd, P, gw, cos_tab, sin_tab, 1, stream));
    G2_TRY(attention(qkv, 3 * qw, rows, qkv + qw, 3 * qw, qkv + 2 * qw, 3 * qw, rows, attn, aw, dh, dh, hp, scale, work,
                     L.n_dec_items, ohc, stream));
    G2_TRY(Gemm(attn, aw, rows, wproj, H, aw, G2VLM_EPI_RESID_F32, x, H).bias(bproj).run(stream));
    if (cross) {
      G2_W(nyw, float, bp + "norm_yw"); G2_W(nyb, float, bp + "norm_yb");
      G2_W(n3w, float, bp + "norm3w"); G2_W(n3b, float, bp + "norm3b");
      G2_W(wckv, bf16, bp + "wckv"); G2_W(bckv, float, bp + "bckv");
      G2_W(wcq, bf16, bp + "wcq"); G2_W(bcq, float, bp + "bcq");
      G2_W(wcproj, bf16, bp + "wcproj"); G2_W(bcproj, float, bp + "bcproj");
      const int n_kv = 2 * dh * (cmp ? ehd : hp), n_q = dh * (cmp ? ehd : hp);
      G2_TRY(g2vlm_layernorm(context, H, yh, H, 1, nyw, nyb, P, H, 1e-6f, 0, 0, stream));
      G2_TRY(Gemm(yh, H, P, wckv, n_kv, H, G2VLM_EPI_STORE_BF16, kvc, 2 * qw).bias(bckv).regroup(grp, strd).run(stream));
      G2_TRY(g2vlm_rope2d(kvc, 2 * qw, P, dh, hp, ehd, P, gw, cos_tab, sin_tab, 1, stream));
      G2_TRY(g2vlm_layernorm(x, H, h, H, 1, n2w, n2b, rows, H, 1e-6f, 0, 0, stream));
      G2_TRY(Gemm(h, H, rows, wcq, n_q, H, G2VLM_EPI_STORE_BF16, qc, qw).bias(bcq).regroup(grp, strd).run(stream));
      G2_TRY(g2vlm_rope2d(qc, qw, rows, dh, hp, ehd, P, gw, cos_tab, sin_tab, 1, stream));
      G2_TRY(attention(qc, qw, rows, kvc, 2 * qw, kvc + qw, 2 * qw, P, attn, aw, dh, dh, hp, scale, cwork, L.n_cross_items,
                       ohc, stream));
      G2_TRY(Gemm(attn, aw, rows, wcproj, H, aw, G2VLM_EPI_RESID_F32, x, H).bias(bcproj).run(stream));
      G2_TRY(g2vlm_layernorm(x, H, h, H, 1, n3w, n3b, rows, H, 1e-6f, 0, 0, stream));
    } else {
      G2_TRY(g2vlm_layernorm(x, H, h, H, 1, n2w, n2b, rows, H, 1e-6f, 0, 0, stream));
    }
    G2_TRY(Gemm(h, H, rows, wfc1, F, H, G2VLM_EPI_STORE_BF16, mid, F).bias(bfc1).flags(G2VLM_GEMM_GELU).run(stream));
    G2_TRY(Gemm(mid, F, rows, wfc2, H, F, G2VLM_EPI_RESID_F32, x, H).bias(bfc2).run(stream));
  }
  G2_W(wout, bf16, pre + "wout");
  G2_W(bout, float, pre + "bout");
  G2_TRY(g2vlm_cast_f32_to_bf16(x, H, h, H, rows, H, stream));
  if (out_f32)   // `.float()` of the bf16 linear_out result
    G2_TRY(Gemm(h, H, rows, wout, out_dim, H, G2VLM_EPI_STORE_F32, out, out_dim).bias(bout).flags(G2VLM_GEMM_ROUND_BF16).run(stream));
  else
    G2_TRY(Gemm(h, H, rows, wout, out_dim, H, G2VLM_EPI_STORE_BF16, out, out_dim).bias(bout).run(stream));
  return G2VLM_OK;
}

// true-fp32 nn.Linear (autocast disabled in the reference) as one split-bf16 GEMM: [hi|hi|lo] x [hi|lo|hi]^T
static int linear_fp32(g2vlm_ctx* ctx, const Ws& ws, const float* x, long long rows, int k, const std::string& wname,
                       const std::string& bname, float* out, bool relu, const float* residual, void* stream) {
  G2_W(w3, bf16, wname);
  G2_W(b, float, bname);
  bf16* xs = ws.p<bf16>("cam.split");
  G2_TRY(g2vlm_split3_f32(x, k, xs, 3LL * k, rows, k, stream));
  Gemm g(xs, 3LL * k, rows, w3, k, 3 * k, G2VLM_EPI_STORE_F32, out, k);
  g.bias(b).flags(relu ? G2VLM_GEMM_RELU : 0u);
  if (residual) g.residual(residual, k);
  G2_TRY(g.run(stream));
  return G2VLM_OK;
}

extern "C" int g2vlm_recon_heads(g2vlm_ctx* ctx, const float* last_hidden, const int64_t* packed_dino_token_indexes,
                                 const float* rope_cos, const float* rope_sin, void* workspace, float* points,
                                 float* local_points, float* global_points, float* camera_poses, float* conf,
                                 void* stream) {
  G2_REQUIRE(ctx && ctx->planned, "g2vlm_recon_plan must run before the stage calls");
  G2_REQUIRE(last_hidden && packed_dino_token_indexes && rope_cos && rope_sin && workspace && points && local_points &&
                 global_points && camera_poses, "recon_heads: null argument");
  const g2vlm_dims& d = ctx->d;
  const Layout& L = ctx->plan;
  G2_REQUIRE(!d.train_conf || conf != nullptr, "recon_heads: the model has a conf branch but conf is NULL");
  const Ws ws{reinterpret_cast<uint8_t*>(workspace), &L};
  const int N = L.n_views, P = L.P, H = d.hidden, C = d.camera_dim, p = d.dino_patch;
  const long long rows = (long long)N * P;
  float* hidden = ws.p<float>("rec.hidden");
  bf16* point_hidden = ws.p<bf16>("rec.point_hidden");
  float* camera_hidden = ws.p<float>("rec.camera_hidden");
  bf16* global_hidden = ws.p<bf16>("rec.global_hidden");
  G2_TRY(g2vlm_gather_rows(last_hidden, (int64_t)H * 4, hidden, (int64_t)H * 4, packed_dino_token_indexes, rows, (int64_t)H * 4, 0, stream));
  if (int rc = run_decoder(ctx, ws, "point_decoder", false, hidden, nullptr, rope_cos, rope_sin, point_hidden, false, d.point_dim, stream)) return rc;
  if (int rc = run_decoder(ctx, ws, "camera_decoder", false, hidden, nullptr, rope_cos, rope_sin, camera_hidden, true, C, stream)) return rc;
  if (int rc = run_decoder(ctx, ws, "global_points_decoder", true, hidden, hidden /* view 0's tokens, g2vlm.py:1196 */,
                           rope_cos, rope_sin, global_hidden, false, d.point_dim, stream)) return rc;
  // camera head (camera_head.py:48-93), fp32
  float* t1 = ws.p<float>("cam.t1");
  float* t2 = ws.p<float>("cam.t2");
  float* f[2] = {ws.p<float>("cam.f2"), ws.p<float>("cam.f3")};
  const float* feat = camera_hidden;
  for (int i = 0; i < 2; ++i) {
    const std::string r = "cam.r" + std::to_string(i);
    if (int rc = linear_fp32(ctx, ws, feat, rows, C, r + "1w", r + "1b", t1, true, nullptr, stream)) return rc;
    if (int rc = linear_fp32(ctx, ws, t1, rows, C, r + "2w", r + "2b", t2, true, nullptr, stream)) return rc;
    if (int rc = linear_fp32(ctx, ws, t2, rows, C, r + "3w", r + "3b", f[i], true, feat, stream)) return rc;
    feat = f[i];
  }
  float* pooled = ws.p<float>("cam.pooled");
  float* m1 = ws.p<float>("cam.m1");
  float* m2 = ws.p<float>("cam.m2");
  G2_TRY(g2vlm_mean_pool(feat, C, pooled, N, P, C, stream));
  if (int rc = linear_fp32(ctx, ws, pooled, N, C, "cam.m0w", "cam.m0b", m1, true, nullptr, stream)) return rc;
  if (int rc = linear_fp32(ctx, ws, m1, N, C, "cam.m2w", "cam.m2b", m2, true, nullptr, stream)) return rc;
  G2_W(fc_tw, float, "cam.fc_tw"); G2_W(fc_tb, float, "cam.fc_tb");
  G2_W(fc_rw, float, "cam.fc_rotw"); G2_W(fc_rb, float, "cam.fc_rotb");
  G2_TRY(g2vlm_camera_pose(m2, C, fc_tw, fc_tb, fc_rw, fc_rb, camera_poses, N, C, stream));
  // point heads: fp32 Linear on bf16-exact activations = two bf16 GEMMs (W = hi + lo), then pixel shuffle + epilogue
  const int nf = 3 * p * p;
  float* feat_pts = ws.p<float>("rec.feat_pts");
  struct { const char* head; const bf16* hid; int mode; float* out0; float* out1; const float* poses; int n_out; } heads[3] = {
      {"point_head", point_hidden, 1, local_points, points, camera_poses, nf},
      {"global_point_head", global_hidden, 0, global_points, nullptr, nullptr, nf},
      {"conf_head", nullptr, 2, conf, nullptr, nullptr, p * p}};
  for (int i = 0; i < (d.train_conf ? 3 : 2); ++i) {
    const bf16* hid = heads[i].hid;
    if (i == 2) {   // confidence branch (g2vlm.py:209-226): a copy of the point decoder + a 1-channel head
      bf16* conf_hidden = ws.p<bf16>("rec.conf_hidden");
      if (int rc = run_decoder(ctx, ws, "conf_decoder", false, hidden, nullptr, rope_cos, rope_sin, conf_hidden, false, d.point_dim, stream)) return rc;
      hid = conf_hidden;
    }
    const std::string hn = std::string("head.") + heads[i].head + ".";
    G2_W(whi, bf16, hn + "whi"); G2_W(wlo, bf16, hn + "wlo"); G2_W(hb, float, hn + "b");
    G2_TRY(Gemm(hid, d.point_dim, rows, whi, heads[i].n_out, d.point_dim, G2VLM_EPI_STORE_F32, feat_pts, heads[i].n_out).bias(hb).run(stream));
    G2_TRY(Gemm(hid, d.point_dim, rows, wlo, heads[i].n_out, d.point_dim, G2VLM_EPI_STORE_F32, feat_pts, heads[i].n_out)
               .flags(G2VLM_GEMM_ACCUMULATE).run(stream));
    G2_TRY(g2vlm_points_epilogue(feat_pts, heads[i].n_out, heads[i].poses, heads[i].out0, heads[i].out1, N, L.H, L.W, p,
                                 heads[i].mode, stream));
  }
  return G2VLM_OK;
}
