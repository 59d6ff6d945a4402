// g2vlm_b200 — elementwise kernels of the fp32 mode (north_star: "fp32 mode <= 1e-4"): fp32 in, fp32 out, no bf16
// rounding point anywhere.  Ground truth: oracle/restate.py mode="fp32" (the reference itself cannot run in fp32:
// hard bf16 casts at modeling/g2vlm/qwen2vl.py:579, 617-619).  Memory-bound, vectorised 16-byte accesses.
#include "common.cuh"

namespace g2 {

constexpr int F32_THREADS = 256;

static inline unsigned f32_blocks(long long items, int per_block) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 32;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

struct Norm3F {
  float mean[3];
  float std[3];
  int enabled;
};

// nn.Conv2d(3, D, 14, 14) input as GEMM rows (dinov2_with_registers/modeling_...:62,71), fp32 patches
__global__ void im2col_f32_kernel(const float* __restrict__ img, float* __restrict__ out, int n, int H, int W, int patch,
                                  int k_pad, Norm3F nrm) {
  const int gh = H / patch, gw = W / patch;
  const long long total = (long long)n * gh * gw * k_pad;
  const int pp = patch * patch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / k_pad;
    const int col = static_cast<int>(i - row * k_pad);
    float v = 0.f;
    if (col < 3 * pp) {
      const int c = col / pp, rem = col - c * pp;
      const int py = rem / patch, px = rem - py * patch;
      const int im = static_cast<int>(row / (gh * gw));
      const int t = static_cast<int>(row - (long long)im * gh * gw);
      const int gy = t / gw, gx = t - gy * gw;
      v = img[(((long long)im * 3 + c) * H + gy * patch + py) * W + gx * patch + px];
      if (nrm.enabled) v = __fdiv_rn(__fsub_rn(v, nrm.mean[c]), nrm.std[c]);   // torchvision Normalize, g2vlm.py:950
    }
    out[i] = v;
  }
}

// Dinov2WithRegistersEmbeddings.forward (:147-171) with an fp32 patch embedding
__global__ void dino_embed_f32_kernel(const float* __restrict__ patch_emb, long long ld_patch, const float* __restrict__ cls,
                                      const float* __restrict__ reg, const float* __restrict__ pos, float* __restrict__ out,
                                      int n, int P, int n_reg, int dim) {
  const int S = 1 + n_reg + P;
  const int d4 = dim >> 2;
  const long long total = (long long)n * S * d4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4) * 4;
    const int im = static_cast<int>(row / S);
    const int local = static_cast<int>(row - (long long)im * S);
    float4 v;
    if (local == 0) {
      const float4 a = *reinterpret_cast<const float4*>(cls + c);
      const float4 b = *reinterpret_cast<const float4*>(pos + c);
      v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    } else if (local <= n_reg) {
      v = *reinterpret_cast<const float4*>(reg + (long long)(local - 1) * dim + c);
    } else {
      const int t = local - 1 - n_reg;
      const float4 a = *reinterpret_cast<const float4*>(patch_emb + ((long long)im * P + t) * ld_patch + c);
      const float4 b = *reinterpret_cast<const float4*>(pos + (long long)(1 + t) * dim + c);
      v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    *reinterpret_cast<float4*>(out + row * dim + c) = v;
  }
}

// one warp per (row, head), head_dim = 128: 4 consecutive elements per lane; rotate_half partner = lane ^ 16
__global__ void qknorm_mrope_f32_kernel(float* __restrict__ qkv, long long ld, long long rows, long long n_first, int n_q,
                                        int n_kv, const float* __restrict__ qw_a, const float* __restrict__ kw_a,
                                        const float* __restrict__ qw_b, const float* __restrict__ kw_b,
                                        const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, float eps) {
  const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int heads = n_q + n_kv;
  if (wid >= rows * heads) return;
  const long long row = wid / heads;
  const int head = static_cast<int>(wid - row * heads);
  const int lane = threadIdx.x & 31;
  float* p = qkv + row * ld + head * 128 + lane * 4;
  const float4 x = *reinterpret_cast<const float4*>(p);
  const float v[4] = {x.x, x.y, x.z, x.w};
  float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / 128.0f + eps);
  const bool first = row < n_first;
  const float* w = head < n_q ? (first ? qw_a : qw_b) : (first ? kw_a : kw_b);
  const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane);
  const float gw[4] = {g.x, g.y, g.z, g.w};
  const int j0 = (lane & 15) * 4;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(cos_tab + row * 64 + j0));
  const float4 s4 = __ldg(reinterpret_cast<const float4*>(sin_tab + row * 64 + j0));
  const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float nv = gw[e] * (v[e] * r);
    const float partner = __shfl_xor_sync(0xffffffffu, nv, 16);
    const float rot = lane < 16 ? -partner : partner;     // rotate_half (modeling_qwen2_vl.py:170-174)
    o[e] = nv * cs[e] + rot * sn[e];
  }
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}

// RoPE2D (pi3/models/layers/pos_embed.py:112-159), one thread per rotation pair, fp32
__global__ void rope2d_f32_kernel(float* __restrict__ buf, long long ld, long long rows, int n_heads, int head_stride,
                                  int head_dim, int tokens_per_view, int grid_w, const float* __restrict__ cos_tab,
                                  const float* __restrict__ sin_tab) {
  const int quarter = head_dim >> 2;
  const int pairs = head_dim >> 1;
  const long long total = rows * n_heads * pairs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pr = static_cast<int>(i % pairs);
    const long long rh = i / pairs;
    const int head = static_cast<int>(rh % n_heads);
    const long long row = rh / n_heads;
    const int tok = static_cast<int>(row % tokens_per_view);
    const int axis = pr / quarter;
    const int j = pr - axis * quarter;
    const int position = axis == 0 ? tok / grid_w : tok % grid_w;
    const float c = cos_tab[position * quarter + j];
    const float s = sin_tab[position * quarter + j];
    float* p = buf + row * ld + (long long)head * head_stride + axis * (head_dim >> 1) + j;
    const float a = p[0], b = p[quarter];
    p[0] = a * c - b * s;
    p[quarter] = b * c + a * s;
  }
}

// fp32 [rows, k] -> bf16 [rows, 6k] = [m | l | h | m | h | h], x = h + m + l EXACTLY (3 x 8 significand bits); pairs with
// weights stored as [m | h | l | h | m | h]: the six products mm + lh + hl + mh + hm + hh (smallest first) drop only
// terms <= 2^-24.
__global__ void split6_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long ldo,
                              long long rows, int k) {
  const long long total = rows * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / k;
    const int c = static_cast<int>(i - r * k);
    const float v = x[r * ldx + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
    __nv_bfloat16* o = out + r * ldo + c;
    o[0] = m;
    o[k] = l;
    o[2 * k] = h;
    o[3 * k] = m;
    o[4 * k] = h;
    o[5 * k] = h;
  }
}

__global__ void swiglu_f32_kernel(const float* __restrict__ gu, long long ld_gu, float* __restrict__ out, long long ldo,
                                  long long rows, int inter) {
  const int c4n = inter >> 2;
  const long long total = rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / c4n;
    const int c = static_cast<int>(i - row * c4n) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gu + row * ld_gu + c);
    const float4 u = *reinterpret_cast<const float4*>(gu + row * ld_gu + inter + c);
    auto f = [](float gg, float uu) { return gg / (1.0f + expf(-gg)) * uu; };
    *reinterpret_cast<float4*>(out + row * ldo + c) = make_float4(f(g.x, u.x), f(g.y, u.y), f(g.z, u.z), f(g.w, u.w));
  }
}

}  // namespace g2

#define F32_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int g2vlm_im2col_patches_f32(const float* images, float* out, int32_t n, int32_t H, int32_t W, int32_t patch,
                                        int32_t k_pad, const float* mean3, const float* std3, void* stream) {
  using namespace g2;
  G2_REQUIRE(images && out, "im2col_f32: null tensor");
  G2_REQUIRE(patch > 0 && H % patch == 0 && W % patch == 0 && k_pad >= 3 * patch * patch, "im2col_f32: bad geometry");
  if (n <= 0) return G2VLM_OK;
  G2_REQUIRE((mean3 == nullptr) == (std3 == nullptr), "im2col_f32: mean3 and std3 must be given together");
  Norm3F nrm;
  nrm.enabled = mean3 != nullptr;
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = mean3 ? mean3[c] : 0.f;
    nrm.std[c] = std3 ? std3[c] : 1.f;
  }
  const long long total = (long long)n * (H / patch) * (W / patch) * k_pad;
  im2col_f32_kernel<<<f32_blocks(total, F32_THREADS * 4), F32_THREADS, 0, (cudaStream_t)stream>>>(images, out, n, H, W, patch,
                                                                                                k_pad, nrm);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

extern "C" int g2vlm_dino_embed_f32(const float* patch_emb, int64_t ld_patch, const float* cls, const float* reg,
                                    const float* pos, float* out, int32_t n, int32_t P, int32_t n_reg, int32_t dim,
                                    void* stream) {
  using namespace g2;
  G2_REQUIRE(patch_emb && cls && reg && pos && out, "dino_embed_f32: null tensor");
  G2_REQUIRE(dim % 4 == 0 && ld_patch % 4 == 0, "dino_embed_f32: dim must be a multiple of 4");
  G2_REQUIRE(F32_ALIGNED16(cls) && F32_ALIGNED16(reg) && F32_ALIGNED16(pos) && F32_ALIGNED16(out) && F32_ALIGNED16(patch_emb),
             "dino_embed_f32: alignment");
  if (n <= 0) return G2VLM_OK;
  const long long total = (long long)n * (1 + n_reg + P) * (dim / 4);
  dino_embed_f32_kernel<<<f32_blocks(total, F32_THREADS * 2), F32_THREADS, 0, (cudaStream_t)stream>>>(
      patch_emb, ld_patch, cls, reg, pos, out, n, P, n_reg, dim);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

extern "C" int g2vlm_qknorm_mrope_f32(float* qkv, int64_t ld, int64_t rows, int64_t n_first, int32_t n_q_heads,
                                      int32_t n_kv_heads, int32_t head_dim, const float* qw_a, const float* kw_a,
                                      const float* qw_b, const float* kw_b, const float* cos_tab, const float* sin_tab,
                                      float eps, void* stream) {
  using namespace g2;
  G2_REQUIRE(qkv && qw_a && kw_a && qw_b && kw_b && cos_tab && sin_tab, "qknorm_mrope_f32: null tensor");
  G2_REQUIRE(head_dim == 128, "qknorm_mrope_f32: head_dim must be 128 (mrope_section is hard-coded)");
  G2_REQUIRE(ld % 4 == 0 && F32_ALIGNED16(qkv) && F32_ALIGNED16(cos_tab) && F32_ALIGNED16(sin_tab) && F32_ALIGNED16(qw_a) &&
                 F32_ALIGNED16(kw_a) && F32_ALIGNED16(qw_b) && F32_ALIGNED16(kw_b), "qknorm_mrope_f32: alignment");
  if (rows <= 0) return G2VLM_OK;
  const long long warps = rows * (n_q_heads + n_kv_heads);
  const long long blocks = (warps + (F32_THREADS / 32) - 1) / (F32_THREADS / 32);
  G2_REQUIRE(blocks < (1LL << 31), "qknorm_mrope_f32: too many rows");
  qknorm_mrope_f32_kernel<<<(unsigned)blocks, F32_THREADS, 0, (cudaStream_t)stream>>>(
      qkv, ld, rows, n_first, n_q_heads, n_kv_heads, qw_a, kw_a, qw_b, kw_b, cos_tab, sin_tab, eps);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

extern "C" int g2vlm_rope2d_f32(float* buf, int64_t ld, int64_t rows, int32_t n_heads_total, int32_t head_stride,
                                int32_t head_dim, int32_t tokens_per_view, int32_t grid_w, const float* cos_tab,
                                const float* sin_tab, void* stream) {
  using namespace g2;
  G2_REQUIRE(buf && cos_tab && sin_tab, "rope2d_f32: null tensor");
  G2_REQUIRE(head_dim > 0 && head_dim % 4 == 0 && head_dim <= head_stride, "rope2d_f32: head_dim % 4, <= head_stride");
  G2_REQUIRE(tokens_per_view > 0 && grid_w > 0, "rope2d_f32: bad grid");
  if (rows <= 0) return G2VLM_OK;
  const long long total = rows * n_heads_total * (head_dim / 2);
  rope2d_f32_kernel<<<f32_blocks(total, F32_THREADS * 4), F32_THREADS, 0, (cudaStream_t)stream>>>(
      buf, ld, rows, n_heads_total, head_stride, head_dim, tokens_per_view, grid_w, cos_tab, sin_tab);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

extern "C" int g2vlm_split6_f32(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t k, void* stream) {
  using namespace g2;
  G2_REQUIRE(x && out && k > 0 && ldo >= 6LL * k, "split6: bad arguments");
  if (rows <= 0) return G2VLM_OK;
  split6_kernel<<<f32_blocks(rows * k, F32_THREADS * 4), F32_THREADS, 0, (cudaStream_t)stream>>>(
      x, ldx, reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, k);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

extern "C" int g2vlm_swiglu_f32(const float* gate_up, int64_t ld_gu, float* out, int64_t ldo, int64_t rows, int32_t inter,
                                void* stream) {
  using namespace g2;
  G2_REQUIRE(gate_up && out, "swiglu_f32: null tensor");
  G2_REQUIRE(inter > 0 && inter % 4 == 0 && ld_gu % 4 == 0 && ldo % 4 == 0 && F32_ALIGNED16(gate_up) && F32_ALIGNED16(out),
             "swiglu_f32: inter / leading dimensions must be multiples of 4, pointers 16-byte aligned");
  if (rows <= 0) return G2VLM_OK;
  const long long total = rows * (inter / 4);
  swiglu_f32_kernel<<<f32_blocks(total, F32_THREADS * 2), F32_THREADS, 0, (cudaStream_t)stream>>>(gate_up, ld_gu, out, ldo,
                                                                                                rows, inter);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}
