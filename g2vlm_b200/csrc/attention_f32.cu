// g2vlm_b200 — fp32 attention for the fp32 mode (north_star: "fp32 mode <= 1e-4"): the same varlen / GQA / per-item
// causal semantics as attention.cu (flash_attn_varlen_func, modeling/g2vlm/qwen2vl.py:643-652, dinov2_model.py:49-58,
// and the Pi3 SDPA calls pi3/models/layers/attention.py:255-259, 370-372) with fp32 operands, fp32 FMA-pipe dot
// products and an exact-precision exp2f softmax.  Bound by the FP32 pipe (no tensor-core type holds 24 mantissa
// bits); it is the verification mode, not the benchmarked path.
//
// CTA = 256 threads = one 64-row query tile of one head.  Thread (ty = tid / 16, tx = tid % 16) owns
//   S[ty + 16 i][tx + 16 j]  (i, j < 4)   and   O[ty + 16 i][4 tx + 64 jj + 0..3]  (jj < D / 64).
// Shared-memory rows are padded by 4 floats, so the 8 threads of an LDS.128 phase (tx..tx+7) touch 8 distinct
// 4-bank groups when they read rows tx + 16 j; the q / p operands are warp-broadcast (2 distinct ty per warp).
#include "common.cuh"

namespace g2 {

constexpr int AF_BM = 64;
constexpr int AF_BN = 64;
constexpr int AF_THREADS = 256;

struct AttnF32Params {
  const float* q;
  const float* k;
  const float* v;
  float* out;
  long long ldq, ldk, ldv, ldo;
  const int* work;
  int q_heads_per_kv, causal, n_heads, out_head_cols;
  float scale_log2;
};

template <int D>
struct AttnF32Cfg {
  static constexpr int kLd = D + 4;
  static constexpr int kPLd = AF_BN + 4;
  static constexpr int kSmem = (AF_BM * kLd + 2 * AF_BN * kLd + AF_BM * kPLd) * 4;
};

template <int D>
__global__ void __launch_bounds__(AF_THREADS) attention_f32_kernel(const AttnF32Params p) {
  using Cfg = AttnF32Cfg<D>;
  constexpr int LD = Cfg::kLd, PLD = Cfg::kPLd;
  constexpr int OJ = (D + 63) / 64;              // float4 column chunks of O per thread (D = 96: second chunk half used)
  extern __shared__ float smem_f[];
  float* Qs = smem_f;
  float* Ks = Qs + AF_BM * LD;
  float* Vs = Ks + AF_BN * LD;
  float* Ps = Vs + AF_BN * LD;

  const int item = blockIdx.x >> 2, sub = blockIdx.x & 3, head = blockIdx.y;
  const int* w = p.work + item * 8;
  const int q0 = w[0] + sub * AF_BM, seg_b = w[1], seg_e = w[2], k_b = w[3], len_k = w[4] - w[3];
  const bool causal = p.causal != 0 || w[5] != 0;
  if (q0 >= seg_e || q0 >= w[0] + 256) return;
  const int rows_here = min(AF_BM, seg_e - q0);
  const int len_q = seg_e - seg_b;
  const int kv_head = head / p.q_heads_per_kv;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;

  // keys this tile needs (causal: bottom-right aligned, as flash-attn)
  int k_needed = len_k;
  if (causal) k_needed = max(0, min(len_k, (q0 + rows_here - 1 - seg_b) + (len_k - len_q) + 1));

  // Q tile -> smem (rows beyond the segment are zero)
  for (int idx = tid; idx < AF_BM * (D / 4); idx += AF_THREADS) {
    const int r = idx / (D / 4), c4 = idx % (D / 4);
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_here) val = *reinterpret_cast<const float4*>(p.q + (long long)(q0 + r) * p.ldq + head * D + c4 * 4);
    *reinterpret_cast<float4*>(Qs + r * LD + c4 * 4) = val;
  }

  float m_run[4], l_run[4], o_acc[4][OJ * 4];
  int limit[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
    const int row = q0 + ty + 16 * i;
    limit[i] = causal ? max(0, min(len_k, (row - seg_b) + (len_k - len_q) + 1)) : len_k;
#pragma unroll
    for (int c = 0; c < OJ * 4; ++c) o_acc[i][c] = 0.f;
  }

  for (int kb0 = 0; kb0 < k_needed; kb0 += AF_BN) {
    __syncthreads();   // previous block's PV is done with Ks / Vs / Ps (and Q is in place on the first pass)
    for (int idx = tid; idx < AF_BN * (D / 4); idx += AF_THREADS) {
      const int r = idx / (D / 4), c4 = idx % (D / 4);
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (kb0 + r < len_k) {
        const long long row = (long long)(k_b + kb0 + r);
        kk = *reinterpret_cast<const float4*>(p.k + row * p.ldk + kv_head * D + c4 * 4);
        vv = *reinterpret_cast<const float4*>(p.v + row * p.ldv + kv_head * D + c4 * 4);
      }
      *reinterpret_cast<float4*>(Ks + r * LD + c4 * 4) = kk;
      *reinterpret_cast<float4*>(Vs + r * LD + c4 * 4) = vv;
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; d += 4) {
      float4 qv[4], kv4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * LD + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv4[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * LD + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv4[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv4[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv4[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv4[j].w, s[i][j]);
        }
    }

    // online softmax per row; a row is spread over the 16 tx lanes of a half-warp
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kb0 + tx + 16 * j;
        s[i][j] = key < limit[i] ? s[i][j] * p.scale_log2 : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[i], mx);
      const float alpha = (m_new == -INFINITY) ? 1.f : exp2f(m_run[i] - m_new);   // m_run = -inf -> alpha = 0
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pj = (m_new == -INFINITY) ? 0.f : exp2f(s[i][j] - m_new);
        sum += pj;
        Ps[(ty + 16 * i) * PLD + tx + 16 * j] = pj;
      }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      l_run[i] = l_run[i] * alpha + sum;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < OJ * 4; ++c) o_acc[i][c] *= alpha;
    }
    __syncthreads();

    // O += P V
#pragma unroll 2
    for (int kk = 0; kk < AF_BN; kk += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * PLD + kk);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int jj = 0; jj < OJ; ++jj) {
          const int col = 4 * tx + 64 * jj;
          if (col < D) {
            const float4 vv = *reinterpret_cast<const float4*>(Vs + (kk + t) * LD + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float pt = t == 0 ? pv[i].x : t == 1 ? pv[i].y : t == 2 ? pv[i].z : pv[i].w;
              o_acc[i][jj * 4 + 0] = fmaf(pt, vv.x, o_acc[i][jj * 4 + 0]);
              o_acc[i][jj * 4 + 1] = fmaf(pt, vv.y, o_acc[i][jj * 4 + 1]);
              o_acc[i][jj * 4 + 2] = fmaf(pt, vv.z, o_acc[i][jj * 4 + 2]);
              o_acc[i][jj * 4 + 3] = fmaf(pt, vv.w, o_acc[i][jj * 4 + 3]);
            }
          }
        }
      }
    }
  }

  // epilogue: O / l (a row without visible keys gets zeros, as flash-attn)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 16 * i;
    if (r >= rows_here) continue;
    const float inv = l_run[i] > 0.f ? 1.0f / l_run[i] : 0.f;
    float* dst = p.out + (long long)(q0 + r) * p.ldo + head * p.out_head_cols;
#pragma unroll
    for (int jj = 0; jj < OJ; ++jj) {
      const int col = 4 * tx + 64 * jj;
      if (col < D && col < p.out_head_cols)
        *reinterpret_cast<float4*>(dst + col) = make_float4(o_acc[i][jj * 4] * inv, o_acc[i][jj * 4 + 1] * inv,
                                                            o_acc[i][jj * 4 + 2] * inv, o_acc[i][jj * 4 + 3] * inv);
    }
  }
}

template <int D>
static int launch_attention_f32(const AttnF32Params& p, int n_items, cudaStream_t stream) {
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_f32_kernel<D>), AttnF32Cfg<D>::kSmem)) return rc;
  attention_f32_kernel<D><<<dim3(n_items * 4, p.n_heads), AF_THREADS, AttnF32Cfg<D>::kSmem, stream>>>(p);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

}  // namespace g2

extern "C" int g2vlm_attention_f32(const g2vlm_attn_args* a, void* stream_) {
  using namespace g2;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  G2_REQUIRE(a != nullptr, "attention_f32: null args");
  G2_REQUIRE(a->q && a->k && a->v && a->out, "attention_f32: null tensor");
  G2_REQUIRE(a->head_dim == 64 || a->head_dim == 96 || a->head_dim == 128 || a->head_dim == 32 || a->head_dim == 16,
             "attention_f32: head_dim must be 16, 32, 64, 96 or 128");
  G2_REQUIRE(a->num_q_heads > 0 && a->num_kv_heads > 0 && a->num_q_heads % a->num_kv_heads == 0,
             "attention_f32: num_q_heads must be a multiple of num_kv_heads");
  G2_REQUIRE(a->ldq % 4 == 0 && a->ldk % 4 == 0 && a->ldv % 4 == 0 && a->ldo % 4 == 0,
             "attention_f32: leading dimensions must be multiples of 4");
  G2_REQUIRE(((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
               reinterpret_cast<uintptr_t>(a->out)) & 15) == 0, "attention_f32: tensors must be 16-byte aligned");
  G2_REQUIRE(a->lse_out == nullptr, "attention_f32: lse_out is not supported");
  G2_REQUIRE(a->n_items >= 0 && a->n_items < (1 << 28), "attention_f32: bad n_items");
  if (a->n_items == 0) return G2VLM_OK;
  G2_REQUIRE(a->work_items != nullptr, "attention_f32: null work table");
  const int D = a->head_dim;
  G2_REQUIRE(a->out_head_cols == 0 || a->out_head_cols == D, "attention_f32: out_head_cols must be 0 or head_dim");
  AttnF32Params p;
  p.q = reinterpret_cast<const float*>(a->q);
  p.k = reinterpret_cast<const float*>(a->k);
  p.v = reinterpret_cast<const float*>(a->v);
  p.out = reinterpret_cast<float*>(a->out);
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.work = a->work_items;
  p.q_heads_per_kv = a->num_q_heads / a->num_kv_heads;
  p.causal = a->causal;
  p.n_heads = a->num_q_heads;
  p.out_head_cols = D;
  p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
  switch (D) {
    case 16: return launch_attention_f32<16>(p, a->n_items, stream);
    case 32: return launch_attention_f32<32>(p, a->n_items, stream);
    case 64: return launch_attention_f32<64>(p, a->n_items, stream);
    case 96: return launch_attention_f32<96>(p, a->n_items, stream);
    default: return launch_attention_f32<128>(p, a->n_items, stream);
  }
}
