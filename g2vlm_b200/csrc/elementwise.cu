// g2vlm_b200 — memory-bound kernels of the recon path: routing permutation copies, routed RMSNorm,
// LayerNorm, M-RoPE table, per-head RMSNorm + M-RoPE, DINO patch/im2col + embedding assembly, RoPE2D,
// pixel-shuffle + exp + unprojection epilogue, pooling, split-bf16 and the SVD camera pose.
// All are HBM-bound: 128-bit vector accesses, one warp per row for the reductions (shuffles only, no
// shared memory), grids sized to cover the rows. Reference call sites: see include/g2vlm_b200.h.
#include "common.cuh"

namespace g2 {

constexpr int EW_THREADS = 256;

static inline unsigned blocks_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

// ------------------------------------------------------------------------------------------------
// gather / scatter rows (16-byte chunks)
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, long long src_pitch,
                                   uint8_t* __restrict__ dst, long long dst_pitch,
                                   const long long* __restrict__ idx, long long n_rows, int chunks,
                                   int scatter) {
  const long long total = n_rows * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks);
    const long long other = idx ? idx[r] : r;
    const long long rs = scatter ? r : other;
    const long long rd = scatter ? other : r;
    const uint4 v = *reinterpret_cast<const uint4*>(src + rs * src_pitch + c * 16LL);
    *reinterpret_cast<uint4*>(dst + rd * dst_pitch + c * 16LL) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// RMSNorm (routed) and LayerNorm: one warp per row
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store4(void* out, bool bf16, long long off, float a, float b, float c, float d) {
  if (bf16) {
    uint2 v = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + off) = v;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off) = make_float4(a, b, c, d);
  }
}

// out_mode bit 1: the module itself is bf16 (training forward, reference Qwen2RMSNorm with a bf16 input): the normalised
// value is cast to bf16 before the weight multiply (modeling_qwen2_vl.py:501 `weight * hidden.to(input_dtype)`)
__device__ __forceinline__ float rn_(float x, int out_mode) { return (out_mode & 2) ? bf16_round(x) : x; }

__global__ void rmsnorm_routed_kernel(const float* __restrict__ x, long long ldx, void* __restrict__ out,
                                      long long ldo, int out_bf16, const float* __restrict__ w_a,
                                      const float* __restrict__ w_b, long long rows, long long n_first,
                                      int dim, float eps) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int n4 = dim >> 2;
  float ss = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / dim + eps);
  const float4* w = reinterpret_cast<const float4*>(row < n_first ? w_a : w_b);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];  // second pass hits L1
    const float4 g = __ldg(w + i);
    store4(out, out_bf16 & 1, row * ldo + 4LL * i, g.x * rn_(v.x * r, out_bf16), g.y * rn_(v.y * r, out_bf16),
           g.z * rn_(v.z * r, out_bf16), g.w * rn_(v.w * r, out_bf16));
  }
}

// Same arithmetic (same per-lane order), row held in registers: ONE global read per element with all NV float4
// loads of a lane in flight at once.  NV = dim / 128 (8 for the 1024-wide DINO / 12 for the 1536-wide streams).
template <int NV>
__global__ void __launch_bounds__(EW_THREADS)
rmsnorm_routed_reg_kernel(const float* __restrict__ x, long long ldx, void* __restrict__ out, long long ldo,
                          int out_bf16, const float* __restrict__ w_a, const float* __restrict__ w_b,
                          long long rows, long long n_first, float eps) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) v[u] = xr[lane + 32 * u];
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) ss += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / (NV * 128) + eps);
  const float4* w = reinterpret_cast<const float4*>(row < n_first ? w_a : w_b);
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int i = lane + 32 * u;
    const float4 g = __ldg(w + i);
    store4(out, out_bf16 & 1, row * ldo + 4LL * i, g.x * rn_(v[u].x * r, out_bf16), g.y * rn_(v[u].y * r, out_bf16),
           g.z * rn_(v[u].z * r, out_bf16), g.w * rn_(v[u].w * r, out_bf16));
  }
}

// few rows (decode steps, 7-token prefill): one BLOCK per row so every element is in flight at once — the
// warp-per-row kernel above is latency-bound there (12 dependent-latency load rounds per lane for H = 1536)
__global__ void __launch_bounds__(256)
rmsnorm_routed_block_kernel(const float* __restrict__ x, long long ldx, void* __restrict__ out, long long ldo,
                            int out_bf16, const float* __restrict__ w_a, const float* __restrict__ w_b,
                            long long n_first, int dim, float eps) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int n4 = dim >> 2;
  float4 v[4];  // dim <= 4096
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.x + 256 * u;
    v[u] = i < n4 ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  ss = red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7];
  const float r = rsqrtf(ss / dim + eps);
  const float4* w = reinterpret_cast<const float4*>(row < n_first ? w_a : w_b);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.x + 256 * u;
    if (i < n4) {
      const float4 g = __ldg(w + i);
      store4(out, out_bf16 & 1, row * ldo + 4LL * i, g.x * rn_(v[u].x * r, out_bf16), g.y * rn_(v[u].y * r, out_bf16),
             g.z * rn_(v[u].z * r, out_bf16), g.w * rn_(v[u].w * r, out_bf16));
    }
  }
}

__global__ void layernorm_kernel(const float* __restrict__ x, long long ldx, void* __restrict__ out,
                                 long long ldo, int out_bf16, const float* __restrict__ w,
                                 const float* __restrict__ b, long long rows, int dim, float eps, int seg_in,
                                 int seg_skip) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long orow = row;
  if (seg_in > 0) {
    const long long s = row / seg_in;
    const int local = static_cast<int>(row - s * seg_in);
    if (local < seg_skip) return;
    orow = s * (seg_in - seg_skip) + (local - seg_skip);
  }
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int n4 = dim >> 2;
  float s1 = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    s1 += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s1) / dim;
  float s2 = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    const float a = v.x - mean, bb = v.y - mean, c = v.z - mean, d = v.w - mean;
    s2 += a * a + bb * bb + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(s2) / dim + eps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    const float4 g = __ldg(w4 + i);
    const float4 bt = __ldg(b4 + i);
    store4(out, out_bf16, orow * ldo + 4LL * i, (v.x - mean) * rstd * g.x + bt.x, (v.y - mean) * rstd * g.y + bt.y,
           (v.z - mean) * rstd * g.z + bt.z, (v.w - mean) * rstd * g.w + bt.w);
  }
}

// Register-cached variant (same arithmetic and per-lane order as above; one global read per element).
template <int NV>
__global__ void __launch_bounds__(EW_THREADS)
layernorm_reg_kernel(const float* __restrict__ x, long long ldx, void* __restrict__ out, long long ldo, int out_bf16,
                     const float* __restrict__ w, const float* __restrict__ b, long long rows, float eps, int seg_in,
                     int seg_skip) {
  constexpr int dim = NV * 128;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long orow = row;
  if (seg_in > 0) {
    const long long s = row / seg_in;
    const int local = static_cast<int>(row - s * seg_in);
    if (local < seg_skip) return;
    orow = s * (seg_in - seg_skip) + (local - seg_skip);
  }
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) v[u] = xr[lane + 32 * u];
  float s1 = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) s1 += (v[u].x + v[u].y) + (v[u].z + v[u].w);
  const float mean = warp_sum(s1) / dim;
  float s2 = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const float a = v[u].x - mean, bb = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
    s2 += a * a + bb * bb + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(s2) / dim + eps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int i = lane + 32 * u;
    const float4 g = __ldg(w4 + i);
    const float4 bt = __ldg(b4 + i);
    store4(out, out_bf16, orow * ldo + 4LL * i, (v[u].x - mean) * rstd * g.x + bt.x, (v[u].y - mean) * rstd * g.y + bt.y,
           (v[u].z - mean) * rstd * g.z + bt.z, (v[u].w - mean) * rstd * g.w + bt.w);
  }
}

// ------------------------------------------------------------------------------------------------
// M-RoPE
// ------------------------------------------------------------------------------------------------
__global__ void mrope_table_kernel(const long long* __restrict__ pos, long long ld_pos,
                                   const float* __restrict__ inv_freq, float* __restrict__ cos_out,
                                   float* __restrict__ sin_out, long long rows, int half, int s0, int s1) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows * half) return;
  const long long t = i / half;
  const int j = static_cast<int>(i - t * half);
  const int axis = j < s0 ? 0 : (j < s0 + s1 ? 1 : 2);
  const float ang = static_cast<float>(pos[axis * ld_pos + t]) * inv_freq[j];
  float s, c;
  sincosf(ang, &s, &c);
  cos_out[i] = c;
  sin_out[i] = s;
}

// one warp per (row, head); head_dim = 128 -> 4 consecutive elements per lane; the rotation partner
// of element d is d +- 64, i.e. the same register slot of lane ^ 16.
__global__ void qknorm_mrope_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, long long rows,
                                    long long n_first, int n_q, int n_kv, const float* __restrict__ qw_a,
                                    const float* __restrict__ kw_a, const float* __restrict__ qw_b,
                                    const float* __restrict__ kw_b, const float* __restrict__ cos_tab,
                                    const float* __restrict__ sin_tab, float eps, int round_normed) {
  const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int heads = n_q + n_kv;
  if (wid >= rows * heads) return;
  const long long row = wid / heads;
  const int head = static_cast<int>(wid - row * heads);
  const int lane = threadIdx.x & 31;
  __nv_bfloat16* p = qkv + row * ld + head * 128 + lane * 4;
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 v01 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
  const __nv_bfloat162 v23 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  float v[4] = {__low2float(v01), __high2float(v01), __low2float(v23), __high2float(v23)};
  const float ss = warp_sum(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
  const float r = rsqrtf(ss / 128.0f + eps);
  const bool first = row < n_first;
  const float* w = head < n_q ? (first ? qw_a : qw_b) : (first ? kw_a : kw_b);
  const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane);
  const float gw[4] = {g.x, g.y, g.z, g.w};
  const int j0 = (lane & 15) * 4;  // frequency index of element 0 of this lane (d % 64)
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(cos_tab + row * 64 + j0));
  const float4 s4 = __ldg(reinterpret_cast<const float4*>(sin_tab + row * 64 + j0));
  const float cs[4] = {c4.x, c4.y, c4.z, c4.w};
  const float sn[4] = {s4.x, s4.y, s4.z, s4.w};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float n = v[e] * r;
    if (round_normed) n = bf16_round(n);
    n = gw[e] * n;
    if (round_normed == 2) n = bf16_round(n);
    const float partner = __shfl_xor_sync(0xffffffffu, n, 16);
    // rotate_half: first half gets -x2, second half gets +x1
    const float rot = lane < 16 ? -partner : partner;
    o[e] = round_normed == 2 ? bf16_round(n * bf16_round(cs[e])) + bf16_round(rot * bf16_round(sn[e]))
                             : n * cs[e] + rot * sn[e];
  }
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
}

// Row-per-warp variant for an even number of heads (the full model: 12 q + 2 k heads -> NU = 7): lane l owns the
// 16-byte chunks l, l+32, ... of the row's q|k columns, i.e. per round u the 8 elements 8k..8k+7 (k = l % 16) of head
// 2u + l/16.  All NU loads of a lane are in flight at once, the row's cos/sin and the norm weights are read once per
// row instead of once per head (that re-read made the per-head kernel above run at 46 % of the HBM roofline), the
// per-head sum of squares is a half-warp reduction and the rotate_half partner (d +- 64) is lane ^ 8.
template <int NU>
__global__ void __launch_bounds__(EW_THREADS)
qknorm_mrope_row_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, long long rows, long long n_first, int n_q,
                        const float* __restrict__ qw_a, const float* __restrict__ kw_a,
                        const float* __restrict__ qw_b, const float* __restrict__ kw_b,
                        const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, float eps,
                        int round_normed) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int k = lane & 15, hsel = lane >> 4;
  uint4* p = reinterpret_cast<uint4*>(qkv + row * ld) + lane;
  uint4 raw[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) raw[u] = p[32 * u];
  const bool first = row < n_first;
  const float4* qw = reinterpret_cast<const float4*>(first ? qw_a : qw_b) + 2 * k;
  const float4* kw = reinterpret_cast<const float4*>(first ? kw_a : kw_b) + 2 * k;
  const float4 q0 = __ldg(qw), q1 = __ldg(qw + 1), k0 = __ldg(kw), k1 = __ldg(kw + 1);
  const float wq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  const float wk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
  const float4* ct = reinterpret_cast<const float4*>(cos_tab + row * 64 + 8 * (k & 7));
  const float4* st = reinterpret_cast<const float4*>(sin_tab + row * 64 + 8 * (k & 7));
  const float4 c0 = __ldg(ct), c1 = __ldg(ct + 1), s0 = __ldg(st), s1 = __ldg(st + 1);
  const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const bool is_q = 2 * u + hsel < n_q;
    const uint32_t rw[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
      v[2 * e] = __low2float(b2);
      v[2 * e + 1] = __high2float(b2);
    }
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);  // stays inside the half-warp
    const float r = rsqrtf(ss / 128.0f + eps);
    float o8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float n = v[e] * r;
      if (round_normed) n = bf16_round(n);
      n = (is_q ? wq[e] : wk[e]) * n;
      if (round_normed == 2) n = bf16_round(n);
      const float partner = __shfl_xor_sync(0xffffffffu, n, 8);
      const float rot = k < 8 ? -partner : partner;  // rotate_half: first half gets -x2, second half gets +x1
      o8[e] = round_normed == 2 ? bf16_round(n * bf16_round(cs[e])) + bf16_round(rot * bf16_round(sn[e]))
                                : n * cs[e] + rot * sn[e];
    }
    p[32 * u] = make_uint4(pack_bf16x2(o8[0], o8[1]), pack_bf16x2(o8[2], o8[3]), pack_bf16x2(o8[4], o8[5]),
                           pack_bf16x2(o8[6], o8[7]));
  }
}

// ------------------------------------------------------------------------------------------------
// DINO input side
// ------------------------------------------------------------------------------------------------
struct Norm3 {
  float mean[3];
  float std[3];
  int enabled;
};

__global__ void im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int n, int H,
                              int W, int patch, int k_pad, Norm3 nrm) {
  const int gh = H / patch, gw = W / patch;
  const long long total = (long long)n * gh * gw * k_pad;
  const int pp = patch * patch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
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
      // torchvision Normalize (g2vlm.py:950): sub then div in fp32 (IEEE-exact, same as the host op)
      if (nrm.enabled) v = __fdiv_rn(__fsub_rn(v, nrm.mean[c]), nrm.std[c]);
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

__global__ void dino_embed_kernel(const __nv_bfloat16* __restrict__ patch_emb, long long ld_patch,
                                  const float* __restrict__ cls, const float* __restrict__ reg,
                                  const float* __restrict__ pos, float* __restrict__ out, int n, int P,
                                  int n_reg, int dim) {
  const int S = 1 + n_reg + P;
  const int d4 = dim >> 2;
  const long long total = (long long)n * S * d4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
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
      const uint2 raw = *reinterpret_cast<const uint2*>(patch_emb + ((long long)im * P + t) * ld_patch + c);
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
      const float4 ps = *reinterpret_cast<const float4*>(pos + (long long)(1 + t) * dim + c);
      v = make_float4(__low2float(a) + ps.x, __high2float(a) + ps.y, __low2float(b) + ps.z, __high2float(b) + ps.w);
    }
    *reinterpret_cast<float4*>(out + row * dim + c) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// RoPE2D: one thread per rotation pair
// ------------------------------------------------------------------------------------------------
__global__ void rope2d_kernel(__nv_bfloat16* __restrict__ buf, long long ld, long long rows, int n_heads,
                              int head_stride, int head_dim, int tokens_per_view, int grid_w,
                              const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                              int bf16_ops) {
  const int quarter = head_dim >> 2;            // distinct frequencies per axis
  const int pairs = head_dim >> 1;              // rotation pairs per head
  const long long total = rows * n_heads * pairs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int pr = static_cast<int>(i % pairs);
    const long long rh = i / pairs;
    const int head = static_cast<int>(rh % n_heads);
    const long long row = rh / n_heads;
    const int tok = static_cast<int>(row % tokens_per_view);
    const int axis = pr / quarter;              // 0: y (first half of the head), 1: x
    const int j = pr - axis * quarter;
    const int position = axis == 0 ? tok / grid_w : tok % grid_w;
    const float c = cos_tab[position * quarter + j];
    const float s = sin_tab[position * quarter + j];
    __nv_bfloat16* p = buf + row * ld + (long long)head * head_stride + axis * (head_dim >> 1) + j;
    const float a = __bfloat162float(p[0]);
    const float b = __bfloat162float(p[quarter]);
    float oa, ob;
    if (bf16_ops) {
      oa = bf16_round(a * c) + bf16_round(-b * s);
      ob = bf16_round(b * c) + bf16_round(a * s);
    } else {
      oa = a * c - b * s;
      ob = b * c + a * s;
    }
    p[0] = __float2bfloat16_rn(oa);
    p[quarter] = __float2bfloat16_rn(ob);
  }
}

// vectorised variant (head_dim/4 a multiple of 8, e.g. 96 -> 24): each thread rotates 8 pairs with
// 16-byte loads/stores
__global__ void rope2d_vec8_kernel(__nv_bfloat16* __restrict__ buf, long long ld, long long rows, int n_heads,
                                   int head_stride, int head_dim, int tokens_per_view, int grid_w,
                                   const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                   int bf16_ops) {
  const int quarter = head_dim >> 2;
  const int chunks = quarter >> 3;              // 8-pair chunks per axis
  const long long total = rows * n_heads * 2 * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % chunks);
    long long r = i / chunks;
    const int axis = static_cast<int>(r & 1);
    r >>= 1;
    const int head = static_cast<int>(r % n_heads);
    const long long row = r / n_heads;
    const int tok = static_cast<int>(row % tokens_per_view);
    const int position = axis == 0 ? tok / grid_w : tok % grid_w;
    const float4* ct = reinterpret_cast<const float4*>(cos_tab + position * quarter + c * 8);
    const float4* st = reinterpret_cast<const float4*>(sin_tab + position * quarter + c * 8);
    const float4 c0 = __ldg(ct), c1 = __ldg(ct + 1), s0 = __ldg(st), s1 = __ldg(st + 1);
    const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    __nv_bfloat16* p = buf + row * ld + (long long)head * head_stride + axis * (head_dim >> 1) + c * 8;
    uint4 ra = *reinterpret_cast<const uint4*>(p);
    uint4 rb = *reinterpret_cast<const uint4*>(p + quarter);
    uint32_t* ua = reinterpret_cast<uint32_t*>(&ra);
    uint32_t* ub = reinterpret_cast<uint32_t*>(&rb);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&ua[e]);
      const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ub[e]);
      const float a[2] = {__low2float(a2), __high2float(a2)};
      const float b[2] = {__low2float(b2), __high2float(b2)};
      float oa[2], ob[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float cc = cs[2 * e + h], ss = sn[2 * e + h];
        if (bf16_ops) {
          oa[h] = bf16_round(a[h] * cc) + bf16_round(-b[h] * ss);
          ob[h] = bf16_round(b[h] * cc) + bf16_round(a[h] * ss);
        } else {
          oa[h] = a[h] * cc - b[h] * ss;
          ob[h] = b[h] * cc + a[h] * ss;
        }
      }
      ua[e] = pack_bf16x2(oa[0], oa[1]);
      ub[e] = pack_bf16x2(ob[0], ob[1]);
    }
    *reinterpret_cast<uint4*>(p) = ra;
    *reinterpret_cast<uint4*>(p + quarter) = rb;
  }
}

// vision rotary (Qwen2-VL ViT): rotate_half over the whole head, per-token cos/sin; one thread per pair
__global__ void rope_vision_kernel(__nv_bfloat16* __restrict__ buf, long long ld, long long rows, int n_heads,
                                   int head_stride, int head_dim, const float* __restrict__ cos_tab,
                                   const float* __restrict__ sin_tab) {
  const int half = head_dim >> 1;
  const long long total = rows * n_heads * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = static_cast<int>(i % half);
    const long long rh = i / half;
    const int head = static_cast<int>(rh % n_heads);
    const long long row = rh / n_heads;
    const float c = cos_tab[row * half + j], s = sin_tab[row * half + j];
    __nv_bfloat16* p = buf + row * ld + (long long)head * head_stride + j;
    const float a = __bfloat162float(p[0]), b = __bfloat162float(p[half]);
    p[0] = __float2bfloat16_rn(a * c - b * s);
    p[half] = __float2bfloat16_rn(b * c + a * s);
  }
}

// ------------------------------------------------------------------------------------------------
// heads
// ------------------------------------------------------------------------------------------------
__global__ void points_epilogue_kernel(const float* __restrict__ feat, long long ld_feat,
                                       const float* __restrict__ poses, float* __restrict__ out0,
                                       float* __restrict__ out1, int n, int H, int W, int patch, int mode) {
  const long long total = (long long)n * H * W;
  const int gw = W / patch, gh = H / patch;
  const int pp = patch * patch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = static_cast<int>(i % W);
    const long long t = i / W;
    const int Y = static_cast<int>(t % H);
    const int v = static_cast<int>(t / H);
    const long long row = (long long)v * gh * gw + (Y / patch) * gw + X / patch;
    const int sub = (Y % patch) * patch + X % patch;
    const float* f = feat + row * ld_feat + sub;
    if (mode == 2) {  // single-channel pixel shuffle (conf head)
      out0[i] = f[0];
      continue;
    }
    const float a = f[0], b = f[pp], c = f[2 * pp];
    if (mode == 0) {
      out0[3 * i] = a; out0[3 * i + 1] = b; out0[3 * i + 2] = c;
    } else {
      const float z = expf(c);
      const float lx = a * z, ly = b * z;
      out0[3 * i] = lx; out0[3 * i + 1] = ly; out0[3 * i + 2] = z;
      const float* P = poses + v * 16;
      out1[3 * i] = P[0] * lx + P[1] * ly + P[2] * z + P[3];
      out1[3 * i + 1] = P[4] * lx + P[5] * ly + P[6] * z + P[7];
      out1[3 * i + 2] = P[8] * lx + P[9] * ly + P[10] * z + P[11];
    }
  }
}

// one block per (view, 128-column slab): coalesced column-wise mean over the tokens
__global__ void mean_pool_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out,
                                 int tokens, int dim) {
  const int v = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= dim) return;
  const float* p = x + (long long)v * tokens * ldx + c;
  float acc = 0.f;
  for (int t = 0; t < tokens; ++t) acc += p[(long long)t * ldx];
  out[(long long)v * dim + c] = acc / tokens;
}

__global__ void split3_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out,
                              long long ldo, long long rows, int k) {
  const long long total = rows * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / k;
    const int c = static_cast<int>(i - r * k);
    const float v = x[r * ldx + c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * ldo + c;
    o[0] = hi;
    o[k] = hi;
    o[2 * k] = lo;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out,
                                     long long ldo, long long rows, int c4) {
  const long long total = rows * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const int c = static_cast<int>(i - r * c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    *reinterpret_cast<uint2*>(out + r * ldo + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ---- 3x3 nearest-rotation via Jacobi eigen-decomposition of A^T A (double precision) -------------
__device__ void jacobi_eig3(double a[3][3], double v[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(a[p][q]) < 1e-300) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
}

// one warp per view: 12 dot products of length dim (warp-reduced), lane 0 finishes the pose
__global__ void camera_pose_kernel(const float* __restrict__ feat, long long ldf, const float* __restrict__ w_t,
                                   const float* __restrict__ b_t, const float* __restrict__ w_r,
                                   const float* __restrict__ b_r, float* __restrict__ poses, int n, int dim) {
  const int view = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (view >= n) return;
  const int lane = threadIdx.x & 31;
  const float* f = feat + (long long)view * ldf;
  float acc[12];
#pragma unroll
  for (int o = 0; o < 12; ++o) acc[o] = 0.f;
  for (int k = lane; k < dim; k += 32) {
    const float x = f[k];
#pragma unroll
    for (int o = 0; o < 3; ++o) acc[o] += x * w_t[o * dim + k];
#pragma unroll
    for (int o = 0; o < 9; ++o) acc[3 + o] += x * w_r[o * dim + k];
  }
#pragma unroll
  for (int o = 0; o < 12; ++o) acc[o] = warp_sum(acc[o]);
  if (lane != 0) return;
  const float t[3] = {acc[0] + b_t[0], acc[1] + b_t[1], acc[2] + b_t[2]};
  // m [3x3] row-major; A = row-normalised m (F.normalize(p=2, dim=-1), eps 1e-12)
  double A[3][3];
  for (int i = 0; i < 3; ++i) {
    float r[3];
    for (int j = 0; j < 3; ++j) r[j] = acc[3 + 3 * i + j] + b_r[3 * i + j];
    const float nrm = fmaxf(sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]), 1e-12f);
    for (int j = 0; j < 3; ++j) A[i][j] = static_cast<double>(r[j] / nrm);
  }
  // The reference computes svd(A^T) = U S V^T and returns R = V diag(1,1,det(V U^T)) U^T, i.e. the
  // rotation nearest to A. With A = P S Q^T: R = p1 q1^T + p2 q2^T + (p1 x p2)(q1 x q2)^T.
  double ata[3][3], Q[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) ata[i][j] = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j];
  jacobi_eig3(ata, Q);
  int i1 = 0;
  for (int i = 1; i < 3; ++i)
    if (ata[i][i] > ata[i1][i1]) i1 = i;
  int i2 = -1;
  for (int i = 0; i < 3; ++i)
    if (i != i1 && (i2 < 0 || ata[i][i] > ata[i2][i2])) i2 = i;
  double q1[3] = {Q[0][i1], Q[1][i1], Q[2][i1]}, q2[3] = {Q[0][i2], Q[1][i2], Q[2][i2]};
  double p1[3], p2[3];
  for (int i = 0; i < 3; ++i) {
    p1[i] = A[i][0] * q1[0] + A[i][1] * q1[1] + A[i][2] * q1[2];
    p2[i] = A[i][0] * q2[0] + A[i][1] * q2[1] + A[i][2] * q2[2];
  }
  double n1 = sqrt(p1[0] * p1[0] + p1[1] * p1[1] + p1[2] * p1[2]);
  for (int i = 0; i < 3; ++i) p1[i] /= (n1 > 0 ? n1 : 1.0);
  const double d12 = p1[0] * p2[0] + p1[1] * p2[1] + p1[2] * p2[2];
  for (int i = 0; i < 3; ++i) p2[i] -= d12 * p1[i];
  double n2 = sqrt(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
  for (int i = 0; i < 3; ++i) p2[i] /= (n2 > 0 ? n2 : 1.0);
  const double p3[3] = {p1[1] * p2[2] - p1[2] * p2[1], p1[2] * p2[0] - p1[0] * p2[2], p1[0] * p2[1] - p1[1] * p2[0]};
  const double q3[3] = {q1[1] * q2[2] - q1[2] * q2[1], q1[2] * q2[0] - q1[0] * q2[2], q1[0] * q2[1] - q1[1] * q2[0]};
  float* P = poses + view * 16;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) P[4 * i + j] = static_cast<float>(p1[i] * q1[j] + p2[i] * q2[j] + p3[i] * q3[j]);
    P[4 * i + 3] = t[i];
  }
  P[12] = 0.f; P[13] = 0.f; P[14] = 0.f; P[15] = 1.f;
}

// ------------------------------------------------------------------------------------------------
// argmax over the vocabulary (greedy decode): one block per row, lowest index wins ties
// ------------------------------------------------------------------------------------------------
// blockIdx.y slices the vocabulary (ARGMAX_SLICES partial results per row, merged by the last slice to finish)
// torch.argmax semantics: the maximum, ties -> lowest index, NaN counts as the maximum (first NaN wins) so that a
// numerical blow-up surfaces in the generated ids instead of being skipped.
__device__ __forceinline__ bool argmax_better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (vn || v == bv) return i < bi;
  return v > bv;
}

constexpr int ARGMAX_SLICES = 8;   // one thread-block cluster per row: the slices merge through distributed shared memory
constexpr int ARGMAX_THREADS = 512;

// grid (rows, SLICES) with cluster dims (1, SLICES, 1) for large vocabularies, grid (rows, 1) otherwise.  No global
// scratch: concurrent launches on different streams (two generations, a graph replay next to an eager step) are
// independent.  16-byte loads when the row is 16-byte aligned.
template <int SLICES>
__global__ void __launch_bounds__(ARGMAX_THREADS)
argmax_bf16_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, int vocab, long long* __restrict__ out) {
  const __nv_bfloat16* row = logits + blockIdx.x * ld;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  const int per = ((vocab + SLICES - 1) / SLICES + 7) & ~7;
  const int lo = blockIdx.y * per, hi = min(vocab, lo + per);
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    for (int i = lo + threadIdx.x * 8; i < hi; i += ARGMAX_THREADS * 8) {
      if (i + 8 <= hi) {
        const uint4 raw = *reinterpret_cast<const uint4*>(row + i);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = __uint_as_float(w[j] << 16), v1 = __uint_as_float(w[j] & 0xffff0000u);
          if (argmax_better(v0, i + 2 * j, best, idx)) { best = v0; idx = i + 2 * j; }
          if (argmax_better(v1, i + 2 * j + 1, best, idx)) { best = v1; idx = i + 2 * j + 1; }
        }
      } else {
        for (int j = i; j < hi; ++j) {
          const float v = __bfloat162float(row[j]);
          if (argmax_better(v, j, best, idx)) { best = v; idx = j; }
        }
      }
    }
  } else {
    for (int i = lo + threadIdx.x; i < hi; i += ARGMAX_THREADS) {
      const float v = __bfloat162float(row[i]);
      if (argmax_better(v, i, best, idx)) { best = v; idx = i; }
    }
  }
  __shared__ float sb[32];
  __shared__ int si[32];
  __shared__ float cb[SLICES];   // rank 0's copy receives every slice's winner
  __shared__ int ci[SLICES];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (argmax_better(ob, oi, best, idx)) { best = ob; idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (ARGMAX_THREADS >> 5) ? sb[threadIdx.x] : -INFINITY;
    idx = threadIdx.x < (ARGMAX_THREADS >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (argmax_better(ob, oi, best, idx)) { best = ob; idx = oi; }
    }
  }
  if constexpr (SLICES == 1) {
    if (threadIdx.x == 0) out[blockIdx.x] = idx == 0x7fffffff ? 0 : idx;
  } else {
    if (threadIdx.x == 0) {   // publish this slice's winner in rank 0's shared memory
      const uint32_t rank = cluster_ctarank();
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(mapa_shared(smem_u32(&cb[rank]), 0)), "f"(best) : "memory");
      asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(mapa_shared(smem_u32(&ci[rank]), 0)), "r"(idx) : "memory");
    }
    cluster_sync_all();
    if (cluster_ctarank() == 0 && threadIdx.x == 0) {
      float b = cb[0];
      int bi = ci[0];
      for (int s2 = 1; s2 < SLICES; ++s2)
        if (argmax_better(cb[s2], ci[s2], b, bi)) { b = cb[s2]; bi = ci[s2]; }
      out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// PLY packing: order-preserving compaction of finite points (3 passes: count, scan, scatter)
// ------------------------------------------------------------------------------------------------
constexpr int PLY_BLOCK = 1024;

__device__ __forceinline__ bool point_finite(const float* p) {
  return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]);
}

__global__ void ply_count_kernel(const float* __restrict__ pts, long long n, int* __restrict__ block_counts, int keep_all) {
  const long long i = blockIdx.x * (long long)PLY_BLOCK + threadIdx.x;
  const int ok = (i < n) && (keep_all || point_finite(pts + 3 * i));
  const int c = __syncthreads_count(ok);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of the per-block counts by ONE block (<= a few thousand entries), total -> n_valid
__global__ void ply_scan_kernel(int* __restrict__ block_counts, int n_blocks, long long* __restrict__ n_valid) {
  __shared__ long long carry;
  __shared__ int warp_tot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = threadIdx.x < (blockDim.x >> 5) ? warp_tot[threadIdx.x] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_tot[threadIdx.x] = w;  // inclusive scan of warp totals
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    const long long excl = carry + warp_off + x - v;
    const int chunk_total = warp_tot[(blockDim.x >> 5) - 1];
    __syncthreads();
    // block offsets fit in int32 for any scene this path can hold (< 2^31 points)
    if (i < n_blocks) block_counts[i] = static_cast<int>(excl);
    if (threadIdx.x == 0) carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_valid = carry;
}

__global__ void ply_scatter_kernel(const float* __restrict__ pts, const float* __restrict__ img, long long n,
                                   int hw, const int* __restrict__ block_offsets, uint8_t* __restrict__ out, int keep_all) {
  __shared__ int warp_cnt[32];
  const long long i = blockIdx.x * (long long)PLY_BLOCK + threadIdx.x;
  const int ok = (i < n) && (keep_all || point_finite(pts + 3 * i));
  const unsigned ballot = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    int w = warp_cnt[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_cnt[lane] = w;
  }
  __syncthreads();
  if (!ok) return;
  const long long dst = block_offsets[blockIdx.x] + (warp ? warp_cnt[warp - 1] : 0) + __popc(ballot & ((1u << lane) - 1));
  uint8_t* rec = out + dst * 27;
  const float* p = pts + 3 * i;
  const long long v = i / hw, pix = i - v * hw;
  const float* c = img + v * 3 * hw + pix;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double d = static_cast<double>(p[k]);
    memcpy(rec + 8 * k, &d, 8);  // records are 27 bytes: unaligned doubles
    // Open3D's PLY writer (the reference's output path, g2vlm_utils.py:146) converts with utility::ColorToUint8 =
    // round(clamp(c, 0, 1) * 255): colours are float32 k/255, so c * 255 often lands at k - 1e-7 and must ROUND to k
    const double col = fmin(1.0, fmax(0.0, static_cast<double>(c[(long long)k * hw]))) * 255.0;
    rec[24 + k] = static_cast<uint8_t>(__double2int_rn(col));
  }
}

// ------------------------------------------------------------------------------------------------
// Pillow-exact 8-bit Lanczos resampling (two passes, 22-bit fixed-point coefficients)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t clip8_fixed(int acc) {
  const int v = acc >> 22;  // arithmetic shift: floor, like Pillow's clip8 lookup
  return static_cast<uint8_t>(min(255, max(0, v)));
}

// horizontal pass: tmp[y][xx][c] from src[y][xmin .. xmin+n)[c]; one thread per (y, xx)
__global__ void resize_h_u8_kernel(const uint8_t* __restrict__ src, long long pitch, int H, int out_w,
                                   const int* __restrict__ bounds, const int* __restrict__ coef, int ksize,
                                   uint8_t* __restrict__ tmp) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)H * out_w) return;
  const int y = static_cast<int>(i / out_w), xx = static_cast<int>(i - (long long)y * out_w);
  const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
  const int* k = coef + (long long)xx * ksize;
  const uint8_t* row = src + y * pitch + 3LL * xmin;
  int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
  for (int x = 0; x < n; ++x) {
    const int w = __ldg(k + x);
    a0 += row[3 * x] * w;
    a1 += row[3 * x + 1] * w;
    a2 += row[3 * x + 2] * w;
  }
  uint8_t* o = tmp + 3 * i;
  o[0] = clip8_fixed(a0);
  o[1] = clip8_fixed(a1);
  o[2] = clip8_fixed(a2);
}

// vertical pass: out[yy][xx][c] from in[ymin .. ymin+n)[xx][c]; one thread per (yy, xx); optional fp32 CHW output
__global__ void resize_v_u8_kernel(const uint8_t* __restrict__ in, long long pitch, int out_h, int out_w,
                                   const int* __restrict__ bounds, const int* __restrict__ coef, int ksize,
                                   uint8_t* __restrict__ out_u8, float* __restrict__ out_f32) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)out_h * out_w) return;
  const int yy = static_cast<int>(i / out_w), xx = static_cast<int>(i - (long long)yy * out_w);
  uint8_t r, g, b;
  if (bounds != nullptr) {
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int* k = coef + (long long)yy * ksize;
    const uint8_t* col = in + ymin * pitch + 3LL * xx;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int y = 0; y < n; ++y) {
      const int w = __ldg(k + y);
      const uint8_t* px = col + y * pitch;
      a0 += px[0] * w;
      a1 += px[1] * w;
      a2 += px[2] * w;
    }
    r = clip8_fixed(a0); g = clip8_fixed(a1); b = clip8_fixed(a2);
  } else {  // the height does not change: Pillow skips the vertical pass
    const uint8_t* px = in + yy * pitch + 3LL * xx;
    r = px[0]; g = px[1]; b = px[2];
  }
  if (out_u8 != nullptr) {
    uint8_t* o = out_u8 + 3 * i;
    o[0] = r; o[1] = g; o[2] = b;
  }
  if (out_f32 != nullptr) {  // ToTensor: uint8 -> float, / 255 (IEEE division, as torch)
    const long long plane = (long long)out_h * out_w;
    out_f32[i] = static_cast<float>(r) / 255.0f;
    out_f32[plane + i] = static_cast<float>(g) / 255.0f;
    out_f32[2 * plane + i] = static_cast<float>(b) / 255.0f;
  }
}

}  // namespace g2

// =================================================================================================
// C ABI
// =================================================================================================
using namespace g2;

#define G2_LAUNCH_CHECK() G2_CUDA_OK(cudaGetLastError())
#define G2_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int g2vlm_gather_rows(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch,
                                 const int64_t* idx, int64_t n_rows, int64_t row_bytes, int32_t scatter,
                                 void* stream) {
  G2_REQUIRE(src && dst, "gather_rows: null tensor");
  G2_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0 && src_pitch % 16 == 0 && dst_pitch % 16 == 0,
             "gather_rows: row_bytes and pitches must be multiples of 16");
  G2_REQUIRE(G2_ALIGNED16(src) && G2_ALIGNED16(dst), "gather_rows: pointers must be 16-byte aligned");
  if (n_rows <= 0) return G2VLM_OK;
  const int chunks = static_cast<int>(row_bytes / 16);
  const unsigned grid = blocks_for(n_rows * chunks, EW_THREADS * 4);
  gather_rows_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)src, src_pitch, (uint8_t*)dst, dst_pitch, (const long long*)idx, n_rows, chunks, scatter);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_rmsnorm_routed(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t out_bf16,
                                    const float* w_a, const float* w_b, int64_t rows, int64_t n_first,
                                    int32_t dim, float eps, void* stream) {
  G2_REQUIRE(x && out && w_a && w_b, "rmsnorm: null tensor");
  G2_REQUIRE(dim > 0 && dim % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "rmsnorm: dim/ld must be multiples of 4");
  G2_REQUIRE(G2_ALIGNED16(x) && G2_ALIGNED16(out) && G2_ALIGNED16(w_a) && G2_ALIGNED16(w_b), "rmsnorm: alignment");
  if (rows <= 0) return G2VLM_OK;
  if (rows <= 64 && dim <= 4096) {
    rmsnorm_routed_block_kernel<<<static_cast<unsigned>(rows), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, out, ldo, out_bf16, w_a, w_b, n_first, dim, eps);
    G2_LAUNCH_CHECK();
    return G2VLM_OK;
  }
  const unsigned grid = blocks_for(rows, EW_THREADS / 32);
  if (dim == 1536)
    rmsnorm_routed_reg_kernel<12><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, out_bf16, w_a, w_b,
                                                                                rows, n_first, eps);
  else if (dim == 1024)
    rmsnorm_routed_reg_kernel<8><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, out_bf16, w_a, w_b,
                                                                               rows, n_first, eps);
  else
    rmsnorm_routed_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, out_bf16, w_a, w_b, rows,
                                                                        n_first, dim, eps);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_layernorm(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t out_bf16,
                               const float* w, const float* b, int64_t rows, int32_t dim, float eps,
                               int32_t seg_in, int32_t seg_skip, void* stream) {
  G2_REQUIRE(x && out && w && b, "layernorm: null tensor");
  G2_REQUIRE(dim > 0 && dim % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "layernorm: dim/ld must be multiples of 4");
  G2_REQUIRE(G2_ALIGNED16(x) && G2_ALIGNED16(out) && G2_ALIGNED16(w) && G2_ALIGNED16(b), "layernorm: alignment");
  G2_REQUIRE(seg_in >= 0 && seg_skip >= 0 && (seg_in == 0 || seg_skip < seg_in), "layernorm: bad segment spec");
  if (rows <= 0) return G2VLM_OK;
  const unsigned grid = blocks_for(rows, EW_THREADS / 32);
  if (dim == 1024)
    layernorm_reg_kernel<8><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, out_bf16, w, b, rows, eps,
                                                                          seg_in, seg_skip);
  else if (dim == 1536)
    layernorm_reg_kernel<12><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, out_bf16, w, b, rows, eps,
                                                                           seg_in, seg_skip);
  else
    layernorm_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, out, ldo, out_bf16, w, b, rows, dim, eps,
                                                                   seg_in, seg_skip);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_mrope_table(const int64_t* position_ids, int64_t ld_pos, const float* inv_freq,
                                 float* cos_out, float* sin_out, int64_t rows, int32_t half, int32_t s0,
                                 int32_t s1, void* stream) {
  G2_REQUIRE(position_ids && inv_freq && cos_out && sin_out, "mrope_table: null tensor");
  G2_REQUIRE(half > 0 && s0 >= 0 && s1 >= 0 && s0 + s1 <= half, "mrope_table: bad sections");
  if (rows <= 0) return G2VLM_OK;
  mrope_table_kernel<<<blocks_for(rows * half, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const long long*)position_ids, ld_pos, inv_freq, cos_out, sin_out, rows, half, s0, s1);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_qknorm_mrope(void* qkv, int64_t ld, int64_t rows, int64_t n_first, int32_t n_q_heads,
                                  int32_t n_kv_heads, int32_t head_dim, const float* qw_a, const float* kw_a,
                                  const float* qw_b, const float* kw_b, const float* cos_tab,
                                  const float* sin_tab, float eps, int32_t round_normed, void* stream) {
  G2_REQUIRE(qkv && qw_a && kw_a && qw_b && kw_b && cos_tab && sin_tab, "qknorm_mrope: null tensor");
  G2_REQUIRE(head_dim == 128, "qknorm_mrope: head_dim must be 128 (mrope_section sums to 64)");
  G2_REQUIRE(ld % 8 == 0 && G2_ALIGNED16(qkv) && G2_ALIGNED16(cos_tab) && G2_ALIGNED16(sin_tab), "qknorm_mrope: alignment");
  if (rows <= 0) return G2VLM_OK;
  if (n_q_heads + n_kv_heads == 14) {
    qknorm_mrope_row_kernel<7><<<blocks_for(rows, EW_THREADS / 32), EW_THREADS, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)qkv, ld, rows, n_first, n_q_heads, qw_a, kw_a, qw_b, kw_b, cos_tab, sin_tab, eps, round_normed);
    G2_LAUNCH_CHECK();
    return G2VLM_OK;
  }
  const long long warps = rows * (n_q_heads + n_kv_heads);
  qknorm_mrope_kernel<<<blocks_for(warps, EW_THREADS / 32), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (__nv_bfloat16*)qkv, ld, rows, n_first, n_q_heads, n_kv_heads, qw_a, kw_a, qw_b, kw_b, cos_tab, sin_tab, eps,
      round_normed);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_im2col_patches(const float* images, void* out, int32_t n, int32_t H, int32_t W, int32_t patch,
                                    int32_t k_pad, const float* mean3, const float* std3, void* stream) {
  G2_REQUIRE(images && out, "im2col: null tensor");
  G2_REQUIRE(patch > 0 && H % patch == 0 && W % patch == 0 && k_pad >= 3 * patch * patch, "im2col: bad geometry");
  if (n <= 0) return G2VLM_OK;
  G2_REQUIRE((mean3 == nullptr) == (std3 == nullptr), "im2col: mean3 and std3 must be given together");
  Norm3 nrm;
  nrm.enabled = mean3 != nullptr;
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = mean3 ? mean3[c] : 0.f;
    nrm.std[c] = std3 ? std3[c] : 1.f;
  }
  const long long total = (long long)n * (H / patch) * (W / patch) * k_pad;
  im2col_kernel<<<blocks_for(total, EW_THREADS * 4), EW_THREADS, 0, (cudaStream_t)stream>>>(
      images, (__nv_bfloat16*)out, n, H, W, patch, k_pad, nrm);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_dino_embed(const void* patch_emb, int64_t ld_patch, const float* cls, const float* reg,
                                const float* pos, float* out, int32_t n, int32_t P, int32_t n_reg, int32_t dim,
                                void* stream) {
  G2_REQUIRE(patch_emb && cls && reg && pos && out, "dino_embed: null tensor");
  G2_REQUIRE(dim % 4 == 0 && ld_patch % 4 == 0, "dino_embed: dim must be a multiple of 4");
  G2_REQUIRE(G2_ALIGNED16(cls) && G2_ALIGNED16(reg) && G2_ALIGNED16(pos) && G2_ALIGNED16(out) &&
                 (reinterpret_cast<uintptr_t>(patch_emb) & 7) == 0, "dino_embed: alignment");
  if (n <= 0) return G2VLM_OK;
  const long long total = (long long)n * (1 + n_reg + P) * (dim / 4);
  dino_embed_kernel<<<blocks_for(total, EW_THREADS * 2), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)patch_emb, ld_patch, cls, reg, pos, out, n, P, n_reg, dim);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_rope2d(void* buf, int64_t ld, int64_t rows, int32_t n_heads_total, int32_t head_stride,
                            int32_t head_dim, int32_t tokens_per_view, int32_t grid_w, const float* cos_tab,
                            const float* sin_tab, int32_t bf16_ops, void* stream) {
  G2_REQUIRE(buf && cos_tab && sin_tab, "rope2d: null tensor");
  G2_REQUIRE(head_dim > 0 && head_dim % 4 == 0 && head_dim <= head_stride, "rope2d: head_dim % 4, <= head_stride");
  G2_REQUIRE(tokens_per_view > 0 && grid_w > 0, "rope2d: bad grid");
  if (rows <= 0) return G2VLM_OK;
  const int quarter = head_dim / 4;
  if (quarter % 8 == 0 && ld % 8 == 0 && head_stride % 8 == 0 && G2_ALIGNED16(buf) && G2_ALIGNED16(cos_tab) &&
      G2_ALIGNED16(sin_tab)) {
    const long long total = rows * n_heads_total * 2 * (quarter / 8);
    rope2d_vec8_kernel<<<blocks_for(total, EW_THREADS * 2), EW_THREADS, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)buf, ld, rows, n_heads_total, head_stride, head_dim, tokens_per_view, grid_w, cos_tab,
        sin_tab, bf16_ops);
  } else {
    const long long total = rows * n_heads_total * (head_dim / 2);
    rope2d_kernel<<<blocks_for(total, EW_THREADS * 4), EW_THREADS, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)buf, ld, rows, n_heads_total, head_stride, head_dim, tokens_per_view, grid_w, cos_tab,
        sin_tab, bf16_ops);
  }
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_points_epilogue(const float* feat, int64_t ld_feat, const float* poses, float* out0,
                                     float* out1, int32_t n, int32_t H, int32_t W, int32_t patch, int32_t mode,
                                     void* stream) {
  G2_REQUIRE(feat && out0, "points_epilogue: null tensor");
  G2_REQUIRE(mode >= 0 && mode <= 2, "points_epilogue: mode must be 0, 1 or 2");
  G2_REQUIRE(mode != 1 || (poses && out1), "points_epilogue: mode 1 needs poses and out1");
  G2_REQUIRE(patch > 0 && H % patch == 0 && W % patch == 0, "points_epilogue: bad geometry");
  if (n <= 0) return G2VLM_OK;
  points_epilogue_kernel<<<blocks_for((long long)n * H * W, EW_THREADS * 2), EW_THREADS, 0, (cudaStream_t)stream>>>(
      feat, ld_feat, poses, out0, out1, n, H, W, patch, mode);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_mean_pool(const float* x, int64_t ldx, float* out, int32_t n_views, int32_t tokens,
                               int32_t dim, void* stream) {
  G2_REQUIRE(x && out && tokens > 0 && dim > 0, "mean_pool: bad arguments");
  if (n_views <= 0) return G2VLM_OK;
  mean_pool_kernel<<<dim3(n_views, cdiv(dim, 128)), 128, 0, (cudaStream_t)stream>>>(x, ldx, out, tokens, dim);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_split3_f32(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t k,
                                void* stream) {
  G2_REQUIRE(x && out && k > 0 && ldo >= 3LL * k, "split3: bad arguments");
  if (rows <= 0) return G2VLM_OK;
  split3_kernel<<<blocks_for(rows * k, EW_THREADS * 4), EW_THREADS, 0, (cudaStream_t)stream>>>(
      x, ldx, (__nv_bfloat16*)out, ldo, rows, k);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_cast_f32_to_bf16(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t rows,
                                      int32_t cols, void* stream) {
  G2_REQUIRE(x && out && cols > 0 && cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "cast: cols/ld must be multiples of 4");
  G2_REQUIRE(G2_ALIGNED16(x) && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "cast: alignment");
  if (rows <= 0) return G2VLM_OK;
  cast_f32_bf16_kernel<<<blocks_for(rows * (cols / 4), EW_THREADS * 4), EW_THREADS, 0, (cudaStream_t)stream>>>(
      x, ldx, (__nv_bfloat16*)out, ldo, rows, cols / 4);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_camera_pose(const float* feat, int64_t ldf, const float* w_t, const float* b_t,
                                 const float* w_r, const float* b_r, float* poses, int32_t n, int32_t dim,
                                 void* stream) {
  G2_REQUIRE(feat && w_t && b_t && w_r && b_r && poses && dim > 0, "camera_pose: bad arguments");
  if (n <= 0) return G2VLM_OK;
  camera_pose_kernel<<<blocks_for(n, 4), 128, 0, (cudaStream_t)stream>>>(feat, ldf, w_t, b_t, w_r, b_r, poses, n, dim);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_ply_pack(const float* points, const float* images, int32_t n_views, int32_t H, int32_t W,
                              void* out, int32_t* block_counts, int64_t* n_valid, int32_t filter_nonfinite, void* stream) {
  G2_REQUIRE(points && images && out && block_counts && n_valid, "ply_pack: null tensor");
  const int keep_all = filter_nonfinite ? 0 : 1;
  G2_REQUIRE(n_views >= 0 && H > 0 && W > 0, "ply_pack: bad geometry");
  const long long n = (long long)n_views * H * W;
  G2_REQUIRE(n < (1LL << 31), "ply_pack: too many points");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_blocks = static_cast<int>((n + PLY_BLOCK - 1) / PLY_BLOCK);
  if (n_blocks > 0) {
    ply_count_kernel<<<n_blocks, PLY_BLOCK, 0, st>>>(points, n, block_counts, keep_all);
    G2_LAUNCH_CHECK();
  }
  ply_scan_kernel<<<1, 1024, 0, st>>>(block_counts, n_blocks, (long long*)n_valid);
  G2_LAUNCH_CHECK();
  if (n_blocks > 0) {
    ply_scatter_kernel<<<n_blocks, PLY_BLOCK, 0, st>>>(points, images, n, H * W, block_counts, (uint8_t*)out, keep_all);
    G2_LAUNCH_CHECK();
  }
  return G2VLM_OK;
}

extern "C" int g2vlm_argmax_bf16(const void* logits, int64_t ld, int64_t rows, int32_t vocab, int64_t* out,
                                 void* stream) {
  G2_REQUIRE(logits && out && vocab > 0, "argmax: bad arguments");
  if (rows <= 0) return G2VLM_OK;
  G2_REQUIRE(rows < (1LL << 31), "argmax: too many rows");
  if (vocab >= 16384) {
    // large vocabularies: one cluster of 8 CTAs per row (a single block is latency-bound: 72 us for 151 936)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(rows), ARGMAX_SLICES);
    cfg.blockDim = dim3(ARGMAX_THREADS);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 1;
    attr.val.clusterDim.y = ARGMAX_SLICES;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    G2_CUDA_OK(cudaLaunchKernelEx(&cfg, argmax_bf16_kernel<ARGMAX_SLICES>, (const __nv_bfloat16*)logits, (long long)ld,
                                  (int)vocab, (long long*)out));
  } else {
    argmax_bf16_kernel<1><<<dim3(static_cast<unsigned>(rows), 1), ARGMAX_THREADS, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)logits, ld, vocab, (long long*)out);
    G2_LAUNCH_CHECK();
  }
  return G2VLM_OK;
}

extern "C" int g2vlm_rope_vision(void* buf, int64_t ld, int64_t rows, int32_t n_heads_total, int32_t head_stride,
                                 int32_t head_dim, const float* cos_tab, const float* sin_tab, void* stream) {
  G2_REQUIRE(buf && cos_tab && sin_tab, "rope_vision: null tensor");
  G2_REQUIRE(head_dim > 0 && head_dim % 2 == 0 && head_dim <= head_stride, "rope_vision: bad head_dim");
  if (rows <= 0) return G2VLM_OK;
  const long long total = rows * n_heads_total * (head_dim / 2);
  rope_vision_kernel<<<blocks_for(total, EW_THREADS * 4), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (__nv_bfloat16*)buf, ld, rows, n_heads_total, head_stride, head_dim, cos_tab, sin_tab);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}

extern "C" int g2vlm_resize_lanczos_u8(const uint8_t* src, int32_t H, int32_t W, int64_t src_pitch,
                                       const int32_t* hbounds, const int32_t* hcoef, int32_t hk,
                                       const int32_t* vbounds, const int32_t* vcoef, int32_t vk, uint8_t* tmp,
                                       int32_t out_h, int32_t out_w, uint8_t* out_u8, float* out_f32, void* stream) {
  using namespace g2;
  G2_REQUIRE(src != nullptr && H > 0 && W > 0 && out_h > 0 && out_w > 0, "resize: bad geometry");
  G2_REQUIRE(src_pitch >= 3LL * W, "resize: src_pitch smaller than a row");
  G2_REQUIRE((hbounds == nullptr) == (hcoef == nullptr) && (vbounds == nullptr) == (vcoef == nullptr),
             "resize: bounds and coefficients must be given together");
  G2_REQUIRE(hbounds != nullptr || W == out_w, "resize: the width changes but no horizontal tables were given");
  G2_REQUIRE(vbounds != nullptr || H == out_h, "resize: the height changes but no vertical tables were given");
  G2_REQUIRE(hbounds == nullptr || (tmp != nullptr && hk > 0), "resize: the horizontal pass needs tmp and hk > 0");
  G2_REQUIRE(vbounds == nullptr || vk > 0, "resize: vk must be positive");
  G2_REQUIRE(out_u8 != nullptr || out_f32 != nullptr, "resize: no output buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* mid = src;
  long long mid_pitch = src_pitch;
  if (hbounds != nullptr) {
    resize_h_u8_kernel<<<blocks_for((long long)H * out_w, EW_THREADS), EW_THREADS, 0, st>>>(src, src_pitch, H, out_w,
                                                                                          hbounds, hcoef, hk, tmp);
    G2_LAUNCH_CHECK();
    mid = tmp;
    mid_pitch = 3LL * out_w;
  }
  resize_v_u8_kernel<<<blocks_for((long long)out_h * out_w, EW_THREADS), EW_THREADS, 0, st>>>(
      mid, mid_pitch, out_h, out_w, vbounds, vcoef, vk, out_u8, out_f32);
  G2_LAUNCH_CHECK();
  return G2VLM_OK;
}
