// g2vlm_b200 — decode-shaped kernels of the chat path: with ONE query token per step the
// work is HBM-bound (3.1 GB of bf16 weights + the KV cache stream through per token), so these are plain
// bandwidth kernels, not tensor-core tiles:
//   gemv_bf16_kernel<EPI>     out[m, n] = epilogue(sum_k x[m,k] * w[n,k]), m <= 8 rows, one warp per output
//                             column, 16-byte loads of the weight row; same epilogues / rounding points as the
//                             tcgen05 GEMM (g2vlm_gemm_bf16 routes here when the call has <= 8 rows)
//   attn_decode_split_kernel  one query token, GQA: keys split across CTAs (flash-decoding); per CTA
//                             scores -> softmax statistics -> P.V for the 6 query heads of a KV head
//   attn_decode_merge_kernel  log-sum-exp merge of the splits
#include "common.cuh"

namespace g2 {

constexpr int GEMV_MAX_ROWS = 8;
constexpr int GEMV_WARPS = 8;

struct GemvParams {
  const __nv_bfloat16* x;  // [rows, K]
  long long ldx;
  const __nv_bfloat16* w;  // [N(, interleaved), K] rows of the selected expert
  long long ldw;
  int rows, N, K;
  uint32_t flags;
  int use_scale;
  void* out;
  long long ldo;
  const float* bias;   // already offset to the expert
  const float* scale;
  const float* residual;
  long long ldr;
  long long row0;      // first output row
};

__device__ __forceinline__ float gelu_erf_d(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void dot8(const uint4& a, const uint4& b, float& acc) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __bfloat1622float2(pa[i]);
    const float2 fb = __bfloat1622float2(pb[i]);
    acc = fmaf(fa.x, fb.x, acc);
    acc = fmaf(fa.y, fb.y, acc);
  }
}

// epilogue of one output element (row m of the call, output column n): same rounding points as the tcgen05 GEMM
template <int EPI>
__device__ __forceinline__ void gemv_store(const GemvParams& p, int n, int m, float f, float f_up) {
  const long long row = p.row0 + m;
  if constexpr (EPI == G2VLM_EPI_SWIGLU_BF16) {
    const float g = bf16_round(f), u = bf16_round(f_up);
    const float sg = bf16_round(g / (1.0f + __expf(-g)));
    reinterpret_cast<__nv_bfloat16*>(p.out)[row * p.ldo + n] = __float2bfloat16_rn(sg * u);
  } else {
    if (p.bias) f += p.bias[n];
    if constexpr (EPI == G2VLM_EPI_STORE_BF16) {
      if (p.flags & G2VLM_GEMM_GELU) f = gelu_erf_d(bf16_round(f));
      else if (p.flags & G2VLM_GEMM_QUICK_GELU) { const float xq = bf16_round(f); f = xq / (1.0f + __expf(-1.702f * xq)); }
      reinterpret_cast<__nv_bfloat16*>(p.out)[row * p.ldo + n] = __float2bfloat16_rn(f);
    } else if constexpr (EPI == G2VLM_EPI_RESID_F32) {
      f = bf16_round(f);
      if (p.use_scale) {
        f *= p.scale[n];
        if (p.flags & G2VLM_GEMM_ROUND_AFTER_SCALE) f = bf16_round(f);
      }
      float* xo = reinterpret_cast<float*>(p.out) + row * p.ldo + n;
      const float xs = *xo + f;
      *xo = (p.flags & G2VLM_GEMM_ROUND_SUM) ? bf16_round(xs) : xs;
    } else {
      if (p.flags & G2VLM_GEMM_ROUND_BF16) f = bf16_round(f);
      float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + n;
      if (p.flags & G2VLM_GEMM_ACCUMULATE) f += *o;
      if (p.flags & G2VLM_GEMM_RELU) f = fmaxf(f, 0.f);
      if (p.residual) f += p.residual[row * p.ldr + n];
      *o = f;
    }
  }
}

// ROWS: 1 (the decode step) or 8; KSPLIT: warps that share one output column (long rows, few columns:
// the down projection K = 8960, N = 1536 would otherwise run on 1536 warps only)
template <int EPI, int ROWS, int KSPLIT>
__global__ void __launch_bounds__(GEMV_WARPS * 32) gemv_bf16_kernel(const GemvParams p) {
  constexpr int COLS = GEMV_WARPS / KSPLIT;  // output columns per block
  __shared__ float part[COLS][KSPLIT][2][ROWS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_out = (EPI == G2VLM_EPI_SWIGLU_BF16) ? p.N / 2 : p.N;
  const int cl = warp / KSPLIT, ks = warp % KSPLIT;
  const int n = blockIdx.x * COLS + cl;  // output column
  const bool active = n < n_out;
  // weight row(s): SwiGLU weights interleave gate/up in blocks of 128 rows
  long long r0 = active ? n : 0, r1 = r0;
  if constexpr (EPI == G2VLM_EPI_SWIGLU_BF16) {
    r0 = (long long)((active ? n : 0) >> 7) * 256 + ((active ? n : 0) & 127);
    r1 = r0 + 128;
  }
  const uint4* w0 = reinterpret_cast<const uint4*>(p.w + r0 * p.ldw);
  const uint4* w1 = reinterpret_cast<const uint4*>(p.w + r1 * p.ldw);
  float acc0[ROWS], acc1[ROWS];
#pragma unroll
  for (int m = 0; m < ROWS; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }
  const int k8 = p.K >> 3;
  const int per = (k8 + KSPLIT - 1) / KSPLIT;
  const int kbeg = ks * per, kend = min(k8, kbeg + per);
  constexpr int U = (ROWS == 1) ? 8 : 4;  // independent 16-byte weight loads in flight per lane
  if (active) {
    for (int c0 = kbeg + lane; c0 < kend; c0 += 32 * U) {
      uint4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + 32 * u;
        a[u] = c < kend ? __ldg(w0 + c) : make_uint4(0, 0, 0, 0);
        if constexpr (EPI == G2VLM_EPI_SWIGLU_BF16) b[u] = c < kend ? __ldg(w1 + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + 32 * u;
        if (c < kend) {
#pragma unroll
          for (int m = 0; m < ROWS; ++m) {
            if (m < p.rows) {
              const uint4 xv = __ldg(reinterpret_cast<const uint4*>(p.x + m * p.ldx) + c);
              dot8(a[u], xv, acc0[m]);
              if constexpr (EPI == G2VLM_EPI_SWIGLU_BF16) dot8(b[u], xv, acc1[m]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < ROWS; ++m) {
    acc0[m] = warp_sum(acc0[m]);
    if constexpr (EPI == G2VLM_EPI_SWIGLU_BF16) acc1[m] = warp_sum(acc1[m]);
  }
  if constexpr (KSPLIT > 1) {
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < ROWS; ++m) { part[cl][ks][0][m] = acc0[m]; part[cl][ks][1][m] = acc1[m]; }
    }
    __syncthreads();
    if (ks != 0) return;
#pragma unroll
    for (int m = 0; m < ROWS; ++m) {
      acc0[m] = 0.f; acc1[m] = 0.f;
#pragma unroll
      for (int q = 0; q < KSPLIT; ++q) { acc0[m] += part[cl][q][0][m]; acc1[m] += part[cl][q][1][m]; }
    }
  }
  if (lane != 0 || !active) return;
  for (int m = 0; m < p.rows && m < ROWS; ++m) gemv_store<EPI>(p, n, m, acc0[m], acc1[m]);
}

// 2..8 rows (the 7-token prompt prefill of every recon call, short chat prompts): still one pass over the weights,
// but 8 scalar dot products per weight would make the kernel issue-bound (measured 0.9 TB/s), so the products go
// through the legacy warp-level tensor-core MMA: D[16 weight rows x 8 x-rows] += W[16 x 16] . X^T[16 x 8]
// (mma.sync m16n8k16 — the right tool for an HBM-bound skinny product; tcgen05 tiles would idle 121 of 128 rows).
// One CTA = 16 output columns, its NW warps split K; lane (g = lane/4, t = lane%4) loads 16 contiguous bytes of
// weight rows g and g+8 and of x row g per 32-element k block (full 32-byte sectors), and because a dot product
// does not care about the order of k, those 8 elements are fed as the k slots {2t,2t+1,2t+8,2t+9} of two MMAs.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int EPI, int NW>
__global__ void __launch_bounds__(NW * 32) gemv_skinny_mma_kernel(const GemvParams p) {
  constexpr bool SW = EPI == G2VLM_EPI_SWIGLU_BF16;
  constexpr int U = 4;  // k blocks in flight per warp
  __shared__ float part[NW][SW ? 2 : 1][16][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 16;  // first output column of this CTA
  long long r0 = n0;
  if constexpr (SW) r0 = (long long)(n0 >> 7) * 256 + (n0 & 127);  // gate rows; up rows sit 128 further
  const uint4* wg0 = reinterpret_cast<const uint4*>(p.w + (r0 + g) * p.ldw) + t;
  const uint4* wg1 = reinterpret_cast<const uint4*>(p.w + (r0 + g + 8) * p.ldw) + t;
  const uint4* wu0 = reinterpret_cast<const uint4*>(p.w + (r0 + 128 + g) * p.ldw) + t;
  const uint4* wu1 = reinterpret_cast<const uint4*>(p.w + (r0 + 128 + g + 8) * p.ldw) + t;
  const bool xrow = g < p.rows;
  const uint4* xp = reinterpret_cast<const uint4*>(p.x + (xrow ? g : 0) * p.ldx) + t;
  const int kb_total = p.K >> 5;
  const int per = (kb_total + NW - 1) / NW;
  const int kb0 = warp * per, kb1 = min(kb_total, kb0 + per);
  float dg[4] = {0.f, 0.f, 0.f, 0.f}, du[4] = {0.f, 0.f, 0.f, 0.f};
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int kb = kb0; kb < kb1; kb += U) {
    uint4 a0[U], a1[U], b0[U], b1[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = kb + u < kb1;
      const int o = (kb + u) * 4;  // uint4 units: 32 elements = 64 B = 4 x 16 B per row
      a0[u] = ok ? __ldg(wg0 + o) : zero;
      a1[u] = ok ? __ldg(wg1 + o) : zero;
      if constexpr (SW) {
        b0[u] = ok ? __ldg(wu0 + o) : zero;
        b1[u] = ok ? __ldg(wu1 + o) : zero;
      }
      xv[u] = (ok && xrow) ? __ldg(xp + o) : zero;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      mma_bf16_16816(dg, a0[u].x, a1[u].x, a0[u].y, a1[u].y, xv[u].x, xv[u].y);
      mma_bf16_16816(dg, a0[u].z, a1[u].z, a0[u].w, a1[u].w, xv[u].z, xv[u].w);
      if constexpr (SW) {
        mma_bf16_16816(du, b0[u].x, b1[u].x, b0[u].y, b1[u].y, xv[u].x, xv[u].y);
        mma_bf16_16816(du, b0[u].z, b1[u].z, b0[u].w, b1[u].w, xv[u].z, xv[u].w);
      }
    }
  }
  // C fragment: d0,d1 = (weight row g, x rows 2t,2t+1), d2,d3 = (weight row g+8, x rows 2t,2t+1)
  part[warp][0][g][2 * t] = dg[0];
  part[warp][0][g][2 * t + 1] = dg[1];
  part[warp][0][g + 8][2 * t] = dg[2];
  part[warp][0][g + 8][2 * t + 1] = dg[3];
  if constexpr (SW) {
    part[warp][1][g][2 * t] = du[0];
    part[warp][1][g][2 * t + 1] = du[1];
    part[warp][1][g + 8][2 * t] = du[2];
    part[warp][1][g + 8][2 * t + 1] = du[3];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int m = threadIdx.x >> 4, c = threadIdx.x & 15;  // 16 consecutive columns per row: coalesced stores
    if (m < p.rows) {
      float f = 0.f, fu = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        f += part[w][0][c][m];
        if constexpr (SW) fu += part[w][1][c][m];
      }
      gemv_store<EPI>(p, n0 + c, m, f, fu);
    }
  }
}


// 9..32 rows (system prompt, short questions of the chat path): the same mapping with RB blocks of 8 x-rows per weight
// fragment — the weights still stream ONCE, every k block feeds 2 x RB MMAs (x RB more for SwiGLU).  The legacy MMA of
// sm_100 retires one m16n8k16 per 2 cycles and SM (tools/micro/mma_sync_bench.cu), so the MMAs stay within
// reach of the HBM rate, while a 128-row tcgen05 tile would put the whole [N, K] weight on N / 256 CTAs (6 CTAs for the
// o / down projections: measured 5 ms for a 20-token prefill, 0.5 TB/s).  8 warps split K; partial tiles meet in shared memory.
constexpr int ROWS_MMA_MAX = 32;   // (RB = 8, 33..64 rows, measured SLOWER than the tile kernel: 7.5 vs 5.7 ms for a 60-token prefill —
                                   //  every CTA re-reads all x rows from L2)
template <int EPI, int RB>
__global__ void __launch_bounds__(256) gemv_rows_mma_kernel(const GemvParams p) {
  constexpr bool SW = EPI == G2VLM_EPI_SWIGLU_BF16;
  constexpr int NW = 8, NH = SW ? 2 : 1, XR = 8 * RB;
  constexpr int U = RB <= 2 ? 4 : 2;  // k blocks in flight per warp
  extern __shared__ float part_dyn[];  // [NW][NH][16][XR]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 16;
  long long r0 = n0;
  if constexpr (SW) r0 = (long long)(n0 >> 7) * 256 + (n0 & 127);
  const uint4* wg0 = reinterpret_cast<const uint4*>(p.w + (r0 + g) * p.ldw) + t;
  const uint4* wg1 = reinterpret_cast<const uint4*>(p.w + (r0 + g + 8) * p.ldw) + t;
  const uint4* wu0 = reinterpret_cast<const uint4*>(p.w + (r0 + 128 + g) * p.ldw) + t;
  const uint4* wu1 = reinterpret_cast<const uint4*>(p.w + (r0 + 128 + g + 8) * p.ldw) + t;
  const uint4* xp[RB];
  bool xok[RB];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    xok[rb] = rb * 8 + g < p.rows;
    xp[rb] = reinterpret_cast<const uint4*>(p.x + (xok[rb] ? rb * 8 + g : 0) * p.ldx) + t;
  }
  const int kb_total = p.K >> 5;
  const int per = (kb_total + NW - 1) / NW;
  const int kb0 = warp * per, kb1 = min(kb_total, kb0 + per);
  float dg[RB][4], du[SW ? RB : 1][4];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dg[rb][e] = 0.f; if constexpr (SW) du[rb][e] = 0.f; }
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int kb = kb0; kb < kb1; kb += U) {
    uint4 a0[U], a1[U], b0[U], b1[U], xv[RB][U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = kb + u < kb1;
      const int o = (kb + u) * 4;
      a0[u] = ok ? __ldg(wg0 + o) : zero;
      a1[u] = ok ? __ldg(wg1 + o) : zero;
      if constexpr (SW) {
        b0[u] = ok ? __ldg(wu0 + o) : zero;
        b1[u] = ok ? __ldg(wu1 + o) : zero;
      }
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) xv[rb][u] = (ok && xok[rb]) ? __ldg(xp[rb] + o) : zero;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        mma_bf16_16816(dg[rb], a0[u].x, a1[u].x, a0[u].y, a1[u].y, xv[rb][u].x, xv[rb][u].y);
        mma_bf16_16816(dg[rb], a0[u].z, a1[u].z, a0[u].w, a1[u].w, xv[rb][u].z, xv[rb][u].w);
        if constexpr (SW) {
          mma_bf16_16816(du[rb], b0[u].x, b1[u].x, b0[u].y, b1[u].y, xv[rb][u].x, xv[rb][u].y);
          mma_bf16_16816(du[rb], b0[u].z, b1[u].z, b0[u].w, b1[u].w, xv[rb][u].z, xv[rb][u].w);
        }
      }
    }
  }
  // C fragment: d0,d1 = (weight row g, x rows 8 rb + 2t, +1), d2,d3 = (weight row g + 8, same x rows)
  float* pw = part_dyn + (long long)warp * NH * 16 * XR;
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    pw[g * XR + rb * 8 + 2 * t] = dg[rb][0];
    pw[g * XR + rb * 8 + 2 * t + 1] = dg[rb][1];
    pw[(g + 8) * XR + rb * 8 + 2 * t] = dg[rb][2];
    pw[(g + 8) * XR + rb * 8 + 2 * t + 1] = dg[rb][3];
    if constexpr (SW) {
      pw[(16 + g) * XR + rb * 8 + 2 * t] = du[rb][0];
      pw[(16 + g) * XR + rb * 8 + 2 * t + 1] = du[rb][1];
      pw[(16 + g + 8) * XR + rb * 8 + 2 * t] = du[rb][2];
      pw[(16 + g + 8) * XR + rb * 8 + 2 * t + 1] = du[rb][3];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * XR; i += 256) {
    const int m = i >> 4, c = i & 15;  // 16 consecutive columns per row: coalesced stores
    if (m < p.rows) {
      float f = 0.f, fu = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        f += part_dyn[((long long)w * NH * 16 + c) * XR + m];
        if constexpr (SW) fu += part_dyn[((long long)w * NH * 16 + 16 + c) * XR + m];
      }
      gemv_store<EPI>(p, n0 + c, m, f, fu);
    }
  }
}

template <int EPI, int RB>
static int launch_rows_mma(const GemvParams& p, int n_out, cudaStream_t stream) {
  constexpr int NH = EPI == G2VLM_EPI_SWIGLU_BF16 ? 2 : 1;
  const int smem = 8 * NH * 16 * 8 * RB * 4;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gemv_rows_mma_kernel<EPI, RB>), smem)) return rc;
  gemv_rows_mma_kernel<EPI, RB><<<n_out / 16, 256, smem, stream>>>(p);
  return G2VLM_OK;
}

template <int EPI>
static int launch_gemv_epi(const GemvParams& p, int n_out, cudaStream_t stream) {
  const bool one = p.rows == 1;
  const bool split = p.K >= 4096;
  if (p.rows > GEMV_MAX_ROWS) {   // 9..32 rows (the caller checked the shape)
    if (p.rows <= 16) return launch_rows_mma<EPI, 2>(p, n_out, stream);
    return launch_rows_mma<EPI, 4>(p, n_out, stream);
  }
  if (!one && n_out % 16 == 0 && p.K % 32 == 0 && p.ldw % 8 == 0 && p.ldx % 8 == 0) {
    if (split) gemv_skinny_mma_kernel<EPI, 16><<<n_out / 16, 16 * 32, 0, stream>>>(p);
    else gemv_skinny_mma_kernel<EPI, 8><<<n_out / 16, 8 * 32, 0, stream>>>(p);
    return G2VLM_OK;
  }
  const int cols = split ? GEMV_WARPS / 4 : GEMV_WARPS;
  const unsigned grid = (n_out + cols - 1) / cols;
  if (one && split) gemv_bf16_kernel<EPI, 1, 4><<<grid, GEMV_WARPS * 32, 0, stream>>>(p);
  else if (one) gemv_bf16_kernel<EPI, 1, 1><<<grid, GEMV_WARPS * 32, 0, stream>>>(p);
  else if (split) gemv_bf16_kernel<EPI, GEMV_MAX_ROWS, 4><<<grid, GEMV_WARPS * 32, 0, stream>>>(p);
  else gemv_bf16_kernel<EPI, GEMV_MAX_ROWS, 1><<<grid, GEMV_WARPS * 32, 0, stream>>>(p);
  return G2VLM_OK;
}

// Called by g2vlm_gemm_bf16 when the call has <= GEMV_MAX_ROWS rows, all in ONE group.
int launch_gemv(const g2vlm_gemm_args* a, int group, cudaStream_t stream) {
  GemvParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->A) + (long long)a->group_row0[group] * a->lda;
  p.ldx = a->lda;
  p.w = reinterpret_cast<const __nv_bfloat16*>(a->B) + (long long)group * a->N * a->ldb;
  p.ldw = a->ldb;
  p.rows = a->group_rows[group];
  p.N = a->N;
  p.K = a->K;
  p.flags = a->flags;
  p.use_scale = a->scale != nullptr && ((a->scale_groups >> group) & 1u);
  p.out = a->out;
  p.ldo = a->ldo;
  p.bias = a->bias ? a->bias + (long long)group * a->N : nullptr;
  p.scale = a->scale;
  p.residual = a->residual;
  p.ldr = a->ldr;
  p.row0 = a->group_row0[group];
  const int n_out = a->epilogue == G2VLM_EPI_SWIGLU_BF16 ? a->N / 2 : a->N;
  int rc;
  switch (a->epilogue) {
    case G2VLM_EPI_STORE_BF16: rc = launch_gemv_epi<G2VLM_EPI_STORE_BF16>(p, n_out, stream); break;
    case G2VLM_EPI_SWIGLU_BF16: rc = launch_gemv_epi<G2VLM_EPI_SWIGLU_BF16>(p, n_out, stream); break;
    case G2VLM_EPI_RESID_F32: rc = launch_gemv_epi<G2VLM_EPI_RESID_F32>(p, n_out, stream); break;
    default: rc = launch_gemv_epi<G2VLM_EPI_STORE_F32>(p, n_out, stream); break;
  }
  if (rc) return rc;
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

// ------------------------------------------------------------------------------------------------
// decode attention (one query token, head_dim 128)
// ------------------------------------------------------------------------------------------------
constexpr int DEC_THREADS = 256;
constexpr int DEC_CHUNK = 160;   // keys per CTA: K and V chunks (2 x 40 KB) are staged in shared memory
constexpr int DEC_MAX_G = 8;     // query heads per KV head
constexpr int DEC_SMEM = 2 * DEC_CHUNK * 256 + DEC_MAX_G * DEC_CHUNK * 4 + 4 * DEC_MAX_G * 128 * 4 + 64;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// One CTA = one KV head x one chunk of <= 160 keys. The whole K and V chunk is pulled into shared memory by
// ONE round of cp.async issued up front (a decode step is latency-bound: the point is to have every byte in
// flight at once), then scores / softmax statistics / P.V run out of shared memory.
__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_split_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, long long ldk,
                         const __nv_bfloat16* __restrict__ v, long long ldv, int L_static,
                         const int* __restrict__ kv_len_dev, int kv_len_extra, int G,
                         float scale_log2, float* __restrict__ part /*[splits][heads][130]*/, int n_heads) {
  extern __shared__ __align__(16) uint8_t dsm[];
  uint8_t* sK = dsm;                                   // [chunk][256 B]
  uint8_t* sV = dsm + DEC_CHUNK * 256;                 // [chunk][256 B]
  float (*ss)[DEC_CHUNK] = reinterpret_cast<float (*)[DEC_CHUNK]>(dsm + 2 * DEC_CHUNK * 256);
  float (*red)[DEC_MAX_G][128] = reinterpret_cast<float (*)[DEC_MAX_G][128]>(dsm + 2 * DEC_CHUNK * 256 + DEC_MAX_G * DEC_CHUNK * 4);
  float* sm = reinterpret_cast<float*>(dsm + 2 * DEC_CHUNK * 256 + DEC_MAX_G * DEC_CHUNK * 4 + 4 * DEC_MAX_G * 128 * 4);
  float* sl = sm + DEC_MAX_G;
  const int split = blockIdx.x, kvh = blockIdx.y, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // the key count may live on the device (CUDA-graph replay of the decode step: same launch, growing cache)
  const int L = kv_len_dev ? (*kv_len_dev + kv_len_extra) : L_static;
  const int chunk = min(DEC_CHUNK, (L + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x));
  const int k0 = split * chunk, k1 = min(L, k0 + chunk), nk = max(0, k1 - k0);

  for (int i = tid; i < nk * 16; i += DEC_THREADS) {   // 16 x 16 B per 256-byte row
    const int r = i >> 4, c = i & 15;
    cp_async16(sK + r * 256 + c * 16, k + (long long)(k0 + r) * ldk + kvh * 128 + c * 8);
    cp_async16(sV + r * 256 + c * 16, v + (long long)(k0 + r) * ldv + kvh * 128 + c * 8);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // the query (pre-scaled by scale*log2e): lane = (key slot kk = lane/8, 16-byte chunks cp and cp+8)
  const int kk = lane >> 3, cp = lane & 7;
  float qr[DEC_MAX_G][16];
#pragma unroll
  for (int h = 0; h < DEC_MAX_G; ++h) {
    if (h < G) {
      const uint4* qp = reinterpret_cast<const uint4*>(q + (kvh * G + h) * 128);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint4 qv = __ldg(qp + cp + 8 * half);
        const __nv_bfloat162* pq = reinterpret_cast<const __nv_bfloat162*>(&qv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(pq[i]);
          qr[h][half * 8 + 2 * i] = f.x * scale_log2;
          qr[h][half * 8 + 2 * i + 1] = f.y * scale_log2;
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // phase 1: scores; a warp handles 4 keys per step, 8 lanes per key, shuffle-reduced
  for (int j0 = warp * 4; j0 < nk; j0 += (DEC_THREADS / 32) * 4) {
    const int j = j0 + kk;
    const uint8_t* kr = sK + min(j, nk - 1) * 256;
    float kf[16];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint4 kv = *reinterpret_cast<const uint4*>(kr + (cp + 8 * half) * 16);
      const __nv_bfloat162* pk = reinterpret_cast<const __nv_bfloat162*>(&kv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(pk[i]);
        kf[half * 8 + 2 * i] = f.x; kf[half * 8 + 2 * i + 1] = f.y;
      }
    }
#pragma unroll
    for (int h = 0; h < DEC_MAX_G; ++h) {
      if (h < G) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) a = fmaf(kf[i], qr[h][i], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (cp == 0 && j < nk) ss[h][j] = a;
      }
    }
  }
  __syncthreads();
  // phase 2: per-head max / exp2 / sum over the chunk (one warp per head)
  if (warp < G) {
    float m = -INFINITY;
    for (int j = lane; j < nk; j += 32) m = fmaxf(m, ss[warp][j]);
    m = warp_max(m);
    float l = 0.f;
    for (int j = lane; j < nk; j += 32) {
      const float pj = exp2f(ss[warp][j] - m);
      ss[warp][j] = pj;
      l += pj;
    }
    l = warp_sum(l);
    if (lane == 0) { sm[warp] = m; sl[warp] = l; }
  }
  __syncthreads();
  // phase 3: o[h][d] = sum_j p[h][j] * v[j][d]; thread = (pair of d, key quarter)
  const int dp = tid & 63, kq = tid >> 6;
  float o[DEC_MAX_G][2];
#pragma unroll
  for (int h = 0; h < DEC_MAX_G; ++h) { o[h][0] = 0.f; o[h][1] = 0.f; }
  for (int j = kq; j < nk; j += 4) {
    const float2 vf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sV + j * 256 + dp * 4));
#pragma unroll
    for (int h = 0; h < DEC_MAX_G; ++h)
      if (h < G) {
        const float pj = ss[h][j];
        o[h][0] = fmaf(pj, vf.x, o[h][0]);
        o[h][1] = fmaf(pj, vf.y, o[h][1]);
      }
  }
#pragma unroll
  for (int h = 0; h < DEC_MAX_G; ++h)
    if (h < G) { red[kq][h][dp * 2] = o[h][0]; red[kq][h][dp * 2 + 1] = o[h][1]; }
  __syncthreads();
  for (int i = tid; i < G * 128; i += DEC_THREADS) {
    const int h = i >> 7, d = i & 127;
    float* dst = part + ((long long)split * n_heads + kvh * G + h) * 130;
    dst[d] = red[0][h][d] + red[1][h][d] + red[2][h][d] + red[3][h][d];
    if (d == 0) { dst[128] = nk > 0 ? sm[h] : -INFINITY; dst[129] = nk > 0 ? sl[h] : 0.f; }
  }
}

// log-sum-exp merge of the splits: one block per head, 4 groups of 128 threads stride over the splits
__global__ void __launch_bounds__(512)
attn_decode_merge_kernel(const float* __restrict__ part, int n_splits, int n_heads, __nv_bfloat16* __restrict__ out) {
  const int h = blockIdx.x, d = threadIdx.x & 127, grp = threadIdx.x >> 7;
  __shared__ float wmax[16];
  __shared__ float racc[4][128], rl[4];
  float M = -INFINITY;
  for (int s = threadIdx.x; s < n_splits; s += 512) M = fmaxf(M, part[((long long)s * n_heads + h) * 130 + 128]);
  M = warp_max(M);
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = M;
  __syncthreads();
  M = wmax[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) M = fmaxf(M, wmax[i]);
  float acc = 0.f, l = 0.f;
#pragma unroll 4
  for (int s = grp; s < n_splits; s += 4) {
    const float* p = part + ((long long)s * n_heads + h) * 130;
    const float ps = p[129];
    const float w = ps > 0.f ? exp2f(p[128] - M) : 0.f;
    acc = fmaf(w, p[d], acc);
    l = fmaf(w, ps, l);
  }
  racc[grp][d] = acc;
  if (d == 0) rl[grp] = l;
  __syncthreads();
  if (grp == 0) {
    const float a = racc[0][d] + racc[1][d] + racc[2][d] + racc[3][d];
    const float lt = rl[0] + rl[1] + rl[2] + rl[3];
    out[h * 128 + d] = __float2bfloat16_rn(lt > 0.f ? a / lt : 0.f);
  }
}

// dst[(*len_dev or static_row) + i, :] = src[i, :] — the in-place append of a step's K|V rows
__global__ void kv_append_kernel(const uint8_t* __restrict__ src, long long src_pitch, uint8_t* __restrict__ dst,
                                 long long dst_pitch, const int* __restrict__ len_dev, long long static_row,
                                 long long rows, int chunks) {
  const long long base = len_dev ? *len_dev : static_row;
  const long long total = rows * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks);
    *reinterpret_cast<uint4*>(dst + (base + r) * dst_pitch + c * 16LL) =
        *reinterpret_cast<const uint4*>(src + r * src_pitch + c * 16LL);
  }
}

}  // namespace g2

extern "C" int g2vlm_kv_append(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes,
                               const int32_t* len_dev, int64_t static_row, int64_t rows, int64_t row_bytes,
                               void* stream) {
  using namespace g2;
  G2_REQUIRE(src && dst && row_bytes > 0 && row_bytes % 16 == 0 && src_pitch_bytes % 16 == 0 && dst_pitch_bytes % 16 == 0,
             "kv_append: bad arguments");
  if (rows <= 0) return G2VLM_OK;
  const int chunks = static_cast<int>(row_bytes / 16);
  long long blocks = (rows * chunks + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  kv_append_kernel<<<static_cast<unsigned>(blocks), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)src, src_pitch_bytes, (uint8_t*)dst, dst_pitch_bytes, len_dev, static_row, rows, chunks);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

extern "C" int g2vlm_attention_decode(const void* q, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                      int64_t kv_len, const int32_t* kv_len_dev, int32_t kv_len_extra, void* out,
                                      int32_t num_q_heads, int32_t num_kv_heads, int32_t head_dim,
                                      float softmax_scale, float* workspace, int64_t workspace_floats,
                                      void* stream) {
  using namespace g2;
  G2_REQUIRE(q && k && v && out && workspace, "attention_decode: null tensor");
  G2_REQUIRE(head_dim == 128, "attention_decode: head_dim must be 128");
  G2_REQUIRE(num_kv_heads > 0 && num_q_heads % num_kv_heads == 0 && num_q_heads / num_kv_heads <= DEC_MAX_G,
             "attention_decode: at most 8 query heads per KV head");
  G2_REQUIRE(kv_len > 0 && kv_len < (1LL << 31), "attention_decode: bad kv_len");
  G2_REQUIRE(ldk % 8 == 0 && ldv % 8 == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0, "attention_decode: k / v must be 16-byte aligned rows");
  // with kv_len_dev the actual key count is *kv_len_dev + kv_len_extra (read on the device, <= kv_len);
  // kv_len then only sizes the split grid, so one captured launch serves a growing cache
  const int L = static_cast<int>(kv_len);
  const int n_splits = (L + DEC_CHUNK - 1) / DEC_CHUNK;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_decode_split_kernel), DEC_SMEM)) return rc;
  G2_REQUIRE((long long)n_splits * num_q_heads * 130 <= workspace_floats, "attention_decode: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  attn_decode_split_kernel<<<dim3(n_splits, num_kv_heads), DEC_THREADS, DEC_SMEM, st>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, L, kv_len_dev,
      kv_len_extra, num_q_heads / num_kv_heads, softmax_scale * 1.4426950408889634f, workspace, num_q_heads);
  G2_CUDA_OK(cudaGetLastError());
  attn_decode_merge_kernel<<<num_q_heads, 512, 0, st>>>(workspace, n_splits, num_q_heads, (__nv_bfloat16*)out);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}
