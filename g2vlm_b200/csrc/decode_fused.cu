// g2vlm_b200 — ONE persistent kernel for a whole greedy decode step (reference
// generate_text, modeling/g2vlm/g2vlm.py:1086-1131, und expert of modeling/g2vlm/qwen2vl.py:555-664, 842-910).
//
// A decode step streams 3.7 GB through the GPU (2.6 GB of und-expert weights, 0.47 GB of lm_head, the K/V cache)
// for ~6 GFLOP: it is HBM-bound, and as ~280 dependent launches of a few microseconds each it ran at 0.24 of the HBM
// roofline — every launch drains the memory pipeline, waits for the grid to retire and ramps up again.  Here the
// step is one cooperative launch of one CTA per SM, 16 warps:
//   * warp 15 of every CTA is a PRODUCER: it walks the CTA's share of every weight matrix of the step, in order —
//     qkv, o_proj, gate/up, down of every layer, then lm_head — and streams it with TMA bulk copies
//     (cp.async.bulk.shared.global; one whole 3 KB weight row, or a 2.5 KB piece of a down-projection row, per copy)
//     into a 4-stage shared-memory ring of 8-row tiles guarded by full / empty mbarriers.  It depends on nothing but free
//     ring slots, so it runs ahead ACROSS the grid barriers and the latency-bound phases.
//   * warps 0-14 are CONSUMERS.  The phases of a layer
//         qkv GEMV | attention partials over a key range | merge | o_proj GEMV | gate/up GEMV + SwiGLU | down GEMV
//     are separated by grid barriers (release-add + acquire-poll on a device counter).  In a GEMV phase warp w < 8
//     multiplies tile row w of every ring stage by the phase's input vector with packed fp32 FMAs (fma.rn.f32x2) and
//     leaves its lanes' partial sums in a table that one thread per output row reduces at the end of the phase.
//   * attention: each CTA owns a key range of one KV head; its K|V rows are pulled with cp.async (issued in front of
//     the barrier wait: the cache rows do not depend on this step); scores and P.V are 16-key tiles on mma.sync.
// Every vector a phase needs (normalised hidden state, attention output, SwiGLU activations) is rebuilt per CTA in
// shared memory from the L2-resident fp32 / bf16 vectors; everything written during the kernel is read back with
// ld.global.cg (L2 only — the other CTAs' updates).
// Inside a GEMV phase the weight stream runs at the HBM rate (6.3 TB/s); the step as a whole reaches 0.41 of it because the
// six grid-wide dependencies of a layer are latency.  Measurements, the earlier versions of this kernel and what was
// tried and rejected: profiles/r02_decode_fused.txt.
//
// Rounding points are those of the multi-launch driver (csrc/decode_step.cu) kernel by kernel: fp32 RMSNorm -> bf16,
// bf16 x bf16 products accumulated in fp32, q/k-norm on the bf16 tensor (normalised value rounded to bf16 before the
// weight multiply), fp32 softmax in base 2, bf16 P (as the prefill kernel and flash-attn), bf16 attention output,
// x += bf16(acc) residual updates, bf16 logits, argmax with torch semantics.
#include "common.cuh"

namespace g2 {

constexpr int DF_CWARPS = 15;                    // consumer warps
constexpr int DF_CTHREADS = DF_CWARPS * 32;
constexpr int DF_THREADS = DF_CTHREADS + 32;     // + the producer warp = 16 warps: 4 per scheduler, 128 registers per
                                                 //   thread (a 17th warp capped the kernel at 96 and it spilled; with 219 KB
                                                 //   of shared memory there is next to no L1 left to catch the spills)
constexpr int DF_MSL = DF_CTHREADS / 16;         // merge phase: split lanes per d
constexpr int DF_MAX_LAYERS = 32;
constexpr int DF_MAX_G = 8;                      // query heads per KV head
constexpr int DF_PART = 132;                     // floats per attention partial: o[128], m, l, pad
constexpr int DF_MAX_VEC = 9216;                 // largest shared-memory input vector (elements): max(H, I, nq*128)
constexpr int DF_NS = 4;                         // ring stages
constexpr int DF_KS_MAX = 1536;                  // columns per stage: whole rows of the 1536-wide matrices (3 KB copies)
constexpr int DF_TILE = 8;                       // weight rows per stage
constexpr int DF_STAGE_BYTES = DF_TILE * (2 * DF_KS_MAX + 64);   // pitch = 2 KS + 64 (conflict-free LDS.128)
constexpr int DF_STAGE = 160;                    // keys per attention stage (K and V rows: 2 x 40 KB)
constexpr int DF_NKV = 1;                        // K|V stage buffers.  (2 x 80 keys, the next stage loading under this one's
                                                 //   softmax: measured 3 % slower — a stage costs ~1500 cycles of barriers and
                                                 //   warp reductions whatever its size, and the load is mostly hidden anyway)
constexpr int DF_SLD = DF_STAGE + 4;             // row pitch of the score matrix (bank-conflict-free C stores)
constexpr int DF_ROUND = 32;                     // stages per round of a phase: the partial table (32 x 8 rows x 33 floats =
                                                 //   33 KB) lives in the K buffer

struct DecFusedParams {
  int num_layers, H, I, nq, nkv, vocab, vpad;
  float eps, scale_log2;
  int s0, s1;
  int ks_h, ks_a, ks_i;            // columns per ring stage for K = H, nq*128, I
  int keep_token;                  // the caller samples the next token: *cur_token is left alone
  int opt;                         // G2VLM_DECODE_OPT, default 1: bit 0 = no explicit __threadfence around the grid barrier's
                                   //   release-add / acquire-poll (0: with fences, +0.1 ms per token)
  g2vlm_und_layer_weights layers[DF_MAX_LAYERS];
  __nv_bfloat16* kv[DF_MAX_LAYERS];
  long long kv_capacity;
  const float* embed;
  const float* final_norm;
  const __nv_bfloat16* lm_head;
  const float* inv_freq;
  long long* cur_token;
  long long* position;
  int* cache_len;
  float* x;
  __nv_bfloat16* qkv;
  __nv_bfloat16* attn;
  __nv_bfloat16* act;
  __nv_bfloat16* logits;
  // fused workspace: [0] barrier arrivals (monotonic), [1] arrivals at the start of the next launch,
  // then the argmax candidates and the attention partials
  unsigned* sync;
  float* cand;                     // [grid][2]  (value, index bits)
  float* part;                     // [grid / nkv splits][nq][DF_PART]
};

// ---- memory helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_df(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void csync() {   // the 16 consumer warps
  asm volatile("bar.sync 1, %0;" ::"n"(DF_CTHREADS) : "memory");
}

// ---- grid barrier (consumer warps), split in two so that independent work sits between ----------------------------
// Monotonic arrival counter; `target` advances by gridDim.x per barrier.  arrive: bar.sync + fence + release add by
// thread 0 publish the CTA's writes; wait: thread 0 polls with acquire loads, the other threads wait at bar.sync.
__device__ __forceinline__ void barrier_arrive(unsigned* ctr, unsigned& target, bool fences = true) {
  csync();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    if (fences) __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
  }
}
__device__ __forceinline__ void barrier_wait(unsigned* ctr, unsigned target, bool fences = true) {
  if (threadIdx.x == 0) {
    unsigned v;
    long long t0 = 0;
    uint32_t spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (static_cast<int>(v - target) >= 0) break;
      if ((++spins & 0xfff) == 0) {
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) {
          printf("g2vlm_b200: decode grid barrier watchdog fired (block %d, %u of %u)\n", blockIdx.x, v, target);
          __trap();
        }
      }
    }
    if (fences) __threadfence();
  }
  csync();
}

// ---- building blocks ------------------------------------------------------------------------------------------
// The input vector of a GEMV phase sits in shared memory as bf16 (what the Linear sees), in the V buffer of the attention
// phase.  (fp32 there saves the conversions but every consumer warp re-read it for every stage: 12 LDS.128 per warp and
// stage next to 6 for the weights — the shared-memory pipe, 128 B / cycle, then capped a stage at ~850 cycles.)
// a bf16 vector of n8 16-byte chunks in global memory (written by other CTAs in the previous phase) -> shared memory
__device__ __forceinline__ void vec_from_bf16(uint4* xs, const __nv_bfloat16* src, int n8) {
  for (int c = threadIdx.x; c < n8; c += DF_CTHREADS) xs[c] = __ldcg(reinterpret_cast<const uint4*>(src) + c);
  csync();
}
// RMSNorm of the fp32 vector src (L2) into shared memory: bf16(w * (x * r)); one round trip to L2 (every thread keeps
// its float4 in registers), block-wide sum of squares through shared memory.
__device__ __forceinline__ void rmsnorm_to_smem(const float* src, const float* w, int dim, float eps, uint4* xs,
                                                float* s_part) {
  const int n4 = dim >> 2;
  const float4* xr = reinterpret_cast<const float4*>(src);
  const int i0 = threadIdx.x;
  float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), g0 = v0;
  if (i0 < n4) {
    v0 = __ldcg(xr + i0);
    g0 = __ldg(reinterpret_cast<const float4*>(w) + i0);   // (cold in HBM: in flight together with x, not after the reduction)
  }
  float ss = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w;
  for (int i = i0 + DF_CTHREADS; i < n4; i += DF_CTHREADS) {
    const float4 v = __ldcg(xr + i);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  csync();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < DF_CWARPS; ++i) tot += s_part[i];
  const float r = rsqrtf(tot / dim + eps);
  for (int i = i0; i < n4; i += DF_CTHREADS) {
    const float4 v = i == i0 ? v0 : __ldcg(xr + i);
    const float4 g = i == i0 ? g0 : __ldg(reinterpret_cast<const float4*>(w) + i);
    reinterpret_cast<uint2*>(xs)[i] = make_uint2(pack_bf16x2(g.x * (v.x * r), g.y * (v.y * r)),
                                                 pack_bf16x2(g.z * (v.z * r), g.w * (v.w * r)));
  }
  csync();
}

__device__ __forceinline__ bool argmax_better_df(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (vn || v == bv) return i < bi;
  return v > bv;
}

// warp-level tensor-core pieces (16-row tiles of an HBM-bound kernel: the legacy mma.sync shape is the right size —
// a tcgen05 tile would idle 120 of its 128 rows)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- the weight stream: what the producer copies and the consumers multiply, in the same order ---------------------
// A GEMV phase of one CTA: `count` rows of a [N, K] weight matrix — rows [first, first + count), or for the gate/up
// matrix the gate rows of units [first, first + nu) followed by their up rows — cut into tiles of 8 rows x K / KS slabs
// of KS columns; stage index = tile * n_slabs + slab.  Rows (units) are dealt to the CTAs as evenly as N allows.
struct DfPhase {
  const __nv_bfloat16* w;
  int K, KS, n_slabs;
  int first, nu, count, n_tiles;
  bool swiglu;
};
__device__ __forceinline__ DfPhase df_phase(const void* w, int N, int K, int KS, bool swiglu) {
  DfPhase ph;
  ph.w = reinterpret_cast<const __nv_bfloat16*>(w);
  ph.K = K; ph.KS = KS; ph.n_slabs = K / KS; ph.swiglu = swiglu;
  const int base = N / gridDim.x, rem = N % gridDim.x, c = blockIdx.x;
  ph.first = c * base + min(c, rem);
  ph.nu = base + (c < rem ? 1 : 0);
  ph.count = swiglu ? 2 * ph.nu : ph.nu;
  ph.n_tiles = (ph.count + DF_TILE - 1) / DF_TILE;
  return ph;
}
// global weight row behind local row r of the phase, or -1
__device__ __forceinline__ long long df_row(const DfPhase& ph, int r) {
  if (r >= ph.count) return -1;
  if (!ph.swiglu) return ph.first + r;
  const int up = r >= ph.nu ? 1 : 0;
  const int j = ph.first + r - up * ph.nu;                                // gate/up interleaved in blocks of 128 rows
  return (long long)(j >> 7) * 256 + (j & 127) + up * 128;
}

struct DfSmem {
  uint8_t ring[DF_NS][DF_STAGE_BYTES];           // the weight stream
  uint8_t kbuf[DF_NKV * DF_STAGE * 256];         // K rows of one KV head, 16-byte chunks XOR-swizzled by row & 7;
  uint8_t vbuf[DF_NKV * DF_STAGE * 256];              //   outside the attention phase kbuf holds the partial table and vbuf
                                                 //   the fp32 input vector of the GEMV phase
  float score[8][DF_SLD];                        // scores, then probabilities (heads >= G stay zero)
  float red[2][8][128];                          // P.V halves; scratch of the merge
  __nv_bfloat16 q[8][128];                       // normalised, rotated query heads (rows >= G zero)
  __nv_bfloat16 new_k[128], new_v[128];          // this step's K / V row of the CTA's KV head
  float m_run[8], l_run[8], alpha[8];
  float part[DF_CWARPS];
  float rope_cs[64], rope_sn[64];                // M-RoPE angles of this step's position (the same for every layer)
  float best_v[DF_CWARPS];
  int best_i[DF_CWARPS];
  uint64_t full[DF_NS], empty[DF_NS];
  long long t_acc[14], t_gu[5], t_at[6];         // TIMING build only
};

// the producer's wait for a free slot backs off between polls: a tight mbarrier.try_wait loop keeps the shared-memory /
// SYNCS path of the SM busy and slows the consumers' LDS / STS / SHFL (measured: the softmax of the attention phase
// took 4100 cycles per stage next to a spinning producer)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      printf("g2vlm_b200: decode producer watchdog fired (block %d parity %u)\n", blockIdx.x, parity);
      __trap();
    }
  }
}
// producer: one phase of the stream.  Lane i < 8 copies the KS-column piece of tile row i (>= 2.5 KB per copy: the TMA
// unit of an SM retires a bulk copy every ~45 ns whatever its size — tools/micro/bulk_copy_bench.cu: 16 x 1.5 KB pieces
// per stage stream at 5.4 TB/s over the chip, 8 x 3 KB at 7.4 TB/s).
__device__ __forceinline__ void df_produce(DfSmem& s, const DfPhase& ph, uint32_t& q, int lane) {
  const uint32_t piece = 2u * ph.KS, pitch = piece + 64u;
  for (int tile = 0; tile < ph.n_tiles; ++tile) {
    const long long row = lane < DF_TILE ? df_row(ph, tile * DF_TILE + lane) : -1;
    const unsigned valid = __ballot_sync(0xffffffffu, row >= 0);
    const uint32_t total = piece * __popc(valid);
    const __nv_bfloat16* src = ph.w + (row >= 0 ? row : 0) * ph.K;
    for (int slab = 0; slab < ph.n_slabs; ++slab, ++q) {
      const int slot = q % DF_NS;
      if (q >= DF_NS) mbar_wait_backoff(&s.empty[slot], ((q / DF_NS) - 1) & 1);
      if (lane == 0) mbar_arrive_expect_tx(&s.full[slot], total);
      __syncwarp();
      if (row >= 0) bulk_g2s(s.ring[slot] + lane * pitch, src + (long long)slab * ph.KS, piece, &s.full[slot]);
    }
  }
}

// consumers: stages [i0, i1) of a phase times the vector in shared memory.  Warp w < 8 takes tile row w of EVERY stage, in
// order (an mbarrier parity wait only tells the current use of a slot from the previous one, so no warp may wait two uses
// ahead), lanes on consecutive 16-byte chunks (conflict-free), products as packed fp32 FMAs (fma.rn.f32x2; bf16 -> fp32
// is a shift / a mask), and releases the slot as soon as its row is in registers: a stage is held ~300 cycles, so the
// slots spend their time being filled.  The row sum of stage i goes to ptab[(i - i0) * 8 + w].  Warps 8-14 idle in these phases.
// Measured on the way here (profiles/r02_decode_fused.txt): mma.sync m16n8k16 with the vector replicated over the 8
// columns was no faster; `c < n ? smem[c] : 0` compiles to a branch per LDS.128 (4400 cycles per stage), hence the
// clamped indices; a warp pair per slot holds a stage for 1700 cycles.
__device__ __forceinline__ unsigned long long df_f2(uint32_t w) {   // bf16x2 -> f32x2
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(w << 16), "r"(w & 0xffff0000u));
  return r;
}
__device__ __forceinline__ unsigned long long df_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long df_ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
constexpr int DF_CH = DF_KS_MAX / 8 / 32;        // 16-byte chunks per lane and row piece (6)
constexpr int DF_PP = 33;                        // pitch of a row's 32 lane partials in the partial table (conflict-free reads)
// one row of stage q + i -> registers: wait for the stage, 6 x LDS.128 (clamped chunk indices: a predicated LDS.128 compiles
// to a branch per load; a clamped chunk is zeroed), release the slot
__device__ __forceinline__ void df_fetch(DfSmem& s, uint32_t qi, int row_off, const int (&cc)[DF_CH], int nch, int lane,
                                         uint4 (&dst)[DF_CH]) {
  const int slot = qi % DF_NS;
  mbar_wait(&s.full[slot], (qi / DF_NS) & 1);
  const uint4* wr = reinterpret_cast<const uint4*>(s.ring[slot] + row_off);
#pragma unroll
  for (int u = 0; u < DF_CH; ++u) {
    dst[u] = wr[cc[u]];
    if (lane + 32 * u >= nch) dst[u] = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(&s.empty[slot]);             // the row is in registers: the producer may refill the slot
}
// the lane's part of the dot product of that row with its chunks of the vector
__device__ __forceinline__ float df_dot(const uint4 (&a)[DF_CH], const uint4 (&x)[DF_CH]) {
  unsigned long long acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
#pragma unroll
  for (int u = 0; u < DF_CH; ++u) {
    acc0 = df_ffma2(df_f2(a[u].x), df_f2(x[u].x), acc0);
    acc1 = df_ffma2(df_f2(a[u].y), df_f2(x[u].y), acc1);
    acc2 = df_ffma2(df_f2(a[u].z), df_f2(x[u].z), acc2);
    acc3 = df_ffma2(df_f2(a[u].w), df_f2(x[u].w), acc3);
  }
  return ((__uint_as_float(static_cast<uint32_t>(acc0)) + __uint_as_float(static_cast<uint32_t>(acc0 >> 32))) +
          (__uint_as_float(static_cast<uint32_t>(acc1)) + __uint_as_float(static_cast<uint32_t>(acc1 >> 32)))) +
         ((__uint_as_float(static_cast<uint32_t>(acc2)) + __uint_as_float(static_cast<uint32_t>(acc2 >> 32))) +
          (__uint_as_float(static_cast<uint32_t>(acc3)) + __uint_as_float(static_cast<uint32_t>(acc3 >> 32))));
}
__device__ __forceinline__ void df_consume(DfSmem& s, const DfPhase& ph, uint32_t q, int i0, int i1, float* ptab,
                                           int warp, int lane, long long* t_dbg = nullptr) {
  if (warp >= DF_TILE || i0 >= i1) return;
  const int row_off = warp * (2 * ph.KS + 64);
  const int nch = ph.KS >> 3, n_slabs = ph.n_slabs;
  const uint4* xs = reinterpret_cast<const uint4*>(s.vbuf);
  int cc[DF_CH];
#pragma unroll
  for (int u = 0; u < DF_CH; ++u) cc[u] = min(lane + 32 * u, nch - 1);
  // the lane's chunks of the vector stay in registers while the slab does not change (K = 1536: the whole phase)
  uint4 x[DF_CH];
  int slab = i0 % n_slabs;
#pragma unroll
  for (int u = 0; u < DF_CH; ++u) x[u] = xs[slab * nch + cc[u]];
  float* pt = ptab + warp * DF_PP + lane;
  for (int i = i0; i < i1; ++i) {
    long long tw0 = 0;
    if (t_dbg) tw0 = clock64();
    uint4 a[DF_CH];
    df_fetch(s, q + i, row_off, cc, nch, lane, a);
    // the lanes' partial sums go to the partial table as they are: a shuffle tree per stage is a 150-cycle dependent
    // chain in the warp that has to keep up with the stream
    pt[(i - i0) * DF_TILE * DF_PP] = df_dot(a, x);
    if (n_slabs > 1) {
      if (++slab == n_slabs) slab = 0;
#pragma unroll
      for (int u = 0; u < DF_CH; ++u) x[u] = xs[slab * nch + cc[u]];
    }
    if (t_dbg) { t_dbg[1] += clock64() - tw0; t_dbg[2] += 1; }
  }
}
// sum of local row r of a round that started at tile t0, by one thread (after a csync()): slabs, then lanes
__device__ __noinline__ float df_row_sum(const float* ptab, int n_slabs, int t0, int r) {
  const int tile = r / DF_TILE - t0, i = r % DF_TILE;
  float f = 0.f;
  for (int sl = 0; sl < n_slabs; ++sl) {
    const float* pp = ptab + ((tile * n_slabs + sl) * DF_TILE + i) * DF_PP;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k += 4) { f0 += pp[k]; f1 += pp[k + 1]; f2 += pp[k + 2]; f3 += pp[k + 3]; }
    f += (f0 + f1) + (f2 + f3);
  }
  return f;
}
// One GEMV phase of the consumers: rounds of <= DF_ROUND stages; after each round one thread per row of the round's
// tiles (per unit for the gate/up matrix: gate row r and up row nu + r) sums the row and calls epi(r, sum, up_sum).
template <typename Epi>
__device__ __forceinline__ void df_gemv(DfSmem& s, const DfPhase& ph, uint32_t& q, float* ptab, int warp, int lane, Epi epi,
                                        long long* t_dbg = nullptr) {
  const int tiles_per_round = max(1, DF_ROUND / ph.n_slabs);
  for (int t0 = 0; t0 < ph.n_tiles; t0 += tiles_per_round) {
    const int t1 = min(ph.n_tiles, t0 + tiles_per_round);
    long long tc0 = 0;
    if (t_dbg) tc0 = clock64();
    df_consume(s, ph, q, t0 * ph.n_slabs, t1 * ph.n_slabs, ptab, warp, lane, t_dbg);
    csync();
    if (t_dbg) t_dbg[4] += clock64() - tc0;
    const int r_end = ph.swiglu ? ph.nu : min(t1 * DF_TILE, ph.count);
    for (int r = t0 * DF_TILE + threadIdx.x; r < r_end; r += DF_CTHREADS)
      epi(r, df_row_sum(ptab, ph.n_slabs, t0, r), ph.swiglu ? df_row_sum(ptab, ph.n_slabs, t0, ph.nu + r) : 0.f);
    if (t1 < ph.n_tiles) csync();
  }
  q += ph.n_tiles * ph.n_slabs;
}

template <bool TIMING>
__global__ void __launch_bounds__(DF_THREADS, 1) und_decode_fused_kernel(const __grid_constant__ DecFusedParams p) {
  extern __shared__ __align__(128) uint8_t df_smem_raw[];
  DfSmem& s = *reinterpret_cast<DfSmem*>(df_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, grid = gridDim.x;
  const int H = p.H, I = p.I, nq = p.nq, nkv = p.nkv, G = nq / nkv;
  const int qkv_w = (nq + 2 * nkv) * 128, kvw = 2 * nkv * 128;
  if (tid == 0) {
    for (int i = 0; i < DF_NS; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], DF_TILE); }
    fence_barrier_init();
  }
  __syncthreads();

  // ================= producer warp: the whole step's weights, in order ==========================================
  if (warp == DF_CWARPS) {
    uint32_t q = 0;
    const int n_phases = 4 * p.num_layers + 1;
    auto phase_at = [&](int idx) {
      if (idx >= 4 * p.num_layers) return df_phase(p.lm_head, p.vocab, H, p.ks_h, false);
      const g2vlm_und_layer_weights& w = p.layers[idx >> 2];
      switch (idx & 3) {
        case 0: return df_phase(w.wqkv, qkv_w, H, p.ks_h, false);
        case 1: return df_phase(w.wo, H, nq * 128, p.ks_a, false);
        case 2: return df_phase(w.wgu, I, H, p.ks_h, true);
        default: return df_phase(w.wdown, H, I, p.ks_i, false);
      }
    };
    for (int idx = 0; idx < n_phases; ++idx) {
      df_produce(s, phase_at(idx), q, lane);
    }
    return;
  }

  // ================= consumer warps ==========================================================================
  uint32_t q = 0;                                   // stages consumed so far (same count as the producer's)
  float* ptab = reinterpret_cast<float*>(s.kbuf);
  uint4* xs = reinterpret_cast<uint4*>(s.vbuf);
  unsigned target = __ldcg(p.sync + 1);
  const bool fences = (p.opt & 1) == 0;
  // TIMING (tools/decode_phase_trace.py): SM cycles CTA 0 and the last CTA spend in each phase / at each barrier
  long long t_prev = 0;
  long long* t_acc = s.t_acc;
  long long* t_gu = s.t_gu;
  unsigned long long ns0 = 0;
  if constexpr (TIMING) {
    if (tid == 0) {
      for (int i = 0; i < 14; ++i) t_acc[i] = 0;
      for (int i = 0; i < 5; ++i) t_gu[i] = 0;
      for (int i = 0; i < 6; ++i) s.t_at[i] = 0;
    }
    t_prev = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
  }
  auto mark = [&](int slot) {
    if constexpr (TIMING) {
      const long long t = clock64();
      if (tid == 0) t_acc[slot] += t - t_prev;
      t_prev = t;
    }
  };
  long long t_at_prev = 0;
  auto amark = [&](int slot) {                      // attention sub-phases (TIMING build)
    if constexpr (TIMING) {
      // (bar.sync does not block at issue: a load of barrier-protected state does, so the clock is read after one)
      const float dummy = *reinterpret_cast<volatile float*>(&s.part[0]);
      long long t = clock64();
      if (dummy == 1.2345e-30f) t += 1;
      if (tid == 0 && slot >= 0) s.t_at[slot] += t - t_at_prev;
      t_at_prev = t;
    }
  };
  const int L = *p.cache_len;                       // keys already in the cache; this step's key is row L
  const long long token = *p.cur_token;
  const long long pos0 = p.position[0], pos1 = p.position[1], pos2 = p.position[2];
  if (tid < 64) {                                   // mrope_table_kernel's arithmetic, once per step instead of per layer
    const long long pos = tid < p.s0 ? pos0 : (tid < p.s0 + p.s1 ? pos1 : pos2);
    sincosf(static_cast<float>(pos) * __ldg(p.inv_freq + tid), &s.rope_sn[tid], &s.rope_cs[tid]);
  }
  const float* x_in = p.embed + token * H;          // layer 0 reads the embedding row; CTA 0 copies it into x
  if (cta == 0)
    for (int i = tid; i < H; i += DF_CTHREADS) p.x[i] = __ldg(x_in + i);

  // attention geometry of this CTA: KV head, key range [k0, k1) of the L + 1 keys
  const int per_kvh = grid / nkv;                  // CTAs per KV head (the grid % nkv last CTAs idle in that phase)
  const bool at_active = cta < per_kvh * nkv;
  const int kvh = at_active ? cta / per_kvh : 0, split = cta % per_kvh;
  const int Lk = L + 1;
  const int span = (Lk + per_kvh - 1) / per_kvh;
  const int k0 = split * span, k1 = at_active ? min(Lk, k0 + span) : 0, nk = max(0, k1 - k0);
  const int nst = (nk + DF_STAGE - 1) / DF_STAGE;
  const int sub = nst > 0 ? (((nk + nst - 1) / nst + 15) & ~15) : 0;   // keys per stage (<= DF_STAGE, multiple of 16)
  const bool owner = at_active && nk > 0 && k1 == Lk;                    // this CTA's range ends with the new key

  for (int l = 0; l < p.num_layers; ++l) {
    const g2vlm_und_layer_weights& w = p.layers[l];
    __nv_bfloat16* kvbuf = p.kv[l];
    const __nv_bfloat16* kc = kvbuf + kvh * 128;
    const __nv_bfloat16* vc = kvbuf + nkv * 128 + kvh * 128;
    // K|V rows [c0, c1) of this CTA's range that already sit in the cache -> shared memory (swizzled), one commit group
    auto load_stage = [&](int st) {
      const int c0 = k0 + st * sub, c1 = min(k1, c0 + sub);
      const int n_old = max(0, min(c1, L) - c0);
      for (int i = tid; i < n_old * 16; i += DF_CTHREADS) {
        const int r = i >> 4, c = i & 15;
        const int off = (st % DF_NKV) * (DF_STAGE * 256) + r * 256 + ((c ^ (r & 7)) << 4);
        cp_async16_df(s.kbuf + off, kc + (long long)(c0 + r) * kvw + c * 8);
        cp_async16_df(s.vbuf + off, vc + (long long)(c0 + r) * kvw + c * 8);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ================= phase 1: RMSNorm + qkv GEMV (+bias) =====================================================
    {
      const DfPhase ph = df_phase(w.wqkv, qkv_w, H, p.ks_h, false);
      // what the epilogue thread of row `tid` needs from global memory is requested now, not after the stream
      const float bias_r = (w.bqkv && tid < ph.count) ? __ldg(w.bqkv + ph.first + tid) : 0.f;
      rmsnorm_to_smem(l == 0 ? x_in : p.x, w.input_norm, H, p.eps, xs, s.part);
      df_gemv(s, ph, q, ptab, warp, lane, [&](int r, float f, float) {
        const float b = r == tid ? bias_r : (w.bqkv ? __ldg(w.bqkv + ph.first + r) : 0.f);
        p.qkv[ph.first + r] = __float2bfloat16_rn(f + b);
      });
    }
    mark(0);
    barrier_arrive(p.sync, target, fences);                    // (its bar.sync also ends the reads of the partial table)
    if (nst > 0) load_stage(0);                        // the cache rows do not depend on this step: load across the barrier
    if (DF_NKV > 1 && nst > 1) load_stage(1);
    barrier_wait(p.sync, target, fences);
    mark(1);
    // ================= phase 2: q/k-norm + M-RoPE, K|V append, attention partials over this CTA's key range =====
    amark(-1);
    {
      // q heads of this KV head (warps 0..G-1), the new key (warp G of the owner), the new value (warp G+1)
      if (at_active && warp <= G) {
        const bool is_k = warp == G;
        if (!is_k || owner) {
          const int col = is_k ? (nq + kvh) * 128 : (kvh * G + warp) * 128;
          const uint2 raw = __ldcg(reinterpret_cast<const uint2*>(p.qkv + col + lane * 4));
          const __nv_bfloat162 v01 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
          const __nv_bfloat162 v23 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
          const float v[4] = {__low2float(v01), __high2float(v01), __low2float(v23), __high2float(v23)};
          const float ss = warp_sum(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
          const float r = rsqrtf(ss / 128.0f + p.eps);
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(is_k ? w.k_norm : w.q_norm) + lane);
          const float gw[4] = {g4.x, g4.y, g4.z, g4.w};
          const int j0 = (lane & 15) * 4;
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float sn = s.rope_sn[j0 + e], cs = s.rope_cs[j0 + e];   // (written before the first grid barrier)
            const float nv = gw[e] * bf16_round(v[e] * r);
            const float partner = __shfl_xor_sync(0xffffffffu, nv, 16);
            const float rot = lane < 16 ? -partner : partner;
            o[e] = nv * cs + rot * sn;
          }
          const uint2 packed = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
          if (is_k) {
            *reinterpret_cast<uint2*>(s.new_k + lane * 4) = packed;
            *reinterpret_cast<uint2*>(kvbuf + (long long)L * kvw + kvh * 128 + lane * 4) = packed;
          } else {
            *reinterpret_cast<uint2*>(&s.q[warp][lane * 4]) = packed;
          }
        }
      } else if (owner && warp == G + 1) {
        const uint2 raw = __ldcg(reinterpret_cast<const uint2*>(p.qkv + (nq + nkv + kvh) * 128 + lane * 4));
        *reinterpret_cast<uint2*>(s.new_v + lane * 4) = raw;
        *reinterpret_cast<uint2*>(kvbuf + (long long)L * kvw + (nkv + kvh) * 128 + lane * 4) = raw;
      }
      for (int i = tid; i < (8 - G) * 64; i += DF_CTHREADS)                          // unused head rows of the B operand
        reinterpret_cast<uint32_t*>(&s.q[G][0])[i] = 0u;
      if (tid < 8) { s.m_run[tid] = -INFINITY; s.l_run[tid] = 0.f; }
      csync();                                     // s.q / s.new_k / s.new_v / m_run are read below
      amark(0);
      float o_acc[4] = {0.f, 0.f, 0.f, 0.f};              // O^T[d = 16 dt + g (+8)][head = 2 t (+1)] of this warp's key half
      const int g8 = lane >> 2, t4 = lane & 3;
      // P.V: warp = (d tile dt, key half khalf); d tile 7 has no second warp (15 consumer warps) and takes every key tile
      const int dt = warp & 7, khalf = warp >> 3, kstep = dt < 7 ? 2 : 1;
      for (int st = 0; st < nst; ++st) {
        const int c0 = k0 + st * sub, c1 = min(k1, c0 + sub), nkc = c1 - c0;
        const int nkc16 = (nkc + 15) & ~15;
        const int n_old = max(0, min(c1, L) - c0);
        uint8_t* kst = s.kbuf + (st % DF_NKV) * (DF_STAGE * 256);
        uint8_t* vst = s.vbuf + (st % DF_NKV) * (DF_STAGE * 256);
        // rows the cp.async round does not write: the new key (owner) and the zero rows that pad the last 16-key tile
        for (int idx = tid; idx < 32 * (nkc16 - n_old); idx += DF_CTHREADS) {
          const int r = n_old + (idx >> 5), c = (idx & 31) >> 1, hsel = idx & 1;   // 16 lanes x 16 B per row, K and V
          {
            uint4 val = make_uint4(0, 0, 0, 0);
            if (r < nkc) val = reinterpret_cast<const uint4*>(hsel ? s.new_v : s.new_k)[c];   // only the new key is >= n_old
            *reinterpret_cast<uint4*>((hsel ? vst : kst) + r * 256 + ((c ^ (r & 7)) << 4)) = val;
          }
        }
        if (DF_NKV > 1 && st + 1 < nst) asm volatile("cp.async.wait_group 1;" ::: "memory");   // stage st + 1 stays in flight
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        csync();
        amark(1);
        // scores S[key, head] = K . q^T on mma.sync m16n8k16: A = 16 keys x 16 dims (ldmatrix), B = q (registers)
        {
          uint32_t qf[8][2];
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            qf[ks][0] = *reinterpret_cast<const uint32_t*>(&s.q[g8][16 * ks + 2 * t4]);
            qf[ks][1] = *reinterpret_cast<const uint32_t*>(&s.q[g8][16 * ks + 2 * t4 + 8]);
          }
          for (int mt = warp; mt * 16 < nkc; mt += DF_CWARPS) {
            // four independent accumulators instead of a chain of 8 dependent MMAs
            float c4[4] = {0.f, 0.f, 0.f, 0.f}, c5[4] = {0.f, 0.f, 0.f, 0.f}, c6[4] = {0.f, 0.f, 0.f, 0.f},
                  c7[4] = {0.f, 0.f, 0.f, 0.f};
            const int row = mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            const uint32_t rbase = smem_u32(kst) + row * 256;
            uint32_t a[8][4];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) ldmatrix_x4(a[ks], rbase + (((2 * ks + (lane >> 4)) ^ (row & 7)) << 4));
            mma_16816(c4, a[0][0], a[0][1], a[0][2], a[0][3], qf[0][0], qf[0][1]);
            mma_16816(c5, a[1][0], a[1][1], a[1][2], a[1][3], qf[1][0], qf[1][1]);
            mma_16816(c6, a[2][0], a[2][1], a[2][2], a[2][3], qf[2][0], qf[2][1]);
            mma_16816(c7, a[3][0], a[3][1], a[3][2], a[3][3], qf[3][0], qf[3][1]);
            mma_16816(c4, a[4][0], a[4][1], a[4][2], a[4][3], qf[4][0], qf[4][1]);
            mma_16816(c5, a[5][0], a[5][1], a[5][2], a[5][3], qf[5][0], qf[5][1]);
            mma_16816(c6, a[6][0], a[6][1], a[6][2], a[6][3], qf[6][0], qf[6][1]);
            mma_16816(c7, a[7][0], a[7][1], a[7][2], a[7][3], qf[7][0], qf[7][1]);
#pragma unroll
            for (int e = 0; e < 4; ++e) c4[e] = (c4[e] + c5[e]) + (c6[e] + c7[e]);
            const int key = mt * 16 + g8;
            if (key < nkc) { s.score[2 * t4][key] = c4[0] * p.scale_log2; s.score[2 * t4 + 1][key] = c4[1] * p.scale_log2; }
            if (key + 8 < nkc) { s.score[2 * t4][key + 8] = c4[2] * p.scale_log2; s.score[2 * t4 + 1][key + 8] = c4[3] * p.scale_log2; }
          }
        }
        csync();
        amark(2);
        // per head: stage max, running max, exp2, sums (one warp per head); rows >= G and keys >= nkc become 0
        if (warp < 8) {
          if (warp < G) {
            float m = -INFINITY;
            for (int j = lane; j < nkc; j += 32) m = fmaxf(m, s.score[warp][j]);
            m = warp_max(m);
            const float m_old = s.m_run[warp];
            const float m_new = fmaxf(m_old, m);
            float lsum = 0.f;
            for (int j = lane; j < nkc16; j += 32) {
              const float pj = j < nkc ? exp2f(s.score[warp][j] - m_new) : 0.f;
              s.score[warp][j] = pj;
              lsum += pj;
            }
            lsum = warp_sum(lsum);
            if (lane == 0) {
              const float al = m_old == -INFINITY ? 0.f : exp2f(m_old - m_new);
              s.alpha[warp] = al;
              s.m_run[warp] = m_new;
              s.l_run[warp] = s.l_run[warp] * al + lsum;
            }
          } else {
            for (int j = lane; j < nkc16; j += 32) s.score[warp][j] = 0.f;
            if (lane == 0) s.alpha[warp] = 0.f;
          }
        }
        csync();
        amark(3);
        // O^T[d, head] += V^T[d, key] . P^T[key, head]: A = V^T (ldmatrix.trans of the [key][d] rows), B = bf16 P
        {
          const float a0 = s.alpha[2 * t4], a1 = s.alpha[2 * t4 + 1];
          o_acc[0] *= a0; o_acc[1] *= a1; o_acc[2] *= a0; o_acc[3] *= a1;
          float o_b[4] = {0.f, 0.f, 0.f, 0.f};              // second accumulator: two independent MMA chains
          int par = 0;
          for (int kt = khalf; kt * 16 < nkc; kt += kstep, par ^= 1) {
            const int row = kt * 16 + (lane & 7) + 8 * (lane >> 4);
            uint32_t a[4];
            ldmatrix_x4_trans(a, smem_u32(vst) + row * 256 + (((2 * dt + ((lane >> 3) & 1)) ^ (row & 7)) << 4));
            const float2 p0 = *reinterpret_cast<const float2*>(&s.score[g8][kt * 16 + 2 * t4]);
            const float2 p1 = *reinterpret_cast<const float2*>(&s.score[g8][kt * 16 + 2 * t4 + 8]);
            if (par == 0) mma_16816(o_acc, a[0], a[1], a[2], a[3], pack_bf16x2(p0.x, p0.y), pack_bf16x2(p1.x, p1.y));
            else mma_16816(o_b, a[0], a[1], a[2], a[3], pack_bf16x2(p0.x, p0.y), pack_bf16x2(p1.x, p1.y));
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) o_acc[e] += o_b[e];
        }
        csync();                                   // every read of this stage is done
        amark(4);
        if (st + DF_NKV < nst) load_stage(st + DF_NKV);
      }
      if (dt == 7) {                                       // no second warp for this d tile: its half stays zero
        s.red[1][2 * t4][112 + g8] = 0.f; s.red[1][2 * t4 + 1][112 + g8] = 0.f;
        s.red[1][2 * t4][120 + g8] = 0.f; s.red[1][2 * t4 + 1][120 + g8] = 0.f;
      }
      s.red[khalf][2 * t4][16 * dt + g8] = o_acc[0];
      s.red[khalf][2 * t4 + 1][16 * dt + g8] = o_acc[1];
      s.red[khalf][2 * t4][16 * dt + g8 + 8] = o_acc[2];
      s.red[khalf][2 * t4 + 1][16 * dt + g8 + 8] = o_acc[3];
      csync();
      if (at_active) {
        for (int i = tid; i < G * 128; i += DF_CTHREADS) {
          const int h = i >> 7, d = i & 127;
          float* dst = p.part + ((long long)split * nq + kvh * G + h) * DF_PART;
          dst[d] = nk > 0 ? s.red[0][h][d] + s.red[1][h][d] : 0.f;
          if (d == 0) { dst[128] = nk > 0 ? s.m_run[h] : -INFINITY; dst[129] = nk > 0 ? s.l_run[h] : 0.f; }
        }
      }
    }
    amark(5);
    mark(2);
    barrier_arrive(p.sync, target, fences);
    barrier_wait(p.sync, target, fences);
    mark(3);

    // ================= phase 3: merge of the partials, spread over the grid: CTA = (head, slice of d) ==========
    {
      const int S = grid / nkv;                          // partials per head
      const int n_slices = grid / nq;                    // CTAs per head
      const int dps = (128 + n_slices - 1) / n_slices;   // d per CTA (<= 16)
      const int head = cta / n_slices, slice = cta % n_slices;
      if (head < nq && slice * dps < 128) {
        // every thread pulls its splits' (m, l, o[d]) in ONE round trip to L2, then the block agrees on the maximum
        const float* base = p.part + (long long)head * DF_PART;
        const long long stride = (long long)nq * DF_PART;
        const int dl = tid & 15, sl = tid >> 4;
        const int d = slice * dps + dl;
        const bool d_ok = dl < dps && d < 128;
        constexpr int MS = 4;                              // splits per thread held in registers (S <= 4 * DF_MSL = 120)
        float pm[MS], pl[MS], po[MS];
#pragma unroll
        for (int k = 0; k < MS; ++k) {
          const int sp = sl + k * DF_MSL;
          const float* pp = base + min(sp, S - 1) * stride;
          pm[k] = sp < S ? __ldcg(pp + 128) : -INFINITY;
          pl[k] = sp < S ? __ldcg(pp + 129) : 0.f;
          po[k] = (sp < S && d_ok) ? __ldcg(pp + d) : 0.f;
        }
        float M = fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3]));
        M = warp_max(M);
        float* wmax = &s.red[0][0][0];                   // [16]
        float* racc = wmax + 32;                         // [DF_MSL][16]
        float* rl = racc + DF_MSL * 16;                  // [DF_MSL]
        if (lane == 0) wmax[warp] = M;
        csync();
        M = wmax[0];
#pragma unroll
        for (int i = 1; i < DF_CWARPS; ++i) M = fmaxf(M, wmax[i]);
        float acc = 0.f, lt = 0.f;
#pragma unroll
        for (int k = 0; k < MS; ++k) {
          const float wgt = pl[k] > 0.f ? exp2f(pm[k] - M) : 0.f;
          acc = fmaf(wgt, po[k], acc);
          lt = fmaf(wgt, pl[k], lt);
        }
        for (int sp = sl + MS * DF_MSL; sp < S; sp += DF_MSL) {   // (more than 120 splits per head: not on a 148-SM part)
          const float* pp = base + sp * stride;
          const float ps = __ldcg(pp + 129);
          const float wgt = ps > 0.f ? exp2f(__ldcg(pp + 128) - M) : 0.f;
          if (d_ok) acc = fmaf(wgt, __ldcg(pp + d), acc);
          lt = fmaf(wgt, ps, lt);
        }
        racc[sl * 16 + dl] = acc;
        if (dl == 0) rl[sl] = lt;
        csync();
        if (tid < 16 && tid < dps && slice * dps + tid < 128) {
          float a = 0.f, lsum = 0.f;
#pragma unroll 6
          for (int i = 0; i < DF_MSL; ++i) { a += racc[i * 16 + tid]; lsum += rl[i]; }
          p.attn[head * 128 + slice * dps + tid] = __float2bfloat16_rn(lsum > 0.f ? a / lsum : 0.f);
        }
      }
    }
    mark(4);
    barrier_arrive(p.sync, target, fences);
    barrier_wait(p.sync, target, fences);
    mark(5);

    // ================= phase 4: o_proj GEMV, x += bf16(acc) ====================================================
    {
      const DfPhase ph = df_phase(w.wo, H, nq * 128, p.ks_a, false);
      const float x_r = tid < ph.count ? __ldcg(p.x + ph.first + tid) : 0.f;   // this CTA's rows: nobody else writes them
      vec_from_bf16(xs, p.attn, (nq * 128) >> 3);
      df_gemv(s, ph, q, ptab, warp, lane, [&](int r, float f, float) {
        const int n = ph.first + r;
        p.x[n] = (r == tid ? x_r : __ldcg(p.x + n)) + bf16_round(f);
      });
    }
    mark(6);
    barrier_arrive(p.sync, target, fences);
    barrier_wait(p.sync, target, fences);
    mark(7);

    // ================= phase 5: RMSNorm + gate/up GEMV + SwiGLU ================================================
    {
      const DfPhase ph = df_phase(w.wgu, I, H, p.ks_h, true);
      long long tg0 = 0;
      if constexpr (TIMING) tg0 = clock64();
      rmsnorm_to_smem(p.x, w.post_norm, H, p.eps, xs, s.part);
      if constexpr (TIMING) { if (tid == 0) t_gu[3] += clock64() - tg0; }
      df_gemv(s, ph, q, ptab, warp, lane, [&](int r, float gate, float up) {   // one round (checked on the host)
        const float g = bf16_round(gate), u = bf16_round(up);
        const float sg = bf16_round(g / (1.0f + __expf(-g)));
        p.act[ph.first + r] = __float2bfloat16_rn(sg * u);
      }, (TIMING && tid == 0) ? t_gu : nullptr);
    }
    mark(8);
    barrier_arrive(p.sync, target, fences);
    barrier_wait(p.sync, target, fences);
    mark(9);

    // ================= phase 6: down GEMV, x += bf16(acc) ======================================================
    {
      const DfPhase ph = df_phase(w.wdown, H, I, p.ks_i, false);
      const float x_r = tid < ph.count ? __ldcg(p.x + ph.first + tid) : 0.f;
      vec_from_bf16(xs, p.act, I >> 3);
      df_gemv(s, ph, q, ptab, warp, lane, [&](int r, float f, float) {
        const int n = ph.first + r;
        p.x[n] = (r == tid ? x_r : __ldcg(p.x + n)) + bf16_round(f);
      });
    }
    mark(10);
    barrier_arrive(p.sync, target, fences);
    barrier_wait(p.sync, target, fences);
    mark(11);
  }

  // ================= final RMSNorm -> bf16, lm_head GEMV, argmax ===============================================
  {
    const DfPhase ph = df_phase(p.lm_head, p.vocab, H, p.ks_h, false);
    rmsnorm_to_smem(p.x, p.final_norm, H, p.eps, xs, s.part);
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    df_gemv(s, ph, q, ptab, warp, lane, [&](int r, float f_row, float) {
      const int n = ph.first + r;
      const __nv_bfloat16 b = __float2bfloat16_rn(f_row);
      p.logits[n] = b;
      const float f = __bfloat162float(b);
      if (argmax_better_df(f, n, best_v, best_i)) { best_v = f; best_i = n; }
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (argmax_better_df(ov, oi, best_v, best_i)) { best_v = ov; best_i = oi; }
    }
    if (lane == 0) { s.best_v[warp] = best_v; s.best_i[warp] = best_i; }
    csync();
    if (tid == 0) {
      for (int i = 1; i < DF_CWARPS; ++i)
        if (argmax_better_df(s.best_v[i], s.best_i[i], best_v, best_i)) { best_v = s.best_v[i]; best_i = s.best_i[i]; }
      p.cand[2 * cta] = best_v;
      p.cand[2 * cta + 1] = __int_as_float(best_i);
    }
  }
  mark(12);
  barrier_arrive(p.sync, target, fences);
  barrier_wait(p.sync, target, fences);
  mark(13);
  if constexpr (TIMING) {
    if (tid == 0 && (cta == 0 || cta == grid - 1)) {
      long long* dst = reinterpret_cast<long long*>(p.sync + 16) + (cta == 0 ? 0 : 15);
      for (int i = 0; i < 14; ++i) dst[i] = t_acc[i];
      unsigned long long ns1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
      dst[14] = static_cast<long long>(ns1 - ns0);
      if (cta == 0) { for (int i = 0; i < 6; ++i) (reinterpret_cast<long long*>(p.sync + 16) + 35)[i] = s.t_at[i]; }
      if (cta == 0) { long long* d2 = reinterpret_cast<long long*>(p.sync + 16) + 30; d2[0] = t_gu[0]; d2[1] = t_gu[1]; d2[2] = t_gu[2]; d2[3] = t_gu[3]; d2[4] = t_gu[4]; }
    }
  }
  if (cta == 0 && warp == 0) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < grid; i += 32) {
      const float v = __ldcg(p.cand + 2 * i);
      const int ix = __float_as_int(__ldcg(p.cand + 2 * i + 1));
      if (argmax_better_df(v, ix, bv, bi)) { bv = v; bi = ix; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (argmax_better_df(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      if (!p.keep_token) *p.cur_token = bi;
      p.position[0] = pos0 + 1; p.position[1] = pos1 + 1; p.position[2] = pos2 + 1;
      *p.cache_len = L + 1;
      p.sync[1] = target;                                  // every CTA of the next launch starts from here
    }
  }
}

}  // namespace g2

extern "C" int64_t g2vlm_und_decode_workspace_bytes(int32_t n_q_heads, int32_t n_kv_heads) {
  using namespace g2;
  if (n_q_heads <= 0 || n_kv_heads <= 0) return 0;
  const long long grid = num_sms();
  return 512 + grid * 2 * 4 + (grid / n_kv_heads + 1) * (long long)n_q_heads * DF_PART * 4;
}

namespace g2 {

// The fused step; called by g2vlm_und_decode_step when the caller provides `fused_ws`.
int launch_decode_fused(const g2vlm_decode_step_args* a, cudaStream_t stream) {
  const int H = a->hidden, I = a->intermediate, nq = a->n_q_heads, nkv = a->n_kv_heads;
  G2_REQUIRE(a->num_layers <= DF_MAX_LAYERS, "decode_step (fused): at most 32 layers");
  G2_REQUIRE(H % 8 == 0 && I % 128 == 0 && H <= DF_MAX_VEC && I <= DF_MAX_VEC && nq * 128 <= DF_MAX_VEC,
             "decode_step (fused): hidden / intermediate size not supported");
  G2_REQUIRE(nq % nkv == 0 && nq / nkv <= DF_MAX_G, "decode_step (fused): at most 8 query heads per KV head");
  const int grid = num_sms();
  G2_REQUIRE(grid >= 8 * nq && grid >= nkv, "decode_step (fused): too few SMs for this head count");
  G2_REQUIRE((reinterpret_cast<uintptr_t>(a->fused_ws) & 15) == 0 &&
                 a->fused_ws_bytes >= g2vlm_und_decode_workspace_bytes(nq, nkv),
             "decode_step (fused): workspace too small (g2vlm_und_decode_workspace_bytes) or misaligned");
  DecFusedParams q;
  memset(&q, 0, sizeof(q));
  q.num_layers = a->num_layers; q.H = H; q.I = I; q.nq = nq; q.nkv = nkv; q.vocab = a->vocab;
  q.vpad = (a->vocab + 7) / 8 * 8;
  q.eps = a->rms_eps;
  q.scale_log2 = static_cast<float>(1.0 / sqrt(static_cast<double>(a->head_dim))) * 1.4426950408889634f;
  q.s0 = a->mrope_s0; q.s1 = a->mrope_s1;
  q.keep_token = a->keep_token;
  // columns per ring stage: the largest divisor of K that is a multiple of 64 and <= 1536
  auto stage_cols = [](int K) {
    for (int ks = DF_KS_MAX; ks >= 64; ks -= 64)
      if (K % ks == 0) return ks;
    return 0;
  };
  q.ks_h = stage_cols(H); q.ks_a = stage_cols(nq * 128); q.ks_i = stage_cols(I);
  G2_REQUIRE(q.ks_h > 0 && q.ks_a > 0 && q.ks_i > 0, "decode_step (fused): hidden / intermediate size must be a multiple of 64");
  // the gate/up phase of a CTA must fit one round of the partial table (its epilogue pairs gate and up rows)
  G2_REQUIRE(2 * ((I + grid - 1) / grid) + DF_TILE <= DF_ROUND * DF_TILE && H / q.ks_h == 1,
             "decode_step (fused): intermediate size too large for the per-CTA partial table");
  for (int l = 0; l < a->num_layers; ++l) {
    q.layers[l] = a->layers[l];
    q.kv[l] = reinterpret_cast<__nv_bfloat16*>(a->kv[l]);
  }
  q.kv_capacity = a->kv_capacity;
  q.embed = a->embed; q.final_norm = a->final_norm;
  q.lm_head = reinterpret_cast<const __nv_bfloat16*>(a->lm_head);
  q.inv_freq = a->inv_freq;
  q.cur_token = reinterpret_cast<long long*>(a->cur_token);
  q.position = reinterpret_cast<long long*>(a->position);
  q.cache_len = a->cache_len;
  q.x = a->x;
  q.qkv = reinterpret_cast<__nv_bfloat16*>(a->qkv);
  q.attn = reinterpret_cast<__nv_bfloat16*>(a->attn);
  q.act = reinterpret_cast<__nv_bfloat16*>(a->act);
  q.logits = reinterpret_cast<__nv_bfloat16*>(a->logits);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->fused_ws);
  q.sync = reinterpret_cast<unsigned*>(ws);
  q.cand = reinterpret_cast<float*>(ws + 512);
  q.part = reinterpret_cast<float*>(ws + 512 + (long long)grid * 8);
  const int smem = static_cast<int>(sizeof(DfSmem));
  const char* opt = getenv("G2VLM_DECODE_OPT");
  q.opt = opt ? atoi(opt) : 1;
  const char* trace = getenv("G2VLM_DECODE_TRACE");   // phase cycle counters at fused_ws + 64 (tools/decode_phase_trace.py)
  const bool timing = trace != nullptr && atoi(trace) != 0;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(und_decode_fused_kernel<false>), smem)) return rc;
  if (timing)
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(und_decode_fused_kernel<true>), smem)) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(DF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (timing) G2_CUDA_OK(cudaLaunchKernelEx(&cfg, und_decode_fused_kernel<true>, q));
  else G2_CUDA_OK(cudaLaunchKernelEx(&cfg, und_decode_fused_kernel<false>, q));
  return G2VLM_OK;
}

}  // namespace g2
