// g2vlm_b200 — varlen / GQA attention on tcgen05 tensor cores (QK^T and PV), accumulators in TMEM.
//
// Replaces flash_attn_varlen_func (modeling/g2vlm/qwen2vl.py:643-652, dinov2_model.py:49-58) and the
// per-view SDPA calls of the Pi3 decoders (modeling/pi3/models/layers/attention.py:255-259,370-372).
//
// Persistent kernel, one CTA per SM.  A unit of work = one work item (<= 256 query rows of one segment) x
// one query head; a CTA walks units blockIdx.x, +gridDim.x, ... (head fastest) with TMEM, barriers and the K/V ring kept
// alive across units (the producer prefetches the next unit while this one drains).  Inside a unit two
// 128-row query tiles are in flight ("ping-pong"): while the softmax warps of tile 0 turn S0 into P0, the
// tensor core runs tile 1's MMAs and vice versa.
//   warp 0        TMA producer: Q tiles once, K/V blocks of 128 keys through an mbarrier ring
//   warps 1, 2    MMA issuers (one per query tile): S_t = Q_t K^T (SS, K-major operands) ; O_t += P_t V (TS: P
//                               read from TMEM, V is an MN-major B operand straight from the [key][d] layout)
//   warps 4-7     softmax of tile 0 (one thread per query row), warps 8-11 softmax of tile 1:
//                 tcgen05.ld S -> running max / exp2 / row sum -> bf16 P written back to TMEM over S
//                 (tcgen05.st); lazy rescale of O in TMEM only when the running max grows by > 2^8
//                 (exact: the stale max cancels in the final 1/l normalisation); final O/l -> global.
//                 P reaches the MMA warp in two pieces, keys [0,96) and [96,128) of the block (own mbarrier
//                 each), so the PV MMAs of the first piece run while the last keys are still in exp2.  Measured and rejected on B200 (r01): moving
//                 25/50 % of the exp2 to a degree-3 FMA-pipe polynomial (-1 % / -6 %: the extra issue
//                 slots cost more than the MUFU relief gains).
// TMEM columns: S0/P0 [0,128)  S1/P1 [128,256)  O0 [256,256+D)  O1 [384,384+D).
#include "common.cuh"

namespace g2 {

constexpr int ATT_BM = 128;       // query rows per tile
constexpr int ATT_BN = 128;       // keys per block
constexpr int ATT_THREADS = 384;  // 12 warps
constexpr float ATT_RESCALE_TAU = 8.0f;
constexpr int ATT_PCHUNKS = 2;    // P_t reaches the MMA warp in two pieces: keys [0,96) and [96,128) of a block
constexpr int ATT_PSPLIT = 96;    // (the six PV MMAs of the first piece run under the exp2 of the last 32 keys)
constexpr int ATT64_POLY = 0;     // head_dim 64: exp2 pairs (of every 8) on the FMA pipe instead of the MUFU — 0: measured slower (r02)

struct AttnKParams {
  CUtensorMap tmQ, tmK, tmV;
  const int* work;  // [n_items][8]: q_tile_begin, q_seg_begin, q_seg_end, k_begin, k_end, -, -, -
  __nv_bfloat16* out;
  long long ldo;
  int q_heads_per_kv;
  int causal;
  float scale_log2;
  int n_items, n_heads;
  int out_head_cols;  // columns of each head written to `out` (= head stride in `out`); D unless compacted
  float* lse;         // optional [rows][n_heads]: ln(sum exp(scaled scores)) per row, for merging partial results
};

template <int D>
struct AttnCfg {
  static constexpr int kBoxes = D / 64;                  // 64-column TMA boxes per row tile
  static constexpr int kTileBytes = ATT_BM * D * 2;      // one 128 x D bf16 tile
  static constexpr int kStages = (D == 128) ? 2 : 4;     // K/V ring depth
  static constexpr int kSmem = 2 * kTileBytes + kStages * 2 * kTileBytes + 1024 + 512;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32 pairs (one FFMA2/FADD2 per two values on sm_100)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

#ifdef G2_ATTN_TRACE
// debug build only (tools/attn_trace.py): clock64 stamps of CTA 0, [role 0..2][block 0..31][tile 0..1][event 0..15]
__device__ long long g_attn_trace[3 * 32 * 2 * 16];
#define G2_TR(role, blk, t, ev) \
  do { if (blockIdx.x == 0 && (blk) < 32) g_attn_trace[(((role) * 32 + (blk)) * 2 + (t)) * 16 + (ev)] = clock64(); } while (0)
#else
#define G2_TR(role, blk, t, ev) do { } while (0)
#endif

// One (work item, head) unit of a persistent CTA, decoded identically by every role.
struct AttnUnit {
  int q_tile_begin, q_seg_begin, q_seg_end, k_begin, len_q, len_k, head, kv_head, nblk;
  bool causal;
};

// P = 2^(s*scale - m) of one 128-key block: packed to bf16 pairs, written to the P columns of TMEM 32 keys per store and handed
// to the MMA warp in two pieces (ATT_PSPLIT keys, then the rest) so PV starts while the last keys are still in exp2.
// POLY (0..8): of every 8 score pairs, POLY take exp2 on the FMA pipe (round-to-nearest split x = n + f, degree-3 minimax
// polynomial for 2^f on [-0.5, 0.5], max rel error 7.5e-5 — far below the bf16 rounding of P — and n added to the
// exponent field) instead of MUFU.EX2.  At head_dim 64 a 128 x 128 tile needs 1024 MUFU clocks (16 exp2 / clk / SM,
// measured: tools/micro/mufu_bench.cu) against 512 tensor clocks, so the kernel is MUFU-bound and the FMA pipe is idle.
template <int POLY>
__device__ __forceinline__ void softmax_block(const uint32_t (&sv)[128], float m_used, float scale_log2, uint64_t& sum2,
                                              uint32_t tP_w, uint64_t* p_full_t, int lane, [[maybe_unused]] bool tr_on = false,
                                              [[maybe_unused]] int tr_t = 0, [[maybe_unused]] uint32_t tr_g = 0) {
  uint64_t neg_m2 = pack2(-m_used, -m_used);
  const uint64_t scale2 = pack2(scale_log2, scale_log2);
  [[maybe_unused]] const uint64_t magic2 = pack2(12582912.f, 12582912.f), nmagic2 = pack2(-12582912.f, -12582912.f);
  [[maybe_unused]] const uint64_t none2 = pack2(-1.f, -1.f);
  [[maybe_unused]] const uint64_t c0 = pack2(0.99992806f, 0.99992806f), c1 = pack2(0.69326097f, 0.69326097f);
  [[maybe_unused]] const uint64_t c2 = pack2(0.24261113f, 0.24261113f), c3 = pack2(0.05517167f, 0.05517167f);
  constexpr int CW = 32;  // keys per TMEM store
#pragma unroll
  for (int c = 0; c < ATT_BN / CW; ++c) {
    if (c * CW == ATT_PSPLIT) {
      // First hand-over: keys [0, ATT_PSPLIT).  ptxas issues every exp2 it can reach in one burst and sinks TMEM stores,
      // waits and arrives behind it (profiles/r02_attention_timeline.txt: with a hand-over per 32 keys all four ended up
      // behind the whole burst, ~120 clocks of exposed store latency each), so the exp2 of the remaining keys are made to
      // depend on a value read AFTER the arrive: +0 or -0 added to -m, numerically the same number.
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full_t[0]);
      if (tr_on) G2_TR(tr_t, tr_g, tr_t, 5);
      const float z = mbar_test_wait(&p_full_t[0], 0) ? 0.f : -0.f;
      neg_m2 = pack2(-m_used + z, -m_used + z);
    }
    uint32_t pk[CW / 2];
#pragma unroll
    for (int i = 0; i < CW; i += 2) {
      const uint64_t x2 = fma2(pack2(__uint_as_float(sv[c * CW + i]), __uint_as_float(sv[c * CW + i + 1])), scale2, neg_m2);
      float x0, x1, p0, p1;
      unpack2(x2, x0, x1);
      if (((i >> 1) & 7) < POLY) {
        const uint64_t xc = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
        const uint64_t t2 = add2(xc, magic2);           // integer part in the low mantissa bits
        const uint64_t f2 = fma2(add2(t2, nmagic2), none2, xc);
        uint64_t q2 = fma2(c3, f2, c2);
        q2 = fma2(q2, f2, c1);
        q2 = fma2(q2, f2, c0);
        float q0, q1, t0, t1;
        unpack2(q2, q0, q1);
        unpack2(t2, t0, t1);
        p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
        p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
      } else {
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
      }
      sum2 = add2(sum2, pack2(p0, p1));
      pk[i >> 1] = pack_bf16x2(p0, p1);
    }
    tmem_st16(tP_w + c * (CW / 2), pk);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&p_full_t[1]);
}

// Units blockIdx.x, blockIdx.x + gridDim.x, ... of a persistent CTA: unit u = (work item u / n_heads, head u % n_heads).
__device__ __forceinline__ AttnUnit decode_unit(const AttnKParams& p, int u) {
  AttnUnit a;
  const int* w = p.work + (u / p.n_heads) * 8;
  a.q_tile_begin = w[0];
  a.q_seg_begin = w[1];
  a.q_seg_end = w[2];
  a.k_begin = w[3];
  a.len_k = w[4] - w[3];
  a.causal = p.causal != 0 || w[5] != 0;  // per launch, or per work item
  a.head = u % p.n_heads;
  a.kv_head = a.head / p.q_heads_per_kv;
  a.len_q = a.q_seg_end - a.q_seg_begin;
  const int rows_here = min(2 * ATT_BM, a.q_seg_end - a.q_tile_begin);
  int k_needed = a.len_k;  // keys needed by this unit (causal: bottom-right aligned, as flash-attn)
  if (a.causal) {
    const int last_row = a.q_tile_begin + rows_here - 1 - a.q_seg_begin;
    k_needed = max(0, min(a.len_k, last_row + (a.len_k - a.len_q) + 1));
  }
  a.nblk = (k_needed + ATT_BN - 1) / ATT_BN;
  return a;
}

// The TMA producer thread: Q tiles once per unit, K/V blocks of 128 keys through the mbarrier ring.
template <int D, int KS>
__device__ __forceinline__ void attn_tma_producer(const AttnKParams& p, uint8_t* sQ, uint8_t* sKV, uint64_t* q_full,
                                                  uint64_t* q_empty, uint64_t* k_full, uint64_t* v_full,
                                                  uint64_t* k_empty, uint64_t* v_empty) {
  constexpr int kTileBytes = ATT_BM * D * 2, kBoxes = D / 64, BOX_BYTES = ATT_BM * 128;
  const int total_units = p.n_items * p.n_heads;
  uint32_t g = 0, n = 0;
  for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
    const AttnUnit a = decode_unit(p, u);
    if (a.nblk == 0) continue;
    mbar_wait(q_empty, (n & 1) ^ 1);  // the previous unit's QK^T are done with the Q tiles
    mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int b = 0; b < kBoxes; ++b)
        tma_load_2d(sQ + t * kTileBytes + b * BOX_BYTES, &p.tmQ, q_full, a.head * D + b * 64, a.q_tile_begin + t * ATT_BM);
    for (int j = 0; j < a.nblk; ++j, ++g) {
      const int s = g % KS;
      const uint32_t par = ((g / KS) & 1) ^ 1;
      uint8_t* sk = sKV + s * 2 * kTileBytes;
      uint8_t* sv = sk + kTileBytes;
      mbar_wait(&k_empty[s], par);
      mbar_arrive_expect_tx(&k_full[s], kTileBytes);
#pragma unroll
      for (int b = 0; b < kBoxes; ++b)
        tma_load_2d(sk + b * BOX_BYTES, &p.tmK, &k_full[s], a.kv_head * D + b * 64, a.k_begin + j * ATT_BN);
      mbar_wait(&v_empty[s], par);
      mbar_arrive_expect_tx(&v_full[s], kTileBytes);
#pragma unroll
      for (int b = 0; b < kBoxes; ++b)
        tma_load_2d(sv + b * BOX_BYTES, &p.tmV, &v_full[s], a.kv_head * D + b * 64, a.k_begin + j * ATT_BN);
    }
    ++n;
  }
}

template <int D, int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int KSTEPS_QK = D / 16;       // MMAs per S tile (K = head dim)
  constexpr int KSTEPS_PV = ATT_BN / 16;  // MMAs per PV block (K = keys)
  constexpr int BOX_BYTES = ATT_BM * 128; // one 64-col box of 128 rows
  constexpr int KS = Cfg::kStages;
  constexpr bool kSepP = D == 64;  // P_t in its own TMEM columns + one MMA issuer per tile (see the TMEM map below)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // 2 tiles
  uint8_t* sKV = smem + 2 * Cfg::kTileBytes;       // kStages x (K tile, V tile)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + KS * 2 * Cfg::kTileBytes);
  uint64_t* q_full = bars;                         // [1]  Q tiles of the current unit landed
  uint64_t* q_empty = bars + 1;                    // [1]  every QK^T of the unit has read them
  uint64_t* k_full = bars + 2;                     // [kStages]
  uint64_t* v_full = k_full + KS;
  uint64_t* k_empty = v_full + KS;
  uint64_t* v_empty = k_empty + KS;
  uint64_t* s_full = v_empty + KS;                 // [2]
  uint64_t* p_full = s_full + 2;                   // [2][ATT_PCHUNKS]: P handed over in key chunks
  uint64_t* o_done = p_full + 2 * ATT_PCHUNKS;     // [2]  PV of a block finished
  uint64_t* o_free = o_done + 2;                   // [2]  the epilogue has read O_t out of TMEM
  uint64_t* s_free = o_free + 2;                   // [2]  (kSepP) the softmax warps hold S_t in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = p.n_items * p.n_heads;

  // Persistent CTA: units blockIdx.x, blockIdx.x + gridDim.x, ... (head index fastest, so CTAs that run
  // side by side work on the heads of the same few query tiles and share their K/V in L2, and the cheap items a
  // caller appends at the end of the work table — the fused prompt rows — land at the end of every CTA's list
  // instead of unbalancing the static round-robin).  TMEM, the
  // barriers and the K/V ring live across units: the producer prefetches the next unit's Q/K/V while the
  // softmax warps still drain this one, which hides the per-unit prologue/epilogue (5-7 us of ~25 us per unit
  // on the 11-block DINO / Pi3-decoder segments when every unit was its own CTA).
  auto decode = [&](int u) { return decode_unit(p, u); };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      mbar_init(q_empty, kSepP ? 2 : 1);   // every MMA issuer is done with the Q tiles
      for (int s = 0; s < KS; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&k_empty[s], kSepP ? 2 : 1);  // released by every MMA issuer
        mbar_init(&v_empty[s], kSepP ? 2 : 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(&s_full[t], 1);
        for (int c = 0; c < ATT_PCHUNKS; ++c) mbar_init(&p_full[t * ATT_PCHUNKS + c], 4);  // one arrive per warp
        mbar_init(&o_done[t], 1);
        mbar_init(&o_free[t], 4);
        mbar_init(&s_free[t], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns of tile t (plain arithmetic: no local arrays).
  //   D = 128: S_t at t*128 with P_t written over its first 64 columns, O_t at 256 + t*128 (all 512 columns used)
  //   D =  64: O_t needs only 64 columns, which leaves room for P_t OUTSIDE S_t: S at t*128, P at 256 + t*64,
  //            O at 384 + t*64.  Then the next QK^T of a tile no longer has to wait for the PV that reads P: it is
  //            issued as soon as the softmax warps hold S_t in registers (s_free), S_{j+1} is ready before
  //            softmax_j ends, and the softmax warps never idle (the per-tile chain QK -> softmax -> PV -> QK that
  //            bounds the D = 128 case is broken).
  auto tS = [&](int t) { return tmem_base + static_cast<uint32_t>(t) * 128u; };
  auto tP = [&](int t) { return kSepP ? tmem_base + 256u + static_cast<uint32_t>(t) * 64u : tS(t); };
  auto tO = [&](int t) {
    return kSepP ? tmem_base + 384u + static_cast<uint32_t>(t) * 64u : tmem_base + 256u + static_cast<uint32_t>(t) * 128u;
  };

  // Every role walks the same unit sequence and keeps the same two running counters, from which all mbarrier
  // parities follow:  g = key blocks processed so far by this CTA,  n = units (with >= 1 block) so far.
  // Register budget (setmaxnreg must sit INSIDE each role branch so ptxas knows which budget governs
  // which code): 4 control warps x 32 x 80 + 8 softmax warps x 32 x 208 = 63488 <= 65536.
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // ------------------------------------ TMA producer ----------------------------------------
    if (elect_one()) attn_tma_producer<D, KS>(p, sQ, sKV, q_full, q_empty, k_full, v_full, k_empty, v_empty);
  } else if (!kSepP && warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // ------------------------------------ MMA issuer (D = 128: one thread issues both tiles) ----
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, D, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t kv_addr = smem_u32(sKV);

      auto issue_qk = [&](int t, int s) {
        const uint32_t qa = q_addr + t * Cfg::kTileBytes;
        const uint32_t ka = kv_addr + s * 2 * Cfg::kTileBytes;
#pragma unroll
        for (int k = 0; k < KSTEPS_QK; ++k) {
          const uint32_t off = (k >> 2) * BOX_BYTES + (k & 3) * 32;
          umma_ss(tS(t), umma_desc_kmajor(qa + off), umma_desc_kmajor(ka + off), idesc_qk, k != 0);
        }
        umma_commit(&s_full[t]);
      };
      // O_t += P_t V, issued piece by piece as the softmax warps hand P over
      auto issue_pv = [&](int t, int s, uint32_t par, bool first_block, [[maybe_unused]] uint32_t gblk) {
        const uint32_t va = kv_addr + s * 2 * Cfg::kTileBytes + Cfg::kTileBytes;
        G2_TR(2, gblk, t, 0);
        // opaque copy: otherwise ptxas precomputes the TMEM address of every chunk of both tiles outside the unit loop and
        // spills them (two local-memory loads per chunk on the issuer's critical path; profiles/r02_attention_rowsplit.txt)
        uint32_t tbase = tmem_base;
        asm volatile("" : "+r"(tbase));
        const uint32_t tS_t = tbase + static_cast<uint32_t>(t) * 128u, tO_t = tbase + 256u + static_cast<uint32_t>(t) * 128u;
#pragma unroll
        for (int c = 0; c < ATT_PCHUNKS; ++c) {
          mbar_wait(&p_full[t * ATT_PCHUNKS + c], par);
          tc_fence_after();
          if (c == 0) G2_TR(2, gblk, t, 1);
          if (c == ATT_PCHUNKS - 1) G2_TR(2, gblk, t, 2);
#pragma unroll
          for (int k = (c == 0 ? 0 : ATT_PSPLIT / 16); k < (c == 0 ? ATT_PSPLIT / 16 : KSTEPS_PV); ++k) {
            // 16 keys = 16 rows of 128 B inside each 64-column box; boxes are BOX_BYTES apart (LBO)
            umma_ts(tO_t, tS_t + k * 8, umma_desc_mnmajor(va + k * 2048, BOX_BYTES), idesc_pv,
                    !(first_block && k == 0));
          }
        }
        umma_commit(&o_done[t]);
        G2_TR(2, gblk, t, 3);
      };

      uint32_t g = 0, n = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const AttnUnit a = decode(u);
        if (a.nblk == 0) continue;
        mbar_wait(q_full, n & 1);
        mbar_wait(&k_full[g % KS], (g / KS) & 1);
        tc_fence_after();
        // S_t is free: the previous unit's last PV_t (which read P_t out of the same columns) is ahead in the pipe
        issue_qk(0, g % KS);
        issue_qk(1, g % KS);
        umma_commit(&k_empty[g % KS]);
        if (a.nblk == 1) umma_commit(q_empty);
        for (int j = 0; j < a.nblk; ++j, ++g) {
          const int s = g % KS;
          const uint32_t par = (g / KS) & 1;
          const int s1 = (g + 1) % KS;
          const uint32_t par1 = ((g + 1) / KS) & 1;
          const bool more = j + 1 < a.nblk;
          mbar_wait(&v_full[s], par);
          if (j == 0 && n > 0) {  // the first PV overwrites O_t: the previous unit's epilogue must have read it
            mbar_wait(&o_free[0], (n - 1) & 1);
            tc_fence_after();
          }
          issue_pv(0, s, g & 1, j == 0, g);
          if (more) {
            mbar_wait(&k_full[s1], par1);
            tc_fence_after();
            issue_qk(0, s1);
            G2_TR(2, g, 0, 4);
          }
          if (j == 0 && n > 0) {
            mbar_wait(&o_free[1], (n - 1) & 1);
            tc_fence_after();
          }
          issue_pv(1, s, g & 1, j == 0, g);
          umma_commit(&v_empty[s]);
          if (more) {
            issue_qk(1, s1);
            G2_TR(2, g, 1, 4);
            umma_commit(&k_empty[s1]);
            if (j + 2 == a.nblk) umma_commit(q_empty);  // that was the unit's last QK^T
          }
        }
        ++n;
      }
    }
  } else if (kSepP && (warp == 1 || warp == 2)) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // ------------------------------------ MMA issuers (D = 64): warp 1 -> query tile 0, warp 2 -> tile 1 ------
    // One issuer per tile: each walks its own tile's sequence in order and is never held up by a barrier of the
    // other tile; the tensor pipe interleaves the two instruction streams.  (Measured: -6.5 % on the DINO shape
    // together with the separate P columns; for D = 128, where P must alias S, two issuers cost +5 %, so that case
    // keeps the single issuer above.)
    if (elect_one()) {
      const int t = warp - 1;
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, D, 0, 1);  // B (= V) is MN-major
      const uint32_t qa = smem_u32(sQ) + t * Cfg::kTileBytes;
      const uint32_t kv_addr = smem_u32(sKV);
      const uint32_t tS_t = tS(t), tP_t = tP(t), tO_t = tO(t);

      auto issue_qk = [&](int s) {
        const uint32_t ka = kv_addr + s * 2 * Cfg::kTileBytes;
#pragma unroll
        for (int k = 0; k < KSTEPS_QK; ++k) {
          const uint32_t off = (k >> 2) * BOX_BYTES + (k & 3) * 32;
          umma_ss(tS_t, umma_desc_kmajor(qa + off), umma_desc_kmajor(ka + off), idesc_qk, k != 0);
        }
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[s]);
      };
      // O_t += P_t V, issued piece by piece as the softmax warps hand P over
      auto issue_pv = [&](int s, uint32_t par, bool first_block) {
        const uint32_t va = kv_addr + s * 2 * Cfg::kTileBytes + Cfg::kTileBytes;
#pragma unroll
        for (int c = 0; c < ATT_PCHUNKS; ++c) {
          mbar_wait(&p_full[t * ATT_PCHUNKS + c], par);
          tc_fence_after();
#pragma unroll
          for (int k = (c == 0 ? 0 : ATT_PSPLIT / 16); k < (c == 0 ? ATT_PSPLIT / 16 : KSTEPS_PV); ++k) {
            // 16 keys = 16 rows of 128 B inside each 64-column box; boxes are BOX_BYTES apart (LBO)
            umma_ts(tO_t, tP_t + k * 8, umma_desc_mnmajor(va + k * 2048, BOX_BYTES), idesc_pv,
                    !(first_block && k == 0));
          }
        }
        umma_commit(&o_done[t]);
        umma_commit(&v_empty[s]);
      };

      uint32_t g = 0, n = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const AttnUnit a = decode(u);
        if (a.nblk == 0) continue;
        mbar_wait(q_full, n & 1);
        mbar_wait(&k_full[g % KS], (g / KS) & 1);
        tc_fence_after();
        // S_t is free: the previous unit's last PV_t (which read P_t) is ahead in this thread's instruction stream
        // (aliased P) / the softmax warps released it through s_free (separate P)
        issue_qk(g % KS);
        if (a.nblk == 1) umma_commit(q_empty);
        for (int j = 0; j < a.nblk; ++j, ++g) {
          const int s = g % KS;
          const uint32_t par = (g / KS) & 1;
          const int s1 = (g + 1) % KS;
          const uint32_t par1 = ((g + 1) / KS) & 1;
          const bool more = j + 1 < a.nblk;
          if constexpr (kSepP) {
            // S_t is free once the softmax warps have loaded it: the next QK^T goes ahead of this block's PV
            mbar_wait(&s_free[t], g & 1);
            if (more) {
              mbar_wait(&k_full[s1], par1);
              tc_fence_after();
              issue_qk(s1);
              if (j + 2 == a.nblk) umma_commit(q_empty);  // that was the unit's last QK^T
            }
          }
          mbar_wait(&v_full[s], par);
          if (j == 0 && n > 0) {  // the first PV overwrites O_t: the previous unit's epilogue must have read it
            mbar_wait(&o_free[t], (n - 1) & 1);
            tc_fence_after();
          }
          issue_pv(s, g & 1, j == 0);
          if (!kSepP && more) {
            mbar_wait(&k_full[s1], par1);
            tc_fence_after();
            issue_qk(s1);
            if (j + 2 == a.nblk) umma_commit(q_empty);  // that was the unit's last QK^T
          }
        }
        ++n;
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ------------------------------------ softmax / correction / epilogue ---------------------
    const int t = (warp - 4) >> 2;              // query tile 0 or 1
    const int sub = warp & 3;                   // TMEM sub-partition
    const int r_in_tile = sub * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t tS_w = tS(t) + lane_off;
    const uint32_t tP_w = tP(t) + lane_off;
    const uint32_t tO_w = tO(t) + lane_off;
    uint32_t g = 0, n = 0;
    [[maybe_unused]] const bool tr_on = sub == 0 && lane == 0;

    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const AttnUnit a = decode(u);
      const int row = a.q_tile_begin + t * ATT_BM + r_in_tile;  // global query row
      const bool row_ok = row < a.q_seg_end;
      if (a.nblk == 0) {
        if (row_ok) {  // no visible keys at all: flash-attn writes zeros
          uint4* dst = reinterpret_cast<uint4*>(p.out + (long long)row * p.ldo + a.head * p.out_head_cols);
#pragma unroll
          for (int q = 0; q < D / 8; ++q)
            if (q * 8 < p.out_head_cols) dst[q] = make_uint4(0, 0, 0, 0);
          if (p.lse != nullptr) p.lse[(long long)row * p.n_heads + a.head] = -INFINITY;
        }
        continue;
      }
      // number of keys this row may see
      int limit = a.len_k;
      if (a.causal) limit = max(0, min(a.len_k, (row - a.q_seg_begin) + (a.len_k - a.len_q) + 1));

      float m_used = 0.f;  // running max in the scaled log2 domain (possibly stale by <= TAU)
      float l_run = 0.f;

      for (int j = 0; j < a.nblk; ++j, ++g) {
        if (tr_on) G2_TR(t, g, t, 0);
        mbar_wait(&s_full[t], g & 1);
        tc_fence_after();
        if (tr_on) G2_TR(t, g, t, 1);
        // all four 32-column loads in flight at once: taking the row maximum of one quarter while the next streams in
        // (a wait::ld per quarter) measured 3-9 % slower on every shape (profiles/r02_attention_rowsplit.txt)
        uint32_t sv[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS_w + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        tmem_wait_ld();
        if (tr_on) G2_TR(t, g, t, 2);
        if constexpr (kSepP) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);
        }

        const int col_base = j * ATT_BN;
        if (col_base + ATT_BN > limit) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (col_base + i >= limit) sv[i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sv[i]));
          mx1 = fmaxf(mx1, __uint_as_float(sv[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sv[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sv[i + 3]));
        }
        const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;  // may be -inf

        if (j == 0) {
          m_used = (m_blk == -INFINITY) ? 0.f : m_blk;
        } else {
          const bool need = m_blk > m_used + ATT_RESCALE_TAU;
          if (__any_sync(0xffffffffu, need)) {
            // rare path: O_t must be multiplied by 2^(m_used - m_new) before the next PV accumulates
            mbar_wait(&o_done[t], (g - 1) & 1);
            tc_fence_after();
            const float m_new = need ? m_blk : m_used;
            const float alpha = ex2_approx(m_used - m_new);
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
              uint32_t ov[32];
              tmem_ld32(tO_w + c * 32, ov);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st32(tO_w + c * 32, ov);
            }
            tmem_wait_st();
            l_run *= alpha;
            m_used = m_new;
          }
        }

        // P = 2^(s*scale - m), packed to bf16 pairs, written over S (columns [0,64) of the S region) and
        // handed to the MMA warp 32 keys at a time so PV starts while the rest of the row is still in exp
        if constexpr (kSepP) {
          if (g > 0) {  // P_t has its own columns: the previous block's PV must be done reading them
            mbar_wait(&o_done[t], (g - 1) & 1);
            tc_fence_after();
          }
        }
        if (tr_on) G2_TR(t, g, t, 3);
        uint64_t sum2 = pack2(0.f, 0.f);
        // fully unmasked block of a non-causal item: part of the exp2 may run on the FMA pipe (never for a block
        // that holds a masked -inf score: the polynomial would turn it into 2^-125 instead of an exact 0)
        const bool poly_ok = POLY > 0 && !a.causal && col_base + ATT_BN <= a.len_k;
        if (poly_ok) softmax_block<POLY>(sv, m_used, p.scale_log2, sum2, tP_w, &p_full[t * ATT_PCHUNKS], lane);
        else softmax_block<0>(sv, m_used, p.scale_log2, sum2, tP_w, &p_full[t * ATT_PCHUNKS], lane, tr_on, t, g);
        float sum0, sum1;
        unpack2(sum2, sum0, sum1);
        l_run += sum0 + sum1;
        if (tr_on) G2_TR(t, g, t, 4);
      }

      // epilogue: O_t leaves TMEM in one go (so the next unit's first PV may overwrite it), then
      // O / l -> bf16 -> global (each thread owns one output row of D contiguous values)
      mbar_wait(&o_done[t], (g - 1) & 1);
      tc_fence_after();
      uint32_t ov[D];
#pragma unroll
      for (int c = 0; c < D / 32; ++c) tmem_ld32(tO_w + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&ov[c * 32]));
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      if (row_ok) {
        // O and l are sums of 2^(s - m_used) terms, so m_used (however stale) + log2(l) is the exact log-sum-exp
        if (p.lse != nullptr)
          p.lse[(long long)row * p.n_heads + a.head] =
              l_run > 0.f ? (m_used + log2f(l_run)) * 0.69314718055994531f : -INFINITY;
        const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
        uint4* dst = reinterpret_cast<uint4*>(p.out + (long long)row * p.ldo + a.head * p.out_head_cols);
#pragma unroll
        for (int q = 0; q < D / 8; ++q) {
          if (q * 8 < p.out_head_cols) dst[q] = make_uint4(
              pack_bf16x2(__uint_as_float(ov[8 * q]) * inv_l, __uint_as_float(ov[8 * q + 1]) * inv_l),
              pack_bf16x2(__uint_as_float(ov[8 * q + 2]) * inv_l, __uint_as_float(ov[8 * q + 3]) * inv_l),
              pack_bf16x2(__uint_as_float(ov[8 * q + 4]) * inv_l, __uint_as_float(ov[8 * q + 5]) * inv_l),
              pack_bf16x2(__uint_as_float(ov[8 * q + 6]) * inv_l, __uint_as_float(ov[8 * q + 7]) * inv_l));
        }
      }
      ++n;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// out = softmax-weighted combination of two partial attention results over disjoint key sets.
// One thread per 8 output columns (16-byte loads/stores); memory-bound: 3 x rows x heads x cols x 2 B.
__global__ void attention_merge_kernel(const __nv_bfloat16* __restrict__ oa, long long lda, const float* __restrict__ lse_a,
                                       const __nv_bfloat16* __restrict__ ob, long long ldb, const float* __restrict__ lse_b,
                                       __nv_bfloat16* __restrict__ out, long long ldo, long long rows, int heads,
                                       int head_cols) {
  const int vec_per_head = head_cols / 8;
  const long long total = rows * heads * vec_per_head;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_head);
    const long long rh = i / vec_per_head;
    const int h = (int)(rh % heads);
    const long long r = rh / heads;
    const float la = lse_a[r * heads + h], lb = lse_b[r * heads + h];
    const float m = fmaxf(la, lb);
    float wa = 0.f, wb = 0.f;
    if (m > -INFINITY) {
      const float ea = __expf(la - m), eb = __expf(lb - m);
      const float inv = 1.0f / (ea + eb);
      wa = ea * inv;
      wb = eb * inv;
    }
    const long long col = (long long)h * head_cols + v * 8;
    const uint4 a4 = *reinterpret_cast<const uint4*>(oa + r * lda + col);
    const uint4 b4 = *reinterpret_cast<const uint4*>(ob + r * ldb + col);
    const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = __uint_as_float(aw[j] << 16), a1 = __uint_as_float(aw[j] & 0xffff0000u);
      const float b0 = __uint_as_float(bw[j] << 16), b1 = __uint_as_float(bw[j] & 0xffff0000u);
      // a weight of exactly 0 must not turn an uninitialised / inf partial into NaN
      const float o0 = (wa > 0.f ? wa * a0 : 0.f) + (wb > 0.f ? wb * b0 : 0.f);
      const float o1 = (wa > 0.f ? wa * a1 : 0.f) + (wb > 0.f ? wb * b1 : 0.f);
      ow[j] = pack_bf16x2(o0, o1);
    }
    *reinterpret_cast<uint4*>(out + r * ldo + col) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

template <int D, int POLY>
static int launch_attention(const AttnKParams& kp, int max_ctas, cudaStream_t stream) {
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_tcgen05_kernel<D, POLY>), AttnCfg<D>::kSmem)) return rc;
  const long long units = (long long)kp.n_items * kp.n_heads;
  long long cap = num_sms();                                                   // one persistent CTA per SM
  if (max_ctas > 0 && max_ctas < cap) cap = max_ctas;
  const unsigned grid = (unsigned)(units < cap ? units : cap);
  attention_tcgen05_kernel<D, POLY><<<grid, ATT_THREADS, AttnCfg<D>::kSmem, stream>>>(kp);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

}  // namespace g2

extern "C" int g2vlm_attention(const g2vlm_attn_args* a, void* stream_) {
  using namespace g2;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  G2_REQUIRE(a != nullptr, "attention: null args");
  G2_REQUIRE(a->q && a->k && a->v && a->out, "attention: null tensor");
  G2_REQUIRE(a->head_dim == 64 || a->head_dim == 128, "attention: head_dim must be 64 or 128 (pad 96 to 128)");
  G2_REQUIRE(a->num_q_heads > 0 && a->num_kv_heads > 0 && a->num_q_heads % a->num_kv_heads == 0,
             "attention: num_q_heads must be a multiple of num_kv_heads");
  G2_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0,
             "attention: leading dimensions must be multiples of 8");
  G2_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "attention: out must be 16-byte aligned");
  G2_REQUIRE(a->n_items >= 0, "attention: negative n_items");
  if (a->n_items == 0) return G2VLM_OK;
  G2_REQUIRE(a->work_items != nullptr, "attention: null work table");
  G2_REQUIRE(a->q_rows > 0 && a->kv_rows > 0, "attention: empty q/kv");

  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  const int D = a->head_dim;
  int rc = make_tmap_2d_bf16(&kp.tmQ, a->q, (uint64_t)a->q_rows, (uint64_t)a->num_q_heads * D,
                             (uint64_t)a->ldq * 2, ATT_BM, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&kp.tmK, a->k, (uint64_t)a->kv_rows, (uint64_t)a->num_kv_heads * D,
                         (uint64_t)a->ldk * 2, ATT_BN, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&kp.tmV, a->v, (uint64_t)a->kv_rows, (uint64_t)a->num_kv_heads * D,
                         (uint64_t)a->ldv * 2, ATT_BN, 64);
  if (rc) return rc;
  kp.work = a->work_items;
  kp.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  kp.ldo = a->ldo;
  kp.q_heads_per_kv = a->num_q_heads / a->num_kv_heads;
  kp.causal = a->causal;
  kp.scale_log2 = a->softmax_scale * 1.4426950408889634f;
  kp.n_items = a->n_items;
  kp.n_heads = a->num_q_heads;
  G2_REQUIRE(a->out_head_cols >= 0 && a->out_head_cols <= D && a->out_head_cols % 8 == 0,
             "attention: out_head_cols must be 0 or a multiple of 8 not above head_dim");
  kp.out_head_cols = a->out_head_cols ? a->out_head_cols : D;
  kp.lse = a->lse_out;
  G2_REQUIRE(a->max_ctas >= 0, "attention: negative max_ctas");
  if (D == 128) return launch_attention<128, 0>(kp, a->max_ctas, stream);
  // head_dim 64 needs 1024 MUFU clocks per 128 x 128 tile against 512 tensor clocks, so part of the exp2 was moved to
  // the FMA pipe (softmax_block<POLY>).  Measured on the DINO shape (profiles/r02_attention64_poly.txt): 224.6 us with
  // 0/8, 247.7 / 239.3 / 247.6 / 312.6 us with 1..4 of 8 pairs offloaded — the softmax warps are bound by their own
  // instruction stream and the per-tile QK -> softmax -> PV chain, not by MUFU throughput.  The default stays 0;
  // G2VLM_ATTN_POLY=2 reproduces the measurement (tools/attn64_poly_ab.py).
  int poly = ATT64_POLY;
  if (const char* e = getenv("G2VLM_ATTN_POLY")) poly = atoi(e);
  if (poly == 2) return launch_attention<64, 2>(kp, a->max_ctas, stream);
  return launch_attention<64, 0>(kp, a->max_ctas, stream);
}

#ifdef G2_ATTN_TRACE
extern "C" int g2vlm_debug_attn_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g2::g_attn_trace, sizeof(g2::g_attn_trace)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int g2vlm_attention_merge(const void* o_a, int64_t lda, const float* lse_a, const void* o_b, int64_t ldb,
                                     const float* lse_b, void* out, int64_t ldo, int64_t rows, int32_t heads,
                                     int32_t head_cols, void* stream_) {
  using namespace g2;
  G2_REQUIRE(o_a && o_b && lse_a && lse_b && out, "attention_merge: null tensor");
  G2_REQUIRE(rows >= 0 && heads > 0 && head_cols > 0 && head_cols % 8 == 0, "attention_merge: head_cols must be a multiple of 8");
  G2_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "attention_merge: leading dimensions must be multiples of 8");
  G2_REQUIRE(((reinterpret_cast<uintptr_t>(o_a) | reinterpret_cast<uintptr_t>(o_b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "attention_merge: tensors must be 16-byte aligned");
  if (rows == 0) return G2VLM_OK;
  const long long total = rows * heads * (head_cols / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  attention_merge_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const __nv_bfloat16*>(o_a), lda, lse_a, reinterpret_cast<const __nv_bfloat16*>(o_b), ldb, lse_b,
      reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, heads, head_cols);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}
