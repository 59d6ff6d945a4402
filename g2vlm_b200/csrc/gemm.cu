// g2vlm_b200 — grouped (token-type-routed) bf16 GEMM on tcgen05 tensor cores.
//
// Two persistent, warp-specialised kernels per epilogue: gemm_bf16_tcgen05_pair_kernel (CTA pairs, cta_group::2,
// 256x256 tiles; used whenever there are enough tiles, see its comment block below) and the 1-CTA kernel:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma 128x256x16, fp32 accumulators in TMEM, 2 accumulator
//                               stages so the epilogue of tile i overlaps the main loop of i+1)
//   warps 2..9  epilogue       (tcgen05.ld -> registers -> fused bias / GELU / SwiGLU / LayerScale+residual
//                               -> swizzled smem transpose -> coalesced 16-byte global stores / RMW)
// Routing: tokens are permuted once per forward so each expert's rows are contiguous ("gather by
// expert"); an M tile therefore belongs to exactly one expert and selects that expert's weight rows
// by a row offset into the stacked weight matrix — each expert runs as one dense GEMM and both run
// in the same launch. See include/g2vlm_b200.h for the reference call sites this replaces.
#include "common.cuh"

namespace g2 {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int STAGING_BYTES = EPI_WARPS * 4096;  // per-warp 32x32 fp32 transpose buffer
constexpr int PANEL_M = 16;  // rasterisation: 16 M tiles x all N tiles per panel (L2 reuse)
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmKParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int n_groups;
  int grp_row0[2];
  int grp_rows[2];
  int grp_mtile0[3];
  int N, K;
  int n_tiles_n;
  int num_tiles;
  int num_kb;
  uint32_t flags;
  uint32_t scale_groups;
  void* out;
  long long ldo;
  const float* bias;
  const float* scale;
  const float* residual;
  long long ldr;
  int out_col_group, out_col_stride;  // STORE_BF16 column regrouping (0 = plain)
  // CTA-pair kernel: B tensor map with a 128-row box (each CTA of the pair loads half of the 256 N rows) and the
  // tile list in units of M-tile PAIRS (every group padded to an even number of M tiles)
  CUtensorMap tmBh;
  int grp_mpair0[3];
  int num_pair_tiles;
  int pair_panel;   // rasterisation: pair_panel M pairs x all N tiles per panel
  CUtensorMap tmOut;   // RESID_F32: fp32 output as a TMA target (32 x 32 boxes, 128-byte swizzle) when use_tma_out
  int use_tma_out;
  int k_chunk_kb;   // > 0 (fp32 mode, STORE_F32): k-blocks per accumulator chunk, chunks summed in fp32 RN by the epilogue
};

struct TileCoord {
  int g, row0, rows_valid, n_tile;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int tile) {
  const int mt_total = p.grp_mtile0[p.n_groups];
  const int per_panel = PANEL_M * p.n_tiles_n;
  const int panel = tile / per_panel;
  const int r = tile - panel * per_panel;
  const int panel_h = min(PANEL_M, mt_total - panel * PANEL_M);
  const int n_tile = r / panel_h;
  const int m = panel * PANEL_M + (r - n_tile * panel_h);
  TileCoord t;
  t.g = (p.n_groups > 1 && m >= p.grp_mtile0[1]) ? 1 : 0;
  const int m_local = m - p.grp_mtile0[t.g];
  t.row0 = p.grp_row0[t.g] + m_local * BM;
  t.rows_valid = min(BM, p.grp_rows[t.g] - m_local * BM);
  t.n_tile = n_tile;
  return t;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// ---- epilogues -------------------------------------------------------------------------------------
// 8 epilogue warps: warp w reads TMEM sub-partition w%4 (32 accumulator rows) and half of the tile's
// columns. tcgen05.ld hands each THREAD one row x 32 columns; writing that straight to global memory
// makes every warp instruction touch 32 different rows (32 L1 wavefronts). Each 32x32 chunk is therefore
// transposed through a per-warp, XOR-swizzled shared-memory staging buffer so that a warp instruction
// covers whole 128-byte row segments of 4 (fp32) / 8 (bf16) rows: coalesced stores and RMW.
__device__ __forceinline__ void stage_store_bf16(uint8_t* stg, int lane, const uint32_t (&pk)[16],
                                                 __nv_bfloat16* out_tile, long long ldo, int rows_ok,
                                                 int cols_ok) {
  // write: row = lane, four 16-byte segments; segment q lives at q ^ ((row >> 1) & 3)  (64-byte rows)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int qs = q ^ ((lane >> 1) & 3);
    *reinterpret_cast<uint4*>(stg + lane * 64 + qs * 16) =
        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
  __syncwarp();
  const int seg = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2);
    const int qs = seg ^ ((row >> 1) & 3);
    const uint4 v = *reinterpret_cast<const uint4*>(stg + row * 64 + qs * 16);
    if (row < rows_ok) {
      __nv_bfloat16* dst = out_tile + row * ldo + seg * 8;
      if (seg * 8 + 8 <= cols_ok) {
        *reinterpret_cast<uint4*>(dst) = v;
      } else {
        const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v);
        for (int j = 0; j < 8; ++j)
          if (seg * 8 + j < cols_ok) dst[j] = e[j];
      }
    }
  }
  __syncwarp();
}

// fp32 variant: thread `lane` writes its row (32 floats) swizzled; afterwards piece i of a lane is
// row i*4 + lane/8, columns (lane%8)*4 .. +3 — 8 lanes cover 128 contiguous bytes of one row.
__device__ __forceinline__ void stage_write_f32(uint8_t* stg, int lane, const float (&f)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int qs = q ^ (lane & 7);
    *reinterpret_cast<float4*>(stg + lane * 128 + qs * 16) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
  }
  __syncwarp();
}
__device__ __forceinline__ float4 stage_read_f32(const uint8_t* stg, int lane, int i) {
  const int row = i * 4 + (lane >> 3);
  return *reinterpret_cast<const float4*>(stg + row * 128 + ((lane & 7) ^ (row & 7)) * 16);
}

// chunk_first / chunk_last (STORE_F32 with k_chunk_kb > 0, fp32 mode): the accumulator of ONE K chunk — bias only with
// the first chunk, `out +=` for every later one (fp32 round-to-nearest on the CUDA cores), activation / scale / residual
// only with the last.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, const TileCoord& t, uint32_t taddr,
                                              int sub, int half, int lane, uint8_t* stg, bool chunk_first = true,
                                              bool chunk_last = true) {
  const int rows_ok = t.rows_valid - sub * 32;  // rows of this warp's 32-row slab that are real
  const long long row_base = t.row0 + sub * 32;
  const int n0 = t.n_tile * BN;

  if constexpr (EPI == G2VLM_EPI_SWIGLU_BF16) {
    // columns [0,128) of the tile = gate, [128,256) = up, for output columns n_tile*128 + [0,128)
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + row_base * p.ldo + t.n_tile * 128;
#pragma unroll 1
    for (int c = half * 2; c < half * 2 + 2; ++c) {
      uint32_t vg[32], vu[32];
      tmem_ld32(taddr + c * 32, vg);
      tmem_ld32(taddr + 128 + c * 32, vu);
      tmem_wait_ld();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float o[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          // reference rounding points: gate_proj -> bf16, silu -> bf16, up_proj -> bf16, mul -> bf16
          const float g = bf16_round(__uint_as_float(vg[j + e]));
          const float u = bf16_round(__uint_as_float(vu[j + e]));
          const float sg = bf16_round(g / (1.0f + __expf(-g)));
          o[e] = sg * u;
        }
        pk[j >> 1] = pack_bf16x2(o[0], o[1]);
      }
      stage_store_bf16(stg, lane, pk, out + c * 32, p.ldo, rows_ok, 32);
    }
    return;
  } else {
    const float* bias = (p.bias && chunk_first) ? p.bias + (long long)t.g * p.N : nullptr;
    const bool use_scale = p.scale != nullptr && ((p.scale_groups >> t.g) & 1u);
#pragma unroll 1
    for (int c = half * 4; c < half * 4 + 4; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) break;  // warp-uniform
      const int cols_ok = min(32, p.N - col0);
      const int c4 = (lane & 7) * 4;  // this lane's 4 columns in the transposed (coalesced) domain
      [[maybe_unused]] float4 old[8];
      [[maybe_unused]] float* out_piece = nullptr;
      // RESID_F32 through the TMA: the chunk goes to the (128-byte-swizzled) staging tile and ONE thread issues a bulk
      // reduce-add of the 32 x 32 box — the fp32 add happens in the L2 reduction units (same single rounding as the
      // SM-side add), the SM never reads the residual stream and never waits for it.  Only whole 32 x 32 chunks of a
      // tile whose 32 rows all belong to this expert group; ragged slabs / column tails keep the read-modify-write.
      [[maybe_unused]] bool tma_chunk = false;
      if constexpr (EPI == G2VLM_EPI_RESID_F32) {
        tma_chunk = p.use_tma_out && rows_ok >= 32 && cols_ok == 32;
      }
      if constexpr (EPI == G2VLM_EPI_RESID_F32) {
        out_piece = reinterpret_cast<float*>(p.out) + (row_base + (lane >> 3)) * p.ldo + col0 + c4;
        if (!tma_chunk) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            old[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i * 4 + (lane >> 3) < rows_ok && c4 + 4 <= cols_ok)
              old[i] = *reinterpret_cast<const float4*>(out_piece + (long long)i * 4 * p.ldo);
          }
        }
        // Pull the NEXT chunk's residual lines (32 rows x 128 B) into L2 now, without holding registers for them:
        // their loads one chunk later then cost an L2 hit instead of an HBM round trip (same-box A/B: DINO dense
        // 69 -> 63.5 us, the other residual GEMMs -1 %; prefetching a whole tile ahead instead was 5-8 % SLOWER).
        if (!tma_chunk && c + 1 < half * 4 + 4 && col0 + 32 < p.N && (lane & 7) == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i * 4 + (lane >> 3) < rows_ok)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(out_piece + (long long)i * 4 * p.ldo + 32));
        }
      }
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_wait_ld();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (bias) {
        if (cols_ok == 32) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = __ldg(b4 + q);
            f[4 * q] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < cols_ok) f[j] += __ldg(bias + col0 + j);
        }
      }

      if constexpr (EPI == G2VLM_EPI_STORE_BF16) {
        if (p.flags & G2VLM_GEMM_GELU) {
          // exact-erf GELU of the bf16-rounded pre-activation (bit-identical to torch's nn.GELU on a bf16 tensor on
          // 100 % of 2 x 10^8 sampled outputs, tools/gemm_gelu_ab.py).  Measured and rejected (r02): tabulating the
          // 3072 bf16 inputs with 2^-9 <= |x| < 8 in shared memory instead — also bit-identical, but 249 us against
          // 184 us on DINO fc1 (divergent range checks + bank-conflicted 2-byte loads cost more than erff's ~25 FMAs).
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_erf(bf16_round(f[j]));
        } else if (p.flags & G2VLM_GEMM_QUICK_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float xq = bf16_round(f[j]);
            f[j] = xq / (1.0f + __expf(-1.702f * xq));
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
        // optional regrouping of the output columns (heads of `group` columns into slots of `stride`); both are
        // multiples of the 32-column chunk, so a chunk never straddles a slot
        const int ocol = p.out_col_group ? col0 + (col0 / p.out_col_group) * (p.out_col_stride - p.out_col_group) : col0;
        stage_store_bf16(stg, lane, pk, reinterpret_cast<__nv_bfloat16*>(p.out) + row_base * p.ldo + ocol, p.ldo,
                         rows_ok, cols_ok);
      } else if constexpr (EPI == G2VLM_EPI_RESID_F32) {
        // x += [bf16]( gamma * bf16(acc + bias) )   (reference: g2vlm/qwen2vl.py:885-887, 907-909)
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = bf16_round(f[j]);
        if (use_scale) {
          if (cols_ok == 32) {
            const float4* s4 = reinterpret_cast<const float4*>(p.scale + col0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 sc = __ldg(s4 + q);
              f[4 * q] *= sc.x; f[4 * q + 1] *= sc.y; f[4 * q + 2] *= sc.z; f[4 * q + 3] *= sc.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < cols_ok) f[j] *= __ldg(p.scale + col0 + j);
          }
          if (p.flags & G2VLM_GEMM_ROUND_AFTER_SCALE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = bf16_round(f[j]);
          }
        }
        if (tma_chunk) {
          if (lane == 0) tma_store_wait_read();   // the previous chunk's bulk read of this staging tile is done
          __syncwarp();
          stage_write_f32(stg, lane, f);          // ends with __syncwarp
          fence_proxy_async_smem();               // generic-proxy writes -> visible to the async proxy
          __syncwarp();
          if (lane == 0) tma_reduce_add_2d(&p.tmOut, stg, col0, (int)row_base);
          continue;
        }
        // RMW of the fp32 residual stream in the transposed (coalesced) domain. The 8 old values were
        // loaded BEFORE the accumulator chunk was read (`old`, below), so their latency is hidden.
        if (p.use_tma_out) {                      // a pending bulk read of the staging tile must finish first
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
        stage_write_f32(stg, lane, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = stage_read_f32(stg, lane, i);
          const int row = i * 4 + (lane >> 3);
          float* d = out_piece + (long long)i * 4 * p.ldo;
          if (row < rows_ok) {
            if (c4 + 4 <= cols_ok) {
              float4 x = old[i];
              x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
              if (p.flags & G2VLM_GEMM_ROUND_SUM) {
                x.x = bf16_round(x.x); x.y = bf16_round(x.y); x.z = bf16_round(x.z); x.w = bf16_round(x.w);
              }
              *reinterpret_cast<float4*>(d) = x;
            } else {
              const float av[4] = {a.x, a.y, a.z, a.w};
              for (int j = 0; j < 4; ++j)
                if (c4 + j < cols_ok) {
                  const float xs = d[j] + av[j];
                  d[j] = (p.flags & G2VLM_GEMM_ROUND_SUM) ? bf16_round(xs) : xs;
                }
            }
          }
        }
        __syncwarp();
      } else {  // G2VLM_EPI_STORE_F32
        if (p.flags & G2VLM_GEMM_ROUND_BF16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = bf16_round(f[j]);
        }
        // order: v = acc [+ bias] [+ out] ; [gelu] ; [* scale] ; [relu] ; [+ residual]   (the last four only with
        // the final K chunk)
        const bool accum = (p.flags & G2VLM_GEMM_ACCUMULATE) != 0 || !chunk_first;
        const bool relu = chunk_last && (p.flags & G2VLM_GEMM_RELU) != 0;
        const bool gelu = chunk_last && (p.flags & G2VLM_GEMM_GELU) != 0;   // fp32 mode: exact erf on the fp32 value
        const bool scl = chunk_last && use_scale;                           // fp32 mode: LayerScale without rounding
        float* outp = reinterpret_cast<float*>(p.out) + (row_base + (lane >> 3)) * p.ldo + col0 + c4;
        const float* resp = (p.residual && chunk_last) ? p.residual + (row_base + (lane >> 3)) * p.ldr + col0 + c4 : nullptr;
        stage_write_f32(stg, lane, f);
        float sc[4] = {1.f, 1.f, 1.f, 1.f};
        if (scl) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c4 + j < cols_ok) sc[j] = __ldg(p.scale + col0 + c4 + j);
        }
        float4 po[8], pr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // batch the loads (accumulate / residual operands)
          const bool ok = i * 4 + (lane >> 3) < rows_ok && c4 + 4 <= cols_ok;
          po[i] = (ok && accum) ? *reinterpret_cast<const float4*>(outp + (long long)i * 4 * p.ldo) : make_float4(0.f, 0.f, 0.f, 0.f);
          pr[i] = (ok && resp) ? *reinterpret_cast<const float4*>(resp + (long long)i * 4 * p.ldr) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = stage_read_f32(stg, lane, i);
          const int row = i * 4 + (lane >> 3);
          float* d = outp + (long long)i * 4 * p.ldo;
          if (row < rows_ok) {
            if (c4 + 4 <= cols_ok) {
              float xv[4] = {a.x + po[i].x, a.y + po[i].y, a.z + po[i].z, a.w + po[i].w};
              if (gelu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) xv[j] = gelu_erf(xv[j]);
              }
              if (scl) {
#pragma unroll
                for (int j = 0; j < 4; ++j) xv[j] *= sc[j];
              }
              if (relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) xv[j] = fmaxf(xv[j], 0.f);
              }
              *reinterpret_cast<float4*>(d) = make_float4(xv[0] + pr[i].x, xv[1] + pr[i].y, xv[2] + pr[i].z, xv[3] + pr[i].w);
            } else {
              const float av[4] = {a.x, a.y, a.z, a.w};
              for (int j = 0; j < 4; ++j) {
                if (c4 + j < cols_ok) {
                  float x = av[j];
                  if (accum) x += d[j];
                  if (gelu) x = gelu_erf(x);
                  if (scl) x *= sc[j];
                  if (relu) x = fmaxf(x, 0.f);
                  if (resp) x += resp[(long long)i * 4 * p.ldr + j];
                  d[j] = x;
                }
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms are 1024 B: align the operand ring on the SHARED address
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], EPI_WARPS * 32);
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

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int b_row = t.g * p.N + t.n_tile * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d(sa, &p.tmA, &full_bar[stage], kb * BK, t.row0);
          tma_load_2d(sa + A_BYTES, &p.tmB, &full_bar[stage], kb * BK, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // k_chunk_kb > 0 (fp32 mode): the accumulator leaves the tensor core every k_chunk_kb k-blocks and the chunks
      // are summed by the epilogue warps in fp32 round-to-nearest.  The tensor core adds into its fp32 accumulator
      // with truncation, so a K of tens of thousands (6 x 8960 in the split form) loses ~1e-5 relative when it is
      // accumulated in one go (measured: 3e-5 per MoT layer); 16 MMA steps per chunk keep that below 1e-6.
      const int chunk = p.k_chunk_kb > 0 ? p.k_chunk_kb : p.num_kb;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kb0 = 0; kb0 < p.num_kb; kb0 += chunk) {
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          const int kb1 = min(p.num_kb, kb0 + chunk);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
            const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_ss(d_tmem, umma_desc_kmajor(a_addr + k * 32), umma_desc_kmajor(b_addr + k * 32),
                      idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tmem_full_bar[acc]);  // accumulator (chunk) complete -> epilogue
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------- epilogue warps -------------------------------------------
    const int sub = warp & 3;          // TMEM sub-partition this warp may access: lanes [32*sub, 32*sub+32)
    const int half = (warp - 2) >> 2;  // which half of the tile's columns
    uint8_t* stg = staging + (warp - 2) * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int chunk = p.k_chunk_kb > 0 ? p.k_chunk_kb : p.num_kb;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      for (int kb0 = 0; kb0 < p.num_kb; kb0 += chunk) {
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BN;
        epilogue_tile<EPI>(p, t, taddr, sub, half, lane, stg, kb0 == 0, kb0 + chunk >= p.num_kb);
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_read();   // outstanding bulk reduce-adds still read this warp's staging tile
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cluster of 2, tcgen05 cta_group::2): the pair computes a 256 x 256 tile.  Each CTA loads its own
// 128 A rows and HALF of the 256 B rows per k block (32 KB instead of 48 KB per CTA and per 128x256x64 of MMA work:
// one third less TMA / L2 / shared-memory operand traffic per FLOP, and a 6-stage ring instead of 4), the leader CTA
// issues one M = 256 MMA per k step that reads both CTAs' operands, and each CTA keeps the accumulator of its own
// 128 rows in its own TMEM and runs the unchanged epilogue on it.
//   full[s]        lives in the leader, whose producer announces the 64 KB of both CTAs; both CTAs' TMA loads
//                  complete their bytes on it (cp.async.bulk.tensor .cta_group::2)
//   empty[s], tmem_full[a]   one per CTA, signalled in both CTAs at once by the leader's multicast tcgen05.commit
//   tmem_empty[a]  lives in the leader; the epilogue warps of both CTAs arrive on it (remote arrive from the peer)
// ------------------------------------------------------------------------------------------------------------------
constexpr int PAIR_STAGES = 6;
constexpr int PAIR_B_BYTES = (BN / 2) * BK * 2;
constexpr int PAIR_STAGE_BYTES = A_BYTES + PAIR_B_BYTES;   // 32 KB
constexpr int PAIR_SMEM = PAIR_STAGES * PAIR_STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ TileCoord decode_pair_tile(const GemmKParams& p, int tile, int rank) {
  const int mp_total = p.grp_mpair0[p.n_groups];
  const int per_panel = p.pair_panel * p.n_tiles_n;
  const int panel = tile / per_panel;
  const int r = tile - panel * per_panel;
  const int panel_h = min(p.pair_panel, mp_total - panel * p.pair_panel);
  const int n_tile = r / panel_h;
  const int mp = panel * p.pair_panel + (r - n_tile * panel_h);
  TileCoord t;
  t.g = (p.n_groups > 1 && mp >= p.grp_mpair0[1]) ? 1 : 0;
  const int m_local = 2 * (mp - p.grp_mpair0[t.g]) + rank;
  t.row0 = p.grp_row0[t.g] + m_local * BM;
  t.rows_valid = min(BM, p.grp_rows[t.g] - m_local * BM);  // <= 0: this CTA's half of the pair is padding
  t.n_tile = n_tile;
  return t;
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* staging = smem + PAIR_STAGES * PAIR_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + PAIR_STAGES;
  uint64_t* tmem_full_bar = empty_bar + PAIR_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmBh);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < PAIR_STAGES; ++s) {
        mbar_init(&full_bar[s], 1);    // the leader announces the bytes of BOTH CTAs (only the leader's copy is used)
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], 2 * EPI_WARPS);   // one arrive per epilogue warp of either CTA
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair<512>(tmem_slot);
  }
  tc_fence_before();
  cluster_sync_all();   // barriers initialised and TMEM allocated in BOTH CTAs before any remote arrive / MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) ---------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < p.num_pair_tiles; tile += n_pairs) {
        const TileCoord t = decode_pair_tile(p, tile, rank);
        const int b_row = t.g * p.N + t.n_tile * BN + rank * (BN / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
          // Only the leader arrives (with the byte count of both CTAs); the peer's bytes may complete on the barrier
          // before that arrive — the transaction count is signed — but never in an earlier phase: the peer reuses a
          // stage only after the leader's MMAs, which waited for that phase, have released it.  (A remote
          // arrive.expect_tx.release.cluster per k block from the peer cost ~1000 cycles each.)
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * PAIR_STAGE_BYTES);
          uint8_t* sa = smem + stage * PAIR_STAGE_BYTES;
          tma_load_2d_pair(sa, &p.tmA, full_leader, kb * BK, t.row0);
          tma_load_2d_pair(sa + A_BYTES, &p.tmBh, full_leader, kb * BK, b_row);
          if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) -----------------------------
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < p.num_pair_tiles; tile += n_pairs) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * PAIR_STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_ss_pair(d_tmem, umma_desc_kmajor(a_addr + k * 32), umma_desc_kmajor(b_addr + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs once these MMAs retire
          if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full_bar[acc]);  // accumulators complete -> both CTAs' epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue warps (both CTAs, own 128 rows) -----------------
    const int sub = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stg = staging + (warp - 2) * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < p.num_pair_tiles; tile += n_pairs) {
      const TileCoord t = decode_pair_tile(p, tile, rank);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * BN;
      if (t.rows_valid > 0) epilogue_tile<EPI>(p, t, taddr, sub, half, lane, stg);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_read();   // outstanding bulk reduce-adds still read this warp's staging tile
  }

  tc_fence_before();
  cluster_sync_all();   // both CTAs are done with TMEM and with each other's shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

template <int EPI>
static int launch_gemm_pair(const GemmKParams& kp, cudaStream_t stream) {
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gemm_bf16_tcgen05_pair_kernel<EPI>), PAIR_SMEM)) return rc;
  const int pairs = min(kp.num_pair_tiles, num_sms() / 2);
  gemm_bf16_tcgen05_pair_kernel<EPI><<<2 * pairs, GEMM_THREADS, PAIR_SMEM, stream>>>(kp);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

template <int EPI>
static int launch_gemm(const GemmKParams& kp, cudaStream_t stream) {
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gemm_bf16_tcgen05_kernel<EPI>), GEMM_SMEM)) return rc;
  const int grid = kp.num_tiles < num_sms() ? kp.num_tiles : num_sms();
  gemm_bf16_tcgen05_kernel<EPI><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(kp);
  G2_CUDA_OK(cudaGetLastError());
  return G2VLM_OK;
}

}  // namespace g2

extern "C" int g2vlm_gemm_bf16(const g2vlm_gemm_args* a, void* stream_) {
  using namespace g2;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  G2_REQUIRE(a != nullptr, "gemm: null args");
  G2_REQUIRE(a->A && a->B && a->out, "gemm: null tensor");
  G2_REQUIRE(a->n_groups == 1 || a->n_groups == 2, "gemm: n_groups must be 1 or 2");
  G2_REQUIRE(a->N > 0 && a->K > 0, "gemm: N and K must be positive");
  G2_REQUIRE(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0, "gemm: K, lda, ldb must be multiples of 8");
  G2_REQUIRE(a->epilogue >= 0 && a->epilogue <= 3, "gemm: unknown epilogue");
  if (a->epilogue == G2VLM_EPI_SWIGLU_BF16) {
    G2_REQUIRE(a->N % 256 == 0, "gemm: SwiGLU epilogue needs N (= 2*intermediate) % 256 == 0");
    G2_REQUIRE(a->ldo % 8 == 0, "gemm: ldo must be a multiple of 8");
  } else if (a->epilogue == G2VLM_EPI_STORE_BF16) {
    G2_REQUIRE(a->ldo % 8 == 0, "gemm: ldo must be a multiple of 8");
  } else {
    G2_REQUIRE(a->ldo % 4 == 0, "gemm: ldo must be a multiple of 4");
    G2_REQUIRE(a->residual == nullptr || a->ldr % 4 == 0, "gemm: ldr must be a multiple of 4");
  }
  G2_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "gemm: out must be 16-byte aligned");
  G2_REQUIRE(a->out_col_group == 0 ||
                 (a->epilogue == G2VLM_EPI_STORE_BF16 && a->out_col_group > 0 && a->out_col_group % 32 == 0 &&
                  a->out_col_stride % 32 == 0 && a->out_col_stride >= a->out_col_group && a->N % a->out_col_group == 0),
             "gemm: out_col_group/out_col_stride need STORE_BF16, multiples of 32, stride >= group, N % group == 0");
  G2_REQUIRE(a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "gemm: bias alignment");
  G2_REQUIRE(a->bias == nullptr || a->n_groups == 1 || a->N % 4 == 0, "gemm: stacked bias needs N % 4 == 0");
  G2_REQUIRE(a->scale == nullptr || (reinterpret_cast<uintptr_t>(a->scale) & 15) == 0, "gemm: scale alignment");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  int mt = 0;
  long long max_row = 0;
  for (int g = 0; g < a->n_groups; ++g) {
    G2_REQUIRE(a->group_rows[g] >= 0 && a->group_row0[g] >= 0, "gemm: negative group extent");
    kp.grp_row0[g] = a->group_row0[g];
    kp.grp_rows[g] = a->group_rows[g];
    kp.grp_mtile0[g] = mt;
    mt += cdiv(a->group_rows[g], BM);
    const long long end = (long long)a->group_row0[g] + a->group_rows[g];
    if (end > max_row) max_row = end;
  }
  kp.grp_mtile0[a->n_groups] = mt;
  if (a->n_groups == 1) kp.grp_mtile0[2] = mt;
  G2_REQUIRE(max_row <= a->a_rows, "gemm: group rows exceed a_rows");
  if (mt == 0) return G2VLM_OK;  // empty input: nothing to do
  {
    // decode-shaped calls (<= 8 rows, one non-empty group): HBM-bound -> GEMV kernel, not a 128-row MMA tile
    int total = 0, nonempty = 0, grp = 0;
    for (int g = 0; g < a->n_groups; ++g)
      if (a->group_rows[g] > 0) { total += a->group_rows[g]; ++nonempty; grp = g; }
    const bool aligned = a->out_col_group == 0 && (reinterpret_cast<uintptr_t>(a->A) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(a->B) & 15) == 0;
    if (nonempty == 1 && total <= 8 && aligned) return launch_gemv(a, grp, stream);
    // 9..32 rows (system prompt / short question prefill): still HBM-bound, and a 128-row tile would put the whole weight matrix
    // on N / 256 CTAs -> warp-level MMA kernel that streams the weights once over N / 16 CTAs (csrc/decode.cu)
    const int n_out = a->epilogue == G2VLM_EPI_SWIGLU_BF16 ? a->N / 2 : a->N;
    if (nonempty == 1 && total <= 32 && aligned && n_out % 16 == 0 && a->K % 32 == 0 && a->k_chunk_blocks == 0 &&
        a->epilogue != G2VLM_EPI_STORE_F32 &&   // (the fp32 heads keep the tile kernel's accumulation order)
        !(a->flags & (G2VLM_GEMM_FORCE_PAIR | G2VLM_GEMM_FORCE_SINGLE)))
      return launch_gemv(a, grp, stream);
  }
  kp.n_groups = a->n_groups;
  kp.N = a->N;
  kp.K = a->K;
  kp.n_tiles_n = cdiv(a->N, BN);
  kp.num_tiles = mt * kp.n_tiles_n;
  kp.num_kb = cdiv(a->K, BK);
  kp.flags = a->flags;
  kp.scale_groups = a->scale_groups;
  kp.out = a->out;
  kp.ldo = a->ldo;
  kp.bias = a->bias;
  kp.scale = a->scale;
  kp.residual = a->residual;
  kp.ldr = a->ldr;
  kp.out_col_group = a->out_col_group;
  kp.out_col_stride = a->out_col_stride;
  // RESID_F32: the residual stream is updated through TMA reduce-add (fp32 add in L2) unless the caller needs the
  // rounded sum (bf16 residual stream of the ViT / training forward) or opts out for A/B timing
  kp.use_tma_out = 0;
  if (a->epilogue == G2VLM_EPI_RESID_F32 && !(a->flags & (G2VLM_GEMM_ROUND_SUM | G2VLM_GEMM_NO_TMA_OUT)) && max_row >= 32) {
    if (int rc_out = make_tmap_2d_f32_box32(&kp.tmOut, a->out, (uint64_t)max_row, (uint64_t)a->N, (uint64_t)a->ldo * 4))
      return rc_out;
    kp.use_tma_out = 1;
  }
  G2_REQUIRE(a->k_chunk_blocks >= 0, "gemm: negative k_chunk_blocks");
  G2_REQUIRE(a->k_chunk_blocks == 0 || (a->epilogue == G2VLM_EPI_STORE_F32 && !(a->flags & G2VLM_GEMM_ROUND_BF16)),
             "gemm: k_chunk_blocks needs the STORE_F32 epilogue without ROUND_BF16");
  // the chunks accumulate IN `out`, so the residual operand (added with the last chunk) must live elsewhere
  G2_REQUIRE(a->k_chunk_blocks == 0 || a->residual == nullptr || a->residual != a->out,
             "gemm: with k_chunk_blocks the residual must not alias out");
  kp.k_chunk_kb = a->k_chunk_blocks;

  int rc = make_tmap_2d_bf16(&kp.tmA, a->A, (uint64_t)a->a_rows, (uint64_t)a->K, (uint64_t)a->lda * 2, BM, BK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&kp.tmB, a->B, (uint64_t)a->n_groups * a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, BN, BK);
  if (rc) return rc;

  // Large problems run on CTA pairs (256 x 256 tiles, cta_group::2); small ones keep one CTA per 128 x 256 tile,
  // which gives twice as many independent tiles to spread over the SMs.  The G2VLM_GEMM_FORCE_PAIR / _FORCE_SINGLE
  // flags choose per call (tests run both kernels on the same shapes; A/B timing).
  int mp = 0;
  for (int g = 0; g < a->n_groups; ++g) {
    kp.grp_mpair0[g] = mp;
    mp += cdiv(cdiv(a->group_rows[g], BM), 2);
  }
  kp.grp_mpair0[a->n_groups] = mp;
  if (a->n_groups == 1) kp.grp_mpair0[2] = mp;
  kp.num_pair_tiles = mp * kp.n_tiles_n;
  // Panel height: the A rows of a panel stay in L2 while every N tile sweeps over them, so short-K problems want
  // tall panels (B = the weights stream from HBM once per panel) and long-K problems short ones (the A panel
  // itself must fit next to B).  Target ~24 MB of A per panel (same-box sweep: gate/up K=1536 857 us at 32 pairs
  // vs 906 at 4; down K=8960 431 us at 4 vs 459 at 32).
  {
    const long long per_pair_bytes = 2LL * BM * a->K * 2;
    long long ph = (24LL << 20) / per_pair_bytes;
    kp.pair_panel = (int)(ph < 4 ? 4 : (ph > 32 ? 32 : ph));
  }
  G2_REQUIRE((a->flags & (G2VLM_GEMM_FORCE_PAIR | G2VLM_GEMM_FORCE_SINGLE)) !=
                 (G2VLM_GEMM_FORCE_PAIR | G2VLM_GEMM_FORCE_SINGLE),
             "gemm: FORCE_PAIR and FORCE_SINGLE are mutually exclusive");
  G2_REQUIRE(!(a->k_chunk_blocks > 0 && (a->flags & G2VLM_GEMM_FORCE_PAIR)), "gemm: k_chunk_blocks runs on the 1-CTA kernel");
  const bool use_pair = a->k_chunk_blocks > 0                  ? false
                        : (a->flags & G2VLM_GEMM_FORCE_PAIR)   ? true
                        : (a->flags & G2VLM_GEMM_FORCE_SINGLE) ? false
                                                               : kp.num_pair_tiles >= num_sms();
  if (use_pair) {
    rc = make_tmap_2d_bf16(&kp.tmBh, a->B, (uint64_t)a->n_groups * a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, BN / 2,
                           BK);
    if (rc) return rc;
    switch (a->epilogue) {
      case G2VLM_EPI_STORE_BF16: return launch_gemm_pair<G2VLM_EPI_STORE_BF16>(kp, stream);
      case G2VLM_EPI_SWIGLU_BF16: return launch_gemm_pair<G2VLM_EPI_SWIGLU_BF16>(kp, stream);
      case G2VLM_EPI_RESID_F32: return launch_gemm_pair<G2VLM_EPI_RESID_F32>(kp, stream);
      default: return launch_gemm_pair<G2VLM_EPI_STORE_F32>(kp, stream);
    }
  }
  switch (a->epilogue) {
    case G2VLM_EPI_STORE_BF16: return launch_gemm<G2VLM_EPI_STORE_BF16>(kp, stream);
    case G2VLM_EPI_SWIGLU_BF16: return launch_gemm<G2VLM_EPI_SWIGLU_BF16>(kp, stream);
    case G2VLM_EPI_RESID_F32: return launch_gemm<G2VLM_EPI_RESID_F32>(kp, stream);
    default: return launch_gemm<G2VLM_EPI_STORE_F32>(kp, stream);
  }
}
