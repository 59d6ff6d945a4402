// Microbenchmark behind csrc/decode_fused.cu: HBM -> shared-memory streaming rate of one producer warp per SM issuing
// cp.async.bulk copies into an mbarrier ring, as a function of the copy size, copies per stage and ring depth.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_bench bulk_copy_bench.cu ; run: ./bulk_copy_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Each CTA streams `bytes_per_cta` contiguous bytes starting at base + cta * bytes_per_cta: stage = copies x piece bytes,
// copy j of a stage reads piece bytes at offset j * row_stride of the stage's window (row_stride >= piece).
__global__ void __launch_bounds__(64) stream_kernel(const uint8_t* base, long long bytes_per_cta, int piece, int copies, long long row_stride,
                                                    int ns, int stage_smem, unsigned long long* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  uint8_t* ring = sm + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ns; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long window = (long long)copies * row_stride;
  const int n_stages = (int)(bytes_per_cta / window);
  const uint8_t* src0 = base + (long long)blockIdx.x * bytes_per_cta;
  if (warp == 0) {
    for (int q = 0; q < n_stages; ++q) {
      const int slot = q % ns;
      if (q >= ns) mbar_wait(empty + slot, ((q / ns) - 1) & 1);
      if (lane == 0) mbar_expect(full + slot, (uint32_t)piece * copies);
      __syncwarp();
      for (int j = lane; j < copies; j += 32)
        bulk(ring + (long long)slot * stage_smem + (long long)j * piece, src0 + q * window + j * row_stride, piece, full + slot);
    }
  } else {
    unsigned long long acc = 0;
    for (int q = 0; q < n_stages; ++q) {
      const int slot = q % ns;
      mbar_wait(full + slot, (q / ns) & 1);
      acc += *reinterpret_cast<const unsigned long long*>(ring + (long long)slot * stage_smem + lane * 8);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
    if (acc == 0x1234567) sink[0] = acc;
  }
}

// register streaming for comparison: every warp reads 16-byte chunks, U loads in flight per lane
template <int U>
__global__ void __launch_bounds__(512) ldg_kernel(const uint4* base, long long chunks_per_cta, unsigned long long* sink) {
  const uint4* p = base + (long long)blockIdx.x * chunks_per_cta;
  unsigned long long acc = 0;
  for (long long c = threadIdx.x; c < chunks_per_cta; c += 512 * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long cc = c + 512LL * u;
      if (cc < chunks_per_cta) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + cc));
      else v[u] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].w;
  }
  if (acc == 0x1234567) sink[0] = acc;
}


// The GEMV consumer of decode_fused.cu on the same ring: stage = 8 rows x 3 KB; warp pair j owns slot j (warp 2j rows 0-3,
// warp 2j+1 rows 4-7), lanes on consecutive 16-byte chunks, packed fp32 FMAs against a bf16 vector in shared memory.
__device__ __forceinline__ unsigned long long f2(uint32_t w) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(w << 16), "r"(w & 0xffff0000u)); return r; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int VARIANT>
__global__ void __launch_bounds__(288) gemv_ring_kernel(const uint8_t* base, long long bytes_per_cta, float* out) {
  constexpr int NS = 4, ROWS = 8, PIECE = 3072, PITCH = 3136, STAGE = ROWS * PITCH;
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  uint4* vec = reinterpret_cast<uint4*>(sm + 256);
  uint8_t* ring = sm + 256 + 4096;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 192; i += blockDim.x) vec[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  __syncthreads();
  const int n_stages = (int)(bytes_per_cta / (ROWS * PIECE));
  const uint8_t* src0 = base + (long long)blockIdx.x * bytes_per_cta;
  if (warp == 8) {
    for (int q = 0; q < n_stages; ++q) {
      const int slot = q % NS;
      if (q >= NS) mbar_wait(empty + slot, ((q / NS) - 1) & 1);
      if (lane == 0) mbar_expect(full + slot, ROWS * PIECE);
      __syncwarp();
      if (lane < ROWS) bulk(ring + slot * STAGE + lane * PITCH, src0 + (long long)q * ROWS * PIECE + lane * PIECE, PIECE, full + slot);
    }
    return;
  }
  const int slot = warp >> 1, r0 = (warp & 1) * 4;
  float total = 0.f;
  for (int q = slot; q < n_stages; q += NS) {
    mbar_wait(full + slot, (q / NS) & 1);
    uint4 x[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) x[u] = vec[lane + 32 * u];
    float sums[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint4* wr = reinterpret_cast<const uint4*>(ring + slot * STAGE + (r0 + r) * PITCH);
      uint4 a[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) a[u] = wr[lane + 32 * u];
      if (VARIANT == 0) {
        unsigned long long acc0 = 0, acc1 = 0;
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          acc0 = ffma2(f2(a[u].x), f2(x[u].x), acc0); acc1 = ffma2(f2(a[u].y), f2(x[u].y), acc1);
          acc0 = ffma2(f2(a[u].z), f2(x[u].z), acc0); acc1 = ffma2(f2(a[u].w), f2(x[u].w), acc1);
        }
        sums[r] = __uint_as_float((uint32_t)acc0) + __uint_as_float((uint32_t)(acc0 >> 32)) + __uint_as_float((uint32_t)acc1) + __uint_as_float((uint32_t)(acc1 >> 32));
      } else {
        uint32_t acc = 0;
#pragma unroll
        for (int u = 0; u < 6; ++u) acc += a[u].x ^ a[u].y ^ a[u].z ^ a[u].w;
        sums[r] = __uint_as_float(acc & 0x3fffffffu);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sums[r] += __shfl_xor_sync(0xffffffffu, sums[r], o);
      total += sums[r];
    }
  }
  if (lane == 0) out[blockIdx.x * 8 + warp] = total;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const long long total = 4LL << 30;
  uint8_t* buf;
  unsigned long long* sink;
  cudaMalloc(&buf, total);
  cudaMalloc(&sink, 8);
  cudaMemset(buf, 1, total);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const long long per_cta = (total / sms) & ~((1LL << 20) - 1);
  struct Cfg { int piece, copies; long long stride; int ns; const char* what; };
  const Cfg cfgs[] = {
      {1536, 16, 3072, 4, "16 x 1536 B of 3 KB rows, 4 stages (the first v3 ring)"},
      {1536, 16, 3072, 8, "same, 8 stages"},
      {3072, 8, 3072, 4, "8 x 3 KB rows, 4 stages"},
      {3072, 16, 3072, 4, "16 x 3 KB rows (48 KB stages), 4 stages"},
      {512, 32, 512, 4, "32 x 512 B contiguous, 4 stages"},
      {8192, 3, 8192, 4, "3 x 8 KB contiguous, 4 stages"},
      {24576, 1, 24576, 4, "1 x 24 KB, 4 stages"},
      {24576, 1, 24576, 8, "1 x 24 KB, 8 stages"},
      {49152, 1, 49152, 4, "1 x 48 KB, 4 stages"},
      {12288, 1, 12288, 16, "1 x 12 KB, 16 stages"},
      {1280, 11, 17920, 4, "11 x 1280 B of 17.9 KB rows (down), 4 stages"},
  };
  for (const Cfg& c : cfgs) {
    const int stage_smem = c.piece * c.copies;
    const size_t smem = 256 + (size_t)stage_smem * c.ns;
    if (smem > 220 * 1024) continue;
    const long long window = c.copies * c.stride;
    const long long bytes_cta = per_cta / window * window;
    float best = 1e9f;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      stream_kernel<<<sms, 64, smem>>>(buf, bytes_cta, c.piece, c.copies, c.stride, c.ns, stage_smem, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double moved = (double)(bytes_cta / window) * c.piece * c.copies * sms;
    printf("bulk ring  %-62s %7.1f GB/s  (%s)\n", c.what, moved / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  {
    const long long chunks = per_cta / 16;
    float best = 1e9f;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      ldg_kernel<8><<<sms, 512>>>(reinterpret_cast<const uint4*>(buf), chunks, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("ld.global.nc 16 B x 8 in flight, 512 threads / SM:                         %7.1f GB/s\n", (double)chunks * 16 * sms / best / 1e6);
    best = 1e9f;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      ldg_kernel<16><<<sms, 512>>>(reinterpret_cast<const uint4*>(buf), chunks, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("ld.global.nc 16 B x 16 in flight, 512 threads / SM:                        %7.1f GB/s\n", (double)chunks * 16 * sms / best / 1e6);
  }
  {
    float* out;
    cudaMalloc(&out, 148 * 8 * 4 * 2);
    const size_t smem = 256 + 4096 + 4 * 8 * 3136;
    cudaFuncSetAttribute(gemv_ring_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(gemv_ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long bytes_cta = per_cta / (8 * 3072) * (8 * 3072);
    for (int v = 0; v < 2; ++v) {
      float best = 1e9f;
      for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        if (v == 0) gemv_ring_kernel<0><<<sms, 288, smem>>>(buf, bytes_cta, out);
        else gemv_ring_kernel<1><<<sms, 288, smem>>>(buf, bytes_cta, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("GEMV ring consumer (8 warps, %s): %7.1f GB/s (%s)\n", v == 0 ? "packed fp32 FMA dot products" : "xor only", (double)bytes_cta * sms / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
