// Throughput and latency of the legacy warp-level MMA (mma.sync.m16n8k16 bf16 -> fp32) on sm_100a: the path the skinny-row
// GEMV / decode attention kernels use.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_bench mma_sync_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(float* out, int iters, long long* cycles) {
  float acc[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  unsigned a = 0x3f803f80u + threadIdx.x, b = 0x3f803f80u;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(acc[c], a, a, a, a, b, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS>
void run(int warps, float* out, long long* cyc) {
  const int iters = 4096;
  k<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_sm = (double)c / ((double)iters * CHAINS * warps);
  printf("%2d warps / SM, %d independent chains per warp: %7.2f cycles per MMA and SM (%6.1f per dependent MMA in a warp), %6.0f dense FLOP / cycle / SM\n",
         warps, CHAINS, per_sm, (double)c / iters, 4096.0 / per_sm);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<1>(1, out, cyc); run<4>(1, out, cyc); run<8>(1, out, cyc);
  run<1>(4, out, cyc); run<4>(4, out, cyc); run<8>(4, out, cyc);
  run<4>(8, out, cyc); run<8>(8, out, cyc); run<4>(16, out, cyc);
  return 0;
}
