// Microbenchmark (builder tool): how close can ONE warp per scheduler come to the MUFU.EX2 rate (4 lanes / clk / scheduler,
// i.e. one warp instruction per 8 clocks), alone and with the instruction mix of the attention softmax
// (per score pair: 1 FFMA2, 2 MUFU.EX2, 1 FADD2, 1 F2FP)?  W = warps per scheduler (block = 128 * W threads, one block per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_warp_bench mufu_warp_bench.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// MODE 0: 32 independent ex2 per iteration, nothing else.  MODE 1: the softmax mix on 32 scores (16 pairs) per iteration.
template <int MODE>
__global__ void k(float* out, long long* clocks, int iters) {
  float s[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = -0.01f * (threadIdx.x % 7 + j);
  uint64_t sum2 = pack2(0.f, 0.f);
  uint32_t keep = 0;
  const uint64_t scale2 = pack2(0.999f, 0.999f), negm2 = pack2(-0.5f, -0.5f);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) s[j] = ex2f(s[j]);
#pragma unroll
      for (int j = 0; j < 32; ++j) s[j] = -s[j];   // FADD/FMUL-pipe filler so the values stay in range (1 op per ex2)
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const uint64_t x2 = fma2(pack2(s[j], s[j + 1]), scale2, negm2);
        float x0, x1;
        unpack2(x2, x0, x1);
        const float p0 = ex2f(x0), p1 = ex2f(x1);
        sum2 = add2(sum2, pack2(p0, p1));
        __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
        keep ^= *reinterpret_cast<uint32_t*>(&b);
        s[j] = -p0;
        s[j + 1] = -p1;
      }
    }
  }
  const long long t1 = clock64();
  float a0, a1, acc = 0.f;
  unpack2(sum2, a0, a1);
#pragma unroll
  for (int j = 0; j < 32; ++j) acc += s[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + a0 + a1 + __uint_as_float(keep);
  if (threadIdx.x == 0 && blockIdx.x == 0) clocks[0] = t1 - t0;
}

int main() {
  float* out; long long* clk; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int w = 1; w <= 4; w *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, 128 * w>>>(out, clk, iters); else k<1><<<148, 128 * w>>>(out, clk, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      const double per = (double)h / ((double)iters * 32.0 * w);   // clocks per MUFU warp instruction and scheduler
      printf("%s  warps/scheduler %d: %.2f clocks per MUFU.EX2 warp instruction per scheduler (8.00 = peak), per warp %.2f\n",
             mode == 0 ? "ex2 only      " : "softmax mix   ", w, per, per * w);
    }
  return 0;
}
