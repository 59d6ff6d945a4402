// Microbenchmark (builder tool): throughput of ex2.approx.ftz.f32 vs ex2.approx.f16x2 vs an FMA-pipe degree-3 exp2
// polynomial (f32x2 packed) on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void k(float* out, int iters) {
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = -0.001f * (threadIdx.x + j);
  uint32_t hacc[4] = {0xb800b800u, 0xb900b900u, 0xba00ba00u, 0xbb00bb00u};
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = ex2f(acc[j]) - 1.5f;          // 8 exps
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) hacc[j] = ex2h2(hacc[j]) ^ 0x80008000u;  // 8 exps in 4 ops
    } else {
      const uint64_t magic = pack2(12582912.f, 12582912.f), nmagic = pack2(-12582912.f, -12582912.f);
      const uint64_t c3 = pack2(0.0555f, 0.0555f), c2 = pack2(0.2402f, 0.2402f), c1 = pack2(0.6931f, 0.6931f), c0 = pack2(1.f, 1.f);
      const uint64_t none = pack2(-1.f, -1.f);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {                                    // 8 exps
        uint64_t x = pack2(acc[j], acc[j + 1]);
        uint64_t t = add2(x, magic);
        uint64_t n = add2(t, nmagic);
        uint64_t f = fma2(n, none, x);
        uint64_t p = fma2(c3, f, c2);
        p = fma2(p, f, c1);
        p = fma2(p, f, c0);
        float p0, p1, t0, t1;
        unpack2(p, p0, p1);
        unpack2(t, t0, t1);
        acc[j] = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)) - 1.5f;
        acc[j + 1] = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)) - 1.5f;
      }
    }
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += acc[j];
  for (int j = 0; j < 4; ++j) s += __uint_as_float(hacc[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const char* names[3] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "poly3 f32x2 (FMA pipe)"};
  for (int m = 0; m < 3; ++m) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<148 * 8, 256>>>(out, iters);
      if (m == 1) k<1><<<148 * 8, 256>>>(out, iters);
      if (m == 2) k<2><<<148 * 8, 256>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double exps = 148.0 * 8 * 256 * 8.0 * iters;
    printf("%-26s %.3f ms  %.1f Gexp/s  = %.2f exp/clk/SM at 1.9 GHz\n", names[m], ms, exps / ms / 1e6, exps / ms / 1e6 / 148 / 1.9);
  }
  return 0;
}
