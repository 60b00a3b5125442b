// Micro-benchmark: throughput of ex2.approx in f32 / f16x2 / bf16x2 form and of an FMA-pipe polynomial exp2,
// per SM sub-partition (decides how fattn's softmax pass 2 should produce its probabilities).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/exp_ex2 scripts/exp_ex2.cu && build/exp_ex2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float ex2poly(float x) {
  x = fmaxf(x, -126.f);
  const float magic = 12582912.f;
  float r = x + magic;
  float f = x - (r - magic);
  float p = fmaf(fmaf(fmaf(0.05550357f, f, 0.24022651f), f, 0.69314720f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b400u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.0f;
      if (MODE == 1) h[i] = ex2h2(h[i]) ^ 0x80008000u;
      if (MODE == 2) h[i] = ex2b2(h[i]) ^ 0x80008000u;
      if (MODE == 3) a[i] = ex2poly(a[i]) - 1.0f;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  const char* names[4] = {"ex2.f32", "ex2.f16x2", "ex2.bf16x2", "poly3 (fma pipe)"};
  for (int threads = 128; threads <= 512; threads *= 2)
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
        if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
      }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double instr = double(iters) * 8 * threads;        // ex2 instructions (thread-level) per CTA
      const double elems = instr * ((mode == 1 || mode == 2) ? 2 : 1);
      printf("%-18s threads/SM %4d: %8lld cycles  %6.2f thread-instr/clk/SM  %6.2f results/clk/SM\n", names[mode], threads, c,
             instr / c, elems / c);
    }
  // accuracy of the polynomial and of the packed forms
  return 0;
}
