// Special-function throughput on sm_100a, per SM: which formulation of SiLU is cheapest for the operand transform of
// igemm_patch.cu (one warp per sub-partition) and for gn_apply (many warps).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exp_mufu scripts/exp_mufu.cu && /tmp/exp_mufu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float ex2_fma(float x) {      // fattn.cu: cubic 2^f on the FMA pipe
  x = fmaxf(x, -126.f);
  const float magic = 12582912.f;
  const float r = x + magic;
  const float f = x - (r - magic);
  const float pl = fmaf(fmaf(fmaf(0.0551716685f, f, 0.242611125f), f, 0.693260968f), f, 0.999928057f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(r) << 23));
}

template <int V>
__device__ __forceinline__ float op(float x) {
  float y;
  if (V == 0) { asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  if (V == 1) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  if (V == 2) { asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  if (V == 3) {   // tanh.approx.f16x2: two results per instruction
    uint32_t xi = __float_as_uint(x), yi;
    asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
    return __uint_as_float(yi);
  }
  if (V == 4) {
    uint32_t xi = __float_as_uint(x), yi;
    asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
    return __uint_as_float(yi);
  }
  if (V == 5) {   // silu via tanh
    const float h = 0.5f * x;
    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(h));
    return fmaf(h, y, h);
  }
  if (V == 6) {   // silu via ex2 + rcp
    float e, r;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return x * r;
  }
  if (V == 7) {   // silu via FMA-pipe ex2 + rcp
    float r;
    const float e = ex2_fma(x * -1.4426950408889634f);
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return x * r;
  }
  if (V == 8) {   // lg2 (another MUFU op, for reference)
    asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
  }
  if (V == 9) {   // rsqrt
    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
  }
  return x;
}

template <int V>
__global__ void k(float* out, int iters, long long* cycles) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i * 37) + 0.5f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<V>(a[i]) * 0.999f + 0.1f;       // 8 independent chains (+1 FFMA each)
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int threads, int results_per_op) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2048;
  k<V><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<V><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += h[i];
  mean /= 148;
  const double ops = (double)iters * 8 * threads * results_per_op;
  printf("%-34s %4d threads/SM : %7.2f results / clk / SM   (%.1f cycles per warp-instruction slot)\n", name, threads,
         ops / mean, mean / ((double)iters * 8 * (threads / 32) / 4.0));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int threads : {128, 512, 1024}) {
    run<0>("tanh.approx.f32", threads, 1);
    run<3>("tanh.approx.f16x2 (2 results)", threads, 2);
    run<1>("ex2.approx.ftz.f32", threads, 1);
    run<4>("ex2.approx.f16x2 (2 results)", threads, 2);
    run<2>("rcp.approx.ftz.f32", threads, 1);
    run<8>("lg2.approx.ftz.f32", threads, 1);
    run<9>("rsqrt.approx.ftz.f32", threads, 1);
    run<5>("silu = h + h*tanh(h)", threads, 1);
    run<6>("silu = x * rcp(1 + ex2(-x*log2e))", threads, 1);
    run<7>("silu, ex2 on the FMA pipe + rcp", threads, 1);
    printf("\n");
  }
  return 0;
}
