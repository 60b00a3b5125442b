// Experiment: can a tcgen05 K-major SWIZZLE_128B A-operand descriptor start at an arbitrary 128-byte
// row of a larger TMA-written tile (needed for 3x3 halo reuse: tap (dy,dx) = row offset into one
// shared patch)?  Tests start-row offsets 0..9 with (a) base_offset = (addr >> 7) & 7 and (b) 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I genpercept_b200/csrc scripts/exp_baseoffset.cu -o build/exp_baseoffset
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace gp;

constexpr int ROWS = 144;   // patch rows in shared memory
constexpr int K = 64, N = 64;

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                            float* out, int start_row, int use_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // ROWS x 128 B
  uint8_t* sB = smem + 20480;               // 64 x 128 B (1024-aligned)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20480 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * 128 + 64 * 128);
    tma_load_3d(sA, &tmA, bar, 0, 0, 0);
    tma_load_3d(sB, &tmB, bar, 0, 0, 0);
    mbar_wait(bar, 0, 1);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + start_row * 128;
    uint64_t ad = make_sw128_kmajor_desc(a_addr);
    if (use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bd = make_sw128_kmajor_desc(smem_u32(sB));
    const uint32_t idesc = make_idesc_f16(128, N, 0);
    for (int kk = 0; kk < 4; ++kk) umma_f16(tm, ad + 2 * kk, bd + 2 * kk, idesc, kk ? 1u : 0u);
    umma_commit(done);
  }
  mbar_wait(done, 0, 2);
  tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) out[(warp * 32 + lane) * N + c + q] = __uint_as_float(r[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 64);
}

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)f;
  std::vector<__half> hA(ROWS * K), hB(N * K);
  std::vector<float> fA(ROWS * K), fB(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2half(v); fA[i] = v; }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 8.0f; hB[i] = __float2half(v); fB[i] = v; }
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)ROWS, 1};
    cuuint64_t str[2] = {(cuuint64_t)K * 2, (cuuint64_t)ROWS * K * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)ROWS, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode A: %d\n", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, 1};
    cuuint64_t str[2] = {(cuuint64_t)K * 2, (cuuint64_t)N * K * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)N, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode B: %d\n", (int)r);
  }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  std::vector<float> ho(128 * N);
  for (int mode = 0; mode < 2; ++mode)
    for (int s = 0; s <= 9; ++s) {
      cudaMemset(dO, 0, 128 * N * 4);
      k<<<1, 128, 40960>>>(tA, tB, dO, s, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("start_row %d base_offset %d: CUDA error %s\n", s, mode, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho.data(), dO, ho.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int kk = 0; kk < K; ++kk) ref += (double)fA[(m + s) * K + kk] * fB[n * K + kk];
          maxerr = fmax(maxerr, fabs(ref - ho[m * N + n]));
        }
      printf("start_row %d  base_offset_field %s : max|err| = %.4f  %s\n", s, mode ? "(addr>>7)&7" : "0", maxerr,
             maxerr < 1e-2 ? "OK" : "MISMATCH");
    }
  return 0;
}
