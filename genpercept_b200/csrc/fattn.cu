// Fused self-attention forward for head_dim 64 (the SD-2.1 UNet's BasicTransformerBlock.attn1):
//   O = softmax(Q K^T) V   per (image, head), non-causal, fp32 softmax state, 16-bit operands.
// (Softmax scale is folded into Wq at load.)  FlashAttention-style online softmax on tcgen05:
//
//   warp 0 lane 0 : TMA producer  — Q tile once; K block [128 keys x 64] + V^T block [64 x 128 keys]
//                                   per iteration into a 3-stage ring
//   warp 1 lane 0 : MMA issuer    — S_j = Q K_j^T (128x128x64, TMEM, double buffered);
//                                   O_j = P_j V_j (128x64x128, fresh TMEM tile, double buffered)
//   warp 2        : TMEM allocator
//   warps 4..11   : softmax       — TWO threads per query row (warps w and w+4 share TMEM lanes
//                                   32*(w%4)..+31): each owns 64 of the 128 key columns of S_j and 32
//                                   of the 64 output columns.  Two passes over the S half-row in TMEM
//                                   (max, then exp2 / sum); the row maximum is exchanged through
//                                   shared memory once per block; P_j is written 16-bit into shared
//                                   memory in the K-major SWIZZLE_128B operand layout; O is
//                                   accumulated in registers: O <- (O + O_{j-1}) * 2^{m_{j-1} - m_j}.
//   Two softmax warps per SM sub-partition hide the ALU/MUFU/TMEM latencies that a single warp per
//   sub-partition left exposed (ncu r1c: issue slots 47 % busy, tensor pipe 17 %).
//
// S and P never touch HBM (the round-1 unfused path wrote both: 4 x T^2 x 2 bytes per head).
#include "fattn.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace gp {
namespace {

constexpr int kThreads = 384;
constexpr int kStages = 3;
constexpr int kQBytes = 128 * 64 * 2;          // 16 KiB
constexpr int kKBytes = 128 * 64 * 2;          // 16 KiB
constexpr int kVBytes = 64 * 128 * 2;          // 16 KiB (two 64-key sub-tiles of 8 KiB)
constexpr int kPBytes = 128 * 128 * 2;         // 32 KiB (two 64-key sub-tiles of 16 KiB)
constexpr int kXchgBytes = 2 * 2 * 128 * 4;    // row-max exchange [parity][half][row] + reused for row sums
constexpr int kSmemBytes = kQBytes + kStages * (kKBytes + kVBytes) + kPBytes + kXchgBytes + 256 + 1024;
constexpr int kTmemCols = 512;
constexpr int kOCol = 256;                     // S0: [0,128) S1: [128,256) O0: [256,320) O1: [320,384)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if constexpr (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pair_sync(int id) {   // the two warps that share a 32-row slice
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1) fattn_kernel(const __grid_constant__ FattnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;                       // [stage][16 KiB]
  uint8_t* sV = sK + kStages * kKBytes;             // [stage][16 KiB]
  uint8_t* sP = sV + kStages * kVBytes;
  float* xchg = reinterpret_cast<float*>(sP + kPBytes);          // [2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes + kXchgBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                     // [3]
  uint64_t* kv_empty = bars + 4;                    // [3]
  uint64_t* s_full = bars + 7;                      // [2]
  uint64_t* s_empty = bars + 9;                     // [2]
  uint64_t* p_full = bars + 11;
  uint64_t* o_full = bars + 12;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int head = bh % p.heads, b = bh / p.heads;
  const int T = p.T;
  const int nblk = (T + 127) >> 7;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 256); mbar_init(&o_full[i], 1); }
    mbar_init(p_full, 256);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    mbar_expect_tx(q_full, kQBytes);
    tma_load_3d(sQ, &p.tmQ, q_full, head * 64, qt * 128, b);
    for (int j = 0; j < nblk; ++j) {
      const int st = j % kStages;
      mbar_wait(&kv_empty[st], ((j / kStages) & 1) ^ 1, 10);
      mbar_expect_tx(&kv_full[st], kKBytes + kVBytes);
      tma_load_3d(sK + st * kKBytes, &p.tmK, &kv_full[st], head * 64, j * 128, b);
      tma_load_3d(sV + st * kVBytes, &p.tmV, &kv_full[st], j * 128, head * 64, b);
      tma_load_3d(sV + st * kVBytes + 8192, &p.tmV, &kv_full[st], j * 128 + 64, head * 64, b);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = make_idesc_f16(128, 128, BF16 ? 1 : 0);
    const uint32_t idesc_o = make_idesc_f16(128, 64, BF16 ? 1 : 0);
    const uint64_t q_desc = make_sw128_kmajor_desc(smem_u32(sQ));
    auto mma_s = [&](int j) {
      const int st = j % kStages;
      mbar_wait(&kv_full[st], (j / kStages) & 1, 11);
      tc_fence_after();
      const uint64_t k_desc = make_sw128_kmajor_desc(smem_u32(sK + st * kKBytes));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base + (j & 1) * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k ? 1u : 0u);
      umma_commit(&s_full[j & 1]);
    };
    mbar_wait(q_full, 0, 12);
    tc_fence_after();
    mma_s(0);
    if (nblk > 1) mma_s(1);
    for (int j = 0; j < nblk; ++j) {
      const int st = j % kStages;
      mbar_wait(p_full, j & 1, 13);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t a = make_sw128_kmajor_desc(smem_u32(sP + (kk >> 2) * 16384)) + 2 * (kk & 3);
        const uint64_t bd = make_sw128_kmajor_desc(smem_u32(sV + st * kVBytes + (kk >> 2) * 8192)) + 2 * (kk & 3);
        umma_f16(tmem_base + kOCol + (j & 1) * 64, a, bd, idesc_o, kk ? 1u : 0u);
      }
      umma_commit(&o_full[j & 1]);
      umma_commit(&kv_empty[st]);
      if (j + 2 < nblk) {
        mbar_wait(&s_empty[j & 1], (j >> 1) & 1, 14);
        tc_fence_after();
        mma_s(j + 2);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output
    const int half = (warp - 4) >> 2;          // which 64 key columns / 32 output columns
    const int wq = (warp - 4) & 3;             // == warp % 4 -> TMEM lanes [32*wq, 32*wq+32)
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float c2 = p.scale_log2e;
    float m = -INFINITY, l = 0.f;
    float O[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) O[i] = 0.f;
    uint8_t* prow = sP + half * 16384 + row * 128;       // this thread's 64-key sub-tile row
    const int sw = row & 7;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1, 15);
      tc_fence_after();
      const uint32_t ts = tmem_base + lane_off + (j & 1) * 128 + half * 64;
      const int kvalid = min(128, T - j * 128) - half * 64;   // valid columns among this thread's 64
      uint32_t ra[32], rb[32];
      // pass 1: maximum over this thread's 64 columns, then exchange with the partner thread
      tmem_ld_32x32(ts, ra);
      tmem_ld_32x32(ts + 32, rb);
      tmem_ld_wait();
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (kvalid >= 64) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(ra[q]), __uint_as_float(rb[q])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(ra[q + 1]), __uint_as_float(rb[q + 1])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(ra[q + 2]), __uint_as_float(rb[q + 2])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(ra[q + 3]), __uint_as_float(rb[q + 3])));
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          if (q < kvalid) mx0 = fmaxf(mx0, __uint_as_float(ra[q]));
          if (32 + q < kvalid) mx1 = fmaxf(mx1, __uint_as_float(rb[q]));
        }
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      float* xs = xchg + (j & 1) * 256;
      xs[half * 128 + row] = mx;
      pair_sync(1 + wq);
      mx = fmaxf(m, fmaxf(mx, xs[(half ^ 1) * 128 + row]));
      const float alpha = ex2((m - mx) * c2);
      if (j > 0) {
        mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1, 16);
        tc_fence_after();
        uint32_t ro[32];
        tmem_ld_32x32(tmem_base + lane_off + kOCol + ((j - 1) & 1) * 64 + half * 32, ro);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) O[q] = (O[q] + __uint_as_float(ro[q])) * alpha;
      }
      l *= alpha;
      // pass 2: probabilities -> shared memory (A operand of P.V), partial row sum
      const float mb = mx * c2;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t* cur = c ? rb : ra;
        float pv[32];
        if (kvalid >= 64) {
#pragma unroll
          for (int q = 0; q < 32; ++q) pv[q] = ex2(__uint_as_float(cur[q]) * c2 - mb);
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) pv[q] = (c * 32 + q < kvalid) ? ex2(__uint_as_float(cur[q]) * c2 - mb) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 32; q += 4) { l0 += pv[q]; l1 += pv[q + 1]; l2 += pv[q + 2]; l3 += pv[q + 3]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack16<BF16>(pv[8 * i + 0], pv[8 * i + 1]);
          u.y = pack16<BF16>(pv[8 * i + 2], pv[8 * i + 3]);
          u.z = pack16<BF16>(pv[8 * i + 4], pv[8 * i + 5]);
          u.w = pack16<BF16>(pv[8 * i + 6], pv[8 * i + 7]);
          *reinterpret_cast<uint4*>(prow + (((c * 4 + i) ^ sw) << 4)) = u;
        }
      }
      l += (l0 + l1) + (l2 + l3);
      tc_fence_before();
      mbar_arrive(&s_empty[j & 1]);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      m = mx;
    }
    // last partial product, total row sum (both halves), normalise, store
    float* xs = xchg + (nblk & 1) * 256;
    xs[half * 128 + row] = l;
    mbar_wait(&o_full[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1, 17);
    tc_fence_after();
    pair_sync(1 + wq);
    const float inv = 1.f / (l + xs[(half ^ 1) * 128 + row]);
    {
      uint32_t ro[32];
      tmem_ld_32x32(tmem_base + lane_off + kOCol + ((nblk - 1) & 1) * 64 + half * 32, ro);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) O[q] = (O[q] + __uint_as_float(ro[q])) * inv;
    }
    const int qrow = qt * 128 + row;
    if (qrow < T) {
      uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.out_b_stride +
                     (long long)qrow * p.out_row_stride + head * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack16<BF16>(O[8 * i + 0], O[8 * i + 1]);
        u.y = pack16<BF16>(O[8 * i + 2], O[8 * i + 3]);
        u.z = pack16<BF16>(O[8 * i + 4], O[8 * i + 5]);
        u.w = pack16<BF16>(O[8 * i + 6], O[8 * i + 7]);
        *reinterpret_cast<uint4*>(op + 8 * i) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

cudaError_t fattn_launch(const FattnParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fattn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fattn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = p.B * p.heads * p.q_tiles;
  if (grid <= 0) return cudaSuccess;
  if (p.bf16)
    fattn_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(p);
  else
    fattn_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace gp
