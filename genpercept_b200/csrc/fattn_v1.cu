// Fused self-attention forward for head_dim 64 (the SD-2.1 UNet's BasicTransformerBlock.attn1):
//   O = softmax(Q K^T) V   per (image, head), non-causal, fp32 softmax state, 16-bit operands.
// (Softmax scale is folded into Wq at load.)  FlashAttention-style online softmax on tcgen05, with
// TWO 128-row query tiles per CTA that ping-pong on the tensor pipe and share every K/V block:
//
//   warp 8 lane 0 : TMA producer  — both Q tiles once; K block [128 keys x 64] + V^T block
//                                   [64 x 128 keys] per iteration into a 4-stage ring
//   warps 9,10 l.0: MMA issuers   — per tile t and block j:  S_t = Q_t K_j^T (128x128x64, one TMEM
//                                   buffer per tile);  O_t,j = P_t,j V_j (128x64x128, fresh TMEM tile,
//                                   double buffered)
//   warp 11       : TMEM allocator (512 columns: S_A, S_B, O_A[2], O_B[2])
//   warps 0..3    : softmax of tile A, warps 4..7  : softmax of tile B (issuers sit in the highest
//                   warp ids: the sub-partition arbiter favours higher ids and must not starve them) — one query row per thread:
//                   two passes over the S row in TMEM (max, then exp2 / sum), P written 16-bit into
//                   shared memory in the K-major SWIZZLE_128B operand layout, O accumulated in
//                   registers:  O <- (O + O_{j-1}) * 2^{m_{j-1} - m_j}.
// Every SM sub-partition hosts one softmax warp of each tile, so while tile A waits on its MMA /
// TMEM / MUFU latencies tile B computes (ncu r1c/r1d: a single tile left issue slots < 50 % busy and
// the tensor pipe at 17 %).  The kernel is MUFU-bound by construction: 128x128 exp2 per 512 tensor
// cycles at 16 exp2/clk/SM.
//
// S and P never touch HBM (the round-1 unfused path wrote both: 4 x T^2 x 2 bytes per head).
#include "fattn.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "ptx.cuh"

namespace gp {
namespace {

constexpr int kThreads = 384;
constexpr int kStages = 4;
constexpr int kQBytes = 128 * 64 * 2;          // 16 KiB per tile
constexpr int kKBytes = 128 * 64 * 2;          // 16 KiB
constexpr int kVBytes = 64 * 128 * 2;          // 16 KiB (two 64-key sub-tiles of 8 KiB)
constexpr int kPBytes = 128 * 128 * 2;         // 32 KiB per tile (two 64-key sub-tiles of 16 KiB)
constexpr int kSmemBytes = 2 * kQBytes + kStages * (kKBytes + kVBytes) + 2 * kPBytes + 256 + 1024;
constexpr int kTmemCols = 512;
constexpr int kOCol = 256;                     // S_A [0,128) S_B [128,256) O_t[buf] at 256 + t*128 + buf*64

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA pipe for x <= 0: round-to-nearest split x = n + f, |f| <= 0.5 (magic-number add), cubic
// minimax 2^f (max relative error 7.5e-5, below the 16-bit rounding of P), n added into the exponent
// field.  Used for one element in four of softmax pass 2, which is otherwise bound by the 16 ex2/clk/SM
// MUFU rate (10 issue slots vs 2, so the split that balances MUFU and issue is ~1:3).
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -126.f);
  const float magic = 12582912.f;              // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float r = x + magic;
  const float f = x - (r - magic);
  const float pl = fmaf(fmaf(fmaf(0.0551716685f, f, 0.242611125f), f, 0.693260968f), f, 0.999928057f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(r) << 23));
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
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <bool BF16, bool POLY>
__global__ void __launch_bounds__(kThreads, 1) fattn_v1_kernel(const __grid_constant__ FattnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // [tile][16 KiB]
  uint8_t* sK = sQ + 2 * kQBytes;                   // [stage][16 KiB]
  uint8_t* sV = sK + kStages * kKBytes;             // [stage][16 KiB]
  uint8_t* sP = sV + kStages * kVBytes;             // [tile][32 KiB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                     // [kStages]
  uint64_t* kv_empty = kv_full + kStages;           // [kStages]
  uint64_t* s_full = kv_empty + kStages;            // [tile]
  uint64_t* p_full = s_full + 2;                    // [tile]
  uint64_t* o_full = p_full + 2;                    // [tile][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int pairs = (p.q_tiles + 1) >> 1;
  const int qp = blockIdx.x % pairs;
  const int bh = blockIdx.x / pairs;
  const int head = bh % p.heads, b = bh / p.heads;
  const int T = p.T;
  const int nblk = (T + 127) >> 7;
  const int ntile = (2 * qp + 1 < p.q_tiles) ? 2 : 1;   // the last pair of an odd tile count is half empty

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], ntile); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
    for (int i = 0; i < 4; ++i) mbar_init(&o_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 11) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer (whole warp waits, one lane issues)
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(q_full, (uint32_t)(ntile * kQBytes));
      for (int t = 0; t < ntile; ++t) tma_load_3d(sQ + t * kQBytes, &p.tmQ, q_full, head * 64, (2 * qp + t) * 128, b);
    }
    for (int j = 0; j < nblk; ++j) {
      const int st = j % kStages;
      mbar_wait(&kv_empty[st], ((j / kStages) & 1) ^ 1, 10);
      if (leader) {
        mbar_expect_tx(&kv_full[st], kKBytes + kVBytes);
        tma_load_3d(sK + st * kKBytes, &p.tmK, &kv_full[st], head * 64, j * 128, b);
        tma_load_3d(sV + st * kVBytes, &p.tmV, &kv_full[st], j * 128, head * 64, b);
        tma_load_3d(sV + st * kVBytes + 8192, &p.tmV, &kv_full[st], j * 128 + 64, head * 64, b);
      }
      __syncwarp();
    }
  } else if ((warp == 9 || warp == 10) && (warp - 9) < ntile) {
    // ------------------------------------------------------------------ MMA issuers: warp 9 -> tile A, warp 10 -> tile B
    // One issuing warp per tile (a shared issuer made each tile wait behind the other's instruction
    // stream).  The whole warp runs the loop and waits on the barriers; one elected lane issues, so the
    // descriptors live in uniform registers.  The next score tile S_t,j+1 is issued before the P.V product
    // of block j (the softmax warps idle until it lands).
    const int t = warp - 9;
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_f16(128, 128, BF16 ? 1 : 0);
    const uint32_t idesc_o = make_idesc_f16(128, 64, BF16 ? 1 : 0);
    const uint64_t q_desc = make_sw128_kmajor_desc(smem_u32(sQ + t * kQBytes));
    const uint64_t p_desc = make_sw128_kmajor_desc(smem_u32(sP + t * kPBytes));
    const uint64_t k_desc0 = make_sw128_kmajor_desc(smem_u32(sK));
    const uint64_t v_desc0 = make_sw128_kmajor_desc(smem_u32(sV));
    const uint32_t s_tmem = tmem_base + t * 128;
    auto mma_s = [&](int st) {                // S_t = Q_t K_j^T
      const uint64_t k_desc = k_desc0 + (uint64_t)(st * (kKBytes >> 4));
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(s_tmem, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 12);
    mbar_wait(&kv_full[0], 0, 11);
    tc_fence_after();
    mma_s(0);
    for (int j = 0; j < nblk; ++j) {
      const int st = j % kStages;
      if (j + 1 < nblk) {
        mbar_wait(&kv_full[(j + 1) % kStages], ((j + 1) / kStages) & 1, 11);
        tc_fence_after();
      }
      const bool trm = p.trace != nullptr && blockIdx.x == 0 && t == 0 && j < 64 && leader;
      if (trm) p.trace[512 + j * 4 + 0] = clock64();
      mbar_wait(&p_full[t], j & 1, 13);       // P_t,j is in shared memory and S_t has been consumed
      tc_fence_after();
      if (trm) p.trace[512 + j * 4 + 1] = clock64();
      if (j + 1 < nblk) mma_s((j + 1) % kStages);
      if (trm) p.trace[512 + j * 4 + 2] = clock64();
      const uint64_t v_desc = v_desc0 + (uint64_t)(st * (kVBytes >> 4));
      const uint32_t o_tmem = tmem_base + kOCol + t * 128 + (j & 1) * 64;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16(o_tmem, p_desc + (uint64_t)((kk >> 2) * (16384 >> 4) + 2 * (kk & 3)),
                   v_desc + (uint64_t)((kk >> 2) * (8192 >> 4) + 2 * (kk & 3)), idesc_o, kk ? 1u : 0u);
        umma_commit(&o_full[t * 2 + (j & 1)]);
        umma_commit(&kv_empty[st]);           // this tile has issued every use of block j (barrier counts both tiles)
      }
      __syncwarp();
      if (trm) p.trace[512 + j * 4 + 3] = clock64();
    }
  } else if (warp < 8 && (warp >> 2) < ntile) {
    // ------------------------------------------------------------------ softmax + output of tile t
    const int t = warp >> 2;
    const int wq = warp & 3;             // == warp % 4 -> TMEM lanes [32*wq, 32*wq+32)
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float c2 = p.scale_log2e;
    float m = -INFINITY, l = 0.f;
    float O[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) O[i] = 0.f;
    const uint32_t prow = smem_u32(sP + t * kPBytes) + row * 128;
    const int sw = row & 7;
    const uint32_t ts = tmem_base + lane_off + t * 128;
    // Ping-pong: the exp2-heavy pass 2 of the two tiles is forced to alternate (token passed through
    // named barriers 2 / 3), so one tile's MUFU phase overlaps the other tile's wait / max / rescale
    // phase instead of both tiles drifting into lock-step (r1f: 2535 cycles per tile-block vs the
    // 1024-cycle MUFU bound).  Tile B hands the first token to tile A.
    // (measured: the forced alternation is slower, 13.0 vs 11.9 ms per step, r1g; kept behind a switch)
    const bool pingpong = false && ntile == 2;
    if (pingpong && t == 1) asm volatile("bar.arrive 2, 256;" ::: "memory");
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && t == 0 && wq == 0 && lane == 0;
    for (int j = 0; j < nblk; ++j) {
      if (tr && j < 64) p.trace[j * 8 + 0] = clock64();
      mbar_wait(&s_full[t], j & 1, 15);
      tc_fence_after();
      if (tr && j < 64) p.trace[j * 8 + 1] = clock64();
      const int kvalid = min(128, T - j * 128);
      uint32_t ra[32], rb[32];
      // pass 1: row maximum (4 independent chains, TMEM loads one chunk ahead)
      float mx0 = m, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      tmem_ld_32x32(ts, ra);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        uint32_t* cur = (c & 1) ? rb : ra;
        if (c < 3) tmem_ld_32x32(ts + (c + 1) * 32, (c & 1) ? ra : rb);
        if (kvalid == 128) {
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(cur[q]));
            mx1 = fmaxf(mx1, __uint_as_float(cur[q + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(cur[q + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(cur[q + 3]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) if (c * 32 + q < kvalid) mx0 = fmaxf(mx0, __uint_as_float(cur[q]));
        }
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float alpha = ex2((m - mx) * c2);
      if (tr && j < 64) p.trace[j * 8 + 2] = clock64();
      tmem_ld_32x32(ts, ra);                       // first chunk of pass 2, in flight during the O update
      if (j > 0) {
        mbar_wait(&o_full[t * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1, 16);
        tc_fence_after();
        if (tr && j < 64) p.trace[j * 8 + 3] = clock64();
        const uint32_t to = tmem_base + lane_off + kOCol + t * 128 + ((j - 1) & 1) * 64;
        tmem_ld_32x32(to, rb);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) O[q] = (O[q] + __uint_as_float(rb[q])) * alpha;
        tmem_ld_32x32(to + 32, rb);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) O[32 + q] = (O[32 + q] + __uint_as_float(rb[q])) * alpha;
      }
      l *= alpha;
      if (tr && j < 64) p.trace[j * 8 + 4] = clock64();
      if (pingpong) {
        if (t == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
        else asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      // pass 2: probabilities -> shared memory (A operand of P.V), row sum (4 partial sums)
      const float mb = mx * c2;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        uint32_t* cur = (c & 1) ? rb : ra;
        if (c < 3) tmem_ld_32x32(ts + (c + 1) * 32, (c & 1) ? ra : rb);
        float pv[32];
        if (kvalid == 128) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const float x = __uint_as_float(cur[q]) * c2 - mb;
            pv[q] = (POLY && (q & 3) == 3) ? ex2_fma(x) : ex2(x);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) pv[q] = (c * 32 + q < kvalid) ? ex2(__uint_as_float(cur[q]) * c2 - mb) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 32; q += 4) { l0 += pv[q]; l1 += pv[q + 1]; l2 += pv[q + 2]; l3 += pv[q + 3]; }
        const uint32_t dst = prow + (c >> 1) * 16384;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_shared_v4(dst + ((((c & 1) * 4 + i) ^ sw) << 4), pack16<BF16>(pv[8 * i + 0], pv[8 * i + 1]),
                       pack16<BF16>(pv[8 * i + 2], pv[8 * i + 3]), pack16<BF16>(pv[8 * i + 4], pv[8 * i + 5]),
                       pack16<BF16>(pv[8 * i + 6], pv[8 * i + 7]));
      }
      l += (l0 + l1) + (l2 + l3);
      if (tr && j < 64) p.trace[j * 8 + 5] = clock64();
      if (pingpong) {                             // hand the MUFU phase to the other tile
        if (t == 0) asm volatile("bar.arrive 3, 256;" ::: "memory");
        else if (j + 1 < nblk) asm volatile("bar.arrive 2, 256;" ::: "memory");
      }
      tc_fence_before();                          // S_t reads are complete before the MMA warp overwrites it
      fence_proxy_async_smem();                   // P_t visible to the tensor core (async proxy)
      mbar_arrive(&p_full[t]);
      if (tr && j < 64) p.trace[j * 8 + 6] = clock64();
      m = mx;
    }
    // last partial product, normalise, store
    mbar_wait(&o_full[t * 2 + ((nblk - 1) & 1)], ((nblk - 1) >> 1) & 1, 17);
    tc_fence_after();
    {
      const uint32_t to = tmem_base + lane_off + kOCol + t * 128 + ((nblk - 1) & 1) * 64;
      const float inv = 1.f / l;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        tmem_ld_32x32(to + h * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) O[h * 32 + q] = (O[h * 32 + q] + __uint_as_float(r[q])) * inv;
      }
    }
    const int qrow = (2 * qp + t) * 128 + row;
    if (qrow < T) {
      uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.out_b_stride +
                     (long long)qrow * p.out_row_stride + head * 64;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
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
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace


cudaError_t fattn_v1_launch(const FattnParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  static bool poly = true;
  if (!attr_set) {
    const void* fns[4] = {(const void*)fattn_v1_kernel<false, false>, (const void*)fattn_v1_kernel<false, true>,
                          (const void*)fattn_v1_kernel<true, false>, (const void*)fattn_v1_kernel<true, true>};
    for (const void* f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) return e;
    }
    const char* env = getenv("GP_FATTN_POLY");   // 0: every exp2 on the MUFU (A/B switch)
    if (env && env[0] == '0') poly = false;
    attr_set = true;
  }
  const int grid = p.B * p.heads * ((p.q_tiles + 1) / 2);
  if (grid <= 0) return cudaSuccess;
  if (p.bf16) {
    if (poly) fattn_v1_kernel<true, true><<<grid, kThreads, kSmemBytes, stream>>>(p);
    else fattn_v1_kernel<true, false><<<grid, kThreads, kSmemBytes, stream>>>(p);
  } else {
    if (poly) fattn_v1_kernel<false, true><<<grid, kThreads, kSmemBytes, stream>>>(p);
    else fattn_v1_kernel<false, false><<<grid, kThreads, kSmemBytes, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace gp
