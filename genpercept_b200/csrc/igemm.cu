// tcgen05 implicit-GEMM kernel (see igemm.h for the operand model).
//
// Warp roles (256 threads, 1 CTA / SM, persistent over a static round-robin tile list).  The single-thread
// producer / issuer roles sit in the HIGHEST warp ids (4..7): the SM sub-partition arbiter favours higher warp
// ids, and an issuer starved by the epilogue warp of its sub-partition stalls the tensor pipe (fattn trace, r1h).
// Each single-thread role is executed by its WHOLE warp (every lane walks the loop and waits on the barriers) and
// one elect.sync lane issues: coordinates and descriptors stay in uniform registers (back-to-back UTCHMMA).
//   warp 4        : TMA producer A (activation box per 64-channel K block, `stages`-deep ring)
//   warp 7        : TMA producer B (weight box per K block) — its own warp: one thread issuing
//                   both boxes plus the barrier traffic could not keep up with BN=128 tiles
//                   (ncu r1a: tensor pipe 45 % active on the 128->128 convs, DRAM/L2 not saturated)
//   warp 5 (and 6): MMA issuer(s), one per accumulator tile (4 tcgen05.mma 128xBNx16 per K block; commit frees the slot)
//   warp 6        : TMEM allocator (512 columns = 2 accumulator buffers x MT tiles)
//   warps 0..3    : epilogue      (tcgen05.ld 32 lanes x 32 columns -> bias / TMA-loaded residual / act -> staged tile
//                   -> TMA store, GroupNorm partial sums from the staged tile)
// MT = 2 (a 256-pixel M tile per CTA, two accumulators sharing every weight box) when BN <= 128:
// halves the weight traffic and the per-byte barrier / TMA issue cost of the narrow-N layers.
#include "igemm.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "ptx.cuh"

namespace gp {

namespace {

constexpr int kThreads = 256;
constexpr int kABytes = kBM * kBK * 2;       // 16 KiB per 128-row tile
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;              // TMEM columns between the two accumulator buffers
constexpr int kMaxSmem = 227 * 1024;

struct TileCoord {
  int n_tile, tx, ty, z0, z1;
};

__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
  TileCoord t;
  t.n_tile = tile % p.n_tiles_n;
  int r = tile / p.n_tiles_n;
  t.tx = r % p.tiles_x;
  r /= p.tiles_x;
  t.ty = r % p.tiles_y;
  r /= p.tiles_y;
  t.z0 = r % p.Z0;
  t.z1 = r / p.Z0;
  return t;
}

template <bool BF16>
__device__ __forceinline__ float cvt16(uint16_t v) {
  if constexpr (BF16) {
    return __bfloat162float(__ushort_as_bfloat16(v));
  } else {
    return __half2float(__ushort_as_half(v));
  }
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
template <bool BF16>
__device__ __forceinline__ void add8(float* v, const uint4& u) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] += cvt16<BF16>((uint16_t)(w[e] & 0xFFFF));
    v[2 * e + 1] += cvt16<BF16>((uint16_t)(w[e] >> 16));
  }
}

// 8 consecutive channels -> one 16-byte store; in the high-precision layout (lo != 0) the rounding residual
// v - float(hi) goes to the lo plane `lo` elements further.
template <bool BF16>
__device__ __forceinline__ void store8_hl(uint16_t* op, long long lo, const float* v) {
  uint4 u;
  u.x = pack16<BF16>(v[0], v[1]);
  u.y = pack16<BF16>(v[2], v[3]);
  u.z = pack16<BF16>(v[4], v[5]);
  u.w = pack16<BF16>(v[6], v[7]);
  *reinterpret_cast<uint4*>(op) = u;
  if (lo) {
    const uint32_t hw[4] = {u.x, u.y, u.z, u.w};
    uint32_t lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      lw[e] = pack16<BF16>(v[2 * e] - cvt16<BF16>((uint16_t)(hw[e] & 0xFFFF)), v[2 * e + 1] - cvt16<BF16>((uint16_t)(hw[e] >> 16)));
    *reinterpret_cast<uint4*>(op + lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// gelu(g) = g * Phi(g), exact-erf form (what diffusers' GEGLU uses), with erf from Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7): one MUFU.RCP, one MUFU.EX2 and a degree-5 Horner instead of erff()'s two-branch
// polynomial — the GEGLU projection is bound by its epilogue (tensor pipe 33 %, ncu r1_final).
//   1 - erf(z) = (a1 t + ... + a5 t^5) e^{-z^2},  t = 1 / (1 + p z),  z = |g| / sqrt(2)
__device__ __forceinline__ float gelu_erf(float g) {
  const float z = fabsf(g) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  q = q * t * e * 0.5f;                                   // = (1 - erf(z)) / 2 = Phi(-|g|)
  return g * (g >= 0.f ? 1.f - q : q);
}

__device__ __forceinline__ void epi_sync() {   // the 128 epilogue threads only
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Staged epilogue (shared by the tap-streaming and the patch-resident main loops): TMEM -> registers
// (bias / residuals / ReLU) -> 16-bit rows in a SWIZZLE_128B shared tile -> one TMA store per
// (warp, 64-channel group), plus the GroupNorm partial sums read back column-wise from the tile.
template <bool BF16>
__device__ __forceinline__ void epilogue_staged(const IgemmParams& p, uint8_t* stg_base, float* sacc, uint64_t* tfull_bar,
                                                uint64_t* tempty_bar, uint64_t* res_bar, uint32_t tmem_base, int warp, int lane) {
  // ===================================================================== epilogue, staged + TMA store
  // TMEM -> registers (bias / residuals / ReLU) -> 16-bit rows in a SWIZZLE_128B shared tile ->
  // one TMA store per (warp, 64-channel group): full-line writes instead of 16-byte pieces at a
  // 2C-byte stride, and image-edge clipping for free.  GroupNorm partial sums are read back
  // column-wise from the staged tile (conflict-free), in a fixed order.
  const int wq = warp;                       // epilogue warps are warps 0..3 (== warp % 4 -> TMEM lanes [32*wq, +32))
  uint8_t* stg = stg_base + wq * 4096;
  const uint32_t stg_addr = smem_u32(stg);
  const uint32_t my_row = stg_addr + lane * 128;
  const bool split = p.out_lo != 0;          // high-precision mode: a second staged tile (+16 KiB) takes the lo plane
  const uint32_t my_row_lo = my_row + 4 * 4096;
  const int sw = lane & 7;
  int acc = 0;
  uint32_t acc_phase = 0, res_phase = 0;
  const bool relu = (p.flags & IG_RELU) != 0;
  const bool geglu = (p.flags & IG_GEGLU) != 0;
  const bool do_stats = p.stats != nullptr;
  const int etid = threadIdx.x;
  int cur_img = -1;
  auto flush_stats = [&](int img) {
    epi_sync();
    float* dst = p.stats + ((long long)img * p.stats_slots + blockIdx.x) * p.Cout * 2;
    for (int i = etid; i < 2 * p.Cout; i += 128) {
      const float tot = (sacc[i] + sacc[2 * p.Cout + i]) + (sacc[4 * p.Cout + i] + sacc[6 * p.Cout + i]);
      dst[i] = tot;
      sacc[i] = 0.f; sacc[2 * p.Cout + i] = 0.f; sacc[4 * p.Cout + i] = 0.f; sacc[6 * p.Cout + i] = 0.f;
    }
    epi_sync();
  };
  if (do_stats) {
    for (int i = etid; i < 8 * p.Cout; i += 128) sacc[i] = 0.f;
    epi_sync();
  }
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const TileCoord t = decode_tile(p, tile);
    const int cls = p.cls_from_z0 ? t.z0 : 0;
    const int n_base = t.n_tile * p.BN;
    bool waited = false;
    if (do_stats) {
      const int img = p.stats_hw ? (t.tx * p.TW) / p.stats_hw : t.z1;
      if (img != cur_img) {
        if (cur_img >= 0) flush_stats(cur_img);
        cur_img = img;
      }
    }
    for (int h = 0; h < p.MT; ++h) {
      const int r0 = h * 128 + wq * 32;                       // first tile row of this warp
      const int row = r0 + lane;
      const int ti = row >> p.tw_shift, tj = row & (p.TW - 1);
      const int gy = t.ty * p.TH + ti, gx = t.tx * p.TW + tj;
      const bool valid = gy < p.gridH && gx < p.gridW;
      const int oy = gy * p.out_sy + p.cls_py[cls], ox = gx * p.out_sx + p.cls_px[cls];
      const long long pix_off = t.z1 * p.out_z1 + (long long)oy * p.out_row_stride + (long long)ox * p.out_pix_stride;
      const int sx = t.tx * p.TW + (r0 & (p.TW - 1)), sy = t.ty * p.TH + (r0 >> p.tw_shift);   // store box origin
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * kAccStride + h * 128;
      for (int c0 = 0; c0 < p.BN; c0 += 64) {
        const int n0 = n_base + c0;
        if (n0 >= p.Cout) break;
        if (geglu) {   // 128 GEMM columns = 4 x [16 values | 16 gates] -> 64 outputs = one staged 128-byte row
          if (c0 & 64) continue;
          if (lane == 0) tma_store_wait_read0();
          __syncwarp();
#pragma unroll
          for (int sub = 0; sub < 4; ++sub) {
            const int ns = n0 + sub * 32;
            float bz[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) bz[q] = 0.f;
            if (p.bias != nullptr) {
#pragma unroll
              for (int q = 0; q < 32; q += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ns + q));
                bz[q] = b4.x; bz[q + 1] = b4.y; bz[q + 2] = b4.z; bz[q + 3] = b4.w;
              }
            }
            if (!waited) {
              mbar_wait(&tfull_bar[acc], acc_phase, 4);
              tc_fence_after();
              waited = true;
            }
            uint32_t r[32];
            tmem_ld_32x32(taddr + c0 + sub * 32, r);
            tmem_ld_wait();
            float g[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float a = __uint_as_float(r[q]) + bz[q], gt = __uint_as_float(r[16 + q]) + bz[16 + q];
              g[q] = valid ? a * gelu_erf(gt) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint32_t a = my_row + (((sub * 2 + i) ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack16<BF16>(g[8 * i], g[8 * i + 1])),
                           "r"(pack16<BF16>(g[8 * i + 2], g[8 * i + 3])), "r"(pack16<BF16>(g[8 * i + 4], g[8 * i + 5])),
                           "r"(pack16<BF16>(g[8 * i + 6], g[8 * i + 7]))
                           : "memory");
            }
          }
          fence_proxy_async_shared();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.tmOut[cls], stg_addr, n0 >> 1, sx, sy, t.z1);
            tma_store_commit();
          }
          continue;
        }
        if (lane == 0) tma_store_wait_read0();                // the previous store has finished reading the tile
        __syncwarp();
        uint4 rt[8];                                          // this thread's residual row (64 channels), res_tma only
        if (p.res_tma) {
          if (lane == 0) {
            mbar_expect_tx(&res_bar[wq], 4096);
            tma_load_4d(stg, &p.tmRes[cls], &res_bar[wq], n0, sx, sy, t.z1);
          }
          mbar_wait(&res_bar[wq], res_phase, 7);
          res_phase ^= 1;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rt[i].x), "=r"(rt[i].y), "=r"(rt[i].z), "=r"(rt[i].w)
                         : "r"(my_row + ((i ^ sw) << 4)));
          __syncwarp();                                       // every row is in registers before the tile is overwritten
        }
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int ns = n0 + sub * 32;
          const long long off = pix_off + ns;
          float bz[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) bz[q] = 0.f;
          if (p.bias != nullptr) {
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ns + q));
              bz[q] = b4.x; bz[q + 1] = b4.y; bz[q + 2] = b4.z; bz[q + 3] = b4.w;
            }
          }
          uint4 r1[4], r2[4];
          const bool has1 = valid && p.res1 != nullptr, has2 = valid && p.res2 != nullptr;
          if (p.res_tma) {
#pragma unroll
            for (int q = 0; q < 4; ++q) r1[q] = rt[sub * 4 + q];
          } else if (has1) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res1) + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) r1[q] = rp[q];
          }
          if (has2) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res2) + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) r2[q] = rp[q];
          }
          if (!waited) {
            mbar_wait(&tfull_bar[acc], acc_phase, 4);
            tc_fence_after();
            waited = true;
          }
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0 + sub * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]) + bz[q];
          if (has1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) add8<BF16>(&v[q * 8], r1[q]);
          }
          if (has2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) add8<BF16>(&v[q * 8], r2[q]);
          }
          if (split) {       // lo planes of the residuals
            if (has1) {
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res1) + off + p.out_lo);
#pragma unroll
              for (int q = 0; q < 4; ++q) add8<BF16>(&v[q * 8], rp[q]);
            }
            if (has2) {
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res2) + off + p.out_lo);
#pragma unroll
              for (int q = 0; q < 4; ++q) add8<BF16>(&v[q * 8], rp[q]);
            }
          }
          if (relu) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
          }
          if (!valid) {      // rows outside the image are clipped by the TMA store; zero them for the statistics
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t a = my_row + (((sub * 4 + i) ^ sw) << 4);
            const uint32_t h0 = pack16<BF16>(v[8 * i], v[8 * i + 1]), h1 = pack16<BF16>(v[8 * i + 2], v[8 * i + 3]);
            const uint32_t h2 = pack16<BF16>(v[8 * i + 4], v[8 * i + 5]), h3 = pack16<BF16>(v[8 * i + 6], v[8 * i + 7]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
            if (split) {     // lo = v - float(hi), rounded to 16 bit
              const uint32_t hw[4] = {h0, h1, h2, h3};
              uint32_t lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                lw[e] = pack16<BF16>(v[8 * i + 2 * e] - cvt16<BF16>((uint16_t)(hw[e] & 0xFFFF)),
                                     v[8 * i + 2 * e + 1] - cvt16<BF16>((uint16_t)(hw[e] >> 16)));
              const uint32_t al = my_row_lo + (((sub * 4 + i) ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(al), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
            }
          }
        }
        fence_proxy_async_shared();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&p.tmOut[cls], stg_addr, n0, sx, sy, t.z1);
          if (split) tma_store_4d(&p.tmOutLo[cls], stg_addr + 4 * 4096, n0, sx, sy, t.z1);
          tma_store_commit();
        }
        if (do_stats) {
          // lane l owns channels n0 + 2l, n0 + 2l + 1: one 32-bit word per staged row
          const uint32_t col = stg_addr + (lane & 3) * 4;
          const int chunk = lane >> 2;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(col + rr * 128 + ((chunk ^ (rr & 7)) << 4)));
            const float a = cvt16<BF16>((uint16_t)(w & 0xFFFF)), b = cvt16<BF16>((uint16_t)(w >> 16));
            s0 += a; q0 += a * a; s1 += b; q1 += b * b;
          }
          float* d = sacc + ((size_t)wq * p.Cout + n0 + 2 * lane) * 2;
          d[0] += s0; d[1] += q0; d[2] += s1; d[3] += q1;
        }
      }
    }
    if (!waited) {
      mbar_wait(&tfull_bar[acc], acc_phase, 4);
      tc_fence_after();
    }
    tc_fence_before();
    mbar_arrive(&tempty_bar[acc]);
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
  if (lane == 0) tma_store_wait_read0();
  if (do_stats && cur_img >= 0) flush_stats(cur_img);
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = kABytes * p.MT;
  const int stage_bytes = a_bytes + p.BN * 128;
  const int stages = p.stages;
  uint8_t* stg_base = smem + stages * stage_bytes;                     // 4 x 4 KiB staging tiles (tma_store only)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + (p.tma_store ? (p.out_lo ? 8 : 4) * 4096 : 0));
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;                                  // [4 epilogue warps] residual tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4);
  float* sacc = reinterpret_cast<float*>(tmem_slot + 4);   // [4 epilogue warps][Cout][2], only with p.stats

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < (p.npass > 1 ? 8 : 4); ++i) tma_prefetch_desc(&p.tmA[i]);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 2);    // producer A + producer B
      mbar_init(&empty_bar[i], p.MT);  // one tcgen05.commit per MMA issuer
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], p.MT);
      mbar_init(&tempty_bar[i], 128);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 6) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Single-thread roles run warp-uniform (every lane walks the loop and waits on the barriers) and one
  // elected lane issues: the TMA coordinates / UMMA descriptors then stay in uniform registers.
  if (warp == 4) {
    // ===================================================================== TMA producer A
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int cls = p.cls_from_z0 ? t.z0 : 0;
      const int a_n = t.z1 * p.a_n_z1 + t.z0 * p.a_n_z0;
      const int a_k0 = t.z0 * p.a_k_z0;
      const int x0 = t.tx * p.TW, y0 = t.ty * p.TH;
      const int ns = p.nseg[cls];
      for (int pass = 0; pass < p.npass; ++pass) {
        const int moff = p.pass_amap[pass];
        for (int s = 0; s < ns; ++s) {
          const IgemmSeg sg = p.seg[cls][s];
          const CUtensorMap* tm = &p.tmA[sg.map + moff];
          const int xs = x0 + sg.dx, ys = y0 + sg.dy;
          for (int c = 0; c < sg.nchunks; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 1);
            if (leader) {
              mbar_expect_tx(&full_bar[stage], (uint32_t)a_bytes);
              tma_load_4d(smem + stage * stage_bytes, tm, &full_bar[stage], a_k0 + c * kBK, xs, ys, a_n);
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 7) {
    // ===================================================================== TMA producer B
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t b_bytes = (uint32_t)p.BN * 128;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int cls = p.cls_from_z0 ? t.z0 : 0;
      const int b_z = t.z1 * p.b_z_z1 + t.z0 * p.b_z_z0;
      const int b_row = t.z0 * p.b_row_z0 + t.n_tile * p.BN;
      const int b_k0 = t.z0 * p.b_k_z0;
      const int nkb = p.nkb[cls];
      for (int pass = 0; pass < p.npass; ++pass) {
        const CUtensorMap* tmb = p.pass_bmap[pass] ? &p.tmB2 : &p.tmB;
        const int bk = b_k0 + p.pass_bk[pass];
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 5);
          if (leader) {
            mbar_expect_tx(&full_bar[stage], b_bytes);
            tma_load_3d(smem + stage * stage_bytes + a_bytes, tmb, &full_bar[stage], bk + kb * kBK, b_row, b_z);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5 || (warp == 6 && p.MT == 2)) {
    const bool leader = elect_one();
    // ===================================================================== MMA issuer(s)
    // With two accumulator tiles (MT = 2, BN <= 128) each tile gets its own issuing thread: a single
    // thread cannot issue one 64-cycle 128x128x16 MMA every 64 cycles once descriptor arithmetic and
    // barrier polls are added (ncu r1h: tensor pipe 57 % busy, issuer never blocked on a barrier).
    const uint32_t idesc = make_idesc_f16(kBM, p.BN, BF16 ? 1 : 0);
    const int h_lo = warp - 5, h_hi = (p.MT == 2) ? warp - 4 : 1;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int cls = p.cls_from_z0 ? t.z0 : 0;
      const int nkb = p.nkb[cls] * p.npass;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase, 3);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
        const uint64_t b_desc = make_sw128_kmajor_desc(a_addr + a_bytes);
        if (leader) {
          for (int h = h_lo; h < h_hi; ++h) {
            const uint64_t a_desc = make_sw128_kmajor_desc(a_addr + h * kABytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // +32 bytes per UMMA_K inside the 128-byte swizzle row -> +2 in the (addr >> 4) field
              umma_f16(d_tmem + h * 128, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp < 4 && p.tma_store) {
    epilogue_staged<BF16>(p, stg_base, sacc, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane);
  } else if (warp < 4) {
    // ===================================================================== epilogue
    const int wq = warp;                     // == warp % 4 -> TMEM lanes [32*wq, 32*wq+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool f32out = (p.flags & IG_OUT_F32_NCHW) != 0;
    const bool relu = (p.flags & IG_RELU) != 0;
    const bool aff = (p.flags & IG_AFFINE_CLAMP01) != 0;
    const bool geglu = (p.flags & IG_GEGLU) != 0;
    const bool do_stats = false;               // statistics are produced by the staged (TMA store) epilogue only
    const int etid = threadIdx.x;              // 0..127 among the epilogue threads
    int cur_img = -1;
    // sum the four warp-private accumulators in a fixed order, publish this CTA's slot, reset
    auto flush_stats = [&](int img) {
      epi_sync();
      float* dst = p.stats + ((long long)img * p.stats_slots + blockIdx.x) * p.Cout * 2;
      for (int i = etid; i < 2 * p.Cout; i += 128) {
        const float tot = (sacc[i] + sacc[2 * p.Cout + i]) + (sacc[4 * p.Cout + i] + sacc[6 * p.Cout + i]);
        dst[i] = tot;
        sacc[i] = 0.f; sacc[2 * p.Cout + i] = 0.f; sacc[4 * p.Cout + i] = 0.f; sacc[6 * p.Cout + i] = 0.f;
      }
      epi_sync();
    };
    if (do_stats) {
      for (int i = etid; i < 8 * p.Cout; i += 128) sacc[i] = 0.f;
      epi_sync();
    }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int cls = p.cls_from_z0 ? t.z0 : 0;
      const int n_base = t.n_tile * p.BN;
      bool waited = false;
      if (do_stats) {
        const int img = p.stats_hw ? (t.tx * p.TW) / p.stats_hw : t.z1;
        if (img != cur_img) {
          if (cur_img >= 0) flush_stats(cur_img);
          cur_img = img;
        }
      }
      for (int h = 0; h < p.MT; ++h) {
        const int row = h * 128 + wq * 32 + lane;
        const int ti = row >> p.tw_shift, tj = row & (p.TW - 1);
        const int gy = t.ty * p.TH + ti, gx = t.tx * p.TW + tj;
        const bool valid = gy < p.gridH && gx < p.gridW;
        const int oy = gy * p.out_sy + p.cls_py[cls], ox = gx * p.out_sx + p.cls_px[cls];
        const long long pix_off = t.z1 * p.out_z1 + t.z0 * p.out_z0 + (long long)oy * p.out_row_stride +
                                  (long long)ox * p.out_pix_stride;
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * kAccStride + h * 128;
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          const int ncols = (p.BN - c0 >= 32) ? 32 : 16;
          const int n0 = n_base + c0;
          const int nvalid = min(ncols, p.Cout - n0);
          const bool live = valid && nvalid > 0;
          const long long off = pix_off + n0;
          const bool vec = !f32out && live && (nvalid == ncols) && ((off & 7) == 0);
          // operands that do not depend on the accumulator are fetched BEFORE waiting on it
          float bz[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) bz[q] = 0.f;
          if (live && p.bias != nullptr) {
            if (nvalid == 32) {
#pragma unroll
              for (int q = 0; q < 32; q += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + q));
                bz[q] = b4.x; bz[q + 1] = b4.y; bz[q + 2] = b4.z; bz[q + 3] = b4.w;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q) if (q < nvalid) bz[q] = __ldg(p.bias + n0 + q);
            }
          }
          uint4 r1[4], r2[4];
          const bool has1 = vec && p.res1 != nullptr, has2 = vec && p.res2 != nullptr;
          if (has1) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res1) + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q * 8 < ncols) r1[q] = rp[q];
          }
          if (has2) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res2) + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q * 8 < ncols) r2[q] = rp[q];
          }
          if (!waited) {
            mbar_wait(&tfull_bar[acc], acc_phase, 4);
            tc_fence_after();
            waited = true;
          }
          uint32_t r[32];
          if (ncols == 32) tmem_ld_32x32(taddr + c0, r); else tmem_ld_32x16(taddr + c0, r);
          tmem_ld_wait();
          if (!live) continue;
          float v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]) + bz[q];
          if (f32out) {
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              if (q < nvalid) {
                float x = v[q];
                if (relu) x = fmaxf(x, 0.f);
                if (aff) x = fminf(fmaxf((x + 1.f) * 0.5f, 0.f), 1.f);
                o[(((long long)t.z1 * p.Cout + (n0 + q)) * p.outH + oy) * p.outW + ox] = x;
              }
            }
            continue;
          }
          if (has1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q * 8 < ncols) add8<BF16>(&v[q * 8], r1[q]);
          } else if (p.res1 != nullptr && live) {
            const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.res1) + off;
#pragma unroll
            for (int q = 0; q < 32; ++q) if (q < nvalid) v[q] += cvt16<BF16>(rp[q]);
          }
          if (has2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q * 8 < ncols) add8<BF16>(&v[q * 8], r2[q]);
          } else if (p.res2 != nullptr && live) {
            const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.res2) + off;
#pragma unroll
            for (int q = 0; q < 32; ++q) if (q < nvalid) v[q] += cvt16<BF16>(rp[q]);
          }
          if (p.out_lo) {      // high-precision layout: lo planes of the residuals
#pragma unroll
            for (int ri = 0; ri < 2; ++ri) {
              const void* rb = ri == 0 ? p.res1 : p.res2;
              if (rb == nullptr || !live) continue;
              const uint16_t* rp = reinterpret_cast<const uint16_t*>(rb) + off + p.out_lo;
              if (vec) {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (q * 8 < ncols) add8<BF16>(&v[q * 8], reinterpret_cast<const uint4*>(rp)[q]);
              } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) if (q < nvalid) v[q] += cvt16<BF16>(rp[q]);
              }
            }
          }
          if (relu) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
          }
          if (geglu) {   // [16 values | 16 gates] -> 16 outputs at column n0/2 (weights are packed interleaved)
            uint16_t* og = reinterpret_cast<uint16_t*>(p.out) + pix_off + (n0 >> 1);
            float g[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) g[q] = v[q] * gelu_erf(v[16 + q]);
#pragma unroll
            for (int q = 0; q < 16; q += 8) store8_hl<BF16>(og + q, p.out_lo, &g[q]);
            continue;
          }
          uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + off;
          if (vec) {
#pragma unroll
            for (int q = 0; q < 32; q += 8) {
              if (q < ncols) store8_hl<BF16>(op + q, p.out_lo, &v[q]);
            }
          } else if (live) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              if (q < nvalid) {
                const uint16_t h = (uint16_t)(pack16<BF16>(v[q], 0.f) & 0xFFFF);
                op[q] = h;
                if (p.out_lo) op[q + p.out_lo] = (uint16_t)(pack16<BF16>(v[q] - cvt16<BF16>(h), 0.f) & 0xFFFF);
              }
            }
          }
        }
      }
      if (!waited) {   // unreachable (BN >= 16), kept so the barrier protocol can never desynchronise
        mbar_wait(&tfull_bar[acc], acc_phase, 4);
        tc_fence_after();
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (do_stats && cur_img >= 0) flush_stats(cur_img);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    __syncwarp();          // reconverge before the .aligned dealloc
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Patch-resident main loop (see IgemmParams::patch).  Same TMEM / epilogue protocol as igemm_kernel;
// only the operand staging differs:
//   shared memory = [2 halo-patch slots][B ring of `stages` weight tiles][staging tiles][barriers]
//   warp 0 lane 0 : one TMA box (64 ch x 130 px x 4 rows) per K chunk into a patch slot
//   warp 3 lane 0 : one weight box per (K chunk, tap) into the B ring
//   warp 1 lane 0 : per K chunk, per tap, per image row h: 4 tcgen05.mma whose A descriptor starts
//                   ((h + dy + 1) * 130 + dx + 1) rows into the patch (absolute-address swizzle makes
//                   any 128-byte row a valid SWIZZLE_128B start: scripts/exp_baseoffset.cu)
template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1) igemm_patch_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.BN * 128;
  const int stages = p.stages;                       // B ring depth
  uint8_t* sB = smem + 2 * p.a_slot_bytes;
  uint8_t* stg_base = sB + stages * b_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(stg_base + 4 * 4096);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + stages;
  uint64_t* tfull_bar = b_empty + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;                                  // [4 epilogue warps] residual tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4);
  float* sacc = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int pw = p.TW + 2;                           // patch width in pixels
  const uint32_t patch_bytes = (uint32_t)(pw * (p.TH + 2) * 128);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmPatch);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2); }        // two MMA issuers
    for (int i = 0; i < stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 2); mbar_init(&tempty_bar[i], 128); }
    for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 6) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================================================================== patch producer
    const bool leader = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int x0 = t.tx * p.TW - 1, y0 = t.ty * p.TH - 1;
      for (int kc = 0; kc < p.kc_count; ++kc) {
        mbar_wait(&a_empty[slot], phase ^ 1, 1);
        if (leader) {
          mbar_expect_tx(&a_full[slot], patch_bytes);
          tma_load_4d(smem + slot * p.a_slot_bytes, &p.tmPatch, &a_full[slot], kc * kBK, x0, y0, t.z1);
        }
        __syncwarp();
        if (++slot == 2) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 7) {
    // ===================================================================== weight producer
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int b_row = t.n_tile * p.BN;
      for (int kc = 0; kc < p.kc_count; ++kc) {
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&b_empty[stage], phase ^ 1, 5);
          if (leader) {
            mbar_expect_tx(&b_full[stage], (uint32_t)b_bytes);
            tma_load_3d(sB + stage * b_bytes, &p.tmB, &b_full[stage], (tap * p.kc_count + kc) * kBK, b_row, 0);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5 || warp == 6) {
    // ===================================================================== MMA issuers (one per image row h)
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(kBM, p.BN, BF16 ? 1 : 0);
    const int h = warp - 5;
    int slot = 0, stage = 0;
    uint32_t a_phase = 0, b_phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = ((p.seg[0][tap].dy + 1) * pw + p.seg[0][tap].dx + 1) * 128;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      for (int kc = 0; kc < p.kc_count; ++kc) {
        mbar_wait(&a_full[slot], a_phase, 3);
        tc_fence_after();
        const uint32_t patch = smem_u32(smem + slot * p.a_slot_bytes);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&b_full[stage], b_phase, 6);
          tc_fence_after();
          const uint64_t b_desc = make_sw128_kmajor_desc(smem_u32(sB + stage * b_bytes));
          if (leader) {
            const uint64_t a_desc = make_sw128_kmajor_desc(patch + tap_off[tap] + h * pw * 128);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_f16(d_tmem + h * 128, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | tap | k) ? 1u : 0u);
            umma_commit(&b_empty[stage]);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; b_phase ^= 1; }
        }
        if (leader) umma_commit(&a_empty[slot]);
        __syncwarp();
        if (++slot == 2) { slot = 0; a_phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp < 4) {
    epilogue_staged<BF16>(p, stg_base, sacc, tfull_bar, tempty_bar, res_bar, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    __syncwarp();          // reconverge before the .aligned dealloc
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(f);
  return fn;
}

int g_num_sms = 0;

}  // namespace

cudaError_t make_tmap_a(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sW,
                        long long sH, long long sN, int TW, int TH, bool bf16) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t make_tmap_b(CUtensorMap* m, const void* base, long long K, long long rows, long long Z,
                        long long sRow, long long sZ, int BN, bool bf16) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)Z};
  cuuint64_t strides[2] = {(cuuint64_t)sRow * 2, (cuuint64_t)sZ * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)BN, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

size_t igemm_smem_bytes(const IgemmParams& p) {
  return (size_t)p.stages * (kABytes * p.MT + p.BN * 128) + (2 * p.stages + 4) * 8 + 16 + 1024;
}

const char* igemm_finalize(IgemmParams* p) {
  if (p->MT == 0) p->MT = 1;
  if (p->npass == 0) p->npass = 1;
  if (p->npass != 1 && p->npass != 3) return "npass must be 1 or 3";
  if (p->npass == 1) { p->pass_amap[0] = 0; p->pass_bk[0] = 0; p->pass_bmap[0] = 0; }
  if (p->out_lo && (p->stats || p->res_tma || p->patch || (p->flags & IG_OUT_F32_NCHW)))
    return "the high-precision output layout excludes epilogue statistics, TMA residuals, the patch loop and fp32 maps";
  if (p->MT != 1 && p->MT != 2) return "MT must be 1 or 2";
  if (p->MT == 2 && p->BN > 128) return "MT=2 needs BN <= 128 (TMEM: 2 buffers x 2 tiles x BN columns)";
  if (p->TW * p->TH != kBM * p->MT) return "TW*TH must be 128*MT";
  if (p->TW > 256 || p->TH > 256) return "TMA box dims are limited to 256";
  if ((1 << p->tw_shift) != p->TW) return "TW must be a power of two";
  if (p->BN < 16 || p->BN > 256 || (p->BN % 16)) return "BN must be a multiple of 16 in [16,256]";
  if (p->Z0 < 1 || p->Z1 < 1) return "bad batch dims";
  p->tiles_x = (p->gridW + p->TW - 1) / p->TW;
  p->tiles_y = (p->gridH + p->TH - 1) / p->TH;
  p->n_tiles_n = (p->Cout + p->BN - 1) / p->BN;
  const int ncls = p->cls_from_z0 ? p->Z0 : 1;
  if (ncls > kMaxClasses) return "too many classes";
  for (int c = 0; c < ncls; ++c) {
    if (p->nseg[c] < 1 || p->nseg[c] > kMaxSegs) return "bad segment count";
    int n = 0;
    for (int s = 0; s < p->nseg[c]; ++s) n += p->seg[c][s].nchunks;
    p->nkb[c] = n;
    if (n < 1) return "empty K loop";
  }
  long long total = (long long)p->n_tiles_n * p->tiles_x * p->tiles_y * p->Z0 * p->Z1;
  if (total > 0x7fffffffLL) return "too many tiles";
  p->total_tiles = (int)total;
  const int stage_bytes = kABytes * p->MT + p->BN * 128;
  const int stats_bytes = p->stats ? 4 * p->Cout * 2 * (int)sizeof(float) : 0;
  if (p->tma_store && ((p->Cout % 64) || (p->BN % 64) || (p->flags & IG_OUT_F32_NCHW) || p->out_z0 != 0))
    return "staged epilogue needs Cout % 64 == 0, BN % 64 == 0, plain 16-bit NHWC output";
  if (p->tma_store && (p->flags & IG_GEGLU) && ((p->Cout % 128) || (p->BN % 128))) return "staged GEGLU needs Cout, BN % 128 == 0";
  if (p->stats && (!p->tma_store || p->Cout > 512 || (p->flags & IG_GEGLU))) return "statistics need the staged epilogue and Cout <= 512";
  int st = (kMaxSmem - 2048 - stats_bytes - (p->tma_store ? (p->out_lo ? 8 : 4) * 4096 : 0)) / stage_bytes;
  if (p->patch) {
    if (!p->tma_store || p->MT != 2 || p->TW != 128 || p->TH != 2 || p->Z0 != 1 || p->nseg[0] != 9 || p->n_tiles_n < 1)
      return "patch mode needs the staged epilogue, TW = 128, MT = 2 and a single-source 3x3 tap table";
    p->a_slot_bytes = ((p->TW + 2) * (p->TH + 2) * 128 + 1023) & ~1023;
    st = (kMaxSmem - 2048 - stats_bytes - 4 * 4096 - 2 * p->a_slot_bytes) / (p->BN * 128);
    if (const char* env = getenv("GP_PATCH_STAGES")) {          // experiment: depth of the weight ring
      const int v = atoi(env);
      if (v >= 2 && v < st) st = v;
    }
  }
  if (st > 8) st = 8;
  if (st < 2) return "tile too large for shared memory";
  p->stages = st;
  return nullptr;
}

static cudaError_t igemm_init();

int igemm_grid(const IgemmParams& p) {
  if (igemm_init() != cudaSuccess) return 0;
  return p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
}

static cudaError_t igemm_init() {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(igemm_patch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(igemm_patch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  return cudaSuccess;
}

cudaError_t igemm_launch(const IgemmParams& p, cudaStream_t stream) {
  cudaError_t ie = igemm_init();
  if (ie != cudaSuccess) return ie;
  if (p.total_tiles <= 0) return cudaSuccess;
  if (p.stats && p.stats_slots < (p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms)) return cudaErrorInvalidValue;
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  // always request the maximum so exactly one CTA (512 TMEM columns) is resident per SM
  const size_t smem = kMaxSmem;
  if (p.patch) {
    if (p.flags & IG_BF16)
      igemm_patch_kernel<true><<<grid, kThreads, smem, stream>>>(p);
    else
      igemm_patch_kernel<false><<<grid, kThreads, smem, stream>>>(p);
  } else if (p.flags & IG_BF16) {
    igemm_kernel<true><<<grid, kThreads, smem, stream>>>(p);
  } else {
    igemm_kernel<false><<<grid, kThreads, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace gp
