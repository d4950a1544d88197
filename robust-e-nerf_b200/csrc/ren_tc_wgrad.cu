// ren_tc_wgrad.cu -- weight gradients of the field MLP on tcgen05 (REN_ENGINE_TC_*).
//
//   dW[n][k] = sum_s Delta[s][n] * H[s][k],   db[n] = sum_s Delta[s][n]
//
// The reduction runs over SAMPLES, so both operands are consumed MN-major straight from the
// fp16 blocked stash / delta matrices (ren_tc.cuh): a 128-sample tile of a matrix is one
// contiguous run in HBM and one bulk-TMA copy into shared memory -- no register staging, no
// transposes.  A job = (layer, 128-row half of n, sample range); its fp32 accumulator
// [128 x H-width] stays in TMEM for the whole sample range and is atomically added into the flat
// gradient buffer once at the end.  The bias gradient is the same MMA against a block of ones.
//   warp 0   producer (2-stage ring of {Delta half tile 32 KB, H tile <= 64 KB})
//   warp 1   MMA issuer (8 K-steps of 16 samples per tile; + the N=16 ones block)
//   warps 2-5 final epilogue (tcgen05.ld -> red.global.add.f32)
// The sigma / radiance head gradients (N = 1..3) are a bandwidth-bound SIMT pass.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "ren_tc.cuh"

#define WG_THREADS 192
#define WG_STAGES 2
#define WG_MAX_MATS 48
#define SMW_D_BYTES 32768
#define SMW_H_BYTES 65536
#define SMW_STAGE (SMW_D_BYTES + SMW_H_BYTES)   // 98304
#define SMW_ONES (WG_STAGES * SMW_STAGE)        // 196608
#define SMW_BAR (SMW_ONES + 4096)
#define SMW_TMEM_PTR (SMW_BAR + 64)
#define SMW_EMPTY (SMW_BAR + 96)               // the job in the accumulator saw no live tile
#define SMW_TOTAL (SMW_BAR + 128)
#define WG_TMEM_ONES 256

struct WgMat {
  long long d_off;   // bytes into the delta buffer: fp16 blocked [Sp x Nd]
  long long h_off;   // bytes into the stash buffer: fp16 blocked [Sp x Hw]
  int Nd, n_half, Hw;
  int w_off, ldw, col0, Ktrue;   // dW[n][col0 + k], k < Ktrue
  int b_off;                     // bias gradient offset or -1
  float h_scale;                 // the stash keeps hidden activations in the log2 domain (ren_tc.cu): ln2/100, else 1
};
struct WgArgs {
  WgMat mats[WG_MAX_MATS];
  int n_mats, splits;
  long long n_tiles;
  const unsigned char* stash;
  const unsigned char* delta;
  const unsigned int* gmax;    // deltas carry the power-of-two scale delta_scale(gmax)
  float* g_params;
  const unsigned char* tile_live;   // per-tile flags or NULL (occupancy-grid marching): dead tiles hold no data
};

extern "C" __global__ void __launch_bounds__(WG_THREADS, 1) ren_tc_wgrad(const WgArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + SMW_BAR;          // [2]
  const uint32_t bar_empty = sbase + SMW_BAR + 16;    // [2]
  const uint32_t bar_accfull = sbase + SMW_BAR + 32;
  const uint32_t bar_accempty = sbase + SMW_BAR + 40;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SMW_TMEM_PTR);
  const int n_jobs = a.n_mats * a.splits;
  const long long tiles_per = (a.n_tiles + a.splits - 1) / a.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accfull, 1);
    mbar_init(bar_accempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // block of fp16 ones: the bias-gradient operand
  for (int i = threadIdx.x; i < 4096 / 4; i += WG_THREADS)
    reinterpret_cast<uint32_t*>(smem + SMW_ONES)[i] = 0x3C003C00u;
  fence_proxy_async();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + SMW_TMEM_PTR) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (REN_ONE_LANE()) {
      uint32_t it = 0;
      for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const WgMat& mt = a.mats[job / a.splits];
        const int split = job % a.splits;
        const long long t0 = split * tiles_per, t1 = min(a.n_tiles, t0 + tiles_per);
        const uint32_t hbytes = (uint32_t)mt.Hw * 256u;
        for (long long t = t0; t < t1; ++t) {
          if (a.tile_live != nullptr && !__ldg(a.tile_live + t)) continue;
          const uint32_t s = it % WG_STAGES, ph = (it / WG_STAGES) & 1u;
          ++it;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(bar_full + 8 * s, SMW_D_BYTES + hbytes);
          const unsigned char* dsrc = a.delta + mt.d_off + (size_t)t * mt.Nd * 256 + (size_t)mt.n_half * SMW_D_BYTES;
          const unsigned char* hsrc = a.stash + mt.h_off + (size_t)t * hbytes;
          bulk_g2s(sbase + s * SMW_STAGE, dsrc, SMW_D_BYTES, bar_full + 8 * s);
          bulk_g2s(sbase + s * SMW_STAGE + SMW_D_BYTES, hsrc, hbytes, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (REN_ONE_LANE()) {
      uint32_t it = 0, jn = 0;
      const uint32_t idesc_ones = umma_idesc(16, 1, 1);
      for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const WgMat& mt = a.mats[job / a.splits];
        const int split = job % a.splits;
        const long long t0 = split * tiles_per, t1 = min(a.n_tiles, t0 + tiles_per);
        if (t0 >= t1) continue;
        mbar_wait(bar_accempty, (jn & 1u) ^ 1u);     // previous job's accumulator has been drained
        tc_fence_after();
        const uint32_t idesc = umma_idesc(mt.Hw, 1, 1);
        uint32_t first = 1;
        for (long long t = t0; t < t1; ++t) {
          if (a.tile_live != nullptr && !__ldg(a.tile_live + t)) continue;
          const uint32_t s = it % WG_STAGES, ph = (it / WG_STAGES) & 1u;
          ++it;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t db = sbase + s * SMW_STAGE, hb = db + SMW_D_BYTES;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            // MN-major operands: LBO = 128 B between 8-sample K blocks, SBO = 2 KB between 8-wide groups
            const uint64_t da = umma_desc(db + ks * 256, 128, 2048);
            const uint64_t dh = umma_desc(hb + ks * 256, 128, 2048);
            const uint64_t d1 = umma_desc(sbase + SMW_ONES + ks * 256, 128, 2048);
            tc_mma_f16(tmem_base, da, dh, idesc, first ? 0u : 1u);
            tc_mma_f16(tmem_base + WG_TMEM_ONES, da, d1, idesc_ones, first ? 0u : 1u);
            first = 0;
          }
          tc_commit(bar_empty + 8 * s);
        }
        *reinterpret_cast<volatile uint32_t*>(smem + SMW_EMPTY) = first;   // nothing accumulated: skip the epilogue
        __threadfence_block();
        tc_commit(bar_accfull);
        ++jn;
      }
    }
  } else {
    const int q = warp & 3;
    const int nl = q * 32 + lane;                       // accumulator lane == local output row n
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t jn = 0;
    const float unscale = 1.f / delta_scale(a.gmax);
    for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
      const WgMat& mt = a.mats[job / a.splits];
      const int split = job % a.splits;
      const long long t0 = split * tiles_per, t1 = min(a.n_tiles, t0 + tiles_per);
      if (t0 >= t1) continue;
      mbar_wait(bar_accfull, jn & 1u);
      tc_fence_after();
      const bool empty_job = *reinterpret_cast<volatile uint32_t*>(smem + SMW_EMPTY) != 0u;
      const int n = mt.n_half * 128 + nl;
      float* row = a.g_params + mt.w_off + (long long)n * mt.ldw + mt.col0;
      for (int cb = 0; cb < (empty_job ? 0 : mt.Hw); cb += 16) {
        uint32_t r[16];
        tc_ld16(tmem_lane + (uint32_t)cb, r);
        tc_ld_wait();
        if (n < mt.Nd) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cb + j < mt.Ktrue) atomicAdd(row + cb + j, __uint_as_float(r[j]) * (unscale * mt.h_scale));
        }
      }
      if (mt.b_off >= 0 && !empty_job) {
        uint32_t r[16];
        tc_ld16(tmem_lane + WG_TMEM_ONES, r);
        tc_ld_wait();
        if (n < mt.Nd) atomicAdd(a.g_params + mt.b_off + n, __uint_as_float(r[0]) * unscale);
      }
      tc_fence_before();
      mbar_arrive(bar_accempty);
      ++jn;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ v2: one job per (matrix, sample range)
// Both 128-row halves of a layer's Delta against ONE fetch of the activation tile (the v1 jobs fetched it once
// per half), 64-sample stages so that a 3-deep ring of {Delta 32 KB, H 32 KB} fits, the bias gradient as column
// sums of the Delta stage taken by four SIMT warps (no room for the ones-block accumulator next to 2 x 256
// columns), and a host-built longest-first job list per CTA instead of a round-robin of unequal jobs.
//   warp 0      producer: two TMA tensor copies per stage (the 64-sample half of all 8-feature groups of the
//               Delta tile and of the activation tile)
//   warp 1      MMA issuer: 4 K-steps of 16 samples per stage, one MMA per Delta half
//   warps 2-5   final epilogue (tcgen05.ld -> red.global.add.f32), one per TMEM lane quadrant
//   warps 6-9   bias gradient: column sums of the Delta stage from shared memory
#define W2_THREADS 320
#define W2_STAGES 3
#define W2_D_BYTES 32768                      // 256 features x 64 samples x 2 B
#define W2_H_BYTES 32768
#define W2_STAGE (W2_D_BYTES + W2_H_BYTES)
#define W2_BAR (W2_STAGES * W2_STAGE)         // 196608
#define W2_TMEM_PTR (W2_BAR + 96)
#define W2_EMPTY (W2_BAR + 104)
#define W2_TOTAL (W2_BAR + 128)
#define W2_MAX_MATS 24
#define W2_MAX_JOBS 1024

struct W2Mat {
  long long d_off, h_off;        // bytes into delta / stash: fp16 blocked [Sp x Nd], [Sp x Hw]
  int Nd, Hw;                    // Nd: 128 or 256 (1 or 2 halves)
  int w_off, ldw, col0, Ktrue;
  int b_off;
  float h_scale;
};
struct W2Job { unsigned short mat, pad; int t0, t1; };
struct W2Args {
  // 3-D tensor maps of the blocked fp16 matrices ([tile][8-feature group][128 samples x 8 features = 2 KB]): one
  // TMA tensor copy brings the 64-sample half (1 KB) of every group of a tile, densely packed, into a stage
  alignas(64) CUtensorMap tm_d[W2_MAX_MATS];
  alignas(64) CUtensorMap tm_h[W2_MAX_MATS];
  W2Mat mats[W2_MAX_MATS];
  W2Job jobs[W2_MAX_JOBS];       // grouped by CTA
  unsigned short cta_first[160]; // jobs of CTA b: [cta_first[b], cta_first[b + 1])
  int n_mats;
  int dbg;                       // timing experiments (env REN_W2_DBG): 1 no bias reads, 2 no MMAs, 4 no drain atomics
  const unsigned char* stash;
  const unsigned char* delta;
  const unsigned int* gmax;
  float* g_params;
  const unsigned char* tile_live;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, uint64_t tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

extern "C" __global__ void __launch_bounds__(W2_THREADS, 1) ren_tc_wgrad2(const __grid_constant__ W2Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + W2_BAR;           // [3]
  const uint32_t bar_empty = sbase + W2_BAR + 32;     // [3]
  const uint32_t bar_accfull = sbase + W2_BAR + 64;
  const uint32_t bar_accempty = sbase + W2_BAR + 72;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + W2_TMEM_PTR);
  const int j0 = a.cta_first[blockIdx.x], j1 = a.cta_first[blockIdx.x + 1];

  if (threadIdx.x == 0) {
    for (int s = 0; s < W2_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 5);               // tcgen05.commit + the four bias warps
    }
    mbar_init(bar_accfull, 1);
    mbar_init(bar_accempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + W2_TMEM_PTR) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== producer =====================
    if (REN_ONE_LANE()) {
      uint32_t it = 0;
      for (int j = j0; j < j1; ++j) {
        const W2Job jb = a.jobs[j];
        const W2Mat& mt = a.mats[jb.mat];
        const uint32_t bytes = (uint32_t)(mt.Nd + mt.Hw) * 128u;
        const uint64_t tmd = reinterpret_cast<uint64_t>(&a.tm_d[jb.mat]);
        const uint64_t tmh = reinterpret_cast<uint64_t>(&a.tm_h[jb.mat]);
        for (int t = jb.t0; t < jb.t1; ++t) {
          if (a.tile_live != nullptr && !__ldg(a.tile_live + t)) continue;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh, ++it) {
            const uint32_t s = it % W2_STAGES, ph = (it / W2_STAGES) & 1u;
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
            const uint32_t dst = sbase + s * W2_STAGE;
            tma_load_3d(dst, tmd, hh * 128, 0, t, bar_full + 8 * s);
            tma_load_3d(dst + W2_D_BYTES, tmh, hh * 128, 0, t, bar_full + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (REN_ONE_LANE()) {
      uint32_t it = 0, jn = 0;
      for (int j = j0; j < j1; ++j) {
        const W2Job jb = a.jobs[j];
        const W2Mat& mt = a.mats[jb.mat];
        mbar_wait(bar_accempty, (jn & 1u) ^ 1u);     // previous job's accumulators have been drained
        tc_fence_after();
        const uint32_t idesc = umma_idesc(mt.Hw, 1, 1);
        const int nh = mt.Nd >> 7;
        uint32_t first = 1;
        for (int t = jb.t0; t < jb.t1; ++t) {
          if (a.tile_live != nullptr && !__ldg(a.tile_live + t)) continue;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh, ++it) {
            const uint32_t s = it % W2_STAGES, ph = (it / W2_STAGES) & 1u;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t db = sbase + s * W2_STAGE, hb = db + W2_D_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (a.dbg & 2) break;
              // MN-major operands: LBO = 128 B between 8-sample K blocks, SBO = 1 KB between 8-wide groups
              const uint64_t dh = umma_desc(hb + ks * 256, 128, 1024);
              tc_mma_f16(tmem_base, umma_desc(db + ks * 256, 128, 1024), dh, idesc, first ? 0u : 1u);
              if (nh > 1)
                tc_mma_f16(tmem_base + 256, umma_desc(db + 16 * 1024 + ks * 256, 128, 1024), dh, idesc, first ? 0u : 1u);
              first = 0;
            }
            tc_commit(bar_empty + 8 * s);
          }
        }
        *reinterpret_cast<volatile uint32_t*>(smem + W2_EMPTY) = first;   // nothing accumulated: skip the drain
        __threadfence_block();
        tc_commit(bar_accfull);
        ++jn;
      }
    }
  } else if (warp < 6) {
    // ===================== accumulator drain =====================
    const int q = warp & 3;
    const int nl = q * 32 + lane;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t jn = 0;
    const float unscale = 1.f / delta_scale(a.gmax);
    for (int j = j0; j < j1; ++j) {
      const W2Job jb = a.jobs[j];
      const W2Mat& mt = a.mats[jb.mat];
      mbar_wait(bar_accfull, jn & 1u);
      tc_fence_after();
      const bool empty_job = *reinterpret_cast<volatile uint32_t*>(smem + W2_EMPTY) != 0u;
      const int nh = mt.Nd >> 7;
      const float sc = unscale * mt.h_scale;
      if (!empty_job && !(a.dbg & 4)) {
        for (int h = 0; h < nh; ++h) {
          float* row = a.g_params + mt.w_off + (long long)(h * 128 + nl) * mt.ldw + mt.col0;
          for (int cb = 0; cb < mt.Hw; cb += 16) {
            uint32_t r[16];
            tc_ld16(tmem_lane + (uint32_t)(h * 256 + cb), r);
            tc_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (cb + jj < mt.Ktrue) atomicAdd(row + cb + jj, __uint_as_float(r[jj]) * sc);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_accempty);
      ++jn;
    }
  } else {
    // ===================== bias gradient: column sums of the Delta stage =====================
    const int bt = threadIdx.x - 192;                // 0..127: feature group bt >> 2, feature pair bt & 3
    const int g = bt >> 2, fp = bt & 3;
    uint32_t it = 0;
    const float unscale = 1.f / delta_scale(a.gmax);
    for (int j = j0; j < j1; ++j) {
      const W2Job jb = a.jobs[j];
      const W2Mat& mt = a.mats[jb.mat];
      const bool want = mt.b_off >= 0 && g < (mt.Nd >> 3) && !(a.dbg & 1);
      float s0 = 0.f, s1 = 0.f;
      for (int t = jb.t0; t < jb.t1; ++t) {
        if (a.tile_live != nullptr && !__ldg(a.tile_live + t)) continue;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh, ++it) {
          const uint32_t s = it % W2_STAGES, ph = (it / W2_STAGES) & 1u;
          mbar_wait(bar_full + 8 * s, ph);
          if (want) {
            const unsigned char* base = smem + s * W2_STAGE + g * 1024 + fp * 4;
#pragma unroll 8
            for (int i = 0; i < 64; ++i) {
              const int sm = (i + g) & 63;           // stagger the rows per group: conflict-free banks
              const float2 v = __half22float2(*reinterpret_cast<const __half2*>(base + sm * 16));
              s0 += v.x;
              s1 += v.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        }
      }
      if (want) {
        atomicAdd(a.g_params + mt.b_off + g * 8 + fp * 2, s0 * unscale);
        atomicAdd(a.g_params + mt.b_off + g * 8 + fp * 2 + 1, s1 * unscale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ head gradients (SIMT)
// which 0: d w_sigma[k] = sum_s g_sigma[s] a_last[s][k] (N = 256), d b_sigma
// which 1: d W_o[c][j]  = sum_s g_c[s][c] g[s][j]        (N = 128), d b_o[c]
struct HeadArgs {
  const unsigned char* h[2];   // fp16 blocked activation matrices
  int N[2];
  int w_off[2], b_off[2], n_ch[2], g_col[2];
  const float* graw;           // [S][C1]
  int C1;
  long long S, n_tiles;
  float* g_params;
  float h_scale;               // log2-domain stash -> true activations
  const unsigned char* tile_live;
};

extern "C" __global__ void __launch_bounds__(256) ren_tc_head_wgrad(const HeadArgs a) {
  // selected with ?: -- a runtime index into the parameter arrays would move them to local memory
  const bool w1 = blockIdx.y != 0;
  const int N = w1 ? a.N[1] : a.N[0], groups = N >> 3, nch = w1 ? a.n_ch[1] : a.n_ch[0];
  const unsigned char* hmat = w1 ? a.h[1] : a.h[0];
  const int g_col = w1 ? a.g_col[1] : a.g_col[0];
  const int w_off = w1 ? a.w_off[1] : a.w_off[0], b_off = w1 ? a.b_off[1] : a.b_off[0];
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int fg = (int)(warp % groups);                  // this warp's 8-feature group
  const long long tstride = n_warps / groups;
  if (tstride == 0 || warp >= tstride * groups) return;
  float acc[3][8];
  float bsum[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  for (long long t = warp / groups; t < a.n_tiles; t += tstride) {
    if (a.tile_live != nullptr && !__ldg(a.tile_live + t)) continue;
    const uint4* blk = reinterpret_cast<const uint4*>(hmat) + ((size_t)t * groups + fg) * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int sl = lane * 4 + i;
      const long long s = t * 128 + sl;
      if (s >= a.S) continue;
      float f[8];
      unpack8(__ldg(blk + sl), f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {                     // static indices keep acc[][] in registers
        if (c < nch) {
          const float g = __ldg(a.graw + s * a.C1 + g_col + c);
          bsum[c] += g;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c][j] = fmaf(g, f[j], acc[c][j]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c >= nch) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[c][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(a.g_params + w_off + c * N + fg * 8 + j, v * a.h_scale);
    }
    float b = bsum[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0 && fg == 0 && b_off >= 0) atomicAdd(a.g_params + b_off + c, b);
  }
}

// ------------------------------------------------------------------ host side
// The head gradients (two skinny reductions over the stash, no shared memory) run on a side stream next to the
// HBM-bound main kernel: fork / join with events on the caller's stream, one side stream per device.
struct WgSide { cudaStream_t s; cudaEvent_t fork, join; bool ok; };
static WgSide* wg_side(int dev) {
  static WgSide side[64];
  static std::mutex mu;
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  WgSide& w = side[dev];
  if (!w.ok) {
    if (cudaStreamCreateWithFlags(&w.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&w.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&w.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    w.ok = true;
  }
  return &w;
}

// v2 job list: every matrix is cut into sample ranges in proportion to the bytes one tile of it moves, then the
// jobs go longest-first to the least-loaded CTA (cost = tiles x bytes per tile + the accumulator drain).
static int launch_wgrad2(const RenPlan& p, long long S, const void* stash, const void* delta, const unsigned int* gmax,
                         float* g_params, int with_bias, const unsigned char* tile_live, int sms, cudaStream_t st) {
  static W2Args a;                         // ~20 KB: too large for the stack of a deep call chain, reused per call
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  memset(&a, 0, sizeof(a));
  static_assert(sizeof(W2Args) <= 32000, "kernel parameter block too large");
  int n = 0;
  const float kap = p.f.hidden_activation == REN_HIDDEN_SOFTPLUS100 ? 0.0069314718055994531f : 1.f;
  auto add = [&](int dmat, int Nd, long long h_off, int Hw, int w_off, int ldw, int col0, int Ktrue, int b_off,
                 float h_scale) {
    W2Mat& m = a.mats[n++];
    m.d_off = (long long)tcs_mat(p, S, dmat); m.Nd = Nd; m.h_off = h_off; m.Hw = Hw;
    m.w_off = w_off; m.ldw = ldw; m.col0 = col0; m.Ktrue = Ktrue; m.b_off = with_bias ? b_off : -1;
    m.h_scale = h_scale;
  };
  auto add_layer = [&](const RenLayer& L, int prev_mat, int prev_w, float prev_scale) {
    const int ldw = L.K0 + L.K1;
    if (L.src0 == 1) add(L.mat, L.N, (long long)tcs_encp(p, S), REN_PE_MAX, L.w_off, ldw, 0, L.K0, L.b_off, 1.f);
    else add(L.mat, L.N, (long long)tcs_mat(p, S, prev_mat), prev_w, L.w_off, ldw, 0, L.K0, L.b_off, prev_scale);
    if (L.K1 > 0) {
      if (L.src1 == 1) add(L.mat, L.N, (long long)tcs_encp(p, S), REN_PE_MAX, L.w_off, ldw, L.K0, L.K1, -1, 1.f);
      else add(L.mat, L.N, (long long)tcs_encd(p, S), REN_DE_MAX, L.w_off, ldw, L.K0, L.K1, -1, 1.f);
    }
  };
  int prev = -1, prevw = 0;
  for (int i = 0; i < p.D; ++i) {
    add_layer(p.trunk[i], prev, prevw, kap);
    prev = p.trunk[i].mat; prevw = p.trunk[i].N;
  }
  add_layer(p.bott, prev, prevw, kap);
  add_layer(p.cond[0], p.bott.mat, p.bott.N, 1.f);
  if (n > W2_MAX_MATS) { ren_set_error("too many wgrad matrices"); return REN_E_INVALID; }
  a.n_mats = n;
  const long long n_tiles = (S + TC_M - 1) / TC_M;
  // splits per matrix
  double bytes[W2_MAX_MATS], total = 0;
  for (int m = 0; m < n; ++m) { bytes[m] = (double)(a.mats[m].Nd + a.mats[m].Hw) * 256.0; total += bytes[m]; }
  const int target = sms * 3;
  struct J { int mat; long long t0, t1; double cost; };
  static J js[W2_MAX_JOBS];
  int nj = 0;
  for (int m = 0; m < n; ++m) {
    long long sp = (long long)(target * bytes[m] / total + 0.5);
    if (sp < 1) sp = 1;
    if (sp > n_tiles) sp = n_tiles;
    const long long per = (n_tiles + sp - 1) / sp;
    for (long long t0 = 0; t0 < n_tiles; t0 += per) {
      if (nj >= W2_MAX_JOBS) { ren_set_error("too many wgrad jobs"); return REN_E_INVALID; }
      const long long t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
      // drain: Nd x Hw atomics, counted like ~6 bytes of stream each
      js[nj++] = {m, t0, t1, (double)(t1 - t0) * bytes[m] + 6.0 * a.mats[m].Nd * a.mats[m].Hw};
    }
  }
  // longest first (insertion sort: a few hundred jobs), then least-loaded CTA
  for (int i = 1; i < nj; ++i) {
    J x = js[i];
    int k = i - 1;
    while (k >= 0 && js[k].cost < x.cost) { js[k + 1] = js[k]; --k; }
    js[k + 1] = x;
  }
  const int grid = nj < sms ? nj : sms;
  static double load[160];
  static int owner[W2_MAX_JOBS], count[160];
  for (int b = 0; b < grid; ++b) { load[b] = 0; count[b] = 0; }
  for (int i = 0; i < nj; ++i) {
    int best = 0;
    for (int b = 1; b < grid; ++b) if (load[b] < load[best]) best = b;
    owner[i] = best; load[best] += js[i].cost; ++count[best];
  }
  int first = 0;
  for (int b = 0; b < grid; ++b) { a.cta_first[b] = (unsigned short)first; first += count[b]; }
  a.cta_first[grid] = (unsigned short)first;
  static int fill[160];
  for (int b = 0; b < grid; ++b) fill[b] = 0;
  for (int i = 0; i < nj; ++i) {
    W2Job& o = a.jobs[a.cta_first[owner[i]] + fill[owner[i]]++];
    o.mat = (unsigned short)js[i].mat; o.pad = 0; o.t0 = (int)js[i].t0; o.t1 = (int)js[i].t1;
  }
  a.stash = (const unsigned char*)stash; a.delta = (const unsigned char*)delta; a.g_params = g_params;
  a.gmax = gmax; a.tile_live = tile_live;
  if (const char* dbg = getenv("REN_W2_DBG")) a.dbg = atoi(dbg);
  // the driver entry point is looked up at run time: the library must load on hosts without a driver (libcuda)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
      ren_set_error("tc wgrad: cuTensorMapEncodeTiled is not available");
      return REN_E_CUDA;
    }
    encode = (EncodeFn)fn;
  }
  for (int m = 0; m < n; ++m) {
    for (int which = 0; which < 2; ++which) {
      const int N = which ? a.mats[m].Hw : a.mats[m].Nd;
      void* base = which ? (void*)((unsigned char*)stash + a.mats[m].h_off) : (void*)((unsigned char*)delta + a.mats[m].d_off);
      // viewed as 8-byte elements so that a box row is the whole contiguous 1 KB run of a group's 64-sample half
      // (with the natural fp16 view the inner dimension is 16 B and the copy degenerates into 16-byte requests:
      // measured 55 % L2 hit rate = every sector asked for twice, 74 % L2 throughput, 3.1 ms)
      const cuuint64_t dims[3] = {256, (cuuint64_t)(N >> 3), (cuuint64_t)n_tiles};
      const cuuint64_t strides[2] = {2048, (cuuint64_t)N * 256};
      const cuuint32_t box[3] = {128, (cuuint32_t)(N >> 3), 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      const CUresult cr = encode(which ? &a.tm_h[m] : &a.tm_d[m], CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, base,
                                                 dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) { ren_set_error("tc wgrad: cuTensorMapEncodeTiled failed (%d)", (int)cr); return REN_E_CUDA; }
    }
  }
  cudaError_t e = cudaFuncSetAttribute(ren_tc_wgrad2, cudaFuncAttributeMaxDynamicSharedMemorySize, W2_TOTAL);
  if (e != cudaSuccess) { ren_set_error("tc wgrad smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  ren_prof_begin(REN_K_WGRAD, st);
  ren_tc_wgrad2<<<grid, W2_THREADS, W2_TOTAL, st>>>(a);
  ren_prof_end(REN_K_WGRAD, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc wgrad launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

// with_bias = 0: the tangent rows of the tangent render (d/dw activations carry no bias)
int ren_tc_launch_wgrad(const RenPlan& p, long long S, const void* stash, const void* delta,
                        const float* graw, const unsigned int* gmax, float* g_params, int with_bias,
                        const unsigned char* tile_live, cudaStream_t st) {
  WgArgs a;
  static_assert(sizeof(WgArgs) <= 4000, "kernel parameter block too large");
  memset(&a, 0, sizeof(a));
  int n = 0;
  const float kap = p.f.hidden_activation == REN_HIDDEN_SOFTPLUS100 ? 0.0069314718055994531f : 1.f;
  auto add = [&](int dmat, int Nd, long long h_off, int Hw, int w_off, int ldw, int col0, int Ktrue, int b_off,
                 float h_scale) {
    for (int h = 0; h < Nd / 128; ++h) {
      WgMat& m = a.mats[n++];
      m.d_off = (long long)tcs_mat(p, S, dmat); m.Nd = Nd; m.n_half = h; m.h_off = h_off; m.Hw = Hw;
      m.w_off = w_off; m.ldw = ldw; m.col0 = col0; m.Ktrue = Ktrue; m.b_off = with_bias ? b_off : -1;
      m.h_scale = h_scale;
    }
  };
  // prev_scale: scale of the stashed main input (hidden activations are kept in the log2 domain)
  auto add_layer = [&](const RenLayer& L, int prev_mat, int prev_w, float prev_scale) {
    const int ldw = L.K0 + L.K1;
    if (L.src0 == 1) add(L.mat, L.N, (long long)tcs_encp(p, S), REN_PE_MAX, L.w_off, ldw, 0, L.K0, L.b_off, 1.f);
    else add(L.mat, L.N, (long long)tcs_mat(p, S, prev_mat), prev_w, L.w_off, ldw, 0, L.K0, L.b_off, prev_scale);
    if (L.K1 > 0) {
      if (L.src1 == 1) add(L.mat, L.N, (long long)tcs_encp(p, S), REN_PE_MAX, L.w_off, ldw, L.K0, L.K1, -1, 1.f);
      else add(L.mat, L.N, (long long)tcs_encd(p, S), REN_DE_MAX, L.w_off, ldw, L.K0, L.K1, -1, 1.f);
    }
  };
  int prev = -1, prevw = 0;
  for (int i = 0; i < p.D; ++i) {
    add_layer(p.trunk[i], prev, prevw, kap);
    prev = p.trunk[i].mat; prevw = p.trunk[i].N;
  }
  const int last_trunk = prev;
  add_layer(p.bott, prev, prevw, kap);
  add_layer(p.cond[0], p.bott.mat, p.bott.N, 1.f);
  if (n > WG_MAX_MATS) { ren_set_error("too many wgrad matrices"); return REN_E_INVALID; }
  a.n_mats = n;
  a.n_tiles = (S + TC_M - 1) / TC_M;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int splits = (2 * sms + n - 1) / n;
  if (splits > a.n_tiles) splits = (int)a.n_tiles;
  if (splits < 1) splits = 1;
  a.splits = splits;
  a.stash = (const unsigned char*)stash; a.delta = (const unsigned char*)delta; a.g_params = g_params;
  a.gmax = gmax; a.tile_live = tile_live;
  const int n_jobs = n * splits;
  const int grid = n_jobs < sms ? n_jobs : sms;
  cudaError_t e = cudaFuncSetAttribute(ren_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, SMW_TOTAL);
  if (e != cudaSuccess) { ren_set_error("tc wgrad smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  HeadArgs h;
  memset(&h, 0, sizeof(h));
  h.h[0] = (const unsigned char*)stash + tcs_mat(p, S, last_trunk); h.N[0] = p.W;
  h.w_off[0] = p.sig_w; h.b_off[0] = with_bias ? p.sig_b : -1; h.n_ch[0] = 1; h.g_col[0] = 0;
  h.h[1] = (const unsigned char*)stash + tcs_mat(p, S, p.cond[0].mat); h.N[1] = p.WC;
  h.w_off[1] = p.rgb_w; h.b_off[1] = with_bias ? p.rgb_b : -1; h.n_ch[1] = p.C; h.g_col[1] = 1;
  h.graw = graw; h.C1 = 1 + p.C; h.S = S; h.n_tiles = a.n_tiles; h.g_params = g_params; h.h_scale = kap;
  h.tile_live = tile_live;
  WgSide* side = wg_side(dev);
  if (side == nullptr) { ren_set_error("tc wgrad: side stream: %s", cudaGetErrorString(cudaGetLastError())); return REN_E_CUDA; }
  cudaEventRecord(side->fork, st);
  cudaStreamWaitEvent(side->s, side->fork, 0);
  ren_prof_count();
  ren_tc_head_wgrad<<<dim3(sms * 2, 2), 256, 0, side->s>>>(h);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc head wgrad launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  cudaEventRecord(side->join, side->s);
  if (getenv("REN_WGRAD_V1") == nullptr) {
    const int rc2 = launch_wgrad2(p, S, stash, delta, gmax, g_params, with_bias, tile_live, sms, st);
    if (rc2) return rc2;
  } else {
    ren_prof_begin(REN_K_WGRAD, st);
    ren_tc_wgrad<<<grid, WG_THREADS, SMW_TOTAL, st>>>(a);
    ren_prof_end(REN_K_WGRAD, st);
    e = cudaGetLastError();
    if (e != cudaSuccess) { ren_set_error("tc wgrad launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  }

  cudaStreamWaitEvent(st, side->join, 0);
  return REN_OK;
}
