// ren_tc.cu -- tcgen05 "chain" kernel of the field MLP: forward and dgrad (REN_ENGINE_TC_*).
//
// One persistent CTA per SM, one 128-sample tile in flight.  Warp-specialised roles:
//   warp 0       weight producer: streams pre-packed fp16 (hi|lo) weight K-chunks from L2 into a
//                3-stage shared-memory ring with 1-D bulk TMA (cp.async.bulk + mbarrier tx bytes)
//   warp 1       MMA issuer: one thread issues tcgen05.mma (kind::f16, M=128, N=256/128/16, K=16);
//                accumulators live in TMEM (512 columns x 128 lanes, fp32)
//   warps 2-17   prologue + epilogue: tcgen05.ld -> bias/activation (fwd) or activation derivative
//                (dgrad) -> fp16 hi/lo split -> the next GEMM's A operand in shared memory (+ the
//                fp16 stash in HBM).  Lane quadrant = warp % 4 (TMEM access rule), 4 warps share
//                a quadrant and split the columns.
// The whole network is a chain of GEMM "steps" described by a table, so forward and dgrad are
// the same kernel: activations / deltas never leave the SM between layers.  The sigma and
// radiance heads (N = 1..3) are N=16 MMAs, and in dgrad the head gradients re-enter as K=16
// rank-1 segments -- there is no per-element head arithmetic.
//
// fp32-class accuracy from fp16 tensor cores: x = hi + lo, hi = fp16(x), lo = fp16(x - hi);
// A*W ~= A_hi*W_hi + A_lo*W_hi + A_hi*W_lo in fp32 (3 MMAs per K step; the dropped lo*lo term is
// 2^-22 relative).  Single-pass fp16/bf16/TF32 misses the 1e-4 bar.
//
// Operand layout (K-major, SWIZZLE_NONE): 8x8 fp16 core matrices of 128 contiguous bytes;
// element (row, k) of a [rows x K] operand at (k/8)*(rows*16) + row*16 + (k%8)*2, so a warp of
// epilogue threads (32 consecutive rows) writes 512 contiguous bytes per 8-feature group.
#include <string.h>

#include "ren_tc.cuh"

#define TC_EPI_WARPS 16
#define TC_THREADS (64 + 32 * TC_EPI_WARPS)   // 576
#define TC_STAGES 3
#define TC_MAX_STEPS 24

// shared-memory map (bytes)
#define SM_A_HI 0          // main activations  [128 x 256] hi
#define SM_A_LO 65536
#define SM_P_HI 131072     // region 1: positional encoding [128 x 64] (dgrad: column 0 = d sigma_raw)
#define SM_P_LO 147456
#define SM_D_HI 163840     // region 2: direction encoding [128 x 32] (dgrad: columns 0..C-1 = d c_raw)
#define SM_D_LO 172032
#define SM_RING 180224
#define SM_STAGE_BYTES 16384
#define SM_BAR (SM_RING + TC_STAGES * SM_STAGE_BYTES)   // 229376
#define SM_TMEM_PTR (SM_BAR + 64)
#define SM_TOTAL (SM_BAR + 128)

#define TMEM_COL_SIGMA 256
#define TMEM_COL_RGB 272

enum { EPI_F_HIDDEN = 0, EPI_F_LINEAR = 1, EPI_F_HEADS = 2, EPI_D_ACT = 3, EPI_D_LINEAR = 4 };

struct TcStep {
  int seg_region[2];   // 0 main, 1 positional-encoding region, 2 direction-encoding region
  int seg_ksteps[2];   // K/16 per segment (0 = unused)
  int N;               // MMA N: 256 / 128 / 16
  int epi;             // EPI_*
  int out_mat;         // stash (fwd) / delta (dgrad) matrix written by the epilogue, -1 none
  int act_mat;         // dgrad: stash matrix of the activation whose derivative multiplies, -1 none
  int write_a;         // result becomes the next step's main A operand
  int bias_off;        // float offset of this step's bias vector in the packed bias area, -1 none
  int tmem_col;        // accumulator base column
  int n_chunks;        // main weight chunks (one per K step, N*64 bytes each)
  int head_chunk;      // fwd: one extra 16 KB chunk = the N=16 sigma-head block over K=256
  long long pack_off;       // byte offset of the main chunks
  long long head_pack_off;  // byte offset of the head chunk
};

struct TcArgs {
  RenPlan p;
  TcStep steps[TC_MAX_STEPS];
  int n_steps;
  int split3;
  long long n_rays;
  int K;
  const float* rays_o;
  const float* rays_d;
  const float* jitter;
  const float* params;
  const unsigned char* pack;     // packed weights
  const float* pack_bias;        // packed (16 B aligned) biases
  unsigned char* stash;          // fp16 stash (tcs_* layout); fwd: written when STASH, dgrad: read
  unsigned char* delta;          // dgrad: fp16 delta mats (written)
  float* raw;                    // fwd: [S][1+C] raw sigma / radiance
  const float* graw;             // dgrad: [S][1+C] d loss / d raw
  const unsigned int* gmax;      // dgrad: max |graw| (float bits) -> power-of-two delta scale
};

__device__ __forceinline__ uint32_t region_hi(int region) {
  return region == 0 ? SM_A_HI : (region == 1 ? SM_P_HI : SM_D_HI);
}
__device__ __forceinline__ uint32_t region_lo(int region) {
  return region == 0 ? SM_A_LO : (region == 1 ? SM_P_LO : SM_D_LO);
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus(beta=100, threshold=20) = ln2/100 * lg2(1 + ex2(100 log2e x)) on the MUFU pipe: 2 MUFU +
// 5 FP32 ops per element, absolute error ~2e-9 (results below 2^-126 flush to zero)
__device__ __forceinline__ float softplus100_fast(float x) {
  const float t = x * 144.26950408889634f;                 // 100 * log2(e)
  const float y = lg2_ftz(1.f + ex2_ftz(t)) * 0.0069314718055994531f;   // ln(2) / 100
  return t > 28.853900817779268f ? x : y;                  // 100 x > 20
}
// d softplus / dx from the OUTPUT a: sigmoid(100 x) = 1 - exp(-100 a)
__device__ __forceinline__ float softplus100_grad_fast(float a) { return 1.f - ex2_ftz(a * -144.26950408889634f); }

// One out-of-line accurate sine keeps the kernel small (84 inlined sinf bodies were ~100 KB of
// SASS and stalled the epilogue warps on instruction fetch).
__device__ __noinline__ float ren_sinf(float x) { return sinf(x); }

// feature j of SinusoidalEncoder(x) with max degree deg (external/mlp.py:227-243)
__device__ __forceinline__ float enc_feature(const float* x, int deg, int j) {
  if (j < 3) return j == 0 ? x[0] : (j == 1 ? x[1] : x[2]);
  if (j >= 3 + 6 * deg) return 0.f;
  const int jj = j - 3;
  const bool shifted = jj >= 3 * deg;
  const int q = shifted ? jj - 3 * deg : jj;
  const int i = q / 3, a = q - 3 * i;
  const float xa = a == 0 ? x[0] : (a == 1 ? x[1] : x[2]);
  const float xb = __fmul_rn(xa, __int_as_float((127 + i) << 23));     // exact 2^i
  return ren_sinf(shifted ? __fadd_rn(xb, 1.5707963267948966f) : xb);
}
// encode k-cores [c0, c1) of an encoding into an operand region (+ fp16 stash)
__device__ __forceinline__ void encode_cores(const float* x, int deg, int kp, int c0, int c1, bool valid,
                                             unsigned char* smem, uint32_t hi_off, uint32_t lo_off, int m,
                                             unsigned char* stash_mat, long long sidx) {
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = valid ? enc_feature(x, deg, c * 8 + j) : 0.f;
    uint4 hi, lo;
    split8(f, &hi, &lo);
    *reinterpret_cast<uint4*>(smem + hi_off + c * (TC_M * 16) + m * 16) = hi;
    *reinterpret_cast<uint4*>(smem + lo_off + c * (TC_M * 16) + m * 16) = lo;
    if (stash_mat != nullptr)
      *reinterpret_cast<uint4*>(stash_mat + 2 * ren_blk(sidx, c * 8, kp)) = hi;
  }
}

template <int MODE, bool STASH>   // MODE 0: forward, 1: dgrad
__global__ void __launch_bounds__(TC_THREADS, 1) ren_tc_chain(const TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const RenPlan& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + SM_BAR;            // [3]
  const uint32_t bar_empty = sbase + SM_BAR + 24;      // [3]
  const uint32_t bar_aready = sbase + SM_BAR + 48;
  const uint32_t bar_accfull = sbase + SM_BAR + 56;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_PTR);
  const long long S = a.n_rays * a.K;
  const long long n_tiles = (S + TC_M - 1) / TC_M;
  const int C1 = 1 + p.C;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_aready, 32 * TC_EPI_WARPS);
    mbar_init(bar_accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + SM_TMEM_PTR) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int g = 0; g < a.n_steps; ++g) {
          const TcStep& st = a.steps[g];
          const uint32_t bytes = (uint32_t)st.N * 64u;
          const unsigned char* src = a.pack + st.pack_off;
          for (int c = 0; c < st.n_chunks; ++c, ++it) {
            const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
            bulk_g2s(sbase + SM_RING + s * SM_STAGE_BYTES, src + (size_t)c * bytes, bytes, bar_full + 8 * s);
          }
          if (st.head_chunk) {
            const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * s, SM_STAGE_BYTES);
            bulk_g2s(sbase + SM_RING + s * SM_STAGE_BYTES, a.pack + st.head_pack_off, SM_STAGE_BYTES,
                     bar_full + 8 * s);
            ++it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, ev = 0;
      const uint32_t idesc16 = umma_idesc(16, 0, 0);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int g = 0; g < a.n_steps; ++g, ++ev) {
          const TcStep& st = a.steps[g];
          mbar_wait(bar_aready, ev & 1u);
          tc_fence_after();
          const uint32_t idesc = umma_idesc(st.N, 0, 0);
          const uint32_t b_lbo = (uint32_t)st.N * 16u;
          const uint32_t d_addr = tmem_base + (uint32_t)st.tmem_col;
          int c = 0;
          for (int seg = 0; seg < 2; ++seg) {
            const uint32_t a_hi = sbase + region_hi(st.seg_region[seg]);
            const uint32_t a_lo = sbase + region_lo(st.seg_region[seg]);
            for (int ks = 0; ks < st.seg_ksteps[seg]; ++ks, ++c, ++it) {
              const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
              mbar_wait(bar_full + 8 * s, ph);
              tc_fence_after();
              const uint32_t wb = sbase + SM_RING + s * SM_STAGE_BYTES;
              const uint64_t da_hi = umma_desc(a_hi + ks * 4096, TC_M * 16, 128);
              const uint64_t db_hi = umma_desc(wb, b_lbo, 128);
              tc_mma_f16(d_addr, da_hi, db_hi, idesc, c > 0 ? 1u : 0u);
              if (a.split3) {
                const uint64_t da_lo = umma_desc(a_lo + ks * 4096, TC_M * 16, 128);
                const uint64_t db_lo = umma_desc(wb + st.N * 32, b_lbo, 128);
                tc_mma_f16(d_addr, da_lo, db_hi, idesc, 1u);
                tc_mma_f16(d_addr, da_hi, db_lo, idesc, 1u);
              }
              tc_commit(bar_empty + 8 * s);   // frees the ring slot when these MMAs have read it
            }
          }
          if (st.head_chunk) {
            // sigma head: [128 x 256] main operand times the N=16 block (row 0 = w_sigma)
            const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t wb = sbase + SM_RING + s * SM_STAGE_BYTES;
            for (int ks = 0; ks < 16; ++ks) {
              const uint64_t da_hi = umma_desc(sbase + SM_A_HI + ks * 4096, TC_M * 16, 128);
              const uint64_t db_hi = umma_desc(wb + ks * 1024, 256, 128);
              tc_mma_f16(tmem_base + TMEM_COL_SIGMA, da_hi, db_hi, idesc16, ks > 0 ? 1u : 0u);
              if (a.split3) {
                const uint64_t da_lo = umma_desc(sbase + SM_A_LO + ks * 4096, TC_M * 16, 128);
                const uint64_t db_lo = umma_desc(wb + ks * 1024 + 512, 256, 128);
                tc_mma_f16(tmem_base + TMEM_COL_SIGMA, da_lo, db_hi, idesc16, 1u);
                tc_mma_f16(tmem_base + TMEM_COL_SIGMA, da_hi, db_lo, idesc16, 1u);
              }
            }
            tc_commit(bar_empty + 8 * s);
            ++it;
          }
          tc_commit(bar_accfull);             // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ===================== prologue + epilogue warps =====================
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;          // column group 0..3
    const int m = q * 32 + lane;             // row of the tile == TMEM lane
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ev = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long sidx = tile * TC_M + m;
      const bool valid = sidx < S;
      bool selector = false;
      if (MODE == 0) {
        // ---- forward prologue: column groups 0-2 encode the position (k-cores 0-2 / 3-5 / 6-7),
        //      group 3 encodes the direction
        float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f};
        long long r = 0;
        int k = 0;
        if (valid) {
          r = sidx / a.K;
          k = (int)(sidx - r * a.K);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            o[c] = a.rays_o[r * 3 + c];
            d[c] = a.rays_d[r * 3 + c];
          }
        }
        if (cg < 3) {
          float z[3] = {0.f, 0.f, 0.f};
          if (valid) {
            float t_min, t_max, t0, t1;
            ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
            ren_march_to(p.r, t_min, k, &t0, &t1);
            selector = ren_field_input(p.f, o, d, t0, t1, z);
            if (!ren_alive(t0, t1, t_max)) selector = false;
          }
          unsigned char* sm = STASH ? a.stash + tcs_encp(p, S) : nullptr;
          const int c0 = cg == 0 ? 0 : (cg == 1 ? 3 : 6), c1 = cg == 0 ? 3 : (cg == 1 ? 6 : 8);
          encode_cores(z, 10, 64, c0, c1, valid, smem, SM_P_HI, SM_P_LO, m, sm, sidx);
        } else {
          const float dz[3] = {__fmul_rn(d[0], 3.141592653589793f), __fmul_rn(d[1], 3.141592653589793f),
                               __fmul_rn(d[2], 3.141592653589793f)};
          encode_cores(dz, 4, 32, 0, 4, valid, smem, SM_D_HI, SM_D_LO, m,
                       STASH ? a.stash + tcs_encd(p, S) : nullptr, sidx);
        }
      } else {
        // ---- dgrad prologue: head gradients enter as rank-1 K=16 segments:
        //      region 1 column 0 = d sigma_raw, region 2 columns 0..C-1 = d c_raw
        if (cg == 0) {
          float gs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float gc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (valid) {
            const float sc = delta_scale(a.gmax);
            gs[0] = a.graw[sidx * C1] * sc;
            for (int c = 0; c < p.C; ++c) gc[c] = a.graw[sidx * C1 + 1 + c] * sc;
          }
          uint4 hi, lo;
          split8(gs, &hi, &lo);
          *reinterpret_cast<uint4*>(smem + SM_P_HI + m * 16) = hi;
          *reinterpret_cast<uint4*>(smem + SM_P_LO + m * 16) = lo;
          split8(gc, &hi, &lo);
          *reinterpret_cast<uint4*>(smem + SM_D_HI + m * 16) = hi;
          *reinterpret_cast<uint4*>(smem + SM_D_LO + m * 16) = lo;
        } else if (cg == 1) {
          const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);   // second K core of both rank-1 segments
          *reinterpret_cast<uint4*>(smem + SM_P_HI + TC_M * 16 + m * 16) = z4;
          *reinterpret_cast<uint4*>(smem + SM_P_LO + TC_M * 16 + m * 16) = z4;
          *reinterpret_cast<uint4*>(smem + SM_D_HI + TC_M * 16 + m * 16) = z4;
          *reinterpret_cast<uint4*>(smem + SM_D_LO + TC_M * 16 + m * 16) = z4;
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_aready);

      // ---- one epilogue per GEMM step
      for (int g = 0; g < a.n_steps; ++g, ++ev) {
        const TcStep& st = a.steps[g];
        const int ncol = st.N >> 2;                  // columns this warp handles (64 / 32 / 4)
        const int col0 = cg * ncol;
        // dgrad: prefetch this thread's slice of the activation stash while the MMAs run
        // byte offset of this thread's 8-feature group g8 inside a blocked fp16 matrix of width N:
        //   (tile * N/8 + g8) * 2048 + m * 16   == the shared-memory operand offset + a per-tile base
        const size_t tile_off = (size_t)tile * (size_t)(st.N >> 3) * 2048 + (size_t)m * 16;
        uint4 act_pref[8];
        if (MODE == 1 && st.epi == EPI_D_ACT && valid) {
          const unsigned char* am = a.stash + tcs_mat(p, S, st.act_mat) + tile_off + (size_t)(col0 >> 3) * 2048;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i * 8 < ncol) act_pref[i] = __ldg(reinterpret_cast<const uint4*>(am + i * 2048));
        }
        mbar_wait(bar_accfull, ev & 1u);
        tc_fence_after();
        if (st.epi == EPI_F_HEADS) {
          if (cg == 0) {
            uint32_t rs[16], rc[16];
            tc_ld16(tmem_lane + TMEM_COL_SIGMA, rs);
            tc_ld16(tmem_lane + TMEM_COL_RGB, rc);
            tc_ld_wait();
            if (valid) {
              const float sraw = __uint_as_float(rs[0]) + __ldg(a.params + p.sig_b);
              a.raw[sidx * C1] = selector ? sraw : -INFINITY;
              for (int c = 0; c < p.C; ++c)
                a.raw[sidx * C1 + 1 + c] = __uint_as_float(rc[c]) + __ldg(a.params + p.rgb_b + c);
            }
          }
        } else {
          unsigned char* out_mat = nullptr;       // already offset to this thread's row of the tile
          if (MODE == 0) {
            if (STASH && st.out_mat >= 0) out_mat = a.stash + tcs_mat(p, S, st.out_mat) + tile_off;
          } else if (st.out_mat >= 0) {      // padded rows are written as zeros (wgrad reads whole tiles)
            out_mat = a.delta + tcs_mat(p, S, st.out_mat) + tile_off;
          }
          unsigned char* a_row = smem + (uint32_t)m * 16;
          const float* bias = (MODE == 0 && st.bias_off >= 0) ? a.pack_bias + st.bias_off : nullptr;
          const uint32_t tcol = tmem_lane + (uint32_t)(st.tmem_col + col0);
          uint32_t r[2][16];
          tc_ld16(tcol, r[0]);
          tc_ld_wait();
#pragma unroll
          for (int cb = 0; cb < 64; cb += 32) {       // fully unrolled: register arrays stay registers
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              const int cbb = cb + hb * 16;
              if (cbb < ncol) {
                if (cbb + 16 < ncol) tc_ld16(tcol + cbb + 16, r[hb ^ 1]);   // prefetch the next block
#pragma unroll
                for (int j8 = 0; j8 < 2; ++j8) {
                  const int n0 = col0 + cbb + j8 * 8;
                  float y[8];
                  if (MODE == 0) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float x = __uint_as_float(r[hb][j8 * 8 + j]) + bb[j];
                      if (st.epi == EPI_F_LINEAR) y[j] = x;
                      else y[j] = (p.f.hidden_activation == REN_HIDDEN_SOFTPLUS100) ? softplus100_fast(x) : fmaxf(x, 0.f);
                    }
                  } else {
                    float av[8];
                    if (st.epi == EPI_D_ACT) unpack8(act_pref[(cbb >> 3) + j8], av);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      float v = __uint_as_float(r[hb][j8 * 8 + j]);
                      if (st.epi == EPI_D_ACT)
                        v *= (p.f.hidden_activation == REN_HIDDEN_SOFTPLUS100) ? softplus100_grad_fast(av[j])
                                                                               : (av[j] > 0.f ? 1.f : 0.f);
                      y[j] = valid ? v : 0.f;
                    }
                  }
                  uint4 hi, lo;
                  split8(y, &hi, &lo);
                  const uint32_t goff = (uint32_t)(n0 >> 3) * (TC_M * 16);
                  if (st.write_a) {
                    *reinterpret_cast<uint4*>(a_row + SM_A_HI + goff) = hi;
                    *reinterpret_cast<uint4*>(a_row + SM_A_LO + goff) = lo;
                  }
                  if (out_mat != nullptr) *reinterpret_cast<uint4*>(out_mat + goff) = hi;
                }
                tc_ld_wait();
              }
            }
          }
        }
        tc_fence_before();
        if (g + 1 < a.n_steps) {
          fence_proxy_async();
          mbar_arrive(bar_aready);
        }
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

// ------------------------------------------------------------------ weight packing
// A pack job writes one B operand: chunks of 16 K' values, each chunk = [hi: N x 16][lo: N x 16]
// in the core-matrix layout.  mode 0: B[n][k'] = W[n*ldw + col(k')] (forward; two input segments);
// mode 1: B[n][k'] = W[k'*ldw + n] (transposed, dgrad); mode 2: fp32 copy of a bias vector.
#define TC_MAX_PACK 64
struct PackJob {
  long long dst_off;   // bytes (modes 0/1) or floats (mode 2)
  int mode, N, n_ksteps;
  int w_off, ldw;
  int Ntrue;           // valid rows n
  int ks0, K0, K1;     // mode 0: segment 0 padded K steps / true widths; mode 1: K0 = valid K'
};
struct PackArgs {
  PackJob jobs[TC_MAX_PACK];
  int n_jobs;
  const float* params;
  unsigned char* pack;
  float* pack_bias;
};

extern "C" __global__ void ren_tc_pack(const PackArgs a) {
  const int g = blockIdx.y;
  if (g >= a.n_jobs) return;
  const PackJob& jb = a.jobs[g];
  if (jb.mode == 2) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < jb.N; n += gridDim.x * blockDim.x)
      a.pack_bias[jb.dst_off + n] = n < jb.Ntrue ? a.params[jb.w_off + n] : 0.f;
    return;
  }
  const int total = jb.n_ksteps * 16 * jb.N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e % jb.N;
    const int kc = e / jb.N;
    const int c = kc >> 4, kk = kc & 15;
    float w = 0.f;
    if (n < jb.Ntrue) {
      if (jb.mode == 0) {
        const int k0pad = jb.ks0 * 16;
        if (kc < k0pad) {
          if (kc < jb.K0) w = a.params[jb.w_off + (long long)n * jb.ldw + kc];
        } else if (kc - k0pad < jb.K1) {
          w = a.params[jb.w_off + (long long)n * jb.ldw + jb.K0 + (kc - k0pad)];
        }
      } else if (kc < jb.K0) {
        w = a.params[jb.w_off + (long long)kc * jb.ldw + n];
      }
    }
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    unsigned char* chunk = a.pack + jb.dst_off + (size_t)c * jb.N * 64;
    const size_t off = (size_t)(kk >> 3) * (jb.N * 16) + (size_t)n * 16 + (kk & 7) * 2;
    *reinterpret_cast<__half*>(chunk + off) = hi;
    *reinterpret_cast<__half*>(chunk + (size_t)jb.N * 32 + off) = lo;
  }
}

// ------------------------------------------------------------------ host side
struct TcProgram {
  TcStep steps[TC_MAX_STEPS];
  int n_steps;
  PackJob jobs[TC_MAX_PACK];
  int n_jobs;
  long long pack_bytes;    // packed weights
  long long bias_floats;   // packed biases
};

static void add_pack(TcProgram& P, int mode, long long dst, int N, int n_ksteps, int w_off, int ldw, int Ntrue,
                     int ks0, int K0, int K1) {
  PackJob& j = P.jobs[P.n_jobs++];
  j.dst_off = dst; j.mode = mode; j.N = N; j.n_ksteps = n_ksteps; j.w_off = w_off; j.ldw = ldw; j.Ntrue = Ntrue;
  j.ks0 = ks0; j.K0 = K0; j.K1 = K1;
}

static void build_forward(const RenPlan& p, TcProgram& P) {
  memset(&P, 0, sizeof(P));
  long long off = 0, boff = 0;
  auto add_layer = [&](const RenLayer& L, int epi) -> TcStep& {
    TcStep& s = P.steps[P.n_steps++];
    s.seg_region[0] = L.src0 == 1 ? 1 : 0;
    s.seg_ksteps[0] = (L.K0 + 15) / 16;
    s.seg_region[1] = L.src1;
    s.seg_ksteps[1] = L.K1 > 0 ? (L.K1 + 15) / 16 : 0;
    s.N = L.N; s.epi = epi; s.out_mat = L.mat; s.act_mat = -1; s.write_a = 1; s.tmem_col = 0;
    s.n_chunks = s.seg_ksteps[0] + s.seg_ksteps[1];
    s.pack_off = off;
    add_pack(P, 0, off, L.N, s.n_chunks, L.w_off, L.K0 + L.K1, L.N, s.seg_ksteps[0], L.K0, L.K1);
    off += (long long)s.n_chunks * s.N * 64;
    s.bias_off = (int)boff;
    add_pack(P, 2, boff, 256, 0, L.b_off, 0, L.N, 0, 0, 0);
    boff += 256;
    return s;
  };
  for (int i = 0; i < p.D; ++i) add_layer(p.trunk[i], EPI_F_HIDDEN);
  {
    TcStep& s = add_layer(p.bott, EPI_F_LINEAR);
    // sigma head rides on the bottleneck step (same A operand = last trunk activations)
    s.head_chunk = 1;
    s.head_pack_off = off;
    add_pack(P, 0, off, 16, 16, p.sig_w, p.W, 1, 16, p.W, 0);
    off += SM_STAGE_BYTES;
  }
  add_layer(p.cond[0], EPI_F_HIDDEN);
  {
    // radiance head: [128 x WC] condition activations times the N=16 block (rows 0..C-1 = W_o)
    TcStep& s = P.steps[P.n_steps++];
    s.seg_region[0] = 0; s.seg_ksteps[0] = p.WC / 16; s.seg_region[1] = 0; s.seg_ksteps[1] = 0;
    s.N = 16; s.epi = EPI_F_HEADS; s.out_mat = -1; s.act_mat = -1; s.write_a = 0; s.bias_off = -1;
    s.tmem_col = TMEM_COL_RGB; s.n_chunks = s.seg_ksteps[0]; s.pack_off = off;
    add_pack(P, 0, off, 16, s.n_chunks, p.rgb_w, p.WC, p.C, s.n_chunks, p.WC, 0);
    off += (long long)s.n_chunks * 16 * 64;
  }
  P.pack_bytes = off;
  P.bias_floats = boff;
}

static void build_dgrad(const RenPlan& p, TcProgram& P) {
  memset(&P, 0, sizeof(P));
  long long off = 0;
  auto add = [&](int region0, int ks0, int region1, int ks1, int N, int epi, int out_mat, int act_mat,
                 int write_a) -> TcStep& {
    TcStep& s = P.steps[P.n_steps++];
    s.seg_region[0] = region0; s.seg_ksteps[0] = ks0; s.seg_region[1] = region1; s.seg_ksteps[1] = ks1;
    s.N = N; s.epi = epi; s.out_mat = out_mat; s.act_mat = act_mat; s.write_a = write_a; s.bias_off = -1;
    s.tmem_col = 0; s.n_chunks = ks0 + ks1; s.pack_off = off;
    return s;
  };
  // t0: d g = (d c_raw) W_o, times act'(g)            [rank-C segment in region 2]
  {
    TcStep& s = add(2, 1, 0, 0, p.WC, EPI_D_ACT, p.cond[0].mat, p.cond[0].mat, 1);
    add_pack(P, 1, off, p.WC, 1, p.rgb_w, p.WC, p.WC, 0, p.C, 0);
    off += (long long)s.n_chunks * s.N * 64;
  }
  // t1: d bott = d g_pre W_r0[:, :W]                   (bottleneck is linear)
  {
    TcStep& s = add(0, p.WC / 16, 0, 0, p.W, EPI_D_LINEAR, p.bott.mat, -1, 1);
    add_pack(P, 1, off, p.W, s.n_chunks, p.cond[0].w_off, p.cond[0].K0 + p.cond[0].K1, p.W, 0, p.WC, 0);
    off += (long long)s.n_chunks * s.N * 64;
  }
  // t2: d a_{D-1} = d bott W_b + (d sigma_raw) w_sigma, times act'(a_{D-1})
  {
    const int last = p.trunk[p.D - 1].mat;
    TcStep& s = add(0, p.W / 16, 1, 1, p.W, EPI_D_ACT, last, last, p.D > 1 ? 1 : 0);
    add_pack(P, 1, off, p.W, p.W / 16, p.bott.w_off, p.bott.K0 + p.bott.K1, p.W, 0, p.W, 0);
    add_pack(P, 1, off + (long long)(p.W / 16) * p.W * 64, p.W, 1, p.sig_w, p.W, p.W, 0, 1, 0);
    off += (long long)s.n_chunks * s.N * 64;
  }
  // trunk layers D-1 .. 1: d a_{i-1} = Delta_i W_i[:, :W], times act'(a_{i-1})
  for (int i = p.D - 1; i >= 1; --i) {
    const RenLayer& L = p.trunk[i];
    TcStep& s = add(0, p.W / 16, 0, 0, p.W, EPI_D_ACT, p.trunk[i - 1].mat, p.trunk[i - 1].mat, i > 1 ? 1 : 0);
    add_pack(P, 1, off, p.W, s.n_chunks, L.w_off, L.K0 + L.K1, p.W, 0, p.W, 0);
    off += (long long)s.n_chunks * s.N * 64;
  }
  P.pack_bytes = off;
  P.bias_floats = 0;
}

int ren_tc_supported(const RenPlan& p) {
  return p.W == 256 && p.WC == 128 && p.CD == 1 && p.hid_K1 == 0 && p.f.pos_deg == 10 && p.f.dir_deg == 4 &&
         p.D >= 1 && p.D + 3 <= TC_MAX_STEPS && p.C <= 3 && p.trunk[0].K1 == 0;
}

static size_t program_bytes(const TcProgram& P) { return (size_t)P.pack_bytes + (size_t)P.bias_floats * 4 + 2048; }

size_t ren_tc_fwd_workspace_bytes(const RenPlan& p) {
  if (!ren_tc_supported(p)) return 0;
  TcProgram P;
  build_forward(p, P);
  return program_bytes(P);
}
size_t ren_tc_bwd_pack_bytes(const RenPlan& p) {
  if (!ren_tc_supported(p)) return 0;
  TcProgram P;
  build_dgrad(p, P);
  return program_bytes(P);
}

static int run_pack(const TcProgram& P, const float* params, unsigned char* pack, float* pack_bias,
                    cudaStream_t st) {
  PackArgs pa;
  for (int i = 0; i < P.n_jobs; ++i) pa.jobs[i] = P.jobs[i];
  pa.n_jobs = P.n_jobs; pa.params = params; pa.pack = pack; pa.pack_bias = pack_bias;
  ren_prof_begin(REN_K_PACK, st);
  ren_tc_pack<<<dim3(16, P.n_jobs), 256, 0, st>>>(pa);
  ren_prof_end(REN_K_PACK, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc pack launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

template <int MODE, bool STASH>
static int launch_chain(const TcArgs& a, long long n_tiles, int kind, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (grid < 1) return REN_OK;
  cudaError_t e = cudaFuncSetAttribute(ren_tc_chain<MODE, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
  if (e != cudaSuccess) { ren_set_error("tc chain smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  ren_prof_begin(kind, st);
  ren_tc_chain<MODE, STASH><<<grid, TC_THREADS, SM_TOTAL, st>>>(a);
  ren_prof_end(kind, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc chain launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

static int carve_pack(const TcProgram& P, void* ws, size_t ws_bytes, unsigned char** pack, float** pack_bias) {
  unsigned char* base = (unsigned char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const size_t need = (size_t)(base - (unsigned char*)ws) + (size_t)P.pack_bytes + (size_t)P.bias_floats * 4 + 16;
  if (ws == nullptr || need > ws_bytes) {
    ren_set_error("workspace too small for packed weights (%zu > %zu)", need, ws_bytes);
    return REN_E_WORKSPACE;
  }
  *pack = base;
  *pack_bias = (float*)(base + ((P.pack_bytes + 15) & ~15LL));
  return REN_OK;
}

int ren_tc_launch_fwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                      const float* jitter, const float* params, void* stash, float* raw, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  TcProgram P;
  build_forward(p, P);
  unsigned char* pack; float* pack_bias;
  int rc = carve_pack(P, ws, ws_bytes, &pack, &pack_bias);
  if (rc) return rc;
  rc = run_pack(P, params, pack, pack_bias, st);
  if (rc) return rc;
  TcArgs a;
  static_assert(sizeof(TcArgs) <= 4000, "kernel parameter block too large");
  memset(&a, 0, sizeof(a));
  a.p = p;
  for (int i = 0; i < P.n_steps; ++i) a.steps[i] = P.steps[i];
  a.n_steps = P.n_steps;
  a.split3 = p.r.engine == REN_ENGINE_TC_F16 ? 0 : 1;
  a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter; a.params = params;
  a.pack = pack; a.pack_bias = pack_bias; a.stash = (unsigned char*)stash; a.raw = raw;
  const long long n_tiles = (n_rays * K + TC_M - 1) / TC_M;
  return stash ? launch_chain<0, true>(a, n_tiles, REN_K_FIELD_FWD, st)
               : launch_chain<0, false>(a, n_tiles, REN_K_FIELD_FWD, st);
}

int ren_tc_launch_dgrad(const RenPlan& p, long long S, const float* params, const void* stash,
                        const float* graw, const unsigned int* gmax, void* delta, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  TcProgram P;
  build_dgrad(p, P);
  unsigned char* pack; float* pack_bias;
  int rc = carve_pack(P, ws, ws_bytes, &pack, &pack_bias);
  if (rc) return rc;
  rc = run_pack(P, params, pack, pack_bias, st);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  for (int i = 0; i < P.n_steps; ++i) a.steps[i] = P.steps[i];
  a.n_steps = P.n_steps;
  a.split3 = p.r.engine == REN_ENGINE_TC_F16 ? 0 : 1;
  a.n_rays = S; a.K = 1;            // dgrad only needs the flat sample count
  a.params = params; a.pack = pack; a.pack_bias = pack_bias;
  a.stash = (unsigned char*)stash; a.delta = (unsigned char*)delta; a.graw = graw; a.gmax = gmax;
  const long long n_tiles = (S + TC_M - 1) / TC_M;
  return launch_chain<1, false>(a, n_tiles, REN_K_FIELD_DGRAD, st);
}
