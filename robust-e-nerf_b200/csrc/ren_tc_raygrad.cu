// ren_tc_raygrad.cu -- tcgen05 version of the ray-gradient kernel (fp32 SIMT form and the
// maths in ren_raygrad.cu).
//
// d loss / d enc = sum over the layers that consume an encoding of Delta_layer W_layer[:, encoding columns]:
// skinny GEMMs (N = 64 positional / 32 direction columns, K = 256 / 128 delta features) over the fp16 delta
// matrices the dgrad chain kernel already wrote for the weight gradients.  Their BLOCKED layout
// ([128 samples][8 features] per 2 KB, ren_tc.cuh) is a K-major UMMA operand as it lies in HBM, so a tile
// moves into shared memory with plain 32 KB bulk copies and feeds tcgen05.mma directly:
//   warp 0      producer: delta tiles, HBM -> 2 x 32 KB ring (cp.async.bulk + mbarrier tx bytes)
//   warp 1      MMA issuer: M = 128 samples, A = delta (fp16, power-of-two scaled), B = W^T slices as fp16
//               hi + lo (2 MMAs per K step), resident in shared memory -- every CTA packs them from the fp32
//               parameters at start; fp32 accumulators in TMEM, two tiles in flight (2 x 192 columns)
//   warps 2-9   two epilogue sets of four warps (TMEM lane quadrant = warp % 4), alternating tiles: thread =
//               sample row; tcgen05.ld of its d loss / d enc, then the reverse pass through the sinusoidal
//               encoder, the contraction (hyper-dual numbers in tangent mode) and x = o + d t to the ray,
//               reduced over the samples of a ray inside the warp before the atomics.
// HBM-bound by construction: 1.25 KB of deltas per sample (2.5 KB in tangent mode), ~0.1 MFLOP per sample.
#include <string.h>

#include "ren_tc.cuh"
#include "ren_dual.cuh"

int ren_tc_supported(const RenPlan& p);      // ren_tc.cu

#define RGT_THREADS 320
#define RGT_STAGE_BYTES 32768            // 128 delta features x 128 samples, fp16
#define RGT_STAGES 2
#define RGT_MAX_SRC 3
#define RGT_B_BYTES 147456               // resident weights: 2 x [64 x 256] + [32 x 128], hi + lo
#define RGT_SM_RING RGT_B_BYTES
#define RGT_SM_BAR (RGT_SM_RING + RGT_STAGES * RGT_STAGE_BYTES)
#define RGT_SM_TOTAL (RGT_SM_BAR + 256)

struct RgtSrc {
  long long d_off;   // byte offset of the layer's delta matrix inside delta / delta_dot
  int N;             // delta features: 128 or 256
  int enc;           // 1 positional (64 columns), 2 direction (32 columns)
  int w_off;         // params offset of W[0][first encoding column]
  int ldw;
  int b_off;         // shared-memory byte offset of the packed W^T slice
};
struct RgtArgs {
  RenPlan p;
  RgtSrc src[RGT_MAX_SRC];
  int n_src;
  long long n_rays;
  int K;
  int tangent;
  const unsigned char* delta;
  const unsigned char* delta_dot;
  const unsigned int* gmax;
  const unsigned char* tile_live;
  const int* sample_idx;
  long long n_samples;
  const float* params;
  const float* rays_o;
  const float* rays_d;
  const float* rays_o_dot;
  const float* rays_d_dot;
  const float* jitter;
  float* g_rays;
};

// 64-bit shared-memory matrix descriptor from its low word (address/16 | LBO/16 << 16); SBO = 128 B, version 1
__device__ __forceinline__ uint64_t mk_desc_rg(uint32_t lo32) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo32), "r"(0x4008u));
  return d;
}

__device__ __noinline__ float rgt_cosf(float x) { return cosf(x); }
__device__ __noinline__ void rgt_sincosf(float x, float* s, float* c) { sincosf(x, s, c); }

// reverse pass of SinusoidalEncoder (external/mlp.py:227-243) with static indices: gv = d loss / d enc,
// gt = d loss / d enc_dot (tangent mode), x / xdot the encoder input and its d/dw
template <int DEG, bool TAN>
__device__ __forceinline__ void rgt_enc_backward(const float* x, const float* xdot, const float* gv, const float* gt,
                                                 float* gx, float* gxd) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gx[c] = gv[c];
    gxd[c] = TAN ? gt[c] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < DEG; ++i) {
    const float sc = __int_as_float((127 + i) << 23);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xb = __fmul_rn(x[c], sc);
      const float xs = __fadd_rn(xb, 1.5707963267948966f);
      const int js = 3 + i * 3 + c, jc = 3 + DEG * 3 + i * 3 + c;
      if (TAN) {
        float s0, c0, s1, c1;
        rgt_sincosf(xb, &s0, &c0);
        rgt_sincosf(xs, &s1, &c1);
        gx[c] = fmaf(gv[js], c0 * sc, fmaf(gv[jc], c1 * sc, gx[c]));
        gx[c] -= (gt[js] * (s0 * sc * sc) + gt[jc] * (s1 * sc * sc)) * xdot[c];
        gxd[c] = fmaf(gt[js], c0 * sc, fmaf(gt[jc], c1 * sc, gxd[c]));
      } else {
        gx[c] = fmaf(gv[js], rgt_cosf(xb) * sc, fmaf(gv[jc], rgt_cosf(xs) * sc, gx[c]));
      }
    }
  }
}

template <int NCOL>
__device__ __forceinline__ void rgt_load(uint32_t taddr, float unscale, float* out) {
  uint32_t r[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; c += 16) tc_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&r[c]));
  tc_ld_wait();
#pragma unroll
  for (int c = 0; c < NCOL; ++c) out[c] = __uint_as_float(r[c]) * unscale;
}

template <bool TAN>
__global__ void __launch_bounds__(RGT_THREADS, 1) ren_tc_raygrad(const RgtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const RenPlan& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + RGT_SM_BAR;            // [RGT_STAGES]
  const uint32_t bar_empty = sbase + RGT_SM_BAR + 32;      // [RGT_STAGES]
  const uint32_t bar_acc = sbase + RGT_SM_BAR + 64;        // [2] accumulators of TMEM set s complete
  const uint32_t bar_free = sbase + RGT_SM_BAR + 96;       // [2] TMEM set s has been read out
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + RGT_SM_BAR + 128);
  const long long S = a.sample_idx != nullptr ? a.n_samples : a.n_rays * a.K;
  const long long n_tiles = (S + TC_M - 1) / TC_M;
  constexpr int NRS = TAN ? 2 : 1;                         // row sets: value deltas (, d/dw deltas)

  if (threadIdx.x == 0) {
    for (int s = 0; s < RGT_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc + 8 * s, 1);
      mbar_init(bar_free + 8 * s, 4);                      // one elected lane per epilogue warp of the set
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W^T slices -> fp16 hi / lo K-major B operands: per 16-deep K step [hi: 2 K-cores x rows x 16 B][lo: same]
  for (int si = 0; si < a.n_src; ++si) {
    const RgtSrc& sc = a.src[si];
    const int rows = sc.enc == 1 ? REN_PE_MAX : REN_DE_MAX, width = sc.enc == 1 ? p.PE : p.DE;
    const float* W = a.params + sc.w_off;
    for (int e = threadIdx.x; e < sc.N * rows; e += blockDim.x) {
      const int n = e % rows, k = e / rows;
      const float w = n < width ? __ldg(W + (long long)k * sc.ldw + n) : 0.f;
      const __half hi = __float2half_rn(w);
      const __half lo = __float2half_rn(w - __half2float(hi));
      unsigned char* chunk = smem + sc.b_off + (k >> 4) * (rows * 64);
      const int off = ((k & 15) >> 3) * (rows * 16) + n * 16 + (k & 7) * 2;
      *reinterpret_cast<__half*>(chunk + off) = hi;
      *reinterpret_cast<__half*>(chunk + rows * 32 + off) = lo;
    }
  }
  fence_proxy_async();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + RGT_SM_BAR + 128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== producer: delta tiles =====================
    if (REN_ONE_LANE()) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (a.tile_live != nullptr && !__ldg(a.tile_live + tile)) continue;
        for (int rs = 0; rs < NRS; ++rs) {
          const unsigned char* base = rs ? a.delta_dot : a.delta;
          for (int si = 0; si < a.n_src; ++si) {
            const RgtSrc& sc = a.src[si];
            const unsigned char* tb = base + sc.d_off + (size_t)tile * (size_t)(sc.N >> 3) * 2048;
            for (int j = 0; j < (sc.N >> 7); ++j, ++it) {
              const uint32_t s = it % RGT_STAGES, ph = (it / RGT_STAGES) & 1u;
              mbar_wait(bar_empty + 8 * s, ph ^ 1u);
              mbar_arrive_expect_tx(bar_full + 8 * s, RGT_STAGE_BYTES);
              bulk_g2s(sbase + RGT_SM_RING + s * RGT_STAGE_BYTES, tb + (size_t)j * RGT_STAGE_BYTES, RGT_STAGE_BYTES,
                       bar_full + 8 * s);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (REN_ONE_LANE()) {
      uint32_t it = 0, n_t = 0;
      const uint32_t idesc_p = umma_idesc(REN_PE_MAX, 0, 0), idesc_d = umma_idesc(REN_DE_MAX, 0, 0);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (a.tile_live != nullptr && !__ldg(a.tile_live + tile)) continue;
        const uint32_t set = n_t & 1u;
        mbar_wait(bar_free + 8 * set, ((n_t >> 1) & 1u) ^ 1u);     // the epilogue of tile n_t - 2 has drained this set
        tc_fence_after();
        for (int rs = 0; rs < NRS; ++rs) {
          uint32_t acc_p = 0u, acc_d = 0u;                           // first MMA into a region overwrites
          for (int si = 0; si < a.n_src; ++si) {
            const RgtSrc& sc = a.src[si];
            const uint32_t rows = sc.enc == 1 ? REN_PE_MAX : REN_DE_MAX;
            const uint32_t idesc = sc.enc == 1 ? idesc_p : idesc_d;
            const uint32_t d_addr = tmem_base + set * 256u + (uint32_t)rs * 96u + (sc.enc == 1 ? 0u : 64u);
            uint32_t& acc = sc.enc == 1 ? acc_p : acc_d;
            for (int j = 0; j < (sc.N >> 7); ++j, ++it) {
              const uint32_t s = it % RGT_STAGES;
              mbar_wait(bar_full + 8 * s, (it / RGT_STAGES) & 1u);
              tc_fence_after();
              const uint32_t a_lo32 = (((sbase + RGT_SM_RING + s * RGT_STAGE_BYTES) >> 4) & 0x3FFFu) | (128u << 16);
              const uint32_t b_lo32 = (((sbase + sc.b_off + (uint32_t)j * 8u * rows * 64u) >> 4) & 0x3FFFu) | (rows << 16);
#pragma unroll
              for (int kl = 0; kl < 8; ++kl) {
                const uint64_t da = mk_desc_rg(a_lo32 + kl * 256u);
                const uint32_t b_k = b_lo32 + (uint32_t)kl * rows * 4u;
                tc_mma_f16(d_addr, da, mk_desc_rg(b_k), idesc, acc);
                tc_mma_f16(d_addr, da, mk_desc_rg(b_k + rows * 2u), idesc, 1u);
                acc = 1u;
              }
              tc_commit(bar_empty + 8 * s);
            }
          }
        }
        tc_commit(bar_acc + 8 * set);
        ++n_t;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const uint32_t set = (uint32_t)(warp - 2) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const float unscale = 1.f / delta_scale(a.gmax);
    uint32_t n_t = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (a.tile_live != nullptr && !__ldg(a.tile_live + tile)) continue;
      if ((n_t & 1u) != set) { ++n_t; continue; }
      const long long sidx = tile * TC_M + m;
      const bool valid = sidx < S;
      // geometry of this sample (overlaps the wait for the accumulators)
      long long r = -1;
      float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f}, od[3] = {0.f, 0.f, 0.f}, dd[3] = {0.f, 0.f, 0.f};
      float z[3] = {0.f, 0.f, 0.f}, pos[3] = {0.f, 0.f, 0.f}, xdot[3] = {0.f, 0.f, 0.f}, zdot[3] = {0.f, 0.f, 0.f};
      float tm = 0.f;
      if (valid) {
        const long long dense = a.sample_idx != nullptr ? (long long)a.sample_idx[sidx] : sidx;
        r = dense / a.K;
        const int k = (int)(dense - r * a.K);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          o[c] = a.rays_o[r * 3 + c];
          d[c] = a.rays_d[r * 3 + c];
          if (TAN) { od[c] = a.rays_o_dot[r * 3 + c]; dd[c] = a.rays_d_dot[r * 3 + c]; }
        }
        float t_min, t_max, t0, t1;
        ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
        ren_march_to(p.r, t_min, k, &t0, &t1);
        ren_field_input(p.f, o, d, t0, t1, z);
        tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          pos[c] = __fadd_rn(o[c], __fmul_rn(__fmul_rn(d[c], __fadd_rn(t0, t1)), 0.5f));
          xdot[c] = fmaf(dd[c], tm, od[c]);
        }
        if (TAN) {
          Dual pd[3], zd[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) pd[c] = Dual(pos[c], xdot[c]);
          ren_contract_t<Dual>(p.f, pd, zd);
#pragma unroll
          for (int c = 0; c < 3; ++c) zdot[c] = zd[c].d;
        }
      }
      mbar_wait(bar_acc + 8 * set, (n_t >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * 256u;
      float gz[3], gzd[3], gdz[3], gdzd[3];
      {
        float gv[REN_PE_MAX], gt[TAN ? REN_PE_MAX : 1];
        rgt_load<REN_PE_MAX>(taddr, unscale, gv);
        if (TAN) rgt_load<REN_PE_MAX>(taddr + 96u, unscale, gt);
        rgt_enc_backward<10, TAN>(z, zdot, gv, gt, gz, gzd);
      }
      {
        float gv[REN_DE_MAX], gt[TAN ? REN_DE_MAX : 1];
        rgt_load<REN_DE_MAX>(taddr + 64u, unscale, gv);
        if (TAN) rgt_load<REN_DE_MAX>(taddr + 160u, unscale, gt);
        // every accumulator of this tile is in registers: hand the TMEM set back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 8 * set);
        const float dz[3] = {__fmul_rn(d[0], 3.141592653589793f), __fmul_rn(d[1], 3.141592653589793f),
                             __fmul_rn(d[2], 3.141592653589793f)};
        const float dzd[3] = {dd[0] * 3.141592653589793f, dd[1] * 3.141592653589793f, dd[2] * 3.141592653589793f};
        rgt_enc_backward<4, TAN>(dz, dzd, gv, gt, gdz, gdzd);
      }
      // through the contraction: gx_k = gz . J e_k + gzd . H[e_k, xdot];  gxd_k = gzd . J e_k
      float out[TAN ? 12 : 6];
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        HDual ph[3], zh[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) ph[c] = HDual(pos[c], c == kk ? 1.f : 0.f, xdot[c], 0.f);
        ren_contract_t<HDual>(p.f, ph, zh);
        float sx = 0.f, sxd = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sx = fmaf(gz[c], zh[c].d1, fmaf(gzd[c], zh[c].d12, sx));
          sxd = fmaf(gzd[c], zh[c].d1, sxd);
        }
        out[kk] = valid ? sx : 0.f;
        out[3 + kk] = valid ? fmaf(tm, sx, 3.141592653589793f * gdz[kk]) : 0.f;
        if (TAN) {
          out[6 + kk] = valid ? sxd : 0.f;
          out[9 + kk] = valid ? fmaf(tm, sxd, 3.141592653589793f * gdzd[kk]) : 0.f;
        }
      }
      // the samples of a ray are consecutive rows: segmented reduction inside the warp, one atomic set per ray
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const long long rk = __shfl_down_sync(0xffffffffu, r, off);
        const bool same = lane + off < 32 && rk == r;
#pragma unroll
        for (int c = 0; c < (TAN ? 12 : 6); ++c) {
          const float t = __shfl_down_sync(0xffffffffu, out[c], off);
          if (same) out[c] += t;
        }
      }
      const long long rprev = __shfl_up_sync(0xffffffffu, r, 1);
      if (valid && (lane == 0 || rprev != r)) {
        float* g = a.g_rays + r * (TAN ? 12 : 6);
#pragma unroll
        for (int c = 0; c < (TAN ? 12 : 6); ++c) atomicAdd(g + c, out[c]);
      }
      ++n_t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// Sources of the tcgen05 engine's network (ren_tc_supported): first trunk layer and the skip layers (positional),
// first condition layer (direction).  Returns 0 when the resident-weight layout does not fit (fp32 kernel instead).
int ren_tc_raygrad_supported(const RenPlan& p) {
  if (!ren_tc_supported(p)) return 0;
  int n_p = 0;
  for (int i = 0; i < p.D; ++i) n_p += (p.trunk[i].src0 == 1) + (p.trunk[i].K1 > 0);
  if (p.bott.src0 == 1 || p.bott.K1 > 0) return 0;
  return n_p <= 2 && p.cond[0].K1 > 0 && p.cond[0].src1 == 2;
}

int ren_tc_launch_raygrad(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                          const float* o_dot, const float* d_dot, const float* jitter, const float* params,
                          const void* delta, const void* delta_dot, const unsigned int* gmax,
                          const unsigned char* tile_live, const int* sample_idx, long long n_samples, float* g_rays,
                          cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  RgtArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p; a.n_rays = n_rays; a.K = K; a.tangent = o_dot != nullptr ? 1 : 0;
  a.delta = (const unsigned char*)delta; a.delta_dot = (const unsigned char*)delta_dot; a.gmax = gmax;
  a.tile_live = tile_live; a.sample_idx = sample_idx; a.n_samples = n_samples; a.params = params;
  a.rays_o = o; a.rays_d = d; a.rays_o_dot = o_dot; a.rays_d_dot = d_dot; a.jitter = jitter; a.g_rays = g_rays;
  if (a.tangent && delta_dot == nullptr) { ren_set_error("tc ray gradient: tangent mode needs delta_dot"); return REN_E_INVALID; }
  const long long S = sample_idx ? n_samples : n_rays * K;
  int n = 0, b_off = 0;
  auto add = [&](const RenLayer& L, int col0, int enc) {
    if (n < RGT_MAX_SRC) {
      RgtSrc& s = a.src[n];
      s.d_off = (long long)tcs_mat(p, S, L.mat); s.N = L.N; s.enc = enc; s.w_off = L.w_off + col0; s.ldw = L.K0 + L.K1;
      s.b_off = b_off;
    }
    ++n;
    b_off += (L.N >> 4) * (enc == 1 ? REN_PE_MAX : REN_DE_MAX) * 64;
  };
  for (int i = 0; i < p.D; ++i) {
    if (p.trunk[i].src0 == 1) add(p.trunk[i], 0, 1);
    if (p.trunk[i].K1 > 0) add(p.trunk[i], p.trunk[i].K0, 1);
  }
  add(p.cond[0], p.cond[0].K0, 2);
  if (n > RGT_MAX_SRC || b_off > RGT_B_BYTES) { ren_set_error("tc ray gradient: layout does not fit"); return REN_E_UNSUPPORTED; }
  a.n_src = n;
  cudaError_t e = cudaMemsetAsync(g_rays, 0, sizeof(float) * (size_t)n_rays * (a.tangent ? 12 : 6), st);
  if (e != cudaSuccess) { ren_set_error("ray gradient memset: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  if (S <= 0) return REN_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (S + TC_M - 1) / TC_M;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  auto kern = a.tangent ? ren_tc_raygrad<true> : ren_tc_raygrad<false>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RGT_SM_TOTAL);
  if (e != cudaSuccess) { ren_set_error("tc ray gradient smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  ren_prof_begin(REN_K_FIELD_DGRAD, st);
  kern<<<grid, RGT_THREADS, RGT_SM_TOTAL, st>>>(a);
  ren_prof_end(REN_K_FIELD_DGRAD, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc ray gradient launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
