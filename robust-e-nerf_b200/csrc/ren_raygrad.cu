// ren_raygrad.cu -- gradient of the loss with respect to the rays.
//
// Needed only when the supervision timestamps are trainable (refractory period,
// models/event_generation_params.py:170-203): the reference's autograd continues from the MLP
// input through SinusoidalEncoder (external/mlp.py:227-243), contract_input_space
// (external/mlp.py:321-335), the sample position o + d (t0+t1)/2 (external/utils.py:87) and the
// direction scaling d*pi (external/mlp.py:353-355) to (o, d), then through LinearTrajectory to t.
//
// Input: the pre-activation gradients Delta of every layer that consumes an encoding (fp32 blocked
// delta workspace of the SIMT engine).  Per row: dEnc = Delta W[:, encoding columns]; value mode
// reduces it to d loss / d (o, d); tangent mode (rows = (value, d/dw) pairs) also needs the second
// derivative of the contraction, taken with hyper-dual numbers, and returns
// d loss / d (o, d, o_dot, d_dot).
#include <stdlib.h>
#include <string.h>

#include "ren_tc.cuh"
#include "ren_dual.cuh"

#define RG_ROWS 64
#define RG_THREADS 256
#define RG_MAX_SRC 8

struct RgSource {
  const void* D;      // Delta matrix of the value rows (blocked [rows x N], fp32 or fp16) or graw (row-major fp32)
  const void* D_dot;  // tcgen05 tangent render: Delta of the d/dw rows in its own matrix (else NULL: the fp32
                      // engine interleaves (value, d/dw) rows in D)
  int N;              // rows of W feeding (== width of D)
  int ldd;            // blocked: N; row-major: 1 + C
  int rowmajor;
  const float* W;     // params + w_off + column offset of the encoding segment
  int ldw;
  int enc;            // 1 positional, 2 direction
};
struct RgArgs {
  RenPlan p;
  RgSource src[RG_MAX_SRC];
  int n_src;
  long long n_rays;
  int K;
  int tangent;
  int fp16;                    // Delta matrices are fp16 blocked, scaled by delta_scale(gmax) (tcgen05 engine)
  const unsigned int* gmax;
  const unsigned char* tile_live;   // per-128-sample-tile flags or NULL: dead tiles hold no data
  const int* sample_idx;            // packed samples: dense index r*K + k of sample s, or NULL
  long long n_samples;
  const float* rays_o;
  const float* rays_d;
  const float* rays_o_dot;
  const float* rays_d_dot;
  const float* jitter;
  float* g_rays;    // [n_rays][6] (value) or [n_rays][12] (tangent); arch ngp: may be NULL (grid gradient only)
  // arch ngp: the positional encoding is the grid (its reverse also scatters the grid gradient), the direction
  // encoding the spherical harmonics of the raw direction
  RenGrid grid;
  const float* table;
  float* g_table;
};

// arch ngp: reverse of one sample through grid + SH encoders.  gvP / gtP: d loss / d (grid features, their d/dw);
// gvD / gtD: same for the SH features.  Scatters the grid gradient; returns d loss / d (x_hat, x_hat_dot) and
// d loss / d (d, d_dot) of the SH branch when `rays` is set.
__device__ __forceinline__ void ngp_enc_backward(const RgArgs& a, const float* xu, const float* xudot,
                                                 const float* d, const float* dd, const float* gvP,
                                                 const float* gtP, const float* gvD, const float* gtD, bool rays,
                                                 float* gxu, float* gxud, float* gd, float* gdd) {
  const bool tan = a.tangent != 0;
  const RenGrid& G = a.grid;
  for (int c = 0; c < 3; ++c) gxu[c] = gxud[c] = gd[c] = gdd[c] = 0.f;
  for (int l = 0; l < G.n_levels; ++l) {
    const RenGridLevel lv = G.lv[l];
    RenCell cell;
    ren_grid_cell(lv, xu, cell);
    const float* gv = gvP + l * G.F;
    const float* gt = gtP + l * G.F;
    bool any = false;
    for (int j = 0; j < G.F; ++j) any = any || gv[j] != 0.f || (tan && gt[j] != 0.f);
    if (!any) continue;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      float w3[3], s3[3];
      uint32_t g3[3];
      ren_corner(cell, corner, w3, s3, g3);
      const float w = w3[0] * w3[1] * w3[2];
      const float dw[3] = {lv.scale * s3[0] * w3[1] * w3[2], lv.scale * s3[1] * w3[0] * w3[2],
                           lv.scale * s3[2] * w3[0] * w3[1]};
      const float wdot = tan ? dw[0] * xudot[0] + dw[1] * xudot[1] + dw[2] * xudot[2] : 0.f;
      const size_t e = ((size_t)lv.offset + ren_grid_index(lv, G.type, g3[0], g3[1], g3[2])) * G.F;
      float A = 0.f, B = 0.f;
      for (int j = 0; j < G.F; ++j) {
        const float t = rays ? __ldg(a.table + e + j) : 0.f;
        A = fmaf(t, gv[j], A);
        if (tan) B = fmaf(t, gt[j], B);
        atomicAdd(a.g_table + e + j, tan ? fmaf(wdot, gt[j], w * gv[j]) : w * gv[j]);
      }
      if (rays) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          gxu[k] = fmaf(dw[k], A, gxu[k]);
          if (tan) {
            gxud[k] = fmaf(dw[k], B, gxud[k]);
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (j != k) gxu[k] = fmaf(lv.scale * lv.scale * s3[k] * s3[j] * w3[3 - k - j] * xudot[j], B, gxu[k]);
          }
        }
      }
    }
  }
  if (rays) {
    const int n = a.p.DE;
    for (int kk = 0; kk < 3; ++kk) {
      HDual dh[3], out[16];
      for (int c = 0; c < 3; ++c) dh[c] = HDual(d[c], c == kk ? 1.f : 0.f, dd[c], 0.f);
      ren_sh_t<HDual>(dh, a.p.f.sh_degree, out);
      float sv = 0.f, st = 0.f;
      for (int j = 0; j < n; ++j) {
        sv = fmaf(gvD[j], out[j].d1, sv);
        if (tan) {
          sv = fmaf(gtD[j], out[j].d12, sv);
          st = fmaf(gtD[j], out[j].d1, st);
        }
      }
      gd[kk] = sv;
      gdd[kk] = st;
    }
  }
}

// d loss / d x and d loss / d xdot of one encoder from d loss / d enc (gv) and d loss / d enc_dot (gt)
__device__ __forceinline__ void enc_backward(const float* x, const float* xdot, int deg, const float* gv,
                                             const float* gt, bool tangent, float* gx, float* gxd) {
  for (int a = 0; a < 3; ++a) {
    gx[a] = gv[a];
    gxd[a] = tangent ? gt[a] : 0.f;
  }
  float sc = 1.f;
  for (int i = 0; i < deg; ++i) {
    for (int a = 0; a < 3; ++a) {
      const float xb = __fmul_rn(x[a], sc);
      const float xs = __fadd_rn(xb, 1.5707963267948966f);
      const int js = 3 + i * 3 + a, jc = 3 + deg * 3 + i * 3 + a;
      const float c0 = cosf(xb) * sc, c1 = cosf(xs) * sc;
      gx[a] = fmaf(gv[js], c0, fmaf(gv[jc], c1, gx[a]));
      if (tangent) {
        const float s0 = sinf(xb) * sc * sc, s1 = sinf(xs) * sc * sc;
        gx[a] -= (gt[js] * s0 + gt[jc] * s1) * xdot[a];
        gxd[a] = fmaf(gt[js], c0, fmaf(gt[jc], c1, gxd[a]));
      }
    }
    sc *= 2.f;
  }
}

// value of Delta[row][n] of one source: separate d/dw matrix (tcgen05) or interleaved rows (fp32 engine)
__device__ __forceinline__ float rg_load(const RgArgs& a, const RgSource& sc, long long row, int n) {
  const void* D = sc.D;
  long long mrow = row;
  if (a.tangent && sc.D_dot != nullptr) {
    D = (row & 1) ? sc.D_dot : sc.D;
    mrow = row >> 1;
  }
  if (sc.rowmajor) return reinterpret_cast<const float*>(D)[mrow * sc.ldd + n];
  if (a.fp16) return __half2float(reinterpret_cast<const __half*>(D)[ren_blk(mrow, n, sc.ldd)]);
  return reinterpret_cast<const float*>(D)[ren_blk(mrow, n, sc.ldd)];
}

// Tile = 64 rows.  Phase 1: dEnc[64 x (64 | 32)] = sum over sources of Delta[64 x N] W[N x enc cols] as a
// register-blocked fp32 GEMM (thread = 4 rows x 4 positional + 2 direction columns, K chunks of 16 staged in
// shared memory).  Phase 2: one thread per sample reduces dEnc through the encoder and the contraction.
extern "C" __global__ void __launch_bounds__(RG_THREADS) ren_raygrad_kernel(const RgArgs a) {
  __shared__ __align__(16) float gP[RG_ROWS][REN_PE_MAX];
  __shared__ __align__(16) float gD[RG_ROWS][REN_DE_MAX];
  __shared__ __align__(16) float As[16][RG_ROWS];        // Delta chunk, k-major
  __shared__ __align__(16) float Ws[16][64];             // weight chunk [k][enc col]
  const RenPlan& p = a.p;
  const long long S = (a.sample_idx ? a.n_samples : a.n_rays * a.K) * (a.tangent ? 2 : 1);   // rows
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;                // rows 4 ty .. 4 ty + 3; columns 4 tx .. (P), 2 tx .. (D)
  const float unscale = a.fp16 ? 1.f / delta_scale(a.gmax) : 1.f;
  for (long long row0 = (long long)blockIdx.x * RG_ROWS; row0 < S; row0 += (long long)gridDim.x * RG_ROWS) {
    if (a.tile_live != nullptr && !__ldg(a.tile_live + ((a.tangent ? (row0 >> 1) : row0) >> 7))) continue;
    float accP[4][4], accD[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int c = 0; c < 4; ++c) accP[i][c] = 0.f;
      accD[i][0] = accD[i][1] = 0.f;
    }
    for (int si = 0; si < a.n_src; ++si) {
      const RgSource& sc = a.src[si];
      const int width = sc.enc == 1 ? p.PE : p.DE;
      const float scale = sc.rowmajor ? 1.f : unscale;
      for (int k0 = 0; k0 < sc.N; k0 += 16) {
        // stage: 64 rows x 16 k of Delta (4 per thread), 16 k x 64 columns of W (4 per thread)
        {
          const int lrow = tid >> 2, kq = (tid & 3) * 4;
          const long long row = row0 + lrow;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = k0 + kq + e;
            As[kq + e][lrow] = (row < S && n < sc.N) ? rg_load(a, sc, row, n) * scale : 0.f;
          }
          const int wk = tid >> 4, wc = (tid & 15) * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = k0 + wk, c = wc + e;
            Ws[wk][c] = (n < sc.N && c < width) ? __ldg(sc.W + (long long)n * sc.ldw + c) : 0.f;
          }
        }
        __syncthreads();
        if (sc.enc == 1) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int c = 0; c < 4; ++c) accP[i][c] = fmaf(ar[i], wr[c], accP[i][c]);
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float2 wv = *reinterpret_cast<const float2*>(&Ws[kk][tx * 2]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              accD[i][0] = fmaf(ar[i], wv.x, accD[i][0]);
              accD[i][1] = fmaf(ar[i], wv.y, accD[i][1]);
            }
          }
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(&gP[ty * 4 + i][tx * 4]) = make_float4(accP[i][0], accP[i][1], accP[i][2], accP[i][3]);
      *reinterpret_cast<float2*>(&gD[ty * 4 + i][tx * 2]) = make_float2(accD[i][0], accD[i][1]);
    }
    __syncthreads();
    const int per = a.tangent ? 2 : 1;
    if (threadIdx.x < RG_ROWS / per) {
      const long long r0 = row0 + threadIdx.x * per;
      if (r0 < S) {
        const long long smp = r0 / per;
        const long long dense = a.sample_idx ? (long long)a.sample_idx[smp] : smp;
        const long long r = dense / a.K;
        const int k = (int)(dense - r * a.K);
        float o[3], d[3], od[3] = {0.f, 0.f, 0.f}, dd[3] = {0.f, 0.f, 0.f};
        for (int c = 0; c < 3; ++c) {
          o[c] = a.rays_o[r * 3 + c];
          d[c] = a.rays_d[r * 3 + c];
          if (a.tangent) { od[c] = a.rays_o_dot[r * 3 + c]; dd[c] = a.rays_d_dot[r * 3 + c]; }
        }
        float t_min, t_max, t0, t1, z[3];
        ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
        ren_march_to(p.r, t_min, k, &t0, &t1);
        ren_field_input(p.f, o, d, t0, t1, z);
        const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
        float pos[3], xdot[3];
        for (int c = 0; c < 3; ++c) {
          pos[c] = __fadd_rn(o[c], __fmul_rn(__fmul_rn(d[c], __fadd_rn(t0, t1)), 0.5f));
          xdot[c] = fmaf(dd[c], tm, od[c]);
        }
        const float* gvP = gP[threadIdx.x * per];
        const float* gtP = gP[threadIdx.x * per + (per - 1)];
        const float* gvD = gD[threadIdx.x * per];
        const float* gtD = gD[threadIdx.x * per + (per - 1)];
        if (p.arch == REN_ARCH_NGP) {
          float xu[3], xudot[3] = {0.f, 0.f, 0.f};
          ren_unit_position(p.f, o, d, t0, t1, xu);
          if (a.tangent) {
            Dual pd[3], xd[3];
            for (int c = 0; c < 3; ++c) pd[c] = Dual(pos[c], xdot[c]);
            ren_contract_unit_t<Dual>(p.f, pd, xd);
            for (int c = 0; c < 3; ++c) xudot[c] = xd[c].d;
          }
          float gxu[3], gxud[3], gdir[3], gdird[3];
          const bool rays = a.g_rays != nullptr;
          ngp_enc_backward(a, xu, xudot, d, dd, gvP, gtP, gvD, gtD, rays, gxu, gxud, gdir, gdird);
          if (rays) {
            const int stride = a.tangent ? 12 : 6;
            float* g = a.g_rays + r * stride;
            for (int kk = 0; kk < 3; ++kk) {
              HDual ph[3], xh[3];
              for (int c = 0; c < 3; ++c) ph[c] = HDual(pos[c], c == kk ? 1.f : 0.f, xdot[c], 0.f);
              ren_contract_unit_t<HDual>(p.f, ph, xh);
              float sx = 0.f, sxd = 0.f;
              for (int c = 0; c < 3; ++c) {
                sx = fmaf(gxu[c], xh[c].d1, fmaf(gxud[c], xh[c].d12, sx));
                sxd = fmaf(gxud[c], xh[c].d1, sxd);
              }
              atomicAdd(g + kk, sx);
              atomicAdd(g + 3 + kk, fmaf(tm, sx, gdir[kk]));
              if (a.tangent) {
                atomicAdd(g + 6 + kk, sxd);
                atomicAdd(g + 9 + kk, fmaf(tm, sxd, gdird[kk]));
              }
            }
          }
        } else {
        float gz[3], gzd[3], gdz[3], gdzd[3];
        float zdot[3] = {0.f, 0.f, 0.f};
        if (a.tangent) {
          Dual pd[3], zd[3];
          for (int c = 0; c < 3; ++c) pd[c] = Dual(pos[c], xdot[c]);
          ren_contract_t<Dual>(p.f, pd, zd);
          for (int c = 0; c < 3; ++c) zdot[c] = zd[c].d;
        }
        enc_backward(z, zdot, p.f.pos_deg, gvP, gtP, a.tangent != 0, gz, gzd);
        const float dz[3] = {__fmul_rn(d[0], 3.141592653589793f), __fmul_rn(d[1], 3.141592653589793f),
                             __fmul_rn(d[2], 3.141592653589793f)};
        const float dzd[3] = {dd[0] * 3.141592653589793f, dd[1] * 3.141592653589793f, dd[2] * 3.141592653589793f};
        enc_backward(dz, dzd, p.f.dir_deg, gvD, gtD, a.tangent != 0, gdz, gdzd);
        // through the contraction: gx_k = gz . J e_k + gzd . H[e_k, xdot];  gxd_k = gzd . J e_k
        float gx[3], gxd[3];
        for (int kk = 0; kk < 3; ++kk) {
          HDual ph[3], zh[3];
          for (int c = 0; c < 3; ++c) ph[c] = HDual(pos[c], c == kk ? 1.f : 0.f, xdot[c], 0.f);
          ren_contract_t<HDual>(p.f, ph, zh);
          float sx = 0.f, sxd = 0.f;
          for (int c = 0; c < 3; ++c) {
            sx = fmaf(gz[c], zh[c].d1, fmaf(gzd[c], zh[c].d12, sx));
            sxd = fmaf(gzd[c], zh[c].d1, sxd);
          }
          gx[kk] = sx;
          gxd[kk] = sxd;
        }
        const int stride = a.tangent ? 12 : 6;
        float* g = a.g_rays + r * stride;
        for (int c = 0; c < 3; ++c) {
          atomicAdd(g + c, gx[c]);
          atomicAdd(g + 3 + c, fmaf(tm, gx[c], 3.141592653589793f * gdz[c]));
          if (a.tangent) {
            atomicAdd(g + 6 + c, gxd[c]);
            atomicAdd(g + 9 + c, fmaf(tm, gxd[c], 3.141592653589793f * gdzd[c]));
          }
        }
        }
      }
    }
    __syncthreads();
  }
}

int ren_tc_raygrad_supported(const RenPlan& p);      // ren_tc_raygrad.cu: tcgen05 version over the fp16 deltas
int ren_tc_launch_raygrad(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                          const float* o_dot, const float* d_dot, const float* jitter, const float* params,
                          const void* delta, const void* delta_dot, const unsigned int* gmax,
                          const unsigned char* tile_live, const int* sample_idx, long long n_samples, float* g_rays,
                          cudaStream_t st);

// fp32 engine: delta = fp32 blocked delta workspace (mats mirror of the stash) followed by graw [rows][1+C];
// tcgen05 engine (fp16 != 0): delta / delta_dot = fp16 blocked delta matrices of the value / d/dw rows
int ren_launch_raygrad(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                       const float* o_dot, const float* d_dot, const float* jitter, const float* params,
                       const void* delta, const void* delta_dot, int fp16, const unsigned int* gmax,
                       const unsigned char* tile_live, const int* sample_idx, long long n_samples, float* g_rays,
                       const RenNgp* ngp, cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  if (fp16 && ren_tc_raygrad_supported(p) && getenv("REN_RAYGRAD_SIMT") == nullptr)
    return ren_tc_launch_raygrad(p, n_rays, K, o, d, o_dot, d_dot, jitter, params, delta, delta_dot, gmax, tile_live,
                                 sample_idx, n_samples, g_rays, st);
  RgArgs a;
  static_assert(sizeof(RgArgs) <= 4000, "kernel parameter block too large");
  memset(&a.grid, 0, sizeof(a.grid));
  a.table = nullptr; a.g_table = nullptr;
  if (p.arch == REN_ARCH_NGP) {
    if (ngp == nullptr || ngp->g_table == nullptr) { ren_set_error("ray gradient: arch ngp without grid"); return REN_E_INVALID; }
    a.grid = ngp->grid; a.table = ngp->table; a.g_table = ngp->g_table;
  }
  a.tile_live = tile_live; a.sample_idx = sample_idx; a.n_samples = n_samples;
  a.p = p; a.n_rays = n_rays; a.K = K; a.tangent = o_dot != nullptr ? 1 : 0; a.fp16 = fp16; a.gmax = gmax;
  a.rays_o = o; a.rays_d = d; a.rays_o_dot = o_dot; a.rays_d_dot = d_dot; a.jitter = jitter; a.g_rays = g_rays;
  const long long Ns = sample_idx ? n_samples : n_rays * K;      // samples
  const long long S = Ns * (a.tangent ? 2 : 1);                  // rows
  const long long Sm = fp16 ? Ns : S;                            // rows of one delta matrix
  const float* graw = fp16 ? nullptr : (const float*)delta + p.act_floats * ren_pad128(S);
  int n = 0;
  auto add = [&](const void* D, const void* Dd, int N, int ldd, int rowmajor, const float* W, int ldw, int enc) {
    if (n < RG_MAX_SRC) {
      RgSource& s = a.src[n];
      s.D = D; s.D_dot = Dd; s.N = N; s.ldd = ldd; s.rowmajor = rowmajor; s.W = W; s.ldw = ldw; s.enc = enc;
    }
    ++n;
  };
  auto mat = [&](const void* base, int m) -> const void* {
    if (base == nullptr) return nullptr;
    return fp16 ? (const void*)((const char*)base + tcs_mat(p, Sm, m))
                : (const void*)((const float*)base + ren_stash_mat(p, S, m));
  };
  auto add_layer = [&](const RenLayer& L) {
    const int ldw = L.K0 + L.K1;
    if (L.src0 == 1) add(mat(delta, L.mat), mat(delta_dot, L.mat), L.N, L.N, 0, params + L.w_off, ldw, 1);
    if (L.K1 > 0)
      add(mat(delta, L.mat), mat(delta_dot, L.mat), L.N, L.N, 0, params + L.w_off + L.K0, ldw, L.src1 == 1 ? 1 : 2);
  };
  for (int i = 0; i < p.D; ++i) add_layer(p.trunk[i]);
  add_layer(p.bott);
  for (int jj = 0; jj < p.CD; ++jj) add_layer(p.cond[jj]);
  const int C1 = 1 + p.C;
  if (!fp16) {      // head layers that consume an encoding directly: shapes the tcgen05 engine does not cover
    if (p.D == 0) add(graw, nullptr, 1, C1, 1, params + p.sig_w, p.hid_K0 + p.hid_K1, 1);
    else if (p.hid_K1 > 0) add(graw, nullptr, 1, C1, 1, params + p.sig_w + p.hid_K0, p.hid_K0 + p.hid_K1, 1);
    if (p.CD == 0) add(graw + 1, nullptr, p.C, C1, 1, params + p.rgb_w + p.W, p.W + p.DE, 2);
  }
  if (n > RG_MAX_SRC) { ren_set_error("ray gradient: too many encoding consumers (%d)", n); return REN_E_UNSUPPORTED; }
  a.n_src = n;
  cudaError_t e = cudaSuccess;
  if (g_rays != nullptr) e = cudaMemsetAsync(g_rays, 0, sizeof(float) * (size_t)n_rays * (a.tangent ? 12 : 6), st);
  if (e != cudaSuccess) { ren_set_error("ray gradient memset: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  long long blocks = (S + RG_ROWS - 1) / RG_ROWS;
  if (blocks > 148 * 32) blocks = 148 * 32;
  ren_prof_begin(REN_K_FIELD_DGRAD, st);
  ren_raygrad_kernel<<<(unsigned)blocks, RG_THREADS, 0, st>>>(a);
  ren_prof_end(REN_K_FIELD_DGRAD, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ray gradient launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
