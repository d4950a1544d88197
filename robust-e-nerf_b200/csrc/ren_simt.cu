// ren_simt.cu -- fp32 FFMA engine of the field MLP (REN_ENGINE_SIMT_FP32).
//
// Parity baseline and any-shape fallback of the tcgen05 engine: one CTA = one tile of
// 64 consecutive samples (ray-major), activations resident in shared memory, weights
// streamed from L2 in 16-wide K chunks, 8x8 register blocks per thread.
//   ren_simt_field_fwd   march -> position -> contraction -> sin/cos encoding -> trunk
//                        (+skip concat) -> sigma head, bottleneck, view-conditioned head
//                        (external/utils.py:83-96, external/mlp.py:99-113,193-205,350-358)
//   ren_simt_field_dgrad reverse pass through the same layers (activation stash in HBM)
//   ren_simt_wgrad       dW = Delta^T H split over samples, db = colsum(Delta)
//
// Tangent mode (TAN): the tile holds 32 samples x {value row, d/dw row} (row = 2*sample + kind), so the
// same GEMMs push the activations AND their derivative along the ray's pose-interpolation weight
// through the network (the reference gets d logI / d t by autograd with create_graph=True,
// models/robust_e_nerf.py:394-398); stash / delta matrices simply have 2S rows.
#include <string.h>

#include "ren_common.cuh"
#include "ren_dual.cuh"

#define TM 64          // rows per tile
#define KC 16          // K chunk
#define NTHREADS 256

struct FwdArgs {
  RenPlan p;
  long long n_rays;
  int K;
  const float* rays_o;
  const float* rays_d;
  const float* jitter;
  const float* params;
  float* stash;     // full stash (training) or NULL
  float* raw;       // [S][1+C] (inside stash when training, else workspace); tangent: [2S][1+C]
  const float* rays_o_dot;   // tangent mode: d o / d w, d d / d w per ray
  const float* rays_d_dot;
  RenGrid grid;              // arch ngp: grid encoding of the contracted position instead of the sinusoidal one
  const float* table;
};

struct DgradArgs {
  RenPlan p;
  long long S;
  const float* params;
  const float* stash;
  const float* graw;  // [S][1+C]   (S = rows: 2 per sample in tangent mode)
  float* delta;       // mats mirror of stash: Delta (pre-activation grads)
};

// ---------------------------------------------------------------- tile GEMM core
// acc[i][j] += sum_k In[ty*8+i][k] * Wt[k][col(j)] for one input segment, K padded to 16.
// FWDW: Wt[k][n] = W[(n)*ldw + k]  (W rows = outputs;   forward)
// else: Wt[k][n] = W[(k)*ldw + n]  (W rows = reduction; dgrad)
// Ktrue: valid reduction length, Ntrue: valid outputs (<=256).  In must be finite on the
// padded K range (weights are zero-filled there).
template <bool FWDW>
__device__ __forceinline__ void gemm_segment(float (&acc)[8][8], const float* __restrict__ In,
                                             int in_stride, const float* __restrict__ W, int ldw,
                                             int Ktrue, int Ntrue, float (*Ws)[KC][256]) {
  const int tid = threadIdx.x;
  const int ty = tid >> 5, tx = tid & 31;
  const int nchunks = (Ktrue + KC - 1) / KC;
  float wreg[KC];
  auto fetch = [&](int c) {
    int k0 = c * KC;
    if (FWDW) {
      int n = tid;
      const float* row = W + (long long)n * ldw + k0;
#pragma unroll
      for (int kk = 0; kk < KC; ++kk)
        wreg[kk] = (n < Ntrue && k0 + kk < Ktrue) ? __ldg(row + kk) : 0.f;
    } else {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk)
        wreg[kk] = (tid < Ntrue && k0 + kk < Ktrue) ? __ldg(W + (long long)(k0 + kk) * ldw + tid) : 0.f;
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) Ws[buf][kk][tid] = wreg[kk];
  };
  fetch(0);
  commit(0);
  __syncthreads();
  const bool hi_half = Ntrue > 128;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) fetch(c + 1);
    const float* in = In + (ty * 8) * in_stride + c * KC;
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; ++k4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        a[i] = *reinterpret_cast<const float4*>(in + i * in_stride + k4 * 4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k4 * 4 + kk][tx * 4]);
        float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hi_half) w1 = *reinterpret_cast<const float4*>(&Ws[buf][k4 * 4 + kk][128 + tx * 4]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, w0.x, acc[i][0]);
          acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]);
          acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          if (hi_half) {
            acc[i][4] = fmaf(av, w1.x, acc[i][4]);
            acc[i][5] = fmaf(av, w1.y, acc[i][5]);
            acc[i][6] = fmaf(av, w1.z, acc[i][6]);
            acc[i][7] = fmaf(av, w1.w, acc[i][7]);
          }
        }
      }
    }
    if (c + 1 < nchunks) commit(buf ^ 1);
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// Shared-memory carve-up (floats).  WS = activation row stride (max width, mult of 16).
struct Smem {
  float* bufA;   // [TM][WS]
  float* bufB;   // [TM][WS]
  float* encP;   // [TM][64]
  float* encD;   // [TM][32]
  float (*Ws)[KC][256];  // [2][KC][256]
  float* misc;   // [TM*8]
};
__device__ __forceinline__ Smem carve(float* base, int WS) {
  Smem s;
  s.bufA = base;
  s.bufB = s.bufA + TM * WS;
  s.encP = s.bufB + TM * WS;
  s.encD = s.encP + TM * REN_PE_MAX;
  s.Ws = reinterpret_cast<float (*)[KC][256]>(s.encD + TM * REN_DE_MAX);
  s.misc = reinterpret_cast<float*>(s.Ws) + 2 * KC * 256;
  return s;
}
static inline size_t simt_smem_bytes(int WS) {
  return sizeof(float) * ((size_t)2 * TM * WS + TM * REN_PE_MAX + TM * REN_DE_MAX + 2 * KC * 256 + TM * 8);
}

// One dense layer on the resident tile: Out = act(W [In0 | In1] + b); optional HBM stash.
// TAN: odd rows carry the tangent: out = act'(pre of the even row) * (W in), no bias.
template <bool TAN>
__device__ __forceinline__ void fwd_layer(const RenPlan& p, const RenLayer& L, const float* params,
                                          const Smem& s, const float* in_main, float* out, int WS,
                                          float* stash_mat, long long s0, long long S) {
  const int tid = threadIdx.x;
  const int ty = tid >> 5, tx = tid & 31;
  const int ldw = L.K0 + L.K1;
  for (int n0 = 0; n0 < L.N; n0 += 256) {
    const int Nt = min(256, L.N - n0);
    float acc[8][8];
    zero_acc(acc);
    const float* in0 = L.src0 == 1 ? s.encP : in_main;
    const int st0 = L.src0 == 1 ? REN_PE_MAX : WS;
    const float* Wn = params + L.w_off + (long long)n0 * ldw;
    gemm_segment<true>(acc, in0, st0, Wn, ldw, L.K0, Nt, s.Ws);
    if (L.K1 > 0) {
      const float* in1 = L.src1 == 1 ? s.encP : s.encD;
      const int st1 = L.src1 == 1 ? REN_PE_MAX : REN_DE_MAX;
      gemm_segment<true>(acc, in1, st1, Wn + L.K0, ldw, L.K1, Nt, s.Ws);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 128 + tx * 4;
      if (n < L.N) {  // N is a multiple of 4
        const float* bp = params + L.b_off + n;   // flat-buffer offsets are not 16 B aligned
        const float4 b = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
        float4 dact = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 v;
          if (TAN && (i & 1)) {
            v = make_float4(acc[i][h * 4 + 0] * dact.x, acc[i][h * 4 + 1] * dact.y,
                            acc[i][h * 4 + 2] * dact.z, acc[i][h * 4 + 3] * dact.w);
          } else {
            v = make_float4(acc[i][h * 4 + 0] + b.x, acc[i][h * 4 + 1] + b.y,
                            acc[i][h * 4 + 2] + b.z, acc[i][h * 4 + 3] + b.w);
            if (L.act) {
              if (TAN) {
                dact.x = ren_hidden_grad_from_pre(p.f.hidden_activation, v.x);
                dact.y = ren_hidden_grad_from_pre(p.f.hidden_activation, v.y);
                dact.z = ren_hidden_grad_from_pre(p.f.hidden_activation, v.z);
                dact.w = ren_hidden_grad_from_pre(p.f.hidden_activation, v.w);
              }
              v.x = ren_hidden_act(p.f.hidden_activation, v.x);
              v.y = ren_hidden_act(p.f.hidden_activation, v.y);
              v.z = ren_hidden_act(p.f.hidden_activation, v.z);
              v.w = ren_hidden_act(p.f.hidden_activation, v.w);
            }
          }
          const int m = ty * 8 + i;
          *reinterpret_cast<float4*>(out + m * WS + n) = v;
          if (stash_mat != nullptr && s0 + m < S)
            *reinterpret_cast<float4*>(stash_mat + ren_blk(s0 + m, n, L.N)) = v;
        }
      }
    }
  }
  __syncthreads();
}

// out[m] = b + sum_k [in0 | in1][m][k] * w[k], 4 threads per row, bank-skewed reads.
__device__ __forceinline__ float row_dot(const float* in0, int st0, int K0, const float* in1, int st1,
                                         int K1, const float* __restrict__ w) {
  const int tid = threadIdx.x;
  const int m = tid >> 2, part = tid & 3;
  float acc = 0.f;
  const int q0 = K0 >> 2;  // K0 % 4 == 0
  for (int j = 0; j < q0; ++j) {
    int k = (((j + m) % q0) << 2) + part;
    acc = fmaf(in0[m * st0 + k], __ldg(w + k), acc);
  }
  if (K1 > 0) {
    const int q1 = (K1 + 3) >> 2;
    for (int j = 0; j < q1; ++j) {
      int k = (((j + m) % q1) << 2) + part;
      if (k < K1) acc = fmaf(in1[m * st1 + k], __ldg(w + K0 + k), acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  return acc;
}

template <bool TAN>
__global__ void __launch_bounds__(NTHREADS, 1) ren_simt_field_fwd(const FwdArgs a) {
  extern __shared__ __align__(16) float smem_f[];
  const RenPlan& p = a.p;
  const int WS = max(p.W, p.WC);
  Smem s = carve(smem_f, WS);
  const long long S = a.n_rays * a.K * (TAN ? 2 : 1);     // rows
  const long long n_tiles = (S + TM - 1) / TM;
  const int tid = threadIdx.x;
  float* sel = s.misc;  // [TM]
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long s0 = tile * TM;
    // ---- prologue: one thread per row
    if (tid < TM) {
      const long long sidx = s0 + tid;     // row
      float* eP = s.encP + tid * REN_PE_MAX;
      float* eD = s.encD + tid * REN_DE_MAX;
      bool selector = false, skipped = false;
      if (sidx < S) {
        const long long smp = TAN ? (sidx >> 1) : sidx;
        const bool tan_row = TAN && (sidx & 1);
        const long long r = smp / a.K;
        const int k = (int)(smp - r * a.K);
        float o[3], d[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          o[c] = a.rays_o[r * 3 + c];
          d[c] = a.rays_d[r * 3 + c];
        }
        float t_min, t_max, t0, t1, z[3];
        ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
        ren_march_to(p.r, t_min, k, &t0, &t1);
        selector = ren_field_input(p.f, o, d, t0, t1, z);
        if (!ren_alive(t0, t1, t_max)) selector = false;  // dead tail sample: masked anyway
        skipped = !ren_occupied(p.f, p.r, o, d, t0, t1);  // unoccupied cell: the marcher emits no sample
        float dz[3] = {__fmul_rn(d[0], 3.141592653589793f), __fmul_rn(d[1], 3.141592653589793f),
                       __fmul_rn(d[2], 3.141592653589793f)};
        if (p.arch == REN_ARCH_NGP) {
          // external/ngp.py:219-252: grid features of the contracted unit position, spherical harmonics of the
          // raw direction; the d/dw row takes the derivative of both along (odot + ddot tm, ddot)
          float xu[3];
          ren_unit_position(p.f, o, d, t0, t1, xu);
          for (int j = 0; j < REN_PE_MAX; ++j) eP[j] = 0.f;
          for (int j = 0; j < REN_DE_MAX; ++j) eD[j] = 0.f;
          // the grid features are filled in by all 256 threads after this block (4 threads per row): leave the
          // unit position (and its derivative along w) in shared memory
          float* gx = s.misc + TM + tid * 7;
          gx[0] = xu[0]; gx[1] = xu[1]; gx[2] = xu[2];
          gx[3] = 0.f; gx[4] = 0.f; gx[5] = 0.f;
          gx[6] = tan_row ? 2.f : 1.f;
          if (!tan_row) {
            float sh[16];
            ren_sh_t<float>(d, p.f.sh_degree, sh);
            for (int j = 0; j < p.DE; ++j) eD[j] = sh[j];
          } else {
            const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
            Dual pd[3], xd[3], dd3[3], shd[16];
            float xdot[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float od = a.rays_o_dot[r * 3 + c], ddc = a.rays_d_dot[r * 3 + c];
              pd[c] = Dual(__fadd_rn(o[c], __fmul_rn(__fmul_rn(d[c], __fadd_rn(t0, t1)), 0.5f)), fmaf(ddc, tm, od));
              dd3[c] = Dual(d[c], ddc);
            }
            ren_contract_unit_t<Dual>(p.f, pd, xd);
#pragma unroll
            for (int c = 0; c < 3; ++c) { xdot[c] = xd[c].d; gx[3 + c] = xdot[c]; }
            ren_sh_t<Dual>(dd3, p.f.sh_degree, shd);
            for (int j = 0; j < p.DE; ++j) eD[j] = shd[j].d;
          }
        } else
        if (!tan_row) {
          ren_encode(z, p.f.pos_deg, REN_PE_MAX, [&](int j, float v) { eP[j] = v; });
          ren_encode(dz, p.f.dir_deg, REN_DE_MAX, [&](int j, float v) { eD[j] = v; });
        } else {
          // x = o + d tm, xdot = odot + ddot tm (t0, t1 carry no gradient: nerfacc marches under
          // no_grad); zdot = J_contract(x) xdot by dual numbers
          const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
          Dual pd[3], zd[3];
          float zdot[3], ddz[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float od = a.rays_o_dot[r * 3 + c], dd = a.rays_d_dot[r * 3 + c];
            pd[c] = Dual(__fadd_rn(o[c], __fmul_rn(__fmul_rn(d[c], __fadd_rn(t0, t1)), 0.5f)), fmaf(dd, tm, od));
            ddz[c] = dd * 3.141592653589793f;
          }
          ren_contract_t<Dual>(p.f, pd, zd);
#pragma unroll
          for (int c = 0; c < 3; ++c) zdot[c] = zd[c].d;
          ren_encode_tangent(z, zdot, p.f.pos_deg, REN_PE_MAX, [&](int j, float v) { eP[j] = v; });
          ren_encode_tangent(dz, ddz, p.f.dir_deg, REN_DE_MAX, [&](int j, float v) { eD[j] = v; });
        }
        if (a.stash) {
          float* gP = a.stash + ren_stash_encp(p, S);
          float* gD = a.stash + ren_stash_encd(p, S);
          if (p.arch != REN_ARCH_NGP)
            for (int j = 0; j < REN_PE_MAX; ++j) gP[ren_blk(sidx, j, REN_PE_MAX)] = eP[j];
          for (int j = 0; j < REN_DE_MAX; ++j) gD[ren_blk(sidx, j, REN_DE_MAX)] = eD[j];
        }
      } else {
        for (int j = 0; j < REN_PE_MAX; ++j) eP[j] = 0.f;
        for (int j = 0; j < REN_DE_MAX; ++j) eD[j] = 0.f;
        if (p.arch == REN_ARCH_NGP) s.misc[TM + tid * 7 + 6] = 0.f;
      }
      sel[tid] = skipped ? 2.f : (selector ? 1.f : 0.f);
    }
    __syncthreads();
    if (p.arch == REN_ARCH_NGP) {
      // grid features: row = tid & 63, levels (tid >> 6), (tid >> 6) + 4, ...  (128 gathers per row spread over 4 threads)
      const int row = tid & (TM - 1), lg = tid >> 6;
      const float* gx = s.misc + TM + row * 7;
      const float mode = gx[6];
      if (mode != 0.f) {
        const float xu[3] = {gx[0], gx[1], gx[2]}, xd[3] = {gx[3], gx[4], gx[5]};
        float* eP = s.encP + row * REN_PE_MAX;
        float* gP = a.stash ? a.stash + ren_stash_encp(p, S) : nullptr;
        const long long sidx = s0 + row;
        ren_grid_encode(a.grid, a.table, xu, mode == 2.f ? xd : nullptr, [&](int j, float v) {
          eP[j] = v;
          if (gP) gP[ren_blk(sidx, j, REN_PE_MAX)] = v;
        }, lg, NTHREADS / TM);
        if (gP && lg == 0)
          for (int j = p.PE; j < REN_PE_MAX; ++j) gP[ren_blk(sidx, j, REN_PE_MAX)] = 0.f;
      }
      __syncthreads();
    }
    // ---- trunk
    float* cur = s.bufA;
    float* nxt = s.bufB;
    for (int i = 0; i < p.D; ++i) {
      const RenLayer& L = p.trunk[i];
      fwd_layer<TAN>(p, L, a.params, s, cur, nxt, WS,
                     a.stash ? a.stash + ren_stash_mat(p, S, L.mat) : nullptr, s0, S);
      float* t = cur; cur = nxt; nxt = t;
    }
    // ---- sigma head (per-row dot product on the hidden features)
    {
      const float* w = a.params + p.sig_w;
      float v = row_dot(p.D > 0 ? cur : s.encP, p.D > 0 ? WS : REN_PE_MAX, p.hid_K0, s.encP,
                        REN_PE_MAX, p.hid_K1, w);
      const int m = tid >> 2;
      if ((tid & 3) == 0 && s0 + m < S) {
        if (TAN && (m & 1)) {
          a.raw[(s0 + m) * (1 + p.C)] = v;                 // d sigma_raw / d w (no bias)
        } else {
          v += a.params[p.sig_b];
          // -inf: outside the contracted box (sigma = 0, still a sample); NaN: no sample at all
          a.raw[(s0 + m) * (1 + p.C)] = sel[m] == 2.f ? __int_as_float(0x7fc00000) : (sel[m] != 0.f ? v : -INFINITY);
        }
      }
    }
    // ---- bottleneck (linear) + view-conditioned hidden layers
    fwd_layer<TAN>(p, p.bott, a.params, s, cur, nxt, WS,
                   a.stash ? a.stash + ren_stash_mat(p, S, p.bott.mat) : nullptr, s0, S);
    { float* t = cur; cur = nxt; nxt = t; }
    for (int j = 0; j < p.CD; ++j) {
      const RenLayer& L = p.cond[j];
      fwd_layer<TAN>(p, L, a.params, s, cur, nxt, WS,
                     a.stash ? a.stash + ren_stash_mat(p, S, L.mat) : nullptr, s0, S);
      float* t = cur; cur = nxt; nxt = t;
    }
    // ---- radiance head
    for (int c = 0; c < p.C; ++c) {
      const int Kin = p.CD > 0 ? p.WC : p.W;
      const float* w = a.params + p.rgb_w + c * (p.CD > 0 ? Kin : Kin + p.DE);
      float v = row_dot(cur, WS, Kin, s.encD, REN_DE_MAX, p.CD > 0 ? 0 : p.DE, w);
      const int m = tid >> 2;
      if ((tid & 3) == 0 && s0 + m < S)
        a.raw[(s0 + m) * (1 + p.C) + 1 + c] = (TAN && (m & 1)) ? v : v + a.params[p.rgb_b + c];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ dgrad
// One reverse GEMM: dIn[m][k] = sum_n Dout[m][n] W[n][k] for k < Kout (first input segment),
// optionally += gs[m] * wsig[k]; then *= act'(a_prev[m][k]) (a_prev from the stash) -> out + HBM.
// TAN: rows come in (value, tangent) pairs with adjoints (abar, abar_dot); through y = act(x),
// ydot = act'(x) xdot:  d/dx = abar act'(x) + abar_dot act''(x) xdot,  d/dxdot = abar_dot act'(x).
template <bool TAN>
__device__ __forceinline__ void bwd_layer(const RenPlan& p, const RenLayer& L, const float* params,
                                          const Smem& s, const float* dout, float* din, int WS,
                                          const float* gs, const float* wsig,
                                          const float* aprev_mat, int prev_act, float* delta_mat,
                                          int Kout, long long s0, long long S) {
  const int tid = threadIdx.x;
  const int ty = tid >> 5, tx = tid & 31;
  const int ldw = L.K0 + L.K1;
  for (int k0 = 0; k0 < Kout; k0 += 256) {
    const int Kt = min(256, Kout - k0);
    float acc[8][8];
    zero_acc(acc);
    gemm_segment<false>(acc, dout, WS, params + L.w_off + k0, ldw, L.N, Kt, s.Ws);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k0 + h * 128 + tx * 4;
      if (k < Kout) {
        float4 ws4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wsig) ws4 = make_float4(__ldg(wsig + k), __ldg(wsig + k + 1), __ldg(wsig + k + 2), __ldg(wsig + k + 3));
        float4 vprev = make_float4(0.f, 0.f, 0.f, 0.f), sprev = vprev, yprev = vprev;
#pragma unroll
        for (int i = 7; i >= 0; --i) {       // descending: the tangent row (odd) of a pair comes first
          const int m = ty * 8 + i;
          float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2],
                                 acc[i][h * 4 + 3]);
          if (wsig) {
            const float g = gs[m];
            v.x = fmaf(g, ws4.x, v.x); v.y = fmaf(g, ws4.y, v.y);
            v.z = fmaf(g, ws4.z, v.z); v.w = fmaf(g, ws4.w, v.w);
          }
          if (prev_act) {
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + m < S) y = *reinterpret_cast<const float4*>(aprev_mat + ren_blk(s0 + m, k, Kout));
            if (TAN && (i & 1)) {
              // tangent row: y = ydot; the derivative factor comes from the value row below it
              float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (s0 + m - 1 < S) yv = *reinterpret_cast<const float4*>(aprev_mat + ren_blk(s0 + m - 1, k, Kout));
              sprev = make_float4(ren_hidden_grad_from_out(p.f.hidden_activation, yv.x),
                                  ren_hidden_grad_from_out(p.f.hidden_activation, yv.y),
                                  ren_hidden_grad_from_out(p.f.hidden_activation, yv.z),
                                  ren_hidden_grad_from_out(p.f.hidden_activation, yv.w));
              vprev = v;                     // abar_dot
              yprev = make_float4(ren_hidden_curv_from_out(p.f.hidden_activation, yv.x, y.x),
                                  ren_hidden_curv_from_out(p.f.hidden_activation, yv.y, y.y),
                                  ren_hidden_curv_from_out(p.f.hidden_activation, yv.z, y.z),
                                  ren_hidden_curv_from_out(p.f.hidden_activation, yv.w, y.w));
              v.x *= sprev.x; v.y *= sprev.y; v.z *= sprev.z; v.w *= sprev.w;
            } else if (TAN) {
              v.x = fmaf(vprev.x, yprev.x, v.x * sprev.x);
              v.y = fmaf(vprev.y, yprev.y, v.y * sprev.y);
              v.z = fmaf(vprev.z, yprev.z, v.z * sprev.z);
              v.w = fmaf(vprev.w, yprev.w, v.w * sprev.w);
            } else {
              v.x *= ren_hidden_grad_from_out(p.f.hidden_activation, y.x);
              v.y *= ren_hidden_grad_from_out(p.f.hidden_activation, y.y);
              v.z *= ren_hidden_grad_from_out(p.f.hidden_activation, y.z);
              v.w *= ren_hidden_grad_from_out(p.f.hidden_activation, y.w);
            }
          }
          if (s0 + m >= S) v = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(din + m * WS + k) = v;
          if (delta_mat != nullptr && s0 + m < S)
            *reinterpret_cast<float4*>(delta_mat + ren_blk(s0 + m, k, Kout)) = v;
        }
      }
    }
  }
  __syncthreads();
}

template <bool TAN>
__global__ void __launch_bounds__(NTHREADS, 1) ren_simt_field_dgrad(const DgradArgs a) {
  extern __shared__ __align__(16) float smem_f[];
  const RenPlan& p = a.p;
  const int WS = max(p.W, p.WC);
  Smem s = carve(smem_f, WS);
  const long long S = a.S;
  const long long n_tiles = (S + TM - 1) / TM;
  const int tid = threadIdx.x;
  float* gsig = s.misc;          // [TM]
  float* gc = s.misc + TM;       // [TM][C] (C <= 4)
  const int C1 = 1 + p.C;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long s0 = tile * TM;
    if (tid < TM) {
      const long long sidx = s0 + tid;
      gsig[tid] = sidx < S ? a.graw[sidx * C1] : 0.f;
      for (int c = 0; c < p.C; ++c) gc[tid * 4 + c] = sidx < S ? a.graw[sidx * C1 + 1 + c] : 0.f;
    }
    __syncthreads();
    float* cur = s.bufA;
    float* nxt = s.bufB;
    // ---- radiance head reverse: d g_last = sum_c gc * W_o[c][:], times act'(g_last)
    {
      const int Kin = p.CD > 0 ? p.WC : p.W;
      const int ldo = p.CD > 0 ? Kin : Kin + p.DE;
      const RenLayer* Llast = p.CD > 0 ? &p.cond[p.CD - 1] : &p.bott;
      const float* amat = a.stash + ren_stash_mat(p, S, Llast->mat);
      float* dmat = a.delta + ren_stash_mat(p, S, Llast->mat);
      if (!TAN) {
        for (int e = tid; e < TM * Kin; e += NTHREADS) {
          const int m = e / Kin, k = e - m * Kin;
          float v = 0.f;
          for (int c = 0; c < p.C; ++c) v = fmaf(gc[m * 4 + c], __ldg(a.params + p.rgb_w + c * ldo + k), v);
          if (s0 + m < S) {
            if (Llast->act) v *= ren_hidden_grad_from_out(p.f.hidden_activation, amat[ren_blk(s0 + m, k, Kin)]);
            dmat[ren_blk(s0 + m, k, Kin)] = v;
          } else {
            v = 0.f;
          }
          cur[m * WS + k] = v;
        }
      } else {
        for (int e = tid; e < (TM / 2) * Kin; e += NTHREADS) {
          const int m = 2 * (e / Kin), k = e % Kin;      // value row m, tangent row m + 1
          float v = 0.f, vt = 0.f;
          for (int c = 0; c < p.C; ++c) {
            const float w = __ldg(a.params + p.rgb_w + c * ldo + k);
            v = fmaf(gc[m * 4 + c], w, v);
            vt = fmaf(gc[(m + 1) * 4 + c], w, vt);
          }
          if (s0 + m + 1 < S) {
            if (Llast->act) {
              const float y = amat[ren_blk(s0 + m, k, Kin)], yd = amat[ren_blk(s0 + m + 1, k, Kin)];
              const float sg = ren_hidden_grad_from_out(p.f.hidden_activation, y);
              v = fmaf(vt, ren_hidden_curv_from_out(p.f.hidden_activation, y, yd), v * sg);
              vt *= sg;
            }
            dmat[ren_blk(s0 + m, k, Kin)] = v;
            dmat[ren_blk(s0 + m + 1, k, Kin)] = vt;
          } else {
            v = 0.f; vt = 0.f;
          }
          cur[m * WS + k] = v;
          cur[(m + 1) * WS + k] = vt;
        }
      }
      __syncthreads();
    }
    // ---- condition layers reverse (cond[j] input main segment = cond[j-1] output or bott)
    for (int j = p.CD - 1; j >= 0; --j) {
      const RenLayer& L = p.cond[j];
      const RenLayer& Lp = j > 0 ? p.cond[j - 1] : p.bott;
      bwd_layer<TAN>(p, L, a.params, s, cur, nxt, WS, nullptr, nullptr,
                     a.stash + ren_stash_mat(p, S, Lp.mat), Lp.act,
                     a.delta + ren_stash_mat(p, S, Lp.mat), L.K0, s0, S);
      float* t = cur; cur = nxt; nxt = t;
    }
    // ---- bottleneck reverse + sigma head, into the last trunk layer's pre-activation grad
    if (p.D > 0) {
      const RenLayer& Lp = p.trunk[p.D - 1];
      bwd_layer<TAN>(p, p.bott, a.params, s, cur, nxt, WS, gsig, a.params + p.sig_w,
                     a.stash + ren_stash_mat(p, S, Lp.mat), Lp.act,
                     a.delta + ren_stash_mat(p, S, Lp.mat), p.hid_K0, s0, S);
      float* t = cur; cur = nxt; nxt = t;
      for (int i = p.D - 1; i >= 1; --i) {
        const RenLayer& L = p.trunk[i];
        const RenLayer& Lq = p.trunk[i - 1];
        bwd_layer<TAN>(p, L, a.params, s, cur, nxt, WS, nullptr, nullptr,
                       a.stash + ren_stash_mat(p, S, Lq.mat), Lq.act,
                       a.delta + ren_stash_mat(p, S, Lq.mat), L.K0, s0, S);
        float* t2 = cur; cur = nxt; nxt = t2;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ wgrad
// dW[n][koff + k] += sum_m D[m][n] * H[m][k];  db[n] += sum_m D[m][n] (k-tile 0 only).
#define WG_MAXP 48
struct WgradProblem {
  const float* D; int ldd;   // N valid columns; blocked [S x ldd] (ren_blk) or row-major [S][ldd]
  const float* H; int ldh;   // Kc valid columns; blocked [S x ldh]
  int d_rowmajor;            // D is the row-major graw matrix
  int N, Kc;
  float* dW; int ldw;        // dW + koff already applied
  float* db;                 // NULL to skip
  int ntn, ntk;              // tiles along n / k (128 each)
};
struct WgradArgs {
  WgradProblem pr[WG_MAXP];
  int tile_prefix[WG_MAXP + 1];
  int n_problems;
  int splits;
  long long S;
  int tangent;               // rows alternate (value, tangent): only value rows feed the bias gradient
};
#define WG_MC 16

// NT: outputs per thread along n and k (8: 128 x 128 tiles; 4: 64 x 64 tiles for narrow networks such as the
// 64-wide MLPs of `arch: ngp`, where a 128 x 128 tile would be 7/8 padding)
template <int NT>
__global__ void __launch_bounds__(NTHREADS, 2) ren_simt_wgrad(const WgradArgs a) {
  constexpr int TW = 16 * NT;                 // tile width along n and along k
  constexpr int PER = WG_MC * TW / NTHREADS;  // staged elements per thread and matrix
  __shared__ __align__(16) float Ds[2][WG_MC][TW];
  __shared__ __align__(16) float Hs[2][WG_MC][TW];
  const int tid = threadIdx.x;
  const int total_tiles = a.tile_prefix[a.n_problems];
  const int tile_id = blockIdx.x % total_tiles;
  const int split = blockIdx.x / total_tiles;
  int pi = 0;
  while (tile_id >= a.tile_prefix[pi + 1]) ++pi;
  const WgradProblem& P = a.pr[pi];
  const int local = tile_id - a.tile_prefix[pi];
  const int tn = local / P.ntk, tk = local - tn * P.ntk;
  const int n0 = tn * TW, k0 = tk * TW;
  const long long per = ((a.S + a.splits - 1) / a.splits + 15) & ~15LL;
  const long long m_begin = (long long)split * per;
  const long long m_end = min(a.S, m_begin + per);
  const int ty = tid >> 4, tx = tid & 15;  // ty: 16 groups of NT n; tx: 16 groups of 4 (+ 4) k
  float acc[NT][NT];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  float bsum[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) bsum[i] = 0.f;
  float dreg[PER], hreg[PER];
  auto fetch = [&](long long m0) {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int idx = e * NTHREADS + tid;
      const int mm = idx / TW, cc = idx % TW;
      const long long m = m0 + mm;
      dreg[e] = (m < m_end && n0 + cc < P.N)
                    ? __ldg(P.D + (P.d_rowmajor ? m * P.ldd + n0 + cc : ren_blk(m, n0 + cc, P.ldd))) : 0.f;
      hreg[e] = (m < m_end && k0 + cc < P.Kc) ? __ldg(P.H + ren_blk(m, k0 + cc, P.ldh)) : 0.f;
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int idx = e * NTHREADS + tid;
      Ds[buf][idx / TW][idx % TW] = dreg[e];
      Hs[buf][idx / TW][idx % TW] = hreg[e];
    }
  };
  if (m_begin < m_end) {
    fetch(m_begin);
    commit(0);
  }
  __syncthreads();
  int buf = 0;
  for (long long m0 = m_begin; m0 < m_end; m0 += WG_MC) {
    const bool more = m0 + WG_MC < m_end;
    if (more) fetch(m0 + WG_MC);
#pragma unroll
    for (int mm = 0; mm < WG_MC; ++mm) {
      float dv[NT], hv[NT];
#pragma unroll
      for (int i = 0; i < NT; i += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(&Ds[buf][mm][ty * NT + i]);
        dv[i] = d4.x; dv[i + 1] = d4.y; dv[i + 2] = d4.z; dv[i + 3] = d4.w;
      }
      {
        const float4 h0 = *reinterpret_cast<const float4*>(&Hs[buf][mm][tx * 4]);
        hv[0] = h0.x; hv[1] = h0.y; hv[2] = h0.z; hv[3] = h0.w;
        if (NT == 8) {
          const float4 h1 = *reinterpret_cast<const float4*>(&Hs[buf][mm][(TW / 2) + tx * 4]);
          hv[NT - 4] = h1.x; hv[NT - 3] = h1.y; hv[NT - 2] = h1.z; hv[NT - 1] = h1.w;
        }
      }
      const bool brow = !(a.tangent && (mm & 1));    // m0 is a multiple of 16: row parity == mm parity
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        if (brow) bsum[i] += dv[i];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(dv[i], hv[j], acc[i][j]);
      }
    }
    if (more) commit(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int n = n0 + ty * NT + i;
    if (n >= P.N) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int k = k0 + (j < 4 ? tx * 4 + j : (TW / 2) + tx * 4 + (j - 4));
      if (k < P.Kc) atomicAdd(P.dW + (long long)n * P.ldw + k, acc[i][j]);
    }
    if (P.db != nullptr && tk == 0 && tx == 0) atomicAdd(P.db + n, bsum[i]);
  }
}

// ------------------------------------------------------------------ launchers (host)
static int g_sm_count = 0;
static int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

int ren_simt_launch_fwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                        const float* o_dot, const float* d_dot, const float* jitter, const float* params,
                        float* stash, float* raw, const RenNgp* ngp, cudaStream_t st) {
  FwdArgs a;
  static_assert(sizeof(FwdArgs) <= 4000, "kernel parameter block too large");
  memset(&a.grid, 0, sizeof(a.grid));
  a.table = nullptr;
  if (p.arch == REN_ARCH_NGP) {
    if (ngp == nullptr) { ren_set_error("simt fwd: arch ngp without grid"); return REN_E_INVALID; }
    a.grid = ngp->grid; a.table = ngp->table;
  }
  a.p = p; a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter;
  a.params = params; a.stash = stash; a.raw = raw; a.rays_o_dot = o_dot; a.rays_d_dot = d_dot;
  const bool tan = o_dot != nullptr;
  const int WS = p.W > p.WC ? p.W : p.WC;
  const size_t smem = simt_smem_bytes(WS);
  cudaError_t e = tan ? cudaFuncSetAttribute(ren_simt_field_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(ren_simt_field_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { ren_set_error("simt fwd smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return REN_E_CUDA; }
  const long long S = n_rays * K * (tan ? 2 : 1);
  const long long n_tiles = (S + TM - 1) / TM;
  long long grid = n_tiles < (long long)sm_count() * 8 ? n_tiles : (long long)sm_count() * 8;
  if (grid < 1) return REN_OK;
  ren_prof_begin(REN_K_FIELD_FWD, st);
  if (tan) ren_simt_field_fwd<true><<<(unsigned)grid, NTHREADS, smem, st>>>(a);
  else ren_simt_field_fwd<false><<<(unsigned)grid, NTHREADS, smem, st>>>(a);
  ren_prof_end(REN_K_FIELD_FWD, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("simt fwd launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

int ren_simt_launch_dgrad(const RenPlan& p, long long S, const float* params, const float* stash,
                          const float* graw, float* delta, int tangent, cudaStream_t st) {
  DgradArgs a;
  a.p = p; a.S = S; a.params = params; a.stash = stash; a.graw = graw; a.delta = delta;
  const int WS = p.W > p.WC ? p.W : p.WC;
  const size_t smem = simt_smem_bytes(WS);
  cudaError_t e = tangent ? cudaFuncSetAttribute(ren_simt_field_dgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(ren_simt_field_dgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { ren_set_error("simt dgrad smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  const long long n_tiles = (S + TM - 1) / TM;
  long long grid = n_tiles < (long long)sm_count() * 8 ? n_tiles : (long long)sm_count() * 8;
  if (grid < 1) return REN_OK;
  ren_prof_begin(REN_K_FIELD_DGRAD, st);
  if (tangent) ren_simt_field_dgrad<true><<<(unsigned)grid, NTHREADS, smem, st>>>(a);
  else ren_simt_field_dgrad<false><<<(unsigned)grid, NTHREADS, smem, st>>>(a);
  ren_prof_end(REN_K_FIELD_DGRAD, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("simt dgrad launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

// dW/db for every layer from (stash activations, delta workspace, graw).
int ren_simt_launch_wgrad(const RenPlan& p, long long S, const float* stash, const float* delta,
                          const float* graw, float* g_params, int tangent, cudaStream_t st) {
  WgradArgs a;
  a.tangent = tangent;
  int np = 0;
  // 64 x 64 output tiles when no layer is wider than 64 (arch ngp), else 128 x 128
  const bool narrow = p.W <= 64 && p.WC <= 64 && p.PE <= 64 && p.DE <= 64;
  const int tw = narrow ? 64 : 128;
  const float* encP = stash + ren_stash_encp(p, S);
  const float* encD = stash + ren_stash_encd(p, S);
  auto add = [&](const float* D, int ldd, int N, const float* H, int ldh, int Kc, float* dW, int ldw,
                 float* db) {
    WgradProblem& P = a.pr[np++];
    P.D = D; P.ldd = ldd; P.N = N; P.H = H; P.ldh = ldh; P.Kc = Kc; P.dW = dW; P.ldw = ldw; P.db = db;
    P.d_rowmajor = (D == graw || D == graw + 1) ? 1 : 0;
    P.ntn = (N + tw - 1) / tw; P.ntk = (Kc + tw - 1) / tw;
  };
  auto add_layer = [&](const RenLayer& L, const float* in_main, int ld_main) {
    const float* D = delta + ren_stash_mat(p, S, L.mat);
    const int ldw = L.K0 + L.K1;
    const float* h0 = L.src0 == 1 ? encP : in_main;
    const int ld0 = L.src0 == 1 ? REN_PE_MAX : ld_main;
    add(D, L.N, L.N, h0, ld0, L.K0, g_params + L.w_off, ldw, g_params + L.b_off);
    if (L.K1 > 0)
      add(D, L.N, L.N, L.src1 == 1 ? encP : encD, L.src1 == 1 ? REN_PE_MAX : REN_DE_MAX, L.K1,
          g_params + L.w_off + L.K0, ldw, nullptr);
  };
  const float* prev = nullptr;
  int ldprev = 0;
  for (int i = 0; i < p.D; ++i) {
    add_layer(p.trunk[i], prev, ldprev);
    prev = stash + ren_stash_mat(p, S, p.trunk[i].mat);
    ldprev = p.trunk[i].N;
  }
  const float* hid = prev; const int ldhid = ldprev;
  add_layer(p.bott, hid, ldhid);
  prev = stash + ren_stash_mat(p, S, p.bott.mat);
  ldprev = p.bott.N;
  for (int j = 0; j < p.CD; ++j) {
    add_layer(p.cond[j], prev, ldprev);
    prev = stash + ren_stash_mat(p, S, p.cond[j].mat);
    ldprev = p.cond[j].N;
  }
  const int C1 = 1 + p.C;
  // sigma head: Delta = graw[:, 0], inputs = hidden features
  {
    const int ldw = p.hid_K0 + p.hid_K1;
    add(graw, C1, 1, p.D > 0 ? hid : encP, p.D > 0 ? ldhid : REN_PE_MAX, p.hid_K0, g_params + p.sig_w, ldw,
        g_params + p.sig_b);
    if (p.hid_K1 > 0) add(graw, C1, 1, encP, REN_PE_MAX, p.hid_K1, g_params + p.sig_w + p.hid_K0, ldw, nullptr);
  }
  // radiance head: Delta = graw[:, 1:], inputs = last condition layer (or bottleneck | encD)
  {
    const int Kin = p.CD > 0 ? p.WC : p.W;
    const int ldw = p.CD > 0 ? Kin : Kin + p.DE;
    add(graw + 1, C1, p.C, prev, ldprev, Kin, g_params + p.rgb_w, ldw, g_params + p.rgb_b);
    if (p.CD == 0) add(graw + 1, C1, p.C, encD, REN_DE_MAX, p.DE, g_params + p.rgb_w + Kin, ldw, nullptr);
  }
  if (np > WG_MAXP) { ren_set_error("too many wgrad problems"); return REN_E_INVALID; }
  a.n_problems = np;
  a.tile_prefix[0] = 0;
  for (int i = 0; i < np; ++i) a.tile_prefix[i + 1] = a.tile_prefix[i] + a.pr[i].ntn * a.pr[i].ntk;
  const int total = a.tile_prefix[np];
  int splits = (sm_count() * 2 * 4 + total - 1) / total;
  const long long max_splits = (S + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  a.splits = splits;
  a.S = S;
  ren_prof_begin(REN_K_WGRAD, st);
  if (narrow) ren_simt_wgrad<4><<<(unsigned)(total * splits), NTHREADS, 0, st>>>(a);
  else ren_simt_wgrad<8><<<(unsigned)(total * splits), NTHREADS, 0, st>>>(a);
  ren_prof_end(REN_K_WGRAD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("simt wgrad launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
