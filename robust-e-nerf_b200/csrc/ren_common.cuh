// ren_common.cuh -- shared host/device definitions of libren_b200.so.
//
// Network plan (host-built from RenFieldDesc), per-sample geometry (march,
// position, contraction, sinusoidal encoding) and activations.  The geometry
// section replicates the reference's fp32 rounding sequence op by op (no FMA
// contraction): SURVEY.md App. A5/A6, external/utils.py:87, external/mlp.py:321-335,
// external/ngp.py:68-106, external/mlp.py:227-243.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ren_b200.h"

#define REN_MAX_TRUNK 16
#define REN_MAX_COND 4
#define REN_PE_MAX 64   // padded positional-encoding width (3*(1+2*10)=63 -> 64)
#define REN_DE_MAX 32   // padded direction-encoding width (3*(1+2*4)=27 -> 32)

// One dense layer of the MLP as the kernels see it: out = act(W [in0 | in1] + b).
struct RenLayer {
  int w_off, b_off;  // offsets (floats) into the flat parameter buffer; W is (N, K0+K1) row-major
  int K0, K1;        // main / extra input widths (K1: 0, PE or DE)
  int src0;          // 0: previous layer output buffer, 1: positional encoding (layer 0)
  int src1;          // 0: none, 1: positional encoding, 2: direction encoding
  int N;             // output width
  int act;           // 0 none, 1 hidden activation
  int mat;           // index of this layer's output matrix in the stash / delta workspace
};

struct RenPlan {
  RenFieldDesc f;
  RenRenderDesc r;
  int D, CD;                 // trunk depth, condition depth
  RenLayer trunk[REN_MAX_TRUNK];
  RenLayer bott;
  RenLayer cond[REN_MAX_COND];
  // heads: sigma (1 x hidden), rgb out (C x WC)
  int sig_w, sig_b, rgb_w, rgb_b;
  int hid_K0, hid_K1;        // hidden feature segments feeding sigma / bottleneck
  int W, WC, C, PE, DE;      // widths: trunk, condition, radiance, enc dims (unpadded)
  int n_mats;                // D + 1 + CD activation matrices
  int mat_width[REN_MAX_TRUNK + 1 + REN_MAX_COND];
  long long mat_off[REN_MAX_TRUNK + 1 + REN_MAX_COND];  // prefix of widths (floats per sample)
  long long act_floats;      // sum of widths
  int n_params;              // arch mlp: the flat parameter buffer; arch ngp: the kernel-side MLP buffer (below)
  // arch ngp (ren_ngp.cu): the kernels run on a kernel-side copy of the MLP parameters laid out like the vanilla
  // topology (trunk = mlp_base hidden layers, sigma head = row 0 of mlp_base's output layer, bottleneck = its
  // geo-feature rows padded to WB, condition layers = mlp_head with input [features | SH]); the caller's flat
  // buffer keeps the reference's state-dict layout [grid | mlp_base | mlp_head]
  int arch, WB;              // WB: bottleneck width (W for arch mlp, geo_feat_dim rounded up to 8 for ngp)
  int geo;                   // geo_feat_dim
  long long n_grid;          // grid parameters at the head of the caller's flat buffer
  long long n_flat;          // elements of the caller's flat buffer
};

// arch ngp: multiresolution grid of tiny-cuda-nn (grid.h: grid_scale / grid_resolution / offset table)
#define REN_MAX_LEVELS 32
struct RenGridLevel {
  float scale;               // fp32(2^(l log2 per_level_scale) base_resolution - 1), evaluated in double on the host
  uint32_t res;              // ceil(scale) + 1
  uint32_t entries;          // table entries of the level
  uint32_t offset;           // first entry of the level
};
struct RenGrid {
  int n_levels, F, type;
  RenGridLevel lv[REN_MAX_LEVELS];
};
// what the NGP-aware kernels need besides the plan (NULL for arch mlp)
struct RenNgp {
  RenGrid grid;
  const float* table;        // grid parameters (caller's flat buffer + 0)
  float* g_table;            // their gradient (accumulated), backward only
};

// Stash layout (floats), S = n_rays*K samples:
//   [mat m : S x mat_width[m]] for m in 0..n_mats-1, then encP [S x 64], encD [S x 32],
//   then raw [S x (1+C)] (sigma_raw or -inf when outside the contracted box, c_raw).
// The delta workspace used by backward mirrors the mats + graw [S x (1+C)].
//
// Activation / delta matrices are stored BLOCKED so that the thread-per-sample-row epilogues
// write full lines: a [S x N] matrix is tiled into 128-sample x 8-feature blocks of 4 KB,
//   index(s, n) = ((s / 128) * (N / 8) + n / 8) * 1024 + (s % 128) * 8 + n % 8
// (the fp32 image of the tcgen05 K-major core-matrix operand layout); S is padded to 128.
__host__ __device__ inline long long ren_pad128(long long S) { return (S + 127) & ~127LL; }
__host__ __device__ inline long long ren_blk(long long s, int n, int N) {
  return (((s >> 7) * (N >> 3) + (n >> 3)) << 10) + ((s & 127) << 3) + (n & 7);
}
__host__ __device__ inline long long ren_stash_mat(const RenPlan& p, long long S, int m) {
  return p.mat_off[m] * ren_pad128(S);
}
__host__ __device__ inline long long ren_stash_encp(const RenPlan& p, long long S) {
  return p.act_floats * ren_pad128(S);
}
__host__ __device__ inline long long ren_stash_encd(const RenPlan& p, long long S) {
  return (p.act_floats + REN_PE_MAX) * ren_pad128(S);
}
__host__ __device__ inline long long ren_stash_raw(const RenPlan& p, long long S) {
  return (p.act_floats + REN_PE_MAX + REN_DE_MAX) * ren_pad128(S);
}
__host__ __device__ inline long long ren_stash_floats(const RenPlan& p, long long S) {
  return (p.act_floats + REN_PE_MAX + REN_DE_MAX + 1 + p.C) * ren_pad128(S);
}
// floats of the backward workspace: delta mats (blocked) + graw [S][1+C]
__host__ __device__ inline long long ren_delta_floats(const RenPlan& p, long long S) {
  return (p.act_floats + 1 + p.C) * ren_pad128(S);
}

#ifdef __CUDACC__

// ------------------------------------------------------------------ activations
// torch.nn.functional.softplus(x, beta, threshold=20): models/nerf.py:17-29
__device__ __forceinline__ float ren_softplus(float x, float beta) {
  float z = x * beta;
  return (z > 20.f) ? x : log1pf(expf(z)) / beta;
}
// d softplus / dx expressed through the OUTPUT y (sigmoid(beta x) = 1 - exp(-beta y))
__device__ __forceinline__ float ren_softplus_grad_from_out(float y, float beta) {
  return -expm1f(-beta * y);
}
__device__ __forceinline__ float ren_hidden_act(int kind, float x) {
  return kind == REN_HIDDEN_SOFTPLUS100 ? ren_softplus(x, 100.f) : fmaxf(x, 0.f);
}
__device__ __forceinline__ float ren_hidden_grad_from_out(int kind, float y) {
  return kind == REN_HIDDEN_SOFTPLUS100 ? ren_softplus_grad_from_out(y, 100.f)
                                        : (y > 0.f ? 1.f : 0.f);
}
// tangent mode (d/dt of the hidden activations): act'(x) from the PRE-activation x
// (torch softplus backward: z/(z+1) with z = exp(beta x), 1 above the threshold)
__device__ __forceinline__ float ren_hidden_grad_from_pre(int kind, float x) {
  if (kind != REN_HIDDEN_SOFTPLUS100) return x > 0.f ? 1.f : 0.f;
  const float z = x * 100.f;
  return (z > 20.f) ? 1.f : 1.f / (1.f + expf(-z));
}
// act''(x) * xdot expressed through the stashed outputs: y = act(x), ydot = act'(x) xdot.
// softplus(beta): act'' = beta s (1 - s) with s = act'  =>  act'' xdot = beta (1 - s) ydot
__device__ __forceinline__ float ren_hidden_curv_from_out(int kind, float y, float ydot) {
  if (kind != REN_HIDDEN_SOFTPLUS100) return 0.f;
  return 100.f * expf(-100.f * y) * ydot;
}
// second derivatives of the output activations wrt their raw value (reverse pass of the tangent
// render; the reference obtains them by double-backward).  _TruncExp.backward multiplies by
// exp(clamp(x, max=15)) (external/ngp.py:54-58), whose own derivative vanishes above the clamp.
__device__ __forceinline__ float ren_density_act2(int kind, float raw) {
  if (kind == REN_DENSITY_SHIFTED_TRUNC_EXP) {
    const float x = raw - 1.f;
    return x < 15.f ? expf(x) : 0.f;
  }
  const float x = (kind == REN_DENSITY_SHIFTED_SOFTPLUS) ? raw - 1.f : raw;
  if (x > 20.f) return 0.f;
  const float s = 1.f / (1.f + expf(-x));
  return s * (1.f - s);
}
__device__ __forceinline__ float ren_radiance_act2(int kind, float raw) {
  if (kind == REN_RADIANCE_SOFTPLUS) {
    if (raw > 20.f) return 0.f;
    const float s = 1.f / (1.f + expf(-raw));
    return s * (1.f - s);
  }
  const float s = 1.f / (1.f + expf(-raw));
  return s * (1.f - s) * (1.f - 2.f * s);
}
// density activation and its derivative wrt the raw value (external/ngp.py:45-65,
// models/nerf.py:8-25).  raw == -inf encodes "outside the box" (selector false).
__device__ __forceinline__ float ren_density_act(int kind, float raw, float* dact) {
  if (kind == REN_DENSITY_SHIFTED_TRUNC_EXP) {
    float x = raw - 1.f;
    *dact = expf(fminf(x, 15.f));
    return expf(x);
  }
  float x = (kind == REN_DENSITY_SHIFTED_SOFTPLUS) ? raw - 1.f : raw;
  float s = ren_softplus(x, 1.f);
  *dact = (x > 20.f) ? 1.f : 1.f / (1.f + expf(-x));
  return s;
}
__device__ __forceinline__ float ren_radiance_act(int kind, float raw, float* dact) {
  if (kind == REN_RADIANCE_SOFTPLUS) {
    *dact = (raw > 20.f) ? 1.f : 1.f / (1.f + expf(-raw));
    return ren_softplus(raw, 1.f);
  }
  float s = 1.f / (1.f + expf(-raw));
  *dact = s * (1.f - s);
  return s;
}

// ------------------------------------------------------------------ marching
// nerfacc v0.3.1 ray_marching over an all-occupied grid (external/utils.py:106-119,
// models/nerf.py:248-251): per-ray [t_min, t_max], one stratified offset per ray.
__device__ __forceinline__ void ren_ray_range(const RenFieldDesc& f, const RenRenderDesc& r,
                                              const float* o, const float* d, float u,
                                              float* t_min_out, float* t_max_out) {
  float t_min = 0.f, t_max = 1e10f;
  if (f.contraction == REN_CONTRACT_AABB) {
    float tn = -INFINITY, tf = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float lo = __fdiv_rn(__fsub_rn(f.aabb[a], o[a]), d[a]);
      float hi = __fdiv_rn(__fsub_rn(f.aabb[3 + a], o[a]), d[a]);
      tn = fmaxf(tn, fminf(lo, hi));
      tf = fminf(tf, fmaxf(lo, hi));
    }
    bool miss = tn > tf;
    tn = fmaxf(tn, 0.f);
    t_min = miss ? 1e10f : tn;
    t_max = miss ? 1e10f : tf;
  }
  if (r.has_near) t_min = fmaxf(t_min, r.near_plane);
  if (r.has_far) t_max = fminf(t_max, r.far_plane);
  if (r.stratified) t_min = __fadd_rn(t_min, __fmul_rn(u, r.render_step_size));
  *t_min_out = t_min;
  *t_max_out = t_max;
}
__device__ __forceinline__ float ren_dt(const RenRenderDesc& r, float t0) {
  return fminf(fmaxf(__fmul_rn(t0, r.cone_angle), r.render_step_size), 1e10f);
}
// t0/t1 of the k-th marching step (sequential fp32 accumulation, like the nerfacc loop).
__device__ __forceinline__ void ren_march_to(const RenRenderDesc& r, float t_min, int k, float* t0o,
                                             float* t1o) {
  float t0 = t_min;
  float t1 = __fadd_rn(t0, ren_dt(r, t0));
  for (int i = 0; i < k; ++i) {
    t0 = t1;
    t1 = __fadd_rn(t0, ren_dt(r, t0));
  }
  *t0o = t0;
  *t1o = t1;
}
__device__ __forceinline__ bool ren_alive(float t0, float t1, float t_max) {
  return __fmul_rn(__fadd_rn(t0, t1), 0.5f) < t_max;
}

// ------------------------------------------------------------------ field input
// positions = o + d*(t0+t1)/2 (external/utils.py:87), contraction to [0,1]^3 + selector,
// then 2*pi*(x-0.5) (external/mlp.py:321-335).  Returns selector.
__device__ __forceinline__ void ren_unit_position(const RenFieldDesc& f, const float* o, const float* d,
                                                  float t0, float t1, float* x) {
  float tm2 = __fadd_rn(t0, t1);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float p = __fadd_rn(o[a], __fmul_rn(__fmul_rn(d[a], tm2), 0.5f));
    x[a] = __fdiv_rn(__fsub_rn(p, f.aabb[a]), __fsub_rn(f.aabb[3 + a], f.aabb[a]));
  }
  if (f.contraction == REN_CONTRACT_SPHERE) {
#pragma unroll
    for (int a = 0; a < 3; ++a) x[a] = __fsub_rn(__fmul_rn(x[a], 2.f), 1.f);
    float mag = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x[0], x[0]), __fmul_rn(x[1], x[1])),
                                     __fmul_rn(x[2], x[2])));
    if (mag > 1.f) {
      float s = __fsub_rn(2.f, __fdiv_rn(1.f, mag));
#pragma unroll
      for (int a = 0; a < 3; ++a) x[a] = __fmul_rn(s, __fdiv_rn(x[a], mag));
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) x[a] = __fadd_rn(__fmul_rn(x[a], 0.25f), 0.5f);
  } else if (f.contraction == REN_CONTRACT_TANH) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
      x[a] = __fmul_rn(__fadd_rn(tanhf(__fsub_rn(x[a], 0.5f)), 1.f), 0.5f);
  }
}
__device__ __forceinline__ bool ren_field_input(const RenFieldDesc& f, const float* o,
                                                const float* d, float t0, float t1, float* z) {
  float x[3];
  ren_unit_position(f, o, d, t0, t1, x);
  bool sel = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    sel = sel && (x[a] > 0.f) && (x[a] < 1.f);
    z[a] = __fmul_rn(6.283185307179586f, __fsub_rn(x[a], 0.5f));
  }
  return sel;
}
// nerfacc grid_occupied_at (ray_marching kernel, call site external/utils.py:106-119): the marching step's
// midpoint, contracted to the unit cube, looked up in the binary grid (cell = floor(x * res), clamped).
__device__ __forceinline__ bool ren_occupied(const RenFieldDesc& f, const RenRenderDesc& r, const float* o,
                                             const float* d, float t0, float t1) {
  if (r.occ_binary == nullptr) return true;
  float x[3];
  ren_unit_position(f, o, d, t0, t1, x);
  int idx = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int i = (int)floorf(x[a] * (float)r.occ_res[a]);
    i = i < 0 ? 0 : (i > r.occ_res[a] - 1 ? r.occ_res[a] - 1 : i);
    idx = idx * r.occ_res[a] + i;
  }
  return r.occ_binary[idx] != 0;
}

// SinusoidalEncoder.forward (external/mlp.py:227-243): [x | sin(2^i x) | sin(2^i x + pi/2)],
// degree-major, dim-minor.  `out` has stride `stride` between features; padded tail zeroed.
template <typename Store>
__device__ __forceinline__ void ren_encode(const float* x, int deg, int padded, Store store) {
  store(0, x[0]);
  store(1, x[1]);
  store(2, x[2]);
  float sc = 1.f;
#pragma unroll
  for (int i = 0; i < deg; ++i) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float xb = __fmul_rn(x[a], sc);
      store(3 + i * 3 + a, sinf(xb));
      store(3 + deg * 3 + i * 3 + a, sinf(__fadd_rn(xb, 1.5707963267948966f)));
    }
    sc *= 2.f;
  }
  for (int j = 3 + 6 * deg; j < padded; ++j) store(j, 0.f);
}

// d/dt of SinusoidalEncoder(x) given xdot: [xdot | 2^i cos(2^i x) xdot | 2^i cos(2^i x + pi/2) xdot]
template <typename Store>
__device__ __forceinline__ void ren_encode_tangent(const float* x, const float* xdot, int deg, int padded,
                                                   Store store) {
  store(0, xdot[0]);
  store(1, xdot[1]);
  store(2, xdot[2]);
  float sc = 1.f;
#pragma unroll
  for (int i = 0; i < deg; ++i) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float xb = __fmul_rn(x[a], sc);
      store(3 + i * 3 + a, cosf(xb) * sc * xdot[a]);
      store(3 + deg * 3 + i * 3 + a, cosf(__fadd_rn(xb, 1.5707963267948966f)) * sc * xdot[a]);
    }
    sc *= 2.f;
  }
  for (int j = 3 + 6 * deg; j < padded; ++j) store(j, 0.f);
}

#endif  // __CUDACC__ (geometry)

// ------------------------------------------------------------------ arch ngp
// `coherent_prime_hash` / `grid_index` of tiny-cuda-nn (grid.h): dense index while the stride fits the level's
// table, else the prime hash; always modulo the table size.
__host__ __device__ inline uint32_t ren_grid_index(const RenGridLevel& lv, int type, uint32_t x, uint32_t y,
                                                   uint32_t z) {
  uint32_t stride = 1, index = 0;
  const uint32_t c[3] = {x, y, z};
  for (int dim = 0; dim < 3 && stride <= lv.entries; ++dim) {
    index += c[dim] * stride;
    stride *= lv.res;
  }
  if (type == REN_GRID_HASH && lv.entries < stride) index = (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);
  return index % lv.entries;
}
#ifdef __CUDACC__
// Corner data of one level at unit position x: integer cell, fractional position (pos_fract with the identity
// interpolant: pos = fmaf(scale, x, 0.5f); floor; frac).
struct RenCell { uint32_t g[3]; float f[3]; };
__device__ __forceinline__ void ren_grid_cell(const RenGridLevel& lv, const float* x, RenCell& c) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pos = fmaf(lv.scale, x[a], 0.5f);
    const float fl = floorf(pos);
    c.g[a] = (uint32_t)(int)fl;
    c.f[a] = pos - fl;
  }
}
// trilinear weight of corner `corner` and its derivative data: a[dim] = f or 1 - f, s[dim] = +1 / -1
__device__ __forceinline__ void ren_corner(const RenCell& c, int corner, float* a, float* s, uint32_t* g) {
#pragma unroll
  for (int dim = 0; dim < 3; ++dim) {
    const bool up = (corner >> dim) & 1;
    a[dim] = up ? c.f[dim] : 1.f - c.f[dim];
    s[dim] = up ? 1.f : -1.f;
    g[dim] = c.g[dim] + (up ? 1u : 0u);
  }
}
// enc[l*F + j] = sum_corners w T[idx][j]  (kernel_grid of tiny-cuda-nn);  xdot != NULL: d enc / d w instead,
// with the weights' derivative along xdot (d frac / d x = scale)
// levels l0, l0 + lstep, ... (lstep = 1: all of them; the fp32 engine's prologue spreads a row's levels over 4 threads)
template <typename Store>
__device__ __forceinline__ void ren_grid_encode(const RenGrid& G, const float* __restrict__ table, const float* x,
                                                const float* xdot, Store store, int l0 = 0, int lstep = 1) {
  for (int l = l0; l < G.n_levels; l += lstep) {
    const RenGridLevel lv = G.lv[l];
    RenCell c;
    ren_grid_cell(lv, x, c);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      float a[3], s[3];
      uint32_t g[3];
      ren_corner(c, corner, a, s, g);
      float w;
      if (xdot == nullptr) {
        w = a[0] * a[1] * a[2];
      } else {
        w = lv.scale * (s[0] * xdot[0] * a[1] * a[2] + s[1] * xdot[1] * a[0] * a[2] + s[2] * xdot[2] * a[0] * a[1]);
      }
      const float* t = table + ((size_t)lv.offset + ren_grid_index(lv, G.type, g[0], g[1], g[2])) * G.F;
      if (G.F == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(t));
        acc[0] = fmaf(w, v.x, acc[0]);
        acc[1] = fmaf(w, v.y, acc[1]);
      } else {
        for (int j = 0; j < G.F; ++j) acc[j] = fmaf(w, __ldg(t + j), acc[j]);
      }
    }
    for (int j = 0; j < G.F; ++j) store(l * G.F + j, acc[j]);
  }
}
#endif

// real spherical harmonics of the raw direction, degree <= 4 (external/sh_encoder.py:27-85), on any number type
template <class T>
__host__ __device__ inline void ren_sh_t(const T* d, int degree, T* out) {
  const T x = d[0], y = d[1], z = d[2];
  const T xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  out[0] = T(0.28209479177387814f);
  if (degree <= 1) return;
  out[1] = T(-0.48860251190291987f) * y;
  out[2] = T(0.48860251190291987f) * z;
  out[3] = T(-0.48860251190291987f) * x;
  if (degree <= 2) return;
  out[4] = T(1.0925484305920792f) * xy;
  out[5] = T(-1.0925484305920792f) * yz;
  out[6] = T(0.94617469575755997f) * z2 - T(0.31539156525251999f);
  out[7] = T(-1.0925484305920792f) * xz;
  out[8] = T(0.54627421529603959f) * x2 - T(0.54627421529603959f) * y2;
  if (degree <= 3) return;
  out[9] = T(0.59004358992664352f) * y * (T(-3.0f) * x2 + y2);
  out[10] = T(2.8906114426405538f) * xy * z;
  out[11] = T(0.45704579946446572f) * y * (T(1.0f) - T(5.0f) * z2);
  out[12] = T(0.3731763325901154f) * z * (T(5.0f) * z2 - T(3.0f));
  out[13] = T(0.45704579946446572f) * x * (T(1.0f) - T(5.0f) * z2);
  out[14] = T(1.4453057213202769f) * z * (x2 - y2);
  out[15] = T(0.59004358992664352f) * x * (T(-1.0f) * x2 + T(3.0f) * y2);
}

// host side (ren_api.cu)
enum { REN_K_FIELD_FWD = 0, REN_K_FIELD_DGRAD, REN_K_WGRAD, REN_K_COMPOSITE_FWD, REN_K_COMPOSITE_BWD,
       REN_K_EVENT, REN_K_LOSS, REN_K_ADAM, REN_K_PACK, REN_K_COUNT };
void ren_prof_begin(int kind, cudaStream_t st);   // counts the launch; records an event when timing is on
void ren_prof_end(int kind, cudaStream_t st);
void ren_prof_count();                            // counts a launch that runs beside a timed one (no events)
int ren_build_plan(const RenFieldDesc* f, const RenRenderDesc* r, RenPlan* plan);
void ren_set_error(const char* fmt, ...);
