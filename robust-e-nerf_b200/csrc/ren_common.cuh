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
  int n_params;
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

#endif  // __CUDACC__

// host side (ren_api.cu)
enum { REN_K_FIELD_FWD = 0, REN_K_FIELD_DGRAD, REN_K_WGRAD, REN_K_COMPOSITE_FWD, REN_K_COMPOSITE_BWD,
       REN_K_EVENT, REN_K_LOSS, REN_K_ADAM, REN_K_PACK, REN_K_COUNT };
void ren_prof_begin(int kind, cudaStream_t st);   // counts the launch; records an event when timing is on
void ren_prof_end(int kind, cudaStream_t st);
void ren_prof_count();                            // counts a launch that runs beside a timed one (no events)
int ren_build_plan(const RenFieldDesc* f, const RenRenderDesc* r, RenPlan* plan);
void ren_set_error(const char* fmt, ...);
