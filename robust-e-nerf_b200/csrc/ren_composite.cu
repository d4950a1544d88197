// ren_composite.cu -- per-ray transmittance compositing, forward and reverse.
//
// Replaces nerfacc v0.3.1 `render_visibility` (the no-grad cull inside
// `ray_marching`, external/utils.py:106-119), `render_weight_from_density` and
// `accumulate_along_rays` as used by `rendering` (external/vol_rendering.py:82-126).
// The density cull is evaluated from the SAME field pass (same weights, same samples
// => same sigma as the reference's extra no-grad trunk pass), so the trunk runs once.
//
// One warp per ray: lanes take consecutive samples, the exclusive transmittance sum is a
// warp scan with a running carry, the early-stop rule is resolved with ballots.
#include "ren_common.cuh"

struct CompArgs {
  RenPlan p;
  long long n_rays;
  int K;
  const float* rays_o;
  const float* rays_d;
  const float* jitter;
  const float* raw;      // [S][1+C]
  const float* bkgd;     // [C] or NULL
  float* radiance;       // [n_rays][C]
  float* opacity;        // [n_rays]
  float* depth;          // [n_rays]
  int* n_kept;           // [n_rays]
  // backward only
  const float* g_radiance;
  const float* g_opacity;
  const float* g_depth;
  float* graw;           // [S][1+C]
  float* g_bkgd;         // [C] or NULL
  unsigned int* gmax;    // max |graw| as float bits (tcgen05 backward loss scaling) or NULL
  // tangent mode: raw / graw have 2 rows per sample (value, d/dw) -- or, when raw_dot / graw_dot are
  // given (tcgen05 engine), the d/dw rows live in their own [S][1+C] arrays
  int tangent;
  const float* raw_dot;
  float* graw_dot;
  // packed samples (occupancy-grid marching on the tcgen05 engine): raw / graw hold only the emitted samples,
  // ray r owns [ray_offsets[r], ray_offsets[r+1]) in marching order
  const int* ray_offsets;
  float* radiance_dot;   // [n_rays][C]
  float* opacity_dot;    // [n_rays]
  const float* g_radiance_dot;
  const float* g_opacity_dot;
};

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-sample quantities of one 32-sample block of a ray (lane = sample).
struct SampleBlk {
  float t0, t1, sdt, alpha, w, Texcl, dact_s, dact_c[3], c[3];
  bool alive, keep;
  long long sidx;       // row of this sample in raw / graw (dense r*K + k, or its packed position)
};

// Walks the ray in blocks of 32 samples.  `Body(k, blk)` is called for every sample index
// k < K (blk.keep false for culled / dead samples).
template <typename Body>
__device__ __forceinline__ void walk_ray(const CompArgs& a, long long r, int lane, Body body) {
  const RenPlan& p = a.p;
  const int C1 = 1 + p.C;
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = a.rays_o[r * 3 + c];
    d[c] = a.rays_d[r * 3 + c];
  }
  float t_min, t_max;
  ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
  // every lane marches to its own first sample, then strides 32 steps per block
  float t0, t1;
  ren_march_to(p.r, t_min, lane, &t0, &t1);
  float carry_sum = 0.f;    // sum of kept sigma*dt before this block
  float carry_T = 1.f;      // product of (1-alpha) of kept samples before this block (cull rule)
  bool stopped = false;
  int packed_base = a.ray_offsets ? a.ray_offsets[r] : 0;
  for (int kb = 0; kb < a.K; kb += 32) {
    const int k = kb + lane;
    SampleBlk b;
    b.t0 = t0; b.t1 = t1;
    const bool in_range = (k < a.K) && ren_alive(t0, t1, t_max);   // geometric: midpoint before t_max
    b.alive = in_range;
    b.sidx = r * a.K + k;
    if (a.ray_offsets != nullptr) {
      // packed: only the steps in occupied cells were emitted; rank them in marching order
      const bool emitted = in_range && ren_occupied(p.f, p.r, o, d, t0, t1);
      const unsigned em = __ballot_sync(0xffffffffu, emitted);
      b.sidx = packed_base + __popc(em & ((1u << lane) - 1u));
      packed_base += __popc(em);
      b.alive = emitted;
    }
    float sraw = -INFINITY;
    b.c[0] = b.c[1] = b.c[2] = 0.f;
    b.dact_c[0] = b.dact_c[1] = b.dact_c[2] = 0.f;
    if (b.alive) {
      const float* rw = a.raw + b.sidx * ((a.tangent && !a.raw_dot) ? 2 : 1) * C1;
      sraw = rw[0];
      if (sraw != sraw) {            // NaN: the marching step fell into an unoccupied grid cell -> no sample
        b.alive = false;
        sraw = -INFINITY;
      } else {
        for (int c = 0; c < p.C; ++c) b.c[c] = ren_radiance_act(p.f.radiance_activation, rw[1 + c], &b.dact_c[c]);
      }
    }
    float sigma = ren_density_act(p.f.density_activation, sraw, &b.dact_s);
    if (!b.alive || sraw == -INFINITY) { sigma = 0.f; b.dact_s = 0.f; }
    const float dt = __fsub_rn(t1, t0);
    b.sdt = __fmul_rn(sigma, dt);
    b.alpha = 1.f - expf(-b.sdt);
    // visibility cull (sequential rule resolved blockwise):
    //   stopped |= T < eps before each sample; keep = alive & !stopped & !(alpha < thre);
    //   T *= (1 - alpha) for kept samples.
    bool cand = b.alive && !(b.alpha < p.r.alpha_thre);
    // prefix product of (1-alpha) over candidates, assuming no stop inside the block
    float f = cand ? (1.f - b.alpha) : 1.f;
    float pp = f;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, pp, off);
      if (lane >= off) pp *= t;
    }
    float Tbefore = __shfl_up_sync(0xffffffffu, pp, 1);
    if (lane == 0) Tbefore = 1.f;
    Tbefore *= carry_T;
    // first lane at which the stop rule fires (T before it < eps); all later lanes are stopped
    unsigned stopmask = __ballot_sync(0xffffffffu, Tbefore < p.r.early_stop_eps);
    int first_stop = stopped ? 0 : (stopmask ? (__ffs(stopmask) - 1) : 32);
    b.keep = cand && (lane < first_stop);
    if (first_stop < 32) stopped = true;
    // exclusive sum of kept sigma*dt
    float sk = b.keep ? b.sdt : 0.f;
    float incl = warp_incl_scan(sk, lane);
    float excl = carry_sum + (incl - sk);
    b.Texcl = expf(-excl);
    b.w = b.keep ? b.Texcl * (1.f - expf(-b.sdt)) : 0.f;
    body(k, b);
    carry_sum += __shfl_sync(0xffffffffu, incl, 31);
    {
      // carry_T: product over kept samples of this block
      float fk = b.keep ? (1.f - b.alpha) : 1.f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) fk *= __shfl_xor_sync(0xffffffffu, fk, off);
      carry_T *= fk;
    }
    // advance 32 marching steps
    for (int i = 0; i < 32; ++i) {
      t0 = t1;
      t1 = __fadd_rn(t0, ren_dt(p.r, t0));
    }
    if (__ballot_sync(0xffffffffu, in_range) == 0u) {
      // the whole block is past t_max: remaining samples are dead too; still let the body
      // see them (backward must zero their grads)
      for (int kb2 = kb + 32; kb2 < a.K; kb2 += 32) {
        SampleBlk z = b;
        z.alive = false; z.keep = false; z.w = 0.f; z.sdt = 0.f;
        z.sidx = a.ray_offsets ? -1 : r * a.K + kb2 + lane;
        body(kb2 + lane, z);
      }
      break;
    }
  }
}

extern "C" __global__ void __launch_bounds__(128) ren_composite_fwd(const CompArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const RenPlan& p = a.p;
  for (long long r = warp; r < a.n_rays; r += n_warps) {
    float accC[3] = {0.f, 0.f, 0.f}, accO = 0.f, accD = 0.f;
    int kept = 0;
    walk_ray(a, r, lane, [&](int k, const SampleBlk& b) {
      if (b.keep) {
        for (int c = 0; c < p.C; ++c) accC[c] = fmaf(b.w, b.c[c], accC[c]);
        accO += b.w;
        accD = fmaf(b.w, __fmul_rn(__fadd_rn(b.t0, b.t1), 0.5f), accD);
        kept += 1;
      }
    });
    accO = warp_sum(accO);
    accD = warp_sum(accD);
    for (int c = 0; c < p.C; ++c) accC[c] = warp_sum(accC[c]);
    kept = __reduce_add_sync(0xffffffffu, kept);
    if (lane == 0) {
      for (int c = 0; c < p.C; ++c) {
        float v = accC[c];
        if (a.bkgd) v = __fadd_rn(v, __fmul_rn(a.bkgd[c], __fsub_rn(1.f, accO)));
        a.radiance[r * p.C + c] = v;
      }
      a.opacity[r] = accO;
      a.depth[r] = accD;
      a.n_kept[r] = kept;
    }
  }
}

// Reverse: with G_i = sum_c gC_c (c_i - bkgd_c) + gO + gD tm_i,
//   dL/dc_i = gC w_i,   dL/ds_i = G_i T_i e^{-s_i} - sum_{j>i kept} G_j w_j   (s = sigma*dt)
extern "C" __global__ void __launch_bounds__(128) ren_composite_bwd(const CompArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const RenPlan& p = a.p;
  const int C1 = 1 + p.C;
  float gb_acc[3] = {0.f, 0.f, 0.f};
  float gabs = 0.f;
  for (long long r = warp; r < a.n_rays; r += n_warps) {
    float gC[3] = {0.f, 0.f, 0.f}, bk[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < p.C; ++c) {
      gC[c] = a.g_radiance ? a.g_radiance[r * p.C + c] : 0.f;
      bk[c] = a.bkgd ? a.bkgd[c] : 0.f;
    }
    const float gO = a.g_opacity ? a.g_opacity[r] : 0.f;
    const float gD = a.g_depth ? a.g_depth[r] : 0.f;
    // pass 1: total of G_j w_j and the opacity
    double tot = 0.0;
    float accO = 0.f;
    walk_ray(a, r, lane, [&](int k, const SampleBlk& b) {
      if (b.keep) {
        float G = gO + gD * __fmul_rn(__fadd_rn(b.t0, b.t1), 0.5f);
        for (int c = 0; c < p.C; ++c) G = fmaf(gC[c], b.c[c] - bk[c], G);
        tot += (double)G * (double)b.w;
        accO += b.w;
      }
    });
    tot = warp_sum_d(tot);
    accO = warp_sum(accO);
    if (a.g_bkgd && a.bkgd)
      for (int c = 0; c < p.C; ++c) gb_acc[c] += gC[c] * (1.f - accO);  // same on all lanes
    // pass 2: per-sample grads with a running inclusive prefix of G w
    double prefix = 0.0;  // sum over samples before this block
    walk_ray(a, r, lane, [&](int k, const SampleBlk& b) {
      float G = 0.f;
      double gw = 0.0;
      if (b.keep) {
        G = gO + gD * __fmul_rn(__fadd_rn(b.t0, b.t1), 0.5f);
        for (int c = 0; c < p.C; ++c) G = fmaf(gC[c], b.c[c] - bk[c], G);
        gw = (double)G * (double)b.w;
      }
      double incl = gw;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      const double suffix = tot - (prefix + incl);
      prefix += __shfl_sync(0xffffffffu, incl, 31);
      if (k < a.K && (a.ray_offsets == nullptr || b.alive)) {
        float* g = a.graw + b.sidx * C1;
        float gs = 0.f;
        if (b.keep) {
          const float dsdt = G * b.Texcl * expf(-b.sdt) - (float)suffix;
          gs = dsdt * __fsub_rn(b.t1, b.t0) * b.dact_s;
        }
        g[0] = gs;
        gabs = fmaxf(gabs, fabsf(gs));
        for (int c = 0; c < p.C; ++c) {
          const float gcv = b.keep ? gC[c] * b.w * b.dact_c[c] : 0.f;
          g[1 + c] = gcv;
          gabs = fmaxf(gabs, fabsf(gcv));
        }
      }
    });
  }
  if (a.g_bkgd && a.bkgd && lane == 0)
    for (int c = 0; c < p.C; ++c)
      if (gb_acc[c] != 0.f) atomicAdd(a.g_bkgd + c, gb_acc[c]);
  if (a.gmax != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gabs = fmaxf(gabs, __shfl_xor_sync(0xffffffffu, gabs, o));
    if (lane == 0 && gabs > 0.f && gabs < INFINITY) atomicMax(a.gmax, __float_as_uint(gabs));
  }
}

// ------------------------------------------------------------------ tangent mode
// With a_i = sigma_i dt_i, E_i = sum_{j<i} a_j, T_i = e^{-E_i}, beta_i = e^{-a_i}, w_i = T_i (1 - beta_i):
//   wdot_i = -Edot_i w_i + T_i beta_i adot_i,   Cdot = sum wdot c + w cdot - bkgd Odot,   Odot = sum wdot
// (the d/dw image of external/vol_rendering.py:82-126; the visibility cull is piecewise constant).
struct TanBlk {
  float adot, cdot[3], sdot_raw, cdot_raw[3], wdot, Edot;
};
// per-block tangent quantities; `carry` = sum of kept adot before this block (updated)
__device__ __forceinline__ TanBlk tan_block(const CompArgs& a, long long r, int k, int lane, const SampleBlk& b,
                                            float& carry) {
  const RenPlan& p = a.p;
  const int C1 = 1 + p.C;
  TanBlk t;
  t.sdot_raw = 0.f;
  t.cdot_raw[0] = t.cdot_raw[1] = t.cdot_raw[2] = 0.f;
  if (b.keep) {
    const float* rw = a.raw_dot ? a.raw_dot + b.sidx * C1 : a.raw + (b.sidx * 2 + 1) * C1;
    t.sdot_raw = rw[0];
    for (int c = 0; c < p.C; ++c) t.cdot_raw[c] = rw[1 + c];
  }
  t.adot = b.keep ? b.dact_s * t.sdot_raw * __fsub_rn(b.t1, b.t0) : 0.f;
  for (int c = 0; c < 3; ++c) t.cdot[c] = b.dact_c[c] * t.cdot_raw[c];
  const float incl = warp_incl_scan(t.adot, lane);
  t.Edot = carry + (incl - t.adot);
  carry += __shfl_sync(0xffffffffu, incl, 31);
  t.wdot = b.keep ? fmaf(-t.Edot, b.w, b.Texcl * expf(-b.sdt) * t.adot) : 0.f;
  return t;
}

extern "C" __global__ void __launch_bounds__(128) ren_composite_tan_fwd(const CompArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const RenPlan& p = a.p;
  for (long long r = warp; r < a.n_rays; r += n_warps) {
    float accC[3] = {0.f, 0.f, 0.f}, accO = 0.f, accD = 0.f, accCd[3] = {0.f, 0.f, 0.f}, accOd = 0.f;
    int kept = 0;
    float carry = 0.f;
    walk_ray(a, r, lane, [&](int k, const SampleBlk& b) {
      const TanBlk t = tan_block(a, r, k, lane, b, carry);
      if (b.keep) {
        for (int c = 0; c < p.C; ++c) {
          accC[c] = fmaf(b.w, b.c[c], accC[c]);
          accCd[c] = fmaf(t.wdot, b.c[c], fmaf(b.w, t.cdot[c], accCd[c]));
        }
        accO += b.w;
        accOd += t.wdot;
        accD = fmaf(b.w, __fmul_rn(__fadd_rn(b.t0, b.t1), 0.5f), accD);
        kept += 1;
      }
    });
    accO = warp_sum(accO);
    accD = warp_sum(accD);
    accOd = warp_sum(accOd);
    for (int c = 0; c < p.C; ++c) { accC[c] = warp_sum(accC[c]); accCd[c] = warp_sum(accCd[c]); }
    kept = __reduce_add_sync(0xffffffffu, kept);
    if (lane == 0) {
      for (int c = 0; c < p.C; ++c) {
        float v = accC[c], vd = accCd[c];
        if (a.bkgd) {
          v = __fadd_rn(v, __fmul_rn(a.bkgd[c], __fsub_rn(1.f, accO)));
          vd = fmaf(-a.bkgd[c], accOd, vd);
        }
        a.radiance[r * p.C + c] = v;
        a.radiance_dot[r * p.C + c] = vd;
      }
      a.opacity[r] = accO;
      a.opacity_dot[r] = accOd;
      a.depth[r] = accD;
      a.n_kept[r] = kept;
    }
  }
}

// Reverse of the tangent compositing.  Per kept sample, with
//   Wv_i = sum_c gC_c (c_i - b_c) + gO + gD tm_i + sum_c gCd_c cdot_i,   Wt_i = sum_c gCd_c (c_i - b_c) + gOd:
//   dL/dcdot_i = gCd w_i,   dL/dc_i = gC w_i + gCd wdot_i
//   dL/dadot_i = Wt_i T_i beta_i - sum_{k>i} Wt_k w_k
//   dL/da_i    = Wv_i T_i beta_i - Wt_i T_i beta_i (Edot_i + adot_i) - sum_{k>i} (Wv_k w_k + Wt_k wdot_k)
// then through sigma = f(raw), sigmadot = f'(raw) rawdot and c = r(raw), cdot = r'(raw) rawdot.
extern "C" __global__ void __launch_bounds__(128) ren_composite_tan_bwd(const CompArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const RenPlan& p = a.p;
  const int C1 = 1 + p.C;
  float gb_acc[3] = {0.f, 0.f, 0.f};
  float gabs = 0.f;
  for (long long r = warp; r < a.n_rays; r += n_warps) {
    float gC[3] = {0.f, 0.f, 0.f}, gCd[3] = {0.f, 0.f, 0.f}, bk[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < p.C; ++c) {
      gC[c] = a.g_radiance ? a.g_radiance[r * p.C + c] : 0.f;
      gCd[c] = a.g_radiance_dot ? a.g_radiance_dot[r * p.C + c] : 0.f;
      bk[c] = a.bkgd ? a.bkgd[c] : 0.f;
    }
    const float gO = a.g_opacity ? a.g_opacity[r] : 0.f;
    const float gOd = a.g_opacity_dot ? a.g_opacity_dot[r] : 0.f;
    const float gD = a.g_depth ? a.g_depth[r] : 0.f;
    auto weights = [&](const SampleBlk& b, const TanBlk& t, float& Wv, float& Wt) {
      Wv = gO + gD * __fmul_rn(__fadd_rn(b.t0, b.t1), 0.5f);
      Wt = gOd;
      for (int c = 0; c < p.C; ++c) {
        Wv = fmaf(gC[c], b.c[c] - bk[c], fmaf(gCd[c], t.cdot[c], Wv));
        Wt = fmaf(gCd[c], b.c[c] - bk[c], Wt);
      }
    };
    // pass 1: totals
    double totA = 0.0, totB = 0.0;     // sum (Wv w + Wt wdot), sum Wt w
    float accO = 0.f, accOd = 0.f, carry = 0.f;
    walk_ray(a, r, lane, [&](int k, const SampleBlk& b) {
      const TanBlk t = tan_block(a, r, k, lane, b, carry);
      if (b.keep) {
        float Wv, Wt;
        weights(b, t, Wv, Wt);
        totA += (double)Wv * (double)b.w + (double)Wt * (double)t.wdot;
        totB += (double)Wt * (double)b.w;
        accO += b.w;
        accOd += t.wdot;
      }
    });
    totA = warp_sum_d(totA);
    totB = warp_sum_d(totB);
    accO = warp_sum(accO);
    accOd = warp_sum(accOd);
    if (a.g_bkgd && a.bkgd)
      for (int c = 0; c < p.C; ++c) gb_acc[c] += gC[c] * (1.f - accO) - gCd[c] * accOd;
    // pass 2
    double prefA = 0.0, prefB = 0.0;
    carry = 0.f;
    walk_ray(a, r, lane, [&](int k, const SampleBlk& b) {
      const TanBlk t = tan_block(a, r, k, lane, b, carry);
      float Wv = 0.f, Wt = 0.f;
      double ga = 0.0, gbv = 0.0;
      if (b.keep) {
        weights(b, t, Wv, Wt);
        ga = (double)Wv * (double)b.w + (double)Wt * (double)t.wdot;
        gbv = (double)Wt * (double)b.w;
      }
      double inclA = ga, inclB = gbv;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        double ta = __shfl_up_sync(0xffffffffu, inclA, off);
        double tb = __shfl_up_sync(0xffffffffu, inclB, off);
        if (lane >= off) { inclA += ta; inclB += tb; }
      }
      const double sufA = totA - (prefA + inclA);
      const double sufB = totB - (prefB + inclB);
      prefA += __shfl_sync(0xffffffffu, inclA, 31);
      prefB += __shfl_sync(0xffffffffu, inclB, 31);
      if (k < a.K && (a.ray_offsets == nullptr || b.alive)) {
        const long long vrow = b.sidx * (a.raw_dot ? 1 : 2);
        float* gv = a.graw + (a.graw_dot ? b.sidx : b.sidx * 2) * C1;
        float* gt = a.graw_dot ? a.graw_dot + b.sidx * C1 : gv + C1;
        float gs = 0.f, gsd = 0.f;
        if (b.keep) {
          const float dt = __fsub_rn(b.t1, b.t0);
          const float Tb = b.Texcl * expf(-b.sdt);
          const float dadot = Wt * Tb - (float)sufB;
          const float da = Wv * Tb - Wt * Tb * (t.Edot + t.adot) - (float)sufA;
          const float sraw = a.raw[vrow * C1];
          gsd = dadot * dt * b.dact_s;
          gs = da * dt * b.dact_s + dadot * dt * ren_density_act2(p.f.density_activation, sraw) * t.sdot_raw;
        }
        gv[0] = gs;
        gt[0] = gsd;
        gabs = fmaxf(gabs, fmaxf(fabsf(gs), fabsf(gsd)));
        for (int c = 0; c < p.C; ++c) {
          float gcv = 0.f, gcd = 0.f;
          if (b.keep) {
            const float craw = a.raw[vrow * C1 + 1 + c];
            const float dc = gC[c] * b.w + gCd[c] * t.wdot;
            const float dcd = gCd[c] * b.w;
            gcd = dcd * b.dact_c[c];
            gcv = dc * b.dact_c[c] + dcd * ren_radiance_act2(p.f.radiance_activation, craw) * t.cdot_raw[c];
          }
          gv[1 + c] = gcv;
          gt[1 + c] = gcd;
          gabs = fmaxf(gabs, fmaxf(fabsf(gcv), fabsf(gcd)));
        }
      }
    });
  }
  if (a.g_bkgd && a.bkgd && lane == 0)
    for (int c = 0; c < p.C; ++c)
      if (gb_acc[c] != 0.f) atomicAdd(a.g_bkgd + c, gb_acc[c]);
  if (a.gmax != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gabs = fmaxf(gabs, __shfl_xor_sync(0xffffffffu, gabs, o));
    if (lane == 0 && gabs > 0.f && gabs < INFINITY) atomicMax(a.gmax, __float_as_uint(gabs));
  }
}

// ------------------------------------------------------------------ live tiles (occupancy-grid marching)
// live[t] = 1 iff any of the 128 dense marching steps of tile t emits a sample (midpoint before t_max and in an
// occupied cell).  The tcgen05 kernels skip dead tiles entirely.
struct LiveArgs {
  RenPlan p;
  long long n_rays;
  int K;
  const float* rays_o;
  const float* rays_d;
  const float* jitter;
  unsigned char* live;
};
extern "C" __global__ void __launch_bounds__(128) ren_tile_live_kernel(const LiveArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const RenPlan& p = a.p;
  for (long long r = warp; r < a.n_rays; r += n_warps) {
    float o[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = a.rays_o[r * 3 + c];
      d[c] = a.rays_d[r * 3 + c];
    }
    float t_min, t_max, t0, t1;
    ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
    ren_march_to(p.r, t_min, lane, &t0, &t1);
    for (int kb = 0; kb < a.K; kb += 32) {
      const int k = kb + lane;
      const bool in_range = (k < a.K) && ren_alive(t0, t1, t_max);
      if (in_range && ren_occupied(p.f, p.r, o, d, t0, t1)) a.live[(r * a.K + k) >> 7] = 1;
      if (__ballot_sync(0xffffffffu, in_range) == 0u) break;
      for (int i = 0; i < 32; ++i) {
        t0 = t1;
        t1 = __fadd_rn(t0, ren_dt(p.r, t0));
      }
    }
  }
}
// ------------------------------------------------------------------ packed samples
// nerfacc.ray_marching emits a packed list of samples (ray_indices, t_starts, t_ends; external/utils.py:106-119).
// Phase 1 counts the emitted steps per ray, phase 2 (after an exclusive scan of the counts) writes the dense
// index r*K + k of every emitted step at its packed position.
struct PackArgs2 {
  RenPlan p;
  long long n_rays;
  int K;
  const float* rays_o;
  const float* rays_d;
  const float* jitter;
  int* counts;               // phase 1 out: [n_rays]
  const int* ray_offsets;    // phase 2 in: exclusive scan of counts, [n_rays + 1]
  int* sample_idx;           // phase 2 out: [total]
};
template <bool FILL>
__global__ void __launch_bounds__(128) ren_march_pack_kernel(const PackArgs2 a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const RenPlan& p = a.p;
  for (long long r = warp; r < a.n_rays; r += n_warps) {
    float o[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = a.rays_o[r * 3 + c];
      d[c] = a.rays_d[r * 3 + c];
    }
    float t_min, t_max, t0, t1;
    ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
    ren_march_to(p.r, t_min, lane, &t0, &t1);
    int base = FILL ? a.ray_offsets[r] : 0;
    for (int kb = 0; kb < a.K; kb += 32) {
      const int k = kb + lane;
      const bool in_range = (k < a.K) && ren_alive(t0, t1, t_max);
      const bool emitted = in_range && ren_occupied(p.f, p.r, o, d, t0, t1);
      const unsigned em = __ballot_sync(0xffffffffu, emitted);
      if (FILL && emitted) a.sample_idx[base + __popc(em & ((1u << lane) - 1u))] = (int)(r * a.K + k);
      base += __popc(em);
      if (__ballot_sync(0xffffffffu, in_range) == 0u) break;
      for (int i = 0; i < 32; ++i) {
        t0 = t1;
        t1 = __fadd_rn(t0, ren_dt(p.r, t0));
      }
    }
    if (!FILL && lane == 0) a.counts[r] = base;
  }
}
int ren_launch_march_pack(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                          const float* jitter, int* counts, const int* ray_offsets, int* sample_idx,
                          cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  PackArgs2 a;
  a.p = p; a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter;
  a.counts = counts; a.ray_offsets = ray_offsets; a.sample_idx = sample_idx;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > 148 * 64) blocks = 148 * 64;
  ren_prof_begin(REN_K_COMPOSITE_FWD, st);
  if (sample_idx) ren_march_pack_kernel<true><<<(unsigned)blocks, 128, 0, st>>>(a);
  else ren_march_pack_kernel<false><<<(unsigned)blocks, 128, 0, st>>>(a);
  ren_prof_end(REN_K_COMPOSITE_FWD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("march pack launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

int ren_launch_tile_live(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                         const float* jitter, unsigned char* live, cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  const long long n_tiles = (n_rays * K + 127) >> 7;
  if (cudaMemsetAsync(live, 0, (size_t)n_tiles, st) != cudaSuccess) { ren_set_error("live-tile memset failed"); return REN_E_CUDA; }
  LiveArgs a;
  a.p = p; a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter; a.live = live;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > 148 * 64) blocks = 148 * 64;
  ren_prof_begin(REN_K_COMPOSITE_FWD, st);
  ren_tile_live_kernel<<<(unsigned)blocks, 128, 0, st>>>(a);
  ren_prof_end(REN_K_COMPOSITE_FWD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("live-tile launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

int ren_launch_composite_tan(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                             const float* jitter, const int* ray_offsets, const float* raw, const float* raw_dot,
                             const float* bkgd, float* radiance,
                             float* opacity, float* depth, float* radiance_dot, float* opacity_dot,
                             int* n_kept, cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  CompArgs a = {};
  a.p = p; a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter; a.raw = raw;
  a.bkgd = bkgd; a.radiance = radiance; a.opacity = opacity; a.depth = depth; a.n_kept = n_kept;
  a.tangent = 1; a.radiance_dot = radiance_dot; a.opacity_dot = opacity_dot; a.raw_dot = raw_dot;
  a.ray_offsets = ray_offsets;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > 148 * 64) blocks = 148 * 64;
  ren_prof_begin(REN_K_COMPOSITE_FWD, st);
  ren_composite_tan_fwd<<<(unsigned)blocks, 128, 0, st>>>(a);
  ren_prof_end(REN_K_COMPOSITE_FWD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("composite tangent fwd launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

int ren_launch_composite_tan_bwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                                 const float* jitter, const int* ray_offsets, const float* raw, const float* raw_dot,
                                 const float* bkgd,
                                 const float* g_radiance, const float* g_opacity, const float* g_depth,
                                 const float* g_radiance_dot, const float* g_opacity_dot, float* graw,
                                 float* graw_dot, float* g_bkgd, unsigned int* gmax, cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  CompArgs a = {};
  a.p = p; a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter; a.raw = raw;
  a.bkgd = bkgd; a.g_radiance = g_radiance; a.g_opacity = g_opacity; a.g_depth = g_depth;
  a.g_radiance_dot = g_radiance_dot; a.g_opacity_dot = g_opacity_dot;
  a.graw = graw; a.g_bkgd = g_bkgd; a.tangent = 1; a.raw_dot = raw_dot; a.graw_dot = graw_dot; a.gmax = gmax;
  a.ray_offsets = ray_offsets;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > 148 * 64) blocks = 148 * 64;
  ren_prof_begin(REN_K_COMPOSITE_BWD, st);
  ren_composite_tan_bwd<<<(unsigned)blocks, 128, 0, st>>>(a);
  ren_prof_end(REN_K_COMPOSITE_BWD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("composite tangent bwd launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

int ren_launch_composite(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                         const float* jitter, const int* ray_offsets, const float* raw, const float* bkgd,
                         float* radiance, float* opacity, float* depth, int* n_kept, cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  CompArgs a = {};
  a.p = p; a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter; a.raw = raw;
  a.bkgd = bkgd; a.radiance = radiance; a.opacity = opacity; a.depth = depth; a.n_kept = n_kept;
  a.ray_offsets = ray_offsets;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > 148 * 64) blocks = 148 * 64;
  ren_prof_begin(REN_K_COMPOSITE_FWD, st);
  ren_composite_fwd<<<(unsigned)blocks, 128, 0, st>>>(a);
  ren_prof_end(REN_K_COMPOSITE_FWD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("composite fwd launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

int ren_launch_composite_bwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                             const float* jitter, const int* ray_offsets, const float* raw, const float* bkgd,
                             const float* g_radiance, const float* g_opacity, const float* g_depth,
                             float* graw, float* g_bkgd, unsigned int* gmax, cudaStream_t st) {
  if (n_rays <= 0) return REN_OK;
  CompArgs a = {};
  a.p = p; a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter; a.raw = raw;
  a.bkgd = bkgd; a.g_radiance = g_radiance; a.g_opacity = g_opacity; a.g_depth = g_depth;
  a.graw = graw; a.g_bkgd = g_bkgd; a.gmax = gmax; a.ray_offsets = ray_offsets;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > 148 * 64) blocks = 148 * 64;
  ren_prof_begin(REN_K_COMPOSITE_BWD, st);
  ren_composite_bwd<<<(unsigned)blocks, 128, 0, st>>>(a);
  ren_prof_end(REN_K_COMPOSITE_BWD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("composite bwd launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
