// ren_event.cu -- per-event prologue / epilogue of the training step and the fused Adam.
//
//   ren_event_rays_kernel   refractory-period correction, supervision timestamps,
//                           searchsorted + lerp/slerp pose, ray generation
//                           (models/event_generation_params.py:196-203,
//                            models/robust_e_nerf.py:322-357, models/trajectories.py:30-91,
//                            utils/tensor_ops.py:83-180, models/nerf.py:206-228)
//   ren_diff_loss_kernel*   log-intensity-difference loss + its gradient
//                           (models/robust_e_nerf.py:865-871,432-443, loss_metric/loss.py:64-74,
//                            utils/modules.py:77-102)
//   ren_adam_kernel         torch.optim.Adam step on a flat buffer
#include "ren_common.cuh"
#include "ren_dual.cuh"

// torch.lerp: |w| < 0.5 ? a + w (b-a) : b - (b-a)(1-w), both as single FMAs
__device__ __forceinline__ double lerp_d(double a, double b, double w) {
  const double diff = b - a;
  return (fabs(w) < 0.5) ? fma(w, diff, a) : fma(-(1.0 - w), diff, b);
}
__device__ __forceinline__ float lerp_f(float a, float b, float w) {
  const float diff = __fsub_rn(b, a);
  return (fabsf(w) < 0.5f) ? fmaf(w, diff, a) : fmaf(-__fsub_rn(1.f, w), diff, b);
}

// Hamilton product, XYZW (roma 1.2.7 quat_product)
__device__ __forceinline__ void quat_mul(const float* p, const float* q, float* out) {
  // vector = p.w*q.xyz + q.w*p.xyz + cross(p.xyz, q.xyz) ; summed left to right like torch
  float cx = __fsub_rn(__fmul_rn(p[1], q[2]), __fmul_rn(p[2], q[1]));
  float cy = __fsub_rn(__fmul_rn(p[2], q[0]), __fmul_rn(p[0], q[2]));
  float cz = __fsub_rn(__fmul_rn(p[0], q[1]), __fmul_rn(p[1], q[0]));
  out[0] = __fadd_rn(__fadd_rn(__fmul_rn(p[3], q[0]), __fmul_rn(q[3], p[0])), cx);
  out[1] = __fadd_rn(__fadd_rn(__fmul_rn(p[3], q[1]), __fmul_rn(q[3], p[1])), cy);
  out[2] = __fadd_rn(__fadd_rn(__fmul_rn(p[3], q[2]), __fmul_rn(q[3], p[2])), cz);
  float dot = __fadd_rn(__fadd_rn(__fmul_rn(p[0], q[0]), __fmul_rn(p[1], q[1])), __fmul_rn(p[2], q[2]));
  out[3] = __fsub_rn(__fmul_rn(p[3], q[3]), dot);
}
__device__ __forceinline__ float norm3(const float* v) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
}

// First and second derivative of the ray (o, d) of a pixel wrt the pose-interpolation weight w
// (models/trajectories.py:63-91 -> utils/tensor_ops.py:157-179 -> roma rotvec_to_unitquat /
// unitquat_to_rotmat -> models/nerf.py:206-228), by second-order jets through the same operation
// sequence the reference differentiates with autograd.  The relative rotation vector does not
// depend on w.  o = lerp(P_l, P_r, w): d o / d w = P_r - P_l, second derivative 0.
__device__ __forceinline__ void quat_mul_j(const J2* p, const J2* q, J2* out) {
  out[0] = p[3] * q[0] + q[3] * p[0] + (p[1] * q[2] - p[2] * q[1]);
  out[1] = p[3] * q[1] + q[3] * p[1] + (p[2] * q[0] - p[0] * q[2]);
  out[2] = p[3] * q[2] + q[3] * p[2] + (p[0] * q[1] - p[1] * q[0]);
  out[3] = p[3] * q[3] - (p[0] * q[0] + p[1] * q[1] + p[2] * q[2]);
}
__device__ __forceinline__ void ray_dir_jet(float w, const float* q0, const float* q1in, const float* vpix,
                                            float* d_dot, float* d_ddot) {
  float q1[4];
  float dot = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0[0], q1in[0]), __fmul_rn(q0[1], q1in[1])),
                                  __fmul_rn(q0[2], q1in[2])), __fmul_rn(q0[3], q1in[3]));
  for (int c = 0; c < 4; ++c) q1[c] = dot < 0.f ? -q1in[c] : q1in[c];
  const float conj[4] = {-q0[0], -q0[1], -q0[2], q0[3]};
  float rel[4];
  rel[0] = conj[3] * q1[0] + q1[3] * conj[0] + (conj[1] * q1[2] - conj[2] * q1[1]);
  rel[1] = conj[3] * q1[1] + q1[3] * conj[1] + (conj[2] * q1[0] - conj[0] * q1[2]);
  rel[2] = conj[3] * q1[2] + q1[3] * conj[2] + (conj[0] * q1[1] - conj[1] * q1[0]);
  rel[3] = conj[3] * q1[3] - (conj[0] * q1[0] + conj[1] * q1[1] + conj[2] * q1[2]);
  const float ang = 2.f * atan2f(sqrtf(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]), rel[3]);
  float scale;
  if (fabsf(ang) <= 1e-3f) {
    const float a2 = ang * ang;
    scale = 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f;
  } else {
    scale = ang / sinf(0.5f * ang);
  }
  const J2 wj(w, 1.f, 0.f);
  J2 rv[3];
  for (int c = 0; c < 3; ++c) rv[c] = wj * J2(scale * rel[c]);
  const J2 n = j2_sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  J2 s;
  if (n.v <= 1e-3f) {
    const J2 n2 = n * n;
    s = J2(0.5f) - n2 * J2(1.f / 48.f) + n2 * n2 * J2(1.f / 3840.f);
  } else {
    s = j2_sin(n * J2(0.5f)) / n;
  }
  const J2 rot[4] = {s * rv[0], s * rv[1], s * rv[2], j2_cos(n * J2(0.5f))};
  const J2 q0j[4] = {J2(q0[0]), J2(q0[1]), J2(q0[2]), J2(q0[3])};
  J2 q[4];
  quat_mul_j(q0j, rot, q);
  const J2 x = q[0], y = q[1], z = q[2], ww = q[3];
  const J2 x2 = x * x, y2 = y * y, z2 = z * z, w2 = ww * ww;
  const J2 xy = x * y, zw = z * ww, xz = x * z, yw = y * ww, yz = y * z, xw = x * ww;
  const J2 two(2.f);
  const J2 R[9] = {x2 - y2 - z2 + w2, two * (xy - zw), two * (xz + yw),
                   two * (xy + zw), y2 - x2 - z2 + w2, two * (yz - xw),
                   two * (xz - yw), two * (yz + xw), z2 - x2 - y2 + w2};
  J2 dv[3];
  for (int r = 0; r < 3; ++r) dv[r] = R[r * 3] * J2(vpix[0]) + R[r * 3 + 1] * J2(vpix[1]) + R[r * 3 + 2] * J2(vpix[2]);
  const J2 nrm = j2_sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
  for (int c = 0; c < 3; ++c) {
    const J2 d = dv[c] / nrm;
    d_dot[c] = d.d;
    d_ddot[c] = d.dd;
  }
}

struct PoseBin { long long left, right; double width; float w; };

// LinearTrajectory.forward for one timestamp -> position p[3], rotation R[9] (row-major)
__device__ __forceinline__ void pose_at(double ts, long long n_poses, const long long* __restrict__ T,
                                        const float* __restrict__ P, const float* __restrict__ Q,
                                        float* p, float* R, PoseBin* bin = nullptr) {
  // searchsorted (left): first i with T[i] >= ts
  long long lo = 0, hi = n_poses;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if ((double)T[mid] < ts) lo = mid + 1; else hi = mid;
  }
  long long right = lo;
  long long left = (ts == (double)T[0]) ? right : right - 1;
  if (left < 0) left = 0;
  if (right > n_poses - 1) right = n_poses - 1;
  long long wl = left < n_poses - 1 ? left : n_poses - 2;
  const double width = (double)(T[wl + 1] - T[wl]);
  const float w = (float)((ts - (double)T[left]) / width);
  if (bin) { bin->left = left; bin->right = right; bin->width = width; bin->w = w; }
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = lerp_f(P[left * 3 + c], P[right * 3 + c], w);
  // slerp (utils/tensor_ops.py:114-180, shortest path)
  float q0[4], q1[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { q0[c] = Q[left * 4 + c]; q1[c] = Q[right * 4 + c]; }
  float dot = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0[0], q1[0]), __fmul_rn(q0[1], q1[1])),
                                  __fmul_rn(q0[2], q1[2])), __fmul_rn(q0[3], q1[3]));
  if (dot < 0.f) {
#pragma unroll
    for (int c = 0; c < 4; ++c) q1[c] = -q1[c];
  }
  float conj[4] = {-q0[0], -q0[1], -q0[2], q0[3]};
  float rel[4];
  quat_mul(conj, q1, rel);
  // unitquat_to_full_rotvec (tensor_ops.py:83-111)
  float ang = __fmul_rn(2.f, atan2f(norm3(rel), rel[3]));
  float scale;
  if (fabsf(ang) <= 1e-3f) {
    float a2 = __fmul_rn(ang, ang);
    float a4 = __fmul_rn(a2, a2);
    scale = __fadd_rn(__fadd_rn(2.f, __fdiv_rn(a2, 12.f)), __fdiv_rn(__fmul_rn(7.f, a4), 2880.f));
  } else {
    scale = __fdiv_rn(ang, sinf(__fmul_rn(ang, 0.5f)));
  }
  float rv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) rv[c] = __fmul_rn(w, __fmul_rn(scale, rel[c]));
  // roma rotvec_to_unitquat
  float n = norm3(rv);
  float s;
  if (n <= 1e-3f) {
    float n2 = __fmul_rn(n, n);
    float n4 = __fmul_rn(n2, n2);
    s = __fadd_rn(__fsub_rn(0.5f, __fdiv_rn(n2, 48.f)), __fdiv_rn(n4, 3840.f));
  } else {
    s = __fdiv_rn(sinf(__fmul_rn(n, 0.5f)), n);
  }
  float rot[4] = {__fmul_rn(s, rv[0]), __fmul_rn(s, rv[1]), __fmul_rn(s, rv[2]), cosf(__fmul_rn(n, 0.5f))};
  float q[4];
  quat_mul(q0, rot, q);
  // unitquat_to_rotmat (roma, no renormalisation)
  const float x = q[0], y = q[1], z = q[2], ww = q[3];
  const float x2 = __fmul_rn(x, x), y2 = __fmul_rn(y, y), z2 = __fmul_rn(z, z), w2 = __fmul_rn(ww, ww);
  const float xy = __fmul_rn(x, y), zw = __fmul_rn(z, ww), xz = __fmul_rn(x, z), yw = __fmul_rn(y, ww);
  const float yz = __fmul_rn(y, z), xw = __fmul_rn(x, ww);
  R[0] = __fadd_rn(__fsub_rn(__fsub_rn(x2, y2), z2), w2);
  R[1] = __fmul_rn(2.f, __fsub_rn(xy, zw));
  R[2] = __fmul_rn(2.f, __fadd_rn(xz, yw));
  R[3] = __fmul_rn(2.f, __fadd_rn(xy, zw));
  R[4] = __fadd_rn(__fsub_rn(__fadd_rn(-x2, y2), z2), w2);
  R[5] = __fmul_rn(2.f, __fsub_rn(yz, xw));
  R[6] = __fmul_rn(2.f, __fsub_rn(xz, yw));
  R[7] = __fmul_rn(2.f, __fadd_rn(yz, xw));
  R[8] = __fadd_rn(__fadd_rn(__fsub_rn(-x2, y2), z2), w2);
}

struct EventRaysArgs {
  long long B;
  const float* position;
  const long long* start_ts;
  const long long* end_ts;
  const double* nz_ts_diff;
  const double* nz_diff_start;
  const double* nz_grad_ts;
  const double* tau;
  long long n_poses;
  const long long* pose_ts;
  const float* pose_pos;
  const float* pose_quat;
  const float* k_inv;
  int n_renders;
  double* out_ts;
  float* out_o;
  float* out_d;
  float* out_rot;
  double* out_ts_diff;
  double* out_start;
  float* out_o_dot;     // (n_renders,B,3) d o / d w          (all four NULL or all set)
  float* out_d_dot;     // (n_renders,B,3) d d / d w
  float* out_d_ddot;    // (n_renders,B,3) d^2 d / d w^2
  double* out_dw_dt;    // (n_renders,B)   d w / d t = 1 / bin width (ns^-1)
};

extern "C" __global__ void ren_event_rays_kernel(const EventRaysArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B) return;
  const double start = (double)a.start_ts[i] + a.tau[0];  // event_generation_params.py:196-203
  const double end = (double)a.end_ts[i];
  // robust_e_nerf.py:322-357
  const double ts_diff = (end - start) * a.nz_ts_diff[i];
  const double d_start = lerp_d(start, fmax(end - ts_diff, start), a.nz_diff_start[i]);
  const double d_end = fmin(d_start + ts_diff, end);
  double ts[3];
  int nr = 0;
  if (a.n_renders == 3) ts[nr++] = lerp_d(d_start, d_end, a.nz_grad_ts[i]);
  ts[nr++] = d_start;
  ts[nr++] = d_end;
  a.out_ts_diff[i] = ts_diff;
  a.out_start[i] = start;
  // models/nerf.py:206-228: d = normalize(R (K^-1 [x, y, 1])); the two mat-vecs are FMA chains
  // over k like the BLAS kernels torch dispatches `@` to
  const float px = a.position[i * 2], py = a.position[i * 2 + 1];
  float v[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    v[r] = fmaf(a.k_inv[r * 3 + 2], 1.f, fmaf(a.k_inv[r * 3 + 1], py, __fmul_rn(a.k_inv[r * 3], px)));
  for (int j = 0; j < nr; ++j) {
    float p[3], R[9];
    PoseBin bin;
    pose_at(ts[j], a.n_poses, a.pose_ts, a.pose_pos, a.pose_quat, p, R, &bin);
    float dv[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      dv[r] = fmaf(R[r * 3 + 2], v[2], fmaf(R[r * 3 + 1], v[1], __fmul_rn(R[r * 3], v[0])));
    const float nrm = norm3(dv);
    const long long o = (long long)j * a.B + i;
    a.out_ts[o] = ts[j];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a.out_o[o * 3 + c] = p[c];
      a.out_d[o * 3 + c] = __fdiv_rn(dv[c], nrm);
    }
    if (a.out_rot)
      for (int c = 0; c < 9; ++c) a.out_rot[o * 9 + c] = R[c];
    if (a.out_o_dot) {
      float dd[3], ddd[3];
      ray_dir_jet(bin.w, a.pose_quat + bin.left * 4, a.pose_quat + bin.right * 4, v, dd, ddd);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        a.out_o_dot[o * 3 + c] = a.pose_pos[bin.right * 3 + c] - a.pose_pos[bin.left * 3 + c];
        a.out_d_dot[o * 3 + c] = dd[c];
        a.out_d_ddot[o * 3 + c] = ddd[c];
      }
      a.out_dw_dt[o] = 1.0 / bin.width;
    }
  }
}

extern "C" int ren_event_rays(int64_t B, const float* position, const int64_t* start_ts,
                              const int64_t* end_ts, const double* nz_ts_diff,
                              const double* nz_diff_start, const double* nz_grad_ts,
                              const double* tau, int64_t n_poses, const int64_t* pose_ts,
                              const float* pose_pos, const float* pose_quat, const float* k_inv,
                              int32_t n_renders, double* out_ts, float* out_o, float* out_d,
                              float* out_rot, double* out_ts_diff, double* out_start, float* out_o_dot,
                              float* out_d_dot, float* out_d_ddot, double* out_dw_dt, void* stream) {
  if (B < 0 || n_poses < 2 || (n_renders != 2 && n_renders != 3)) {
    ren_set_error("ren_event_rays: bad sizes (B=%lld, n_poses=%lld, n_renders=%d)", (long long)B,
                  (long long)n_poses, n_renders);
    return REN_E_INVALID;
  }
  if (n_renders == 3 && nz_grad_ts == nullptr) {
    ren_set_error("ren_event_rays: n_renders=3 needs nz_grad_ts");
    return REN_E_INVALID;
  }
  const int n_jet = (out_o_dot != nullptr) + (out_d_dot != nullptr) + (out_d_ddot != nullptr) + (out_dw_dt != nullptr);
  if (n_jet != 0 && n_jet != 4) {
    ren_set_error("ren_event_rays: the four ray-derivative outputs must be all NULL or all set");
    return REN_E_INVALID;
  }
  if (B == 0) return REN_OK;
  EventRaysArgs a;
  a.out_o_dot = out_o_dot; a.out_d_dot = out_d_dot; a.out_d_ddot = out_d_ddot; a.out_dw_dt = out_dw_dt;
  a.B = B; a.position = position; a.start_ts = (const long long*)start_ts;
  a.end_ts = (const long long*)end_ts; a.nz_ts_diff = nz_ts_diff; a.nz_diff_start = nz_diff_start;
  a.nz_grad_ts = nz_grad_ts; a.tau = tau; a.n_poses = n_poses; a.pose_ts = (const long long*)pose_ts;
  a.pose_pos = pose_pos; a.pose_quat = pose_quat; a.k_inv = k_inv; a.n_renders = n_renders;
  a.out_ts = out_ts; a.out_o = out_o; a.out_d = out_d; a.out_rot = out_rot;
  a.out_ts_diff = out_ts_diff; a.out_start = out_start;
  ren_prof_begin(REN_K_EVENT, (cudaStream_t)stream);
  ren_event_rays_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  ren_prof_end(REN_K_EVENT, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ren_event_rays launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

extern "C" __global__ void ren_trajectory_kernel(long long n, const double* __restrict__ ts,
                                                 long long n_poses, const long long* __restrict__ T,
                                                 const float* __restrict__ P, const float* __restrict__ Q,
                                                 float* __restrict__ out_pos, float* __restrict__ out_rot) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3], R[9];
  pose_at(ts[i], n_poses, T, P, Q, p, R);
  for (int c = 0; c < 3; ++c) out_pos[i * 3 + c] = p[c];
  for (int c = 0; c < 9; ++c) out_rot[i * 9 + c] = R[c];
}

extern "C" int ren_trajectory(int64_t n, const double* ts, int64_t n_poses, const int64_t* pose_ts,
                              const float* pose_pos, const float* pose_quat, float* out_pos,
                              float* out_rot, void* stream) {
  if (n < 0 || n_poses < 2) { ren_set_error("ren_trajectory: bad sizes"); return REN_E_INVALID; }
  if (n == 0) return REN_OK;
  ren_prof_begin(REN_K_EVENT, (cudaStream_t)stream);
  ren_trajectory_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      n, ts, n_poses, (const long long*)pose_ts, pose_pos, pose_quat, out_pos, out_rot);
  ren_prof_end(REN_K_EVENT, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ren_trajectory launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

// ------------------------------------------------------------------ diff loss
struct DiffLossArgs {
  long long B;
  int error_fn, all_valid;
  float imin;
  const float *c0, *c1, *o0, *o1, *ev_diff;
  const double *ts_diff, *start;
  const long long* end_ts;
  float *log_i0, *log_i1, *loss;
  int* n_valid;
  float *g_c0, *g_c1;
  double* acc;   // scratch[0]: sum of errors ; scratch[1] (as int): valid count
};

__device__ __forceinline__ float diff_err(int fn, float pred, float tgt, float* dpred) {
  const float e = pred - tgt;
  if (fn == REN_ERR_MSE) { *dpred = 2.f * e; return e * e; }
  const float sgn = (e > 0.f) ? 1.f : (e < 0.f ? -1.f : 0.f);
  if (fn == REN_ERR_L1) { *dpred = sgn; return fabsf(e); }
  const float den = fmaxf(fabsf(tgt), 2.220446049250313e-16f);  // utils/modules.py:86,96
  *dpred = sgn / den;
  return fabsf(e) / den;
}

extern "C" __global__ void ren_diff_loss_kernel(const DiffLossArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float err = 0.f;
  int valid = 0;
  if (i < a.B) {
    const float l0 = logf(a.c0[i] + a.imin), l1 = logf(a.c1[i] + a.imin);
    a.log_i0[i] = l0;
    a.log_i1[i] = l1;
    valid = a.all_valid ? 1 : ((a.o0[i] > 0.f) || (a.o1[i] > 0.f));
    const double rate = (double)a.ev_diff[i] / ((double)a.end_ts[i] - a.start[i]);  // loss.py:38-41
    const float tgt = (float)(a.ts_diff[i] * rate);
    float dp;
    err = valid ? diff_err(a.error_fn, l1 - l0, tgt, &dp) : 0.f;
  }
  // block reduce
  __shared__ double s_err[8];
  __shared__ int s_cnt[8];
  double e = (double)err;
  int c = valid;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_err[w] = e; s_cnt[w] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double te = 0.0; int tc = 0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) { te += s_err[j]; tc += s_cnt[j]; }
    atomicAdd(a.acc, te);
    atomicAdd(reinterpret_cast<int*>(a.acc + 1), tc);
  }
}

extern "C" __global__ void ren_diff_loss_finish_kernel(const DiffLossArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = *reinterpret_cast<const int*>(a.acc + 1);
  if (i == 0) {
    a.loss[0] = (float)(a.acc[0] / (double)nv);   // empty mask -> NaN, as torch's mean of empty
    a.n_valid[0] = nv;
  }
  if (i >= a.B || a.g_c0 == nullptr) return;
  const float i0 = a.c0[i] + a.imin, i1 = a.c1[i] + a.imin;
  const int valid = a.all_valid ? 1 : ((a.o0[i] > 0.f) || (a.o1[i] > 0.f));
  float g = 0.f;
  if (valid) {
    const double rate = (double)a.ev_diff[i] / ((double)a.end_ts[i] - a.start[i]);
    const float tgt = (float)(a.ts_diff[i] * rate);
    float dp;
    diff_err(a.error_fn, a.log_i1[i] - a.log_i0[i], tgt, &dp);
    g = dp / (float)nv;
  }
  a.g_c1[i] = g / i1;
  a.g_c0[i] = -g / i0;
}

extern "C" int ren_event_diff_loss(int64_t B, int32_t error_fn, int32_t all_valid,
                                   float min_modeled_intensity, const float* c0, const float* c1,
                                   const float* o0, const float* o1, const float* ev_diff,
                                   const double* ts_diff, const double* start, const int64_t* end_ts,
                                   float* log_i0, float* log_i1, float* loss, int32_t* n_valid,
                                   float* g_c0, float* g_c1, void* scratch, void* stream) {
  if (B <= 0 || error_fn < 0 || error_fn > 2 || scratch == nullptr) {
    ren_set_error("ren_event_diff_loss: bad arguments");
    return REN_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DiffLossArgs a;
  a.B = B; a.error_fn = error_fn; a.all_valid = all_valid; a.imin = min_modeled_intensity;
  a.c0 = c0; a.c1 = c1; a.o0 = o0; a.o1 = o1; a.ev_diff = ev_diff; a.ts_diff = ts_diff;
  a.start = start; a.end_ts = (const long long*)end_ts; a.log_i0 = log_i0; a.log_i1 = log_i1;
  a.loss = loss; a.n_valid = n_valid; a.g_c0 = g_c0; a.g_c1 = g_c1; a.acc = (double*)scratch;
  cudaError_t e = cudaMemsetAsync(scratch, 0, 16, st);
  if (e != cudaSuccess) { ren_set_error("diff loss memset: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  const unsigned blocks = (unsigned)((B + 255) / 256);
  ren_prof_begin(REN_K_LOSS, st);
  ren_diff_loss_kernel<<<blocks, 256, 0, st>>>(a);
  ren_prof_end(REN_K_LOSS, st);
  ren_prof_begin(REN_K_LOSS, st);
  ren_diff_loss_finish_kernel<<<blocks, 256, 0, st>>>(a);
  ren_prof_end(REN_K_LOSS, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("diff loss launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

// ------------------------------------------------------------------ Adam
extern "C" __global__ void ren_adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                           float* __restrict__ m, float* __restrict__ v, float lr,
                                           float b1, float b2, float eps, float wd, float gscale,
                                           float bc1, float bc2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float grad = g[i] * gscale;
  const float param = p[i];
  if (wd != 0.f) grad = fmaf(wd, param, grad);
  const float mi = m[i] + (grad - m[i]) * (1.f - b1);       // torch: exp_avg.lerp_(grad, 1-beta1)
  const float vi = v[i] * b2 + (1.f - b2) * grad * grad;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = param - (lr / bc1) * (mi / denom);
}

extern "C" int ren_adam_step(int64_t n, float* params, const float* grads, float* exp_avg,
                             float* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float grad_scale, int64_t step, void* stream) {
  if (n <= 0 || step < 1) { ren_set_error("ren_adam_step: bad arguments"); return REN_E_INVALID; }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  ren_prof_begin(REN_K_ADAM, (cudaStream_t)stream);
  ren_adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, grad_scale,
      (float)bc1, (float)sqrt(bc2));
  ren_prof_end(REN_K_ADAM, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ren_adam_step launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

// Same update with the 1-based step count read from DEVICE memory, so that a captured CUDA graph of the whole
// training step advances Adam's bias correction on every replay (the host only increments *step on the stream).
extern "C" __global__ void ren_adam_dev_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                               float* __restrict__ m, float* __restrict__ v, float lr,
                                               float b1, float b2, float eps, float wd, float gscale,
                                               const long long* __restrict__ step) {
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double t = (double)*step;
    s_bc[0] = (float)(1.0 - pow((double)b1, t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float grad = g[i] * gscale;
  const float param = p[i];
  if (wd != 0.f) grad = fmaf(wd, param, grad);
  const float mi = m[i] + (grad - m[i]) * (1.f - b1);
  const float vi = v[i] * b2 + (1.f - b2) * grad * grad;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / s_bc[1] + eps;
  p[i] = param - (lr / s_bc[0]) * (mi / denom);
}

extern "C" int ren_adam_step_dev(int64_t n, float* params, const float* grads, float* exp_avg,
                                 float* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, float grad_scale, const int64_t* step, void* stream) {
  if (n <= 0 || step == nullptr) { ren_set_error("ren_adam_step_dev: bad arguments"); return REN_E_INVALID; }
  ren_prof_begin(REN_K_ADAM, (cudaStream_t)stream);
  ren_adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, grad_scale,
      (const long long*)step);
  ren_prof_end(REN_K_ADAM, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ren_adam_step_dev launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
