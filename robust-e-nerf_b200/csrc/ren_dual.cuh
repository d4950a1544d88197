// ren_dual.cuh -- forward-mode number types for the tangent-mode render (the log-intensity-
// gradient loss: d logI / d t of models/robust_e_nerf.py:394-398, utils/autograd.py:4-34).
//
//   Dual   (v, d)            first derivative along one direction (sample position -> contracted
//                            position, external/mlp.py:321-335, external/ngp.py:68-106)
//   HDual  (v, d1, d2, d12)  hyper-dual: two directions and the mixed second derivative (reverse pass
//                            of the tangent render with respect to the ray: d/dx [mu . J(x) xdot])
//   J2     (v, d, dd)        value, first and second derivative along ONE variable (the pose
//                            interpolation weight w of models/trajectories.py:63-65)
// Derivative components use plain fp32 arithmetic; the VALUE components of the product path are
// always taken from the bit-faithful routines in ren_common.cuh, never from these types.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// ------------------------------------------------------------------ Dual
struct Dual {
  float v, d;
  __host__ __device__ Dual() : v(0.f), d(0.f) {}
  __host__ __device__ Dual(float v_) : v(v_), d(0.f) {}
  __host__ __device__ Dual(float v_, float d_) : v(v_), d(d_) {}
};
__host__ __device__ inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__host__ __device__ inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__host__ __device__ inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__host__ __device__ inline Dual operator/(Dual a, Dual b) {
  const float r = 1.f / b.v;
  const float q = a.v * r;
  return Dual(q, (a.d - q * b.d) * r);
}
__host__ __device__ inline Dual t_sqrt(Dual a) {
  const float s = sqrtf(a.v);
  return Dual(s, s > 0.f ? a.d / (2.f * s) : 0.f);
}
__host__ __device__ inline Dual t_tanh(Dual a) {
  const float t = tanhf(a.v);
  return Dual(t, (1.f - t * t) * a.d);
}
__host__ __device__ inline float t_val(Dual a) { return a.v; }

// ------------------------------------------------------------------ HDual
struct HDual {
  float v, d1, d2, d12;
  __host__ __device__ HDual() : v(0.f), d1(0.f), d2(0.f), d12(0.f) {}
  __host__ __device__ HDual(float v_) : v(v_), d1(0.f), d2(0.f), d12(0.f) {}
  __host__ __device__ HDual(float v_, float a, float b, float c) : v(v_), d1(a), d2(b), d12(c) {}
};
__host__ __device__ inline HDual operator+(HDual a, HDual b) {
  return HDual(a.v + b.v, a.d1 + b.d1, a.d2 + b.d2, a.d12 + b.d12);
}
__host__ __device__ inline HDual operator-(HDual a, HDual b) {
  return HDual(a.v - b.v, a.d1 - b.d1, a.d2 - b.d2, a.d12 - b.d12);
}
__host__ __device__ inline HDual operator*(HDual a, HDual b) {
  return HDual(a.v * b.v, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2,
               a.d12 * b.v + a.d1 * b.d2 + a.d2 * b.d1 + a.v * b.d12);
}
// f(a) with f', f'' given at a.v
__host__ __device__ inline HDual hd_apply(HDual a, float f, float f1, float f2) {
  return HDual(f, f1 * a.d1, f1 * a.d2, f1 * a.d12 + f2 * a.d1 * a.d2);
}
__host__ __device__ inline HDual operator/(HDual a, HDual b) {
  const float r = 1.f / b.v;
  return a * hd_apply(b, r, -r * r, 2.f * r * r * r);
}
__host__ __device__ inline HDual t_sqrt(HDual a) {
  const float s = sqrtf(a.v);
  if (!(s > 0.f)) return HDual(s);
  return hd_apply(a, s, 0.5f / s, -0.25f / (s * a.v));
}
__host__ __device__ inline HDual t_tanh(HDual a) {
  const float t = tanhf(a.v);
  const float s = 1.f - t * t;
  return hd_apply(a, t, s, -2.f * t * s);
}
__host__ __device__ inline float t_val(HDual a) { return a.v; }

// ------------------------------------------------------------------ J2
struct J2 {
  float v, d, dd;
  __host__ __device__ J2() : v(0.f), d(0.f), dd(0.f) {}
  __host__ __device__ J2(float v_) : v(v_), d(0.f), dd(0.f) {}
  __host__ __device__ J2(float v_, float d_, float dd_) : v(v_), d(d_), dd(dd_) {}
};
__host__ __device__ inline J2 operator+(J2 a, J2 b) { return J2(a.v + b.v, a.d + b.d, a.dd + b.dd); }
__host__ __device__ inline J2 operator-(J2 a, J2 b) { return J2(a.v - b.v, a.d - b.d, a.dd - b.dd); }
__host__ __device__ inline J2 operator-(J2 a) { return J2(-a.v, -a.d, -a.dd); }
__host__ __device__ inline J2 operator*(J2 a, J2 b) {
  return J2(a.v * b.v, a.d * b.v + a.v * b.d, a.dd * b.v + 2.f * a.d * b.d + a.v * b.dd);
}
__host__ __device__ inline J2 j2_apply(J2 a, float f, float f1, float f2) {
  return J2(f, f1 * a.d, f2 * a.d * a.d + f1 * a.dd);
}
__host__ __device__ inline J2 operator/(J2 a, J2 b) {
  const float r = 1.f / b.v;
  return a * j2_apply(b, r, -r * r, 2.f * r * r * r);
}
__host__ __device__ inline J2 j2_sqrt(J2 a) {
  const float s = sqrtf(a.v);
  if (!(s > 0.f)) return J2(s);          // torch.norm backward at 0 -> 0
  return j2_apply(a, s, 0.5f / s, -0.25f / (s * a.v));
}
__host__ __device__ inline J2 j2_sin(J2 a) { return j2_apply(a, sinf(a.v), cosf(a.v), -sinf(a.v)); }
__host__ __device__ inline J2 j2_cos(J2 a) { return j2_apply(a, cosf(a.v), -sinf(a.v), -cosf(a.v)); }

// ------------------------------------------------------------------ contraction on any of the above
// VanillaNeRFRadianceField.contract_input_space (external/mlp.py:321-335) with
// contract_to_unisphere / contract_tanh (external/ngp.py:68-106): p -> z = 2 pi (x_hat - 1/2).
#include "../../include/ren_b200.h"
// p -> x_hat in the unit cube (what `NGPradianceField.query_density` feeds the grid encoding, external/ngp.py:219-228)
template <class T>
__host__ __device__ inline void ren_contract_unit_t(const RenFieldDesc& f, const T* p, T* x) {
  for (int a = 0; a < 3; ++a) x[a] = (p[a] - T(f.aabb[a])) / T(f.aabb[3 + a] - f.aabb[a]);
  if (f.contraction == REN_CONTRACT_SPHERE) {
    for (int a = 0; a < 3; ++a) x[a] = x[a] * T(2.f) - T(1.f);
    const T mag = t_sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    if (t_val(mag) > 1.f) {
      const T s = T(2.f) - T(1.f) / mag;
      for (int a = 0; a < 3; ++a) x[a] = s * (x[a] / mag);
    }
    for (int a = 0; a < 3; ++a) x[a] = x[a] * T(0.25f) + T(0.5f);
  } else if (f.contraction == REN_CONTRACT_TANH) {
    for (int a = 0; a < 3; ++a) x[a] = (t_tanh(x[a] - T(0.5f)) + T(1.f)) * T(0.5f);
  }
}
template <class T>
__host__ __device__ inline void ren_contract_t(const RenFieldDesc& f, const T* p, T* z) {
  T x[3];
  ren_contract_unit_t<T>(f, p, x);
  for (int a = 0; a < 3; ++a) z[a] = T(6.283185307179586f) * (x[a] - T(0.5f));
}
