// ren_tc.cuh -- PTX wrappers (mbarrier, bulk TMA, tcgen05) and the fp16 stash layout shared by the
// tcgen05 kernels (ren_tc.cu: forward / dgrad chain; ren_tc_wgrad.cu: weight gradients).
#pragma once
#include <cuda_fp16.h>

#include "ren_common.cuh"

#define TC_M 128   // samples per tile == TMEM lanes

// ------------------------------------------------------------------ fp16 stash (tcgen05 engine)
// Every activation / delta matrix [S x N] is kept as fp16 in the BLOCKED layout of ren_blk():
// per 128-sample tile and 8-feature group one 2 KB block [128 samples][8 features].  The same
// bytes are (a) a K-major UMMA operand for the chain kernels (K = features) and (b) an MN-major
// UMMA operand for the weight-gradient kernel (K = samples), so tiles move HBM <-> SMEM with
// plain bulk copies.  Layout in bytes (Sp = S padded to 128):
//   mats      : mat m at 2 * mat_off[m] * Sp,   size 2 * width * Sp
//   encP, encD: 2 * act_floats * Sp, then + 2 * 64 * Sp
//   raw (f32) : after encD, [Sp][1 + C]
__host__ __device__ inline size_t tcs_mat(const RenPlan& p, long long S, int m) {
  return (size_t)2 * p.mat_off[m] * ren_pad128(S);
}
__host__ __device__ inline size_t tcs_encp(const RenPlan& p, long long S) {
  return (size_t)2 * p.act_floats * ren_pad128(S);
}
__host__ __device__ inline size_t tcs_encd(const RenPlan& p, long long S) {
  return (size_t)2 * (p.act_floats + REN_PE_MAX) * ren_pad128(S);
}
__host__ __device__ inline size_t tcs_raw(const RenPlan& p, long long S) {
  return (size_t)2 * (p.act_floats + REN_PE_MAX + REN_DE_MAX) * ren_pad128(S);
}
// per-tile "live" bytes (occupancy-grid marching): a tile none of whose 128 marching steps emits a sample is
// skipped by every kernel of the step
__host__ __device__ inline size_t tcs_live(const RenPlan& p, long long S) {
  return tcs_raw(p, S) + (size_t)4 * (1 + p.C) * ren_pad128(S);
}
__host__ __device__ inline size_t tcs_bytes(const RenPlan& p, long long S) {
  return tcs_live(p, S) + (((size_t)(ren_pad128(S) >> 7) + 255) & ~(size_t)255);
}
// backward workspace: delta mats (fp16 blocked) then graw (f32 [Sp][1+C])
__host__ __device__ inline size_t tcd_graw(const RenPlan& p, long long S) {
  return (size_t)2 * p.act_floats * ren_pad128(S);
}
// max |graw| (float bits) written by the compositing reverse pass: the deltas are scaled by a power
// of two so that they sit in fp16's range (loss scaling); the weight-gradient epilogue unscales.
__host__ __device__ inline size_t tcd_gmax(const RenPlan& p, long long S) {
  return tcd_graw(p, S) + (size_t)4 * (1 + p.C) * ren_pad128(S);
}
__host__ __device__ inline size_t tcd_bytes(const RenPlan& p, long long S) { return tcd_gmax(p, S) + 256; }

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The suspend-time hint lets the hardware park the thread until the phase completes (or the hint expires) instead
// of returning after the short default window: without it the wait loops of the chain kernel re-issued
// try_wait + branch every ~125 cycles (a quarter of all executed instructions, ncu r02).
#ifndef REN_MBAR_HINT_NS
#define REN_MBAR_HINT_NS 20000
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if REN_MBAR_HINT_NS == 0
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"((uint32_t)REN_MBAR_HINT_NS) : "memory");
#endif
}
// the same wait polling without a suspend-time hint (single-thread roles on a latency-critical hand-off)
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// REN_ELECT: the single-thread roles (TMA producers, MMA issuers) run on a lane chosen with elect.sync (1) or on
// lane 0 (0, the A/B baseline)
#ifndef REN_ELECT
#define REN_ELECT 1
#endif
// One lane of the (converged) warp; unlike `lane == 0`, ptxas knows the guarded region runs on exactly one thread, so
// the uniform-register operands of tcgen05.mma / commit need no per-lane serialisation loop around every instruction
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0u;
}
#define REN_ONE_LANE() (REN_ELECT ? elect_one() : lane == 0)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane (issue only; see tc_ld_wait)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor, version 1, no
// swizzle).  K-major operand:  LBO = byte stride between the two K cores of one MMA,
//                               SBO = byte stride between 8-row (M/N) groups.
//                 MN-major operand: LBO = byte stride between 8-deep K blocks,
//                               SBO = byte stride between 8-wide M/N groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32, A = B = f16, M = 128; major bits select MN-major
__device__ __forceinline__ uint32_t umma_idesc(int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

// 8 values -> 16 B of fp16 hi and 16 B of fp16 lo (x = hi + lo to ~2^-22)
__device__ __forceinline__ void split8(const float* y, uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(y[2 * i] - hf.x, y[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// 8 values -> 16 B of fp16 (round to nearest)
__device__ __forceinline__ void pack8(const float* y, uint4* hi) {
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
}
// power-of-two scale that brings max |graw| to ~16 (fp16 headroom 2^12 above, 2^28 below)
__device__ __forceinline__ float delta_scale(const unsigned int* gmax) {
  const float mx = __uint_as_float(*gmax);
  if (!(mx > 0.f)) return 1.f;
  int e = (int)floorf(log2f(16.f / mx));
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return exp2f((float)e);
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
#endif  // __CUDACC__
