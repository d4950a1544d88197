// ren_tc.cu -- tcgen05 engine of the field MLP (REN_ENGINE_TC_F16X3 / REN_ENGINE_TC_F16).
//
// One persistent CTA per SM.  Warp-specialised roles:
//   warp 0      weight producer: streams pre-packed fp16 (hi|lo) weight K-chunks from L2 into a
//               3-stage shared-memory ring with 1-D bulk TMA (cp.async.bulk + mbarrier tx)
//   warp 1      MMA issuer: one thread issues tcgen05.mma (kind::f16, M=128, N=256/128, K=16),
//               accumulator in TMEM (256 columns x 128 lanes fp32)
//   warps 2-9   prologue (march -> position -> contraction -> sin/cos encoding) and epilogue
//               (tcgen05.ld -> bias + softplus(beta=100) -> fp16 hi/lo split -> next layer's A
//               operand in shared memory; sigma / radiance heads as per-row dot products)
// A tile is 128 consecutive samples (= TMEM lanes).  Activations never leave the SM (except the
// optional fp32 stash the backward pass reads).
//
// fp32-class accuracy from fp16 tensor cores: x = hi + lo with hi = fp16(x), lo = fp16(x - hi);
// A*W ~= A_hi*W_hi + A_lo*W_hi + A_hi*W_lo accumulated in fp32 (3 MMAs per K step; the dropped
// lo*lo term is 2^-22 relative).  SURVEY.md section 7.3-1: single-pass fp16/bf16/TF32 misses the 1e-4 bar.
//
// Operand layout (both operands K-major, SWIZZLE_NONE): 8x8 fp16 core matrices of 128 contiguous
// bytes; element (row, k) of a [rows x K] operand at (k/8)*(rows*16) + row*16 + (k%8)*2.  A warp of
// epilogue threads (32 consecutive rows) therefore writes 512 contiguous bytes per k-core.
#include <cuda_fp16.h>

#include "ren_common.cuh"

#define TC_M 128
#define TC_THREADS 320
#define TC_STAGES 3
#define TC_MAX_STEPS 24

// shared-memory map (bytes)
#define SM_A_HI 0
#define SM_A_LO 65536
#define SM_P_HI 131072
#define SM_P_LO 147456
#define SM_D_HI 163840
#define SM_D_LO 172032
#define SM_RING 180224
#define SM_STAGE_BYTES 16384
#define SM_BAR (SM_RING + TC_STAGES * SM_STAGE_BYTES)   // 229376
#define SM_TMEM_PTR (SM_BAR + 64)
#define SM_SIG_PART (SM_BAR + 128)                      // [128] f32
#define SM_RGB_PART (SM_SIG_PART + 512)                 // [3][128] f32
#define SM_TOTAL (SM_RGB_PART + 1536)                   // 231552

struct TcStep {
  int seg_region[2];   // 0 main activations, 1 positional encoding, 2 direction encoding
  int seg_ksteps[2];   // K/16 per segment (0 = unused)
  int seg_K[2];        // true (unpadded) K per segment
  int N;               // 256 or 128
  int w_off, ldw, bias_off;
  int epi;             // 0 hidden activation, 1 linear (bottleneck), 2 condition layer (+ radiance head)
  int mat;             // stash matrix index
  int last_trunk;      // accumulate the sigma head in this epilogue
  long long pack_off;  // byte offset of this step's packed weights
  int n_chunks;        // total K steps
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
  const unsigned char* pack;
  float* stash;
  float* raw;
};

// ------------------------------------------------------------------ PTX wrappers
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle: LBO = byte stride between K cores,
// SBO = byte stride between 8-row groups (cute/arch/mma_sm100_desc.hpp SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128
__device__ __forceinline__ uint32_t umma_idesc(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ uint32_t region_hi(int region) {
  return region == 0 ? SM_A_HI : (region == 1 ? SM_P_HI : SM_D_HI);
}
__device__ __forceinline__ uint32_t region_lo(int region) {
  return region == 0 ? SM_A_LO : (region == 1 ? SM_P_LO : SM_D_LO);
}

// softplus(beta=100, threshold=20) on the MUFU pipe (ex2 + lg2); absolute error ~2e-9
__device__ __forceinline__ float softplus100_fast(float x) {
  const float z = x * 100.f;
  const float e = __expf(z);
  const float y = __logf(1.f + e) * 0.01f;
  return z > 20.f ? x : y;
}

// 8 activations -> 16 B of fp16 hi and 16 B of fp16 lo
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

// Writes an encoding of NF (<= KP) features into the blocked operand region (hi/lo) and the stash.
template <int DEG, int KP>
__device__ __forceinline__ void encode_to_operand(const float* x, bool valid, unsigned char* smem, uint32_t hi_off,
                                                  uint32_t lo_off, int m, float* stash_mat, long long sidx) {
  float f[KP];
  if (valid) {
    ren_encode(x, DEG, KP, [&](int j, float v) { f[j] = v; });
  } else {
#pragma unroll
    for (int j = 0; j < KP; ++j) f[j] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < KP / 8; ++c) {
    uint4 hi, lo;
    split8(f + c * 8, &hi, &lo);
    *reinterpret_cast<uint4*>(smem + hi_off + c * (TC_M * 16) + m * 16) = hi;
    *reinterpret_cast<uint4*>(smem + lo_off + c * (TC_M * 16) + m * 16) = lo;
  }
  if (stash_mat != nullptr && valid) {
#pragma unroll
    for (int j = 0; j < KP; j += 4)
      *reinterpret_cast<float4*>(stash_mat + ren_blk(sidx, j, KP)) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
  }
}

template <bool STASH>
__global__ void __launch_bounds__(TC_THREADS, 1) ren_tc_field_fwd(const TcArgs a) {
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

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_aready, 256);
    mbar_init(bar_accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sbase + SM_TMEM_PTR) : "memory");
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
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, ev = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int g = 0; g < a.n_steps; ++g, ++ev) {
          const TcStep& st = a.steps[g];
          mbar_wait(bar_aready, ev & 1u);
          tc_fence_after();
          const uint32_t idesc = umma_idesc(st.N);
          const uint32_t b_lbo = (uint32_t)st.N * 16u;
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
              tc_mma_f16(tmem_base, da_hi, db_hi, idesc, c > 0 ? 1u : 0u);
              if (a.split3) {
                const uint64_t da_lo = umma_desc(a_lo + ks * 4096, TC_M * 16, 128);
                const uint64_t db_lo = umma_desc(wb + st.N * 32, b_lbo, 128);
                tc_mma_f16(tmem_base, da_lo, db_hi, idesc, 1u);
                tc_mma_f16(tmem_base, da_hi, db_lo, idesc, 1u);
              }
              tc_commit(bar_empty + 8 * s);   // frees the ring slot when these MMAs have read it
            }
          }
          tc_commit(bar_accfull);             // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ===================== prologue + epilogue warps =====================
    const int ew = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                // column half
    const int m = q * 32 + lane;             // row of the tile == TMEM lane
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* sig_part = reinterpret_cast<float*>(smem + SM_SIG_PART);
    float* rgb_part = reinterpret_cast<float*>(smem + SM_RGB_PART);
    const int C1 = 1 + p.C;
    uint32_t ev = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long sidx = tile * TC_M + m;
      const bool valid = sidx < S;
      // ---- prologue: half 0 -> positional encoding, half 1 -> direction encoding
      bool selector = false;
      {
        float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f};
        float t0 = 0.f, t1 = 0.f, t_max = 0.f;
        if (valid) {
          const long long r = sidx / a.K;
          const int k = (int)(sidx - r * a.K);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            o[c] = a.rays_o[r * 3 + c];
            d[c] = a.rays_d[r * 3 + c];
          }
          if (half == 0) {
            float t_min;
            ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
            ren_march_to(p.r, t_min, k, &t0, &t1);
          }
        }
        if (half == 0) {
          float z[3] = {0.f, 0.f, 0.f};
          if (valid) {
            selector = ren_field_input(p.f, o, d, t0, t1, z);
            if (!ren_alive(t0, t1, t_max)) selector = false;
          }
          encode_to_operand<10, 64>(z, valid, smem, SM_P_HI, SM_P_LO, m,
                                    STASH ? a.stash + ren_stash_encp(p, S) : nullptr, sidx);
        } else {
          float dz[3] = {__fmul_rn(d[0], 3.141592653589793f), __fmul_rn(d[1], 3.141592653589793f),
                         __fmul_rn(d[2], 3.141592653589793f)};
          encode_to_operand<4, 32>(dz, valid, smem, SM_D_HI, SM_D_LO, m,
                                   STASH ? a.stash + ren_stash_encd(p, S) : nullptr, sidx);
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_aready);
      float sig_acc = 0.f;
      float rgb_acc[3] = {0.f, 0.f, 0.f};
      // ---- one epilogue per GEMM step
      for (int g = 0; g < a.n_steps; ++g, ++ev) {
        const TcStep& st = a.steps[g];
        mbar_wait(bar_accfull, ev & 1u);
        tc_fence_after();
        const int ncol = st.N >> 1;                  // columns this warp handles
        const int col0 = half * ncol;
        float* stash_row = (STASH && valid) ? a.stash + ren_stash_mat(p, S, st.mat) : nullptr;
        for (int cb = 0; cb < ncol; cb += 32) {
          uint32_t r[32];
          tc_ld32(tmem_lane + (uint32_t)(col0 + cb), r);
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float y[8];
            const int n0 = col0 + cb + j8 * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = __uint_as_float(r[j8 * 8 + j]) + __ldg(a.params + st.bias_off + n0 + j);
              if (st.epi == 1) {
                y[j] = x;
              } else {
                y[j] = (p.f.hidden_activation == REN_HIDDEN_SOFTPLUS100) ? softplus100_fast(x) : fmaxf(x, 0.f);
              }
            }
            if (st.last_trunk) {
#pragma unroll
              for (int j = 0; j < 8; ++j) sig_acc = fmaf(y[j], __ldg(a.params + p.sig_w + n0 + j), sig_acc);
            }
            if (st.epi == 2) {
              for (int c = 0; c < p.C; ++c) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  rgb_acc[c] = fmaf(y[j], __ldg(a.params + p.rgb_w + c * p.WC + n0 + j), rgb_acc[c]);
              }
            } else {
              uint4 hi, lo;
              split8(y, &hi, &lo);
              const uint32_t off = (uint32_t)(n0 >> 3) * (TC_M * 16) + (uint32_t)m * 16;
              *reinterpret_cast<uint4*>(smem + SM_A_HI + off) = hi;
              *reinterpret_cast<uint4*>(smem + SM_A_LO + off) = lo;
            }
            if (STASH && stash_row != nullptr) {
              float* dst = stash_row + ren_blk(sidx, n0, st.N);   // 32 contiguous bytes, lanes contiguous
              *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
          }
        }
        tc_fence_before();
        if (g + 1 < a.n_steps) {
          fence_proxy_async();
          mbar_arrive(bar_aready);
        }
      }
      // ---- heads: combine the two column halves, write raw sigma / radiance
      if (half == 1) {
        sig_part[m] = sig_acc;
        for (int c = 0; c < p.C; ++c) rgb_part[c * TC_M + m] = rgb_acc[c];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && valid) {
        const float sraw = sig_acc + sig_part[m] + __ldg(a.params + p.sig_b);
        a.raw[sidx * C1] = selector ? sraw : -INFINITY;
        for (int c = 0; c < p.C; ++c)
          a.raw[sidx * C1 + 1 + c] = rgb_acc[c] + rgb_part[c * TC_M + m] + __ldg(a.params + p.rgb_b + c);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ weight packing
struct PackArgs {
  TcStep steps[TC_MAX_STEPS];
  int n_steps;
  const float* params;
  unsigned char* pack;
};

extern "C" __global__ void ren_tc_pack_weights(const PackArgs a) {
  const int g = blockIdx.y;
  if (g >= a.n_steps) return;
  const TcStep& st = a.steps[g];
  const int total = st.n_chunks * 16 * st.N;    // elements (k-major sweep)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e % st.N;
    const int kc = e / st.N;                   // concatenated (padded) k index
    const int c = kc >> 4, kk = kc & 15;
    const int k0pad = st.seg_ksteps[0] * 16;
    float w = 0.f;
    if (kc < k0pad) {
      if (kc < st.seg_K[0]) w = a.params[st.w_off + (long long)n * st.ldw + kc];
    } else {
      const int k1 = kc - k0pad;
      if (k1 < st.seg_K[1]) w = a.params[st.w_off + (long long)n * st.ldw + st.seg_K[0] + k1];
    }
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    unsigned char* chunk = a.pack + st.pack_off + (size_t)c * st.N * 64;
    const size_t off = (size_t)(kk >> 3) * (st.N * 16) + (size_t)n * 16 + (kk & 7) * 2;
    *reinterpret_cast<__half*>(chunk + off) = hi;
    *reinterpret_cast<__half*>(chunk + (size_t)st.N * 32 + off) = lo;
  }
}

// ------------------------------------------------------------------ host side
static int build_steps(const RenPlan& p, TcStep* steps, long long* pack_bytes) {
  int n = 0;
  long long off = 0;
  auto add = [&](const RenLayer& L, int epi, int last_trunk) {
    TcStep& s = steps[n++];
    s.seg_region[0] = L.src0 == 1 ? 1 : 0;
    s.seg_K[0] = L.K0;
    s.seg_ksteps[0] = (L.K0 + 15) / 16;
    s.seg_region[1] = L.src1;     // 0 none / 1 encP / 2 encD
    s.seg_K[1] = L.K1;
    s.seg_ksteps[1] = L.K1 > 0 ? (L.K1 + 15) / 16 : 0;
    s.N = L.N;
    s.w_off = L.w_off;
    s.ldw = L.K0 + L.K1;
    s.bias_off = L.b_off;
    s.epi = epi;
    s.mat = L.mat;
    s.last_trunk = last_trunk;
    s.n_chunks = s.seg_ksteps[0] + s.seg_ksteps[1];
    s.pack_off = off;
    off += (long long)s.n_chunks * s.N * 64;
  };
  for (int i = 0; i < p.D; ++i) add(p.trunk[i], 0, i == p.D - 1);
  add(p.bott, 1, 0);
  add(p.cond[0], 2, 0);
  *pack_bytes = off;
  return n;
}

int ren_tc_supported(const RenPlan& p) {
  return p.W == 256 && p.WC == 128 && p.CD == 1 && p.hid_K1 == 0 && p.f.pos_deg == 10 && p.f.dir_deg == 4 &&
         p.D + 2 <= TC_MAX_STEPS && p.C <= 3;
}

size_t ren_tc_workspace_bytes(const RenPlan& p, long long S) {
  (void)S;
  if (!ren_tc_supported(p)) return 0;
  TcStep steps[TC_MAX_STEPS];
  long long bytes = 0;
  build_steps(p, steps, &bytes);
  return (size_t)bytes + 1024;
}

int ren_tc_launch_fwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                      const float* jitter, const float* params, float* stash, float* raw, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  TcArgs a;
  static_assert(sizeof(TcArgs) <= 4000, "kernel parameter block too large");
  long long pack_bytes = 0;
  a.p = p;
  a.n_steps = build_steps(p, a.steps, &pack_bytes);
  unsigned char* pack = (unsigned char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  if (ws == nullptr || (size_t)(pack - (unsigned char*)ws) + (size_t)pack_bytes > ws_bytes) {
    ren_set_error("workspace too small for packed weights (%lld B needed)", pack_bytes);
    return REN_E_WORKSPACE;
  }
  PackArgs pa;
  for (int i = 0; i < a.n_steps; ++i) pa.steps[i] = a.steps[i];
  pa.n_steps = a.n_steps;
  pa.params = params;
  pa.pack = pack;
  ren_prof_begin(REN_K_PACK, st);
  ren_tc_pack_weights<<<dim3(32, a.n_steps), 256, 0, st>>>(pa);
  ren_prof_end(REN_K_PACK, st);
  a.split3 = p.r.engine == REN_ENGINE_TC_F16 ? 0 : 1;
  a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter; a.params = params;
  a.pack = pack; a.stash = stash; a.raw = raw;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (n_rays * K + TC_M - 1) / TC_M;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (grid < 1) return REN_OK;
  cudaError_t e;
  if (stash) {
    e = cudaFuncSetAttribute(ren_tc_field_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e != cudaSuccess) { ren_set_error("tc fwd smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
    ren_prof_begin(REN_K_FIELD_FWD, st);
    ren_tc_field_fwd<true><<<grid, TC_THREADS, SM_TOTAL, st>>>(a);
    ren_prof_end(REN_K_FIELD_FWD, st);
  } else {
    e = cudaFuncSetAttribute(ren_tc_field_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e != cudaSuccess) { ren_set_error("tc fwd smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
    ren_prof_begin(REN_K_FIELD_FWD, st);
    ren_tc_field_fwd<false><<<grid, TC_THREADS, SM_TOTAL, st>>>(a);
    ren_prof_end(REN_K_FIELD_FWD, st);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc fwd launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
