// ren_tc.cu -- tcgen05 "chain" kernel of the field MLP: forward and dgrad (REN_ENGINE_TC_*).
//
// One persistent CTA per SM, one 128-sample tile in flight, ALL activations resident in tensor
// memory: the 512 TMEM columns are four 128-column regions used in ping-pong pairs.  A layer
// reads its A operand (fp16 hi/lo, packed two features per 32-bit column) from one pair and
// accumulates fp32 into the other; the epilogue converts each accumulator region IN PLACE into the
// next layer's A operand (tcgen05.ld -> bias/activation -> hi/lo split -> tcgen05.st), so no
// activation ever touches shared memory and the two 128-column halves of a layer pipeline against
// the tensor pipe: while the epilogue warps convert N-half 0, the MMAs of N-half 1 run, and the
// next layer's K-half-0 MMAs start as soon as half 0 is converted.
//   warp 0       weight producer: pre-packed fp16 (hi|lo) weight chunks, L2 -> 3 x 32 KB shared-memory
//                ring with 1-D bulk TMA (cp.async.bulk + mbarrier tx bytes)
//   warp 1       MMA issuer: one thread issues tcgen05.mma kind::f16, M=128, K=16; A from TMEM
//                (activations) or shared memory (positional / direction encodings).  The first K-half
//                of a layer covers both N-halves (N=256 operands, issued as interleaved N=128 pairs), the
//                last K-half runs per N-half (N=128) so that accumulator half 0 completes -- and its
//                epilogue starts -- a quarter layer early.  The thread is chosen with elect.sync
//                (REN_ONE_LANE): under `lane == 0` ptxas wraps every tcgen05 / bulk-copy instruction in a
//                per-lane serialisation loop and the single issuing thread, not the tensor pipe, sets the
//                pace
//   warps 2-17   epilogues.  Lane quadrant = warp % 4 (TMEM access rule); the 4 warps of a quadrant split
//                the 128 columns of a half 32 each and write their 32 packed columns back over their own.
//                dgrad: the activation-stash slices they need arrive by cp.async through the (idle)
//                encoding buffers, issued one half-epilogue ahead
//   warps 18-19  forward only: prologue (march, contraction, 84 sines -> encodings of the NEXT tile in the
//                free one of two shared-memory buffers)
// fp32-class accuracy from fp16 tensor cores: x = hi + lo; A*W ~= A_hi*W_hi + A_lo*W_hi + A_hi*W_lo
// (3 MMAs per K step).  Single-pass fp16/bf16/TF32 misses the 1e-4 bar.
//
// Hidden softplus(beta=100) runs in the log2 domain: the packed weights carry the factor
// 100*log2(e), the accumulator is u = 100 log2e (W h + b), the stored activation is
// g = max(u,0) + lg2(1 + 2^-|u|) = 100 log2e * softplus_100(pre), and the next layer's weights absorb
// the inverse factor -- 4 FP32 + 2 MUFU instructions per element.  The fp16 stash keeps g for the
// hidden layers (the weight-gradient kernels unscale).
//
// The sigma / radiance heads (N = 1..3) are per-thread dot products in the epilogue of the layer that
// feeds them; in dgrad the head gradients re-enter as rank-1 updates in the epilogue.
#include <stdlib.h>
#include <string.h>

#include "ren_tc.cuh"
#include "ren_dual.cuh"

#define TC_EPI_WARPS 16
#define TC_PRO_WARPS 2                         // forward prologue warps: two sample rows per thread (640 threads keep
                                               // 96 registers for the epilogue warps)
#define TC_THREADS_FWD (64 + 32 * TC_EPI_WARPS + 32 * TC_PRO_WARPS)   // 640
#define TC_THREADS_BWD (64 + 32 * TC_EPI_WARPS)                       // 576: dgrad has no prologue warps
#define TC_STAGES 3            // ring stages of up to TC_STAGE_KS K steps (32 KB)
#define TC_STAGE_KS 4
#define TC_MAX_STEPS 20
#ifndef FLAT_AHEAD
#define FLAT_AHEAD 1            // softplus groups (of 8 columns) computed ahead of the split / store of a group
#endif
// Straight-line epilogues of the common layer kinds (hidden softplus layer of the value pass; plain dgrad through a
// hidden softplus layer): the per-group runtime tests of the generic loop (layer kind, operand precision, stash /
// operand destinations, row validity) are decided once per half-epilogue.  0 = generic loop only (A/B builds).
#ifndef REN_FLAT_FWD
#define REN_FLAT_FWD 1
#endif
#ifndef REN_FLAT_BWD
#define REN_FLAT_BWD 1
#endif
// Biases of the hidden softplus layers of the value pass as one more K step of the layer's GEMM instead of one
// FADD (+ a quarter of a shared-memory load) per element in the epilogue, which is the limiter: A = a resident
// [128 x 16] tile whose first column is 1, B = [N x 16] with the scaled bias (fp16 hi | lo) in its first column, two
// MMAs (1 * b_hi + 1 * b_lo) that also initialise the accumulator.  Needs REN_FLAT_FWD.  0 = bias in the epilogue.
#ifndef REN_BIAS_MMA
#define REN_BIAS_MMA 0
#endif
#if REN_BIAS_MMA && !REN_FLAT_FWD
#error "REN_BIAS_MMA needs REN_FLAT_FWD"
#endif
// Every N = 256 MMA of a layer's first K-half is issued as two N = 128 MMAs on the two accumulator halves,
// interleaved, so that consecutive MMAs never accumulate into the same columns (same arithmetic, bit-identical
// results; measured -1 % forward / -3 % dgrad once the issue loop runs on an elect.sync lane, +10 % before that,
// when every extra tcgen05.mma cost a per-lane serialisation loop).  0 = one N = 256 MMA.
#ifndef REN_WIDE_SPLIT
#define REN_WIDE_SPLIT 1
#endif
// Issue experiment (A/B builds only, off): operand set-up of an op hoisted above the issuer's barrier waits, which poll
// without the suspend hint.  Same MMAs in the same order: bit-identical results.
#ifndef REN_ISSUE_FAST
#define REN_ISSUE_FAST 0
#endif

// shared-memory map (bytes)
#define SM_P_HI 0          // positional encoding [128 x 64] hi (K-major core matrices)
#define SM_P_LO 16384
#define SM_D_HI 32768      // direction encoding [128 x 32]
#define SM_D_LO 40960
#define SM_ENC_BYTES 49152 // one encoding buffer; two of them (the prologue warps run one tile ahead)
// dgrad has no encodings: the same 96 KB stage the activation stash of the NEXT half-epilogue (cp.async, one
// 16-byte slot per epilogue thread and 8-feature group): two activation buffers + one for the tangent / E matrix
#define SM_DSTAGE_ACT 0        // [2][4][512] x 16 B
#define SM_DSTAGE_AUX 65536    // [4][512] x 16 B
#define SM_RING (2 * SM_ENC_BYTES)
#define SM_CHUNK_BYTES 8192          // one [128 x 16] hi|lo weight chunk (one K step of one N-half)
#define SM_STAGE_BYTES (TC_STAGE_KS * SM_CHUNK_BYTES)
#define SM_VEC (SM_RING + TC_STAGES * SM_STAGE_BYTES)   // 196608: float vectors (biases, head weights)
#define VEC_BIAS 0                                      // [TC_MAX_STEPS][256] -> only n_steps used
#define VEC_WSIG (12 * 256)
#define VEC_WO (VEC_WSIG + 256)                         // [3][128]
#define VEC_HB (VEC_WO + 384)                           // sigma bias, rgb biases
#define VEC_FLOATS (VEC_HB + 4)                         // 3716
#define SM_PART (SM_VEC + 16384)                        // head partial sums: float4 [4][128]
#define SM_PROG (SM_PART + 8192)                        // stage program: TC_MAX_OPS x 32 B
#define TC_MAX_OPS 128
#define SM_SEL (SM_PROG + TC_MAX_OPS * 32)              // selector flags of the two buffered tiles: float [2][128]
#define SM_BAR (SM_SEL + 1024)
#define SM_TMEM_PTR (SM_BAR + 320)
#define SM_ONES (SM_BAR + 512)                          // [128 x 16] fp16 K-major operand, column 0 = 1 (bias K step)
#define SM_TOTAL (SM_ONES + 4096)
#define OP_ONES 0x400u                                  // StageOp flag: A = the ones tile, B = the bias chunk

#define LOG2E_100 144.26950408889634f      // 100 * log2(e)
#define LN2_OVER_100 0.0069314718055994531f

enum { EPI_F_HIDDEN = 0, EPI_F_LINEAR = 1, EPI_F_LAST = 2, EPI_D_ACT = 3, EPI_D_LINEAR = 4 };

struct TcStep {
  int N;          // 256 (two halves) or 128 (one)
  int acc_col;    // accumulator base column (0 or 256)
  int a_col;      // TMEM A base column
  int a_k;        // K taken from TMEM: 0, 128 or 256
  int seg;        // shared-memory K segment: 0 none, 1 positional encoding (4 K steps), 2 direction (2)
  int epi;        // EPI_*
  int out_mat;    // stash (fwd) / delta (dgrad) matrix written by the epilogue, -1 none
  int act_mat;    // dgrad: stash matrix of the activation whose derivative multiplies, -1 none
  int write_a;    // result becomes the next step's A operand
  int bias_idx;   // row of the shared bias table, -1 none
  int rank1;      // dgrad: += g_sigma * w_sigma[n] before the activation derivative
  int sig_head;   // fwd: this epilogue also accumulates the sigma head
  float acc_scale;   // EPI_F_LINEAR: out = acc_scale * acc + bias (keeps the packed weights out of fp16 subnormals)
  int bias_mma;      // fwd: the bias enters as the first K step of the GEMM (ones tile x bias chunk), not in the epilogue
  long long pack_off;   // byte offset of this step's weight chunks (in MMA issue order)
};

// One ring stage of the tile program (host-built, identical for every tile): what the producer
// copies and what the MMA issuer does with it.  32 bytes, read as two 128-bit shared-memory loads.
struct StageOp {
  uint32_t pack_off;    // byte offset of the stage's weight chunks in the pack buffer
  uint32_t bytes;       // stage bytes
  uint32_t d_col;       // accumulator TMEM column
  uint32_t a_addr;      // A of the first K step: TMEM column (a_smem = 0) or descriptor low word minus base (a_smem = 1)
  uint32_t a_lo_delta;  // from the hi to the lo part of A (TMEM columns / descriptor units)
  uint32_t flags;       // [0:2] K steps, [3] wide (N=256), [4] a_smem, [5] first MMA overwrites, [6:7] wait: 1 A-half 0,
                        // 2 A-half 1, 3 encodings; [8:9] commit accumulator half (1, 2) after the stage; [10] bias K step
  uint32_t pad0, pad1;
};

struct TcArgs {
  RenPlan p;
  TcStep steps[TC_MAX_STEPS];
  int n_steps;
  int split3;                    // operand precision: 0 single fp16 pass; 1 A and W split hi/lo (3 MMAs); 2 TMEM-resident
                                 // activations single fp16, weights split (2 MMAs; the shared-memory encodings keep 3)
  int softplus;                  // hidden activation is softplus(beta=100) (log2-domain epilogue)
  int dbg;                       // timing experiments only (env REN_TC_DBG): 1 no epilogue math, 2 no MMAs, 4 no sines
  long long n_rays;
  int K;
  const float* rays_o;
  const float* rays_d;
  const float* jitter;
  const unsigned char* pack;     // packed weights
  const float* vec;              // packed float vectors (VEC_* layout)
  const StageOp* prog;           // stage program (global copy)
  int n_ops;
  unsigned char* stash;          // fp16 stash (tcs_* layout); fwd: written when STASH, dgrad: read
  unsigned char* delta;          // dgrad: fp16 delta mats (written)
  float* raw;                    // fwd: [S][1+C] raw sigma / radiance
  const float* graw;             // dgrad: [S][1+C] d loss / d raw
  const unsigned int* gmax;      // dgrad: max |graw| (float bits) -> power-of-two delta scale
  // tangent render (d/dw rows): three more passes of the same chain
  int tan;                       // fwd: tangent pass -- A = d enc/dw, epilogue = act'(value stash) * acc, no biases
  int pass;                      // dgrad: 0 plain, 1 tangent adjoint (also writes E = abar_dot * act'' * pre_dot),
                                 //        2 value adjoint (adds E)
  const float* rays_o_dot;
  const float* rays_d_dot;
  const unsigned char* stash_val;   // fwd tangent pass: the value pass' stash (read)
  const unsigned char* stash_dot;   // dgrad pass 1: the tangent pass' stash (read)
  unsigned char* extra;             // dgrad: E matrices (fp16 blocked; written by pass 1, read by pass 2)
  const unsigned char* tile_live;   // per-tile flags (occupancy-grid marching) or NULL: dead tiles are skipped by all roles
  const int* sample_idx;            // packed samples: dense index r*K + k of every emitted marching step, or NULL
  long long n_samples;              // number of packed samples (when sample_idx != NULL)
};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2-domain softplus: u = 100 log2e x  ->  100 log2e * softplus_100(x)
__device__ __forceinline__ float softplus_log2(float u) { return fmaxf(u, 0.f) + lg2_ftz(1.f + ex2_ftz(-fabsf(u))); }
// d softplus / dx from g = 100 log2e * softplus_100(x): sigmoid(100 x) = 1 - 2^-g
__device__ __forceinline__ float softplus_grad_log2(float g) { return 1.f - ex2_ftz(-g); }

// dgrad through one hidden activation, from the stashed log2-domain output g (and, for the tangent render,
// the stashed tangent gdot):  pass 0/1: y = v act'(g);  pass 1 also E = v act''(pre) pre_dot = v ln2 2^-g gdot;
// pass 2: y = v act'(g) + E (x = E).
__device__ __forceinline__ float dgrad_act(int softplus, int pass, float v, float g, float x, float* e) {
  float s, e2 = 0.f;
  if (softplus) { e2 = ex2_ftz(-g); s = 1.f - e2; } else { s = g > 0.f ? 1.f : 0.f; }
  *e = pass == 1 ? v * (0.6931471805599453f * e2 * x) : 0.f;
  return pass == 2 ? fmaf(v, s, x) : v * s;
}

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_st4(uint32_t taddr, const uint4& v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16-byte asynchronous global -> shared copy (no register staging), group commit / wait
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One out-of-line accurate sine keeps the kernel small (84 inlined sinf bodies were ~100 KB of
// SASS and stalled the epilogue warps on instruction fetch).
__device__ __noinline__ float ren_sinf(float x) { return sinf(x); }

// feature j of SinusoidalEncoder(x) with max degree deg (external/mlp.py:227-243)
__device__ __forceinline__ float enc_feature(const float* x, int deg, int j) {
  if (j < 3) return j == 0 ? x[0] : (j == 1 ? x[1] : x[2]);
  if (j >= 3 + 6 * deg) return 0.f;
  const int jj = j - 3;
  const bool shifted = jj >= 3 * deg;
  const int q = shifted ? jj - 3 * deg : jj;
  const int i = q / 3, a = q - 3 * i;
  const float xa = a == 0 ? x[0] : (a == 1 ? x[1] : x[2]);
  const float xb = __fmul_rn(xa, __int_as_float((127 + i) << 23));     // exact 2^i
  return ren_sinf(shifted ? __fadd_rn(xb, 1.5707963267948966f) : xb);
}
__device__ __noinline__ float ren_cosf(float x) { return cosf(x); }
// d/dw of the same feature given xdot = dx/dw
__device__ __forceinline__ float enc_feature_tan(const float* x, const float* xd, int deg, int j) {
  if (j < 3) return j == 0 ? xd[0] : (j == 1 ? xd[1] : xd[2]);
  if (j >= 3 + 6 * deg) return 0.f;
  const int jj = j - 3;
  const bool shifted = jj >= 3 * deg;
  const int q = shifted ? jj - 3 * deg : jj;
  const int i = q / 3, a = q - 3 * i;
  const float xa = a == 0 ? x[0] : (a == 1 ? x[1] : x[2]);
  const float xda = a == 0 ? xd[0] : (a == 1 ? xd[1] : xd[2]);
  const float sc = __int_as_float((127 + i) << 23);
  const float xb = __fmul_rn(xa, sc);
  return ren_cosf(shifted ? __fadd_rn(xb, 1.5707963267948966f) : xb) * sc * xda;
}
// encode k-cores [c0, c1) of an encoding into an operand region (+ fp16 stash)
__device__ __forceinline__ void encode_cores(const float* x, const float* xd, int deg, int kp, int c0, int c1,
                                             bool valid, unsigned char* smem, uint32_t hi_off, uint32_t lo_off, int m,
                                             unsigned char* stash_mat, long long sidx) {
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      f[j] = !valid ? 0.f : (xd ? enc_feature_tan(x, xd, deg, c * 8 + j) : enc_feature(x, deg, c * 8 + j));
    uint4 hi, lo;
    split8(f, &hi, &lo);
    *reinterpret_cast<uint4*>(smem + hi_off + c * (TC_M * 16) + m * 16) = hi;
    *reinterpret_cast<uint4*>(smem + lo_off + c * (TC_M * 16) + m * 16) = lo;
    if (stash_mat != nullptr)
      *reinterpret_cast<uint4*>(stash_mat + 2 * ren_blk(sidx, c * 8, kp)) = hi;
  }
}

// K sources of a step in MMA issue order: 0 = shared-memory segment, 1 = TMEM K-half 0, 2 = TMEM K-half 1
__host__ __device__ inline int step_src_ksteps(const TcStep& st, int src) {
  if (src == 0) return st.seg == 1 ? 4 : (st.seg == 2 ? 2 : 0);
  if (src == 1) return st.a_k >= 128 ? 8 : 0;
  return st.a_k >= 256 ? 8 : 0;
}

__host__ __device__ inline int step_last_src(const TcStep& st) {
  return st.a_k >= 256 ? 2 : (st.a_k >= 128 ? 1 : 0);
}
// 64-bit shared-memory matrix descriptor from its low word (address/16 | LBO/16 << 16); SBO = 128 B, version 1
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo32) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo32), "r"(0x4008u));
  return d;
}

// MODE 0: forward (VAR 0 value pass, 1 tangent pass); MODE 1: dgrad (VAR = pass 0 / 1 / 2)
template <int MODE, bool STASH, int VAR>
__global__ void __launch_bounds__(MODE == 0 ? TC_THREADS_FWD : TC_THREADS_BWD, 1) ren_tc_chain(const TcArgs a) {
  constexpr bool TAN = MODE == 0 && VAR == 1;
  constexpr int PASS = MODE == 1 ? VAR : 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  const RenPlan& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + SM_BAR;                 // [TC_STAGES <= 16]
  const uint32_t bar_empty = sbase + SM_BAR + 128;          // [TC_STAGES <= 16]
  const uint32_t bar_p = sbase + SM_BAR + 256;              // [2] encodings of a tile are in shared-memory buffer b
  const uint32_t bar_pfree = sbase + SM_BAR + 304;          // [2] the MMAs that read buffer b have completed
  const uint32_t bar_a = sbase + SM_BAR + 272;              // [2] A K-half h of the next step is in TMEM
  const uint32_t bar_acc = sbase + SM_BAR + 288;            // [2] accumulator N-half h is complete
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_PTR);
  float* vec = reinterpret_cast<float*>(smem + SM_VEC);
  const long long S = a.sample_idx != nullptr ? a.n_samples : a.n_rays * a.K;
  const long long n_tiles = (S + TC_M - 1) / TC_M;
  const int C1 = 1 + p.C;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_p + 8 * h, TC_PRO_WARPS);  // one elected lane per prologue warp arrives
      mbar_init(bar_pfree + 8 * h, 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_a + 8 * h, TC_EPI_WARPS);
      mbar_init(bar_acc + 8 * h, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < VEC_FLOATS; i += blockDim.x) vec[i] = a.vec[i];
  for (int i = threadIdx.x; i < a.n_ops * 8; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem + SM_PROG)[i] = reinterpret_cast<const uint32_t*>(a.prog)[i];
  if (MODE == 0 && REN_BIAS_MMA) {
    // K core 0: row m = (1, 0, 0, 0, 0, 0, 0, 0); K core 1: zeros
    for (int i = threadIdx.x; i < 2 * TC_M; i += blockDim.x)
      reinterpret_cast<uint4*>(smem + SM_ONES)[i] = make_uint4(i < TC_M ? 0x00003C00u : 0u, 0u, 0u, 0u);
    fence_proxy_async();          // read by the tensor core (async proxy)
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + SM_TMEM_PTR) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (REN_ONE_LANE()) {
      const uint4* prog = reinterpret_cast<const uint4*>(smem + SM_PROG);
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (a.tile_live != nullptr && !__ldg(a.tile_live + tile)) continue;
        for (int i = 0; i < a.n_ops; ++i, ++it) {
          const uint4 o0 = prog[2 * i];
          const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          if (a.dbg & 8) {
            mbar_arrive(bar_full + 8 * s);
          } else {
            mbar_arrive_expect_tx(bar_full + 8 * s, o0.y);
            bulk_g2s(sbase + SM_RING + s * SM_STAGE_BYTES, a.pack + o0.x, o0.y, bar_full + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (REN_ONE_LANE()) {
      const uint4* prog = reinterpret_cast<const uint4*>(smem + SM_PROG);
      uint32_t it = 0, n_p = 0, n_a0 = 0, n_a1 = 0;
      const uint32_t idesc128 = umma_idesc(128, 0, 0), idesc256 = umma_idesc(256, 0, 0);
      const uint32_t split3 = a.split3 != 0, a_lo = a.split3 == 1, nomma = a.dbg & 2;
      // low words of the shared-memory descriptors are linear in the address (all < 256 KB)
      const uint32_t ring_lo = ((sbase + SM_RING) >> 4) & 0x3FFFu;
      const uint32_t smem_lo = ((sbase >> 4) & 0x3FFFu) | (128u << 16);      // encodings: K-core stride 2048 B
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (a.tile_live != nullptr && !__ldg(a.tile_live + tile)) continue;
        const uint32_t eb = n_p & 1u;                                   // encoding buffer of this tile
        const uint32_t enc_lo = smem_lo + eb * (SM_ENC_BYTES >> 4);
        uint4 o0 = prog[0], o1 = prog[1];
        for (int i = 0; i < a.n_ops; ++i, ++it) {
          const uint32_t flags = o1.y, d_addr = tmem_base + o0.z, a0 = o0.w, a_ld = o1.x;
          // prefetch the next op while this one runs
          const int nx = (i + 1 < a.n_ops) ? i + 1 : 0;
          const uint4 p0 = prog[2 * nx], p1 = prog[2 * nx + 1];
          const uint32_t wt = (flags >> 6) & 3u;
          const uint32_t s = it % TC_STAGES;
#if REN_ISSUE_FAST
          // everything the MMAs of this op need that does not depend on the hand-off is computed BEFORE the waits (the
          // empty asm pins it there), and the issuer -- one thread -- polls its barriers without the suspend hint: the
          // epilogue -> first-MMA hand-off is on the critical cycle of every layer
          const bool wide = (flags & 8u) != 0u;
          uint32_t idesc = wide ? idesc256 : idesc128;
          const uint32_t chunk16 = wide ? 1024u : 512u;               // chunk bytes / 16
          uint32_t b_lo32 = (ring_lo + s * (SM_STAGE_BYTES >> 4)) | ((wide ? 256u : 128u) << 16);
          const int cnt = (int)(flags & 7u);
          uint32_t acc = (flags & 32u) ? 0u : 1u;
          asm volatile("" : "+r"(idesc), "+r"(b_lo32), "+r"(acc));
          if (wt == 1u) { mbar_wait_spin(bar_a, n_a0 & 1u); ++n_a0; tc_fence_after(); }
          else if (wt == 2u) { mbar_wait_spin(bar_a + 8, n_a1 & 1u); ++n_a1; tc_fence_after(); }
          else if (wt == 3u) { mbar_wait_spin(bar_p + 8 * eb, (n_p >> 1) & 1u); tc_fence_after(); }
          mbar_wait_spin(bar_full + 8 * s, (it / TC_STAGES) & 1u);
          tc_fence_after();
#else
          if (wt == 1u) { mbar_wait(bar_a, n_a0 & 1u); ++n_a0; tc_fence_after(); }
          else if (wt == 2u) { mbar_wait(bar_a + 8, n_a1 & 1u); ++n_a1; tc_fence_after(); }
          else if (wt == 3u) { mbar_wait(bar_p + 8 * eb, (n_p >> 1) & 1u); tc_fence_after(); }
          mbar_wait(bar_full + 8 * s, (it / TC_STAGES) & 1u);
          tc_fence_after();
          const bool wide = (flags & 8u) != 0u;
          const uint32_t idesc = wide ? idesc256 : idesc128;
          const uint32_t chunk16 = wide ? 1024u : 512u;               // chunk bytes / 16
          const uint32_t b_lo32 = (ring_lo + s * (SM_STAGE_BYTES >> 4)) | ((wide ? 256u : 128u) << 16);
          const int cnt = (int)(flags & 7u);
          uint32_t acc = (flags & 32u) ? 0u : 1u;
#endif
          if (!nomma) {
            if (flags & OP_ONES) {
              const uint64_t da = mk_desc(smem_lo + (uint32_t)(SM_ONES >> 4));
              tc_mma_f16(d_addr, da, mk_desc(b_lo32), idesc, acc);
              if (split3) tc_mma_f16(d_addr, da, mk_desc(b_lo32 + (chunk16 >> 1)), idesc, 1u);
            } else if (flags & 16u) {
              const uint32_t a_lo32 = enc_lo + a0;
#pragma unroll
              for (int kl = 0; kl < TC_STAGE_KS; ++kl) {
                if (kl < cnt) {
                  const uint64_t da_hi = mk_desc(a_lo32 + kl * 256);
                  if (REN_WIDE_SPLIT && wide) {
                    // rows 128..255 of the [256 x 16] chunk start 2048 B into each K core
                    const uint32_t bh = b_lo32 + kl * chunk16, bl = bh + (chunk16 >> 1);
                    tc_mma_f16(d_addr, da_hi, mk_desc(bh), idesc128, acc);
                    tc_mma_f16(d_addr + 128u, da_hi, mk_desc(bh + 128u), idesc128, acc);
                    if (split3) {
                      const uint64_t da_lo = mk_desc(a_lo32 + kl * 256 + a_ld);
                      tc_mma_f16(d_addr, da_lo, mk_desc(bh), idesc128, 1u);
                      tc_mma_f16(d_addr + 128u, da_lo, mk_desc(bh + 128u), idesc128, 1u);
                      tc_mma_f16(d_addr, da_hi, mk_desc(bl), idesc128, 1u);
                      tc_mma_f16(d_addr + 128u, da_hi, mk_desc(bl + 128u), idesc128, 1u);
                    }
                  } else {
                    const uint64_t db_hi = mk_desc(b_lo32 + kl * chunk16);
                    tc_mma_f16(d_addr, da_hi, db_hi, idesc, acc);
                    if (split3) {
                      tc_mma_f16(d_addr, mk_desc(a_lo32 + kl * 256 + a_ld), db_hi, idesc, 1u);
                      tc_mma_f16(d_addr, da_hi, mk_desc(b_lo32 + kl * chunk16 + (chunk16 >> 1)), idesc, 1u);
                    }
                  }
                  acc = 1u;
                }
              }
            } else {
              const uint32_t a_tm = tmem_base + a0;
#pragma unroll
              for (int kl = 0; kl < TC_STAGE_KS; ++kl) {
                if (kl < cnt) {
                  // K step ks0 + kl (ks0 even): hi columns at +32 (kl/2) + 8 (kl%2), lo 16 further
                  const uint32_t a_hi = a_tm + (uint32_t)(32 * (kl >> 1) + 8 * (kl & 1));
                  if (REN_WIDE_SPLIT && wide) {
                    const uint32_t bh = b_lo32 + kl * chunk16, bl = bh + (chunk16 >> 1);
                    tc_mma_f16_ts(d_addr, a_hi, mk_desc(bh), idesc128, acc);
                    tc_mma_f16_ts(d_addr + 128u, a_hi, mk_desc(bh + 128u), idesc128, acc);
                    if (split3) {
                      if (a_lo) {
                        tc_mma_f16_ts(d_addr, a_hi + a_ld, mk_desc(bh), idesc128, 1u);
                        tc_mma_f16_ts(d_addr + 128u, a_hi + a_ld, mk_desc(bh + 128u), idesc128, 1u);
                      }
                      tc_mma_f16_ts(d_addr, a_hi, mk_desc(bl), idesc128, 1u);
                      tc_mma_f16_ts(d_addr + 128u, a_hi, mk_desc(bl + 128u), idesc128, 1u);
                    }
                  } else {
                    const uint64_t db_hi = mk_desc(b_lo32 + kl * chunk16);
                    tc_mma_f16_ts(d_addr, a_hi, db_hi, idesc, acc);
                    if (split3) {
                      if (a_lo) tc_mma_f16_ts(d_addr, a_hi + a_ld, db_hi, idesc, 1u);
                      tc_mma_f16_ts(d_addr, a_hi, mk_desc(b_lo32 + kl * chunk16 + (chunk16 >> 1)), idesc, 1u);
                    }
                  }
                  acc = 1u;
                }
              }
            }
          }
          tc_commit(bar_empty + 8 * s);          // frees the ring slot when these MMAs have read it
          const uint32_t ca = (flags >> 8) & 3u;
          if (ca) tc_commit(bar_acc + 8 * (ca - 1u));   // accumulator half complete -> epilogue
          o0 = p0; o1 = p1;
        }
        if (MODE == 0) tc_commit(bar_pfree + 8 * eb);   // every MMA reading this tile's encodings is done
        ++n_p;                                          // live tiles only
      }
    }
  } else if (warp >= 2 + TC_EPI_WARPS) {
    // ===================== forward prologue warps (one tile ahead of the MMAs) =====================
    // march -> sample position -> contraction -> sin/cos encodings of one sample row per thread, written
    // as fp16 hi/lo K-major operands into the free encoding buffer (+ the fp16 stash when training)
    if (MODE == 0) {
      float* sel = reinterpret_cast<float*>(smem + SM_SEL);
      uint32_t n_t = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (a.tile_live != nullptr && !__ldg(a.tile_live + tile)) continue;
        const uint32_t eb = n_t & 1u;
        mbar_wait(bar_pfree + 8 * eb, ((n_t >> 1) & 1u) ^ 1u);       // the MMAs of tile n_t - 2 are done with it
#pragma unroll 1
        for (int m = (warp - 2 - TC_EPI_WARPS) * 32 + lane; m < TC_M; m += 32 * TC_PRO_WARPS) {
        const long long sidx = tile * TC_M + m;
        const bool valid = sidx < S && !(a.dbg & 4);
        float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f}, z[3] = {0.f, 0.f, 0.f};
        bool selector = false, skipped = false;
        const long long dense = (valid && a.sample_idx != nullptr) ? (long long)a.sample_idx[sidx] : sidx;
        if (valid) {
          const long long r = dense / a.K;
          const int k = (int)(dense - r * a.K);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            o[c] = a.rays_o[r * 3 + c];
            d[c] = a.rays_d[r * 3 + c];
          }
          float t_min, t_max, t0, t1;
          ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
          ren_march_to(p.r, t_min, k, &t0, &t1);
          selector = ren_field_input(p.f, o, d, t0, t1, z);
          if (!ren_alive(t0, t1, t_max)) selector = false;
          skipped = !ren_occupied(p.f, p.r, o, d, t0, t1);      // unoccupied cell: the marcher emits no sample
        }
        const float dz[3] = {__fmul_rn(d[0], 3.141592653589793f), __fmul_rn(d[1], 3.141592653589793f),
                             __fmul_rn(d[2], 3.141592653589793f)};
        float zdot[3] = {0.f, 0.f, 0.f}, dzdot[3] = {0.f, 0.f, 0.f};
        if (TAN && valid) {
          // x = o + d tm, xdot = odot + ddot tm (the march carries no gradient); zdot = J_contract(x) xdot
          const long long r = dense / a.K;
          const int k = (int)(dense - r * a.K);
          float t_min, t_max, t0, t1;
          ren_ray_range(p.f, p.r, o, d, a.jitter ? a.jitter[r] : 0.f, &t_min, &t_max);
          ren_march_to(p.r, t_min, k, &t0, &t1);
          const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
          Dual pd[3], zd[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float od = a.rays_o_dot[r * 3 + c], dd = a.rays_d_dot[r * 3 + c];
            pd[c] = Dual(__fadd_rn(o[c], __fmul_rn(__fmul_rn(d[c], __fadd_rn(t0, t1)), 0.5f)), fmaf(dd, tm, od));
            dzdot[c] = dd * 3.141592653589793f;
          }
          ren_contract_t<Dual>(p.f, pd, zd);
#pragma unroll
          for (int c = 0; c < 3; ++c) zdot[c] = zd[c].d;
        }
        unsigned char* buf = smem + eb * SM_ENC_BYTES;
        encode_cores(z, TAN ? zdot : nullptr, 10, 64, 0, 8, valid, buf, SM_P_HI, SM_P_LO, m,
                     STASH ? a.stash + tcs_encp(p, S) : nullptr, sidx);
        encode_cores(dz, TAN ? dzdot : nullptr, 4, 32, 0, 4, valid, buf, SM_D_HI, SM_D_LO, m,
                     STASH ? a.stash + tcs_encd(p, S) : nullptr, sidx);
        sel[eb * 128 + m] = skipped ? 2.f : (selector ? 1.f : 0.f);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p + 8 * eb);
        ++n_t;                                          // live tiles only
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const bool a_single = a.split3 != 1;     // the activations go back to tensor memory as single fp16 values
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;          // column group 0..3: 32 of the 128 columns of a half
    const int m = q * 32 + lane;             // row of the tile == TMEM lane
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float4* part = reinterpret_cast<float4*>(smem + SM_PART);   // [4][128]
    uint32_t n_acc0 = 0u, n_acc1 = 0u;       // scalars: an indexed array would live in local memory
    uint32_t n_t = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long sidx = tile * TC_M + m;
      if (a.tile_live != nullptr && !__ldg(a.tile_live + tile)) {
        // dead tile: no marching step of it emits a sample -- NaN in the raw sigma column means "no sample" to
        // the compositing kernels; stash / delta tiles stay untouched and are skipped by every consumer
        if (MODE == 0 && !TAN && cg == 0 && sidx < S) a.raw[sidx * C1] = __int_as_float(0x7fc00000);
        continue;
      }
      const bool valid = sidx < S && !(MODE == 0 && (a.dbg & 4));
      float selector_code = 0.f;             // 0 outside the contracted box, 1 inside, 2 unoccupied cell (no sample)
      float gs = 0.f;                        // dgrad: scaled d loss / d sigma_raw of this row
      float head_acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (MODE == 0) {
        // the encodings of this tile come from the prologue warps (one tile ahead)
      } else {
        // ---- dgrad prologue: Delta of the view-conditioned layer, elementwise:
        //      d g[n] = act'(g[n]) * sum_c (d c_raw[c]) W_o[c][n]  -> A K-half 0 of the first step
        const TcStep& st0 = a.steps[0];
        const int mat = p.cond[0].mat;
        const size_t tile_off = (size_t)tile * (size_t)(p.WC >> 3) * 2048 + (size_t)m * 16 + (size_t)(cg * 4) * 2048;
        uint4 act[4], aux[4];          // aux: tangent activations (pass 1) or E (pass 2)
        if (valid) {
          const unsigned char* am = a.stash + tcs_mat(p, S, mat) + tile_off;
#pragma unroll
          for (int i = 0; i < 4; ++i) act[i] = __ldg(reinterpret_cast<const uint4*>(am + i * 2048));
          if (PASS) {
            const unsigned char* xm = (PASS == 1 ? a.stash_dot : a.extra) + tcs_mat(p, S, mat) + tile_off;
#pragma unroll
            for (int i = 0; i < 4; ++i) aux[i] = __ldg(reinterpret_cast<const uint4*>(xm + i * 2048));
          }
        }
        float gc[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const float sc = delta_scale(a.gmax);
          gs = a.graw[sidx * C1] * sc;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < p.C) gc[c] = a.graw[sidx * C1 + 1 + c] * sc;
        }
        uint32_t hi16[16], lo16[16];
        unsigned char* dm = a.delta + tcs_mat(p, S, mat) + tile_off;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          float av[8], xv[8], y[8], e[8];
          unpack8(act[j8], av);
          if (PASS) unpack8(aux[j8], xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = cg * 32 + j8 * 8 + j;
            float v = gc[0] * vec[VEC_WO + n];
            if (p.C > 1) v = fmaf(gc[1], vec[VEC_WO + 128 + n], v);
            if (p.C > 2) v = fmaf(gc[2], vec[VEC_WO + 256 + n], v);
            y[j] = valid ? dgrad_act(a.softplus, PASS, v, av[j], xv[j], &e[j]) : 0.f;
            if (!valid) e[j] = 0.f;
          }
          uint4 hi, lo;
          if (a_single) { pack8(y, &hi); lo = make_uint4(0u, 0u, 0u, 0u); } else split8(y, &hi, &lo);
          hi16[j8 * 4 + 0] = hi.x; hi16[j8 * 4 + 1] = hi.y; hi16[j8 * 4 + 2] = hi.z; hi16[j8 * 4 + 3] = hi.w;
          lo16[j8 * 4 + 0] = lo.x; lo16[j8 * 4 + 1] = lo.y; lo16[j8 * 4 + 2] = lo.z; lo16[j8 * 4 + 3] = lo.w;
          *reinterpret_cast<uint4*>(dm + j8 * 2048) = hi;
          if (PASS == 1) {
            split8(e, &hi, &lo);
            *reinterpret_cast<uint4*>(a.extra + tcs_mat(p, S, mat) + tile_off + j8 * 2048) = hi;
          }
        }
        const uint32_t tcol = tmem_lane + (uint32_t)(st0.a_col + 32 * cg);
        tc_st16(tcol, hi16);
        if (!a_single) tc_st16(tcol + 16, lo16);
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a);
      }

      // dgrad: staging state of the activation-stash copies (this thread's 16-byte slots)
      const uint32_t d_stage = sbase + (uint32_t)((warp - 2) * 32 + lane) * 16u;
      const unsigned char* d_stage_gen = smem + ((warp - 2) * 32 + lane) * 16;
      uint32_t d_buf = 0u;
      bool d_have = false;
      // ---- one epilogue per GEMM step and N-half
      for (int g = 0; g < a.n_steps; ++g) {
        const TcStep& st = a.steps[g];
        const int nh = st.N >> 7;
        for (int h = 0; h < nh; ++h) {
          const int n0 = 128 * h + 32 * cg;            // first of this thread's 32 features
          // byte offset of this thread's first 8-feature group inside a blocked fp16 matrix of width N
          const size_t tile_off = (size_t)tile * (size_t)(st.N >> 3) * 2048 + (size_t)m * 16 + (size_t)(n0 >> 3) * 2048;
          uint4 act[4];
          bool staged_next = false;
          if (MODE == 1 && st.epi == EPI_D_ACT && valid) {
            // this thread's slice of the activation stash arrives through shared memory: the copy for THIS
            // half-epilogue was issued during the previous one (a whole epilogue of lead time, no registers
            // held across the accumulator wait); issue the next one now
            if (!d_have) {
              const unsigned char* am = a.stash + tcs_mat(p, S, st.act_mat) + tile_off;
#pragma unroll
              for (int i = 0; i < 4; ++i) cp_async16(d_stage + SM_DSTAGE_ACT + d_buf * 32768 + i * 8192, am + i * 2048);
              cp_async_commit();
            }
            if (PASS) {
              const unsigned char* xm = (PASS == 1 ? a.stash_dot : a.extra) + tcs_mat(p, S, st.act_mat) + tile_off;
#pragma unroll
              for (int i = 0; i < 4; ++i) cp_async16(d_stage + SM_DSTAGE_AUX + i * 8192, xm + i * 2048);
              cp_async_commit();
            }
            const int g2 = h + 1 < nh ? g : g + 1, h2 = h + 1 < nh ? h + 1 : 0;
            if (g2 < a.n_steps && a.steps[g2].epi == EPI_D_ACT) {
              const TcStep& s2 = a.steps[g2];
              const unsigned char* am = a.stash + tcs_mat(p, S, s2.act_mat) + (size_t)tile * (size_t)(s2.N >> 3) * 2048 +
                                        (size_t)m * 16 + (size_t)((128 * h2 + 32 * cg) >> 3) * 2048;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                cp_async16(d_stage + SM_DSTAGE_ACT + (d_buf ^ 1u) * 32768 + i * 8192, am + i * 2048);
              cp_async_commit();
              staged_next = true;
            }
          }
          if (MODE == 0 && TAN && st.epi != EPI_F_LINEAR && valid) {
            // tangent pass: act'(pre) comes from the value pass' stash of this layer's output
            const unsigned char* am = a.stash_val + tcs_mat(p, S, st.out_mat) + tile_off;
#pragma unroll
            for (int i = 0; i < 4; ++i) act[i] = __ldg(reinterpret_cast<const uint4*>(am + i * 2048));
          }
          if (h == 0) { mbar_wait(bar_acc, n_acc0 & 1u); ++n_acc0; }
          else { mbar_wait(bar_acc + 8, n_acc1 & 1u); ++n_acc1; }
          tc_fence_after();
          // the prologue warps published the selector before bar_p of this tile, which the first MMA waited
          // for; read it now -- the buffer may be recycled as soon as the tile's last MMA completes
          if (MODE == 0 && g == 0 && h == 0)
            selector_code = reinterpret_cast<const volatile float*>(smem + SM_SEL)[(n_t & 1u) * 128 + m];
          const uint32_t tcol = tmem_lane + (uint32_t)(st.acc_col + n0);
          uint32_t r[32];
          if (a.dbg & 16) {
            if (st.write_a) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bar_a + 8 * h); }
            continue;
          }
          tc_ld16(tcol, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
          tc_ld16(tcol + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
          tc_ld_wait();
          if (MODE == 1 && st.epi == EPI_D_ACT && valid) {
            if (staged_next) cp_async_wait<1>(); else cp_async_wait<0>();   // all but the copy for the next half
          }
          unsigned char* out_mat = nullptr;        // already offset to this thread's row / first group
          if (MODE == 0) {
            if (STASH && st.out_mat >= 0) out_mat = a.stash + tcs_mat(p, S, st.out_mat) + tile_off;
          } else if (st.out_mat >= 0) {            // padded rows are written as zeros (wgrad reads whole tiles)
            out_mat = a.delta + tcs_mat(p, S, st.out_mat) + tile_off;
          }
          const float* bias = (MODE == 0 && st.bias_idx >= 0) ? vec + VEC_BIAS + st.bias_idx * 256 + n0 : nullptr;
          // the common case (hidden softplus layer of the value pass, split operands) as ONE flat pass over the
          // thread's 32 columns: 32 independent bias -> ex2 -> lg2 chains in flight instead of 8, and no per-group
          // tests of the layer kind / destinations.  Software pipeline: the MUFU chains of group j8 + FLAT_AHEAD are
          // issued before the split / stores of group j8, so the special-function pipe and the ALU work overlap
          // inside every warp
          bool done = false;
          if (REN_FLAT_FWD && MODE == 0 && !TAN) {
            // (with REN_BIAS_MMA the host marks exactly these steps -- the bias is already in the accumulator)
            if (REN_BIAS_MMA ? st.bias_mma != 0
                             : (st.epi == EPI_F_HIDDEN && a.softplus && !a_single && !(a.dbg & 1) &&
                                (!STASH || st.out_mat >= 0))) {
              auto sp8 = [&](int j8) {
#if REN_BIAS_MMA
#pragma unroll
                for (int j = 8 * j8; j < 8 * j8 + 8; ++j) r[j] = __float_as_uint(softplus_log2(__uint_as_float(r[j])));
#else
#pragma unroll
                for (int j4 = 2 * j8; j4 < 2 * j8 + 2; ++j4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bias + j4 * 4);
                  r[j4 * 4 + 0] = __float_as_uint(softplus_log2(__uint_as_float(r[j4 * 4 + 0]) + b4.x));
                  r[j4 * 4 + 1] = __float_as_uint(softplus_log2(__uint_as_float(r[j4 * 4 + 1]) + b4.y));
                  r[j4 * 4 + 2] = __float_as_uint(softplus_log2(__uint_as_float(r[j4 * 4 + 2]) + b4.z));
                  r[j4 * 4 + 3] = __float_as_uint(softplus_log2(__uint_as_float(r[j4 * 4 + 3]) + b4.w));
                }
#endif
              };
#pragma unroll
              for (int j8 = 0; j8 < FLAT_AHEAD; ++j8) sp8(j8);
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                if (j8 + FLAT_AHEAD < 4) sp8(j8 + FLAT_AHEAD);
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[j8 * 8 + j]);
                if (st.sig_head) {
                  const float4 w0 = *reinterpret_cast<const float4*>(vec + VEC_WSIG + n0 + j8 * 8);
                  const float4 w1 = *reinterpret_cast<const float4*>(vec + VEC_WSIG + n0 + j8 * 8 + 4);
                  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) head_acc[0] = fmaf(y[j], ww[j], head_acc[0]);
                }
                uint4 hi, lo;
                split8(y, &hi, &lo);
                tc_st4(tcol + j8 * 4, hi);                     // a hidden layer always feeds the next step
                tc_st4(tcol + 16 + j8 * 4, lo);
                if (STASH) *reinterpret_cast<uint4*>(out_mat + j8 * 2048) = hi;
              }
              done = true;
            }
          }
          // plain dgrad through a hidden softplus layer: y = v (1 - 2^-g) = fma(-v, 2^-g, v) with g from the staged
          // activation stash; one pass over the 32 columns, the ex2 chains of the next group ahead of the split / stores
          // of this one.  Rows past the last sample (last tile only) write zero deltas (the weight-gradient kernel
          // reads whole tiles); what they leave in tensor memory stays inside their own row of every later MMA
          if (REN_FLAT_BWD && MODE == 1 && PASS == 0) {
            if (st.epi == EPI_D_ACT && a.softplus && !a_single && st.out_mat >= 0 && !(a.dbg & 1)) {
              const unsigned char* sa = d_stage_gen + SM_DSTAGE_ACT + d_buf * 32768;
              const float* wsig = vec + VEC_WSIG + n0;
              auto da8 = [&](int j8) {
                float av[8];
                unpack8(*reinterpret_cast<const uint4*>(sa + j8 * 8192), av);
                if (st.rank1) {
                  const float4 w0 = *reinterpret_cast<const float4*>(wsig + j8 * 8);
                  const float4 w1 = *reinterpret_cast<const float4*>(wsig + j8 * 8 + 4);
                  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    r[j8 * 8 + j] = __float_as_uint(fmaf(gs, ww[j], __uint_as_float(r[j8 * 8 + j])));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float v = __uint_as_float(r[j8 * 8 + j]);
                  r[j8 * 8 + j] = __float_as_uint(fmaf(-v, ex2_ftz(-av[j]), v));
                }
              };
              da8(0);
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                if (j8 + 1 < 4) da8(j8 + 1);
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[j8 * 8 + j]);
                uint4 hi, lo;
                split8(y, &hi, &lo);
                if (st.write_a) {
                  tc_st4(tcol + j8 * 4, hi);
                  tc_st4(tcol + 16 + j8 * 4, lo);
                }
                if (!valid) hi = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(out_mat + j8 * 2048) = hi;
              }
              done = true;
            }
          }
          if (!done) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float y[8];
            float e8[8];
            if (MODE == 0 && TAN) {
              // d/dw rows: no bias; gdot = act'(pre) * (W gdot_prev) with act' from the value stash
              float av[8];
              if (st.epi != EPI_F_LINEAR) unpack8(act[j8], av);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float x = __uint_as_float(r[j8 * 8 + j]);
                if (st.epi == EPI_F_LINEAR) y[j] = x * st.acc_scale;
                else y[j] = valid ? x * (a.softplus ? softplus_grad_log2(av[j]) : (av[j] > 0.f ? 1.f : 0.f)) : 0.f;
                // d/dw activations scale with the camera speed: keep them inside fp16 (the hi part is the operand)
                y[j] = fminf(fmaxf(y[j], -60000.f), 60000.f);
              }
            } else if (MODE == 0) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias + j8 * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bias + j8 * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (st.epi == EPI_F_LINEAR) {
                  y[j] = fmaf(__uint_as_float(r[j8 * 8 + j]), st.acc_scale, bb[j]);
                } else {
                  const float x = __uint_as_float(r[j8 * 8 + j]) + bb[j];
                  y[j] = (a.dbg & 1) ? x : (a.softplus ? softplus_log2(x) : fmaxf(x, 0.f));
                }
              }
            }
            if (MODE == 0) {
              if (st.sig_head) {
                const float4 w0 = *reinterpret_cast<const float4*>(vec + VEC_WSIG + n0 + j8 * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(vec + VEC_WSIG + n0 + j8 * 8 + 4);
                const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) head_acc[0] = fmaf(y[j], ww[j], head_acc[0]);
              }
              if (st.epi == EPI_F_LAST) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  if (c < p.C) {
                    const float* wo = vec + VEC_WO + c * 128 + n0 + j8 * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) head_acc[1 + c] = fmaf(y[j], wo[j], head_acc[1 + c]);
                  }
                }
              }
            } else {
              float av[8], xv[8];
              if (st.epi == EPI_D_ACT) {
                unpack8(*reinterpret_cast<const uint4*>(d_stage_gen + SM_DSTAGE_ACT + d_buf * 32768 + j8 * 8192), av);
                if (PASS) unpack8(*reinterpret_cast<const uint4*>(d_stage_gen + SM_DSTAGE_AUX + j8 * 8192), xv);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float v = __uint_as_float(r[j8 * 8 + j]);
                if (st.rank1) v = fmaf(gs, vec[VEC_WSIG + n0 + j8 * 8 + j], v);
                e8[j] = 0.f;
                if (st.epi == EPI_D_ACT) v = dgrad_act(a.softplus, PASS, v, av[j], xv[j], &e8[j]);
                y[j] = valid ? v : 0.f;
                if (!valid) e8[j] = 0.f;
              }
            }
            uint4 hi, lo;
            if (a_single) pack8(y, &hi); else split8(y, &hi, &lo);
            // every accumulator column of this thread is already in registers: the packed operand goes back over
            // them group by group (keeps 32 registers free compared with one 16-column store at the end)
            if (st.write_a) {
              tc_st4(tcol + j8 * 4, hi);
              if (!a_single) tc_st4(tcol + 16 + j8 * 4, lo);
            }
            if (out_mat != nullptr) *reinterpret_cast<uint4*>(out_mat + j8 * 2048) = hi;
            if (MODE == 1 && PASS == 1 && st.epi == EPI_D_ACT && st.out_mat >= 0) {
              uint4 eh;                               // fp16 only: the value-adjoint pass adds E in fp32
              __half2* e2 = reinterpret_cast<__half2*>(&eh);
#pragma unroll
              for (int j = 0; j < 4; ++j) e2[j] = __floats2half2_rn(e8[2 * j], e8[2 * j + 1]);
              *reinterpret_cast<uint4*>(a.extra + tcs_mat(p, S, st.out_mat) + tile_off + j8 * 2048) = eh;
            }
          }
          }   // !done
          if (MODE == 1 && st.epi == EPI_D_ACT && valid) {
            d_have = staged_next;
            if (staged_next) d_buf ^= 1u;
          }
          if (st.write_a) {
            tc_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a + 8 * h);
          }
        }
      }
      if (MODE == 0) {
        // ---- heads: combine the four column groups of each row
        part[cg * 128 + m] = make_float4(head_acc[0], head_acc[1], head_acc[2], head_acc[3]);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        if (cg == 0 && sidx < S) {
          const float4 p0 = part[m], p1 = part[128 + m], p2 = part[256 + m], p3 = part[384 + m];
          const float hb = TAN ? 0.f : 1.f;          // the d/dw rows carry no bias
          const float sraw = (p0.x + p1.x) + (p2.x + p3.x) + hb * vec[VEC_HB];
          a.raw[sidx * C1] = TAN ? sraw
                                 : (selector_code == 2.f ? __int_as_float(0x7fc00000)
                                                         : (selector_code != 0.f ? sraw : -INFINITY));
          a.raw[sidx * C1 + 1] = (p0.y + p1.y) + (p2.y + p3.y) + hb * vec[VEC_HB + 1];
          if (p.C > 1) a.raw[sidx * C1 + 2] = (p0.z + p1.z) + (p2.z + p3.z) + hb * vec[VEC_HB + 2];
          if (p.C > 2) a.raw[sidx * C1 + 3] = (p0.w + p1.w) + (p2.w + p3.w) + hb * vec[VEC_HB + 3];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
      }
      ++n_t;                                            // live tiles only
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ weight packing
// A pack job writes consecutive [rows x 16] B chunks (rows = 128 or 256):
// chunk = [hi: 2 K-cores x rows x 16 B][lo: same].
//   mode 0: B[n][k] = scale * W[(n0+n)*ldw + col0 + k]   (forward;  k < Ktrue, n0+n < Ntrue)
//   mode 1: B[n][k] = W[(col0+k)*ldw + n0 + n]           (dgrad: transposed)
//   mode 2: vec[dst + i] = scale * params[w_off + i], i < Ntrue (zero up to n_ksteps entries)
#define TC_MAX_PACK 64
struct PackJob {
  long long dst_off;   // bytes (modes 0/1) or floats (mode 2)
  int mode, n0, Ntrue, n_ksteps, k0, Ktrue, rows;
  int w_off, ldw, col0;
  float scale;
};
struct PackArgs {
  PackJob jobs[TC_MAX_PACK];
  int n_jobs;
  const float* params;
  unsigned char* pack;
  float* vec;
};

extern "C" __global__ void ren_tc_pack(const PackArgs a) {
  const int g = blockIdx.y;
  if (g >= a.n_jobs) return;
  const PackJob& jb = a.jobs[g];
  if (jb.mode == 2) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < jb.n_ksteps; n += gridDim.x * blockDim.x)
      a.vec[jb.dst_off + n] = n < jb.Ntrue ? a.params[jb.w_off + n] * jb.scale : 0.f;
    return;
  }
  const int total = jb.n_ksteps * 16 * jb.rows;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int n, kc;
    if (jb.mode == 0) { kc = e % (jb.n_ksteps * 16); n = e / (jb.n_ksteps * 16); }   // reads run along k
    else { n = e % jb.rows; kc = e / jb.rows; }                                      // reads run along n
    const int c = kc >> 4, kk = kc & 15;
    const int k = jb.k0 + kc;
    float w = 0.f;
    if (jb.n0 + n < jb.Ntrue && k < jb.Ktrue) {
      if (jb.mode == 0) w = a.params[jb.w_off + (long long)(jb.n0 + n) * jb.ldw + jb.col0 + k] * jb.scale;
      else w = a.params[jb.w_off + (long long)(jb.col0 + k) * jb.ldw + jb.n0 + n];
    }
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    unsigned char* chunk = a.pack + jb.dst_off + (size_t)c * jb.rows * 64;
    const size_t off = (size_t)(kk >> 3) * (jb.rows * 16) + (size_t)n * 16 + (kk & 7) * 2;
    *reinterpret_cast<__half*>(chunk + off) = hi;
    *reinterpret_cast<__half*>(chunk + (size_t)jb.rows * 32 + off) = lo;
  }
}

// ------------------------------------------------------------------ host side
struct TcProgram {
  TcStep steps[TC_MAX_STEPS];
  int n_steps;
  StageOp ops[TC_MAX_OPS];
  int n_ops;
  PackJob jobs[TC_MAX_PACK];
  int n_jobs;
  long long pack_bytes;    // packed weights
};

static void add_vec(TcProgram& P, int dst, int count, int n_valid, int w_off, float scale) {
  PackJob& j = P.jobs[P.n_jobs++];
  memset(&j, 0, sizeof(j));
  j.dst_off = dst; j.mode = 2; j.n_ksteps = count; j.Ntrue = n_valid; j.w_off = w_off; j.scale = scale;
}

// chunks of one step in MMA issue order: source (segment, K-half 0, K-half 1) -> N-half -> K step
static void add_step_pack(TcProgram& P, TcStep& s, long long& off, int mode, int w_off, int ldw, int Ntrue,
                          int main_col0, int main_K, float main_scale, int seg_col0, int seg_K, float seg_scale,
                          int b_off = -1, float b_scale = 1.f) {
  s.pack_off = off;
  const int nh = s.N >> 7;
  const int last_src = step_last_src(s);
  if (s.bias_mma) {
    // one [N x 16] chunk whose k = 0 column is the scaled bias: the "weight" matrix is the bias vector itself
    // (row stride 1, one true column)
    PackJob& j = P.jobs[P.n_jobs++];
    memset(&j, 0, sizeof(j));
    j.dst_off = off; j.mode = 0; j.n0 = 0; j.Ntrue = Ntrue; j.n_ksteps = 1; j.rows = s.N;
    j.w_off = b_off; j.ldw = 1; j.col0 = 0; j.k0 = 0; j.Ktrue = 1; j.scale = b_scale;
    off += (long long)s.N * 64;
  }
  for (int src = 0; src < 3; ++src) {
    const int nks = step_src_ksteps(s, src);
    if (nks == 0) continue;
    const bool wide = src != last_src && nh == 2;     // one [256 x 16] operand instead of two halves
    for (int h = 0; h < (wide ? 1 : nh); ++h) {
      PackJob& j = P.jobs[P.n_jobs++];
      memset(&j, 0, sizeof(j));
      j.dst_off = off; j.mode = mode; j.n0 = 128 * h; j.Ntrue = Ntrue; j.n_ksteps = nks; j.rows = wide ? 256 : 128;
      j.w_off = w_off; j.ldw = ldw;
      if (src == 0) { j.k0 = 0; j.Ktrue = seg_K; j.col0 = seg_col0; j.scale = seg_scale; }
      else { j.k0 = 128 * (src - 1); j.Ktrue = main_K; j.col0 = main_col0; j.scale = main_scale; }
      off += (long long)nks * j.rows * 64;
    }
  }
}

// bias_mma: the hidden softplus layers take their bias as a K step of the GEMM -- exactly the steps the flat forward
// epilogue handles (value pass, split operands; see fwd_bias_mma())
static void build_forward(const RenPlan& p, TcProgram& P, bool bias_mma) {
  memset(&P, 0, sizeof(P));
  const bool sp = p.f.hidden_activation == REN_HIDDEN_SOFTPLUS100;
  const float cs = sp ? LOG2E_100 : 1.f, kap = sp ? LN2_OVER_100 : 1.f;
  long long off = 0;
  int acc = 0, prev_acc = 0;
  auto add_layer = [&](const RenLayer& L, int epi, int N, float out_scale, float main_in_scale) -> TcStep& {
    TcStep& s = P.steps[P.n_steps];
    memset(&s, 0, sizeof(s));
    s.N = N; s.acc_col = acc; s.a_col = prev_acc; s.epi = epi; s.out_mat = L.mat; s.act_mat = -1;
    s.write_a = epi != EPI_F_LAST; s.bias_idx = P.n_steps;
    s.bias_mma = (bias_mma && epi == EPI_F_HIDDEN && sp && N == 256 && L.mat >= 0) ? 1 : 0;
    int seg_col0 = 0, seg_K = 0, main_K = 0;
    if (L.src0 == 1) { s.a_k = 0; s.seg = 1; seg_col0 = 0; seg_K = L.K0; }          // first layer: encoding only
    else {
      s.a_k = L.K0; main_K = L.K0;
      if (L.K1 > 0) { s.seg = L.src1; seg_col0 = L.K0; seg_K = L.K1; }
    }
    const float main_scale = (out_scale * main_in_scale > 0.999f && out_scale * main_in_scale < 1.001f)
                                 ? 1.f : out_scale * main_in_scale;     // 100 log2e * ln2/100 == 1 exactly
    add_step_pack(P, s, off, 0, L.w_off, L.K0 + L.K1, L.N, 0, main_K, main_scale, seg_col0, seg_K, out_scale,
                  L.b_off, out_scale);
    add_vec(P, VEC_BIAS + P.n_steps * 256, 256, L.N, L.b_off, out_scale);
    ++P.n_steps;
    prev_acc = acc;
    acc = 256 - acc;
    return s;
  };
  for (int i = 0; i < p.D; ++i) {
    TcStep& s = add_layer(p.trunk[i], EPI_F_HIDDEN, 256, cs, 1.f / cs);
    if (i == p.D - 1) s.sig_head = 1;
  }
  add_layer(p.bott, EPI_F_LINEAR, 256, 1.f, 1.f).acc_scale = kap;   // bott = kap * (W_b g) + b_b
  add_layer(p.cond[0], EPI_F_LAST, 128, cs, 1.f);
  add_vec(P, VEC_WSIG, 256, p.W, p.sig_w, kap);
  for (int c = 0; c < 3; ++c) add_vec(P, VEC_WO + 128 * c, 128, c < p.C ? p.WC : 0, p.rgb_w + (c < p.C ? c : 0) * p.WC, kap);
  add_vec(P, VEC_HB, 1, 1, p.sig_b, 1.f);
  add_vec(P, VEC_HB + 1, 3, p.C, p.rgb_b, 1.f);
  P.pack_bytes = off;
}

static void build_dgrad(const RenPlan& p, TcProgram& P) {
  memset(&P, 0, sizeof(P));
  long long off = 0;
  int acc = 256, prev_acc = 0;     // the prologue writes the first A operand (Delta of the condition layer) at column 0
  auto add = [&](int a_k, int epi, int out_mat, int act_mat, int write_a, int rank1, int w_off, int ldw) {
    TcStep& s = P.steps[P.n_steps++];
    memset(&s, 0, sizeof(s));
    s.N = 256; s.acc_col = acc; s.a_col = prev_acc; s.a_k = a_k; s.seg = 0; s.epi = epi; s.out_mat = out_mat;
    s.act_mat = act_mat; s.write_a = write_a; s.bias_idx = -1; s.rank1 = rank1;
    add_step_pack(P, s, off, 1, w_off, ldw, p.W, 0, a_k, 1.f, 0, 0, 1.f);
    prev_acc = acc;
    acc = 256 - acc;
  };
  // t1: d bott = d g_pre W_r0[:, :W]                    (bottleneck is linear)
  add(p.WC, EPI_D_LINEAR, p.bott.mat, -1, 1, 0, p.cond[0].w_off, p.cond[0].K0 + p.cond[0].K1);
  // t2: d a_{D-1} = d bott W_b + (d sigma_raw) w_sigma, times act'(a_{D-1})
  {
    const int last = p.trunk[p.D - 1].mat;
    add(p.W, EPI_D_ACT, last, last, p.D > 1 ? 1 : 0, 1, p.bott.w_off, p.bott.K0 + p.bott.K1);
  }
  // trunk layers D-1 .. 1: d a_{i-1} = Delta_i W_i[:, :W], times act'(a_{i-1})
  for (int i = p.D - 1; i >= 1; --i) {
    const RenLayer& L = p.trunk[i];
    add(p.W, EPI_D_ACT, p.trunk[i - 1].mat, p.trunk[i - 1].mat, i > 1 ? 1 : 0, 0, L.w_off, L.K0 + L.K1);
  }
  add_vec(P, VEC_WSIG, 256, p.W, p.sig_w, 1.f);
  for (int c = 0; c < 3; ++c) add_vec(P, VEC_WO + 128 * c, 128, c < p.C ? p.WC : 0, p.rgb_w + (c < p.C ? c : 0) * p.WC, 1.f);
  P.pack_bytes = off;
}

// Flatten the steps into ring stages, in exactly the order the weights were packed:
// step -> K source (segment, TMEM K-half 0, TMEM K-half 1) -> N-half (unless wide) -> stage.
__host__ __device__ inline int build_ops_impl(const TcStep* steps, int n_steps, int mode, StageOp* ops) {
  int n_ops = 0;
  for (int g = 0; g < n_steps; ++g) {
    const TcStep& st = steps[g];
    const int nh = st.N >> 7;
    const int last_src = step_last_src(st);
    long long off = st.pack_off;
    bool fresh = true;
    int moved_wait_src = -1;       // source whose first-op wait the bias op already carries
    if (mode == 0 && st.bias_mma) {
      if (n_ops >= TC_MAX_OPS) return -1;
      StageOp& o = ops[n_ops++];
      o.pad0 = 0; o.pad1 = 0;
      o.pack_off = (uint32_t)off; o.bytes = (uint32_t)(st.N * 64);
      off += st.N * 64;
      o.d_col = (uint32_t)st.acc_col; o.a_addr = 0; o.a_lo_delta = 0;
      uint32_t flags = 1u | (nh == 2 ? 8u : 0u) | 16u | 32u | OP_ONES;
      // the accumulator region may only be overwritten once the step's first source could be issued (for the
      // second step of a tile that is what orders it after the previous tile's last epilogue)
      for (int src = 0; src < 3; ++src) {
        if (step_src_ksteps(st, src) == 0) continue;
        if (src > 0) flags |= (uint32_t)src << 6;
        else if (g == 0) flags |= 3u << 6;
        moved_wait_src = src;
        break;
      }
      o.flags = flags;
      fresh = false;
    }
    for (int src = 0; src < 3; ++src) {
      const int nks = step_src_ksteps(st, src);
      if (nks == 0) continue;
      const bool wide = src != last_src && nh == 2;
      const int per = wide ? 2 : 4, reps = wide ? 1 : nh;
      const long long chunk = wide ? 2 * SM_CHUNK_BYTES : SM_CHUNK_BYTES;
      for (int h = 0; h < reps; ++h) {
        for (int ks0 = 0; ks0 < nks; ks0 += per) {
          if (n_ops >= TC_MAX_OPS) return -1;
          StageOp& o = ops[n_ops++];
          o.pad0 = 0; o.pad1 = 0;
          const int cnt = nks - ks0 < per ? nks - ks0 : per;
          o.pack_off = (uint32_t)off; o.bytes = (uint32_t)(cnt * chunk);
          off += cnt * chunk;
          o.d_col = (uint32_t)(st.acc_col + 128 * h);
          uint32_t flags = (uint32_t)cnt | (wide ? 8u : 0u);
          if (src == 0) {
            flags |= 16u;
            const int base = st.seg == 1 ? SM_P_HI : SM_D_HI;
            o.a_addr = (uint32_t)((base + ks0 * 4096) >> 4);
            o.a_lo_delta = (uint32_t)((st.seg == 1 ? (SM_P_LO - SM_P_HI) : (SM_D_LO - SM_D_HI)) >> 4);
          } else {
            o.a_addr = (uint32_t)(st.a_col + 128 * (src - 1) + 32 * (ks0 >> 1));
            o.a_lo_delta = 16u;
          }
          if (fresh && ks0 == 0) flags |= 32u;
          if (ks0 == 0 && h == 0 && src != moved_wait_src) {
            if (src > 0) flags |= (uint32_t)src << 6;
            else if (mode == 0 && g == 0) flags |= 3u << 6;          // first stage of the tile: encodings ready
          }
          if (src == last_src && ks0 + per >= nks) flags |= (uint32_t)(h + 1) << 8;
          o.flags = flags;
        }
      }
      fresh = false;
    }
  }
  return n_ops;
}
struct ProgArgs {
  TcStep steps[TC_MAX_STEPS];
  int n_steps, mode;
  StageOp* prog;
};
// the device writes its own copy of the stage program (no host->device copy on the launch path)
extern "C" __global__ void ren_tc_prog(const ProgArgs a) {
  if (blockIdx.x == 0 && threadIdx.x == 0) build_ops_impl(a.steps, a.n_steps, a.mode, a.prog);
}
static int build_ops(TcProgram& P, int mode) {
  P.n_ops = build_ops_impl(P.steps, P.n_steps, mode, P.ops);
  return P.n_ops < 0 ? -1 : 0;
}

int ren_tc_supported(const RenPlan& p) {
  return p.arch == REN_ARCH_MLP && p.W == 256 && p.WC == 128 && p.CD == 1 && p.hid_K1 == 0 && p.f.pos_deg == 10 && p.f.dir_deg == 4 &&
         p.D >= 1 && p.D + 2 <= 12 && p.C <= 3 && p.trunk[0].K1 == 0;
}

static size_t program_bytes(const TcProgram& P) {
  return (size_t)P.pack_bytes + (size_t)VEC_FLOATS * 4 + sizeof(StageOp) * TC_MAX_OPS + 2048;
}

size_t ren_tc_fwd_workspace_bytes(const RenPlan& p) {
  if (!ren_tc_supported(p)) return 0;
  TcProgram P;
  build_forward(p, P, REN_BIAS_MMA != 0);      // the largest variant of the forward program
  return program_bytes(P);
}
size_t ren_tc_bwd_pack_bytes(const RenPlan& p) {
  if (!ren_tc_supported(p)) return 0;
  TcProgram P;
  build_dgrad(p, P);
  return program_bytes(P);
}

static int run_pack(const TcProgram& P, int mode, const float* params, unsigned char* pack, float* vec,
                    StageOp* prog, cudaStream_t st) {
  {
    ProgArgs ga;
    for (int i = 0; i < P.n_steps; ++i) ga.steps[i] = P.steps[i];
    ga.n_steps = P.n_steps; ga.mode = mode; ga.prog = prog;
    ren_prof_begin(REN_K_PACK, st);
    ren_tc_prog<<<1, 32, 0, st>>>(ga);
    ren_prof_end(REN_K_PACK, st);
  }
  PackArgs pa;
  static_assert(sizeof(PackArgs) <= 4000, "pack parameter block too large");
  if (P.n_jobs > TC_MAX_PACK) { ren_set_error("tc pack: too many jobs (%d)", P.n_jobs); return REN_E_INVALID; }
  for (int i = 0; i < P.n_jobs; ++i) pa.jobs[i] = P.jobs[i];
  pa.n_jobs = P.n_jobs; pa.params = params; pa.pack = pack; pa.vec = vec;
  ren_prof_begin(REN_K_PACK, st);
  ren_tc_pack<<<dim3(16, P.n_jobs), 256, 0, st>>>(pa);
  ren_prof_end(REN_K_PACK, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc pack launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

template <int MODE, bool STASH, int VAR>
static int launch_chain(const TcArgs& a, long long n_tiles, int kind, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (grid < 1) return REN_OK;
  cudaError_t e = cudaFuncSetAttribute(ren_tc_chain<MODE, STASH, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
  if (e != cudaSuccess) { ren_set_error("tc chain smem attr: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  ren_prof_begin(kind, st);
  ren_tc_chain<MODE, STASH, VAR><<<grid, MODE == 0 ? TC_THREADS_FWD : TC_THREADS_BWD, SM_TOTAL, st>>>(a);
  ren_prof_end(kind, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("tc chain launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

static int carve_pack(const TcProgram& P, void* ws, size_t ws_bytes, unsigned char** pack, float** vec,
                      StageOp** prog) {
  unsigned char* base = (unsigned char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const size_t need = (size_t)(base - (unsigned char*)ws) + (size_t)P.pack_bytes + (size_t)VEC_FLOATS * 4 +
                      sizeof(StageOp) * TC_MAX_OPS + 64;
  if (ws == nullptr || need > ws_bytes) {
    ren_set_error("workspace too small for packed weights (%zu > %zu)", need, ws_bytes);
    return REN_E_WORKSPACE;
  }
  *pack = base;
  *vec = (float*)(base + ((P.pack_bytes + 15) & ~15LL));
  *prog = (StageOp*)((unsigned char*)*vec + (((size_t)VEC_FLOATS * 4 + 15) & ~(size_t)15));
  return REN_OK;
}

// operand precision, activation and the timing switches of a launch
static void fill_modes(TcArgs& a, const RenPlan& p) {
  a.split3 = p.r.engine == REN_ENGINE_TC_F16 ? 0 : (p.r.engine == REN_ENGINE_TC_F16A ? 2 : 1);
  if (const char* sp = getenv("REN_TC_SPLIT")) a.split3 = atoi(sp);      // precision experiments
  a.softplus = p.f.hidden_activation == REN_HIDDEN_SOFTPLUS100 ? 1 : 0;
  const char* dbg = getenv("REN_TC_DBG");
  a.dbg = dbg ? atoi(dbg) : 0;
}
static void fill_args(TcArgs& a, const RenPlan& p, const TcProgram& P) {
  static_assert(sizeof(TcArgs) <= 4000, "kernel parameter block too large");
  memset(&a, 0, sizeof(a));
  a.p = p;
  for (int i = 0; i < P.n_steps; ++i) a.steps[i] = P.steps[i];
  a.n_steps = P.n_steps;
  fill_modes(a, p);
}

// The forward value pass with split operands runs the flat epilogue on its hidden softplus layers; with
// REN_BIAS_MMA those layers take their bias through the GEMM (the tangent pass carries no biases; the reduced
// precision modes and the timing switches keep the generic epilogue, which adds the bias itself)
static bool fwd_bias_mma(const TcArgs& a, bool tan) {
  return REN_BIAS_MMA != 0 && !tan && a.split3 == 1 && a.softplus && !(a.dbg & 1);
}

// o_dot / d_dot / stash_val != NULL: the tangent pass of the tangent render (stash = its own stash, raw = raw_dot)
int ren_tc_launch_fwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                      const float* o_dot, const float* d_dot, const void* stash_val,
                      const float* jitter, const float* params, void* stash, float* raw, const unsigned char* tile_live,
                      const int* sample_idx, long long n_samples, void* ws, size_t ws_bytes, cudaStream_t st) {
  TcProgram P;
  TcArgs a;
  fill_modes(a, p);            // operand precision / timing switches decide which epilogue the hidden layers take
  build_forward(p, P, fwd_bias_mma(a, o_dot != nullptr));
  unsigned char* pack; float* vec; StageOp* prog;
  if (build_ops(P, 0)) { ren_set_error("tc forward: stage program too long"); return REN_E_UNSUPPORTED; }
  int rc = carve_pack(P, ws, ws_bytes, &pack, &vec, &prog);
  if (rc) return rc;
  rc = run_pack(P, 0, params, pack, vec, prog, st);
  if (rc) return rc;
  fill_args(a, p, P);
  a.n_rays = n_rays; a.K = K; a.rays_o = o; a.rays_d = d; a.jitter = jitter;
  a.pack = pack; a.vec = vec; a.prog = prog; a.n_ops = P.n_ops; a.stash = (unsigned char*)stash; a.raw = raw;
  a.tan = o_dot != nullptr ? 1 : 0; a.rays_o_dot = o_dot; a.rays_d_dot = d_dot;
  a.stash_val = (const unsigned char*)stash_val; a.tile_live = tile_live;
  a.sample_idx = sample_idx; a.n_samples = n_samples;
  const long long n_tiles = ((sample_idx ? n_samples : n_rays * K) + TC_M - 1) / TC_M;
  if (a.tan) return launch_chain<0, true, 1>(a, n_tiles, REN_K_FIELD_FWD, st);     // the tangent pass always stashes
  return stash ? launch_chain<0, true, 0>(a, n_tiles, REN_K_FIELD_FWD, st)
               : launch_chain<0, false, 0>(a, n_tiles, REN_K_FIELD_FWD, st);
}

// pass 0: plain dgrad; 1: tangent-adjoint chain of the tangent render (reads stash_dot, writes extra);
// 2: its value-adjoint chain (reads extra)
int ren_tc_launch_dgrad(const RenPlan& p, long long S, const float* params, const void* stash,
                        const float* graw, const unsigned int* gmax, void* delta, int pass, const void* stash_dot,
                        void* extra, const unsigned char* tile_live, void* ws, size_t ws_bytes, cudaStream_t st) {
  TcProgram P;
  build_dgrad(p, P);
  unsigned char* pack; float* vec; StageOp* prog;
  if (build_ops(P, 1)) { ren_set_error("tc dgrad: stage program too long"); return REN_E_UNSUPPORTED; }
  int rc = carve_pack(P, ws, ws_bytes, &pack, &vec, &prog);
  if (rc) return rc;
  rc = run_pack(P, 1, params, pack, vec, prog, st);
  if (rc) return rc;
  TcArgs a;
  fill_args(a, p, P);
  a.n_rays = S; a.K = 1;            // dgrad only needs the flat sample count
  a.pack = pack; a.vec = vec; a.prog = prog; a.n_ops = P.n_ops;
  a.stash = (unsigned char*)stash; a.delta = (unsigned char*)delta; a.graw = graw; a.gmax = gmax;
  a.pass = pass; a.stash_dot = (const unsigned char*)stash_dot; a.extra = (unsigned char*)extra;
  a.tile_live = tile_live;
  const long long n_tiles = (S + TC_M - 1) / TC_M;
  if (pass == 1) return launch_chain<1, false, 1>(a, n_tiles, REN_K_FIELD_DGRAD, st);
  if (pass == 2) return launch_chain<1, false, 2>(a, n_tiles, REN_K_FIELD_DGRAD, st);
  return launch_chain<1, false, 0>(a, n_tiles, REN_K_FIELD_DGRAD, st);
}

// ------------------------------------------------------------------ program dump (tests; no GPU needed)
// The tile program of the chain kernel as the host builds it: steps (12 ints each: N, acc_col, a_col, a_k, seg,
// epi, out_mat, act_mat, write_a, bias_idx, rank1 | sig_head << 8, bias_mma), stage ops (8 words each, the
// device layout) and weight-pack jobs (12 int64 each: dst_off, mode, n0, Ntrue, n_ksteps, k0, Ktrue, rows, w_off,
// ldw, col0, scale bits).  mode 0 forward (bias_mma: 0 / 1), 1 dgrad.  Not part of the C ABI in include/: internal.
extern "C" int ren_tc_debug_program(const RenFieldDesc* f, const RenRenderDesc* r, int32_t mode, int32_t bias_mma,
                                    int32_t* steps, int32_t* n_steps, uint32_t* ops, int32_t* n_ops, int64_t* jobs,
                                    int32_t* n_jobs, int64_t* pack_bytes) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  if (!ren_tc_supported(p)) { ren_set_error("tc engine does not support this network"); return REN_E_UNSUPPORTED; }
  TcProgram P;
  if (mode == 0) build_forward(p, P, bias_mma != 0); else build_dgrad(p, P);
  if (build_ops(P, mode)) { ren_set_error("stage program too long"); return REN_E_UNSUPPORTED; }
  if (P.n_jobs > TC_MAX_PACK) { ren_set_error("too many pack jobs (%d)", P.n_jobs); return REN_E_INVALID; }
  for (int i = 0; i < P.n_steps; ++i) {
    const TcStep& s = P.steps[i];
    const int v[12] = {s.N, s.acc_col, s.a_col, s.a_k, s.seg, s.epi, s.out_mat, s.act_mat, s.write_a, s.bias_idx,
                       s.rank1 | (s.sig_head << 8), s.bias_mma};
    for (int k = 0; k < 12; ++k) steps[12 * i + k] = v[k];
  }
  *n_steps = P.n_steps;
  memcpy(ops, P.ops, sizeof(StageOp) * P.n_ops);
  *n_ops = P.n_ops;
  for (int i = 0; i < P.n_jobs; ++i) {
    const PackJob& j = P.jobs[i];
    uint32_t sb;
    memcpy(&sb, &j.scale, 4);
    const int64_t v[12] = {j.dst_off, j.mode, j.n0, j.Ntrue, j.n_ksteps, j.k0, j.Ktrue, j.rows, j.w_off, j.ldw,
                           j.col0, (int64_t)sb};
    for (int k = 0; k < 12; ++k) jobs[12 * i + k] = v[k];
  }
  *n_jobs = P.n_jobs;
  *pack_bytes = P.pack_bytes;
  return REN_OK;
}
