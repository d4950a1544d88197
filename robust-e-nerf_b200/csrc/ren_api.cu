// ren_api.cu -- C ABI of libren_b200.so (include/ren_b200.h): plan building, argument
// validation, workspace carving and kernel sequencing.  No allocation, no host sync.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "ren_tc.cuh"

// engines
int ren_simt_launch_fwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                        const float* o_dot, const float* d_dot, const float* jitter, const float* params,
                        float* stash, float* raw, const RenNgp* ngp, cudaStream_t st);
int ren_simt_launch_dgrad(const RenPlan& p, long long S, const float* params, const float* stash,
                          const float* graw, float* delta, int tangent, cudaStream_t st);
int ren_simt_launch_wgrad(const RenPlan& p, long long S, const float* stash, const float* delta,
                          const float* graw, float* g_params, int tangent, cudaStream_t st);
int ren_launch_composite_tan(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                             const float* jitter, const int* ray_offsets, const float* raw, const float* raw_dot,
                             const float* bkgd,
                             float* radiance, float* opacity, float* depth, float* radiance_dot, float* opacity_dot,
                             int* n_kept, cudaStream_t st);
int ren_launch_composite_tan_bwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                                 const float* jitter, const int* ray_offsets, const float* raw, const float* raw_dot,
                                 const float* bkgd,
                                 const float* g_radiance, const float* g_opacity, const float* g_depth,
                                 const float* g_radiance_dot, const float* g_opacity_dot, float* graw,
                                 float* graw_dot, float* g_bkgd, unsigned int* gmax, cudaStream_t st);
int ren_launch_raygrad(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                       const float* o_dot, const float* d_dot, const float* jitter, const float* params,
                       const void* delta, const void* delta_dot, int fp16, const unsigned int* gmax,
                       const unsigned char* tile_live, const int* sample_idx, long long n_samples, float* g_rays,
                       const RenNgp* ngp, cudaStream_t st);
int ren_launch_composite(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                         const float* jitter, const int* ray_offsets, const float* raw, const float* bkgd,
                         float* radiance, float* opacity, float* depth, int* n_kept, cudaStream_t st);
int ren_launch_march_pack(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                          const float* jitter, int* counts, const int* ray_offsets, int* sample_idx,
                          cudaStream_t st);
int ren_launch_composite_bwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                             const float* jitter, const int* ray_offsets, const float* raw, const float* bkgd,
                             const float* g_radiance, const float* g_opacity, const float* g_depth,
                             float* graw, float* g_bkgd, unsigned int* gmax, cudaStream_t st);
// arch ngp (ren_ngp.cu)
int ren_build_grid(const RenFieldDesc* f, RenGrid* g, long long* n_entries);
int ren_ngp_pack(const RenPlan& p, const float* flat, float* kparams, cudaStream_t st);
int ren_ngp_unpack_grads(const RenPlan& p, const float* g_kparams, float* g_flat, cudaStream_t st);
// tcgen05 engine (ren_tc.cu)
int ren_tc_supported(const RenPlan& p);
size_t ren_tc_fwd_workspace_bytes(const RenPlan& p);
size_t ren_tc_bwd_pack_bytes(const RenPlan& p);
int ren_tc_launch_fwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                      const float* o_dot, const float* d_dot, const void* stash_val,
                      const float* jitter, const float* params, void* stash, float* raw, const unsigned char* tile_live,
                      const int* sample_idx, long long n_samples, void* ws, size_t ws_bytes, cudaStream_t st);
int ren_tc_launch_dgrad(const RenPlan& p, long long S, const float* params, const void* stash,
                        const float* graw, const unsigned int* gmax, void* delta, int pass, const void* stash_dot,
                        void* extra, const unsigned char* tile_live, void* ws, size_t ws_bytes, cudaStream_t st);
int ren_tc_launch_wgrad(const RenPlan& p, long long S, const void* stash, const void* delta,
                        const float* graw, const unsigned int* gmax, float* g_params, int with_bias,
                        const unsigned char* tile_live, cudaStream_t st);
int ren_launch_tile_live(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                         const float* jitter, unsigned char* live, cudaStream_t st);

static thread_local char g_err[512] = "";

void ren_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- launch accounting
#include <atomic>
#include <mutex>
#include <vector>
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_timing{0};
static std::mutex g_prof_mu;
struct ProfPair { cudaEvent_t a, b; };
static std::vector<ProfPair> g_prof[REN_K_COUNT];
static thread_local cudaEvent_t g_prof_open = nullptr;

void ren_prof_begin(int kind, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_timing.load(std::memory_order_relaxed)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_prof_open = e;
  (void)kind;
}
void ren_prof_count() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void ren_prof_end(int kind, cudaStream_t st) {
  if (!g_timing.load(std::memory_order_relaxed) || g_prof_open == nullptr) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof[kind].push_back({g_prof_open, e});
  g_prof_open = nullptr;
}

extern "C" int64_t ren_launch_count(void) { return g_launches.load(); }
extern "C" int ren_timing_enable(int32_t on) { g_timing.store(on ? 1 : 0); return REN_OK; }
extern "C" int ren_timing_collect(int32_t kind, double* total_ms, int64_t* count) {
  if (kind < 0 || kind >= REN_K_COUNT || !total_ms || !count) { ren_set_error("ren_timing_collect: bad arguments"); return REN_E_INVALID; }
  std::vector<ProfPair> v;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    v.swap(g_prof[kind]);
  }
  double tot = 0;
  for (auto& p : v) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    tot += ms;
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  *total_ms = tot;
  *count = (int64_t)v.size();
  return REN_OK;
}

extern "C" int ren_version(void) { return REN_ABI_VERSION; }
extern "C" const char* ren_last_error(void) { return g_err; }

static int validate_field(const RenFieldDesc* f) {
  if (f == nullptr) { ren_set_error("null RenFieldDesc"); return REN_E_INVALID; }
  if (f->net_depth < 1 || f->net_depth > REN_MAX_TRUNK) { ren_set_error("net_depth %d outside [1,%d]", f->net_depth, REN_MAX_TRUNK); return REN_E_INVALID; }
  if (f->net_width < 16 || f->net_width % 16) { ren_set_error("net_width %d must be a positive multiple of 16", f->net_width); return REN_E_INVALID; }
  if (f->net_depth_condition < 0 || f->net_depth_condition > REN_MAX_COND) { ren_set_error("net_depth_condition %d outside [0,%d]", f->net_depth_condition, REN_MAX_COND); return REN_E_INVALID; }
  if (f->net_depth_condition > 0 && (f->net_width_condition < 16 || f->net_width_condition % 16)) { ren_set_error("net_width_condition %d must be a positive multiple of 16", f->net_width_condition); return REN_E_INVALID; }
  if (f->radiance_dim < 1 || f->radiance_dim > 3) { ren_set_error("radiance_dim %d outside [1,3]", f->radiance_dim); return REN_E_INVALID; }
  if (f->arch != REN_ARCH_MLP && f->arch != REN_ARCH_NGP) { ren_set_error("arch id %d out of range", f->arch); return REN_E_INVALID; }
  if (f->arch == REN_ARCH_NGP) {
    if (f->net_depth_condition < 1) { ren_set_error("arch ngp: mlp_head needs at least one hidden layer"); return REN_E_UNSUPPORTED; }
    if (f->sh_degree < 1 || f->sh_degree > 4) { ren_set_error("arch ngp: sh degree %d outside [1,4]", f->sh_degree); return REN_E_UNSUPPORTED; }
    if (f->geo_feat_dim < 1 || f->geo_feat_dim > 63) { ren_set_error("arch ngp: geo_feat_dim %d outside [1,63]", f->geo_feat_dim); return REN_E_UNSUPPORTED; }
    RenGrid g;
    int grc = ren_build_grid(f, &g, nullptr);
    if (grc) return grc;
  } else
  if (f->pos_deg < 0 || f->pos_deg > 10 || f->dir_deg < 0 || f->dir_deg > 4) { ren_set_error("encoder degrees (%d,%d) outside [0,10]x[0,4]", f->pos_deg, f->dir_deg); return REN_E_INVALID; }
  if (f->skip_layer < 0) { ren_set_error("skip_layer must be >= 0 (0 = none)"); return REN_E_INVALID; }
  if (f->hidden_activation < 0 || f->hidden_activation > 1 || f->density_activation < 0 || f->density_activation > 2 ||
      f->radiance_activation < 0 || f->radiance_activation > 1 || f->contraction < 0 || f->contraction > 2) {
    ren_set_error("activation / contraction id out of range");
    return REN_E_INVALID;
  }
  for (int a = 0; a < 3; ++a)
    if (!(f->aabb[3 + a] > f->aabb[a])) { ren_set_error("aabb max must exceed min on every axis"); return REN_E_INVALID; }
  return REN_OK;
}

int ren_build_plan(const RenFieldDesc* f, const RenRenderDesc* r, RenPlan* p) {
  int rc = validate_field(f);
  if (rc) return rc;
  memset(p, 0, sizeof(*p));
  p->f = *f;
  if (r) p->r = *r;
  p->D = f->net_depth; p->CD = f->net_depth_condition;
  p->W = f->net_width; p->WC = f->net_depth_condition > 0 ? f->net_width_condition : f->net_width;
  p->C = f->radiance_dim;
  p->PE = 3 * (1 + 2 * f->pos_deg); p->DE = 3 * (1 + 2 * f->dir_deg);
  p->arch = f->arch; p->WB = p->W; p->geo = 0; p->n_grid = 0;
  const bool ngp = f->arch == REN_ARCH_NGP;
  const int skip = ngp ? 0 : f->skip_layer;
  if (ngp) {
    RenGrid g;
    long long entries = 0;
    rc = ren_build_grid(f, &g, &entries);
    if (rc) return rc;
    p->PE = g.n_levels * g.F;                 // grid features take the place of the positional encoding
    p->DE = f->sh_degree * f->sh_degree;      // spherical harmonics that of the direction encoding
    p->geo = f->geo_feat_dim;
    p->WB = (f->geo_feat_dim + 7) & ~7;
    p->n_grid = entries * g.F;
  }
  int off = 0, mat = 0;
  long long moff = 0;
  int in0 = p->PE, in1 = 0;  // input of the next trunk layer: main width, extra width
  bool first = true;
  for (int i = 0; i < p->D; ++i) {
    RenLayer& L = p->trunk[i];
    L.K0 = in0; L.K1 = in1; L.src0 = first ? 1 : 0; L.src1 = in1 > 0 ? 1 : 0;
    L.N = p->W; L.act = 1;
    L.w_off = off; off += L.N * (L.K0 + L.K1);
    L.b_off = off; off += L.N;
    L.mat = mat; p->mat_width[mat] = L.N; p->mat_off[mat] = moff; moff += L.N; ++mat;
    first = false;
    in0 = p->W;
    in1 = (skip > 0 && i % skip == 0 && i > 0) ? p->PE : 0;
  }
  p->hid_K0 = in0; p->hid_K1 = in1;
  p->sig_w = off; off += p->hid_K0 + p->hid_K1;
  p->sig_b = off; off += 1;
  {
    RenLayer& L = p->bott;
    L.K0 = p->hid_K0; L.K1 = p->hid_K1; L.src0 = 0; L.src1 = p->hid_K1 > 0 ? 1 : 0;
    L.N = p->WB; L.act = 0;
    L.w_off = off; off += L.N * (L.K0 + L.K1);
    L.b_off = off; off += L.N;
    L.mat = mat; p->mat_width[mat] = L.N; p->mat_off[mat] = moff; moff += L.N; ++mat;
  }
  int c0 = p->WB, c1 = p->DE;
  for (int j = 0; j < p->CD; ++j) {
    RenLayer& L = p->cond[j];
    L.K0 = c0; L.K1 = c1; L.src0 = 0; L.src1 = c1 > 0 ? 2 : 0;
    L.N = p->WC; L.act = 1;
    L.w_off = off; off += L.N * (L.K0 + L.K1);
    L.b_off = off; off += L.N;
    L.mat = mat; p->mat_width[mat] = L.N; p->mat_off[mat] = moff; moff += L.N; ++mat;
    c0 = p->WC; c1 = 0;
  }
  p->rgb_w = off; off += p->C * (c0 + c1);
  p->rgb_b = off; off += p->C;
  p->n_params = off;
  p->n_mats = mat;
  p->act_floats = moff;
  p->n_flat = off;
  if (ngp) {
    // the reference's own tensors: grid, mlp_base (hidden, output (1 + geo, W)), mlp_head (hidden, output (C, WC))
    long long n = p->n_grid;
    int in = p->PE;
    for (int i = 0; i < p->D; ++i) { n += (long long)p->W * in + p->W; in = p->W; }
    n += (long long)(1 + p->geo) * p->W + (1 + p->geo);
    in = p->DE + p->geo;
    for (int j = 0; j < p->CD; ++j) { n += (long long)p->WC * in + p->WC; in = p->WC; }
    n += (long long)p->C * p->WC + p->C;
    p->n_flat = n;
  }
  return REN_OK;
}

extern "C" int ren_param_count(const RenFieldDesc* f, int64_t* n_params) {
  RenPlan p;
  int rc = ren_build_plan(f, nullptr, &p);
  if (rc) return rc;
  if (n_params) *n_params = p.n_flat;
  return REN_OK;
}

extern "C" int ren_grid_param_count(const RenFieldDesc* f, int64_t* n_grid_params) {
  RenPlan p;
  int rc = ren_build_plan(f, nullptr, &p);
  if (rc) return rc;
  if (n_grid_params) *n_grid_params = p.n_grid;
  return REN_OK;
}

extern "C" int ren_param_layout(const RenFieldDesc* f, int32_t* n_layers, int64_t* w_off, int64_t* b_off,
                                int32_t* out_features, int32_t* in_features) {
  RenPlan p;
  int rc = ren_build_plan(f, nullptr, &p);
  if (rc) return rc;
  int n = 0;
  auto put = [&](long long wo, long long bo, int outf, int inf) {
    if (w_off) w_off[n] = wo;
    if (b_off) b_off[n] = bo;
    if (out_features) out_features[n] = outf;
    if (in_features) in_features[n] = inf;
    ++n;
  };
  if (p.arch == REN_ARCH_NGP) {
    long long o = p.n_grid;
    auto lin = [&](int outf, int inf) { put(o, o + (long long)outf * inf, outf, inf); o += (long long)outf * inf + outf; };
    int in = p.PE;
    for (int i = 0; i < p.D; ++i) { lin(p.W, in); in = p.W; }
    lin(1 + p.geo, p.W);
    in = p.DE + p.geo;
    for (int j = 0; j < p.CD; ++j) { lin(p.WC, in); in = p.WC; }
    lin(p.C, p.WC);
    if (n_layers) *n_layers = n;
    return REN_OK;
  }
  for (int i = 0; i < p.D; ++i) put(p.trunk[i].w_off, p.trunk[i].b_off, p.trunk[i].N, p.trunk[i].K0 + p.trunk[i].K1);
  put(p.sig_w, p.sig_b, 1, p.hid_K0 + p.hid_K1);
  put(p.bott.w_off, p.bott.b_off, p.bott.N, p.bott.K0 + p.bott.K1);
  for (int j = 0; j < p.CD; ++j) put(p.cond[j].w_off, p.cond[j].b_off, p.cond[j].N, p.cond[j].K0 + p.cond[j].K1);
  put(p.rgb_w, p.rgb_b, p.C, p.CD > 0 ? p.WC : p.W + p.DE);
  if (n_layers) *n_layers = n;
  return REN_OK;
}

static int validate_render(const RenRenderDesc* r) {
  if (r == nullptr) { ren_set_error("null RenRenderDesc"); return REN_E_INVALID; }
  if (!(r->render_step_size > 0.f)) { ren_set_error("render_step_size must be > 0"); return REN_E_INVALID; }
  if (r->cone_angle < 0.f) { ren_set_error("cone_angle must be >= 0"); return REN_E_INVALID; }
  if (r->engine < 0 || r->engine > 4) { ren_set_error("engine id out of range"); return REN_E_INVALID; }
  if (r->occ_binary != nullptr && (r->occ_res[0] < 1 || r->occ_res[1] < 1 || r->occ_res[2] < 1)) {
    ren_set_error("occupancy grid resolution must be positive");
    return REN_E_INVALID;
  }
  return REN_OK;
}
// nerfacc skips unoccupied space by voxel traversal for the `aabb` grid type: sample positions leave the
// fixed-step lattice the dense layout assumes
// Occupancy grid: every contraction is marched on the fixed-step lattice with a per-step cell test.  For `aabb`
// nerfacc skips empty voxels by stepping the midpoint in whole dt units until it leaves the voxel
// (advance_to_next_voxel), so its surviving samples are exactly the lattice steps whose midpoint cell is occupied.
static int check_occupancy(const RenFieldDesc* f, const RenRenderDesc* r) {
  (void)f;
  if (r->occ_binary != nullptr && (r->occ_res[0] < 1 || r->occ_res[1] < 1 || r->occ_res[2] < 1)) {
    ren_set_error("occupancy grid: bad resolution (%d, %d, %d)", r->occ_res[0], r->occ_res[1], r->occ_res[2]);
    return REN_E_INVALID;
  }
  return REN_OK;
}

extern "C" int ren_samples_per_ray(const RenFieldDesc* f, const RenRenderDesc* r, int32_t* K) {
  int rc = validate_field(f);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  const int cap = 1 << 20;
  double best = 1e30;
  if (f->contraction == REN_CONTRACT_AABB) {
    // marching starts at the box entry: the chord bounds the count (cone steps are only longer)
    double d2 = 0;
    for (int a = 0; a < 3; ++a) { double e = (double)f->aabb[3 + a] - f->aabb[a]; d2 += e * e; }
    best = sqrt(d2) / r->render_step_size + 2.0;
  }
  if (r->has_far) {
    // exact fp32 replay of the device march for the un-jittered ray starting at the near
    // plane: fl() is monotone, so any jitter / later entry only yields fewer samples.
    const float t_max = r->far_plane;
    volatile float t0 = r->has_near ? (r->near_plane > 0.f ? r->near_plane : 0.f) : 0.f;
    int n = 0;
    for (;;) {
      volatile float dt = t0 * r->cone_angle;
      if (dt < r->render_step_size) dt = r->render_step_size;
      if (dt > 1e10f) dt = 1e10f;
      volatile float t1 = t0 + dt;
      volatile float sum = t0 + t1;
      volatile float mid = sum * 0.5f;
      if (!(mid < t_max) || n >= cap) break;
      ++n;
      t0 = t1;
    }
    if ((double)n < best) best = n;
  }
  if (best > 1e29) {
    ren_set_error("unbounded marching (no far plane, non-aabb contraction) needs the occupancy-grid path");
    return REN_E_UNSUPPORTED;
  }
  if (best >= cap) { ren_set_error("dense sample bound %g too large", best); return REN_E_UNSUPPORTED; }
  *K = best < 1.0 ? 1 : (int32_t)best;
  return REN_OK;
}

// AUTO resolves to the tcgen05 engine whenever the network shape is covered.
static int pick_engine(const RenPlan& p) {
  int e = p.r.engine;
  if (e == REN_ENGINE_AUTO) e = ren_tc_supported(p) ? REN_ENGINE_TC_F16X3 : REN_ENGINE_SIMT_FP32;
  return e;
}
static int check_engine(const RenPlan& p, int engine) {
  if (engine != REN_ENGINE_SIMT_FP32 && !ren_tc_supported(p)) {
    ren_set_error("tcgen05 engine covers net_width 256 / net_width_condition 128 / 1 condition layer / "
                  "pos 10 / dir 4 encoders only; use engine simt_fp32 or auto");
    return REN_E_UNSUPPORTED;
  }
  return REN_OK;
}
static inline void* align256(void* p) {
  return (void*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
}

// arch ngp: the kernels run on a kernel-side copy of the MLP parameters (ren_ngp.cu); both it and its gradient
// live at the head of the caller's workspace
struct NgpCtx {
  bool on;
  RenNgp ngp;
  const float* kparams;     // what the field kernels read (== params for arch mlp)
  float* g_kparams;         // backward: kernel-side gradient (== g_params for arch mlp)
};
static size_t ngp_ws_bytes(const RenPlan& p) {
  return p.arch == REN_ARCH_NGP ? 2 * ((sizeof(float) * (size_t)p.n_params + 255) & ~(size_t)255) + 512 : 0;
}
static int ngp_prepare(const RenFieldDesc* f, const RenPlan& p, const float* params, float* g_params,
                       void* workspace, size_t workspace_bytes, char** ws, size_t* ws_left, bool backward,
                       cudaStream_t st, NgpCtx* c) {
  c->on = p.arch == REN_ARCH_NGP;
  c->kparams = params;
  c->g_kparams = g_params;
  if (!c->on) return REN_OK;
  const size_t one = (sizeof(float) * (size_t)p.n_params + 255) & ~(size_t)255;
  if (workspace == nullptr || *ws_left < 2 * one + 256) { ren_set_error("workspace too small for the ngp parameter copy"); return REN_E_WORKSPACE; }
  float* kp = (float*)*ws;
  float* gkp = (float*)(*ws + one);
  *ws = (char*)align256(*ws + 2 * one);
  *ws_left = workspace_bytes - (size_t)(*ws - (char*)workspace);
  int rc = ren_build_grid(f, &c->ngp.grid, nullptr);
  if (rc) return rc;
  c->ngp.table = params;
  c->ngp.g_table = g_params;
  rc = ren_ngp_pack(p, params, kp, st);
  if (rc) return rc;
  c->kparams = kp;
  if (backward) {
    if (cudaMemsetAsync(gkp, 0, sizeof(float) * (size_t)p.n_params, st) != cudaSuccess) { ren_set_error("ngp gradient memset failed"); return REN_E_CUDA; }
    c->g_kparams = gkp;
  }
  return REN_OK;
}

static size_t stash_bytes_for(const RenPlan& p, int engine, long long S) {
  return engine == REN_ENGINE_SIMT_FP32 ? sizeof(float) * (size_t)ren_stash_floats(p, S) : tcs_bytes(p, S);
}

extern "C" int ren_stash_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                               size_t* bytes) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  if (n_rays < 0 || K < 1) { ren_set_error("bad n_rays/K"); return REN_E_INVALID; }
  const int engine = pick_engine(p);
  rc = check_engine(p, engine);
  if (rc) return rc;
  *bytes = stash_bytes_for(p, engine, (long long)n_rays * K);
  return REN_OK;
}

// workspace:
//   forward : raw [S][1+C] when there is no stash to hold it, + packed weights (tcgen05)
//   backward: delta matrices + graw [S][1+C] (+ packed transposed weights)
extern "C" int ren_workspace_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                   int32_t has_stash, int32_t for_backward, size_t* bytes) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  if (n_rays < 0 || K < 1) { ren_set_error("bad n_rays/K"); return REN_E_INVALID; }
  const long long S = (long long)n_rays * K;
  const int engine = pick_engine(p);
  rc = check_engine(p, engine);
  if (rc) return rc;
  size_t n;
  if (engine == REN_ENGINE_SIMT_FP32) {
    n = for_backward ? sizeof(float) * (size_t)ren_delta_floats(p, S)
                     : (has_stash ? 0 : sizeof(float) * (size_t)((1 + p.C) * ren_pad128(S)));
  } else {
    n = for_backward ? tcd_bytes(p, S) + ren_tc_bwd_pack_bytes(p)
                     : (has_stash ? 0 : sizeof(float) * (size_t)((1 + p.C) * ren_pad128(S)) +
                                            (size_t)(ren_pad128(S) >> 7) + 512) +
                           ren_tc_fwd_workspace_bytes(p);
  }
  *bytes = n + 2048 + ngp_ws_bytes(p);
  return REN_OK;
}

// packed samples are a tcgen05-engine layout; returns the sample count the field kernels run over
static int check_packed(const RenPacked* packed, int engine, int64_t n_rays, int32_t K, long long* S) {
  *S = (long long)n_rays * K;
  if (packed == nullptr) return REN_OK;
  if (engine == REN_ENGINE_SIMT_FP32) { ren_set_error("packed samples need a tcgen05 engine"); return REN_E_UNSUPPORTED; }
  if (packed->n_samples < 0 || packed->n_samples > *S || (packed->n_samples > 0 && !packed->sample_idx) ||
      !packed->ray_offsets) {
    ren_set_error("bad RenPacked");
    return REN_E_INVALID;
  }
  *S = packed->n_samples;
  return REN_OK;
}

extern "C" int ren_render_forward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                  const float* rays_o, const float* rays_d, const float* jitter,
                                  const float* params, const float* bkgd, float* radiance, float* opacity,
                                  float* depth, int32_t* n_kept, const RenPacked* packed, void* stash,
                                  size_t stash_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  if (n_rays < 0 || K < 1) { ren_set_error("bad n_rays/K"); return REN_E_INVALID; }
  if (n_rays == 0) return REN_OK;
  if (!rays_o || !rays_d || !params || !radiance || !opacity || !depth || !n_kept) {
    ren_set_error("ren_render_forward: null pointer argument");
    return REN_E_INVALID;
  }
  if (r->stratified && jitter == nullptr) { ren_set_error("stratified marching needs a jitter buffer"); return REN_E_INVALID; }
  rc = check_occupancy(f, r);
  if (rc) return rc;
  const int engine = pick_engine(p);
  rc = check_engine(p, engine);
  if (rc) return rc;
  p.r.engine = engine;
  long long S;
  rc = check_packed(packed, engine, n_rays, K, &S);
  if (rc) return rc;
  const int* ray_offsets = packed ? packed->ray_offsets : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  float* raw;
  char* ws = (char*)align256(workspace);
  size_t ws_left = workspace ? workspace_bytes - (size_t)(ws - (char*)workspace) : 0;
  NgpCtx nc;
  rc = ngp_prepare(f, p, params, nullptr, workspace, workspace_bytes, &ws, &ws_left, false, st, &nc);
  if (rc) return rc;
  if (stash) {
    if (stash_bytes < stash_bytes_for(p, engine, S)) { ren_set_error("stash too small"); return REN_E_WORKSPACE; }
    raw = engine == REN_ENGINE_SIMT_FP32 ? (float*)stash + ren_stash_raw(p, S)
                                         : (float*)((char*)stash + tcs_raw(p, S));
  } else {
    const size_t need = sizeof(float) * (size_t)((1 + p.C) * ren_pad128(S));
    if (!workspace || ws_left < need) { ren_set_error("workspace too small for raw samples (%zu < %zu)", ws_left, need); return REN_E_WORKSPACE; }
    raw = (float*)ws;
    ws = (char*)align256(ws + need);
    ws_left = workspace_bytes - (size_t)(ws - (char*)workspace);
  }
  if (engine == REN_ENGINE_SIMT_FP32)
    rc = ren_simt_launch_fwd(p, n_rays, K, rays_o, rays_d, nullptr, nullptr, jitter, nc.kparams, (float*)stash, raw,
                             nc.on ? &nc.ngp : nullptr, st);
  else {
    unsigned char* live = nullptr;
    if (r->occ_binary != nullptr && packed == nullptr) {   // dense layout: find the tiles that emit no sample at all
      if (stash) {
        live = (unsigned char*)stash + tcs_live(p, S);
      } else {
        const size_t lb = (((size_t)(ren_pad128(S) >> 7)) + 255) & ~(size_t)255;
        if (ws_left < lb) { ren_set_error("workspace too small for the live-tile flags"); return REN_E_WORKSPACE; }
        live = (unsigned char*)ws;
        ws += lb;
        ws_left -= lb;
      }
      rc = ren_launch_tile_live(p, n_rays, K, rays_o, rays_d, jitter, live, st);
      if (rc) return rc;
    }
    if (S > 0)
      rc = ren_tc_launch_fwd(p, n_rays, K, rays_o, rays_d, nullptr, nullptr, nullptr, jitter, params, stash, raw, live,
                             packed ? packed->sample_idx : nullptr, S, ws, ws_left, st);
  }
  if (rc) return rc;
  return ren_launch_composite(p, n_rays, K, rays_o, rays_d, jitter, ray_offsets, raw, bkgd, radiance, opacity,
                              depth, n_kept, st);
}

extern "C" int ren_render_backward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                   const float* rays_o, const float* rays_d, const float* jitter,
                                   const float* params, const float* bkgd, const float* g_radiance,
                                   const float* g_opacity, const float* g_depth, const RenPacked* packed,
                                   const void* stash, size_t stash_bytes, float* g_params, float* g_bkgd,
                                   float* g_rays, void* workspace, size_t workspace_bytes, void* stream) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  if (n_rays < 0 || K < 1) { ren_set_error("bad n_rays/K"); return REN_E_INVALID; }
  if (n_rays == 0) return REN_OK;
  if (!rays_o || !rays_d || !params || !stash || !g_params || !workspace) {
    ren_set_error("ren_render_backward: null pointer argument");
    return REN_E_INVALID;
  }
  const int engine = pick_engine(p);
  rc = check_engine(p, engine);
  if (rc) return rc;
  p.r.engine = engine;
  long long S;
  rc = check_packed(packed, engine, n_rays, K, &S);
  if (rc) return rc;
  const int* ray_offsets = packed ? packed->ray_offsets : nullptr;
  const int* sample_idx = packed ? packed->sample_idx : nullptr;
  if (stash_bytes < stash_bytes_for(p, engine, S)) { ren_set_error("stash too small"); return REN_E_WORKSPACE; }
  char* ws = (char*)align256(workspace);
  size_t ws_left = workspace_bytes - (size_t)(ws - (char*)workspace);
  cudaStream_t st = (cudaStream_t)stream;
  NgpCtx nc;
  rc = ngp_prepare(f, p, params, g_params, workspace, workspace_bytes, &ws, &ws_left, true, st, &nc);
  if (rc) return rc;
  if (engine == REN_ENGINE_SIMT_FP32) {
    const size_t need = sizeof(float) * (size_t)ren_delta_floats(p, S);
    if (ws_left < need) { ren_set_error("workspace too small for backward (%zu < %zu)", ws_left, need); return REN_E_WORKSPACE; }
    float* delta = (float*)ws;
    float* graw = delta + p.act_floats * ren_pad128(S);
    const float* st_f = (const float*)stash;
    const float* raw = st_f + ren_stash_raw(p, S);
    rc = ren_launch_composite_bwd(p, n_rays, K, rays_o, rays_d, jitter, nullptr, raw, bkgd, g_radiance, g_opacity,
                                  g_depth, graw, g_bkgd, nullptr, st);
    if (rc) return rc;
    rc = ren_simt_launch_dgrad(p, S, nc.kparams, st_f, graw, delta, 0, st);
    if (rc) return rc;
    if (g_rays || nc.on) {      // arch ngp: the same pass scatters the grid gradient
      rc = ren_launch_raygrad(p, n_rays, K, rays_o, rays_d, nullptr, nullptr, jitter, nc.kparams, delta, nullptr, 0,
                              nullptr, nullptr, nullptr, 0, g_rays, nc.on ? &nc.ngp : nullptr, st);
      if (rc) return rc;
    }
    rc = ren_simt_launch_wgrad(p, S, st_f, delta, graw, nc.g_kparams, 0, st);
    if (rc) return rc;
    return nc.on ? ren_ngp_unpack_grads(p, nc.g_kparams, g_params, st) : REN_OK;
  }
  const size_t need = tcd_bytes(p, S);
  if (ws_left < need + 1024) { ren_set_error("workspace too small for backward (%zu < %zu)", ws_left, need); return REN_E_WORKSPACE; }
  char* delta = ws;
  float* graw = (float*)(ws + tcd_graw(p, S));
  char* pack_ws = (char*)align256(ws + need);
  const size_t pack_left = workspace_bytes - (size_t)(pack_ws - (char*)workspace);
  const float* raw = (const float*)((const char*)stash + tcs_raw(p, S));
  unsigned int* gmax = (unsigned int*)(ws + tcd_gmax(p, S));
  if (cudaMemsetAsync(gmax, 0, 256, st) != cudaSuccess) { ren_set_error("gmax memset failed"); return REN_E_CUDA; }
  rc = ren_launch_composite_bwd(p, n_rays, K, rays_o, rays_d, jitter, ray_offsets, raw, bkgd, g_radiance, g_opacity,
                                g_depth, graw, g_bkgd, gmax, st);
  if (rc) return rc;
  if (S == 0) {
    if (g_rays && cudaMemsetAsync(g_rays, 0, sizeof(float) * 6 * (size_t)n_rays, st) != cudaSuccess) return REN_E_CUDA;
    return REN_OK;
  }
  const unsigned char* live = (r->occ_binary != nullptr && packed == nullptr)
                                  ? (const unsigned char*)stash + tcs_live(p, S) : nullptr;
  rc = ren_tc_launch_dgrad(p, S, params, stash, graw, gmax, delta, 0, nullptr, nullptr, live, pack_ws, pack_left, st);
  if (rc) return rc;
  if (g_rays) {
    rc = ren_launch_raygrad(p, n_rays, K, rays_o, rays_d, nullptr, nullptr, jitter, params, delta, nullptr, 1, gmax,
                            live, sample_idx, S, g_rays, nullptr, st);
    if (rc) return rc;
  }
  return ren_tc_launch_wgrad(p, S, stash, delta, graw, gmax, g_params, 1, live, st);
}

// ------------------------------------------------------------------ tangent mode
// fp32 FFMA engine: (value, d/dw) row pairs through one pass.  tcgen05 engine: the value pass, then a
// tangent pass of the same chain kernel whose epilogue multiplies by act'(pre) from the value stash; the
// reverse is two dgrad passes (tangent adjoint, which also emits E = abar_dot act'' pre_dot, then value
// adjoint, which adds E) and two weight-gradient passes.
static int tangent_plan(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K, RenPlan* p) {
  int rc = ren_build_plan(f, r, p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  rc = check_occupancy(f, r);
  if (rc) return rc;
  if (n_rays < 0 || K < 1) { ren_set_error("bad n_rays/K"); return REN_E_INVALID; }
  const int engine = pick_engine(*p);
  rc = check_engine(*p, engine);
  if (rc) return rc;
  p->r.engine = engine;
  return REN_OK;
}
// tcgen05 tangent layouts (bytes).  stash: [value stash | tangent stash].  backward workspace:
// [delta_val mats | graw_val | gmax 256 | delta_dot mats | graw_dot | extra mats | packed weights]
static size_t tct_mats(const RenPlan& p, long long S) { return (size_t)2 * p.act_floats * ren_pad128(S); }
static size_t tct_graw(const RenPlan& p, long long S) { return (size_t)4 * (1 + p.C) * ren_pad128(S); }
static size_t tct_bwd_bytes(const RenPlan& p, long long S) {
  return 3 * tct_mats(p, S) + 2 * tct_graw(p, S) + 256 + 1024;
}

extern "C" int ren_tangent_stash_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                       size_t* bytes) {
  RenPlan p;
  int rc = tangent_plan(f, r, n_rays, K, &p);
  if (rc) return rc;
  const long long S = (long long)n_rays * K;
  *bytes = p.r.engine == REN_ENGINE_SIMT_FP32 ? sizeof(float) * (size_t)ren_stash_floats(p, 2 * S)
                                               : 2 * tcs_bytes(p, S) + 512;
  return REN_OK;
}

extern "C" int ren_tangent_workspace_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays,
                                           int32_t K, int32_t has_stash, int32_t for_backward, size_t* bytes) {
  RenPlan p;
  int rc = tangent_plan(f, r, n_rays, K, &p);
  if (rc) return rc;
  const long long S = (long long)n_rays * K;
  size_t n;
  if (p.r.engine == REN_ENGINE_SIMT_FP32) {
    n = for_backward ? sizeof(float) * (size_t)ren_delta_floats(p, 2 * S)
                     : (has_stash ? 0 : sizeof(float) * (size_t)((1 + p.C) * ren_pad128(2 * S)));
  } else {
    // without a caller stash the forward still needs both stashes (the tangent pass reads the value pass')
    n = for_backward ? tct_bwd_bytes(p, S) + ren_tc_bwd_pack_bytes(p)
                     : (has_stash ? 0 : 2 * tcs_bytes(p, S) + 512) + ren_tc_fwd_workspace_bytes(p);
  }
  *bytes = n + 2048 + ngp_ws_bytes(p);
  return REN_OK;
}

extern "C" int ren_render_tangent_forward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                          const float* rays_o, const float* rays_d, const float* rays_o_dot,
                                          const float* rays_d_dot, const float* jitter, const float* params,
                                          const float* bkgd, float* radiance, float* opacity, float* depth,
                                          float* radiance_dot, float* opacity_dot, int32_t* n_kept,
                                          const RenPacked* packed, void* stash, size_t stash_bytes, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  RenPlan p;
  int rc = tangent_plan(f, r, n_rays, K, &p);
  if (rc) return rc;
  if (n_rays == 0) return REN_OK;
  if (!rays_o || !rays_d || !rays_o_dot || !rays_d_dot || !params || !radiance || !opacity || !depth ||
      !radiance_dot || !opacity_dot || !n_kept) {
    ren_set_error("ren_render_tangent_forward: null pointer argument");
    return REN_E_INVALID;
  }
  if (r->stratified && jitter == nullptr) { ren_set_error("stratified marching needs a jitter buffer"); return REN_E_INVALID; }
  long long S;
  rc = check_packed(packed, p.r.engine, n_rays, K, &S);
  if (rc) return rc;
  const int* ray_offsets = packed ? packed->ray_offsets : nullptr;
  const int* sample_idx = packed ? packed->sample_idx : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align256(workspace);
  size_t ws_left = workspace ? workspace_bytes - (size_t)(ws - (char*)workspace) : 0;
  NgpCtx nc;
  rc = ngp_prepare(f, p, params, nullptr, workspace, workspace_bytes, &ws, &ws_left, false, st, &nc);
  if (rc) return rc;
  if (p.r.engine != REN_ENGINE_SIMT_FP32) {
    const size_t one = (tcs_bytes(p, S) + 255) & ~(size_t)255;
    char* sv;
    if (stash) {
      if (stash_bytes < 2 * tcs_bytes(p, S) + 512) { ren_set_error("stash too small"); return REN_E_WORKSPACE; }
      sv = (char*)align256(stash);
    } else {
      if (!workspace || ws_left < 2 * one) { ren_set_error("workspace too small for the tangent stashes"); return REN_E_WORKSPACE; }
      sv = ws;
      ws += 2 * one;
      ws_left -= 2 * one;
    }
    char* sd = sv + one;
    float* raw_val = (float*)(sv + tcs_raw(p, S));
    float* raw_dot = (float*)(sd + tcs_raw(p, S));
    unsigned char* live = nullptr;
    if (r->occ_binary != nullptr && packed == nullptr) {
      live = (unsigned char*)sv + tcs_live(p, S);
      rc = ren_launch_tile_live(p, n_rays, K, rays_o, rays_d, jitter, live, st);
      if (rc) return rc;
    }
    if (S > 0) {
      rc = ren_tc_launch_fwd(p, n_rays, K, rays_o, rays_d, nullptr, nullptr, nullptr, jitter, params, sv, raw_val,
                             live, sample_idx, S, ws, ws_left, st);
      if (rc) return rc;
      rc = ren_tc_launch_fwd(p, n_rays, K, rays_o, rays_d, rays_o_dot, rays_d_dot, sv, jitter, params, sd, raw_dot,
                             live, sample_idx, S, ws, ws_left, st);
      if (rc) return rc;
    }
    return ren_launch_composite_tan(p, n_rays, K, rays_o, rays_d, jitter, ray_offsets, raw_val, raw_dot, bkgd,
                                    radiance, opacity, depth, radiance_dot, opacity_dot, n_kept, st);
  }
  const long long S2 = 2 * S;
  float* raw;
  if (stash) {
    if (stash_bytes < sizeof(float) * (size_t)ren_stash_floats(p, S2)) { ren_set_error("stash too small"); return REN_E_WORKSPACE; }
    raw = (float*)stash + ren_stash_raw(p, S2);
  } else {
    const size_t need = sizeof(float) * (size_t)((1 + p.C) * ren_pad128(S2));
    if (!workspace || ws_left < need) { ren_set_error("workspace too small for raw samples (%zu < %zu)", ws_left, need); return REN_E_WORKSPACE; }
    raw = (float*)ws;
  }
  rc = ren_simt_launch_fwd(p, n_rays, K, rays_o, rays_d, rays_o_dot, rays_d_dot, jitter, nc.kparams, (float*)stash,
                           raw, nc.on ? &nc.ngp : nullptr, st);
  if (rc) return rc;
  return ren_launch_composite_tan(p, n_rays, K, rays_o, rays_d, jitter, nullptr, raw, nullptr, bkgd, radiance, opacity,
                                  depth, radiance_dot, opacity_dot, n_kept, st);
}

extern "C" int ren_render_tangent_backward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                           const float* rays_o, const float* rays_d, const float* rays_o_dot,
                                           const float* rays_d_dot, const float* jitter, const float* params,
                                           const float* bkgd, const float* g_radiance, const float* g_opacity,
                                           const float* g_depth, const float* g_radiance_dot,
                                           const float* g_opacity_dot, const RenPacked* packed, const void* stash,
                                           size_t stash_bytes, float* g_params, float* g_bkgd, float* g_rays,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  RenPlan p;
  int rc = tangent_plan(f, r, n_rays, K, &p);
  if (rc) return rc;
  if (n_rays == 0) return REN_OK;
  if (!rays_o || !rays_d || !rays_o_dot || !rays_d_dot || !params || !stash || !g_params || !workspace) {
    ren_set_error("ren_render_tangent_backward: null pointer argument");
    return REN_E_INVALID;
  }
  long long S;
  rc = check_packed(packed, p.r.engine, n_rays, K, &S);
  if (rc) return rc;
  const int* ray_offsets = packed ? packed->ray_offsets : nullptr;
  const int* sample_idx = packed ? packed->sample_idx : nullptr;
  char* ws = (char*)align256(workspace);
  size_t ws_left = workspace_bytes - (size_t)(ws - (char*)workspace);
  cudaStream_t st = (cudaStream_t)stream;
  NgpCtx nc;
  rc = ngp_prepare(f, p, params, g_params, workspace, workspace_bytes, &ws, &ws_left, true, st, &nc);
  if (rc) return rc;
  if (p.r.engine != REN_ENGINE_SIMT_FP32) {
    if (stash_bytes < 2 * tcs_bytes(p, S) + 512) { ren_set_error("stash too small"); return REN_E_WORKSPACE; }
    const size_t need = tct_bwd_bytes(p, S);
    if (ws_left < need) { ren_set_error("workspace too small for backward (%zu < %zu)", ws_left, need); return REN_E_WORKSPACE; }
    const size_t one = (tcs_bytes(p, S) + 255) & ~(size_t)255;
    const char* sv = (const char*)align256((void*)stash);
    const char* sd = sv + one;
    const float* raw_val = (const float*)(sv + tcs_raw(p, S));
    const float* raw_dot = (const float*)(sd + tcs_raw(p, S));
    char* delta_val = ws;
    float* graw_val = (float*)(delta_val + tct_mats(p, S));
    unsigned int* gmax = (unsigned int*)((char*)graw_val + tct_graw(p, S));
    char* delta_dot = (char*)gmax + 256;
    float* graw_dot = (float*)(delta_dot + tct_mats(p, S));
    char* extra = (char*)graw_dot + tct_graw(p, S);
    char* pack_ws = (char*)align256(extra + tct_mats(p, S));
    const size_t pack_left = workspace_bytes - (size_t)(pack_ws - (char*)workspace);
    if (cudaMemsetAsync(gmax, 0, 256, st) != cudaSuccess) { ren_set_error("gmax memset failed"); return REN_E_CUDA; }
    rc = ren_launch_composite_tan_bwd(p, n_rays, K, rays_o, rays_d, jitter, ray_offsets, raw_val, raw_dot, bkgd,
                                      g_radiance, g_opacity, g_depth, g_radiance_dot, g_opacity_dot, graw_val, graw_dot,
                                      g_bkgd, gmax, st);
    if (rc) return rc;
    if (S == 0) {
      if (g_rays && cudaMemsetAsync(g_rays, 0, sizeof(float) * 12 * (size_t)n_rays, st) != cudaSuccess) return REN_E_CUDA;
      return REN_OK;
    }
    const unsigned char* live = (r->occ_binary != nullptr && packed == nullptr)
                                    ? (const unsigned char*)sv + tcs_live(p, S) : nullptr;
    rc = ren_tc_launch_dgrad(p, S, params, sv, graw_dot, gmax, delta_dot, 1, sd, extra, live, pack_ws, pack_left, st);
    if (rc) return rc;
    rc = ren_tc_launch_dgrad(p, S, params, sv, graw_val, gmax, delta_val, 2, nullptr, extra, live, pack_ws, pack_left,
                             st);
    if (rc) return rc;
    if (g_rays) {
      rc = ren_launch_raygrad(p, n_rays, K, rays_o, rays_d, rays_o_dot, rays_d_dot, jitter, params, delta_val,
                              delta_dot, 1, gmax, live, sample_idx, S, g_rays, nullptr, st);
      if (rc) return rc;
    }
    rc = ren_tc_launch_wgrad(p, S, sv, delta_val, graw_val, gmax, g_params, 1, live, st);
    if (rc) return rc;
    return ren_tc_launch_wgrad(p, S, sd, delta_dot, graw_dot, gmax, g_params, 0, live, st);
  }
  const long long S2 = 2 * S;
  if (stash_bytes < sizeof(float) * (size_t)ren_stash_floats(p, S2)) { ren_set_error("stash too small"); return REN_E_WORKSPACE; }
  const size_t need = sizeof(float) * (size_t)ren_delta_floats(p, S2);
  if (ws_left < need) { ren_set_error("workspace too small for backward (%zu < %zu)", ws_left, need); return REN_E_WORKSPACE; }
  float* delta = (float*)ws;
  float* graw = delta + p.act_floats * ren_pad128(S2);
  const float* st_f = (const float*)stash;
  const float* raw = st_f + ren_stash_raw(p, S2);
  rc = ren_launch_composite_tan_bwd(p, n_rays, K, rays_o, rays_d, jitter, nullptr, raw, nullptr, bkgd, g_radiance,
                                    g_opacity, g_depth, g_radiance_dot, g_opacity_dot, graw, nullptr, g_bkgd, nullptr, st);
  if (rc) return rc;
  rc = ren_simt_launch_dgrad(p, S2, nc.kparams, st_f, graw, delta, 1, st);
  if (rc) return rc;
  if (g_rays || nc.on) {
    rc = ren_launch_raygrad(p, n_rays, K, rays_o, rays_d, rays_o_dot, rays_d_dot, jitter, nc.kparams, delta, nullptr, 0,
                            nullptr, nullptr, nullptr, 0, g_rays, nc.on ? &nc.ngp : nullptr, st);
    if (rc) return rc;
  }
  rc = ren_simt_launch_wgrad(p, S2, st_f, delta, graw, nc.g_kparams, 1, st);
  if (rc) return rc;
  return nc.on ? ren_ngp_unpack_grads(p, nc.g_kparams, g_params, st) : REN_OK;
}

// ------------------------------------------------------------------ density at free points
// The field kernels take rays; a point is the ray (o = x, d = 0): every marching step then sits at x exactly.
// One sample per point, no jitter, no grid, planes wide open; the raw sigma column becomes the density.
struct DensityArgs {
  long long n;
  int C1, density_activation;
  const float* raw;
  float* sigma;
};
extern "C" __global__ void ren_density_finish(const DensityArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const float sraw = a.raw[i * a.C1];
  float dact;
  float s = ren_density_act(a.density_activation, sraw, &dact);
  if (sraw == -INFINITY || sraw != sraw) s = 0.f;
  a.sigma[i] = s;
}
static void density_descs(const RenRenderDesc* r, RenRenderDesc* rr) {
  *rr = *r;
  rr->has_near = 1; rr->near_plane = 0.f; rr->has_far = 1; rr->far_plane = 1e9f;
  rr->render_step_size = 1.f; rr->cone_angle = 0.f; rr->early_stop_eps = 0.f; rr->alpha_thre = 0.f;
  rr->stratified = 0; rr->occ_binary = nullptr;
}
extern "C" int ren_density_workspace_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_points,
                                           size_t* bytes) {
  if (!r || !bytes || n_points < 0) { ren_set_error("ren_density_workspace_bytes: bad arguments"); return REN_E_INVALID; }
  RenRenderDesc rr;
  density_descs(r, &rr);
  size_t ws = 0;
  int rc = ren_workspace_bytes(f, &rr, n_points, 1, 0, 0, &ws);
  if (rc) return rc;
  *bytes = ws + sizeof(float) * 3 * (size_t)(n_points + 64) + 1024;     // (ren_workspace_bytes counts the ngp copy)
  return REN_OK;
}
extern "C" int ren_field_density(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_points, const float* points,
                                 const float* params, float* sigma, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  if (!r) { ren_set_error("null RenRenderDesc"); return REN_E_INVALID; }
  RenRenderDesc rr;
  density_descs(r, &rr);
  RenPlan p;
  int rc = ren_build_plan(f, &rr, &p);
  if (rc) return rc;
  rc = validate_render(&rr);
  if (rc) return rc;
  if (n_points < 0) { ren_set_error("bad n_points"); return REN_E_INVALID; }
  if (n_points == 0) return REN_OK;
  if (!points || !params || !sigma || !workspace) { ren_set_error("ren_field_density: null pointer argument"); return REN_E_INVALID; }
  const int engine = pick_engine(p);
  rc = check_engine(p, engine);
  if (rc) return rc;
  p.r.engine = engine;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align256(workspace);
  size_t ws_left = workspace_bytes - (size_t)(ws - (char*)workspace);
  NgpCtx nc;
  rc = ngp_prepare(f, p, params, nullptr, workspace, workspace_bytes, &ws, &ws_left, false, st, &nc);
  if (rc) return rc;
  const size_t zero_bytes = sizeof(float) * 3 * (size_t)n_points;
  const size_t raw_bytes = sizeof(float) * (size_t)((1 + p.C) * ren_pad128(n_points));
  if (ws_left < zero_bytes + raw_bytes + 512) { ren_set_error("ren_field_density: workspace too small"); return REN_E_WORKSPACE; }
  float* zeros = (float*)ws;
  if (cudaMemsetAsync(zeros, 0, zero_bytes, st) != cudaSuccess) { ren_set_error("density memset failed"); return REN_E_CUDA; }
  float* raw = (float*)align256(ws + zero_bytes);
  char* rest = (char*)align256((char*)raw + raw_bytes);
  ws_left = workspace_bytes - (size_t)(rest - (char*)workspace);
  if (engine == REN_ENGINE_SIMT_FP32)
    rc = ren_simt_launch_fwd(p, n_points, 1, points, zeros, nullptr, nullptr, nullptr, nc.kparams, nullptr, raw,
                             nc.on ? &nc.ngp : nullptr, st);
  else
    rc = ren_tc_launch_fwd(p, n_points, 1, points, zeros, nullptr, nullptr, nullptr, nullptr, params, nullptr, raw,
                           nullptr, nullptr, 0, rest, ws_left, st);
  if (rc) return rc;
  DensityArgs da;
  da.n = n_points; da.C1 = 1 + p.C; da.density_activation = f->density_activation; da.raw = raw; da.sigma = sigma;
  ren_prof_begin(REN_K_COMPOSITE_FWD, st);
  ren_density_finish<<<(unsigned)((n_points + 255) / 256), 256, 0, st>>>(da);
  ren_prof_end(REN_K_COMPOSITE_FWD, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ren_field_density launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}

// ------------------------------------------------------------------ packed march
#include <cub/device/device_scan.cuh>
extern "C" int ren_resolve_engine(const RenFieldDesc* f, const RenRenderDesc* r, int32_t* engine) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  if (!engine) { ren_set_error("ren_resolve_engine: null output"); return REN_E_INVALID; }
  const int e = pick_engine(p);
  rc = check_engine(p, e);
  if (rc) return rc;
  *engine = e;
  return REN_OK;
}
extern "C" int ren_march_workspace_bytes(int64_t n_rays, size_t* bytes) {
  if (n_rays < 0 || !bytes) { ren_set_error("ren_march_workspace_bytes: bad arguments"); return REN_E_INVALID; }
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (const int*)nullptr, (int*)nullptr, (int)(n_rays + 1));
  *bytes = scan + sizeof(int) * (size_t)(n_rays + 1) + 1024;
  return REN_OK;
}
extern "C" int ren_march_count(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                               const float* rays_o, const float* rays_d, const float* jitter, int32_t* ray_offsets,
                               void* workspace, size_t workspace_bytes, void* stream) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  rc = check_occupancy(f, r);
  if (rc) return rc;
  if (n_rays < 0 || K < 1 || (long long)n_rays * K >= (1LL << 31)) { ren_set_error("ren_march_count: bad n_rays/K"); return REN_E_INVALID; }
  if (!rays_o || !rays_d || !ray_offsets || !workspace) { ren_set_error("ren_march_count: null pointer argument"); return REN_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align256(workspace);
  int* counts = (int*)ws;
  char* scan_ws = (char*)align256(ws + sizeof(int) * (size_t)(n_rays + 1));
  size_t scan_bytes = workspace_bytes - (size_t)(scan_ws - (char*)workspace);
  if (cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)(n_rays + 1), st) != cudaSuccess) return REN_E_CUDA;
  rc = ren_launch_march_pack(p, n_rays, K, rays_o, rays_d, jitter, counts, nullptr, nullptr, st);
  if (rc) return rc;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, counts, ray_offsets, (int)(n_rays + 1), st);
  if (e != cudaSuccess) { ren_set_error("ren_march_count scan: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
extern "C" int ren_march_fill(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                              const float* rays_o, const float* rays_d, const float* jitter,
                              const int32_t* ray_offsets, int32_t* sample_idx, void* stream) {
  RenPlan p;
  int rc = ren_build_plan(f, r, &p);
  if (rc) return rc;
  rc = validate_render(r);
  if (rc) return rc;
  if (n_rays < 0 || K < 1) { ren_set_error("ren_march_fill: bad n_rays/K"); return REN_E_INVALID; }
  if (n_rays == 0) return REN_OK;
  if (!rays_o || !rays_d || !ray_offsets || !sample_idx) { ren_set_error("ren_march_fill: null pointer argument"); return REN_E_INVALID; }
  return ren_launch_march_pack(p, n_rays, K, rays_o, rays_d, jitter, nullptr, ray_offsets, sample_idx,
                               (cudaStream_t)stream);
}
