// ren_ngp.cu -- `arch: ngp` (external/ngp.py:109-280): host-side grid table, and the mapping between the
// caller's flat parameter buffer (the reference's state-dict layout) and the kernel-side MLP buffer.
//
//   caller's flat buffer   [ grid (mlp_base.0.params) | mlp_base.1.hidden_layers.* | mlp_base.1.output_layer
//                            | mlp_head.hidden_layers.* | mlp_head.output_layer ]        (W (out,in) then b, each)
//   kernel-side buffer     the vanilla topology of ren_build_plan: trunk layers = mlp_base hidden layers;
//                          sigma head = row 0 of mlp_base's output layer; bottleneck = its rows 1..geo, padded
//                          with zero rows to WB; condition layer 0 = mlp_head layer 0 with its input columns
//                          reordered from [SH | features] (external/ngp.py:246-252) to [features (WB) | SH];
//                          further condition layers and the radiance head copied as they are.
// The fp32 engine (ren_simt.cu), the compositing kernels and the weight-gradient kernel then run unchanged;
// only the encoders differ (grid + spherical harmonics instead of the two sinusoidal ones: ren_common.cuh,
// prologue of ren_simt_field_fwd, reverse in ren_raygrad.cu).  Gradients come back through the same map.
#include <math.h>
#include <string.h>

#include "ren_common.cuh"

int ren_build_grid(const RenFieldDesc* f, RenGrid* g, long long* n_entries) {
  memset(g, 0, sizeof(*g));
  if (f->grid_levels < 1 || f->grid_levels > REN_MAX_LEVELS) { ren_set_error("grid_levels %d outside [1,%d]", f->grid_levels, REN_MAX_LEVELS); return REN_E_INVALID; }
  const int F = f->grid_features;
  if (!(F == 1 || F == 2 || F == 4 || F == 8) || f->grid_levels * F > REN_PE_MAX) { ren_set_error("grid_features %d (x %d levels) unsupported", F, f->grid_levels); return REN_E_INVALID; }
  if (f->grid_type < 0 || f->grid_type > 2 || f->grid_base_resolution < 1 || !(f->grid_per_level_scale > 0.f) ||
      f->grid_log2_hashmap_size < 1 || f->grid_log2_hashmap_size > 28) {
    ren_set_error("bad grid encoding configuration");
    return REN_E_INVALID;
  }
  g->n_levels = f->grid_levels; g->F = F; g->type = f->grid_type;
  const double log2_pls = log2((double)f->grid_per_level_scale);
  unsigned long long offset = 0;
  for (int l = 0; l < g->n_levels; ++l) {
    RenGridLevel& lv = g->lv[l];
    lv.scale = (float)(exp2((double)l * log2_pls) * (double)f->grid_base_resolution - 1.0);
    lv.res = (uint32_t)ceilf(lv.scale) + 1u;
    unsigned long long n = (unsigned long long)lv.res * lv.res * lv.res;
    const unsigned long long cap = 0xFFFFFFFFull / 2;
    if (n > cap) n = cap;
    n = (n + 7) / 8 * 8;
    if (g->type == REN_GRID_TILED) {
      const unsigned long long b = (unsigned long long)f->grid_base_resolution * f->grid_base_resolution * f->grid_base_resolution;
      if (n > b) n = b;
    } else if (g->type == REN_GRID_HASH) {
      const unsigned long long t = 1ull << f->grid_log2_hashmap_size;
      if (n > t) n = t;
    }
    lv.entries = (uint32_t)n;
    lv.offset = (uint32_t)offset;
    offset += n;
    if (offset > 0x7FFFFFFFull) { ren_set_error("grid too large"); return REN_E_UNSUPPORTED; }
  }
  if (n_entries) *n_entries = (long long)offset;
  return REN_OK;
}

// ------------------------------------------------------------------ flat <-> kernel-side parameter map
// One tensor of the kernel-side buffer: dst[n][k] (n < rows, k < cols, row stride cols) takes
// src[(row0 + n) * ld + col(k)] when n < src_rows and k lies in one of two column segments, else 0.
struct NgpMap {
  int dst_off, rows, cols;
  long long src_off;
  int ld, row0, src_rows;
  int seg_dst[2], seg_n[2], seg_src[2];
};
#define NGP_MAX_MAPS 32
struct NgpMapArgs {
  NgpMap m[NGP_MAX_MAPS];
  int n_maps;
  const float* flat;     // pack: source;   unpack: unused
  float* kp;             // pack: destination; unpack: source (kernel-side gradient)
  float* g_flat;         // unpack: destination (accumulated)
};

static int build_maps(const RenPlan& p, NgpMapArgs* a) {
  int n = 0;
  long long src = p.n_grid;
  auto plain = [&](int dst_w, int dst_b, int rows, int cols) {
    NgpMap& w = a->m[n++];
    memset(&w, 0, sizeof(w));
    w.dst_off = dst_w; w.rows = rows; w.cols = cols; w.src_off = src; w.ld = cols; w.src_rows = rows;
    w.seg_n[0] = cols;
    src += (long long)rows * cols;
    NgpMap& b = a->m[n++];
    memset(&b, 0, sizeof(b));
    b.dst_off = dst_b; b.rows = rows; b.cols = 1; b.src_off = src; b.ld = 1; b.src_rows = rows; b.seg_n[0] = 1;
    src += rows;
  };
  // mlp_base hidden layers
  for (int i = 0; i < p.D; ++i) plain(p.trunk[i].w_off, p.trunk[i].b_off, p.trunk[i].N, p.trunk[i].K0);
  // mlp_base output layer (1 + geo, W): row 0 -> sigma head, rows 1.. -> bottleneck
  {
    const long long w_src = src, b_src = src + (long long)(1 + p.geo) * p.W;
    NgpMap m;
    memset(&m, 0, sizeof(m));
    m.dst_off = p.sig_w; m.rows = 1; m.cols = p.W; m.src_off = w_src; m.ld = p.W; m.src_rows = 1; m.seg_n[0] = p.W;
    a->m[n++] = m;
    memset(&m, 0, sizeof(m));
    m.dst_off = p.sig_b; m.rows = 1; m.cols = 1; m.src_off = b_src; m.ld = 1; m.src_rows = 1; m.seg_n[0] = 1;
    a->m[n++] = m;
    memset(&m, 0, sizeof(m));
    m.dst_off = p.bott.w_off; m.rows = p.WB; m.cols = p.W; m.src_off = w_src; m.ld = p.W; m.row0 = 1; m.src_rows = p.geo;
    m.seg_n[0] = p.W;
    a->m[n++] = m;
    memset(&m, 0, sizeof(m));
    m.dst_off = p.bott.b_off; m.rows = p.WB; m.cols = 1; m.src_off = b_src; m.ld = 1; m.row0 = 1; m.src_rows = p.geo;
    m.seg_n[0] = 1;
    a->m[n++] = m;
    src = b_src + 1 + p.geo;
  }
  // mlp_head layer 0 (WC, DE + geo): [features | SH] <- [SH | features]
  {
    const int in_src = p.DE + p.geo;
    NgpMap m;
    memset(&m, 0, sizeof(m));
    m.dst_off = p.cond[0].w_off; m.rows = p.WC; m.cols = p.WB + p.DE; m.src_off = src; m.ld = in_src; m.src_rows = p.WC;
    m.seg_dst[0] = 0; m.seg_n[0] = p.geo; m.seg_src[0] = p.DE;
    m.seg_dst[1] = p.WB; m.seg_n[1] = p.DE; m.seg_src[1] = 0;
    a->m[n++] = m;
    src += (long long)p.WC * in_src;
    memset(&m, 0, sizeof(m));
    m.dst_off = p.cond[0].b_off; m.rows = p.WC; m.cols = 1; m.src_off = src; m.ld = 1; m.src_rows = p.WC; m.seg_n[0] = 1;
    a->m[n++] = m;
    src += p.WC;
  }
  for (int j = 1; j < p.CD; ++j) plain(p.cond[j].w_off, p.cond[j].b_off, p.WC, p.WC);
  plain(p.rgb_w, p.rgb_b, p.C, p.WC);
  if (src != p.n_flat) { ren_set_error("ngp parameter map: internal size mismatch (%lld != %lld)", src, p.n_flat); return REN_E_INVALID; }
  a->n_maps = n;
  return REN_OK;
}

__device__ __forceinline__ long long ngp_src_index(const NgpMap& m, int nrow, int k) {
  if (nrow >= m.src_rows) return -1;
#pragma unroll
  for (int s = 0; s < 2; ++s)
    if (k >= m.seg_dst[s] && k < m.seg_dst[s] + m.seg_n[s])
      return m.src_off + (long long)(m.row0 + nrow) * m.ld + m.seg_src[s] + (k - m.seg_dst[s]);
  return -1;
}

extern "C" __global__ void ren_ngp_pack_kernel(const NgpMapArgs a) {
  const NgpMap& m = a.m[blockIdx.y];
  const int total = m.rows * m.cols;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const long long si = ngp_src_index(m, e / m.cols, e % m.cols);
    a.kp[m.dst_off + e] = si >= 0 ? a.flat[si] : 0.f;
  }
}
extern "C" __global__ void ren_ngp_unpack_kernel(const NgpMapArgs a) {
  const NgpMap& m = a.m[blockIdx.y];
  const int total = m.rows * m.cols;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const long long si = ngp_src_index(m, e / m.cols, e % m.cols);
    if (si >= 0) a.g_flat[si] += a.kp[m.dst_off + e];       // every flat element has exactly one image
  }
}

// flat (caller) -> kernel-side MLP parameters
int ren_ngp_pack(const RenPlan& p, const float* flat, float* kparams, cudaStream_t st) {
  NgpMapArgs a;
  static_assert(sizeof(NgpMapArgs) <= 4000, "parameter block too large");
  int rc = build_maps(p, &a);
  if (rc) return rc;
  a.flat = flat; a.kp = kparams; a.g_flat = nullptr;
  ren_prof_begin(REN_K_PACK, st);
  ren_ngp_pack_kernel<<<dim3(8, a.n_maps), 256, 0, st>>>(a);
  ren_prof_end(REN_K_PACK, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ngp pack launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
// kernel-side gradient -> accumulated into the caller's flat gradient
int ren_ngp_unpack_grads(const RenPlan& p, const float* g_kparams, float* g_flat, cudaStream_t st) {
  NgpMapArgs a;
  int rc = build_maps(p, &a);
  if (rc) return rc;
  a.flat = nullptr; a.kp = const_cast<float*>(g_kparams); a.g_flat = g_flat;
  ren_prof_begin(REN_K_PACK, st);
  ren_ngp_unpack_kernel<<<dim3(8, a.n_maps), 256, 0, st>>>(a);
  ren_prof_end(REN_K_PACK, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ren_set_error("ngp unpack launch: %s", cudaGetErrorString(e)); return REN_E_CUDA; }
  return REN_OK;
}
