/* ren_b200.h -- C ABI of libren_b200.so: the B200 (sm_100a) event-NeRF hot path.
 *
 * The reference (wengflow/robust-e-nerf) has NO native code and NO FFI: its hot
 * path is Python calling third-party CUDA extensions (nerfacc, ATen/cuBLAS).
 * Each entry point below therefore cites the reference *Python* interface whose
 * arithmetic it replaces; `INTEGRATION.md` shows the ctypes binding a reference
 * maintainer would add at `RobustENeRF._build_nerf`
 * (robust_e_nerf/models/robust_e_nerf.py:204-260).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers into caller-owned, contiguous,
 *     row-major buffers (no allocation, no ownership transfer inside the library).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Return 0 on success, non-zero REN_E_* otherwise; `ren_last_error()` returns a
 *     thread-local message.  Nothing throws across the ABI.  No host sync inside.
 *   - MLP parameters are ONE flat fp32 buffer in the reference's state-dict order
 *     and `nn.Linear` layout ((out,in) row-major weight, then bias) -- see
 *     `ren_param_layout`.
 *   - Samples are laid out dense and ray-major: sample (r, k) at r*K + k, K =
 *     `ren_samples_per_ray` (an upper bound; dead tail samples are masked).
 */
#ifndef REN_B200_H
#define REN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REN_ABI_VERSION 5

enum {
  REN_OK = 0,
  REN_E_INVALID = 1,     /* bad argument / unsupported configuration */
  REN_E_CUDA = 2,        /* CUDA runtime error (message has the cudaError string) */
  REN_E_WORKSPACE = 3,   /* caller-provided workspace / stash too small */
  REN_E_UNSUPPORTED = 4  /* valid reference configuration this build does not cover */
};

/* models/nerf.py:17-29 activation tables */
enum { REN_HIDDEN_SOFTPLUS100 = 0, REN_HIDDEN_RELU = 1 };
enum { REN_DENSITY_SHIFTED_TRUNC_EXP = 0, REN_DENSITY_SOFTPLUS = 1, REN_DENSITY_SHIFTED_SOFTPLUS = 2 };
enum { REN_RADIANCE_SOFTPLUS = 0, REN_RADIANCE_SIGMOID = 1 };
/* models/robust_e_nerf.py:214-218 contraction_type */
enum { REN_CONTRACT_AABB = 0, REN_CONTRACT_SPHERE = 1, REN_CONTRACT_TANH = 2 };
/* GEMM engine for the MLP */
enum {
  REN_ENGINE_AUTO = 0,    /* tcgen05 when the network shape is covered, else SIMT */
  REN_ENGINE_SIMT_FP32 = 1, /* fp32 FFMA tiles: parity baseline, any shape */
  REN_ENGINE_TC_F16X3 = 2,  /* tcgen05 kind::f16, split fp16 hi/lo operands, 3 MMAs (fp32-class accuracy) */
  REN_ENGINE_TC_F16 = 3,    /* tcgen05 kind::f16, single pass (throughput mode; misses the 1e-4 bar) */
  REN_ENGINE_TC_F16A = 4    /* tcgen05 kind::f16, activations single fp16, weights split hi/lo: 2 MMAs */
};
/* models/nerf.py:104-167 `arch` */
enum { REN_ARCH_MLP = 0, REN_ARCH_NGP = 1 };
/* tiny-cuda-nn grid encodings (`otype` in the configs/train YAMLs, line 63: HashGrid | DenseGrid | TiledGrid) */
enum { REN_GRID_HASH = 0, REN_GRID_DENSE = 1, REN_GRID_TILED = 2 };
/* loss_metric/loss.py:16-30 error functions */
enum { REN_ERR_L1 = 0, REN_ERR_MSE = 1, REN_ERR_MAPE = 2 };

/* `VanillaNeRFRadianceField.__init__` (external/mlp.py:247-319) + the
 * `model.nerf.mlp` YAML block (configs/train/synthetic.yaml:85-96). */
typedef struct RenFieldDesc {
  int32_t net_depth;          /* trunk layers (8) */
  int32_t net_width;          /* trunk width (256), multiple of 16 */
  int32_t skip_layer;         /* concat encoding after layer i when i%skip==0 && i>0; 0 = none */
  int32_t net_depth_condition;/* view-conditioned hidden layers (1) */
  int32_t net_width_condition;/* (128), multiple of 16 */
  int32_t radiance_dim;       /* C: 1 (monochrome) or 3 (Bayer) */
  int32_t pos_deg;            /* L_pos (10) -> 3*(1+2L) = 63 features */
  int32_t dir_deg;            /* L_dir (4)  -> 27 features */
  int32_t hidden_activation;  /* REN_HIDDEN_* */
  int32_t density_activation; /* REN_DENSITY_* */
  int32_t radiance_activation;/* REN_RADIANCE_* */
  int32_t contraction;        /* REN_CONTRACT_* */
  float aabb[6];              /* xmin,ymin,zmin,xmax,ymax,zmax (external/mlp.py:269) */
  /* `arch: ngp` -- `NGPradianceField.__init__` (external/ngp.py:109-207) + the `model.nerf.ngp` YAML block
   * (configs/train/mocap-desk2.yaml:61-88): multiresolution grid encoding of the contracted position
   * (tiny-cuda-nn `Encoding`, interpolation Linear), real spherical harmonics of the raw direction
   * (external/sh_encoder.py:15-193), mlp_base grid -> [density | geo_feat_dim features], mlp_head
   * [SH | features] -> radiance.  With arch = REN_ARCH_NGP the fields above are read as: net_depth / net_width =
   * mlp_base n_hidden_layers / n_neurons, net_depth_condition / net_width_condition = mlp_head
   * n_hidden_layers (>= 1) / n_neurons, hidden_activation = both MLPs'; skip_layer, pos_deg, dir_deg unused.
   * All zero (arch = REN_ARCH_MLP) for the vanilla field. */
  int32_t arch;               /* REN_ARCH_* */
  int32_t grid_type;          /* REN_GRID_* */
  int32_t grid_levels;        /* n_levels (16), <= 32 */
  int32_t grid_features;      /* n_features_per_level (2): 1, 2, 4 or 8; levels * features <= 64 */
  int32_t grid_log2_hashmap_size; /* (19) */
  int32_t grid_base_resolution;   /* (16) */
  float grid_per_level_scale;     /* (1.4472692) */
  int32_t sh_degree;          /* dir_encoding.degree (4), 1..4 */
  int32_t geo_feat_dim;       /* (15), 1..63 */
  int32_t reserved_ngp;
} RenFieldDesc;

/* `NeRF.__init__` rendering options (models/nerf.py:31-48) as consumed by
 * `render_image` / nerfacc.ray_marching (external/utils.py:106-119). */
typedef struct RenRenderDesc {
  float near_plane;      /* used iff has_near */
  float far_plane;       /* used iff has_far */
  float render_step_size;
  float cone_angle;
  float early_stop_eps;
  float alpha_thre;
  int32_t has_near;
  int32_t has_far;
  int32_t stratified;    /* training: t_min += u*step, one u per ray */
  int32_t engine;        /* REN_ENGINE_* */
  /* nerfacc.OccupancyGrid binary grid (models/nerf.py:98-102, external/utils.py:106-119): device pointer to
   * occ_res[0]*occ_res[1]*occ_res[2] bytes (0 / 1), index (ix*res_y + iy)*res_z + iz over the contracted unit
   * cube of the field's aabb / contraction; NULL = every cell occupied.  Marching steps whose midpoint falls
   * in an unoccupied cell emit no sample.  All three contractions: for `aabb` nerfacc leaves an empty voxel in
   * whole marching steps, so its samples are the same lattice steps this per-step test keeps. */
  const uint8_t* occ_binary;
  int32_t occ_res[3];
  int32_t reserved;
} RenRenderDesc;

/* Packed samples of one render call, as nerfacc.ray_marching hands them to the field
 * (external/utils.py:106-128): only the marching steps that land in occupied grid cells.  Built by
 * ren_march_count + (host: n_samples = ray_offsets[n_rays]) + ren_march_fill; passed to the four render
 * entry points (NULL = dense (ray, step) layout).  With it, stash / workspace sizes are those of
 * (n_rays = n_samples, K = 1).  tcgen05 engine only. */
typedef struct RenPacked {
  int64_t n_samples;            /* total emitted samples (host value) */
  const int32_t* sample_idx;    /* (n_samples) dense index ray*K + step of each sample, ray-major marching order */
  const int32_t* ray_offsets;   /* (n_rays + 1) exclusive scan of the per-ray counts */
} RenPacked;

int ren_version(void);
const char* ren_last_error(void);

/* Measurement hooks (bench.py): kernels launched by this library so far; optional CUDA-event
 * timing of every launch on its own stream, summed per kernel kind
 * (0 field fwd, 1 field dgrad, 2 wgrad, 3 composite fwd, 4 composite bwd, 5 event rays,
 *  6 loss, 7 adam, 8 weight pack).  `ren_timing_collect` waits for the recorded events. */
int64_t ren_launch_count(void);
int ren_timing_enable(int32_t on);
int ren_timing_collect(int32_t kind, double* total_ms, int64_t* count);

/* Number of fp32 elements of the flat parameter buffer (arch ngp: grid parameters first, then the MLPs). */
int ren_param_count(const RenFieldDesc* f, int64_t* n_params);

/* arch ngp: number of grid parameters (= `tcnn.Encoding.params.numel()`, state-dict key
 * `radiance_field.mlp_base.0.params`, external/ngp.py:166-188) at the head of the flat buffer; 0 for arch mlp.
 * Per level l: scale_l = fp32(2^(l log2 per_level_scale) base_resolution - 1), resolution ceil(scale_l) + 1,
 * entries min(roundup8(resolution^3), 2^log2_hashmap_size) (Hash), entry = features floats. */
int ren_grid_param_count(const RenFieldDesc* f, int64_t* n_grid_params);

/* Layout of the flat buffer: for layer l (state-dict order: base.hidden_layers.0..D-1,
 * sigma_layer.output_layer, bottleneck_layer.output_layer, rgb_layer.hidden_layers.*,
 * rgb_layer.output_layer; external/mlp.py:56-76,143-186) writes weight offset, bias
 * offset, out and in features.  Returns the number of layers via *n_layers; arrays may
 * be NULL to query the count.  arch ngp: mlp_base.1.hidden_layers.*, mlp_base.1.output_layer,
 * mlp_head.hidden_layers.*, mlp_head.output_layer (external/ngp.py:171-207), offsets after the grid. */
int ren_param_layout(const RenFieldDesc* f, int32_t* n_layers, int64_t* w_off, int64_t* b_off,
                     int32_t* out_features, int32_t* in_features);

/* Upper bound K on samples per ray for the dense layout (host arithmetic only). */
int ren_samples_per_ray(const RenFieldDesc* f, const RenRenderDesc* r, int32_t* K);

/* Bytes of the activation stash `ren_render_forward` fills when gradients are wanted,
 * and of the scratch workspace `ren_render_forward` (given has_stash) or
 * `ren_render_backward` (for_backward != 0) needs. */
int ren_stash_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                    size_t* bytes);
int ren_workspace_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                        int32_t has_stash, int32_t for_backward, size_t* bytes);

/* The engine REN_ENGINE_AUTO resolves to for this network (REN_ENGINE_SIMT_FP32 or REN_ENGINE_TC_F16X3). */
int ren_resolve_engine(const RenFieldDesc* f, const RenRenderDesc* r, int32_t* engine);

/* Phase 1 / 2 of the packed march (replaces the sample enumeration of `nerfacc.ray_marching`,
 * external/utils.py:106-119).  ren_march_count: ray_offsets (n_rays + 1) i32 = exclusive scan of the per-ray
 * emitted-step counts (workspace: ren_march_workspace_bytes).  ren_march_fill: sample_idx (ray_offsets[n_rays]). */
int ren_march_workspace_bytes(int64_t n_rays, size_t* bytes);
int ren_march_count(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                    const float* rays_o, const float* rays_d, const float* jitter, int32_t* ray_offsets,
                    void* workspace, size_t workspace_bytes, void* stream);
int ren_march_fill(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                   const float* rays_o, const float* rays_d, const float* jitter, const int32_t* ray_offsets,
                   int32_t* sample_idx, void* stream);

/* Replaces `NeRF.forward` (models/nerf.py:230-286) -> `render_image`
 * (external/utils.py:38-140) -> nerfacc.ray_marching + `rendering`
 * (external/vol_rendering.py:16-128) -> `VanillaNeRFRadianceField.forward`
 * (external/mlp.py:350-358).
 *   rays_o, rays_d  (n_rays,3) f32      jitter (n_rays) f32 in [0,1) or NULL
 *   params          flat fp32 MLP parameters (ren_param_layout)
 *   bkgd            (C) f32 background radiance (already soft-plussed) or NULL
 *   radiance        (n_rays,C) f32  = sum w c + bkgd (1-O)
 *   opacity         (n_rays)   f32  = sum w
 *   depth           (n_rays)   f32  = sum w (t0+t1)/2      (NOT yet divided by O+eps)
 *   n_kept          (n_rays)   i32  samples surviving march + visibility cull
 *   stash           NULL for no-grad rendering, else ren_stash_bytes() bytes */
int ren_render_forward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                       const float* rays_o, const float* rays_d, const float* jitter,
                       const float* params, const float* bkgd, float* radiance, float* opacity,
                       float* depth, int32_t* n_kept, const RenPacked* packed, void* stash, size_t stash_bytes,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Replaces `VanillaNeRFRadianceField.query_density` (external/mlp.py:344-348) on free points, as the occupancy
 * grid update evaluates it (models/nerf.py:171-198): sigma (n_points) f32 = density_activation(sigma_raw) *
 * selector.  points (n_points,3) f32.  workspace: ren_density_workspace_bytes. */
int ren_density_workspace_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_points, size_t* bytes);
int ren_field_density(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_points, const float* points,
                      const float* params, float* sigma, void* workspace, size_t workspace_bytes, void* stream);

/* Reverse pass of the above (replaces autograd through the same reference files).
 *   g_radiance (n_rays,C), g_opacity (n_rays), g_depth (n_rays): upstream grads (NULL = 0)
 *   g_params   flat fp32, ACCUMULATED into (caller zeroes)
 *   g_bkgd     (C) accumulated, or NULL
 *   g_rays     (n_rays,6) f32 = d loss / d [o | d] (written) or NULL -- needed only when the
 *              timestamps are trainable (refractory period) */
int ren_render_backward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                        const float* rays_o, const float* rays_d, const float* jitter,
                        const float* params, const float* bkgd, const float* g_radiance,
                        const float* g_opacity, const float* g_depth, const RenPacked* packed,
                        const void* stash, size_t stash_bytes, float* g_params, float* g_bkgd, float* g_rays,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Tangent-mode render: the same render PLUS its derivative along one scalar per ray (the pose
 * interpolation weight w of the ray's timestamp), replacing
 * `autograd.gradient(log_intensity, ts, create_graph=True)` (utils/autograd.py:4-34, call site
 * models/robust_e_nerf.py:394-398) by a forward-mode pass: value and tangent rows go through the
 * same GEMMs, so d logI / d t = radiance_dot / (radiance + min_modeled_intensity) * dw/dt.
 *   rays_o_dot, rays_d_dot (n_rays,3) f32: d o / d w, d d / d w (ren_event_rays)
 *   radiance_dot (n_rays,C), opacity_dot (n_rays): d radiance / d w, d opacity / d w
 * The march (t0, t1) carries no gradient, as in the reference (nerfacc marches under no_grad).
 * Stash / workspace sizes come from ren_tangent_stash_bytes / ren_tangent_workspace_bytes. */
int ren_tangent_stash_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                            size_t* bytes);
int ren_tangent_workspace_bytes(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                int32_t has_stash, int32_t for_backward, size_t* bytes);
int ren_render_tangent_forward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                               const float* rays_o, const float* rays_d, const float* rays_o_dot,
                               const float* rays_d_dot, const float* jitter, const float* params,
                               const float* bkgd, float* radiance, float* opacity, float* depth,
                               float* radiance_dot, float* opacity_dot, int32_t* n_kept,
                               const RenPacked* packed, void* stash, size_t stash_bytes, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Reverse pass of the tangent-mode render (the reference's double-backward through
 * external/mlp.py, external/vol_rendering.py): upstream grads of the five per-ray outputs ->
 * g_params / g_bkgd (accumulated) and, when g_rays != NULL, (n_rays,12) f32 =
 * d loss / d [o | d | o_dot | d_dot] (written), which chain to the timestamp / refractory period. */
int ren_render_tangent_backward(const RenFieldDesc* f, const RenRenderDesc* r, int64_t n_rays, int32_t K,
                                const float* rays_o, const float* rays_d, const float* rays_o_dot,
                                const float* rays_d_dot, const float* jitter, const float* params,
                                const float* bkgd, const float* g_radiance, const float* g_opacity,
                                const float* g_depth, const float* g_radiance_dot,
                                const float* g_opacity_dot, const RenPacked* packed, const void* stash,
                                size_t stash_bytes, float* g_params, float* g_bkgd, float* g_rays,
                                void* workspace, size_t workspace_bytes, void* stream);

/* Replaces `RefractoryPeriod.forward` (models/event_generation_params.py:196-203), the
 * supervision-timestamp derivation (models/robust_e_nerf.py:322-357),
 * `LinearTrajectory.forward` (models/trajectories.py:30-91, utils/tensor_ops.py:83-180)
 * and `NeRF.pixel_params_to_ray` (models/nerf.py:206-228) for the two renders of the
 * log-intensity-difference loss (and optionally the gradient-loss timestamp).
 *   position (B,2) f32; start_ts,end_ts (B) i64; nz_* (B) f64 normalised samples
 *   tau (1) f64 device scalar (refractory period, ns)
 *   pose_ts (P) i64, pose_pos (P,3) f32, pose_quat (P,4) f32 XYZW; k_inv (9) f32
 *   n_renders = 2: rays for [diff_start | diff_end]; 3: [grad | diff_start | diff_end]
 *   out_ts (n_renders,B) f64; out_o,out_d (n_renders,B,3) f32; out_rot (n_renders,B,9) f32 or NULL
 *   out_ts_diff (B) f64 (the diff interval length), out_start (B) f64 = start_ts + tau
 *   out_o_dot, out_d_dot, out_d_ddot (n_renders,B,3) f32 and out_dw_dt (n_renders,B) f64, all NULL
 *   or all set: first / second derivative of the ray wrt the pose-interpolation weight w and
 *   dw/dt = 1 / (pose bin width) -- what autograd propagates through
 *   models/trajectories.py:63-91 for the log-intensity-gradient loss and a trainable tau */
int ren_event_rays(int64_t B, const float* position, const int64_t* start_ts,
                   const int64_t* end_ts, const double* nz_ts_diff, const double* nz_diff_start,
                   const double* nz_grad_ts, const double* tau, int64_t n_poses,
                   const int64_t* pose_ts, const float* pose_pos, const float* pose_quat,
                   const float* k_inv, int32_t n_renders, double* out_ts, float* out_o,
                   float* out_d, float* out_rot, double* out_ts_diff, double* out_start,
                   float* out_o_dot, float* out_d_dot, float* out_d_ddot, double* out_dw_dt,
                   void* stream);

/* Replaces `LinearTrajectory.forward` (models/trajectories.py:30-91) alone: pose at n
 * timestamps.  ts (n) f64 -> out_pos (n,3) f32, out_rot (n,9) f32 row-major. */
int ren_trajectory(int64_t n, const double* ts, int64_t n_poses, const int64_t* pose_ts,
                   const float* pose_pos, const float* pose_quat, float* out_pos, float* out_rot,
                   void* stream);

/* Replaces `Loss.log_intensity_diff` (loss_metric/loss.py:64-74) on top of
 * `render_pixels` post-processing (models/robust_e_nerf.py:865-871,432-443):
 *   I = C + min_modeled_intensity, L = log I, valid = (O0>0)|(O1>0) (all-true with bkgd),
 *   err = error_fn(L1 - L0, f32(ts_diff * ev_diff/(end-start'))), loss = mean_valid(err).
 *   c0,c1,o0,o1 (B) f32; ev_diff (B) f32; ts_diff,start,(B) f64; end_ts (B) i64
 *   out: log_i0, log_i1 (B) f32, loss (1) f32 (unweighted masked mean), n_valid (1) i32,
 *        g_c0, g_c1 (B) f32 = d loss / d c (NULL to skip), scratch (>= 16 bytes, zeroed inside) */
int ren_event_diff_loss(int64_t B, int32_t error_fn, int32_t all_valid, float min_modeled_intensity,
                        const float* c0, const float* c1, const float* o0, const float* o1,
                        const float* ev_diff, const double* ts_diff, const double* start,
                        const int64_t* end_ts, float* log_i0, float* log_i1, float* loss,
                        int32_t* n_valid, float* g_c0, float* g_c1, void* scratch, void* stream);

/* Raw event stream -> per-pixel event intervals: replaces the per-event Python loops
 * `Event.queue_raw_events` (data/datasets.py:189-284) and `Event.extract_max_refractory_period`
 * (data/datasets.py:132-187) of the reference's dataset preparation (`raw_events.npz` ->
 * `events.pt`, `max_refractory_period.pt`).  The stream must be time-ordered, as the reference assumes.
 *   position (n,2) u16 (x,y); timestamp (n) i64 ns; polarity (n) u8 {0,1}
 *   out_* sized for n events; the first *out_n_valid entries are written, in stream order:
 *   out_position (n,2) i64, out_start_ts / out_end_ts / out_num_pos / out_num_neg (n) i64
 *   out_n_valid (1) i64 device scalar; out_min_interval (1) i64 device scalar = the maximum
 *   refractory period (smallest positive same-pixel interval), INT64_MAX when there is none. */
int ren_event_queue_workspace_bytes(int64_t n_events, int32_t img_width, int32_t img_height, size_t* bytes);
int ren_event_queue(int64_t n_events, int32_t img_width, int32_t img_height, const uint16_t* position,
                    const int64_t* timestamp, const uint8_t* polarity, int64_t* out_position,
                    int64_t* out_start_ts, int64_t* out_end_ts, int64_t* out_num_pos, int64_t* out_num_neg,
                    int64_t* out_n_valid, int64_t* out_min_interval, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Fused Adam step on a flat fp32 buffer (torch.optim.Adam semantics incl. L2
 * weight_decay added to the gradient; models/robust_e_nerf.py:800-813).
 * `step` is the 1-based step count; grad_scale multiplies the gradient first
 * (1/world after a sum all-reduce). */
int ren_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  float lr, float beta1, float beta2, float eps, float weight_decay,
                  float grad_scale, int64_t step, void* stream);
/* Same update with the 1-based step count read from device memory (`step`: (1) i64 device scalar the caller
 * increments on the same stream), so that a CUDA graph capturing the whole training step stays valid across
 * replays.  torch.optim.Adam(capturable=True) keeps its step on the device for the same reason. */
int ren_adam_step_dev(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                      float lr, float beta1, float beta2, float eps, float weight_decay,
                      float grad_scale, const int64_t* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REN_B200_H */
