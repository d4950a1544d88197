"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Self-contained torch-CPU restatement of the reference's event-NeRF hot path
(SURVEY.md section 8 rows a1-a19), written functionally over plain tensors so it
travels to the GPU box where /root/reference does not exist.  Every function cites
the reference file:line it follows.  All arithmetic is done in the dtype of the
inputs / parameters (fp32 to mirror the reference, fp64 to print the fp32 noise
floor), with differentiable torch ops so first- and second-order autograd work.

It is validated bit-for-bit (same torch ops, same order) against the UNMODIFIED
reference running on the shims in tests/test_oracle_vs_reference.py (this
container only) and against the committed golden vectors everywhere.

PARITY UNPINNED at the third-party seams: nerfacc==0.3.1 and roma==1.2.7 are not
vendored in /root/reference and the reference ships no golden vectors; their
published algorithms are restated here and in oracle/shims.
"""
import math
from dataclasses import dataclass, field as dc_field
from typing import List, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
@dataclass
class FieldConfig:
    """Mirror of `VanillaNeRFRadianceField.__init__` (external/mlp.py:247-264) and
    the `model.nerf.mlp` YAML block (configs/train/synthetic.yaml:85-96)."""
    aabb: List[float] = dc_field(default_factory=lambda: [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
    contraction: str = "sphere"            # aabb | sphere | tanh
    net_depth: int = 8
    net_width: int = 256
    skip_layer: Optional[int] = 4
    net_depth_condition: int = 1
    net_width_condition: int = 128
    radiance_dim: int = 1
    hidden_activation: str = "softplus"     # softplus(beta=100) | relu  (models/nerf.py:17-20)
    density_activation: str = "shifted_trunc_exp"  # | softplus | shifted_softplus (nerf.py:21-25)
    radiance_activation: str = "softplus"   # | sigmoid (nerf.py:26-29)
    pos_deg: int = 10
    dir_deg: int = 4
    # `arch: ngp` (robust_e_nerf/external/ngp.py:109-152, configs/train/*.yaml:61-88): multiresolution hash grid
    # + SH directions + two small MLPs.  `hidden_activation` is mlp_base's, `head_hidden_activation` mlp_head's
    # (None = the same); net_* / pos_deg / dir_deg above are unused.
    arch: str = "mlp"                        # mlp | ngp
    n_levels: int = 16
    n_features_per_level: int = 2
    log2_hashmap_size: int = 19
    base_resolution: int = 16
    per_level_scale: float = 1.4472692012786865
    grid_type: str = "Hash"                  # Hash | Dense | Tiled  (otype HashGrid | DenseGrid | TiledGrid)
    sh_degree: int = 4
    base_n_neurons: int = 64
    base_n_hidden_layers: int = 1
    geo_feat_dim: int = 15
    head_n_neurons: int = 64
    head_n_hidden_layers: int = 2
    head_hidden_activation: Optional[str] = None

    def grid_meta(self):
        from . import hashgrid
        return hashgrid.grid_meta(self.n_levels, self.n_features_per_level, self.log2_hashmap_size,
                                  self.base_resolution, self.per_level_scale, self.grid_type)

    @property
    def pos_dim(self):
        return 3 * (1 + 2 * self.pos_deg)

    @property
    def dir_dim(self):
        return 3 * (1 + 2 * self.dir_deg)

    def layer_shapes(self):
        """[(name, out, in)] in the reference's state-dict order
        (external/mlp.py:56-76,143-186; `arch: ngp`: external/ngp.py:171-207 after the grid parameters)."""
        if self.arch == "ngp":
            shapes, in_f = [], self.n_levels * self.n_features_per_level
            for i in range(self.base_n_hidden_layers):
                shapes.append((f"mlp_base.1.hidden_layers.{i}", self.base_n_neurons, in_f))
                in_f = self.base_n_neurons
            shapes.append(("mlp_base.1.output_layer", 1 + self.geo_feat_dim, in_f))
            in_f = self.sh_degree ** 2 + self.geo_feat_dim
            for i in range(self.head_n_hidden_layers):
                shapes.append((f"mlp_head.hidden_layers.{i}", self.head_n_neurons, in_f))
                in_f = self.head_n_neurons
            shapes.append(("mlp_head.output_layer", self.radiance_dim, in_f))
            return shapes
        shapes = []
        in_f = self.pos_dim
        for i in range(self.net_depth):
            shapes.append((f"mlp.base.hidden_layers.{i}", self.net_width, in_f))
            if self.skip_layer is not None and i % self.skip_layer == 0 and i > 0:
                in_f = self.net_width + self.pos_dim
            else:
                in_f = self.net_width
        hidden = in_f
        shapes.append(("mlp.sigma_layer.output_layer", 1, hidden))
        shapes.append(("mlp.bottleneck_layer.output_layer", self.net_width, hidden))
        in_f = self.net_width + self.dir_dim
        for i in range(self.net_depth_condition):
            shapes.append((f"mlp.rgb_layer.hidden_layers.{i}", self.net_width_condition, in_f))
            in_f = self.net_width_condition
        shapes.append(("mlp.rgb_layer.output_layer", self.radiance_dim, in_f))
        return shapes


@dataclass
class RenderConfig:
    """Mirror of the `NeRF.__init__` rendering options (models/nerf.py:31-48)."""
    near_plane: Optional[float] = 0.5
    far_plane: Optional[float] = 4.5
    render_step_size: float = 4.0 / 128
    cone_angle: float = 0.0
    early_stop_eps: float = 0.0
    alpha_thre: float = 0.0
    opacity_eps: float = 1e-10
    stratified: bool = True


def field_param_prefix(cfg: FieldConfig):
    """State-dict prefix of the field's trainable tensors under `RobustENeRF`: `nerf.radiance_field.mlp.*` for
    the vanilla field (external/mlp.py:266-299), `nerf.radiance_field.mlp_base.*` / `.mlp_head.*` for ngp."""
    return "nerf.radiance_field.mlp_" if cfg.arch == "ngp" else "nerf.radiance_field.mlp."


def init_field_params(cfg: FieldConfig, seed=0, dtype=torch.float32):
    """Default `nn.Linear` init (the reference passes hidden_init/output_init/
    bias_init = None, external/mlp.py:297-299): kaiming_uniform(a=sqrt(5)) weights,
    U(+-1/sqrt(fan_in)) biases, drawn layer by layer in construction order so the
    result equals `torch.manual_seed(seed); VanillaNeRFRadianceField(...)`."""
    g_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    params = {}
    if cfg.arch == "ngp":
        # `tcnn.Encoding` is constructed first (external/ngp.py:166-170); the oracle shim draws its grid from
        # torch's global generator (oracle/shims/tinycudann.py), then the Linear layers follow in order
        from . import hashgrid
        params["mlp_base.0.params"] = hashgrid.init_params(cfg.grid_meta()).to(dtype)
    for name, out_f, in_f in cfg.layer_shapes():
        lin = torch.nn.Linear(in_f, out_f)
        params[name + ".weight"] = lin.weight.detach().to(dtype).clone()
        params[name + ".bias"] = lin.bias.detach().to(dtype).clone()
    torch.random.set_rng_state(g_state)
    return params


# ----------------------------------------------------------------------- activations
def softplus(x, beta=1.0, threshold=20.0):
    return F.softplus(x, beta, threshold)


class _TruncExp(torch.autograd.Function):
    """external/ngp.py:45-58: exp forward, backward uses exp(clamp(x, max=15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


def hidden_act(cfg, x, head=False):
    kind = cfg.hidden_activation
    if head and getattr(cfg, "head_hidden_activation", None) is not None:
        kind = cfg.head_hidden_activation
    if kind == "softplus":
        return softplus(x, 100.0)            # models/nerf.py:18
    return torch.relu(x)


def density_act(cfg, x):
    if cfg.density_activation == "shifted_trunc_exp":
        return _TruncExp.apply(x - 1)        # external/ngp.py:64-65
    if cfg.density_activation == "softplus":
        return softplus(x, 1.0)
    return softplus(x - 1, 1.0)              # models/nerf.py:8-13


def radiance_act(cfg, x):
    if cfg.radiance_activation == "softplus":
        return softplus(x, 1.0)
    return torch.sigmoid(x)


# ----------------------------------------------------------------------------- field
def contract_unit(cfg: FieldConfig, x):
    """World -> unit cube [0,1]^3 part of `contract_input_space` (also what nerfacc's grid lookup applies)."""
    return (contract_input_space(cfg, x)[0] / (2 * math.pi)) + 0.5


def contract_inv(cfg: FieldConfig, u):
    """Unit cube -> world: inverse of the above (nerfacc `contract_inv`, used by OccupancyGrid._update)."""
    aabb = torch.tensor(cfg.aabb, dtype=u.dtype)
    a_min, a_max = aabb[:3], aabb[3:]
    if cfg.contraction == "sphere":
        y = u * 4 - 2
        mag = y.norm(dim=-1, keepdim=True)
        y = torch.where(mag > 1, y / mag / (2 - mag), y)
        u = y * 0.5 + 0.5
    elif cfg.contraction == "tanh":
        u = torch.atanh((u * 2 - 1).clamp(-1 + 1e-6, 1 - 1e-6)) + 0.5
    return u * (a_max - a_min) + a_min


@torch.no_grad()
def grid_occupied(cfg: FieldConfig, grid_binary, x):
    """nerfacc `grid_occupied_at` (ray_marching kernel): cell = floor(unit position * resolution), clamped."""
    # same op order as the field input: normalise, contract (no 2 pi round trip)
    aabb = torch.tensor(cfg.aabb, dtype=x.dtype)
    a_min, a_max = aabb[:3], aabb[3:]
    u = (x - a_min) / (a_max - a_min)
    if cfg.contraction == "sphere":
        u = u * 2 - 1
        mag = u.norm(dim=-1, keepdim=True)
        u = torch.where(mag > 1, (2 - 1 / mag) * (u / mag), u)
        u = u / 4 + 0.5
    elif cfg.contraction == "tanh":
        u = (torch.tanh(u - 0.5) + 1) / 2
    res = torch.tensor(grid_binary.shape)
    idx = torch.floor(u * res).long()
    idx = torch.minimum(torch.maximum(idx, torch.zeros_like(idx)), res - 1)
    return grid_binary[idx[..., 0], idx[..., 1], idx[..., 2]]


def contract_input_space(cfg: FieldConfig, x):
    """external/mlp.py:321-335 + external/ngp.py:68-106."""
    aabb = torch.tensor(cfg.aabb, dtype=x.dtype)
    a_min, a_max = aabb[:3], aabb[3:]
    if cfg.contraction == "sphere":
        x = (x - a_min) / (a_max - a_min)
        x = x * 2 - 1
        mag = x.norm(dim=-1, keepdim=True)
        mask = mag > 1
        x = torch.where(mask, (2 - 1 / mag) * (x / mag), x)
        x = x / 4 + 0.5
    elif cfg.contraction == "tanh":
        x = (x - a_min) / (a_max - a_min)
        x = x - 0.5
        x = (torch.tanh(x) + 1) / 2
    else:
        x = (x - a_min) / (a_max - a_min)
    selector = ((x > 0.0) & (x < 1.0)).all(dim=-1)
    x = 2 * math.pi * (x - 0.5)
    return x, selector


def sinusoidal_encode(x, max_deg):
    """external/mlp.py:227-243 (min_deg 0, identity on; `scales` is int64)."""
    if max_deg == 0:
        return x
    scales = torch.tensor([2 ** i for i in range(0, max_deg)])
    xb = torch.reshape(x[..., None, :] * scales[:, None],
                       list(x.shape[:-1]) + [max_deg * x.shape[-1]])
    latent = torch.sin(torch.cat([xb, xb + 0.5 * math.pi], dim=-1))
    return torch.cat([x, latent], dim=-1)


def _linear(params, name, x):
    """`nn.Linear`, or its `torch.nn.utils.weight_norm` form (`weight_g`, `weight_v`: external/mlp.py:301-319,
    external/ngp.py:209-217) when the parameter dict carries those keys."""
    if name + ".weight_g" in params:
        w = torch._weight_norm(params[name + ".weight_v"], params[name + ".weight_g"], 0)
    else:
        w = params[name + ".weight"]
    return F.linear(x, w, params[name + ".bias"])


def mlp_trunk(params, cfg, enc):
    """external/mlp.py:99-113 with the skip concat *after* layer idx % skip == 0."""
    x = enc
    for i in range(cfg.net_depth):
        x = hidden_act(cfg, _linear(params, f"mlp.base.hidden_layers.{i}", x))
        if cfg.skip_layer is not None and i % cfg.skip_layer == 0 and i > 0:
            x = torch.cat([x, enc], dim=-1)
    return x


def contract_unit_selector(cfg: FieldConfig, x):
    """World -> [0,1]^3 and the in-box selector, as `NGPradianceField.query_density` applies them
    (external/ngp.py:219-228; same contraction formulas as the mlp field, external/ngp.py:68-106)."""
    aabb = torch.tensor(cfg.aabb, dtype=x.dtype)
    a_min, a_max = aabb[:3], aabb[3:]
    x = (x - a_min) / (a_max - a_min)
    if cfg.contraction == "sphere":
        x = x * 2 - 1
        mag = x.norm(dim=-1, keepdim=True)
        x = torch.where(mag > 1, (2 - 1 / mag) * (x / mag), x)
        x = x / 4 + 0.5
    elif cfg.contraction == "tanh":
        x = (torch.tanh(x - 0.5) + 1) / 2
    return x, ((x > 0.0) & (x < 1.0)).all(dim=-1)


def ngp_base(params, cfg, x):
    """`NGPradianceField.mlp_base` on world points (external/ngp.py:171-188,219-240): hash-grid features ->
    MLP -> [density before activation | geometry features], and the selector."""
    from . import hashgrid
    u, selector = contract_unit_selector(cfg, x)
    h = hashgrid.encode(u, params["mlp_base.0.params"], cfg.grid_meta())
    for i in range(cfg.base_n_hidden_layers):
        h = hidden_act(cfg, _linear(params, f"mlp_base.1.hidden_layers.{i}", h))
    out = _linear(params, "mlp_base.1.output_layer", h)
    return out[..., :1], out[..., 1:], selector


def ngp_forward(params, cfg, x, d):
    """`NGPradianceField.forward` (external/ngp.py:266-280): `_query_rgb` concatenates [SH(dir) | geometry
    features] (:246-252) and mlp_head applies the radiance activation as its output activation (:189-207)."""
    from . import hashgrid
    raw_sigma, geo, selector = ngp_base(params, cfg, x)
    h = torch.cat([hashgrid.sh_encode(d, cfg.sh_degree), geo], dim=-1)
    for i in range(cfg.head_n_hidden_layers):
        h = hidden_act(cfg, _linear(params, f"mlp_head.hidden_layers.{i}", h), head=True)
    raw_rgb = _linear(params, "mlp_head.output_layer", h)
    return radiance_act(cfg, raw_rgb), density_act(cfg, raw_sigma) * selector[..., None]


def field_density(params, cfg, x):
    """`VanillaNeRFRadianceField.query_density` (external/mlp.py:344-348) /
    `NGPradianceField.query_density` (external/ngp.py:219-244)."""
    if cfg.arch == "ngp":
        raw_sigma, _, selector = ngp_base(params, cfg, x)
        return density_act(cfg, raw_sigma) * selector[..., None]
    z, selector = contract_input_space(cfg, x)
    h = mlp_trunk(params, cfg, sinusoidal_encode(z, cfg.pos_deg))
    raw_sigma = _linear(params, "mlp.sigma_layer.output_layer", h)
    return density_act(cfg, raw_sigma) * selector[..., None]


def field_forward(params, cfg, x, d):
    """`VanillaNeRFRadianceField.forward` (external/mlp.py:350-358) +
    `NerfMLP.forward` (:193-205)."""
    if cfg.arch == "ngp":
        return ngp_forward(params, cfg, x, d)
    z, selector = contract_input_space(cfg, x)
    h = mlp_trunk(params, cfg, sinusoidal_encode(z, cfg.pos_deg))
    raw_sigma = _linear(params, "mlp.sigma_layer.output_layer", h)
    cond = sinusoidal_encode(d * math.pi, cfg.dir_deg)
    bott = _linear(params, "mlp.bottleneck_layer.output_layer", h)
    g = torch.cat([bott, cond], dim=-1)
    for i in range(cfg.net_depth_condition):
        g = hidden_act(cfg, _linear(params, f"mlp.rgb_layer.hidden_layers.{i}", g))
    raw_rgb = _linear(params, "mlp.rgb_layer.output_layer", g)
    return radiance_act(cfg, raw_rgb), density_act(cfg, raw_sigma) * selector[..., None]


# ------------------------------------------------------------------------- sampling
@torch.no_grad()
def ray_aabb_intersect(o, d, aabb):
    """nerfacc v0.3.1 slab test (miss -> 1e10, entry clamped at 0)."""
    t_lo = (aabb[:3] - o) / d
    t_hi = (aabb[3:] - o) / d
    t_near = torch.minimum(t_lo, t_hi).amax(dim=-1)
    t_far = torch.maximum(t_lo, t_hi).amin(dim=-1)
    miss = t_near > t_far
    t_near = torch.clamp(t_near, min=0.0)
    big = torch.full_like(t_near, 1e10)
    return torch.where(miss, big, t_near), torch.where(miss, big, t_far)


@torch.no_grad()
def march(cfg: FieldConfig, rcfg: RenderConfig, o, d, jitter):
    """Dense (n_rays, K) sample grid: t0, t1, alive mask.  nerfacc v0.3.1
    `ray_marching` over an all-occupied grid, external/utils.py:106-119."""
    if cfg.contraction == "aabb":                                  # models/nerf.py:248-251
        t_min, t_max = ray_aabb_intersect(o, d, torch.tensor(cfg.aabb, dtype=o.dtype))
    else:
        t_min = torch.zeros_like(o[..., 0])
        t_max = torch.ones_like(o[..., 0]) * 1e10
    if rcfg.near_plane is not None:
        t_min = torch.clamp(t_min, min=rcfg.near_plane)
    if rcfg.far_plane is not None:
        t_max = torch.clamp(t_max, max=rcfg.far_plane)
    step = torch.tensor(rcfg.render_step_size, dtype=torch.float32).to(o.dtype)
    if rcfg.stratified:
        t_min = t_min + jitter.to(o.dtype) * step
    t0 = t_min.clone()
    t1 = t0 + torch.clamp(t0 * rcfg.cone_angle, min=step, max=1e10)
    t0s, t1s, alive = [], [], []
    while True:
        a = (t0 + t1) * 0.5 < t_max
        if not bool(a.any()):
            break
        t0s.append(t0)
        t1s.append(t1)
        alive.append(a)
        t0 = t1
        t1 = t0 + torch.clamp(t0 * rcfg.cone_angle, min=step, max=1e10)
    if not t0s:
        z = o.new_zeros((o.shape[0], 0))
        return z, z, z.bool()
    return torch.stack(t0s, 1), torch.stack(t1s, 1), torch.stack(alive, 1)


def sample_positions(o, d, t0, t1):
    """external/utils.py:87: `t_origins + t_dirs * (t_starts + t_ends) / 2.0`
    == o + ((d * (t0 + t1)) / 2), each op rounded separately."""
    return o[:, None, :] + d[:, None, :] * (t0 + t1)[..., None] / 2.0


@torch.no_grad()
def visibility(alpha, alive, early_stop_eps, alpha_thre):
    """nerfacc v0.3.1 `render_visibility` on the dense grid."""
    T = torch.ones_like(alpha[:, 0])
    stopped = torch.zeros_like(alive[:, 0])
    vis = torch.zeros_like(alive)
    for k in range(alpha.shape[1]):
        stopped = stopped | (T < early_stop_eps)
        a = alpha[:, k]
        keep = alive[:, k] & ~stopped & ~(a < alpha_thre)
        vis[:, k] = keep
        T = torch.where(keep, T * (1.0 - a), T)
    return vis


# ------------------------------------------------------------------------ rendering
def render(params, cfg: FieldConfig, rcfg: RenderConfig, o, d, jitter, bkgd=None,
           return_aux=False, grid_binary=None):
    """`NeRF.forward` (models/nerf.py:230-286) -> `render_image`
    (external/utils.py:38-140) -> `rendering` (external/vol_rendering.py:82-126).
    Returns radiance (B[,C]), opacity (B), depth (B), mean samples per ray."""
    B = o.shape[0]
    t0, t1, alive = march(cfg, rcfg, o.detach(), d.detach(), jitter)
    K = t0.shape[1]
    if grid_binary is not None and K > 0:
        # marching steps whose midpoint lies in an unoccupied cell emit no sample.  sphere / tanh grids advance by
        # fixed steps; for aabb nerfacc's advance_to_next_voxel moves the midpoint by whole steps until it has left
        # the empty voxel -- either way the surviving samples stay on the same lattice.
        alive = alive & grid_occupied(cfg, grid_binary, sample_positions(o.detach(), d.detach(), t0, t1))
    if K == 0:
        keep = alive
    else:
        with torch.no_grad():                                      # utils.py:68-81 (no-grad pass)
            pos = sample_positions(o.detach(), d.detach(), t0, t1)
            sig = field_density(params, cfg, pos.reshape(-1, 3)).reshape(B, K)
            alpha = 1.0 - torch.exp(-sig * (t1 - t0))
            keep = visibility(alpha, alive, rcfg.early_stop_eps, rcfg.alpha_thre)
    # packed, ray-major samples exactly as nerfacc hands them over
    ray_idx = torch.arange(B)[:, None].expand(B, K)[keep]
    ts, te = t0[keep][:, None], t1[keep][:, None]
    n_samples = int(ray_idx.numel())
    positions = o[ray_idx] + d[ray_idx] * (ts + te) / 2.0          # utils.py:85-87
    rgb, sigma = field_forward(params, cfg, positions, d[ray_idx])
    # weights: exclusive segmented sum of sigma*delta, then exp (vol_rendering.py:89-95)
    sdt = sigma * (te - ts)
    dense = sdt.new_zeros((B, K)).index_put(torch.nonzero(keep, as_tuple=True), sdt[:, 0])
    excl = (torch.cumsum(dense, dim=1) - dense)[keep][:, None]
    w = torch.exp(-excl) * (1.0 - torch.exp(-sdt))
    C = cfg.radiance_dim
    colors = rgb.new_zeros((B, C)).index_add(0, ray_idx, w * rgb)   # vol_rendering.py:111-122
    opac = rgb.new_zeros((B, 1)).index_add(0, ray_idx, w)
    depth = rgb.new_zeros((B, 1)).index_add(0, ray_idx, w * ((ts + te) / 2.0))
    if bkgd is not None:
        colors = colors + bkgd * (1.0 - opac)                       # vol_rendering.py:125-126
    radiance = colors.squeeze(-1)
    opacity = opac.squeeze(-1)
    depth = depth.squeeze(-1) / (opacity + rcfg.opacity_eps)        # nerf.py:279-284
    mean_spp = n_samples / B
    if return_aux:
        return radiance, opacity, depth, mean_spp, {"t0": t0, "t1": t1, "keep": keep,
                                                    "weights": w, "sigma": sigma, "rgb": rgb}
    return radiance, opacity, depth, mean_spp


# ------------------------------------------------------------------------- poses
def lerp(a, b, w):
    return torch.lerp(a, b, w)


def quat_product(p, q):
    v = (p[..., None, 3] * q[..., :3] + q[..., None, 3] * p[..., :3]
         + torch.cross(p[..., :3], q[..., :3], dim=-1))
    s = p[..., 3] * q[..., 3] - torch.sum(p[..., :3] * q[..., :3], dim=-1)
    return torch.cat((v, s[..., None]), dim=-1)


def unitquat_to_full_rotvec(q):
    """utils/tensor_ops.py:83-111."""
    angle = 2 * torch.atan2(torch.norm(q[:, :3], dim=1), q[:, 3])
    small = torch.abs(angle) <= 1e-3
    scale = torch.empty_like(angle)
    scale[small] = 2 + angle[small] ** 2 / 12 + 7 * angle[small] ** 4 / 2880
    scale[~small] = angle[~small] / torch.sin(angle[~small] / 2)
    return scale[:, None] * q[:, :3]


def rotvec_to_unitquat(rv):
    """roma 1.2.7 `rotvec_to_unitquat`."""
    n = torch.norm(rv, dim=-1)
    small = n <= 1e-3
    scale = torch.empty_like(n)
    scale[small] = 0.5 - n[small] ** 2 / 48 + n[small] ** 4 / 3840
    scale[~small] = torch.sin(n[~small] / 2) / n[~small]
    return torch.cat([scale[:, None] * rv, torch.cos(n / 2)[:, None]], dim=-1)


def unitquat_slerp(q0, q1, steps):
    """utils/tensor_ops.py:114-180 with shortest_path=True, one step per pair."""
    q1 = torch.where(torch.sum(q0 * q1, dim=-1, keepdim=True) < 0, -q1, q1)
    conj = q0.clone()
    conj[..., :3] *= -1
    rel_q = quat_product(conj, q1)
    rel_rv = unitquat_to_full_rotvec(rel_q)
    rots = rotvec_to_unitquat(steps.unsqueeze(-1) * rel_rv)
    return quat_product(q0, rots)


def unitquat_to_rotmat(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    rows = [
        torch.stack([x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)], -1),
        torch.stack([2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)], -1),
        torch.stack([2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2], -1),
    ]
    return torch.stack(rows, -2)


def trajectory(scene, ts):
    """`LinearTrajectory.forward` (models/trajectories.py:30-91).  `ts` f64 (B,).
    Pose tables: T (C,) i64, p (C,3), q (C,4) XYZW in scene dtype."""
    T = scene["T_wc_timestamp"]
    P = scene["T_wc_position"]
    Q = scene["T_wc_orientation"]
    right = torch.searchsorted(T, ts.detach())
    corner = ts.detach() == T[0]
    left = torch.where(corner, right, right - 1)
    assert bool(((left >= 0) & (right < len(T))).all())
    bin_width = T.diff()
    w = (ts - T[left]) / bin_width[left]
    w = w.to(P.dtype)
    p = lerp(P[left], P[right], w.unsqueeze(-1))
    q = unitquat_slerp(Q[left], Q[right], w)
    return p, unitquat_to_rotmat(q)


def pixel_params_to_ray(K_inv, pix, p, R):
    """`NeRF.pixel_params_to_ray` (models/nerf.py:206-228)."""
    h = torch.cat([pix, torch.ones_like(pix[..., :1])], dim=-1).unsqueeze(-1)
    dvec = (R @ (K_inv @ h)).squeeze(-1)
    dvec = dvec / torch.linalg.vector_norm(dvec, dim=-1, keepdim=True)
    return p, dvec


def render_pixels(params, cfg, rcfg, K_inv, pix, p, R, jitter, bkgd, min_modeled_intensity):
    """`RobustENeRF.render_pixels` (models/robust_e_nerf.py:849-885)."""
    o, d = pixel_params_to_ray(K_inv, pix, p, R)
    rad, opac, depth, spp = render(params, cfg, rcfg, o, d, jitter, bkgd)
    intensity = rad + min_modeled_intensity
    if bkgd is None:
        valid = opac > 0
    else:
        valid = torch.ones_like(opac, dtype=torch.bool)
    depth = depth * torch.sum(d * R[..., 2], dim=-1)
    return intensity, opac, depth, spp, valid


# --------------------------------------------------------------------- training step
@dataclass
class StepConfig:
    w_grad: float = 0.0          # loss.weight.log_intensity_grad
    w_diff: float = 1.0          # loss.weight.log_intensity_diff
    err_grad: str = "mape"       # loss.error_fn.*
    err_diff: str = "mse"
    pw_grad: Optional[str] = None                       # loss.param_weight.*
    pw_diff: Optional[str] = "mean_contrast_reciprocal_sq"
    min_modeled_intensity: float = 1e-3


def _err(kind, inp, tgt):
    if kind == "l1":
        return (inp - tgt).abs()
    if kind == "mse":
        return (inp - tgt) ** 2
    eps = 2.220446049250313e-16                                   # utils/modules.py:86
    return (inp - tgt).abs() / tgt.abs().clamp(min=eps)           # MAPE, modules.py:96


def training_step(params, cfg: FieldConfig, rcfg: RenderConfig, scfg: StepConfig, scene,
                  batch, jitter, theta_ratio, theta_tau, theta_bkgd=None):
    """`RobustENeRF.training_step` (models/robust_e_nerf.py:301-517) for the
    monochrome case.  `theta_ratio` (f32 scalar): raw p2n contrast-threshold-ratio
    parameter; `theta_tau` (f64 scalar): raw refractory-period parameter;
    `theta_bkgd` (C,) raw background parameter or None.  Returns a dict of the
    scalars the reference logs plus per-ray intermediates."""
    ev = {k: v.squeeze(0) for k, v in batch["event"].items()}
    nz = {k: v.squeeze(0) for k, v in batch["normalized"].items()}
    fdt = next(iter(params.values())).dtype
    out = {}
    # a2: contrast threshold (models/event_generation_params.py:59-84)
    C_n = scene["neg_contrast_threshold"].to(fdt)
    ratio = softplus(theta_ratio, 1.0)
    C_p = ratio * C_n
    C_mean = (C_p + C_n) / 2
    ev_diff = ev["num_pos"] * C_p - ev["num_neg"] * C_n
    # a3: refractory period (event_generation_params.py:170-203, modules.py:58-74)
    tau_max = scene["max_refractory_period"]
    with torch.no_grad():
        lim = torch.tensor(1e-4).logit().abs()
        theta_tau.copy_(tau_max * (theta_tau / tau_max).clamp(min=-lim, max=lim))
    tau = tau_max * torch.sigmoid(theta_tau / tau_max) + 0
    start_ts = ev["start_ts"] + tau
    end_ts = ev["end_ts"]
    # a4: supervision timestamps (robust_e_nerf.py:322-357)
    if scfg.w_diff > 0:
        ts_diff = (end_ts - start_ts) * nz["ts_diff"]
        d_start = torch.lerp(start_ts, torch.max(end_ts - ts_diff, start_ts), nz["diff_start_ts"])
        d_end = torch.min(d_start + ts_diff, end_ts)
    if scfg.w_grad > 0:
        s, e = (start_ts, end_ts) if scfg.w_diff == 0 else (d_start, d_end)
        g_ts = torch.lerp(s, e, nz["grad_ts"])
        if not g_ts.requires_grad:
            g_ts.requires_grad_()
    K_inv = torch.linalg.inv(scene["intrinsics"])                  # robust_e_nerf.py:78-84
    bkgd = None if theta_bkgd is None else softplus(theta_bkgd, 1.0)
    pix = ev["position"].to(K_inv.dtype)      # no-op in fp32; fp64 noise-floor runs promote
    spps = []
    losses = {}
    rate = ev_diff / (end_ts - start_ts)                           # loss.py:38-41
    if scfg.w_grad > 0:
        p, R = trajectory(scene, g_ts)
        I, opac, _, spp, valid = render_pixels(params, cfg, rcfg, K_inv, pix, p, R,
                                               jitter["grad"], bkgd, scfg.min_modeled_intensity)
        logI = I.log()
        g = torch.autograd.grad(logI, g_ts, torch.ones_like(logI), create_graph=True)[0]
        spps.append(spp)
        out.update(grad_log_intensity=logI, log_intensity_grad=g, grad_valid=valid,
                   grad_opacity=opac)
        losses["log_intensity_grad"] = _err(scfg.err_grad, g, rate)[valid].mean()
    if scfg.w_diff > 0:
        p0, R0 = trajectory(scene, d_start)
        p1, R1 = trajectory(scene, d_end)
        I0, op0, _, spp0, v0 = render_pixels(params, cfg, rcfg, K_inv, pix, p0, R0,
                                             jitter["diff_start"], bkgd, scfg.min_modeled_intensity)
        I1, op1, _, spp1, v1 = render_pixels(params, cfg, rcfg, K_inv, pix, p1, R1,
                                             jitter["diff_end"], bkgd, scfg.min_modeled_intensity)
        L0, L1 = I0.log(), I1.log()
        dL = L1 - L0
        valid = v0 | v1
        spps += [spp0, spp1]
        tgt = (ts_diff * rate).to(dL.dtype)
        out.update(start_log_intensity=L0, end_log_intensity=L1, diff_valid=valid,
                   diff_target=tgt, start_opacity=op0, end_opacity=op1)
        losses["log_intensity_diff"] = _err(scfg.err_diff, dL, tgt)[valid].mean()
    mcr = 1 / C_mean
    pw = {None: 1.0, "mean_contrast_reciprocal": mcr, "mean_contrast_reciprocal_sq": mcr ** 2}
    weights = {"log_intensity_grad": scfg.w_grad, "log_intensity_diff": scfg.w_diff}
    pws = {"log_intensity_grad": scfg.pw_grad, "log_intensity_diff": scfg.pw_diff}
    total = 0
    for name, val in losses.items():
        losses[name] = pw[pws[name]] * val
        total = total + losses[name] * weights[name]
    out["loss"] = total
    out["losses"] = losses
    out["mean_num_samples_per_ray"] = sum(spps) / len(spps)
    return out
