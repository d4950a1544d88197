"""`NeRF` -- drop-in for the reference's `robust_e_nerf.models.nerf.NeRF`
(robust_e_nerf/models/nerf.py:16-286) for `arch: mlp` (tcgen05 engine) and `arch: ngp` (grid encoding + small
MLPs on the fp32 engine), backed by libren_b200.so.

Same constructor arguments, same `forward(ray_origin, ray_direction)` ->
`(radiance, opacity, depth, mean_num_samples_per_ray)`, same `update_occ_grid`,
`pixel_params_to_ray` and state-dict keys
(`radiance_field.mlp.base.hidden_layers.{i}.{weight,bias}`, ... -- SURVEY.md section 5), so
`RobustENeRF._build_nerf` (robust_e_nerf/models/robust_e_nerf.py:204-260) can return it.

All MLP parameters live in ONE flat fp32 buffer (the order `ren_param_layout`
documents); the named `nn.Parameter`s are views into it, so the C ABI gets a single
pointer, the gradient is a single vector (one NCCL all-reduce) and Adam is one kernel.
"""
import ctypes as C
import math

import torch

from . import _lib


def _contraction_name(contraction_type):
    if isinstance(contraction_type, str):
        return contraction_type
    name = getattr(contraction_type, "name", str(contraction_type))      # nerfacc.ContractionType
    return {"AABB": "aabb", "UN_BOUNDED_SPHERE": "sphere", "UN_BOUNDED_TANH": "tanh"}[name]


def _cfg_get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class _Holder(torch.nn.Module):
    """Empty container used to reproduce the reference's module tree / key names."""


class _Encoder(torch.nn.Module):
    def __init__(self, max_deg):
        super().__init__()
        # external/mlp.py:217-219: an int64 buffer that is part of the state dict
        self.register_buffer("scales", torch.tensor([2 ** i for i in range(0, max_deg)]))


class OccupancyGrid(torch.nn.Module):
    """`nerfacc.OccupancyGrid` (nerfacc==0.3.1; constructed at robust_e_nerf/models/nerf.py:98-102, updated at
    :200-204): EMA-max occupancy estimate per cell of the contracted unit cube + its thresholded binary image,
    which the marcher consults (external/utils.py:106-119).  State-dict keys `_roi_aabb`, `resolution`, `occs`,
    `_binary` as in nerfacc.  Everything is device-side torch; the density query is `ren_field_density`."""

    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution=128, contraction="aabb"):
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * self.NUM_DIM
        self.contraction = contraction
        self.register_buffer("_roi_aabb", torch.as_tensor(roi_aabb, dtype=torch.float32))
        self.register_buffer("resolution", torch.as_tensor(resolution, dtype=torch.int32))
        self.num_cells = int(torch.as_tensor(resolution).prod())
        self.register_buffer("occs", torch.zeros(self.num_cells))
        self.register_buffer("_binary", torch.zeros(list(resolution), dtype=torch.bool))
        self._u8 = None
        self._u8_version = None

    @property
    def binary(self):
        return self._binary

    def binary_u8(self):
        """Contiguous uint8 image of `_binary` for the kernels (re-made when `_binary` changed)."""
        key = (self._binary.data_ptr(), self._binary._version, self._binary.device)
        if self._u8 is None or self._u8_version != key:
            self._u8 = self._binary.to(torch.uint8).contiguous()
            self._u8_version = key
        return self._u8

    def cell_coords(self, indices):
        ry, rz = int(self.resolution[1]), int(self.resolution[2])
        return torch.stack((indices // (ry * rz), (indices // rz) % ry, indices % rz), dim=-1)

    def contract_inv(self, x):
        """Unit cube -> world (inverse of the field's input contraction, external/ngp.py:68-106)."""
        a_min, a_max = self._roi_aabb[:3], self._roi_aabb[3:]
        if self.contraction == "sphere":
            y = x * 4 - 2
            mag = y.norm(dim=-1, keepdim=True)
            y = torch.where(mag > 1, y / mag / (2 - mag).clamp(min=1e-12), y)
            x = y * 0.5 + 0.5
        elif self.contraction == "tanh":
            x = torch.atanh((x * 2 - 1).clamp(-1 + 1e-6, 1 - 1e-6)) + 0.5
        return x * (a_max - a_min) + a_min

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n, generator=None):
        dev = self.occs.device
        uniform = torch.randint(self.num_cells, (n,), device=dev, generator=generator)
        occupied = torch.nonzero(self._binary.flatten())[:, 0]
        if n < len(occupied):
            occupied = occupied[torch.randint(len(occupied), (n,), device=dev, generator=generator)]
        return torch.cat([uniform, occupied], dim=0)

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre=0.01, ema_decay=0.95, warmup_steps=256, generator=None):
        dev = self.occs.device
        if step < warmup_steps:
            indices = torch.arange(self.num_cells, device=dev)
        else:
            indices = self._sample_uniform_and_occupied_cells(self.num_cells // 4, generator)
        coords = self.cell_coords(indices)
        jitter = torch.rand(coords.shape, device=dev, dtype=torch.float32, generator=generator)
        x = (coords + jitter) / self.resolution
        if self.contraction == "sphere":                 # only the inscribed ball of the unit cube is ever reached
            mask = (x - 0.5).norm(dim=1) < 0.5
            x, indices = x[mask], indices[mask]
        occ = occ_eval_fn(self.contract_inv(x)).reshape(-1)
        self.occs[indices] = torch.maximum(self.occs[indices] * ema_decay, occ)
        self._binary = (self.occs > torch.clamp(self.occs.mean(), max=occ_thre)).view(self._binary.shape)

    @torch.no_grad()
    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, generator=None):
        if not self.training:
            raise RuntimeError("OccupancyGrid.every_n_step() is a training-time call")
        if step % n == 0:
            self._update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps, generator)


class RenderFunction(torch.autograd.Function):
    """ren_render_forward / ren_render_backward as one differentiable op.

    inputs : owner (NeRF), want_grad (bool: grad mode of the caller), o (B,3), d (B,3),
             jitter (B)|None, bkgd (C)|None, *params
    outputs: radiance (B,C), opacity (B), depth_raw (B), n_kept (B) int32
    Gradients flow to the MLP parameters and `bkgd`; when `o` / `d` require grad (trainable
    refractory period) the reverse pass also returns d loss / d (o, d).
    """

    @staticmethod
    def forward(ctx, owner, want_grad, o, d, jitter, bkgd, *params):
        lib = _lib.load()
        if not o.is_cuda:
            raise RuntimeError("ren_b200 renders on CUDA tensors only (no CPU fallback)")
        B = o.shape[0]
        C_ = owner.radiance_dim
        ray_grad = bool(want_grad) and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        fdesc, rdesc, K = owner._descs()
        flat = owner._flat_for(params)
        need_grad = bool(want_grad) and (ray_grad or any(ctx.needs_input_grad[5:]))
        dev = o.device
        o = o.contiguous().float()
        d = d.contiguous().float()
        if jitter is not None:
            jitter = jitter.contiguous().float()
        if bkgd is not None:
            bkgd = bkgd.contiguous().float()
        radiance = torch.empty((B, C_), device=dev, dtype=torch.float32)
        opacity = torch.empty((B,), device=dev, dtype=torch.float32)
        depth = torch.empty((B,), device=dev, dtype=torch.float32)
        n_kept = torch.empty((B,), device=dev, dtype=torch.int32)
        # occupancy grid on the tcgen05 engine: march first, run the field on the emitted samples only
        packed, keep = owner._march_pack(fdesc, rdesc, B, K, o, d, jitter)
        nS, nK = (max(int(packed.n_samples), 1), 1) if packed is not None else (B, K)
        stash = None
        stash_n = 0
        if need_grad:
            stash_n = _lib.stash_bytes(fdesc, rdesc, nS, nK)
            stash = torch.empty((stash_n,), device=dev, dtype=torch.uint8)
        ws_n = _lib.workspace_bytes(fdesc, rdesc, nS, nK, need_grad, False)
        ws = torch.empty((ws_n,), device=dev, dtype=torch.uint8)
        _lib.check(lib.ren_render_forward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o), _lib.ptr(d), _lib.ptr(jitter),
            _lib.ptr(flat), _lib.ptr(bkgd), _lib.ptr(radiance), _lib.ptr(opacity), _lib.ptr(depth),
            _lib.ptr(n_kept), None if packed is None else C.byref(packed), _lib.ptr(stash), stash_n, _lib.ptr(ws),
            ws.numel(), _lib.stream_ptr()), "ren_render_forward")
        if need_grad:
            ctx.owner = owner
            ctx.descs = (fdesc, rdesc, K)
            ctx.packed, ctx.packed_keep = packed, keep
            ctx.save_for_backward(o, d, jitter if jitter is not None else torch.empty(0, device=dev),
                                  bkgd if bkgd is not None else torch.empty(0, device=dev), stash,
                                  flat, *params)
            ctx.has_jitter = jitter is not None
            ctx.has_bkgd = bkgd is not None
            ctx.n_params = len(params)
            ctx.ray_grad = ray_grad
        ctx.mark_non_differentiable(n_kept)
        return radiance, opacity, depth, n_kept

    @staticmethod
    def backward(ctx, g_rad, g_op, g_depth, _g_kept):
        if torch.is_grad_enabled():
            # create_graph=True: the caller will differentiate this gradient again -- the reference does so for
            # d log I / d t (utils/autograd.py:4-34, models/robust_e_nerf.py:394-398).  The gradient with respect
            # to the rays then becomes a differentiable op of its own (RenderRayGradFunction).
            o, d, jitter, bkgd = ctx.saved_tensors[:4]
            params = ctx.saved_tensors[6:]
            if ctx.packed is not None:
                raise RuntimeError("ren_b200: create_graph=True through the packed occupancy-grid march is not "
                                   "supported; use RobustENeRF.training_step (tangent render)")
            res = RenderRayGradFunction.apply(ctx.owner, ctx.descs, o, d, jitter if ctx.has_jitter else None,
                                              bkgd if ctx.has_bkgd else None, g_rad, g_op, g_depth, *params)
            # (g_o, g_d) carry a graph; the parameter / background gradients are returned as plain values
            return (None, None, res[0], res[1], None, res[2] if ctx.has_bkgd else None) + tuple(res[3:])
        lib = _lib.load()
        o, d, jitter, bkgd, stash, flat = ctx.saved_tensors[:6]
        fdesc, rdesc, K = ctx.descs
        owner = ctx.owner
        B = o.shape[0]
        dev = o.device
        jitter = jitter if ctx.has_jitter else None
        bkgd = bkgd if ctx.has_bkgd else None
        g_flat = torch.zeros_like(flat)
        g_bkgd = torch.zeros_like(bkgd) if bkgd is not None else None
        packed = ctx.packed
        nS, nK = (max(int(packed.n_samples), 1), 1) if packed is not None else (B, K)
        ws_n = _lib.workspace_bytes(fdesc, rdesc, nS, nK, True, True)
        ws = torch.empty((ws_n,), device=dev, dtype=torch.uint8)
        cg = lambda t: None if t is None else t.contiguous().float()
        g_rad, g_op, g_depth = cg(g_rad), cg(g_op), cg(g_depth)
        g_rays = torch.empty((B, 6), device=dev, dtype=torch.float32) if ctx.ray_grad else None
        _lib.check(lib.ren_render_backward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o), _lib.ptr(d), _lib.ptr(jitter),
            _lib.ptr(flat), _lib.ptr(bkgd), _lib.ptr(g_rad), _lib.ptr(g_op), _lib.ptr(g_depth),
            None if packed is None else C.byref(packed), _lib.ptr(stash), stash.numel(), _lib.ptr(g_flat),
            _lib.ptr(g_bkgd), _lib.ptr(g_rays), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "ren_render_backward")
        grads = owner._split_flat(g_flat)
        g_o = g_rays[:, 0:3] if ctx.ray_grad else None
        g_d = g_rays[:, 3:6] if ctx.ray_grad else None
        return (None, None, g_o, g_d, None, g_bkgd) + tuple(grads)


class RenderRayGradFunction(torch.autograd.Function):
    """d (sum_rays g_rad . C + g_op O) / d (o, d) as a differentiable op -- what `autograd.grad(log_intensity, ts,
    create_graph=True)` needs from the render when the reference's own `training_step` runs on `ren_b200.NeRF`
    (the NeRF-level plug of INTEGRATION.md section 1 with the log-intensity-gradient loss on).

    forward : one ordinary reverse pass (`ren_render_backward` with ray gradients).
    backward: for cotangents (v_o, v_d) of (g_o, g_d), the scalar  sum v_o . g_o + v_d . g_d  is the directional
              derivative of  sum g_rad . C + g_op O  along (o_dot, d_dot) = (v_o, v_d), i.e.
              sum g_rad . C_dot + g_op O_dot  of the TANGENT render: its gradient with respect to the parameters and
              rays is `ren_render_tangent_backward` with (g_rad, g_op) as the upstream of (C_dot, O_dot), and with
              respect to (g_rad, g_op) it is (C_dot, O_dot) itself.  No second-order kernel is needed.
    inputs  : owner, descs, o, d, jitter|None, bkgd|None, g_rad (B,C), g_op (B)|None, g_depth|None, *params
    outputs : g_o, g_d (B,3) [differentiable], then g_bkgd and the parameter gradients as plain values"""

    @staticmethod
    def forward(ctx, owner, descs, o, d, jitter, bkgd, g_rad, g_op, g_depth, *params):
        lib = _lib.load()
        fdesc, rdesc, K = descs
        if g_depth is not None and bool((g_depth != 0).any()):
            raise RuntimeError("ren_b200: create_graph=True with a depth gradient is not supported")
        B, dev = o.shape[0], o.device
        flat = owner._flat_for(params)
        cg = lambda t: None if t is None else t.detach().contiguous().float()
        o_, d_, g_rad_, g_op_ = cg(o), cg(d), cg(g_rad), cg(g_op)
        jit, bk = cg(jitter), cg(bkgd)
        # value pass with a stash, then its reverse with ray gradients
        C_ = owner.radiance_dim
        outs = [torch.empty((B, C_), device=dev), torch.empty((B,), device=dev), torch.empty((B,), device=dev)]
        n_kept = torch.empty((B,), device=dev, dtype=torch.int32)
        stash_n = _lib.stash_bytes(fdesc, rdesc, B, K)
        stash = torch.empty((stash_n,), device=dev, dtype=torch.uint8)
        ws = torch.empty((_lib.workspace_bytes(fdesc, rdesc, B, K, True, False),), device=dev, dtype=torch.uint8)
        _lib.check(lib.ren_render_forward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o_), _lib.ptr(d_), _lib.ptr(jit), _lib.ptr(flat),
            _lib.ptr(bk), _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(n_kept), None,
            _lib.ptr(stash), stash_n, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "ren_render_forward")
        g_flat = torch.zeros_like(flat)
        g_bk = torch.zeros_like(bk) if bk is not None else None
        g_rays = torch.empty((B, 6), device=dev, dtype=torch.float32)
        ws = torch.empty((_lib.workspace_bytes(fdesc, rdesc, B, K, True, True),), device=dev, dtype=torch.uint8)
        _lib.check(lib.ren_render_backward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o_), _lib.ptr(d_), _lib.ptr(jit), _lib.ptr(flat),
            _lib.ptr(bk), _lib.ptr(g_rad_), _lib.ptr(g_op_), None, None, _lib.ptr(stash), stash.numel(),
            _lib.ptr(g_flat), _lib.ptr(g_bk), _lib.ptr(g_rays), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
            "ren_render_backward")
        ctx.owner, ctx.descs = owner, descs
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(o_, d_, jit if jit is not None else empty, bk if bk is not None else empty,
                              g_rad_, g_op_ if g_op_ is not None else empty)
        ctx.flags = (jit is not None, bk is not None, g_op_ is not None)
        ctx.flat = flat
        ctx.n_params = len(params)
        plain = (g_bk if g_bk is not None else torch.zeros(0, device=dev),) + tuple(owner._split_flat(g_flat))
        ctx.mark_non_differentiable(*plain)
        return (g_rays[:, 0:3].contiguous(), g_rays[:, 3:6].contiguous()) + plain

    @staticmethod
    def backward(ctx, v_o, v_d, *_plain):
        _lib.refuse_double_backward()
        lib = _lib.load()
        owner = ctx.owner
        fdesc, rdesc, K = ctx.descs
        o, d, jit, bk, g_rad, g_op = ctx.saved_tensors
        has_jit, has_bk, has_gop = ctx.flags
        jit = jit if has_jit else None
        bk = bk if has_bk else None
        g_op = g_op if has_gop else None
        B, dev = o.shape[0], o.device
        C_ = owner.radiance_dim
        zero3 = lambda t: torch.zeros((B, 3), device=dev) if t is None else t.contiguous().float()
        v_o, v_d = zero3(v_o), zero3(v_d)
        # everything below is linear in each ray's (v_o, v_d): run the tangent kernels on per-ray directions of
        # unit size (the tcgen05 engine keeps tangent activations in fp16) and move the scale into the upstream
        # gradients / back onto the outputs
        scale = torch.maximum(v_o.abs().amax(dim=1), v_d.abs().amax(dim=1)).clamp(min=1e-30)
        v_o, v_d = v_o / scale[:, None], v_d / scale[:, None]
        g_rad = (g_rad * scale[:, None]).contiguous()
        if has_gop:
            g_op = (g_op * scale).contiguous()
        flat = ctx.flat
        rad, rad_dot = torch.empty((B, C_), device=dev), torch.empty((B, C_), device=dev)
        opac, opac_dot, depth = (torch.empty((B,), device=dev) for _ in range(3))
        n_kept = torch.empty((B,), device=dev, dtype=torch.int32)
        stash_n = _lib.tangent_stash_bytes(fdesc, rdesc, B, K)
        stash = torch.empty((stash_n,), device=dev, dtype=torch.uint8)
        ws = torch.empty((_lib.tangent_workspace_bytes(fdesc, rdesc, B, K, True, False),), device=dev, dtype=torch.uint8)
        _lib.check(lib.ren_render_tangent_forward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o), _lib.ptr(d), _lib.ptr(v_o), _lib.ptr(v_d),
            _lib.ptr(jit), _lib.ptr(flat), _lib.ptr(bk), _lib.ptr(rad), _lib.ptr(opac), _lib.ptr(depth),
            _lib.ptr(rad_dot), _lib.ptr(opac_dot), _lib.ptr(n_kept), None, _lib.ptr(stash), stash_n, _lib.ptr(ws),
            ws.numel(), _lib.stream_ptr()), "ren_render_tangent_forward")
        g_flat = torch.zeros_like(flat)
        g_bk = torch.zeros_like(bk) if bk is not None else None
        g_rays = torch.empty((B, 12), device=dev, dtype=torch.float32)
        ws = torch.empty((_lib.tangent_workspace_bytes(fdesc, rdesc, B, K, True, True),), device=dev, dtype=torch.uint8)
        _lib.check(lib.ren_render_tangent_backward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o), _lib.ptr(d), _lib.ptr(v_o), _lib.ptr(v_d),
            _lib.ptr(jit), _lib.ptr(flat), _lib.ptr(bk), None, None, None, _lib.ptr(g_rad), _lib.ptr(g_op), None,
            _lib.ptr(stash), stash.numel(), _lib.ptr(g_flat), _lib.ptr(g_bk), _lib.ptr(g_rays), _lib.ptr(ws),
            ws.numel(), _lib.stream_ptr()), "ren_render_tangent_backward")
        grads = owner._split_flat(g_flat)
        return (None, None, g_rays[:, 0:3], g_rays[:, 3:6], None, g_bk, rad_dot * scale[:, None],
                opac_dot * scale if has_gop else None, None) + tuple(grads)


class RenderTangentFunction(torch.autograd.Function):
    """ren_render_tangent_forward / _backward: the render and its derivative along the ray's pose
    interpolation weight w (forward mode), differentiable once more in reverse mode -- what the
    reference obtains with `autograd.gradient(..., create_graph=True)` followed by `backward()`
    (robust_e_nerf/utils/autograd.py:4-34, robust_e_nerf/models/robust_e_nerf.py:394-398).

    inputs : owner, want_grad, o, d, o_dot, d_dot (B,3), jitter (B)|None, bkgd (C)|None, *params
    outputs: radiance (B,C), opacity (B), depth_raw (B), radiance_dot (B,C), opacity_dot (B), n_kept
    """

    @staticmethod
    def forward(ctx, owner, want_grad, o, d, o_dot, d_dot, jitter, bkgd, *params):
        lib = _lib.load()
        if not o.is_cuda:
            raise RuntimeError("ren_b200 renders on CUDA tensors only (no CPU fallback)")
        B = o.shape[0]
        C_ = owner.radiance_dim
        ray_grad = bool(want_grad) and any(ctx.needs_input_grad[2:6])
        fdesc, rdesc, K = owner._descs()
        flat = owner._flat_for(params)
        need_grad = bool(want_grad) and (ray_grad or any(ctx.needs_input_grad[7:]))
        dev = o.device
        o, d, o_dot, d_dot = (t.contiguous().float() for t in (o, d, o_dot, d_dot))
        if jitter is not None:
            jitter = jitter.contiguous().float()
        if bkgd is not None:
            bkgd = bkgd.contiguous().float()
        radiance = torch.empty((B, C_), device=dev, dtype=torch.float32)
        radiance_dot = torch.empty((B, C_), device=dev, dtype=torch.float32)
        opacity = torch.empty((B,), device=dev, dtype=torch.float32)
        opacity_dot = torch.empty((B,), device=dev, dtype=torch.float32)
        depth = torch.empty((B,), device=dev, dtype=torch.float32)
        n_kept = torch.empty((B,), device=dev, dtype=torch.int32)
        packed, keep = owner._march_pack(fdesc, rdesc, B, K, o, d, jitter)
        nS, nK = (max(int(packed.n_samples), 1), 1) if packed is not None else (B, K)
        stash, stash_n = None, 0
        if need_grad:
            stash_n = _lib.tangent_stash_bytes(fdesc, rdesc, nS, nK)
            stash = torch.empty((stash_n,), device=dev, dtype=torch.uint8)
        ws = torch.empty((_lib.tangent_workspace_bytes(fdesc, rdesc, nS, nK, need_grad, False),),
                         device=dev, dtype=torch.uint8)
        _lib.check(lib.ren_render_tangent_forward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o), _lib.ptr(d), _lib.ptr(o_dot),
            _lib.ptr(d_dot), _lib.ptr(jitter), _lib.ptr(flat), _lib.ptr(bkgd), _lib.ptr(radiance),
            _lib.ptr(opacity), _lib.ptr(depth), _lib.ptr(radiance_dot), _lib.ptr(opacity_dot),
            _lib.ptr(n_kept), None if packed is None else C.byref(packed), _lib.ptr(stash), stash_n, _lib.ptr(ws),
            ws.numel(), _lib.stream_ptr()), "ren_render_tangent_forward")
        if need_grad:
            ctx.owner = owner
            ctx.descs = (fdesc, rdesc, K)
            ctx.packed, ctx.packed_keep = packed, keep
            empty = torch.empty(0, device=dev)
            ctx.save_for_backward(o, d, o_dot, d_dot, jitter if jitter is not None else empty,
                                  bkgd if bkgd is not None else empty, stash, flat)
            ctx.has_jitter = jitter is not None
            ctx.has_bkgd = bkgd is not None
            ctx.ray_grad = ray_grad
        ctx.mark_non_differentiable(n_kept)
        return radiance, opacity, depth, radiance_dot, opacity_dot, n_kept

    @staticmethod
    def backward(ctx, g_rad, g_op, g_depth, g_rad_dot, g_op_dot, _g_kept):
        _lib.refuse_double_backward()
        lib = _lib.load()
        o, d, o_dot, d_dot, jitter, bkgd, stash, flat = ctx.saved_tensors
        fdesc, rdesc, K = ctx.descs
        owner = ctx.owner
        B = o.shape[0]
        dev = o.device
        jitter = jitter if ctx.has_jitter else None
        bkgd = bkgd if ctx.has_bkgd else None
        g_flat = torch.zeros_like(flat)
        g_bkgd = torch.zeros_like(bkgd) if bkgd is not None else None
        packed = ctx.packed
        nS, nK = (max(int(packed.n_samples), 1), 1) if packed is not None else (B, K)
        ws = torch.empty((_lib.tangent_workspace_bytes(fdesc, rdesc, nS, nK, True, True),), device=dev,
                         dtype=torch.uint8)
        cg = lambda t: None if t is None else t.contiguous().float()
        g_rad, g_op, g_depth, g_rad_dot, g_op_dot = (cg(t) for t in (g_rad, g_op, g_depth, g_rad_dot, g_op_dot))
        g_rays = torch.empty((B, 12), device=dev, dtype=torch.float32) if ctx.ray_grad else None
        _lib.check(lib.ren_render_tangent_backward(
            C.byref(fdesc), C.byref(rdesc), B, K, _lib.ptr(o), _lib.ptr(d), _lib.ptr(o_dot),
            _lib.ptr(d_dot), _lib.ptr(jitter), _lib.ptr(flat), _lib.ptr(bkgd), _lib.ptr(g_rad),
            _lib.ptr(g_op), _lib.ptr(g_depth), _lib.ptr(g_rad_dot), _lib.ptr(g_op_dot),
            None if packed is None else C.byref(packed), _lib.ptr(stash), stash.numel(), _lib.ptr(g_flat),
            _lib.ptr(g_bkgd), _lib.ptr(g_rays), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
            "ren_render_tangent_backward")
        grads = owner._split_flat(g_flat)
        gr = [g_rays[:, 3 * i:3 * i + 3] if ctx.ray_grad else None for i in range(4)]
        return (None, None, gr[0], gr[1], gr[2], gr[3], None, g_bkgd) + tuple(grads)


class _GridParams(torch.nn.Module):
    """Stands where `tcnn.Encoding` stands in the reference's `mlp_base` Sequential
    (robust_e_nerf/external/ngp.py:166-188): ONE flat fp32 parameter `params` (state-dict key
    `radiance_field.mlp_base.0.params`), initialised U(-1e-4, 1e-4) like tiny-cuda-nn's grids."""

    def __init__(self, n_params):
        super().__init__()
        self.params = torch.nn.Parameter((torch.rand(n_params) * 2 - 1) * 1e-4)


class NeRF(torch.nn.Module):
    """See module docstring.  `arch: mlp` is the 8x256 network the north star names; `arch: ngp`
    (robust_e_nerf/external/ngp.py:109-280) is what the shipped configs select."""

    HIDDEN_ACTIVATIONS = ("softplus", "relu")                               # models/nerf.py:17-20
    DENSITY_ACTIVATIONS = ("shifted_trunc_exp", "softplus", "shifted_softplus")   # :21-25
    RADIANCE_ACTIVATIONS = ("softplus", "sigmoid")                          # :26-29

    def __init__(self, aabb, contraction_type, occ_grid_config, near_plane, far_plane,
                 render_step_size, render_bkgd, cone_angle, early_stop_eps, alpha_thre,
                 test_chunk_size, arch, arch_config, num_dim=3, radiance_dim=1,
                 opacity_eps=1e-10, engine="auto"):
        super().__init__()
        # the reference's argument checks (models/nerf.py:52-69)
        if (near_plane is not None) and (far_plane is not None):
            assert 0 <= near_plane <= far_plane
        assert render_step_size > 0
        assert (render_bkgd is None) or (isinstance(render_bkgd, str) and render_bkgd == "parameter") \
            or isinstance(render_bkgd, torch.Tensor)
        assert cone_angle >= 0
        assert 0 <= early_stop_eps <= 1
        assert 0 <= alpha_thre <= 1
        assert test_chunk_size > 0
        assert num_dim == 3, "ren_b200 renders 3-D scenes"
        assert radiance_dim > 0
        assert isinstance(opacity_eps, (int, float)) and opacity_eps > 0
        if arch not in ("mlp", "ngp"):
            raise NotImplementedError(f"arch={arch!r}")                      # models/nerf.py:166-167
        self.arch = arch
        sub = [arch_config] if arch == "mlp" else [_cfg_get(arch_config, "mlp_base"), _cfg_get(arch_config, "mlp_head")]
        wn = [bool(_cfg_get(c, "weight_norm")) for c in sub]
        if any(wn) and not all(wn):
            raise NotImplementedError("ren_b200 applies weight_norm to mlp_base and mlp_head together or not at all")
        # external/mlp.py:301-319, external/ngp.py:209-217: `torch.nn.utils.weight_norm` on every Linear
        # (state-dict keys `weight_g` / `weight_v`); the kernels see the effective weights g v / |v|
        self.weight_norm = all(wn)

        self.register_buffer("aabb", torch.tensor(aabb, dtype=torch.float32), persistent=False)
        self.contraction_type = contraction_type
        self.contraction = _contraction_name(contraction_type)
        self.occ_grid_config = occ_grid_config
        # models/nerf.py:98-102.  `enabled: false` (an extra key, absent from the reference YAMLs) keeps every cell
        # occupied -- the fixed-N regime of SURVEY.md section 8d that the golden vectors and bench.py use.
        cfg_get = (lambda k, dflt: occ_grid_config.get(k, dflt)) if isinstance(occ_grid_config, dict) \
            else (lambda k, dflt: getattr(occ_grid_config, k, dflt))
        self.occ_grid_enabled = bool(cfg_get("enabled", True)) and occ_grid_config is not None
        self.occupancy_grid = None
        if self.occ_grid_enabled:
            self.occupancy_grid = OccupancyGrid(aabb, cfg_get("resolution", 128), self.contraction)
        self.near_plane = near_plane
        self.far_plane = far_plane
        self.register_buffer("render_step_size", torch.tensor(render_step_size, dtype=torch.float32),
                             persistent=False)
        self.cone_angle = cone_angle
        self.early_stop_eps = early_stop_eps
        self.alpha_thre = alpha_thre
        self.test_chunk_size = test_chunk_size
        self.opacity_eps = opacity_eps
        self.radiance_dim = radiance_dim
        self.engine = engine

        # background (models/nerf.py:79-93): softplus-parametrised parameter, buffer or None
        self._bkgd_mode = None
        if render_bkgd is None:
            self.render_bkgd = None
        elif isinstance(render_bkgd, str):
            self._bkgd_mode = "parameter"
            holder = _Holder()
            holder.render_bkgd = _Holder()
            # parametrize stores right_inverse(init): softplus(original) == ones, with the reference's own
            # formula log(exp(x) - 1) (utils/modules.py:46-54) so that the initial value is bit-identical
            holder.render_bkgd.original = torch.nn.Parameter(
                torch.log(torch.exp(torch.ones(radiance_dim)) - 1))
            self.parametrizations = holder          # key: parametrizations.render_bkgd.original
        else:
            self._bkgd_mode = "buffer"
            self.register_buffer("render_bkgd", render_bkgd, persistent=False)

        g = lambda k: _cfg_get(arch_config, k)
        if arch == "ngp":
            as_dict = lambda c: dict(c) if isinstance(c, dict) else {k: v for k, v in vars(c).items()}
            blocks = {k: as_dict(g(k)) for k in ("pos_encoding", "dir_encoding", "mlp_base", "mlp_head")}
            self.hidden_activation = blocks["mlp_base"]["hidden_activation"]
            self.density_activation = blocks["mlp_base"]["density_activation"]
            self.radiance_activation = blocks["mlp_head"]["radiance_activation"]
            assert self.hidden_activation in self.HIDDEN_ACTIVATIONS
            assert blocks["mlp_head"]["hidden_activation"] in self.HIDDEN_ACTIVATIONS
            assert self.density_activation in self.DENSITY_ACTIVATIONS
            assert self.radiance_activation in self.RADIANCE_ACTIVATIONS
            self._fdesc = _lib.make_ngp_field_desc(radiance_dim=radiance_dim, contraction=self.contraction, aabb=aabb,
                                                   **blocks)
            self._layout = _lib.param_layout(self._fdesc)
            self._n_params = _lib.param_count(self._fdesc)
            self._n_grid = _lib.grid_param_count(self._fdesc)
            self._build_ngp_tree(blocks)
            self._desc_cache = {}
            return
        self._n_grid = 0
        self.hidden_activation = g("hidden_activation")
        self.density_activation = g("density_activation")
        self.radiance_activation = g("radiance_activation")
        assert self.hidden_activation in self.HIDDEN_ACTIVATIONS
        assert self.density_activation in self.DENSITY_ACTIVATIONS
        assert self.radiance_activation in self.RADIANCE_ACTIVATIONS
        self._fdesc = _lib.make_field_desc(
            net_depth=g("net_depth"), net_width=g("net_width"), skip_layer=g("skip_layer"),
            net_depth_condition=g("net_depth_condition"), net_width_condition=g("net_width_condition"),
            radiance_dim=radiance_dim, pos_deg=g("pos_encoder_max_deg"),
            dir_deg=g("view_encoder_max_deg"), hidden_activation=self.hidden_activation,
            density_activation=self.density_activation, radiance_activation=self.radiance_activation,
            contraction=self.contraction, aabb=aabb)
        self._layout = _lib.param_layout(self._fdesc)
        self._n_params = _lib.param_count(self._fdesc)
        self._build_tree(g)
        self._desc_cache = {}

    # ------------------------------------------------------------------ parameters
    def _build_tree(self, g):
        """Module tree with the reference's key names (external/mlp.py:143-186,
        external/mlp.py:266-299); default `nn.Linear` init in construction order so that
        `torch.manual_seed(s); NeRF(...)` draws the same weights as the reference."""
        depth, cdepth = g("net_depth"), g("net_depth_condition")
        names = [f"base.hidden_layers.{i}" for i in range(depth)]
        names += ["sigma_layer.output_layer", "bottleneck_layer.output_layer"]
        names += [f"rgb_layer.hidden_layers.{j}" for j in range(cdepth)]
        names += ["rgb_layer.output_layer"]
        assert len(names) == len(self._layout)
        rf = _Holder()
        rf.register_buffer("aabb", torch.tensor(self.aabb.tolist(), dtype=torch.float32))  # mlp.py:269
        rf.posi_encoder = _Encoder(g("pos_encoder_max_deg"))
        rf.view_encoder = _Encoder(g("view_encoder_max_deg"))
        mlp = _Holder()
        mlp.base = _Holder()
        mlp.base.hidden_layers = torch.nn.ModuleList()
        mlp.sigma_layer = _Holder()
        mlp.bottleneck_layer = _Holder()
        mlp.rgb_layer = _Holder()
        mlp.rgb_layer.hidden_layers = torch.nn.ModuleList()
        self._linears = []
        for name, (w_off, b_off, out_f, in_f) in zip(names, self._layout):
            lin = torch.nn.Linear(in_f, out_f)
            self._linears.append(lin)
            if name.startswith("base.hidden_layers"):
                mlp.base.hidden_layers.append(lin)
            elif name.startswith("rgb_layer.hidden_layers"):
                mlp.rgb_layer.hidden_layers.append(lin)
            else:
                head, leaf = name.split(".")
                setattr(getattr(mlp, head), leaf, lin)
        rf.mlp = mlp
        self.radiance_field = rf
        self._apply_weight_norm()
        self._flatten()

    def _build_ngp_tree(self, blocks):
        """Module tree with the reference's key names for `arch: ngp` (external/ngp.py:152-207):
        `radiance_field.mlp_base.0.params` (grid), `mlp_base.1.{hidden_layers.i, output_layer}`,
        `mlp_head.{hidden_layers.j, output_layer}`; grid first, then default `nn.Linear` inits in construction
        order (the reference passes hidden_init / output_init / bias_init = None)."""
        rf = _Holder()
        rf.register_buffer("aabb", torch.tensor(self.aabb.tolist(), dtype=torch.float32))   # ngp.py:150
        grid = _GridParams(self._n_grid)
        base, head = _Holder(), _Holder()
        base.hidden_layers = torch.nn.ModuleList()
        head.hidden_layers = torch.nn.ModuleList()
        n_base = int(blocks["mlp_base"]["n_hidden_layers"])
        n_head = int(blocks["mlp_head"]["n_hidden_layers"])
        assert len(self._layout) == n_base + 1 + n_head + 1
        self._linears = []
        for i, (w_off, b_off, out_f, in_f) in enumerate(self._layout):
            lin = torch.nn.Linear(in_f, out_f)
            self._linears.append(lin)
            if i < n_base:
                base.hidden_layers.append(lin)
            elif i == n_base:
                base.output_layer = lin
            elif i < n_base + 1 + n_head:
                head.hidden_layers.append(lin)
            else:
                head.output_layer = lin
        rf.mlp_base = torch.nn.Sequential(grid, base)
        rf.mlp_head = head
        self.radiance_field = rf
        object.__setattr__(self, "_grid", grid)          # alias, not a second registration of the module
        self._apply_weight_norm()
        self._flatten()

    def _apply_weight_norm(self):
        if self.weight_norm:
            for lin in self._linears:
                torch.nn.utils.weight_norm(lin)

    def _flat_effective(self):
        """weight_norm: the flat buffer of EFFECTIVE parameters as a differentiable function of
        (weight_g, weight_v, bias) [and the grid]: autograd carries the render's flat gradient back to them."""
        parts = [self._grid.params] if self._n_grid else []
        for lin in self._linears:
            parts.append(torch._weight_norm(lin.weight_v, lin.weight_g, 0).reshape(-1))
            parts.append(lin.bias)
        return torch.cat(parts).float()

    def _render_params(self):
        return [self._flat_effective()] if self.weight_norm else self.mlp_parameters()

    def _flat_for(self, params):
        """The flat parameter buffer a render call reads: the caller's effective-weight tensor under weight_norm,
        else the buffer the named parameters are views of."""
        if self.weight_norm:
            return params[0].detach().contiguous()
        return self.flat_params()

    def _flatten(self):
        """(Re)pack the grid and every Linear's weight/bias into one flat buffer and re-point the
        parameters at views of it."""
        if self.weight_norm:
            self._flat = None                      # parameters are (g, v, bias) leaves: no shared buffer
            return
        ref = self._linears[0].weight
        flat = torch.empty((self._n_params,), device=ref.device, dtype=torch.float32)
        with torch.no_grad():
            if self._n_grid:
                flat[:self._n_grid].copy_(self._grid.params.detach())
                self._grid.params.data = flat[:self._n_grid]
            for lin, (w_off, b_off, out_f, in_f) in zip(self._linears, self._layout):
                flat[w_off:w_off + out_f * in_f].copy_(lin.weight.detach().reshape(-1))
                flat[b_off:b_off + out_f].copy_(lin.bias.detach())
                lin.weight.data = flat[w_off:w_off + out_f * in_f].view(out_f, in_f)
                lin.bias.data = flat[b_off:b_off + out_f]
        self._flat = flat

    def _views_intact(self):
        base = self._flat.data_ptr()
        if self._flat.device != self._linears[0].weight.device:
            return False
        if self._n_grid and self._grid.params.data_ptr() != base:
            return False
        for lin, (w_off, b_off, _, _) in zip(self._linears, self._layout):
            if lin.weight.data_ptr() != base + 4 * w_off or lin.bias.data_ptr() != base + 4 * b_off:
                return False
        return True

    def flat_params(self):
        """The flat fp32 parameter buffer (rebuilt if `.to()` / `load_state_dict` moved
        the parameters)."""
        if self.weight_norm:
            with torch.no_grad():
                return self._flat_effective().contiguous()
        if not self._views_intact():
            self._flatten()
        return self._flat

    def mlp_parameters(self):
        out = [self._grid.params] if self._n_grid else []
        for lin in self._linears:
            out += [lin.weight_g, lin.weight_v, lin.bias] if self.weight_norm else [lin.weight, lin.bias]
        return out

    def _split_flat(self, g_flat):
        if self.weight_norm:
            return [g_flat]                        # one input (the effective flat buffer), one gradient
        grads = [g_flat[:self._n_grid]] if self._n_grid else []
        for (w_off, b_off, out_f, in_f) in self._layout:
            grads.append(g_flat[w_off:w_off + out_f * in_f].view(out_f, in_f))
            grads.append(g_flat[b_off:b_off + out_f])
        return grads

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._flatten()
        return out

    def background(self):
        """Current background radiance (C,) or None (models/nerf.py:79-93)."""
        if self._bkgd_mode == "parameter":
            return torch.nn.functional.softplus(self.parametrizations.render_bkgd.original, beta=1)
        if self._bkgd_mode == "buffer":
            return self.render_bkgd
        return None

    # ------------------------------------------------------------------ rendering
    def _descs(self, engine=None):
        engine = self.engine if engine is None else engine
        key = (self.training, engine)
        hit = self._desc_cache.get(key)
        if hit is None:
            rdesc = _lib.make_render_desc(
                near_plane=self.near_plane, far_plane=self.far_plane,
                render_step_size=float(self.render_step_size), cone_angle=self.cone_angle,
                early_stop_eps=self.early_stop_eps, alpha_thre=self.alpha_thre,
                stratified=self.training,                    # external/utils.py:115
                engine=engine)
            K = _lib.samples_per_ray(self._fdesc, rdesc)
            hit = (self._fdesc, rdesc, K)
            self._desc_cache[key] = hit
        _lib.set_occupancy(hit[1], None if self.occupancy_grid is None else self.occupancy_grid.binary_u8())
        return hit

    packed_march = True      # occupancy grid + tcgen05 engine: run the field on the emitted samples only

    def _march_pack(self, fdesc, rdesc, n_rays, K, o, d, jitter):
        """(RenPacked, keep-alive tensors) when the render goes through the packed march, else (None, None)."""
        if self.occupancy_grid is None or not self.packed_march or n_rays == 0:
            return None, None
        if _lib.resolved_engine(fdesc, rdesc) == _lib.REN_ENGINE["simt_fp32"]:
            return None, None
        return _lib.march_pack(fdesc, rdesc, n_rays, K, o, d, jitter)

    def query_density(self, x):
        """`VanillaNeRFRadianceField.query_density` (external/mlp.py:344-348) at free points (n,3) -> (n,)."""
        lib = _lib.load()
        if not x.is_cuda:
            raise RuntimeError("ren_b200 evaluates the field on CUDA tensors only (no CPU fallback)")
        x = x.detach().contiguous().float()
        n = x.shape[0]
        fdesc, rdesc, _ = self._descs()
        sigma = torch.empty((n,), device=x.device, dtype=torch.float32)
        nb = C.c_size_t(0)
        _lib.check(lib.ren_density_workspace_bytes(C.byref(fdesc), C.byref(rdesc), n, C.byref(nb)),
                   "ren_density_workspace_bytes")
        ws = torch.empty((int(nb.value),), device=x.device, dtype=torch.uint8)
        _lib.check(lib.ren_field_density(C.byref(fdesc), C.byref(rdesc), n, _lib.ptr(x), _lib.ptr(self.flat_params()),
                                         _lib.ptr(sigma), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "ren_field_density")
        return sigma

    @torch.no_grad()
    def update_occ_grid(self, step, T_wc_position, generator=None):
        """models/nerf.py:170-204: the nerfacc occupancy-grid update, with the density sweep over the sampled cells
        on the field kernels.  `occ_grid.enabled: false` keeps every cell occupied.  (For `aabb` nerfacc skips empty
        voxels by whole marching steps, so the kept samples are the lattice steps in occupied cells -- the same
        per-step test the sphere / tanh grids use.)"""
        if self.occupancy_grid is None:
            return None
        cfg = self.occ_grid_config
        get = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
        step_size = float(self.render_step_size)

        def occ_eval_fn(x):
            if self.cone_angle > 0.0:
                cam = torch.randint(0, len(T_wc_position), (x.shape[0],), device=x.device, generator=generator)
                t = (T_wc_position.to(x.device)[cam, :] - x).norm(dim=-1)
                ss = torch.clamp(t * self.cone_angle, min=step_size)
                if (self.near_plane is not None) and (self.far_plane is not None):
                    ss = torch.where((t > self.near_plane) & (t < self.far_plane), ss, torch.zeros_like(ss))
            else:
                ss = step_size
            return self.query_density(x) * ss

        self.occupancy_grid.every_n_step(step, occ_eval_fn, get("occ_thre"), get("ema_decay"), get("warmup_steps"),
                                         get("n"), generator=generator)

    @staticmethod
    def pixel_params_to_ray(intrinsics_inverse, pixel_position, T_wc_position, T_wc_orientation):
        """models/nerf.py:206-228 (plain torch: a handful of (N,3) ops; the training path
        uses the fused `ren_event_rays` kernel instead)."""
        h = torch.cat((pixel_position, torch.ones_like(pixel_position[..., 0].unsqueeze(dim=-1))), dim=-1)
        ray_direction = (T_wc_orientation @ (intrinsics_inverse @ h.unsqueeze(dim=-1))).squeeze(dim=-1)
        ray_direction = ray_direction / torch.linalg.vector_norm(ray_direction, dim=-1, keepdim=True)
        return T_wc_position, ray_direction

    def render_raw(self, o, d, jitter=None):
        """(radiance (B,C), opacity (B), un-normalised depth (B), n_kept (B) i32)."""
        if self.training and jitter is None:
            jitter = torch.rand((o.shape[0],), device=o.device, dtype=torch.float32)
        if not self.training:
            jitter = None
        return RenderFunction.apply(self, torch.is_grad_enabled(), o, d, jitter, self.background(),
                                    *self._render_params())

    def render_tangent_raw(self, o, d, o_dot, d_dot, jitter=None):
        """Render + derivative along the pose interpolation weight: (radiance (B,C), opacity (B),
        un-normalised depth (B), radiance_dot (B,C), opacity_dot (B), n_kept (B) i32)."""
        if self.training and jitter is None:
            jitter = torch.rand((o.shape[0],), device=o.device, dtype=torch.float32)
        if not self.training:
            jitter = None
        return RenderTangentFunction.apply(self, torch.is_grad_enabled(), o, d, o_dot, d_dot, jitter,
                                           self.background(), *self._render_params())

    def forward(self, ray_origin, ray_direction, jitter=None):
        """models/nerf.py:230-286.  `jitter` (optional, (N,) in [0,1)) injects the per-ray
        stratified offsets nerfacc draws internally."""
        lead = ray_origin.shape[:-1]
        o = ray_origin.reshape(-1, 3)
        d = ray_direction.reshape(-1, 3)
        n_rays = o.shape[0]
        rad, opac, depth, n_kept = [], [], [], []
        chunk = n_rays if self.training else self.test_chunk_size         # external/utils.py:99-104
        for i in range(0, max(n_rays, 1), max(chunk, 1)):
            j = None if jitter is None else jitter.reshape(-1)[i:i + chunk]
            r, op, dp, nk = self.render_raw(o[i:i + chunk], d[i:i + chunk], j)
            rad.append(r); opac.append(op); depth.append(dp); n_kept.append(nk)
        radiance = torch.cat(rad).reshape(*lead, self.radiance_dim).squeeze(dim=-1)
        opacity = torch.cat(opac).reshape(lead)
        depth = torch.cat(depth).reshape(lead)
        depth = depth / (opacity + self.opacity_eps)
        mean_num_samples_per_ray = torch.cat(n_kept).sum() / n_rays        # device scalar: no host sync
        return radiance, opacity, depth, mean_num_samples_per_ray
