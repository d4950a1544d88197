"""ctypes binding of `libren_b200.so` (include/ren_b200.h).

The product path has NO fallback: if the shared library is missing or a call returns
a non-zero code, a `RuntimeError` is raised (the reference fails the same way when
its native dependency is absent: robust_e_nerf/external/ngp.py:34-42 exits when
`tinycudann` cannot be imported; nerfacc raises on non-CUDA tensors).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# REN_B200_LIB selects another build of the same library (A/B timing of kernel variants); default: the in-tree one
LIB_PATH = os.environ.get("REN_B200_LIB") or os.path.join(_HERE, "lib", "libren_b200.so")

REN_ENGINE = {"auto": 0, "simt_fp32": 1, "tc_f16x3": 2, "tc_f16": 3, "tc_f16a": 4}
HIDDEN_ACT = {"softplus": 0, "relu": 1}
DENSITY_ACT = {"shifted_trunc_exp": 0, "softplus": 1, "shifted_softplus": 2}
RADIANCE_ACT = {"softplus": 0, "sigmoid": 1}
CONTRACTION = {"aabb": 0, "sphere": 1, "tanh": 2}
ERROR_FN = {"l1": 0, "mse": 1, "mape": 2}
ARCH = {"mlp": 0, "ngp": 1}
GRID_TYPE = {"HashGrid": 0, "DenseGrid": 1, "TiledGrid": 2}
ABI_VERSION = 5


class RenFieldDesc(C.Structure):
    _fields_ = [
        ("net_depth", C.c_int32), ("net_width", C.c_int32), ("skip_layer", C.c_int32),
        ("net_depth_condition", C.c_int32), ("net_width_condition", C.c_int32),
        ("radiance_dim", C.c_int32), ("pos_deg", C.c_int32), ("dir_deg", C.c_int32),
        ("hidden_activation", C.c_int32), ("density_activation", C.c_int32),
        ("radiance_activation", C.c_int32), ("contraction", C.c_int32),
        ("aabb", C.c_float * 6),
        # arch ngp (include/ren_b200.h)
        ("arch", C.c_int32), ("grid_type", C.c_int32), ("grid_levels", C.c_int32), ("grid_features", C.c_int32),
        ("grid_log2_hashmap_size", C.c_int32), ("grid_base_resolution", C.c_int32),
        ("grid_per_level_scale", C.c_float), ("sh_degree", C.c_int32), ("geo_feat_dim", C.c_int32),
        ("reserved_ngp", C.c_int32),
    ]


class RenPacked(C.Structure):
    _fields_ = [("n_samples", C.c_int64), ("sample_idx", C.c_void_p), ("ray_offsets", C.c_void_p)]


class RenRenderDesc(C.Structure):
    _fields_ = [
        ("near_plane", C.c_float), ("far_plane", C.c_float), ("render_step_size", C.c_float),
        ("cone_angle", C.c_float), ("early_stop_eps", C.c_float), ("alpha_thre", C.c_float),
        ("has_near", C.c_int32), ("has_far", C.c_int32), ("stratified", C.c_int32),
        ("engine", C.c_int32),
        ("occ_binary", C.c_void_p), ("occ_res", C.c_int32 * 3), ("reserved", C.c_int32),
    ]


_P = C.c_void_p
_SIGNATURES = {
    "ren_version": (C.c_int, []),
    "ren_last_error": (C.c_char_p, []),
    "ren_launch_count": (C.c_int64, []),
    "ren_timing_enable": (C.c_int, [C.c_int32]),
    "ren_timing_collect": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "ren_param_count": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(C.c_int64)]),
    "ren_grid_param_count": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(C.c_int64)]),
    "ren_param_layout": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ren_samples_per_ray": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc),
                                      C.POINTER(C.c_int32)]),
    "ren_stash_bytes": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64,
                                  C.c_int32, C.POINTER(C.c_size_t)]),
    "ren_workspace_bytes": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64,
                                      C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ren_render_forward": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64,
                                     C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(RenPacked), _P,
                                     C.c_size_t, _P, C.c_size_t, _P]),
    "ren_resolve_engine": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.POINTER(C.c_int32)]),
    "ren_march_workspace_bytes": (C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    "ren_march_count": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64, C.c_int32,
                                  _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "ren_march_fill": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64, C.c_int32,
                                 _P, _P, _P, _P, _P, _P]),
    "ren_density_workspace_bytes": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64,
                                              C.POINTER(C.c_size_t)]),
    "ren_field_density": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64, _P, _P, _P, _P,
                                    C.c_size_t, _P]),
    "ren_render_backward": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64,
                                      C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(RenPacked), _P,
                                      C.c_size_t, _P, _P, _P, _P, C.c_size_t, _P]),
    "ren_tangent_stash_bytes": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64,
                                          C.c_int32, C.POINTER(C.c_size_t)]),
    "ren_tangent_workspace_bytes": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc),
                                              C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                              C.POINTER(C.c_size_t)]),
    "ren_render_tangent_forward": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc), C.c_int64,
                                             C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                             _P, C.POINTER(RenPacked), _P, C.c_size_t, _P, C.c_size_t, _P]),
    "ren_render_tangent_backward": (C.c_int, [C.POINTER(RenFieldDesc), C.POINTER(RenRenderDesc),
                                              C.c_int64, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                              _P, _P, _P, C.POINTER(RenPacked), _P, C.c_size_t, _P, _P, _P, _P,
                                              C.c_size_t, _P]),
    "ren_event_rays": (C.c_int, [C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P,
                                 C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ren_trajectory": (C.c_int, [C.c_int64, _P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "ren_event_diff_loss": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_float, _P, _P, _P, _P, _P,
                                      _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ren_event_queue_workspace_bytes": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ren_event_queue": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  C.c_size_t, _P]),
    "ren_adam_step": (C.c_int, [C.c_int64, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_int64, _P]),
    "ren_adam_step_dev": (C.c_int, [C.c_int64, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_float, _P, _P]),
}

_lib = None


def exported_symbols():
    """Names `include/ren_b200.h` declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def load():
    """dlopen the library (idempotent); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). ren_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ren_version() != ABI_VERSION:
        raise RuntimeError(f"libren_b200 ABI version {lib.ren_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ren_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor, None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "ren_b200 passes raw pointers: tensors must be contiguous"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_field_desc(net_depth=8, net_width=256, skip_layer=4, net_depth_condition=1,
                    net_width_condition=128, radiance_dim=1, pos_deg=10, dir_deg=4,
                    hidden_activation="softplus", density_activation="shifted_trunc_exp",
                    radiance_activation="softplus", contraction="sphere",
                    aabb=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)):
    d = RenFieldDesc()
    d.net_depth, d.net_width = int(net_depth), int(net_width)
    d.skip_layer = 0 if skip_layer is None else int(skip_layer)
    d.net_depth_condition, d.net_width_condition = int(net_depth_condition), int(net_width_condition)
    d.radiance_dim, d.pos_deg, d.dir_deg = int(radiance_dim), int(pos_deg), int(dir_deg)
    d.hidden_activation = HIDDEN_ACT[hidden_activation]
    d.density_activation = DENSITY_ACT[density_activation]
    d.radiance_activation = RADIANCE_ACT[radiance_activation]
    d.contraction = CONTRACTION[contraction]
    for i, v in enumerate(aabb):
        d.aabb[i] = float(v)
    return d


def make_ngp_field_desc(pos_encoding, dir_encoding, mlp_base, mlp_head, radiance_dim=1, contraction="sphere",
                        aabb=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)):
    """RenFieldDesc of `arch: ngp` from the four YAML blocks (configs/train/*.yaml:61-88; plain dicts)."""
    if pos_encoding.get("otype") not in GRID_TYPE:
        raise NotImplementedError(f"pos_encoding.otype {pos_encoding.get('otype')!r}")
    if pos_encoding.get("interpolation", "Linear") != "Linear":
        raise NotImplementedError("ren_b200 implements `interpolation: Linear` of the grid encoding")
    if mlp_base["hidden_activation"] != mlp_head["hidden_activation"]:
        raise NotImplementedError("ren_b200 expects mlp_base and mlp_head to share the hidden activation "
                                  "(every shipped config does)")
    d = RenFieldDesc()
    d.arch = ARCH["ngp"]
    d.net_depth, d.net_width = int(mlp_base["n_hidden_layers"]), int(mlp_base["n_neurons"])
    d.skip_layer = 0
    d.net_depth_condition, d.net_width_condition = int(mlp_head["n_hidden_layers"]), int(mlp_head["n_neurons"])
    d.radiance_dim = int(radiance_dim)
    d.hidden_activation = HIDDEN_ACT[mlp_base["hidden_activation"]]
    d.density_activation = DENSITY_ACT[mlp_base["density_activation"]]
    d.radiance_activation = RADIANCE_ACT[mlp_head["radiance_activation"]]
    d.contraction = CONTRACTION[contraction]
    for i, v in enumerate(aabb):
        d.aabb[i] = float(v)
    d.grid_type = GRID_TYPE[pos_encoding["otype"]]
    d.grid_levels = int(pos_encoding["n_levels"])
    d.grid_features = int(pos_encoding["n_features_per_level"])
    d.grid_log2_hashmap_size = int(pos_encoding.get("log2_hashmap_size", 19))
    d.grid_base_resolution = int(pos_encoding["base_resolution"])
    d.grid_per_level_scale = float(pos_encoding["per_level_scale"])
    d.sh_degree = int(dir_encoding["degree"])
    d.geo_feat_dim = int(mlp_base["geo_feat_dim"])
    return d


def grid_param_count(fdesc):
    n = C.c_int64(0)
    check(load().ren_grid_param_count(C.byref(fdesc), C.byref(n)), "ren_grid_param_count")
    return int(n.value)


def make_render_desc(near_plane=None, far_plane=None, render_step_size=1e-2, cone_angle=0.0,
                     early_stop_eps=0.0, alpha_thre=0.0, stratified=True, engine="auto"):
    import numpy as np
    d = RenRenderDesc()
    d.has_near = int(near_plane is not None)
    d.has_far = int(far_plane is not None)
    d.near_plane = float(near_plane) if near_plane is not None else 0.0
    d.far_plane = float(far_plane) if far_plane is not None else 0.0
    d.render_step_size = float(np.float32(render_step_size))
    d.cone_angle, d.early_stop_eps, d.alpha_thre = float(cone_angle), float(early_stop_eps), float(alpha_thre)
    d.stratified = int(bool(stratified))
    d.engine = REN_ENGINE[engine] if isinstance(engine, str) else int(engine)
    d.occ_binary = None
    return d


def set_occupancy(rdesc, binary):
    """Point a render descriptor at a (rx, ry, rz) uint8 device tensor (None = all occupied).  The caller keeps
    the tensor alive for as long as the descriptor is used."""
    if binary is None:
        rdesc.occ_binary = None
        return
    assert binary.dtype == __import__("torch").uint8 and binary.is_contiguous() and binary.dim() == 3
    rdesc.occ_binary = binary.data_ptr()
    for i in range(3):
        rdesc.occ_res[i] = int(binary.shape[i])


def param_layout(fdesc):
    """[(weight_offset, bias_offset, out_features, in_features)] in state-dict order."""
    lib = load()
    n = C.c_int32(0)
    check(lib.ren_param_layout(C.byref(fdesc), C.byref(n), None, None, None, None), "ren_param_layout")
    w = (C.c_int64 * n.value)()
    b = (C.c_int64 * n.value)()
    o = (C.c_int32 * n.value)()
    i = (C.c_int32 * n.value)()
    check(lib.ren_param_layout(C.byref(fdesc), C.byref(n), w, b, o, i), "ren_param_layout")
    return [(int(w[k]), int(b[k]), int(o[k]), int(i[k])) for k in range(n.value)]


def param_count(fdesc):
    n = C.c_int64(0)
    check(load().ren_param_count(C.byref(fdesc), C.byref(n)), "ren_param_count")
    return int(n.value)


def samples_per_ray(fdesc, rdesc):
    k = C.c_int32(0)
    check(load().ren_samples_per_ray(C.byref(fdesc), C.byref(rdesc), C.byref(k)), "ren_samples_per_ray")
    return int(k.value)


def stash_bytes(fdesc, rdesc, n_rays, K):
    n = C.c_size_t(0)
    check(load().ren_stash_bytes(C.byref(fdesc), C.byref(rdesc), n_rays, K, C.byref(n)), "ren_stash_bytes")
    return int(n.value)


def workspace_bytes(fdesc, rdesc, n_rays, K, has_stash, for_backward):
    n = C.c_size_t(0)
    check(load().ren_workspace_bytes(C.byref(fdesc), C.byref(rdesc), n_rays, K, int(has_stash),
                                     int(for_backward), C.byref(n)), "ren_workspace_bytes")
    return int(n.value)


def tangent_stash_bytes(fdesc, rdesc, n_rays, K):
    n = C.c_size_t(0)
    check(load().ren_tangent_stash_bytes(C.byref(fdesc), C.byref(rdesc), n_rays, K, C.byref(n)),
          "ren_tangent_stash_bytes")
    return int(n.value)


def tangent_workspace_bytes(fdesc, rdesc, n_rays, K, has_stash, for_backward):
    n = C.c_size_t(0)
    check(load().ren_tangent_workspace_bytes(C.byref(fdesc), C.byref(rdesc), n_rays, K, int(has_stash),
                                             int(for_backward), C.byref(n)), "ren_tangent_workspace_bytes")
    return int(n.value)


def resolved_engine(fdesc, rdesc):
    e = C.c_int32(0)
    check(load().ren_resolve_engine(C.byref(fdesc), C.byref(rdesc), C.byref(e)), "ren_resolve_engine")
    return int(e.value)


def march_pack(fdesc, rdesc, n_rays, K, o, d, jitter):
    """Packed march of one render call: (RenPacked, keep-alive tensors).  One host sync (the sample count), like
    nerfacc's packed ray_marching."""
    import torch
    lib = load()
    nb = C.c_size_t(0)
    check(lib.ren_march_workspace_bytes(n_rays, C.byref(nb)), "ren_march_workspace_bytes")
    ws = torch.empty((int(nb.value),), device=o.device, dtype=torch.uint8)
    offsets = torch.empty((n_rays + 1,), device=o.device, dtype=torch.int32)
    check(lib.ren_march_count(C.byref(fdesc), C.byref(rdesc), n_rays, K, ptr(o), ptr(d), ptr(jitter), ptr(offsets),
                              ptr(ws), ws.numel(), stream_ptr()), "ren_march_count")
    n = int(offsets[-1].item())
    idx = torch.empty((max(n, 1),), device=o.device, dtype=torch.int32)
    if n > 0:
        check(lib.ren_march_fill(C.byref(fdesc), C.byref(rdesc), n_rays, K, ptr(o), ptr(d), ptr(jitter), ptr(offsets),
                                 ptr(idx), stream_ptr()), "ren_march_fill")
    pk = RenPacked()
    pk.n_samples, pk.sample_idx, pk.ray_offsets = n, idx.data_ptr(), offsets.data_ptr()
    return pk, (idx, offsets)


def refuse_double_backward():
    """First statement of every autograd.Function.backward here: grad mode is on inside a backward only under
    `create_graph=True`, i.e. when the caller wants to differentiate the fused kernels a second time (the
    reference does so for d log I / d t, utils/autograd.py:4-34).  They are once differentiable -- say so instead
    of handing back gradients that silently carry no graph; `NeRF.render_tangent_raw` is the supported route."""
    import torch
    if torch.is_grad_enabled():
        raise RuntimeError("ren_b200: the fused render is once differentiable (create_graph=True is not supported); "
                           "use NeRF.render_tangent_raw / RobustENeRF.training_step for the log-intensity-gradient "
                           "loss")

