"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/ren_b200.h declares, and its host-only entry points (layout / sizing / argument
validation -- no compute, no GPU) agree with the oracle's description of the network."""
import ctypes as C
import os
import re

import pytest

from oracle import restatement as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ren_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ren_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    lib = C.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ren_b200.h but not exported"


def test_binding_covers_header(built_lib):
    from ren_b200 import _lib
    assert _lib.exported_symbols() == _declared_symbols()
    assert _lib.load().ren_version() == 5


@pytest.mark.parametrize("kw", [
    dict(), dict(net_depth=4, net_width=64, skip_layer=4, net_width_condition=128),
    dict(net_depth=4, net_width=64, skip_layer=2, net_width_condition=64),
    dict(net_depth=5, net_width=32, skip_layer=4, net_width_condition=16, radiance_dim=3),
    dict(net_depth=2, net_width=16, skip_layer=None, net_depth_condition=2, net_width_condition=32),
])
def test_param_layout_matches_reference_order(built_lib, kw):
    """Flat-buffer layout == the reference's state-dict order and nn.Linear shapes
    (external/mlp.py:56-76,143-186) as restated by the oracle."""
    from ren_b200 import _lib
    fcfg = R.FieldConfig(**kw)
    desc = _lib.make_field_desc(
        net_depth=fcfg.net_depth, net_width=fcfg.net_width, skip_layer=fcfg.skip_layer,
        net_depth_condition=fcfg.net_depth_condition, net_width_condition=fcfg.net_width_condition,
        radiance_dim=fcfg.radiance_dim)
    layout = _lib.param_layout(desc)
    shapes = fcfg.layer_shapes()
    assert len(layout) == len(shapes)
    off = 0
    for (w_off, b_off, out_f, in_f), (_, o, i) in zip(layout, shapes):
        assert (out_f, in_f) == (o, i)
        assert w_off == off
        off += o * i
        assert b_off == off
        off += o
    assert _lib.param_count(desc) == off
    if not kw:
        assert off == 595586


def test_samples_per_ray_bounds(built_lib):
    import torch
    from ren_b200 import _lib
    f = _lib.make_field_desc()
    for n in (32, 64, 128, 256):
        step = 4.0 / n
        r = _lib.make_render_desc(near_plane=0.5, far_plane=0.5 + (n + 0.5) * step, render_step_size=step)
        assert _lib.samples_per_ray(f, r) == n
    # the bound dominates the oracle's marcher for any jitter
    fc = R.FieldConfig()
    for far, step in ((4.5, 4 / 128), (3.7, 0.013), (6.0, 4 / 192)):
        rc = R.RenderConfig(near_plane=0.5, far_plane=far, render_step_size=step)
        o = torch.zeros(64, 3)
        d = torch.nn.functional.normalize(torch.randn(64, 3), dim=-1)
        t0, _, alive = R.march(fc, rc, o, d, torch.rand(64))
        r = _lib.make_render_desc(near_plane=0.5, far_plane=far, render_step_size=step)
        assert _lib.samples_per_ray(f, r) >= int(alive.sum(1).max())
        t0, _, alive = R.march(fc, rc, o, d, torch.zeros(64))
        assert _lib.samples_per_ray(f, r) == int(alive.sum(1).max())


def test_invalid_arguments_fail_loudly(built_lib):
    from ren_b200 import _lib
    lib = _lib.load()
    bad = _lib.make_field_desc(net_width=100)
    n = C.c_int64(0)
    assert lib.ren_param_count(C.byref(bad), C.byref(n)) == 1
    assert b"net_width" in lib.ren_last_error()
    f = _lib.make_field_desc()
    r = _lib.make_render_desc(near_plane=0.5, far_plane=None, render_step_size=0.01)   # unbounded sphere march
    k = C.c_int32(0)
    assert lib.ren_samples_per_ray(C.byref(f), C.byref(r), C.byref(k)) == 4
    with pytest.raises(RuntimeError):
        _lib.samples_per_ray(f, r)
    # null pointers are rejected before any launch
    r = _lib.make_render_desc(near_plane=0.5, far_plane=4.5, render_step_size=4 / 128)
    rc = lib.ren_render_forward(C.byref(f), C.byref(r), 8, 128, None, None, None, None, None, None, None,
                                None, None, None, None, 0, None, 0, None)
    assert rc == 1


def test_sizes(built_lib):
    from ren_b200 import _lib
    f = _lib.make_field_desc()
    S = 1024 * 128
    acts = 8 * 256 + 256 + 128
    # fp32 engine: blocked fp32 stash (activations + encodings + raw sigma/radiance)
    r = _lib.make_render_desc(near_plane=0.5, far_plane=4.5, render_step_size=4 / 128, engine="simt_fp32")
    assert _lib.stash_bytes(f, r, 1024, 128) == 4 * S * (acts + 64 + 32 + 2)
    assert _lib.workspace_bytes(f, r, 1024, 128, True, True) >= 4 * S * (acts + 2)
    assert _lib.workspace_bytes(f, r, 1024, 128, False, False) >= 4 * S * 2
    # tcgen05 engine (auto for the 8x256 network): fp16 stash, fp32 raw
    r = _lib.make_render_desc(near_plane=0.5, far_plane=4.5, render_step_size=4 / 128, engine="auto")
    # ... plus one live-tile byte per 128 samples (occupancy-grid marching), padded to 256
    assert _lib.stash_bytes(f, r, 1024, 128) == S * (2 * (acts + 64 + 32) + 4 * 2) + S // 128
    assert _lib.workspace_bytes(f, r, 1024, 128, True, True) >= S * (2 * acts + 4 * 2)
    # a shape the tcgen05 engine does not cover falls back to fp32 under auto and is refused when forced
    f64 = _lib.make_field_desc(net_depth=4, net_width=64)
    assert _lib.stash_bytes(f64, r, 8, 128) == 4 * 1024 * (4 * 64 + 64 + 128 + 64 + 32 + 2)
    forced = _lib.make_render_desc(near_plane=0.5, far_plane=4.5, render_step_size=4 / 128, engine="tc_f16x3")
    with pytest.raises(RuntimeError, match="tcgen05"):
        _lib.stash_bytes(f64, forced, 8, 128)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C11 on its own (no C++ or torch types)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use.c"
    src.write_text('#include "ren_b200.h"\n'
                   "int main(void) { RenFieldDesc f; RenRenderDesc r; RenPacked p; (void)f; (void)r; (void)p;\n"
                   "  return REN_ABI_VERSION == 5 ? 0 : 1; }\n")
    out = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                          str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


@pytest.mark.parametrize("kw", [
    dict(), dict(n_levels=8, log2_hashmap_size=12, base_resolution=4, per_level_scale=1.6),
    dict(n_levels=16, log2_hashmap_size=22, base_resolution=32, per_level_scale=1.3819),
    dict(grid_type="Dense", n_levels=4, base_resolution=8, per_level_scale=2.0, n_features_per_level=4),
    dict(grid_type="Tiled", n_levels=6, base_resolution=16, per_level_scale=1.5, n_features_per_level=1),
])
def test_ngp_grid_and_layout_match_the_oracle(built_lib, kw):
    """`arch: ngp` host arithmetic: the per-level table sizes (tiny-cuda-nn's offset table) and the flat parameter
    layout (grid, mlp_base, mlp_head in state-dict order) agree with the oracle's restatement."""
    from ren_b200 import _lib
    fcfg = R.FieldConfig(arch="ngp", **kw)
    meta = fcfg.grid_meta()
    otype = {"Hash": "HashGrid", "Dense": "DenseGrid", "Tiled": "TiledGrid"}[fcfg.grid_type]
    desc = _lib.make_ngp_field_desc(
        pos_encoding=dict(otype=otype, n_levels=fcfg.n_levels, n_features_per_level=fcfg.n_features_per_level,
                          log2_hashmap_size=fcfg.log2_hashmap_size, base_resolution=fcfg.base_resolution,
                          per_level_scale=fcfg.per_level_scale, interpolation="Linear"),
        dir_encoding=dict(degree=fcfg.sh_degree),
        mlp_base=dict(hidden_activation="softplus", density_activation="shifted_trunc_exp", n_neurons=64,
                      n_hidden_layers=1, geo_feat_dim=15),
        mlp_head=dict(hidden_activation="softplus", radiance_activation="softplus", n_neurons=64, n_hidden_layers=2))
    assert _lib.grid_param_count(desc) == meta["n_params"]
    layout = _lib.param_layout(desc)
    shapes = fcfg.layer_shapes()
    assert len(layout) == len(shapes)
    off = meta["n_params"]
    for (w_off, b_off, out_f, in_f), (_, o, i) in zip(layout, shapes):
        assert (out_f, in_f) == (o, i) and w_off == off
        off += o * i
        assert b_off == off
        off += o
    assert _lib.param_count(desc) == off
