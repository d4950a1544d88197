"""The drop-in boundary, pinned against the UNMODIFIED reference module (build container only: needs
/root/reference): the INTEGRATION.md section 1 recipe -- a `RobustENeRF` subclass whose `_build_nerf`
(robust_e_nerf/models/robust_e_nerf.py:204-260) returns `ren_b200.NeRF` -- is executed for real, on CPU, up to
the first CUDA launch: construction from the reference's own hparams, the state dict of a stock reference model
loads into it key for key, the reference's `configure_optimizers` partitions its parameters like the stock
model's, `update_occ_grid` / `forward` keep their signatures, and the first render call refuses to run on CPU
(no fallback).  Both `arch: mlp` and `arch: ngp`."""
import inspect

import pytest
import torch

import helpers as H
from oracle import ref_harness, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference")


def _plugged_class(ref):
    import ren_b200

    class RobustENeRFB200(ref.models.robust_e_nerf.RobustENeRF):
        def _build_nerf(self, camera_poses, gpus):
            h = self.hparams.nerf
            return ren_b200.NeRF(h.aabb, h.contraction_type, h.occ_grid, h.near_plane, h.far_plane,
                                 h.render_step_size, self.render_bkgd, h.cone_angle, h.early_stop_eps,
                                 h.alpha_thre, h.test_chunk_size, h.arch, arch_config=h[h.arch],
                                 num_dim=self.NUM_DIM, radiance_dim=3 if self.has_bayer_filter else 1)

    return RobustENeRFB200


@pytest.mark.parametrize("name", ["step_plumbing", "step_ngp"])
def test_subclass_recipe_runs_against_the_unmodified_reference(name, built_lib, dataset_dir):
    import ren_b200
    g = H.load_golden(name)
    fcfg, rcfg, scfg = H.golden_configs(g)
    ref = ref_loader.load_reference()
    kw = dict(learn_bkgd=g["learn_bkgd"], freeze_ct=g["freeze_ct"], freeze_rp=g["freeze_rp"], seed=0)
    stock = ref_harness.build_reference_model(dataset_dir, fcfg, rcfg, scfg, **kw)
    stock_cls = ref.models.robust_e_nerf.RobustENeRF
    ref.models.robust_e_nerf.RobustENeRF = _plugged_class(ref)       # the harness instantiates this name
    try:
        plugged = ref_harness.build_reference_model(dataset_dir, fcfg, rcfg, scfg, **kw)
    finally:
        ref.models.robust_e_nerf.RobustENeRF = stock_cls
    assert isinstance(plugged.nerf, ren_b200.NeRF) and not isinstance(stock.nerf, ren_b200.NeRF)
    assert isinstance(plugged, stock_cls)

    # state dict: same keys, shapes and dtypes (the occupancy grid aside: the oracle's nerfacc shim keeps an
    # all-occupied grid with fewer buffers than nerfacc's); same seeded initial weights
    occ = "nerf.occupancy_grid."
    sd_s = {k: v for k, v in stock.state_dict().items() if not k.startswith(occ)}
    sd_p = {k: v for k, v in plugged.state_dict().items() if not k.startswith(occ)}
    assert set(sd_s) == set(sd_p)
    for k, v in sd_s.items():
        assert sd_p[k].shape == v.shape and sd_p[k].dtype == v.dtype, k
        assert torch.equal(sd_p[k], v), k                              # same construction order, same RNG draws
    # a reference-produced checkpoint loads
    with torch.no_grad():
        altered = {k: (v + 0.125 if v.is_floating_point() else v) for k, v in sd_s.items()}
    missing, unexpected = plugged.load_state_dict(altered, strict=False)
    assert not unexpected and all(k.startswith(occ) for k in missing)
    flat = plugged.nerf.flat_params()
    first = plugged.nerf._layout[0]
    key = [k for k in sd_s if k.endswith("hidden_layers.0.weight")][0]
    assert torch.equal(flat[first[0]:first[0] + first[2] * first[3]].view(first[2], first[3]), altered[key])

    # the reference's own configure_optimizers on the plugged module: same partition as the stock module
    groups_s = stock.configure_optimizers()["optimizer"].param_groups
    groups_p = plugged.configure_optimizers()["optimizer"].param_groups
    assert [len(gr["params"]) for gr in groups_s] == [len(gr["params"]) for gr in groups_p]
    assert [sum(p.numel() for p in gr["params"]) for gr in groups_s] == \
           [sum(p.numel() for p in gr["params"]) for gr in groups_p]
    assert groups_p[2]["weight_decay"] == groups_s[2]["weight_decay"] == 1e-6

    # method surface the reference's training_step touches (robust_e_nerf.py:375-389,856-860)
    ref_nerf = ref.models.nerf.NeRF
    want = list(inspect.signature(ref_nerf.update_occ_grid).parameters)
    have = list(inspect.signature(ren_b200.NeRF.update_occ_grid).parameters)
    assert have[:len(want)] == want
    want = list(inspect.signature(ref_nerf.forward).parameters)
    have = list(inspect.signature(ren_b200.NeRF.forward).parameters)
    assert have[:len(want)] == want
    assert list(inspect.signature(ref_nerf.pixel_params_to_ray).parameters) == \
           list(inspect.signature(ren_b200.NeRF.pixel_params_to_ray).parameters)

    # the reference's training_step runs unmodified up to the first render and stops there on a CPU-only box
    batch = {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in g["batch"].items()}
    with pytest.raises(RuntimeError, match="CUDA"):
        plugged.training_step(batch, 0)
