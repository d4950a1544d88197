"""Pins the oracle (oracle/restatement.py, torch CPU) to the golden vectors minted by the
UNMODIFIED reference (oracle/make_golden.py): loss, per-ray log-intensities, per-render
outputs and parameter gradients.  fp32 throughout; bars are fp32 round-off, not 1e-4."""
import pytest
import torch

import helpers as H
from oracle import restatement as R
from oracle import synth


def _run_restatement(g, dataset_dir):
    fcfg, rcfg, scfg = H.golden_configs(g)
    params = {k: v.clone().requires_grad_(True) for k, v in H.golden_params(g).items()}
    scene = synth.load_scene(dataset_dir)
    th_r = torch.log(torch.expm1(scene["pos_contrast_threshold"] / scene["neg_contrast_threshold"]))
    th_r = th_r.clone().requires_grad_(not g["freeze_ct"])
    tau_max = scene["max_refractory_period"]
    tau0 = scene["refractory_period"].to(torch.float64)
    th_t = (tau_max * torch.logit(tau0 / tau_max)).clone().requires_grad_(not g["freeze_rp"])
    th_b = torch.log(torch.expm1(torch.ones(fcfg.radiance_dim))).requires_grad_(True) if g["learn_bkgd"] else None
    out = R.training_step(params, fcfg, rcfg, scfg, scene, g["batch"], g["jitter"], th_r, th_t, th_b)
    return out, params, th_b


@pytest.mark.parametrize("name", H.GOLDEN_NAMES + H.NGP_GOLDEN_NAMES)
def test_restatement_matches_reference_golden(name, dataset_dir):
    g = H.load_golden(name)
    torch.set_num_threads(8)
    out, params, th_b = _run_restatement(g, dataset_dir)
    prefix = R.field_param_prefix(H.golden_configs(g)[0])
    assert abs(float(out["loss"].detach()) - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    spp = float(g["logged"]["train/mean_num_samples_per_ray"])
    assert abs(float(out["mean_num_samples_per_ray"]) - spp) < 1e-5 * spp
    out["loss"].backward()
    checked = 0
    for k, gref in g["grads"].items():
        if gref is None or not k.startswith(prefix):
            continue
        got = params[k[len("nerf.radiance_field."):]].grad
        nw, _, _ = H.rel_err(got, gref)
        assert nw < 5e-4, (k, nw)
        checked += 1
    assert checked >= (9 if name in H.NGP_GOLDEN_NAMES else 10)
    if th_b is not None:
        gref = g["grads"]["nerf.parametrizations.render_bkgd.original"]
        assert H.rel_err(th_b.grad, gref)[0] < 5e-4


@pytest.mark.parametrize("name", ["step_8x256", "step_aabb_cull"])
def test_render_matches_reference_golden(name):
    """`NeRF.forward` outputs of the reference at fixed (o, d, jitter)."""
    g = H.load_golden(name)
    fcfg, rcfg, _ = H.golden_configs(g)
    params = H.golden_params(g)
    bk = torch.ones(fcfg.radiance_dim) if g["learn_bkgd"] else None
    bk = None if bk is None else torch.nn.functional.softplus(torch.log(torch.expm1(bk)))
    for tag in ("start", "end"):
        rr = g["renders"][tag]
        with torch.no_grad():
            rad, opac, depth, spp = R.render(params, fcfg, rcfg, rr["o"], rr["d"], g["jitter"]["diff_" + tag], bk)
        assert H.rel_err(rad, rr["radiance"])[1] < 2e-6
        assert H.rel_err(opac, rr["opacity"])[1] < 2e-6
        assert H.rel_err(depth, rr["depth"])[0] < 1e-5
        assert abs(spp - float(rr["mean_spp"])) < 1e-6


def test_fp32_floor_is_recorded():
    """The golden keeps the fp64 run so parity reports can print the fp32 noise floor."""
    g = H.load_golden("step_8x256")
    l32, l64 = float(g["loss"]), float(g["fp64"]["loss"])
    assert abs(l32 - l64) / abs(l64) < 1e-3
