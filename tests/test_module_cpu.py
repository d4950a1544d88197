"""Host-side logic of the drop-in modules (no GPU): constructor surface, state-dict key
names, flat parameter buffer, optimiser groups, loud failure on CPU."""
import pytest
import torch

import helpers as H
from oracle import restatement as R


def _model(dataset_dir, name="step_8x256"):
    g = H.load_golden(name)
    return g, H.build_product(g, dataset_dir)


def test_state_dict_keys_match_reference(built_lib, dataset_dir):
    _, m = _model(dataset_dir, "step_plumbing")
    keys = set(m.state_dict().keys())
    expect = {"contrast_threshold.parametrizations.p2n_contrast_threshold_ratio.original",
              "refractory_period.parametrizations._refractory_period.original",
              "nerf.parametrizations.render_bkgd.original", "nerf.radiance_field.aabb",
              "nerf.radiance_field.posi_encoder.scales", "nerf.radiance_field.view_encoder.scales",
              "nerf.radiance_field.mlp.sigma_layer.output_layer.weight",
              "nerf.radiance_field.mlp.sigma_layer.output_layer.bias",
              "nerf.radiance_field.mlp.bottleneck_layer.output_layer.weight",
              "nerf.radiance_field.mlp.bottleneck_layer.output_layer.bias",
              "nerf.radiance_field.mlp.rgb_layer.hidden_layers.0.weight",
              "nerf.radiance_field.mlp.rgb_layer.hidden_layers.0.bias",
              "nerf.radiance_field.mlp.rgb_layer.output_layer.weight",
              "nerf.radiance_field.mlp.rgb_layer.output_layer.bias"}
    expect |= {f"nerf.radiance_field.mlp.base.hidden_layers.{i}.{w}" for i in range(4) for w in ("weight", "bias")}
    assert keys == expect
    sd = m.state_dict()
    assert sd["refractory_period.parametrizations._refractory_period.original"].dtype == torch.float64
    assert sd["nerf.radiance_field.posi_encoder.scales"].dtype == torch.int64
    assert sd["nerf.radiance_field.posi_encoder.scales"].tolist() == [2 ** i for i in range(10)]
    # golden grads were recorded under the reference's parameter names: same set
    g = H.load_golden("step_plumbing")
    assert set(g["grads"].keys()) == {n for n, _ in m.named_parameters()}


def test_default_init_equals_reference_init(built_lib, dataset_dir):
    """torch.manual_seed(s); RobustENeRF(...) draws the reference's weights (external/mlp.py:297-299)."""
    from ren_b200 import RobustENeRF
    g = H.load_golden("step_8x256")
    torch.manual_seed(0)
    m = RobustENeRF(dataset_directory=dataset_dir, **H.product_config(g))
    ref = H.golden_params(g)
    sd = m.state_dict()
    for k, v in ref.items():
        assert torch.equal(sd["nerf.radiance_field." + k], v), k


def test_flat_parameter_views(built_lib, dataset_dir):
    _, m = _model(dataset_dir)
    nerf = m.nerf
    flat = nerf.flat_params()
    assert flat.numel() == 595586
    w = nerf.radiance_field.mlp.base.hidden_layers[3].weight
    with torch.no_grad():
        w[5, 7] = 123.0
    off = nerf._layout[3][0]
    assert float(nerf.flat_params()[off + 5 * 256 + 7]) == 123.0
    # .to() / load_state_dict keep the views consistent
    m2 = m.double().float()
    assert m2.nerf._views_intact()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["nerf.radiance_field.mlp.sigma_layer.output_layer.bias"].fill_(0.5)
    m.load_state_dict(sd)
    assert float(m.nerf.flat_params()[m.nerf._layout[8][1]]) == 0.5


def test_sensor_parametrisations(built_lib, dataset_dir):
    _, m = _model(dataset_dir)
    assert abs(float(m.contrast_threshold.p2n_contrast_threshold_ratio) - 1.0) < 1e-6
    assert abs(float(m.contrast_threshold.mean_contrast_threshold) - 0.25) < 1e-7
    assert m.refractory_period.refractory_period.dtype == torch.float64
    assert abs(float(m.refractory_period.refractory_period) - 1000.0) < 1e-6
    # logit clamp (models/event_generation_params.py:170-185)
    with torch.no_grad():
        m.refractory_period.parametrizations._refractory_period.original.fill_(1e9)
    tau = float(m.refractory_period.refractory_period)
    assert abs(tau - 50000 * (1 - 1e-4)) < 1e-3


def test_optimizer_groups(built_lib, dataset_dir):
    g, m = _model(dataset_dir, "step_plumbing")
    cfg = m.configure_optimizers()
    opt = cfg["optimizer"]
    assert isinstance(opt, torch.optim.Adam) and len(opt.param_groups) == 4
    ct, rp, mlp, other = opt.param_groups
    assert ct["lr"] == 0.1 and rp["lr"] == 50000 * 50 and mlp["lr"] == 0.01
    assert mlp["weight_decay"] == 1e-6 and ct["weight_decay"] == 0
    assert len(mlp["params"]) == 2 * (4 + 4) and len(other["params"]) == 1      # bkgd
    assert cfg["lr_scheduler"]["interval"] == "epoch"
    assert cfg["lr_scheduler"]["scheduler"].milestones == {20: 1, 30: 1, 36: 1}


def test_no_cpu_fallback(built_lib, dataset_dir):
    """The product path must fail loudly without CUDA (no oracle / torch fallback)."""
    g, m = _model(dataset_dir)
    o = torch.zeros(4, 3)
    d = torch.tensor([[0.0, 0.0, 1.0]]).repeat(4, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.nerf(o, d, torch.zeros(4))
    import ren_b200.nerf as nmod, ren_b200.module as mmod, inspect
    for mod in (nmod, mmod):
        assert "oracle" not in inspect.getsource(mod)


def test_unsupported_configurations_raise(built_lib, dataset_dir):
    from ren_b200 import NeRF
    g = H.load_golden("step_8x256")
    cfg = H.product_config(g)["nerf"]
    with pytest.raises(NotImplementedError):                      # models/nerf.py:166-167
        NeRF(cfg["aabb"], "sphere", cfg["occ_grid"], 0.5, 4.5, 0.03, None, 0.0, 0.0, 0.0, 16384, "tensorf",
             arch_config=cfg["mlp"])
    ngp = H.product_config(H.load_golden("step_ngp"))["nerf"]["ngp"]
    with pytest.raises(NotImplementedError):                      # grid interpolation other than Linear
        NeRF(cfg["aabb"], "sphere", cfg["occ_grid"], 0.5, 4.5, 0.03, None, 0.0, 0.0, 0.0, 16384, "ngp",
             arch_config=dict(ngp, pos_encoding=dict(ngp["pos_encoding"], interpolation="Smoothstep")))


@pytest.mark.parametrize("name", ["step_8x256", "step_ngp"])
def test_weight_norm_keys_and_effective_weights(name, built_lib, dataset_dir):
    """`weight_norm: true` (external/mlp.py:301-319, external/ngp.py:209-217): `weight_g` / `weight_v` state-dict
    keys like `torch.nn.utils.weight_norm`, the effective weights equal the plain initialisation, the optimiser's
    MLP group holds the (g, v, bias) leaves."""
    from ren_b200 import RobustENeRF
    g = H.load_golden(name)
    cfg = H.product_config(g)
    cfg_wn = H.product_config(g)
    if name == "step_ngp":
        cfg_wn["nerf"]["ngp"]["mlp_base"]["weight_norm"] = True
        cfg_wn["nerf"]["ngp"]["mlp_head"]["weight_norm"] = True
    else:
        cfg_wn["nerf"]["mlp"]["weight_norm"] = True
    torch.manual_seed(0)
    plain = RobustENeRF(dataset_directory=dataset_dir, **cfg)
    torch.manual_seed(0)
    wn = RobustENeRF(dataset_directory=dataset_dir, **cfg_wn)
    n_lin = len(plain.nerf._linears)
    keys = set(wn.state_dict())
    assert sum(k.endswith(".weight_g") for k in keys) == n_lin and sum(k.endswith(".weight_v") for k in keys) == n_lin
    assert not any(k.endswith(".weight") for k in keys if "radiance_field" in k)
    assert float((wn.nerf.flat_params() - plain.nerf.flat_params()).abs().max()) < 1e-6
    groups = wn.configure_optimizers()["optimizer"].param_groups
    assert len(groups[2]["params"]) == 3 * n_lin + (1 if name == "step_ngp" else 0)
    flat = wn.nerf._flat_effective()
    assert flat.requires_grad and flat.numel() == plain.nerf.flat_params().numel()


def test_ngp_module_tree_matches_the_reference_keys(built_lib, dataset_dir):
    """`arch: ngp`: state-dict keys / shapes of the reference's `NGPradianceField` (external/ngp.py:152-207) and
    the oracle's seeded initial weights (same construction order: grid, mlp_base, mlp_head)."""
    g = H.load_golden("step_ngp")
    from ren_b200 import RobustENeRF
    from oracle import restatement as R
    torch.manual_seed(g["weight_seed"])
    m = RobustENeRF(dataset_directory=dataset_dir, **H.product_config(g))
    want = {k for k in g["grads"] if k.startswith("nerf.radiance_field.")}
    have = {k for k, _ in m.named_parameters() if k.startswith("nerf.radiance_field.")}
    assert want == have
    init = R.init_field_params(H.golden_configs(g)[0], seed=g["weight_seed"])
    sd = m.state_dict()
    for k, v in init.items():
        assert torch.equal(sd["nerf.radiance_field." + k], v), k
    groups = m.configure_optimizers()["optimizer"].param_groups
    assert len(groups[2]["params"]) == len(init)               # grid + MLPs all carry the L2 weight decay


def test_device_sampler_distributions():
    """`ren_b200.data.DeviceEventSampler` (on CPU here: same code path, `device` argument only): batch layout,
    dtypes and the three normalised-timestamp distributions of data/datamodule.py:149-196."""
    from ren_b200.data import DeviceEventSampler, trunc_normal
    M = 1000
    ev = {"position": torch.randint(0, 64, (M, 2)), "start_ts": torch.arange(M) * 10, "end_ts": torch.arange(M) * 10 + 5,
          "num_pos": torch.ones(M, dtype=torch.int64), "num_neg": torch.zeros(M, dtype=torch.int64)}
    s = DeviceEventSampler(ev, 4096, device="cpu", seed=3, rank=1)
    b = next(iter(s))
    assert b["event"]["position"].shape == (1, 4096, 2) and b["event"]["position"].dtype == torch.float32
    assert b["event"]["start_ts"].dtype == torch.int64 and b["event"]["end_ts"].shape == (1, 4096)
    assert bool((b["event"]["end_ts"] - b["event"]["start_ts"] == 5).all())        # rows stay aligned
    nz = b["normalized"]
    assert all(v.dtype == torch.float64 and v.shape == (1, 4096) for v in nz.values())
    assert bool((nz["ts_diff"] == 1).all())
    assert 0.45 < float(nz["diff_start_ts"].mean()) < 0.55 and float(nz["diff_start_ts"].min()) >= 0
    g = nz["grad_ts"]
    assert float(g.min()) >= 0 and float(g.max()) <= 1 and abs(float(g.mean()) - 0.5) < 0.02
    assert 0.18 < float(g.std()) < 0.24                       # N(0.5, 0.25) truncated to +-2 sigma: std ~0.22
    s.batch_size = 17                                          # dynamic batch size
    assert next(s)["event"]["end_ts"].shape == (1, 17)
    assert float(trunc_normal(0, 1, (10,), 0.5, 0.25, dtype=torch.float64).max()) <= 1


def test_affine_log_intensity_alignment():
    """`align_log_intensity` vs the reference's formulation (robust_e_nerf.py:641-674): lstsq over the
    (N*H*W, 2) design matrix [log pred | 1] per channel, float64."""
    import numpy as np
    from ren_b200 import RobustENeRF
    gen = torch.Generator().manual_seed(0)
    target = torch.rand(3, 2, 12, 16, generator=gen) * 0.9 + 0.05
    # an affine distortion in log space + noise, different per channel
    a = torch.tensor([0.7, 1.6]).view(1, 2, 1, 1)
    b = torch.tensor([-0.4, 0.9]).view(1, 2, 1, 1)
    pred = torch.exp((target.log() - b) / a + 0.01 * torch.randn(target.shape, generator=gen))
    aligned, scale, offset = RobustENeRF.align_log_intensity(pred, target)
    for c in range(2):
        A = np.stack([pred[:, c].log().double().numpy().ravel(), np.ones(3 * 12 * 16)], axis=1)
        x, *_ = np.linalg.lstsq(A, target[:, c].log().double().numpy().ravel(), rcond=None)
        assert abs(float(scale[c]) - x[0]) < 1e-10 and abs(float(offset[c]) - x[1]) < 1e-10
        want = np.exp((A @ x).astype(np.float32)).reshape(3, 12, 16)
        assert np.allclose(aligned[:, c].numpy(), want, rtol=1e-6, atol=0)
    assert torch.allclose(scale, a.reshape(-1).double(), atol=0.05)
    # exact affine relation -> exact recovery
    pred = torch.exp(2.0 * target.log() + 0.3)
    aligned, scale, offset = RobustENeRF.align_log_intensity(pred, target)
    assert torch.allclose(aligned, target, rtol=1e-5) and torch.allclose(scale, torch.full((2,), 0.5).double())
    # a constant prediction: lstsq's minimum-norm solution of the rank-1 system
    pred = torch.full_like(target, 0.5)
    _, scale, offset = RobustENeRF.align_log_intensity(pred, target)
    A = np.stack([np.full(3 * 12 * 16, float(torch.tensor(0.5).log())), np.ones(3 * 12 * 16)], axis=1)
    x, *_ = np.linalg.lstsq(A, target[:, 0].log().double().numpy().ravel(), rcond=None)
    assert abs(float(scale[0]) - x[0]) < 1e-9 and abs(float(offset[0]) - x[1]) < 1e-9


def test_dynamic_batch_size_follows_the_reference_rule(built_lib, dataset_dir):
    """robust_e_nerf.py:907-950: new batch = int(per-GPU ray-sample budget / mean samples per ray), written to the
    datamodule's event dataset and its normalised-timestamp samplers; with gradient accumulation only on the
    second-to-last micro-batch."""
    from types import SimpleNamespace as NS
    _, m = _model(dataset_dir, "step_plumbing")
    samplers = [NS(size=7), NS(size=7), NS(size=7)]
    dm = NS(train_dataset=NS(batch_size=7), train_normalized_sampler=NS(datasets=samplers))
    m.trainer = NS(accumulate_grad_batches=1, datamodule=dm)
    budget = m.train_ray_sample_batch_size
    out = m.update_train_batch_size(torch.tensor(37.5), 0)
    assert float(out) == 37.5
    assert dm.train_dataset.batch_size == int(budget / 37.5) and all(s.size == int(budget / 37.5) for s in samplers)
    # accumulation over 4 micro-batches: only index 2 (mod 4) resizes
    m.trainer.accumulate_grad_batches = 4
    for idx, expect_change in ((0, False), (1, False), (2, True), (3, False), (6, True)):
        dm.train_dataset.batch_size = -1
        m.update_train_batch_size(torch.tensor(100.0), idx)
        assert (dm.train_dataset.batch_size == int(budget / 100.0)) == expect_change, idx
    m.trainer = None


def test_epoch_end_writes_prediction_pngs(built_lib, dataset_dir, tmp_path):
    """robust_e_nerf.py:733-779: aligned predictions -> 8-bit PNGs named by sample id under <log_dir>/predictions."""
    import cv2
    import numpy as np
    from types import SimpleNamespace as NS
    _, m = _model(dataset_dir, "step_plumbing")
    m.eval_save_pred_intensity_img = True
    m.trainer = NS(log_dir=str(tmp_path))
    m.set_eval_stage("val", [[10.0, 0, 3.5], [0, 10.0, 2.5], [0, 0, 1]], 6, 8, 0.5 / 256, 1 - 0.5 / 256)
    gen = torch.Generator().manual_seed(0)
    target = torch.rand(2, 6, 8, generator=gen) * 0.9 + 0.05
    pred = torch.exp(1.3 * target.log() + 0.2)                       # affine in log space: aligned == target
    ids = [torch.tensor([ord(c) for c in name.ljust(16)]) for name in ("r_000", "r_001")]
    outs = [{"pred_intensity_img": pred[i], "target_intensity_img": target[i], "sample_id": ids[i]} for i in range(2)]
    res = m.validation_epoch_end(outs)
    assert float(res["psnr"]) > 60
    lo, hi = 0.5 / 256, 1 - 0.5 / 256
    for i, name in enumerate(("r_000", "r_001")):
        path = tmp_path / "predictions" / f"{name}.png"
        assert path.is_file()
        got = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
        want = np.round(255 * np.clip((res["pred_intensity_img"][i, 0].numpy() - lo) / (hi - lo), 0, 1)).astype(np.uint8)
        assert got.dtype == np.uint8 and got.shape == (6, 8) and np.array_equal(got, want)
    m.trainer = None


def test_ssim_properties():
    """`RobustENeRF.ssim` (the torchmetrics 0.6.2 definition, restated): identity, symmetry, monotone in noise, and
    equal to an independent scipy.ndimage evaluation of the same windowed statistics."""
    import numpy as np
    from scipy import ndimage
    from ren_b200 import RobustENeRF
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(2, 1, 40, 48, generator=gen, dtype=torch.float64)
    assert abs(float(RobustENeRF.ssim(x, x, 1.0)) - 1.0) < 1e-12
    noisy = [(x + s * torch.randn(x.shape, generator=gen, dtype=torch.float64)).clamp(0, 1) for s in (0.02, 0.1, 0.3)]
    vals = [float(RobustENeRF.ssim(n, x, 1.0)) for n in noisy]
    assert 1 > vals[0] > vals[1] > vals[2] > 0
    assert abs(float(RobustENeRF.ssim(noisy[1], x, 1.0)) - float(RobustENeRF.ssim(x, noisy[1], 1.0))) < 1e-12
    # independent evaluation: Gaussian-window means with mirror ("reflect" without edge repeat) boundaries
    ax = np.arange(11) - 5.0
    g = np.exp(-(ax / 1.5) ** 2 / 2)
    g /= g.sum()
    k = np.outer(g, g)
    f = lambda a: ndimage.correlate(a, k, mode="mirror")
    total = []
    for i in range(2):
        p, t = noisy[1][i, 0].numpy(), x[i, 0].numpy()
        mp, mt = f(p), f(t)
        vp, vt, cv = f(p * p) - mp * mp, f(t * t) - mt * mt, f(p * t) - mp * mt
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        m = ((2 * mp * mt + c1) * (2 * cv + c2)) / ((mp * mp + mt * mt + c1) * (vp + vt + c2))
        total.append(m[5:-5, 5:-5].mean())
    assert abs(float(RobustENeRF.ssim(noisy[1], x, 1.0)) - float(np.mean(total))) < 1e-10


def test_trainer_flat_gradient_detection(built_lib, dataset_dir):
    """`FlatTrainer._flat_mlp_grad`: the fused backward hands out the MLP gradients as views of one flat buffer;
    the trainer copies that buffer once, and falls back to per-parameter copies when the views were replaced."""
    from ren_b200.trainer import FlatTrainer
    _, m = _model(dataset_dir, "step_plumbing")
    tr = FlatTrainer(m, use_fused_adam=False)
    params = m.nerf.mlp_parameters()
    n = sum(p.numel() for p in params)
    flat = torch.arange(n, dtype=torch.float32) * 1e-3
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view(p.shape)
        off += p.numel()
    got = tr._flat_mlp_grad()
    assert got is not None and got.data_ptr() == flat.data_ptr() and torch.equal(got, flat)
    tr._gather_grads()
    assert torch.equal(tr.gbuf[:n], flat)
    # one gradient replaced by a fresh tensor (e.g. a second backward accumulated out of place)
    params[3].grad = params[3].grad.clone()
    assert tr._flat_mlp_grad() is None
    tr.gbuf.zero_()
    tr._gather_grads()
    assert torch.equal(tr.gbuf[:n], flat)                        # the per-parameter path gives the same vector
    params[0].grad = None
    assert tr._flat_mlp_grad() is None
