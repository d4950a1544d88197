"""Shared test helpers: golden loading, product-model construction from a golden spec."""
import os

import torch

from oracle import restatement as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = ["step_plumbing", "step_8x256", "step_8x256_gain2", "step_aabb_cull"]
NGP_GOLDEN_NAMES = ["step_ngp"]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def golden_configs(g):
    return R.FieldConfig(**g["fcfg"]), R.RenderConfig(**g["rcfg"]), R.StepConfig(**g["scfg"])


def golden_params(g, dtype=torch.float32):
    """MLP weights of a golden case: seeded default init (+ recorded overrides), verified
    against the stored checksums."""
    fcfg, _, _ = golden_configs(g)
    params = R.init_field_params(fcfg, seed=g["weight_seed"], dtype=torch.float32)
    if g.get("weight_gain") is not None:
        for k in params:
            if k.endswith(".weight"):
                params[k] = params[k] * g["weight_gain"]
    if g.get("table_gain") is not None:
        params["mlp_base.0.params"] = params["mlp_base.0.params"] * g["table_gain"]
    if g.get("sigma_bias_override") is not None:
        params["mlp.sigma_layer.output_layer.bias"].fill_(g["sigma_bias_override"])
    for k, (s, a) in g["weight_checksums"].items():
        v = params[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(s)), f"weight RNG drift in {k}"
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), f"weight RNG drift in {k}"
    return {k: v.to(dtype) for k, v in params.items()}


def product_config(g):
    """ren_b200.RobustENeRF constructor kwargs equivalent to the golden's reference hparams
    (oracle/ref_harness.py:build_reference_model)."""
    f, r, s = g["fcfg"], g["rcfg"], g["scfg"]
    nerf = dict(
        aabb=list(f["aabb"]), contraction_type=f["contraction"],
        occ_grid=dict(resolution=8, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, enabled=False),
        near_plane=r["near_plane"], far_plane=r["far_plane"], render_step_size=r["render_step_size"],
        cone_angle=r["cone_angle"], early_stop_eps=r["early_stop_eps"], alpha_thre=r["alpha_thre"],
        test_chunk_size=16384, arch=f.get("arch", "mlp"), load_state_dict=False, freeze=False,
        ngp=dict(
            pos_encoding=dict(otype={"Hash": "HashGrid", "Dense": "DenseGrid", "Tiled": "TiledGrid"}[f.get("grid_type", "Hash")],
                              n_levels=f.get("n_levels", 16), n_features_per_level=f.get("n_features_per_level", 2),
                              log2_hashmap_size=f.get("log2_hashmap_size", 19), base_resolution=f.get("base_resolution", 16),
                              per_level_scale=f.get("per_level_scale", 1.4472692012786865), interpolation="Linear"),
            dir_encoding=dict(degree=f.get("sh_degree", 4)),
            mlp_base=dict(hidden_activation=f["hidden_activation"], density_activation=f["density_activation"],
                          n_neurons=f.get("base_n_neurons", 64), n_hidden_layers=f.get("base_n_hidden_layers", 1),
                          geo_feat_dim=f.get("geo_feat_dim", 15), weight_norm=False),
            mlp_head=dict(hidden_activation=f.get("head_hidden_activation") or f["hidden_activation"],
                          radiance_activation=f["radiance_activation"], n_neurons=f.get("head_n_neurons", 64),
                          n_hidden_layers=f.get("head_n_hidden_layers", 2), weight_norm=False)),
        mlp=dict(net_depth=f["net_depth"], net_width=f["net_width"], skip_layer=f["skip_layer"],
                 net_depth_condition=f["net_depth_condition"],
                 net_width_condition=f["net_width_condition"],
                 hidden_activation=f["hidden_activation"], density_activation=f["density_activation"],
                 radiance_activation=f["radiance_activation"], pos_encoder_max_deg=f["pos_deg"],
                 view_encoder_max_deg=f["dir_deg"], weight_norm=False))
    return dict(
        git_head_hash="test", eval_target=["novel_view"], num_nodes=1, gpus=[0],
        min_modeled_intensity=s["min_modeled_intensity"], eval_save_pred_intensity_img=False,
        checkpoint_filepath=None,
        contrast_threshold=dict(load_state_dict=False, freeze=g["freeze_ct"]),
        refractory_period=dict(load_state_dict=False, freeze=g["freeze_rp"]), nerf=nerf,
        loss=dict(error_fn=dict(log_intensity_grad=s["err_grad"], log_intensity_diff=s["err_diff"]),
                  weight=dict(log_intensity_grad=s["w_grad"], log_intensity_diff=s["w_diff"],
                              nerf_mlp_weight_decay=1e-6),
                  param_weight=dict(log_intensity_grad=s["pw_grad"], log_intensity_diff=s["pw_diff"])),
        metric=dict(lpips_net="alex"),
        optimizer=dict(algo="adam", lr=dict(contrast_threshold=0.1, default=0.01),
                       relative_lr=dict(refractory_period=50)),
        lr_scheduler=dict(algo="multi_step_lr", interval="epoch",
                          multi_step_lr=dict(milestones=[20, 30, 36], gamma=0.33)),
        alpha_over_white_bg=g["learn_bkgd"], train_eff_ray_sample_batch_size=1 << 20)


def build_product(g, dataset_dir, device="cpu", engine="auto"):
    from ren_b200 import RobustENeRF
    model = RobustENeRF(dataset_directory=dataset_dir, engine=engine, **product_config(g))
    params = golden_params(g)
    sd = {"nerf.radiance_field." + k: v for k, v in params.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return model.to(device).train()


def rel_err(a, b, floor=1e-3):
    """(normwise rel, max abs, max elementwise rel with |ref| floored)."""
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    nw = float((a - b).norm() / b.norm().clamp(min=1e-30))
    ma = float((a - b).abs().max())
    er = float(((a - b).abs() / b.abs().clamp(min=floor)).max())
    return nw, ma, er
