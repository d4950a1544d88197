"""ORACLE (test infrastructure, build container only): drive the UNMODIFIED
reference `RobustENeRF` (robust_e_nerf/models/robust_e_nerf.py:16) on CPU through
the shims, on a synthetic dataset directory written by `oracle.synth`.

Used by `oracle/make_golden.py` (mints tests/golden/*.pt) and by
tests/test_oracle_vs_reference.py (pins `oracle.restatement` to the reference).
Needs /root/reference, so it never runs on the GPU box.
"""
import torch

from . import ref_loader
from .restatement import FieldConfig, RenderConfig, StepConfig


def build_reference_model(dataset_dir, fcfg: FieldConfig, rcfg: RenderConfig, scfg: StepConfig,
                          learn_bkgd=False, freeze_ct=True, freeze_rp=True, seed=0,
                          train_eff_ray_sample_batch_size=1 << 20):
    """Instantiate the reference LightningModule with the YAML-equivalent hparams
    (configs/train/synthetic.yaml:24-128) for `arch: mlp` or `arch: ngp`."""
    ref = ref_loader.load_reference()
    from easydict import EasyDict
    import pytorch_lightning as pl

    nerf_cfg = EasyDict(
        aabb=list(fcfg.aabb), contraction_type=fcfg.contraction,
        occ_grid=dict(resolution=8, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16),
        near_plane=rcfg.near_plane, far_plane=rcfg.far_plane,
        render_step_size=rcfg.render_step_size, cone_angle=rcfg.cone_angle,
        early_stop_eps=rcfg.early_stop_eps, alpha_thre=rcfg.alpha_thre,
        test_chunk_size=16384, arch=fcfg.arch, load_state_dict=False, freeze=False,
        # configs/train/*.yaml:61-88
        ngp=dict(
            pos_encoding=dict(otype={"Hash": "HashGrid", "Dense": "DenseGrid", "Tiled": "TiledGrid"}[fcfg.grid_type],
                              n_levels=fcfg.n_levels, n_features_per_level=fcfg.n_features_per_level,
                              log2_hashmap_size=fcfg.log2_hashmap_size, base_resolution=fcfg.base_resolution,
                              per_level_scale=fcfg.per_level_scale, interpolation="Linear"),
            dir_encoding=dict(degree=fcfg.sh_degree),
            mlp_base=dict(hidden_activation=fcfg.hidden_activation, density_activation=fcfg.density_activation,
                          n_neurons=fcfg.base_n_neurons, n_hidden_layers=fcfg.base_n_hidden_layers,
                          geo_feat_dim=fcfg.geo_feat_dim, weight_norm=False),
            mlp_head=dict(hidden_activation=fcfg.head_hidden_activation or fcfg.hidden_activation,
                          radiance_activation=fcfg.radiance_activation, n_neurons=fcfg.head_n_neurons,
                          n_hidden_layers=fcfg.head_n_hidden_layers, weight_norm=False)),
        mlp=dict(net_depth=fcfg.net_depth, net_width=fcfg.net_width, skip_layer=fcfg.skip_layer,
                 net_depth_condition=fcfg.net_depth_condition,
                 net_width_condition=fcfg.net_width_condition,
                 hidden_activation=fcfg.hidden_activation,
                 density_activation=fcfg.density_activation,
                 radiance_activation=fcfg.radiance_activation,
                 pos_encoder_max_deg=fcfg.pos_deg, view_encoder_max_deg=fcfg.dir_deg,
                 weight_norm=False),
    )
    loss_cfg = EasyDict(
        error_fn=dict(log_intensity_grad=scfg.err_grad, log_intensity_diff=scfg.err_diff),
        weight=dict(log_intensity_grad=scfg.w_grad, log_intensity_diff=scfg.w_diff,
                    nerf_mlp_weight_decay=1e-6),
        param_weight=dict(log_intensity_grad=scfg.pw_grad, log_intensity_diff=scfg.pw_diff),
    )
    torch.manual_seed(seed)
    model = ref.models.robust_e_nerf.RobustENeRF(
        git_head_hash="oracle", eval_target=["novel_view"], num_nodes=1, gpus=[0],
        min_modeled_intensity=scfg.min_modeled_intensity, eval_save_pred_intensity_img=False,
        checkpoint_filepath=None,
        contrast_threshold=EasyDict(load_state_dict=False, freeze=freeze_ct),
        refractory_period=EasyDict(load_state_dict=False, freeze=freeze_rp),
        nerf=nerf_cfg, loss=loss_cfg, metric=EasyDict(lpips_net="alex"),
        optimizer=EasyDict(algo="adam", lr=dict(contrast_threshold=0.1, default=0.01),
                           relative_lr=dict(refractory_period=50)),
        lr_scheduler=EasyDict(algo="multi_step_lr", interval="epoch",
                              multi_step_lr=dict(milestones=[20, 30, 36], gamma=0.33)),
        dataset_directory=dataset_dir, alpha_over_white_bg=learn_bkgd,
        train_eff_ray_sample_batch_size=train_eff_ray_sample_batch_size,
    )

    class _Sampler:
        size = 0

    class _DS:
        batch_size = 0

    class _DM:
        train_dataset = _DS()

        class train_normalized_sampler:
            datasets = [_Sampler(), _Sampler(), _Sampler()]

    model.trainer = pl.Trainer(datamodule=_DM())
    model.train()
    return model


def reference_training_step(model, batch, jitter, scfg: StepConfig):
    """Run the reference's own `training_step` with injected per-ray jitter (order
    of renders: grad, diff-start, diff-end -- robust_e_nerf.py:383-424)."""
    import nerfacc

    order = []
    if scfg.w_grad > 0:
        order.append(jitter["grad"])
    if scfg.w_diff > 0:
        order += [jitter["diff_start"], jitter["diff_end"]]
    queue = list(order)

    def jitter_fn(n, device, dtype):
        u = queue.pop(0)
        assert u.shape[0] == n
        return u.to(device=device, dtype=dtype)

    nerfacc.set_jitter_fn(jitter_fn)
    try:
        batch = {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in batch.items()}
        loss = model.training_step(batch, 0)
    finally:
        nerfacc.set_jitter_fn(None)
    assert not queue
    return loss


def reference_render(model, o, d, jitter):
    """The reference's `NeRF.forward` (models/nerf.py:230-286) with injected jitter."""
    import nerfacc

    nerfacc.set_jitter_fn(lambda n, device, dtype: jitter.to(device=device, dtype=dtype))
    try:
        return model.nerf(o, d)
    finally:
        nerfacc.set_jitter_fn(None)
