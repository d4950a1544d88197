"""`RobustENeRF` task module on top of libren_b200.so -- same constructor, hooks,
hyper-parameter keys, logged scalars and state-dict keys as the reference's
`robust_e_nerf.models.robust_e_nerf.RobustENeRF`
(robust_e_nerf/models/robust_e_nerf.py:16-961), so `scripts/run.py` can instantiate it
(robust_e_nerf/scripts/run.py:45-63).  Scope: the training hot path;
validation / test image metrics are out of scope.

A training step is: one `ren_event_rays` launch (refractory correction, supervision
timestamps, pose lerp/slerp, ray generation for both renders), ONE `ren_render_forward`
over the concatenated start|end rays, one `ren_event_diff_loss` launch; backward is
`ren_render_backward`.  The sensor parameters stay tiny PyTorch scalars.
"""
import ctypes as C
import math
import os
import warnings

import numpy as np
import torch

from . import _lib
from ._pl import LightningModule
from .nerf import NeRF


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


# ------------------------------------------------------------------ dataset files
def load_camera_calibration(dataset_directory):
    """`datasets.Event.load_camera_calibration` (robust_e_nerf/data/datasets.py:15-34,66-80)."""
    return dict(np.load(os.path.join(dataset_directory, "camera_calibration.npz")))


def load_camera_poses(dataset_directory):
    """`datasets.CameraPose.load_camera_poses` (robust_e_nerf/data/datasets.py:714-727)."""
    poses = np.load(os.path.join(dataset_directory, "camera_poses.npz"))
    return {k: torch.tensor(poses[k]) for k in ("T_wc_position", "T_wc_orientation", "T_wc_timestamp")}


# ------------------------------------------------------------------ parametrisations
class Softplus(torch.nn.Module):
    """utils/modules.py:38-55"""

    def __init__(self, beta=1, threshold=20):
        super().__init__()
        self.beta, self.threshold = beta, threshold

    def forward(self, x):
        return torch.nn.functional.softplus(x, self.beta, self.threshold)

    def right_inverse(self, x):
        inv = torch.log(torch.exp(self.beta * x) - 1) / self.beta
        return torch.where(x * self.beta > self.threshold, x, inv)


class ScaledShiftedSigmoid(torch.nn.Module):
    """utils/modules.py:58-74"""

    def __init__(self, low=0, high=1):
        super().__init__()
        self.low, self.scale = low, high - low

    def forward(self, x):
        return self.scale * torch.sigmoid(x / self.scale) + self.low

    def right_inverse(self, x):
        return self.scale * torch.logit((x - self.low) / self.scale)


class ContrastThreshold(torch.nn.Module):
    """models/event_generation_params.py:8-84 (state-dict key
    `parametrizations.p2n_contrast_threshold_ratio.original`)."""

    def __init__(self, dataset_directory):
        super().__init__()
        calib = load_camera_calibration(dataset_directory)
        pos = torch.from_numpy(calib["pos_contrast_threshold"])
        neg = torch.from_numpy(calib["neg_contrast_threshold"])
        ratio = pos / neg
        assert ratio > 0
        self.register_buffer("init_p2n_contrast_threshold_ratio", ratio, persistent=False)
        self.register_buffer("neg_contrast_threshold", neg, persistent=False)
        self.p2n_contrast_threshold_ratio = torch.nn.Parameter(ratio.clone())
        torch.nn.utils.parametrize.register_parametrization(
            self, "p2n_contrast_threshold_ratio", Softplus(beta=1))

    @property
    def pos_contrast_threshold(self):
        return self.p2n_contrast_threshold_ratio * self.neg_contrast_threshold

    @property
    def mean_contrast_threshold(self):
        return (self.pos_contrast_threshold + self.neg_contrast_threshold) / 2

    def forward(self, num_pos, num_neg):
        """Effective log-intensity difference of the event interval (:72-84)."""
        return num_pos * self.pos_contrast_threshold - num_neg * self.neg_contrast_threshold


class RefractoryPeriod(torch.nn.Module):
    """models/event_generation_params.py:87-203 (f64 parameter, state-dict key
    `parametrizations._refractory_period.original`)."""
    MIN_SCALED_SHIFTED_SIGMOID_GRAD_MAGNITUDE = 0.0001
    REDEFINED_CALIBRATED_REFRACTORY_PERIOD_FACTOR = 0.999

    def __init__(self, dataset_directory):
        super().__init__()
        calib = load_camera_calibration(dataset_directory)
        tau = torch.from_numpy(np.asarray(calib["refractory_period"]))
        # event_generation_params.py:101-114: the cached maximum refractory period, else extracted from the raw
        # event stream (`ren_event_queue` on the current CUDA device) and cached for the next run
        from . import events as _events
        tau_max = _events.Event.load_max_refractory_period(dataset_directory)
        if tau_max is None:
            raw_path = os.path.join(dataset_directory, _events.RAW_EVENTS_FILENAME)
            if not os.path.isfile(raw_path):
                raise FileNotFoundError(
                    f"{dataset_directory}: neither {_events.MAX_REFRACTORY_PERIOD_FILENAME} nor "
                    f"{_events.RAW_EVENTS_FILENAME} (robust_e_nerf/data/datasets.py:97-130)")
            tau_max = _events.extract_max_refractory_period(_events.load_raw_events(dataset_directory), calib)
            _events.Event.save_max_refractory_period(tau_max, dataset_directory)
        if not (0 <= tau < tau_max):
            warnings.warn(f"Calibrated refractory period ({tau}) >= Max. possible refractory period ({tau_max}).")
            tau = self.REDEFINED_CALIBRATED_REFRACTORY_PERIOD_FACTOR * tau_max
            warnings.warn(f"Redefining calibrated refractory period to "
                          f"{self.REDEFINED_CALIBRATED_REFRACTORY_PERIOD_FACTOR} of max. possible refractory "
                          f"period ({tau}).")
        self.register_buffer("init_refractory_period", tau, persistent=False)
        self.register_buffer("max_refractory_period", tau_max, persistent=False)
        self.register_buffer(
            "max_scaled_logit_magnitude",
            torch.tensor(self.MIN_SCALED_SHIFTED_SIGMOID_GRAD_MAGNITUDE).logit().abs(), persistent=False)
        self._refractory_period = torch.nn.Parameter(tau.to(torch.float64))
        torch.nn.utils.parametrize.register_parametrization(
            self, "_refractory_period", ScaledShiftedSigmoid(low=0, high=tau_max))
        self.clamp_refractory_period()

    @torch.no_grad()
    def clamp_refractory_period(self):
        orig = self.parametrizations._refractory_period.original
        scaled = orig / self.max_refractory_period
        clamped = scaled.clamp(min=-self.max_scaled_logit_magnitude, max=self.max_scaled_logit_magnitude)
        orig.copy_(self.max_refractory_period * clamped)

    @property
    def refractory_period(self):
        self.clamp_refractory_period()
        return self._refractory_period


class LinearTrajectory(torch.nn.Module):
    """models/trajectories.py:8-91: continuous-time pose by lerp + slerp (`ren_trajectory`)."""

    def __init__(self, camera_poses):
        super().__init__()
        self.register_buffer("T_wc_timestamp", camera_poses["T_wc_timestamp"].to(torch.int64).contiguous(),
                             persistent=False)
        self.register_buffer("T_wc_position", camera_poses["T_wc_position"].float().contiguous(),
                             persistent=False)
        self.register_buffer("T_wc_orientation", camera_poses["T_wc_orientation"].float().contiguous(),
                             persistent=False)

    def forward(self, ts):
        ts = ts.detach().to(torch.float64).contiguous().reshape(-1)
        n = ts.numel()
        pos = torch.empty((n, 3), device=ts.device, dtype=torch.float32)
        rot = torch.empty((n, 3, 3), device=ts.device, dtype=torch.float32)
        _lib.check(_lib.load().ren_trajectory(
            n, _lib.ptr(ts), self.T_wc_timestamp.numel(), _lib.ptr(self.T_wc_timestamp),
            _lib.ptr(self.T_wc_position), _lib.ptr(self.T_wc_orientation), _lib.ptr(pos), _lib.ptr(rot),
            _lib.stream_ptr()), "ren_trajectory")
        return pos, rot


# ------------------------------------------------------------------ loss op
class DiffLossFunction(torch.autograd.Function):
    """`ren_event_diff_loss`: masked-mean event loss on the two rendered radiances.
    Differentiable wrt c0, c1 (kernel) and ev_diff (the contrast-threshold path)."""

    @staticmethod
    def forward(ctx, c0, c1, o0, o1, ev_diff, ts_diff, start, end_ts, error_fn, all_valid, imin):
        B = c0.numel()
        dev = c0.device
        c0, c1, o0, o1 = (t.contiguous().float() for t in (c0, c1, o0, o1))
        ev_diff = ev_diff.contiguous().float()
        log_i0 = torch.empty_like(c0)
        log_i1 = torch.empty_like(c1)
        g_c0 = torch.empty_like(c0)
        g_c1 = torch.empty_like(c1)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        n_valid = torch.empty((), device=dev, dtype=torch.int32)
        scratch = torch.empty((2,), device=dev, dtype=torch.float64)
        _lib.check(_lib.load().ren_event_diff_loss(
            B, _lib.ERROR_FN[error_fn], int(all_valid), float(imin), _lib.ptr(c0), _lib.ptr(c1),
            _lib.ptr(o0), _lib.ptr(o1), _lib.ptr(ev_diff), _lib.ptr(ts_diff), _lib.ptr(start),
            _lib.ptr(end_ts), _lib.ptr(log_i0), _lib.ptr(log_i1), _lib.ptr(loss), _lib.ptr(n_valid),
            _lib.ptr(g_c0), _lib.ptr(g_c1), _lib.ptr(scratch), _lib.stream_ptr()), "ren_event_diff_loss")
        ctx.save_for_backward(g_c0, g_c1, c1, ts_diff, start, end_ts, log_i0, log_i1, ev_diff, o0, o1, n_valid)
        ctx.imin = imin
        ctx.error_fn = error_fn
        ctx.all_valid = bool(all_valid)
        ctx.mark_non_differentiable(log_i0, log_i1, n_valid)
        return loss, log_i0, log_i1, n_valid

    @staticmethod
    def backward(ctx, g_loss, _a, _b, _c):
        _lib.refuse_double_backward()
        g_c0, g_c1, c1, ts_diff, start, end_ts, log_i0, log_i1, ev_diff, o0, o1, n_valid = ctx.saved_tensors
        g_ev = None
        if ctx.needs_input_grad[4]:
            # tgt = f32(ts_diff * ev / (end - start)).  l1 / mse depend on (pred - tgt) only, so
            # d loss / d tgt = -(d loss / d L1) with L1 = log(c1 + imin); mape = |pred - tgt| / max(|tgt|, eps)
            # has the target in the denominator as well (utils/modules.py:96)
            dtgt_dev = ts_diff / (end_ts.double() - start)
            g_tgt = -(g_c1 * (c1 + ctx.imin)).double()
            if ctx.error_fn == "mape":
                eps = 2.220446049250313e-16
                tgt = (ev_diff.double() * dtgt_dev).float().double()
                pred = (log_i1 - log_i0).double()
                valid = torch.ones_like(o0, dtype=torch.bool) if ctx.all_valid else ((o0 > 0) | (o1 > 0))
                extra = -(pred - tgt).abs() * torch.sign(tgt) / (tgt * tgt)
                extra = torch.where(valid & (tgt.abs() > eps), extra, torch.zeros_like(extra))
                g_tgt = g_tgt + extra / n_valid.double()
            g_ev = (g_tgt * dtgt_dev * g_loss.double()).float()
        return (g_c0 * g_loss, g_c1 * g_loss, None, None, g_ev, None, None, None, None, None, None)


class RayChainFunction(torch.autograd.Function):
    """Ties the rays `ren_event_rays` produced to the (differentiable, f64) supervision
    timestamp they were generated at, so that d loss / d ts -- and from there the refractory
    period -- follows the reference's autograd path through `LinearTrajectory.forward` and
    `pixel_params_to_ray` (robust_e_nerf/models/trajectories.py:63-91, models/nerf.py:206-228):
        d loss/d ts = (g_o . o' + g_d . d' + g_ddot . d'') * dw/dt      (o'' = 0 for a lerp)
    with ' = d/dw taken by second-order jets inside the kernel."""

    @staticmethod
    def forward(ctx, ts, o, d, o_dot, d_dot, d_ddot, dw_dt):
        ctx.save_for_backward(o_dot, d_dot, d_ddot, dw_dt)
        return o.clone(), d.clone(), o_dot.clone(), d_dot.clone()

    @staticmethod
    def backward(ctx, g_o, g_d, g_odot, g_ddot):
        _lib.refuse_double_backward()
        o_dot, d_dot, d_ddot, dw_dt = ctx.saved_tensors
        g_w = torch.zeros_like(dw_dt)
        if g_o is not None:
            g_w = g_w + (g_o * o_dot).sum(-1).double()
        if g_d is not None:
            g_w = g_w + (g_d * d_dot).sum(-1).double()
        if g_ddot is not None:
            g_w = g_w + (g_ddot * d_ddot).sum(-1).double()
        return g_w * dw_dt, None, None, None, None, None, None


# ------------------------------------------------------------------ task module
class RobustENeRF(LightningModule):
    INTRINSICS_KEY = "intrinsics"
    NUM_DIM = 3
    MAX_NUM_SAMPLES_PER_RAY = 1024                                   # robust_e_nerf.py:19
    MODEL_COMPONENTS = ["contrast_threshold", "refractory_period", "nerf"]

    def __init__(self, git_head_hash, eval_target, num_nodes, gpus, min_modeled_intensity,
                 eval_save_pred_intensity_img, checkpoint_filepath, contrast_threshold,
                 refractory_period, nerf, loss, metric, optimizer, lr_scheduler, dataset_directory,
                 alpha_over_white_bg, train_eff_ray_sample_batch_size, engine="auto"):
        super().__init__()
        assert isinstance(min_modeled_intensity, (int, float)) and min_modeled_intensity > 0
        for cfg in (contrast_threshold, refractory_period, nerf):
            assert isinstance(_get(cfg, "load_state_dict"), bool)
            assert isinstance(_get(cfg, "freeze"), bool)
        assert isinstance(train_eff_ray_sample_batch_size, int) and train_eff_ray_sample_batch_size > 0
        if _get(nerf, "freeze"):
            assert _get(nerf, "load_state_dict")
        num_gpus = num_nodes * len(gpus)
        self.train_ray_sample_batch_size = train_eff_ray_sample_batch_size // num_gpus   # :63-66
        self.eval_save_pred_intensity_img = eval_save_pred_intensity_img
        self.engine = engine

        calib = load_camera_calibration(dataset_directory)
        self.has_bayer_filter = str(calib["bayer_pattern"]) != ""
        self.register_buffer("train_intrinsics_inv",
                             torch.linalg.inv(torch.from_numpy(calib[self.INTRINSICS_KEY])).float(),
                             persistent=False)
        self.render_bkgd = "parameter" if alpha_over_white_bg else None          # :155-159
        self.save_hyperparameters("git_head_hash", "min_modeled_intensity", "checkpoint_filepath",
                                  "contrast_threshold", "refractory_period", "nerf", "loss", "metric",
                                  "optimizer", "lr_scheduler")
        self.contrast_threshold = ContrastThreshold(dataset_directory)
        self.refractory_period = RefractoryPeriod(dataset_directory)
        camera_poses = load_camera_poses(dataset_directory)
        self.nerf = self._build_nerf(camera_poses, gpus)
        self.trajectory = LinearTrajectory(camera_poses)
        self._load_model_component_state_dicts()
        self._freeze_model_components()
        self._load_eval_stages(dataset_directory, eval_target)
        lw = _get(loss, "weight")
        for v in (lw.values() if isinstance(lw, dict) else vars(lw).values()):
            assert isinstance(v, (int, float)) and v >= 0                         # loss.py:16-20
        self._opt_state = None

    def _load_eval_stages(self, dataset_directory, eval_target):
        """robust_e_nerf.py:86-152: frame size, intrinsics and target value range of the validation / test
        renders come from the posed target images (`event_view`: the training views for both stages;
        `novel_view`: `transforms_val.json` / `transforms_test.json`).  A dataset without a `views` folder (the
        synthetic benchmark streams) simply has no evaluation stage until `set_eval_stage` is called; like the
        reference, a missing test split is not an error."""
        from .posed_images import PosedImage
        if set(eval_target) == {"event_view"}:
            stages = {"val": "train", "test": "train"}
        elif set(eval_target) == {"novel_view"}:
            stages = {"val": "val", "test": "test"}
        else:
            raise NotImplementedError(eval_target)
        if PosedImage.posed_img_folder_path(dataset_directory) is None:
            return
        loaded = {}
        for name, stage in stages.items():
            try:
                if stage not in loaded:
                    loaded[stage] = PosedImage(dataset_directory, stage, permutation_seed=None)
            except FileNotFoundError:
                if name == "val":
                    raise
                continue
            self.set_eval_stage_from_dataset(name, loaded[stage])

    def set_eval_stage_from_dataset(self, name, posed_images):
        h, w = posed_images.posed_imgs.img.shape[-2:]
        self.set_eval_stage(name, posed_images.posed_imgs.intrinsics, h, w,
                            float(posed_images.min_normalized_pixel_value),
                            float(posed_images.max_normalized_pixel_value))

    # -------------------------------------------------------------- builders
    def _build_nerf(self, camera_poses, gpus):
        """robust_e_nerf.py:204-260 -- THE drop-in boundary: returns `ren_b200.NeRF`."""
        cfg = self.hparams["nerf"]
        aabb = _get(cfg, "aabb")
        if isinstance(aabb, str) and aabb == "auto":
            p = camera_poses["T_wc_position"]
            aabb = torch.cat((p.min(dim=0).values, p.max(dim=0).values)).tolist()
        step = _get(cfg, "render_step_size")
        if isinstance(step, str) and step == "auto":
            ext = torch.tensor(aabb[3:]) - torch.tensor(aabb[:3])
            step = math.sqrt(self.NUM_DIM) * torch.max(ext).item() / self.MAX_NUM_SAMPLES_PER_RAY
        arch = _get(cfg, "arch")
        return NeRF(aabb, _get(cfg, "contraction_type"), _get(cfg, "occ_grid"), _get(cfg, "near_plane"),
                    _get(cfg, "far_plane"), step, self.render_bkgd, _get(cfg, "cone_angle"),
                    _get(cfg, "early_stop_eps"), _get(cfg, "alpha_thre"), _get(cfg, "test_chunk_size"),
                    arch, arch_config=_get(cfg, arch), num_dim=self.NUM_DIM,
                    radiance_dim=3 if self.has_bayer_filter else 1, engine=self.engine)

    def _load_model_component_state_dicts(self):
        """robust_e_nerf.py:265-286 + utils/modules.py:16-35 (partial component restore)."""
        if not any(_get(self.hparams[c], "load_state_dict") for c in self.MODEL_COMPONENTS):
            return
        ckpt = torch.load(self.hparams["checkpoint_filepath"], map_location="cpu")
        for c in self.MODEL_COMPONENTS:
            if _get(self.hparams[c], "load_state_dict"):
                prefix = c + "."
                sub = {k[len(prefix):]: v for k, v in ckpt["state_dict"].items() if k.startswith(prefix)}
                if c == "nerf" and getattr(self.nerf, "occupancy_grid", None) is None:
                    sub = {k: v for k, v in sub.items() if not k.startswith("occupancy_grid.")}
                getattr(self, c).load_state_dict(sub)

    def _freeze_model_components(self):
        for c in self.MODEL_COMPONENTS:
            if _get(self.hparams[c], "freeze"):
                for p in getattr(self, c).parameters():
                    p.requires_grad_(False)

    def forward(self, batch):
        pass

    # -------------------------------------------------------------- hot path
    def _loss_cfg(self, group, name):
        return _get(_get(self.hparams["loss"], group), name)

    def event_rays(self, event, normalized, n_renders=2, jets=False):
        """Steps b-d of SURVEY.md section 3.2 + `pixel_params_to_ray` in one launch.  `jets`: also
        return d/dw and d^2/dw^2 of the rays and dw/dt (tangent render, trainable tau)."""
        B = event["start_ts"].numel()
        dev = event["start_ts"].device
        tau = self.refractory_period.refractory_period.detach().reshape(1).contiguous()
        out_ts = torch.empty((n_renders, B), device=dev, dtype=torch.float64)
        out_o = torch.empty((n_renders, B, 3), device=dev, dtype=torch.float32)
        out_d = torch.empty((n_renders, B, 3), device=dev, dtype=torch.float32)
        out_rot = torch.empty((n_renders, B, 3, 3), device=dev, dtype=torch.float32)
        ts_diff = torch.empty((B,), device=dev, dtype=torch.float64)
        start = torch.empty((B,), device=dev, dtype=torch.float64)
        o_dot = d_dot = d_ddot = dw_dt = None
        if jets:
            o_dot = torch.empty((n_renders, B, 3), device=dev, dtype=torch.float32)
            d_dot = torch.empty((n_renders, B, 3), device=dev, dtype=torch.float32)
            d_ddot = torch.empty((n_renders, B, 3), device=dev, dtype=torch.float32)
            dw_dt = torch.empty((n_renders, B), device=dev, dtype=torch.float64)
        tr = self.trajectory
        c = lambda t: t.contiguous()
        _lib.check(_lib.load().ren_event_rays(
            B, _lib.ptr(c(event["position"].float())), _lib.ptr(c(event["start_ts"])),
            _lib.ptr(c(event["end_ts"])), _lib.ptr(c(normalized["ts_diff"])),
            _lib.ptr(c(normalized["diff_start_ts"])),
            _lib.ptr(c(normalized["grad_ts"])) if n_renders == 3 else None, _lib.ptr(tau),
            tr.T_wc_timestamp.numel(), _lib.ptr(tr.T_wc_timestamp), _lib.ptr(tr.T_wc_position),
            _lib.ptr(tr.T_wc_orientation), _lib.ptr(c(self.train_intrinsics_inv)), n_renders,
            _lib.ptr(out_ts), _lib.ptr(out_o), _lib.ptr(out_d), _lib.ptr(out_rot), _lib.ptr(ts_diff),
            _lib.ptr(start), _lib.ptr(o_dot), _lib.ptr(d_dot), _lib.ptr(d_ddot), _lib.ptr(dw_dt),
            _lib.stream_ptr()), "ren_event_rays")
        return dict(ts=out_ts, o=out_o, d=out_d, rot=out_rot, ts_diff=ts_diff, start=start,
                    o_dot=o_dot, d_dot=d_dot, d_ddot=d_ddot, dw_dt=dw_dt)

    def _supervision_timestamps(self, event, normalized, with_grad_ts):
        """robust_e_nerf.py:322-357 in differentiable f64 torch ((B,) ops): used only when the
        refractory period is trainable, to chain d loss / d ts to tau."""
        tau = self.refractory_period.refractory_period
        start = event["start_ts"] + tau
        end = event["end_ts"]
        ts_diff = (end - start) * normalized["ts_diff"]
        d_start = torch.lerp(start, torch.max(end - ts_diff, start), normalized["diff_start_ts"])
        d_end = torch.min(d_start + ts_diff, end)
        ts = [d_start, d_end]
        if with_grad_ts:
            ts = [torch.lerp(d_start, d_end, normalized["grad_ts"])] + ts
        return start, ts

    def render_pixels(self, intrinsics_inverse, pixel_position, T_wc_position, T_wc_orientation,
                      jitter=None):
        """robust_e_nerf.py:849-885."""
        o, d = self.nerf.pixel_params_to_ray(intrinsics_inverse, pixel_position, T_wc_position,
                                              T_wc_orientation)
        intensity, opacity, depth, spp = self.nerf(o, d, jitter)
        if intensity.dim() > opacity.dim():
            intensity = intensity.permute(-1, *range(opacity.dim()))
        intensity = intensity + self.hparams["min_modeled_intensity"]
        if self.render_bkgd is None:
            is_valid = opacity > 0
        else:
            is_valid = torch.ones_like(opacity, dtype=torch.bool)
        depth = depth * torch.sum(d * T_wc_orientation[..., 2], dim=-1)
        return intensity, opacity, depth, spp, is_valid

    def training_step(self, batch, batch_index, jitter=None):
        """robust_e_nerf.py:301-517.  `jitter` (optional dict with (B,) tensors `grad`,
        `diff_start`, `diff_end`) injects the per-ray stratified offsets (tests / parity runs)."""
        event = {k: v.squeeze(dim=0) for k, v in batch["event"].items()}
        normalized = {k: v.squeeze(dim=0) for k, v in batch["normalized"].items()}
        B = event["start_ts"].numel()
        for v in normalized.values():
            assert v.shape[0] == B
        w_grad = self._loss_cfg("weight", "log_intensity_grad")
        w_diff = self._loss_cfg("weight", "log_intensity_diff")
        assert w_grad > 0 or w_diff > 0                                                   # loss.py:16-20
        tau_train = any(p.requires_grad for p in self.refractory_period.parameters()) and torch.is_grad_enabled()
        dev = event["start_ts"].device
        imin = self.hparams["min_modeled_intensity"]
        all_valid = self.render_bkgd is not None
        # a: contrast threshold (tiny torch ops on (B,))
        ev_diff = self.contrast_threshold(event["num_pos"], event["num_neg"])
        # b-d + raygen: [grad |] diff_start | diff_end
        nr = 3 if w_grad > 0 else 2
        if w_diff <= 0:
            # l_grad alone (:346-357): the gradient timestamp is drawn over the whole corrected interval
            # [start + tau, end] -- the diff interval of normalised length 1 starting at 0
            normalized = dict(normalized, ts_diff=torch.ones_like(normalized["ts_diff"]),
                              diff_start_ts=torch.zeros_like(normalized["diff_start_ts"]))
        rays = self.event_rays(event, normalized, nr, jets=(w_grad > 0 or tau_train))
        acc = getattr(self.trainer, "accumulate_grad_batches", 1) if self.trainer is not None else 1
        if self.global_step is not None and batch_index % max(int(acc or 1), 1) == 0:     # :375-379
            self.nerf.update_occ_grid(self.global_step, self.trajectory.T_wc_position)
        ro = [rays["o"][j] for j in range(nr)]
        rd = [rays["d"][j] for j in range(nr)]
        rod = [rays["o_dot"][j] if rays["o_dot"] is not None else None for j in range(nr)]
        rdd = [rays["d_dot"][j] if rays["d_dot"] is not None else None for j in range(nr)]
        start = rays["start"]
        if tau_train:
            # differentiable timestamps -> rays (d loss / d ts chains to tau)
            start, ts = self._supervision_timestamps(event, normalized, w_grad > 0)   # (normalised samples as above)
            for j in range(nr):
                ro[j], rd[j], rod[j], rdd[j] = RayChainFunction.apply(
                    ts[j], rays["o"][j], rays["d"][j], rays["o_dot"][j], rays["d_dot"][j],
                    rays["d_ddot"][j], rays["dw_dt"][j])
        n_kept_all, opac_all = [], []
        losses = {}
        # f: gradient render (tangent mode) + log-intensity-gradient loss (f64, (B,) torch ops)
        if w_grad > 0:
            u = jitter["grad"].to(dev, torch.float32) if jitter is not None else None
            rad, opac_g, _d, rad_dot, _od, n_kept = self.nerf.render_tangent_raw(ro[0], rd[0], rod[0], rdd[0], u)
            if self.has_bayer_filter:
                ch = event["channel_idx"].to(torch.int64).unsqueeze(-1)                 # :887-890
                cg, cg_dot = rad.gather(1, ch).squeeze(-1), rad_dot.gather(1, ch).squeeze(-1)
            else:
                cg, cg_dot = rad[:, 0], rad_dot[:, 0]
            # d logI / d ts = (dC/dw) / (C + I_min) [f32] * dw/dt [f64]   (trajectories.py:63-65 cast)
            log_i_grad = (cg_dot / (cg + imin)).double() * rays["dw_dt"][0]
            rate = ev_diff / (event["end_ts"] - start)                                  # loss.py:38-41
            err = self._error(self._loss_cfg("error_fn", "log_intensity_grad"), log_i_grad, rate)
            valid_g = torch.ones_like(opac_g, dtype=torch.bool) if all_valid else opac_g > 0
            # masked mean without boolean indexing (no host sync); 0 / 0 = nan like `err[valid].mean()`
            losses["log_intensity_grad"] = torch.where(valid_g, err, torch.zeros_like(err)).sum() / valid_g.sum()
            n_kept_all.append(n_kept)
            opac_all.append(opac_g)
            self.last_grad = dict(log_intensity_grad=log_i_grad, valid=valid_g, target=rate,
                                  log_intensity=torch.log(cg + imin))
        # f: both difference renders in one launch (start rays | end rays)
        log_i0 = log_i1 = n_valid = opac = None
        if w_diff > 0:
            if jitter is None:
                u = torch.rand((2 * B,), device=dev, dtype=torch.float32)
            else:
                u = torch.cat((jitter["diff_start"], jitter["diff_end"])).to(dev, torch.float32)
            rad, opac, _depth, n_kept = self.nerf.render_raw(torch.cat((ro[nr - 2], ro[nr - 1])),
                                                             torch.cat((rd[nr - 2], rd[nr - 1])), u)
            n_kept_all += [n_kept[:B], n_kept[B:]]
            opac_all += [opac[:B], opac[B:]]
            if self.has_bayer_filter:
                ch = event["channel_idx"].to(torch.int64).repeat(2).unsqueeze(-1)         # :887-890
                c = rad.gather(1, ch).squeeze(-1)
            else:
                c = rad[:, 0]
            c0, c1 = c[:B], c[B:]
            o0, o1 = opac[:B], opac[B:]
            # h-j: loss
            loss_raw, log_i0, log_i1, n_valid = DiffLossFunction.apply(
                c0, c1, o0, o1, ev_diff, rays["ts_diff"], rays["start"], event["end_ts"].contiguous(),
                self._loss_cfg("error_fn", "log_intensity_diff"), all_valid, imin)
            losses["log_intensity_diff"] = loss_raw
        mcr = 1 / self.contrast_threshold.mean_contrast_threshold
        pw = {None: 1.0, "mean_contrast_reciprocal": mcr, "mean_contrast_reciprocal_sq": mcr ** 2}
        train_loss = 0
        for name in list(losses):
            losses[name] = pw[self._loss_cfg("param_weight", name)] * losses[name]
            train_loss = train_loss + losses[name] * self._loss_cfg("weight", name)
        # i: dynamic batch size (host sync, as in the reference; only with a trainer attached)
        mean_spp = sum(nk.sum() / B for nk in n_kept_all) / len(n_kept_all)
        if self.trainer is not None:
            mean_spp = self.update_train_batch_size(mean_spp, batch_index)
        # k: logging
        self.log("train/loss", train_loss, prog_bar=True)
        for name, val in losses.items():
            self.log(f"train/{name}", val)
        if not _get(self.hparams["contrast_threshold"], "freeze"):
            self.log("train/p2n_contrast_threshold_ratio", self.contrast_threshold.p2n_contrast_threshold_ratio)
        if not _get(self.hparams["refractory_period"], "freeze"):
            self.log("train/refractory_period", self.refractory_period.refractory_period)
        self.log("train/batch_size", B)
        self.log("train/mean_num_samples_per_ray", mean_spp)
        occ_rate = sum((op > 0).float().mean() for op in opac_all) / len(opac_all)
        self.log("train/mean_ray_occ_rate", occ_rate)
        if self.render_bkgd is None:
            self.log("train/mean_valid_rate", occ_rate)
        else:
            self.log("train/mean_valid_rate", torch.ones((), device=dev))
        self.last = dict(start_log_intensity=log_i0, end_log_intensity=log_i1, n_valid=n_valid,
                         rays=rays, opacity=opac)
        return train_loss

    @staticmethod
    def _error(kind, inp, tgt):
        """loss_metric/loss.py:16-30, utils/modules.py:77-102 (MAPE eps = float64 machine eps)."""
        if kind == "l1":
            return (inp - tgt).abs()
        if kind == "mse":
            return (inp - tgt) ** 2
        if kind == "mape":
            return (inp - tgt).abs() / tgt.abs().clamp(min=2.220446049250313e-16)
        raise NotImplementedError(kind)

    def update_train_batch_size(self, mean_num_samples_per_ray, batch_index):
        """robust_e_nerf.py:907-950."""
        m = torch.mean(self.all_gather(mean_num_samples_per_ray.detach().float()))
        acc = getattr(self.trainer, "accumulate_grad_batches", 1)
        if acc > 1 and (batch_index % acc) != (acc - 2):
            return m
        dm = getattr(self.trainer, "datamodule", None)
        if dm is not None:
            new_bs = int(self.train_ray_sample_batch_size / m)
            dm.train_dataset.batch_size = new_bs
            for s in dm.train_normalized_sampler.datasets:
                s.size = new_bs
        return m

    # -------------------------------------------------------------- optimiser
    def configure_optimizers(self):
        """robust_e_nerf.py:782-832: Adam with 4 parameter groups + MultiStepLR."""
        ct = list(self.contrast_threshold.parameters())
        rp = list(self.refractory_period.parameters())
        mlp = [p for n, p in self.named_parameters() if n.startswith("nerf.radiance_field.mlp")]
        ids = {id(p) for p in ct + rp + mlp}
        other = [p for p in self.parameters() if id(p) not in ids]
        opt_cfg = self.hparams["optimizer"]
        if _get(opt_cfg, "algo") != "adam":
            raise NotImplementedError
        lr = _get(opt_cfg, "lr")
        optimizer = torch.optim.Adam(
            [{"params": ct, "lr": _get(lr, "contrast_threshold")},
             {"params": rp, "lr": float(self.refractory_period.max_refractory_period)
              * _get(_get(opt_cfg, "relative_lr"), "refractory_period")},
             {"params": mlp, "weight_decay": self._loss_cfg("weight", "nerf_mlp_weight_decay")},
             {"params": other}], lr=_get(lr, "default"))
        sch_cfg = self.hparams["lr_scheduler"]
        if _get(sch_cfg, "algo") != "multi_step_lr":
            raise NotImplementedError
        ms = _get(sch_cfg, "multi_step_lr")
        scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=_get(ms, "milestones"),
                                                         gamma=_get(ms, "gamma"))
        return {"optimizer": optimizer,
                "lr_scheduler": {"scheduler": scheduler, "interval": _get(sch_cfg, "interval")}}

    # -------------------------------------------------------------- evaluation render
    def set_eval_stage(self, name, intrinsics, img_height, img_width, min_normalized_pixel_value=0.0,
                       max_normalized_pixel_value=1.0):
        """Registers `{name}_intrinsics_inv` / `{name}_img_pixel_pos` and the target value range exactly as
        the reference's constructor does from its PosedImage datasets (robust_e_nerf.py:97-147); `name` is
        "val" or "test".  The datasets themselves (views/transforms_*.json, PNG targets) are
        SURVEY.md section 8f #3."""
        assert name in ("val", "test")
        setattr(self, f"{name}_min_normalized_pixel_value", float(min_normalized_pixel_value))
        setattr(self, f"{name}_max_normalized_pixel_value", float(max_normalized_pixel_value))
        dev = self.train_intrinsics_inv.device
        k = torch.as_tensor(intrinsics, dtype=torch.float32)
        self.register_buffer(f"{name}_intrinsics_inv", torch.linalg.inv(k).to(dev), persistent=False)
        pix = torch.stack(torch.meshgrid(torch.arange(img_width), torch.arange(img_height), indexing="xy"),
                          dim=2).to(torch.float32)
        self.register_buffer(f"{name}_img_pixel_pos", pix.to(dev), persistent=False)

    @torch.no_grad()
    def evaluation_step(self, batch, batch_index, stage):
        """robust_e_nerf.py:533-571: full-frame intensity image at one camera pose.  In eval mode
        `NeRF.forward` renders in `test_chunk_size` ray chunks without jitter (external/utils.py:99-115)."""
        batch = {k: (v.squeeze(dim=0) if torch.is_tensor(v) else v) for k, v in batch.items()}
        pix = stage["img_pixel_pos"]
        h, w = pix.shape[:2]
        if "img" in batch:
            assert batch["img"].shape[-2:] == (h, w)
        pos = batch["T_wc_position"].view(1, 1, 3).expand(h, w, -1)
        rot = batch["T_wc_orientation"].view(1, 1, 3, 3).expand(h, w, -1, -1)
        pred, _, _, _, _ = self.render_pixels(stage["intrinsics_inv"], pix, pos, rot)
        return {"sample_id": batch.get("sample_id"), "pred_intensity_img": pred,
                "target_intensity_img": batch.get("img")}

    def _eval_stage(self, name):
        if not hasattr(self, f"{name}_intrinsics_inv"):
            raise RuntimeError(f"call set_eval_stage({name!r}, intrinsics, height, width) first: the posed-image "
                               "datasets that provide them are SURVEY.md section 8f #3")
        return {"name": name, "intrinsics_inv": getattr(self, f"{name}_intrinsics_inv"),
                "img_pixel_pos": getattr(self, f"{name}_img_pixel_pos"),
                "min_normalized_pixel_value": getattr(self, f"{name}_min_normalized_pixel_value"),
                "max_normalized_pixel_value": getattr(self, f"{name}_max_normalized_pixel_value")}

    @staticmethod
    @torch.no_grad()
    def align_log_intensity(pred_intensity_img, target_intensity_img):
        """robust_e_nerf.py:636-683: the event-only reconstruction is affine-ambiguous in log intensity, so the
        prediction is mapped through the per-channel least-squares `scale * log(pred) + offset ~ log(target)`
        (over all images of the epoch, float64) before any metric.  Images are (N, C, H, W).  The reference
        materialises the (N*H*W, 2) design matrix on the CPU for `torch.linalg.lstsq`; here the same minimiser
        comes from the five centred sums, on whatever device the images live on (a rank-deficient system -- a
        constant prediction -- gets lstsq's minimum-norm solution).  Returns (aligned intensity, scale (C,),
        offset (C,))."""
        assert pred_intensity_img.shape == target_intensity_img.shape and pred_intensity_img.dim() == 4
        p = pred_intensity_img.log().to(torch.float64)
        t = target_intensity_img.log().to(torch.float64)
        dims = (0, 2, 3)
        pm, tm = p.mean(dims, keepdim=True), t.mean(dims, keepdim=True)
        spp = ((p - pm) ** 2).sum(dims, keepdim=True)
        spt = ((p - pm) * (t - tm)).sum(dims, keepdim=True)
        degenerate = spp <= 0
        scale = torch.where(degenerate, pm * tm / (pm * pm + 1), spt / spp.clamp(min=1e-300))
        offset = torch.where(degenerate, tm / (pm * pm + 1), tm - scale * pm)
        aligned = (scale * p + offset).to(target_intensity_img.dtype).exp()
        return aligned, scale.reshape(-1), offset.reshape(-1)

    @staticmethod
    @torch.no_grad()
    def ssim(pred_img, target_img, data_range, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
        """Structural similarity as `torchmetrics.functional.ssim` of torchmetrics==0.6.2 computes it
        (`loss_metric/metric.py:75-82`: defaults, `data_range = max_target_val`, mean over the map): 11 x 11
        Gaussian window (sigma 1.5) applied per channel to reflect-padded images, the map cropped back by the pad
        width, then averaged.  torchmetrics is absent here: restated from its published implementation, parity
        unpinned; `tests/test_module_cpu.py::test_ssim_properties` checks it against an independent
        scipy.ndimage evaluation of the same definition.  Images are (N, C, H, W)."""
        assert pred_img.shape == target_img.shape and pred_img.dim() == 4
        C_ = pred_img.shape[1]
        c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
        ax = torch.arange(kernel_size, dtype=pred_img.dtype, device=pred_img.device) - (kernel_size - 1) / 2
        g = torch.exp(-(ax / sigma) ** 2 / 2)
        g = g / g.sum()
        kernel = (g[:, None] * g[None, :]).expand(C_, 1, kernel_size, kernel_size).contiguous()
        pad = (kernel_size - 1) // 2
        p = torch.nn.functional.pad(pred_img, (pad, pad, pad, pad), mode="reflect")
        t = torch.nn.functional.pad(target_img, (pad, pad, pad, pad), mode="reflect")
        n = p.shape[0]
        out = torch.nn.functional.conv2d(torch.cat((p, t, p * p, t * t, p * t)), kernel, groups=C_)
        mu_p, mu_t, e_pp, e_tt, e_pt = (out[i * n:(i + 1) * n] for i in range(5))
        var_p, var_t, cov = e_pp - mu_p * mu_p, e_tt - mu_t * mu_t, e_pt - mu_p * mu_t
        ssim_map = ((2 * mu_p * mu_t + c1) * (2 * cov + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (var_p + var_t + c2))
        # the window is centred on every pixel of the ORIGINAL image after the valid convolution of the padded one;
        # torchmetrics 0.6 additionally drops a pad-wide border of that map
        if ssim_map.shape[-1] > 2 * pad and ssim_map.shape[-2] > 2 * pad:
            ssim_map = ssim_map[..., pad:-pad, pad:-pad]
        return ssim_map.mean()

    @torch.no_grad()
    def evaluation_epoch_end(self, outputs, stage):
        """robust_e_nerf.py:589-696: gather the per-view outputs of `evaluation_step` across ranks, and on the
        global-zero rank align the predicted log intensities to the targets, then mean L1, PSNR
        (`loss_metric/metric.py:60-73`: per image with data_range = max - min of the target range) and SSIM,
        and dump the aligned predictions as PNGs.  LPIPS (pretrained AlexNet weights) is absent from this image.
        Returns the metrics and the aligned images (None on the other ranks, like the reference's bare return)."""
        pred = torch.stack([o["pred_intensity_img"] for o in outputs])
        target = torch.stack([o["target_intensity_img"] for o in outputs]).to(pred.device)
        ids = None
        if outputs[0].get("sample_id") is not None:
            ids = torch.stack([o["sample_id"] for o in outputs]).reshape(len(outputs), -1).to(pred.device)
        # :591-606 gather & merge across devices (a sharded validation set under DDP)
        distributed = torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size() > 1
        if distributed:
            pred = self.all_gather(pred).flatten(0, 1)
            target = self.all_gather(target).flatten(0, 1)
            if ids is not None:
                ids = self.all_gather(ids).flatten(0, 1)
        log_dir = getattr(self.trainer, "log_dir", None)              # :613-616 (read on every rank)
        if not getattr(self.trainer, "is_global_zero", True):         # :618-621
            return None
        if not self.has_bayer_filter:                                   # (N, H, W) -> (N, 1, H, W)
            pred, target = pred.unsqueeze(1), target.unsqueeze(1)
        aligned, scale, offset = self.align_log_intensity(pred, target)
        rng = stage["max_normalized_pixel_value"] - stage["min_normalized_pixel_value"]
        l1 = torch.stack([(a - b).abs().mean() for a, b in zip(aligned, target)]).mean()
        mse = ((aligned - target) ** 2).mean(dim=(1, 2, 3))
        psnr = (10.0 * torch.log10(rng ** 2 / mse)).mean()
        h, w = aligned.shape[-2:]
        ssim = None
        if min(h, w) > 5:                                            # reflect padding needs pad < size
            ssim = torch.stack([self.ssim(a[None], b[None], stage["max_normalized_pixel_value"])
                                for a, b in zip(aligned, target)]).mean()
        name = stage.get("name", "eval")
        self.log(f"{name}/l1", l1, prog_bar=True, rank_zero_only=True)
        self.log(f"{name}/psnr", psnr, prog_bar=True, rank_zero_only=True)
        if ssim is not None:
            self.log(f"{name}/ssim", ssim, prog_bar=True, rank_zero_only=True)
        self.log(f"{name}/epoch", float(self.current_epoch), rank_zero_only=True)          # :690-696
        if self.eval_save_pred_intensity_img and log_dir is not None and ids is not None:
            self.save_pred_intensity_imgs(aligned, ids, stage, os.path.join(log_dir, self.PREDICTIONS_FOLDER_NAME))
        return {"l1": l1, "psnr": psnr, "ssim": ssim, "log_intensity_scale": scale, "log_intensity_offset": offset,
                "pred_intensity_img": aligned}

    PREDICTIONS_FOLDER_NAME = "predictions"                          # robust_e_nerf.py:20-22
    PREDICTION_FILE_EXTENSION = ".png"
    PREDICTION_BIT_DEPTH = 8

    def save_pred_intensity_imgs(self, pred_intensity_img, sample_id, stage, folder):
        """robust_e_nerf.py:733-779: the aligned predictions (N, C, H, W), mapped from the target value range to
        [0, 2^bits - 1], clipped, rounded and written as `<sample id>.png` (gray, or BGR for a colour sensor)."""
        import cv2
        top = 2 ** self.PREDICTION_BIT_DEPTH - 1
        lo, hi = stage["min_normalized_pixel_value"], stage["max_normalized_pixel_value"]
        img = (top * ((pred_intensity_img.detach().cpu() - lo) / (hi - lo)).clamp(min=0, max=1)).round()
        img = img.numpy().astype(np.uint8 if self.PREDICTION_BIT_DEPTH == 8 else np.uint16).transpose(0, 2, 3, 1)
        os.makedirs(folder, exist_ok=True)
        names = ["".join(map(chr, row)).rstrip() for row in sample_id.tolist()]
        for name, im in zip(names, img):
            if self.has_bayer_filter:
                im = cv2.cvtColor(im, cv2.COLOR_RGB2BGR)
            cv2.imwrite(os.path.join(folder, name + self.PREDICTION_FILE_EXTENSION), im)
        return names

    def validation_epoch_end(self, outputs):
        """robust_e_nerf.py:573-579."""
        return self.evaluation_epoch_end(outputs, self._eval_stage("val"))

    def test_epoch_end(self, outputs):
        """robust_e_nerf.py:581-587."""
        return self.evaluation_epoch_end(outputs, self._eval_stage("test"))

    def validation_step(self, batch, batch_index):
        """robust_e_nerf.py:519-524."""
        return self.evaluation_step(batch, batch_index, self._eval_stage("val"))

    def test_step(self, batch, batch_index):
        """robust_e_nerf.py:526-531."""
        return self.evaluation_step(batch, batch_index, self._eval_stage("test"))
