#!/usr/bin/env python
"""bench.py -- ray-samples/s of the event-NeRF training hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one `RobustENeRF.training_step` + backward + gradient all-reduce (N>1) + Adam on one
synthetic event batch.  Workloads (BASELINE.json `configs`, shapes per SURVEY.md section 8d):

  2  (headline) 4096 rays/GPU x 128 samples x 2 renders, 8Lx256 MLP, l_diff, frozen C_p / tau
  3  16384 rays x 128 samples x 3 renders (one in tangent mode), l_diff + l_grad, C_p and tau trainable
  4  mocap-desk2-like: 8192 rays/GPU x 192 steps x 3 renders, aabb / near / far of the YAML, occupancy grid on
  5  office-maze-like: 8192 rays/GPU x 256 steps x 3 renders, occupancy grid on

One ray-sample = one (ray, sample) pair pushed through field + compositing in one render; with an occupancy
grid the emitted samples (what the reference's dynamic batch size counts) are reported next to the dense steps.

  value     whole-job ray-samples/s, event batches already resident in HBM; the timed block of K steps is
            repeated until >= 2 s of GPU time: median block (min / max in `blocks`)
  e2e       same through the public API with HOST (pinned) batches: H2D copy of every batch and a D2H read of
            the loss inside the timed region
  roofline  dominant kernel (field forward) against the measured bf16 tensor peak: 1 186 304 algorithmic FLOP
            per ray-sample / CUDA-event time
  parity    the second half of BASELINE.json's metric: after the timed region, a fixed 128-ray subset of
            batch 0 rendered by the CUDA path with the TRAINED weights vs the oracle on the same weights, rays
            and jitter: per-ray log I (normwise, max-abs) and loss relative error
  secondary (N=1, --config 2 only) short runs of configs 3, 4, 5 (and of config 4's shape with `arch: ngp`) in the
            same JSON line
  cpu_baseline  the oracle port of the reference's own Python path on the host cores, bounded sample

`--impl reference` times that CPU path alone (the upstream reference has no CPU mode and cannot be installed
here: nerfacc / tiny-cuda-nn are CUDA-only and absent) -- a PORT, not the reference.
"""
import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "robust-e-nerf_b200"))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_FWD = 1_186_304          # 593 152 MAC, 8x256 MLP, C = 1
MIN_TIMED_MS = 2000.0
BENCH_LR = 5e-4

# rays per GPU, marching steps per ray, renders, l_grad weight, trainable sensor parameters, scene box
CONFIGS = {
    2: dict(name="config2: ficus-easy-like synthetic 640x480 event stream, 4096 rays/GPU x 128 samples x 2 renders, "
                 "8Lx256 MLP, l_diff (mse), frozen Cp/tau, fwd+bwd+Adam",
            rays=4096, n=128, renders=2, w_grad=0.0, trainable=False, aabb=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5),
            near=0.5, far=None, span=4.0, grid=False),
    3: dict(name="config3: ficus-hard-like, 16384 rays x 128 samples x 3 renders (2 value + 1 tangent), 8Lx256 MLP, "
                 "l_diff + 1e-3 l_grad (mape), Cp and tau trainable from a poor init, fwd+bwd+Adam",
            rays=16384, n=128, renders=3, w_grad=1e-3, trainable=True, aabb=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5),
            near=0.5, far=None, span=4.0, grid=False),
    4: dict(name="config4: mocap-desk2-like, 8192 rays/GPU x 192 marching steps x 3 renders, aabb / near / far of "
                 "configs/train/mocap-desk2.yaml:38-48, sphere contraction, occupancy grid 128^3 on, l_diff + l_grad, "
                 "Cp and tau trainable, fwd+bwd+Adam",
            rays=8192, n=192, renders=3, w_grad=1e-3, trainable=True, aabb=(0.5, -2.1, 0.6, 2.0, -0.6, 1.6),
            near=0.05, far=3.0, span=None, grid=True),
    5: dict(name="config5: office-maze-like, 8192 rays/GPU x 256 marching steps x 3 renders, aabb / near / far of "
                 "configs/train/office-maze.yaml:38-48, sphere contraction, occupancy grid 128^3 on, l_diff + l_grad, "
                 "Cp and tau trainable, fwd+bwd+Adam",
            rays=8192, n=256, renders=3, w_grad=1e-3, trainable=True, aabb=(-2.3, -3.2, 0.0, 1.7, 2.8, 3.0),
            near=0.01, far=4.0, span=None, grid=True),
    # not a BASELINE config (its metric names the 8Lx256 MLP): the architecture the reference's YAMLs select
    6: dict(name="config4 shape with `arch: ngp` (configs/train/mocap-desk2.yaml:56-88: 16-level 2^19 hash grid, SH-4, "
                 "64-wide softplus MLPs) on the fp32 engine: 8192 rays/GPU x 192 marching steps x 3 renders, occupancy "
                 "grid 128^3 on, l_diff + l_grad, Cp and tau trainable, fwd+bwd+Adam",
            rays=8192, n=192, renders=3, w_grad=1e-3, trainable=True, aabb=(0.5, -2.1, 0.6, 2.0, -0.6, 1.6),
            near=0.05, far=3.0, span=None, grid=True, arch="ngp"),
}
NGP_BLOCK = dict(
    pos_encoding=dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                      base_resolution=16, per_level_scale=1.4472692012786865, interpolation="Linear"),
    dir_encoding=dict(degree=4),
    mlp_base=dict(hidden_activation="softplus", density_activation="shifted_trunc_exp", n_neurons=64,
                  n_hidden_layers=1, geo_feat_dim=15, weight_norm=False),
    mlp_head=dict(hidden_activation="softplus", radiance_activation="softplus", n_neurons=64, n_hidden_layers=2,
                  weight_norm=False))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), hbm=p["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json, bf16 sustained)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_metrics():
    """Tensor-pipe activity and DRAM bytes of the dominant kernel come from a COMMITTED ncu capture of this
    command (profiles/r02_ncu_metrics.json names the report it was read from); absent file -> null fields."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            d["file"] = "profiles/r02_ncu_metrics.json"
            return d
        except Exception:
            pass
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.rows = []

    def _run_nvml(self):
        """NVML directly (nvidia_ml_py): sub-millisecond samples instead of one nvidia-smi process per sample."""
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.gpu)
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            N.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        bits = ((0x8, 4), (0x40, 5), (0x20, 6), (0x4, 7))       # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        while not self.stop_flag.is_set():
            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            row = [str(self.gpu), str(sm), str(mx), "0", "", "", "", ""]
            for bit, col in bits:
                row[col] = "Active" if r & bit else "Not Active"
            self.rows.append(row)
            self.stop_flag.wait(0.01)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# --------------------------------------------------------------------------- workload construction
def scene_geometry(wc):
    """Orbit centre / scale of the synthetic trajectory: the unit scene for configs 2-3, re-centred on the aabb
    centre and scaled to its extent for the real-sequence boxes."""
    a = wc["aabb"]
    if tuple(a) == (-1.5, -1.5, -1.5, 1.5, 1.5, 1.5):
        return (0.0, 0.0, 0.0), 1.0
    center = tuple((a[i] + a[i + 3]) / 2 for i in range(3))
    ext = max(a[i + 3] - a[i] for i in range(3))
    return center, 0.3 * ext / 3.0            # orbit radius = 0.3 x the largest extent: the camera is inside


def product_config(wc, world):
    from ren_b200 import synthetic
    if wc["span"] is not None:
        cfg = synthetic.default_config(n_samples=wc["n"], aabb=wc["aabb"], near=wc["near"], span=wc["span"])
    else:
        step = (wc["far"] - wc["near"]) / (wc["n"] + 0.5)
        cfg = synthetic.default_config(n_samples=wc["n"], aabb=wc["aabb"], near=wc["near"], span=step * wc["n"])
    cfg["num_nodes"], cfg["gpus"] = 1, list(range(world))
    # the synthetic events carry random polarities (there is nothing to converge to): at the YAML's lr 0.01 Adam
    # drives the trunc-exp density to inf within a few hundred steps -- in the oracle as well.  A NeRF-MLP-sized
    # lr keeps the benchmarked network finite for the whole run; the step's arithmetic does not depend on it.
    cfg["optimizer"]["lr"]["default"] = BENCH_LR
    cfg["loss"]["weight"]["log_intensity_grad"] = wc["w_grad"]
    if wc["trainable"]:
        cfg["contrast_threshold"]["freeze"] = False
        cfg["refractory_period"]["freeze"] = False
    if wc.get("arch") == "ngp":
        cfg["nerf"]["arch"] = "ngp"
        cfg["nerf"]["ngp"] = NGP_BLOCK
    if wc["grid"]:
        cfg["nerf"]["occ_grid"] = dict(resolution=128, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16,
                                       enabled=True)
    return cfg


def build_model(wc, world, rank, dev, engine):
    import torch
    from ren_b200 import RobustENeRF, synthetic
    root = os.path.join(tempfile.gettempdir(), f"ren_b200_bench_{rank}_{wc['n']}_{int(wc['grid'])}")
    center, scale = scene_geometry(wc)
    synthetic.write_dataset(root, center=center, scale=scale)
    cfg = product_config(wc, world)
    torch.manual_seed(0)
    model = RobustENeRF(dataset_directory=root, engine=engine, **cfg).to(dev).train()
    if wc["trainable"]:
        with torch.no_grad():       # poor initialisation: ratio x1.5, tau x0.5
            ct = model.contrast_threshold.parametrizations.p2n_contrast_threshold_ratio
            ct.original.copy_(ct[0].right_inverse(torch.tensor(1.5)))
            rp = model.refractory_period.parametrizations._refractory_period
            rp.original.copy_(rp[0].right_inverse(model.refractory_period.init_refractory_period.double() * 0.5))
    return model, cfg, root


def oracle_configs(cfg):
    """oracle.restatement configs equivalent to a `synthetic.default_config` dict."""
    from oracle import restatement as R
    n, mlp, loss = cfg["nerf"], cfg["nerf"]["mlp"], cfg["loss"]
    fcfg = R.FieldConfig(aabb=list(n["aabb"]), contraction=n["contraction_type"], net_depth=mlp["net_depth"],
                         net_width=mlp["net_width"], skip_layer=mlp["skip_layer"],
                         net_depth_condition=mlp["net_depth_condition"],
                         net_width_condition=mlp["net_width_condition"])
    rcfg = R.RenderConfig(near_plane=n["near_plane"], far_plane=n["far_plane"],
                          render_step_size=n["render_step_size"], cone_angle=n["cone_angle"],
                          early_stop_eps=n["early_stop_eps"], alpha_thre=n["alpha_thre"])
    scfg = R.StepConfig(w_grad=loss["weight"]["log_intensity_grad"], w_diff=loss["weight"]["log_intensity_diff"],
                        err_grad=loss["error_fn"]["log_intensity_grad"], err_diff=loss["error_fn"]["log_intensity_diff"],
                        pw_grad=loss["param_weight"]["log_intensity_grad"],
                        pw_diff=loss["param_weight"]["log_intensity_diff"],
                        min_modeled_intensity=cfg["min_modeled_intensity"])
    return fcfg, rcfg, scfg


def parity_check(model, cfg, root, host_batch, host_jitter, dev, n_rays=128):
    """CUDA path vs the oracle on the first `n_rays` rays of batch 0 with the model's CURRENT weights."""
    import torch
    from oracle import restatement as R
    from oracle import synth as osynth
    torch.set_num_threads(os.cpu_count() or 1)
    sub = {k: {kk: vv[:, :n_rays].clone() for kk, vv in v.items()} for k, v in host_batch.items()}
    jit = {k: v[:n_rays].clone() for k, v in host_jitter.items()}
    grid = getattr(model.nerf, "occupancy_grid", None)
    step0 = model.global_step
    try:
        model.global_step = None            # no occupancy-grid update inside the parity render
    except AttributeError:
        pass
    with torch.no_grad():
        loss = model.training_step({k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in sub.items()}, 0,
                                   {k: v.to(dev) for k, v in jit.items()})
    try:
        model.global_step = step0
    except AttributeError:
        pass
    got0, got1 = model.last["start_log_intensity"].double().cpu(), model.last["end_log_intensity"].double().cpu()
    sd = model.state_dict()
    params = {k[len("nerf.radiance_field."):]: v.detach().float().cpu() for k, v in sd.items()
              if k.startswith("nerf.radiance_field.mlp.")}
    scene = osynth.load_scene(root)
    fcfg, rcfg, scfg = oracle_configs(cfg)
    th_r = sd["contrast_threshold.parametrizations.p2n_contrast_threshold_ratio.original"].detach().cpu()
    th_t = sd["refractory_period.parametrizations._refractory_period.original"].detach().cpu().clone()
    if grid is not None or cfg["nerf"]["arch"] != "mlp":
        return {"rays": n_rays, "note": "occupancy grid on / arch ngp: the oracle comparisons of these paths live in "
                                        "tests/test_occupancy_gpu.py and tests/test_ngp_gpu.py", "loss": float(loss)}
    if scfg.w_grad > 0:
        out = R.training_step(params, fcfg, rcfg, scfg, scene, sub, jit, th_r, th_t)
    else:
        with torch.no_grad():
            out = R.training_step(params, fcfg, rcfg, scfg, scene, sub, jit, th_r, th_t)
    ref0, ref1 = out["start_log_intensity"].detach().double(), out["end_log_intensity"].detach().double()
    got, ref = torch.cat((got0, got1)), torch.cat((ref0, ref1))
    ref_loss = float(out["loss"].detach())
    res = {"rays": n_rays, "weights": "after the timed training steps",
           "logI_normwise": float((got - ref).norm() / ref.norm()), "logI_maxabs": float((got - ref).abs().max()),
           "logI_std_over_rays": float(ref.std()),
           "loss_rel_err": abs(float(loss) - ref_loss) / abs(ref_loss), "loss": float(loss), "oracle_loss": ref_loss,
           "bar": 1e-4}
    if scfg.w_grad > 0:
        g, r = model.last_grad["log_intensity_grad"].double().cpu(), out["log_intensity_grad"].detach().double()
        res["dlogI_dt_normwise"] = float((g - r).norm() / r.norm())
    return res


def cpu_oracle_rate(wc, n_rays, steps, warmup, dataset_dir):
    """Oracle port of the reference training step (fwd + bwd) on the host cores."""
    import torch
    from oracle import restatement as R
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    scene = synth.load_scene(dataset_dir)
    fcfg = R.FieldConfig()
    step = 4.0 / wc["n"]
    rcfg = R.RenderConfig(near_plane=0.5, far_plane=0.5 + (wc["n"] + 0.5) * step, render_step_size=step)
    scfg = R.StepConfig()
    params = {k: v.requires_grad_(True) for k, v in R.init_field_params(fcfg, seed=0).items()}
    batch, jitter = synth.make_batch(n_rays, scene["T_wc_timestamp"])
    th_r = torch.log(torch.expm1(torch.tensor(1.0)))
    tau_max = scene["max_refractory_period"]
    th_t = tau_max * torch.logit(scene["refractory_period"].double() / tau_max)
    times = []
    loss = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in params.values():
            p.grad = None
        out = R.training_step(params, fcfg, rcfg, scfg, scene, batch, jitter, th_r, th_t.clone())
        out["loss"].backward()
        dt = time.perf_counter() - t0
        loss = float(out["loss"].detach())
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    return wc["renders"] * n_rays * wc["n"] / per_step, per_step, torch.get_num_threads(), loss


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    wc = CONFIGS[2]
    root = synth.write_dataset(os.path.join(tempfile.gettempdir(), "ren_b200_bench_ref"))
    n_rays = 256
    rate, per_step, cores, loss = cpu_oracle_rate(wc, n_rays, args.steps, args.warmup, root)
    sample = (f"PORT of the reference path (oracle/restatement.py, torch CPU fp32): {n_rays} of the 4096 rays x "
              f"{wc['n']} samples x {wc['renders']} renders per step (fwd+bwd)")
    print(json.dumps({
        "impl": "reference", "metric": "ray-samples/s", "value": rate, "unit": "ray-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wc["name"], "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "ray-samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": loss}))


# --------------------------------------------------------------------------- product arm
class Runner:
    """One workload on this rank: model, trainer, resident + pinned host batches, timed blocks."""

    def __init__(self, wc, args, rank, world, dev, graph):
        import torch
        from ren_b200 import synthetic
        from ren_b200.trainer import FlatTrainer
        self.wc, self.world, self.dev, self.rank = wc, world, dev, rank
        self.model, self.cfg, self.root = build_model(wc, world, rank, dev, args.engine)
        self.graph = bool(graph) and not wc["grid"]
        self.trainer = FlatTrainer(self.model, world_size=world, cuda_graph=self.graph)
        pose_ts = self.model.trajectory.T_wc_timestamp.cpu()
        self.n_batches = 4
        self.host = [synthetic.make_batch(wc["rays"], pose_ts, rank=rank, seed=1234 + 17 * i, pin=True)
                     for i in range(self.n_batches)]
        self.resident = [(synthetic.to_device(b, dev), {k: v.to(dev) for k, v in j.items()}) for b, j in self.host]
        self.h2d = synthetic.batch_nbytes(self.host[0][0]) + sum(v.numel() * 4 for v in self.host[0][1].values())
        self.it = 0
        self.spp_sum = None
        self.torch = torch

    def sync_all(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            self.torch.cuda.synchronize()

    def block(self, n_steps, e2e):
        """K steps between two CUDA events (barrier + synchronize on both sides), max over ranks."""
        import torch.distributed as dist
        from ren_b200 import synthetic
        torch = self.torch
        self.sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = None
        for _ in range(n_steps):
            i = self.it
            self.it += 1
            if e2e:
                hb, hj = self.host[i % self.n_batches]
                batch = synthetic.to_device(hb, self.dev)
                jit = {k: v.to(self.dev, non_blocking=True) for k, v in hj.items()}
                last = float(self.trainer.step(batch, i, jit))            # D2H read of the loss
            else:
                batch, jit = self.resident[i % self.n_batches]
                last = self.trainer.step(batch, i, jit)
        b.record()
        self.sync_all()
        ms = a.elapsed_time(b)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, float(last)

    def timed(self, n_steps, e2e, min_ms):
        """Blocks of exactly `n_steps` steps, repeated until `min_ms` of GPU time: (median, [all], loss)."""
        blocks, total, loss = [], 0.0, None
        while True:
            ms, loss = self.block(n_steps, e2e)
            blocks.append(ms)
            total += ms
            if total >= min_ms or len(blocks) >= 200:
                break
        return statistics.median(blocks), blocks, loss

    def mean_spp(self):
        v = self.model.logged.get("train/mean_num_samples_per_ray")
        return float(v) if v is not None else float(self.wc["n"])


def forward_only_rate(run, n_rays=8192, iters=10):
    """the forward-only (inference) rate beside the training rate -- a no-grad render of
    `n_rays` rays through the public `NeRF.forward` with the weights the timed steps left behind (no stash is written),
    CUDA events around `iters` calls after 3 warm-up calls.  Rays are the training rays of the resident batches' kind:
    origins on the camera orbit, directions through the scene."""
    torch = run.torch
    dev = run.dev
    g = torch.Generator(device=dev).manual_seed(4321)
    o = torch.randn((n_rays, 3), device=dev, generator=g)
    o = 3.0 * o / o.norm(dim=-1, keepdim=True)
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn((n_rays, 3), device=dev, generator=g), dim=-1)
    u = torch.rand((n_rays,), device=dev, generator=g)
    nerf = run.model.nerf
    with torch.no_grad():
        for _ in range(3):
            out = nerf(o, d, u)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            out = nerf(o, d, u)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    spp = float(out[3])
    return {"rays": n_rays, "mean_samples_per_ray": spp, "ms_per_render": ms,
            "ray_samples_per_s": n_rays * spp / (ms * 1e-3),
            "tflops_algorithmic": n_rays * spp * FLOP_PER_SAMPLE_FWD / (ms * 1e-3) / 1e12,
            "note": "no-grad NeRF.forward (march, encode, MLP, composite), no activation stash"}


def release(run):
    """Drop a Runner's CUDA graphs (their private memory pools) and cached blocks before the next workload."""
    import gc
    import torch
    run.trainer._graphs.clear()
    run.trainer = run.model = run.resident = None
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def run_secondary(cfg_id, args, rank, world, dev, graph):
    """A short run of another BASELINE config (warm-up 3, then one block of 5 steps) for the same JSON line."""
    wc = CONFIGS[cfg_id]
    try:
        r = Runner(wc, args, rank, world, dev, graph)
        r.block(3, False)
        if r.graph:
            r.block(2, False)                       # first replay after the capture
        ms, blocks, loss = r.timed(5, False, 0.0)
        spp = r.mean_spp()
        dense = wc["renders"] * wc["rays"] * wc["n"] * world
        emitted = wc["renders"] * wc["rays"] * spp * world
        out = {"config": cfg_id, "workload": wc["name"], "ms_per_step": ms / 5, "steps": 5,
               "ray_samples_per_s": emitted * 5 / (ms * 1e-3), "dense_steps_per_s": dense * 5 / (ms * 1e-3),
               "mean_samples_per_ray": spp, "loss": loss, "cuda_graph": r.graph}
        try:
            out["parity"] = parity_check(r.model, r.cfg, r.root, r.host[0][0], r.host[0][1], dev, 64)
        except Exception as e:                      # the secondary record must not take the headline down
            out["parity"] = {"error": repr(e)[:200]}
        release(r)
        return out
    except Exception as e:
        return {"config": cfg_id, "workload": wc["name"], "error": repr(e)[:300]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    from ren_b200 import _lib
    lib = _lib.load()
    wc = CONFIGS[args.config]
    graph = args.graph != "off"
    run = Runner(wc, args, rank, world, dev, graph)
    model, trainer = run.model, run.trainer
    names = ["field_fwd", "field_dgrad", "wgrad", "composite_fwd", "composite_bwd", "event_rays", "loss",
             "adam", "pack"]

    # warm-up (eager: the first three steps are never captured), then the per-kernel CUDA-event split and the
    # launch count of one eager step
    run.block(max(args.warmup, 3), False)
    lib.ren_timing_enable(1)
    for k in range(len(names)):
        lib.ren_timing_collect(k, C.byref(C.c_double()), C.byref(C.c_int64()))
    trainer.cuda_graph = False
    n_prof = 3
    launches0 = lib.ren_launch_count()
    run.block(n_prof, False)
    launches_per_step = (lib.ren_launch_count() - launches0) / n_prof
    kern = {}
    for k, nme in enumerate(names):
        tot, cnt = C.c_double(), C.c_int64()
        lib.ren_timing_collect(k, C.byref(tot), C.byref(cnt))
        if cnt.value:
            kern[nme] = {"ms_per_launch": tot.value / cnt.value, "launches": cnt.value,
                         "ms_per_step": tot.value / n_prof}
    lib.ren_timing_enable(0)
    trainer.cuda_graph = run.graph
    graph_note = None
    if run.graph:
        try:
            run.block(2, False)                     # capture + first replays, untimed
        except Exception as e:                      # capture refused: stay eager and say so
            graph_note = repr(e)[:200]
            trainer.cuda_graph = run.graph = False
            torch.cuda.synchronize()
            run.block(1, False)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, blocks, loss = run.timed(args.steps, False, args.min_ms)
    run.block(min(args.warmup, 2), True)
    ms_e2e, blocks_e2e, _ = run.timed(args.steps, True, args.min_ms / 2)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    spp = run.mean_spp()
    samples_per_step = wc["renders"] * wc["rays"] * spp * world
    value = samples_per_step * args.steps / (ms * 1e-3)
    e2e = samples_per_step * args.steps / (ms_e2e * 1e-3)

    params_equal = None
    if world > 1:
        flat = model.nerf.flat_params()
        chk = torch.stack((flat.double().sum(), flat.double().abs().sum(), flat.double().pow(2).sum()))
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        params_equal = all(bool(torch.equal(allc[0], c)) for c in allc)
    if rank != 0:
        leave(world)
        return

    pk = peaks()
    engine = model.nerf.engine
    engine_is_tc = engine != "simt_fp32"
    ncu = ncu_metrics() if (engine_is_tc and args.config == 2) else None
    roof = None
    if "field_fwd" in kern:
        per_launch_samples = wc["renders"] * wc["rays"] * wc["n"] if args.config == 2 else None
        if per_launch_samples:
            ach = per_launch_samples * FLOP_PER_SAMPLE_FWD / (kern["field_fwd"]["ms_per_launch"] * 1e-3) / 1e12
            # ~100 B per ray per render (o, d, jitter in; C, O, D out) + the 2.38 MB of weights
            alg = wc["renders"] * wc["rays"] * 100 + 595_586 * 4
            roof = {"bound": "tensor", "kernel": "field_fwd (ren_tc_chain<0,1,0>)", "achieved": ach,
                    "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                    "traffic": ncu.get("field_fwd_dram_bytes") if ncu else None, "traffic_unit": "bytes/launch",
                    "algorithmic_bytes": alg,
                    "stash_bytes": per_launch_samples * 5096,
                    "tensor_pipe_active_pct_ncu": ncu.get("field_fwd_tensor_pipe_pct") if ncu else None,
                    "ncu_source": ncu.get("source") if ncu else None,
                    "peak_source": pk["source"], "samples_per_launch": per_launch_samples,
                    "flop_per_sample": FLOP_PER_SAMPLE_FWD,
                    "note": "fp16 hi/lo split executes 3 MMAs per algorithmic MMA: ceiling of frac is 1/3; "
                            "traffic is the activation stash the training forward must write (stash_bytes), "
                            "not algorithmic traffic"}
    # the HBM-bound kernel of the step beside the tensor-bound one: weight gradients read the activation stash and the
    # deltas once each (DRAM bytes of the committed ncu capture, this run's CUDA-event time, measured copy peak)
    roof_wgrad = None
    try:
        wk = (ncu or {}).get("kernels", {}).get("ren_tc_wgrad2")
        if wk and "wgrad" in kern and args.config == 2:
            gb = wk["dram_read_gb"] + wk["dram_write_gb"]
            ach = gb / (kern["wgrad"]["ms_per_launch"] * 1e-3)
            roof_wgrad = {"bound": "hbm", "kernel": "wgrad (ren_tc_wgrad2)", "achieved": ach, "peak": pk["hbm"],
                          "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": gb * 1e9, "traffic_unit": "bytes/launch",
                          "algorithmic_bytes": per_launch_samples * 2 * 2 * 2432 if per_launch_samples else None,
                          "note": "algorithmic bytes = fp16 stash + fp16 deltas of the 2432 hidden activations per "
                                  "sample, read once each; traffic from the committed ncu capture"}
    except Exception:
        roof_wgrad = None
    parity = None
    try:
        parity = parity_check(model, run.cfg, run.root, run.host[0][0], run.host[0][1], dev)
    except Exception as e:
        parity = {"error": repr(e)[:300]}
    fwd_only = None
    if world == 1:
        try:
            fwd_only = forward_only_rate(run)
        except Exception as e:                      # an extra record must not take the headline down
            fwd_only = {"error": repr(e)[:300]}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import synth as osynth
        oroot = osynth.write_dataset(os.path.join(tempfile.gettempdir(), "ren_b200_bench_ref"))
        n_rays = 256
        rate, per_step, cores, _ = cpu_oracle_rate(CONFIGS[2], n_rays, 2, 1, oroot)
        cpu = {"value": rate, "unit": "ray-samples/s", "cores": cores, "kind": "port",
               "sample": f"{n_rays} rays x 128 samples x 2 renders per step, 2 timed steps "
                         "(fwd+bwd, fp32, torch CPU oracle port of the reference path)"}
    h2d_bytes, graph_on, run_cfg = run.h2d, run.graph, None
    secondary = None
    if world == 1 and args.config == 2 and not args.no_secondary:
        del trainer
        release(run)
        secondary = [run_secondary(c, args, rank, world, dev, graph) for c in (3, 4, 5, 6)]
    out = {
        "metric": "ray-samples/s", "value": value, "unit": "ray-samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if engine == "simt_fp32" else
                 "fwd + dgrad: f16 hi/lo split operands (3 MMAs), f32 accumulate; wgrad: single f16 operands, "
                 "f32 accumulate; geometry / compositing / loss / Adam f32 (timestamps f64)",
        "data": "synthetic",
        "config": {"workload": wc["name"], "config_id": args.config, "engine": engine, "rays_per_gpu": wc["rays"],
                   "samples_per_ray": spp, "marching_steps_per_ray": wc["n"], "renders_per_step": wc["renders"],
                   "cuda_graph": graph_on if graph_note is None else {"on": False, "capture_error": graph_note},
                   "adam_lr": BENCH_LR, "finite": bool(math.isfinite(loss)),
                   "l2": "per-step working set (activation stash ~11 GB) >> 126 MB L2; 4 rotating input batches",
                   "parallelism": f"dp{world}"},
        "blocks": {"n": len(blocks), "steps_per_block": args.steps, "ms_per_step_median": ms / args.steps,
                   "ms_per_step_min": min(blocks) / args.steps, "ms_per_step_max": max(blocks) / args.steps,
                   "timed_ms_total": sum(blocks)},
        "e2e": {"value": e2e, "unit": "ray-samples/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "blocks": len(blocks_e2e)},
        "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step,
        "roofline": roof, "parity": parity, "cpu_baseline": cpu, "kernels": kern, "forward_only": fwd_only,
        "roofline_wgrad": roof_wgrad,
        "params_equal_across_ranks": params_equal, "secondary": secondary,
        "clocks": sampler.summary(), "loss": loss,
    }
    print(json.dumps(out), flush=True)
    leave(world)


def leave(world):
    """Multi-rank exit without communicator teardown: every collective of the run has completed on every rank (the
    last one is the parameter-checksum all-gather), and `destroy_process_group` with captured graphs that hold
    NCCL kernels still alive was seen to hang at 8 ranks.  torchrun only needs exit code 0 from every rank."""
    if world > 1:
        import torch
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))   # 6: config-4 shape, arch ngp
    ap.add_argument("--graph", default="auto", choices=["auto", "off"])
    ap.add_argument("--min-ms", type=float, default=MIN_TIMED_MS,
                    help="repeat the timed block until this much GPU time (profiler runs pass 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
