#!/usr/bin/env python
"""bench.py -- ray-samples/s of the event-NeRF training hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one `RobustENeRF.training_step` + backward + gradient all-reduce (N>1) + Adam on
one synthetic event batch of BASELINE config 2: 4096 rays/GPU x 128 samples x 2 renders
(= 1 048 576 ray-samples per GPU per step), 8-layer x 256-wide MLP.  One ray-sample = one
(ray, sample) pair pushed through field + compositing in one render.

  value     whole-job ray-samples/s, event batch already resident in HBM
  e2e       same through the public API with HOST (pinned) batches: H2D copy of every batch
            and a D2H read of the loss inside the timed region
  roofline  dominant kernel (field forward) against the measured bf16 tensor peak:
            1 186 304 algorithmic FLOP per ray-sample / CUDA-event time
  cpu_baseline  the oracle port of the reference's own Python path on the host cores,
            bounded sample of the same workload

`--impl reference` times that CPU path alone (the upstream reference has no CPU mode and
cannot be installed here: nerfacc / tiny-cuda-nn are CUDA-only and absent).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "robust-e-nerf_b200"))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
N_SAMPLES = 128
RENDERS = 2
FLOP_PER_SAMPLE_FWD = 1_186_304
WORKLOAD = ("config2: ficus-easy-like synthetic 640x480 event stream, 4096 rays/GPU x 128 samples x 2 "
            "renders, 8Lx256 MLP, l_diff (mse), frozen Cp/tau, fwd+bwd+Adam")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), hbm=p["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json, bf16 sustained)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


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


def cpu_oracle_rate(n_rays, steps, warmup, dataset_dir):
    """Oracle port of the reference training step (fwd + bwd) on the host cores."""
    import torch
    from oracle import restatement as R
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    scene = synth.load_scene(dataset_dir)
    fcfg = R.FieldConfig()
    step = 4.0 / N_SAMPLES
    rcfg = R.RenderConfig(near_plane=0.5, far_plane=0.5 + (N_SAMPLES + 0.5) * step, render_step_size=step)
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
    return RENDERS * n_rays * N_SAMPLES / per_step, per_step, torch.get_num_threads(), loss


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    root = synth.write_dataset(os.path.join(tempfile.gettempdir(), "ren_b200_bench_ref"))
    n_rays = 256
    rate, per_step, cores, loss = cpu_oracle_rate(n_rays, args.steps, args.warmup, root)
    sample = f"{n_rays} rays x {N_SAMPLES} samples x {RENDERS} renders per step (fwd+bwd, fp32, torch CPU)"
    print(json.dumps({
        "impl": "reference", "metric": "ray-samples/s", "value": rate, "unit": "ray-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "ray-samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": loss}))


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
    from ren_b200 import RobustENeRF, _lib, synthetic
    from ren_b200.trainer import FlatTrainer
    lib = _lib.load()

    root = os.path.join(tempfile.gettempdir(), f"ren_b200_bench_{rank}")
    synthetic.write_dataset(root)
    cfg = synthetic.default_config(n_samples=N_SAMPLES)
    cfg["num_nodes"], cfg["gpus"] = 1, list(range(world))
    torch.manual_seed(0)
    model = RobustENeRF(dataset_directory=root, engine=args.engine, **cfg).to(dev).train()
    trainer = FlatTrainer(model, world_size=world)
    pose_ts = model.trajectory.T_wc_timestamp.cpu()
    n_batches = 4
    host = [synthetic.make_batch(RAYS_PER_GPU, pose_ts, rank=rank, seed=1234 + 17 * i, pin=True)
            for i in range(n_batches)]
    resident = [(synthetic.to_device(b, dev), {k: v.to(dev) for k, v in j.items()}) for b, j in host]
    h2d = synthetic.batch_nbytes(host[0][0]) + sum(v.numel() * 4 for v in host[0][1].values())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(n_steps, e2e):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = None
        for i in range(n_steps):
            if e2e:
                hb, hj = host[i % n_batches]
                batch = synthetic.to_device(hb, dev)
                jit = {k: v.to(dev, non_blocking=True) for k, v in hj.items()}
                last = float(trainer.step(batch, i, jit))            # D2H read of the loss
            else:
                batch, jit = resident[i % n_batches]
                last = trainer.step(batch, i, jit)
        b.record()
        sync_all()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, float(last)

    timed(args.warmup, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.ren_timing_enable(1)
    for k in range(9):
        import ctypes as C
        lib.ren_timing_collect(k, C.byref(C.c_double()), C.byref(C.c_int64()))
    launches0 = lib.ren_launch_count()
    ms, loss = timed(args.steps, False)
    launches = lib.ren_launch_count() - launches0
    kern = {}
    import ctypes as C
    names = ["field_fwd", "field_dgrad", "wgrad", "composite_fwd", "composite_bwd", "event_rays", "loss",
             "adam", "pack"]
    for k, nme in enumerate(names):
        tot, cnt = C.c_double(), C.c_int64()
        lib.ren_timing_collect(k, C.byref(tot), C.byref(cnt))
        if cnt.value:
            kern[nme] = {"ms_per_launch": tot.value / cnt.value, "launches": cnt.value,
                         "ms_per_step": tot.value / args.steps}
    lib.ren_timing_enable(0)
    timed(min(args.warmup, 2), True)
    ms_e2e, _ = timed(args.steps, True)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    samples_per_step = RENDERS * RAYS_PER_GPU * N_SAMPLES * world
    value = samples_per_step * args.steps / (ms * 1e-3)
    e2e = samples_per_step * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    engine_is_tc = model.nerf.engine != "simt_fp32"
    roof = None
    if "field_fwd" in kern:
        per_launch_samples = RENDERS * RAYS_PER_GPU * N_SAMPLES
        ach = per_launch_samples * FLOP_PER_SAMPLE_FWD / (kern["field_fwd"]["ms_per_launch"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "field_fwd", "achieved": ach, "peak": pk["tflops"],
                "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                # dram__bytes_read.sum + dram__bytes_write.sum of ren_tc_chain<0,1> for this exact launch shape
                # (ncu --set full, profiles/r01_tc_v32_ncu_summary.txt); algorithmic = 5.1 KB stash per sample
                "traffic": 5.391e9 if engine_is_tc else None, "traffic_unit": "bytes/launch",
                "algorithmic_bytes": per_launch_samples * 5096,
                "tensor_pipe_active_pct_ncu": 50.8 if engine_is_tc else None,
                "peak_source": pk["source"], "samples_per_launch": per_launch_samples,
                "flop_per_sample": FLOP_PER_SAMPLE_FWD,
                "note": "fp16 hi/lo split executes 3 MMAs per algorithmic MMA: ceiling of frac is 1/3"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import synth as osynth
        oroot = osynth.write_dataset(os.path.join(tempfile.gettempdir(), "ren_b200_bench_ref"))
        n_rays = 256
        rate, per_step, cores, _ = cpu_oracle_rate(n_rays, 2, 1, oroot)
        cpu = {"value": rate, "unit": "ray-samples/s", "cores": cores, "kind": "port",
               "sample": f"{n_rays} rays x {N_SAMPLES} samples x {RENDERS} renders per step, 2 timed steps "
                         "(fwd+bwd, fp32, torch CPU oracle)"}
    engine = model.nerf.engine
    out = {
        "metric": "ray-samples/s", "value": value, "unit": "ray-samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if engine == "simt_fp32" else "f16x3-split operands, f32 accumulate (fwd); f32 (bwd)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "engine": engine, "rays_per_gpu": RAYS_PER_GPU,
                   "samples_per_ray": N_SAMPLES, "renders_per_step": RENDERS,
                   "l2": "per-step working set (activation stash ~11 GB) >> 126 MB L2; 4 rotating input batches",
                   "parallelism": f"dp{world}"},
        "e2e": {"value": e2e, "unit": "ray-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "kernels": kern,
        "clocks": sampler.summary(), "loss": loss,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
