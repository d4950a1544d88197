"""Which part of the training step refuses CUDA-graph capture?  Captures growing prefixes of one step."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from ren_b200 import RobustENeRF, synthetic, _lib
from ren_b200.trainer import FlatTrainer

root = synthetic.write_dataset("/tmp/ren_b200_graph_probe")
cfg = synthetic.default_config(n_samples=128)
torch.manual_seed(0)
model = RobustENeRF(dataset_directory=root, **cfg).cuda().train()
tr = FlatTrainer(model, cuda_graph=False)
pose_ts = model.trajectory.T_wc_timestamp.cpu()
b, j = synthetic.make_batch(1024, pose_ts)
b, j = synthetic.to_device(b, "cuda"), {k: v.cuda() for k, v in j.items()}
for i in range(3):
    tr.step(b, i, j)
torch.cuda.synchronize()


def stage_rays():
    ev = {k: v.squeeze(0) for k, v in b["event"].items()}
    nz = {k: v.squeeze(0) for k, v in b["normalized"].items()}
    return model.event_rays(ev, nz, 2)


def stage_fwd_nograd():
    with torch.no_grad():
        return model.training_step(b, 0, j)


def stage_fwd():
    return model.training_step(b, 0, j)


def stage_fwd_bwd():
    for p in model.parameters():
        p.grad = None
    model.training_step(b, 0, j).backward()


def stage_full():
    tr._step_body(b, 0, j)


for mode in ("global", "thread_local", "relaxed"):
    for name, fn in (("event_rays", stage_rays), ("forward no-grad", stage_fwd_nograd), ("forward", stage_fwd),
                     ("forward+backward", stage_fwd_bwd), ("full step", stage_full)):
        g = torch.cuda.CUDAGraph()
        try:
            torch.cuda.synchronize()
            with torch.cuda.graph(g, capture_error_mode=mode):
                fn()
            g.replay()
            torch.cuda.synchronize()
            print(f"[{mode}] {name}: captured + replayed OK", flush=True)
        except Exception as e:
            print(f"[{mode}] {name}: FAILED: {repr(e)[:300]}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("   sync after failure:", repr(e2)[:200])
        print("   last lib error:", _lib.load().ren_last_error().decode()[:200], flush=True)
