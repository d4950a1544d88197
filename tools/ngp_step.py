"""One `arch: ngp` training step at config 4's rendering options (see bench.py CONFIGS[6]) for profiling."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "robust-e-nerf_b200"))
import torch
import bench
from ren_b200 import synthetic
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wc = dict(bench.CONFIGS[6], rays=rays)
model, cfg, root = bench.build_model(wc, 1, 0, torch.device("cuda", 0), "auto")
from ren_b200.trainer import FlatTrainer
tr = FlatTrainer(model)
pose_ts = model.trajectory.T_wc_timestamp.cpu()
b, j = synthetic.make_batch(rays, pose_ts)
b, j = synthetic.to_device(b, "cuda"), {k: v.cuda() for k, v in j.items()}
model.global_step = 0          # step 0 updates the occupancy grid (every cell of the random-init field stays occupied)
tr.step(b, 0, j)
model.global_step = 1          # no occupancy-grid update in the timed steps
for i in range(steps):
    torch.cuda.synchronize(); t = time.perf_counter()
    loss = tr.step(b, i + 1, j)
    model.global_step = 1
    torch.cuda.synchronize()
    print("step", i, "ms", (time.perf_counter() - t) * 1e3, "loss", float(loss), flush=True)
