"""Timing of the tcgen05 weight-gradient kernel variants on the config-2 shape (per-kernel CUDA-event times)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from ren_b200 import RobustENeRF, synthetic, _lib
lib = _lib.load()
root = synthetic.write_dataset("/tmp/ren_b200_wgprobe")
cfg = synthetic.default_config(n_samples=128)
torch.manual_seed(0)
model = RobustENeRF(dataset_directory=root, **cfg).cuda().train()
b, j = synthetic.make_batch(4096, model.trajectory.T_wc_timestamp.cpu())
b, j = synthetic.to_device(b, "cuda"), {k: v.cuda() for k, v in j.items()}
def run(n):
    for _ in range(n):
        for p in model.parameters():
            p.grad = None
        model.training_step(b, 0, j).backward()
    torch.cuda.synchronize()
run(3)
lib.ren_timing_enable(1)
for k in range(9):
    lib.ren_timing_collect(k, C.byref(C.c_double()), C.byref(C.c_int64()))
run(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
tot, cnt = C.c_double(), C.c_int64()
lib.ren_timing_collect(2, C.byref(tot), C.byref(cnt))
print(os.environ.get("REN_WGRAD_V1"), os.environ.get("REN_W2_DBG"), "wgrad ms/launch", tot.value / max(cnt.value, 1), flush=True)
