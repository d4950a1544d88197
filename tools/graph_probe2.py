"""Bench-like flow at 4096 rays: which ingredient breaks CUDA-graph capture of the step?"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from ren_b200 import RobustENeRF, synthetic, _lib
from ren_b200.trainer import FlatTrainer

variant = sys.argv[1]
lib = _lib.load()
root = synthetic.write_dataset("/tmp/ren_b200_graph_probe2")
cfg = synthetic.default_config(n_samples=128)
cfg["optimizer"]["lr"]["default"] = 5e-4
torch.manual_seed(0)
model = RobustENeRF(dataset_directory=root, **cfg).cuda().train()
tr = FlatTrainer(model, cuda_graph=(variant in ("trainer", "timing+trainer")))
pose_ts = model.trajectory.T_wc_timestamp.cpu()
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
MODE = sys.argv[3] if len(sys.argv) > 3 else "global"
batches = [synthetic.make_batch(B, pose_ts, seed=1234 + 17 * i, pin=True) for i in range(4)]
res = [(synthetic.to_device(b, "cuda"), {k: v.cuda() for k, v in j.items()}) for b, j in batches]
tr.cuda_graph_saved = tr.cuda_graph
tr.cuda_graph = False
for i in range(3):
    tr.step(res[i % 4][0], i, res[i % 4][1])
torch.cuda.synchronize()
if "timing" in variant:
    lib.ren_timing_enable(1)
    for i in range(3):
        tr.step(res[i % 4][0], i, res[i % 4][1])
    for k in range(9):
        lib.ren_timing_collect(k, C.byref(C.c_double()), C.byref(C.c_int64()))
    lib.ren_timing_enable(0)
torch.cuda.synchronize()
print(variant, "mem allocated GB", torch.cuda.memory_allocated() / 2**30, "reserved", torch.cuda.memory_reserved() / 2**30, flush=True)
try:
    if variant in ("trainer", "timing+trainer"):
        tr.cuda_graph = True
        for i in range(4):
            loss = tr.step(res[i % 4][0], 10 + i, res[i % 4][1])
        torch.cuda.synchronize()
        print(variant, "graph steps OK, loss", float(loss), "reserved GB", torch.cuda.memory_reserved() / 2**30)
    else:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=MODE):
            tr._step_body(res[0][0], 0, res[0][1])
        g.replay()
        torch.cuda.synchronize()
        print(variant, B, MODE, "captured + replayed OK; reserved GB", torch.cuda.memory_reserved() / 2**30)
except Exception as e:
    import traceback
    print(variant, B, MODE, "FAILED:", repr(e)[:200])
    print("".join(traceback.format_exc().splitlines(True)[-24:]))
    print("lib error:", lib.ren_last_error().decode()[:300])
