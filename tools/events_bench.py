"""GPU timing of the raw-event conversion (ren_event_queue) next to the oracle and the reference loop rate."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np, torch
from oracle import events_prep as EP
from ren_b200 import events as E, _lib
import ctypes as C
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
raw = EP.synth_raw_events(n, 640, 480, seed=3)
dev = torch.device("cuda")
pos = torch.from_numpy(raw["position"].view(np.uint8).reshape(-1)).to(dev)
ts = torch.from_numpy(raw["timestamp"]).to(dev)
pol = torch.from_numpy(raw["polarity"].astype(np.uint8)).to(dev)
out = [torch.empty((n,), dtype=torch.int64, device=dev) for _ in range(4)]
opos = torch.empty((n, 2), dtype=torch.int64, device=dev)
nv = torch.zeros(1, dtype=torch.int64, device=dev); mi = torch.zeros(1, dtype=torch.int64, device=dev)
lib = _lib.load(); nb = C.c_size_t(0)
_lib.check(lib.ren_event_queue_workspace_bytes(n, 640, 480, C.byref(nb)), "ws")
ws = torch.empty((nb.value,), dtype=torch.uint8, device=dev)
def run():
    _lib.check(lib.ren_event_queue(n, 640, 480, _lib.ptr(pos), _lib.ptr(ts), _lib.ptr(pol), _lib.ptr(opos), _lib.ptr(out[0]),
                                   _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]), _lib.ptr(nv), _lib.ptr(mi),
                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "run")
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
# algorithmic bytes per event: read 4 (pos) + 8 (ts) + 1 (pol); write up to 6 x 8 = 48 for a kept event
kept = int(nv.item())
alg = n * 13 + kept * 48
print(f"ren_event_queue: {n} events ({kept} kept) in {ms:.3f} ms -> {n/ms*1e3:.3e} events/s, "
      f"algorithmic {alg/ms/1e6:.0f} GB/s")
t = time.perf_counter(); EP.queue_raw_events(raw, {"img_width": 640}); dt = time.perf_counter() - t
print(f"numpy oracle (1 core, vectorised sort): {n/dt:.3e} events/s")
