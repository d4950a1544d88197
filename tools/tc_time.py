"""GPU timing probe: no-grad tcgen05 render at config-2 size under the REN_TC_DBG experiment flags."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers as H
from oracle import synth
root = synth.write_dataset("/tmp/ren_tc_time")
g = H.load_golden("step_8x256")
m = H.build_product(g, root, "cuda", "tc_f16x3")
B = 8192
o = torch.randn(B, 3, device="cuda"); o = 3 * o / o.norm(dim=-1, keepdim=True)
d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(B, 3, device="cuda"), dim=-1)
u = torch.rand(B, device="cuda")
with torch.no_grad():
    for _ in range(3): m.nerf(o, d, u)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): m.nerf(o, d, u)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print(f"REN_TC_DBG={os.environ.get('REN_TC_DBG','0')}: {B}x128 no-grad render {dt*1e3:.3f} ms", flush=True)
