"""GPU profiling target: a few no-grad renders (config-2 size) with one engine."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers as H
from oracle import synth
root = synth.write_dataset("/tmp/ren_tc_prof")
g = H.load_golden("step_8x256")
eng = sys.argv[1] if len(sys.argv) > 1 else "tc_f16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = H.build_product(g, root, "cuda", eng)
o = torch.randn(B, 3, device="cuda"); o = 3 * o / o.norm(dim=-1, keepdim=True)
d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(B, 3, device="cuda"), dim=-1)
u = torch.rand(B, device="cuda")
with torch.no_grad():
    for _ in range(n):
        m.nerf(o, d, u)
torch.cuda.synchronize()
print("done")
