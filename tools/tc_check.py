"""GPU diagnostic: tcgen05 engine vs SIMT engine vs reference golden on one render."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers as H
from oracle import synth
root = synth.write_dataset("/tmp/ren_tc_check")
g = H.load_golden(sys.argv[1] if len(sys.argv) > 1 else "step_8x256_gain2")
li = lambda r: torch.log(r.double() + 1e-3)
rr = g["renders"]["start"]
outs = {}
for eng in ("simt_fp32", "tc_f16x3", "tc_f16"):
    m = H.build_product(g, root, "cuda", eng)
    with torch.no_grad():
        rad, opac, depth, spp = m.nerf(rr["o"].cuda(), rr["d"].cuda(), g["jitter"]["diff_start"].cuda())
    torch.cuda.synchronize()
    outs[eng] = rad
    nw, ma, _ = H.rel_err(li(rad), li(rr["radiance"]))
    print(f"{eng}: logI normwise={nw:.3e} maxabs={ma:.3e} opacity maxabs={H.rel_err(opac, rr['opacity'])[1]:.2e} spp={float(spp)}", flush=True)
# timing at config-2 size, no grad
m = H.build_product(g, root, "cuda", "tc_f16x3")
B = 8192
o = torch.randn(B, 3, device="cuda"); o = 3 * o / o.norm(dim=-1, keepdim=True)
d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(B, 3, device="cuda"), dim=-1)
u = torch.rand(B, device="cuda")
for eng in ("tc_f16x3", "tc_f16", "simt_fp32"):
    m.nerf.engine = eng; m.nerf._desc_cache.clear()
    with torch.no_grad():
        for _ in range(2): m.nerf(o, d, u)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5): m.nerf(o, d, u)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"{eng}: {B}x128 no-grad render {dt*1e3:.2f} ms -> {B*128/dt:.3e} ray-samples/s", flush=True)
