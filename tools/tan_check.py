"""GPU diagnostic: tangent render on the tcgen05 engine vs the fp32 engine (outputs, gradients, time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers as H
from oracle import synth
from ren_b200 import _lib
root = synth.write_dataset("/tmp/ren_tan_check")
g = dict(H.load_golden("step_8x256_gain2"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gen = torch.Generator().manual_seed(11)
ang = torch.rand(B, generator=gen) * 6.2831853
o = torch.stack([3 * torch.cos(ang), 3 * torch.sin(ang), 0.2 * torch.sin(3 * ang)], -1).cuda()
d = torch.nn.functional.normalize(-o + 0.5 * torch.randn(B, 3, generator=gen).cuda(), dim=-1)
od = 0.05 * torch.randn(B, 3, generator=gen).cuda()
dd = 0.05 * torch.randn(B, 3, generator=gen).cuda()
u = torch.rand(B, generator=gen).cuda()
res = {}
for eng in ("simt_fp32", "auto"):
    m = H.build_product(g, root, "cuda", eng)
    lib = _lib.load()
    n0 = lib.ren_launch_count()
    out = m.nerf.render_tangent_raw(o, d, od, dd, u)
    rad, opac, dep, rad_dot, opac_dot, nk = out
    loss = (rad_dot ** 2).sum() + (rad * opac_dot).sum() + rad.sum()
    loss.backward()
    torch.cuda.synchronize()
    grads = torch.cat([p.grad.reshape(-1) for p in m.nerf.mlp_parameters()])
    res[eng] = (rad.detach(), rad_dot.detach(), opac_dot.detach(), grads, m.nerf.parametrizations.render_bkgd.original.grad)
    for p in m.nerf.parameters(): p.grad = None
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3):
        out = m.nerf.render_tangent_raw(o, d, od, dd, u)
        ((out[3] ** 2).sum() + out[0].sum()).backward()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print(f"{eng}: launches {lib.ren_launch_count() - n0}, tangent render fwd+bwd {dt*1e3:.2f} ms for {B}x128 samples", flush=True)
a, b = res["auto"], res["simt_fp32"]
for name, x, y in zip(("radiance", "radiance_dot", "opacity_dot", "mlp grads", "bkgd grad"), a, b):
    print(f"{name}: tc vs simt normwise={H.rel_err(x, y, floor=1e-9)[0]:.2e} maxabs={H.rel_err(x, y)[1]:.2e}")
