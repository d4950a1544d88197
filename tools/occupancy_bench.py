"""GPU timing of a real-sequence-like render step with an occupancy grid (BASELINE config 4 rendering options:
non-cubic aabb + sphere contraction, render_step_size auto (~1100 dense steps per ray), cone_angle 0.004): a sparse
synthetic grid stands in for a trained scene; dense march vs dead-tile skipping vs packed samples (fwd + bwd)."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers as H
from oracle import synth
root = synth.write_dataset("/tmp/ren_occ_bench")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
aabb = [0.5, -2.1, 0.6, 2.0, -0.6, 1.6]
g = dict(H.load_golden("step_8x256_gain2"))
g["fcfg"] = dict(g["fcfg"], aabb=aabb, contraction="sphere")
step = math.sqrt(3) * 1.5 / 1024
g["rcfg"] = dict(g["rcfg"], near_plane=0.05, far_plane=3.0, render_step_size=step, cone_angle=0.004,
                 early_stop_eps=1e-4)
cfg = H.product_config(g)
cfg["nerf"]["occ_grid"] = dict(resolution=128, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, enabled=True)
from ren_b200 import RobustENeRF
m = RobustENeRF(dataset_directory=root, engine="auto", **cfg).cuda().train()
gen = torch.Generator().manual_seed(0)
# a "surface": occupied cells in a thin shell of the contracted cube (~3 % of the cells)
idx = torch.stack(torch.meshgrid(*[torch.arange(128)] * 3, indexing="ij"), -1).float()
rad = ((idx + 0.5) / 128 - 0.5).norm(dim=-1)
grid = ((rad > 0.10) & (rad < 0.115)).cuda()
ang = torch.rand(B, generator=gen) * 6.2831853
o = (torch.tensor(aabb[:3]) + torch.tensor(aabb[3:])) / 2 + 0.9 * torch.stack([torch.cos(ang), torch.sin(ang), 0.2 * torch.sin(3 * ang)], -1)
d = torch.nn.functional.normalize(-0.9 * torch.stack([torch.cos(ang), torch.sin(ang), 0 * ang], -1) + 0.3 * torch.randn(B, 3, generator=gen), dim=-1)
o, d, u = o.cuda(), d.cuda(), torch.rand(B, generator=gen).cuda()
res = {}
for mode in ("dense, all cells occupied", "grid, dead-tile skipping", "grid, packed samples"):
    if mode.startswith("dense"):
        m.nerf.occupancy_grid._binary = torch.ones_like(grid)
        m.nerf.packed_march = False
    else:
        m.nerf.occupancy_grid._binary = grid
        m.nerf.packed_march = mode.endswith("packed samples")
    def run():
        rad_, op, dep, nk = m.nerf.render_raw(o, d, u)
        rad_.sum().backward()
        return nk
    for _ in range(2): nk = run()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): nk = run()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    K = m.nerf._descs()[2]
    res[mode] = dict(ms_fwd_bwd=dt * 1e3, emitted_samples=int(nk.sum()), dense_steps=B * K)
    print(mode, res[mode], flush=True)
print(json.dumps({"workload": f"{B} rays, config-4 rendering options, 128^3 grid with {float(grid.float().mean())*100:.1f} % occupied cells, "
                              "8x256 MLP, one render fwd+bwd", "results": res}))
