"""compute-sanitizer target: one small training step per engine (value + tangent renders) and the event prep."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import helpers as H
from oracle import synth, events_prep as EP
root = synth.write_dataset("/tmp/ren_sanitize")
scene = synth.load_scene(root)
g = dict(H.load_golden("step_8x256"))
g["scfg"] = dict(g["scfg"], w_grad=1e-3)
g["learn_bkgd"] = True
g["freeze_rp"] = False
g["rcfg"] = dict(g["rcfg"], render_step_size=4.0 / 24, far_plane=0.5 + 24.5 * (4.0 / 24))
for eng in ("auto", "simt_fp32"):
    m = H.build_product(g, root, "cuda", eng)
    batch, jitter = synth.make_batch(37, scene["T_wc_timestamp"], rank=0, seed=5)
    loss = m.training_step({k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in batch.items()}, 0,
                           {k: v.cuda() for k, v in jitter.items()})
    loss.backward()
    torch.cuda.synchronize()
    print(eng, float(loss))
# occupancy-grid marching: packed samples (tcgen05) and the masked dense march (fp32)
for eng in ("auto", "simt_fp32"):
    g2 = dict(g)
    g2["fcfg"] = dict(g["fcfg"], contraction="sphere")
    cfg = H.product_config(g2)
    cfg["nerf"]["occ_grid"] = dict(resolution=8, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, enabled=True)
    from ren_b200 import RobustENeRF
    m = RobustENeRF(dataset_directory=root, engine=eng, **cfg).cuda().train()
    m.nerf.occupancy_grid._binary = (torch.rand(8, 8, 8, generator=torch.Generator().manual_seed(4)) > 0.6).cuda()
    o = torch.tensor([[3.0, 0.1, 0.2]]).repeat(19, 1).cuda().requires_grad_(True)
    d = torch.nn.functional.normalize(-o.detach() + 0.3 * torch.randn(19, 3, device="cuda"), dim=-1).requires_grad_(True)
    rad, opac, depth, spp = m.nerf(o, d, torch.rand(19, device="cuda"))
    (rad.sum() + opac.sum()).backward()
    torch.cuda.synchronize()
    print(eng, "grid render", float(spp))
from ren_b200 import events as E
raw = EP.synth_raw_events(20000, 31, 17, seed=2)
q, mi = E._run(raw, {"img_width": 31, "img_height": 17}, "cuda")
print("events", q["end_ts"].numel(), mi)
