"""Occupancy grid on the GPU: density queries, grid-masked marching and the
nerfacc-style grid update, against the torch-CPU oracle.  nerfacc==0.3.1 is not vendored under /root/reference
(parity unpinned at that seam): the oracle restates its published algorithm."""
import pytest
import torch

import helpers as H
from oracle import restatement as R
from oracle import synth

pytestmark = pytest.mark.gpu
ENGINES = ["simt_fp32", "auto"]


def _model(dataset_dir, engine, contraction="sphere", res=8, enabled=True, **occ):
    g = dict(H.load_golden("step_8x256_gain2"))
    g["fcfg"] = dict(g["fcfg"], contraction=contraction)
    cfg = H.product_config(g)
    cfg["nerf"]["occ_grid"] = dict(resolution=res, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16,
                                   enabled=enabled, **occ)
    from ren_b200 import RobustENeRF
    m = RobustENeRF(dataset_directory=dataset_dir, engine=engine, **cfg)
    m.load_state_dict({"nerf.radiance_field." + k: v for k, v in H.golden_params(g).items()}, strict=False)
    return m.cuda().train(), g


def _rays(B, seed):
    gen = torch.Generator().manual_seed(seed)
    ang = torch.rand(B, generator=gen) * 6.2831853
    o = torch.stack([3 * torch.cos(ang), 3 * torch.sin(ang), 0.2 * torch.sin(3 * ang)], -1)
    d = torch.nn.functional.normalize(-o + 0.4 * torch.randn(B, 3, generator=gen), dim=-1)
    return o, d, torch.rand(B, generator=gen)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("contraction", ["sphere", "tanh", "aabb"])
def test_query_density_matches_oracle(contraction, engine, built_lib, dataset_dir):
    m, g = _model(dataset_dir, engine, contraction, enabled=False)
    fcfg, _, _ = H.golden_configs(g)
    x = (torch.rand(1000, 3, generator=torch.Generator().manual_seed(1)) - 0.5) * 5.0
    got = m.nerf.query_density(x.cuda()).cpu()
    with torch.no_grad():
        ref = R.field_density(H.golden_params(g), fcfg, x).reshape(-1)
    nw, ma, _ = H.rel_err(got, ref)
    print(f"[density {contraction}/{engine}] normwise={nw:.2e} maxabs={ma:.2e}")
    assert nw < 1e-4
    assert torch.equal(got == 0, ref == 0)                    # the selector mask agrees exactly


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("contraction", ["sphere", "tanh", "aabb"])
def test_grid_masked_render_matches_oracle(contraction, engine, built_lib, dataset_dir):
    """A random binary grid: the dense marcher drops the steps that fall in unoccupied cells exactly like the
    packed nerfacc marcher would (sample counts, log I, gradients)."""
    m, g = _model(dataset_dir, engine, contraction, res=8)
    fcfg, rcfg, _ = H.golden_configs(g)
    grid = torch.rand(8, 8, 8, generator=torch.Generator().manual_seed(4)) > 0.45
    m.nerf.occupancy_grid._binary = grid.cuda()
    o, d, u = _rays(64, 7)
    rad, opac, depth, spp = m.nerf(o.cuda(), d.cuda(), u.cuda())
    wts = torch.randn(64, generator=torch.Generator().manual_seed(2))
    (rad * wts.cuda()).sum().backward()
    params = {k: v.clone().requires_grad_(True) for k, v in H.golden_params(g).items()}
    th_b = torch.log(torch.expm1(torch.ones(1))).requires_grad_(True)
    rad_o, opac_o, depth_o, spp_o = R.render(params, fcfg, rcfg, o, d, u, R.softplus(th_b, 1.0), grid_binary=grid)
    (rad_o * wts).sum().backward()
    nw, ma, _ = H.rel_err(torch.log(rad.double() + 1e-3), torch.log(rad_o.double() + 1e-3))
    print(f"[grid {contraction}/{engine}] logI normwise={nw:.2e} maxabs={ma:.2e} spp={float(spp):.3f}/{spp_o:.3f}")
    assert nw < 1e-4 and ma < 3e-4 and H.rel_err(opac, opac_o)[1] < 1e-4
    assert abs(float(spp) - spp_o) < 1e-6 * spp_o + 1e-9 and spp_o < 128         # samples really were dropped
    named = dict(m.nerf.named_parameters())
    for k, v in params.items():
        a, b = named["radiance_field." + k].grad.double().cpu().reshape(-1), v.grad.double().reshape(-1)
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp(min=1e-30))
        assert cos > 0.9999 and float((a - b).norm() / b.norm()) < 1e-2, (k, cos)
    # an all-true grid is the ungridded render, bit for bit
    m.nerf.occupancy_grid._binary = torch.ones(8, 8, 8, dtype=torch.bool).cuda()
    with torch.no_grad():
        full = m.nerf(o.cuda(), d.cuda(), u.cuda())
        m.nerf.occupancy_grid = None
        m.nerf._desc_cache.clear()
        none = m.nerf(o.cuda(), d.cuda(), u.cuda())
    assert torch.equal(full[0], none[0]) and float(full[3]) == float(none[3])
    assert float(full[3]) == 128.0 or contraction == "aabb"                      # aabb: only the chord in the box


def test_grid_update_matches_restatement(built_lib, dataset_dir):
    """`update_occ_grid` at step 0 (warm-up: every cell, jittered) vs the same update restated on the CPU with the
    oracle's density and the same jitter stream; then a later step keeps the EMA / threshold rule."""
    m, g = _model(dataset_dir, "auto", "sphere", res=8)
    fcfg, rcfg, _ = H.golden_configs(g)
    T_pos = synth.load_scene(dataset_dir)["T_wc_position"].cuda()
    gen = torch.Generator(device="cuda").manual_seed(5)
    m.nerf.update_occ_grid(0, T_pos, generator=gen)
    grid = m.nerf.occupancy_grid
    # restatement
    gen2 = torch.Generator(device="cuda").manual_seed(5)
    idx = torch.arange(512)
    coords = torch.stack((idx // 64, (idx // 8) % 8, idx % 8), -1)
    jit = torch.rand((512, 3), device="cuda", dtype=torch.float32, generator=gen2).cpu()
    x = (coords + jit) / 8
    mask = (x - 0.5).norm(dim=1) < 0.5
    with torch.no_grad():
        occ = R.field_density(H.golden_params(g), fcfg, R.contract_inv(fcfg, x[mask])).reshape(-1) * rcfg.render_step_size
    occs = torch.zeros(512)
    occs[idx[mask]] = torch.maximum(occs[idx[mask]] * 0.95, occ)
    binary = (occs > torch.clamp(occs.mean(), max=1e-2)).view(8, 8, 8)
    nw = H.rel_err(grid.occs, occs)[0]
    print(f"[grid update] occs normwise={nw:.2e} occupied {int(grid.binary.sum())}/{int(binary.sum())} of 512")
    assert nw < 1e-4
    assert int((grid.binary.cpu() != binary).sum()) <= 2          # cells within fp32 noise of the threshold
    assert 0 < int(grid.binary.sum()) < 512
    # not a multiple of n: no update; the next multiple decays unvisited cells and keeps the maximum
    before = grid.occs.clone()
    m.nerf.update_occ_grid(5, T_pos, generator=gen)
    assert torch.equal(before, grid.occs)
    m.nerf.update_occ_grid(16, T_pos, generator=gen)
    assert bool((grid.occs >= before * 0.95 - 1e-12).all())


def test_training_with_occupancy_grid(built_lib, dataset_dir):
    """A few full training steps (l_diff + l_grad, tcgen05 engine) with the grid updating at step 0: fewer samples
    than the dense march, finite loss, state-dict round trip of the grid buffers."""
    m, g = _model(dataset_dir, "auto", "sphere", res=16)
    m.hparams["loss"]["weight"]["log_intensity_grad"] = 1e-3
    # a thin medium with strong contrast: density * step stays below occ_thre, so the threshold is the mean occupancy
    # (over all cells) and the many cells far below it are switched off; a uniform random network would mark every
    # reachable cell occupied
    with torch.no_grad():
        m.nerf.radiance_field.mlp.sigma_layer.output_layer.bias.fill_(-3.0)
        m.nerf.radiance_field.mlp.sigma_layer.output_layer.weight.mul_(20.0)
    scene = synth.load_scene(dataset_dir)
    from ren_b200.trainer import FlatTrainer
    tr = FlatTrainer(m)
    m.global_step = 0
    for i in range(3):
        batch, jitter = synth.make_batch(256, scene["T_wc_timestamp"], rank=0, seed=40 + i)
        loss = tr.step({k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in batch.items()}, i,
                       {k: v.cuda() for k, v in jitter.items()})
        assert bool(torch.isfinite(loss))
    spp = float(m.logged["train/mean_num_samples_per_ray"])
    print(f"[train with grid] mean samples per ray {spp:.2f} of 128, occupied cells "
          f"{int(m.nerf.occupancy_grid.binary.sum())}/4096")
    assert 0 < spp < 128
    sd = m.state_dict()
    for k in ("nerf.occupancy_grid._roi_aabb", "nerf.occupancy_grid.resolution", "nerf.occupancy_grid.occs",
              "nerf.occupancy_grid._binary"):
        assert k in sd, k


@pytest.mark.parametrize("contraction", ["sphere", "aabb"])
@pytest.mark.parametrize("packed", [True, False])
def test_sparse_grid_matches_masked_march(packed, contraction, built_lib, dataset_dir):
    """A sparse grid over long rays (512 marching steps).  The tcgen05 engine either packs the emitted samples
    (ren_march_count / ren_march_fill, the nerfacc layout) or, on the dense layout, skips every 128-step tile that
    emits no sample; both must equal the fp32 engine, which evaluates and masks every step -- value render,
    tangent render, MLP and ray gradients."""
    import time
    grid = torch.zeros(16, 16, 16, dtype=torch.bool)
    gen = torch.Generator().manual_seed(8)
    grid.view(-1)[torch.randperm(4096, generator=gen)[:300]] = True           # ~7 % of the cells
    o, d, u = _rays(192, 3)
    od = 0.05 * torch.randn(192, 3, generator=gen)
    dd = 0.05 * torch.randn(192, 3, generator=gen)
    out = {}
    for engine in ENGINES:
        m, g = _model(dataset_dir, engine, contraction, res=16)
        m.nerf.render_step_size.fill_(4.0 / 512)
        m.nerf.far_plane = 0.5 + 512.5 * (4.0 / 512)
        m.nerf._desc_cache.clear()
        m.nerf.packed_march = packed
        m.nerf.occupancy_grid._binary = grid.cuda()
        oc, dc = o.cuda().requires_grad_(True), d.cuda().requires_grad_(True)
        odc, ddc = od.cuda().requires_grad_(True), dd.cuda().requires_grad_(True)
        torch.cuda.synchronize(); t = time.perf_counter()
        rad, opac, dep, rad_dot, opac_dot, nk = m.nerf.render_tangent_raw(oc, dc, odc, ddc, u.cuda())
        rad2, opac2, _, nk2 = m.nerf.render_raw(oc, dc, u.cuda())
        ((rad_dot ** 2).sum() + rad.sum() + (rad2 * opac2.unsqueeze(-1)).sum()).backward()
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        grads = torch.cat([p.grad.reshape(-1) for p in m.nerf.mlp_parameters()])
        out[engine] = (rad.detach(), rad_dot.detach(), nk.clone(), grads, oc.grad.clone(), ddc.grad.clone(), dt)
    a, b = out["auto"], out["simt_fp32"]
    assert torch.equal(a[2], b[2]) and int(a[2].sum()) < 0.2 * 192 * 512       # same kept samples, sparse
    for name, x, y, bar in (("radiance", a[0], b[0], 1e-5), ("radiance_dot", a[1], b[1], 2e-4),
                            ("mlp grads", a[3], b[3], 5e-3), ("d loss/d o", a[4], b[4], 1e-2),
                            ("d loss/d d_dot", a[5], b[5], 1e-2)):
        nw = H.rel_err(x, y, floor=1e-9)[0]
        print(f"[sparse grid packed={packed}] {name}: tcgen05 vs fp32 normwise={nw:.2e}")
        assert nw < bar, name
    print(f"[sparse grid packed={packed}] kept {int(a[2].sum())} of {192 * 512} steps; tcgen05 {a[6] * 1e3:.1f} ms, fp32 {b[6] * 1e3:.1f} ms")


@pytest.mark.parametrize("packed", [True, False])
def test_empty_grid_renders_background(packed, built_lib, dataset_dir):
    """No occupied cell at all (the state of a fresh nerfacc grid before its first update): zero samples, the
    render is the background, gradients reach the background only."""
    m, g = _model(dataset_dir, "auto", "sphere", res=8)
    m.nerf.packed_march = packed
    m.nerf.occupancy_grid._binary = torch.zeros(8, 8, 8, dtype=torch.bool).cuda()
    o, d, u = _rays(33, 5)
    rad, opac, depth, spp = m.nerf(o.cuda(), d.cuda(), u.cuda())
    bk = float(m.nerf.background().detach())
    assert float(spp) == 0.0 and float(opac.abs().max()) == 0.0
    assert torch.allclose(rad, torch.full_like(rad, bk))
    rad.sum().backward()
    assert float(m.nerf.parametrizations.render_bkgd.original.grad.abs().sum()) > 0
    assert all(float(p.grad.abs().sum()) == 0.0 for p in m.nerf.mlp_parameters())
