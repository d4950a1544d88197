"""Host logic of the occupancy grid on the CPU: the nerfacc-style update rule
(models/nerf.py:98-102,170-204; nerfacc==0.3.1 `OccupancyGrid`, un-vendored: restated from its published
algorithm), the inverse contractions, state-dict keys.  The density sweep and the grid-masked march run on the
GPU (tests/test_occupancy_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle import restatement as R


def _grid(contraction, res=8, aabb=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)):
    from ren_b200.nerf import OccupancyGrid
    return OccupancyGrid(list(aabb), res, contraction)


@pytest.mark.parametrize("contraction", ["aabb", "sphere", "tanh"])
def test_contract_inv_inverts_the_field_contraction(contraction):
    """unit cube -> world (`contract_inv`, used to place the grid's sample points) followed by the field's input
    contraction (external/mlp.py:321-335 as restated by the oracle) is the identity."""
    g = _grid(contraction, aabb=(-0.5, -2.1, 0.6, 1.5, 0.4, 2.2))
    fcfg = R.FieldConfig(contraction=contraction, aabb=[-0.5, -2.1, 0.6, 1.5, 0.4, 2.2])
    u = torch.rand(4000, 3, generator=torch.Generator().manual_seed(0)) * 0.96 + 0.02
    if contraction == "sphere":                      # only the inscribed ball of the unit cube is reachable
        u = u[(u - 0.5).norm(dim=1) < 0.49]
    x = g.contract_inv(u)
    back = R.contract_unit(fcfg, x)
    assert float((back - u).abs().max()) < 2e-5
    # and the oracle's own inverse agrees with the product's
    assert float((R.contract_inv(fcfg, u) - x).abs().max()) < 1e-4 * float(x.abs().max())


def test_update_rule_matches_numpy_restatement():
    """EMA-max update + threshold min(mean, occ_thre) over all cells (warm-up), against numpy on the same jitter."""
    g = _grid("aabb", res=6).train()
    dens = lambda x: torch.exp(-4.0 * (x ** 2).sum(-1)) * 0.05
    gen = torch.Generator().manual_seed(3)
    state = gen.get_state()
    g.every_n_step(0, dens, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, generator=gen)
    # restatement
    gen2 = torch.Generator().manual_seed(0)
    gen2.set_state(state)
    idx = np.arange(6 ** 3)
    coords = np.stack([idx // 36, (idx // 6) % 6, idx % 6], -1)
    jitter = torch.rand((6 ** 3, 3), generator=gen2, dtype=torch.float32).numpy()
    x = ((coords + jitter) / 6.0) * 3.0 - 1.5
    occ = np.exp(-4.0 * (x.astype(np.float32) ** 2).sum(-1)) * 0.05
    occs = np.maximum(np.zeros(6 ** 3, np.float32) * 0.95, occ.astype(np.float32))
    binary = occs > min(occs.mean(), 1e-2)
    assert np.allclose(g.occs.numpy(), occs, rtol=1e-5, atol=1e-9)
    assert np.array_equal(g.binary.numpy().reshape(-1), binary)
    assert 0 < int(g.binary.sum()) < 6 ** 3
    # a second update decays the old estimate and keeps the maximum
    before = g.occs.clone()
    g.every_n_step(16, lambda x: torch.zeros(x.shape[0]), occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16,
                   generator=gen)
    assert torch.allclose(g.occs, before * 0.95)
    # steps that are not multiples of n leave the grid alone
    g.every_n_step(17, dens, n=16, generator=gen)
    assert torch.allclose(g.occs, before * 0.95)


def test_update_after_warmup_samples_a_quarter_plus_occupied():
    g = _grid("tanh", res=8).train()
    seen = []

    def dens(x):
        seen.append(x.shape[0])
        return torch.full((x.shape[0],), 0.5)

    gen = torch.Generator().manual_seed(1)
    g._binary = torch.zeros(8, 8, 8, dtype=torch.bool)
    g._binary.view(-1)[:10] = True
    g.every_n_step(512, dens, warmup_steps=256, n=16, generator=gen)
    assert seen == [8 ** 3 // 4 + 10]                # N/4 uniform cells + every occupied cell (fewer than N/4)
    assert int((g.occs > 0).sum()) <= seen[0]


def test_eval_mode_update_raises_and_state_dict_keys():
    g = _grid("sphere", res=4)
    assert sorted(g.state_dict()) == ["_binary", "_roi_aabb", "occs", "resolution"]
    g.eval()
    with pytest.raises(RuntimeError):
        g.every_n_step(0, lambda x: torch.zeros(x.shape[0]))
    # sphere grids never query the corners of the unit cube (outside the inscribed ball)
    g.train()
    pts = []
    g.every_n_step(0, lambda x: (pts.append(x), torch.ones(x.shape[0]))[1], generator=torch.Generator().manual_seed(2))
    assert pts[0].shape[0] < 4 ** 3 and bool(torch.isfinite(pts[0]).all())
    u8 = g.binary_u8()
    assert u8.dtype == torch.uint8 and u8.shape == (4, 4, 4) and g.binary_u8() is u8      # cached until _binary changes
