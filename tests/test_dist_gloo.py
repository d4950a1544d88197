"""world_size-2 `gloo` test (CPU) of the data-parallel host logic: per-rank event batches, ONE
all-reduce of the flat gradient vector per step, identical parameters on every rank afterwards and
equal to single-process training on the mean gradient (DDP semantics of scripts/run.py:81-86)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _ToyNeRF(torch.nn.Module):
    """Stands in for ren_b200.NeRF on CPU: flat parameter buffer + named views."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self._flat = torch.randn(12 + 4)
        self.w = torch.nn.Parameter(self._flat[:12].view(4, 3))
        self.b = torch.nn.Parameter(self._flat[12:])

    def flat_params(self):
        return self._flat

    def mlp_parameters(self):
        return [self.w, self.b]


class _ToyModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.nerf = _ToyNeRF()
        self.scale = torch.nn.Parameter(torch.tensor(1.0))     # a "small" non-MLP parameter
        # an f64 parameter like the refractory period (event_generation_params.py:126): its gradient must
        # survive the fp32 all-reduce bucket (hi / lo pair)
        self.tau = torch.nn.Parameter(torch.tensor(1.0 + 1e-9, dtype=torch.float64))
        self.global_step = 0
        self.logged = {}

    def configure_optimizers(self):
        opt = torch.optim.Adam([{"params": [], "lr": 0.1}, {"params": [self.tau], "lr": 1.0},
                                {"params": self.nerf.mlp_parameters(), "weight_decay": 1e-6},
                                {"params": [self.scale]}], lr=0.01)
        sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[20], gamma=0.33)
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sch, "interval": "epoch"}}

    def training_step(self, batch, batch_index, jitter=None):
        x, y = batch
        pred = torch.nn.functional.linear(x, self.nerf.w, self.nerf.b) * self.scale
        self.logged["train/mean_num_samples_per_ray"] = torch.tensor(100.0 + float(x[0, 0]))
        return ((pred.double() * self.tau - y) ** 2).mean()


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)        # per-rank seed = seed + rank (datamodule.py:85-89)
    return torch.randn(16, 3, generator=g), torch.randn(16, 4, generator=g)


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "robust-e-nerf_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ren_b200.trainer import FlatTrainer
    model = _ToyModule()
    tr = FlatTrainer(model, world_size=world, use_fused_adam=False)
    for step in range(3):
        tr.step(_data(rank), step)
    # the dynamic-batch scalar rode in the same all-reduce: cross-rank mean of the per-rank values
    want = sum(100.0 + float(_data(r)[0][0, 0]) for r in range(world)) / world
    assert abs(float(tr.mean_spp) - want) < 1e-4
    tr.epoch_end()
    flat = torch.cat([p.detach().reshape(-1).double() for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save(gathered, out)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2(tmp_path, built_lib):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "params.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    assert torch.equal(got[0], got[1]), "ranks diverged"
    # single-process reference: average the two ranks' gradients by hand
    model = _ToyModule()
    opt = model.configure_optimizers()["optimizer"]
    for step in range(3):
        grads = []
        for rank in range(2):
            for p in model.parameters():
                p.grad = None
            model.training_step(_data(rank), step).backward()
            grads.append([p.grad.clone() for p in model.parameters()])
        for p, g0, g1 in zip(model.parameters(), grads[0], grads[1]):
            p.grad = (g0 + g1) / 2
        opt.step()
    ref = torch.cat([p.detach().reshape(-1).double() for p in model.parameters()])
    assert torch.allclose(got[0], ref, rtol=1e-5, atol=1e-6)


def test_f64_gradient_survives_the_fp32_bucket(built_lib):
    """The refractory period's f64 gradient crosses the fp32 all-reduce vector as a (hi, lo) pair."""
    sys.path.insert(0, os.path.join(ROOT, "robust-e-nerf_b200"))
    from ren_b200.trainer import FlatTrainer
    model = _ToyModule()
    tr = FlatTrainer(model, world_size=1, use_fused_adam=False)
    g = 1.2345678901234567e-7
    model.tau.grad = torch.tensor(g, dtype=torch.float64)
    model.scale.grad = torch.tensor(0.5)
    for p in model.nerf.mlp_parameters():
        p.grad = torch.zeros_like(p)
    tr._gather_grads()
    model.tau.grad = None
    tr._scatter_small(0.5)
    assert model.tau.grad.dtype == torch.float64
    assert abs(float(model.tau.grad) - 0.5 * g) <= 1e-13 * g          # fp32 alone would be off by ~3e-8 * g
    assert float(model.scale.grad) == 0.25
