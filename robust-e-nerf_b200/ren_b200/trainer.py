"""Minimal optimisation loop for `RobustENeRF` when Lightning is absent (bench / smoke /
multi-GPU runs): the reference's automatic-optimisation step
(`scripts/run.py:88-111` -> PL -> `training_step` -> `backward` -> DDP all-reduce ->
`Adam.step`, robust_e_nerf/models/robust_e_nerf.py:782-832) with the data-parallel
exchange reduced to ONE NCCL all-reduce of one flat fp32 gradient vector per step
 and Adam on the flat MLP buffer as one kernel (`ren_adam_step`).
"""
import torch

from . import _lib


class FlatTrainer:
    def __init__(self, model, world_size=1, use_fused_adam=True):
        self.model = model
        self.world = world_size
        cfg = model.configure_optimizers()
        opt = cfg["optimizer"]
        self.scheduler = cfg["lr_scheduler"]["scheduler"]
        self.torch_opt = opt
        self.mlp_group = opt.param_groups[2]            # configure_optimizers group order
        self.mlp_params = model.nerf.mlp_parameters()
        self.small_params = [p for i, g in enumerate(opt.param_groups) if i != 2 for p in g["params"]
                             if p.requires_grad]
        self.use_fused_adam = use_fused_adam and all(p.requires_grad for p in self.mlp_params)
        flat = model.nerf.flat_params()
        n_small = sum(p.numel() for p in self.small_params)
        self.gbuf = torch.zeros(flat.numel() + n_small, device=flat.device, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = 0
        if self.use_fused_adam:
            # the torch optimiser keeps the small (scalar) groups only
            self.mlp_group["params"] = []

    def _flat_mlp_grad(self):
        """The fused backward hands out the MLP gradients as views of ONE flat buffer in parameter order
        (`NeRF._split_flat`); autograd keeps those views as `.grad`.  Returns that buffer as a (n,) view when
        this is still the case (one copy instead of one per parameter), else None."""
        g0 = self.mlp_params[0].grad
        if g0 is None or g0.dtype != torch.float32:
            return None
        base, off = g0.untyped_storage().data_ptr(), g0.storage_offset()
        for p in self.mlp_params:
            g = p.grad
            if (g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.storage_offset() != off
                    or g.untyped_storage().data_ptr() != base):
                return None
            off += g.numel()
        n = off - g0.storage_offset()
        return torch.empty(0, dtype=torch.float32, device=g0.device).set_(
            g0.untyped_storage(), g0.storage_offset(), (n,), (1,))

    def _gather_grads(self):
        off = 0
        params = self.mlp_params + self.small_params
        flat = self._flat_mlp_grad()
        if flat is not None:
            self.gbuf[:flat.numel()].copy_(flat)
            off = flat.numel()
            params = self.small_params
        for p in params:
            n = p.numel()
            if p.grad is None:
                self.gbuf[off:off + n].zero_()
            else:
                self.gbuf[off:off + n].copy_(p.grad.reshape(-1).float())
            off += n

    def _scatter_small(self, scale):
        off = sum(p.numel() for p in self.mlp_params)
        for p in self.small_params:
            n = p.numel()
            p.grad = (self.gbuf[off:off + n] * scale).reshape(p.shape).to(p.dtype)
            off += n

    def step(self, batch, batch_index=0, jitter=None):
        m = self.model
        for p in self.mlp_params + self.small_params:
            p.grad = None
        loss = m.training_step(batch, batch_index, jitter)
        loss.backward()
        self._gather_grads()
        scale = 1.0
        if self.world > 1:
            torch.distributed.all_reduce(self.gbuf)       # one collective: DDP's gradient mean
            scale = 1.0 / self.world
        self.step_count += 1
        flat = m.nerf.flat_params()
        if self.use_fused_adam:
            g = self.mlp_group
            b1, b2 = g["betas"]
            _lib.check(_lib.load().ren_adam_step(
                flat.numel(), _lib.ptr(flat), _lib.ptr(self.gbuf), _lib.ptr(self.exp_avg),
                _lib.ptr(self.exp_avg_sq), float(g["lr"]), float(b1), float(b2), float(g["eps"]),
                float(g["weight_decay"]), float(scale), self.step_count, _lib.stream_ptr()),
                "ren_adam_step")
            if self.small_params:
                self._scatter_small(scale)
                self.torch_opt.step()
        else:
            off = 0
            for p in self.mlp_params:
                n = p.numel()
                p.grad = (self.gbuf[off:off + n] * scale).reshape(p.shape)
                off += n
            self._scatter_small(scale)
            self.torch_opt.step()
        try:
            m.global_step = self.step_count
        except AttributeError:      # Lightning: read-only, owned by the real Trainer
            pass
        return loss
