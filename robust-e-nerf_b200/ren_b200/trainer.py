"""Minimal optimisation loop for `RobustENeRF` when Lightning is absent (bench / smoke /
multi-GPU runs): the reference's automatic-optimisation step
(`scripts/run.py:88-111` -> PL -> `training_step` -> `backward` -> DDP all-reduce ->
`Adam.step`, robust_e_nerf/models/robust_e_nerf.py:782-832) with the data-parallel
exchange reduced to ONE NCCL all-reduce of one flat fp32 vector per step
 and Adam on the flat MLP buffer as one kernel (`ren_adam_step`).

Layout of the all-reduced vector `gbuf` (fp32):
    [ MLP gradients (n_mlp) | small fp32 parameters | (hi, lo) fp32 pairs of the f64 parameters | mean spp ]
* the refractory period is an f64 parameter (event_generation_params.py:126): its gradient travels as an
  fp32 (hi, lo) pair and is re-assembled in f64 after the sum, so no precision is lost to the fp32 bucket;
* the last slot carries this rank's mean number of samples per ray: after the sum it is the cross-rank mean the
  reference obtains with a second collective (`all_gather`, models/robust_e_nerf.py:916-919) for the dynamic
  batch size.

`cuda_graph=True` captures the whole static-shape step (training_step, backward, all-reduce, both Adam
updates) in one CUDA graph per batch size and replays it: the ~70 small ATen launches of the sensor
parametrisations, loss weighting and logging scalars stop costing host launch latency.  Steps whose shapes
depend on device data (occupancy grid with the packed march) cannot be captured and run eagerly.
"""
import torch

from . import _lib


class FlatTrainer:
    def __init__(self, model, world_size=1, use_fused_adam=True, cuda_graph=False):
        self.model = model
        self.world = world_size
        cfg = model.configure_optimizers()
        opt = cfg["optimizer"]
        self.scheduler = cfg["lr_scheduler"]["scheduler"]
        self.scheduler_interval = cfg["lr_scheduler"].get("interval", "epoch")
        self.torch_opt = opt
        self.mlp_group = opt.param_groups[2]            # configure_optimizers group order
        self.mlp_params = model.nerf.mlp_parameters()
        self.small_params = [p for i, g in enumerate(opt.param_groups) if i != 2 for p in g["params"]
                             if p.requires_grad]
        # the fused Adam runs on the flat buffer the parameters are views of; under weight_norm they are separate
        # (weight_g, weight_v, bias) leaves and torch's Adam takes them
        self.use_fused_adam = use_fused_adam and all(p.requires_grad for p in self.mlp_params) \
            and not getattr(model.nerf, "weight_norm", False)
        flat = model.nerf.flat_params()
        self.n_mlp = sum(p.numel() for p in self.mlp_params)
        # small parameters: fp32 ones take one slot per element, f64 ones a (hi, lo) pair
        self._small_slots = []
        off = self.n_mlp
        for p in self.small_params:
            width = 2 if p.dtype == torch.float64 else 1
            self._small_slots.append((off, p.numel(), width))
            off += p.numel() * width
        self.spp_slot = off
        self.gbuf = torch.zeros(off + 1, device=flat.device, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(flat) if self.use_fused_adam else None
        self.exp_avg_sq = torch.zeros_like(flat) if self.use_fused_adam else None
        self.step_count = 0
        self.epoch = 0
        self.mean_spp = None                            # device scalar: cross-rank mean samples per ray
        if self.use_fused_adam:
            # the torch optimiser keeps the small (scalar) groups only
            self.mlp_group["params"] = []
            # Adam's step count lives on the device so that a captured graph advances it
            self._dev_step = torch.zeros((), device=flat.device, dtype=torch.int64)
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        self._eager_steps = 0
        self.stream = None
        if self.cuda_graph:
            for g in opt.param_groups:                  # torch Adam inside a graph needs capturable state
                g["capturable"] = True
            # every step -- eager warm-up and capture alike -- runs on ONE side stream: autograd's AccumulateGrad
            # nodes are tied to the stream they were created on, and a capture must not make the legacy default
            # stream depend on the capturing stream (cudaErrorStreamCaptureImplicit)
            self.stream = torch.cuda.Stream(device=flat.device)

    # ------------------------------------------------------------------ gradient vector
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
        flat = self._flat_mlp_grad()
        if flat is not None:
            self.gbuf[:self.n_mlp].copy_(flat)
        else:
            off = 0
            for p in self.mlp_params:
                n = p.numel()
                if p.grad is None:
                    self.gbuf[off:off + n].zero_()
                else:
                    self.gbuf[off:off + n].copy_(p.grad.reshape(-1).float())
                off += n
        for p, (off, n, width) in zip(self.small_params, self._small_slots):
            dst = self.gbuf[off:off + n * width]
            if p.grad is None:
                dst.zero_()
            elif width == 1:
                dst.copy_(p.grad.reshape(-1).float())
            else:
                g = p.grad.reshape(-1).double()
                hi = g.float()
                dst[:n].copy_(hi)
                dst[n:].copy_((g - hi.double()).float())
        spp = self.model.logged.get("train/mean_num_samples_per_ray") if hasattr(self.model, "logged") else None
        if torch.is_tensor(spp):
            self.gbuf[self.spp_slot].copy_(spp.detach().float())
        else:
            self.gbuf[self.spp_slot] = float(spp) if spp is not None else 0.0

    def _scatter_small(self, scale):
        for p, (off, n, width) in zip(self.small_params, self._small_slots):
            if width == 1:
                g = self.gbuf[off:off + n] * scale
            else:
                g = (self.gbuf[off:off + n].double() + self.gbuf[off + n:off + 2 * n].double()) * scale
            p.grad = g.reshape(p.shape).to(p.dtype)

    # ------------------------------------------------------------------ one step
    def _step_body(self, batch, batch_index, jitter):
        m = self.model
        for p in self.mlp_params + self.small_params:
            p.grad = None
        loss = m.training_step(batch, batch_index, jitter)
        loss.backward()
        self._gather_grads()
        scale = 1.0
        if self.world > 1:
            torch.distributed.all_reduce(self.gbuf)       # THE collective of the step: DDP's gradient mean
            scale = 1.0 / self.world
        self.mean_spp = self.gbuf[self.spp_slot] * scale
        flat = m.nerf.flat_params()
        if self.use_fused_adam:
            g = self.mlp_group
            b1, b2 = g["betas"]
            self._dev_step += 1
            _lib.check(_lib.load().ren_adam_step_dev(
                flat.numel(), _lib.ptr(flat), _lib.ptr(self.gbuf), _lib.ptr(self.exp_avg),
                _lib.ptr(self.exp_avg_sq), float(g["lr"]), float(b1), float(b2), float(g["eps"]),
                float(g["weight_decay"]), float(scale), _lib.ptr(self._dev_step), _lib.stream_ptr()),
                "ren_adam_step_dev")
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
        return loss

    @staticmethod
    def _signature(batch, jitter):
        sig = []
        for grp in (batch["event"], batch["normalized"], jitter or {}):
            for k in sorted(grp):
                v = grp[k]
                sig.append((k, tuple(v.shape), v.dtype))
        return tuple(sig)

    def _graphable(self):
        nerf = self.model.nerf
        return getattr(nerf, "occupancy_grid", None) is None

    def step(self, batch, batch_index=0, jitter=None):
        m = self.model
        if self.stream is not None:
            caller = torch.cuda.current_stream()
            self.stream.wait_stream(caller)             # the batch was produced on the caller's stream
            with torch.cuda.stream(self.stream):
                loss = self._dispatch(batch, batch_index, jitter)
            caller.wait_stream(self.stream)             # loss / parameters are visible to the caller again
        else:
            loss = self._dispatch(batch, batch_index, jitter)
        self.step_count += 1
        if self.scheduler_interval == "step":
            self.scheduler.step()
        try:
            m.global_step = self.step_count
        except AttributeError:      # Lightning: read-only, owned by the real Trainer
            pass
        return loss

    def _dispatch(self, batch, batch_index, jitter):
        if self.cuda_graph and jitter is not None and self._graphable() and self._eager_steps >= 3:
            return self._graph_step(batch, batch_index, jitter)
        self._eager_steps += 1
        return self._step_body(batch, batch_index, jitter)

    def _graph_step(self, batch, batch_index, jitter):
        # learning rates are baked into the captured launches: a scheduler step re-captures
        key = (self._signature(batch, jitter), tuple(float(g["lr"]) for g in self.torch_opt.param_groups))
        entry = self._graphs.get(key)
        if entry is None:
            static_b = {g: {k: v.clone() for k, v in grp.items()} for g, grp in batch.items()}
            static_j = {k: v.clone() for k, v in jitter.items()}
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=self.stream):
                loss = self._step_body(static_b, batch_index, static_j)
            entry = (graph, static_b, static_j, loss)
            self._graphs[key] = entry
            # capture does not execute: fall through to the first replay with this batch
        graph, static_b, static_j, loss = entry
        for g, grp in batch.items():
            for k, v in grp.items():
                static_b[g][k].copy_(v, non_blocking=True)
        for k, v in jitter.items():
            static_j[k].copy_(v, non_blocking=True)
        graph.replay()
        return loss

    # ------------------------------------------------------------------ epoch / dynamic batch
    def epoch_end(self):
        """`interval: epoch` learning-rate schedule (configs/train/*.yaml:121-128: MultiStepLR, milestones
        20 / 30 / 36 epochs, gamma 0.33) -- Lightning steps it at every training epoch end."""
        self.epoch += 1
        if self.scheduler_interval == "epoch":
            self.scheduler.step()
        try:
            self.model.current_epoch = self.epoch
        except AttributeError:
            pass

    def new_batch_size(self):
        """models/robust_e_nerf.py:907-950: rays per rank for the next step so that rays x mean samples per ray
        stays at the configured budget; the cross-rank mean rode in the gradient all-reduce.  One host read."""
        if self.mean_spp is None:
            return None
        m = float(self.mean_spp)
        if not m > 0:
            return None
        return int(self.model.train_ray_sample_batch_size / m)
