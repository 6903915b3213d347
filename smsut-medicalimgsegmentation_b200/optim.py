"""Fused optimisers with the torch.optim surface the trainers use (param_groups[...]['lr'], step(), zero_grad()).

Parameters and gradients of a network are re-homed into two flat fp32 buffers, so one kernel updates the whole
network (the reference loops over 175 / 46 tensors: ~1300 aten calls per step) and one NCCL all-reduce covers
every gradient.  The learning rate lives in a device scalar so a captured CUDA graph follows the poly schedule.
Reference semantics: torch.optim.SGD(momentum=0.9, weight_decay) / torch.optim.Adam(betas, weight_decay) as
constructed at trainer/uganShp0Trainer.py:72-74 and trainer/unetTrainer.py:47.
"""
import torch

from . import ops


class FlatParams:
    """Re-homes a parameter list into one flat fp32 buffer (no gradients): the EMA teacher's weights."""

    def __init__(self, params):
        self.params = [p for p in params]
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p, sz in zip(self.params, sizes):
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)
                off += sz

    def live_tensors(self):
        return [self.flat]


class _FlatOptimizer:
    def __init__(self, params, lr, early=None):
        """early: optional list of parameters whose gradients are complete EARLY in the iteration (before the rest of the
        backward has run).  They are laid out as the tail of the flat buffers -- [late ..., early ...] -- so that their
        gradients are one contiguous range that a data-parallel run can all-reduce while the remaining backward is
        still running (parallel.DataParallelContext.all_reduce_early)."""
        self.params = [p for p in params]
        if early:
            ids = {id(p) for p in early}
            self.params = [p for p in self.params if id(p) not in ids] + [p for p in self.params if id(p) in ids]
        self.n_early = len(early) if early else 0
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        total = sum(sizes)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        self.early_offset = total      # first element of the early bucket in the flat buffers
        with torch.no_grad():
            for i, (p, sz) in enumerate(zip(self.params, sizes)):
                if self.n_early and i == len(self.params) - self.n_early:
                    self.early_offset = off
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)
                p.grad = self.grad[off:off + n].view(p.shape)
                off += sz
        self.early_done = False        # the early bucket of this iteration has already been completed (and reduced)
        # tap-major scratch for the tensor-core weight gradients (ops.WgradScratch); folded into `grad` by finish_grads()
        # deterministic mode: the parameter gradients the kernels accumulate with atomics (norm gamma / beta, direct
        # convs) go through a fixed-point shadow of the flat gradient, folded in by finish_grads()
        self.grad_shadow = ops.det_register(self.grad) if ops.DET[0] and self.grad.is_cuda else None
        self.wgrad_scratch = ops.WgradScratch(self.params, [p.grad for p in self.params], grad_flat=self.grad,
                                              n_early=self.n_early, early_offset=self.early_offset)
        self.param_groups = [dict(params=self.params, lr=lr)]
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        self.grad_scale = 1.0      # 1/world_size after a summing all-reduce of `grad`

    def _detached(self):
        """parameters whose .grad no longer aliases the flat gradient buffer (net.zero_grad(set_to_none=True),
        optimizer-unaware code assigning p.grad, ...)"""
        lo = self.grad.data_ptr()
        hi = lo + self.grad.numel() * 4
        return [p for p in self.params if p.grad is None or not (lo <= p.grad.data_ptr() < hi)]

    def zero_grad(self, set_to_none=False):
        self.early_done = False
        self.grad.zero_()
        if self.grad_shadow is not None:
            self.grad_shadow.zero_()
        self.wgrad_scratch.discard()     # a backward without a step must not leak into the next one
        if self._detached():             # autograd may have swapped a .grad tensor in: re-attach the flat views
            self._reattach()

    def finish_grads(self):
        """complete the flat gradient: adopt gradients that landed in foreign .grad tensors, fold the tensor-core
        weight-gradient scratch and (deterministic mode) the fixed-point shadow in (no-op when nothing is pending)"""
        stray = self._detached()
        if stray:
            # the fused step reads only the flat buffer: a detached .grad would silently train on zeros
            held = [(p, p.grad) for p in stray]
            self._reattach()
            with torch.no_grad():
                for p, g in held:
                    if g is not None:
                        p.grad.add_(g.to(p.grad.dtype).view_as(p.grad))
        self.wgrad_scratch.flush(part='late' if self.early_done else 'all')

    def finish_early_grads(self):
        """complete the gradients of the early bucket (fold its part of the tensor-core scratch and of the fixed-point
        shadow in); the final finish_grads() then only completes the rest"""
        if not self.n_early or self.early_done:
            return False
        self.wgrad_scratch.flush(part='early')
        self.early_done = True
        return True

    def _reattach(self):
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.grad[off:off + n].view(p.shape)
            off += (n + 3) // 4 * 4

    def live_tensors(self):
        """every device buffer a step mutates (graph.StateSnapshot)"""
        out = [self.flat, self.grad, self.lr_dev] + [getattr(self, k) for k in self._STATE]
        if self.grad_shadow is not None:
            out.append(self.grad_shadow)
        return out

    def _sync_lr(self):
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    # resume state (an extension: the reference saves weights only, trainer/uganShp0Trainer.py:94-107): the flat
    # optimizer buffers in parameter order, copied in place so captured graphs keep their addresses
    _STATE = ()

    def state_dict(self):
        sd = {k: getattr(self, k).detach().cpu().clone() for k in self._STATE}
        sd['lr'] = float(self.lr_dev.item())
        sd['numel'] = int(self.flat.numel())
        return sd

    def load_state_dict(self, sd):
        if int(sd['numel']) != self.flat.numel():
            raise ValueError(f"optimizer state holds {sd['numel']} elements, this network has {self.flat.numel()}")
        with torch.no_grad():
            for k in self._STATE:
                getattr(self, k).copy_(sd[k])
            self.lr_dev.fill_(float(sd['lr']))
        self.param_groups[0]['lr'] = self._lr_host = float(sd['lr'])


class SGD(_FlatOptimizer):
    _STATE = ('mom',)

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, early=None):
        super().__init__(params, lr, early=early)
        self.momentum, self.weight_decay = momentum, weight_decay
        self.mom = torch.zeros_like(self.flat)

    def step(self):
        self.finish_grads()
        self._sync_lr()
        ops.sgd_step(self.flat, self.grad, self.mom, self.lr_dev, self.momentum, self.weight_decay, self.grad_scale)


class Adam(_FlatOptimizer):
    _STATE = ('m', 'v', 'state')

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, lr)
        self.betas, self.eps, self.weight_decay = tuple(betas), eps, weight_decay
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.state = torch.zeros(1, dtype=torch.float32, device=self.flat.device)

    def step(self):
        self.finish_grads()
        self._sync_lr()
        ops.adam_step(self.flat, self.grad, self.m, self.v, self.lr_dev, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.state, self.grad_scale)


class PolyLR:
    """lr_k = base * (1 - (k-1)/max_iter)^0.9 for step k >= 1 (step 0 runs at base): the reference sets the LR
    after step k from iter = k (trainer/uganConsisTrainer.py:198-203).  tick() runs on the device."""

    def __init__(self, optimizers, base_lr, max_iter, power=0.9):
        self.opts, self.base, self.max_iter, self.power = list(optimizers), base_lr, max_iter, power
        dev = self.opts[0].flat.device
        self.iter_state = torch.zeros(1, dtype=torch.float32, device=dev)
        for o in self.opts[1:]:
            o.lr_dev = self.opts[0].lr_dev    # one shared device scalar

    def live_tensors(self):
        return [self.iter_state]

    def tick(self):
        ops.poly_lr_tick(self.iter_state, self.opts[0].lr_dev, self.base, float(self.max_iter), self.power)

    def host_lr(self, it):
        return self.base * (1.0 - max(it - 1, 0) / self.max_iter) ** self.power

    def state_dict(self):
        return dict(iter_state=self.iter_state.detach().cpu().clone())

    def load_state_dict(self, sd):
        with torch.no_grad():
            self.iter_state.copy_(sd['iter_state'])
