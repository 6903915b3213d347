"""Data parallelism for the SMSUT step: one process per GPU (torchrun), replicated G / D, each rank running the
per-GPU iteration on its own 8 labelled + 8 unlabelled slices.  The reference uses single-process
nn.DataParallel (trainer/uganShp0Trainer.py:66-68) and computes its losses on the gathered global batch; the
three exchange steps that reproduce that arithmetic are:
  1. sum-all-reduce of D's flat gradient buffer (then 1/W inside the fused Adam kernel),
  2. sum-all-reduce of G's flat gradient buffer (1/W inside the fused SGD kernel) -- in two buckets: the gradients that
     are complete after the generator's early backward stage (segmentation halves + netF, ~45 % of the bytes) are
     reduced on a communication stream WHILE the discriminator phase and the rest of the backward run
     (all_reduce_early), the remainder after the last weight-gradient kernel (all_reduce_grads),
  3. sum-all-reduce of the 3x5 Dice statistics inside the loss (batch-Dice is non-linear in batch-wide sums).
All of them are NCCL collectives on the compute stream (NVLink 5 / NVSwitch), capturable in the step's CUDA graph.
InstanceNorm, PatchNCE groups and the gradient penalty are per-sample: no exchange.
"""
import os

import torch
import torch.distributed as dist

from . import functional as Fn


class DataParallelContext:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.owns_group = False
        self._comm = None          # communication stream of all_reduce_early
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend, rank=self.rank, world_size=self.world, **kw)
            self.owns_group = True
        if self.world > 1:
            Fn.set_data_parallel(self.all_reduce_sum, self.world)

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def all_reduce_grads(self, optimizer):
        """sum the flat gradient buffer over ranks; the fused optimizer kernel applies 1/world.  If the early bucket
        of this iteration was already reduced (all_reduce_early) only the rest is, and the current stream is ordered
        behind the communication stream."""
        if self.world > 1:
            optimizer.finish_grads()
            if getattr(optimizer, 'early_done', False):
                if optimizer.early_offset > 0:
                    dist.all_reduce(optimizer.grad[:optimizer.early_offset], op=dist.ReduceOp.SUM)
                if self._comm is not None:
                    torch.cuda.current_stream().wait_stream(self._comm)
            else:
                dist.all_reduce(optimizer.grad, op=dist.ReduceOp.SUM)
            optimizer.grad_scale = 1.0 / self.world

    def all_reduce_early(self, optimizer, producers=()):
        """Overlap: complete and sum-all-reduce the optimizer's EARLY bucket (optim._FlatOptimizer(early=...)) on a
        communication stream that waits only for `producers` -- the streams whose kernels write those gradients -- so
        the collective runs beside whatever the calling stream does next (NCCL over NVLink, captured in the step's
        graph as a parallel branch).  all_reduce_grads() later reduces the rest and joins the communication stream."""
        if self.world <= 1 or not getattr(optimizer, 'n_early', 0):
            return False
        if optimizer.grad.is_cuda:
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=optimizer.grad.device, priority=-1)
            cur = torch.cuda.current_stream()
            self._comm.wait_stream(cur)
            for s in producers:
                self._comm.wait_stream(s)
            with torch.cuda.stream(self._comm):
                if optimizer.finish_early_grads():
                    dist.all_reduce(optimizer.grad[optimizer.early_offset:], op=dist.ReduceOp.SUM)
        elif optimizer.finish_early_grads():         # CPU test double (gloo): same arithmetic, no streams
            dist.all_reduce(optimizer.grad[optimizer.early_offset:], op=dist.ReduceOp.SUM)
        return True

    def broadcast_params(self, *optimizers):
        """replicas start from rank 0's weights (nn.DataParallel's replicate())"""
        if self.world > 1:
            for o in optimizers:
                dist.broadcast(o.flat, src=0)

    def max_over_ranks(self, value):
        if self.world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier(self):
        if self.world > 1:
            dist.barrier()

    def close(self):
        Fn.set_data_parallel(None, 1)
        if self.owns_group and dist.is_initialized():
            dist.destroy_process_group()
