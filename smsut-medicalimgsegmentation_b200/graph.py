"""CUDA-graph capture of a whole training iteration.

The reference issues ~9 300 aten calls and 11 host syncs per UGANConsisTrainer iteration;
this path launches ~1 400 kernels with no sync, and captures forward + double backward + both optimizer steps +
the LR tick into ONE graph, so the host cost of an iteration is a handful of async copies and one graph launch.
The iteration's inputs (slices, labels, modality vectors, alpha, patch ids, lambda_semi) live in static device
buffers that are refreshed before each replay.
"""
import os

import torch

from . import _lib


class GraphedStep:
    """step_fn(*tensors) -> tensor(s); all arguments must be device tensors (they become the static inputs)."""

    def __init__(self, step_fn, example_inputs, warmup=3, refresh=None):
        self.step_fn = step_fn
        self.static_in = [t.clone() if isinstance(t, torch.Tensor) else t for t in example_inputs]
        # the capture stream carries the critical path: same (high) priority as the branch streams, above the wgrad
        # side streams
        from . import ops
        s = torch.cuda.Stream(priority=ops._BRANCH_PRIORITY)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                if refresh is not None:
                    refresh(self.static_in)
                step_fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        dot = os.environ.get("SMSUT_GRAPH_DOT")      # development: dump the captured graph (nodes + dependency edges)
        if dot:
            self.graph.enable_debug_mode()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph, stream=s):
            self.static_out = step_fn(*self.static_in)
        self.launches_per_replay = _lib.launch_count() - before
        if dot:
            self.graph.debug_dump(dot)

    def accepts(self, inputs):
        """True if `inputs` have the shapes / dtypes the graph was captured for (a ragged last batch does not)"""
        if len(inputs) != len(self.static_in):
            return False
        for dst, src in zip(self.static_in, inputs):
            if isinstance(dst, torch.Tensor):
                if not isinstance(src, torch.Tensor) or src.shape != dst.shape or src.dtype != dst.dtype:
                    return False
            elif dst != src:
                return False
        return True

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if isinstance(dst, torch.Tensor) and src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays = getattr(self, 'replays', 0) + 1
        # the replay ended with optimizer steps the host-side bookkeeping has not seen: mark the bf16 weight copies
        # stale so that an eager forward after it (validation, sampling) repacks them
        from . import ops
        ops.param_generation[0] += 1
        return self.static_out


class StateSnapshot:
    """Clones of a trainer's live state tensors (weights, optimizer moments, schedules, running statistics); restore()
    copies them back IN PLACE, so buffers captured in a CUDA graph keep their addresses.  Capturing a step inside a
    training run costs warm-up iterations on the example batch: the snapshot makes the capture side-effect free."""

    def __init__(self, tensors):
        seen, self.pairs = set(), []
        for t in tensors:
            if not isinstance(t, torch.Tensor) or t.numel() == 0:
                continue
            key = (t.data_ptr(), t.numel(), t.dtype)
            if key in seen:
                continue
            seen.add(key)
            self.pairs.append((t, t.detach().clone()))

    def restore(self):
        with torch.no_grad():
            for t, c in self.pairs:
                t.copy_(c)
