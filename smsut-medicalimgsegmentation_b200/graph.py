"""CUDA-graph capture of a whole training iteration.

The reference issues ~9 300 aten calls and 11 host syncs per UGANConsisTrainer iteration;
this path launches ~1 400 kernels with no sync, and captures forward + double backward + both optimizer steps +
the LR tick into ONE graph, so the host cost of an iteration is a handful of async copies and one graph launch.
The iteration's inputs (slices, labels, modality vectors, alpha, patch ids, lambda_semi) live in static device
buffers that are refreshed before each replay.
"""
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
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph, stream=s):
            self.static_out = step_fn(*self.static_in)
        self.launches_per_replay = _lib.launch_count() - before

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if isinstance(dst, torch.Tensor) and src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
