"""SMSUT training-step hot path on B200: hand-written sm_100a kernels behind the reference's nn.Module /
trainer API.  The directory name is not a Python identifier; import it through
`__graft_entry__.load_package()` (registers it as `smsut_b200`) or run the mirrored CLI entry points
(`trainer/uganConsisTrainer.py -p train -f 0`) from inside this directory.
"""
from . import _lib  # noqa: F401  (fails loudly when libsmsut_b200.so is missing)

__all__ = ["_lib"]

# Branch streams (ops.parallel_branch) make some leaf gradients arrive from another stream than the one their
# AccumulateGrad node was created on; the engine synchronises them correctly, the warning is only advice.
try:
    import torch.autograd.graph as _ag
    if hasattr(_ag, "set_warn_on_accumulate_grad_stream_mismatch"):
        _ag.set_warn_on_accumulate_grad_stream_mismatch(False)
except Exception:      # pragma: no cover
    pass
