"""Eager-mode timing of the UGANConsisTrainer step (development aid; bench.py is the contract)."""
import random
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import _lib  # noqa: E402
from smsut_b200.data_loader import syntheticLoader as synlod  # noqa: E402
from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer  # noqa: E402

torch.manual_seed(0)
tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=256))
lb = synlod.get_loader(None, 'train', 0, 8, pool_batches=2)
ul = synlod.get_loader(None, 'val', 0, 8, pool_batches=2)
(x1, y, m1, _), (x2, _, m2, _) = next(iter(lb)), next(iter(ul))
batch = tr.prepare_batch(x1, y, m1, x2, m2, 2)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for i in range(3):
    a, ids = tr.draw(16)
    losses = tr.train_step(*batch, a, ids, 0.5, True)
torch.cuda.synchronize()
print("warm losses", [round(v, 4) for v in losses.tolist()])
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for i in range(steps):
    a, ids = tr.draw(16)
    losses = tr.train_step(*batch, a, ids, 0.5, True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"eager step: {ms:.2f} ms/step  ({16 / ms * 1e3:.1f} slices/s), wall {1e3 * (time.time() - t0) / steps:.2f} ms, "
      f"{(_lib.launch_count() - l0) / steps:.0f} smsut launches/step, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
print("losses", [round(v, 4) for v in losses.tolist()])
