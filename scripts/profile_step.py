"""Two eager UGANConsisTrainer iterations (warm-up + one to profile) for ncu's launch list / --set full capture."""
import sys
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
lb = synlod.get_loader(None, 'train', 0, 8, pool_batches=1)
ul = synlod.get_loader(None, 'val', 0, 8, pool_batches=1)
(x1, y, m1, _), (x2, _, m2, _) = next(iter(lb)), next(iter(ul))
batch = tr.prepare_batch(x1, y, m1, x2, m2, 2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i in range(n):
    l0 = _lib.launch_count()
    a, ids = tr.draw(16)
    losses = tr.train_step(*batch, a, ids, 0.5, True)
    torch.cuda.synchronize()
    print("step", i, "smsut launches", _lib.launch_count() - l0, flush=True)
print([round(v, 4) for v in losses.tolist()])
