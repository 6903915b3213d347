"""Graph-replay timing of the UGANConsisTrainer step (development aid): prints ms/step."""
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200.data_loader import syntheticLoader as synlod  # noqa: E402
from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer  # noqa: E402

torch.manual_seed(0)
tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=256))
lb = synlod.get_loader(None, 'train', 0, 8, pool_batches=2)
ul = synlod.get_loader(None, 'val', 0, 8, pool_batches=2)
(x1, y, m1, _), (x2, _, m2, _) = next(iter(lb)), next(iter(ul))
batch = tr.prepare_batch(x1, y, m1, x2, m2, 2)
a, ids = tr.draw(16)
lam = torch.full((1,), 5.0, device="cuda")
step = tr.graphed_step([*batch, a, ids[0], lam], use_semi=True)
flush = torch.empty(160 * 2 ** 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    step(*batch, a, ids[0], lam)
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tot = 0.0
for _ in range(n):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = step(*batch, a, ids[0], lam)
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"graph step: {tot / n:.3f} ms  ({16 / (tot / n) * 1e3:.1f} slices/s)  launches {step.launches_per_replay}", flush=True)
