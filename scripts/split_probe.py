"""Development: is the staged generator backward (SMSUT_SPLIT_G_BACKWARD) the same computation as the single
g_loss.backward()?  Runs the same UGANConsisTrainer iteration from identical weights R times per mode and prints
the relative distance of G's flat gradient between runs of one mode (run-to-run spread: fp32 atomics order) and
between the modes, plus the distance of the ten losses."""
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from oracle import smsut_oracle as O  # noqa: E402
from smsut_b200.trainer import uganConsisTrainer as T  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 3
size, bs = 256, 4
Gw = {k: v.cuda() for k, v in O.make_weights(O.ugan_shapes(), 7).items()}
Gw["tsl_decoder.fc.weight"] *= 0.05      # keep tanh out of saturation; with lambda_gp = 0 below the step is well conditioned,
# so run-to-run spread (atomics order) is small and a real ordering bug between the stages would stand out
Dw = {k: v.cuda() for k, v in O.make_weights(O.disc_shapes(size), 8).items()}
x1, y = O.synthetic_batch(bs, size, 11)
x2, _ = O.synthetic_batch(bs, size, 12)
m1, m2 = torch.full((bs,), 1), torch.full((bs,), 3)
gen = torch.Generator().manual_seed(3)
alpha = torch.randn(2 * bs, generator=gen).cuda()
ids = [torch.randperm(256, generator=gen)[:64].cuda()]


def run(split):
    T.SPLIT_G_BACKWARD[0] = split
    tr = T.UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=size))
    tr.net.load_state_dict(Gw)
    tr.D.load_state_dict(Dw)
    tr.lambda_gp = 0.0
    batch = tr.prepare_batch(x1, y, m1, x2, m2, 2)
    losses = tr.train_step(*batch, alpha, ids, 0.7, True)
    torch.cuda.synchronize()
    return tr.optimizer.grad.clone(), tr.d_optimizer.grad.clone(), losses.clone()


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


res = {s: [run(s) for _ in range(R)] for s in (False, True)}
for s in (False, True):
    r = res[s]
    print(f"split={s}: run-to-run  G grad rel {max(rel(r[i][0], r[0][0]) for i in range(1, R)):.3e}  "
          f"D grad rel {max(rel(r[i][1], r[0][1]) for i in range(1, R)):.3e}  losses rel {max(rel(r[i][2], r[0][2]) for i in range(1, R)):.3e}")
a, b = res[False], res[True]
print(f"between modes: G grad rel {max(rel(b[i][0], a[j][0]) for i in range(R) for j in range(R)):.3e} (min "
      f"{min(rel(b[i][0], a[j][0]) for i in range(R) for j in range(R)):.3e})  D grad rel {rel(b[0][1], a[0][1]):.3e}  "
      f"losses rel {rel(b[0][2], a[0][2]):.3e}")
