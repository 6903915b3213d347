"""Run the same work twice and report how far the results move (kernels without atomics must be bit-identical)."""
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from oracle import smsut_oracle as O  # noqa: E402
from smsut_b200 import ops  # noqa: E402
from smsut_b200.network.ugan import Discriminator, UGANnce  # noqa: E402
from smsut_b200.trainer.uganShp0Trainer import UGANShp0Trainer  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def diff(a, b):
    return (a.float() - b.float()).abs().max().item()


for cins, cout, h in (([16], 16, 256), ([16, 16], 16, 256), ([32], 32, 128), ([64], 64, 64), ([256], 256, 16)):
    xs = [torch.randn(8, h, h, c, device=dev).to(torch.bfloat16) for c in cins]
    dy = torch.randn(8, h, h, cout, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, sum(cins), 3, 3, device=dev) * 0.05
    pw = ops.PackedWeight(w)
    ops.PackTable([pw]).refresh()
    r = []
    for _ in range(3):
        y = ops.conv_fprop(xs, pw)
        dx = ops.conv_dgrad(dy, pw, [x.shape[3] for x in xs])
        dw = ops.conv_wgrad(xs, dy, pw)
        r.append((y, dx[0], dw))
    print("conv", cins, cout, h, "fprop", max(diff(r[0][0], r[i][0]) for i in (1, 2)), "dgrad",
          max(diff(r[0][1], r[i][1]) for i in (1, 2)), "wgrad rel",
          max(diff(r[0][2], r[i][2]) for i in (1, 2)) / r[0][2].abs().max().item())

x = torch.randn(8, 128, 128, 32, device=dev).to(torch.bfloat16)
ga, ba = torch.ones(32, device=dev), torch.zeros(32, device=dev)
st = [ops.in_stats(x) for _ in range(3)]
print("in_stats rel", max(diff(st[0], s) for s in st[1:]) / st[0].abs().max().item())
out = [ops.in_apply(x, st[0], ga, ba, act=ops.ACT_LRELU) for _ in range(3)]
print("in_apply", max(diff(out[0], o) for o in out[1:]))

D = Discriminator(256, 4, 16, max_width=256).to(dev)
D.load_state_dict({k: v.to(dev) for k, v in O.make_weights(O.disc_shapes(256), 8).items()})
img, _ = O.synthetic_batch(8, 256, 6, device=dev)
xh = (img + 0.1 * torch.randn_like(img))
vals = []
for _ in range(3):
    xr = xh.clone().requires_grad_(True)
    src, cls = D(xr)
    gp = UGANShp0Trainer.gradient_penalty(None, src, xr)
    D.zero_grad()
    gp.backward()
    vals.append((src.detach().clone(), gp.item(), D.main[2].conv1.weight.grad.clone()))
print("D forward", max(diff(vals[0][0], v[0]) for v in vals[1:]), "gp", [v[1] for v in vals], "dgrad(gp) rel",
      max(diff(vals[0][2], v[2]) for v in vals[1:]) / vals[0][2].abs().max().item())

G = UGANnce(1, 5, 4, 16).to(dev)
G.load_state_dict({k: v.to(dev) for k, v in O.make_weights(O.ugan_shapes(), 7).items()})
m = torch.tensor([[1., 0, -1, 0]] * 8, device=dev)
ids = [torch.randperm(256, device=dev)[:64]]
outs = [G(img, m, sample_ids=ids) for _ in range(3)]
print("G seg", max(diff(outs[0][0], o[0]) for o in outs[1:]), "tsl", max(diff(outs[0][1], o[1]) for o in outs[1:]),
      "feat", max(diff(outs[0][2][0], o[2][0]) for o in outs[1:]))
