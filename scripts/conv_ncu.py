"""The 13 3x3 conv classes of one generator forward (16 slices), one warm-up + one measured launch each, for
`ncu --set full -k regex:conv_(band|tc)_kernel`: the measured launches are the odd ones (1, 3, 5 ...)."""
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
import bench  # noqa: E402
from smsut_b200 import ops  # noqa: E402

torch.manual_seed(0)
for cins, cout, h, ks, count in bench.CONV_CLASSES:
    xs = [torch.randn(16, h, h, c, device="cuda").to(torch.bfloat16) for c in cins]
    w = torch.randn(cout, sum(cins), ks, ks, device="cuda") * 0.05
    pw = ops.PackedWeight(w)
    ops.PackTable([pw]).refresh()
    for _ in range(2):
        ops.conv_fprop(xs, pw, want_stats=True)
    torch.cuda.synchronize()
print("ok")
