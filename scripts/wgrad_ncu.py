"""The wide-layer weight gradients (wgrad_band_kernel) of one generator forward for `ncu --set full -k
regex:wgrad_band_kernel`: one warm-up + one measured launch per class (the measured launches are the odd ones)."""
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import ops  # noqa: E402

torch.manual_seed(0)
for cin, cout, h, ks in ((16, 16, 256, 3), (16, 32, 128, 3), (32, 32, 128, 3), (16, 16, 256, 5), (32, 16, 256, 1)):
    x = torch.randn(16, h, h, cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(16, h, h, cout, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, ks, ks, device="cuda") * 0.05
    pw = ops.PackedWeight(w)
    ops.PackTable([pw]).refresh()
    for _ in range(2):
        ops.conv_wgrad([x], dy, pw)
    torch.cuda.synchronize()
print("ok")
