"""A few conv_tc / wgrad_tc launches on representative layer classes, for an ncu --set full capture."""
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import ops  # noqa: E402

torch.manual_seed(0)
for cin, cout, h in ((16, 16, 256), (64, 64, 64), (256, 256, 16)):
    x = torch.randn(16, h, h, cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(16, h, h, cout, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    pw = ops.PackedWeight(w)
    ops.PackTable([pw]).refresh()
    for _ in range(2):
        y = ops.conv_fprop([x], pw)
        dw = ops.conv_wgrad([x], dy, pw)
torch.cuda.synchronize()
print("ok")
