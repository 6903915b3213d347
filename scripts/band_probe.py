import sys
import torch
sys.path.insert(0, ".")
import __graft_entry__ as g
g.load_package()
from smsut_b200 import ops
torch.manual_seed(0)
for cins, cout, h in (([16], 16, 256), ([16, 16], 16, 256), ([32], 32, 128)):
    xs = [torch.randn(16, h, h, c, device="cuda").to(torch.bfloat16) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3, device="cuda") * 0.05
    pw = ops.PackedWeight(w)
    ops.PackTable([pw]).refresh()
    for _ in range(2):
        ops.conv_fprop(xs, pw)
torch.cuda.synchronize()
print("ok")
