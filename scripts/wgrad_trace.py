"""Development: clock stamps of wgrad_band_kernel (SMSUT_WGRAD_TRACE=1) on one layer class.
Usage: wgrad_trace.py [cin cout h]"""
import os
import sys

os.environ["SMSUT_WGRAD_TRACE"] = "1"
import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import ops  # noqa: E402

cin, cout, h = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 16, 256))]
x = torch.randn(16, h, h, cin, device="cuda").to(torch.bfloat16)
dy = torch.randn(16, h, h, cout, device="cuda").to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
pw = ops.PackedWeight(w)
ops.PackTable([pw]).refresh()
flush = torch.empty(192 * 2 ** 20, dtype=torch.uint8, device="cuda")
for i in range(2):
    flush.zero_()
    torch.cuda.synchronize()
    ops.conv_wgrad([x], dy, pw)
torch.cuda.synchronize()
