"""Two-branch InstanceNorm backward at the 256x256 / 16-channel / 16-slice shape, for an ncu --set full capture."""
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import ops  # noqa: E402

torch.manual_seed(0)
x = torch.randn(16, 256, 256, 16, device="cuda").to(torch.bfloat16)
xb = torch.randn(16, 256, 256, 16, device="cuda").to(torch.bfloat16)
d = torch.randn(16, 256, 256, 16, device="cuda").to(torch.bfloat16)
ga, ba = torch.ones(16, device="cuda"), torch.zeros(16, device="cuda")
for _ in range(2):
    st, stb = ops.in_stats(x), ops.in_stats(xb)
    out = ops.in_apply(x, st, ga, ba, xb, stb, ga, ba, act=ops.ACT_LRELU)
    r = ops.in_bwd(d, out, x, st, ga, xb, stb, ga, act=ops.ACT_LRELU)                   # mask read from `out`
    r = ops.in_bwd(d, out, x, st, ga, xb, stb, ga, act=ops.ACT_LRELU, betas=(ba, ba))   # mask recomputed from x
torch.cuda.synchronize()
print("ok")
