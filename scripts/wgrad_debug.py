import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as g; g.load_package()
from smsut_b200 import ops
torch.manual_seed(1)
n, c, h = 2, 16, 256
x = torch.randn(n, c, h, h, device='cuda').to(torch.bfloat16).float()
dy = torch.randn(n, c, h, h, device='cuda').to(torch.bfloat16).float()
w = torch.randn(c, c, 3, 3, device='cuda')
pw = ops.PackedWeight(w); ops.PackTable([pw]).refresh()
dw = ops.conv_wgrad([x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)], dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), pw)
ref = torch.nn.grad.conv2d_weight(x, w.shape, dy, padding=1)
for ty in range(3):
    print([round(((dw[:, :, ty, tx] - ref[:, :, ty, tx]).norm() / ref[:, :, ty, tx].norm()).item(), 4) for tx in range(3)])
# does some dw tap equal another ref tap?
for ty in range(3):
    for ty2 in range(3):
        r = ((dw[:, :, ty, 1] - ref[:, :, ty2, 1]).norm() / ref[:, :, ty2, 1].norm()).item()
        if r < 0.05: print("dw ty", ty, "matches ref ty", ty2, round(r, 4))
