"""Hardware experiment: does a UMMA K-major smem descriptor whose start address is shifted by ONE row inside a
swizzled tile read the right data?  (Decides whether a halo tile loaded once can serve all 3x3 taps.)"""
import os
import subprocess
import sys

CODE = """
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
import __graft_entry__ as g; g.load_package()
from smsut_b200 import ops
torch.manual_seed(0)
for c in (16, 32, 64):
    x = torch.randn(2, 8, 256, c, device='cuda').to(torch.bfloat16)     # W = 256 -> tile = 128 pixels of one row
    w = torch.randn(c, c, 3, 3, device='cuda') * 0.1
    pw = ops.PackedWeight(w); ops.PackTable([pw]).refresh()
    y = ops.conv_fprop([x], pw).float()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    keep = torch.ones(256, dtype=torch.bool, device='cuda'); keep[127] = False; keep[255] = False
    d = (y - ref)[:, :, keep, :]
    print('c', c, 'rel err (excluding the last pixel of each tile):', (d.norm() / ref[:, :, keep, :].norm()).item())
"""
for mode in ("0", "1", "2"):
    env = dict(os.environ, SMSUT_DEBUG_ROWSHIFT=mode)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    print(f"== SMSUT_DEBUG_ROWSHIFT={mode} exit {r.returncode}")
    print(r.stdout.strip())
    if r.returncode:
        print(r.stderr.strip()[-600:])
