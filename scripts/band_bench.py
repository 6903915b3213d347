"""Micro-benchmark of the conv kernels on the wide layer classes under a few tuning knobs (development aid)."""
import os
import subprocess
import sys

CODE = """
import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as g; g.load_package()
from smsut_b200 import ops
torch.manual_seed(0)
flush = torch.empty(192 * 2**20, dtype=torch.uint8, device='cuda')
out = []
for cins, cout, h in (([16], 16, 256), ([16, 16], 16, 256), ([32], 32, 128), ([32, 32], 32, 128)):
    xs = [torch.randn(16, h, h, c, device='cuda').to(torch.bfloat16) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3, device='cuda') * 0.05
    pw = ops.PackedWeight(w); ops.PackTable([pw]).refresh()
    ops.conv_fprop(xs, pw)
    ms = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv_fprop(xs, pw); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    out.append(round(ms / 5 * 1e3, 1))
print(out)
"""
SETTINGS = [{}, {"SMSUT_NO_BAND": "1"}, {"SMSUT_BAND_ROWS": "8"}, {"SMSUT_BAND_ROWS": "32"}, {"SMSUT_BAND_ROWS": "64"},
            {"SMSUT_BAND_SLOTS": "5"}, {"SMSUT_BAND_SLOTS": "12"}, {"SMSUT_BAND_ROWS": "32", "SMSUT_BAND_SLOTS": "5"}]
for st in SETTINGS:
    r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, **st), capture_output=True, text=True, timeout=300)
    print(st, "->", r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-300:], flush=True)
