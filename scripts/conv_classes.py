"""Development: time fprop / dgrad / wgrad of the 13 3x3 conv classes of one generator forward (16 slices), each launch
alone after an L2 flush (CUDA events).  Usage: conv_classes.py [reps]"""
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import ops  # noqa: E402

CLASSES = [([16], 16, 256), ([16, 16], 16, 256), ([16], 32, 128), ([32], 32, 128), ([32, 32], 32, 128), ([32], 64, 64),
           ([64], 64, 64), ([64, 64], 64, 64), ([64], 128, 32), ([128], 128, 32), ([128, 128], 128, 32), ([128], 256, 16),
           ([256], 256, 16)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
flush = torch.empty(192 * 2 ** 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    fn()
    ms = 0.0
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(300000)      # keep the GPU busy while the host queues the launch: no launch latency in the timing
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps * 1e3


tot = [0.0, 0.0, 0.0]
print(f"{'class':>18} | {'fprop':>7} {'+stats':>7} {'dgrad':>7} {'wgrad':>7}  us   (TFLOP/s fprop)")
for cins, cout, h in CLASSES:
    xs = [torch.randn(16, h, h, c, device="cuda").to(torch.bfloat16) for c in cins]
    dy = torch.randn(16, h, h, cout, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, sum(cins), 3, 3, device="cuda") * 0.05
    pw = ops.PackedWeight(w)
    ops.PackTable([pw]).refresh()
    dw = torch.zeros_like(w)
    t_f = timed(lambda: ops.conv_fprop(xs, pw))
    t_s = timed(lambda: ops.conv_fprop(xs, pw, want_stats=True))
    t_d = timed(lambda: ops.conv_dgrad(dy, pw, [x.shape[3] for x in xs]))
    t_w = timed(lambda: ops.conv_wgrad(xs, dy, pw, out=dw))
    flop = 2.0 * 16 * h * h * cout * sum(cins) * 9
    print(f"{sum(cins):4d}->{cout:3d} @{h:3d} x{len(cins)} | {t_f:7.1f} {t_s:7.1f} {t_d:7.1f} {t_w:7.1f}       ({flop / t_f / 1e6:6.1f})",
          flush=True)
    tot[0] += t_f
    tot[1] += t_d
    tot[2] += t_w
print("sum us: fprop %.0f dgrad %.0f wgrad %.0f" % tuple(tot))
