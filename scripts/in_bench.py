"""Development: InstanceNorm kernels per level (16 slices), each launch alone after an L2 flush; us and GB/s of the
algorithmic bytes (bf16 tensors read + written)."""
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
flush = torch.empty(192 * 2 ** 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    fn()
    ms = 0.0
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps * 1e3


print(f"{'level':>10} | stats  apply1 apply2  bwd1(red+app)  bwd2(red+app)  bwd1r  bwd2r [us]   | GB/s: apply2 bwd2 bwd2r"
      "   (bwdNr: activation sign recomputed from x instead of read from out)")
for h, c in ((256, 16), (128, 32), (64, 64), (32, 128), (16, 256)):
    x = torch.randn(16, h, h, c, device="cuda").to(torch.bfloat16)
    xb = torch.randn(16, h, h, c, device="cuda").to(torch.bfloat16)
    d = torch.randn(16, h, h, c, device="cuda").to(torch.bfloat16)
    ga, ba = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    st, stb = ops.in_stats(x), ops.in_stats(xb)
    out1 = ops.in_apply(x, st, ga, ba, act=ops.ACT_LRELU)
    out2 = ops.in_apply(x, st, ga, ba, xb, stb, ga, ba, act=ops.ACT_LRELU)
    nb = x.numel() * 2
    t_s = timed(lambda: ops.in_stats(x))
    t_a1 = timed(lambda: ops.in_apply(x, st, ga, ba, act=ops.ACT_LRELU))
    t_a2 = timed(lambda: ops.in_apply(x, st, ga, ba, xb, stb, ga, ba, act=ops.ACT_LRELU))
    t_b1 = timed(lambda: ops.in_bwd(d, out1, x, st, ga, act=ops.ACT_LRELU))
    t_b2 = timed(lambda: ops.in_bwd(d, out2, x, st, ga, xb, stb, ga, act=ops.ACT_LRELU))
    t_b1r = timed(lambda: ops.in_bwd(d, out1, x, st, ga, act=ops.ACT_LRELU, betas=(ba, None)))
    t_b2r = timed(lambda: ops.in_bwd(d, out2, x, st, ga, xb, stb, ga, act=ops.ACT_LRELU, betas=(ba, ba)))
    print(f"{h:4d}x{c:<4d} | {t_s:6.1f} {t_a1:6.1f} {t_a2:6.1f} {t_b1:14.1f} {t_b2:14.1f} {t_b1r:6.1f} {t_b2r:6.1f}       | "
          f"{3 * nb / t_a2 / 1e3:6.0f} {10 * nb / t_b2 / 1e3:6.0f} {8 * nb / t_b2r / 1e3:6.0f}", flush=True)
