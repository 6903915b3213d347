"""Development: achieved read bandwidth of in_stats (1 read stream) vs tensor size, vs torch reductions / copies."""
import sys

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import ops  # noqa: E402

flush = torch.empty(192 * 2 ** 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
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


for n in (16, 64, 256):
    x = torch.randn(n, 256, 256, 16, device="cuda").to(torch.bfloat16)
    nb = x.numel() * 2
    y = torch.empty_like(x)
    t1 = timed(lambda: ops.in_stats(x))
    t2 = timed(lambda: x.view(torch.int32).sum())
    t3 = timed(lambda: y.copy_(x))
    t4 = timed(lambda: ops.act_fwd(x, ops.ACT_LRELU))
    print(f"{nb / 2**20:7.1f} MiB: in_stats {t1:7.1f} us {nb / t1 / 1e3:6.0f} GB/s | torch int32 sum {t2:7.1f} us {nb / t2 / 1e3:6.0f} GB/s | "
          f"torch copy {t3:7.1f} us {2 * nb / t3 / 1e3:6.0f} GB/s (r+w) | act_fwd {t4:7.1f} us {2 * nb / t4 / 1e3:6.0f} GB/s (r+w)", flush=True)
