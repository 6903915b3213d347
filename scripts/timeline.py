"""Development: kernel timeline of the captured UGANConsisTrainer iteration through torch.profiler (CUPTI activity
records carry per-stream start / end device timestamps also for kernels launched by a graph replay).  Writes
gpurun_out/timeline.json = [[name, stream, start_us, dur_us], ...] of ONE replay and prints an occupancy summary."""
import json
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200.data_loader import syntheticLoader as synlod  # noqa: E402
from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer  # noqa: E402

torch.manual_seed(0)
tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=256))
lb = synlod.get_loader(None, 'train', 0, 8, pool_batches=1)
ul = synlod.get_loader(None, 'val', 0, 8, pool_batches=1)
(x1, y, m1, _), (x2, _, m2, _) = next(iter(lb)), next(iter(ul))
batch = tr.prepare_batch(x1, y, m1, x2, m2, 2)
a, ids = tr.draw(16)
lam = torch.full((1,), 0.5, device="cuda")
step = tr.graphed_step([*batch, a, ids[0], lam], use_semi=True)
for _ in range(3):
    step(*batch, a, ids[0], lam)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step(*batch, a, ids[0], lam)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/timeline_trace.json")
tr_json = json.load(open("gpurun_out/timeline_trace.json"))
ev = [e for e in tr_json["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
print("gpu events", len(ev))
# split into replays by the big gaps (host sync between them is absent; use the memset of arena_begin as the marker)
marks = [i for i, e in enumerate(ev) if e["cat"] == "gpu_memset" and e.get("args", {}).get("bytes", 0) >= 40 << 20]
print("arena memsets at", marks[:6])
if len(marks) >= 3:
    seg = ev[marks[1]:marks[2]]
else:
    seg = ev[len(ev) // 3: 2 * len(ev) // 3]
t0 = seg[0]["ts"]
rows = [[e["name"][:60], e.get("args", {}).get("stream", -1), round(e["ts"] - t0, 2), round(e["dur"], 2)] for e in seg]
json.dump(rows, open("gpurun_out/timeline.json", "w"))
os.remove("gpurun_out/timeline_trace.json")
end = max(r[2] + r[3] for r in rows)
busy = sum(r[3] for r in rows)
streams = sorted({r[1] for r in rows})
print(f"one replay: {len(rows)} gpu ops on {len(streams)} streams, span {end / 1e3:.2f} ms, sum of durations {busy / 1e3:.2f} ms")
# concurrency profile: time with k kernels in flight
pts = []
for r in rows:
    pts.append((r[2], 1)); pts.append((r[2] + r[3], -1))
pts.sort()
cur, last, hist = 0, 0.0, {}
for t, d in pts:
    hist[cur] = hist.get(cur, 0.0) + (t - last)
    cur += d
    last = t
print("time (ms) with k kernels in flight:", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
for s in streams:
    rs = [r for r in rows if r[1] == s]
    print(f"stream {s}: {len(rs)} ops, busy {sum(r[3] for r in rs) / 1e3:.2f} ms, first {rs[0][2] / 1e3:.2f} last end {(rs[-1][2] + rs[-1][3]) / 1e3:.2f}")
