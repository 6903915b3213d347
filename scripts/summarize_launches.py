"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST profiled
iteration (the script under ncu ran warm-up + one iteration).  Usage: summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import re
import sys

src = sys.argv[1]
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
data = [(x[ki], float(x[vi].replace(',', '')), x[gi]) for x in r]
half = data[len(data) // 2:]
tot = sum(v for _, v, _ in half)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, g in half:
    k = re.sub(r'\(.*', '', k).replace('smsut::', '').replace('void ', '')
    k = re.sub(r'at::native::', 'aten::', k)
    agg[k][0] += 1
    agg[k][1] += v
out = [f"# ncu launch list summary ({src})", "",
       f"one UGANConsisTrainer iteration (16 slices, 256x256), eager, cold-cache serialised launches: "
       f"{len(half)} launches, {tot / 1e6:.2f} ms total", "",
       "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot < 0.002:
        continue
    out.append(f"| `{k[:80]}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {t / c / 1e3:.1f} |")
# aten's reductions are `at::native::reduce_kernel<...>` (already renamed to aten:: above); a bare 'reduce_kernel' test
# would also swallow the library's own in_bwd_reduce_kernel / in_bwd2_reduce_kernel
ours = sum(t for k, (c, t) in agg.items() if not k.startswith('aten') and 'elementwise' not in k and 'cub' not in k
           and 'Memset' not in k and 'nccl' not in k.lower())
out += ["", f"libsmsut_b200 kernels: {100 * ours / tot:.1f}% of the device time; the rest is aten glue "
        "(scalar loss arithmetic, skip-gradient adds, memsets)."]
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
