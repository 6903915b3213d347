"""Summarise an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel totals for the LAST profiled iteration (the script under ncu ran a warm-up iteration + one iteration).
Usage: summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import re
import sys

src = sys.argv[1]
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ii, ki, mi, ui, vi = (hdr.index(k) for k in ('ID', 'Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value'))
launches = collections.OrderedDict()           # id -> {name, ns, rd, wr}
for x in r:
    rec = launches.setdefault(x[ii], dict(name=x[ki], ns=0.0, rd=0.0, wr=0.0))
    v = float(x[vi].replace(',', ''))
    unit = x[ui].lower()
    scale = {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9, 'nsecond': 1.0, 'usecond': 1e3, 'msecond': 1e6, 'second': 1e9,
             'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(unit, 1.0)
    if x[mi].startswith('gpu__time_duration'):
        rec['ns'] = v * scale
    elif x[mi].startswith('dram__bytes_read'):
        rec['rd'] = v * scale
    elif x[mi].startswith('dram__bytes_write'):
        rec['wr'] = v * scale
data = list(launches.values())
half = data[len(data) // 2:]
tot = sum(d['ns'] for d in half)
have_dram = any(d['rd'] or d['wr'] for d in half)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in half:
    k = re.sub(r'\(.*', '', d['name']).replace('smsut::', '').replace('void ', '')
    k = re.sub(r'at::native::', 'aten::', k)
    k = re.sub(r'^native::', 'aten::', k)
    agg[k][0] += 1
    agg[k][1] += d['ns']
    agg[k][2] += d['rd'] + d['wr']
dram = sum(v[2] for v in agg.values())
out = [f"# ncu launch list summary ({src})", "",
       f"one UGANConsisTrainer iteration (16 slices, 256x256), eager, cold-cache serialised launches: "
       f"{len(half)} launches, {tot / 1e6:.2f} ms total" + (f", {dram / 1e9:.2f} GB of DRAM traffic" if have_dram else ""), "",
       "| kernel | launches | total ms | share | avg us |" + (" DRAM GB | GB/s while running |" if have_dram else ""),
       "|---|---|---|---|---|" + ("---|---|" if have_dram else "")]
for k, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot < 0.002:
        continue
    row = f"| `{k[:80]}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {t / c / 1e3:.1f} |"
    if have_dram:
        row += f" {b / 1e9:.2f} | {b / t:.0f} |"
    out.append(row)
# aten's reductions are `at::native::reduce_kernel<...>` (renamed to aten:: above); a bare 'reduce_kernel' test would also
# swallow the library's own in_bwd_reduce_kernel / in_bwd2_reduce_kernel
ours = sum(t for k, (c, t, b) in agg.items() if not k.startswith('aten') and 'elementwise' not in k and 'cub' not in k
           and 'Memset' not in k and 'nccl' not in k.lower())
out += ["", f"libsmsut_b200 kernels: {100 * ours / tot:.1f}% of the device time; the rest is aten glue "
        "(scalar loss arithmetic, skip-gradient adds, memsets)."]
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
