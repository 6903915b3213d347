"""Summarise an `ncu --page raw --csv` export: one row per profiled launch with the metrics the roofline uses.
Usage: summarize_ncu.py raw.csv out.md [--conv-classes out.json]   (conv classes: the measured launches are the odd ones)"""
import csv
import json
import sys

src, out_md = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]


def col(prefix):
    for i, h in enumerate(hdr):
        if h == prefix:
            return i
    for i, h in enumerate(hdr):
        if h.startswith(prefix):
            return i
    return None


def num(r, name, default=0.0):
    i = col(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    return float(r[i].replace(",", ""))


def scaled(r, name):
    """value in base units (bytes / seconds) using the unit row"""
    i = col(name)
    if i is None:
        return 0.0
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    table = {"tbyte": 1e12, "gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
             "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
    return v * table.get(u, 1.0)


lines = ["| # | kernel | grid | us | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % | L2->SM MB | tensor pipe % | warps active % | regs |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
recs = []
for n, r in enumerate(rows[2:]):
    name = r[col("Kernel Name")].split("(")[0].replace("smsut::", "").replace("void ", "")
    dur = scaled(r, "gpu__time_duration.sum")
    rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
    l2 = scaled(r, "l1tex__m_xbar2l1tex_read_bytes.sum")
    tens = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
               num(r, "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"))
    rec = dict(kernel=name, grid=r[col("Grid Size")], us=dur * 1e6, dram_read=rd, dram_write=wr,
               dram_pct=num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), l2_bytes=l2, tensor_pct=tens,
               warps_pct=num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
               regs=num(r, "launch__registers_per_thread"))
    recs.append(rec)
    lines.append(f"| {n} | `{name[:48]}` | {rec['grid']} | {rec['us']:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                 f"{(rd + wr) / dur / 1e9:.0f} | {rec['dram_pct']:.1f} | {l2 / 1e6:.1f} | {tens:.1f} | {rec['warps_pct']:.1f} | {rec['regs']:.0f} |")
open(out_md, "w").write("# ncu --set full summary (" + src + ")\n\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
if "--conv-classes" in sys.argv:
    sys.path.insert(0, ".")
    import bench
    meas = recs[1::2]
    assert len(meas) == len(bench.CONV_CLASSES), (len(meas), len(bench.CONV_CLASSES))
    tot_b = tot_n = 0.0
    per = []
    for (cins, cout, h, ks, count), m in zip(bench.CONV_CLASSES, meas):
        per.append(dict(cin=sum(cins), cout=cout, hw=h, kernel=m["kernel"], us_under_ncu=m["us"],
                        dram_bytes=m["dram_read"] + m["dram_write"], tensor_pct=m["tensor_pct"]))
        tot_b += count * (m["dram_read"] + m["dram_write"])
        tot_n += count
    json.dump(dict(source=src, dram_bytes_per_launch_weighted=tot_b / tot_n, per_class=per),
              open(sys.argv[sys.argv.index("--conv-classes") + 1], "w"), indent=1)
