"""Condense the parity evidence the `-m gpu` tests write under gpurun_out/parity_*.json (scratch) into ONE committed
summary: profiles/<prefix>_parity_summary.{json,md}.  Usage: python scripts/collect_parity.py [prefix=r2]"""
import glob
import json
import os
import sys

prefix = sys.argv[1] if len(sys.argv) > 1 else "r2"
src = "gpurun_out"
out = {}


def load(name):
    p = os.path.join(src, name)
    return json.load(open(p)) if os.path.exists(p) else None


def stats(vals):
    v = sorted(float(x) for x in vals)
    return dict(n=len(v), median=v[len(v) // 2], p90=v[min(len(v) - 1, int(0.9 * len(v)))], max=v[-1]) if v else {}


lines = [f"# Parity evidence, round {prefix[1:]} (from the `-m gpu` tests' own output on a B200)", ""]

# per-layer, teacher-forced
lines += ["## Per-layer comparison, teacher-forced (tests/test_parity_layers_gpu.py)", "",
          "Every layer gets the ORACLE's input (activations forward, cotangent backward) and the oracle's recorded "
          "LeakyReLU / max-pool selections; relative L2 error of the layer's own output / input gradient / parameter gradients.",
          "`flips` = fraction of a layer's LeakyReLU / max-pool selections where ours (free) differ from the oracle's.",
          "", "| network | layers | fwd median | fwd max | dx median | dx max | param-grad median | param-grad max | max flips |",
          "|---|---|---|---|---|---|---|---|---|"]
for f in sorted(glob.glob(os.path.join(src, "parity_layers_*.json"))):
    d = json.load(open(f))
    name = os.path.basename(f)[len("parity_layers_"):-5]
    layers = d.get("layers", {})
    rows = layers.values() if isinstance(layers, dict) else layers
    rows = [r for r in rows if isinstance(r, dict)]
    fw = [r["fwd"] for r in rows if r.get("fwd") is not None]
    dx = [x for r in rows for x in (r.get("dx") or [])]
    pg = [x for r in rows for x in (r.get("params") or {}).values()]
    fl = [r["flips"] for r in rows if r.get("flips") is not None]
    sf, sd, sp = stats(fw), stats(dx), stats(pg)
    out["layers_" + name] = dict(fwd=sf, dx=sd, params=sp, flips_max=max(fl) if fl else None, worst=d.get("worst"))
    lines.append(f"| {name} | {sf.get('n', 0)} | {sf.get('median', 0):.2e} | {sf.get('max', 0):.2e} | "
                 f"{sd.get('median', 0):.2e} | {sd.get('max', 0):.2e} | {sp.get('median', 0):.2e} | {sp.get('max', 0):.2e} | "
                 f"{max(fl) if fl else 0:.2e} |")
lines.append("")

# end-to-end with forced selections
lines += ["## End-to-end gradients with the oracle's selections forced vs free-running", "",
          "`forced`: our backward with the oracle's LeakyReLU masks / max-pool indices; `free`: our own selections. The gap is "
          "what bf16 storage does by flipping selections of near-zero pre-activations, not kernel error.", "",
          "| case | masks compared | flip fraction | grads forced: median / p90 / max | grads free: median / max |", "|---|---|---|---|---|"]
for f in sorted(glob.glob(os.path.join(src, "parity_forced_*.json"))):
    d = json.load(open(f))
    name = os.path.basename(f)[len("parity_forced_"):-5]
    out["forced_" + name] = {k: d.get(k) for k in ("masks_compared", "mask_flip_fraction", "median_forced", "p90_forced",
                                                   "max_forced", "median_free", "max_free", "loss")}
    mf = d.get("mask_flip_fraction")
    mf = f"{mf:.2e}" if isinstance(mf, float) else json.dumps(mf)[:60]
    lines.append(f"| {name} | {d.get('masks_compared')} | {mf} | {d.get('median_forced', 0):.2e} / {d.get('p90_forced', 0):.2e} / "
                 f"{d.get('max_forced', 0):.2e} | {d.get('median_free', 0):.2e} / {d.get('max_free', 0):.2e} |")
lines.append("")

# trajectories
lines += ["## Free-running U-Net trajectories (200 SGD steps, deterministic mode, vs the fp32 oracle on CUDA)", "",
          "| size | mean rel | worst rel | worst rel, first 40 steps |", "|---|---|---|---|"]
for size in ("128", "256"):
    d = load(f"parity_unet_trajectory_{size}.json")
    if d:
        out["trajectory_" + size] = {k: d.get(k) for k in ("mean_rel", "worst_rel", "worst_rel_first_40", "steps")}
        lines.append(f"| {size}² | {d['mean_rel']:.2e} | {d['worst_rel']:.2e} | {d['worst_rel_first_40']:.2e} |")
lines.append("")

# the headline iteration
for key, title in (("parity_headline_8p8.json", "Headline iteration at the benchmarked size (8 + 8 slices, 256²) vs the fp32 oracle on CUDA"),
                   ("parity_deterministic.json", "Deterministic mode: eager == eager == graph replay == streams off (bitwise)"),
                   ("parity_graph_vs_eager.json", "Graph replay vs eager step")):
    d = load(key)
    if d is None:
        continue
    out[key[len("parity_"):-5]] = {k: v for k, v in d.items() if not isinstance(v, dict) or len(v) <= 12}
    lines += [f"## {title}", "", "```", json.dumps({k: v for k, v in d.items() if not isinstance(v, dict) or len(v) <= 12}, separators=(",", ":"))[:1500], "```", ""]

for key in ("parity_consis_step_semi0.json", "parity_consis_step_semi1.json", "parity_discriminator_gp.json"):
    d = load(key)
    if d is None:
        continue
    small = {k: v for k, v in d.items() if not isinstance(v, dict)}
    for k, v in d.items():
        if isinstance(v, dict) and v and all(isinstance(x, (int, float)) for x in v.values()):
            small[k + "_stats"] = stats(v.values())
    out[key[len("parity_"):-5]] = small

os.makedirs("profiles", exist_ok=True)
json.dump(out, open(f"profiles/{prefix}_parity_summary.json", "w"), indent=1)
open(f"profiles/{prefix}_parity_summary.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
