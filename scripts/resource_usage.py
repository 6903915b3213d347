"""Registers / spills / shared memory per kernel of the built library (cuobjdump --dump-resource-usage; no GPU needed)
-> profiles/<tag>_resource_usage.md.  Template instances of one kernel are folded into one row (ranges)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "smsut-medicalimgsegmentation_b200", "libsmsut_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["/usr/local/cuda/bin/cu++filt"], input="\n".join(re.findall(r"Function (\S+):", out)),
                       capture_output=True, text=True).stdout.split("\n")
rows = collections.OrderedDict()
it = iter(names)
for m in re.finditer(r"Function (\S+):\n\s*(.*)", out):
    demangled = next(it)
    base = re.sub(r"^(void )?(smsut::)?(<unnamed>::|\(anonymous namespace\)::)?", "", demangled)
    base = re.split(r"[<(]", base)[0]
    f = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
    r = rows.setdefault(base, dict(n=0, reg=[], stack=[], shared=[], local=[]))
    r["n"] += 1
    r["reg"].append(int(f.get("REG", 0)))
    r["stack"].append(int(f.get("STACK", 0)))
    r["shared"].append(int(f.get("SHARED", 0)))
    r["local"].append(int(f.get("LOCAL", 0)))


def rng(v):
    return str(v[0]) if min(v) == max(v) else f"{min(v)}–{max(v)}"


path = os.path.join(ROOT, "profiles", f"{tag}_resource_usage.md")
with open(path, "w") as f:
    f.write("# Per-kernel resources of libsmsut_b200.so (sm_100a, `cuobjdump --dump-resource-usage`)\n\n"
            "`stack` = the per-thread local-memory frame: the argument area of device `printf` in the bounded-wait traps "
            "(16-96 B in the tensor-core kernels) and register spills where a `__launch_bounds__` occupancy target was "
            "preferred to a spill-free build (the InstanceNorm backward instances, DESIGN.md section 3); `local` = "
            "statically declared local arrays.  Static shared memory only (the tensor-core kernels take their stage "
            "rings as dynamic shared memory).\n\n| kernel | instances | registers / thread | stack B | static shared B | local B |\n|---|---|---|---|---|---|\n")
    for k, r in sorted(rows.items(), key=lambda kv: -max(kv[1]["reg"])):
        f.write(f"| `{k}` | {r['n']} | {rng(r['reg'])} | {rng(r['stack'])} | {rng(r['shared'])} | {rng(r['local'])} |\n")
print("wrote", path, len(rows), "kernels")
