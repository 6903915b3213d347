"""Development: dump the captured iteration's CUDA graph as DOT and print, for a few kernels of interest, the chain of
predecessors (which node each one waits for).  Usage: graph_dot.py [kernel substring ...]"""
import os
import re
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import graph as G  # noqa: E402
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
os.makedirs("gpurun_out", exist_ok=True)
os.environ["SMSUT_GRAPH_DOT"] = "gpurun_out/step_graph.dot"
step = tr.graphed_step([*batch, a, ids[0], lam], use_semi=True)
txt = open("gpurun_out/step_graph.dot").read()
print("dot bytes", len(txt))
nodes = {}
for m in re.finditer(r'"?(graph_\w+|\w+)"?\s*\[[^\]]*?label\s*=\s*"([^"]*)"', txt):
    nodes[m.group(1)] = m.group(2)
edges = re.findall(r'"?(\w+)"?\s*->\s*"?(\w+)"?', txt)
print("nodes", len(nodes), "edges", len(edges))
open("gpurun_out/step_graph_edges.txt", "w").write("\n".join(f"{a} {b}" for a, b in edges))
import json
json.dump(nodes, open("gpurun_out/step_graph_nodes.json", "w"))
os.remove("gpurun_out/step_graph.dot") if len(txt) > 30 << 20 else None
