"""Golden index streams of the reference's InTurn samplers (data_loader/inTurnLoader.py:15-80).

The reference module cannot be imported offline (its imports pull medpy / elasticdeform through balanceLoader ->
misc.utils), so the two sampler classes are lifted out of the reference SOURCE FILE by name with `ast` and executed
unchanged in an empty namespace that provides only what they use (`Sampler`, `List`, `random`).  Run in the build
container:  python tests/golden/make_golden_sampler.py  -> tests/golden/inturn_sampler.json"""
import ast
import json
import os
import random
from typing import List

from torch.utils.data import Sampler

SRC = "/root/reference/data_loader/inTurnLoader.py"
HERE = os.path.dirname(os.path.abspath(__file__))

tree = ast.parse(open(SRC).read())
wanted = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("InTurnTrainBatchSampler", "InTurnTestBatchSampler")]
ns = {"Sampler": Sampler, "List": List, "random": random}
exec(compile(ast.Module(body=wanted, type_ignores=[]), SRC, "exec"), ns)

cases = []
for seed, sizes, bs, shuffle in ((1, (37, 20, 53, 16), 8, False), (2, (37, 20, 53, 16), 8, True), (3, (9, 9, 9, 9), 4, False),
                                 (4, (64, 8, 24, 40), 8, False), (5, (5, 17), 2, True)):
    samples, n = [], 0
    for s in sizes:
        samples.append(list(range(n, n + s)))
        n += s
    random.seed(seed)
    sampler = ns["InTurnTrainBatchSampler"]([list(x) for x in samples], bs, shuffle)
    epochs = [[list(b) for b in sampler] for _ in range(2)]            # state carries over between epochs
    test = [list(b) for b in ns["InTurnTestBatchSampler"]([list(x) for x in samples], bs)]
    cases.append(dict(seed=seed, sizes=sizes, batch_size=bs, shuffle=shuffle, train_len=len(sampler), epochs=epochs,
                      test=test, test_len=len(ns["InTurnTestBatchSampler"]([list(x) for x in samples], bs))))
json.dump(cases, open(os.path.join(HERE, "inturn_sampler.json"), "w"))
print("wrote", len(cases), "cases")

# ---- data_loader/balanceLoader.py:80-109, lifted the same way -> tests/golden/balance_sampler.json ------------------
SRC_B = "/root/reference/data_loader/balanceLoader.py"
tree_b = ast.parse(open(SRC_B).read())
wanted_b = [n for n in tree_b.body if isinstance(n, ast.ClassDef) and n.name == "ModalityBalanceBatchSampler"]
exec(compile(ast.Module(body=wanted_b, type_ignores=[]), SRC_B, "exec"), ns)
cases_b = []
for seed, sizes, bs in ((1, (37, 20, 53, 16), 8), (2, (9, 9, 9, 9), 4), (3, (64, 8, 24, 40), 8), (4, (5, 17), 2), (5, (12, 12, 12, 12), 16)):
    samples, n = [], 0
    for s in sizes:
        samples.append(list(range(n, n + s)))
        n += s
    random.seed(seed)
    sampler = ns["ModalityBalanceBatchSampler"]([list(x) for x in samples], bs)
    cases_b.append(dict(seed=seed, sizes=sizes, batch_size=bs, length=len(sampler),
                        epochs=[[list(b) for b in sampler] for _ in range(3)]))      # cursors carry over between epochs
json.dump(cases_b, open(os.path.join(HERE, "balance_sampler.json"), "w"))
print("wrote", len(cases_b), "balance cases")
