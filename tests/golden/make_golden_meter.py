"""Golden traces of the reference's epoch Meter and modality-organ Dice matrix (misc/utils.py:58-160, 180-203).

misc/utils.py cannot be imported offline (medpy, skimage), so `Meter` and `get_mo_matrix` are lifted out of the
reference SOURCE FILE by name with `ast` and executed unchanged in a namespace that provides what they use: the
reference's own `config` and `misc.loss.get_tp_fp_fn_tn` (both importable), numpy with the `np.int` alias numpy 2
removed, and medpy's `dc` restated from its documentation (2 |p & g| / (|p| + |g|), 0 if both are empty).
`collect_dice_by` moves its one-hot tensor with `.cuda(index)`; there is no GPU in the build container, so that one call
is a no-op here.  Run in the build container:  python tests/golden/make_golden_meter.py -> tests/golden/meter.json"""
import ast
import json
import os
import sys
import types
from collections import OrderedDict
from copy import deepcopy

import numpy as np
import torch

REF = "/root/reference"
SRC = os.path.join(REF, "misc", "utils.py")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import config as cfg                                  # noqa: E402  (the reference's)
from misc.loss import get_tp_fp_fn_tn                 # noqa: E402


def dc(result, reference):
    result, reference = np.atleast_1d(result.astype(bool)), np.atleast_1d(reference.astype(bool))
    inter = np.count_nonzero(result & reference)
    size = np.count_nonzero(result) + np.count_nonzero(reference)
    return 2.0 * inter / float(size) if size else 0.0


np_compat = types.ModuleType("np_compat")
np_compat.__dict__.update(np.__dict__)
np_compat.int = int
tree = ast.parse(open(SRC).read())
wanted = [n for n in tree.body if (isinstance(n, ast.ClassDef) and n.name == "Meter")
          or (isinstance(n, ast.FunctionDef) and n.name == "get_mo_matrix")]
ns = dict(OrderedDict=OrderedDict, deepcopy=deepcopy, torch=torch, cfg=cfg, get_tp_fp_fn_tn=get_tp_fp_fn_tn, np=np_compat,
          dc=dc)
exec(compile(ast.Module(body=wanted, type_ignores=[]), SRC, "exec"), ns)
Meter, get_mo_matrix = ns["Meter"], ns["get_mo_matrix"]
torch.Tensor.cuda = lambda self, *a, **k: self        # see the docstring

rng = np.random.default_rng(7)
out = dict(traces=[], dice=[], mo=[])
min_keys = [f"loss_{i}" for i in range(cfg.n_modal)] + ["loss"]
max_keys = [f"dice_{i}" for i in range(cfg.n_modal)] + ["dice"]
for alpha, with_max, reset_at in ((1.0, True, None), (0.7, False, None), (0.5, True, 2)):
    m = Meter(min_better_keys=min_keys, max_better_keys=max_keys if with_max else [], alpha=alpha)
    epochs = []
    for epoch in range(4):
        m.reset_cur()
        calls = []
        for _ in range(6):
            loss, modal, n = float(rng.uniform(0.2, 3.0)), int(rng.integers(0, cfg.n_modal - (epoch == 1))), int(rng.integers(1, 9))
            v, k = Meter.collect_loss_by(loss, modal, n)
            m.accumulate(v, k)
            calls.append(dict(loss=loss, modal=modal, n=n, v=v, k=k))
        dice = None
        if with_max:
            dice = {k: float(rng.uniform(0.1, 0.9)) for k in max_keys}
            m.accumulate(dice, {k: 1. for k in dice})
        m.update_cur(reset_best=(reset_at == epoch))
        epochs.append(dict(calls=calls, dice=dice, reset_best=(reset_at == epoch), cur=dict(m.cur_values),
                           best=dict(m.best_values), pre=dict(m.pre_values), text=str(m)))
    out["traces"].append(dict(alpha=alpha, min_keys=min_keys, max_keys=max_keys if with_max else [], epochs=epochs))

gen = torch.Generator().manual_seed(11)
for b, c, h in ((5, 5, 10), (3, 5, 8), (4, 3, 6)):
    vals = np.round(torch.randn(b, c, h, h, generator=gen).double().numpy(), 3).tolist()      # short decimals in the JSON
    logits = torch.tensor(vals, dtype=torch.float32)
    gt = torch.randint(0, c, (b, h, h), generator=gen)
    gt[0] = 0                                           # a slice without foreground: the smooth term decides
    modal = torch.randint(0, cfg.n_modal, (b,), generator=gen)
    a, n = Meter.collect_dice_by(logits, gt, modal, cfg.n_modal)
    out["dice"].append(dict(logits=vals, gt=gt.tolist(), modal=modal.tolist(), a=a, n=n))

names = list(cfg.Modality.__members__)
for seed in (0, 1):
    r = np.random.default_rng(seed)
    gt, prd = {}, {}
    for m_i, vols in ((0, 2), (2, 1), (3, 3)):          # one modality without volumes: the 1e-8 divisor
        for p in range(vols):
            g = r.integers(0, cfg.n_label + 1, size=(3, 8, 8))
            q = np.where(r.uniform(size=g.shape) < 0.7, g, r.integers(0, cfg.n_label + 1, size=g.shape))
            if p == 0:
                g[g == 2] = 0                            # an organ missing from the label: dc's empty rule
                q[q == 2] = 0
            gt[f"{names[m_i]}_{p}"], prd[f"{names[m_i]}_{p}"] = g, q
    mat = get_mo_matrix(prd, gt)
    out["mo"].append(dict(gt={k: v.tolist() for k, v in gt.items()}, prd={k: v.tolist() for k, v in prd.items()},
                          matrix=mat.tolist()))

json.dump(out, open(os.path.join(HERE, "meter.json"), "w"))
print("wrote", len(out["traces"]), "meter traces,", len(out["dice"]), "dice cases,", len(out["mo"]), "matrices")
