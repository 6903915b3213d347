"""misc/utils.py (the epoch Meter of BaseTrainer.fit, the modality-organ Dice matrix, directory / yaml helpers) against
traces of the reference's own classes (tests/golden/meter.json, made by make_golden_meter.py from the reference's
source), and the fit loop's use of them on the fp32 test double."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "meter.json")))


def close(a, b, tol=1e-9):
    assert set(a) == set(b), (sorted(a), sorted(b))
    for k in a:
        assert abs(a[k] - b[k]) <= tol * max(1.0, abs(b[k])), (k, a[k], b[k])


@pytest.mark.parametrize("trace", GOLD["traces"], ids=lambda t: f"alpha{t['alpha']}")
def test_meter_epochs_match_reference_trace(pkg, trace):
    from smsut_b200.misc.utils import Meter
    m = Meter(min_better_keys=trace["min_keys"], max_better_keys=trace["max_keys"], alpha=trace["alpha"])
    assert list(m.configs.items()) == [(k, "min") for k in trace["min_keys"]] + [(k, "max") for k in trace["max_keys"]]
    assert m.pre_values is None
    for ep in trace["epochs"]:
        m.reset_cur()
        assert all(v == 0 for v in m.cur_values.values()) and all(v == 0 for v in m.n.values())
        for c in ep["calls"]:
            v, n = Meter.collect_loss_by(c["loss"], c["modal"], c["n"])
            assert v == c["v"] and n == c["k"]
            m.accumulate(v, n)
        if ep["dice"] is not None:
            m.accumulate(ep["dice"], {k: 1. for k in ep["dice"]})
        m.update_cur(reset_best=ep["reset_best"])
        close(m.cur_values, ep["cur"])
        close(m.best_values, ep["best"])
        close(m.pre_values, ep["pre"])
        assert str(m) == ep["text"]


@pytest.mark.parametrize("case", GOLD["dice"], ids=lambda c: f"b{len(c['modal'])}")
def test_collect_dice_by_matches_reference(pkg, case):
    from smsut_b200 import config as cfg
    from smsut_b200.misc.utils import Meter
    logits, gt = torch.tensor(case["logits"], dtype=torch.float32), torch.tensor(case["gt"])
    a, n = Meter.collect_dice_by(logits, gt, torch.tensor(case["modal"]), cfg.n_modal)
    close(a, case["a"], 1e-6)           # the reference sums float32 per-slice values
    assert n == case["n"]


@pytest.mark.parametrize("case", GOLD["mo"], ids=["seed0", "seed1"])
def test_modality_organ_matrix_matches_reference(pkg, case):
    from smsut_b200.misc.utils import connected_components, get_all_matrix, get_mo_matrix
    gt = {k: np.array(v) for k, v in case["gt"].items()}
    prd = {k: np.array(v) for k, v in case["prd"].items()}
    assert np.abs(get_mo_matrix(prd, gt) - np.array(case["matrix"])).max() < 1e-12
    # get_all_matrix: the Dice block differs from get_mo_matrix only through the connected-component clean-up
    dice, hd, sd = get_all_matrix({k: v.copy() for k, v in prd.items()}, gt)
    cleaned = {}
    for k, v in prd.items():
        c = connected_components(v)
        cleaned[k] = np.stack([connected_components(sl) for sl in c])
    assert np.abs(dice - get_mo_matrix(cleaned, gt)).max() < 1e-12 and np.array_equal(dice, hd)
    assert sd.shape == dice.shape and np.isfinite(sd).all() and (sd >= 0).all()


def test_directory_yaml_and_label_volume_helpers(pkg, tmp_path, monkeypatch):
    from smsut_b200 import config as cfg
    from smsut_b200.misc import utils
    a, b = tmp_path / "a", tmp_path / "a" / "b"
    utils.maybe_mkdir(str(a), str(b))
    utils.maybe_mkdir(str(a))                       # existing directories are left alone
    assert a.is_dir() and b.is_dir()
    with pytest.raises(FileNotFoundError):          # like os.mkdir: no parents
        utils.maybe_mkdir(str(tmp_path / "x" / "y"))
    split = {m: dict(train=["1"], val=["2"], test=["3", "4"] if m == "ct" else ["3"]) for m in cfg.Modality.__members__}
    utils.write_yaml(split, str(tmp_path / cfg.split_yaml))
    assert utils.read_yaml(str(tmp_path / cfg.split_yaml)) == split
    rng = np.random.default_rng(0)
    for m in cfg.Modality.__members__:
        for p, z in (("3", 5), ("4", 2)):
            os.makedirs(tmp_path / m / p, exist_ok=True)
            np.save(tmp_path / m / p / f"{m}_{p}.npy", rng.integers(0, 5, size=(z, 8, 8)).astype(np.uint8))
    n, vols = utils.get_label_npys(str(tmp_path), "all", "test")
    assert n == 5 * 4 + 2 and set(vols) == {f"{m}_3" for m in cfg.Modality.__members__} | {"ct_4"}
    n, vols = utils.get_label_npys(str(tmp_path), "t2", "test")
    assert n == 5 and list(vols) == ["t2_3"] and vols["t2_3"].shape == (5, 8, 8)


# ----------------------------------------------------------------------------------------------------------------------
# the meters inside the epoch loop (fp32 test double of the kernel layer)
# ----------------------------------------------------------------------------------------------------------------------
sys.path.insert(0, HERE)
import cpu_ops_mock                                   # noqa: E402
from oracle import smsut_oracle as O                  # noqa: E402

os.environ["SMSUT_ALLOW_CPU_TEST_DOUBLE"] = "1"


@pytest.fixture()
def exact(pkg):
    with cpu_ops_mock.installed(exact=True) as ops:
        yield ops


def _loader(batches):
    class L(list):
        dataset = None
    return L(batches)


def test_fit_fills_the_reference_meters(exact, tmp_path, monkeypatch, capsys):
    """BaseTrainer.fit (baseTrainer.py:147-199): the train meter holds the slice-weighted mean of the iterations' losses
    per modality, the test meter the per-modality validation loss (weighted with the PADDED size of a ragged last batch,
    as the reference does) and the modality-organ Dice; `best` follows the test meter; [TRN] / [TST] lines are logged."""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    monkeypatch.setattr(cfg, "batch_size", 2)
    size = 32
    tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name="m", input_size=size))
    tr.expr_root = str(tmp_path)
    sd = O.make_weights(O.unet_shapes(), 3)
    tr.net.load_state_dict(sd)
    mods = [0, 2, 2, 3]
    train = []
    for i, m in enumerate(mods):
        x, y = O.synthetic_batch(2, size, 40 + i)
        train.append((x, y, torch.full((2,), m), [f"{m}_p_{z}" for z in range(2)]))
    test = []
    for i, (m, n) in enumerate(((0, 2), (0, 1), (3, 2))):          # a ragged batch of one slice
        x, y = O.synthetic_batch(n, size, 60 + i)
        test.append((x, y, torch.full((n,), m), [f"{m}_q{i}_{z}" for z in range(n)]))
    tr.fit(loaders=(_loader(train), _loader(train), _loader(test)), max_epoch=2, iters_per_epoch=4)
    train_meter, test_meter = tr.meters

    # the same two epochs on the oracle: per-iteration losses, then the validation loss of the weights after each epoch
    st, w = {}, dict(sd)
    it = 0
    for epoch in range(2):
        sums, cnt = {}, {}
        for (x, y, mdl, _) in train:
            loss, _ = O.unet_step(w, st, x, y, O.poly_lr(1e-2, max(it - 1, 0), cfg.max_epoch * cfg.num_iter_per_epoch))
            it += 1
            for k in ("loss", f"loss_{int(mdl[0])}"):
                sums[k] = sums.get(k, 0.0) + loss.item() * 2
                cnt[k] = cnt.get(k, 0) + 2
        means = {k: sums[k] / cnt[k] for k in sums}
    # cfg.exp_alpha = 1: no smoothing, the meter holds the last epoch.  2e-3: eight free-running SGD steps of the fp32 test
    # double and the oracle (test_host_logic.py allows the weights 1e-3 after three)
    for k, v in means.items():
        assert abs(train_meter.cur_values[k] - v) < 2e-3 * max(1.0, abs(v)), (k, train_meter.cur_values[k], v)
    assert train_meter.cur_values["loss_1"] == 0 and train_meter.n["loss_1"] == 0
    assert train_meter.n["loss"] == 8 and train_meter.n["loss_2"] == 4

    sums, cnt = {}, {}
    with torch.no_grad():
        for (x, y, mdl, _) in test:
            loss = O.dice_ce_loss(O.unet_forward({k: v for k, v in tr.net.state_dict().items()}, x), y).item()
            for k in ("loss", f"loss_{int(mdl[0])}"):
                sums[k] = sums.get(k, 0.0) + loss * cfg.batch_size          # padded size, baseTrainer.py:214-230
                cnt[k] = cnt.get(k, 0) + cfg.batch_size
    for k in sums:
        assert abs(test_meter.cur_values[k] - sums[k] / cnt[k]) < 1e-5, (k, test_meter.cur_values[k], sums[k] / cnt[k])
    dices = tr.validate_dice()[0]
    for k, v in dices.items():
        assert abs(test_meter.cur_values[k] - v) < 1e-12
    assert test_meter.best_values["dice"] >= test_meter.cur_values["dice"]
    assert test_meter.best_values["loss"] <= test_meter.cur_values["loss"]
    out = capsys.readouterr().out
    assert out.count("[TRN] Epoch:") == 2 and out.count("[TST] Epoch:") == 2
    assert " loss_ct:" in out and " dice_t2:" in out and "lr: " in out
    for c in ("best.ckpt", "last.ckpt"):
        assert os.path.exists(os.path.join(str(tmp_path), "000", "ckpt", c))


def test_consis_epoch_notes_the_segmentation_loss(exact, monkeypatch):
    """UGANConsisTrainer.train_epoch (uganConsisTrainer.py:96,112,157-158): G_seg under the labelled batch's modality,
    weighted with cfg.batch_size; one flush per epoch"""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.misc.utils import Meter
    from smsut_b200.trainer.uganConsisTrainer import LOSS_KEYS, UGANConsisTrainer
    monkeypatch.setattr(cfg, "batch_size", 2)
    size = 64
    tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=size))
    lb, ul = [], []
    for i, m in enumerate((1, 3)):
        x, y = O.synthetic_batch(2, size, 80 + i)
        lb.append((x, y, torch.full((2,), m), None))
        x2, _ = O.synthetic_batch(2, size, 90 + i)
        ul.append((x2, None, torch.full((2,), (m + 1) % 4), None))
    seen = []
    step = tr.train_step
    monkeypatch.setattr(tr, "train_step", lambda *a, **k: seen.append(step(*a, **k)) or seen[-1])
    meter = tr.make_meters()[0]
    tr.train_epoch(_loader(lb), _loader(ul), meter, num_iter=2)
    assert tr._meter_queue == [] and len(seen) == 2
    seg = [float(s[LOSS_KEYS.index("G_seg")]) for s in seen]
    # the loaders' first batches are the epoch's fixed sample images (L82-93): iteration 0 trains on the SECOND labelled
    # batch (modality 3), iteration 1 on the first one again (modality 1) after the loader has been restarted
    assert abs(meter.cur_values["loss_3"] - seg[0] * 2) < 1e-6 and abs(meter.cur_values["loss_1"] - seg[1] * 2) < 1e-6
    assert abs(meter.cur_values["loss"] - (seg[0] + seg[1]) * 2) < 1e-6 and meter.n["loss"] == 4
    assert isinstance(meter, Meter) and meter.n["loss_0"] == 0


def test_run_directory_layout_train_log_and_tensorboard_scalars(exact, tmp_path, monkeypatch):
    """init_train_env (baseTrainer.py:81-98) + the scalars of fit (:165-172, :187-193): <run>/{ckpt,tb,result,sample},
    train.log with the epoch lines, `train/<key>`, `train/lr`, `test/<key>` per epoch with modality names"""
    from types import SimpleNamespace
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    from smsut_b200 import config as cfg
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    monkeypatch.setattr(cfg, "batch_size", 2)
    tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name="tb", input_size=32))
    tr.expr_root = str(tmp_path)
    batches = []
    for i, m in enumerate((0, 1)):
        x, y = O.synthetic_batch(2, 32, 20 + i)
        batches.append((x, y, torch.full((2,), m), [f"{m}_v_{z}" for z in range(2)]))
    tr.fit(loaders=(_loader(batches), _loader(batches), _loader(batches)), max_epoch=2, iters_per_epoch=2)
    run = os.path.join(str(tmp_path), "000")
    assert sorted(os.listdir(run)) == ["ckpt", "result", "sample", "tb", "train.log"]
    log = open(os.path.join(run, "train.log")).read()
    assert log.count("[TRN] Epoch:") == 2 and log.count("[TST] Epoch:") == 2 and "Save model to" in log
    tr.close_run_logs()
    ea = EventAccumulator(os.path.join(run, "tb"))
    ea.Reload()
    tags = set(ea.Tags()["scalars"])
    want = {f"train/loss_{n}" for n in cfg.Modality.__members__} | {"train/loss", "train/lr", "test/loss", "test/dice"} \
        | {f"test/dice_{n}" for n in cfg.Modality.__members__} | {f"test/loss_{n}" for n in cfg.Modality.__members__}
    assert tags == want, tags ^ want
    train_meter, test_meter = tr.meters
    ev = ea.Scalars("test/dice")
    assert [e.step for e in ev] == [0, 1] and abs(ev[1].value - test_meter.cur_values["dice"]) < 1e-6
    assert abs(ea.Scalars("train/loss_ct")[1].value - train_meter.cur_values["loss_0"]) < 1e-6
    # a second run under the same experiment gets the next number (baseTrainer.py:83)
    tr2 = UnetTrainer('train', SimpleNamespace(fold=0, expr_name="tb", input_size=32))
    tr2.expr_root = str(tmp_path)
    assert tr2.model_idx == "001" and os.path.isdir(os.path.join(str(tmp_path), "001", "tb"))
    tr2.close_run_logs()


def test_reference_helper_names_are_importable(exact, tmp_path):
    """names a caller of the reference's modules may import: the class name of meanTeacherTrainer.py:35, loss.py:23's
    get_tp_fp_fn_tn, externalTransforms.py:12's MaskToTensor, init_train_env / register_experiment_args"""
    from types import SimpleNamespace
    from PIL import Image
    from smsut_b200.data_loader.externalTransforms import MaskToTensor
    from smsut_b200.misc.loss import get_tp_fp_fn_tn
    from smsut_b200.trainer.meanTeacherTrainer import MeanTeacherTrainer, meanTeacherTrainer
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    assert meanTeacherTrainer is MeanTeacherTrainer
    lab = np.arange(12, dtype=np.uint8).reshape(3, 4) % 5
    t = MaskToTensor()(Image.fromarray(lab))
    assert t.dtype == torch.int64 and torch.equal(t, torch.from_numpy(lab).long())
    gen = torch.Generator().manual_seed(0)
    prob = torch.softmax(torch.randn(2, 5, 6, 6, generator=gen), dim=1)
    gt = torch.randint(0, 5, (2, 6, 6), generator=gen)
    tp, fp, fn, tn = get_tp_fp_fn_tn(prob, gt)
    onehot = torch.nn.functional.one_hot(gt, 5).permute(0, 3, 1, 2).float()
    assert tp.shape == (2, 5) and torch.allclose(tp, (prob * onehot).sum((2, 3)))
    assert torch.allclose(tp + fn, onehot.sum((2, 3))) and torch.allclose(tp + fp + fn + tn, torch.full((2, 5), 36.))
    assert get_tp_fp_fn_tn(prob, gt, dims=(0, 2, 3))[0].shape == (5,)
    tr = UnetTrainer('train', SimpleNamespace(fold=1, expr_name="e", input_size=32))
    assert tr.init_train_env(str(tmp_path)) == "000" and os.path.isdir(os.path.join(str(tmp_path), "000", "sample"))
    tr.register_experiment_args(str(tmp_path))
    reg = open(os.path.join(str(tmp_path), "expriments.log")).read()
    assert reg.startswith("UnetTrainer, " + os.path.join(str(tmp_path), "000")) and "fold=1" in reg
    tr.close_run_logs()


def test_validate_epoch_and_dice_the_reference_way(exact, monkeypatch):
    """baseTrainer.py:207-252 called as the reference's fit / test call it: validate_epoch(loader, npys) returns (number
    of slices, prediction volumes), validate_dice(prd_npys, gt_npys) the Dice dict -- equal to the dict the device-side
    confusion counts give, incl. a ragged last batch and a volume split over two batches"""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    monkeypatch.setattr(cfg, "batch_size", 4)
    tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=32))
    sd = O.make_weights(O.unet_shapes(), 5)
    tr.net.load_state_dict(sd)
    layout = [('ct', 'a', 4), ('ct', 'b', 4), ('ct', 'b', 2), ('t1out', 'c', 3)]
    batches, z0, gt = [], {}, {}
    for bi, (m, pid, n) in enumerate(layout):
        img, lab = O.synthetic_batch(n, 32, 500 + bi)
        s = z0.get((m, pid), 0)
        batches.append((img, lab, torch.full((n,), cfg.Modality[m].value), [f"{m}_{pid}_{s + z}" for z in range(n)]))
        z0[(m, pid)] = s + n
        key = f"{m}_{pid}"
        gt[key] = np.concatenate([gt[key], lab.numpy()]) if key in gt else lab.numpy()
    n_slices, prd = tr.validate_epoch(batches, gt)
    assert n_slices == 13 and set(prd) == set(gt) and all(prd[k].shape == gt[k].shape for k in gt)
    for img, lab, mdl, names in batches:
        pred = O.unet_forward(sd, img).argmax(1).numpy()
        for i, nm in enumerate(names):
            m, pid, z = nm.split('_')
            assert np.array_equal(prd[f"{m}_{pid}"][int(z)], pred[i]), nm
    ref_way = tr.validate_dice(prd, gt)
    device_way, matrix = tr.validate_dice()
    assert set(ref_way) == set(device_way)
    for k in ref_way:
        assert abs(ref_way[k] - device_way[k]) < 1e-12, k
    assert isinstance(tr.validate_epoch(batches), float)


def _flood_components(mask, full):
    """connected components by flood fill: neighbours differ by at most 1 per axis, in at most 2 axes when `full`"""
    import itertools
    offsets = [o for o in itertools.product((-1, 0, 1), repeat=mask.ndim)
               if any(o) and sum(abs(x) for x in o) <= (2 if full else 1)]
    seen, comps = np.zeros(mask.shape, bool), []
    for start in zip(*np.nonzero(mask)):
        if seen[start]:
            continue
        comp, stack = [], [start]
        seen[start] = True
        while stack:
            cur = stack.pop()
            comp.append(cur)
            for o in offsets:
                nb = tuple(c + d for c, d in zip(cur, o))
                if all(0 <= x < s for x, s in zip(nb, mask.shape)) and mask[nb] and not seen[nb]:
                    seen[nb] = True
                    stack.append(nb)
        comps.append(comp)
    return comps


@pytest.mark.parametrize("shape", [(24, 24), (5, 12, 12)])
def test_connected_components_against_flood_fill(pkg, shape):
    """utils.py:18-37 on scipy.ndimage.label: per label keep the components with more than 10 % of its voxels"""
    from smsut_b200 import config as cfg
    from smsut_b200.misc.utils import connected_components
    rng = np.random.default_rng(3)
    pred = np.zeros(shape, dtype=np.int64)
    for lab in range(1, cfg.n_label + 1):                    # blobs of very different sizes + diagonal contacts + specks
        for _ in range(4):
            c = [int(rng.integers(1, s - 1)) for s in shape]
            r = int(rng.integers(1, 4))
            sl = tuple(slice(max(x - r, 0), x + r) for x in c)
            pred[sl] = lab
        for _ in range(6):
            pred[tuple(int(rng.integers(0, s)) for s in shape)] = lab
    got = connected_components(pred)
    want = np.zeros(shape, dtype=np.uint8)
    for lab in range(1, cfg.n_modal + 1):
        comps = _flood_components(pred == lab, full=True)
        total = sum(len(c) for c in comps)
        for comp in comps:
            if len(comp) > 0.1 * total:
                for v in comp:
                    want[v] += lab
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    assert (got != 0).sum() < (pred != 0).sum()               # the specks went


def test_assd_against_brute_force_and_known_cases(pkg):
    """medpy's assd restated (border = object minus its 1-connected erosion; mean of the two directed mean distances)"""
    from smsut_b200.misc.utils import _surface_distances, assd
    a = np.zeros((20, 20), bool); a[4:10, 4:10] = True
    assert assd(a, a) == 0.0
    b = np.roll(a, 3, axis=1)                                 # the same square three pixels to the right
    rng = np.random.default_rng(5)
    c = np.zeros((6, 16, 16), bool); c[1:5, 3:12, 4:11] = True
    d = np.zeros((6, 16, 16), bool); d[2:6, 5:14, 2:9] = True; d[0, 0, 0] = True
    e = rng.uniform(size=(12, 12)) < 0.4

    def border(x):
        out = np.zeros_like(x)
        for v in zip(*np.nonzero(x)):
            for ax in range(x.ndim):
                for step in (-1, 1):
                    nb = list(v); nb[ax] += step
                    if not (0 <= nb[ax] < x.shape[ax]) or not x[tuple(nb)]:
                        # scipy's erosion treats outside-the-array as background: a voxel on the array edge is border
                        out[v] = True
        return out

    def directed(x, y):
        bx, by = np.argwhere(border(x)), np.argwhere(border(y))
        return np.mean([np.sqrt(((by - p) ** 2).sum(1)).min() for p in bx])

    for x, y in ((a, b), (c, d), (e, np.roll(e, 2, axis=0)), (e, ~e)):
        want = (directed(x, y) + directed(y, x)) / 2
        assert abs(assd(x, y) - want) < 1e-9, (assd(x, y), want)
        assert abs(assd(x, y) - assd(y, x)) < 1e-12
    assert 0 < assd(a, b) <= 3.0
    assert abs(_surface_distances(a, b).max() - 3.0) < 1e-12  # left edge of a -> left edge of b
    with pytest.raises(RuntimeError):
        assd(np.zeros((4, 4), bool), a[:4, :4] | True)
    with pytest.raises(RuntimeError):
        assd(a, np.zeros_like(a))


def test_test_entry_point_with_label_volumes_writes_dice_and_assd_blocks(exact, tmp_path, monkeypatch):
    """BaseTrainer.test the reference's way (baseTrainer.py:254-318): label volumes in, predictions assembled on the
    host, `<run>/all_trois_matrix.csv` = Dice block, empty line, ASSD block; the Dice block equals what the device-side
    confusion counts give"""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.misc.utils import connected_components, get_all_matrix
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    monkeypatch.setattr(cfg, "batch_size", 4)
    tr = UnetTrainer('test', SimpleNamespace(fold=0, expr_name=None, input_size=32))
    tr.net.load_state_dict(O.make_weights(O.unet_shapes(), 5))
    batches, gt = [], {}
    for bi, (m, pid, n) in enumerate((('ct', 'a', 4), ('t1in', 'b', 4), ('t1out', 'c', 4), ('t2', 'd', 3))):
        img, lab = O.synthetic_batch(n, 32, 700 + bi)
        batches.append((img, lab, torch.full((n,), cfg.Modality[m].value), [f"{m}_{pid}_{z}" for z in range(n)]))
        gt[f"{m}_{pid}"] = lab.numpy()
    # labels = the network's own predictions shifted by one pixel: every organ a volume predicts is in its labels too
    # (an organ absent from a label volume raises, medpy's rule: checked at the end), Dice < 1 and ASSD > 0
    _, prd = tr.validate_epoch(batches, gt)
    gt_used = {k: np.roll(v, 1, axis=2) for k, v in prd.items()}
    batches = [(img, torch.from_numpy(gt_used['_'.join(names[0].split('_')[:2])]).long(), mdl, names)
               for img, _, mdl, names in batches]          # the loader's labels ARE the label volumes
    run = os.path.join(str(tmp_path), "000")
    matrix = tr.test('inTurn', run, loader=batches, gt_npys=gt_used)
    blocks = open(os.path.join(run, "all_trois_matrix.csv")).read().split("\n\n")
    assert len(blocks) == 2
    rows = [[float(x) for x in r.split(",")] for r in blocks[0].strip().split("\n")]
    assert np.abs(np.array(rows) - matrix).max() < 5e-5 and np.array(rows).shape == (5, 5)
    device_way = tr.validate_dice()[1]
    assert np.abs(device_way - matrix).max() < 1e-12
    srows = np.array([[float(x) for x in r.split(",")] for r in blocks[1].strip().split("\n")])
    _, prd2 = tr.validate_epoch(batches, gt_used)
    assert np.abs(srows - get_all_matrix(prd2, gt_used)[2]).max() < 5e-5
    assert 0 < matrix[-1, -1] < 1 and srows[-1, -1] > 0
    k = next(iter(prd2))
    organ = next(j for j in range(1, 5) if (connected_components(prd2[k]) == j).any())
    bad = dict(gt_used)
    bad[k] = np.where(bad[k] == organ, 0, bad[k])
    with pytest.raises(RuntimeError):
        get_all_matrix(prd2, bad)
