# -*- coding: utf-8 -*-
"""Counterpart of the host-side helpers of the reference's misc/utils.py that the hot path's callers use: the
directory / yaml helpers (utils.py:40-55), the epoch `Meter` of `BaseTrainer.fit` (utils.py:58-160), the label-volume
loader (utils.py:163-177) and the modality-organ Dice matrix (utils.py:180-203).

`connected_components` / `get_all_matrix` (utils.py:18-37, 206-279), the CPU post-processing and surface-distance
extras of `-p test`, are restated on scipy.ndimage (skimage / medpy are not importable offline and the reference pins
no version: those two are checked against brute-force statements, not against the libraries).  The trainers compute the
Dice matrix from device-side confusion counts instead (`BaseTrainer.validate_dice`)."""
import os
from collections import OrderedDict
from os.path import join as pjoin

import numpy as np
import torch
import yaml

from .. import config as cfg


def maybe_mkdir(*paths):
    """create each directory that does not exist yet (parents must exist, like os.mkdir: utils.py:40-44)"""
    for p in paths:
        if not os.path.exists(p):
            os.mkdir(p)


def read_yaml(path):
    with open(path, 'r') as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def write_yaml(data, path):
    with open(path, 'w') as f:
        yaml.dump(data, f)


class Meter(object):
    """Per-epoch bookkeeping of `fit` (baseTrainer.py:147-199).  Keys are registered as 'min' (losses) or 'max' (Dice)
    better; within an epoch `accumulate` adds weighted sums and weights, `update_cur` turns them into means, smooths
    them against the previous epoch with `alpha` (1 = no smoothing) and tracks the best value per key.

    Same attributes as the reference's class (`configs`, `cur_values`, `best_values`, `pre_values`, `n`): the trainers
    read `cur_values['dice'] >= best_values['dice']` to decide on the `best` checkpoint."""

    def __init__(self, min_better_keys, max_better_keys, alpha=1.):
        self.alpha = alpha
        self.configs = OrderedDict([(k, 'min') for k in min_better_keys] + [(k, 'max') for k in max_better_keys])
        self.cur_values = self.get_empty_dict()
        self.n = self.get_empty_dict()
        self.best_values = self.get_empty_dict()
        self.pre_values = None          # None until the first update_cur: that epoch seeds `best` and `pre`

    def get_empty_dict(self):
        return dict.fromkeys(self.configs, 0)

    def reset_cur(self):
        self.cur_values = self.get_empty_dict()
        self.n = self.get_empty_dict()

    def accumulate(self, values, n):
        """values[k]: weighted sum to add to key k, n[k]: its weight"""
        for k, v in values.items():
            self.cur_values[k] += v
            self.n[k] += n[k]

    def update_cur(self, reset_best=False):
        first = self.pre_values is None
        for k in self.configs:
            v = self.cur_values[k]
            if self.n[k] != 0:
                v = v / self.n[k]
            if not first:
                v = (1. - self.alpha) * self.pre_values[k] + self.alpha * v
            self.cur_values[k] = v
        if first or reset_best:
            self.best_values = dict(self.cur_values)
            self.pre_values = dict(self.cur_values)
            return
        for k, direction in self.configs.items():
            v = self.cur_values[k]
            if (direction == 'min' and v < self.best_values[k]) or (direction == 'max' and v > self.best_values[k]):
                self.best_values[k] = v
            self.pre_values[k] = v

    @staticmethod
    def collect_loss_by(sample_loss, modal_id, n):
        """a batch mean `sample_loss` over `n` slices of modality `modal_id` -> (sums, weights) for accumulate():
        counted under the overall key 'loss' and under the modality's own 'loss_<id>'"""
        k = 'loss_' + str(modal_id)
        total = sample_loss * n
        return {'loss': total, k: total}, {'loss': n, k: n}

    @staticmethod
    def collect_dice_by(output, gt, modal_idxs, n_modal, smooth=1e-5):
        """output: (B, C, H, W) logits, gt: (B, H, W) labels, modal_idxs: (B,) modality of each slice.  Per slice the
        foreground-mean Dice of the argmax mask, (2 tp + s) / (2 tp + fp + fn + s); summed per modality (values) with
        the slice counts (weights).  Runs on whatever device `output` lives on (utils.py:119-150 assumes CUDA)."""
        c = output.shape[1]
        pred = torch.argmax(output, dim=1)
        classes = torch.arange(c, device=output.device).view(1, c, 1, 1)
        p = pred.unsqueeze(1) == classes
        g = gt.to(output.device).long().unsqueeze(1) == classes
        tp = (p & g).flatten(2).sum(-1).double()
        size_p, size_g = p.flatten(2).sum(-1).double(), g.flatten(2).sum(-1).double()
        # fp = |p| - tp, fn = |g| - tp
        dice = (2 * tp + smooth) / (size_p + size_g + smooth)
        per_slice = (dice[:, 1:].sum(dim=1) / (c - 1)).tolist()
        sums, counts = [0.] * n_modal, [0] * n_modal
        for d, mi in zip(per_slice, torch.as_tensor(modal_idxs).tolist()):
            sums[int(mi)] += d
            counts[int(mi)] += 1
        a = {f'dice_{i}': sums[i] for i in range(n_modal)}
        b = {f'dice_{i}': counts[i] for i in range(n_modal)}
        a['dice'], b['dice'] = sum(sums), sum(counts)
        return a, b

    @staticmethod
    def display_key(k):
        """'loss_1' -> 'loss_t1in' (the log / TensorBoard tag of baseTrainer.py:165-170)"""
        if '_' in k:
            typ, m = k.split('_')
            return f'{typ}_{cfg.Modality(int(m)).name}'
        return k

    def __repr__(self):
        return ''.join(' %s: %.4f/%.4f,' % (self.display_key(k), self.cur_values[k], self.best_values[k])
                       for k in self.configs)


def get_label_npys(png_root, modal, phase):
    """(number of slices, {'<modality>_<patient>': (Z, H, W) label volume}) of a split: the `<m>/<p>/<m>_<p>.npy`
    files the pre-processing writes next to the PNG slices (utils.py:163-177)"""
    split = read_yaml(pjoin(png_root, cfg.split_yaml))
    modal = list(cfg.Modality.__members__) if modal == 'all' else [modal]
    volumes, n = {}, 0
    for m in modal:
        for p in split[m][phase]:
            vol = np.load(pjoin(png_root, m, p, f'{m}_{p}.npy'))
            volumes[f'{m}_{p}'] = vol
            n += vol.shape[0]
    return n, volumes


def dice_coefficient(p, g):
    """medpy.metric.dc on boolean arrays: 2 |p & g| / (|p| + |g|), 0 when both are empty"""
    p, g = np.asarray(p, dtype=bool), np.asarray(g, dtype=bool)
    denom = int(p.sum()) + int(g.sum())
    return 2.0 * int((p & g).sum()) / denom if denom > 0 else 0.0


def get_mo_matrix(prd_npys, gt_npys):
    """Modality-organ Dice matrix from host volumes (utils.py:180-203): Dice per volume and organ, averaged over the
    volumes of a modality; last row / column = means over modalities / organs.  `BaseTrainer.validate_dice` computes
    the same matrix from per-volume confusion counts that never leave the device."""
    n_modal, n_label = cfg.n_modal, cfg.n_label
    matrix = np.zeros((n_modal + 1, n_label + 1))
    n = np.zeros((n_modal, 1))
    for k, g in gt_npys.items():
        m = cfg.Modality[k.split('_')[0]].value
        p = prd_npys[k]
        for j in range(1, n_label + 1):
            matrix[m, j - 1] += dice_coefficient(p == j, g == j)
        n[m] += 1
    n[n == 0] += 1e-8
    matrix[:n_modal, :n_label] /= n
    matrix[-1, :] = matrix[:n_modal].mean(axis=0)
    matrix[:, -1] = matrix[:, :n_label].mean(axis=1)
    return matrix


def connected_components(pred):
    """The clean-up `-p test` applies to a predicted label map before the surface metric (utils.py:18-37): per label
    1..cfg.n_modal (sic: the reference loops over the modality count, which equals the organ count on CHAOS), keep the
    connected components that hold more than 10 % of that label's voxels.  skimage.measure.label(connectivity=2) is
    restated with scipy.ndimage.label: neighbours within squared distance 2 (8-connected in 2-D, 18-connected in 3-D).
    skimage is not importable offline, so this restatement is checked against a flood fill, not against skimage."""
    from scipy import ndimage
    pred = np.asarray(pred)
    structure = ndimage.generate_binary_structure(pred.ndim, min(2, pred.ndim))
    out = np.zeros_like(pred)
    for i in range(cfg.n_modal):
        labels, num = ndimage.label(pred == i + 1, structure=structure)
        if num == 0:
            continue
        sizes = np.bincount(labels.ravel(), minlength=num + 1)
        keep = sizes > 0.1 * sizes[1:].sum()
        keep[0] = False
        out += (keep[labels] * (i + 1)).astype(out.dtype)
    return np.uint8(out)


def _surface_distances(result, reference, voxelspacing=None, connectivity=1):
    """distances from the border voxels of `result` to the nearest border voxel of `reference` (medpy.metric.binary:
    border = object minus its erosion, Euclidean distance transform of the complement of the reference border)"""
    from scipy import ndimage
    result, reference = np.atleast_1d(np.asarray(result).astype(bool)), np.atleast_1d(np.asarray(reference).astype(bool))
    if not result.any():
        raise RuntimeError('The first supplied array does not contain any binary object.')
    if not reference.any():
        raise RuntimeError('The second supplied array does not contain any binary object.')
    footprint = ndimage.generate_binary_structure(result.ndim, connectivity)
    result_border = result ^ ndimage.binary_erosion(result, structure=footprint, iterations=1)
    reference_border = reference ^ ndimage.binary_erosion(reference, structure=footprint, iterations=1)
    dt = ndimage.distance_transform_edt(~reference_border, sampling=voxelspacing)
    return dt[result_border]


def assd(result, reference, voxelspacing=None, connectivity=1):
    """medpy.metric.assd restated from its documentation (medpy is not importable offline; no version is pinned by the
    reference): the mean of the two directed average surface distances"""
    return float(np.mean((_surface_distances(result, reference, voxelspacing, connectivity).mean(),
                          _surface_distances(reference, result, voxelspacing, connectivity).mean())))


def _with_means(matrix):
    n_modal, n_label = matrix.shape
    full = np.zeros((n_modal + 1, n_label + 1))
    full[:n_modal, :n_label] = matrix
    full[-1, :] = full[:n_modal].mean(axis=0)
    full[:, -1] = full[:, :n_label].mean(axis=1)
    return full


def get_all_matrix(prd_npys, gt_npys):
    """(Dice, "Hausdorff", ASSD) modality-organ matrices of `-p test` (utils.py:206-279) on predictions cleaned by
    connected_components (the volume, then every slice).  As in the reference the second matrix repeats the Dice
    values (`t = s`), an organ missing from the prediction scores the largest ASSD seen so far in its volume, and an
    organ missing from the LABELS of a volume raises (medpy's RuntimeError).  CPU evaluation: nothing here runs on the
    device."""
    n_modal, n_label = cfg.n_modal, cfg.n_label
    dice, hd, sd = (np.zeros((n_modal, n_label)) for _ in range(3))
    n = np.zeros((n_modal, 1))
    for k, g in gt_npys.items():
        m = cfg.Modality[k.split('_')[0]].value
        p = connected_components(prd_npys[k])
        for z in range(p.shape[0]):
            p[z] = connected_components(p[z])
        worst = 0
        for j in range(1, n_label + 1):
            predx, gx = p == j, np.asarray(g) == j
            s = dice_coefficient(predx, gx)
            r = worst if not predx.any() else assd(predx, gx)
            worst = max(worst, r)
            dice[m, j - 1] += s
            hd[m, j - 1] += s
            sd[m, j - 1] += r
        n[m] += 1
    n[n == 0] += 1e-8
    return _with_means(dice / n), _with_means(hd / n), _with_means(sd / n)
