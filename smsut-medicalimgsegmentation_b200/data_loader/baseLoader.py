# -*- coding: utf-8 -*-
"""Counterpart of the reference's data_loader/baseLoader.py (and balanceLoader.py's BalanceDataset) for the GPU input
pipeline: the PNG slice tree  <root>/<modality>/<patient>/{images,labels}/<modality>_<patient>_<z>.png  with the
split file cfg.split_yaml (baseLoader.py:31-48, balanceLoader.py:31-54) is decoded ONCE into two u8 arrays that live
in HBM; a batch is the sampler's index draw + the parameter draws of externalTransforms.py + ONE launch
(ops.augment_batch).  The reference decodes / augments per item in 6 DataLoader workers (baseLoader.py:50-60).

Loaders yield what the reference's do -- (img fp32 (B,1,H,W) in [-1,1], msk int64 (B,H,W), mdl int64 (B,), names) --
with img / msk already on the device (the trainers' `.to(device)` is then a no-op); mdl stays on the host like the
reference's collated tensor (label2onehot indexes with it on the CPU)."""
import os
import random
from os.path import join as pjoin

import numpy as np
import torch

from .. import config as cfg
from .. import ops
from . import externalTransforms as extt


def read_yaml(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


class BaseDataset(object):
    """All slices of the requested modalities / phase / fold, decoded to u8 (baseLoader.py:16-60).  `modal`: 'all' or
    one modality name.  modal_sample_ids[m] lists the dataset indices of modality m (balanceLoader.py:33-53)."""

    def __init__(self, data_root, phase, modal='all', fold=0, load_in_ram=True, joint_transform=None, device=None):
        from PIL import Image
        self.data_root, self.phase, self.fold = data_root, phase, fold
        self.modal = list(cfg.Modality.__members__) if modal == 'all' else [modal]
        self.joint_transform = joint_transform
        split = read_yaml(pjoin(data_root, getattr(cfg, 'split_yaml', 'semi-1910.yaml')))
        imgs, msks, self.modalities, self.names = [], [], [], []
        self.modal_sample_ids = [[] for _ in cfg.Modality.__members__]
        for m in self.modal:
            modal_root = pjoin(data_root, m)
            pids = split[m][phase] if phase == 'test' else split[m][phase][fold]
            for pid in pids:
                pid_root = pjoin(modal_root, str(pid), 'images')
                for png in sorted(os.listdir(pid_root)):
                    img = pjoin(pid_root, png)           # eg. /path/to/ct/001/images/ct_001_000.png
                    imgs.append(np.asarray(Image.open(img).convert('L'), dtype=np.uint8))
                    msks.append(np.asarray(Image.open(img.replace('images', 'labels')), dtype=np.uint8))
                    self.modal_sample_ids[cfg.Modality[m].value].append(len(self.names))
                    self.modalities.append(cfg.Modality[m].value)
                    self.names.append(png.replace('.png', ''))
        if not imgs:
            raise RuntimeError(f'no slices under {data_root} for phase {phase!r}, fold {fold}')
        shapes = {a.shape for a in imgs} | {a.shape for a in msks}
        if len(shapes) != 1:
            raise NotImplementedError(f'all slices must share one size (found {sorted(shapes)})')
        self.h, self.w = imgs[0].shape
        dev = device if device is not None else ('cuda' if torch.cuda.is_available() else 'cpu')
        self.images = torch.from_numpy(np.stack(imgs)).to(dev)        # (n, h, w) u8, resident
        self.labels = torch.from_numpy(np.stack(msks)).to(dev)
        self.device = torch.device(dev)

    def __len__(self):
        return len(self.names)

    def __repr__(self):
        return self.__class__.__name__ + '(samples={0}, phase={1} {2}, modality={3})'.format(
            len(self), self.phase, self.fold, self.modal)


BalanceDataset = BaseDataset        # balanceLoader.py's dataset is the same tree plus modal_sample_ids


class GpuBatchLoader(object):
    """Iterable over (img, msk, mdl, names) batches: `batch_sampler` yields lists of dataset indices."""

    def __init__(self, dataset, batch_sampler):
        self.dataset, self.batch_sampler = dataset, batch_sampler
        self._params = None

    def __len__(self):
        return len(self.batch_sampler)

    def __iter__(self):
        ds = self.dataset
        for idx in self.batch_sampler:
            n = len(idx)
            params = torch.zeros((n, extt.PARAM_FLOATS), dtype=torch.float32)
            if ds.joint_transform is not None:
                view = params.numpy()
                for r in range(n):
                    extt.pack_params(ds.joint_transform.draw(ds.h, ds.w), view[r])
            index = torch.tensor(idx, dtype=torch.int64)
            if ds.device.type == 'cuda':
                params = params.pin_memory().to(ds.device, non_blocking=True)
                index = index.pin_memory().to(ds.device, non_blocking=True)
            img, msk = ops.augment_batch(ds.images, ds.labels, index, params)
            mdl = torch.tensor([ds.modalities[i] for i in idx], dtype=torch.int64)
            yield img, msk, mdl, [ds.names[i] for i in idx]


class _ShuffledBatches(object):
    """DataLoader(batch_size, shuffle, drop_last) as an index sampler (baseLoader.py:81-82)"""

    def __init__(self, n, batch_size, shuffle, drop_last):
        self.n, self.batch_size, self.shuffle, self.drop_last = n, batch_size, shuffle, drop_last

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else -(-self.n // self.batch_size)

    def __iter__(self):
        order = list(range(self.n))
        if self.shuffle:
            random.shuffle(order)
        for i in range(0, self.n, self.batch_size):
            b = order[i:i + self.batch_size]
            if len(b) == self.batch_size or not self.drop_last:
                yield b


def get_loader(data_root, phase, fold, batch_size, data_aug=None, modal='all', load_in_ram=True, device=None):
    """baseLoader.py:71-84 (argument order of the trainers' call: root, phase, fold, batch size, augmentation)"""
    joint_augs = parse_aug(data_aug) if phase in ('train', 'val') else None
    dataset = BaseDataset(data_root, phase, modal, fold, load_in_ram, joint_augs, device)
    print(dataset)
    return GpuBatchLoader(dataset, _ShuffledBatches(len(dataset), batch_size, phase == 'train', phase == 'train'))


def parse_aug(data_aug):
    """baseLoader.py:87-112 -> ONE JointCompose holding the joint and the image-only draws (None = normalise only).
    colorJitter has no counterpart here (off in the reference's config.py:63)."""
    if data_aug is None or not data_aug:
        return None
    augs = []
    if data_aug.get('rotate'):
        augs.append(extt.JointRotate(data_aug['rotate_degrees']))
    if data_aug.get('elasticDeform'):
        augs.append(extt.JointElasticDeform(data_aug['elasticDeform_sigmas'], data_aug['elasticDeform_points']))
    if data_aug.get('resizeCrop'):
        augs.append(extt.JointRandomResizedCrop(data_aug['resizeCrop_size']))
    if data_aug.get('colorJitter'):
        raise NotImplementedError("colorJitter is off in the reference's config (config.py:63) and not built")
    if data_aug.get('gammaCorrect'):
        augs.append(extt.RandomGammaCorrection(data_aug['gammaCorrect_gammas']))
    return extt.JointCompose(augs)
