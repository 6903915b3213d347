# -*- coding: utf-8 -*-
"""Counterpart of the reference's data_loader/balanceLoader.py: batches that hold the same number of slices of every
modality.  No trainer of the reference builds this loader (they import the module and use inTurnLoader / baseLoader,
baseTrainer.py:128-135); it is here so that `from data_loader import balanceLoader` keeps working, on top of the
GPU-resident dataset of baseLoader.py."""
import random
from typing import List

from .. import config as cfg
from . import baseLoader as bslod

BalanceDataset = bslod.BalanceDataset


class ModalityBalanceBatchSampler(object):
    """balanceLoader.py:80-109.  samples[m]: dataset indices of modality m, each list shuffled once.  A batch takes the
    next batch_size / num_modality indices of EVERY modality; a modality whose cursor has moved past the end of its
    list is reshuffled and starts over (the batch that ran into the end came up short and is dropped: only full
    batches are yielded).  One pass makes ceil(longest list / per-modality share) attempts."""

    def __init__(self, samples: List[List[int]], batch_size: int):
        self.samples = samples
        self.num_modality = len(samples)
        self.batch_size = batch_size
        self.num_samples_per_modality = batch_size // self.num_modality
        self.starts = [0] * self.num_modality
        self.n = max([len(spl) for spl in samples] + [0])
        for spl in self.samples:
            random.shuffle(spl)

    def __iter__(self):
        share = self.num_samples_per_modality
        for _ in range(0, self.n, share):
            batch = []
            for j, spl in enumerate(self.samples):
                s = self.starts[j]
                batch += spl[s:s + share]
                self.starts[j] = s + share
                if self.starts[j] > len(spl):
                    random.shuffle(spl)
                    self.starts[j] = 0
            if len(batch) == self.batch_size:
                yield batch

    def __len__(self):
        return self.n // self.num_samples_per_modality


def get_loader(data_root, phase, fold, batch_size, data_aug=None, load_in_ram: bool = True, device=None):
    """balanceLoader.py:112-125: training / validation splits only"""
    if phase not in ('train', 'val'):
        raise ValueError(phase)
    if batch_size % cfg.n_modal != 0:
        raise AssertionError('Batch size must be an integral multiple of #modality.')
    dataset = BalanceDataset(data_root, phase, 'all', fold, load_in_ram, bslod.parse_aug(data_aug), device)
    print(dataset)
    return bslod.GpuBatchLoader(dataset, ModalityBalanceBatchSampler(dataset.modal_sample_ids, batch_size))
