# -*- coding: utf-8 -*-
"""Counterpart of the reference's data_loader/inTurnLoader.py: batches that hold ONE modality each, the modalities
taking turns (the UGAN trainers rely on it: one modal_org per half batch, uganConsisTrainer.py:110-127), on top of the
GPU-resident dataset of baseLoader.py."""
import random
from typing import List

from . import baseLoader as bslod


class InTurnTrainBatchSampler(object):
    """inTurnLoader.py:15-60.  samples[m]: dataset indices of modality m.  Every modality's list is shuffled once;
    modalities are visited round robin (in a reshuffled order per round when `shuffle`); a modality whose next batch
    would run past its list starts over with a fresh shuffle; only full batches are yielded.
    len = num_modality * max_m(full batches of m, minus one when its list does not divide evenly)."""

    def __init__(self, samples: List[List[int]], batch_size: int, shuffle: bool):
        self.samples = samples
        self.num_modality = len(samples)
        self.batch_size = batch_size
        self.starts = [0 for _ in range(self.num_modality)]
        self.shuffle = shuffle
        self.queue = [i for i in range(self.num_modality)]
        self.cur_modality = 0
        most = 0
        for i, spl in enumerate(self.samples):
            n = len(spl) // batch_size - 1 if len(spl) % batch_size else len(spl) // batch_size
            most = max(n, most)
            random.shuffle(self.samples[i])
        self.n = self.num_modality * most

    def __iter__(self):
        for _ in range(self.n):
            cur = self.queue[self.cur_modality] if self.shuffle else self.cur_modality
            s = self.starts[cur]
            if s + self.batch_size >= len(self.samples[cur]):
                self.starts[cur] = 0
                s = 0
                random.shuffle(self.samples[cur])
            else:
                self.starts[cur] += self.batch_size
            batch = self.samples[cur][s:s + self.batch_size]
            if len(batch) == self.batch_size:
                yield batch
            if self.shuffle and self.cur_modality + 1 == self.num_modality:
                random.shuffle(self.queue)
            self.cur_modality = (self.cur_modality + 1) % self.num_modality

    def __len__(self):
        return self.n


class InTurnTestBatchSampler(object):
    """inTurnLoader.py:63-80: modality after modality in dataset order, the last batch of a modality may be ragged"""

    def __init__(self, samples: List[List[int]], batch_size: int):
        self.samples, self.num_modality, self.batch_size = samples, len(samples), batch_size
        self.n = sum(len(spl) // batch_size for spl in samples)

    def __iter__(self):
        for spl in self.samples:
            for i in range(0, len(spl), self.batch_size):
                yield spl[i:i + self.batch_size]

    def __len__(self):
        return self.n


def get_loader(data_root, phase, fold, batch_size, data_aug=None, load_in_ram: bool = True, device=None):
    """inTurnLoader.py:83-97"""
    if phase == 'train' or phase == 'val':
        dataset = bslod.BalanceDataset(data_root, phase, 'all', fold, load_in_ram, bslod.parse_aug(data_aug), device)
        sampler = InTurnTrainBatchSampler(dataset.modal_sample_ids, batch_size, shuffle=False)
    else:
        dataset = bslod.BalanceDataset(data_root, phase, 'all', fold, load_in_ram, None, device)
        sampler = InTurnTestBatchSampler(dataset.modal_sample_ids, batch_size)
    print(dataset)
    return bslod.GpuBatchLoader(dataset, sampler)
