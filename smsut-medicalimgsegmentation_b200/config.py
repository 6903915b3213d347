# -*- coding: utf-8 -*-
"""Module-level constants of the hot path; same names and values as the reference's config.py (the values are
the contract: config.py:7-11 modalities, :23-33 training constants, :36-37 network, :49 input size, :57 batch,
:74-78 optimiser and NCE layers).  Raw-dataset roots and the resampling spacing of the offline pre-processing are out of
scope."""
import os as _os
from enum import Enum

# modality codes of the translation input (config.py:7-11); member order = code
Modality = Enum('Modality', ['ct', 't1in', 't1out', 't2'], start=0)
n_modal = len(Modality)
mod_type = 'ct, t1in, t1out, t2'
n_label = 4                      # foreground organs (CHAOS); the networks have n_label + 1 output channels
seed = 2020

# training schedule and loss weights (config.py:23-33)
num_iter_per_epoch, max_epoch = 150, 200
exp_alpha = 1.0
weight_dc = weight_ce = 0.5

# network (config.py:36-37) and slice size (config.py:49)
img_channels, base_width = 1, 16
input_size = 256

# data loader (config.py:44-72): the processed PNG dataset root ('***/bimod' in the reference: a placeholder) and the
# joint augmentation of the training loaders, applied on the GPU (data_loader/externalTransforms.py)
base_root = png_root = _os.environ.get('SMSUT_BASE_ROOT', '***/bimod')
expr_root = _os.environ.get('SMSUT_EXPR_ROOT', './smsut-out')      # output root (config.py:46); env override for tests
split_yaml = 'semi-1910.yaml'
batch_size, num_workers = 8, 6
data_aug = {
    'rotate': True, 'rotate_degrees': 15,
    'resizeCrop': True, 'resizeCrop_size': input_size,
    'elasticDeform': True, 'elasticDeform_sigmas': (9., 13.), 'elasticDeform_points': 3,
    'colorJitter': False,
    'gammaCorrect': False, 'gammaCorrect_gammas': (0.7, 1.5),
}

# optimiser (config.py:74-75) and the encoder layers PatchNCE samples (config.py:78)
lr, weight_decay = 1e-2, 1e-3
nce_layers = [5]

# coraNet (config.py:80-95).  The reference ships the 2-class SAML vectors (`default_w = [1, 1]`, `w_con = [1, 5]`,
# `w_rad = [5, 1]`) next to n_label = 4, for which nn.CrossEntropyLoss(weight) raises on the 5-class heads; the CHAOS
# vectors of its comments are the ones consistent with n_label = 4.  Plain lists here: the trainer puts them on its
# device.
thres = 0.5
default_w = [1.] * (n_label + 1)
w_con = [1.] + [5.] * n_label
w_rad = [5.] + [1.] * n_label
pre_epoch, cora_epoch, pred_step = 100, 200, 10
