# -*- coding: utf-8 -*-
"""Module-level constants of the hot path; same names and values as the reference's config.py (the values are
the contract: config.py:7-11 modalities, :23-33 training constants, :36-37 network, :49 input size, :57 batch,
:74-78 optimiser and NCE layers).  Dataset roots and augmentation settings are out of scope (synthetic data)."""
import os as _os
from enum import Enum


class Modality(Enum):
    ct = 0
    t1in = 1
    t1out = 2
    t2 = 3


seed = 2020
n_modal = len(Modality.__members__)
n_label = 4

num_iter_per_epoch = 150
max_epoch = 200
exp_alpha = 1.
weight_dc = 0.5
weight_ce = 0.5

img_channels = 1
base_width = 16

input_size = 256
mod_type = ('ct, t1in, t1out, t2')

# Data loader (config.py:44-72): the processed PNG dataset root ('***/bimod' in the reference: a placeholder) and the
# joint augmentation of the training loaders, applied on the GPU (data_loader/externalTransforms.py)
base_root = _os.environ.get('SMSUT_BASE_ROOT', '***/bimod')
png_root = base_root
split_yaml = 'semi-1910.yaml'
batch_size = 8
num_workers = 6
data_aug = dict(
    rotate=True,
    rotate_degrees=15,
    resizeCrop=True,
    resizeCrop_size=input_size,
    elasticDeform=True,
    elasticDeform_sigmas=(9., 13.),
    elasticDeform_points=3,
    colorJitter=False,
    gammaCorrect=False,
    gammaCorrect_gammas=(0.7, 1.5),
)

lr = 1e-2
weight_decay = 1e-3

nce_layers = [5]

# coraNet (config.py:80-95).  The reference ships the 2-class SAML vectors (`default_w = [1, 1]`, `w_con = [1, 5]`,
# `w_rad = [5, 1]`) next to n_label = 4, for which nn.CrossEntropyLoss(weight) raises on the 5-class heads; the CHAOS
# vectors of its comments are the ones consistent with n_label = 4.  Plain lists here: the trainer puts them on its
# device.
thres = 0.5
default_w = [1., 1., 1., 1., 1.]
w_con = [1., 5., 5., 5., 5.]
w_rad = [5., 1., 1., 1., 1.]
pre_epoch = 100
cora_epoch = 200
pred_step = 10

expr_root = _os.environ.get('SMSUT_EXPR_ROOT', './smsut-out')      # the reference's expr_root (config.py:46); env override for tests
