# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/uganTrainer.py: the UGAN generator without the
PatchNCE head (network/ugan.py:84-124), trained on labelled slices with the shape loss
lambda_shp * Dice+CE(seg(G(G(x))), y) (build_network :51-66, train_epoch :115-215; lambda_shp ramps with the epoch,
:122-123).  The iteration itself is UGANShp0Trainer.shape_train_step -- the same kernels as the headline path."""
import argparse
import os
import random
import sys

if __package__ in (None, ""):      # `python trainer/uganTrainer.py -p train -f 0` from the package directory
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import __graft_entry__ as _g
    _g.load_package()
    __package__ = "smsut_b200.trainer"

import numpy as np
import torch

from .. import config as cfg
from ..network.ugan import UGAN, Discriminator
from ..optim import SGD, Adam, PolyLR
from .uganShp0Trainer import UGANShp0Trainer


class UGANTrainer(UGANShp0Trainer):
    def __init__(self, phase, args):
        self.lambda_shp = 10
        self.lambda_shp_lazy = 20
        self.parallel = None           # parallel.DataParallelContext under torchrun
        super(UGANTrainer, self).__init__(phase, args)

    def build_network(self):
        self.net = UGAN(cfg.img_channels, cfg.n_label + 1, cfg.n_modal, cfg.base_width)
        self.net.to(self.device)
        self.D = Discriminator(self.input_size, cfg.n_modal, cfg.base_width,
                               max_width=256 if cfg.base_width == 16 else 512)
        self.D.to(self.device)
        if self.phase == 'train':
            beta1, beta2 = self.beta1, self.beta2
            self.optimizer = SGD(self.net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
            self.d_optimizer = Adam(self.D.parameters(), cfg.lr, [beta1, beta2], weight_decay=cfg.weight_decay)
            self.lr_sched = PolyLR([self.optimizer, self.d_optimizer], cfg.lr, cfg.max_epoch * cfg.num_iter_per_epoch)

    def translate(self, x, m):
        return self.net(x, m)

    def segment(self, img):
        seg, _ = self.net(img)
        return seg

    def epoch_lambda_shp(self):
        # uganTrainer.py:122-123
        lambda_shp = self.epoch * (self.lambda_shp / self.lambda_shp_lazy)
        return float(min(lambda_shp, self.lambda_seg))


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('-p', '--phase', type=str, default='train')
    parser.add_argument('-f', '--fold', type=int, default=0)
    parser.add_argument('-nm', '--expr_name', type=str, default=None)
    parser.add_argument('-i', '--model_id', type=str, default=None)
    parser.add_argument('-wh', '--which_ckpt', type=str, default='last')
    parser.add_argument('--epochs', type=int, default=None, help='(extension) shorten the run')
    parser.add_argument('--iters', type=int, default=None, help='(extension) iterations per epoch')
    args = parser.parse_args()
    random.seed(cfg.seed); np.random.seed(cfg.seed)
    torch.manual_seed(cfg.seed); torch.cuda.manual_seed(cfg.seed)
    if args.phase == 'train':
        trainer = UGANTrainer('train', args)
        trainer.fit('inTurn', max_epoch=args.epochs, iters_per_epoch=args.iters)
    elif args.phase == 'test':
        trainer = UGANTrainer('test', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.test('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    elif args.phase == 'pseudo':
        trainer = UGANTrainer('pseudo', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.saving_pseudo('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    else:
        raise NotImplementedError
