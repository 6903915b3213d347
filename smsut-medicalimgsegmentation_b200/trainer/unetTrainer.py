# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/unetTrainer.py: supervised U-Net step (unetTrainer.py:51-85)."""
import argparse
import os
import random
import sys

if __package__ in (None, ""):      # `python trainer/unetTrainer.py -p train -f 0` from the package directory
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import __graft_entry__ as _g
    _g.load_package()
    __package__ = "smsut_b200.trainer"

import numpy as np
import torch

from .. import config as cfg
from .. import functional as Fn
from .. import ops
from ..network.unet import UNet
from ..optim import SGD, PolyLR
from .baseTrainer import BaseTrainer


class UnetTrainer(BaseTrainer):
    def __init__(self, phase, args=None):
        super(UnetTrainer, self).__init__(phase, args)
        self.parallel = None           # parallel.DataParallelContext under torchrun

    def build_network(self):
        self.net = UNet(cfg.img_channels, cfg.n_label + 1, cfg.base_width, norm_type='instance', act_type='lrelu')
        self.net.to(self.device)
        if self.phase == 'train':
            self.optimizer = SGD(self.net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
            self.lr_sched = PolyLR([self.optimizer], cfg.lr, cfg.max_epoch * cfg.num_iter_per_epoch)

    def train_step(self, img, msk):
        """One iteration of unetTrainer.py:71-85 on device tensors; returns the loss as a device scalar."""
        ops.arena_begin(img.device)
        self.lr_sched.tick()
        out = self.net(img)
        sample_loss = self.loss(out, msk)
        self.optimizer.zero_grad()
        with Fn.accumulate_param_grads():
            sample_loss.backward()
        if self.parallel is not None:
            self.parallel.all_reduce_grads(self.optimizer)
        self.optimizer.step()
        ops.arena_end()
        self.iter += 1
        return sample_loss.detach().clone()

    def train_epoch(self, lb_loader, ul_loader, meter, num_iter=None):
        self.net.train()
        lb_itr = iter(lb_loader)
        for _ in range(num_iter or cfg.num_iter_per_epoch):
            try:
                img, msk, mdl, _ = next(lb_itr)
            except StopIteration:
                lb_itr = iter(lb_loader)
                img, msk, mdl, _ = next(lb_itr)
            img = img.to(self.device, non_blocking=True)
            msk = msk.to(self.device, non_blocking=True)
            step = self.graphed('unet', self.train_step, [img, msk]) if self.graph_enabled() else None
            if step is not None:
                loss = step(img, msk)
                self.iter += 1
            else:
                loss = self.train_step(img, msk)
            self.meter_note(meter, loss, mdl[0].item(), img.size(0))      # unetTrainer.py:69-76
            for param_group in self.optimizer.param_groups:   # host mirror of the device-side schedule
                param_group['lr'] = self.optimizer._lr_host = self.lr_sched.host_lr(self.iter)
        self.meter_flush()
        return loss


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
        trainer = UnetTrainer('train', args)
        trainer.fit('inTurn', max_epoch=args.epochs, iters_per_epoch=args.iters)
    elif args.phase == 'test':
        trainer = UnetTrainer('test', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.test('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    elif args.phase == 'pseudo':
        trainer = UnetTrainer('pseudo', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.saving_pseudo('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    else:
        raise NotImplementedError
