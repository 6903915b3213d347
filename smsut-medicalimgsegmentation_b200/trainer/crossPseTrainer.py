# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/crossPseTrainer.py (cross pseudo supervision):
two U-Nets trained on the same labelled + unlabelled slices, each supervised on the unlabelled half by the other
one's argmax (build_network :45-58, train_epoch :74-146).  No new kernels: the pseudo-label argmax is taken inside
the fused Dice/CE kernel (misc/loss.py), and the second network runs on a branch stream beside the first."""
import argparse
import os
import random
import sys

if __package__ in (None, ""):      # `python trainer/crossPseTrainer.py -p train -f 0` from the package directory
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


class crossPseTrainer(BaseTrainer):
    def __init__(self, phase, args=None):
        self.parallel = None           # parallel.DataParallelContext under torchrun
        super(crossPseTrainer, self).__init__(phase, args)
        self.lambda_semi = 0.1
        self.log_step = 50

    def build_network(self):
        self.net = UNet(cfg.img_channels, cfg.n_label + 1, cfg.base_width, norm_type='instance', act_type='lrelu')
        self.net2 = UNet(cfg.img_channels, cfg.n_label + 1, cfg.base_width, norm_type='instance', act_type='lrelu')
        self.net.to(self.device)
        self.net2.to(self.device)
        if self.phase == 'train':
            self.optimizer1 = SGD(self.net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
            self.optimizer2 = SGD(self.net2.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
            self.lr_sched = PolyLR([self.optimizer1, self.optimizer2], cfg.lr, cfg.max_epoch * cfg.num_iter_per_epoch)

    def train_step(self, img, msk, lambda_semi):
        """One iteration of crossPseTrainer.py:96-131 on device tensors: img = cat(labelled, unlabelled)
        (2*bs,1,H,W), msk (bs,H,W).  Returns (seg1, seg2, semi1, semi2) as one device vector."""
        bs = msk.shape[0]
        if isinstance(lambda_semi, torch.Tensor):
            lambda_semi = lambda_semi.reshape(())      # device scalar: one captured graph serves every epoch
        ops.arena_begin(img.device)
        self.lr_sched.tick()
        with ops.parallel_branch(5) as b2:          # the two networks share nothing until the cross losses
            out2 = self.net2(img)
            sample2_loss = self.loss(out2[:bs], msk)
        out1 = self.net(img)
        sample1_loss = self.loss(out1[:bs], msk)
        b2.join(out2, sample2_loss)
        # self.loss(out_a[bs:], argmax(out_b[bs:]).detach()): the argmax happens inside the loss kernel
        semi1_loss = self.loss(out1[bs:], out2[bs:].detach())
        semi2_loss = self.loss(out2[bs:], out1[bs:].detach())
        total_loss = sample1_loss + sample2_loss + lambda_semi * semi1_loss + lambda_semi * semi2_loss
        self.optimizer1.zero_grad()
        self.optimizer2.zero_grad()
        with Fn.accumulate_param_grads():
            total_loss.backward()
        if self.parallel is not None:
            self.parallel.all_reduce_grads(self.optimizer1)
            self.parallel.all_reduce_grads(self.optimizer2)
        self.optimizer1.step()
        self.optimizer2.step()
        ops.arena_end()
        self.iter += 1
        return torch.stack([sample1_loss.detach(), sample2_loss.detach(), semi1_loss.detach(), semi2_loss.detach()])

    def train_epoch(self, lb_loader, ul_loader, meter, num_iter=None):
        self.net.train()
        self.net2.train()
        lb_itr = iter(lb_loader)
        ul_itr = iter(ul_loader)
        lambda_semi = self.lambda_semi * self.sigmoid_rampup(self.epoch, cfg.max_epoch)
        losses = None
        lam_dev = torch.zeros(1, device=self.device)
        for i in range(num_iter or cfg.num_iter_per_epoch):
            try:
                img1, msk, mdl1, _ = next(lb_itr)
            except StopIteration:
                lb_itr = iter(lb_loader)
                img1, msk, mdl1, _ = next(lb_itr)
            try:
                img2, _, mdl2, _ = next(ul_itr)
            except StopIteration:
                ul_itr = iter(ul_loader)
                img2, _, mdl2, _ = next(ul_itr)
            img = torch.cat([img1, img2], dim=0).to(self.device, non_blocking=True)
            msk = msk.to(self.device, non_blocking=True)
            step = None
            if self.graph_enabled():
                lam_dev.fill_(float(lambda_semi))
                step = self.graphed('cross_pse', self.train_step, [img, msk, lam_dev])
            if step is not None:
                losses = step(img, msk, lam_dev)
                self.iter += 1
            else:
                losses = self.train_step(img, msk, lambda_semi)
            # crossPseTrainer.py:106-119: both networks' supervised losses go to the same meter
            self.meter_note(meter, losses[0], mdl1[0].item(), img.size(0))
            self.meter_note(meter, losses[1], mdl1[0].item(), img.size(0))
            if (i + 1) % self.log_step == 0:
                s1, s2, c1, c2 = losses.tolist()
                self.info('Iter %d, global_iter: %d, crossPse1_loss: %.4f, crossPse2_loss: %.4f, '
                          'seg1_loss: %.4f, seg2_loss: %.4f, lambda_semi: %f' % (i, self.iter, c1, c2, s1, s2, lambda_semi))
            lr_ = self.lr_sched.host_lr(self.iter)
            for opt in (self.optimizer1, self.optimizer2):
                for param_group in opt.param_groups:
                    param_group['lr'] = opt._lr_host = lr_
        self.meter_flush()
        return losses


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
        trainer = crossPseTrainer('train', args)
        trainer.fit('inTurn', max_epoch=args.epochs, iters_per_epoch=args.iters)
    elif args.phase == 'test':
        trainer = crossPseTrainer('test', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.test('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    elif args.phase == 'pseudo':
        trainer = crossPseTrainer('pseudo', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.saving_pseudo('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    else:
        raise NotImplementedError
