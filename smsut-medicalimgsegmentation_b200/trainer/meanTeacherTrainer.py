# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/meanTeacherTrainer.py: semi-supervised U-Net with an EMA teacher
(build_network :48-61, update_ema_variable :63-69, train_epoch :71-153)."""
import argparse
import os
import random
import sys

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import __graft_entry__ as _g
    _g.load_package()
    __package__ = "smsut_b200.trainer"

import numpy as np
import torch

from .. import config as cfg
from .. import functional as Fn
from .. import ops
from ..misc.loss import _flat_logits
from ..network.unet import UNet
from ..optim import SGD, FlatParams, PolyLR
from .baseTrainer import BaseTrainer


class MeanTeacherTrainer(BaseTrainer):
    def __init__(self, phase, args=None):
        self.lambda_semi = 1
        self.ema_decay = 0.99
        self.epoch_rampup = 30
        self.alpha = 0
        self.semi_from_iter = 100
        self.log_step = 50
        self.parallel = None
        super(MeanTeacherTrainer, self).__init__(phase, args)

    def build_network(self):
        self.net = UNet(cfg.img_channels, cfg.n_label + 1, cfg.base_width, norm_type='instance', act_type='lrelu')
        self.net.to(self.device)
        if self.phase == 'train':
            self.ema = UNet(cfg.img_channels, cfg.n_label + 1, cfg.base_width, norm_type='instance', act_type='lrelu')
            for param in self.ema.parameters():
                param.detach_()
            self.ema.to(self.device)
            self.optimizer = SGD(self.net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
            self.ema_flat = FlatParams(self.ema.parameters())     # same tensor order as the student's flat buffer
            self.alpha_dev = torch.zeros(1, device=self.device)
            self.lr_sched = PolyLR([self.optimizer], cfg.lr, cfg.max_epoch * cfg.num_iter_per_epoch)

    def update_ema_variable(self):
        # alpha = min(1 - 1/(iter+1), 0.99), 0 while iter < 100; ONE fused launch over the flat buffers
        self.alpha = self.host_alpha()
        self.alpha_dev.fill_(float(self.alpha))
        ops.ema_update(self.ema_flat.flat, self.optimizer.flat, self.alpha_dev)

    def host_alpha(self):
        # alpha = min(1 - 1/(iter+1), 0.99), 0 while iter < 100 (meanTeacherTrainer.py:63-69)
        return 0 if self.iter < self.semi_from_iter else min(1 - 1 / (self.iter + 1), self.ema_decay)

    def train_step(self, img, msk, noise, lambda_semi, alpha=None, use_semi=None):
        """One iteration of meanTeacherTrainer.py:95-153: img = cat(labelled, unlabelled) (2*bs,1,H,W); `noise` is the
        clamped N(0, 0.01^2) perturbation of the teacher's input (drawn outside, L106).  alpha (device scalar) /
        use_semi: the iteration-dependent EMA coefficient and consistency switch handed in from outside (the captured
        graph of train_epoch); by default both follow self.iter."""
        bs = msk.shape[0]
        if use_semi is None:
            use_semi = self.iter >= self.semi_from_iter
        if isinstance(lambda_semi, torch.Tensor):
            lambda_semi = lambda_semi.reshape(())
        ops.arena_begin(img.device)
        self.lr_sched.tick()
        with ops.parallel_branch(5) as b_ema:        # the teacher's forward runs beside the student's
            with torch.no_grad():
                ema_outputs = self.ema(img[bs:] + noise)
        out = self.net(img)
        b_ema.join(ema_outputs)
        sample_loss = self.loss(out[:bs], msk)
        if not use_semi:
            semi_loss = torch.zeros((), dtype=torch.float32, device=img.device)
        else:
            semi_loss = Fn.SoftmaxMSEFn.apply(_flat_logits(out[bs:]), _flat_logits(ema_outputs))
        total_loss = sample_loss + lambda_semi * semi_loss
        self.optimizer.zero_grad()
        with Fn.accumulate_param_grads():
            total_loss.backward()
        if self.parallel is not None:
            self.parallel.all_reduce_grads(self.optimizer)
        self.optimizer.step()
        if alpha is None:
            self.update_ema_variable()
        else:
            ops.ema_update(self.ema_flat.flat, self.optimizer.flat, alpha)
        ops.arena_end()
        self.iter += 1
        return torch.stack([sample_loss.detach(), semi_loss.detach()])

    def train_epoch(self, lb_loader, ul_loader, meter, num_iter=None):
        self.net.train()
        lb_itr = iter(lb_loader)
        ul_itr = iter(ul_loader)
        lambda_semi = self.lambda_semi * self.sigmoid_rampup(self.epoch, self.epoch_rampup)
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
            noise = torch.clamp(torch.randn_like(img[cfg.batch_size:]) * 0.01, -0.02, 0.02)
            step = None
            if self.graph_enabled():
                use_semi = self.iter >= self.semi_from_iter
                self.alpha = self.host_alpha()
                self.alpha_dev.fill_(float(self.alpha))
                lam_dev.fill_(float(lambda_semi))
                inputs = [img, msk, noise, lam_dev, self.alpha_dev]
                step = self.graphed(('mean_teacher', bool(use_semi)),
                                    lambda *a: self.train_step(*a, use_semi=use_semi), inputs)
            if step is not None:
                losses = step(*inputs)
                self.iter += 1
            else:
                losses = self.train_step(img, msk, noise, lambda_semi)
            # meanTeacherTrainer.py:103,121: the supervised loss under the labelled batch's modality, weighted with the
            # size of the concatenated batch
            self.meter_note(meter, losses[0], mdl1[0].item(), img.size(0))
            if (i + 1) % self.log_step == 0:
                seg, semi = losses.tolist()
                self.info('Iter %d, global_iter: %d, semi_loss: %.4f, seg_loss: %.4f, lambda_semi: %f self.alpha: %f' %
                          (i, self.iter, semi, seg, lambda_semi, self.alpha))
            for param_group in self.optimizer.param_groups:
                param_group['lr'] = self.optimizer._lr_host = self.lr_sched.host_lr(self.iter)
        self.meter_flush()
        return losses


meanTeacherTrainer = MeanTeacherTrainer        # the reference's class name (meanTeacherTrainer.py:35)


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
        trainer = MeanTeacherTrainer('train', args)
        trainer.fit('inTurn', max_epoch=args.epochs, iters_per_epoch=args.iters)
    elif args.phase == 'test':
        trainer = MeanTeacherTrainer('test', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.test('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    elif args.phase == 'pseudo':
        trainer = MeanTeacherTrainer('pseudo', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.saving_pseudo('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    else:
        raise NotImplementedError
