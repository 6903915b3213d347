# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/coraNetTrainer.py: a U-Net with a
(1 + 3 * n_label)-channel output read as three heads that share the background channel -- a plain head and a
conservative / a radical one trained with class-weighted cross entropies (build_network :151-165) -- pre-trained on
the labelled slices (pre_epoch :426-524, prefit :526-602), then trained with pseudo labels of the unlabelled slices
where the conservative and radical heads agree ("certain" areas: Dice + masked CE, pred_unlabel :177-226) and with a
mean-teacher consistency where they do not ("uncertain" areas: masked softmax-MSE against the EMA teacher,
train_epoch :228-424, fit :604-690).  Same names, attributes and checkpoint files as the reference.

The arithmetic runs on the library's kernels: the U-Net of every other trainer, `smsut_heads_split_*`, the batch
Dice + CE of misc/loss.py, `smsut_wce_*` (class weights, certainty mask), the per-sample Dice, and
`smsut_softmax_mse_masked_*` (csrc/coranet.cu).  The pseudo-label set stays on the GPU (the reference round-trips
every slice through numpy and a DataLoader).

As shipped, the reference's config pairs n_label = 4 with the 2-class weight vectors of its SAML configuration
(config.py:82-88), with which nn.CrossEntropyLoss raises on the 5-class heads; config.py of this package carries the
CHAOS vectors the reference lists in its comments.  `cfg.thres` / get_mask (:131-135) are dead code there (every use
is commented out) and are not reproduced."""
import argparse
import os
import random
import sys
import time
from os.path import join as pjoin

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import __graft_entry__ as _g
    _g.load_package()
    __package__ = "smsut_b200.trainer"

import numpy as np
import torch
import torch.nn as nn

from .. import config as cfg
from .. import functional as Fn
from .. import ops
from ..misc.loss import _flat_logits
from ..network.unet import UNet
from ..optim import SGD, FlatParams, PolyLR
from .baseTrainer import BaseTrainer


class DiceAndCrossEntropyLoss(nn.Module):
    """The trainer's own loss class (coraNetTrainer.py:44-58): Dice (batch or per sample) + cross entropy with class
    weights; `reduc=True` is CrossEntropyLoss(reduction='none'), which the trainer only ever multiplies with a mask
    and normalises by `mask.sum() + 1e-16` (:301-303) -- here `forward(x, y, mask)` returns that quotient."""

    def __init__(self, weight_ce=1., weight_dc=1., batch_dice=False, weight=None, reduc=False):
        super(DiceAndCrossEntropyLoss, self).__init__()
        self.weight_ce, self.weight_dc, self.batch_dice, self.reduc = weight_ce, weight_dc, batch_dice, reduc
        w = cfg.default_w if weight is None else weight
        self.register_buffer('weight', torch.as_tensor(w, dtype=torch.float32))
        self.unit_weight = all(float(v) == 1.0 for v in w)      # decided on the host: forward() must stay capturable

    def forward(self, x, y, mask=None):
        """x: (B, C, H, W) logits or the (B*H*W, C) rows of a head; y: (B, H, W) labels; mask: (B, H, W) floats"""
        b = y.shape[0]
        logits = x if x.dim() == 2 else _flat_logits(x)
        yy = y.reshape(-1)
        yy = yy if yy.dtype == torch.int64 else yy.long()
        if self.batch_dice and mask is None and self.unit_weight:
            return Fn.DiceCEFn.apply(logits, yy, None, self.weight_ce, self.weight_dc)      # one fused pass (misc/loss.py)
        loss = 0.
        if self.weight_dc != 0:
            if self.batch_dice:
                dc = Fn.DiceCEFn.apply(logits, yy, None, 0., 1.)
            else:
                # per-sample Dice (SoftDiceLoss(batch_dice=False), misc/loss.py:52-63): the batch kernel on one sample
                # at a time; the mean over (sample, class) is the mean of the per-sample means
                per = logits.shape[0] // b
                dc = sum(Fn.DiceCEFn.apply(logits[i * per:(i + 1) * per], yy[i * per:(i + 1) * per], None, 0., 1.)
                         for i in range(b)) / b
            loss = loss + self.weight_dc * dc
        if self.weight_ce != 0:
            if self.reduc and mask is None:
                raise NotImplementedError("reduction='none' is only used under a certainty mask (coraNetTrainer.py:301)")
            m = None if mask is None else mask.reshape(-1).float().contiguous()
            loss = loss + self.weight_ce * Fn.WeightedCEFn.apply(logits, yy, self.weight.to(logits.device), m)
        return loss


class PseudoLabelSet(object):
    """The `make_data` dataset + DataLoader(batch_size, shuffle=True, drop_last=True) of coraNetTrainer.py:82-97,222-225
    as device tensors: iterating yields (img (B,H,W), plab, mask, lab, mdl) batches in a fresh random order."""

    def __init__(self, imgs, plabs, masks, labs, mdls, batch_size):
        self.img, self.plab, self.mask, self.lab, self.mdl = imgs, plabs, masks, labs, mdls
        self.batch_size = batch_size
        self.num = imgs.shape[0]

    def __len__(self):
        return self.num // self.batch_size

    def __iter__(self):
        perm = torch.as_tensor(np.random.permutation(self.num), device=self.img.device)
        for i in range(len(self)):
            idx = perm[i * self.batch_size:(i + 1) * self.batch_size]
            yield self.img[idx], self.plab[idx], self.mask[idx], self.lab[idx], self.mdl[idx]


class coraNetTrainer(BaseTrainer):
    def __init__(self, phase, args=None):
        self.lambda_semi = 1
        self.ema_decay = 0.99
        self.epoch_rampup = 30
        self.alpha = 0
        self.unsup_from_iter = 1000          # "if self.iter < 1000" (:340)
        self.model_id = getattr(args, 'model_id', None)
        self.log_step = 50
        self.n_heads = 3
        super(coraNetTrainer, self).__init__(phase, args)
        dev = self.device
        self.loss = DiceAndCrossEntropyLoss(weight_ce=cfg.weight_ce, weight_dc=cfg.weight_dc, batch_dice=True).to(dev)
        self.conloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., weight=cfg.w_con).to(dev)
        self.radloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., weight=cfg.w_rad).to(dev)
        self.CAceloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., reduc=True).to(dev)
        self.diceloss = DiceAndCrossEntropyLoss(weight_ce=0., weight_dc=1.).to(dev)
        self.CAconloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., weight=cfg.w_con, reduc=True).to(dev)
        self.CAradloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., weight=cfg.w_rad, reduc=True).to(dev)

    # ---- checkpoints of the teacher (:119-129) --------------------------------------------------------------------
    def load_ema_model(self, model_idx=None, which_ckpt='last'):
        if model_idx is None:
            model_idx = self.model_idx
        path = pjoin(self.expr_root, model_idx, 'ckpt', f'{which_ckpt}.ckpt')
        self.ema.load_state_dict(torch.load(path, map_location=lambda storage, loc: storage))
        ops.param_generation[0] += 1
        self.info(f'Load model from {path}.')

    def save_ema_model(self, prefix):
        path = pjoin(self.expr_root, self.model_idx, 'ckpt', f'{prefix}.ckpt')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({k: v.detach().float().cpu() for k, v in self.ema.state_dict().items()}, path)
        self.info(f'Save model to {path}.')

    def softmax_mse_loss(self, input_logits, target_logits):
        """(:137-149) elementwise (softmax - softmax)^2; the trainer's hot path uses the fused masked form"""
        assert input_logits.size() == target_logits.size()
        return (torch.softmax(input_logits, dim=1) - torch.softmax(target_logits, dim=1)) ** 2

    def build_network(self):
        n_out = cfg.n_label * 3 + 1
        self.net = UNet(cfg.img_channels, n_out, cfg.base_width, norm_type='instance', act_type='lrelu')
        self.net.to(self.device)
        if self.phase == 'train':
            self.ema = UNet(cfg.img_channels, n_out, cfg.base_width, norm_type='instance', act_type='lrelu')
            for param in self.ema.parameters():
                param.detach_()
            self.ema.to(self.device)
            self.optimizer = SGD(self.net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
            self.ema_flat = FlatParams(self.ema.parameters())
            self.alpha_dev = torch.zeros(1, device=self.device)
            # lr_ = cfg.lr * (1 - iter / (cora_epoch * num_iter_per_epoch)) ** 0.9 (:414)
            self.lr_sched = PolyLR([self.optimizer], cfg.lr, cfg.cora_epoch * cfg.num_iter_per_epoch)

    def host_alpha(self):
        return 0 if self.iter < 100 else min(1 - 1 / (self.iter + 1), self.ema_decay)

    def update_ema_variable(self):
        """(:168-175) one fused launch over the flat parameter buffers"""
        self.alpha = self.host_alpha()
        self.alpha_dev.fill_(float(self.alpha))
        ops.ema_update(self.ema_flat.flat, self.optimizer.flat, self.alpha_dev)

    # ---- the three heads ---------------------------------------------------------------------------------------------
    def heads(self, out):
        """(B, 1 + 3L, H, W) -> three (B*H*W, 1 + L) row tensors (`torch.cat([out_back, out_h], dim=1)`, :279-286)"""
        return Fn.HeadsSplitFn.apply(_flat_logits(out), cfg.n_label, self.n_heads)

    def segment(self, img):
        """validate_epoch / test / pseudo read head 0 (:715-733): channels 0 .. n_label of the output"""
        out = self.net(img)
        self._val_out = [out]           # for validation_loss (a list: not part of the trainer's live tensors)
        return out[:, :cfg.n_label + 1]

    def validation_loss(self, out, msk, b):
        """(:715-731) validate_epoch reports the three-head supervised loss, not head 0's Dice + CE alone"""
        return self.supervised_loss(self._val_out.pop()[:b], msk)[0]

    def supervised_loss(self, out, msk):
        h = self.heads(out)
        cedc_loss = self.loss(h[0], msk)
        loss_con = self.conloss(h[1], msk)
        loss_rad = self.radloss(h[2], msk)
        return (cedc_loss + loss_con + loss_rad) / 4, (cedc_loss, loss_con, loss_rad)

    @torch.no_grad()
    def pred_unlabel(self, ul_loader):
        """(:177-226) pseudo labels of every unlabelled slice: argmax of head 0; certainty mask: heads 1 and 2 agree.
        Returns (PseudoLabelSet, mean binary Dice of the pseudo labels against the withheld labels -- medpy's dc on the
        label maps, i.e. foreground vs background)."""
        self.net.eval()
        imgs, plabs, masks, labs, mdls = [], [], [], [], []
        dice_sum, n = 0.0, 0
        for img, lab, mdl, _ in ul_loader:
            img, lab = img.to(self.device, non_blocking=True), lab.to(self.device, non_blocking=True)
            b, _, hh, ww = img.shape
            h = self.heads(self.net(img))
            p0, p1, p2 = (ops.argmax_c(h[k].contiguous()).view(b, hh, ww) for k in range(3))
            imgs.append(img[:, 0]); plabs.append(p0); masks.append((p1 == p2).float()); labs.append(lab)
            mdls.append(torch.as_tensor(mdl).to(self.device).reshape(-1))
            fg_p, fg_l = p0 > 0, lab > 0
            inter = (fg_p & fg_l).flatten(1).sum(1).double()
            size = (fg_p.flatten(1).sum(1) + fg_l.flatten(1).sum(1)).double()
            dice_sum += torch.where(size > 0, 2 * inter / size.clamp_min(1), torch.zeros_like(inter)).sum().item()
            n += b
        self.net.train()
        plab_dice = dice_sum / max(n, 1)
        self.info('Pseudo label dice : {}'.format(plab_dice))
        new_loader = PseudoLabelSet(torch.cat(imgs), torch.cat(plabs), torch.cat(masks), torch.cat(labs), torch.cat(mdls),
                                    cfg.batch_size)
        return new_loader, plab_dice

    # ---- iterations ----------------------------------------------------------------------------------------------------
    def pre_step(self, img1, msk, alpha=None):
        """one iteration of pre_epoch (:461-499).  The reference also pushes the unlabelled half through the network and
        drops its output; only the labelled slices reach the loss (InstanceNorm: per sample), so only they are run."""
        ops.arena_begin(img1.device)
        loss, parts = self.supervised_loss(self.net(img1), msk)
        self.optimizer.zero_grad()
        with Fn.accumulate_param_grads():
            loss.backward()
        self.optimizer.step()
        if alpha is None:
            self.update_ema_variable()
        else:
            ops.ema_update(self.ema_flat.flat, self.optimizer.flat, alpha)
        ops.arena_end()
        self.iter += 1
        return torch.stack([loss.detach()] + [p.detach() for p in parts])

    def train_step(self, img1, msk, img2, plab2, mask, consistency_weight, alpha=None, use_unsup=None):
        """one iteration of train_epoch (:264-352): img1 / msk labelled, img2 (B,1,H,W) / plab2 / mask (B,H,W) from the
        pseudo-label set.  Before iter 1000 the certain / uncertain terms are replaced by zeros (:340-342): their forward
        passes are then not run at all."""
        if use_unsup is None:
            use_unsup = self.iter >= self.unsup_from_iter
        if isinstance(consistency_weight, torch.Tensor):
            consistency_weight = consistency_weight.reshape(())
        ops.arena_begin(img1.device)
        self.lr_sched.tick()
        zero = torch.zeros((), dtype=torch.float32, device=img1.device)
        if use_unsup:
            with ops.parallel_branch(5) as b_ema:        # the teacher's forward runs beside the student's
                with torch.no_grad():
                    ema_heads = self.heads(self.ema(img2))
        supervised_loss, _ = self.supervised_loss(self.net(img1), msk)
        if use_unsup:
            h2 = self.heads(self.net(img2))
            b_ema.join(ema_heads)
            dice_loss2 = self.diceloss(h2[0], plab2)
            loss_ce2 = self.CAceloss(h2[0], plab2, mask)
            certain_loss = (loss_ce2 + dice_loss2) / 2
            m = mask.reshape(-1).float().contiguous()
            consts = [consistency_weight * Fn.SoftmaxMSEMaskedFn.apply(h2[k], ema_heads[k].contiguous(), m, True)
                      for k in range(3)]
            uncertain_loss = (consts[0] + consts[1] + consts[2]) / 3
        else:
            certain_loss, uncertain_loss = zero, zero
        loss = supervised_loss + certain_loss + uncertain_loss * 0.1
        self.optimizer.zero_grad()
        with Fn.accumulate_param_grads():
            loss.backward()
        self.optimizer.step()
        if alpha is None:
            self.update_ema_variable()
        else:
            ops.ema_update(self.ema_flat.flat, self.optimizer.flat, alpha)
        ops.arena_end()
        self.iter += 1
        return torch.stack([supervised_loss.detach(), certain_loss.detach(), uncertain_loss.detach()])

    @staticmethod
    def _next(itr, loader):
        try:
            return next(itr), itr
        except StopIteration:
            itr = iter(loader)
            return next(itr), itr

    def pre_epoch(self, lb_loader, ul_loader, meter, num_iter=None):
        self.net.train()
        lb_itr = iter(lb_loader)
        losses = None
        for i in range(num_iter or cfg.num_iter_per_epoch):
            (img1, msk, mdl1, _), lb_itr = self._next(lb_itr, lb_loader)
            img1, msk = img1.to(self.device, non_blocking=True), msk.to(self.device, non_blocking=True)
            step = None
            if self.graph_enabled():
                self.alpha = self.host_alpha()
                self.alpha_dev.fill_(float(self.alpha))
                inputs = [img1, msk, self.alpha_dev]
                step = self.graphed(('cora_pre',), lambda *a: self.pre_step(*a), inputs)
            if step is not None:
                losses = step(*inputs)
                self.iter += 1
            else:
                losses = self.pre_step(img1, msk)
            # (:457,494-495) the iteration's loss under the labelled batch's modality; the reference weights it with the
            # size of its concatenated labelled + unlabelled batch
            self.meter_note(meter, losses[0], mdl1[0].item(), 2 * img1.size(0))
            if (i + 1) % self.log_step == 0:
                tl, cd, lc, lr_ = losses.tolist()
                self.info(f'Iter %d, global_iter: %d, train_loss: %.4f cedc_loss: %.4f, loss_con: %.4f, loss_rad: %.4f' %
                          (i, self.iter, tl, cd, lc, lr_))
        self.meter_flush()
        return losses

    def train_epoch(self, lb_loader, ul_loader, new_loader, meter=None, num_iter=None):
        self.net.train()
        self.ema.train()
        lb_itr, pse_itr = iter(lb_loader), iter(new_loader)
        consistency_weight = self.lambda_semi * self.sigmoid_rampup(self.epoch, self.epoch_rampup)
        cw_dev = torch.zeros(1, device=self.device)
        losses = None
        for i in range(num_iter or cfg.num_iter_per_epoch):
            (img1, msk, mdl1, _), lb_itr = self._next(lb_itr, lb_loader)
            (img2, plab2, mask, _, mdl2), pse_itr = self._next(pse_itr, new_loader)
            img1, msk = img1.to(self.device, non_blocking=True), msk.to(self.device, non_blocking=True)
            img2 = img2.to(self.device).unsqueeze(dim=1).contiguous()
            plab2, mask = plab2.to(self.device).contiguous(), mask.to(self.device).contiguous()
            step = None
            if self.graph_enabled():
                use_unsup = self.iter >= self.unsup_from_iter
                self.alpha = self.host_alpha()
                self.alpha_dev.fill_(float(self.alpha))
                cw_dev.fill_(float(consistency_weight))
                inputs = [img1, msk, img2, plab2, mask, cw_dev, self.alpha_dev]
                step = self.graphed(('cora_train', bool(use_unsup)),
                                    lambda *a: self.train_step(*a, use_unsup=use_unsup), inputs)
            if step is not None:
                losses = step(*inputs)
                self.iter += 1
            else:
                losses = self.train_step(img1, msk, img2, plab2, mask, consistency_weight)
            if meter is not None:       # (:283,344-351) supervised + certain + 0.1 uncertain
                self.meter_note(meter, losses[0] + losses[1] + 0.1 * losses[2], mdl1[0].item(),
                                img1.size(0) + img2.size(0))
            if (i + 1) % self.log_step == 0:
                sup, cer, unc = losses.tolist()
                self.info(f'Iter %d, global_iter: %d, supervised_loss: %.4f, certain_loss: %.4f, uncertain_loss: %f' %
                          (i, self.iter, sup, cer, unc))
            for param_group in self.optimizer.param_groups:
                param_group['lr'] = self.optimizer._lr_host = self.lr_sched.host_lr(self.iter)
        self.meter_flush()
        return losses

    # ---- the two training phases -------------------------------------------------------------------------------------
    def prefit(self, loader_type='inTurn', pre_epoch=None, iters_per_epoch=None, loaders=None):
        """(:526-602) supervised pre-training; keeps `pre_best` / `pre_ema_best` on the validation Dice and writes
        `pre_last` / `pre_ema_last` at the end"""
        self.setup_data_parallel()      # raises under torchrun: this trainer's steps exchange nothing
        train_lb_loader, train_ul_loader, test_loader = loaders if loaders is not None else self.make_loaders(loader_type)
        self.info_loader_sizes(train_lb_loader, train_ul_loader, test_loader)
        train_meter, test_meter = self.make_meters()
        self.init_train_env()
        self.open_writer()
        best_epoch = -1
        n_epoch = pre_epoch if pre_epoch is not None else cfg.pre_epoch
        tic = time.time()
        for epoch in range(n_epoch):
            train_meter.reset_cur()
            self.pre_epoch(train_lb_loader, train_ul_loader, train_meter, num_iter=iters_per_epoch)
            self.epoch += 1
            tic = self.log_train_stage(train_meter, epoch, best_epoch, n_epoch, tic, tag='pre ')
            tic = self.test_stage(test_loader, test_meter, epoch, n_epoch, tic, tag='pre ')
            if test_meter.cur_values['dice'] >= test_meter.best_values['dice']:
                self.save_model(prefix='pre_best')
                self.save_ema_model(prefix='pre_ema_best')
                best_epoch = epoch
        self.save_model(prefix='pre_last')
        self.save_ema_model(prefix='pre_ema_last')

    def fit(self, loader_type='inTurn', max_epoch=None, iters_per_epoch=None, loaders=None):
        """(:604-690) loads `pre_best` / `pre_ema_best` of run `model_id`, predicts the pseudo labels (again every
        cfg.pred_step epochs) and trains; `best` on the validation Dice, `last` at the end"""
        self.setup_data_parallel()      # raises under torchrun: this trainer's steps exchange nothing
        train_lb_loader, train_ul_loader, test_loader = loaders if loaders is not None else self.make_loaders(loader_type)
        self.info_loader_sizes(train_lb_loader, train_ul_loader, test_loader)
        train_meter, test_meter = self.make_meters()
        best_epoch = -1
        self.load_model(self.model_id, 'pre_best')
        self.load_ema_model(self.model_id, 'pre_ema_best')
        self.model_idx = None       # this run gets its own directory (the reference allocates it at construction)
        self.init_train_env()
        self.open_writer()
        new_loader, plab_dice = self.pred_unlabel(train_ul_loader)
        n_epoch = max_epoch if max_epoch is not None else cfg.cora_epoch
        tic = time.time()
        for epoch in range(n_epoch):
            if epoch % cfg.pred_step == 0:
                new_loader, plab_dice = self.pred_unlabel(train_ul_loader)
            train_meter.reset_cur()
            self.train_epoch(train_lb_loader, train_ul_loader, new_loader, train_meter, num_iter=iters_per_epoch)
            self.epoch += 1
            tic = self.log_train_stage(train_meter, epoch, best_epoch, n_epoch, tic)
            tic = self.test_stage(test_loader, test_meter, epoch, n_epoch, tic)
            if test_meter.cur_values['dice'] >= test_meter.best_values['dice']:
                self.save_model(prefix='best')
                best_epoch = epoch
        self.save_model(prefix='last')


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('-p', '--phase', type=str, choices=('train', 'pretrain', 'test', 'pseudo'))
    parser.add_argument('-f', '--fold', type=int, default=0)
    parser.add_argument('-nm', '--expr_name', type=str)
    parser.add_argument('-i', '--model_id', type=str, help='run whose pre_best checkpoints `-p train` starts from; the run to test')
    parser.add_argument('-wh', '--which_ckpt', type=str, default='last')
    parser.add_argument('--epochs', type=int, default=None)
    parser.add_argument('--iters', type=int, default=None)
    parser.add_argument('--input_size', type=int, default=None)
    args = parser.parse_args()

    random.seed(cfg.seed)
    np.random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(cfg.seed)

    trainer = coraNetTrainer('train' if args.phase == 'pretrain' else args.phase, args)
    if args.phase == 'pretrain':
        # the reference keeps `trainer.prefit('inTurn')` as a line to un-comment before `fit` ("firstly pretrain, then
        # train", :768); here it is its own phase
        trainer.prefit('inTurn', pre_epoch=args.epochs, iters_per_epoch=args.iters)
    elif args.phase == 'train':
        trainer.fit('inTurn', max_epoch=args.epochs, iters_per_epoch=args.iters)
    elif args.phase == 'test':
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.test('inTurn', pjoin(trainer.expr_root, args.model_id or '000'))
    elif args.phase == 'pseudo':
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.saving_pseudo('inTurn', pjoin(trainer.expr_root, args.model_id or '000'))
    else:
        raise NotImplementedError
