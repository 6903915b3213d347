# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/uganShp0Trainer.py: hyper-parameters (:37-50), build_network (:52-74),
G/D checkpoints (:76-107), label2onehot / create_vectors (:109-120), denorm (:122-125), gradient_penalty
(:127-134), the labelled-only GAN iteration (:136-235; shared with trainer/uganTrainer.py, which adds the shape
loss) and the device-side validate path (:250-287)."""
import os
import random
import time
from os.path import join as pjoin

import numpy as np
import torch
import torch.nn as nn

from .. import config as cfg
from .. import functional as Fn
from .. import ops
from ..network.blocks import refresh_packs
from ..network.patchnce import PatchNCELoss
from ..network.ugan import Discriminator, UGANnce
from ..optim import SGD, Adam, PolyLR
from .baseTrainer import BaseTrainer


class UGANShp0Trainer(BaseTrainer):
    def __init__(self, phase, args):
        # Hyper params.
        self.lambda_cls = 1
        self.lambda_rec = 10
        self.lambda_gp = 10
        self.lambda_seg = 10

        self.log_step = 50
        self.n_critic = 1

        self.beta1 = 0.9
        self.beta2 = 0.999

        super(UGANShp0Trainer, self).__init__(phase, args)

    def build_network(self):
        self.net = UGANnce(cfg.img_channels, cfg.n_label + 1, cfg.n_modal, cfg.base_width)
        self.net.to(self.device)

        self.criterionNCE = []
        for nce_layer in cfg.nce_layers:
            self.criterionNCE.append(PatchNCELoss(cfg.batch_size).to(self.device))

        self.D = Discriminator(self.input_size, cfg.n_modal, cfg.base_width,
                               max_width=256 if cfg.base_width == 16 else 512)
        self.D.to(self.device)

        # data parallelism is one process per GPU (parallel.py), not nn.DataParallel: nothing to wrap here
        if self.phase == 'train':
            beta1, beta2 = self.beta1, self.beta2
            # early bucket: the segmentation halves and netF get their whole gradient in the generator's early backward
            # stage (UGANConsisTrainer.train_step) -- a data-parallel run reduces them beside the discriminator phase
            early = [p for n, p in self.net.named_parameters() if n.startswith(('seg_encoder.', 'seg_decoder.', 'netF.'))]
            self.optimizer = SGD(self.net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay, early=early)
            self.d_optimizer = Adam(self.D.parameters(), cfg.lr, [beta1, beta2], weight_decay=cfg.weight_decay)
            self.lr_sched = PolyLR([self.optimizer, self.d_optimizer], cfg.lr, cfg.max_epoch * cfg.num_iter_per_epoch)

    def load_model(self, model_idx, which_ckpt):
        G_path = pjoin(self.expr_root, model_idx, 'ckpt', f'{which_ckpt}_G.ckpt')
        D_path = pjoin(self.expr_root, model_idx, 'ckpt', f'{which_ckpt}_D.ckpt')
        self.net.load_state_dict(torch.load(G_path, map_location='cpu'))
        self.D.load_state_dict(torch.load(D_path, map_location='cpu'))
        self._model_idx = model_idx
        print(f'[*] Load G and D from {G_path}.')

    def save_model(self, prefix):
        assert self.phase == 'train'
        if not self.is_main:
            return
        G_path = pjoin(self.expr_root, self.model_idx, 'ckpt', f'{prefix}_G.ckpt')
        D_path = pjoin(self.expr_root, self.model_idx, 'ckpt', f'{prefix}_D.ckpt')
        os.makedirs(os.path.dirname(G_path), exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in self.net.state_dict().items()}, G_path)
        torch.save({k: v.detach().cpu() for k, v in self.D.state_dict().items()}, D_path)
        self.info(f'[*] Save G and D to {G_path}.')

    def label2onehot(self, modals, dim=cfg.n_modal):
        batch_size = modals.size(0)
        out = torch.zeros(batch_size, dim)
        out[np.arange(batch_size), modals.long()] = 1
        return out

    def create_vectors(self, vec_org, dim):
        vec_trg_list = []
        for i in range(dim):
            vec_trg = self.label2onehot(torch.ones(vec_org.size(0)) * i, dim)
            vec_trg_list.append(vec_trg.to(self.device))
        return vec_trg_list

    @staticmethod
    def denorm(x):
        out = (x + 1.) / 2.
        return out.clamp_(0, 1)

    def sample_translations(self, x_fixed, modal_org, save_path=None):
        """The per-epoch debug grid of uganShp0Trainer.py:219-228 / uganConsisTrainer.py:205-214: the fixed slices next
        to their translation into every modality, concatenated along the width and de-normalised to [0, 1]
        ((B, 1, H, (n_modal + 1) * W) fp32 on the host).  save_path: also written as one image (nrow=1, padding=0
        like save_image) when PIL is importable."""
        was_training = self.net.training
        self.net.eval()
        with torch.no_grad():
            x_fixed = x_fixed.to(self.device)
            vec_fixed_org = self.label2onehot(modal_org, cfg.n_modal).to(self.device)
            x_fake_list = [x_fixed.float()]
            for vec_fixed in self.create_vectors(vec_fixed_org, cfg.n_modal):
                _, x_fake = self.translate(x_fixed, vec_fixed - vec_fixed_org)
                x_fake_list.append(x_fake.float())
            grid = self.denorm(torch.cat(x_fake_list, dim=3).cpu())
        self.net.train(was_training)
        if save_path is not None:
            try:
                from PIL import Image
            except ImportError:
                Image = None
            if Image is not None:
                os.makedirs(os.path.dirname(save_path) or '.', exist_ok=True)
                rows = (grid[:, 0] * 255.0 + 0.5).clamp(0, 255).to(torch.uint8)       # save_image's quantisation
                Image.fromarray(torch.cat(list(rows), dim=0).numpy()).save(save_path)
                print(f'[*] Saved real and fake images into {save_path}.')
        return grid

    def gradient_penalty(self, y, x):
        """mean_b (|| d sum(y) / d x_b ||_2 - 1)^2 with a differentiable first-order pass: the backward of every
        op of D emits its hand-written second-order kernels (functional.py)."""
        weight = torch.ones(y.size(), device=y.device)
        with Fn.inputs_only():
            dydx = torch.autograd.grad(outputs=y, inputs=x, grad_outputs=weight,
                                       retain_graph=True, create_graph=True,
                                       only_inputs=True)[0]
        return Fn.GradPenaltyFn.apply(dydx.contiguous())

    def segment(self, img):
        seg, _ = self.net(img, val_phase=True)
        return seg

    def translate(self, x, m):
        """(seg, tsl) of the generator.  The reference's own loop unpacks two values from UGANnce.forward, which
        returns four outside val_phase (network/ugan.py:195), so uganShp0Trainer.py:180 can only ever have run with
        the two-output form: that is what is computed here."""
        return self.net(x, m, val_phase=True)

    SHP_LOSS_KEYS = ('D_real', 'D_fake', 'D_cls', 'D_gp', 'G_fake', 'G_rec', 'G_cls', 'G_seg', 'G_shp')

    def shape_train_step(self, x_real, y_real, modal_org, modal_trg, vec_ot, vec_to, alpha, lambda_shp=None):
        """One iteration (n_critic = 1) of uganShp0Trainer.py:162-217 on labelled device tensors, or, with
        `lambda_shp` (host float or device scalar), of uganTrainer.py:141-196 whose G loss adds
        lambda_shp * Dice+CE(y_rec, y_real).  As in UGANConsisTrainer.train_step, one generator forward serves the D
        phase (detached) and the G phase, the cycle pass runs on a branch stream beside the D phase, and D's weight
        gradients are not computed in the G phase.  Returns the losses ordered like SHP_LOSS_KEYS (G_shp = 0 without
        the shape term) as one device vector."""
        lambda_cls, lambda_gp = self.lambda_cls, self.lambda_gp
        lambda_seg, lambda_rec = self.lambda_seg, self.lambda_rec
        ops.arena_begin(x_real.device)
        self.lr_sched.tick()
        y_fake, x_fake = self.translate(x_real, vec_ot)
        if isinstance(lambda_shp, torch.Tensor):
            lambda_shp = lambda_shp.reshape(())
        with ops.parallel_branch(4) as b_cyc:
            g_loss_seg = self.loss(y_fake, y_real)
            y_rec, x_rec = self.translate(x_fake, vec_to)
            g_loss_rec = Fn.L1MeanFn.apply(x_rec.contiguous(), x_real)
            g_partial = lambda_rec * g_loss_rec + lambda_seg * g_loss_seg
            if lambda_shp is not None:
                g_loss_shp = self.loss(y_rec, y_real)
                g_partial = g_partial + lambda_shp * g_loss_shp
            else:
                g_loss_shp = torch.zeros((), device=x_real.device)

        x_fake_d = x_fake.detach()
        refresh_packs(self.D)
        with ops.parallel_branch(1) as b_fake:
            out_src_f, _ = self.D(x_fake_d)
            d_loss_fake = Fn.MeanFn.apply(out_src_f, 1.0)
        with ops.parallel_branch(2) as b_hat:
            x_hat = ops.lerp_rows(alpha, x_real, x_fake_d.contiguous()).requires_grad_(True)
            out_src_h, _ = self.D(x_hat)
            d_loss_gp = self.gradient_penalty(out_src_h, x_hat)
        out_src, out_cls = self.D(x_real)
        d_loss_real = Fn.MeanFn.apply(out_src, -1.0)
        d_loss_cls = Fn.CERowsFn.apply(out_cls.contiguous(), modal_org)
        b_fake.join(d_loss_fake)
        b_hat.join(d_loss_gp)
        d_loss = d_loss_real + d_loss_fake + lambda_cls * d_loss_cls + lambda_gp * d_loss_gp
        self.d_optimizer.zero_grad()
        with Fn.accumulate_param_grads():
            d_loss.backward()
        if getattr(self, 'parallel', None) is not None:
            self.parallel.all_reduce_grads(self.d_optimizer)
        self.d_optimizer.step()

        for p in self.d_optimizer.params:
            p.requires_grad_(False)
        out_src, out_cls = self.D(x_fake)
        for p in self.d_optimizer.params:
            p.requires_grad_(True)
        g_loss_fake = Fn.MeanFn.apply(out_src, -1.0)
        g_loss_cls = Fn.CERowsFn.apply(out_cls.contiguous(), modal_trg)
        b_cyc.join(g_partial, g_loss_seg, g_loss_rec, g_loss_shp)
        g_loss = g_loss_fake + lambda_cls * g_loss_cls + g_partial
        self.optimizer.zero_grad()
        with Fn.accumulate_param_grads():
            g_loss.backward()
        if getattr(self, 'parallel', None) is not None:
            self.parallel.all_reduce_grads(self.optimizer)
        self.optimizer.step()
        ops.arena_end()
        return torch.stack([d_loss_real.detach(), d_loss_fake.detach(), d_loss_cls.detach(), d_loss_gp.detach(),
                            g_loss_fake.detach(), g_loss_rec.detach(), g_loss_cls.detach(), g_loss_seg.detach(),
                            g_loss_shp.detach()])

    def epoch_lambda_shp(self):
        return None                     # uganShp0Trainer: no shape term

    def train_epoch(self, lb_loader, ul_loader, meter, num_iter=None):
        """uganShp0Trainer.py:136-235 / uganTrainer.py:115-215: labelled slices only, one target modality per
        iteration (random.randint), alpha ~ N(0,1) per slice."""
        self.net.train()
        self.D.train()
        lambda_shp = self.epoch_lambda_shp()
        self.info(f'\nlambda_seg: {self.lambda_seg}' + ('.' if lambda_shp is None else f', lambda_shp: {lambda_shp}.'))
        itr = iter(lb_loader)
        tic = time.time()
        losses = None
        # fixed images of the epoch's sample grid (uganShp0Trainer.py:149-155): the loader's first batch, taken off the
        # iterator before the loop as the reference does
        x_fixed, _, modal_fixed, inm = next(itr)
        if inm is not None:
            self.info(list(inm))
        lam_dev = torch.zeros(1, device=self.device)
        for i in range(self.n_critic * (num_iter or cfg.num_iter_per_epoch)):
            try:
                x_real, y_real, modal_org, _ = next(itr)
            except StopIteration:
                itr = iter(lb_loader)
                x_real, y_real, modal_org, _ = next(itr)
            mj = self.target_modality_rng.randint(0, cfg.n_modal - 1)
            modal_trg = torch.zeros_like(modal_org).fill_(mj)
            vec_org = self.label2onehot(modal_org, cfg.n_modal)
            vec_trg = self.label2onehot(modal_trg, cfg.n_modal)
            dev = self.device
            alpha = torch.randn(x_real.size(0), device=dev)
            batch = [x_real.to(dev, non_blocking=True), y_real.to(dev, non_blocking=True), modal_org.to(dev),
                     modal_trg.to(dev), (vec_trg - vec_org).to(dev), (vec_org - vec_trg).to(dev), alpha]
            step = None
            if self.graph_enabled():
                # lambda_shp (uganTrainer.py:122-123, per epoch) rides in a device scalar; None = no shape term
                if lambda_shp is not None:
                    lam_dev.fill_(float(lambda_shp))
                    inputs = batch + [lam_dev]
                    step = self.graphed('shape', self.shape_train_step, inputs)
                else:
                    inputs = batch
                    step = self.graphed('shp0', self.shape_train_step, inputs)
            if step is not None:
                losses = step(*inputs)
            else:
                losses = self.shape_train_step(*batch, lambda_shp)
            # uganShp0Trainer.py:162,206-207: the segmentation loss under the batch's modality
            self.meter_note(meter, losses[self.SHP_LOSS_KEYS.index('G_seg')], modal_org[0].item(), x_real.size(0))
            if (i + 1) % (self.n_critic * self.log_step) == 0:
                log = 'Iter: %d/%d(%d), elapsed: %.2fs,' % (i, self.n_critic * cfg.num_iter_per_epoch, self.iter,
                                                            time.time() - tic)
                tic = time.time()
                for k, v in zip(self.SHP_LOSS_KEYS, losses.tolist()):
                    if k != 'G_shp' or lambda_shp is not None:
                        log += ' %s: %.4f,' % (k, v)
                self.info(log)
            lr_ = self.lr_sched.host_lr(self.iter + 1)
            for opt in (self.optimizer, self.d_optimizer):
                for param_group in opt.param_groups:
                    param_group['lr'] = lr_
                opt._lr_host = lr_
            self.iter += 1
        if getattr(self, 'save_samples', False) and self.is_main:      # uganShp0Trainer.py:219-228 (off by default)
            self.sample_translations(x_fixed, modal_fixed, save_path=os.path.join(
                self.expr_root, self.model_idx, 'sample', f'train-{self.epoch + 1}-images.png'))
        self.meter_flush()
        return losses
