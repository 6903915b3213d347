# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/uganShp0Trainer.py: hyper-parameters (:37-50), build_network (:52-74),
G/D checkpoints (:76-107), label2onehot / create_vectors (:109-120), denorm (:122-125), gradient_penalty
(:127-134) and the device-side validate path (:250-287)."""
import os
from os.path import join as pjoin

import numpy as np
import torch
import torch.nn as nn

from .. import config as cfg
from .. import functional as Fn
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
            self.optimizer = SGD(self.net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
            self.d_optimizer = Adam(self.D.parameters(), cfg.lr, [beta1, beta2], weight_decay=cfg.weight_decay)
            self.lr_sched = PolyLR([self.optimizer, self.d_optimizer], cfg.lr, cfg.max_epoch * cfg.num_iter_per_epoch)

    def load_model(self, model_idx, which_ckpt):
        G_path = pjoin(self.expr_root, model_idx, 'ckpt', f'{which_ckpt}_G.ckpt')
        D_path = pjoin(self.expr_root, model_idx, 'ckpt', f'{which_ckpt}_D.ckpt')
        self.net.load_state_dict(torch.load(G_path, map_location='cpu'))
        self.D.load_state_dict(torch.load(D_path, map_location='cpu'))
        print(f'[*] Load G and D from {G_path}.')

    def save_model(self, prefix):
        assert self.phase == 'train'
        G_path = pjoin(self.expr_root, self.model_idx, 'ckpt', f'{prefix}_G.ckpt')
        D_path = pjoin(self.expr_root, self.model_idx, 'ckpt', f'{prefix}_D.ckpt')
        os.makedirs(os.path.dirname(G_path), exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in self.net.state_dict().items()}, G_path)
        torch.save({k: v.detach().cpu() for k, v in self.D.state_dict().items()}, D_path)
        print(f'[*] Save G and D to {G_path}.')

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
