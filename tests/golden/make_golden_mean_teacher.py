"""Generates tests/golden/mean_teacher.npz: three iterations of the reference's mean-teacher loop around the REAL
reference modules (UNet, DiceAndCrossEntropyLoss from /root/reference, CPU, fp32) with torch.optim.SGD -- a literal
transcription of trainer/meanTeacherTrainer.py:95-153 (+ update_ema_variable :63-69, LR rule :148-151), started at
self.iter = 99 so that the run crosses the `iter < 100` switch of both the consistency loss and the EMA decay.  The
noise draw is injected.  Run in the build container only; tests/test_oracle.py pins oracle.mean_teacher_step to it."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import config as cfg  # noqa: E402  (the reference's)
from misc.loss import DiceAndCrossEntropyLoss  # noqa: E402
from network.unet import UNet  # noqa: E402

from oracle import smsut_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)
size, bs, it0 = 64, 2, 99
net = UNet(cfg.img_channels, cfg.n_label + 1, cfg.base_width, norm_type='instance', act_type='lrelu')
ema = UNet(cfg.img_channels, cfg.n_label + 1, cfg.base_width, norm_type='instance', act_type='lrelu')
net.load_state_dict(O.make_weights(O.unet_shapes(), 71))
ema.load_state_dict(O.make_weights(O.unet_shapes(), 72))
for param in ema.parameters():
    param.detach_()
loss_fn = DiceAndCrossEntropyLoss(weight_ce=cfg.weight_ce, weight_dc=cfg.weight_dc, batch_dice=True)
optimizer = torch.optim.SGD(net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
max_iter = cfg.max_epoch * cfg.num_iter_per_epoch
for g in optimizer.param_groups:                                   # the LR the loop left behind at the end of iter 98
    g['lr'] = cfg.lr * (1.0 - (it0 - 1) / max_iter) ** 0.9
lambda_semi = 0.8
fix = {}
self_iter = it0
for k in range(3):
    img1, msk = O.synthetic_batch(bs, size, 81 + k)
    img2, _ = O.synthetic_batch(bs, size, 91 + k)
    img = torch.cat([img1, img2], dim=0)
    ul_img = img[bs:]
    noise = torch.clamp(torch.randn(ul_img.shape, generator=torch.Generator().manual_seed(k)) * 0.01, -0.02, 0.02)
    ema_inputs = ul_img + noise
    out = net(img)
    out_soft = torch.softmax(out, dim=1)
    with torch.no_grad():
        ema_outputs = ema(ema_inputs)
        ema_outputs_soft = torch.softmax(ema_outputs, dim=1)
    sample_loss = loss_fn(out[:bs], msk)
    if self_iter < 100:
        semi_loss = torch.tensor(0., dtype=torch.float32)
    else:
        semi_loss = torch.mean((out_soft[bs:] - ema_outputs_soft) ** 2)
    total_loss = sample_loss + lambda_semi * semi_loss
    optimizer.zero_grad()
    total_loss.backward()
    optimizer.step()
    alpha = 0 if self_iter < 100 else min(1 - 1 / (self_iter + 1), 0.99)
    for ema_param, param in zip(ema.parameters(), net.parameters()):
        ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)
    lr_ = cfg.lr * (1.0 - self_iter / max_iter) ** 0.9
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr_
    self_iter += 1
    fix[f"losses{k}"] = np.array([sample_loss.item(), semi_loss.item()])
    fix[f"net_sum{k}"] = np.array([p.detach().double().sum().item() for p in net.parameters()])
    fix[f"ema_sum{k}"] = np.array([p.detach().double().sum().item() for p in ema.parameters()])
    fix[f"net_norm{k}"] = np.array([p.detach().double().norm().item() for p in net.parameters()])
    fix[f"ema_norm{k}"] = np.array([p.detach().double().norm().item() for p in ema.parameters()])
fix["fc"] = net.decoder.fc.weight.detach().numpy().astype(np.float32)
fix["ema_fc"] = ema.decoder.fc.weight.detach().numpy().astype(np.float32)
np.savez_compressed(os.path.join(HERE, "mean_teacher.npz"), **fix)
print("mean_teacher", sum(a.nbytes for a in fix.values()) // 1024, "KiB raw")
