"""Generates tests/golden/*.npz by running the REAL reference modules (imported from /root/reference, CPU, fp32)
on seeded inputs with the deterministic weights of oracle.smsut_oracle.make_weights.

Run in the build container only (`python tests/golden/make_golden.py`); /root/reference does not exist on the GPU
box, so the outputs are committed and tests/test_oracle.py pins the oracle against them.  The reference's trainers
cannot be imported (medpy / skimage / elasticdeform are not installed), so the one-iteration
fixture drives the reference's nn.Modules and losses with torch.optim.SGD / Adam through a literal transcription
of trainer/uganConsisTrainer.py:129-180 with the random draws injected.
"""
import os
import random
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
nn.Module.cuda = lambda self, *a, **k: self          # PatchSampleF.forward calls mlp.cuda() unconditionally (ugan.py:329)

import config as cfg  # noqa: E402  (the reference's)
from misc.loss import DiceAndCrossEntropyLoss  # noqa: E402
from network.patchnce import PatchNCELoss  # noqa: E402
from network.ugan import Discriminator, UGANnce  # noqa: E402
from network.unet import UNet  # noqa: E402

from oracle import smsut_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def npy(t):
    return t.detach().cpu().numpy().astype(np.float32)


def grad_norms(net):
    return {("gn." + k): np.float32(p.grad.norm().item()) for k, p in net.named_parameters()}


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, sum(a.nbytes for a in arrs.values()) // 1024, "KiB raw")


# ---- U-Net (network/unet.py) fwd + Dice/CE + bwd
net = UNet(1, 5, 16, norm_type='instance', act_type='lrelu')
net.load_state_dict(O.make_weights(O.unet_shapes(), 1))
x, y = O.synthetic_batch(2, 64, 3)
out = net(x)
crit = DiceAndCrossEntropyLoss(weight_ce=0.5, weight_dc=0.5, batch_dice=True)
loss = crit(out, y)
loss.backward()
save("unet", logits=npy(out), loss=npy(loss), fc_grad=npy(net.decoder.fc.weight.grad),
     pre_grad=npy(net.encoder.pre_conv.weight.grad), **grad_norms(net))

# ---- UGANnce (network/ugan.py) forward with injected patch ids + a backward
net = UGANnce(1, 5, 4, 16)
net.load_state_dict(O.make_weights(O.ugan_shapes(), 4))
x, _ = O.synthetic_batch(2, 64, 4)
m = torch.tensor([[1., 0, -1, 0], [0, 1., -1, 0]])
ids = [torch.randperm(16, generator=torch.Generator().manual_seed(0))]
seg, tsl, feats, _ = net(x, m, sample_ids=ids)
(seg.mean() + tsl.mean() + (feats[0] ** 3).sum()).backward()
save("ugannce", seg=npy(seg), tsl=npy(tsl), feat=npy(feats[0]), ids=ids[0].numpy(), **grad_norms(net))

# ---- Discriminator + gradient penalty (network/ugan.py:198-229, trainer/uganShp0Trainer.py:127-134)
D = Discriminator(64, 4, 16, max_width=256)
D.load_state_dict(O.make_weights(O.disc_shapes(64), 5))
x, _ = O.synthetic_batch(3, 64, 6)
x_hat = (x + 0.1 * torch.randn(x.shape, generator=torch.Generator().manual_seed(1))).requires_grad_(True)
out_src, out_cls = D(x_hat)
dydx = torch.autograd.grad(outputs=out_src, inputs=x_hat, grad_outputs=torch.ones(out_src.size()), retain_graph=True,
                           create_graph=True, only_inputs=True)[0]
gp = torch.mean((torch.sqrt(torch.sum(dydx.view(dydx.size(0), -1) ** 2, dim=1)) - 1) ** 2)
(10 * gp + out_src.mean() + out_cls.pow(2).mean()).backward()
save("discriminator", x_hat=npy(x_hat), out_src=npy(out_src), out_cls=npy(out_cls), gp=npy(gp), **grad_norms(D))

# ---- losses on random tensors (misc/loss.py, network/patchnce.py)
g = torch.Generator().manual_seed(2)
logits = (torch.randn(2, 5, 32, 32, generator=g) * 2).requires_grad_(True)
labels = torch.randint(0, 5, (2, 32, 32), generator=g)
l = crit(logits, labels)
l.backward()
q = torch.randn(64, 256, generator=g).requires_grad_(True)
k = torch.randn(64, 256, generator=g)
nce = PatchNCELoss(8)(q, k)
nce.mean().backward()
save("losses", logits=npy(logits), labels=labels.numpy(), dice_ce=npy(l), dlogits=npy(logits.grad), q=npy(q), k=npy(k),
     nce_rows=npy(nce), dq=npy(q.grad))

# ---- one UGANConsisTrainer iteration (trainer/uganConsisTrainer.py:129-180) with injected draws
size, bs = 64, 2
net = UGANnce(1, 5, 4, 16)
net.load_state_dict(O.make_weights(O.ugan_shapes(), 7))
D = Discriminator(size, 4, 16, max_width=256)
D.load_state_dict(O.make_weights(O.disc_shapes(size), 8))
optimizer = torch.optim.SGD(net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
d_optimizer = torch.optim.Adam(D.parameters(), cfg.lr, [0.9, 0.999], weight_decay=cfg.weight_decay)
criterionNCE = PatchNCELoss(cfg.batch_size)
x1, y_real = O.synthetic_batch(bs, size, 11)
x2, _ = O.synthetic_batch(bs, size, 12)
x_real = torch.cat([x1, x2])
modal_org = torch.cat([torch.full((bs,), 1), torch.full((bs,), 3)])
gen = torch.Generator().manual_seed(3)
fix = {}
for use_semi in (0, 1):
    mj = 2
    alpha = torch.randn(2 * bs, generator=gen).view(-1, 1, 1, 1)
    ids = [torch.randperm(16, generator=gen)]
    modal_trg = torch.zeros_like(modal_org).fill_(mj)
    vec_org, vec_trg = O.label2onehot(modal_org, 4), O.label2onehot(modal_trg, 4)
    vec_ot, vec_to = vec_trg - vec_org, vec_org - vec_trg
    out_src, out_cls = D(x_real)
    d_loss_real = - torch.mean(out_src)
    d_loss_cls = F.cross_entropy(out_cls, modal_org)
    _, x_fake, feat_x_pool, sample_ids = net(x_real, vec_ot, sample_ids=ids)
    out_src, out_cls = D(x_fake.detach())
    d_loss_fake = torch.mean(out_src)
    x_hat = (alpha * x_real.data + (1 - alpha) * x_fake.data).requires_grad_(True)
    out_src, _ = D(x_hat)
    dydx = torch.autograd.grad(outputs=out_src, inputs=x_hat, grad_outputs=torch.ones(out_src.size()),
                               retain_graph=True, create_graph=True, only_inputs=True)[0]
    d_loss_gp = torch.mean((torch.sqrt(torch.sum(dydx.view(dydx.size(0), -1) ** 2, dim=1)) - 1) ** 2)
    d_loss = d_loss_real + d_loss_fake + 1 * d_loss_cls + 10 * d_loss_gp
    d_optimizer.zero_grad(); optimizer.zero_grad()
    d_loss.backward()
    d_gn = {f"s{use_semi}.dgn." + k: np.float32(p.grad.norm().item()) for k, p in D.named_parameters()}
    d_optimizer.step()
    y_fake, x_fake, feat_x_pool, sample_ids = net(x_real, vec_ot, sample_ids=ids)
    out_src, out_cls = D(x_fake)
    g_loss_fake = - torch.mean(out_src)
    g_loss_cls = F.cross_entropy(out_cls, modal_trg)
    g_loss_seg = crit(y_fake[:bs], y_real)
    y_rec, x_rec, feat_f_pool, _ = net(x_fake, vec_to, sample_ids=sample_ids)
    g_loss_rec = torch.mean(torch.abs(x_real - x_rec))
    g_loss_semi = crit(y_rec, torch.argmax(y_fake, dim=1)) if use_semi else torch.tensor(0.)
    g_loss_nce = (criterionNCE(feat_f_pool[0], feat_x_pool[0]) * 1.0).mean()
    g_loss = g_loss_fake + 10 * g_loss_rec + 1 * g_loss_cls + 10 * g_loss_seg + 0.7 * g_loss_semi + 1.0 * g_loss_nce
    d_optimizer.zero_grad(); optimizer.zero_grad()
    g_loss.backward()
    g_gn = {f"s{use_semi}.ggn." + k: np.float32(p.grad.norm().item()) for k, p in net.named_parameters()}
    optimizer.step()
    fix.update(d_gn); fix.update(g_gn)
    fix[f"s{use_semi}.alpha"] = npy(alpha); fix[f"s{use_semi}.ids"] = ids[0].numpy()
    fix[f"s{use_semi}.losses"] = np.array([v.item() for v in (d_loss_real, d_loss_fake, d_loss_cls, d_loss_gp, g_loss_fake,
                                                              g_loss_rec, g_loss_cls, g_loss_seg, g_loss_semi, g_loss_nce)],
                                          dtype=np.float64)
    fix[f"s{use_semi}.G_checksum"] = np.array([p.detach().double().sum().item() for p in net.parameters()])
    fix[f"s{use_semi}.D_checksum"] = np.array([p.detach().double().sum().item() for p in D.parameters()])
    fix[f"s{use_semi}.x_fake"] = npy(x_fake)
save("consis_step", **fix)
