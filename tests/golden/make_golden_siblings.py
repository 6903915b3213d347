"""Generates tests/golden/siblings.npz: one iteration each of the sibling trainers of SURVEY.md section 8(f) N4, run on
the REAL reference modules (UNet, UGAN, Discriminator, DiceAndCrossEntropyLoss from /root/reference, CPU, fp32) with
torch.optim.SGD / Adam through a literal transcription of
  * trainer/crossPseTrainer.py:96-131 (two iterations),
  * trainer/uganTrainer.py:159-196 (shape loss; one iteration) and trainer/uganShp0Trainer.py:180-217 (no shape loss),
with the random draws injected (the trainer files cannot be imported offline: medpy / skimage are absent).
Run in the build container only; tests/test_oracle.py pins oracle.cross_pse_step / ugan_shape_step against it."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import config as cfg  # noqa: E402  (the reference's)
from misc.loss import DiceAndCrossEntropyLoss  # noqa: E402
from network.ugan import UGAN, Discriminator  # noqa: E402
from network.unet import UNet  # noqa: E402

from oracle import smsut_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)
crit = DiceAndCrossEntropyLoss(weight_ce=0.5, weight_dc=0.5, batch_dice=True)
fix = {}


def gn(prefix, net):
    """gradient norms as ONE array + the parameter names they belong to (a zip entry per scalar costs 200 bytes)"""
    names = [k for k, _ in net.named_parameters()]
    return {prefix + "norms": np.array([p.grad.norm().item() for _, p in net.named_parameters()], dtype=np.float32),
            prefix + "names": np.array(",".join(names))}


def checksum(net):
    return np.array([p.detach().double().sum().item() for p in net.parameters()])


# ---- cross pseudo supervision
size, bs = 64, 2
net = UNet(1, 5, 16, norm_type='instance', act_type='lrelu'); net.load_state_dict(O.make_weights(O.unet_shapes(), 31))
net2 = UNet(1, 5, 16, norm_type='instance', act_type='lrelu'); net2.load_state_dict(O.make_weights(O.unet_shapes(), 32))
optimizer1 = torch.optim.SGD(net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
optimizer2 = torch.optim.SGD(net2.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
lambda_semi = 0.1 * 0.5
for it in range(2):
    img1, msk = O.synthetic_batch(bs, size, 41 + it)
    img2, _ = O.synthetic_batch(bs, size, 51 + it)
    img = torch.cat([img1, img2], dim=0)
    out1 = net(img)
    sample1_loss = crit(out1[:bs], msk)
    out2 = net2(img)
    sample2_loss = crit(out2[:bs], msk)
    pred1 = torch.argmax(out1[bs:], dim=1).detach()
    pred2 = torch.argmax(out2[bs:], dim=1).detach()
    semi1_loss = crit(out1[bs:], pred2)
    semi2_loss = crit(out2[bs:], pred1)
    total_loss = sample1_loss + sample2_loss + lambda_semi * semi1_loss + lambda_semi * semi2_loss
    optimizer1.zero_grad(); optimizer2.zero_grad()
    total_loss.backward()
    fix.update(gn(f"cps{it}.g1.", net)); fix.update(gn(f"cps{it}.g2.", net2))
    optimizer1.step(); optimizer2.step()
    fix[f"cps{it}.losses"] = np.array([v.item() for v in (sample1_loss, sample2_loss, semi1_loss, semi2_loss)])
    fix[f"cps{it}.sum1"], fix[f"cps{it}.sum2"] = checksum(net), checksum(net2)

# ---- UGAN with / without the shape loss
ugan_shapes = {k: v for k, v in O.ugan_shapes().items() if not k.startswith("netF.")}
for name, lambda_shp in (("shp", 3.5), ("shp0", None)):
    G = UGAN(1, 5, 4, 16); G.load_state_dict(O.make_weights(ugan_shapes, 61))
    D = Discriminator(size, 4, 16, max_width=256); D.load_state_dict(O.make_weights(O.disc_shapes(size), 62))
    optimizer = torch.optim.SGD(G.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
    d_optimizer = torch.optim.Adam(D.parameters(), cfg.lr, [0.9, 0.999], weight_decay=cfg.weight_decay)
    x_real, y_real = O.synthetic_batch(3, size, 63)
    modal_org = torch.full((3,), 1)
    mj = 3
    modal_trg = torch.zeros_like(modal_org).fill_(mj)
    vec_org, vec_trg = O.label2onehot(modal_org, 4), O.label2onehot(modal_trg, 4)
    vec_ot, vec_to = vec_trg - vec_org, vec_org - vec_trg
    alpha = torch.randn(3, generator=torch.Generator().manual_seed(64)).view(-1, 1, 1, 1)
    out_src, out_cls = D(x_real)
    d_loss_real = - torch.mean(out_src)
    d_loss_cls = F.cross_entropy(out_cls, modal_org)
    _, x_fake = G(x_real, vec_ot)
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
    fix.update(gn(f"{name}.dgn.", D))
    d_optimizer.step()
    y_fake, x_fake = G(x_real, vec_ot)
    out_src, out_cls = D(x_fake)
    g_loss_fake = - torch.mean(out_src)
    g_loss_cls = F.cross_entropy(out_cls, modal_trg)
    g_loss_seg = crit(y_fake, y_real)
    y_rec, x_rec = G(x_fake, vec_to)
    g_loss_rec = torch.mean(torch.abs(x_real - x_rec))
    g_loss = g_loss_fake + 10 * g_loss_rec + 1 * g_loss_cls + 10 * g_loss_seg
    vals = [d_loss_real, d_loss_fake, d_loss_cls, d_loss_gp, g_loss_fake, g_loss_rec, g_loss_cls, g_loss_seg]
    if lambda_shp is not None:
        g_loss_shp = crit(y_rec, y_real)
        g_loss = g_loss + lambda_shp * g_loss_shp
        vals.append(g_loss_shp)
    d_optimizer.zero_grad(); optimizer.zero_grad()
    g_loss.backward()
    fix.update(gn(f"{name}.ggn.", G))
    optimizer.step()
    fix[f"{name}.alpha"] = alpha.numpy().astype(np.float32)
    fix[f"{name}.losses"] = np.array([v.item() for v in vals])
    fix[f"{name}.G_checksum"], fix[f"{name}.D_checksum"] = checksum(G), checksum(D)

np.savez_compressed(os.path.join(HERE, "siblings.npz"), **fix)
print("siblings", sum(a.nbytes for a in fix.values()) // 1024, "KiB raw")
