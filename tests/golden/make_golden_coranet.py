"""Generates tests/golden/coranet.npz: iterations of the reference's coraNetTrainer (trainer/coraNetTrainer.py) run on
the REAL reference pieces, CPU fp32:
  * network/unet.py UNet with 1 + 3 * n_label output channels (build_network, :151-165),
  * the trainer's OWN `DiceAndCrossEntropyLoss` (class weights / reduction='none', :44-58) and `softmax_mse_loss`
    (:137-149), lifted out of the trainer's SOURCE FILE with `ast` and executed unchanged -- the module itself cannot be
    imported offline (medpy / tensorboard) -- with `Tensor.cuda()` made the identity (the class calls `.cuda()` on its
    weight vector; there is no GPU in the build container),
  * a literal transcription of one pre_epoch iteration (:461-499), pred_unlabel (:186-207) and two train_epoch
    iterations (:264-352: one before and one after the iter-1000 switch) around them, torch.optim.SGD, the EMA rule
    of update_ema_variable (:168-175).
The class-weight vectors are the CHAOS ones of the reference config's comments (config.py:82-90): the shipped values
are the 2-class SAML vectors, with which nn.CrossEntropyLoss raises on the 5-class heads of n_label = 4.
Run in the build container only; tests/test_oracle.py pins the oracle's coranet_* functions against the output."""
import ast
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import config as cfg  # noqa: E402  (the reference's)
from misc.loss import SoftDiceLoss  # noqa: E402
from network.unet import UNet  # noqa: E402

from oracle import smsut_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)
torch.Tensor.cuda = lambda self, *a, **k: self          # see the docstring

cfg.default_w = torch.FloatTensor(O.CORA_W["default"])
cfg.w_con = torch.FloatTensor(O.CORA_W["con"])
cfg.w_rad = torch.FloatTensor(O.CORA_W["rad"])

SRC = "/root/reference/trainer/coraNetTrainer.py"
tree = ast.parse(open(SRC).read())
loss_cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DiceAndCrossEntropyLoss"]
trainer = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "coraNetTrainer"][0]
mse_fn = [n for n in trainer.body if isinstance(n, ast.FunctionDef) and n.name == "softmax_mse_loss"]
ns = {"nn": nn, "torch": torch, "F": F, "cfg": cfg, "SoftDiceLoss": SoftDiceLoss}
exec(compile(ast.Module(body=loss_cls + mse_fn, type_ignores=[]), SRC, "exec"), ns)
DiceAndCrossEntropyLoss, softmax_mse_loss = ns["DiceAndCrossEntropyLoss"], ns["softmax_mse_loss"]

# the loss objects of coraNetTrainer.__init__ (:110-117)
loss_ = DiceAndCrossEntropyLoss(weight_ce=cfg.weight_ce, weight_dc=cfg.weight_dc, batch_dice=True)
conloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., weight=cfg.w_con)
radloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., weight=cfg.w_rad)
CAceloss = DiceAndCrossEntropyLoss(weight_ce=1., weight_dc=0., reduc=True)
diceloss = DiceAndCrossEntropyLoss(weight_ce=0., weight_dc=1.)

size, bs = 64, 2
n_out = cfg.n_label * 3 + 1
net = UNet(cfg.img_channels, n_out, cfg.base_width, norm_type='instance', act_type='lrelu')
ema = UNet(cfg.img_channels, n_out, cfg.base_width, norm_type='instance', act_type='lrelu')
net.load_state_dict(O.make_weights(O.unet_shapes(out_ch=n_out), 71))
ema.load_state_dict(O.make_weights(O.unet_shapes(out_ch=n_out), 72))
for param in ema.parameters():
    param.detach_()
optimizer = torch.optim.SGD(net.parameters(), lr=cfg.lr, momentum=0.9, weight_decay=cfg.weight_decay)
fix = {"n_out": np.array(n_out)}


def heads(out, b):
    out_back = out[:b, 0, :, :].unsqueeze(dim=1)
    r = []
    for h in range(3):
        r.append(torch.cat([out_back, out[:b, (h * cfg.n_label + 1): (h + 1) * cfg.n_label + 1, :, :]], dim=1))
    return r


def update_ema_variable(it):
    alpha = 0 if it < 100 else min(1 - 1 / (it + 1), 0.99)
    for ema_param, param in zip(ema.parameters(), net.parameters()):
        ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)


def record(tag, losses):
    names = [k for k, _ in net.named_parameters()]
    fix[tag + ".losses"] = np.array([float(v.detach()) for v in losses])
    fix[tag + ".gnorms"] = np.array([p.grad.norm().item() if p.grad is not None else 0.0 for _, p in net.named_parameters()],
                                    dtype=np.float32)
    fix[tag + ".names"] = np.array(",".join(names))


def checksums(tag):
    fix[tag + ".sum_net"] = np.array([p.detach().double().sum().item() for p in net.parameters()])
    fix[tag + ".sum_ema"] = np.array([p.detach().double().sum().item() for p in ema.parameters()])


# ---- pre_epoch iteration (:461-499), at iter 200 (EMA coefficient 0.99: the teacher keeps its own weights)
img1, msk = O.synthetic_batch(bs, size, 81)
img2, _ = O.synthetic_batch(bs, size, 82)
out = net(torch.cat([img1, img2], dim=0))
out0, out1, out2 = heads(out, bs)
cedc_loss = loss_(out0, msk)
loss_con = conloss(out1, msk)
loss_rad = radloss(out2, msk)
loss = (cedc_loss + loss_con + loss_rad) / 4
optimizer.zero_grad()
loss.backward()
record("pre", (loss, cedc_loss, loss_con, loss_rad))
optimizer.step()
update_ema_variable(200)
checksums("pre")

# ---- pred_unlabel (:186-207) on a batch of unlabelled slices
imgu, _ = O.synthetic_batch(bs, size, 83)
with torch.no_grad():
    o0, o1, o2 = heads(net(imgu), bs)
    plab = torch.argmax(o0, dim=1)
    mask = (torch.argmax(o1, dim=1) == torch.argmax(o2, dim=1)).float()
fix["pred.plab"] = plab.numpy().astype(np.uint8)
fix["pred.mask"] = mask.numpy().astype(np.uint8)

# ---- train_epoch iterations (:264-352): iter 300 (certain / uncertain zeroed) and iter 1500
for tag, it, cw in (("trn_early", 300, 0.3), ("trn_late", 1500, 0.3)):
    img1, msk = O.synthetic_batch(bs, size, 91 + it)
    out_s = net(img1)
    out0, out1, out2 = heads(out_s, bs)
    supervised_loss = (loss_(out0, msk) + conloss(out1, msk) + radloss(out2, msk)) / 4
    out_p = net(imgu)
    out20, out21, out22 = heads(out_p, bs)
    dice_loss2 = diceloss(out20, plab)
    loss_ce2 = (CAceloss(out20, plab) * mask).sum() / (mask.sum() + 1e-16)
    certain_loss = (loss_ce2 + dice_loss2) / 2
    mask_u = (1 - mask).unsqueeze(1)
    with torch.no_grad():
        out_ema = ema(imgu)
    ema0, ema1, ema2 = heads(out_ema, bs)
    const = []
    for a, b in ((out20, ema0), (out21, ema1), (out22, ema2)):
        dist = softmax_mse_loss(None, a, b)
        const.append(cw * ((dist * mask_u).sum() / (mask_u.sum() + 1e-16)))
    uncertain_loss = (const[0] + const[1] + const[2]) / 3
    if it < 1000:
        certain_loss = torch.tensor(0.)
        uncertain_loss = torch.tensor(0.)
    loss = supervised_loss + certain_loss + uncertain_loss * 0.1
    optimizer.zero_grad()
    loss.backward()
    record(tag, (supervised_loss, certain_loss, uncertain_loss))
    optimizer.step()
    update_ema_variable(it)
    checksums(tag)

np.savez_compressed(os.path.join(HERE, "coranet.npz"), **fix)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in fix.items() if "losses" in k}, {k: fix[k] for k in fix if "losses" in k})
