# -*- coding: utf-8 -*-
"""Drop-in counterpart of the reference's misc/loss.py (DiceAndCrossEntropyLoss / SoftDiceLoss, loss.py:8-63):
softmax, one-hot, tp/fp/fn and the cross entropy are one fused kernel pass over the logits (forward) and one
(backward), with 128-bit accesses on channels-last fp32 logits."""
import torch
import torch.nn as nn

from .. import functional as Fn


def _flat_logits(x):
    """(N,C,H,W) logits -> (N*H*W, C) fp32, zero-copy for the channels-last heads of our networks"""
    v = x.permute(0, 2, 3, 1)
    if v.dtype != torch.float32:
        v = v.float()
    if not v.is_contiguous():
        v = v.contiguous()
    return v.reshape(-1, x.shape[1])


def get_tp_fp_fn_tn(output, gt, dims=(2, 3)):
    """Soft true / false positive / negative sums of (B, C, H, W) probabilities `output` against (B, H, W) labels `gt`
    over `dims` (loss.py:23-36).  The loss kernels compute the same sums inside their fused pass; this host-side form
    is for callers of the reference's helper (misc/utils.py Meter.collect_dice_by)."""
    with torch.no_grad():
        onehot = torch.zeros_like(output).scatter_(1, gt.to(output.device).long().unsqueeze(1), 1)
    tp = (output * onehot).sum(dim=dims)
    fp = (output * (1. - onehot)).sum(dim=dims)
    fn = ((1. - output) * onehot).sum(dim=dims)
    tn = ((1. - output) * (1. - onehot)).sum(dim=dims)
    return tp, fp, fn, tn


class DiceAndCrossEntropyLoss(nn.Module):
    def __init__(self, weight_ce=1., weight_dc=1., batch_dice=False):
        super(DiceAndCrossEntropyLoss, self).__init__()
        if not batch_dice:
            raise NotImplementedError("the trainers construct the loss with batch_dice=True (baseTrainer.py:57)")
        self.weight_ce = weight_ce
        self.weight_dc = weight_dc
        self.batch_dice = batch_dice

    def forward(self, x, y):
        """x: (B, C, H, W) logits.  y: (B, H, W) int64 labels, or (B, C, H, W) logits whose argmax is the target
        (the fused form of `self.loss(source, torch.argmax(target, dim=1))`)."""
        logits = _flat_logits(x)
        if y.dim() == 4:
            return Fn.DiceCEFn.apply(logits, None, _flat_logits(y.detach()), self.weight_ce, self.weight_dc)
        y = y.reshape(-1)
        if y.dtype != torch.int64:
            y = y.long()
        return Fn.DiceCEFn.apply(logits, y.contiguous(), None, self.weight_ce, self.weight_dc)


class SoftDiceLoss(DiceAndCrossEntropyLoss):
    def __init__(self, batch_dice=False, smooth=1e-5):
        super(SoftDiceLoss, self).__init__(weight_ce=0., weight_dc=1., batch_dice=batch_dice)
        if abs(smooth - 1e-5) > 1e-12:
            raise NotImplementedError("smooth is fixed at the reference's 1e-5")
