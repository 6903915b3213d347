"""Drop-in counterpart of the reference's network/patchnce.py (PatchNCELoss, network/patchnce.py:6-51)."""
import torch
from torch import nn

from .. import functional as Fn


class PatchNCELoss(nn.Module):
    def __init__(self, batch_size):
        super().__init__()
        self.batch_size = batch_size

    def forward(self, feat_q, feat_k):
        # l_pos = <q_r, k_r>; negatives = the other rows of the same group of N/batch_size rows, own row -> -10;
        # logits / 0.07; cross entropy against class 0, reduction 'none'.  One fused kernel (no logits tensor).
        feat_k = feat_k.detach()
        q = feat_q if feat_q.is_contiguous() else feat_q.contiguous()
        k = feat_k if feat_k.is_contiguous() else feat_k.contiguous()
        if q.shape[0] % self.batch_size:
            raise ValueError("number of feature rows must be divisible by batch_size")
        return Fn.PatchNCEFn.apply(q.float(), k.float(), self.batch_size)
