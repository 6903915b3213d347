# -*- coding: utf-8 -*-
"""Single-branch U-Net segmenter with the constructor signature, attribute names (`encoder`, `decoder`) and
state_dict keys of the reference's network/unet.py:13-32; the forward runs on libsmsut_b200's kernels (NHWC bf16
activations handed from block to block, fp32 logits out)."""
import torch.nn as nn

from . import blocks


def _init_parameters(net, act_type):
    """He-normal (fan_out) for every conv / transposed conv with the gain of the network's activation, unit gain and
    zero shift for every affine norm -- the initialisation rule of network/unet.py:21-27."""
    gain_of = 'relu' if act_type == 'relu' else 'leaky_relu'
    convs = [m for m in net.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))]
    norms = [m for m in net.modules() if isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d))]
    for conv in convs:
        nn.init.kaiming_normal_(conv.weight, mode='fan_out', nonlinearity=gain_of)
    for norm in norms:
        nn.init.ones_(norm.weight)
        nn.init.zeros_(norm.bias)


class UNet(nn.Module):
    def __init__(self, in_ch, out_ch, base_width=64, norm_type='batch', act_type='relu'):
        super().__init__()
        kind = dict(norm=norm_type, act=act_type)
        self.encoder = blocks.Encoder(in_ch, blocks.BasicBlock, base_width, **kind)
        self.decoder = blocks.Decoder(out_ch, blocks.BasicBlock, base_width, **kind)
        _init_parameters(self, act_type)

    def forward(self, x):
        blocks.refresh_packs(self)          # bf16 weight copies follow the fp32 masters (one launch, usually a no-op)
        bottleneck, skips = self.encoder(x)
        return self.decoder(bottleneck, skips)
