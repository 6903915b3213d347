# -*- coding: utf-8 -*-
"""Drop-in counterpart of the reference's network/unet.py (UNet, network/unet.py:13-32)."""
import torch
import torch.nn as nn

from .blocks import BasicBlock, Encoder, Decoder, refresh_packs


class UNet(nn.Module):
    def __init__(self, in_ch, out_ch, base_width=64,
                 norm_type='batch', act_type='relu'):
        super(UNet, self).__init__()

        self.encoder = Encoder(in_ch, BasicBlock, base_width, norm=norm_type, act=act_type)
        self.decoder = Decoder(out_ch, BasicBlock, base_width, norm=norm_type, act=act_type)

        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight, mode='fan_out',
                                        nonlinearity='relu' if act_type == 'relu' else 'leaky_relu')
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        refresh_packs(self)
        x, skips = self.encoder(x)
        x = self.decoder(x, skips)
        return x
