# -*- coding: utf-8 -*-
"""Drop-in counterpart of the reference's network/ugan.py: Encoder, Decoder, UGAN, UGANnce, Discriminator,
PatchSampleF, define_F, init_net with the same constructor signatures, forward return arity and state_dict keys
(network/ugan.py:22-339); all arithmetic runs on libsmsut_b200's sm_100a kernels.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import config as cfg
from .. import functional as Fn
from .. import ops
from ..functional import ACT_LRELU, ACT_NONE, ACT_TANH, to_nchw, to_nhwc
from . import networks
from .blocks import (BasicBlock, BottleBlock, CatPair, Conv2d, UpSampleAndConcat, _act_code, _image_nhwc, _stem,
                     conv1x1, conv3x3, get_act, get_norm, refresh_packs)
from .patchnce import PatchNCELoss  # noqa: F401  (re-exported like the reference)


class _TslInputFn(torch.autograd.Function):
    """cat([x, m.view(B,n,1,1).repeat(1,1,H,W)], 1) as an NHWC bf16 tensor (network/ugan.py:154-159); the gradient
    of the image is channel 0 of the input gradient."""

    @staticmethod
    def forward(ctx, x, m):
        return ops.build_tsl_input(x.contiguous(), m.contiguous().float(), 16)

    @staticmethod
    def backward(ctx, d):
        return d[..., 0:1].float().permute(0, 3, 1, 2), None


class Encoder(nn.Module):
    def __init__(self, in_ch, base_width=32, norm_type='batch', act_type='relu'):
        super(Encoder, self).__init__()
        self.pre = nn.Sequential(
            Conv2d(in_ch, base_width // 2, kernel_size=5, stride=1, padding=2, bias=False),
            get_norm(base_width // 2, norm_type),
            get_act(act_type)
        )
        self.enc1 = BasicBlock(base_width // 2, base_width, norm_type, act_type)
        self.pool1 = nn.MaxPool2d(2, stride=2)  # x2
        self.enc2 = BasicBlock(base_width, 2 * base_width, norm_type, act_type)
        self.pool2 = nn.MaxPool2d(2, stride=2)  # x4
        self.enc3 = BasicBlock(2 * base_width, 4 * base_width, norm_type, act_type)
        self.pool3 = nn.MaxPool2d(2, stride=2)  # x8
        self.enc4 = BasicBlock(4 * base_width, 8 * base_width, norm_type, act_type)
        self.pool4 = nn.MaxPool2d(2, stride=2)  # x16

    def forward_nhwc(self, xin):
        """xin: (N,H,W,16) bf16 zero-padded image (ImageInputFn) or translation input (_TslInputFn)"""
        retn = []
        h = _stem(self.pre[0], self.pre[1], self.pre[2], xin)
        for enc in (self.enc1, self.enc2, self.enc3, self.enc4):
            h = enc.forward_nhwc([h])
            h, skip = Fn.MaxPoolSkipFn.apply(h)
            retn.append(to_nchw(skip))
        retn.reverse()
        return h, retn

    def forward(self, x):
        h, retn = self.forward_nhwc(to_nhwc(x))
        return to_nchw(h), retn


class Decoder(nn.Module):
    def __init__(self, out_ch, base_width=32, norm_type='batch', act_type='relu', tranposed=True, use_tanh=False):
        super(Decoder, self).__init__()
        self.up4 = UpSampleAndConcat(16 * base_width, 8 * base_width, transposed=tranposed)  # x8
        self.dec4 = BasicBlock(16 * base_width, 8 * base_width, norm_type, act_type)
        self.up3 = UpSampleAndConcat(8 * base_width, 4 * base_width, transposed=tranposed)  # x4
        self.dec3 = BasicBlock(8 * base_width, 4 * base_width, norm_type, act_type)
        self.up2 = UpSampleAndConcat(4 * base_width, 2 * base_width, transposed=tranposed)  # x2
        self.dec2 = BasicBlock(4 * base_width, 2 * base_width, norm_type, act_type)
        self.up1 = UpSampleAndConcat(2 * base_width, base_width, transposed=tranposed)  # x1
        self.dec1 = BasicBlock(2 * base_width, base_width, norm_type, act_type)

        # 1x1 head with bias; tanh (when asked for) is fused into the head kernel's epilogue
        self.fc = Conv2d(base_width, out_ch, 1, bias=True, out_f32=True, fused_act=ACT_TANH if use_tanh else ACT_NONE)
        self.tanh = None
        if use_tanh:
            self.tanh = nn.Tanh()

    def forward(self, e5, x_ens):
        d4 = self.dec4(self.up4(e5, x_ens[0]))
        d3 = self.dec3(self.up3(d4, x_ens[1]))
        d2 = self.dec2(self.up2(d3, x_ens[2]))
        d1 = self.dec1(self.up1(d2, x_ens[3]))
        out = self.fc(d1)
        return out


def _kaiming_init(module):
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='leaky_relu')
        elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class UGAN(nn.Module):
    def __init__(self, in_ch, out_ch, n_modal, base_width=32):
        super(UGAN, self).__init__()
        self.n_modal = n_modal
        self.tsl_encoder = Encoder(in_ch + n_modal, base_width, norm_type='instance', act_type='lrelu')
        self.seg_encoder = Encoder(in_ch, base_width, norm_type='instance', act_type='lrelu')

        self.enc5 = BasicBlock(8 * base_width, 16 * base_width, norm='instance', act='lrelu')

        self.tsl_decoder = Decoder(1, base_width, norm_type='instance', act_type='lrelu',
                                   tranposed=False, use_tanh=True)
        self.seg_decoder = Decoder(out_ch, base_width, norm_type='instance', act_type='lrelu',
                                   tranposed=True, use_tanh=False)
        _kaiming_init(self)

    def _seg_half(self, xin):
        seg_out, seg_ens = self.seg_encoder.forward_nhwc(xin)
        seg_out = self.enc5.forward_nhwc([seg_out])
        return self.seg_decoder(to_nchw(seg_out), seg_ens)

    def _branches(self, x, m, seg_rows=None, seg_rest=True, want_seg=True):
        """seg_rows = r: the segmentation half runs as two independent sub-batches -- slices [:r] with an autograd
        graph, slices [r:] without one (or not at all when seg_rest is False) -- and `seg` is returned as the pair.
        Every layer of the generator is per-sample (InstanceNorm), so the values equal the full-batch forward; the
        trainer uses it where only the labelled slices' logits are differentiated (uganConsisTrainer.py:151-155:
        g_loss_seg = loss(y_fake[:bs], y_real); the other slices' logits are only argmax targets), which halves that
        branch's backward instead of pushing zeros through it."""
        if x.shape[1] != 1:
            raise NotImplementedError("the path translates single-channel slices (cfg.img_channels = 1)")
        refresh_packs(self)
        if m is None:
            m = torch.zeros(x.size(0), self.n_modal, device=x.device)
        x = x.float()
        # the segmentation half runs on a branch stream beside the translation half (they share only weights); a
        # forward that itself runs inside a branch (the cycle pass of the trainer) uses another stream, so that the
        # backward of the first pass's segmentation half is not queued behind the second pass's
        if not want_seg:
            seg, joins = None, []       # the caller discards the logits (cycle pass while the consistency loss is off)
        elif seg_rows is None:
            with ops.parallel_branch(0 if ops.current_branch() is None else 3) as br:
                seg = self._seg_half(Fn.ImageInputFn.apply(x))
            joins = [(br, seg)]
        else:
            xin = Fn.ImageInputFn.apply(x)            # before the forks: both sub-batches read it
            with ops.parallel_branch(0 if ops.current_branch() is None else 3) as br:
                seg_a = self._seg_half(xin[:seg_rows])
            joins, seg_b = [(br, seg_a)], None
            if seg_rest:
                with ops.parallel_branch(6) as br2:
                    with torch.no_grad():
                        seg_b = self._seg_half(xin[seg_rows:])
                joins.append((br2, seg_b))
            seg = (seg_a, seg_b)

        tsl_in = _TslInputFn.apply(x, m)
        tsl_out, tsl_ens = self.tsl_encoder.forward_nhwc(tsl_in)
        tsl_out_1 = self.enc5.forward_nhwc([tsl_out])
        tsl = self.tsl_decoder(to_nchw(tsl_out_1), tsl_ens)
        for b, t in joins:
            b.join(t)
        return seg, tsl, tsl_out_1

    def forward(self, x, m=None):
        seg, tsl, _ = self._branches(x, m)
        return seg, tsl


class UGANnce(UGAN):
    def __init__(self, in_ch, out_ch, n_modal, base_width=32, val_phase=False):
        nn.Module.__init__(self)
        self.n_modal = n_modal
        self.val_phase = val_phase
        self.tsl_encoder = Encoder(in_ch + n_modal, base_width, norm_type='instance', act_type='lrelu')
        self.seg_encoder = Encoder(in_ch, base_width, norm_type='instance', act_type='lrelu')

        self.enc5 = BasicBlock(8 * base_width, 16 * base_width, norm='instance', act='lrelu')

        self.netF = define_F(in_ch, netF_nc=256)
        if not self.netF.mlp_init:
            self.netF.create_mlp(cfg.nce_layers, input_nc=16 * base_width)

        self.tsl_decoder = Decoder(1, base_width, norm_type='instance', act_type='lrelu',
                                   tranposed=False, use_tanh=True)
        self.seg_decoder = Decoder(out_ch, base_width, norm_type='instance', act_type='lrelu',
                                   tranposed=True, use_tanh=False)
        _kaiming_init(self)

    def forward(self, x, m=None, sample_ids=None, val_phase=False, seg_rows=None, seg_rest=True, want_seg=True):
        """seg_rows / seg_rest (extension, see UGAN._branches): `seg` comes back as (seg[:r] with graph, seg[r:]
        without graph or None).  want_seg=False: the segmentation half is not run and `seg` is None."""
        seg, tsl, tsl_out_1 = self._branches(x, m, seg_rows, seg_rest, want_seg)
        if val_phase:
            return seg, tsl
        feats = [to_nchw(tsl_out_1)]
        if sample_ids is None:
            feat_pool, sample_ids = self.netF(feats)
        else:
            feat_pool, _ = self.netF(feats, patch_ids=sample_ids)
        return seg, tsl, feat_pool, sample_ids


class Discriminator(nn.Module):
    def __init__(self, input_size, n_modal, base_width=32, max_width=512):
        super(Discriminator, self).__init__()
        blocks = []
        # stem: 4x4 s2 conv + bias with the LeakyReLU fused into the kernel epilogue (main.1 stays for the
        # state_dict / Sequential index layout)
        blocks += [Conv2d(1, base_width, kernel_size=4, stride=2, padding=1, out_pad=ops.pad16(base_width),
                          fused_act=ACT_LRELU),
                   nn.LeakyReLU(inplace=True)]

        repeat_num = int(np.log2(input_size)) - 2
        in_width = base_width
        for _ in range(1, repeat_num):
            out_width = min(in_width * 2, max_width)
            blocks += [BottleBlock(in_width, out_width, norm_type='instance', act_type='lrelu', stride=2)]
            in_width = out_width
        self.main = nn.Sequential(*blocks)

        kernel_size = int(input_size / np.power(2, repeat_num))
        self.conv_src = Conv2d(out_width, 1, kernel_size=3, stride=1, padding=1, bias=False, out_f32=True)
        self.conv_cls = Conv2d(out_width, n_modal, kernel_size=kernel_size, bias=False, out_f32=True)
        _kaiming_init(self)

    def forward(self, x):
        if x.shape[1] != 1:
            raise NotImplementedError("the discriminator judges single-channel slices")
        refresh_packs(self)
        out = self.main[0].forward_nhwc([_image_nhwc(x.float())])   # conv + bias + LeakyReLU
        for blk in list(self.main)[2:]:
            out = blk.forward_nhwc(out)
        out_src = to_nchw(self.conv_src.forward_nhwc([out]))
        out_cls = to_nchw(self.conv_cls.forward_nhwc([out]))
        return out_src, out_cls.reshape(out_cls.size(0), out_cls.size(1))


def define_F(input_nc, netF='mlp_sample', norm='batch', use_dropout=False, init_type='normal',
             init_gain=0.02,
             no_antialias=False, gpu_ids=None, netF_nc=256):
    if gpu_ids is None:
        gpu_ids = []
    net = PatchSampleF(use_mlp=True, init_type=init_type, init_gain=init_gain, gpu_ids=gpu_ids, nc=netF_nc)
    return init_net(net, init_type, init_gain, gpu_ids)


def init_net(net, init_type='normal', init_gain=0.02, gpu_ids=[], debug=False, initialize_weights=True):
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
        net.to(gpu_ids[0])

    if initialize_weights:
        networks.init_weights(net, init_type, init_gain=init_gain, debug=debug)
    return net


class _Linear(nn.Linear):
    """nn.Linear parameters; the GEMM runs as a 1x1 tcgen05 conv over the sampled rows."""

    def packed(self):
        w = self.weight
        pw = self.__dict__.get("_pw")
        if pw is None or pw.src_ptr != w.data_ptr() or pw.fprop.device != w.device:
            pw = ops.PackedWeight(w.detach().view(w.shape[0], w.shape[1], 1, 1))
            pw.src_ptr = w.data_ptr()
            self.__dict__["_pw"] = pw
            self.__dict__["_table"] = ops.PackTable([pw])
        if pw.stale():
            self.__dict__["_table"].refresh(force=True)
        return pw


class PatchSampleF(nn.Module):
    def __init__(self, use_mlp=False, init_type='normal', init_gain=0.02, nc=256, gpu_ids=[]):
        super(PatchSampleF, self).__init__()
        self.l2norm = networks.Normalize(2)
        self.use_mlp = use_mlp
        self.nc = nc
        self.mlp_init = False
        self.init_type = init_type
        self.init_gain = init_gain
        self.gpu_ids = gpu_ids

    def create_mlp(self, nce_layers, input_nc=256):
        for mlp_id, layer in enumerate(nce_layers):
            mlp = nn.Sequential(*[_Linear(input_nc, self.nc), nn.ReLU(), _Linear(self.nc, self.nc)])
            if len(self.gpu_ids) > 0:
                mlp.cuda()
            setattr(self, 'mlp_%d' % mlp_id, mlp)
        init_net(self, self.init_type, self.init_gain, self.gpu_ids)
        self.mlp_init = True

    def forward(self, feats, num_patches=64, patch_ids=None):
        return_ids = []
        return_feats = []
        if not self.use_mlp or num_patches <= 0:
            raise NotImplementedError("the path samples patches and projects them with the MLP (use_mlp=True)")
        if not self.mlp_init:
            self.create_mlp(cfg.nce_layers)
        for feat_id, feat in enumerate(feats):
            f = to_nhwc(feat)
            hw = f.shape[1] * f.shape[2]
            if patch_ids is not None:
                patch_id = patch_ids[feat_id]
            else:
                patch_id = torch.randperm(hw, device=feat.device)
                patch_id = patch_id[:int(min(num_patches, patch_id.shape[0]))]
            mlp = getattr(self, 'mlp_%d' % feat_id)
            if mlp[0].weight.device != feat.device:
                mlp.to(feat.device)
            l1, l2 = mlp[0], mlp[2]
            x_sample = Fn.PatchSampleFn.apply(f, patch_id.contiguous(), l1.packed(), l1.weight, l1.bias, l2.packed(),
                                              l2.weight, l2.bias)
            return_ids.append(patch_id)
            return_feats.append(x_sample)
        return return_feats, return_ids
