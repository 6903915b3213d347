# -*- coding: utf-8 -*-
"""Drop-in counterpart of the reference's network/blocks.py: same helper / class names, constructor signatures and
state_dict keys (conv1, bn1, conv2, bn2, shortcut1, shortcut2, downsample.{0,1}, up, up.1, pre_conv, pre_bn,
layer1..5, fc), with every forward running on the sm_100a kernels of libsmsut_b200 (NHWC bf16 internally).

Module boundaries speak logical NCHW like the reference; tensors handed between these modules are bf16 views
with channels-last strides, so crossing a boundary costs nothing.
"""
import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops
from ..functional import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, to_nchw, to_nhwc


def _act_code(act):
    if act is None:
        return ACT_NONE
    if isinstance(act, nn.LeakyReLU):
        if abs(act.negative_slope - Fn.SLOPE) > 1e-12:
            raise NotImplementedError("the fused kernels implement LeakyReLU with slope 0.01 (the reference's value)")
        return ACT_LRELU
    if isinstance(act, nn.ReLU):
        return ACT_RELU
    raise NotImplementedError(type(act))


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameters; forward on tcgen05 (1x1 / 3x3, stride 1, 'same', Cout % 16 == 0) or on the direct
    CUDA-core kernel (stems and heads: tiny K, HBM-bound).  The choice is static per layer."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, out_pad=None,
                 out_f32=False, fused_act=ACT_NONE, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias, **kw)
        k = self.kernel_size[0]
        assert self.kernel_size[0] == self.kernel_size[1] and self.groups == 1 and self.dilation == (1, 1)
        self.tensor_core = (k in (1, 3, 5) and self.stride == (1, 1) and self.padding == (k // 2, k // 2)
                            and out_channels % 8 == 0 and not bias and fused_act == ACT_NONE and not out_f32)
        self.out_pad, self.out_f32, self.fused_act = out_pad, out_f32, fused_act
        self._pw = None

    def packed(self):
        w = self.weight
        if self._pw is None or self._pw.weight is not w or self._pw.fprop.device != w.device:
            self._pw = ops.PackedWeight(w)
            self._table = ops.PackTable([self._pw])
        return self._pw

    def ensure_packed(self):
        pw = self.packed()
        if pw.stale():
            self._table.refresh(force=True)
        return pw

    def forward_nhwc(self, xs, with_stats=False):
        """xs: list of NHWC tensors (channel-concatenated input without the concat).  with_stats: also return the
        InstanceNorm statistics of the output (fused into the conv epilogue on the wide layers), or None."""
        if self.tensor_core:
            y, st = Fn.ConvFn.apply(with_stats, self.ensure_packed(), self.weight, None, None, *xs)
            return (y, st) if with_stats else y
        assert len(xs) == 1
        y = Fn.DirectConvFn.apply(xs[0], self.weight, self.bias, self.stride[0], self.padding[0], self.fused_act,
                                  self.out_pad, self.out_f32)
        if Fn.ACT_TAPS[0] is not None and self.fused_act in (ACT_LRELU, ACT_RELU):
            Fn.ACT_TAPS[0].append((self, y))        # parity instrumentation: the fused activation's output
        return (y, None) if with_stats else y

    def forward(self, x):
        if not self.tensor_core and self.in_channels == 1 and x.dtype == torch.float32:
            xin = _image_nhwc(x)        # the direct kernel reads the fp32 single-channel image in place
        else:
            xin = to_nhwc(x)
        return to_nchw(self.forward_nhwc([xin]))


class ConvTranspose2d(nn.ConvTranspose2d):
    """nn.ConvTranspose2d(k=2, s=2, bias=False) as a tcgen05 GEMM with a pixel-shuffle store."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, bias=True, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, bias=bias, **kw)
        if self.kernel_size != (2, 2) or self.stride != (2, 2) or bias or in_channels % 16 or out_channels % 16:
            raise NotImplementedError("only ConvTranspose2d(k=2, s=2, bias=False) with channels % 16 == 0 is on the path")
        self._pw = None

    def packed(self):
        w = self.weight
        if self._pw is None or self._pw.weight is not w or self._pw.fprop.device != w.device:
            self._pw = ops.PackedWeight(w, transposed=True)
            self._table = ops.PackTable([self._pw])
        return self._pw

    def ensure_packed(self):
        pw = self.packed()
        if pw.stale():
            self._table.refresh(force=True)
        return pw

    def forward_nhwc(self, x):
        return Fn.ConvTFn.apply(self.ensure_packed(), self.weight, x)

    def forward(self, x):
        return to_nchw(self.forward_nhwc(to_nhwc(x)))


class InstanceNorm2d(nn.InstanceNorm2d):
    """nn.InstanceNorm2d(C, affine=True), eps 1e-5, biased variance, no running stats."""

    def __init__(self, num_features, affine=True, **kw):
        super().__init__(num_features, affine=affine, **kw)
        if not affine or self.track_running_stats or abs(self.eps - 1e-5) > 1e-12:
            raise NotImplementedError("the path uses InstanceNorm2d(affine=True, eps=1e-5) without running stats")

    def forward(self, x):
        xin = to_nhwc(x)
        cp = self.num_features if xin.shape[3] != self.num_features else None
        return to_nchw(Fn.INActFn.apply(xin, self.weight, self.bias, None, None, None, None, ACT_NONE, cp))


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d(C) (affine, running statistics, momentum 0.1, eps 1e-5: `get_norm(..., 'batch')`, the default
    norm of the reference's UNet / Encoder / Decoder signatures) on the InstanceNorm kernels: the batch statistics
    are the per-sample sums pooled over the batch (Fn._bn_stats), so forward and backward reuse the fused
    norm + activation + residual kernels.  Same parameters and buffers as torch's (state_dict compatible)."""
    smsut_batch_norm = True

    def __init__(self, num_features, **kw):
        super().__init__(num_features, **kw)
        if not self.affine or abs(self.eps - 1e-5) > 1e-12:
            raise NotImplementedError("the kernels implement BatchNorm2d(affine=True, eps=1e-5)")

    def forward(self, x):
        xin = to_nhwc(x)
        cp = self.num_features if xin.shape[3] != self.num_features else None
        return to_nchw(Fn.in_act(xin, self, act=ACT_NONE, c_params=cp))


class LeakyReLU(nn.LeakyReLU):
    def forward(self, x):
        return to_nchw(_LReluFn.apply(to_nhwc(x), _act_code(self)))


class ReLU(nn.ReLU):
    def forward(self, x):
        return to_nchw(_LReluFn.apply(to_nhwc(x), ACT_RELU))


class _LReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        y = ops.act_fwd(x, act, Fn.SLOPE)
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy if dy.is_contiguous() else dy.contiguous()
        if torch.is_grad_enabled():
            return Fn.LReluBwdFn.apply(dy, y, ctx.act), None
        return ops.act_bwd(dy, y, act=ctx.act, slope=Fn.SLOPE), None


def refresh_packs(root):
    """Refresh the bf16 weight copies of every tensor-core conv under `root` with ONE launch (no-op when no master
    weight changed).  Root networks call this at the top of forward; the per-layer checks then all pass."""
    packs = [m.packed() for m in root.modules()
             if (isinstance(m, Conv2d) and m.tensor_core) or isinstance(m, ConvTranspose2d)]
    key = tuple(id(p) for p in packs)
    cache = root.__dict__.get("_smsut_pack_table")
    if cache is None or cache[0] != key:
        cache = (key, ops.PackTable(packs))
        root.__dict__["_smsut_pack_table"] = cache
    cache[1].refresh()


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    if stride != 1 or groups != 1 or dilation != 1:
        raise NotImplementedError("the path only uses conv3x3(stride=1, groups=1, dilation=1)")
    return Conv2d(in_planes, out_planes, kernel_size=3, stride=1, padding=1, bias=False)


def conv1x1(in_planes, out_planes, stride=1):
    if stride != 1:
        raise NotImplementedError("the path only uses conv1x1(stride=1)")
    return Conv2d(in_planes, out_planes, kernel_size=1, stride=1, bias=False)


def get_norm(channels, norm_type):
    if norm_type == 'instance':
        return InstanceNorm2d(channels, affine=True)
    elif norm_type == 'batch':
        return BatchNorm2d(channels)
    else:
        raise NotImplementedError


def get_act(act_type, inplace=True, negative=1e-2):
    if act_type == 'relu':
        return ReLU(inplace=inplace)
    elif act_type == 'lrelu':
        return LeakyReLU(negative_slope=negative, inplace=inplace)
    else:
        raise NotImplementedError


class CatPair(tuple):
    """The value of `torch.cat([up, skip], dim=1)` kept as its two halves: the consuming conv reads both sources
    in one K loop, so the concatenated tensor never exists.  `.cat()` materialises it for foreign consumers."""

    def cat(self):
        return torch.cat(list(self), dim=1)


def _sources(x):
    return [to_nhwc(t) for t in x] if isinstance(x, CatPair) else [to_nhwc(x)]


class UpSampleAndConcat(nn.Module):
    def __init__(self, in_ch, out_ch, transposed=True):
        super(UpSampleAndConcat, self).__init__()
        self.transposed = transposed
        if transposed:
            self.up = ConvTranspose2d(in_ch, out_ch, kernel_size=2, stride=2, bias=False)
        else:
            self.up = nn.Sequential(
                nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False),
                conv1x1(in_ch, out_ch)
            )

    def forward(self, x, skip):
        xin = to_nhwc(x)
        if self.transposed:
            up = self.up.forward_nhwc(xin)
        else:
            up = self.up[1].forward_nhwc([Fn.BilinearFn.apply(xin)])
        return CatPair((to_nchw(up), skip))


class BasicBlock(nn.Module):
    def __init__(self, in_ch, out_ch, norm, act, **kwargs):
        super(BasicBlock, self).__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.bn1 = get_norm(out_ch, norm)
        self.relu = get_act(act)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.bn2 = get_norm(out_ch, norm)
        self.downsample = (in_ch != out_ch)
        if self.downsample:
            self.shortcut1 = conv1x1(in_ch, out_ch)
            self.shortcut2 = get_norm(out_ch, norm)

    def forward_nhwc(self, xs):
        act = _act_code(self.relu)
        if self.downsample:
            c1, s1, cs, ss = Fn.ConvFn.apply(True, self.conv1.ensure_packed(), self.conv1.weight,
                                             self.shortcut1.ensure_packed(), self.shortcut1.weight, *xs)
        else:
            assert len(xs) == 1
            (c1, s1), cs, ss = self.conv1.forward_nhwc(xs, with_stats=True), None, None
        a1 = Fn.in_act(c1, self.bn1, act=act, stats_a=s1)
        c2, s2 = self.conv2.forward_nhwc([a1], with_stats=True)
        if self.downsample:
            return Fn.in_act(c2, self.bn2, xb=cs, norm_b=self.shortcut2, act=act, stats_a=s2, stats_b=ss)
        return Fn.in_act(c2, self.bn2, res=xs[0], act=act, stats_a=s2)

    def forward(self, x):
        return to_nchw(self.forward_nhwc(_sources(x)))


class BottleBlock(nn.Module):
    def __init__(self, in_channels, out_channels, norm_type='batch', act_type='relu', stride=1):
        super(BottleBlock, self).__init__()
        assert stride in (1, 2)
        self.conv1 = conv3x3(in_channels, out_channels)
        self.bn1 = get_norm(out_channels, norm_type)
        self.relu = get_act(act_type)
        self.conv2 = conv3x3(out_channels, out_channels)
        self.bn2 = get_norm(out_channels, norm_type)
        self.stride = stride
        self.downsample = None
        if in_channels != out_channels:
            self.downsample = nn.Sequential(
                conv1x1(in_channels, out_channels),
                get_norm(out_channels, norm_type))

    def forward_nhwc(self, x):
        act = _act_code(self.relu)
        identity = Fn.AvgPoolFn.apply(x) if self.stride == 2 else x
        c1, s1 = self.conv1.forward_nhwc([x], with_stats=True)
        out = Fn.in_act(c1, self.bn1, act=act, stats_a=s1)
        if self.stride == 2:
            out = Fn.AvgPoolFn.apply(out)
        out, s2 = self.conv2.forward_nhwc([out], with_stats=True)
        if self.downsample is not None:
            cs, ss = self.downsample[0].forward_nhwc([identity], with_stats=True)
            return Fn.in_act(out, self.bn2, xb=cs, norm_b=self.downsample[1], act=act, stats_a=s2, stats_b=ss)
        return Fn.in_act(out, self.bn2, res=identity, act=act, stats_a=s2)

    def forward(self, x):
        return to_nchw(self.forward_nhwc(to_nhwc(x)))


def _stem(conv, bn, relu, x):
    """5x5 stem conv (tcgen05, 25 taps over a 16-channel zero-padded bf16 input) -> IN -> act; the 8-channel result
    lives in a 16-channel tensor (channels 8..15 are zero) so the next conv reads it directly."""
    y, st = conv.forward_nhwc([x], with_stats=True)
    return Fn.in_act(y, bn, act=_act_code(relu), c_params=bn.num_features if y.shape[3] != bn.num_features else None,
                     stats_a=st)


def _image_nhwc(x):
    """(N,1,H,W) fp32 image -> (N,H,W,1) fp32 view (same memory)"""
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


class Encoder(nn.Module):
    def __init__(self, in_ch, block, width=32, norm='batch', act='lrelu', **kwargs):
        super(Encoder, self).__init__()
        self.pre_conv = Conv2d(in_ch, width // 2, kernel_size=5, stride=1, padding=2, bias=False)
        self.pre_bn = get_norm(width // 2, norm)
        self.pre_relu = get_act(act)

        self.layer1 = block(width // 2, 1 * width, norm, act, **kwargs)
        self.pool1 = nn.MaxPool2d(2, 2)
        self.layer2 = block(1 * width, 2 * width, norm, act, **kwargs)
        self.pool2 = nn.MaxPool2d(2, 2)
        self.layer3 = block(2 * width, 4 * width, norm, act, **kwargs)
        self.pool3 = nn.MaxPool2d(2, 2)
        self.layer4 = block(4 * width, 8 * width, norm, act, **kwargs)
        self.pool4 = nn.MaxPool2d(2, 2)
        self.layer5 = block(8 * width, 16 * width, norm, act, **kwargs)

    def forward(self, x):
        skips = []
        h = _stem(self.pre_conv, self.pre_bn, self.pre_relu, to_nhwc(x))
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            h = layer.forward_nhwc([h])
            h, skip = Fn.MaxPoolSkipFn.apply(h)
            skips.append(to_nchw(skip))
        h = self.layer5.forward_nhwc([h])
        return to_nchw(h), skips


class Decoder(nn.Module):
    def __init__(self, out_ch, block, width=32, norm='batch', act='lrelu', **kwargs):
        super(Decoder, self).__init__()
        self.up4 = UpSampleAndConcat(16 * width, 8 * width)
        self.layer4 = block(16 * width, 8 * width, norm, act, **kwargs)
        self.up3 = UpSampleAndConcat(8 * width, 4 * width)
        self.layer3 = block(8 * width, 4 * width, norm, act, **kwargs)
        self.up2 = UpSampleAndConcat(4 * width, 2 * width)
        self.layer2 = block(4 * width, 2 * width, norm, act, **kwargs)
        self.up1 = UpSampleAndConcat(2 * width, 1 * width)
        self.layer1 = block(2 * width, 1 * width, norm, act, **kwargs)
        self.fc = Conv2d(width, out_ch, kernel_size=1, stride=1, bias=False, out_f32=True)

    def forward(self, x, skips):
        x = self.layer4(self.up4(x, skips[3]))
        x = self.layer3(self.up3(x, skips[2]))
        x = self.layer2(self.up2(x, skips[1]))
        x = self.layer1(self.up1(x, skips[0]))
        x = self.fc(x)
        return x
