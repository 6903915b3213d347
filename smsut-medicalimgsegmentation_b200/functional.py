"""Autograd wiring of the C-ABI kernels.

Every Function works on NHWC bf16 activations (fp32 at the network heads) and launches only libsmsut_b200
kernels.  First-order backwards use the fused kernels.  When autograd runs a backward with create_graph=True
(the WGAN-GP term, trainer/uganShp0Trainer.py:127-134) the same Functions emit a *differentiable* backward made
of the `*BwdFn` Functions below, whose own backwards are the hand-derived second-order kernels
(smsut_in_bwd2_*, conv fprop/wgrad of the cotangent, avg-pool forward of the cotangent ...).
"""
import os
import threading

import torch
from torch.autograd import Function

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, BF16, F32

SLOPE = 0.01


class _Mode(threading.local):
    inputs_only = False  # create_graph backward: skip parameter gradients (only d/d(input) is requested)


mode = _Mode()
_inputs_only_global = [False]  # the autograd engine runs backward on its own thread: use a process-wide flag


class inputs_only:
    """Context for torch.autograd.grad(..., inputs=x, create_graph=True): parameter gradients of the first-order
    pass are not requested (gradient_penalty), so the differentiable backward skips them."""

    def __enter__(self):
        self.prev = _inputs_only_global[0]
        _inputs_only_global[0] = True

    def __exit__(self, *a):
        _inputs_only_global[0] = self.prev


_accumulate = [False]


class accumulate_param_grads:
    """Inside this context (the trainers' backward calls) the weight-gradient kernels accumulate straight into the
    parameters' existing `.grad` tensors (the flat gradient buffer of optim.py) and autograd is handed `None`:
    no zero-fill + add pair per parameter.  Outside it the Functions return fresh gradient tensors as usual."""

    def __init__(self, side_streams=True, join=True, flush=True, side_group=0):
        """join=False leaves the side / branch streams of this backward running (no wait, no scratch flush): a later
        `accumulate_param_grads()` block -- or the optimizer step -- completes the gradients.  flush=False joins but
        leaves the tap-major scratch to the optimizer step (another network's weight-gradient kernels may still be
        in flight, and flush_all would fold THEIR scratch too).  side_group: see ops.side_streams_enable."""
        self.side = side_streams
        self.join = join
        self.flush = flush
        self.group = side_group

    def __enter__(self):
        self.prev = _accumulate[0]
        _accumulate[0] = True
        ops.side_streams_enable(self.side, self.group)     # wgrad kernels run beside the dgrad chain (ops._Side)

    def __exit__(self, *a):
        _accumulate[0] = self.prev
        ops.side_streams_enable(False)
        if not self.join:
            return
        ops.branch_join_all()
        ops.side_join()
        if self.flush:
            ops.WgradScratch.flush_all()       # after the block every .grad is complete (tap-major scratch folded in)


def _target(param):
    """the tensor to accumulate a parameter gradient into, or None (return the gradient to autograd)"""
    if not _accumulate[0] or param is None:
        return None
    g = param.grad
    if g is None or g.dtype != F32 or not g.is_contiguous() or g.shape != param.shape:
        return None
    return g


def _ret(grad, target):
    return None if target is not None else grad


def _c(t):
    """contiguous view of a gradient (autograd may hand over expanded / permuted tensors)"""
    return t if t.is_contiguous() else t.contiguous()


def to_nhwc(x):
    """logical NCHW tensor -> NHWC bf16 contiguous.  Zero-copy for the channels-last bf16 views our modules
    return; fp32 NCHW inputs go through the layout kernel."""
    if x.dim() != 4:
        raise ValueError("expected a 4-D NCHW tensor")
    if x.dtype == BF16:
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous():
            return v
        return v.contiguous()
    if x.dtype == F32:
        n, c, h, w = x.shape
        if c <= 16:
            return ImageInputFn.apply(x)
        if x.requires_grad:
            return x.permute(0, 2, 3, 1).to(BF16).contiguous()
        return ops.nchw_to_nhwc(x.contiguous(), ops.pad16(c))
    raise TypeError(f"unsupported activation dtype {x.dtype}")


class ImageInputFn(Function):
    """(N, C<=16, H, W) fp32 network input -> (N, H, W, 16) bf16 (zero channel padding) for the tensor-core stem;
    the gradient of the image is the leading C channels of the stem's input gradient."""

    @staticmethod
    def forward(ctx, x):
        ctx.c = x.shape[1]
        return ops.nchw_to_nhwc(x.contiguous(), 16)

    @staticmethod
    def backward(ctx, d):
        return d[..., :ctx.c].float().permute(0, 3, 1, 2)


def to_nchw(y):
    """NHWC tensor -> logical NCHW view (channels-last strides, no copy)"""
    return y.permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------
# tensor-core convolutions
# ----------------------------------------------------------------------------------------------
class ConvDgradFn(Function):
    """dx = dgrad(dy; W) as a differentiable op (single source)."""

    @staticmethod
    def forward(ctx, pw, weight, dy):
        ctx.pw = pw
        ctx.pw_weight = weight
        ctx.save_for_backward(dy)
        return ops.conv_dgrad(dy, pw)[0]

    @staticmethod
    def backward(ctx, u):
        (dy,) = ctx.saved_tensors
        u = _c(u)
        g_dy = ops.conv_fprop([u], ctx.pw) if ctx.needs_input_grad[2] else None
        g_w = None
        if ctx.needs_input_grad[1]:
            t = _target(ctx.pw_weight)
            g_w = _ret(ops.conv_wgrad([u], dy, ctx.pw, out=t), t)
        return None, g_w, g_dy


class ConvFn(Function):
    """y = conv(cat(xs), W): 1x1 / 3x3, stride 1, 'same'.  A second (pw2, weight2) pair computes a second conv of
    the same input in the same node (BasicBlock's conv1 + shortcut1) so the input gets ONE gradient."""

    @staticmethod
    def forward(ctx, want_stats, pw, weight, pw2, weight2, *xs):
        """returns (y, stats) or (y, stats, y2, stats2): stats = InstanceNorm statistics of the conv output
        (fused into the conv epilogue on the wide layers; non-differentiable), or an empty tensor if not wanted"""
        ctx.pw, ctx.pw2 = pw, pw2
        ctx.save_for_backward(weight, weight2, *xs)
        # the statistics outputs never carry a gradient: without this autograd materialises a zero tensor for each of
        # them in every backward (~200 fill launches per iteration in the ncu launch list)
        ctx.set_materialize_grads(False)
        if want_stats:
            y, st = ops.conv_fprop(list(xs), pw, want_stats=True)
        else:
            y, st = ops.conv_fprop(list(xs), pw), torch.empty(0, device=xs[0].device)
        ctx.mark_non_differentiable(st)
        if pw2 is None:
            return y, st
        y2, st2 = ops.conv_fprop(list(xs), pw2, want_stats=True)
        ctx.mark_non_differentiable(st2)
        return y, st, y2, st2

    @staticmethod
    def backward(ctx, dy, _ds=None, dy2=None, _ds2=None):
        weight, weight2, *xs = ctx.saved_tensors
        pw, pw2 = ctx.pw, ctx.pw2
        if dy is None:      # (set_materialize_grads(False)) y itself unused: only the second conv contributes
            dy = torch.zeros((*xs[0].shape[:3], pw.cout_pad), dtype=BF16, device=xs[0].device)
        if pw2 is not None and dy2 is None:
            dy2 = torch.zeros((*xs[0].shape[:3], pw2.cout_pad), dtype=BF16, device=xs[0].device)
        dy = _c(dy)
        need_x = any(ctx.needs_input_grad[5:])
        splits = [x.shape[3] for x in xs]
        if torch.is_grad_enabled():
            assert pw2 is None and len(xs) == 1, "double backward is implemented for single-source convs"
            dx = ConvDgradFn.apply(pw, weight, dy) if need_x else None
            dw = None
            if ctx.needs_input_grad[2] and not _inputs_only_global[0]:
                dw = ops.conv_wgrad(xs, dy.detach(), pw)
            return None, None, dw, None, None, dx
        dxs = [None] * len(xs)
        if need_x:
            dxs = ops.conv_dgrad(dy, pw, splits)
        dw = dw2 = None
        if ctx.needs_input_grad[2]:
            t = _target(weight)
            dw = _ret(ops.conv_wgrad(xs, dy, pw, out=t), t)
        if pw2 is not None:
            dy2 = _c(dy2)
            if need_x:
                ops.conv_dgrad_accumulate(dy2, pw2, dxs)
            if ctx.needs_input_grad[4]:
                t = _target(weight2)
                dw2 = _ret(ops.conv_wgrad(xs, dy2, pw2, out=t), t)
        return (None, None, dw, None, dw2, *dxs)


class ConvTFn(Function):
    """nn.ConvTranspose2d(k=2, s=2, bias=False) (network/blocks.py:41)"""

    @staticmethod
    def forward(ctx, pw, weight, x):
        ctx.pw = pw
        ctx.save_for_backward(weight, x)
        return ops.convt_fprop(x, pw)

    @staticmethod
    def backward(ctx, dy):
        weight, x = ctx.saved_tensors
        dy = _c(dy)
        dx = ops.convt_dgrad(dy, ctx.pw) if ctx.needs_input_grad[2] else None
        dw = None
        if ctx.needs_input_grad[1]:
            t = _target(weight)
            dw = _ret(ops.convt_wgrad(x, dy, ctx.pw, out=t), t)
        return None, dw, dx


# ----------------------------------------------------------------------------------------------
# direct convolutions (stems and heads)
# ----------------------------------------------------------------------------------------------
class DirectDgradFn(Function):
    @staticmethod
    def forward(ctx, weight, g, x_shape, x_dtype, stride, pad):
        ctx.meta = (stride, pad, g.shape[3], g.dtype)
        ctx.save_for_backward(weight, g)
        return ops.conv_direct_dgrad(g, weight, x_shape, x_dtype, stride, pad)

    @staticmethod
    def backward(ctx, u):
        weight, g = ctx.saved_tensors
        stride, pad, gc, gdt = ctx.meta
        u = _c(u)
        g_g = None
        if ctx.needs_input_grad[1]:
            g_g = ops.conv_direct_fprop(u, weight, stride, pad, out_c=gc, out_f32=(gdt == F32))
        g_w = None
        if ctx.needs_input_grad[0]:
            t = _target(weight)
            g_w, _ = ops.conv_direct_wgrad(u, g, weight, stride, pad, want_bias=False, dw=t)
            g_w = _ret(g_w, t)
        return g_w, g_g, None, None, None, None


class DirectConvFn(Function):
    """Stem / head convolutions on CUDA cores (tiny K, HBM-bound): x NHWC (bf16 or fp32), fp32 OIHW weights,
    optional bias and fused LeakyReLU / tanh."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, act, out_c, out_f32):
        y = ops.conv_direct_fprop(x, weight, stride, pad, bias=bias, act=act, slope=SLOPE, out_c=out_c, out_f32=out_f32)
        ctx.meta = (stride, pad, act)
        ctx.save_for_backward(x, weight, y if act != ACT_NONE else None, bias)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y, bias = ctx.saved_tensors
        stride, pad, act = ctx.meta
        dy = _c(dy)
        diff = torch.is_grad_enabled()
        tw, tb = _target(weight), _target(bias)
        if (not diff and weight.shape[2] == 1 and weight.shape[3] == 1 and stride == 1 and pad == 0 and x.dtype == BF16
                and x.shape[3] == 16 and weight.shape[1] == 16 and weight.shape[0] <= 8 and dy.dtype == F32
                and act in (ACT_NONE, ACT_TANH)):
            # 1x1 head: dx, dW and dbias in one fused pass
            dx, dw, db = ops.head1x1_bwd(x, dy, y if act == ACT_TANH else None, weight, ctx.needs_input_grad[0], dw=tw,
                                         db=tb, want_bias=ctx.has_bias)
            return dx, _ret(dw, tw), (_ret(db, tb) if ctx.has_bias else None), None, None, None, None, None
        if act == ACT_TANH:
            assert not diff, "double backward through the tanh head is not on the path"
            g = ops.tanh_bwd(dy, y)
        elif act != ACT_NONE:
            g = LReluBwdFn.apply(dy, y, act) if diff else ops.act_bwd(dy, y, act=act, slope=SLOPE)
        else:
            g = dy
        dx = dw = db = None
        if diff:
            if ctx.needs_input_grad[0]:
                dx = DirectDgradFn.apply(weight, g, tuple(x.shape), x.dtype, stride, pad)
            if not _inputs_only_global[0] and ctx.needs_input_grad[1]:
                dw, db = ops.conv_direct_wgrad(x, g.detach(), weight, stride, pad, want_bias=ctx.has_bias)
        else:
            if ctx.needs_input_grad[0]:
                dx = ops.conv_direct_dgrad(g, weight, tuple(x.shape), x.dtype, stride, pad)
            if ctx.needs_input_grad[1]:
                dw, db = ops.conv_direct_wgrad(x, g, weight, stride, pad, want_bias=ctx.has_bias, dw=tw, db=tb)
                dw, db = _ret(dw, tw), (_ret(db, tb) if ctx.has_bias else None)
        return dx, dw, db, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# InstanceNorm + activation + residual
# ----------------------------------------------------------------------------------------------
RECOMPUTE_SIGN = [os.environ.get("SMSUT_IN_BWD_RECOMPUTE", "1") != "0"]   # A/B knob for the InstanceNorm backward
BN_OFF, BN_TRAIN, BN_EVAL = 0, 1, 2     # INActFn `batch` modes: InstanceNorm / BatchNorm training / BatchNorm eval


class LReluBwdFn(Function):
    """g = dy * act'(ref)  (ref = the activation's output; piecewise linear => second derivative is zero)"""

    @staticmethod
    def forward(ctx, dy, ref, act):
        ctx.act = act
        ctx.save_for_backward(ref)
        return ops.act_bwd(dy, ref, act=act, slope=SLOPE)

    @staticmethod
    def backward(ctx, u):
        (ref,) = ctx.saved_tensors
        return ops.act_bwd(_c(u), ref, act=ctx.act, slope=SLOPE), None, None


class INBwdFn(Function):
    """dx = gamma*rstd*(g - mean g - xhat*mean(g*xhat)) as a differentiable op (stats are functions of x)."""

    @staticmethod
    def forward(ctx, g, x, stats, gamma):
        ctx.save_for_backward(g, x, stats, gamma)
        return ops.in_bwd(g, None, x, stats, gamma, act=ACT_NONE)[0]

    @staticmethod
    def backward(ctx, u):
        g, x, stats, gamma = ctx.saved_tensors
        g_g, g_x, g_gamma = ops.in_bwd2(_c(u), g, x, stats, gamma)
        return g_g, g_x, None, g_gamma


class INActFn(Function):
    """out = act( IN(xa; ga, ba) [+ IN(xb; gb, bb)] [+ res] )  (network/blocks.py:66-80, 99-117)"""

    @staticmethod
    def forward(ctx, xa, ga, ba, xb, gb, bb, res, act, c_params, sa=None, sb=None, batch=BN_OFF):
        if sa is None:
            sa = ops.in_stats(xa)
        if xb is not None and sb is None:
            sb = ops.in_stats(xb)
        out = ops.in_apply(xa, sa, ga, ba, xb, sb, gb, bb, res=res, act=act, slope=SLOPE, c_params=c_params)
        ctx.act, ctx.cp, ctx.has_res, ctx.batch = act, c_params, res is not None, batch
        ctx.save_for_backward(xa, sa, ga, xb, sb, gb, out, ba, bb)
        return out

    @staticmethod
    def backward(ctx, dout):
        xa, sa, ga, xb, sb, gb, out, ba, bb = ctx.saved_tensors
        dout = _c(dout)
        act, cp = ctx.act, ctx.cp
        want_res = ctx.has_res and ctx.needs_input_grad[6]
        if ctx.batch == BN_EVAL:
            raise NotImplementedError("backward through BatchNorm2d in eval mode (running statistics) is not on the path")
        if torch.is_grad_enabled():
            if ctx.batch != BN_OFF:
                raise NotImplementedError("double backward through BatchNorm2d is not on the path (the WGAN-GP "
                                          "discriminator uses InstanceNorm)")
            assert cp is None or cp == xa.shape[3]
            g = LReluBwdFn.apply(dout, out, act) if act != ACT_NONE else dout
            dxa = INBwdFn.apply(g, xa, sa, ga)
            dxb = INBwdFn.apply(g, xb, sb, gb) if xb is not None else None
            dga = dba = dgb = dbb = None
            if not _inputs_only_global[0]:
                _, dga, dba, _, dgb, dbb, _ = ops.in_bwd(dout.detach(), out, xa, sa, ga, xb, sb, gb, False, act, SLOPE, cp)
            return dxa, dga, dba, dxb, dgb, dbb, (g if want_res else None), None, None, None, None, None
        targets = None
        need = ctx.needs_input_grad
        if not (need[1] or need[2] or need[4] or need[5]):
            targets = [None, None, None, None]          # frozen parameters: no parameter gradients at all
        else:
            tg = [_target(ga), _target(ba), _target(gb) if xb is not None else None,
                  _target(bb) if xb is not None else None]
            if tg[0] is not None and tg[1] is not None and (xb is None or (tg[2] is not None and tg[3] is not None)):
                targets = tg
        # no residual input in the forward: the sign of the activation's input is a function of xa / xb alone
        betas = (ba, bb) if (not ctx.has_res and RECOMPUTE_SIGN[0]) else None
        dxa, dga, dba, dxb, dgb, dbb, dres = ops.in_bwd(dout, out, xa, sa, ga, xb, sb, gb, want_res, act, SLOPE, cp,
                                                        targets=targets, batch=ctx.batch == BN_TRAIN, betas=betas)
        return dxa, dga, dba, dxb, dgb, dbb, dres, None, None, None, None, None


def _bn_stats(x, norm, stats):
    """Statistics table for a BatchNorm2d layer (network/blocks.py:19-26, norm_type='batch'): training = the
    per-sample sums pooled over the batch (+ the running-estimate update of torch.nn.BatchNorm2d), eval = a table
    built from the running estimates.  Returns (table, mode)."""
    n, h, w, c = x.shape
    with torch.no_grad():
        if norm.training or not norm.track_running_stats:
            pooled = ops.bn_pool(stats if stats is not None else ops.in_stats(x))
            if norm.training and norm.track_running_stats:
                norm.num_batches_tracked.add_(1)
                if norm.momentum is None:
                    raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative average) is not on the path")
                ops.bn_running_update(pooled, h * w, norm.running_mean, norm.running_var, norm.momentum)
            return pooled, BN_TRAIN
        return ops.bn_eval_stats(norm.running_mean, norm.running_var, n, h * w, c), BN_EVAL


# Parity instrumentation (tests/parity_layers.py): when a list is installed here, every fused norm + activation call
# appends (norm module, activation output).  The sign of that output is the LeakyReLU / ReLU mask the kernels used;
# the per-layer parity protocol forces the oracle onto the same masks.
ACT_TAPS = [None]


def in_act(xa, norm_a, xb=None, norm_b=None, res=None, act=ACT_LRELU, c_params=None, stats_a=None, stats_b=None):
    batch = BN_OFF
    if getattr(norm_a, "smsut_batch_norm", False):
        stats_a, batch = _bn_stats(xa, norm_a, stats_a)
        if xb is not None:
            stats_b, _ = _bn_stats(xb, norm_b, stats_b)
    out = INActFn.apply(xa, norm_a.weight, norm_a.bias, xb, norm_b.weight if norm_b is not None else None,
                        norm_b.bias if norm_b is not None else None, res, act, c_params, stats_a, stats_b, batch)
    if ACT_TAPS[0] is not None and act != ACT_NONE:
        ACT_TAPS[0].append((norm_a, out))
    return out


# ----------------------------------------------------------------------------------------------
# pooling / resampling
# ----------------------------------------------------------------------------------------------
class MaxPoolSkipFn(Function):
    """nn.MaxPool2d(2, 2) that also hands back its input as the skip tensor, so the pooled path and the skip
    path deliver their gradients to ONE backward (fused route + add)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        ctx.set_materialize_grads(False)
        return ops.maxpool2_fwd(x), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        (x,) = ctx.saved_tensors
        if dy is None:
            return dskip
        return ops.maxpool2_bwd(x, _c(dy), add=_c(dskip) if dskip is not None else None)


class MaxPoolFn(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.maxpool2_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool2_bwd(x, _c(dy))


class AvgPoolBwdFn(Function):
    @staticmethod
    def forward(ctx, dy):
        return ops.avgpool2_bwd(dy)

    @staticmethod
    def backward(ctx, u):
        return ops.avgpool2_fwd(_c(u))


class AvgPoolFn(Function):
    """F.avg_pool2d(x, 2) (network/blocks.py:101-112)"""

    @staticmethod
    def forward(ctx, x):
        return ops.avgpool2_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        if torch.is_grad_enabled():
            return AvgPoolBwdFn.apply(dy)
        return ops.avgpool2_bwd(dy)


class BilinearFn(Function):
    """nn.Upsample(x2, bilinear, align_corners=False) (network/blocks.py:44)"""

    @staticmethod
    def forward(ctx, x):
        return ops.bilinear2_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.bilinear2_bwd(_c(dy))


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
_dice_allreduce = [None]  # callable(acc[:3c]) summing the Dice statistics over data-parallel ranks, or None
_world = [1]


def set_data_parallel(allreduce_fn, world_size):
    """Data-parallel hook: batch-Dice is a non-linear function of batch-wide sums, so tp/fp/fn are summed over
    ranks inside the loss (the reference's DataParallel computes the loss on the gathered batch)."""
    _dice_allreduce[0] = allreduce_fn
    _world[0] = world_size


class DiceCEFn(Function):
    """DiceAndCrossEntropyLoss (misc/loss.py:8-63, batch dice) on fp32 NHWC logits (npix, C); targets are int64
    labels or the argmax of `label_logits` (consistency_loss, trainer/uganConsisTrainer.py:45-53)."""

    @staticmethod
    def forward(ctx, logits, labels, label_logits, w_ce, w_dc):
        npix, c = logits.shape
        acc = ops.zeros(3 * c + 1, logits.device)
        ops.dice_ce_fwd(logits, labels, label_logits, acc)
        if _dice_allreduce[0] is not None:
            _dice_allreduce[0](acc[:3 * c])
        loss = ops.dice_ce_finish(acc, npix, c, w_dc, w_ce)
        ctx.save_for_backward(logits, labels, label_logits, acc)
        ctx.w = (w_ce, w_dc)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        logits, labels, label_logits, acc = ctx.saved_tensors
        w_ce, w_dc = ctx.w
        d = ops.dice_ce_bwd(logits, labels, label_logits, acc, _c(g).view(1), 1.0, logits.shape[0], w_dc * _world[0], w_ce)
        return d, None, None, None, None


class SoftmaxMSEFn(Function):
    """mean((softmax(zs) - softmax(zt))**2) with gradient to zs (trainer/meanTeacherTrainer.py:124-130)"""

    @staticmethod
    def forward(ctx, zs, zt):
        out = ops.zeros(1, zs.device)
        ops.softmax_mse_fwd(zs, zt, out)
        ctx.save_for_backward(zs, zt)
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        zs, zt = ctx.saved_tensors
        return ops.softmax_mse_bwd(zs, zt, _c(g).view(1)), None


class HeadsSplitFn(Function):
    """The three 5-class heads of coraNet's 13-channel output (trainer/coraNetTrainer.py:279-297): (npix, 1 + H*L) ->
    (H, npix, 1 + L) with the background channel shared; backward sums its H gradients."""

    @staticmethod
    def forward(ctx, z, nlab, nheads):
        ctx.dims = (nlab, nheads)
        return ops.heads_split_fwd(z, nlab, nheads)

    @staticmethod
    def backward(ctx, d):
        nlab, nheads = ctx.dims
        return ops.heads_split_bwd(_c(d), nlab, nheads), None, None


class WeightedCEFn(Function):
    """nn.CrossEntropyLoss(weight=cw) on (npix, C) logits: 'mean' form sum w[y] nll / sum w[y] (mask None), or the
    masked form `(CE(reduction='none') * mask).sum() / (mask.sum() + 1e-16)` (trainer/coraNetTrainer.py:44-58,301-303)"""

    @staticmethod
    def forward(ctx, z, y, cw, mask):
        acc = ops.zeros(3, z.device)
        ops.wce_fwd(z, y, cw, mask, acc)
        ctx.save_for_backward(z, y, cw, mask, acc)
        den = acc[2] + 1e-16 if mask is not None else acc[1]
        return (acc[0] / den).view(())

    @staticmethod
    def backward(ctx, g):
        z, y, cw, mask, acc = ctx.saved_tensors
        return ops.wce_bwd(z, y, cw, mask, acc, _c(g).view(1), mask is not None), None, None, None


class SoftmaxMSEMaskedFn(Function):
    """`(softmax_mse_loss(zs, zt) * m).sum() / (m.sum() + 1e-16)` with m = 1 - mask (invert) or mask, broadcast over the
    classes (trainer/coraNetTrainer.py:137-149,331-337); gradient to zs only."""

    @staticmethod
    def forward(ctx, zs, zt, mask, invert):
        acc = ops.zeros(2, zs.device)
        ops.softmax_mse_masked_fwd(zs, zt, mask, invert, acc)
        ctx.save_for_backward(zs, zt, mask, acc)
        ctx.invert = invert
        return (acc[0] / (acc[1] + 1e-16)).view(())

    @staticmethod
    def backward(ctx, g):
        zs, zt, mask, acc = ctx.saved_tensors
        return ops.softmax_mse_masked_bwd(zs, zt, mask, ctx.invert, acc, _c(g).view(1)), None, None, None


class L1MeanFn(Function):
    """mean |a - b| with gradient to a (g_loss_rec, trainer/uganConsisTrainer.py:162)"""

    @staticmethod
    def forward(ctx, a, b):
        out = ops.zeros(1, a.device)
        ops.l1_fwd(a, b, out, 1.0 / a.numel())
        ctx.save_for_backward(a, b)
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return ops.l1_bwd(a, b, _c(g).view(1), 1.0 / a.numel()), None


class MeanFn(Function):
    """scale * mean(x) (adversarial terms, trainer/uganConsisTrainer.py:130,136,154)"""

    @staticmethod
    def forward(ctx, x, scale):
        out = ops.zeros(1, x.device)
        ops.sum_f32(x, out, scale / x.numel())
        ctx.meta = (tuple(x.shape), scale / x.numel())
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        shape, s = ctx.meta
        return ops.fill_scaled(shape, _c(g).view(1), s, g.device), None


class CERowsFn(Function):
    """F.cross_entropy on (rows, C<=8) logits (modality classification, uganConsisTrainer.py:131,155)"""

    @staticmethod
    def forward(ctx, logits, target):
        out = ops.zeros(1, logits.device)
        ops.ce_rows_fwd(logits, target, out, 1.0)
        ctx.save_for_backward(logits, target)
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        logits, target = ctx.saved_tensors
        return ops.ce_rows_bwd(logits, target, _c(g).view(1), 1.0), None


class GradPenaltyFn(Function):
    """mean_b (||g_b||_2 - 1)^2 (trainer/uganShp0Trainer.py:131-134); g is the (differentiable) dD/dx_hat."""

    @staticmethod
    def forward(ctx, g):
        out = ops.zeros(1, g.device)
        norm = ops.gp_fwd(g, out, 1.0)
        ctx.save_for_backward(g, norm)
        return out.view(())

    @staticmethod
    def backward(ctx, gs):
        g, norm = ctx.saved_tensors
        return ops.gp_bwd(g, norm, _c(gs).view(1), 1.0)


class PatchSampleFn(Function):
    """PatchSampleF.forward for one feature map (network/ugan.py:316-334): gather `ids` rows of the NHWC
    bottleneck, Linear-ReLU-Linear on the tensor cores, L2 normalise (network/networks.py:241-242)."""

    @staticmethod
    def forward(ctx, feat, ids, pw1, w1, b1, pw2, w2, b2):
        rows = ops.gather_rows(feat, ids)                      # (R, C) bf16
        r, c = rows.shape
        tw = 128
        while r % tw:
            tw //= 2
        x4 = rows.view(1, r // tw, tw, c)                      # GEMM rows laid out as (1, r/tw, tw) "pixels"
        h = ops.conv_fprop([x4], pw1, bias=b1, act=ACT_RELU)   # bf16
        y = ops.conv_fprop([h], pw2, bias=b2, out_f32=True)    # fp32
        q, norm = ops.l2norm_fwd(y.view(r, -1))
        ctx.pw = (pw1, pw2)
        ctx.biases = (b1, b2)
        ctx.fshape = tuple(feat.shape)
        ctx.save_for_backward(ids, x4, h, q, norm, w1, w2)
        return q

    @staticmethod
    def backward(ctx, dq):
        ids, x4, h, q, norm, w1, w2 = ctx.saved_tensors
        pw1, pw2 = ctx.pw
        r = q.shape[0]
        dy = ops.l2norm_bwd(_c(dq), q, norm).view(1, x4.shape[1], x4.shape[2], -1)      # bf16
        # parameter gradients go straight into the flat .grad views inside accumulate_param_grads (autograd gets None):
        # an AccumulateGrad node runs on the stream its parameter was FIRST used on -- the main stream, for netF -- and
        # would make that stream wait for this backward (it held the whole discriminator phase back by 1.7 ms)
        b1, b2 = ctx.biases
        tw1, tb1, tw2, tb2 = _target(w1), _target(b1), _target(w2), _target(b2)
        dw2 = ops.conv_wgrad([h], dy, pw2, out=tw2.view(pw2.weight.shape) if tw2 is not None else None)
        db2 = ops.colsum(dy.view(r, -1), out=tb2)
        dh = ops.conv_dgrad(dy, pw2)[0]
        dh = ops.act_bwd(dh, h, act=ACT_RELU)
        dw1 = ops.conv_wgrad([x4], dh, pw1, out=tw1.view(pw1.weight.shape) if tw1 is not None else None)
        db1 = ops.colsum(dh.view(r, -1), out=tb1)
        dfeat = None
        if ctx.needs_input_grad[0]:
            drows = ops.conv_dgrad(dh, pw1)[0].view(r, -1)
            dfeat = torch.zeros(ctx.fshape, dtype=BF16, device=dq.device)
            ops.scatter_rows_add(drows, ids, dfeat)
        return (dfeat, None, None, _ret(dw1.view_as(w1), tw1), _ret(db1, tb1), None, _ret(dw2.view_as(w2), tw2),
                _ret(db2, tb2))


class PatchNCEFn(Function):
    """PatchNCELoss.forward (network/patchnce.py:13-51): per-row loss; feat_k is detached."""

    @staticmethod
    def forward(ctx, q, k, groups):
        rows = q.shape[0]
        out = ops.zeros(1, q.device)
        loss_rows = ops.patchnce_fwd(q, k, groups, rows // groups, 1.0 / 0.07, out, 1.0)
        ctx.groups = groups
        ctx.save_for_backward(q, k)
        return loss_rows

    @staticmethod
    def backward(ctx, g_rows):
        q, k = ctx.saved_tensors
        rows = q.shape[0]
        ones = torch.ones(1, dtype=F32, device=q.device)
        # d(loss_rows[r])/dq[r] * g_rows[r]: the kernel bakes 1/rows in, undo it and apply the per-row cotangent
        dq = ops.patchnce_bwd(q, k, ctx.groups, rows // ctx.groups, 1.0 / 0.07, ones, float(rows))
        return dq * _c(g_rows).view(rows, 1), None, None
