"""The two symbols of the reference's network/networks.py that the hot path reaches (network/ugan.py:266,274):
`Normalize` (networks.py:234-243) and `init_weights` (networks.py:163-195).  The rest of that vendored CUT file
(ResnetGenerator, NLayerDiscriminator, ...) is dead code for this path and is not reproduced."""
import torch
import torch.nn as nn
from torch.nn import init

from .. import ops


class Normalize(nn.Module):
    """x / (||x||_p + 1e-7) per row; p = 2 runs on the l2norm kernel."""

    def __init__(self, power=2):
        super(Normalize, self).__init__()
        self.power = power

    def forward(self, x):
        if self.power != 2 or x.dim() != 2:
            raise NotImplementedError("the path only uses Normalize(2) on (rows, C) features")
        return _L2NormFn.apply(x.float().contiguous())


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, norm = ops.l2norm_fwd(x)
        ctx.save_for_backward(y, norm)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, norm = ctx.saved_tensors
        return ops.l2norm_bwd(dy.contiguous(), y, norm).float()


def init_weights(net, init_type='normal', init_gain=0.02, debug=False):
    """N(0, init_gain) for Conv / Linear weights, zero biases (the reference's 'normal' branch)."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, 'weight') and (classname.find('Conv') != -1 or classname.find('Linear') != -1):
            if init_type == 'normal':
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == 'xavier':
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == 'kaiming':
                init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            if hasattr(m, 'bias') and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find('BatchNorm2d') != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    net.apply(init_func)
