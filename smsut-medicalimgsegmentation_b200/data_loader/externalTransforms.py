# -*- coding: utf-8 -*-
"""Counterpart of the reference's data_loader/externalTransforms.py for the GPU input pipeline (SURVEY.md section 8f
N2).  The reference's classes transform ONE PIL slice on the host; here the same classes only DRAW the random
parameters of their transform (`get_params`, same distributions as the reference / torchvision) and the pixels are
moved by ONE CUDA launch per batch (csrc/augment.cu, smsut_augment_batch) on the u8 dataset resident in HBM.

    JointRotate(degrees)                  externalTransforms.py:58-68   angle ~ U(-degrees, degrees)
    JointElasticDeform(sigmas, points, p) externalTransforms.py:71-94   sigma ~ U(sigmas), applied with probability p,
                                                                        control displacements ~ N(0, sigma) on a
                                                                        points x points grid (elasticdeform)
    JointRandomResizedCrop(size, scale, ratio)  externalTransforms.py:46-55   torchvision RandomResizedCrop.get_params
    RandomGammaCorrection(gammas, p)      externalTransforms.py:22-43
    JointCompose                          externalTransforms.py:97-104
`pack_params` lays a slice's draws out as the SMSUT_AUG_PARAM_FLOATS floats the kernel reads.
"""
import math
import random

import numpy as np

PARAM_FLOATS = 66      # == SMSUT_AUG_PARAM_FLOATS (include/smsut_b200.h)
MAX_POINTS = 5


class MaskToTensor(object):
    """label image -> int64 tensor (externalTransforms.py:12-20).  The GPU pipeline emits int64 labels itself; this is
    for host-side callers of the reference's transform."""

    def __call__(self, x):
        import torch
        return torch.from_numpy(np.array(x)).long()

    def __repr__(self):
        return self.__class__.__name__ + '()\n'


class JointRotate(object):
    def __init__(self, degrees, resample=False, expand=False, center=None):
        if expand or center is not None:
            raise NotImplementedError("the path rotates about the image centre without expanding the canvas")
        self.degrees = (-degrees, degrees) if not isinstance(degrees, (tuple, list)) else tuple(degrees)

    @staticmethod
    def get_params(degrees):
        return random.uniform(degrees[0], degrees[1])

    def draw(self, h, w, rec):
        angle = self.get_params(self.degrees)
        rec['rotate'] = inverse_rotation(angle, h, w)

    def __repr__(self):
        return self.__class__.__name__ + '(degrees={0})'.format(self.degrees)


def inverse_rotation(angle_deg, h, w):
    """Output-to-source matrix of PIL's Image.rotate(angle) about the image centre (counter-clockwise positive, y
    down): source = R(-angle) applied about the centre, in pixel-CENTRE coordinates (x + 0.5, y + 0.5)."""
    a = -math.radians(angle_deg)
    cx, cy = w * 0.5, h * 0.5
    c, s = math.cos(a), math.sin(a)
    # PIL: matrix = [cos, sin, 0, -sin, cos, 0] for angle -> maps output (x, y) to input
    m = [c, s, 0.0, -s, c, 0.0]
    m[2] = cx - (m[0] * cx + m[1] * cy)
    m[5] = cy - (m[3] * cx + m[4] * cy)
    return m


class JointElasticDeform(object):
    def __init__(self, sigmas, points, p=0.5):
        if points > MAX_POINTS:
            raise NotImplementedError(f"at most {MAX_POINTS} control points per axis")
        self.sigmas, self.points, self.p = sigmas, points, p

    @staticmethod
    def get_params(sigmas):
        return random.uniform(sigmas[0], sigmas[1])

    def draw(self, h, w, rec):
        s = self.get_params(self.sigmas)
        if random.random() < self.p:
            disp = np.random.normal(0.0, s, size=(2, self.points, self.points))      # (dy, dx) per control point
            rec['elastic'] = (self.points, bspline_coefficients(disp))

    def __repr__(self):
        return self.__class__.__name__ + '(sigma={0}, points={1}, p={2})'.format(self.sigmas, self.points, self.p)


def bspline_coefficients(values):
    """Cubic B-spline coefficients c (mirror boundary) such that the spline INTERPOLATES `values` at the control
    points, along the last two axes (the prefilter of scipy.ndimage.spline_filter(order=3, mode='mirror'), solved
    directly: the grids are 3..5 points wide).  values: (..., n, n)."""
    v = np.asarray(values, dtype=np.float64)
    n = v.shape[-1]
    if n == 1:
        return v.astype(np.float32)
    A = np.zeros((n, n))
    for i in range(n):
        for k, wgt in ((i - 1, 1.0 / 6.0), (i, 4.0 / 6.0), (i + 1, 1.0 / 6.0)):
            period = 2 * (n - 1)
            k = k % period
            k = k if k < n else period - k
            A[i, k] += wgt
    Ainv = np.linalg.inv(A)
    c = np.einsum('ij,...jk->...ik', Ainv, v)          # along rows
    c = np.einsum('ij,...kj->...ki', Ainv, c)          # along columns
    return c.astype(np.float32)


class JointRandomResizedCrop(object):
    def __init__(self, size, scale=(0.6, 1.0), ratio=(3. / 4., 4. / 3.), interpolation=None):
        self.size = (size, size) if isinstance(size, int) else tuple(size)
        self.scale, self.ratio = scale, ratio

    @staticmethod
    def get_params(h, w, scale, ratio):
        """torchvision.transforms.RandomResizedCrop.get_params: (top, left, height, width)"""
        area = h * w
        log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
        for _ in range(10):
            target_area = area * random.uniform(scale[0], scale[1])
            aspect = math.exp(random.uniform(log_ratio[0], log_ratio[1]))
            cw = int(round(math.sqrt(target_area * aspect)))
            ch = int(round(math.sqrt(target_area / aspect)))
            if 0 < cw <= w and 0 < ch <= h:
                return random.randint(0, h - ch), random.randint(0, w - cw), ch, cw
        in_ratio = float(w) / float(h)
        if in_ratio < min(ratio):
            cw, ch = w, int(round(w / min(ratio)))
        elif in_ratio > max(ratio):
            ch, cw = h, int(round(h * max(ratio)))
        else:
            cw, ch = w, h
        return (h - ch) // 2, (w - cw) // 2, ch, cw

    def draw(self, h, w, rec):
        if self.size != (h, w):
            raise NotImplementedError("the crop is resized back to the slice size (cfg.data_aug['resizeCrop_size'] == input_size)")
        rec['crop'] = self.get_params(h, w, self.scale, self.ratio)


class RandomGammaCorrection(object):
    def __init__(self, gammas, p=0.5):
        if len(gammas) != 2:
            raise ValueError("Argument gammas must be a sequence of len 2.")
        self.gammas, self.p = gammas, p

    @staticmethod
    def get_params(gammas):
        return random.uniform(gammas[0], gammas[1])

    def draw(self, h, w, rec):
        gamma = self.get_params(self.gammas)
        if random.random() < self.p:
            rec['gamma'] = gamma

    def __repr__(self):
        return self.__class__.__name__ + '(gammas={0}, p={1})'.format(self.gammas, self.p)


class JointCompose(object):
    """The joint transforms of one loader, in the reference's order (rotate, elastic, resized crop) followed by the
    image-only ones; `draw(h, w)` returns one slice's parameter record."""

    def __init__(self, tfsm):
        order = {JointRotate: 0, JointElasticDeform: 1, JointRandomResizedCrop: 2, RandomGammaCorrection: 3}
        for t in tfsm:
            if type(t) not in order:
                raise NotImplementedError(type(t))
        if [order[type(t)] for t in tfsm] != sorted(order[type(t)] for t in tfsm):
            raise NotImplementedError("the fused kernel applies rotate -> elastic -> resized crop -> gamma (the order of "
                                      "data_loader/baseLoader.py:93-100)")
        self.transforms = list(tfsm)

    def draw(self, h, w):
        rec = {}
        for t in self.transforms:
            t.draw(h, w, rec)
        return rec


def pack_params(rec, out):
    """one slice's record -> the PARAM_FLOATS floats of csrc/augment.cu::AugParams (out: float32 view of that length)"""
    out[:] = 0.0
    if 'rotate' in rec:
        out[0] = 1.0
        out[1:7] = rec['rotate']
    if 'elastic' in rec:
        points, coef = rec['elastic']
        out[7] = 1.0
        out[15] = points
        out[16:16 + 2 * points * points] = np.asarray(coef, dtype=np.float32).reshape(-1)
    if 'crop' in rec:
        out[8] = 1.0
        out[9:13] = rec['crop']
    if 'gamma' in rec:
        out[13] = 1.0
        out[14] = rec['gamma']
    return out
