"""CPU check of the per-layer parity protocol itself (tests/parity_layers.py): with the kernel layer swapped for its
fp32 PyTorch test double, every layer of the three networks must match the oracle to rounding (1e-4) forward and
backward, teacher-forced per layer and end to end with forced selections.  The GPU twin
(tests/test_parity_layers_gpu.py) runs the same protocol on the sm_100a kernels with the north-star tolerance."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cpu_ops_mock  # noqa: E402
import parity_layers as PL  # noqa: E402
from oracle import smsut_oracle as O  # noqa: E402

os.environ.setdefault("SMSUT_ALLOW_CPU_TEST_DOUBLE", "1")
F32 = torch.float32


def _check(results, tol):
    w = PL.summarize(results)
    assert w["fwd"][0] < tol and w["dx"][0] < tol and w["params"][0] < tol, w
    return w


def test_recording_style_is_transparent():
    sd = O.make_weights(O.unet_shapes(), 1)
    x, y = O.synthetic_batch(2, 32, 3)
    a = O.unet_forward(sd, x)
    st = O.RecordingStyle()
    b = O.unet_forward(sd, x, style=st)
    assert torch.equal(a, b)
    assert "encoder.layer3.act2" in st.masks and "encoder.pool2" in st.pools and "decoder.fc.out" in st.taps
    # forcing the run's own selections reproduces it
    c = O.unet_forward(sd, x, style=O.RecordingStyle(forced_masks=dict(st.masks), forced_pool=dict(st.pools)))
    assert torch.allclose(a, c, atol=1e-6)


def test_unet_layers_protocol(pkg):
    from smsut_b200 import functional as Fn
    from smsut_b200.network.unet import UNet
    sd = O.make_weights(O.unet_shapes(), 1)
    with cpu_ops_mock.installed(exact=True):
        net = UNet(1, 5, 16, 'instance', 'lrelu')
        net.load_state_dict(sd)
        x, y = O.synthetic_batch(2, 32, 3)
        st = O.RecordingStyle()
        leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        O.dice_ce_loss(O.unet_forward(leaf, x, style=st), y).backward()
        res = {L.name: PL.run_layer(Fn, L, sd, act_dtype=F32) for L in PL.unet_layers(Fn, net, sd, st)}
        assert len(res) == 15
        _check(res, 1e-4)
        # experiment 2: free drop-in forward, its selections forced onto the oracle
        Fn.ACT_TAPS[0] = []
        out = net(x)
        taps, Fn.ACT_TAPS[0] = Fn.ACT_TAPS[0], None
        masks, pools = PL.collect_selections(taps, PL.selection_keys(net, "unet"),
                                             {f"encoder.layer{i}.act2": f"encoder.pool{i}" for i in range(1, 5)})
        assert len(masks) == 19 and len(pools) == 4
        fs = PL.ForcedStyle(forced_masks=masks, forced_pool=pools)
        leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        ref = O.unet_forward(leaf, x, style=fs)
        net.zero_grad()
        O.dice_ce_loss(ref, y).backward()
        from smsut_b200.misc.loss import DiceAndCrossEntropyLoss
        DiceAndCrossEntropyLoss(0.5, 0.5, batch_dice=True)(out, y).backward()
        assert PL.rel(out, ref) < 1e-4
        worst = max(PL.rel(p.grad, leaf[k].grad) for k, p in net.named_parameters())
        assert worst < 1e-3, worst


def test_ugan_and_discriminator_layers_protocol(pkg):
    from smsut_b200 import functional as Fn
    from smsut_b200.network.ugan import Discriminator, UGANnce
    size = 64
    with cpu_ops_mock.installed(exact=True):
        sd = O.make_weights(O.ugan_shapes(), 4)
        net = UGANnce(1, 5, 4, 16)
        net.load_state_dict(sd)
        x, _ = O.synthetic_batch(2, size, 4)
        m = torch.tensor([[1., 0, -1, 0], [0, 1., -1, 0]])
        ids = [torch.randperm(16, generator=torch.Generator().manual_seed(0))[:16]]
        st = O.RecordingStyle()
        leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        seg, tsl, feats, _ = O.ugannce_forward(leaf, x, m, sample_ids=ids, style=st)
        feats[0].retain_grad()
        w = torch.randn(seg.shape, generator=torch.Generator().manual_seed(1))
        ((seg * w).mean() + tsl.mean() + (feats[0] ** 3).sum()).backward()
        layers = PL.ugan_layers(Fn, net, sd, st, m, ids, feats[0].grad)
        res = {L.name: PL.run_layer(Fn, L, sd, act_dtype=F32) for L in layers}
        assert len(res) == 2 * 5 + 2 + 2 * 9 + 1
        _check(res, 1e-4)

        dsd = O.make_weights(O.disc_shapes(size), 5)
        D = Discriminator(size, 4, 16, max_width=256)
        D.load_state_dict(dsd)
        st = O.RecordingStyle()
        leaf = {k: v.clone().requires_grad_(True) for k, v in dsd.items()}
        src, cls = O.discriminator_forward(leaf, x, style=st)
        src.retain_grad(); cls.retain_grad()
        (src.mean() + cls.pow(2).mean()).backward()
        res = {L.name: PL.run_layer(Fn, L, dsd, act_dtype=F32) for L in PL.disc_layers(Fn, D, dsd, st, src.grad, cls.grad)}
        assert len(res) == 1 + 3 + 2
        _check(res, 1e-4)
