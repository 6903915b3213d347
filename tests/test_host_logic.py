"""Host-logic tests that run without a GPU: the module tree, the autograd wiring (including the differentiable
double backward of the discriminator for WGAN-GP), the fused-optimizer plumbing and the full UGANConsisTrainer
step are executed with the kernel layer replaced by the PyTorch test double of tests/cpu_ops_mock.py in `exact`
(fp32) mode, and must reproduce the oracle to rounding error.  (The kernels themselves are checked on the GPU in
test_kernels_gpu.py / test_modules_gpu.py.)"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cpu_ops_mock  # noqa: E402
from oracle import smsut_oracle as O  # noqa: E402

os.environ["SMSUT_ALLOW_CPU_TEST_DOUBLE"] = "1"


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.fixture()
def exact(pkg):
    with cpu_ops_mock.installed(exact=True) as ops:
        yield ops


def test_state_dict_layouts_match_reference_contract(pkg):
    from smsut_b200.network.ugan import Discriminator, UGANnce
    from smsut_b200.network.unet import UNet
    for net, shapes in ((UNet(1, 5, 16, 'instance', 'lrelu'), O.unet_shapes()),
                        (UGANnce(1, 5, 4, 16), O.ugan_shapes()),
                        (Discriminator(256, 4, 16, max_width=256), O.disc_shapes())):
        sd = net.state_dict()
        assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v) for k, v in shapes.items()}
    assert len(O.ugan_shapes()) == 175 and len(O.disc_shapes()) == 46 and len(O.unet_shapes()) == 89


def test_unet_forward_backward_matches_oracle(exact):
    from smsut_b200.misc.loss import DiceAndCrossEntropyLoss
    from smsut_b200.network.unet import UNet
    net = UNet(1, 5, 16, 'instance', 'lrelu')
    sd = O.make_weights(O.unet_shapes(), 1)
    net.load_state_dict(sd)
    x, y = O.synthetic_batch(2, 64, 3)
    out = net(x)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.unet_forward(ref_sd, x)
    assert out.shape == ref.shape and rel(out, ref) < 1e-5
    loss = DiceAndCrossEntropyLoss(0.5, 0.5, batch_dice=True)(out, y)
    lref = O.dice_ce_loss(ref, y)
    assert abs(loss.item() - lref.item()) < 1e-5
    loss.backward()
    lref.backward()
    for k, p in net.named_parameters():
        assert rel(p.grad, ref_sd[k].grad) < 1e-4, k


def test_ugannce_forward_backward_matches_oracle(exact):
    from smsut_b200.network.ugan import UGANnce
    net = UGANnce(1, 5, 4, 16)
    sd = O.make_weights(O.ugan_shapes(), 4)  # a seed without |pre-activation| ~ 1e-6 (LeakyReLU mask flips)
    net.load_state_dict(sd)
    x, _ = O.synthetic_batch(2, 64, 4)
    m = torch.tensor([[1., 0, -1, 0], [0, 1., -1, 0]])
    ids = [torch.randperm(16, generator=torch.Generator().manual_seed(0))]
    seg, tsl, feats, _ = net(x, m, sample_ids=ids)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rseg, rtsl, rfeats, _ = O.ugannce_forward(ref_sd, x, m, sample_ids=ids)
    assert rel(seg, rseg) < 1e-4 and rel(tsl, rtsl) < 1e-4 and rel(feats[0], rfeats[0]) < 1e-4
    # val_phase arity
    assert len(net(x, val_phase=True)) == 2
    w = torch.randn(rseg.shape, generator=torch.Generator().manual_seed(1))
    (seg * w).sum().add(tsl.sum()).add((feats[0] ** 3).sum()).backward()
    (rseg * w).sum().add(rtsl.sum()).add((rfeats[0] ** 3).sum()).backward()
    for k, p in net.named_parameters():
        assert rel(p.grad, ref_sd[k].grad) < 1e-3, k


def test_discriminator_gradient_penalty_double_backward(exact):
    from smsut_b200.network.ugan import Discriminator
    from smsut_b200.trainer.uganShp0Trainer import UGANShp0Trainer
    D = Discriminator(64, 4, 16, max_width=256)
    sd = O.make_weights(O.disc_shapes(64), 5)
    D.load_state_dict(sd)
    x, _ = O.synthetic_batch(3, 64, 6)
    # seeded noise: about one draw in twelve puts a pre-activation within rounding of zero, where the test double and the
    # oracle (same fp32 arithmetic, different summation order) pick different LeakyReLU sides (error 1e-3 instead of 1e-6)
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(0))
    x_hat = (x + 0.1 * noise).requires_grad_(True)
    out_src, out_cls = D(x_hat)
    gp = UGANShp0Trainer.gradient_penalty(None, out_src, x_hat)
    (gp * 10 + out_src.mean() + out_cls.pow(2).mean()).backward()

    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x_hat.detach().clone().requires_grad_(True)
    rsrc, rcls = O.discriminator_forward(ref_sd, xr)
    assert rel(out_src, rsrc) < 1e-5 and rel(out_cls, rcls) < 1e-5
    rgp = O.gradient_penalty(rsrc, xr)
    assert abs(gp.item() - rgp.item()) < 1e-4 * max(1.0, abs(rgp.item()))
    (rgp * 10 + rsrc.mean() + rcls.pow(2).mean()).backward()
    for k, p in D.named_parameters():
        assert rel(p.grad, ref_sd[k].grad) < 5e-4, k


def _consis_trainer(size):
    from types import SimpleNamespace
    from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer
    tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=size))
    G = O.make_weights(O.ugan_shapes(), 7)
    D = O.make_weights(O.disc_shapes(size), 8)
    tr.net.load_state_dict(G)
    tr.D.load_state_dict(D)
    return tr, G, D


@pytest.mark.parametrize("use_semi", [False, True])
def test_full_ugan_consis_step_matches_oracle(exact, use_semi):
    """One teacher-forced iteration (the free-running GAN trajectory is chaotic even
    fp32-vs-fp32 -- Adam's first update is lr*sign(g) -- so every quantity downstream of D's Adam step gets a
    looser bound than the D-phase quantities)."""
    from smsut_b200.trainer.uganConsisTrainer import LOSS_KEYS
    size, bs = 64, 2
    tr, G, D = _consis_trainer(size)
    x1, y = O.synthetic_batch(bs, size, 11)
    x2, _ = O.synthetic_batch(bs, size, 12)
    mod1, mod2 = torch.full((bs,), 1), torch.full((bs,), 3)
    gen = torch.Generator().manual_seed(3)
    mj = 2
    alpha = torch.randn(2 * bs, generator=gen)
    ids = [torch.randperm(16, generator=gen)]
    batch = tr.prepare_batch(x1, y, mod1, x2, mod2, mj)
    got = tr.train_step(*batch, alpha, ids, 0.7, use_semi)
    xr, mr = torch.cat([x1, x2]), torch.cat([mod1, mod2])
    ref, d_grads = O.ugan_d_phase(G, D, {}, xr, mr, mj, alpha.view(-1, 1, 1, 1), ids, 1e-2)
    for k, p in tr.D.named_parameters():       # gradients of d_loss (incl. the double backward of the GP term)
        assert rel(p.grad, d_grads[k]) < 1e-3, k
        assert rel(p, D[k]) < 2e-2, k          # after Adam (sign-like first update: see docstring)
    # teacher-force the G phase: the oracle continues from the trainer's updated discriminator
    D = {k: v.detach().clone() for k, v in tr.D.state_dict().items()}
    g_ref, g_grads = O.ugan_g_phase(G, D, {}, xr, y, mr, mj, ids, 1e-2, 1000 if use_semi else 0, 0.7, nce_batch=8)
    ref.update(g_ref)
    for k, v in zip(LOSS_KEYS, got.tolist()):
        assert abs(v - ref[k]) < 2e-4 * max(1.0, abs(ref[k])), (k, v, ref[k])
    if use_semi:
        assert ref['G_semi'] > 0
    # G gradients chain D(G(x)) and G(G(x)): a 1e-6 perturbation of x_fake flips the LeakyReLU mask of the odd
    # near-zero pre-activation (each flip moves a gradient tensor by ~1%), so the bound is looser than for the
    # single-network tests above, which are exact
    for k, p in tr.net.named_parameters():
        assert rel(p.grad, g_grads[k]) < 8e-2, k
        assert rel(p, G[k]) < 8e-2, k          # after SGD


def test_unet_trainer_steps_match_oracle(exact):
    from types import SimpleNamespace
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=64))
    sd = O.make_weights(O.unet_shapes(), 21)
    tr.net.load_state_dict(sd)
    st = {}
    for it in range(3):
        x, y = O.synthetic_batch(2, 64, 30 + it)
        loss = tr.train_step(x, y)
        ref, _ = O.unet_step(sd, st, x, y, O.poly_lr(1e-2, max(it - 1, 0), 30000))
        assert abs(loss.item() - ref.item()) < 1e-5
        for k, p in tr.net.named_parameters():
            assert rel(p, sd[k]) < 1e-3, (it, k)


def test_bf16_double_tracks_oracle_within_tolerance(pkg):
    """same wiring with bf16 activations (what the kernels store): logits within the 2e-2 budget"""
    from smsut_b200.network.unet import UNet
    with cpu_ops_mock.installed(exact=False):
        net = UNet(1, 5, 16, 'instance', 'lrelu')
        sd = O.make_weights(O.unet_shapes(), 1)
        net.load_state_dict(sd)
        x, _ = O.synthetic_batch(2, 64, 3)
        assert rel(net(x), O.unet_forward(sd, x)) < 3e-2


def test_mean_teacher_steps_match_oracle(exact):
    from types import SimpleNamespace
    from smsut_b200.trainer.meanTeacherTrainer import MeanTeacherTrainer
    tr = MeanTeacherTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=64))
    tr.semi_from_iter = 1
    sd, ema = O.make_weights(O.unet_shapes(), 31), O.make_weights(O.unet_shapes(), 32)
    tr.net.load_state_dict(sd)
    tr.ema.load_state_dict(ema)
    st = {}
    for it in range(3):
        x1, y = O.synthetic_batch(2, 64, 40 + it)
        x2, _ = O.synthetic_batch(2, 64, 50 + it)
        x = torch.cat([x1, x2])
        noise = torch.clamp(torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(it)) * 0.01, -0.02, 0.02)
        got = tr.train_step(x, y, noise, 0.8).tolist()
        ref = O.mean_teacher_step(sd, ema, st, x, y, noise, O.poly_lr(1e-2, max(it - 1, 0), 30000), it, 0.8, warm=1)
        assert abs(got[0] - ref[0]) < 1e-4 and abs(got[1] - ref[1]) < 1e-5, (it, got, ref)
        for k, p in tr.net.named_parameters():
            assert rel(p, sd[k]) < 1e-3, (it, k)
        for k, p in tr.ema.named_parameters():
            assert rel(p, ema[k]) < 1e-3, (it, k)


def test_coranet_steps_match_oracle(exact, tmp_path, monkeypatch):
    """coraNetTrainer: a pre_epoch iteration, pred_unlabel, train_epoch iterations before / after the
    iter-1000 switch and the pre_best -> fit checkpoint hand-over, on the fp32 test double vs the oracle (which
    tests/test_oracle.py pins to the fixture made from the reference's own UNet and loss class)."""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.trainer.coraNetTrainer import PseudoLabelSet, coraNetTrainer
    monkeypatch.setattr(cfg, "expr_root", str(tmp_path))
    tr = coraNetTrainer('train', SimpleNamespace(fold=0, expr_name='cora', input_size=64, model_id=None))
    n_out = 3 * cfg.n_label + 1
    sd, ema = O.make_weights(O.unet_shapes(out_ch=n_out), 71), O.make_weights(O.unet_shapes(out_ch=n_out), 72)
    assert set(tr.net.state_dict()) == set(sd)
    tr.net.load_state_dict(sd)
    tr.ema.load_state_dict(ema)
    st = {}
    tr.iter = 200
    img1, msk = O.synthetic_batch(2, 64, 81)
    got = tr.pre_step(img1, msk).tolist()
    ref, _ = O.coranet_pre_step(sd, ema, st, img1, msk, 1e-2, 200)
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5), (got, ref)
    for k, p in tr.net.named_parameters():
        assert rel(p, sd[k]) < 1e-3, k
    for k, p in tr.ema.named_parameters():
        assert rel(p, ema[k]) < 1e-3, k
    # pseudo labels and certainty mask of a batch of unlabelled slices
    imgu, labu = O.synthetic_batch(2, 64, 83)
    new_loader, plab_dice = tr.pred_unlabel([(imgu, labu, torch.zeros(2, dtype=torch.long), None)])
    plab, mask = O.coranet_pred_unlabel(sd, imgu)
    assert isinstance(new_loader, PseudoLabelSet) and new_loader.num == 2 and 0.0 <= plab_dice <= 1.0
    assert (new_loader.plab != plab).float().mean() < 1e-3 and (new_loader.mask != mask).float().mean() < 1e-3
    fg_p, fg_l = plab > 0, labu > 0
    ref_dice = np.mean([2.0 * (fg_p[i] & fg_l[i]).sum().item() / max((fg_p[i].sum() + fg_l[i].sum()).item(), 1) for i in range(2)])
    assert abs(plab_dice - ref_dice) < 1e-6
    # train iterations: before (certain / uncertain terms are zeros) and after iter 1000
    for it in (300, 1500):
        tr.iter = it
        img1, msk = O.synthetic_batch(2, 64, 91 + it)
        got = tr.train_step(img1, msk, imgu, plab, mask, 0.3).tolist()
        ref, _ = O.coranet_train_step(sd, ema, st, img1, msk, imgu, plab, mask, 1e-2, it, 0.3)
        assert np.allclose(got, ref, rtol=2e-4, atol=1e-5), (it, got, ref)
        for k, p in tr.net.named_parameters():
            assert rel(p, sd[k]) < 2e-3, (it, k)
        for k, p in tr.ema.named_parameters():
            assert rel(p, ema[k]) < 2e-3, (it, k)
    # prefit writes pre_best / pre_ema_best, fit of a fresh trainer starts from them
    loaders = ([(img1, msk, torch.zeros(2, dtype=torch.long), None)], [(imgu, labu, torch.zeros(2, dtype=torch.long), None)],
               [(imgu, labu, torch.zeros(2, dtype=torch.long), ["0_1_0", "0_1_1"])])
    monkeypatch.setattr(cfg, "batch_size", 2)
    tr.prefit('inTurn', pre_epoch=1, iters_per_epoch=1, loaders=loaders)
    ck = os.path.join(str(tmp_path), 'cora', tr.model_idx, 'ckpt')
    assert {'pre_best.ckpt', 'pre_ema_best.ckpt', 'pre_last.ckpt', 'pre_ema_last.ckpt'} <= set(os.listdir(ck))
    tr2 = coraNetTrainer('train', SimpleNamespace(fold=0, expr_name='cora', input_size=64, model_id=tr.model_idx))
    tr2.fit('inTurn', max_epoch=1, iters_per_epoch=2, loaders=loaders)
    assert {'best.ckpt', 'last.ckpt'} <= set(os.listdir(os.path.join(str(tmp_path), 'cora', tr2.model_idx, 'ckpt')))
    assert tr2.iter == 2 and tr2.epoch == 1


def test_unet_batchnorm_relu_matches_oracle(exact):
    """get_norm('batch') / get_act('relu') (the UNet signature's defaults, network/unet.py:14): training passes with
    running-estimate updates, then eval; state_dict carries nn.BatchNorm2d's buffers."""
    from smsut_b200.misc.loss import DiceAndCrossEntropyLoss
    from smsut_b200.network.unet import UNet
    net = UNet(1, 5, 16)
    sd = O.add_bn_buffers(O.make_weights(O.unet_shapes(), 11))
    assert set(net.state_dict()) == set(sd)
    net.load_state_dict(sd)
    ref_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in sd.items()}
    crit = DiceAndCrossEntropyLoss(0.5, 0.5, batch_dice=True)
    net.train()
    for seed in (21, 22):
        x, y = O.synthetic_batch(2, 48, seed)
        net.zero_grad()
        for v in ref_sd.values():
            v.grad = None
        out = net(x)
        ref = O.unet_forward(ref_sd, x, style=O.Style("batch", "relu", training=True))
        assert rel(out, ref) < 1e-5
        crit(out, y).backward()
        O.dice_ce_loss(ref, y).backward()
        for k, p in net.named_parameters():
            assert rel(p.grad, ref_sd[k].grad) < 2e-4, k
    for k, v in net.state_dict().items():
        if "running" in k:
            assert rel(v, ref_sd[k]) < 1e-5, k
        if "num_batches" in k:
            assert int(v) == 2
    net.eval()
    x, _ = O.synthetic_batch(2, 48, 23)
    with torch.no_grad():
        out = net(x)
    ref = O.unet_forward(ref_sd, x, style=O.Style("batch", "relu", training=False))
    assert rel(out, ref) < 1e-5


def test_cross_pse_trainer_steps_match_oracle(exact):
    """crossPseTrainer: two networks, cross pseudo-labels through the fused loss's argmax"""
    from types import SimpleNamespace
    from smsut_b200.trainer.crossPseTrainer import crossPseTrainer
    tr = crossPseTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=64))
    sd1, sd2 = O.make_weights(O.unet_shapes(), 31), O.make_weights(O.unet_shapes(), 32)
    tr.net.load_state_dict(sd1)
    tr.net2.load_state_dict(sd2)
    st1, st2 = {}, {}
    for it in range(3):
        x1, y = O.synthetic_batch(2, 64, 41 + it)
        x2, _ = O.synthetic_batch(2, 64, 51 + it)
        x = torch.cat([x1, x2])
        got = tr.train_step(x, y, 0.05).tolist()
        ref, _, _ = O.cross_pse_step(sd1, sd2, st1, st2, x, y, O.poly_lr(1e-2, max(it - 1, 0), 30000), 0.05)
        for v, k in zip(got, ("seg1", "seg2", "semi1", "semi2")):
            # the cross losses see argmax pseudo-labels: after an update, rounding-level weight differences flip
            # the label of the odd near-tied pixel, which moves the loss by more than rounding error
            tol = 1e-4 if it == 0 or k.startswith("seg") else 2e-3
            assert abs(v - ref[k]) < tol * max(1.0, abs(ref[k])), (it, k, v, ref[k])
        for net, sd in ((tr.net, sd1), (tr.net2, sd2)):
            for k, p in net.named_parameters():
                assert rel(p, sd[k]) < 1e-3, (it, k)


@pytest.mark.parametrize("lambda_shp", [3.5, None])
def test_ugan_shape_trainer_step_matches_oracle(exact, lambda_shp):
    """UGANTrainer (shape loss) / the UGANShp0Trainer iteration on the `UGAN` generator, teacher-forced across D's Adam
    step like test_full_ugan_consis_step_matches_oracle."""
    from types import SimpleNamespace
    from smsut_b200.trainer.uganTrainer import UGANTrainer
    size = 64
    tr = UGANTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=size))
    shapes = {k: v for k, v in O.ugan_shapes().items() if not k.startswith("netF.")}
    G, D = O.make_weights(shapes, 61), O.make_weights(O.disc_shapes(size), 62)
    assert set(tr.net.state_dict()) == set(G)
    tr.net.load_state_dict(G)
    tr.D.load_state_dict(D)
    x, y = O.synthetic_batch(3, size, 63)
    modal_org, mj = torch.full((3,), 1), 3
    modal_trg = torch.full_like(modal_org, mj)
    vo, vt = tr.label2onehot(modal_org, 4), tr.label2onehot(modal_trg, 4)
    alpha = torch.randn(3, generator=torch.Generator().manual_seed(64))
    got = tr.shape_train_step(x, y, modal_org, modal_trg, vt - vo, vo - vt, alpha, lambda_shp).tolist()
    # the oracle's D phase from the same start ...
    D0 = {k: v.clone() for k, v in D.items()}
    G0 = {k: v.clone() for k, v in G.items()}
    ref, d_grads, _ = O.ugan_shape_step(G0, D0, {}, {}, x, y, modal_org, mj, alpha.view(-1, 1, 1, 1), 1e-2, lambda_shp)
    for k, p in tr.D.named_parameters():
        assert rel(p.grad, d_grads[k]) < 3e-3, k     # dominated by the double backward of a GP term of O(1e3)
    for v, k in list(zip(got, tr.SHP_LOSS_KEYS))[:4]:
        assert abs(v - ref[k]) < 2e-4 * max(1.0, abs(ref[k])), (k, v, ref[k])
    # ... and its G phase against the trainer's updated discriminator (teacher forcing)
    import unittest.mock as um
    Dt = {k: v.detach().clone() for k, v in tr.D.state_dict().items()}
    with um.patch.object(O, "adam_update", lambda *a, **k: None):
        ref2, _, g_grads = O.ugan_shape_step(G, Dt, {}, {}, x, y, modal_org, mj, alpha.view(-1, 1, 1, 1), 1e-2, lambda_shp)
    for v, k in list(zip(got, tr.SHP_LOSS_KEYS))[4:]:
        if k in ref2:
            assert abs(v - ref2[k]) < 2e-4 * max(1.0, abs(ref2[k])), (k, v, ref2[k])
    assert ("G_shp" in ref2) == (lambda_shp is not None)
    for k, p in tr.net.named_parameters():
        assert rel(p.grad, g_grads[k]) < 8e-2, k
        assert rel(p, G[k]) < 8e-2, k


def test_resume_state_round_trip(exact, tmp_path):
    """save_state / load_state (extension): a fresh trainer restored from the state file
    continues exactly like the original -- weights, EMA teacher, SGD momentum, LR-schedule position, counters."""
    from types import SimpleNamespace
    from smsut_b200.trainer.meanTeacherTrainer import MeanTeacherTrainer

    def batch(it):
        x1, y = O.synthetic_batch(2, 64, 40 + it)
        x2, _ = O.synthetic_batch(2, 64, 50 + it)
        noise = torch.clamp(torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(it)) * 0.01, -0.02, 0.02)
        return torch.cat([x1, x2]), y, noise

    def make():
        tr = MeanTeacherTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=64))
        tr.semi_from_iter = 1
        tr.expr_root = str(tmp_path)
        return tr

    a = make()
    a.net.load_state_dict(O.make_weights(O.unet_shapes(), 31))
    a.ema.load_state_dict(O.make_weights(O.unet_shapes(), 32))
    for it in range(2):
        a.train_step(*batch(it), 0.8)
    path = a.save_state()
    want = a.train_step(*batch(2), 0.8)
    b = make()                      # fresh random weights, zero momentum, schedule at 0
    b.load_state(path)
    assert b.iter == 2
    got = b.train_step(*batch(2), 0.8)
    assert torch.equal(got, want)
    for (k, p), (_, q) in zip(a.net.named_parameters(), b.net.named_parameters()):
        assert torch.equal(p, q), k
    for (k, p), (_, q) in zip(a.ema.named_parameters(), b.ema.named_parameters()):
        assert torch.equal(p, q), k
    assert torch.equal(a.optimizer.mom, b.optimizer.mom) and torch.equal(a.lr_sched.iter_state, b.lr_sched.iter_state)


def test_fit_validate_and_reference_format_checkpoints(exact, tmp_path, monkeypatch):
    """`-p train` path of BaseTrainer (trainer/baseTrainer.py:125-201): fit = train_epoch + validate_epoch + best/last
    checkpoints.  The checkpoints are the reference's format -- a plain state_dict of fp32 OIHW tensors without a
    `module.` prefix (trainer/uganShp0Trainer.py:94-107) -- and load back into a fresh trainer (`-p test`)."""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    monkeypatch.setattr(cfg, "batch_size", 2)
    args = SimpleNamespace(fold=0, expr_name="t", input_size=32)
    tr = UnetTrainer('train', args)
    tr.expr_root = str(tmp_path)
    tr.fit('inTurn', max_epoch=1, iters_per_epoch=2)
    assert tr.iter == 2 and tr.epoch == 1
    ck = torch.load(os.path.join(str(tmp_path), '000', 'ckpt', 'last.ckpt'))
    assert set(ck) == set(O.unet_shapes()) and all(v.dtype == torch.float32 and v.device.type == 'cpu' for v in ck.values())
    assert {k: tuple(v.shape) for k, v in ck.items()} == {k: tuple(v) for k, v in O.unet_shapes().items()}
    assert os.path.exists(os.path.join(str(tmp_path), '000', 'ckpt', 'best.ckpt'))
    te = UnetTrainer('test', args)
    te.expr_root = str(tmp_path)
    te.load_model('000', 'last')
    for (k, p), (_, q) in zip(tr.net.state_dict().items(), te.net.state_dict().items()):
        assert torch.equal(p.cpu(), q.cpu()), k
    # the oracle evaluates the checkpoint to the same logits as the restored network
    x, _ = O.synthetic_batch(2, 32, 5)
    with torch.no_grad():
        assert rel(te.net(x), O.unet_forward(ck, x)) < 1e-5


def test_validation_modality_organ_matrix_matches_oracle(exact, monkeypatch):
    """validate_epoch + validate_dice: the modality-organ Dice matrix of get_mo_matrix
    (misc/utils.py:180-203) from per-volume confusion counts, incl. a ragged last batch (padded to cfg.batch_size) and
    a volume that spans two batches."""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    monkeypatch.setattr(cfg, "batch_size", 4)
    tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=32))
    sd = O.make_weights(O.unet_shapes(), 5)
    tr.net.load_state_dict(sd)
    # modality 0: volumes a (4 slices) and b (6 slices: one full batch + a ragged batch of 2); modality 2: volume c (4)
    layout = [(0, 'a', 4), (0, 'b', 4), (0, 'b', 2), (2, 'c', 4)]
    batches, z0 = [], {}
    for bi, (m, pid, n) in enumerate(layout):
        img, lab = O.synthetic_batch(n, 32, 300 + bi)
        s = z0.get((m, pid), 0)
        names = [f"{m}_{pid}_{s + z}" for z in range(n)]
        z0[(m, pid)] = s + n
        batches.append((img, lab, torch.full((n,), m), names))
    dice = tr.validate_epoch(batches)
    dices, matrix = tr.validate_dice()
    prd, gt = {}, {}
    for img, lab, mdl, names in batches:
        pred = O.unet_forward(sd, img).argmax(1)
        key = '_'.join(names[0].split('_')[:2])
        prd[key] = torch.cat([prd[key], pred]) if key in prd else pred
        gt[key] = torch.cat([gt[key], lab]) if key in gt else lab
    ref = O.mo_matrix({k: v.numpy() for k, v in prd.items()}, {k: v.numpy() for k, v in gt.items()})
    assert matrix.shape == (5, 5) and abs(matrix - ref).max() < 1e-12
    assert abs(dices['dice'] - ref[-1, -1]) < 1e-12 and abs(dices['dice_2'] - ref[2, -1]) < 1e-12
    assert dices['dice_1'] == 0.0 and sorted(tr.volume_confusion) == ['0_a', '0_b', '2_c']
    assert int(tr.confusion.sum()) == 14 * 32 * 32 and 0.0 <= dice <= 1.0
    assert int(sum(v.sum() for v in tr.volume_confusion.values())) == 14 * 32 * 32


def test_validate_epoch_without_slice_names(exact, monkeypatch):
    """loaders that yield no names (inm=None): global counts only, as tests/test_modules_gpu.py uses it"""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    monkeypatch.setattr(cfg, "batch_size", 4)
    tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=32))
    batches = []
    for i, n in enumerate((4, 3)):
        x, y = O.synthetic_batch(n, 32, 70 + i)
        batches.append((x, y, torch.zeros(n, dtype=torch.int64), None))
    dice = tr.validate_epoch(batches)
    assert int(tr.confusion.sum()) == 7 * 32 * 32 and tr.volume_confusion == {} and 0.0 <= dice <= 1.0


def test_translation_sample_grid_matches_oracle(exact, tmp_path):
    """per-epoch sample grid (uganConsisTrainer.py:205-214): [x, G(x -> modality 0), ..., G(x -> modality 3)] along the
    width, de-normalised"""
    size = 64
    tr, G, D = _consis_trainer(size)
    x, _ = O.synthetic_batch(2, size, 17)
    modal = torch.tensor([1, 3])
    path = os.path.join(str(tmp_path), 'sample', 'train-1-images.png')
    grid = tr.sample_translations(x, modal, save_path=path)
    assert grid.shape == (2, 1, size, 5 * size) and grid.min() >= 0 and grid.max() <= 1
    org = O.label2onehot(modal, 4)
    cols = [x]
    for j in range(4):
        trg = O.label2onehot(torch.full((2,), j), 4)
        cols.append(O.ugannce_forward(G, x, trg - org, val_phase=True)[1])
    ref = ((torch.cat(cols, dim=3) + 1) / 2).clamp(0, 1)
    assert rel(grid, ref) < 1e-4          # saturated tanh outputs: fp32 rounding of a sign-like head
    try:
        import PIL  # noqa: F401
        assert os.path.exists(path)
    except ImportError:
        pass
    assert tr.net.training


def test_consis_train_epoch_loop_and_sample_file(exact, tmp_path, monkeypatch):
    """the epoch loop around train_step (loader cycling, random draws, LR mirror, iteration counter, sample grid)"""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.data_loader import syntheticLoader as synlod
    from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer
    monkeypatch.setattr(cfg, "batch_size", 2)
    tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=64))
    tr.expr_root, tr.save_samples, tr.semi_from_iter = str(tmp_path), True, 1
    lb = synlod.get_loader(None, 'train', 0, 2, size=64, pool_batches=1)       # one batch each: the loaders must cycle
    ul = synlod.get_loader(None, 'val', 0, 2, size=64, pool_batches=1)
    w0 = tr.net.seg_decoder.fc.weight.detach().clone()
    losses = tr.train_epoch(lb, ul, None, num_iter=2)
    assert tr.iter == 2 and losses.shape == (10,) and torch.isfinite(losses).all()
    assert losses[8] > 0                                                        # consistency loss on from iter 1
    assert not torch.equal(w0, tr.net.seg_decoder.fc.weight)
    assert abs(tr.optimizer.param_groups[0]['lr'] - O.poly_lr(1e-2, 1, cfg.max_epoch * cfg.num_iter_per_epoch)) < 1e-12
    try:
        import PIL  # noqa: F401
        assert os.path.exists(os.path.join(str(tmp_path), '000', 'sample', 'train-1-images.png'))
    except ImportError:
        pass


def test_ugan_and_cross_pse_epoch_loops(exact, tmp_path, monkeypatch):
    """train_epoch of UGANTrainer (lambda_shp ramp of uganTrainer.py:122-123, labelled-only loop) and crossPseTrainer"""
    from types import SimpleNamespace
    from smsut_b200 import config as cfg
    from smsut_b200.data_loader import syntheticLoader as synlod
    from smsut_b200.trainer.crossPseTrainer import crossPseTrainer
    from smsut_b200.trainer.uganTrainer import UGANTrainer
    monkeypatch.setattr(cfg, "batch_size", 2)
    args = SimpleNamespace(fold=0, expr_name=None, input_size=64)
    lb = synlod.get_loader(None, 'train', 0, 2, size=64, pool_batches=1)
    ul = synlod.get_loader(None, 'val', 0, 2, size=64, pool_batches=1)
    tr = UGANTrainer('train', args)
    assert tr.epoch_lambda_shp() == 0.0
    tr.epoch = 4
    assert tr.epoch_lambda_shp() == 2.0                    # epoch * (10 / 20), capped at lambda_seg
    tr.epoch = 100
    assert tr.epoch_lambda_shp() == 10.0
    tr.epoch = 4
    losses = tr.train_epoch(lb, ul, None, num_iter=2)
    assert tr.iter == 2 and losses.shape == (9,) and torch.isfinite(losses).all() and losses[8] > 0
    cp = crossPseTrainer('train', args)
    losses = cp.train_epoch(lb, ul, None, num_iter=2)
    assert cp.iter == 2 and losses.shape == (4,) and torch.isfinite(losses).all()
    assert abs(cp.optimizer2.param_groups[0]['lr'] - O.poly_lr(1e-2, 1, cfg.max_epoch * cfg.num_iter_per_epoch)) < 1e-12


@pytest.mark.parametrize("module", ["unetTrainer", "meanTeacherTrainer", "crossPseTrainer", "uganConsisTrainer", "uganTrainer"])
def test_cli_train_then_test_entry_points(exact, tmp_path, monkeypatch, capsys, module):
    """`python trainer/<x>Trainer.py -p train -f 0` then `-p test -f 0 -i 000 -wh last` (the reference's entry points,
    e.g. trainer/unetTrainer.py:150-171), through the modules' own __main__ blocks on the CPU test double."""
    import runpy
    from smsut_b200 import config as cfg
    monkeypatch.setattr(cfg, "batch_size", 2)
    monkeypatch.setattr(cfg, "input_size", 32)
    monkeypatch.setattr(cfg, "expr_root", str(tmp_path))
    name = f"smsut_b200.trainer.{module}"
    monkeypatch.setattr(sys, "argv", [module + ".py", "-p", "train", "-f", "0", "-nm", "cli", "--epochs", "1", "--iters", "1"])
    sys.modules.pop(name, None)
    runpy.run_module(name, run_name="__main__")
    ckpts = ["last_G.ckpt", "last_D.ckpt"] if module.startswith("ugan") else ["last.ckpt"]     # uganShp0Trainer.py:94-107
    for c in ckpts:
        assert os.path.exists(os.path.join(str(tmp_path), "cli", "000", "ckpt", c)), c
    monkeypatch.setattr(sys, "argv", [module + ".py", "-p", "test", "-f", "0", "-nm", "cli", "-i", "000", "-wh", "last"])
    runpy.run_module(name, run_name="__main__")
    out = capsys.readouterr().out
    assert "dice:" in out
    # BaseTrainer.test (baseTrainer.py:254-318): the modality-organ Dice matrix as <run>/all_trois_matrix.csv
    rows = [r for r in open(os.path.join(str(tmp_path), "cli", "000", "all_trois_matrix.csv")).read().split("\n") if r]
    assert len(rows) == cfg.n_modal + 1 and all(len(r.split(",")) == cfg.n_label + 1 for r in rows)
    assert ("dice: " + rows[-1].split(",")[-1]) in out


@pytest.mark.parametrize("module", ["unetTrainer", "uganConsisTrainer"])
def test_cli_pseudo_entry_point(exact, tmp_path, monkeypatch, capsys, module):
    """`-p pseudo -i 000` (uganConsisTrainer.py:329-332 -> saving_pseudo, baseTrainer.py:320-378 /
    uganConsisTrainer.py:216-306): prediction / label / input images of the test split, plus the translation strip
    for the GAN trainer, written under <run>/pseudo."""
    import runpy
    from PIL import Image
    from smsut_b200 import config as cfg
    monkeypatch.setattr(cfg, "batch_size", 2)
    monkeypatch.setattr(cfg, "input_size", 32)
    monkeypatch.setattr(cfg, "expr_root", str(tmp_path))
    name = f"smsut_b200.trainer.{module}"
    monkeypatch.setattr(sys, "argv", [module + ".py", "-p", "train", "-f", "0", "-nm", "cli", "--epochs", "1", "--iters", "1"])
    sys.modules.pop(name, None)
    runpy.run_module(name, run_name="__main__")
    monkeypatch.setattr(cfg, "pseudo_volumes", None, raising=False)      # the synthetic volumes are not the four named ones
    monkeypatch.setattr(sys, "argv", [module + ".py", "-p", "pseudo", "-f", "0", "-nm", "cli", "-i", "000", "-wh", "last"])
    runpy.run_module(name, run_name="__main__")
    out = os.path.join(str(tmp_path), "cli", "000", "pseudo")
    files = sorted(os.listdir(out))
    kinds = ("pse.jpg", "gt.jpg", "ori.jpg") + (("fk.jpg",) if module == "uganConsisTrainer" else ())
    assert len(files) == 8 * len(kinds)            # 4 test batches of 2 slices
    for k in kinds:
        assert sum(f.endswith(k) for f in files) == 8
    pse = Image.open(os.path.join(out, [f for f in files if f.endswith("pse.jpg")][0]))
    assert pse.size == (32, 32) and pse.mode == "RGB"
    if module == "uganConsisTrainer":
        fk = Image.open(os.path.join(out, [f for f in files if f.endswith("fk.jpg")][0]))
        assert fk.size == (32 * (cfg.n_modal + 1), 32)


def test_cli_coranet_pretrain_train_test(exact, tmp_path, monkeypatch, capsys):
    """coraNetTrainer's entry points (trainer/coraNetTrainer.py:746-776): `-p pretrain` (the reference's commented
    `prefit` line as a phase of its own) writes pre_best / pre_ema_best, `-p train -i 000` starts from them, predicts
    the pseudo labels and trains, `-p test -i 001` reloads the result."""
    import runpy
    from smsut_b200 import config as cfg
    monkeypatch.setattr(cfg, "batch_size", 2)
    monkeypatch.setattr(cfg, "input_size", 32)
    monkeypatch.setattr(cfg, "expr_root", str(tmp_path))
    name = "smsut_b200.trainer.coraNetTrainer"
    base = ["coraNetTrainer.py", "-f", "0", "-nm", "cli"]
    monkeypatch.setattr(sys, "argv", base + ["-p", "pretrain", "--epochs", "1", "--iters", "1"])
    sys.modules.pop(name, None)
    runpy.run_module(name, run_name="__main__")
    ck = os.path.join(str(tmp_path), "cli", "000", "ckpt")
    assert {"pre_best.ckpt", "pre_ema_best.ckpt", "pre_last.ckpt", "pre_ema_last.ckpt"} <= set(os.listdir(ck))
    monkeypatch.setattr(sys, "argv", base + ["-p", "train", "-i", "000", "--epochs", "1", "--iters", "1"])
    runpy.run_module(name, run_name="__main__")
    assert {"best.ckpt", "last.ckpt"} <= set(os.listdir(os.path.join(str(tmp_path), "cli", "001", "ckpt")))
    assert "Pseudo label dice" in capsys.readouterr().out
    monkeypatch.setattr(sys, "argv", base + ["-p", "test", "-i", "001", "-wh", "last"])
    runpy.run_module(name, run_name="__main__")
    assert "dice:" in capsys.readouterr().out
