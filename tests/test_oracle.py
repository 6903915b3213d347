"""Pins the oracle (oracle/smsut_oracle.py) against fixtures produced by the REAL reference modules
(tests/golden/make_golden.py, run once in the build container where /root/reference is mounted)."""
import os

import numpy as np
import pytest
import torch

from oracle import smsut_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name + ".npz")))


def close(a, b, tol=1e-4):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / (b.norm() + 1e-12)).item() < tol


def leaf(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def check_grad_norms(fix, prefix, leafs, tol=1e-3):
    n = 0
    for k, v in leafs.items():
        key = prefix + k
        if key in fix:
            n += 1
            assert abs(v.grad.norm().item() - float(fix[key])) < tol * max(1e-3, float(fix[key])), k
    assert n > 10


def test_unet_matches_reference_fixture():
    f = load("unet")
    sd = leaf(O.make_weights(O.unet_shapes(), 1))
    x, y = O.synthetic_batch(2, 64, 3)
    out = O.unet_forward(sd, x)
    assert close(out.detach(), f["logits"], 1e-5)
    loss = O.dice_ce_loss(out, y)
    assert abs(loss.item() - float(f["loss"])) < 1e-6
    loss.backward()
    assert close(sd["decoder.fc.weight"].grad, f["fc_grad"]) and close(sd["encoder.pre_conv.weight"].grad, f["pre_grad"])
    check_grad_norms(f, "gn.", sd)


def test_ugannce_matches_reference_fixture():
    f = load("ugannce")
    sd = leaf(O.make_weights(O.ugan_shapes(), 4))
    x, _ = O.synthetic_batch(2, 64, 4)
    m = torch.tensor([[1., 0, -1, 0], [0, 1., -1, 0]])
    ids = [torch.as_tensor(f["ids"])]
    seg, tsl, feats, _ = O.ugannce_forward(sd, x, m, sample_ids=ids)
    assert close(seg.detach(), f["seg"], 1e-5) and close(tsl.detach(), f["tsl"], 1e-5)
    assert close(feats[0].detach(), f["feat"], 1e-5)
    (seg.mean() + tsl.mean() + (feats[0] ** 3).sum()).backward()
    check_grad_norms(f, "gn.", sd)
    assert len(O.ugannce_forward(sd, x, val_phase=True)) == 2


def test_discriminator_and_gradient_penalty_match_reference_fixture():
    f = load("discriminator")
    sd = leaf(O.make_weights(O.disc_shapes(64), 5))
    x_hat = torch.as_tensor(f["x_hat"]).requires_grad_(True)
    out_src, out_cls = O.discriminator_forward(sd, x_hat)
    assert close(out_src.detach(), f["out_src"], 1e-5) and close(out_cls.detach(), f["out_cls"], 1e-5)
    gp = O.gradient_penalty(out_src, x_hat)
    assert abs(gp.item() - float(f["gp"])) < 1e-4 * float(f["gp"])
    (10 * gp + out_src.mean() + out_cls.pow(2).mean()).backward()
    check_grad_norms(f, "gn.", sd)


def test_losses_match_reference_fixture():
    f = load("losses")
    logits = torch.as_tensor(f["logits"]).requires_grad_(True)
    l = O.dice_ce_loss(logits, torch.as_tensor(f["labels"]))
    assert abs(l.item() - float(f["dice_ce"])) < 1e-6
    l.backward()
    assert close(logits.grad, f["dlogits"], 1e-5)
    q = torch.as_tensor(f["q"]).requires_grad_(True)
    rows = O.patchnce_loss(q, torch.as_tensor(f["k"]), 8)
    assert close(rows.detach(), f["nce_rows"], 1e-5)
    rows.mean().backward()
    assert close(q.grad, f["dq"], 1e-5)


def test_two_consis_iterations_match_reference_fixture():
    """two consecutive iterations (consistency loss off, then on) with torch.optim SGD/Adam state carried over"""
    f = load("consis_step")
    size, bs = 64, 2
    Gw, Dw = O.make_weights(O.ugan_shapes(), 7), O.make_weights(O.disc_shapes(size), 8)
    g_state, d_state = {}, {}
    x1, y = O.synthetic_batch(bs, size, 11)
    x2, _ = O.synthetic_batch(bs, size, 12)
    x_real = torch.cat([x1, x2])
    modal = torch.cat([torch.full((bs,), 1), torch.full((bs,), 3)])
    keys = ('D_real', 'D_fake', 'D_cls', 'D_gp', 'G_fake', 'G_rec', 'G_cls', 'G_seg', 'G_semi', 'G_nce')
    for s in (0, 1):
        losses, d_grads, g_grads = O.ugan_consis_step(
            Gw, Dw, g_state, d_state, x_real, y, modal, 2, torch.as_tensor(f[f"s{s}.alpha"]),
            [torch.as_tensor(f[f"s{s}.ids"])], 1e-2, 1000 if s else 0, 0.7, nce_batch=8)
        # step 1 starts from weights that went through Adam's sign-like first update: fp32 round-off is amplified
        tol = 1e-4 if s == 0 else 3e-2
        for k, ref in zip(keys, f[f"s{s}.losses"]):
            assert abs(losses[k] - ref) < tol * max(1.0, abs(ref)), (s, k, losses[k], ref)
        for k, v in d_grads.items():
            ref = float(f[f"s{s}.dgn.{k}"])
            assert abs(v.norm().item() - ref) < (1e-3 if s == 0 else 0.2) * max(ref, 1e-3), (s, k)
        if s == 0:
            for k, v in g_grads.items():
                ref = float(f[f"s{s}.ggn.{k}"])
                assert abs(v.norm().item() - ref) < 5e-2 * max(ref, 1e-3), (s, k)
            dsum = np.array([v.double().sum().item() for v in Dw.values()])
            assert np.abs(dsum - f["s0.D_checksum"]).max() < 0.5      # Adam: lr * sign(g) on near-zero gradients
            gsum = np.array([v.double().sum().item() for v in Gw.values()])
            assert np.allclose(gsum, f["s0.G_checksum"], rtol=2e-2, atol=0.5)


def test_schedules_and_helpers():
    assert O.poly_lr(1e-2, 0, 30000) == 1e-2 and abs(O.poly_lr(1e-2, 15000, 30000) - 1e-2 * 0.5 ** 0.9) < 1e-12
    assert O.sigmoid_rampup(0, 200) == pytest.approx(float(np.exp(-5.0)))
    assert O.sigmoid_rampup(200, 200) == 1.0 and O.sigmoid_rampup(5, 0) == 1.0
    assert O.ema_alpha(50) == 0.0 and O.ema_alpha(100) == 0.99 and O.ema_alpha(10 ** 6) == 0.99
    oh = O.label2onehot(torch.tensor([0, 3, 1]), 4)
    assert oh.tolist() == [[1, 0, 0, 0], [0, 0, 0, 1], [0, 1, 0, 0]]
    img, lab = O.synthetic_batch(2, 64, 0)
    assert img.shape == (2, 1, 64, 64) and img.min() >= -1 and img.max() <= 1 and set(lab.unique().tolist()) <= {0, 1, 2, 3, 4}
    assert 0.5 < (lab == 0).float().mean() < 0.95


def test_unet_batchnorm_relu_matches_reference_fixture():
    """UNet's default norm / activation (network/unet.py:14: batch norm + ReLU): two training passes (running
    estimates updated twice) and an eval pass against the real reference's outputs."""
    f = load("unet_bn")
    sd = leaf(O.make_weights(O.unet_shapes(), 11))
    sd = O.add_bn_buffers(sd)
    st = O.Style("batch", "relu", training=True)
    for it, seed in enumerate((21, 22)):
        x, y = O.synthetic_batch(2, 48, seed)
        for v in sd.values():
            v.grad = None
        out = O.unet_forward(sd, x, style=st)
        assert close(out.detach(), f[f"logits{it}"], 1e-5)
        loss = O.dice_ce_loss(out, y)
        assert abs(loss.item() - float(f[f"loss{it}"])) < 1e-6
        loss.backward()
        check_grad_norms(f, f"gn{it}.", {k: v for k, v in sd.items() if v.requires_grad})
    assert close(sd["decoder.fc.weight"].grad, f["fc_grad"]) and close(sd["encoder.pre_conv.weight"].grad, f["pre_grad"])
    bufs = [k for k in f if k.startswith("buf.")]
    assert len(bufs) == 2 * 28
    for k in bufs:
        assert close(sd[k[4:]], f[k], 1e-5), k
    x, _ = O.synthetic_batch(2, 48, 23)
    out = O.unet_forward(sd, x, style=O.Style("batch", "relu", training=False))
    assert close(out.detach(), f["logits_eval"], 1e-5)


def _norms(f, prefix):
    return dict(zip(str(f[prefix + "names"]).split(","), f[prefix + "norms"].tolist()))


def test_cross_pseudo_supervision_steps_match_reference_fixture():
    """two consecutive crossPseTrainer iterations (trainer/crossPseTrainer.py:96-131) with SGD momentum carried over"""
    f = load("siblings")
    sd1, sd2 = O.make_weights(O.unet_shapes(), 31), O.make_weights(O.unet_shapes(), 32)
    st1, st2 = {}, {}
    for it in range(2):
        x1, y = O.synthetic_batch(2, 64, 41 + it)
        x2, _ = O.synthetic_batch(2, 64, 51 + it)
        losses, g1, g2 = O.cross_pse_step(sd1, sd2, st1, st2, torch.cat([x1, x2]), y, 1e-2, 0.05)
        got = [losses[k] for k in ("seg1", "seg2", "semi1", "semi2")]
        assert np.allclose(got, f[f"cps{it}.losses"], rtol=2e-4 if it == 0 else 2e-3, atol=1e-5), (it, got)
        for g, name in ((g1, "g1"), (g2, "g2")):
            for k, ref in _norms(f, f"cps{it}.{name}.").items():
                assert abs(g[k].norm().item() - ref) < (1e-3 if it == 0 else 2e-2) * max(ref, 1e-3), (it, name, k)
        for sd, name in ((sd1, "sum1"), (sd2, "sum2")):
            s = np.array([v.double().sum().item() for v in sd.values()])
            assert np.allclose(s, f[f"cps{it}.{name}"], rtol=1e-3, atol=5e-2)


@pytest.mark.parametrize("name,lambda_shp", [("shp", 3.5), ("shp0", None)])
def test_ugan_shape_step_matches_reference_fixture(name, lambda_shp):
    """one UGANTrainer iteration (trainer/uganTrainer.py:159-196, shape loss) / UGANShp0Trainer iteration
    (trainer/uganShp0Trainer.py:180-217) on the reference's `UGAN` generator (no PatchNCE head)"""
    f = load("siblings")
    shapes = {k: v for k, v in O.ugan_shapes().items() if not k.startswith("netF.")}
    G, D = O.make_weights(shapes, 61), O.make_weights(O.disc_shapes(64), 62)
    x, y = O.synthetic_batch(3, 64, 63)
    losses, d_grads, g_grads = O.ugan_shape_step(G, D, {}, {}, x, y, torch.full((3,), 1), 3,
                                                 torch.as_tensor(f[f"{name}.alpha"]), 1e-2, lambda_shp=lambda_shp)
    keys = ["D_real", "D_fake", "D_cls", "D_gp", "G_fake", "G_rec", "G_cls", "G_seg"] + (["G_shp"] if lambda_shp else [])
    for k, ref in zip(keys, f[f"{name}.losses"]):
        assert abs(losses[k] - ref) < 1e-4 * max(1.0, abs(ref)), (k, losses[k], ref)
    for k, ref in _norms(f, f"{name}.dgn.").items():
        assert abs(d_grads[k].norm().item() - ref) < 1e-3 * max(ref, 1e-3), k
    for k, ref in _norms(f, f"{name}.ggn.").items():
        assert abs(g_grads[k].norm().item() - ref) < 5e-2 * max(ref, 1e-3), k
    gsum = np.array([v.double().sum().item() for v in G.values()])
    assert np.allclose(gsum, f[f"{name}.G_checksum"], rtol=2e-2, atol=0.5)
    dsum = np.array([v.double().sum().item() for v in D.values()])
    assert np.abs(dsum - f[f"{name}.D_checksum"]).max() < 0.5      # Adam: lr * sign(g) on near-zero gradients


def test_mean_teacher_steps_match_reference_fixture():
    """three meanTeacherTrainer iterations (trainer/meanTeacherTrainer.py:95-153) starting at iter 99: the consistency
    loss and the EMA decay switch on at iter 100 (config 4 of BASELINE.json)"""
    f = load("mean_teacher")
    sd, ema = O.make_weights(O.unet_shapes(), 71), O.make_weights(O.unet_shapes(), 72)
    st = {}
    for k in range(3):
        it = 99 + k
        x1, y = O.synthetic_batch(2, 64, 81 + k)
        x2, _ = O.synthetic_batch(2, 64, 91 + k)
        noise = torch.clamp(torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(k)) * 0.01, -0.02, 0.02)
        seg, semi = O.mean_teacher_step(sd, ema, st, torch.cat([x1, x2]), y, noise, O.poly_lr(1e-2, it - 1, 30000), it, 0.8)
        assert abs(seg - f[f"losses{k}"][0]) < 1e-5 * (1 + k * 20) and abs(semi - f[f"losses{k}"][1]) < 1e-6 * (1 + k * 20)
        assert (semi == 0.0) == (it < 100)
        for name, d in (("net", sd), ("ema", ema)):
            norms = np.array([v.double().norm().item() for v in d.values()])
            sums = np.array([v.double().sum().item() for v in d.values()])
            assert np.allclose(norms, f[f"{name}_norm{k}"], rtol=1e-4, atol=1e-5), (k, name)
            assert np.allclose(sums, f[f"{name}_sum{k}"], rtol=1e-3, atol=2e-3), (k, name)
    assert close(sd["decoder.fc.weight"], f["fc"], 1e-4) and close(ema["decoder.fc.weight"], f["ema_fc"], 1e-4)


def test_coranet_iterations_match_reference_fixture():
    """coraNetTrainer (trainer/coraNetTrainer.py): a pre_epoch iteration (:461-499), pred_unlabel (:186-207) and two
    train_epoch iterations (:264-352, before / after the iter-1000 switch) against tests/golden/coranet.npz, which
    make_golden_coranet.py produced with the reference's UNet and the trainer's own loss class lifted from its source"""
    f = load("coranet")
    n_out = int(f["n_out"])
    sd, ema = O.make_weights(O.unet_shapes(out_ch=n_out), 71), O.make_weights(O.unet_shapes(out_ch=n_out), 72)
    st = {}

    def check(tag, losses, grads, rtol_l, rtol_g):
        assert np.allclose(losses, f[tag + ".losses"], rtol=rtol_l, atol=1e-6), (tag, losses, f[tag + ".losses"])
        for k, ref in zip(str(f[tag + ".names"]).split(","), f[tag + ".gnorms"].tolist()):
            assert abs(grads[k].norm().item() - ref) < rtol_g * max(ref, 1e-3), (tag, k, grads[k].norm().item(), ref)
        s = np.array([v.double().sum().item() for k, v in sd.items() if k in dict.fromkeys(str(f[tag + ".names"]).split(","))])
        assert np.allclose(s, f[tag + ".sum_net"], rtol=1e-3, atol=5e-2), tag
        e = np.array([v.double().sum().item() for k, v in ema.items() if k in dict.fromkeys(str(f[tag + ".names"]).split(","))])
        assert np.allclose(e, f[tag + ".sum_ema"], rtol=1e-3, atol=5e-2), tag

    img1, msk = O.synthetic_batch(2, 64, 81)
    losses, grads = O.coranet_pre_step(sd, ema, st, img1, msk, 1e-2, 200)
    check("pre", losses, grads, 2e-5, 1e-3)
    imgu, _ = O.synthetic_batch(2, 64, 83)
    plab, mask = O.coranet_pred_unlabel(sd, imgu)
    assert (plab.numpy() != f["pred.plab"]).mean() < 1e-3 and (mask.numpy() != f["pred.mask"]).mean() < 1e-3
    plab, mask = torch.as_tensor(f["pred.plab"]).long(), torch.as_tensor(f["pred.mask"]).float()
    for tag, it in (("trn_early", 300), ("trn_late", 1500)):
        img1, msk = O.synthetic_batch(2, 64, 91 + it)
        losses, grads = O.coranet_train_step(sd, ema, st, img1, msk, imgu, plab, mask, 1e-2, it, 0.3)
        check(tag, losses, grads, 2e-3, 2e-2)
