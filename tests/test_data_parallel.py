"""world_size-2 data-parallel host logic on CPU (gloo): gradient all-reduce over the flat buffers and the Dice
statistics all-reduce inside the loss must reproduce the single-process step on the concatenated batch -- what the
reference's nn.DataParallel computes on its gathered batch.  Kernels are the PyTorch test
double (exact mode); the NCCL path itself is exercised by bench.py --gpus N on the GPU box."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), SMSUT_ALLOW_CPU_TEST_DOUBLE="1")
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    torch.set_num_threads(2)
    import __graft_entry__ as g
    g.load_package()
    from types import SimpleNamespace

    import cpu_ops_mock
    from oracle import smsut_oracle as O
    from smsut_b200.parallel import DataParallelContext
    from smsut_b200.trainer.unetTrainer import UnetTrainer
    with cpu_ops_mock.installed(exact=True):
        par = DataParallelContext(backend="gloo")
        tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=64))
        tr.parallel = par
        tr.net.load_state_dict(O.make_weights(O.unet_shapes(), 21 + rank))   # replicas differ until the broadcast
        par.broadcast_params(tr.optimizer)
        losses = []
        for it in range(2):
            x, y = O.synthetic_batch(2, 64, 100 + 10 * it + rank)
            losses.append(tr.train_step(x, y).item())
        torch.save((rank, losses, {k: v.detach().clone() for k, v in tr.net.state_dict().items()}),
                   os.path.join(outdir, f"rank{rank}.pt"))
        par.close()


@pytest.mark.timeout(300)
def test_two_rank_unet_step_equals_single_process_on_concatenated_batch(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import smsut_oracle as O
    ctx = mp.get_context("spawn")
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(2)]
    sd = O.make_weights(O.unet_shapes(), 21)         # rank 0's weights win the broadcast
    st = {}
    ref_losses = []
    for it in range(2):
        xs, ys = zip(*[O.synthetic_batch(2, 64, 100 + 10 * it + r) for r in range(2)])
        loss, _ = O.unet_step(sd, st, torch.cat(xs), torch.cat(ys), O.poly_lr(1e-2, max(it - 1, 0), 30000))
        ref_losses.append(loss.item())

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-12)).item()
    # both ranks hold the same weights, equal to the single-process result
    for k in sd:
        assert rel(res[0][2][k], res[1][2][k]) < 1e-6, k
        assert rel(res[0][2][k], sd[k]) < 1e-3, k
    # each rank's loss = global Dice term + its LOCAL cross-entropy mean; their average is the global loss
    for it in range(2):
        avg = 0.5 * (res[0][1][it] + res[1][1][it])
        assert abs(avg - ref_losses[it]) < 1e-4, (it, avg, ref_losses[it])


def _gan_inputs(rank, bs=2, size=64):
    from oracle import smsut_oracle as O
    x1, y = O.synthetic_batch(bs, size, 11 + 10 * rank)
    x2, _ = O.synthetic_batch(bs, size, 12 + 10 * rank)
    alpha = torch.randn(2 * bs, generator=torch.Generator().manual_seed(3 + rank))
    return x1, y, x2, alpha


def _gan_worker(rank, world, port, outdir, early="0"):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), SMSUT_ALLOW_CPU_TEST_DOUBLE="1", SMSUT_EARLY_ALLREDUCE=early)
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    torch.set_num_threads(2)
    import __graft_entry__ as g
    g.load_package()
    from types import SimpleNamespace

    import cpu_ops_mock
    from oracle import smsut_oracle as O
    from smsut_b200.parallel import DataParallelContext
    from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer
    size, bs = 64, 2
    with cpu_ops_mock.installed(exact=True):
        par = DataParallelContext(backend="gloo")
        tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=size))
        tr.parallel = par
        tr.net.load_state_dict(O.make_weights(O.ugan_shapes(), 7 + rank))        # replicas differ until the broadcast
        tr.D.load_state_dict(O.make_weights(O.disc_shapes(size), 8 + rank))
        par.broadcast_params(tr.optimizer, tr.d_optimizer)
        x1, y, x2, alpha = _gan_inputs(rank)
        ids = [torch.randperm(16, generator=torch.Generator().manual_seed(0))]
        batch = tr.prepare_batch(x1, y, torch.full((bs,), 1), x2, torch.full((bs,), 3), 2)
        losses = tr.train_step(*batch, alpha, ids, 0.7, True)
        w = float(world)
        torch.save(dict(losses=losses.tolist(),
                        D={k: v.detach().clone() for k, v in tr.D.state_dict().items()},
                        G={k: v.detach().clone() for k, v in tr.net.state_dict().items()},
                        d_grads={k: p.grad.detach().clone() / w for k, p in tr.D.named_parameters()},
                        g_grads={k: p.grad.detach().clone() / w for k, p in tr.net.named_parameters()}),
                   os.path.join(outdir, f"gan{rank}.pt"))
        par.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("early", ["0", "1"])
def test_two_rank_ugan_consis_step_equals_single_process_on_global_batch(tmp_path, early):
    """The headline path at world_size 2 (gloo, fp32 test double): per-rank 2 labelled + 2 unlabelled slices, the two
    flat-gradient all-reduces and the Dice-statistic all-reduces inside both Dice/CE losses (segmentation and
    consistency) -- including the staged generator backward, whose first stage runs before the D phase -- against the
    oracle's single-process iteration on the global batch [lb0, lb1, ul0, ul1] (what nn.DataParallel computes).
    Teacher-forced across D's Adam step like tests/test_host_logic.py.  early = "1": the opt-in overlapped all-reduce of
    the early gradient bucket (SMSUT_EARLY_ALLREDUCE, trainer/uganConsisTrainer.py) must give the same step."""
    sys.path.insert(0, ROOT)
    from oracle import smsut_oracle as O
    ctx = mp.get_context("spawn")
    port = 29900 + os.getpid() % 90 + (100 if early == "1" else 0)
    procs = [ctx.Process(target=_gan_worker, args=(r, 2, port, str(tmp_path), early)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), f"gan{r}.pt")) for r in range(2)]

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-12)).item()

    size, bs = 64, 2
    ins = [_gan_inputs(r) for r in range(2)]
    x = torch.cat([ins[0][0], ins[1][0], ins[0][2], ins[1][2]])
    y = torch.cat([ins[0][1], ins[1][1]])
    modal = torch.cat([torch.full((2 * bs,), 1), torch.full((2 * bs,), 3)])
    alpha = torch.cat([ins[0][3][:bs], ins[1][3][:bs], ins[0][3][bs:], ins[1][3][bs:]]).view(-1, 1, 1, 1)
    ids = [torch.randperm(16, generator=torch.Generator().manual_seed(0))]
    G, D = O.make_weights(O.ugan_shapes(), 7), O.make_weights(O.disc_shapes(size), 8)      # rank 0's weights
    ref, d_grads = O.ugan_d_phase(G, D, {}, x, modal, 2, alpha, ids, 1e-2)
    # replicas stay identical, and their averaged D gradient is the global-batch gradient
    for k in D:
        assert rel(res[0]["D"][k], res[1]["D"][k]) < 1e-6, k
    # the GP double backward dominates these gradients (O(1e3 - 1e4) values).  Convolutions over 4 and over 8 slices round
    # differently; an fp32-rounding-level difference flips the LeakyReLU mask of the odd near-zero pre-activation near
    # D's output, which moves that sample's input gradient -- and with it every layer's GP gradient -- by a fraction of
    # a percent (measured: median 2.6e-3, max 6.9e-3; a missing all-reduce or a wrong 1/W would be off by 50-100 %)
    d_rel = sorted(rel(res[0]["d_grads"][k], d_grads[k]) for k in D)
    assert d_rel[len(d_rel) // 2] < 1e-2 and d_rel[-1] < 3e-2, (d_rel[len(d_rel) // 2], d_rel[-1])
    Dt = {k: v.clone() for k, v in res[0]["D"].items()}                                    # teacher-force the G phase
    g_ref, g_grads = O.ugan_g_phase(G, Dt, {}, x, y, modal, 2, ids, 1e-2, 1000, 0.7, nce_batch=2 * 8)      # PatchNCELoss(cfg.batch_size * world): the groups stay the per-rank ones
    ref.update(g_ref)
    keys = ('D_real', 'D_fake', 'D_cls', 'D_gp', 'G_fake', 'G_rec', 'G_cls', 'G_seg', 'G_semi', 'G_nce')
    for i, k in enumerate(keys):
        avg = 0.5 * (res[0]["losses"][i] + res[1]["losses"][i])      # means and (global Dice + local CE) both average
        assert abs(avg - ref[k]) < 5e-4 * max(1.0, abs(ref[k])), (k, avg, ref[k])
    for k in g_grads:
        assert rel(res[0]["G"][k], res[1]["G"][k]) < 1e-6, k
        assert rel(res[0]["g_grads"][k], g_grads[k]) < 8e-2, k


def _semi_worker(rank, world, port, outdir, which):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), SMSUT_ALLOW_CPU_TEST_DOUBLE="1")
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    torch.set_num_threads(2)
    import __graft_entry__ as g
    g.load_package()
    from types import SimpleNamespace

    import cpu_ops_mock
    from oracle import smsut_oracle as O
    from smsut_b200.parallel import DataParallelContext
    args = SimpleNamespace(fold=0, expr_name=None, input_size=64)
    with cpu_ops_mock.installed(exact=True):
        par = DataParallelContext(backend="gloo")
        if which == "mean_teacher":
            from smsut_b200.trainer.meanTeacherTrainer import MeanTeacherTrainer
            tr = MeanTeacherTrainer('train', args)
            tr.semi_from_iter = 1
            nets = dict(net=tr.net, ema=tr.ema)
            opts = [tr.optimizer]
        else:
            from smsut_b200.trainer.crossPseTrainer import crossPseTrainer
            tr = crossPseTrainer('train', args)
            nets = dict(net=tr.net, net2=tr.net2)
            opts = [tr.optimizer1, tr.optimizer2]
        tr.parallel = par
        for i, n in enumerate(nets.values()):
            n.load_state_dict(O.make_weights(O.unet_shapes(), 31 + i))          # same weights on both ranks
        par.broadcast_params(*opts)
        losses = []
        for it in range(2):
            x1, y = O.synthetic_batch(2, 64, 40 + it + 100 * rank)
            x2, _ = O.synthetic_batch(2, 64, 50 + it + 100 * rank)
            x = torch.cat([x1, x2])
            if which == "mean_teacher":
                noise = torch.clamp(torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(it + 10 * rank)) * 0.01,
                                    -0.02, 0.02)
                losses.append(tr.train_step(x, y, noise, 0.8).tolist())
            else:
                losses.append(tr.train_step(x, y, 0.05).tolist())
        torch.save(dict(losses=losses, nets={k: {n: v.detach().clone() for n, v in m.state_dict().items()}
                                             for k, m in nets.items()}), os.path.join(outdir, f"{which}{rank}.pt"))
        par.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("which", ["mean_teacher", "cross_pse"])
def test_two_rank_semi_supervised_steps_equal_single_process_on_global_batch(tmp_path, which):
    """config 4 (mean teacher: gradient + Dice-statistic all-reduce, EMA update local and identical on every rank) and
    the cross-pseudo-supervision trainer (two networks, four Dice/CE losses) at world_size 2 on gloo vs the oracle's
    single-process steps on the global batch [lb0, lb1, ul0, ul1]."""
    sys.path.insert(0, ROOT)
    from oracle import smsut_oracle as O
    ctx = mp.get_context("spawn")
    port = 29700 + os.getpid() % 90 + (0 if which == "mean_teacher" else 100)
    procs = [ctx.Process(target=_semi_worker, args=(r, 2, port, str(tmp_path), which)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), f"{which}{r}.pt")) for r in range(2)]

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-12)).item()

    sd1, sd2 = O.make_weights(O.unet_shapes(), 31), O.make_weights(O.unet_shapes(), 32)
    st1, st2 = {}, {}
    for it in range(2):
        xs1, ys, xs2 = [], [], []
        for r in range(2):
            x1, y = O.synthetic_batch(2, 64, 40 + it + 100 * r)
            x2, _ = O.synthetic_batch(2, 64, 50 + it + 100 * r)
            xs1.append(x1); ys.append(y); xs2.append(x2)
        x, y = torch.cat(xs1 + xs2), torch.cat(ys)
        lr = O.poly_lr(1e-2, max(it - 1, 0), 30000)
        if which == "mean_teacher":
            noise = torch.cat([torch.clamp(torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(it + 10 * r))
                                           * 0.01, -0.02, 0.02) for r in range(2)])
            ref = O.mean_teacher_step(sd1, sd2, st1, x, y, noise, lr, it, 0.8, warm=1)
        else:
            d, _, _ = O.cross_pse_step(sd1, sd2, st1, st2, x, y, lr, 0.05)
            ref = [d[k] for k in ("seg1", "seg2", "semi1", "semi2")]
        for i, v in enumerate(ref):
            avg = 0.5 * (res[0]["losses"][it][i] + res[1]["losses"][it][i])
            tol = 1e-4 if it == 0 else 2e-3     # iteration 1 sees argmax pseudo-labels / rounding of updated weights
            assert abs(avg - float(v)) < tol * max(1.0, abs(float(v))), (it, i, avg, float(v))
    names = list(res[0]["nets"])
    for name, sd in zip(names, (sd1, sd2)):
        for k in sd:
            assert rel(res[0]["nets"][name][k], res[1]["nets"][name][k]) < 1e-6, (name, k)
            assert rel(res[0]["nets"][name][k], sd[k]) < 2e-3, (name, k)


# ----------------------------------------------------------------------------------------------------------------------
# data parallelism behind the trainer API: `torchrun ... trainer/<x>Trainer.py -p train` = fit() on every rank
# ----------------------------------------------------------------------------------------------------------------------
def _fit_worker(rank, world, port, outdir, trainer_name, size):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), SMSUT_ALLOW_CPU_TEST_DOUBLE="1", SMSUT_TENSORBOARD="0")
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    torch.set_num_threads(2)
    import random
    import numpy as np
    import __graft_entry__ as g
    g.load_package()
    from types import SimpleNamespace

    import cpu_ops_mock
    from smsut_b200 import config as cfg
    cfg.batch_size, cfg.expr_root = 2, os.path.join(outdir, "expr")
    module = {"UnetTrainer": "unetTrainer", "UGANConsisTrainer": "uganConsisTrainer"}[trainer_name]
    mod = __import__(f"smsut_b200.trainer.{module}", fromlist=[trainer_name])
    with cpu_ops_mock.installed(exact=True):
        random.seed(cfg.seed); np.random.seed(cfg.seed); torch.manual_seed(cfg.seed)      # what __main__ does on every rank
        tr = getattr(mod, trainer_name)('train', SimpleNamespace(fold=0, expr_name="dp", input_size=size))
        seen = []
        step = tr.train_step
        tr.train_step = lambda *a, **k: seen.append([t.detach().clone() if isinstance(t, torch.Tensor) else t for t in a]) or step(*a, **k)
        tr.fit('synthetic', max_epoch=1, iters_per_epoch=2)
        first = {k: v.detach().clone() for k, v in tr.net.state_dict().items()}
        torch.save(dict(rank=rank, weights=first, inputs=seen, meters=[m.cur_values for m in tr.meters],
                        is_main=tr.is_main, world=tr.parallel.world), os.path.join(outdir, f"fit{rank}.pt"))
        tr.parallel.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("trainer_name,size", [("UnetTrainer", 32), ("UGANConsisTrainer", 64)])
def test_two_rank_fit_through_the_trainer_api(tmp_path, trainer_name, size):
    """fit() under WORLD_SIZE = 2 (what `torchrun --nproc-per-node 2 trainer/<x>Trainer.py -p train` runs): the replicas
    attach the data-parallel context themselves, train on DIFFERENT slices, end with identical weights, agree on the
    target modality of every iteration, and only rank 0 writes the run directory."""
    ctx = mp.get_context("spawn")
    port = 29500 + (os.getpid() + 137 + len(trainer_name)) % 400
    procs = [ctx.Process(target=_fit_worker, args=(r, 2, port, str(tmp_path), trainer_name, size)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), f"fit{r}.pt"), weights_only=False) for r in range(2)]
    assert res[0]["is_main"] and not res[1]["is_main"] and res[0]["world"] == 2
    for k, v in res[0]["weights"].items():
        assert torch.equal(v, res[1]["weights"][k]), k                   # replicas stay in step
    a, b = res[0]["inputs"], res[1]["inputs"]
    assert len(a) == len(b) == 2
    assert not torch.equal(a[0][0], b[0][0])                             # different slices per replica
    if trainer_name == "UGANConsisTrainer":
        # modal_trg (argument 3 of train_step, prepare_batch's order): ONE target modality per iteration, shared
        for it in range(2):
            assert int(a[it][3][0]) == int(b[it][3][0]), "the replicas drew different target modalities"
            assert len(set(a[it][3].tolist())) == 1
    run = os.path.join(str(tmp_path), "expr", "dp")
    assert sorted(os.listdir(run)) == ["000"]                            # one run directory: rank 0's
    ck = os.listdir(os.path.join(run, "000", "ckpt"))
    assert any(c.startswith("last") for c in ck) and any(c.startswith("best") for c in ck)
    log = open(os.path.join(run, "000", "train.log")).read()
    assert log.count("[TRN] Epoch:") == 1 and log.count("[TST] Epoch:") == 1
    # the train meter holds the mean over BOTH replicas' slices (summed before update_cur): equal on both ranks, and
    # weighted with twice one replica's slice count
    for k, v in res[0]["meters"][0].items():
        assert abs(v - res[1]["meters"][0][k]) < 1e-9, k
    assert res[0]["meters"][0]["loss"] > 0
    # the test stage runs on every replica with the same weights and the same test slices
    for k, v in res[0]["meters"][1].items():
        assert abs(v - res[1]["meters"][1][k]) < 1e-6, k
