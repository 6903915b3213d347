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
