"""Development: graph-replay time of growing prefixes / parts of the UGANConsisTrainer iteration (which phase costs what)."""
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from smsut_b200 import functional as Fn, ops  # noqa: E402
from smsut_b200.data_loader import syntheticLoader as synlod  # noqa: E402
from smsut_b200.graph import GraphedStep  # noqa: E402
from smsut_b200.network.blocks import refresh_packs  # noqa: E402
from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer  # noqa: E402

torch.manual_seed(0)
tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=256))
lb = synlod.get_loader(None, 'train', 0, 8, pool_batches=1)
ul = synlod.get_loader(None, 'val', 0, 8, pool_batches=1)
(x1, y, m1, _), (x2, _, m2, _) = next(iter(lb)), next(iter(ul))
batch = tr.prepare_batch(x1, y, m1, x2, m2, 2)
alpha, ids = tr.draw(16)
lam = torch.full((1,), 5.0, device="cuda")


def part(which):
    def fn(x_real, y_real, modal_org, modal_trg, vec_ot, vec_to, alpha, ids, lam):
        self = tr
        bs = y_real.shape[0]
        ops.arena_begin(x_real.device)
        y_fake, x_fake, feat_x_pool, sid = self.net(x_real, vec_ot, sample_ids=[ids])
        outs = [y_fake.float().sum()]
        if which == "fwd1":
            return torch.stack(outs)
        if which in ("fwd1+cycle", "fwd1+cycle+gbwd"):
            g_loss_seg = self.loss(y_fake[:bs], y_real)
            y_rec, x_rec, feat_f_pool, _ = self.net(x_fake, vec_to, sample_ids=sid)
            g_loss_rec = Fn.L1MeanFn.apply(x_rec.contiguous(), x_real)
            g_loss_semi = self.consistency_loss(y_rec, y_fake)
            g_loss_nce = self.nce_loss(feat_x_pool, feat_f_pool)
            gl = 10 * g_loss_rec + 10 * g_loss_seg + lam.reshape(()) * g_loss_semi + g_loss_nce
            if which == "fwd1+cycle+gbwd":
                self.optimizer.zero_grad()
                with Fn.accumulate_param_grads():
                    gl.backward()
                self.optimizer.finish_grads()
            return torch.stack([gl.detach()])
        # D phase
        x_fake_d = x_fake.detach()
        refresh_packs(self.D)
        with ops.parallel_branch(1) as b_fake:
            out_src_f, _ = self.D(x_fake_d)
            d_loss_fake = Fn.MeanFn.apply(out_src_f, 1.0)
        with ops.parallel_branch(2) as b_hat:
            x_hat = ops.lerp_rows(alpha, x_real, x_fake_d.contiguous()).requires_grad_(True)
            out_src_h, _ = self.D(x_hat)
            d_loss_gp = self.gradient_penalty(out_src_h, x_hat)
        out_src, out_cls = self.D(x_real)
        d_loss_real = Fn.MeanFn.apply(out_src, -1.0)
        d_loss_cls = Fn.CERowsFn.apply(out_cls.contiguous(), modal_org)
        b_fake.join(d_loss_fake)
        b_hat.join(d_loss_gp)
        d_loss = d_loss_real + d_loss_fake + d_loss_cls + 10 * d_loss_gp
        if which == "fwd1+Dfwd":
            return torch.stack([d_loss.detach()])
        self.d_optimizer.zero_grad()
        with Fn.accumulate_param_grads():
            d_loss.backward()
        self.d_optimizer.finish_grads()
        return torch.stack([d_loss.detach()])
    return fn


flush = torch.empty(160 * 2 ** 20, dtype=torch.uint8, device="cuda")
for which in ("fwd1", "fwd1+Dfwd", "fwd1+Dphase", "fwd1+cycle", "fwd1+cycle+gbwd"):
    step = GraphedStep(part(which), [*batch, alpha, ids[0], lam], warmup=2)
    for _ in range(2):
        step(*batch, alpha, ids[0], lam)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(*batch, alpha, ids[0], lam)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f"{which:18s} {tot / 5:7.3f} ms   launches {step.launches_per_replay}", flush=True)
    del step
