# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/uganConsisTrainer.py: the UGANConsisTrainer iteration
(uganConsisTrainer.py:66-214) -- D step (WGAN-GP + modality classification) and G step (adversarial,
classification, Dice/CE segmentation, L1 cycle, pseudo-label consistency, PatchNCE) -- on libsmsut_b200 kernels.

Differences that do not change the arithmetic:
  * the D-phase generator forward (L133) and the G-phase one (L151) see the same weights and inputs, so ONE
    forward serves both (its detached translation feeds D);
  * the weight gradients of D produced by g_loss.backward() are zeroed unread by the reference (L144): they are
    not computed;
  * the ten losses stay on the device (one D2H copy when logging) instead of eleven .item() syncs per step;
  * the poly LR schedule ticks on the device.
"""
import argparse
import os
import random
import sys
import time

if __package__ in (None, ""):      # `python trainer/uganConsisTrainer.py -p train -f 0` from the package directory
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import __graft_entry__ as _g
    _g.load_package()
    __package__ = "smsut_b200.trainer"

import numpy as np
import torch
import torch.nn as nn

from .. import config as cfg
from .. import functional as Fn
from .. import ops
from ..network.blocks import refresh_packs
from .uganShp0Trainer import UGANShp0Trainer

# stage the generator's backward so that the cycle pass's half runs beside the discriminator phase (train_step)
SPLIT_G_BACKWARD = [os.environ.get("SMSUT_SPLIT_G_BACKWARD", "1") != "0"]

# run the first pass's segmentation half as labelled (with graph) + unlabelled (no graph) sub-batches (train_step)
SPLIT_SEG_ROWS = [os.environ.get("SMSUT_SPLIT_SEG_ROWS", "1") != "0"]

# D(x_real) and D(G(x).detach()) of the D phase as one 2B-slice pass (train_step); measured: see DESIGN.md
BATCH_D_REAL_FAKE = [os.environ.get("SMSUT_BATCH_D", "0") != "0"]

# data-parallel runs: all-reduce the early gradient bucket (segmentation halves + netF) on a communication stream beside
# the discriminator phase (train_step).  Built, correct (2-rank gloo test, NCCL parity check of bench.py) and MEASURED
# SLOWER on 2 x B200: 10.76 ms / step with the overlapped bucket against 10.31 ms with one all-reduce after the backward
# (N = 1: 10.22 ms) -- the collective's CTAs sit on SMs through the busiest window of the iteration, and every NCCL call
# shares one stream, so the discriminator's all-reduce queues behind it.  Off by default (SMSUT_EARLY_ALLREDUCE=1).
EARLY_ALLREDUCE = [os.environ.get("SMSUT_EARLY_ALLREDUCE", "0") != "0"]
# the weight-gradient kernels of stage A (the generator's early backward, beside the discriminator phase) are queued and
# issued after the discriminator's backward instead: they then fill the SMs the latency-bound tail of the iteration
# (D step -> D(G(x)) -> the first pass's translation-half backward) leaves idle (ops.defer_begin / run_deferred)
DEFER_STAGE_A_WGRADS = [os.environ.get("SMSUT_DEFER_WGRAD", "1") != "0"]

LOSS_KEYS = ('D_real', 'D_fake', 'D_cls', 'D_gp', 'G_fake', 'G_rec', 'G_cls', 'G_seg', 'G_semi', 'G_nce')


class UGANConsisTrainer(UGANShp0Trainer):
    def __init__(self, phase, args):
        super(UGANConsisTrainer, self).__init__(phase, args)
        self.lambda_semi = 10
        self.semi_from_iter = 1000
        self.parallel = None           # parallel.DataParallelContext when launched under torchrun

    def consistency_loss(self, source, target):
        # Dice+CE of `source` against argmax(target): the argmax is taken inside the fused loss kernel
        return self.loss(source, target)

    def nce_loss(self, feat_x_pool, feat_f_pool):
        n_layers = cfg.nce_layers
        n = len(n_layers)
        total_nce_loss = 0.0
        for f_f, f_x, crit, nce_layer in zip(feat_f_pool, feat_x_pool, self.criterionNCE, n_layers):
            loss = crit(f_f, f_x) * 1.0
            total_nce_loss += loss.mean()
        return total_nce_loss / n

    # ------------------------------------------------------------------------------------------
    def train_step(self, x_real, y_real, modal_org, modal_trg, vec_ot, vec_to, alpha, sample_ids, lambda_semi,
                   use_semi):
        """The body of the hot loop (uganConsisTrainer.py:129-203) on device tensors.  `alpha` (B,) ~ N(0,1) and
        `sample_ids` ([ids (64,) int64]) are the iteration's random draws (L138, network/ugan.py:321).
        Returns the ten losses as one device vector ordered like LOSS_KEYS."""
        lambda_cls, lambda_gp = self.lambda_cls, self.lambda_gp
        lambda_seg, lambda_rec = self.lambda_seg, self.lambda_rec
        bs = y_real.shape[0]
        ops.arena_begin(x_real.device)      # one memset serves every zeroed accumulator of the iteration
        self.lr_sched.tick()

        # generator forward shared by the D phase (detached) and the G phase
        split = SPLIT_G_BACKWARD[0]
        if SPLIT_SEG_ROWS[0]:
            # only the labelled slices' logits are differentiated (L153: loss(y_fake[:bs], y_real)); the unlabelled ones
            # are argmax targets of the consistency loss (needed only when it is on): no autograd graph for them
            (y_fake_lb, y_fake_ul), x_fake, feat_x_pool, sample_ids = self.net(
                x_real, vec_ot, sample_ids=sample_ids, seg_rows=bs, seg_rest=bool(use_semi))
        else:
            y_fake, x_fake, feat_x_pool, sample_ids = self.net(x_real, vec_ot, sample_ids=sample_ids)
            y_fake_lb, y_fake_ul = y_fake[:bs], None

        # ---------------- the cycle pass of the G phase (L157-168) needs only x_fake and G's weights: it runs on a
        # branch stream beside the whole D phase (forward, gradient penalty, backward, Adam step).
        # g_loss = [adversarial + classification terms, which need the UPDATED discriminator]
        #        + g_partial [cycle L1, segmentation, consistency, PatchNCE: functions of G alone].
        # Gradients add, so with `split` the backward of g_partial starts as soon as the cycle forward is done --
        # through the whole cycle pass down to d g_partial / d x_fake and through the first pass's segmentation
        # half -- beside the D phase; only the first pass's translation half waits for the discriminator (it is
        # back-propagated ONCE, with both contributions to d/d x_fake summed).  That takes the cycle-pass backward,
        # half of the generator's backward work, off the serial path behind the D phase.
        if isinstance(lambda_semi, torch.Tensor):
            lambda_semi = lambda_semi.reshape(())      # device scalar: one captured graph serves every epoch
        if split:
            self.optimizer.zero_grad()                 # before the fork: stage A accumulates into G's gradients
        x_fake_c = x_fake.detach().requires_grad_(True) if split else x_fake
        with ops.parallel_branch(4) as b_cyc:
            g_loss_seg = self.loss(y_fake_lb, y_real)
            # y_rec only feeds the consistency loss (L162-168): while that is off (iter < semi_from_iter) the cycle
            # pass's segmentation half is not run at all (the reference computes it and drops it)
            y_rec, x_rec, feat_f_pool, _ = self.net(x_fake_c, vec_to, sample_ids=sample_ids,
                                                    want_seg=bool(use_semi) or not SPLIT_SEG_ROWS[0])
            g_loss_rec = Fn.L1MeanFn.apply(x_rec.contiguous(), x_real)
            if use_semi:
                if y_fake_ul is not None:
                    y_fake = torch.cat([y_fake_lb.detach(), y_fake_ul], dim=0)
                g_loss_semi = self.consistency_loss(y_rec, y_fake)
            else:
                g_loss_semi = torch.zeros((), device=x_real.device)
            g_loss_nce = self.nce_loss(feat_x_pool, feat_f_pool)
            g_partial = lambda_rec * g_loss_rec + lambda_seg * g_loss_seg + lambda_semi * g_loss_semi + 1.0 * g_loss_nce
            if split:
                with Fn.accumulate_param_grads(join=False, side_group=1):
                    if DEFER_STAGE_A_WGRADS[0] and not EARLY_ALLREDUCE[0]:
                        ops.defer_begin()
                    g_partial.backward()
                    deferred = ops.defer_end()
                dx_fake_cyc = x_fake_c.grad
        stage_a = ops.pending_detach() if split else None   # the D phase joins only what it forks itself
        if split and self.parallel is not None and EARLY_ALLREDUCE[0]:
            # overlap: the segmentation halves and netF have their complete gradient once stage A is through -- their
            # bucket is all-reduced on a communication stream while the discriminator phase and the rest of the
            # generator's backward run; the final all_reduce_grads() handles the other half
            producers = list(stage_a[0]) + list(stage_a[2])
            if b_cyc.on:
                producers.append(b_cyc.stream)
            self.parallel.all_reduce_early(self.optimizer, producers)

        # ---------------- D phase (L129-146): the three discriminator passes are independent chains of small
        # kernels -> real on the current stream, fake and interpolated on branch streams (ops.parallel_branch)
        x_fake_d = x_fake.detach()
        refresh_packs(self.D)       # before the fork: all three passes read the bf16 weight copies this launch writes
        with ops.parallel_branch(2) as b_hat:
            x_hat = ops.lerp_rows(alpha, x_real, x_fake_d.contiguous()).requires_grad_(True)
            out_src_h, _ = self.D(x_hat)
            d_loss_gp = self.gradient_penalty(out_src_h, x_hat)
        if BATCH_D_REAL_FAKE[0]:
            # D(x_real) and D(G(x).detach()) as ONE pass over 2B slices: every layer of D is per-sample (InstanceNorm),
            # so the values are those of two passes; the discriminator's kernels are far smaller than the machine
            nb = x_real.shape[0]
            out_src, out_cls = self.D(torch.cat([x_real, x_fake_d], dim=0))
            d_loss_real = Fn.MeanFn.apply(out_src[:nb].contiguous(), -1.0)
            d_loss_fake = Fn.MeanFn.apply(out_src[nb:].contiguous(), 1.0)
            d_loss_cls = Fn.CERowsFn.apply(out_cls[:nb].contiguous(), modal_org)
        else:
            with ops.parallel_branch(1) as b_fake:
                out_src_f, _ = self.D(x_fake_d)
                d_loss_fake = Fn.MeanFn.apply(out_src_f, 1.0)
            out_src, out_cls = self.D(x_real)
            d_loss_real = Fn.MeanFn.apply(out_src, -1.0)
            d_loss_cls = Fn.CERowsFn.apply(out_cls.contiguous(), modal_org)
            b_fake.join(d_loss_fake)
        b_hat.join(d_loss_gp)

        d_loss = d_loss_real + d_loss_fake + lambda_cls * d_loss_cls + lambda_gp * d_loss_gp
        self.d_optimizer.zero_grad()
        # flush=False: G's weight-gradient scratch is being written by stage A; d_optimizer.step() folds D's own
        with Fn.accumulate_param_grads(flush=not split):   # wgrad kernels add straight into the flat gradient buffer
            d_loss.backward()
        if split and deferred:
            held = ops.pending_detach()                     # the D step's own join must not wait for them
            ops.run_deferred(deferred)
            late = ops.pending_detach()
            ops.pending_attach(held)
            stage_a = (stage_a[0] + late[0], stage_a[1] + late[1], stage_a[2] + late[2])
        if self.parallel is not None:
            self.parallel.all_reduce_grads(self.d_optimizer)
        self.d_optimizer.step()

        # ---------------- G phase (L151-180); D's parameters are constants here
        for p in self.d_optimizer.params:
            p.requires_grad_(False)
        out_src, out_cls = self.D(x_fake)
        for p in self.d_optimizer.params:
            p.requires_grad_(True)
        g_loss_fake = Fn.MeanFn.apply(out_src, -1.0)
        g_loss_cls = Fn.CERowsFn.apply(out_cls.contiguous(), modal_trg)
        b_cyc.join(g_partial, g_loss_seg, g_loss_rec, g_loss_semi, g_loss_nce)

        if split:
            ops.pending_attach(stage_a)
            if dx_fake_cyc.is_cuda:
                dx_fake_cyc.record_stream(torch.cuda.current_stream())
            with Fn.accumulate_param_grads():
                torch.autograd.backward([g_loss_fake + lambda_cls * g_loss_cls, x_fake], [None, dx_fake_cyc])
        else:
            g_loss = g_loss_fake + lambda_cls * g_loss_cls + g_partial
            self.optimizer.zero_grad()
            with Fn.accumulate_param_grads():
                g_loss.backward()
        if self.parallel is not None:
            self.parallel.all_reduce_grads(self.optimizer)
        self.optimizer.step()

        ops.arena_end()
        return torch.stack([d_loss_real.detach(), d_loss_fake.detach(), d_loss_cls.detach(), d_loss_gp.detach(),
                            g_loss_fake.detach(), g_loss_rec.detach(), g_loss_cls.detach(), g_loss_seg.detach(),
                            g_loss_semi.detach(), g_loss_nce.detach()])

    def graphed_step(self, example, use_semi):
        """Capture train_step for the given example inputs (x_real, y_real, modal_org, modal_trg, vec_ot, vec_to,
        alpha, ids, lambda_semi[device scalar]) into a CUDA graph; returns a callable with the same arguments."""
        from ..graph import GraphedStep

        def fn(x_real, y_real, modal_org, modal_trg, vec_ot, vec_to, alpha, ids, lam):
            return self.train_step(x_real, y_real, modal_org, modal_trg, vec_ot, vec_to, alpha, [ids], lam, use_semi)
        return GraphedStep(fn, example)

    def draw(self, batch, generator=None):
        """The random draws of one iteration, made outside the (capturable) step: alpha ~ N(0,1) per sample
        (L138, sic: randn, not rand) and the 64 shared patch positions (network/ugan.py:321-322)."""
        hw = (self.input_size // 16) ** 2
        alpha = torch.randn(batch, device=self.device, generator=generator)
        ids = torch.randperm(hw, device=self.device, generator=generator)[:min(64, hw)]
        return alpha, [ids]

    def prepare_batch(self, x_real1, y_real, modal_org1, x_real2, modal_org2, mj):
        """Host-side assembly of L110-127: concatenate the labelled and unlabelled halves, build the modality
        difference vectors and move everything to the device."""
        dev = self.device
        modal_org = torch.cat([modal_org1, modal_org2], dim=0)
        modal_trg = torch.zeros_like(modal_org).fill_(mj)
        vec_org = self.label2onehot(modal_org, cfg.n_modal)
        vec_trg = self.label2onehot(modal_trg, cfg.n_modal)
        if torch.device(dev).type != 'cuda':
            x_real = torch.cat([x_real1, x_real2], dim=0)
            return (x_real.to(dev), y_real.to(dev), modal_org.to(dev), modal_trg.to(dev), (vec_trg - vec_org).to(dev),
                    (vec_org - vec_trg).to(dev))
        # asynchronous copies straight from the (pinned) loader tensors: no host-side concatenation of the slices, and
        # the few-byte modality tensors go through the caching pinned allocator so that nothing blocks the host
        b1, b2 = x_real1.shape[0], x_real2.shape[0]
        x_real = torch.empty((b1 + b2, *x_real1.shape[1:]), dtype=x_real1.dtype, device=dev)
        x_real[:b1].copy_(x_real1, non_blocking=True)
        x_real[b1:].copy_(x_real2, non_blocking=True)

        def small(t):
            return t.pin_memory().to(dev, non_blocking=True)
        return (x_real, y_real.to(dev, non_blocking=True), small(modal_org), small(modal_trg),
                small(vec_trg - vec_org), small(vec_org - vec_trg))

    # volumes whose slices the reference keeps (uganConsisTrainer.py:290): "selected visualization samples"
    pseudo_volumes = ('ct_028', 't1in_037', 't1out_015', 't2_032')

    def saving_pseudo(self, loader_type, expr_root, loader=None):
        """uganConsisTrainer.py:216-306: like BaseTrainer.saving_pseudo, plus `<name>fk.jpg` -- the slice next to its
        translation into every modality (B, 1, H, (n_modal + 1) W) -- and only for the volumes in `pseudo_volumes`
        (None: every volume).  Returns (slices seen, slices written)."""
        from PIL import Image
        self.net.eval()
        pred_root = os.path.join(expr_root, 'pseudo')
        os.makedirs(pred_root, exist_ok=True)
        loader = loader if loader is not None else self.pseudo_loader(loader_type)
        self.info(f'Predict and save in {pred_root}.')
        count = written = 0
        volumes = getattr(cfg, 'pseudo_volumes', self.pseudo_volumes)      # cfg.pseudo_volumes = None: keep every volume
        with torch.no_grad():
            for img, msk, mdl, inm in loader:
                b = img.shape[0]
                count += b
                img = img.to(self.device, non_blocking=True)
                vec_fixed_org = self.label2onehot(mdl, cfg.n_modal).to(self.device)
                x_fake_list = [img.float()]
                for vec_fixed in self.create_vectors(vec_fixed_org, cfg.n_modal):
                    _, x_fake = self.translate(img, vec_fixed - vec_fixed_org)
                    x_fake_list.append(x_fake.float())
                img_fake = torch.cat(x_fake_list, dim=3)
                out = self.segment(img)
                logits = out.permute(0, 2, 3, 1).reshape(-1, out.shape[1])
                pred = ops.argmax_c(logits if logits.is_contiguous() else logits.contiguous())
                pred = pred.view(b, *out.shape[2:]).cpu().numpy()
                img_np, msk_np = img.reshape(b, *img.shape[2:]).cpu().numpy(), msk.cpu().numpy()
                fk = ((img_fake.cpu() + 1.) / 2 * 255).numpy()
                for i in range(b):
                    mod, pid = str(inm[i]).split('_')[:2]
                    if volumes is not None and mod + '_' + pid not in volumes:
                        continue
                    written += 1
                    Image.fromarray(self.colorize(pred[i]).astype(np.uint8)).save(os.path.join(pred_root, inm[i] + 'pse.jpg'))
                    Image.fromarray(self.colorize(msk_np[i]).astype(np.uint8)).save(os.path.join(pred_root, inm[i] + 'gt.jpg'))
                    # `((a + 1) * 255).astype(uint8)` wraps above 255: the reference's own arithmetic (L283-287)
                    Image.fromarray(((img_np[i] + 1) * 255).astype(np.uint8)).save(os.path.join(pred_root, inm[i] + 'ori.jpg'))
                    Image.fromarray(fk[i, 0].astype(np.uint8)).save(os.path.join(pred_root, inm[i] + 'fk.jpg'))
        self.net.train()
        print(count)
        return count, written

    def train_epoch(self, lb_loader, ul_loader, meter, num_iter=None):
        self.net.train()
        self.D.train()
        lambda_semi = self.lambda_semi * self.sigmoid_rampup(self.epoch, cfg.max_epoch)
        n_critic = self.n_critic
        self.info(f'\nlambda_seg: {self.lambda_seg}, lambda_semi: {lambda_semi}.')

        lb_itr = iter(lb_loader)
        ul_itr = iter(ul_loader)
        tic = time.time()
        losses = None
        # fixed images of the epoch's sample grid (L82-93): the first batch of either loader, taken off the iterators
        # before the loop as the reference does (the epoch trains on the batches after them)
        x_fixed1, _, modal_fixed1, inm1 = next(lb_itr)
        x_fixed2, _, modal_fixed2, inm2 = next(ul_itr)
        if inm1 is not None and inm2 is not None:
            self.info(list(inm1) + list(inm2))
        lam_dev = torch.zeros(1, device=self.device)
        timing = self._iter_events = [] if os.environ.get('SMSUT_TIMING') and torch.cuda.is_available() else None
        for i in range(n_critic * (num_iter or cfg.num_iter_per_epoch)):
            try:
                x_real1, y_real, modal_org1, _ = next(lb_itr)
            except StopIteration:
                lb_itr = iter(lb_loader)
                x_real1, y_real, modal_org1, _ = next(lb_itr)
            try:
                x_real2, _, modal_org2, _ = next(ul_itr)
            except StopIteration:
                ul_itr = iter(ul_loader)
                x_real2, _, modal_org2, _ = next(ul_itr)

            mj = self.target_modality_rng.randint(0, cfg.n_modal - 1)
            batch = self.prepare_batch(x_real1, y_real, modal_org1, x_real2, modal_org2, mj)
            alpha, sample_ids = self.draw(batch[0].size(0))
            use_semi = self.iter >= self.semi_from_iter
            step = None
            if self.graph_enabled():
                # one captured graph per state of the consistency switch (iter < / >= semi_from_iter); the epoch's
                # lambda_semi and the learning rate are device scalars, so the graphs serve the whole run
                lam_dev.fill_(float(lambda_semi))
                inputs = [*batch, alpha, sample_ids[0], lam_dev]
                step = self.graphed(('consis', bool(use_semi)),
                                    lambda *a: self.train_step(*a[:7], [a[7]], a[8], use_semi), inputs)
            if step is not None:
                losses = step(*inputs)
            else:
                losses = self.train_step(*batch, alpha, sample_ids, lambda_semi, use_semi)
            # uganConsisTrainer.py:96,112,157-158: the segmentation loss under the labelled batch's modality, weighted
            # with cfg.batch_size
            self.meter_note(meter, losses[LOSS_KEYS.index('G_seg')], modal_org1[0].item(), cfg.batch_size)
            if timing is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                timing.append(e)

            if (i + 1) % (n_critic * self.log_step) == 0:
                vals = losses.tolist()      # the only device->host sync of the loop
                log = 'Iter: %d/%d(%d), elapsed: %.2fs,' \
                    % (i, n_critic * cfg.num_iter_per_epoch, self.iter, time.time() - tic)
                tic = time.time()
                for k, v in zip(LOSS_KEYS, vals):
                    log += ' %s: %.4f,' % (k, v)
                self.info(log)

            lr_ = self.lr_sched.host_lr(self.iter + 1)
            for opt in (self.optimizer, self.d_optimizer):
                for param_group in opt.param_groups:
                    param_group['lr'] = lr_
                opt._lr_host = lr_
            self.iter += 1
        if getattr(self, 'save_samples', False) and self.is_main:      # L205-214 (off by default: an image file per epoch)
            self.sample_translations(torch.cat([x_fixed1, x_fixed2], dim=0), torch.cat([modal_fixed1, modal_fixed2], dim=0),
                                     save_path=os.path.join(self.expr_root, self.model_idx, 'sample',
                                                                    f'train-{self.epoch + 1}-images.png'))
        self.meter_flush()
        return losses


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('-p', '--phase', type=str, default='train')
    parser.add_argument('-f', '--fold', type=int, default=0)
    parser.add_argument('-nm', '--expr_name', type=str, default=None)
    parser.add_argument('-i', '--model_id', type=str, default=None)
    parser.add_argument('-wh', '--which_ckpt', type=str, default='last')
    parser.add_argument('--epochs', type=int, default=None, help='(extension) shorten the run')
    parser.add_argument('--iters', type=int, default=None, help='(extension) iterations per epoch')
    args = parser.parse_args()

    random.seed(cfg.seed)
    np.random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)
    torch.cuda.manual_seed(cfg.seed)

    if args.phase == 'train':
        trainer = UGANConsisTrainer('train', args)
        trainer.fit('inTurn', max_epoch=args.epochs, iters_per_epoch=args.iters)
    elif args.phase == 'test':
        trainer = UGANConsisTrainer('test', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.test('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    elif args.phase == 'pseudo':
        trainer = UGANConsisTrainer('pseudo', args)
        trainer.load_model(args.model_id or '000', args.which_ckpt)
        trainer.saving_pseudo('inTurn', os.path.join(trainer.expr_root, args.model_id or '000'))
    else:
        raise NotImplementedError
