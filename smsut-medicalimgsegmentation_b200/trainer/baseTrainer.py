# -*- coding: utf-8 -*-
"""Counterpart of the reference's trainer/baseTrainer.py restricted to what the hot path needs: construction
(device, network, Dice+CE loss: baseTrainer.py:35-62), sigmoid_rampup (:64-72), save_model (:120-123), the epoch
loop `fit` (:125-201) with its train / test meters and [TRN] / [TST] log lines, and a device-side validate_epoch
(:207-244).  The run directory gets the reference's layout (ckpt / tb / result / sample, `train.log`, TensorBoard
scalars of both meters: :81-98, :165-172, :187-193) except the copy of the working directory into `code/`.  The medpy
surface metrics are out of scope."""
import abc
import os
import random
import time
from os.path import join as pjoin

import numpy as np
import torch

from .. import config as cfg
from .. import ops
from ..data_loader import syntheticLoader as synlod
from ..misc.loss import DiceAndCrossEntropyLoss
from ..misc.utils import Meter


class BaseTrainer(object):
    def __init__(self, phase, args=None):
        self.args = args
        self.writer = None      # TensorBoard writer of the run (opened by fit)
        self._log_file = None   # <run>/train.log once the run directory exists
        if not torch.cuda.is_available() and os.environ.get("SMSUT_ALLOW_CPU_TEST_DOUBLE") != "1":
            raise RuntimeError("the SMSUT B200 trainers need a CUDA device: there is no CPU fallback")
        if self.launched_data_parallel() and torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))      # one process per GPU (torchrun)
        self.target_modality_rng = random       # shared by the replicas of a data-parallel run (setup_data_parallel)
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.phase = phase
        self.fold = 0 if args is None else getattr(args, 'fold', 0)
        if args is None or getattr(args, 'expr_name', None) is None or len(args.expr_name) == 0:
            expr_name = self.__class__.__name__
        else:
            expr_name = args.expr_name
        # the run directory <expr_root>/<NNN> is allocated like the reference's init_train_env (baseTrainer.py:74-77:
        # NNN = number of entries already under expr_root), but lazily -- at the first save / fit -- so that building a
        # trainer has no side effects on disk
        self.expr_root, self._model_idx = pjoin(cfg.expr_root, expr_name), None
        self.modality = 'all'
        self.input_size = getattr(args, 'input_size', None) or cfg.input_size
        self.net = None
        self.build_network()
        self.loss = DiceAndCrossEntropyLoss(weight_ce=cfg.weight_ce, weight_dc=cfg.weight_dc, batch_dice=True)
        self.epoch = 0
        self.iter = 0
        self._graphs = {}       # lazily captured CUDA graphs of the iteration, keyed by its host-side mode switches
        self._meter_queue = []  # (meter, device scalar, modality, slices) noted since the last meter_flush

    # ---- CUDA graph behind the epoch loops: `fit` runs the iteration the benchmark measures -------------------------
    def graph_enabled(self):
        """The epoch loops replay ONE captured CUDA graph per iteration (forward, backward, optimizer steps, LR tick:
        graph.py) -- the eager iteration pays ~1 300 host launches.  SMSUT_GRAPH=0 keeps the eager iteration."""
        return torch.device(self.device).type == 'cuda' and os.environ.get('SMSUT_GRAPH', '1') != '0'

    def _live_tensors(self):
        out = []
        for obj in vars(self).values():
            if isinstance(obj, torch.nn.Module):
                out += list(obj.state_dict().values())
            elif hasattr(obj, 'live_tensors'):
                out += obj.live_tensors()
            elif isinstance(obj, torch.Tensor):
                out.append(obj)
        return out

    def graphed(self, key, fn, inputs):
        """The captured graph of `fn(*inputs)` (device tensors only) for mode `key`, captured at first use with the
        given inputs as the example batch.  The capture's warm-up iterations run on a snapshot: weights, optimizer
        state, schedules and counters are put back afterwards, so training continues as if nothing had run.
        Returns None when the inputs do not match the captured shapes (ragged batch): run that iteration eagerly."""
        from ..graph import GraphedStep, StateSnapshot
        g = self._graphs.get(key)
        if g is None:
            snap = StateSnapshot(self._live_tensors())
            counters = (self.iter, self.epoch)
            rng = torch.cuda.get_rng_state(self.device), torch.get_rng_state()
            g = GraphedStep(fn, list(inputs))
            snap.restore()
            self.iter, self.epoch = counters
            torch.cuda.set_rng_state(rng[0], self.device)
            torch.set_rng_state(rng[1])
            ops.param_generation[0] += 1
            self._graphs[key] = g
        return g if g.accepts(inputs) else None

    # ---- data parallelism behind the trainer API: `torchrun --nproc-per-node N trainer/<x>Trainer.py -p train` ------
    @staticmethod
    def launched_data_parallel():
        return int(os.environ.get("WORLD_SIZE", "1")) > 1

    @property
    def is_main(self):
        """rank 0 (or a single process): the replica that logs, writes checkpoints, samples and TensorBoard files"""
        par = getattr(self, 'parallel', None)
        return par is None or par.rank == 0

    def setup_data_parallel(self):
        """Called by fit().  The reference wraps its networks in nn.DataParallel when it sees several GPUs
        (uganShp0Trainer.py:66-68); here every GPU has its own process (torchrun sets RANK / WORLD_SIZE /
        LOCAL_RANK): attach a parallel.DataParallelContext -- the trainers' steps all-reduce their flat gradients and
        the Dice statistics through it --, start every replica from rank 0's weights, and give the replicas different
        data: `random` / numpy / torch are re-seeded with cfg.seed + rank (sampler shuffles, augmentation draws,
        alpha, patch positions), while the target modality of an iteration comes from ONE stream shared by all
        replicas (the reference draws one per iteration for its whole batch, uganConsisTrainer.py:114)."""
        if not self.launched_data_parallel() or self.phase != 'train':
            return None
        if not hasattr(self, 'parallel'):
            raise NotImplementedError(f'{self.__class__.__name__} has no data-parallel step: launch it as one process')
        if self.parallel is None:
            from ..parallel import DataParallelContext
            self.parallel = DataParallelContext()
        par = self.parallel
        flats = [o for o in vars(self).values() if hasattr(o, 'flat') and isinstance(o.flat, torch.Tensor)]
        par.broadcast_params(*flats)                       # optimizers' master weights and EMA teachers
        ops.param_generation[0] += 1
        seed = cfg.seed + par.rank
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        self.target_modality_rng = random.Random(cfg.seed)
        return par

    # ---- the epoch meters of `fit` (baseTrainer.py:147-199, misc/utils.py:58-160) -----------------------------------
    def meter_note(self, meter, loss, modal_id, n):
        """`meter.accumulate(*meter.collect_loss_by(loss.item(), m, n))` of the reference's loops (unetTrainer.py:75-76,
        uganConsisTrainer.py:157-158, ...) without its device->host sync per iteration: the scalar stays on the device
        (as a copy: a graph replay overwrites its outputs) until meter_flush."""
        if meter is not None:
            self._meter_queue.append((meter, loss.detach().reshape(()).clone(), int(modal_id), int(n)))

    def meter_flush(self):
        """hand everything noted since the last flush to its meter: ONE device->host copy"""
        queue, self._meter_queue = self._meter_queue, []
        if queue:
            values = torch.stack([t for _, t, _, _ in queue]).tolist()
            for (meter, _, modal_id, n), v in zip(queue, values):
                meter.accumulate(*meter.collect_loss_by(v, modal_id, n))

    def meter_all_reduce(self, meter):
        """data-parallel runs: sum the epoch's weighted sums and weights over the replicas before update_cur, so the
        [TRN] line reports the mean over the global batch (what the reference's meter sees behind nn.DataParallel)"""
        par = getattr(self, 'parallel', None)
        if par is None or par.world <= 1:
            return
        keys = list(meter.configs)
        t = torch.tensor([float(meter.cur_values[k]) for k in keys] + [float(meter.n[k]) for k in keys],
                         dtype=torch.float64, device=self.device)
        par.all_reduce_sum(t)
        vals = t.tolist()
        for i, k in enumerate(keys):
            meter.cur_values[k], meter.n[k] = vals[i], vals[len(keys) + i]

    @staticmethod
    def make_meters():
        """(train meter, test meter) of baseTrainer.py:147-151"""
        min_better_keys = [f'loss_{i}' for i in range(cfg.n_modal)] + ['loss']
        max_better_keys = [f'dice_{i}' for i in range(cfg.n_modal)] + ['dice']
        return (Meter(min_better_keys=min_better_keys, max_better_keys=[], alpha=cfg.exp_alpha),
                Meter(min_better_keys=min_better_keys, max_better_keys=max_better_keys, alpha=1.))

    @property
    def model_idx(self):
        if self._model_idx is None:
            if self.phase != 'train':
                raise RuntimeError("no run id yet: pass -i <model_id> / call load_model(model_idx, which_ckpt)")
            os.makedirs(self.expr_root, exist_ok=True)
            self._model_idx = str(len(os.listdir(self.expr_root))).rjust(3, '0')
            model_root = pjoin(self.expr_root, self._model_idx)
            for d in ('ckpt', 'tb', 'result', 'sample'):                 # init_train_env, baseTrainer.py:84-90
                os.makedirs(pjoin(model_root, d), exist_ok=True)
            self._log_file = open(pjoin(model_root, 'train.log'), 'a', encoding='utf-8')
            self.info(f'Create train environment in {model_root}.')
        return self._model_idx

    def init_train_env(self, expr_root=None):
        """baseTrainer.py:81-98: allocate the run directory now (this class does it lazily, at the first save / fit, so
        that building a trainer leaves nothing on disk); returns the run id"""
        if expr_root is not None:
            self.expr_root = expr_root
        return self.model_idx

    def register_experiment_args(self, path, filename='expriments.log'):
        """baseTrainer.py:74-79: append the trainer class, its run directory and its arguments to <path>/<filename>"""
        assert self.phase == 'train'
        with open(pjoin(path, filename), 'a') as f:
            f.write(self.__class__.__name__ + ', ' + pjoin(self.expr_root, self.model_idx) + '\n')
            f.write(str(self.args) + '\n\n')

    @model_idx.setter
    def model_idx(self, value):
        if value != self._model_idx:        # another run directory: its own train.log and TensorBoard files
            self.close_run_logs()
        self._model_idx = value

    def close_run_logs(self):
        if self.writer is not None:
            self.writer.close()
            self.writer = None
        if self._log_file is not None:
            self._log_file.close()
            self._log_file = None

    @staticmethod
    def sigmoid_rampup(current, rampup_length):
        """Exponential rampup from https://arxiv.org/abs/1610.02242"""
        if rampup_length == 0:
            return 1.0
        else:
            current = np.clip(current, 0.0, rampup_length)
            phase = 1.0 - current / rampup_length
            return float(np.exp(-5.0 * phase * phase))

    def info(self, msg):
        """console + <run>/train.log (the reference's FileLogger, baseTrainer.py:94-103)"""
        if not self.is_main:
            return
        print(msg, flush=True)
        if self._log_file is not None and not self._log_file.closed:
            self._log_file.write(time.strftime('%Y-%m-%d %H:%M:%S') + f' - INFO: {msg}\n')
            self._log_file.flush()

    def open_writer(self):
        """TensorBoard scalars under <run>/tb like baseTrainer.py:92 (SMSUT_TENSORBOARD=0 switches them off)"""
        if self.writer is None and self.is_main and os.environ.get('SMSUT_TENSORBOARD', '1') != '0':
            try:
                from torch.utils.tensorboard import SummaryWriter
            except ImportError as e:          # logging only: the training path does not depend on it
                self.info(f'TensorBoard writer unavailable ({e}); scalars are not written.')
                return None
            self.writer = SummaryWriter(pjoin(self.expr_root, self.model_idx, 'tb'))
        return self.writer

    def write_scalars(self, stage, meter, epoch, lr=None):
        """`<stage>/<key>` per meter key with modality names (baseTrainer.py:165-172, 187-193)"""
        if self.writer is not None:
            for k, v in meter.cur_values.items():
                self.writer.add_scalar(f'{stage}/{Meter.display_key(k)}', v, epoch)
            if lr is not None:
                self.writer.add_scalar(f'{stage}/lr', lr, epoch)
            self.writer.flush()

    @abc.abstractmethod
    def build_network(self):
        pass

    def save_model(self, prefix):
        if not self.is_main:
            return
        path = pjoin(self.expr_root, self.model_idx, 'ckpt', f'{prefix}.ckpt')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in self.net.state_dict().items()}, path)
        self.info(f'Save model to {path}.')

    def load_model(self, model_idx=None, which_ckpt='last'):
        if model_idx is None:               # baseTrainer.py:112-117: default = this run's own checkpoint
            model_idx = self.model_idx
        path = pjoin(self.expr_root, model_idx, 'ckpt', f'{which_ckpt}.ckpt')
        self.net.load_state_dict(torch.load(path, map_location='cpu'))
        self._model_idx = model_idx
        self.info(f'Load model from {path}.')

    # ---- resume state: an extension.  The reference checkpoints weights only, so a
    # restarted run loses SGD momentum, Adam moments, the LR schedule position and the EMA teacher.
    def _stateful(self):
        """name -> object with state_dict()/load_state_dict(): every network, optimizer and schedule the trainer owns"""
        out = {}
        for name, obj in vars(self).items():
            if name in ('loss', 'parallel'):
                continue
            if hasattr(obj, 'state_dict') and hasattr(obj, 'load_state_dict'):
                out[name] = obj
        return out

    def save_state(self, prefix='resume'):
        if not self.is_main:
            return None
        path = pjoin(self.expr_root, self.model_idx, 'ckpt', f'{prefix}.state')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        state = {k: ({n: t.detach().cpu() if isinstance(t, torch.Tensor) else t for n, t in o.state_dict().items()})
                 for k, o in self._stateful().items()}
        state['__counters__'] = dict(epoch=self.epoch, iter=self.iter)
        torch.save(state, path)
        self.info(f'Save training state to {path}.')
        return path

    def load_state(self, path):
        state = torch.load(path, map_location='cpu')
        objs = self._stateful()
        missing = sorted(set(objs) - set(state))
        if missing:
            raise KeyError(f'training state {path} lacks {missing}')
        for k, o in objs.items():
            o.load_state_dict(state[k])
        ops.param_generation[0] += 1        # master weights changed outside an optimizer step: refresh the bf16 packs
        self.epoch, self.iter = state['__counters__']['epoch'], state['__counters__']['iter']

    def make_loaders(self, loader_type, phases=('train', 'val', 'test')):
        """(labelled, unlabelled, test) loaders of baseTrainer.py:125-136 (or the subset named by `phases`).  'inTurn' /
        'base' read the PNG slice tree under cfg.base_root (data_loader/inTurnLoader.py, baseLoader.py); when no dataset
        is there (this repository ships none: CHAOS / Synapse are not redistributable) the synthetic abdominal slices
        are used instead -- LOUDLY, a run on them is a smoke run, not a trained model.  'synthetic' asks for them
        explicitly."""
        if loader_type not in ('inTurn', 'base', 'synthetic'):
            raise NotImplementedError(loader_type)
        root = getattr(cfg, 'base_root', None)
        if loader_type != 'synthetic' and root and os.path.isdir(root):
            from ..data_loader import baseLoader as bslod, inTurnLoader as inlod
            lod = inlod if loader_type == 'inTurn' else bslod
            aug = getattr(cfg, 'data_aug', None)
            make = dict(train=lambda: lod.get_loader(root, 'train', self.fold, cfg.batch_size, aug, device=self.device),
                        val=lambda: lod.get_loader(root, 'val', self.fold, cfg.batch_size, aug, device=self.device),
                        test=lambda: lod.get_loader(root, 'test', 0, cfg.batch_size, device=self.device))
            return tuple(make[ph]() for ph in phases)
        if loader_type != 'synthetic':
            self.info(f'*** WARNING: no dataset under cfg.base_root = {root!r}: {self.phase}({loader_type!r}) falls back '
                      'to SYNTHETIC abdominal-like slices (data_loader/syntheticLoader.py). Checkpoints and scores of '
                      'this run are NOT those of a model trained / tested on CHAOS / Synapse data. ***')
        par = getattr(self, 'parallel', None)
        rank = par.rank if par is not None else 0
        make = dict(train=lambda: synlod.get_loader(None, 'train', self.fold, cfg.batch_size, size=self.input_size, rank=rank),
                    val=lambda: synlod.get_loader(None, 'val', self.fold, cfg.batch_size, size=self.input_size, rank=rank),
                    test=lambda: synlod.get_loader(None, 'test', 0, cfg.batch_size, size=self.input_size, pool_batches=4))
        return tuple(make[ph]() for ph in phases)

    def test(self, loader_type, expr_root, loader=None, gt_npys=None):
        """`-p test` (baseTrainer.py:254-318): segment the test split, collect the modality-organ matrices (row =
        modality, column = organ, last row / column = means) and write them to
        `<expr_root>/<modality>_trois_matrix.csv`: the Dice block, an empty line, the ASSD block.

        With the label volumes of the processed dataset (`get_label_npys` under cfg.base_root, or `gt_npys`) this is the
        reference's route: predictions assembled into host volumes, Dice by get_mo_matrix, ASSD by get_all_matrix after
        the connected-component clean-up (CPU work on the host volumes).  Without them (the synthetic fallback) the Dice
        block comes from the device-side confusion counts and there is no ASSD block.  Returns the Dice matrix."""
        from ..misc.utils import get_all_matrix, get_label_npys, get_mo_matrix
        root = getattr(cfg, 'base_root', None)
        if loader is None:
            if loader_type != 'inTurn':
                raise NotImplementedError(loader_type)
            loader = self.make_loaders(loader_type, phases=('test',))[0]
            if gt_npys is None and root and os.path.isdir(root):
                try:
                    n_gt_slic, gt_npys = get_label_npys(root, self.modality, 'test')
                except FileNotFoundError as e:      # a PNG tree without the pre-processing's <m>_<p>.npy label volumes
                    self.info(f'No label volumes ({e}): Dice from the device-side confusion counts, no ASSD block.')
        self.info(f"Predict and score the test split ({pjoin(expr_root, 'result')}).")
        fmt = lambda mat: ''.join(','.join('%.4f' % v for v in row) + '\n' for row in mat)      # noqa: E731
        if gt_npys is not None:
            n_prd_slic, prd_npys = self.validate_epoch(loader, gt_npys, None, save_path=pjoin(expr_root, 'result'))
            assert n_prd_slic == sum(v.shape[0] for v in gt_npys.values()), 'a slice of the test split was not predicted'
            matrix = get_mo_matrix(prd_npys, gt_npys)
            _, _, assd_matrix = get_all_matrix(prd_npys, gt_npys)
            log = fmt(matrix) + '\n' + fmt(assd_matrix)
        else:
            self.validate_epoch(loader)
            matrix = self.validate_dice()[1]
            log = fmt(matrix) + '\n'
        os.makedirs(expr_root, exist_ok=True)
        with open(pjoin(expr_root, f'{self.modality}_trois_matrix.csv'), 'w') as f:
            f.write(log)
        self.info(log)
        self.info('dice: %.4f' % matrix[-1, -1])
        return matrix

    def fit(self, loader_type='inTurn', max_epoch=None, iters_per_epoch=None, loaders=None):
        """loaders: optional (labelled, unlabelled, test) loaders injected by the caller instead of make_loaders.
        Under torchrun every process runs this loop on its own data (setup_data_parallel); the test stage runs on every
        replica (same test split, same weights: the Dice-statistics exchange inside the loss stays in step), rank 0
        alone logs and writes files."""
        self.setup_data_parallel()
        if self.is_main:
            self.init_train_env()           # the run directory and train.log exist before the first epoch line
        train_lb_loader, train_ul_loader, test_loader = loaders if loaders is not None else self.make_loaders(loader_type)
        self.info_loader_sizes(train_lb_loader, train_ul_loader, test_loader)
        train_meter, test_meter = self.make_meters()
        self.open_writer()
        best_epoch = -1
        n_epoch = max_epoch or cfg.max_epoch
        tic = time.time()
        for epoch in range(n_epoch):
            train_meter.reset_cur()
            self.train_epoch(train_lb_loader, train_ul_loader, train_meter, num_iter=iters_per_epoch)
            self._write_timing(time.time() - tic, iters_per_epoch or cfg.num_iter_per_epoch)
            self.epoch += 1
            tic = self.log_train_stage(train_meter, epoch, best_epoch, n_epoch, tic)
            tic = self.test_stage(test_loader, test_meter, epoch, n_epoch, tic)
            if test_meter.cur_values['dice'] >= test_meter.best_values['dice']:
                self.save_model(prefix='best')
                best_epoch = epoch
        self.save_model(prefix='last')
        self.meters = (train_meter, test_meter)
        if getattr(self, 'parallel', None) is not None:
            self.parallel.barrier()             # rank 0 has written `last` before any replica leaves

    def info_loader_sizes(self, train_lb_loader, train_ul_loader, test_loader):
        """baseTrainer.py:139-141"""
        for what, loader in (('train labeled', train_lb_loader), ('train unlabel', train_ul_loader), ('test ', test_loader)):
            dataset = getattr(loader, 'dataset', None)
            if dataset is not None:
                self.info(f'{what} images: {len(dataset)}')

    def log_train_stage(self, train_meter, epoch, best_epoch, n_epoch, tic, tag=''):
        """the train logs and scalars of an epoch (baseTrainer.py:158-172); returns the new tic"""
        self.meter_flush()
        self.meter_all_reduce(train_meter)
        train_meter.update_cur()
        opt = getattr(self, 'optimizer', None) or getattr(self, 'optimizer1')      # crossPseTrainer.py:185-188
        self.info('')
        lr = opt.param_groups[0]['lr']
        self.info(f'lr: {lr}.')
        self.info('[TRN] %sEpoch: %d(%d)/%d, elapsed: %.2fs,' % (tag, epoch, best_epoch, n_epoch, time.time() - tic)
                  + str(train_meter))
        self.write_scalars('train', train_meter, epoch, lr)
        return time.time()

    def test_stage(self, test_loader, test_meter, epoch, n_epoch, tic, tag=''):
        """the test stage of an epoch (baseTrainer.py:177-193): per-modality loss from validate_epoch, Dice from the
        modality-organ matrix -- the reference's model-selection metric; returns the new tic"""
        test_meter.reset_cur()
        self.validate_epoch(test_loader, meter=test_meter)
        v = self.validate_dice()[0]
        test_meter.accumulate(v, {k: 1. for k in v.keys()})
        test_meter.update_cur()
        self.info('[TST] %sEpoch: %d/%d, elapsed: %.2fs,' % (tag, epoch, n_epoch, time.time() - tic) + str(test_meter))
        self.write_scalars('test', test_meter, epoch)
        return time.time()

    def _write_timing(self, epoch_s, iters):
        """SMSUT_TIMING=<path>: per-epoch iteration timing as JSON (tests/test_parity_layers_gpu.py checks that the
        epoch loop runs at the benchmarked speed).  steady_ms_per_iter = device time of the last epoch's final
        iterations (after graph capture), measured by train_epoch with CUDA events when it keeps them."""
        path = os.environ.get('SMSUT_TIMING')
        if not path or not self.is_main:
            return
        import json
        ev = getattr(self, '_iter_events', None)
        steady = None
        if ev and len(ev) >= 8:
            torch.cuda.synchronize()
            tail = ev[len(ev) // 2:]
            steady = tail[0].elapsed_time(tail[-1]) / (len(tail) - 1)
        replays = sum(getattr(g, 'replays', 0) for g in self._graphs.values())
        with open(path, 'w') as f:
            json.dump(dict(epoch_s=epoch_s, iters=iters, ms_per_iter_wall=epoch_s * 1e3 / max(iters, 1),
                           steady_ms_per_iter=steady, graph_replays=replays, graphs=len(self._graphs)), f)

    # ---- `-p pseudo` (baseTrainer.py:320-378): predictions of the test split written out as images -----------------
    @staticmethod
    def colorize(img):
        """label map -> RGB with the reference's palette (baseTrainer.py:322-329)"""
        colors = [(255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0)]
        h, w = img.shape
        color_img = np.zeros((h, w, 3))
        for i in range(1, 5):
            color_img[img == i, :] = colors[i - 1][:]
        return color_img

    def pseudo_loader(self, loader_type):
        if loader_type != 'inTurn':
            raise NotImplementedError(loader_type)
        return self.make_loaders('inTurn', phases=('test',))[0]

    def saving_pseudo(self, loader_type, expr_root, loader=None):
        """baseTrainer.py:320-378: segment every slice of the test loader and save `<name>pse.jpg` (prediction),
        `<name>gt.jpg` (label), `<name>ori.jpg` (input) under <expr_root>/pseudo.  The forward + argmax run on the
        kernels; only the image files are host work.  Returns the number of slices written."""
        from PIL import Image
        self.net.eval()
        pred_root = pjoin(expr_root, 'pseudo')
        os.makedirs(pred_root, exist_ok=True)
        loader = loader if loader is not None else self.pseudo_loader(loader_type)
        self.info(f'Predict and save in {pred_root}.')
        count = 0
        with torch.no_grad():
            for img, msk, mdl, inm in loader:
                b = img.shape[0]
                count += b
                img = img.to(self.device, non_blocking=True)
                out = self.segment(img)
                logits = out.permute(0, 2, 3, 1).reshape(-1, out.shape[1])
                pred = ops.argmax_c(logits if logits.is_contiguous() else logits.contiguous())
                pred = pred.view(b, *out.shape[2:]).cpu().numpy()
                img_np, msk_np = img.reshape(b, *img.shape[2:]).cpu().numpy(), msk.cpu().numpy()
                for i in range(b):
                    self._save_pseudo_item(pred_root, inm[i], pred[i], msk_np[i], img_np[i], None)
        self.net.train()
        print(count)
        return count

    def _save_pseudo_item(self, pred_root, name, pred, msk, img, fake):
        from PIL import Image
        Image.fromarray(self.colorize(pred).astype(np.uint8)).save(pjoin(pred_root, name + 'pse.jpg'))
        Image.fromarray(self.colorize(msk).astype(np.uint8)).save(pjoin(pred_root, name + 'gt.jpg'))
        # (a + 1) * 255 on a [-1, 1] image, then mode 'F' -> 'RGB' (clips at 255): the reference's own arithmetic
        Image.fromarray(((img + 1) * 255).astype(np.float32)).convert('RGB').save(pjoin(pred_root, name + 'ori.jpg'))

    def validate_dice(self, volume_confusion=None, gt_npys=None):
        """The reference's selection metric (baseTrainer.py:246-252 -> misc/utils.py:180-203 get_mo_matrix): Dice per
        volume and organ, averaged over the volumes of a modality, then over organs / modalities; from the per-volume
        confusion counts of the last validate_epoch instead of medpy on host volumes.  medpy.metric.dc(p, g) =
        2 |p & g| / (|p| + |g|), and 0 when both are empty.  Returns (dict like the reference's, matrix).

        Called the reference's way, `validate_dice(prd_npys, gt_npys)` with host volumes, it returns the dict alone,
        computed by misc.utils.get_mo_matrix."""
        if gt_npys is not None:
            from ..misc.utils import get_mo_matrix
            mo = get_mo_matrix(volume_confusion, gt_npys)
            dices = {f'dice_{i}': mo[i, -1] for i in range(cfg.n_modal)}
            dices['dice'] = mo[-1, -1]
            return dices
        confs = self.volume_confusion if volume_confusion is None else volume_confusion
        n_modal, n_label = cfg.n_modal, cfg.n_label
        matrix = np.zeros((n_modal, n_label))
        n = np.zeros((n_modal, 1))
        for key, cv in confs.items():
            m = key.split('_')[0]
            m = int(m) if m.isdigit() else cfg.Modality[m].value
            c = cv.cpu().numpy().astype(np.float64)
            inter, size_p, size_g = np.diag(c), c.sum(0), c.sum(1)
            for i in range(n_label):
                j = i + 1
                denom = size_p[j] + size_g[j]
                matrix[m][i] += 2.0 * inter[j] / denom if denom > 0 else 0.0
            n[m] += 1
        n[n == 0] += 1e-8
        matrix /= n
        full = np.zeros((n_modal + 1, n_label + 1))
        full[:n_modal, :n_label] = matrix
        full[-1, :] = np.mean(full[0:n_modal], axis=0)
        full[:, -1] = np.mean(full[:, 0:n_label], axis=1)
        dices = {f'dice_{i}': full[i, -1] for i in range(n_modal)}
        dices['dice'] = full[-1, -1]
        return dices, full

    @abc.abstractmethod
    def train_epoch(self, lb_loader, ul_loader, meter, num_iter=None):
        pass

    def segment(self, img):
        return self.net(img)

    def validation_loss(self, out, msk, b):
        """the loss validate_epoch reports per batch: Dice + CE of the segmentation logits (baseTrainer.py:228)"""
        return self.loss(out, msk)

    def validate_epoch(self, loader, npys=None, meter=None, save_path=None):
        """Mean foreground Dice of argmax predictions over the loader (pads a ragged last batch to cfg.batch_size
        like baseTrainer.py:214-219); the confusion counts stay on the device (`confusion`, `volume_confusion`:
        validate_dice() builds the modality-organ matrix from them).

        Called the reference's way -- with `npys`, the {'<modality>_<patient>': (Z, H, W)} label volumes of
        get_label_npys -- it ALSO copies the predicted label maps into host volumes of those shapes and returns the
        reference's `(number of predicted slices, prediction volumes)` (baseTrainer.py:207-244)."""
        self.net.eval()
        n_cls = cfg.n_label + 1
        prd_npys, n_prd_slic = None, 0
        if npys is not None:
            prd_npys = {k: np.zeros(v.shape, dtype=v.dtype) for k, v in npys.items()}
        conf = torch.zeros((n_cls, n_cls), dtype=torch.int64, device=self.device)     # conf[label, prediction]
        vol_conf = {}
        with torch.no_grad():
            for img, msk, mdl, inm in loader:
                b, c, h, w = img.shape
                if b != cfg.batch_size:
                    img = torch.cat([img, torch.zeros((cfg.batch_size - b, c, h, w), dtype=img.dtype)], dim=0)
                img = img.to(self.device, non_blocking=True)
                msk = msk.to(self.device, non_blocking=True)
                out = self.segment(img)[:b]
                if meter is not None:
                    # baseTrainer.py:224-231: the batch's Dice+CE under the modality of its first slice, weighted with
                    # img.size(0) -- the PADDED batch size for a ragged last batch, as in the reference
                    self.meter_note(meter, self.validation_loss(out, msk, b), mdl[0].item(), img.size(0))
                # fp32 logits as (pixels, classes): zero-copy for the channels-last tensors the networks return
                logits = out.permute(0, 2, 3, 1).reshape(-1, n_cls)
                logits = logits if logits.is_contiguous() else logits.contiguous()
                ops.confusion_counts(logits, msk.reshape(-1), conf)
                if prd_npys is not None:
                    pred = ops.argmax_c(logits).view(b, h, w).cpu().numpy()
                    for i in range(b):
                        m_name, pid, z = str(inm[i]).split('_')
                        prd_npys[f'{m_name}_{pid}'][int(z)] = pred[i]
                        n_prd_slic += 1
                # per-volume counts for the modality-organ matrix: slices are named '<modality>_<patient>_<z>'
                # (baseTrainer.py:241); consecutive slices of one volume share one accumulator
                if inm is None:
                    continue                # unnamed slices: only the global counts
                i0 = 0
                keys = ['_'.join(str(nm).split('_')[:2]) for nm in inm][:b]
                for i in range(1, b + 1):
                    if i == b or keys[i] != keys[i0]:
                        cv = vol_conf.get(keys[i0])
                        if cv is None:
                            cv = vol_conf[keys[i0]] = torch.zeros((n_cls, n_cls), dtype=torch.int64, device=self.device)
                        ops.confusion_counts(logits[i0 * h * w:i * h * w], msk[i0:i].reshape(-1), cv)
                        i0 = i
        self.meter_flush()
        self.confusion = conf
        self.volume_confusion = vol_conf
        inter = conf.diagonal().double()
        denom = (conf.sum(0) + conf.sum(1)).double()
        dice = (2 * inter[1:] / denom[1:].clamp_min(1)).mean().item()
        self.net.train()
        return dice if prd_npys is None else (n_prd_slic, prd_npys)
