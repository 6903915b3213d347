"""Smallest GPU exercise of the epoch bookkeeping around the captured iteration: UGANConsisTrainer.fit for one short
epoch on synthetic 64x64 slices (meter notes on graph-replay outputs, validation loss under no_grad, [TRN] / [TST]
lines, best / last checkpoints), then BaseTrainer.test.  Writes its progress to gpurun_out/meter_probe.log."""
import os
import sys
import tempfile
import time
from types import SimpleNamespace

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "meter_probe.log"), "w")


def say(msg):
    LOG.write("%.1fs %s\n" % (time.time() - t0, msg))
    LOG.flush()
    print(msg, flush=True)


os.environ.setdefault("SMSUT_TENSORBOARD", "0")
import torch                                              # noqa: E402
import __graft_entry__ as g                               # noqa: E402
g.load_package()
from smsut_b200 import config as cfg                      # noqa: E402
from smsut_b200.trainer.uganConsisTrainer import UGANConsisTrainer      # noqa: E402
say("imports done")
cfg.expr_root = tempfile.mkdtemp()
tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name="probe", input_size=64))
say("trainer built")
tr.fit('synthetic', max_epoch=1, iters_per_epoch=int(os.environ.get("PROBE_ITERS", "4")))
torch.cuda.synchronize()
train_meter, test_meter = tr.meters
say("fit done: train %s | test %s" % (train_meter, test_meter))
assert train_meter.cur_values["loss"] > 0 and test_meter.cur_values["loss"] > 0
assert all(torch.isfinite(torch.tensor(list(m.cur_values.values()))).all() for m in tr.meters)
replays = sum(getattr(gr, "replays", 0) for gr in tr._graphs.values())
say("graph replays: %d" % replays)
tr.phase = 'test'
tr.test('synthetic' if False else 'inTurn', os.path.join(tr.expr_root, tr.model_idx))
say("meter probe ok")
