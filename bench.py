"""bench.py -- the SMSUT hot path on B200: one UGANConsisTrainer iteration (D step + G step) per "step".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[1]): uganConsisTrainer full step, 8 labelled + 8 unlabelled synthetic 1x256x256
slices per GPU, 5 classes, 4 modalities, bf16 activations / fp32 accumulation and master weights.
  value  : slices/s with the step's inputs already resident in HBM (CUDA-graph replays, CUDA events, max over ranks)
  e2e    : the same through the trainer API with HOST (pinned) batches: H2D copies of the slices / labels / modality
           vectors and a D2H read of the ten losses inside the timed region, every step
  roofline: the dominant kernel (conv_tc_kernel, tcgen05 implicit GEMM) timed alone with CUDA events on the conv
           layer classes of one generator forward at the same batch (L2 flushed between launches)
  cpu_baseline / --impl reference: the oracle (oracle/smsut_oracle.py: the reference's arithmetic in PyTorch fp32)
           on the box's host cores -- the reference is a PyTorch program, so this is its CPU path ("port" of the
           step body around the same aten ops; the reference's trainer file cannot be imported offline).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class _StdoutToStderr:
    """stdout carries exactly ONE JSON line, but libraries write to file descriptor 1 behind Python's back (NCCL prints
    "NCCL version ..." at communicator creation): point fd 1 at stderr for the whole run and give it back only for
    the result line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(text, flush=True)
        os.dup2(2, 1)


_OUT = None

FLOP_PER_SLICE = 1.0763e11      # reference-as-executed conv+GEMM FLOPs per slice
METRIC, UNIT = "train slices/sec (256x256)", "slices/s"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), p["hbm_gbs"], "measured"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference / CPU arm: the oracle on host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_rate(bs, steps, warmup, threads=None):
    import torch
    from oracle import smsut_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    G, D = O.make_weights(O.ugan_shapes(), 7), O.make_weights(O.disc_shapes(256), 8)
    g_state, d_state = {}, {}
    x1, y = O.synthetic_batch(bs, 256, 11)
    x2, _ = O.synthetic_batch(bs, 256, 12)
    x = torch.cat([x1, x2])
    modal = torch.cat([torch.full((bs,), 1), torch.full((bs,), 3)])
    gen = torch.Generator().manual_seed(0)
    times = []
    for it in range(warmup + steps):
        alpha = torch.randn(2 * bs, 1, 1, 1, generator=gen)
        ids = [torch.randperm(256, generator=gen)[:64]]
        t0 = time.perf_counter()
        O.ugan_consis_step(G, D, g_state, d_state, x, y, modal, it % 4, alpha, ids, 1e-2, 1000 + it, 0.5, nce_batch=8)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return 2 * bs / sec, sec, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = 4                              # bounded sample: 4 labelled + 4 unlabelled slices per step (~1.3 s on 16 cores)
    steps, warmup = min(args.steps, 8), min(max(args.warmup, 1), 2)
    rate, sec, threads = cpu_step_rate(bs, steps, warmup)
    sample = f"{steps} timed iterations (after {warmup} warm-up) of the full uganConsis step at {bs}+{bs} 256x256 slices"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "uganConsisTrainer full step (UGAN G/D + U-Net segmenter + PatchNCE + consistency), "
                                   "1x256x256 slices, 5 classes", "per_step_slices": 2 * bs, "device": "host CPU"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _OUT.emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
CONV_CLASSES = [  # (cin list, cout, H, ksize, count per G forward): SURVEY.md section 8(d)
    ([16], 16, 256, 3, 6), ([16, 16], 16, 256, 3, 2), ([16], 32, 128, 3, 2), ([32], 32, 128, 3, 4),
    ([32, 32], 32, 128, 3, 2), ([32], 64, 64, 3, 2), ([64], 64, 64, 3, 4), ([64, 64], 64, 64, 3, 2),
    ([64], 128, 32, 3, 2), ([128], 128, 32, 3, 4), ([128, 128], 128, 32, 3, 2), ([128], 256, 16, 3, 2),
    ([256], 256, 16, 3, 2),
]


def conv_roofline(torch, ops, batch, launches_per_class=24):
    """The implicit-GEMM conv kernels (conv_band_kernel on the W % 128 == 0 layers, conv_tc_kernel below) alone on
    the 13 3x3 conv classes of one generator forward.  Per class: `launches_per_class` launches captured in a CUDA
    graph (no host launch latency in the timing), each reading its own input set -- the sets together are larger
    than L2 (>= 160 MB, or 64 sets) so no launch finds its input cached -- timed with CUDA events around the replay
    on the replay stream.  Returns (FLOP-weighted TFLOP/s over one G forward, per-class rows, launches timed)."""
    tot_flop = tot_ms = 0.0
    per = []
    launches = 0
    stream = torch.cuda.Stream()
    for cins, cout, h, ks, count in CONV_CLASSES:
        bytes_in = batch * h * h * sum(cins) * 2
        nsets = max(2, min(64, -(-160 * 2 ** 20 // bytes_in)))
        sets = [[torch.randn(batch, h, h, c, device="cuda").to(torch.bfloat16) for c in cins] for _ in range(nsets)]
        w = torch.randn(cout, sum(cins), ks, ks, device="cuda") * 0.05
        pw = ops.PackedWeight(w)
        ops.PackTable([pw]).refresh()
        ops.conv_fprop(sets[0], pw, want_stats=True)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            with torch.cuda.graph(graph, stream=stream):
                keep = [ops.conv_fprop(sets[r % nsets], pw, want_stats=True) for r in range(launches_per_class)]
            graph.replay()                         # warm-up replay
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            graph.replay()
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / launches_per_class
        launches += launches_per_class
        del keep, graph
        flop = 2.0 * batch * h * h * cout * sum(cins) * ks * ks
        ideal_us = max(flop / 1.6304e15, (bytes_in + batch * h * h * cout * 2) / 6.4846e12) * 1e6
        per.append({"cin": sum(cins), "cout": cout, "hw": h, "k": ks, "us": ms * 1e3, "tflops": flop / ms / 1e9,
                    "gbs": (bytes_in + batch * h * h * cout * 2) / ms / 1e6, "ideal_us": ideal_us})
        tot_flop += flop * count
        tot_ms += ms * count
    return tot_flop / tot_ms / 1e9, per, launches


def run_ours(args):
    import torch
    import __graft_entry__ as g
    g.load_package()
    from smsut_b200 import _lib, ops
    from smsut_b200 import config as cfg
    from smsut_b200.data_loader import syntheticLoader as synlod
    from smsut_b200.parallel import DataParallelContext
    from smsut_b200.trainer.uganConsisTrainer import LOSS_KEYS, UGANConsisTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the SMSUT hot path has no CPU fallback")
    par = DataParallelContext()
    torch.cuda.set_device(par.local_rank)
    dev = torch.device("cuda", par.local_rank)
    random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)                 # same draw of patch ids / target modality on every rank
    tr = UGANConsisTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=256))
    if par.world > 1:
        tr.parallel = par
        par.broadcast_params(tr.optimizer, tr.d_optimizer)
    bs = cfg.batch_size
    lb = synlod.get_loader(None, 'train', 0, bs, seed=2020 + 100 * par.rank, pool_batches=4)
    ul = synlod.get_loader(None, 'val', 0, bs, seed=4040 + 100 * par.rank, pool_batches=4)
    lb_pool, ul_pool = list(lb), list(ul)

    def host_batch(i):
        x1, y, m1, _ = lb_pool[i % len(lb_pool)]
        x2, _, m2, _ = ul_pool[i % len(ul_pool)]
        return x1, y, m1, x2, m2

    gen_a = torch.Generator(device=dev).manual_seed(1234 + par.rank)      # alpha: per-rank stream
    gen_i = torch.Generator(device=dev).manual_seed(99)                   # patch ids: shared by all ranks

    def draws():
        alpha = torch.randn(2 * bs, device=dev, generator=gen_a)
        ids = torch.randperm(256, device=dev, generator=gen_i)[:64]
        return alpha, ids

    lam = torch.full((1,), 10 * tr.sigmoid_rampup(100, cfg.max_epoch), device=dev)
    batch = tr.prepare_batch(*host_batch(0), 2)
    alpha, ids = draws()
    step = tr.graphed_step([*batch, alpha, ids, lam], use_semi=True)       # 3 eager warm-up iterations + capture
    launches_per_step = step.launches_per_replay

    # ---- device-resident timing
    flush = torch.empty(160 * 2 ** 20, dtype=torch.uint8, device=dev)
    for _ in range(max(args.warmup, 3)):
        step(*batch, *draws(), lam)
    torch.cuda.synchronize()
    par.barrier()
    sampler = ClockSampler(par.local_rank) if par.rank == 0 else None
    evs = []
    for i in range(args.steps):
        a, ix = draws()
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step(*batch, a, ix, lam)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    par.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    ms = par.max_over_ranks(ms)
    clocks = sampler.stop() if sampler else None
    last_losses = out.tolist()

    # ---- end to end through the trainer API with host batches: every step copies its own pinned host batch to the
    # device and reads its ten losses back; the copy of batch i+1 is issued on a copy stream while step i runs
    h2d = d2h = 0
    copy_stream = torch.cuda.Stream()

    def stage(i):
        with torch.cuda.stream(copy_stream):
            hb = tr.prepare_batch(*host_batch(i), i % 4)           # pinned host -> device copies (non_blocking)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return hb, ev

    for i in range(2):
        hb, ev = stage(i)
        torch.cuda.current_stream().wait_event(ev)
        step(*hb, *draws(), lam).tolist()
    torch.cuda.synchronize()
    par.barrier()
    t0 = time.perf_counter()
    nxt = stage(0)
    for i in range(args.steps):
        hb, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        losses = step(*hb, *draws(), lam)
        if i + 1 < args.steps:
            nxt = stage(i + 1)
        vals = losses.tolist()                                  # D2H read of the ten losses (syncs the step)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = par.max_over_ranks(e2e_ms)
    x1, y, m1, x2, m2 = host_batch(0)
    h2d = (x1.numel() + x2.numel()) * 4 + y.numel() * 8 + 2 * (m1.numel() + m2.numel()) * 8 + 2 * 2 * bs * 4 * 4
    d2h = len(LOSS_KEYS) * 4

    n_slices = 2 * bs * par.world
    value = n_slices / ms * 1e3
    burst, sustained, hbm, how = peaks()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": par.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "uganConsisTrainer full step (UGAN G/D + U-Net segmenter + PatchNCE + consistency), "
                                   "1x256x256 slices, 5 classes", "per_gpu_slices": 2 * bs, "global_batch": n_slices,
                       "parallelism": f"dp{par.world}", "l2": "192 MiB buffer written between timed iterations",
                       "cuda_graph": True},
            "e2e": {"value": n_slices / e2e_ms * 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "losses": dict(zip(LOSS_KEYS, [round(v, 5) for v in last_losses])),
            "step_roofline": {"bound": "tensor", "achieved": value / par.world * FLOP_PER_SLICE / 1e12, "peak": sustained,
                              "unit": "TFLOP/s", "frac": value / par.world * FLOP_PER_SLICE / 1e12 / sustained,
                              "note": "reference-as-executed FLOPs per slice (1.0763e11) x slices/s/GPU vs sustained "
                                      f"bf16 GEMM peak ({how})"}}
    if par.rank == 0 and par.world == 1:
        tf, per, _ = conv_roofline(torch, ops, 2 * bs)
        traffic = None
        try:   # DRAM bytes per launch of the same launches under `ncu --set full` (profiles/r1_conv_classes_ncu.json)
            t = json.load(open(os.path.join(ROOT, "profiles", "r1_conv_classes_ncu.json")))
            traffic = t["dram_bytes_per_launch_weighted"]
        except Exception:
            pass
        algo_bytes = sum(c[4] * (2 * bs) * c[2] * c[2] * (sum(c[0]) + c[1]) * 2 for c in CONV_CLASSES) / \
            sum(c[4] for c in CONV_CLASSES)
        line["roofline"] = {"bound": "tensor", "achieved": tf, "peak": burst, "unit": "TFLOP/s", "frac": tf / burst,
                            "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes,
                            "kernel": "conv_band_kernel / conv_tc_kernel (tcgen05 implicit GEMM + fused IN statistics)",
                            "how": "FLOP-weighted over the 13 3x3 conv classes of one generator forward at 16 slices: "
                                   "24 launches per class captured in a CUDA graph, each on its own input set (sets "
                                   "together > L2), CUDA events around the replay on its stream; peak = burst "
                                   f"({how}); 8 of the 13 classes are HBM-bound (SURVEY.md section 8d): see gbs / ideal_us",
                            "per_class": per}
        rate, sec, threads = cpu_step_rate(4, 6, 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "6 timed iterations (after 1 warm-up) of the oracle's full uganConsis step at "
                                          f"4+4 256x256 slices ({sec:.2f} s each), fp32, torch CPU, {threads} threads"}
    if par.rank == 0:
        _OUT.emit(json.dumps(line))
    if par.world > 1:
        # tearing down a NCCL communicator that live CUDA graphs still reference can block at exit: everything is
        # measured and printed, so synchronise, agree that every rank is done, and leave without the teardown
        torch.cuda.synchronize()
        par.barrier()
        sys.stdout.flush()
        os._exit(0)
    par.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    _OUT = _StdoutToStderr()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
