"""bench.py -- the SMSUT hot path on B200: one UGANConsisTrainer iteration (D step + G step) per "step".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_cuda]
                  [--workload ugan_consis|mean_teacher_512|unet_infer] [--no-check]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Other workloads (BASELINE.json configs[3], configs[4]; same JSON contract, `config.workload` names them):
  mean_teacher_512 : meanTeacherTrainer iteration (student + EMA teacher U-Net, Dice/CE + softmax-MSE consistency)
                     on 8 labelled + 8 unlabelled 1x512x512 slices per GPU
  unet_infer       : U-Net segmentation (forward + argmax) of slice batches 1..256 at 256x256 on one GPU
--impl torch_cuda: context only -- the oracle's PyTorch statement of the step on the same GPU through cuDNN / cuBLAS
  (fp32, TF32, bf16-autocast + channels_last): "the existing Blackwell kernels" of SURVEY.md section 8(d).
N > 1 also runs a parity check before the timing (one iteration of the N-rank step against the oracle's single-process
  iteration on the global batch) and reports it as `parity` on the JSON line.

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
def cpu_model():
    """the host CPU's model name (SURVEY.md section 8d asks the CPU arm to name it), '' if it cannot be read"""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return ""


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


WORKLOADS = {
    "ugan_consis": "uganConsisTrainer full step (UGAN G/D + U-Net segmenter + PatchNCE + consistency), "
                   "1x256x256 slices, 5 classes",
    "mean_teacher_512": "meanTeacherTrainer step (student + EMA-teacher U-Net, Dice/CE + softmax-MSE consistency), "
                        "1x512x512 slices, 5 classes",
    "unet_infer": "U-Net segmentation (forward + argmax) of 1x256x256 slice batches 1..256",
}


def run_reference(args):
    """The reference's own CPU path for the SAME configuration as our arm: 8 labelled + 8 unlabelled slices per step
    (BASELINE.md section 4), every host thread, the requested warm-up / step counts -- each step is the bounded sample
    (~2-3 s on 16 cores).  Only if the whole run would exceed ~4 minutes are the timed steps cut (and reported)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "ugan_consis":
        _OUT.emit(json.dumps({"impl": "reference", "unavailable": f"the CPU arm times the ugan_consis workload only "
                                                                  f"(asked for {args.workload})"}))
        return
    bs = 8
    warmup = max(1, min(args.warmup, 3))
    rate1, sec1, threads = cpu_step_rate(bs, 1, 1)                   # one timed step sizes the run
    steps = max(2, min(args.steps, int(240.0 / max(sec1, 1e-3)) - warmup))
    rate, sec, threads = cpu_step_rate(bs, steps, warmup)
    sample = (f"{steps} timed iterations (after {warmup} warm-up) of the full uganConsis step at {bs}+{bs} 256x256 "
              f"slices, {sec:.2f} s each, fp32, torch CPU, {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["ugan_consis"], "per_gpu_slices": 2 * bs, "global_batch": 2 * bs,
                       "device": "host CPU"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "cpu": cpu_model()},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _OUT.emit(json.dumps(line))


def torch_cuda_rates(torch, bs=8, steps=3, warmup=2):
    """Context, not a target: the oracle's PyTorch statement of the full uganConsis step (the reference's own aten
    ops: cuDNN convolutions, native InstanceNorm, autograd double backward) on THIS GPU, in the three precisions a
    user of the reference could pick: strict fp32, TF32 (the reference's default, baseTrainer.py:40) and bf16 autocast
    with channels_last weights.  Returns {mode: slices/s}."""
    from oracle import smsut_oracle as O
    dev = "cuda"
    out = {}
    for mode in ("fp32", "tf32", "bf16_autocast_channels_last"):
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        torch.backends.cudnn.benchmark = True
        G = {k: v.to(dev) for k, v in O.make_weights(O.ugan_shapes(), 7).items()}
        D = {k: v.to(dev) for k, v in O.make_weights(O.disc_shapes(256), 8).items()}
        if mode.startswith("bf16"):
            G = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in G.items()}
            D = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in D.items()}
        x1, y = O.synthetic_batch(bs, 256, 11, device=dev)
        x2, _ = O.synthetic_batch(bs, 256, 12, device=dev)
        x = torch.cat([x1, x2])
        if mode.startswith("bf16"):
            x = x.contiguous(memory_format=torch.channels_last)
        modal = torch.cat([torch.full((bs,), 1), torch.full((bs,), 3)]).to(dev)
        gen = torch.Generator(device=dev).manual_seed(0)
        g_state, d_state = {}, {}
        try:
            evs = []
            for it in range(warmup + steps):
                alpha = torch.randn(2 * bs, 1, 1, 1, device=dev, generator=gen)
                ids = [torch.randperm(256, device=dev, generator=gen)[:64]]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16")):
                    O.ugan_consis_step(G, D, g_state, d_state, x, y, modal, it % 4, alpha, ids, 1e-2, 1000 + it, 0.5,
                                       nce_batch=8)
                e1.record()
                if it >= warmup:
                    evs.append((e0, e1))
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
            out[mode] = {"slices_per_s": 2 * bs / ms * 1e3, "ms_per_step": ms}
        except Exception as e:      # noqa: BLE001  (context arm: report, never fail the bench)
            out[mode] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        del G, D
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return out


def run_torch_cuda(args):
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        _OUT.emit(json.dumps({"impl": "torch_cuda", "unavailable": "no CUDA device"}))
        return
    rates = torch_cuda_rates(torch, steps=max(3, min(args.steps, 10)), warmup=max(2, min(args.warmup, 3)))
    best = max((v.get("slices_per_s", 0.0) for v in rates.values()), default=0.0)
    _OUT.emit(json.dumps({"impl": "torch_cuda", "metric": METRIC, "value": best, "unit": UNIT, "n_gpus": 1,
                          "higher_is_better": True, "data": "synthetic", "modes": rates,
                          "config": {"workload": WORKLOADS["ugan_consis"], "per_gpu_slices": 16,
                                     "what": "oracle (PyTorch eager: cuDNN / cuBLAS / aten kernels) on this GPU"}}))


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


def _timed_replays(torch, par, fn, steps, warmup, flush):
    """W warm-up calls, then K calls each bracketed by CUDA events on the current stream with an L2 flush (a 192 MiB
    buffer written, untimed) before it; barrier + synchronize on both sides; returns (mean ms, max over ranks)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    par.barrier()
    evs = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    par.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    return par.max_over_ranks(ms)


def _teardown(torch, par, *holders):
    """Leave without the NCCL-teardown hang of round 1: captured graphs hold references to the communicator's streams,
    so they are destroyed first (graphs, then the process group); a watchdog forces the exit if that still blocks --
    by then the result line is printed and every rank has passed the final barrier."""
    import gc
    torch.cuda.synchronize()
    par.barrier()
    sys.stdout.flush()
    if par.world > 1:
        t = threading.Timer(20.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
    for h in holders:
        if hasattr(h, "_graphs"):
            h._graphs.clear()
    del holders
    gc.collect()
    torch.cuda.synchronize()
    par.close()


def parity_check(torch, tr, par, host_batch, bs, dev):
    """ONE iteration of the N-rank step (NCCL all-reduces of both flat gradients and
    of the Dice statistics) from identical weights, each rank on its own 8 + 8 slices, against the oracle's
    single-process iteration on the global batch [lb_0 .. lb_{N-1}, ul_0 .. ul_{N-1}] with PatchNCELoss(8 N) -- what
    the reference's nn.DataParallel computes.  The oracle is the checker here, never the thing timed.  The trainer's
    state is restored afterwards.  Returns the `parity` object of the JSON line."""
    import torch.distributed as dist
    from smsut_b200.graph import StateSnapshot
    from smsut_b200.trainer.uganConsisTrainer import LOSS_KEYS
    W = par.world
    snap = StateSnapshot(tr._live_tensors())
    G0 = {k: v.detach().clone() for k, v in tr.net.state_dict().items()}
    D0 = {k: v.detach().clone() for k, v in tr.D.state_dict().items()}
    x1, y, m1, x2, m2 = host_batch(0)
    mj = 2
    gen = torch.Generator(device=dev).manual_seed(777 + par.rank)
    alpha = torch.randn(2 * bs, device=dev, generator=gen)
    ids = torch.randperm(256, device=dev, generator=torch.Generator(device=dev).manual_seed(5))[:64]
    batch = tr.prepare_batch(x1, y, m1, x2, m2, mj)
    losses = tr.train_step(*batch, alpha, [ids], 0.7, True)
    torch.cuda.synchronize()

    def gather(t):
        out = [torch.empty_like(t) for _ in range(W)]
        dist.all_gather(out, t.contiguous())
        return out
    xs, ys = gather(batch[0]), gather(batch[1])
    ms, als, ls = gather(batch[2]), gather(alpha), gather(losses)
    d_sum, g_sum = tr.d_optimizer.grad.clone(), tr.optimizer.grad.clone()       # summed over ranks by the all-reduce
    # replicas must stay bit-identical: max - min of the updated weights over ranks
    spread = []
    for flat in (tr.optimizer.flat, tr.d_optimizer.flat):
        hi, lo = flat.clone(), flat.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        spread.append(float((hi - lo).abs().max()))
    result = None
    if par.rank == 0:
        try:
            from oracle import smsut_oracle as O
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            x = torch.cat([t[:bs] for t in xs] + [t[bs:] for t in xs])
            yy = torch.cat(ys)
            modal = torch.cat([t[:bs] for t in ms] + [t[bs:] for t in ms])
            al = torch.cat([t[:bs] for t in als] + [t[bs:] for t in als]).view(-1, 1, 1, 1)
            ref, d_grads = O.ugan_d_phase(G0, D0, {}, x, modal, mj, al, [ids], 1e-2)
            Dt = {k: v.detach().clone() for k, v in tr.D.state_dict().items()}      # teacher-force the G phase
            g_ref, g_grads = O.ugan_g_phase(G0, Dt, {}, x, yy, modal, mj, [ids], 1e-2, 1000, 0.7, nce_batch=bs * W)
            ref.update(g_ref)
            ours = torch.stack(ls).mean(0).tolist()        # means, and (global Dice + local CE), both average over ranks

            def flat_of(named, grads):
                return torch.cat([grads[k].flatten().float() for k, _ in named])

            def mine(named, flat):
                off, parts = 0, []
                for _, p_ in named:
                    n = p_.numel()
                    parts.append(flat[off:off + n])
                    off += (n + 3) // 4 * 4
                return torch.cat(parts) / W

            def in_flat_order(net, opt):      # (name, parameter) in the optimizer's flat-buffer order
                names = {id(p_): k for k, p_ in net.named_parameters()}
                return [(names[id(p_)], p_) for p_ in opt.params]
            dn, gn = in_flat_order(tr.D, tr.d_optimizer), in_flat_order(tr.net, tr.optimizer)
            a_d, b_d = mine(dn, d_sum), flat_of(dn, d_grads)
            a_g, b_g = mine(gn, g_sum), flat_of(gn, g_grads)
            cos = lambda a, b: float(a @ b / (a.norm() * b.norm() + 1e-30))
            result = {"world": W, "global_batch": 2 * bs * W,
                      "losses_rel": {k: abs(v - ref[k]) / max(1.0, abs(ref[k])) for k, v in zip(LOSS_KEYS, ours)},
                      "d_grad_cosine": cos(a_d, b_d), "g_grad_cosine": cos(a_g, b_g),
                      "d_grad_norm_rel": abs(float(a_d.norm() / b_d.norm()) - 1.0),
                      "g_grad_norm_rel": abs(float(a_g.norm() / b_g.norm()) - 1.0),
                      "d_grad_checksum_rel": abs(float(a_d.sum() - b_d.sum())) / (abs(float(b_d.sum())) + 1e-30),
                      "g_grad_checksum_rel": abs(float(a_g.sum() - b_g.sum())) / (abs(float(b_g.sum())) + 1e-30),
                      "replica_weight_spread_max_abs": max(spread),
                      "against": "oracle (fp32, TF32 off) single-process iteration on the global batch, this GPU"}
            del x, yy, ref, d_grads, g_grads
        except Exception as e:      # noqa: BLE001
            result = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.empty_cache()
    snap.restore()
    from smsut_b200 import ops
    ops.param_generation[0] += 1
    par.barrier()
    return result


def run_ours(args):
    import torch
    import __graft_entry__ as g
    g.load_package()
    from smsut_b200 import config as cfg
    from smsut_b200.parallel import DataParallelContext

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the SMSUT hot path has no CPU fallback")
    par = DataParallelContext()
    torch.cuda.set_device(par.local_rank)
    random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)                 # same draw of patch ids / target modality on every rank
    if args.workload == "mean_teacher_512":
        return run_mean_teacher(args, torch, par)
    if args.workload == "unet_infer":
        return run_unet_infer(args, torch, par)
    return run_ugan_consis(args, torch, par)


def run_ugan_consis(args, torch, par):
    from smsut_b200 import _lib, ops
    from smsut_b200 import config as cfg
    from smsut_b200.data_loader import syntheticLoader as synlod
    from smsut_b200.trainer.uganConsisTrainer import LOSS_KEYS, UGANConsisTrainer
    dev = torch.device("cuda", par.local_rank)
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

    parity = None
    if par.world > 1 and not args.no_check:
        parity = parity_check(torch, tr, par, host_batch, bs, dev)

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
    flush = torch.empty(192 * 2 ** 20, dtype=torch.uint8, device=dev)
    out = [None]

    def one():
        out[0] = step(*batch, *draws(), lam)
    sampler = ClockSampler(par.local_rank) if par.rank == 0 else None
    ms = _timed_replays(torch, par, one, args.steps, max(args.warmup, 3), flush)
    clocks = sampler.stop() if sampler else None
    last_losses = out[0].tolist()

    # ---- end to end through the trainer API with host batches: every step copies its own pinned host batch to the
    # device and reads its ten losses back; the copy of batch i+1 is issued on a copy stream while step i runs
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
        losses.tolist()                                  # D2H read of the ten losses (syncs the step)
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
            "config": {"workload": WORKLOADS["ugan_consis"], "per_gpu_slices": 2 * bs, "global_batch": n_slices,
                       "parallelism": f"dp{par.world}", "l2": "192 MiB buffer written between timed iterations",
                       "cuda_graph": True},
            "e2e": {"value": n_slices / e2e_ms * 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "losses": dict(zip(LOSS_KEYS, [round(v, 5) for v in last_losses])),
            "step_roofline": {"bound": "tensor", "achieved": value / par.world * FLOP_PER_SLICE / 1e12, "peak": sustained,
                              "unit": "TFLOP/s", "frac": value / par.world * FLOP_PER_SLICE / 1e12 / sustained,
                              "note": "reference-as-executed FLOPs per slice (1.0763e11) x slices/s/GPU vs sustained "
                                      f"bf16 GEMM peak ({how})"}}
    if parity is not None:
        line["parity"] = parity
    if par.rank == 0 and par.world == 1:
        tf, per, _ = conv_roofline(torch, ops, 2 * bs)
        traffic, traffic_src = None, None
        for name in ("r2_conv_classes_ncu.json", "r1_conv_classes_ncu.json"):
            try:   # DRAM bytes per launch of the same launches under `ncu --set full` (a profiler capture cannot run
                   # inside the timed process: the committed summary of the same command is quoted, with its source)
                t = json.load(open(os.path.join(ROOT, "profiles", name)))
                traffic, traffic_src = t["dram_bytes_per_launch_weighted"], "profiles/" + name
                break
            except Exception:
                pass
        algo_bytes = sum(c[4] * (2 * bs) * c[2] * c[2] * (sum(c[0]) + c[1]) * 2 for c in CONV_CLASSES) / \
            sum(c[4] for c in CONV_CLASSES)
        line["roofline"] = {"bound": "tensor", "achieved": tf, "peak": burst, "unit": "TFLOP/s", "frac": tf / burst,
                            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": algo_bytes,
                            "kernel": "conv_band_kernel / conv_tc_kernel (tcgen05 implicit GEMM + fused IN statistics)",
                            "how": "FLOP-weighted over the 13 3x3 conv classes of one generator forward at 16 slices: "
                                   "24 launches per class captured in a CUDA graph, each on its own input set (sets "
                                   "together > L2), CUDA events around the replay on its stream; peak = burst "
                                   f"({how}); 8 of the 13 classes are HBM-bound (SURVEY.md section 8d): see gbs / ideal_us",
                            "per_class": per}
        rate, sec, threads = cpu_step_rate(bs, 4, 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "4 timed iterations (after 1 warm-up) of the oracle's full uganConsis step at "
                                          f"{bs}+{bs} 256x256 slices ({sec:.2f} s each), fp32, torch CPU, {threads} threads",
                                "cpu": cpu_model()}
        if not args.no_context:
            del step
            tr._graphs.clear()
            torch.cuda.empty_cache()
            line["torch_cuda_context"] = torch_cuda_rates(torch)
    if par.rank == 0:
        _OUT.emit(json.dumps(line))
    step = None
    _teardown(torch, par, tr)


def run_mean_teacher(args, torch, par):
    """BASELINE.json configs[3]: meanTeacherTrainer (trainer/meanTeacherTrainer.py:95-153) at 512x512, 8 labelled + 8
    unlabelled slices per GPU, consistency term on (iter >= 100), data-parallel over the ranks."""
    from smsut_b200 import config as cfg
    from smsut_b200.data_loader import syntheticLoader as synlod
    from smsut_b200.graph import GraphedStep
    from smsut_b200.trainer.meanTeacherTrainer import MeanTeacherTrainer
    dev = torch.device("cuda", par.local_rank)
    size, bs = 512, cfg.batch_size
    tr = MeanTeacherTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=size))
    if par.world > 1:
        tr.parallel = par
        par.broadcast_params(tr.optimizer)
    tr.iter = 1000
    lb = list(synlod.get_loader(None, 'train', 0, bs, size=size, seed=2020 + 100 * par.rank, pool_batches=2))
    ul = list(synlod.get_loader(None, 'val', 0, bs, size=size, seed=4040 + 100 * par.rank, pool_batches=2))

    def host_batch(i):
        return lb[i % len(lb)][0], lb[i % len(lb)][1], ul[i % len(ul)][0]

    def to_device(i):
        x1, y, x2 = host_batch(i)
        img = torch.empty((2 * bs, 1, size, size), dtype=x1.dtype, device=dev)
        img[:bs].copy_(x1, non_blocking=True)
        img[bs:].copy_(x2, non_blocking=True)
        return img, y.to(dev, non_blocking=True)

    gen = torch.Generator(device=dev).manual_seed(1234 + par.rank)
    lam = torch.full((1,), 10 * tr.sigmoid_rampup(20, tr.epoch_rampup), device=dev)
    alpha = torch.full((1,), 0.99, device=dev)

    def noise():
        return torch.clamp(torch.randn((bs, 1, size, size), device=dev, generator=gen) * 0.01, -0.02, 0.02)
    img, msk = to_device(0)
    step = GraphedStep(lambda *a: tr.train_step(*a, use_semi=True), [img, msk, noise(), lam, alpha])
    flush = torch.empty(192 * 2 ** 20, dtype=torch.uint8, device=dev)
    out = [None]

    def one():
        out[0] = step(img, msk, noise(), lam, alpha)
    sampler = ClockSampler(par.local_rank) if par.rank == 0 else None
    ms = _timed_replays(torch, par, one, args.steps, max(args.warmup, 3), flush)
    clocks = sampler.stop() if sampler else None
    # end to end: host batch -> device every step, the two losses read back
    copy_stream = torch.cuda.Stream()

    def stage(i):
        with torch.cuda.stream(copy_stream):
            hb = to_device(i)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return hb, ev
    for i in range(2):                                   # untimed: first-use allocations of the copy stream
        (im, mk), ev = stage(i)
        torch.cuda.current_stream().wait_event(ev)
        step(im, mk, noise(), lam, alpha).tolist()
    torch.cuda.synchronize()
    par.barrier()
    t0 = time.perf_counter()
    nxt = stage(0)
    for i in range(args.steps):
        (im, mk), ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        losses = step(im, mk, noise(), lam, alpha)
        if i + 1 < args.steps:
            nxt = stage(i + 1)
        losses.tolist()
    torch.cuda.synchronize()
    e2e_ms = par.max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    n_slices = 2 * bs * par.world
    value = n_slices / ms * 1e3
    burst, sustained, hbm, how = peaks()
    flop_pair = 1.8309e11            # mean-teacher 512x512 step per (labelled, unlabelled) pair
    ach = value / par.world / 2 * flop_pair / 1e12
    line = {"metric": "train slices/sec (512x512, mean teacher)", "value": value, "unit": UNIT, "n_gpus": par.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOADS["mean_teacher_512"], "per_gpu_slices": 2 * bs, "global_batch": n_slices,
                       "parallelism": f"dp{par.world}", "l2": "192 MiB buffer written between timed iterations",
                       "cuda_graph": True},
            "e2e": {"value": n_slices / e2e_ms * 1e3, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * bs * size * size * 4 + bs * size * size * 8, "d2h_bytes_per_step": 8},
            "gpu_launches": step.launches_per_replay * args.steps, "launches_per_step": step.launches_per_replay,
            "clocks": clocks, "losses": [round(v, 5) for v in out[0].tolist()],
            "step_roofline": {"bound": "tensor", "achieved": ach, "peak": sustained, "unit": "TFLOP/s",
                              "frac": ach / sustained,
                              "note": f"1.8309e11 reference-as-executed FLOP per slice pair vs sustained bf16 GEMM peak ({how})"}}
    if par.rank == 0:
        _OUT.emit(json.dumps(line))
    step = None
    _teardown(torch, par, tr)


def run_unet_infer(args, torch, par):
    """BASELINE.json configs[4]: U-Net segmentation of slice batches 1..256 at 256x256 (validate_epoch's body,
    trainer/baseTrainer.py:207-244: forward, argmax) on one GPU.  Per batch size: resident slices/s (CUDA graph of
    forward + argmax, CUDA events, L2 flushed) and end to end (pinned host images -> device, masks -> host)."""
    from smsut_b200 import ops
    from smsut_b200.data_loader import syntheticLoader as synlod
    from smsut_b200.network.unet import UNet
    dev = torch.device("cuda", par.local_rank)
    net = UNet(1, 5, 16, 'instance', 'lrelu').to(dev).eval()
    flush = torch.empty(192 * 2 ** 20, dtype=torch.uint8, device=dev)
    rows, launches = [], 0
    steps = max(5, min(args.steps, 20))
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        pool = next(iter(synlod.get_loader(None, 'test', 0, min(n, 16), seed=90 + n, pool_batches=1)))[0]
        host = pool.repeat((n + 15) // 16, 1, 1, 1)[:n].contiguous().pin_memory()
        x = host.to(dev)

        def fwd(inp):
            with torch.no_grad():
                out = net(inp)
                return ops.argmax_c(out.permute(0, 2, 3, 1).reshape(-1, 5))
        for _ in range(3):
            fwd(x)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        from smsut_b200 import _lib
        before = _lib.launch_count()
        with torch.cuda.stream(s):
            with torch.cuda.graph(graph, stream=s):
                pred = fwd(x)
        per_replay = _lib.launch_count() - before
        ms = _timed_replays(torch, par, graph.replay, steps, 3, flush)
        # end to end, as a serving loop is built: two (input, graph, mask) sets in ping-pong -- the H2D copy of batch
        # i+1 and the D2H copy of mask i-1 run on their own streams beside the forward of batch i; every batch is
        # copied in from pinned host memory and every mask is copied out to pinned host memory
        sets = [(x, graph, pred)]
        x2 = host.to(dev)
        graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            for _ in range(2):
                fwd(x2)
            s.synchronize()
            with torch.cuda.graph(graph2, stream=s):
                pred2 = fwd(x2)
        sets.append((x2, graph2, pred2))
        host_out = [torch.empty(pred.shape, dtype=pred.dtype).pin_memory() for _ in range(2)]
        h2d, d2h, main = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
        in_ready = [torch.cuda.Event() for _ in range(2)]
        computed = [torch.cuda.Event() for _ in range(2)]
        out_done = [torch.cuda.Event() for _ in range(2)]

        def pipeline(count):
            for i in range(count):
                k = i & 1
                xi, gi, pi = sets[k]
                with torch.cuda.stream(h2d):
                    h2d.wait_event(computed[k])          # the forward that last read this input buffer
                    xi.copy_(host, non_blocking=True)
                    in_ready[k].record(h2d)
                main.wait_event(in_ready[k])
                main.wait_event(out_done[k])             # the copy-out that last read this mask buffer
                gi.replay()
                computed[k].record(main)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(computed[k])
                    host_out[k].copy_(pi, non_blocking=True)
                    out_done[k].record(d2h)
            torch.cuda.synchronize()
        pipeline(4)
        t0 = time.perf_counter()
        pipeline(steps)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
        del graph2, pred2, sets
        rows.append({"batch": n, "ms": ms, "slices_per_s": n / ms * 1e3, "e2e_slices_per_s": n / e2e_ms * 1e3,
                     "launches": per_replay})
        launches += per_replay * steps
        del graph, pred
    best = max(rows, key=lambda r: r["slices_per_s"])
    best_e2e = max(rows, key=lambda r: r["e2e_slices_per_s"])
    burst, sustained, hbm, how = peaks()
    ach = best["slices_per_s"] * 6.546e9 / 1e12
    line = {"metric": "inference slices/sec (256x256 U-Net)", "value": best["slices_per_s"], "unit": UNIT, "n_gpus": 1,
            "steps": steps, "warmup": 3, "ms_per_step": best["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOADS["unet_infer"], "best_batch": best["batch"],
                       "l2": "192 MiB buffer written between timed iterations", "cuda_graph": True},
            "e2e": {"value": best_e2e["e2e_slices_per_s"], "unit": UNIT, "batch": best_e2e["batch"],
                    "h2d_bytes_per_step": best_e2e["batch"] * 256 * 256 * 4, "d2h_bytes_per_step": best_e2e["batch"] * 256 * 256 * 8},
            "gpu_launches": launches, "sweep": rows,
            "step_roofline": {"bound": "tensor", "achieved": ach, "peak": sustained, "unit": "TFLOP/s", "frac": ach / sustained,
                              "note": f"6.546e9 FLOP per slice forward (SURVEY.md section 8d) vs sustained bf16 GEMM peak ({how})"}}
    _OUT.emit(json.dumps(line))
    par.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_cuda"])
    ap.add_argument("--workload", default="ugan_consis", choices=sorted(WORKLOADS))
    ap.add_argument("--no-check", dest="no_check", action="store_true",
                    help="skip the N-rank parity check against the oracle (N > 1 only)")
    ap.add_argument("--no-context", dest="no_context", action="store_true",
                    help="skip the torch-on-CUDA context measurement of the default N = 1 run")
    a = ap.parse_args()
    _OUT = _StdoutToStderr()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch_cuda":
        run_torch_cuda(a)
    else:
        run_ours(a)
