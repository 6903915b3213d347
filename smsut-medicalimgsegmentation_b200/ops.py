"""Tensor-level wrappers over the C ABI (include/smsut_b200.h).  No autograd here, no fallbacks.

Conventions: activations are contiguous bf16 torch tensors of shape (N, H, W, C) (NHWC) living on the
current CUDA device; every call launches on torch's current stream, so the wrappers are safe inside CUDA-graph
capture and on the autograd engine thread.  Shapes are validated in Python, the library validates the rest and
reports through smsut_last_error().
"""
import ctypes as C
import os
import weakref

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, TC_CONV, TC_CONVT_DGRAD, TC_CONVT_FWD, ConvDirectArgs,
                   ConvTcArgs, PackEntry, UnpackEntry, WgradTcArgs, call)

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _chk(t, dtype, name):
    if t.device.type != "cuda":
        raise _lib.SmsutError(f"{name}: SMSUT kernels need CUDA tensors (there is no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


def pad16(c):
    return (c + 15) // 16 * 16


# bumped by the fused optimisers (they write parameters through raw pointers, which does not advance
# torch's per-tensor version counters); PackTable repacks when it or any tensor version changes
param_generation = [0]


# ----------------------------------------------------------------------------------------------
# zero-initialised scratch (statistics, reduction buffers, loss accumulators)
# ----------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------
# deterministic mode (SMSUT_DETERMINISTIC=1, csrc/det.cu): every accumulator the kernels reduce into with atomics gets
# a registered 64-bit fixed-point shadow; integer atomics are order-independent, `resolve` folds the shadow into the
# fp32 tensor before its first reader.  Two runs of a step -- eager or graph replay, any stream schedule -- are then
# bit-identical, which turns "is it a race or is it atomics noise?" into an equality test.
# ----------------------------------------------------------------------------------------------
DET = [os.environ.get("SMSUT_DETERMINISTIC", "0") == "1"]


def set_deterministic(on=True):
    """Switch deterministic accumulation on / off for accumulators allocated FROM NOW ON (arena, optimizers built
    afterwards); call it before constructing trainers / optimizers."""
    DET[0] = bool(on)
    if _Arena.buf is not None:      # rebuilt (with / without its shadow) at the next arena_begin
        if _Arena.shadow is not None:
            call("smsut_det_unregister", _p(_Arena.buf))
        _Arena.buf = _Arena.shadow = None


def det_register(t):
    """Allocate, register and return the fixed-point shadow of the fp32 accumulator tensor `t` (contiguous)."""
    shadow = torch.zeros(t.numel(), dtype=torch.int64, device=t.device)
    call("smsut_det_register", _p(t), t.numel() * 4, _p(shadow))
    # the registry holds raw addresses: drop the entry when the accumulator dies (its memory may be handed to an
    # unrelated tensor); the finalizer keeps the shadow alive exactly as long as the accumulator
    weakref.finalize(t, _det_release, t.data_ptr(), shadow)
    return shadow


def det_unregister(ptr):
    _lib.lib.smsut_det_unregister(C.c_void_p(ptr))


def resolve(t):
    """Fold the pending fixed-point partial sums of accumulator `t` into it (no-op outside deterministic mode and
    for unregistered tensors).  Launched on the current stream, right after the kernels that accumulate into `t`."""
    if DET[0] and t is not None and t.numel() > 0:
        call("smsut_det_resolve", _p(t), t.numel(), _stream())
    return t


class _Arena:
    """One buffer cleared by ONE memset at the start of a training iteration; the hundreds of small zeroed
    accumulators of the iteration are slices of it (each would otherwise be its own fill launch)."""
    buf = None
    shadow = None      # deterministic mode: int64 fixed-point shadow of `buf` (one value per fp32 slot)
    off = 0
    active = False


def arena_begin(device, nbytes=48 << 20):
    if _Arena.buf is None or _Arena.buf.device != torch.device(device) or _Arena.buf.numel() != nbytes:
        if _Arena.shadow is not None:
            call("smsut_det_unregister", _p(_Arena.buf))
            _Arena.shadow = None
        _Arena.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if DET[0] and _Arena.buf.is_cuda:
            _Arena.shadow = torch.zeros(nbytes // 4, dtype=torch.int64, device=device)
            call("smsut_det_register", _p(_Arena.buf), nbytes, _p(_Arena.shadow))
    _Arena.buf.zero_()
    if _Arena.shadow is not None:
        _Arena.shadow.zero_()
    _Arena.off = 0
    _Arena.active = True
    del _branch_used[:]       # a new iteration: branch streams are (re)forked from here on


def arena_end():
    _Arena.active = False


def zeros(shape, device, dtype=F32):
    if isinstance(shape, int):
        shape = (shape,)
    n = 1
    for d in shape:
        n *= d
    nbytes = n * (4 if dtype == F32 else 2)
    if _Arena.active and _Arena.buf.device == torch.device(device) and _Arena.off + nbytes <= _Arena.buf.numel():
        # a tensor over the arena's storage that is NOT an autograd view of it (own version counter): it can be
        # saved for backward and returned from Functions while other slices are written in place
        item = 4 if dtype == F32 else 2
        strides, acc = [], 1
        for d in reversed(shape):
            strides.append(acc)
            acc *= d
        t = torch.empty(0, dtype=dtype, device=_Arena.buf.device)
        t.set_(_Arena.buf.untyped_storage(), _Arena.off // item, tuple(shape), tuple(reversed(strides)))
        _Arena.off += (nbytes + 255) // 256 * 256
        return t
    t = torch.zeros(shape, dtype=dtype, device=device)
    if DET[0] and dtype == F32 and t.is_cuda and t.numel() > 0:
        det_register(t)      # outside an iteration's arena: a shadow of its own, unregistered when the tensor dies
    return t


def _det_release(ptr, shadow):
    det_unregister(ptr)
    del shadow


# ----------------------------------------------------------------------------------------------
# side streams: weight-gradient kernels run beside the dgrad chain
# ----------------------------------------------------------------------------------------------
class _Side:
    """The weight gradients of a backward pass are leaves of the dependency graph: nothing reads them before the
    optimizer step.  Inside `side_streams()` they are launched on side streams forked from the current stream
    (captured as parallel branches of the step's CUDA graph), so they fill the SMs the narrow dgrad / norm kernels
    of the critical path leave idle.  Their operands are held until the join, so the caching allocator cannot hand
    the memory to a later kernel of the main stream while a side kernel still reads it."""
    # side streams per pool.  Measured on B200 (ms per 16-slice step, round-2 final build): 1 -> 10.65, 2 -> 9.91,
    # 3 -> 9.88, 4 -> 9.79 (twice, on two boxes), 6 -> 9.85, 8 -> 9.83.  The default stays 2: with 4, one 4-GPU bench run
    # (of ~15 runs on 1 / 2 / 4 / 8 GPUs with that setting) aborted with an unexplained `unspecified launch failure`
    # during the timed replays and there was no GPU budget left to chase it; 2 is the setting every recorded
    # measurement, test run and multi-GPU line of the round was taken with before that.  SMSUT_SIDE_STREAMS=4 opts in.
    n_streams = int(os.environ.get("SMSUT_SIDE_STREAMS", "2"))
    streams = {}        # device index -> [streams]
    active = False
    group = 0
    nxt = 0
    held = []
    used = []
    defer = None        # list of (fn, keep, event) while weight-gradient launches are being deferred (defer_begin)


def side_streams_enable(on=True, group=0):
    """group: which pool of side streams the following weight-gradient kernels use.  Two backward passes that are in
    flight at the same time (the generator's cycle-pass backward beside the discriminator phase) use different pools,
    so that joining one does not wait for the other's queue."""
    _Side.active = bool(on) and _Side.n_streams > 0
    _Side.group = group


def pending_detach():
    """Hand the bookkeeping of everything forked so far (side streams with weight-gradient kernels in flight, the
    tensors they read, the branch streams used) to the caller and start afresh: the next join then only waits for
    work issued after this call.  `pending_attach` gives it back before the join that must cover it."""
    st = (_Side.used, _Side.held, list(_branch_used))
    _Side.used, _Side.held = [], []
    del _branch_used[:]
    return st


def pending_attach(st):
    for s in st[0]:
        if s not in _Side.used:
            _Side.used.append(s)
    _Side.held.extend(st[1])
    for b in st[2]:
        if b not in _branch_used:
            _branch_used.append(b)


def defer_begin():
    """From now on side_run() only QUEUES its launches (with an event that marks where their operands are ready); they
    are issued by run_deferred().  Weight gradients are leaves of the dependency graph, so WHEN they run is free: the
    trainer moves those of the generator's early backward out of the window it shares with the discriminator phase
    (every SM busy) into the tail of the iteration (one latency-bound chain, SMs idle)."""
    if _Side.active:
        _Side.defer = []


def defer_end():
    """stop deferring; returns the queued launches for run_deferred()"""
    items, _Side.defer = _Side.defer, None
    return items or []


def run_deferred(items, group=2):
    """Issue the queued weight-gradient launches on side streams ordered after the CURRENT stream (so they start no
    earlier than this point of the iteration) and after their own operands; the next side_join() waits for them."""
    if not items:
        return
    main = torch.cuda.current_stream()
    pool = _Side.streams.get((main.device_index, group))
    if pool is None:
        pool = [torch.cuda.Stream(device=main.device, priority=0) for _ in range(max(_Side.n_streams, 1))]
        _Side.streams[(main.device_index, group)] = pool
    started = set()
    for i, (fn, keep, ev) in enumerate(items):
        s = pool[i % len(pool)]
        if i < len(pool) and id(s) not in started:
            s.wait_stream(main)
            started.add(id(s))
        s.wait_event(ev)
        with torch.cuda.stream(s):
            fn()
        if s not in _Side.used:
            _Side.used.append(s)
        _Side.held.append(keep)


def side_run(fn, keep):
    """fn() on a side stream ordered after everything issued so far on the current stream (or inline when side
    streams are off).  `keep`: the tensors fn's kernels read."""
    if not _Side.active:
        fn()
        return
    main = torch.cuda.current_stream()
    if _Side.defer is not None:
        ev = torch.cuda.Event()
        ev.record(main)
        _Side.defer.append((fn, keep, ev))
        return
    pool = _Side.streams.get((main.device_index, _Side.group))
    if pool is None:
        # lowest priority: the dgrad / norm chain on the forking streams is the critical path
        pool = [torch.cuda.Stream(device=main.device, priority=0) for _ in range(_Side.n_streams)]
        _Side.streams[(main.device_index, _Side.group)] = pool
    s = pool[_Side.nxt % len(pool)]
    _Side.nxt += 1
    s.wait_stream(main)
    with torch.cuda.stream(s):
        fn()
    if s not in _Side.used:
        _Side.used.append(s)
    _Side.held.append(keep)


_branch_streams = {}
_BRANCH_PRIORITY = int(os.environ.get("SMSUT_BRANCH_PRIORITY", "-1"))   # above the wgrad side streams (0 = lowest)
# Branches 3 / 4 carry the generator's cycle pass and its early backward ("stage A" of UGANConsisTrainer.train_step).
# Kernel nodes of a captured graph keep their stream's priority; running stage A BELOW the discriminator phase
# (SMSUT_STAGE_A_PRIORITY=0) measured slower (10.46 vs 10.30 ms / step): its kernels are what fills the SMs the
# discriminator's small kernels leave idle, and the final backward waits for it anyway.
_BRANCH_PRIORITIES = {3: int(os.environ.get("SMSUT_STAGE_A_PRIORITY", "-1")),
                      4: int(os.environ.get("SMSUT_STAGE_A_PRIORITY", "-1")),
                      # branches 1 / 2: D(G(x)) and the interpolated pass with the gradient penalty's double backward --
                      # the longest dependent chain of the iteration (140 small kernels)
                      1: int(os.environ.get("SMSUT_D_PRIORITY", "-1")),
                      2: int(os.environ.get("SMSUT_D_PRIORITY", "-1"))}
_branch_used = []         # branch streams forked since the iteration began (arena_begin)
_branch_stack = []        # ids of the parallel_branch blocks the calling thread is currently inside (forward only)
branch_parallel = [os.environ.get("SMSUT_BRANCH_STREAMS", "1") != "0"]


def current_branch():
    """id of the innermost active parallel_branch block, or None on the forking (main) stream"""
    return _branch_stack[-1] if _branch_stack else None


class parallel_branch:
    """`with parallel_branch(k) as br:` runs the enclosed launches on branch stream k, forked from the current stream;
    `br.join(*outputs)` (after the block) makes the current stream wait for them.  Used for data-independent chains
    (the translation / segmentation halves of the generator, the three discriminator passes): most of their kernels
    are far smaller than the machine, so two chains side by side cost little more than one.  Autograd replays each
    node's backward on its forward stream, so the backward chains overlap as well.  Capturable: the branch joins the
    capture at the fork and is joined back before the step ends."""

    def __init__(self, k=0):
        self.k = k
        self.on = branch_parallel[0] and torch.cuda.is_available()

    def __enter__(self):
        _branch_stack.append(self.k)
        if not self.on:
            return self
        self.main = torch.cuda.current_stream()
        key = (self.main.device_index, self.k)
        st = _branch_streams.get(key)
        if st is None:
            st = _branch_streams[key] = torch.cuda.Stream(device=self.main.device,
                                                          priority=_BRANCH_PRIORITIES.get(self.k, _BRANCH_PRIORITY))
        self.stream = st
        if st not in _branch_used:
            _branch_used.append(st)
        st.wait_stream(self.main)
        self.ctx = torch.cuda.stream(st)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        _branch_stack.pop()
        if self.on:
            self.ctx.__exit__(*a)

    def join(self, *outputs):
        """call on the forking stream after the block: orders it after the branch and tells the caching allocator
        that `outputs` (allocated on the branch stream) are consumed here"""
        if not self.on:
            return
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.stream)
        for t in outputs:
            if isinstance(t, torch.Tensor):
                t.record_stream(cur)


def branch_join_all():
    """the current stream waits for every branch stream of its device (end of a backward pass: autograd only orders
    streams where a gradient tensor crosses them, but parameter gradients are also written in place -- InstanceNorm
    gamma / beta atomics, the wgrad scratch -- by kernels whose Functions hand autograd `None`)"""
    if not _branch_used:
        return
    cur = torch.cuda.current_stream()
    for st in _branch_used:     # only the streams forked in this iteration (others are not part of a capture)
        if st.device_index == cur.device_index and st != cur:
            cur.wait_stream(st)


def side_join():
    """the current stream waits for every side kernel issued since the last join"""
    if _Side.used:
        main = torch.cuda.current_stream()
        for s in _Side.used:
            main.wait_stream(s)
    _Side.used = []
    _Side.held = []


# ----------------------------------------------------------------------------------------------
# weight packing
# ----------------------------------------------------------------------------------------------
class PackedWeight:
    """bf16 GEMM-ready copies of one fp32 master weight (see smsut_pack_entry)."""

    def __init__(self, weight, transposed=False, need_dgrad=True):
        self.weight = weight
        self.transposed = transposed
        if transposed:
            cin, cout, kh, kw = weight.shape
            assert kh == 2 and kw == 2 and cin % 16 == 0 and cout % 16 == 0
            self.cin_pad, self.cout_pad = cin, cout
        else:
            cout, cin, kh, kw = weight.shape
            self.cin_pad, self.cout_pad = pad16(cin), pad16(cout)
        self.cin, self.cout, self.kh, self.kw = cin, cout, kh, kw
        # weight-gradient scratch geometry (taps, rows, cols) of the tensor-core wgrad kernels: see WgradScratch
        weight._smsut_tc = (kh * kw, cin, cout) if transposed else (kh * kw, cout, cin)
        n = self.cin_pad * self.cout_pad * kh * kw
        dev = weight.device
        self.fprop = torch.empty(n, dtype=BF16, device=dev)
        self.dgrad = torch.empty(n, dtype=BF16, device=dev) if need_dgrad else None
        self.version = None
        self.ptr = 0

    def stale(self):
        w = self.weight
        return self.version != (w._version, param_generation[0]) or self.ptr != w.data_ptr()

    def entry(self):
        e = PackEntry()
        e.w = self.weight.data_ptr()
        e.fprop = self.fprop.data_ptr()
        e.dgrad = self.dgrad.data_ptr() if self.dgrad is not None else 0
        e.cout, e.cin, e.kh, e.kw = self.cout, self.cin, self.kh, self.kw
        e.transposed = 1 if self.transposed else 0
        e.cout_pad, e.cin_pad = self.cout_pad, self.cin_pad
        return e


class PackTable:
    """Device table of pack entries: one launch refreshes every bf16 weight copy of a network."""

    def __init__(self, packs):
        self.packs = list(packs)
        self._table = None
        self._ptrs = None

    def _build(self):
        arr = (PackEntry * len(self.packs))(*[p.entry() for p in self.packs])
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._table = raw.to(self.packs[0].weight.device)
        self._ptrs = [p.weight.data_ptr() for p in self.packs]

    def refresh(self, force=False):
        """Repack if any master weight changed (tensor version counter) or moved."""
        if not self.packs:
            return
        ptrs = [p.weight.data_ptr() for p in self.packs]
        if self._table is None or ptrs != self._ptrs:
            self._build()
            force = True
        if not force and not any(p.stale() for p in self.packs):
            return
        call("smsut_pack_weights", _p(self._table), len(self.packs), _stream())
        gen = param_generation[0]
        for p in self.packs:
            p.version = (p.weight._version, gen)
            p.ptr = p.weight.data_ptr()


class WgradScratch:
    """Tap-major fp32 scratch [tap][rows][cols] beside a flat gradient buffer (optim._FlatOptimizer).

    In the OIHW gradient the 16 accumulator columns a wgrad thread owns are 36 bytes apart and the lanes of a warp
    another 36 * Cin bytes: every fp32 atomic is its own L2 transaction, and those atomics were 43 % of the wgrad
    time (57 % on the 16x16 / 32x32 levels).  In the tap-major scratch the same 16 columns are 64 contiguous bytes
    (four red.global.add.v4.f32) and the band kernel's lanes are contiguous channels.  `flush()` -- one
    smsut_unpack_wgrads launch per network -- folds the scratch into the OIHW gradient and clears it."""
    owners = weakref.WeakSet()

    def __init__(self, params, grad_views, grad_flat=None, n_early=0, early_offset=None):
        """params: the network's parameters; grad_views[i]: the flat-gradient view of params[i].  The scratch is built
        on first use: the conv modules create their PackedWeight (which marks a weight as tensor-core) lazily.
        grad_flat: the flat gradient buffer the views live in (deterministic mode: its shadow is folded in by flush)."""
        self.params, self.grads = list(params), list(grad_views)
        self.grad_flat = grad_flat
        # the last n_early parameters form the "early bucket" (optim._FlatOptimizer): flush(part=...) can complete the
        # two parts of the gradient separately
        self.n_early, self.early_offset = n_early, early_offset
        self.n_late_entries, self.scratch_split = 0, 0
        self.dirty = False
        self.flat = None
        self.shadow = None
        self.views = {}
        for p in self.params:
            p._smsut_scratch_owner = weakref.ref(self)
        WgradScratch.owners.add(self)

    def _build(self):
        self._flush_scratch()     # NOT flush(): folding the gradient's fixed-point shadow mid-backward would split a sum
        items = [(p, g) for p, g in zip(self.params, self.grads) if getattr(p, "_smsut_tc", None) is not None]
        self.views = {}
        if not items:
            return
        dev = items[0][0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p, _ in items]
        self.flat = torch.zeros(sum(sizes), dtype=F32, device=dev)
        self.shadow = det_register(self.flat) if DET[0] and self.flat.is_cuda else None
        entries, off = [], 0
        early_ids = {id(p) for p in self.params[len(self.params) - self.n_early:]} if self.n_early else set()
        self.n_late_entries = sum(1 for p, _ in items if id(p) not in early_ids)
        self.scratch_split = sum(sz for (p, _), sz in zip(items, sizes) if id(p) not in early_ids)
        for (p, g), sz in zip(items, sizes):
            taps, rows, cols = p._smsut_tc
            view = self.flat[off:off + p.numel()].view(taps, rows, cols)
            self.views[id(p)] = (view, g.data_ptr())
            e = UnpackEntry()
            e.scratch, e.grad, e.rows, e.cols, e.taps = view.data_ptr(), g.data_ptr(), rows, cols, taps
            entries.append(e)
            off += sz
        arr = (UnpackEntry * len(entries))(*entries)
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.n = len(entries)
        # the zero fill above ran on whatever stream the first weight-gradient call came from; wgrad kernels of other
        # branch streams must not overtake it (first iteration only, never during capture: the warm-up builds it)
        if dev.type == "cuda" and not torch.cuda.is_current_stream_capturing():
            torch.cuda.synchronize(dev)

    def view_for(self, weight, out):
        """the scratch view of `weight` if `out` is its flat-gradient view, else None"""
        ent = self.views.get(id(weight))
        if ent is None:
            if getattr(weight, "_smsut_tc", None) is None or not any(p is weight for p in self.params):
                return None
            self._build()
            ent = self.views.get(id(weight))
            if ent is None:
                return None
        if out.data_ptr() != ent[1]:
            return None
        self.dirty = True
        return ent[0]

    def discard(self):
        """drop pending weight gradients (optimizer.zero_grad)"""
        if self.flat is not None and self.dirty:
            self.flat.zero_()
            if self.shadow is not None:
                self.shadow.zero_()
        self.dirty = False

    def flush(self, part='all'):
        """part: 'all', or 'early' / 'late' = only the early bucket's parameters / only the others"""
        if self.grad_flat is not None:       # deterministic mode: atomically accumulated parameter gradients
            if part == 'all' or not self.n_early:
                resolve(self.grad_flat)
            elif part == 'early':
                resolve(self.grad_flat[self.early_offset:])
            elif self.early_offset > 0:
                resolve(self.grad_flat[:self.early_offset])
        self._flush_scratch(part if self.n_early else 'all')

    def _flush_scratch(self, part='all'):
        if self.flat is None or not self.dirty:
            return
        entry = ctypes_sizeof_unpack_entry()
        if part == 'all':
            first, count, lo, hi = 0, self.n, 0, self.flat.numel()
        elif part == 'late':
            first, count, lo, hi = 0, self.n_late_entries, 0, self.scratch_split
        else:
            first, count, lo, hi = self.n_late_entries, self.n - self.n_late_entries, self.scratch_split, self.flat.numel()
        if count > 0:
            resolve(self.flat[lo:hi])
            call("smsut_unpack_wgrads", C.c_void_p(self.table.data_ptr() + first * entry), count, _stream())
            self.flat[lo:hi].zero_()
        if part != 'early':
            self.dirty = False       # 'early' leaves the late part pending

    @staticmethod
    def flush_all():
        for o in list(WgradScratch.owners):
            o.flush()


def ctypes_sizeof_unpack_entry():
    return C.sizeof(UnpackEntry)


def _wgrad_scratch(weight, out):
    """the tap-major scratch to accumulate into when `out` is the weight's flat-gradient view, else None"""
    ref = getattr(weight, "_smsut_scratch_owner", None)
    owner = ref() if ref is not None else None
    if owner is None or out is None:
        return None
    return owner.view_for(weight, out)


# ----------------------------------------------------------------------------------------------
# tensor-core convolutions
# ----------------------------------------------------------------------------------------------
def _conv_args(kind, ksize, n, h, w, srcs, wpack, ncols, ncols_pad, out0, out0_ld, out0_coff=0, out1=None, out1_ld=0,
               out1_coff=0, split=0, bias=None, act=ACT_NONE, slope=0.01, accumulate=False, out_f32=False, bn=0):
    a = ConvTcArgs()
    a.kind, a.ksize, a.n, a.h, a.w, a.nsrc = kind, ksize, n, h, w, len(srcs)
    for i, (t, c, ld) in enumerate(srcs):
        _chk(t, BF16, "conv_tc src")
        a.src[i] = t.data_ptr()
        a.src_c[i] = c
        a.src_ld[i] = ld
    a.wpack = wpack.data_ptr()
    a.ncols, a.ncols_pad = ncols, ncols_pad
    a.out0, a.out0_ld, a.out0_coff = out0.data_ptr(), out0_ld, out0_coff
    if out1 is not None:
        a.out1, a.out1_ld, a.out1_coff, a.split = out1.data_ptr(), out1_ld, out1_coff, split
    a.bias = bias.data_ptr() if bias is not None else 0
    a.act, a.slope, a.accumulate, a.out_f32, a.bn = act, slope, int(accumulate), int(out_f32), bn
    a.stats = 0
    return a


def conv_tc(*args, **kw):
    """srcs: list of (tensor, channels_used, channel_pitch)."""
    call("smsut_conv_tc", C.byref(_conv_args(*args, **kw)), _stream())


def conv_fprop(xs, pw, bias=None, act=ACT_NONE, out_f32=False, want_stats=False):
    """y = conv(cat(xs, channel), W) for a 1x1 / 3x3 / 5x5 stride-1 'same' conv.  xs: NHWC bf16 tensors.
    want_stats: also return the InstanceNorm statistics (n, 2, C) of y -- fused into the conv epilogue where the
    library supports it (smsut_conv_tc_fuses_stats), otherwise by the statistics kernel."""
    n, h, w, _ = xs[0].shape
    srcs = [(x, x.shape[3], x.shape[3]) for x in xs]
    assert sum(s[1] for s in srcs) == pw.cin_pad, (sum(s[1] for s in srcs), pw.cin_pad)
    y = torch.empty((n, h, w, pw.cout_pad), dtype=F32 if out_f32 else BF16, device=xs[0].device)
    a = _conv_args(TC_CONV, pw.kh, n, h, w, srcs, pw.fprop, pw.cout_pad, pw.cout_pad, y, pw.cout_pad, bias=bias, act=act,
                   out_f32=out_f32)
    stats = None
    if want_stats and _lib.lib.smsut_conv_tc_fuses_stats(C.byref(a)) == 1:
        stats = zeros((n, 2, pw.cout_pad), y.device)
        a.stats = stats.data_ptr()
    call("smsut_conv_tc", C.byref(a), _stream())
    if not want_stats:
        return y
    return y, (resolve(stats) if stats is not None else in_stats(y))


def conv_dgrad(dy, pw, splits=None):
    """dx (or one dx per concatenated source) of conv_fprop."""
    n, h, w, c = dy.shape
    assert c == pw.cout_pad
    if splits is None or len(splits) == 1:
        dx = torch.empty((n, h, w, pw.cin_pad), dtype=BF16, device=dy.device)
        conv_tc(TC_CONV, pw.kh, n, h, w, [(dy, c, c)], pw.dgrad, pw.cin_pad, pw.cin_pad, dx, pw.cin_pad)
        return [dx]
    c0, c1 = splits
    dx0 = torch.empty((n, h, w, c0), dtype=BF16, device=dy.device)
    dx1 = torch.empty((n, h, w, c1), dtype=BF16, device=dy.device)
    conv_tc(TC_CONV, pw.kh, n, h, w, [(dy, c, c)], pw.dgrad, pw.cin_pad, pw.cin_pad, dx0, c0, 0, dx1, c1, 0, c0)
    return [dx0, dx1]


def conv_dgrad_accumulate(dy, pw, dxs):
    """dxs[i] += the slice of dgrad(dy; W) that belongs to source i (fused read-modify-write epilogue)."""
    n, h, w, c = dy.shape
    if len(dxs) == 1:
        conv_tc(TC_CONV, pw.kh, n, h, w, [(dy, c, c)], pw.dgrad, pw.cin_pad, pw.cin_pad, dxs[0], pw.cin_pad,
                accumulate=True)
    else:
        c0, c1 = dxs[0].shape[3], dxs[1].shape[3]
        conv_tc(TC_CONV, pw.kh, n, h, w, [(dy, c, c)], pw.dgrad, pw.cin_pad, pw.cin_pad, dxs[0], c0, 0, dxs[1], c1, 0,
                c0, accumulate=True)


def wgrad_tc(kind, ksize, n, h, w, x, x_c, dy, dy_c, dw, cin_total, ci_off, cout_total, c_valid=0, tap_major=False):
    a = WgradTcArgs()
    a.kind, a.ksize, a.n, a.h, a.w = kind, ksize, n, h, w
    a.x, a.x_c, a.x_ld = _chk(x, BF16, "wgrad x").data_ptr(), x_c, x.shape[-1]
    a.dy, a.dy_c, a.dy_ld = _chk(dy, BF16, "wgrad dy").data_ptr(), dy_c, dy.shape[-1]
    a.dw = _chk(dw, F32, "wgrad dw").data_ptr()
    a.cin_total, a.ci_off, a.cout_total, a.c_valid = cin_total, ci_off, cout_total, c_valid
    a.dw_layout = 1 if tap_major else 0
    call("smsut_wgrad_tc", C.byref(a), _stream())


def conv_wgrad(xs, dy, pw, out=None):
    """fp32 OIHW weight gradient of conv_fprop, accumulated (atomics) into `out` or a fresh zeroed tensor."""
    n, h, w, _ = xs[0].shape
    dw = out if out is not None else zeros(tuple(pw.weight.shape), pw.weight.device)
    scratch = _wgrad_scratch(pw.weight, out)
    dst = scratch if scratch is not None else dw

    def launch():
        off = 0
        for x in xs:
            c = x.shape[3]
            valid = min(c, pw.cin - off)
            wgrad_tc(TC_CONV, pw.kh, n, h, w, x, c, dy, pw.cout_pad, dst, pw.cin, off, pw.cout, c_valid=valid,
                     tap_major=scratch is not None)
            off += c
    if out is not None:
        side_run(launch, (list(xs), dy))
    else:
        launch()
        resolve(dw)
    return dw


def convt_fprop(x, pw):
    n, h, w, c = x.shape
    assert c == pw.cin
    y = torch.empty((n, 2 * h, 2 * w, pw.cout), dtype=BF16, device=x.device)
    conv_tc(TC_CONVT_FWD, 1, n, h, w, [(x, c, c)], pw.fprop, 4 * pw.cout, 4 * pw.cout, y, pw.cout)
    return y


def convt_dgrad(dy, pw):
    n, h2, w2, c = dy.shape
    assert c == pw.cout
    dx = torch.empty((n, h2 // 2, w2 // 2, pw.cin), dtype=BF16, device=dy.device)
    conv_tc(TC_CONVT_DGRAD, 1, n, h2 // 2, w2 // 2, [(dy, c, c)], pw.dgrad, pw.cin, pw.cin, dx, pw.cin)
    return dx


def convt_wgrad(x, dy, pw, out=None):
    n, h, w, c = x.shape
    dw = out if out is not None else zeros(tuple(pw.weight.shape), pw.weight.device)
    scratch = _wgrad_scratch(pw.weight, out)
    dst = scratch if scratch is not None else dw
    if out is not None:
        side_run(lambda: wgrad_tc(TC_CONVT_FWD, 1, n, h, w, x, c, dy, pw.cout, dst, pw.cin, 0, pw.cout,
                                  tap_major=scratch is not None), (x, dy))
    else:
        wgrad_tc(TC_CONVT_FWD, 1, n, h, w, x, c, dy, pw.cout, dst, pw.cin, 0, pw.cout)
        resolve(dw)
    return dw


# ----------------------------------------------------------------------------------------------
# direct convolutions (stems / heads)
# ----------------------------------------------------------------------------------------------
def _direct_args(x, weight, y, stride, pad, bias, act, slope, accumulate, cin=None):
    cout, wcin, kh, kw = weight.shape
    n, h, w, x_ld = x.shape
    _, ho, wo, y_ld = y.shape
    a = ConvDirectArgs()
    a.n, a.h, a.w, a.cin = n, h, w, wcin if cin is None else cin
    a.cout, a.kh, a.kw, a.stride, a.pad, a.ho, a.wo = cout, kh, kw, stride, pad, ho, wo
    a.x, a.x_ld, a.x_f32 = x.data_ptr(), x_ld, int(x.dtype == F32)
    a.wt = _chk(weight, F32, "direct conv weight").data_ptr()
    a.bias = bias.data_ptr() if bias is not None else 0
    a.y, a.y_ld, a.y_f32 = y.data_ptr(), y_ld, int(y.dtype == F32)
    a.act, a.slope, a.accumulate = act, slope, int(accumulate)
    return a


def direct_out_hw(h, w, k, stride, pad):
    return (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1


def conv_direct_fprop(x, weight, stride, pad, bias=None, act=ACT_NONE, slope=0.01, out_c=None, out_f32=False):
    """x (N,H,W,x_ld) bf16|fp32 with weight.shape[1] live channels -> (N,Ho,Wo,out_c); channels >= Cout are zero."""
    cout, _, kh, kw = weight.shape
    n, h, w, _ = x.shape
    ho, wo = direct_out_hw(h, w, kh, stride, pad)
    y = torch.empty((n, ho, wo, out_c or cout), dtype=F32 if out_f32 else BF16, device=x.device)
    a = _direct_args(x, weight, y, stride, pad, bias, act, slope, False)
    call("smsut_conv_direct_fprop", C.byref(a), _stream())
    return y


def conv_direct_dgrad(dy, weight, x_shape, x_dtype, stride, pad):
    dx = torch.empty(x_shape, dtype=x_dtype, device=dy.device)
    a = _direct_args(dx, weight, dy, stride, pad, None, ACT_NONE, 0.0, False)
    call("smsut_conv_direct_dgrad", C.byref(a), _stream())
    return dx


def head1x1_bwd(x, dy, y, weight, want_dx, dw=None, db=None, want_bias=False):
    """fused backward of a 1x1 head: returns dx (bf16 or None), dW, dbias"""
    n, h, w, c = _chk(x, BF16, "head x").shape
    cout = weight.shape[0]
    dx = torch.empty_like(x) if want_dx else None
    fresh_w, fresh_b = dw is None, want_bias and db is None
    dw = dw if dw is not None else zeros(tuple(weight.shape), x.device)
    if fresh_b:
        db = zeros(cout, x.device)
    call("smsut_head1x1_bwd", _p(x), _p(_chk(dy, F32, "head dy")), _p(y), _p(weight), _p(dx), _p(dw), _p(db), n * h * w, c,
         cout, _stream())
    if fresh_w:
        resolve(dw)
    if fresh_b:
        resolve(db)
    return dx, dw, db


def conv_direct_wgrad(x, dy, weight, stride, pad, want_bias, dw=None, db=None):
    side = dw is not None and (db is not None or not want_bias)     # accumulating into the flat gradient buffer
    fresh_w, fresh_b = dw is None, want_bias and db is None
    dw = dw if dw is not None else zeros(tuple(weight.shape), weight.device)
    if fresh_b:
        db = zeros(weight.shape[0], weight.device)
    if not want_bias:
        db = None
    a = _direct_args(x, weight, dy, stride, pad, None, ACT_NONE, 0.0, False)
    if side:
        side_run(lambda: call("smsut_conv_direct_wgrad", C.byref(a), _p(dw), _p(db), _stream()), (x, dy, a))
    else:
        call("smsut_conv_direct_wgrad", C.byref(a), _p(dw), _p(db), _stream())
        if fresh_w:
            resolve(dw)
        if fresh_b and db is not None:
            resolve(db)
    return dw, db


# ----------------------------------------------------------------------------------------------
# InstanceNorm / activations
# ----------------------------------------------------------------------------------------------
def in_stats(x):
    n, h, w, c = _chk(x, BF16, "in_stats x").shape
    stats = zeros((n, 2, c), x.device)
    call("smsut_in_stats", _p(x), n, h * w, c, _p(stats), _stream())
    return resolve(stats)


def in_apply(xa, sa, ga, ba, xb=None, sb=None, gb=None, bb=None, res=None, act=ACT_NONE, slope=0.01, c_params=None):
    n, h, w, c = xa.shape
    out = torch.empty_like(xa)
    call("smsut_in_apply", _p(xa), _p(sa), _p(ga), _p(ba), _p(xb), _p(sb), _p(gb), _p(bb), _p(res), _p(out), n, h * w, c,
         c if c_params is None else c_params, act, slope, _stream())
    return out


def bn_pool(rows):
    """(n, k, c) fp32 table -> the same shape with every row = the mean over n (BatchNorm = pooled InstanceNorm sums)"""
    n, k, c = _chk(rows, F32, "bn_pool rows").shape
    out = torch.empty_like(rows)
    call("smsut_bn_pool", _p(rows), _p(out), n, k, c, _stream())
    return out


def bn_running_update(pooled, hw, running_mean, running_var, momentum):
    n, _, c = pooled.shape
    call("smsut_bn_running_update", _p(pooled), n, hw, c, running_mean.numel(), float(momentum),
         _p(_chk(running_mean, F32, "running_mean")), _p(_chk(running_var, F32, "running_var")), _stream())


def bn_eval_stats(running_mean, running_var, n, hw, c):
    stats = torch.empty((n, 2, c), dtype=F32, device=running_mean.device)
    call("smsut_bn_eval_stats", _p(_chk(running_mean, F32, "running_mean")), _p(_chk(running_var, F32, "running_var")),
         _p(stats), n, hw, c, running_mean.numel(), _stream())
    return stats


def in_bwd(dout, out, xa, sa, ga, xb=None, sb=None, gb=None, want_res=False, act=ACT_NONE, slope=0.01, c_params=None,
           targets=None, batch=False, betas=None):
    """returns dxa, dgamma_a, dbeta_a, dxb, dgamma_b, dbeta_b, dres.  `targets` = (dga, dba, dgb, dbb) fp32 tensors the
    parameter gradients are accumulated INTO (the flat .grad views); they are then returned as None.
    batch: BatchNorm -- the reductions are pooled over the samples between the two passes.
    betas = (beta_a, beta_b | None): the forward had no residual input, so the kernels recompute the activation's
    sign from xa / xb instead of streaming `out` (one tensor less in each pass)."""
    n, h, w, c = xa.shape
    cp = c if c_params is None else c_params
    dev = xa.device
    ba = bb = None
    if betas is not None and act != ACT_NONE:
        ba, bb = betas
        out = None
    red = zeros((n, 3, c), dev)
    dxa = torch.empty_like(xa)
    dxb = torch.empty_like(xa) if xb is not None else None
    dres = torch.empty_like(xa) if want_res else None
    pg = None
    if targets is not None:
        t = list(targets)
        ret = [None, None, None, None]
    else:
        pg = zeros((4, cp), dev)
        t = [pg[0], pg[1], pg[2] if xb is not None else None, pg[3] if xb is not None else None]
        ret = t
    if not batch:
        # InstanceNorm: both passes in one launch (the second one re-reads its strip from L2)
        counters = zeros(n, dev)
        call("smsut_in_bwd_fused", _p(dout), _p(out), _p(xa), _p(sa), _p(ga), _p(ba), _p(dxa), _p(t[0]), _p(t[1]), _p(xb),
             _p(sb), _p(gb), _p(bb), _p(dxb), _p(t[2]), _p(t[3]), _p(dres), _p(red), _p(counters), n, h * w, c, cp, act,
             slope, _stream())
    else:
        # BatchNorm: the reductions are pooled over the samples between the two passes
        call("smsut_in_bwd_reduce", _p(dout), _p(out), _p(xa), _p(sa), _p(ga), _p(ba), _p(xb), _p(sb), _p(gb), _p(bb),
             _p(red), n, h * w, c, cp, act, slope, _stream())
        resolve(red)
        call("smsut_bn_pool", _p(red), _p(red), n, 3, c, _stream())
        call("smsut_in_bwd_apply", _p(dout), _p(out), _p(xa), _p(sa), _p(ga), _p(ba), _p(dxa), _p(t[0]), _p(t[1]), _p(xb),
             _p(sb), _p(gb), _p(bb), _p(dxb), _p(t[2]), _p(t[3]), _p(dres), _p(red), n, h * w, c, cp, act, slope,
             _stream())
    if pg is not None:
        resolve(pg)
    return dxa, ret[0], ret[1], dxb, ret[2], ret[3], dres


def in_bwd2(u, dy, x, stats, gamma):
    """double backward of plain InstanceNorm: returns g_dy, g_x, dgamma"""
    n, h, w, c = x.shape
    red2 = zeros((n, 5, c), x.device)
    counters = zeros(n, x.device)
    g_dy, g_x = torch.empty_like(x), torch.empty_like(x)
    dgamma = zeros(c, x.device)
    call("smsut_in_bwd2_fused", _p(u), _p(dy), _p(x), _p(stats), _p(gamma), _p(red2), _p(counters), _p(g_dy), _p(g_x),
         _p(dgamma), n, h * w, c, _stream())
    return g_dy, g_x, resolve(dgamma)


def act_fwd(x, act, slope=0.01):
    y = torch.empty_like(x)
    call("smsut_act_fwd", _p(_chk(x, BF16, "act x")), _p(y), x.numel(), act, slope, _stream())
    return y


def act_bwd(dy, ref, add=None, act=ACT_LRELU, slope=0.01):
    dx = torch.empty_like(dy)
    call("smsut_act_bwd", _p(_chk(dy, BF16, "act dy")), _p(ref), _p(add), _p(dx), dy.numel(), act, slope, _stream())
    return dx


def add_bf16(a, b):
    out = torch.empty_like(a)
    call("smsut_add_bf16", _p(_chk(a, BF16, "add a")), _p(_chk(b, BF16, "add b")), _p(out), a.numel(), _stream())
    return out


def colsum(x, out=None):
    """column sums of a (rows, c) bf16 matrix, accumulated INTO `out` (fp32, e.g. a flat-gradient view) when given"""
    rows, c = x.shape
    fresh = out is None
    if fresh:
        out = zeros(c, x.device)
    call("smsut_colsum_bf16", _p(_chk(x, BF16, "colsum x")), rows, c, _p(_chk(out, F32, "colsum out")), _stream())
    return resolve(out) if fresh else out


# ----------------------------------------------------------------------------------------------
# pooling / resampling / layout
# ----------------------------------------------------------------------------------------------
def maxpool2_fwd(x):
    n, h, w, c = _chk(x, BF16, "maxpool x").shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=BF16, device=x.device)
    call("smsut_maxpool2_fwd", _p(x), _p(y), n, h, w, c, _stream())
    return y


def maxpool2_bwd(x, dy, add=None):
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    call("smsut_maxpool2_bwd", _p(x), _p(_chk(dy, BF16, "maxpool dy")), _p(add), _p(dx), n, h, w, c, _stream())
    return dx


def avgpool2_fwd(x):
    n, h, w, c = _chk(x, BF16, "avgpool x").shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=BF16, device=x.device)
    call("smsut_avgpool2_fwd", _p(x), _p(y), n, h, w, c, _stream())
    return y


def avgpool2_bwd(dy, add=None):
    n, ho, wo, c = _chk(dy, BF16, "avgpool dy").shape
    dx = torch.empty((n, 2 * ho, 2 * wo, c), dtype=BF16, device=dy.device)
    call("smsut_avgpool2_bwd", _p(dy), _p(add), _p(dx), n, 2 * ho, 2 * wo, c, _stream())
    return dx


def bilinear2_fwd(x):
    n, h, w, c = _chk(x, BF16, "bilinear x").shape
    y = torch.empty((n, 2 * h, 2 * w, c), dtype=BF16, device=x.device)
    call("smsut_bilinear2_fwd", _p(x), _p(y), n, h, w, c, _stream())
    return y


def bilinear2_bwd(dy):
    n, h2, w2, c = _chk(dy, BF16, "bilinear dy").shape
    dx = torch.empty((n, h2 // 2, w2 // 2, c), dtype=BF16, device=dy.device)
    call("smsut_bilinear2_bwd", _p(dy), _p(dx), n, h2 // 2, w2 // 2, c, _stream())
    return dx


def nchw_to_nhwc(x, c_pad):
    n, c, h, w = _chk(x, F32, "nchw x").shape
    y = torch.empty((n, h, w, c_pad), dtype=BF16, device=x.device)
    call("smsut_nchw_f32_to_nhwc_bf16", _p(x), _p(y), n, c, h, w, c_pad, _stream())
    return y


def nhwc_to_nchw(x, c):
    n, h, w, ld = _chk(x, BF16, "nhwc x").shape
    y = torch.empty((n, c, h, w), dtype=F32, device=x.device)
    call("smsut_nhwc_bf16_to_nchw_f32", _p(x), _p(y), n, c, h, w, ld, _stream())
    return y


def build_tsl_input(x, m, c_pad):
    n, _, h, w = _chk(x, F32, "tsl x").shape
    y = torch.empty((n, h, w, c_pad), dtype=BF16, device=x.device)
    call("smsut_build_tsl_input", _p(x), _p(_chk(m, F32, "tsl m")), _p(y), n, h * w, m.shape[1], c_pad, _stream())
    return y


def augment_batch(images, labels, index, params):
    """images / labels: the resident u8 dataset (slices, h, w); index int64 (n,); params fp32 (n, 66): see
    smsut_augment_batch.  Returns (x fp32 (n,1,h,w) in [-1,1], y int64 (n,h,w))."""
    if images.device.type != "cuda":
        raise _lib.SmsutError("augment_batch: SMSUT kernels need CUDA tensors (there is no CPU fallback)")
    slices, h, w = images.shape
    assert images.dtype == torch.uint8 and labels.dtype == torch.uint8 and labels.shape == images.shape
    assert images.is_contiguous() and labels.is_contiguous()
    n = index.numel()
    assert index.dtype == torch.int64 and params.dtype == F32 and tuple(params.shape) == (n, 66) and params.is_contiguous()
    x = torch.empty((n, 1, h, w), dtype=F32, device=images.device)
    y = torch.empty((n, h, w), dtype=torch.int64, device=images.device)
    call("smsut_augment_batch", _p(images), _p(labels), _p(index), _p(params), _p(x), _p(y), n, h, w, _stream())
    return x, y


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
def dice_ce_fwd(logits, labels, label_logits, acc):
    """logits fp32 (npix, c) NHWC-flattened; accumulates tp/fp/fn/ce sums into acc[3c+1]."""
    npix, c = logits.shape
    call("smsut_dice_ce_fwd", _p(_chk(logits, F32, "dice logits")), _p(labels), _p(label_logits), _p(acc), npix, c,
         _stream())
    resolve(acc)


def dice_ce_finish(acc, npix_total, c, w_dc, w_ce):
    loss = torch.empty(1, dtype=F32, device=acc.device)
    call("smsut_dice_ce_finish", _p(acc), _p(loss), npix_total, c, w_dc, w_ce, _stream())
    return loss


def dice_ce_bwd(logits, labels, label_logits, acc, gscale, scale, npix_total, w_dc, w_ce):
    npix, c = logits.shape
    d = torch.empty_like(logits)
    call("smsut_dice_ce_bwd", _p(logits), _p(labels), _p(label_logits), _p(acc), _p(gscale), scale, _p(d), npix,
         npix_total, c, w_dc, w_ce, _stream())
    return d


def softmax_mse_fwd(zs, zt, out):
    npix, c = zs.shape
    call("smsut_softmax_mse_fwd", _p(_chk(zs, F32, "mse zs")), _p(_chk(zt, F32, "mse zt")), _p(out), npix, c, _stream())
    resolve(out)


def softmax_mse_bwd(zs, zt, gscale):
    d = torch.empty_like(zs)
    call("smsut_softmax_mse_bwd", _p(zs), _p(zt), _p(gscale), _p(d), zs.shape[0], zs.shape[1], _stream())
    return d


# ---- coraNet losses (csrc/coranet.cu; trainer/coraNetTrainer.py) --------------------------------------------------
def heads_split_fwd(z, nlab, nheads):
    """(npix, 1 + nheads*nlab) fp32 -> (nheads, npix, 1 + nlab): head h = [channel 0, channels 1 + h*nlab .. (h+1)*nlab]"""
    npix, c = z.shape
    assert c == 1 + nheads * nlab
    out = torch.empty((nheads, npix, 1 + nlab), dtype=F32, device=z.device)
    call("smsut_heads_split_fwd", _p(_chk(z, F32, "heads z")), _p(out), npix, nlab, nheads, _stream())
    return out


def heads_split_bwd(dheads, nlab, nheads):
    npix = dheads.shape[1]
    dz = torch.empty((npix, 1 + nheads * nlab), dtype=F32, device=dheads.device)
    call("smsut_heads_split_bwd", _p(_chk(dheads, F32, "heads grad")), _p(dz), npix, nlab, nheads, _stream())
    return dz


def wce_fwd(z, y, cw, mask, acc):
    """acc[0] += sum m*cw[y]*nll, acc[1] += sum cw[y], acc[2] += sum m (cw / mask None = ones)"""
    npix, c = z.shape
    call("smsut_wce_fwd", _p(_chk(z, F32, "wce z")), _p(y), _p(cw), _p(mask), _p(acc), npix, c, _stream())
    resolve(acc)


def wce_bwd(z, y, cw, mask, acc, gscale, mask_den):
    d = torch.empty_like(z)
    call("smsut_wce_bwd", _p(z), _p(y), _p(cw), _p(mask), _p(acc), _p(gscale), 1 if mask_den else 0, _p(d), z.shape[0],
         z.shape[1], _stream())
    return d


def softmax_mse_masked_fwd(zs, zt, mask, invert, acc):
    """acc[0] += sum_p m_p sum_c (softmax(zs) - softmax(zt))^2, acc[1] += sum_p m_p, m = 1 - mask if invert else mask"""
    npix, c = zs.shape
    call("smsut_softmax_mse_masked_fwd", _p(_chk(zs, F32, "mse zs")), _p(_chk(zt, F32, "mse zt")), _p(_chk(mask, F32, "mse mask")),
         1 if invert else 0, _p(acc), npix, c, _stream())
    resolve(acc)


def softmax_mse_masked_bwd(zs, zt, mask, invert, acc, gscale):
    d = torch.empty_like(zs)
    call("smsut_softmax_mse_masked_bwd", _p(zs), _p(zt), _p(mask), 1 if invert else 0, _p(acc), _p(gscale), _p(d),
         zs.shape[0], zs.shape[1], _stream())
    return d


def argmax_c(logits):
    npix, c = logits.shape
    out = torch.empty(npix, dtype=torch.int64, device=logits.device)
    call("smsut_argmax_c", _p(_chk(logits, F32, "argmax logits")), _p(out), npix, c, _stream())
    return out


def confusion_counts(logits, labels, conf):
    """conf[label, argmax(logits)] += 1; logits fp32 (npix, c), labels int64 (npix,), conf int64 (c, c) accumulated"""
    npix, c = _chk(logits, F32, "confusion logits").shape
    assert labels.dtype == torch.int64 and labels.numel() == npix and conf.dtype == torch.int64 and conf.numel() == c * c
    call("smsut_confusion_counts", _p(logits), _p(labels.contiguous()), _p(conf), npix, c, _stream())
    return conf


def l1_fwd(a, b, out, scale):
    call("smsut_l1_fwd", _p(_chk(a, F32, "l1 a")), _p(_chk(b, F32, "l1 b")), _p(out), a.numel(), scale, _stream())
    resolve(out)


def l1_bwd(a, b, gscale, scale):
    da = torch.empty_like(a)
    call("smsut_l1_bwd", _p(a), _p(b), _p(gscale), scale, _p(da), a.numel(), _stream())
    return da


def sum_f32(x, out, scale):
    call("smsut_sum_f32", _p(_chk(x, F32, "sum x")), _p(out), x.numel(), scale, _stream())
    resolve(out)


def fill_f32(x, value):
    call("smsut_fill_f32", _p(x), x.numel(), value, _stream())


def fill_scaled(shape, gscale, scale, device):
    x = torch.empty(shape, dtype=F32, device=device)
    call("smsut_fill_scaled_f32", _p(x), x.numel(), _p(gscale), scale, _stream())
    return x


def tanh_bwd(dy, y):
    dx = torch.empty_like(y)
    call("smsut_tanh_bwd", _p(_chk(dy, F32, "tanh dy")), _p(_chk(y, F32, "tanh y")), _p(dx), y.numel(), _stream())
    return dx


def lerp_rows(alpha, x, y):
    out = torch.empty_like(x)
    rows = x.shape[0]
    call("smsut_lerp_rows_f32", _p(_chk(alpha, F32, "lerp alpha")), _p(_chk(x, F32, "lerp x")), _p(_chk(y, F32, "lerp y")),
         _p(out), rows, x.numel() // rows, _stream())
    return out


def ce_rows_fwd(logits, target, out, scale):
    rows, c = logits.shape
    call("smsut_ce_rows_fwd", _p(_chk(logits, F32, "ce logits")), _p(target), _p(out), rows, c, scale, _stream())
    resolve(out)


def ce_rows_bwd(logits, target, gscale, scale):
    rows, c = logits.shape
    d = torch.empty_like(logits)
    call("smsut_ce_rows_bwd", _p(logits), _p(target), _p(gscale), scale, _p(d), rows, c, _stream())
    return d


def gp_fwd(g, out, scale):
    b = g.shape[0]
    per = g.numel() // b
    norm = torch.empty(b, dtype=F32, device=g.device)
    norm2 = zeros(b, g.device)
    call("smsut_gp_fwd", _p(_chk(g, F32, "gp g")), _p(norm), _p(norm2), _p(out), b, per, scale, _stream())
    resolve(out)
    return norm


def gp_bwd(g, norm, gscale, scale):
    b = g.shape[0]
    u = torch.empty_like(g)
    call("smsut_gp_bwd", _p(g), _p(norm), _p(gscale), scale, _p(u), b, g.numel() // b, _stream())
    return u


def gather_rows(feat, ids):
    n, h, w, c = _chk(feat, BF16, "gather feat").shape
    out = torch.empty((n * ids.numel(), c), dtype=BF16, device=feat.device)
    call("smsut_gather_rows", _p(feat), _p(ids), _p(out), n, h * w, c, ids.numel(), _stream())
    return out


def scatter_rows_add(dout, ids, dfeat):
    n, h, w, c = dfeat.shape
    call("smsut_scatter_rows_add", _p(_chk(dout, BF16, "scatter dout")), _p(ids), _p(dfeat), n, h * w, c, ids.numel(),
         _stream())


def l2norm_fwd(x):
    rows, c = _chk(x, F32, "l2norm x").shape
    y = torch.empty_like(x)
    norm = torch.empty(rows, dtype=F32, device=x.device)
    call("smsut_l2norm_fwd", _p(x), _p(y), _p(norm), rows, c, _stream())
    return y, norm


def l2norm_bwd(dy, y, norm):
    rows, c = y.shape
    dx = torch.empty((rows, c), dtype=BF16, device=y.device)
    call("smsut_l2norm_bwd", _p(_chk(dy, F32, "l2norm dy")), _p(y), _p(norm), _p(dx), rows, c, _stream())
    return dx


def patchnce_fwd(q, k, groups, np_, inv_t, out, scale):
    rows, c = q.shape
    loss_rows = torch.empty(rows, dtype=F32, device=q.device)
    call("smsut_patchnce_fwd", _p(_chk(q, F32, "nce q")), _p(_chk(k, F32, "nce k")), _p(loss_rows), _p(out), groups,
         np_, c, inv_t, scale, _stream())
    resolve(out)
    return loss_rows


def patchnce_bwd(q, k, groups, np_, inv_t, gscale, scale):
    dq = torch.empty_like(q)
    call("smsut_patchnce_bwd", _p(q), _p(k), _p(gscale), scale, _p(dq), groups, np_, q.shape[1], inv_t, _stream())
    return dq


# ----------------------------------------------------------------------------------------------
# optimisers over flat fp32 buffers
# ----------------------------------------------------------------------------------------------
def sgd_step(p, g, mom, lr, momentum, weight_decay, grad_scale=1.0):
    param_generation[0] += 1
    call("smsut_sgd_step", _p(p), _p(g), _p(mom), p.numel(), _p(lr), momentum, weight_decay, grad_scale, _stream())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, state, grad_scale=1.0):
    param_generation[0] += 1
    call("smsut_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), _p(lr), beta1, beta2, eps, weight_decay, _p(state),
         grad_scale, _stream())


def ema_update(ema, p, alpha):
    param_generation[0] += 1
    call("smsut_ema_update", _p(ema), _p(p), p.numel(), _p(alpha), _stream())


def poly_lr_tick(iter_state, lr_out, base_lr, max_iter, power):
    call("smsut_poly_lr_tick", _p(iter_state), _p(lr_out), base_lr, max_iter, power, _stream())
