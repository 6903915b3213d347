"""GPU parity of every C-ABI kernel against a plain PyTorch fp32 statement of the same op.

Inputs are rounded to bf16 first so the only differences are accumulation order and the final bf16 rounding:
tolerance is relative L2 <= 1e-2 for bf16 outputs (north_star: 2e-2 per layer) and 1e-4..1e-3 for fp32 outputs.
All calls go through the C ABI (smsut_b200.ops -> ctypes -> libsmsut_b200.so).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"
# the torch statements must be true fp32 (cuDNN / cuBLAS default to TF32 for fp32 inputs)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def bf(t):
    return t.to(torch.bfloat16)


def nhwc(t):  # NCHW fp32 -> NHWC bf16 contiguous
    return bf(t).permute(0, 2, 3, 1).contiguous()


def nchw(t):  # NHWC -> NCHW fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def rnd(*shape, scale=1.0, seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    return bf(torch.randn(*shape, device=DEV) * scale).float()


@pytest.fixture(scope="module")
def ops(pkg):
    from smsut_b200 import ops as o
    return o


def make_pack(ops, weight, transposed=False):
    pw = ops.PackedWeight(weight, transposed=transposed)
    ops.PackTable([pw]).refresh()
    return pw


# (cin list, cout, h, n, ksize)  -- the layer classes of SURVEY.md section 8(d) plus the discriminator's
CONV_CASES = [
    ([16], 16, 256, 2, 3), ([16, 16], 16, 256, 1, 3), ([16], 32, 128, 2, 3), ([32], 32, 128, 2, 3),
    ([32, 32], 32, 128, 2, 3), ([32], 64, 64, 2, 3), ([64], 64, 64, 3, 3), ([64, 64], 64, 64, 2, 3),
    ([64], 128, 32, 2, 3), ([128], 128, 32, 2, 3), ([128, 128], 128, 32, 2, 3), ([128], 256, 16, 2, 3),
    ([256], 256, 16, 2, 3), ([256], 256, 8, 4, 3), ([256], 256, 4, 16, 3), ([256], 256, 4, 3, 3),
    ([16], 32, 128, 2, 1), ([32, 32], 32, 128, 1, 1), ([128], 256, 16, 2, 1), ([16], 16, 256, 1, 1),
    ([64], 32, 64, 2, 1), ([256], 128, 32, 2, 1),
]


@pytest.mark.parametrize("cins,cout,h,n,ks", CONV_CASES)
def test_conv_tc_fprop_dgrad_wgrad(ops, cins, cout, h, n, ks):
    torch.manual_seed(1)
    cin = sum(cins)
    xs = [rnd(n, c, h, h) for c in cins]
    wt = rnd(cout, cin, ks, ks, scale=(2.0 / (cin * ks * ks)) ** 0.5)
    pw = make_pack(ops, wt)
    x_cat = torch.cat(xs, 1)
    y_ref = F.conv2d(x_cat, wt, padding=ks // 2)
    y = ops.conv_fprop([nhwc(x) for x in xs], pw)
    assert rel(nchw(y), y_ref) < 1e-2

    dy = rnd(n, cout, h, h)
    dx_ref = torch.nn.grad.conv2d_input(x_cat.shape, wt, dy, padding=ks // 2)
    dxs = ops.conv_dgrad(nhwc(dy), pw, splits=cins)
    dx = torch.cat([nchw(d) for d in dxs], 1)
    assert rel(dx, dx_ref) < 1e-2

    dw_ref = torch.nn.grad.conv2d_weight(x_cat, wt.shape, dy, padding=ks // 2)
    dw = ops.conv_wgrad([nhwc(x) for x in xs], nhwc(dy), pw)
    assert rel(dw, dw_ref) < 1e-2


VT_CASES = [([32], 64, 64, 2), ([64], 64, 64, 3), ([64, 64], 64, 64, 2), ([64], 128, 32, 2), ([128, 128], 128, 32, 2),
            ([128], 256, 16, 2), ([256], 256, 16, 3), ([16], 32, 64, 2), ([32, 16], 48, 32, 2), ([256], 256, 8, 4)]


@pytest.mark.parametrize("cins,cout,h,n", VT_CASES)
def test_conv_tc_vertical_tap_sharing(ops, monkeypatch, cins, cout, h, n):
    """SMSUT_TC_VT=1: a K step of conv_tc_kernel is one (dx, source, chunk) whose A box of th + 2 rows is fetched once
    and serves the three vertical taps through descriptor offsets.  Same results as the one-box-per-tap path
    (different accumulation order only) and as the fp32 statement; 8x8 (tile spans two images) must fall back."""
    torch.manual_seed(21)
    cin = sum(cins)
    xs = [rnd(n, c, h, h) for c in cins]
    wt = rnd(cout, cin, 3, 3, scale=(2.0 / (cin * 9)) ** 0.5)
    pw = make_pack(ops, wt)
    x_cat, dy = torch.cat(xs, 1), rnd(n, cout, h, h)
    y_ref = F.conv2d(x_cat, wt, padding=1)
    dx_ref = torch.nn.grad.conv2d_input(x_cat.shape, wt, dy, padding=1)
    got = {}
    for vt in ("0", "2", "1"):          # off / wherever legal / the default rule (th >= 4, >= 2 stages)
        monkeypatch.setenv("SMSUT_TC_VT", vt)
        y, st = ops.conv_fprop([nhwc(x) for x in xs], pw, want_stats=True)
        dxs = ops.conv_dgrad(nhwc(dy), pw, splits=cins)
        got[vt] = (nchw(y), torch.cat([nchw(d) for d in dxs], 1), st)
        assert rel(got[vt][0], y_ref) < 1e-2 and rel(got[vt][1], dx_ref) < 1e-2, vt
    for vt in ("1", "2"):
        assert rel(got[vt][0], got["0"][0]) < 2e-3 and rel(got[vt][1], got["0"][1]) < 2e-3
        assert rel(got[vt][2], got["0"][2]) < 1e-3


@pytest.mark.parametrize("mc", ["2", "4"])
@pytest.mark.parametrize("cins,cout,h,n,ks", [([32], 64, 64, 2, 3), ([64, 64], 64, 64, 2, 3), ([64], 128, 32, 2, 3),
                                              ([128, 128], 128, 32, 4, 3), ([128], 256, 16, 4, 3), ([256], 256, 16, 6, 3),
                                              ([256], 256, 8, 4, 3), ([64], 32, 64, 2, 1), ([256], 128, 32, 2, 1)])
def test_conv_tc_weight_multicast(ops, monkeypatch, mc, cins, cout, h, n, ks):
    """SMSUT_TC_MCAST=2|4: a cluster of M tiles of conv_tc_kernel loads each weight tile once -- CTA r fetches its slice
    of the rows and TMA-multicasts it into every CTA of the cluster; the stage is released by a multicast commit of
    every CTA.  Same results (bitwise: same accumulation order) as the un-clustered kernel, with and without the
    vertical-tap sharing, incl. shapes where the cluster does not divide the tiles (falls back)."""
    torch.manual_seed(31)
    cin = sum(cins)
    xs = [rnd(n, c, h, h) for c in cins]
    wt = rnd(cout, cin, ks, ks, scale=(2.0 / (cin * ks * ks)) ** 0.5)
    pw = make_pack(ops, wt)
    x_cat, dy = torch.cat(xs, 1), rnd(n, cout, h, h)
    y_ref = F.conv2d(x_cat, wt, padding=ks // 2)
    dx_ref = torch.nn.grad.conv2d_input(x_cat.shape, wt, dy, padding=ks // 2)
    got = {}
    for key in ("0", mc):
        monkeypatch.setenv("SMSUT_TC_MCAST", key)
        y, st = ops.conv_fprop([nhwc(x) for x in xs], pw, want_stats=True)
        dxs = ops.conv_dgrad(nhwc(dy), pw, splits=cins)
        got[key] = (nchw(y), torch.cat([nchw(d) for d in dxs], 1), st)
        assert rel(got[key][0], y_ref) < 1e-2 and rel(got[key][1], dx_ref) < 1e-2, key
    assert torch.equal(got[mc][0], got["0"][0]) and torch.equal(got[mc][1], got["0"][1])
    assert rel(got[mc][2], got["0"][2]) < 1e-5


@pytest.mark.parametrize("cins,cout,h,n,ks", [([16], 16, 256, 1, 3), ([16, 16], 16, 128, 2, 3), ([32], 64, 64, 2, 3),
                                              ([64, 64], 64, 64, 2, 3), ([128], 256, 16, 2, 3), ([256], 256, 8, 4, 3),
                                              ([64], 32, 64, 2, 1)])
def test_wgrad_tap_major_scratch(ops, cins, cout, h, n, ks):
    """The weight-gradient path of the trainers: wgrad kernels accumulate into the tap-major scratch beside the flat
    gradient buffer (ops.WgradScratch, vector reductions) and smsut_unpack_wgrads folds it into the OIHW gradient."""
    import types
    torch.manual_seed(5)
    cin = sum(cins)
    xs = [rnd(n, c, h, h) for c in cins]
    wt = rnd(cout, cin, ks, ks, scale=0.1)
    pw = make_pack(ops, wt)
    dy = rnd(n, cout, h, h)
    dw_ref = torch.nn.grad.conv2d_weight(torch.cat(xs, 1), wt.shape, dy, padding=ks // 2)
    grad = torch.full_like(wt, 0.5)                       # the flat-gradient view already holds something
    sc = ops.WgradScratch([wt], [grad])
    ops.conv_wgrad([nhwc(x) for x in xs], nhwc(dy), pw, out=grad)
    ops.conv_wgrad([nhwc(x) for x in xs], nhwc(dy), pw, out=grad)      # two autograd nodes of one weight
    ops.side_join()
    assert sc.dirty and grad.eq(0.5).all()
    sc.flush()
    assert not sc.dirty and sc.flat.abs().max().item() == 0.0
    assert rel(grad - 0.5, 2 * dw_ref) < 1e-2
    sc.flush()                                            # idempotent
    assert rel(grad - 0.5, 2 * dw_ref) < 1e-2


@pytest.mark.parametrize("cin,cout,h,w,n,ks,live", [
    (16, 16, 256, 256, 2, 3, 16), (32, 16, 128, 128, 2, 3, 32), (16, 32, 128, 128, 3, 3, 16), (32, 32, 40, 128, 2, 3, 32),
    (16, 64, 24, 128, 2, 3, 16), (16, 16, 64, 256, 1, 5, 5), (16, 16, 32, 128, 2, 5, 1), (32, 32, 128, 128, 2, 1, 32),
    (16, 32, 17, 128, 5, 1, 16), (32, 64, 16, 256, 2, 1, 32), (16, 16, 512, 512, 1, 3, 8)])
def test_wgrad_warp_mma_kernel(ops, monkeypatch, cin, cout, h, w, n, ks, live):
    """wgrad_hmma_kernel (the wide, narrow-channel layers: W % 128 == 0, 16 / 32 input channels) on every template
    instance, ragged row segments, non-square images, zero-padded input channels (`live` of `cin` exist in the weight)
    and both destination layouts, vs torch.nn.grad.conv2d_weight on the same bf16-exact values."""
    monkeypatch.setenv("SMSUT_WGRAD_HMMA", "1")          # opt-in kernel (measured slower than the tcgen05 band kernel)
    torch.manual_seed(13)
    x = rnd(n, cin, h, w)
    x[:, live:] = 0
    dy = rnd(n, cout, h, w)
    wt = rnd(cout, live, ks, ks, scale=0.1)
    pw = make_pack(ops, wt)
    assert pw.cin_pad == cin
    dw_ref = torch.nn.grad.conv2d_weight(x[:, :live], wt.shape, dy, padding=ks // 2)
    dw = ops.conv_wgrad([nhwc(x)], nhwc(dy), pw)                      # OIHW destination
    assert rel(dw, dw_ref) < 5e-3, rel(dw, dw_ref)
    grad = torch.zeros_like(wt)
    sc = ops.WgradScratch([wt], [grad])
    ops.conv_wgrad([nhwc(x)], nhwc(dy), pw, out=grad)                 # tap-major scratch, folded by unpack_wgrads
    ops.side_join()
    sc.flush()
    assert rel(grad, dw_ref) < 5e-3, rel(grad, dw_ref)


@pytest.mark.parametrize("m64,multi,cluster", [("2", "1", "4"), ("0", "1", "2"), ("2", "0", "4"), ("0", "0", "1"), ("1", "1", "1")])
@pytest.mark.parametrize("cin,cout,h,w,n,ks,live", [
    (16, 16, 256, 256, 2, 3, 16), (16, 32, 128, 128, 3, 3, 16), (16, 64, 24, 128, 2, 3, 16), (16, 16, 40, 128, 2, 3, 8),
    (16, 32, 17, 128, 5, 1, 16), (16, 16, 64, 256, 1, 5, 5), (16, 16, 3, 128, 2, 5, 1), (32, 16, 128, 128, 2, 3, 32),
    (32, 64, 2, 128, 2, 3, 32), (16, 16, 512, 512, 1, 3, 16)])
def test_wgrad_band_kernel_variants(ops, monkeypatch, m64, multi, cluster, cin, cout, h, w, n, ks, live):
    """wgrad_band_kernel with M = 64 UMMAs (16-channel x chunks of the 1x1 / 3x3 layers: accumulator rows in TMEM lanes
    0..15 of each warp quarter) and with M = 128 (SMSUT_WGRAD_M64=0; 32-channel chunks and 5x5 always), with one
    MMA-issuing warp per vertical tap and with a single issuer (SMSUT_WGRAD_MULTI=0), with the accumulators of a
    thread-block cluster summed through distributed shared memory before the atomics (SMSUT_WGRAD_CLUSTER=4 / 2) and
    without (=1), ragged and very short row segments, zero-padded input channels, both destination layouts, vs torch.nn.grad.conv2d_weight."""
    monkeypatch.setenv("SMSUT_WGRAD_M64", m64)
    monkeypatch.setenv("SMSUT_WGRAD_MULTI", multi)
    monkeypatch.setenv("SMSUT_WGRAD_CLUSTER", cluster)
    monkeypatch.delenv("SMSUT_WGRAD_HMMA", raising=False)
    torch.manual_seed(17)
    x = rnd(n, cin, h, w)
    x[:, live:] = 0
    dy = rnd(n, cout, h, w)
    wt = rnd(cout, live, ks, ks, scale=0.1)
    pw = make_pack(ops, wt)
    assert pw.cin_pad == cin
    dw_ref = torch.nn.grad.conv2d_weight(x[:, :live], wt.shape, dy, padding=ks // 2)
    dw = ops.conv_wgrad([nhwc(x)], nhwc(dy), pw)                      # OIHW destination
    assert rel(dw, dw_ref) < 5e-3, rel(dw, dw_ref)
    grad = torch.zeros_like(wt)
    sc = ops.WgradScratch([wt], [grad])
    ops.conv_wgrad([nhwc(x)], nhwc(dy), pw, out=grad)                 # tap-major scratch, folded by unpack_wgrads
    ops.side_join()
    sc.flush()
    assert rel(grad, dw_ref) < 5e-3, rel(grad, dw_ref)


def test_convt_wgrad_tap_major_scratch(ops):
    x = rnd(2, 64, 32, 32, seed=6)
    wt = rnd(64, 32, 2, 2, scale=0.1)
    pw = make_pack(ops, wt, transposed=True)
    dy = rnd(2, 32, 64, 64)
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, stride=2).backward(dy)
    grad = torch.zeros_like(wt)
    sc = ops.WgradScratch([wt], [grad])
    ops.convt_wgrad(nhwc(x), nhwc(dy), pw, out=grad)
    ops.side_join()
    sc.flush()
    assert rel(grad, wr.grad) < 1e-2


def test_conv_tc_padded_input_channels(ops):
    """8 live channels in a 16-channel tensor (the 5x5 stem feeds enc1.conv1 this way)."""
    n, h, cin, cout = 2, 64, 8, 16
    x = rnd(n, cin, h, h, seed=2)
    wt = rnd(cout, cin, 3, 3, scale=0.1)
    pw = make_pack(ops, wt)
    xp = torch.cat([x, torch.zeros_like(x)], 1)
    y = ops.conv_fprop([nhwc(xp)], pw)
    assert rel(nchw(y), F.conv2d(x, wt, padding=1)) < 1e-2
    dy = rnd(n, cout, h, h)
    dx = nchw(ops.conv_dgrad(nhwc(dy), pw)[0])
    assert dx[:, cin:].abs().max().item() == 0.0
    assert rel(dx[:, :cin], torch.nn.grad.conv2d_input(x.shape, wt, dy, padding=1)) < 1e-2
    dw = ops.conv_wgrad([nhwc(xp)], nhwc(dy), pw)
    assert rel(dw, torch.nn.grad.conv2d_weight(x, wt.shape, dy, padding=1)) < 1e-2


def test_conv_tc_bias_relu_f32_linear(ops):
    """netF's Linear(256,256)+ReLU as a 1x1 conv over (1, 8, 128) 'pixels' with fp32 output."""
    rows, c = 1024, 256
    x = rnd(rows, c, seed=3)
    wt = rnd(c, c, scale=0.05)
    b = torch.randn(c, device=DEV)
    pw = make_pack(ops, wt.view(c, c, 1, 1))
    y = ops.conv_fprop([bf(x).view(1, 8, 128, c)], pw, bias=b, act=ops.ACT_RELU, out_f32=True)
    assert y.dtype == torch.float32
    assert rel(y.view(rows, c), F.relu(x @ wt.t() + b)) < 1e-4


@pytest.mark.parametrize("cin,cout,h,n", [(32, 16, 128, 2), (64, 32, 64, 2), (128, 64, 32, 2), (256, 128, 16, 2)])
def test_convt_tc(ops, cin, cout, h, n):
    x = rnd(n, cin, h, h, seed=4)
    wt = rnd(cin, cout, 2, 2, scale=(1.0 / cin) ** 0.5)
    pw = make_pack(ops, wt, transposed=True)
    y_ref = F.conv_transpose2d(x, wt, stride=2)
    y = ops.convt_fprop(nhwc(x), pw)
    assert rel(nchw(y), y_ref) < 1e-2
    dy = rnd(n, cout, 2 * h, 2 * h)
    dx_ref = F.conv2d(dy, wt, stride=2)
    assert rel(nchw(ops.convt_dgrad(nhwc(dy), pw)), dx_ref) < 1e-2
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, stride=2).backward(dy)
    assert rel(ops.convt_wgrad(nhwc(x), nhwc(dy), pw), wr.grad) < 1e-2


DIRECT_CASES = [
    # cin, cout, k, stride, pad, h, x_f32, bias
    (1, 8, 5, 1, 2, 64, False, False), (5, 8, 5, 1, 2, 64, False, False), (1, 16, 4, 2, 1, 64, True, True),
    (16, 1, 1, 1, 0, 64, False, True), (16, 5, 1, 1, 0, 64, False, True), (256, 1, 3, 1, 1, 4, False, False),
    (256, 4, 4, 1, 0, 4, False, False),
]


@pytest.mark.parametrize("cin,cout,k,stride,pad,h,x_f32,bias", DIRECT_CASES)
def test_conv_direct(ops, cin, cout, k, stride, pad, h, x_f32, bias):
    n = 3
    torch.manual_seed(5)
    x = rnd(n, cin, h, h)
    wt = torch.randn(cout, cin, k, k, device=DEV) * (1.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, device=DEV) if bias else None
    xin = x.permute(0, 2, 3, 1).contiguous() if x_f32 else F.pad(nhwc(x), (0, (-cin) % 8))
    y_ref = F.conv2d(x, wt, b, stride=stride, padding=pad)
    y = ops.conv_direct_fprop(xin, wt, stride, pad, bias=b, out_f32=True)
    assert rel(nchw(y), y_ref) < 1e-4
    ypad = ops.conv_direct_fprop(xin, wt, stride, pad, bias=b, out_c=16)
    assert ypad.dtype == torch.bfloat16 and ypad.shape[-1] == 16
    assert rel(nchw(ypad)[:, :cout], y_ref) < 1e-2
    if cout < 16:
        assert ypad[..., cout:].abs().max().item() == 0.0

    dy = torch.randn_like(y_ref)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=stride, padding=pad)
    dx = ops.conv_direct_dgrad(dyn, wt, (n, h, h, cin), torch.float32, stride, pad)
    assert rel(nchw(dx), dx_ref) < 1e-4
    dw_ref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=stride, padding=pad)
    dw, db = ops.conv_direct_wgrad(xin, dyn, wt, stride, pad, want_bias=bias)
    assert rel(dw, dw_ref) < 1e-4
    if bias:
        assert rel(db, dy.sum((0, 2, 3))) < 1e-4


def test_discriminator_stem_backward_kernels(ops):
    """the specialised backward of D's stem (Conv2d(1, 16, 4, s2, p1) + bias, network/ugan.py:202): dgrad to the fp32
    image and wgrad / dbias from bf16 gradients, vs torch.nn.grad on the same (bf16-exact) values; also a second size"""
    for n, h in ((4, 256), (3, 64)):
        torch.manual_seed(7)
        x = rnd(n, 1, h, h)
        wt = torch.randn(16, 1, 4, 4, device=DEV) * 0.25
        dy = torch.randn(n, 16, h // 2, h // 2, device=DEV).bfloat16().float()
        dyn = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
        xin = x.permute(0, 2, 3, 1).contiguous()
        b = torch.randn(16, device=DEV)
        y = ops.conv_direct_fprop(xin, wt, 2, 1, bias=b, act=2, slope=0.01, out_c=16)      # bias + LeakyReLU fused
        assert y.dtype == torch.bfloat16 and rel(nchw(y), F.leaky_relu(F.conv2d(x, wt, b, stride=2, padding=1), 0.01)) < 5e-3
        y0 = ops.conv_direct_fprop(xin, wt, 2, 1, out_c=16)                                 # the cotangent form: no bias / act
        assert rel(nchw(y0), F.conv2d(x, wt, None, stride=2, padding=1)) < 5e-3
        dx = ops.conv_direct_dgrad(dyn, wt, (n, h, h, 1), torch.float32, 2, 1)
        assert rel(nchw(dx), torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=2, padding=1)) < 1e-5
        dw, db = ops.conv_direct_wgrad(xin, dyn, wt, 2, 1, want_bias=True)
        assert rel(dw, torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=2, padding=1)) < 1e-4
        assert rel(db, dy.sum((0, 2, 3))) < 1e-4
        dw2 = torch.zeros_like(wt)
        ops.conv_direct_wgrad(xin, dyn, wt, 2, 1, want_bias=False, dw=dw2)          # accumulate form, no bias
        ops.conv_direct_wgrad(xin, dyn, wt, 2, 1, want_bias=False, dw=dw2)
        assert rel(dw2, 2 * dw) < 1e-5


def in_ref(x, g, b):
    return F.instance_norm(x, weight=g, bias=b, eps=1e-5)


@pytest.mark.parametrize("c,h,n", [(16, 64, 3), (32, 32, 2), (256, 16, 2), (64, 8, 5)])
def test_instance_norm_fwd_bwd(ops, c, h, n):
    torch.manual_seed(6)
    xa, xb, res = bf(rnd(n, c, h, h) * 2 + 0.5).float(), rnd(n, c, h, h), rnd(n, c, h, h)
    ga, ba, gb, bb = [torch.randn(c, device=DEV) for _ in range(4)]
    a, b2 = nhwc(xa), nhwc(xb)
    sa, sb = ops.in_stats(a), ops.in_stats(b2)
    assert rel(sa[:, 0], xa.sum((2, 3))) < 1e-4
    # plain IN + lrelu
    out = ops.in_apply(a, sa, ga, ba, act=ops.ACT_LRELU)
    xr = xa.clone().requires_grad_(True)
    gr, br = ga.clone().requires_grad_(True), ba.clone().requires_grad_(True)
    ref = F.leaky_relu(in_ref(xr, gr, br), 0.01)
    assert rel(nchw(out), ref) < 1e-2
    dout = rnd(n, c, h, h)
    ref.backward(dout)
    dxa, dga, dba, *_ = ops.in_bwd(nhwc(dout), out, a, sa, ga, act=ops.ACT_LRELU)
    assert rel(nchw(dxa), xr.grad) < 2e-2
    assert rel(dga, gr.grad) < 2e-2 and rel(dba, br.grad) < 2e-2
    # two normalised branches + residual + lrelu
    out2 = ops.in_apply(a, sa, ga, ba, b2, sb, gb, bb, res=nhwc(res), act=ops.ACT_LRELU)
    leaves = [t.clone().requires_grad_(True) for t in (xa, ga, ba, xb, gb, bb, res)]
    ref2 = F.leaky_relu(in_ref(*leaves[0:3]) + in_ref(*leaves[3:6]) + leaves[6], 0.01)
    assert rel(nchw(out2), ref2) < 1e-2
    ref2.backward(dout)
    got = ops.in_bwd(nhwc(dout), out2, a, sa, ga, b2, sb, gb, want_res=True, act=ops.ACT_LRELU)
    for g_, l_ in zip(got, leaves):
        g_ = nchw(g_) if g_.dim() == 4 else g_
        assert rel(g_, l_.grad) < 2e-2
    # no residual in the forward: the backward may recompute the activation's sign from xa / xb instead of reading
    # `out` (betas=...).  Both modes must use the SAME mask: one flipped element would move dx by ~3e-3 relative and
    # dbeta (= sum g) by ~1e-3, while the different summation order of the two modes moves them by ~1e-6
    for act in (ops.ACT_LRELU, ops.ACT_RELU):
        o1 = ops.in_apply(a, sa, ga, ba, act=act)
        r_out = ops.in_bwd(nhwc(dout), o1, a, sa, ga, act=act)
        r_rec = ops.in_bwd(nhwc(dout), o1, a, sa, ga, act=act, betas=(ba, None))
        assert rel(r_rec[0], r_out[0]) < 2e-4 and rel(r_rec[1], r_out[1]) < 1e-5 and rel(r_rec[2], r_out[2]) < 1e-5
        o2 = ops.in_apply(a, sa, ga, ba, b2, sb, gb, bb, act=act)
        r_out = ops.in_bwd(nhwc(dout), o2, a, sa, ga, b2, sb, gb, act=act)
        r_rec = ops.in_bwd(nhwc(dout), o2, a, sa, ga, b2, sb, gb, act=act, betas=(ba, bb))
        assert rel(r_rec[0], r_out[0]) < 2e-4 and rel(r_rec[3], r_out[3]) < 2e-4
        for i in (1, 2, 4, 5):
            assert rel(r_rec[i], r_out[i]) < 1e-5, i
    leaves = [t.clone().requires_grad_(True) for t in (xa, ga, ba, xb, gb, bb)]
    ref3 = F.leaky_relu(in_ref(*leaves[0:3]) + in_ref(*leaves[3:6]), 0.01)
    ref3.backward(dout)
    o2 = ops.in_apply(a, sa, ga, ba, b2, sb, gb, bb, act=ops.ACT_LRELU)
    got = ops.in_bwd(nhwc(dout), None, a, sa, ga, b2, sb, gb, act=ops.ACT_LRELU, betas=(ba, bb))
    for g_, l_ in zip(got[:6], leaves):
        g_ = nchw(g_) if g_.dim() == 4 else g_
        assert rel(g_, l_.grad) < 2e-2


@pytest.mark.parametrize("c,cp,h,n", [(16, 16, 64, 3), (16, 8, 32, 4), (128, 128, 16, 5)])
def test_batch_norm_on_pooled_statistics(ops, c, cp, h, n):
    """nn.BatchNorm2d (training: batch statistics + running-estimate update; eval: running estimates) from the
    InstanceNorm kernels + smsut_bn_pool / smsut_bn_running_update / smsut_bn_eval_stats.  cp < c: the stem's
    8 channels in a 16-channel tensor."""
    torch.manual_seed(16)
    x = rnd(n, cp, h, h) * 1.7 + 0.4
    x[1] *= 2.5                                     # samples with different statistics
    x = bf(x).float()
    xp = F.pad(x, (0, 0, 0, 0, 0, c - cp))
    g, b = torch.randn(cp, device=DEV), torch.randn(cp, device=DEV)
    bn = torch.nn.BatchNorm2d(cp).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(g); bn.bias.copy_(b)
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    a = nhwc(xp)
    pooled = ops.bn_pool(ops.in_stats(a))
    assert torch.equal(pooled[0], pooled[n - 1])
    ops.bn_running_update(pooled, h * h, rm, rv, 0.1)
    out = ops.in_apply(a, pooled, g, b, act=ops.ACT_RELU, c_params=cp if cp != c else None)
    xr = x.clone().requires_grad_(True)
    ref = F.relu(bn(xr))
    assert rel(nchw(out)[:, :cp], ref) < 1e-2
    assert nchw(out)[:, cp:].abs().max().item() == 0 if cp < c else True
    assert rel(rm, bn.running_mean) < 1e-4 and rel(rv, bn.running_var) < 1e-4
    dout = rnd(n, cp, h, h)
    ref.backward(dout)
    dxa, dga, dba, *_ = ops.in_bwd(nhwc(F.pad(dout, (0, 0, 0, 0, 0, c - cp))), out, a, pooled, g, act=ops.ACT_RELU,
                                   c_params=cp if cp != c else None, batch=True)
    assert rel(nchw(dxa)[:, :cp], xr.grad) < 2e-2
    assert rel(dga, bn.weight.grad) < 2e-2 and rel(dba, bn.bias.grad) < 2e-2
    # eval mode
    bn.eval()
    st = ops.bn_eval_stats(bn.running_mean, bn.running_var, n, h * h, c)
    out_e = ops.in_apply(a, st, g, b, act=ops.ACT_NONE, c_params=cp if cp != c else None)
    assert rel(nchw(out_e)[:, :cp], bn(x)) < 1e-2


@pytest.mark.parametrize("c,h,n", [(16, 32, 2), (64, 16, 3)])
def test_instance_norm_double_backward(ops, c, h, n):
    torch.manual_seed(7)
    x = bf(rnd(n, c, h, h) * 1.5 + 0.3).float().requires_grad_(True)
    g = (torch.rand(c, device=DEV) + 0.5).requires_grad_(True)
    b = torch.zeros(c, device=DEV, requires_grad=True)
    dy = rnd(n, c, h, h).requires_grad_(True)
    u = rnd(n, c, h, h)
    y = in_ref(x, g, b)
    (dx,) = torch.autograd.grad(y, x, dy, create_graph=True)
    g_dy, g_x, g_g = torch.autograd.grad(dx, (dy, x, g), u)
    xs = nhwc(x.detach())
    st = ops.in_stats(xs)
    k_dy, k_x, k_g = ops.in_bwd2(nhwc(u), nhwc(dy.detach()), xs, st, g.detach())
    assert rel(nchw(k_dy), g_dy) < 2e-2
    assert rel(nchw(k_x), g_x) < 3e-2
    assert rel(k_g, g_g) < 3e-2


def test_elementwise_and_colsum(ops):
    x, y = rnd(2, 16, 32, 32, seed=8), rnd(2, 16, 32, 32)
    a, b = nhwc(x), nhwc(y)
    assert rel(nchw(ops.act_fwd(a, ops.ACT_LRELU)), F.leaky_relu(x, 0.01)) < 1e-2
    assert rel(nchw(ops.add_bf16(a, b)), x + y) < 1e-2
    ref = y * torch.where(x > 0, 1.0, 0.01) + x
    assert rel(nchw(ops.act_bwd(b, a, add=a, act=ops.ACT_LRELU)), ref) < 1e-2
    m = rnd(1024, 256)
    assert rel(ops.colsum(bf(m)), m.sum(0)) < 1e-3


@pytest.mark.parametrize("c,h,n", [(16, 64, 2), (128, 16, 3)])
def test_pool_and_resample(ops, c, h, n):
    torch.manual_seed(9)
    x = rnd(n, c, h, h)
    a = nhwc(x)
    xr = x.clone().requires_grad_(True)
    mp = F.max_pool2d(xr, 2, 2)
    assert torch.equal(nchw(ops.maxpool2_fwd(a)), mp.detach())
    dy = rnd(n, c, h // 2, h // 2)
    mp.backward(dy)
    assert rel(nchw(ops.maxpool2_bwd(a, nhwc(dy))), xr.grad) < 1e-6
    assert rel(nchw(ops.maxpool2_bwd(a, nhwc(dy), add=a)), xr.grad + x) < 1e-2
    assert rel(nchw(ops.avgpool2_fwd(a)), F.avg_pool2d(x, 2)) < 1e-2
    assert rel(nchw(ops.avgpool2_bwd(nhwc(dy))), 0.25 * F.interpolate(dy, scale_factor=2, mode="nearest")) < 1e-2
    xr2 = x.clone().requires_grad_(True)
    up = F.interpolate(xr2, scale_factor=2, mode="bilinear", align_corners=False)
    assert rel(nchw(ops.bilinear2_fwd(a)), up) < 1e-2
    d2 = rnd(n, c, 2 * h, 2 * h)
    up.backward(d2)
    assert rel(nchw(ops.bilinear2_bwd(nhwc(d2))), xr2.grad) < 1e-2


def test_maxpool_ties_route_to_first(ops):
    x = torch.zeros(1, 8, 4, 4, device=DEV)
    dy = torch.ones(1, 8, 2, 2, device=DEV)
    xr = x.clone().requires_grad_(True)
    F.max_pool2d(xr, 2, 2).backward(dy)
    assert torch.equal(nchw(ops.maxpool2_bwd(nhwc(x), nhwc(dy))), xr.grad)


def test_layout_kernels(ops):
    x = torch.randn(3, 5, 16, 16, device=DEV)
    y = ops.nchw_to_nhwc(x, 8)
    assert torch.equal(y[..., :5], nhwc(x)) and y[..., 5:].abs().max().item() == 0
    assert torch.equal(ops.nhwc_to_nchw(y, 5), bf(x).float())
    img = torch.randn(3, 1, 16, 16, device=DEV)
    m = torch.tensor([[1., 0, -1, 0], [0, 0, 0, 0], [-1, 1, 0, 0]], device=DEV)
    t = ops.build_tsl_input(img, m, 8)
    ref = torch.cat([img, m.view(3, 4, 1, 1).repeat(1, 1, 16, 16)], 1)
    assert torch.equal(t[..., :5], nhwc(ref)) and t[..., 5:].abs().max().item() == 0


def dice_ce_ref(logits, labels, w_ce=0.5, w_dc=0.5):
    # restates misc/loss.py:16-63 (batch dice, background dropped, double epsilon)
    p = torch.softmax(logits, 1)
    oh = F.one_hot(labels, logits.shape[1]).permute(0, 3, 1, 2).float()
    tp, fp, fn = (p * oh).sum((0, 2, 3)), (p * (1 - oh)).sum((0, 2, 3)), ((1 - p) * oh).sum((0, 2, 3))
    dc = (2 * tp + 1e-5) / (2 * tp + fp + fn + 1e-5 + 1e-8)
    return w_dc * (1 - dc[1:].mean()) + w_ce * F.cross_entropy(logits, labels)


@pytest.mark.parametrize("pseudo", [False, True])
def test_dice_ce(ops, pseudo):
    torch.manual_seed(10)
    n, c, h = 4, 5, 64
    logits = (torch.randn(n, c, h, h, device=DEV) * 2).requires_grad_(True)
    other = torch.randn(n, c, h, h, device=DEV)
    labels = other.argmax(1) if pseudo else torch.randint(0, c, (n, h, h), device=DEV)
    ref = dice_ce_ref(logits, labels)
    ref.backward()
    lg = logits.detach().permute(0, 2, 3, 1).reshape(-1, c).contiguous()
    ol = other.permute(0, 2, 3, 1).reshape(-1, c).contiguous() if pseudo else None
    lab = None if pseudo else labels.reshape(-1).contiguous()
    acc = torch.zeros(3 * c + 1, device=DEV)
    ops.dice_ce_fwd(lg, lab, ol, acc)
    loss = ops.dice_ce_finish(acc, lg.shape[0], c, 0.5, 0.5)
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    one = torch.ones(1, device=DEV)
    d = ops.dice_ce_bwd(lg, lab, ol, acc, one, 1.0, lg.shape[0], 0.5, 0.5)
    assert rel(d.view(n, h, h, c).permute(0, 3, 1, 2), logits.grad) < 1e-3
    if pseudo:
        assert torch.equal(ops.argmax_c(ol), labels.reshape(-1))


def test_small_losses(ops):
    torch.manual_seed(11)
    a = torch.randn(4, 1, 32, 32, device=DEV, requires_grad=True)
    b = torch.randn(4, 1, 32, 32, device=DEV)
    out = torch.zeros(1, device=DEV)
    ops.l1_fwd(a.detach(), b, out, 1.0 / a.numel())
    ref = (a - b).abs().mean()
    ref.backward()
    assert abs(out.item() - ref.item()) < 1e-5
    one = torch.ones(1, device=DEV)
    assert rel(ops.l1_bwd(a.detach(), b, one, 1.0 / a.numel()), a.grad) < 1e-6
    # modality CE
    z = torch.randn(16, 4, device=DEV, requires_grad=True)
    t = torch.randint(0, 4, (16,), device=DEV)
    out.zero_()
    ops.ce_rows_fwd(z.detach(), t, out, 1.0)
    ref = F.cross_entropy(z, t)
    ref.backward()
    assert abs(out.item() - ref.item()) < 1e-5
    assert rel(ops.ce_rows_bwd(z.detach(), t, one, 1.0), z.grad) < 1e-5
    # gradient penalty (trainer/uganShp0Trainer.py:127-134)
    g = torch.randn(8, 1, 32, 32, device=DEV, requires_grad=True)
    out.zero_()
    norm = ops.gp_fwd(g.detach(), out, 1.0)
    ref = ((g.view(8, -1).pow(2).sum(1).sqrt() - 1) ** 2).mean()
    ref.backward()
    assert abs(out.item() - ref.item()) < 1e-4 * ref.item()
    assert rel(ops.gp_bwd(g.detach(), norm, one, 1.0), g.grad) < 1e-5
    # mean
    out.zero_()
    ops.sum_f32(b, out, 1.0 / b.numel())
    assert abs(out.item() - b.mean().item()) < 1e-5


def test_patchnce_pipeline(ops):
    """gather -> L2 normalise -> PatchNCE (network/ugan.py:316-334, network/patchnce.py:13-51)."""
    torch.manual_seed(12)
    n, c, hw, nid, groups = 4, 256, 16, 64, 2
    feat = rnd(n, c, hw, hw)
    ids = torch.randperm(hw * hw, device=DEV)[:nid]
    rows = ops.gather_rows(nhwc(feat), ids)
    ref_rows = feat.permute(0, 2, 3, 1).flatten(1, 2)[:, ids, :].flatten(0, 1)
    assert torch.equal(rows.float(), ref_rows)
    d = torch.zeros(n, hw, hw, c, device=DEV, dtype=torch.bfloat16)
    ops.scatter_rows_add(rows, ids, d)
    refd = torch.zeros(n, hw * hw, c, device=DEV)
    refd[:, ids, :] = ref_rows.view(n, nid, c)
    assert torch.equal(d.float().view(n, hw * hw, c), refd)

    q0 = torch.randn(n * nid, c, device=DEV, requires_grad=True)
    k0 = torch.randn(n * nid, c, device=DEV)
    qn = q0 / (q0.pow(2).sum(1, keepdim=True).sqrt() + 1e-7)
    kn = k0 / (k0.pow(2).sum(1, keepdim=True).sqrt() + 1e-7)
    y, norm = ops.l2norm_fwd(q0.detach())
    assert rel(y, qn) < 1e-5
    # reference PatchNCE restated
    B = n * nid
    l_pos = (qn * kn).sum(1, keepdim=True)
    qg, kg = qn.view(groups, -1, c), kn.view(groups, -1, c)
    npg = qg.shape[1]
    l_neg = torch.bmm(qg, kg.transpose(2, 1))
    eye = torch.eye(npg, device=DEV, dtype=torch.bool)[None]
    l_neg = l_neg.masked_fill(eye, -10.0).view(-1, npg)
    outl = torch.cat((l_pos, l_neg), 1) / 0.07
    ref_rows_loss = F.cross_entropy(outl, torch.zeros(B, dtype=torch.long, device=DEV), reduction="none")
    ref = ref_rows_loss.mean()
    ref.backward()
    out = torch.zeros(1, device=DEV)
    kk, _ = ops.l2norm_fwd(k0)
    lr = ops.patchnce_fwd(y, kk, groups, npg, 1 / 0.07, out, 1.0)
    assert rel(lr, ref_rows_loss) < 1e-4 and abs(out.item() - ref.item()) < 1e-4
    one = torch.ones(1, device=DEV)
    dq = ops.patchnce_bwd(y, kk, groups, npg, 1 / 0.07, one, 1.0)
    dx = ops.l2norm_bwd(dq, y, norm)
    assert rel(dx, q0.grad) < 1e-2


def test_optimisers(ops):
    torch.manual_seed(13)
    n = 100003
    p0, g = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    lr = torch.tensor([1e-2], device=DEV)
    # SGD momentum 0.9 wd 1e-3, two steps
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], lr=1e-2, momentum=0.9, weight_decay=1e-3)
    p, mom = p0.clone(), torch.zeros(n, device=DEV)
    for _ in range(2):
        pr.grad = g.clone()
        opt.step()
        ops.sgd_step(p, g, mom, lr, 0.9, 1e-3)
    assert rel(p, pr.detach()) < 1e-6
    # Adam with L2 weight decay, three steps
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-2, betas=(0.9, 0.999), weight_decay=1e-3)
    p, m, v, st = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(1, device=DEV)
    for _ in range(3):
        pr.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, lr, 0.9, 0.999, 1e-8, 1e-3, st)
    assert rel(p, pr.detach()) < 1e-5
    # EMA
    ema, alpha = torch.randn(n, device=DEV), torch.tensor([0.99], device=DEV)
    ref = 0.99 * ema + 0.01 * p0
    ops.ema_update(ema, p0, alpha)
    assert rel(ema, ref) < 1e-6
    # poly LR: lr used at step k+1 is base*(1-k/max)^0.9
    it, lro = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    got = []
    for _ in range(4):
        ops.poly_lr_tick(it, lro, 1e-2, 30000.0, 0.9)
        got.append(lro.item())
    want = [1e-2 * (1 - max(k - 1, 0) / 30000.0) ** 0.9 for k in range(4)]
    assert max(abs(a - b) for a, b in zip(got, want)) < 1e-8


@pytest.mark.parametrize("cin", [1, 5])
def test_stem_5x5_on_tensor_cores(ops, cin):
    """5x5 stem (network/ugan.py:26): Cin in {1,5} zero-padded to 16 channels, Cout = 8 padded to 16"""
    n, h, cout = 2, 128, 8
    x = rnd(n, cin, h, h, seed=14)
    wt = rnd(cout, cin, 5, 5, scale=0.1)
    pw = make_pack(ops, wt)
    assert pw.cin_pad == 16 and pw.cout_pad == 16
    xp = nhwc(F.pad(x, (0, 0, 0, 0, 0, 16 - cin)))
    y = ops.conv_fprop([xp], pw)
    y_ref = F.conv2d(x, wt, padding=2)
    assert rel(nchw(y)[:, :cout], y_ref) < 1e-2 and y[..., cout:].abs().max().item() == 0
    dy = F.pad(rnd(n, cout, h, h), (0, 0, 0, 0, 0, 16 - cout))
    dx = nchw(ops.conv_dgrad(nhwc(dy), pw)[0])
    assert rel(dx[:, :cin], torch.nn.grad.conv2d_input(x.shape, wt, dy[:, :cout], padding=2)) < 1e-2
    assert dx[:, cin:].abs().max().item() == 0
    dw = ops.conv_wgrad([xp], nhwc(dy), pw)
    assert dw.shape == wt.shape
    assert rel(dw, torch.nn.grad.conv2d_weight(x, wt.shape, dy[:, :cout], padding=2)) < 1e-2
    acc = torch.ones_like(wt)
    ops.conv_wgrad([xp], nhwc(dy), pw, out=acc)          # accumulate-into-grad form
    assert rel(acc - 1, dw) < 1e-5


@pytest.mark.parametrize("cout,tanh", [(5, False), (1, True)])
def test_fused_head_backward(ops, cout, tanh):
    torch.manual_seed(15)
    n, h = 2, 64
    x = rnd(n, 16, h, h)
    wt = torch.randn(cout, 16, 1, 1, device=DEV) * 0.3
    b = torch.randn(cout, device=DEV)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = F.conv2d(xr, wr, br)
    y_ref = torch.tanh(y_ref) if tanh else y_ref
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    y = ops.conv_direct_fprop(nhwc(x), wt, 1, 0, bias=b, act=ops.ACT_TANH if tanh else ops.ACT_NONE, out_f32=True)
    assert rel(nchw(y), y_ref) < 1e-4
    dx, dw, db = ops.head1x1_bwd(nhwc(x), dy.permute(0, 2, 3, 1).contiguous(), y if tanh else None, wt, True,
                                 want_bias=True)
    assert rel(nchw(dx), xr.grad) < 1e-2 and rel(dw, wr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("cins,cout,h,n", [([16], 16, 256, 2), ([16, 16], 16, 256, 2), ([32], 32, 128, 2),
                                            ([16], 32, 128, 2), ([64], 64, 64, 2), ([64, 64], 64, 64, 3),
                                            ([128], 128, 32, 2), ([128, 128], 128, 32, 2), ([256], 256, 16, 3),
                                            ([256], 256, 8, 5), ([256], 256, 4, 16), ([64], 128, 32, 2)])
def test_conv_fused_instance_norm_statistics(ops, cins, cout, h, n):
    """statistics from the conv epilogue (band kernel: carried down the strip; conv_tc_kernel: per tile, also when a
    tile holds several whole images) or the statistics kernel == sums over the stored output"""
    torch.manual_seed(16)
    xs = [rnd(n, c, h, h) for c in cins]
    wt = rnd(cout, sum(cins), 3, 3, scale=0.1)
    pw = make_pack(ops, wt)
    y, st = ops.conv_fprop([nhwc(x) for x in xs], pw, want_stats=True)
    yf = y.float()
    assert rel(st[:, 0], yf.sum((1, 2))) < 1e-4 and rel(st[:, 1], (yf * yf).sum((1, 2))) < 1e-4


def test_confusion_counts_bit_exact(ops):
    """validation metric path: conf[label, argmax] counts equal torch's, ties go to the first maximum, labels outside
    [0, C) are ignored, repeated calls accumulate"""
    torch.manual_seed(21)
    npix, c = 3 * 97 * 101, 5
    logits = torch.randn(npix, c, device=DEV)
    logits[::7] = logits[::7].round()            # plenty of exact ties
    logits[::11, 1] = logits[::11, 3]
    labels = torch.randint(-1, c + 1, (npix,), device=DEV)
    conf = torch.zeros(c, c, dtype=torch.int64, device=DEV)
    ops.confusion_counts(logits, labels, conf)
    ops.confusion_counts(logits, labels, conf)
    pred = logits.argmax(1)
    ok = (labels >= 0) & (labels < c)
    ref = torch.zeros(c * c, dtype=torch.int64, device=DEV)
    ref.index_add_(0, labels[ok] * c + pred[ok], torch.ones_like(pred[ok]))
    assert torch.equal(conf.view(-1), 2 * ref)


def test_conv_tc_cluster_split_k(ops, monkeypatch):
    """opt-in split-K over a thread-block cluster with a DSMEM reduce-scatter (SMSUT_TC_KSPLIT): same results as the
    un-split kernel, fused statistics included.  (The knob is read once per process: run in a subprocess.)"""
    import os
    import subprocess
    import sys
    code = """
import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as g; g.load_package()
from smsut_b200 import ops
import torch.nn.functional as F
torch.manual_seed(3)
for cin, cout, h, n in ((256, 256, 16, 16), (128, 256, 16, 16), (256, 256, 8, 16)):
    x = torch.randn(n, cin, h, h, device='cuda').to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device='cuda') * (2.0 / (cin * 9)) ** 0.5).to(torch.bfloat16).float()
    pw = ops.PackedWeight(w); ops.PackTable([pw]).refresh()
    y, st = ops.conv_fprop([x.permute(0, 2, 3, 1).contiguous()], pw, want_stats=True)
    ref = F.conv2d(x.float(), w, padding=1)
    got = y.permute(0, 3, 1, 2).float()
    assert ((got - ref).norm() / ref.norm()).item() < 1e-2
    yf = y.float()
    assert ((st[:, 0] - yf.sum((1, 2))).norm() / yf.sum((1, 2)).norm()).item() < 1e-4
print('ok')
"""
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SMSUT_TC_KSPLIT="4"), capture_output=True,
                       text=True, timeout=300, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


# ---- coraNet losses (csrc/coranet.cu; trainer/coraNetTrainer.py) ---------------------------------------------------
@pytest.mark.parametrize("npix,nlab", [(4096, 4), (50001, 4), (777, 1)])
def test_coranet_heads_split(ops, npix, nlab):
    """(npix, 1 + 3L) -> three (npix, 1 + L) heads sharing channel 0 (`torch.cat([out_back, out_h], 1)`,
    coraNetTrainer.py:279-286), bit-exact; the backward sums the three background gradients."""
    z = torch.randn(npix, 1 + 3 * nlab, device=DEV)
    h = ops.heads_split_fwd(z, nlab, 3)
    ref = torch.stack([torch.cat([z[:, :1], z[:, 1 + k * nlab:1 + (k + 1) * nlab]], 1) for k in range(3)])
    assert torch.equal(h, ref)
    d = torch.randn_like(h)
    dz = ops.heads_split_bwd(d, nlab, 3)
    zr = z.clone().requires_grad_(True)
    torch.stack([torch.cat([zr[:, :1], zr[:, 1 + k * nlab:1 + (k + 1) * nlab]], 1) for k in range(3)]).backward(d)
    assert rel(dz, zr.grad) < 1e-6


@pytest.mark.parametrize("npix,c,weighted,masked", [(65536, 5, True, False), (65536, 5, False, True), (30011, 5, True, True),
                                                     (4099, 2, True, True), (131072, 5, False, False)])
def test_coranet_weighted_masked_ce(pkg, ops, npix, c, weighted, masked):
    """nn.CrossEntropyLoss(weight) in 'mean' form and the masked `(CE_none * mask).sum() / (mask.sum() + 1e-16)` form
    (coraNetTrainer.py:44-58,301-303), forward and backward, vs PyTorch fp32"""
    from smsut_b200 import functional as Fn
    torch.manual_seed(npix)
    z = (torch.randn(npix, c, device=DEV) * 3).requires_grad_(True)
    y = torch.randint(0, c, (npix,), device=DEV)
    cw = (torch.rand(c, device=DEV) * 4 + 0.5) if weighted else None
    mask = (torch.rand(npix, device=DEV) > 0.4).float() if masked else None
    loss = Fn.WeightedCEFn.apply(z, y, cw, mask)
    zr = z.detach().clone().requires_grad_(True)
    if masked:
        ref = (F.cross_entropy(zr, y, weight=cw, reduction="none") * mask).sum() / (mask.sum() + 1e-16)
    else:
        ref = F.cross_entropy(zr, y, weight=cw)
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item())), (loss.item(), ref.item())
    (loss * 1.7).backward()
    (ref * 1.7).backward()
    assert rel(z.grad, zr.grad) < 1e-5
    if masked:      # an empty mask gives 0 / 1e-16 = 0 and a zero gradient, like the reference's expression
        z2 = z.detach().clone().requires_grad_(True)
        l0 = Fn.WeightedCEFn.apply(z2, y, cw, torch.zeros_like(mask))
        l0.backward()
        assert l0.item() == 0.0 and z2.grad.abs().max().item() == 0.0


@pytest.mark.parametrize("npix,c,invert", [(65536, 5, True), (30011, 5, False), (4099, 2, True)])
def test_coranet_masked_softmax_mse(pkg, ops, npix, c, invert):
    """`(softmax_mse_loss(zs, zt) * m).sum() / (m.sum() + 1e-16)`, m = (1 - mask) broadcast over the classes
    (coraNetTrainer.py:137-149,327-337), forward and gradient to the student logits, vs PyTorch fp32"""
    from smsut_b200 import functional as Fn
    torch.manual_seed(npix + 1)
    zs = (torch.randn(npix, c, device=DEV) * 2).requires_grad_(True)
    zt = torch.randn(npix, c, device=DEV) * 2
    mask = (torch.rand(npix, device=DEV) > 0.6).float()
    loss = Fn.SoftmaxMSEMaskedFn.apply(zs, zt, mask, invert)
    zr = zs.detach().clone().requires_grad_(True)
    m = (1 - mask) if invert else mask
    ref = (((torch.softmax(zr, 1) - torch.softmax(zt, 1)) ** 2) * m[:, None]).sum() / (m.sum() + 1e-16)
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1e-3, abs(ref.item())), (loss.item(), ref.item())
    (loss * 0.3).backward()
    (ref * 0.3).backward()
    assert rel(zs.grad, zr.grad) < 1e-5
