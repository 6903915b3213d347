"""ORACLE -- test infrastructure only (imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product path).

A plain PyTorch fp32 restatement of the arithmetic of the SMSUT hot path, written functionally over the
reference's own `state_dict` layout (the compatibility contract) instead of as nn.Modules.
Each function cites the reference file:line it follows.  The reference is pure Python on top of PyTorch (an
un-vendored third-party dependency with no pinned version), and ships no tests or golden vectors, so parity is
pinned by us: tests/golden/make_golden.py imports the *real* reference modules from /root/reference in the
build container, runs them on seeded inputs with the deterministic weights of `make_weights`, and commits the
outputs under tests/golden/; tests/test_oracle.py checks this restatement against those fixtures.
"""
import math

import torch
import torch.nn.functional as F

SLOPE = 0.01  # LeakyReLU slope everywhere on the path (network/blocks.py:28-32, network/ugan.py:203)
EPS = 1e-5    # InstanceNorm2d eps (network/blocks.py:22-23)


# --------------------------------------------------------------------------------------------------
# deterministic weights shared by the golden generator, the oracle tests and the GPU parity tests
# --------------------------------------------------------------------------------------------------
def unet_shapes(in_ch=1, out_ch=5, w=16):
    """state_dict layout of network/unet.py:13-32 + network/blocks.py:120-174 (norm='instance')."""
    s = {}

    def block(p, cin, cout):
        s[p + "conv1.weight"] = (cout, cin, 3, 3)
        s[p + "bn1.weight"] = (cout,); s[p + "bn1.bias"] = (cout,)
        s[p + "conv2.weight"] = (cout, cout, 3, 3)
        s[p + "bn2.weight"] = (cout,); s[p + "bn2.bias"] = (cout,)
        if cin != cout:
            s[p + "shortcut1.weight"] = (cout, cin, 1, 1)
            s[p + "shortcut2.weight"] = (cout,); s[p + "shortcut2.bias"] = (cout,)

    s["encoder.pre_conv.weight"] = (w // 2, in_ch, 5, 5)
    s["encoder.pre_bn.weight"] = (w // 2,); s["encoder.pre_bn.bias"] = (w // 2,)
    chans = [w // 2, w, 2 * w, 4 * w, 8 * w, 16 * w]
    for i in range(1, 6):
        block(f"encoder.layer{i}.", chans[i - 1], chans[i])
    for i in (4, 3, 2, 1):
        s[f"decoder.up{i}.up.weight"] = (chans[i + 1], chans[i], 2, 2)
        block(f"decoder.layer{i}.", chans[i + 1], chans[i])
    s["decoder.fc.weight"] = (out_ch, w, 1, 1)
    return s


def ugan_shapes(in_ch=1, out_ch=5, n_modal=4, w=16, nc=256):
    """state_dict layout of network/ugan.py:126-151 (UGANnce)."""
    s = {}

    def block(p, cin, cout):
        s[p + "conv1.weight"] = (cout, cin, 3, 3)
        s[p + "bn1.weight"] = (cout,); s[p + "bn1.bias"] = (cout,)
        s[p + "conv2.weight"] = (cout, cout, 3, 3)
        s[p + "bn2.weight"] = (cout,); s[p + "bn2.bias"] = (cout,)
        if cin != cout:
            s[p + "shortcut1.weight"] = (cout, cin, 1, 1)
            s[p + "shortcut2.weight"] = (cout,); s[p + "shortcut2.bias"] = (cout,)

    chans = [w // 2, w, 2 * w, 4 * w, 8 * w, 16 * w]
    for enc, cin in (("tsl_encoder.", in_ch + n_modal), ("seg_encoder.", in_ch)):
        s[enc + "pre.0.weight"] = (w // 2, cin, 5, 5)
        s[enc + "pre.1.weight"] = (w // 2,); s[enc + "pre.1.bias"] = (w // 2,)
        for i in range(1, 5):
            block(f"{enc}enc{i}.", chans[i - 1], chans[i])
    block("enc5.", chans[4], chans[5])
    s["netF.mlp_0.0.weight"] = (nc, 16 * w); s["netF.mlp_0.0.bias"] = (nc,)
    s["netF.mlp_0.2.weight"] = (nc, nc); s["netF.mlp_0.2.bias"] = (nc,)
    for dec, oc, transposed in (("tsl_decoder.", 1, False), ("seg_decoder.", out_ch, True)):
        for i in (4, 3, 2, 1):
            if transposed:
                s[f"{dec}up{i}.up.weight"] = (chans[i + 1], chans[i], 2, 2)
            else:
                s[f"{dec}up{i}.up.1.weight"] = (chans[i], chans[i + 1], 1, 1)
            block(f"{dec}dec{i}.", chans[i + 1], chans[i])
        s[dec + "fc.weight"] = (oc, w, 1, 1); s[dec + "fc.bias"] = (oc,)
    return s


def disc_shapes(input_size=256, n_modal=4, w=16, max_width=256):
    """state_dict layout of network/ugan.py:198-215 (Discriminator)."""
    s = {"main.0.weight": (w, 1, 4, 4), "main.0.bias": (w,)}
    repeat = int(math.log2(input_size)) - 2
    cin = w
    for i in range(1, repeat):
        cout = min(cin * 2, max_width)
        p = f"main.{i + 1}."
        s[p + "conv1.weight"] = (cout, cin, 3, 3)
        s[p + "bn1.weight"] = (cout,); s[p + "bn1.bias"] = (cout,)
        s[p + "conv2.weight"] = (cout, cout, 3, 3)
        s[p + "bn2.weight"] = (cout,); s[p + "bn2.bias"] = (cout,)
        if cin != cout:
            s[p + "downsample.0.weight"] = (cout, cin, 1, 1)
            s[p + "downsample.1.weight"] = (cout,); s[p + "downsample.1.bias"] = (cout,)
        cin = cout
    k = int(input_size / 2 ** repeat)
    s["conv_src.weight"] = (1, cin, 3, 3)
    s["conv_cls.weight"] = (n_modal, cin, k, k)
    return s


def make_weights(shapes, seed, device="cpu"):
    """Deterministic weights for a shape table: conv-like tensors ~ N(0, 2/fan_out) (the reference's
    kaiming_normal_(fan_out) scale, network/ugan.py:145-151), affine gains 1 +- 0.1, biases +- 0.1, so that every
    parameter matters in a parity check.  Generated on CPU (identical on every machine), then moved."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if len(shp) == 4:
            fan_out = shp[0] * shp[2] * shp[3]
            if ".up.weight" in k:  # ConvTranspose2d (Cin, Cout, 2, 2): fan_out counts dim 0 in PyTorch
                fan_out = shp[0] * shp[2] * shp[3]
            t = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_out)
        elif len(shp) == 2:
            t = torch.randn(shp, generator=g) * 0.02
        elif k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.1 * torch.randn(shp, generator=g)
        out[k] = t.to(device)
    return out


def synthetic_batch(n, size, seed, n_label=4, device="cpu"):
    """Abdominal-like synthetic slices: body ellipse + 4 organ ellipses, u8-quantised,
    normalised to [-1, 1] like ToTensor+Normalize(0.5, 0.5) (data_loader/baseLoader.py:89).  Returns
    image (n,1,size,size) fp32 and labels (n,size,size) int64 in {0..n_label}."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
    img = torch.zeros(n, 1, size, size)
    lab = torch.zeros(n, size, size, dtype=torch.int64)
    organs = [(-0.35, -0.15, 0.30, 0.22), (-0.25, 0.35, 0.10, 0.13), (0.25, 0.35, 0.10, 0.13), (0.42, -0.10, 0.14, 0.18)]
    for i in range(n):
        j = torch.rand(12, generator=g) * 0.1 - 0.05
        body = ((xx / (0.85 + j[0])) ** 2 + (yy / (0.65 + j[1])) ** 2) < 1
        v = torch.full((size, size), 0.05)
        v[body] = 0.35 + j[2].item()
        for c, (cx, cy, rx, ry) in enumerate(organs[:n_label]):
            m = (((xx - cx - j[3 + c]) / rx) ** 2 + ((yy - cy - j[7 + c]) / ry) ** 2) < 1
            m &= body
            v[m] = 0.5 + 0.1 * c + j[11].item()
            lab[i][m] = c + 1
        v = v + 0.05 * torch.randn(size, size, generator=g)
        u8 = (v.clamp(0, 1) * 255).round()
        img[i, 0] = (u8 / 255 - 0.5) / 0.5
    return img.to(device), lab.to(device)


# --------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------
def inorm(x, sd, p):
    # nn.InstanceNorm2d(C, affine=True): biased variance, eps 1e-5, no running stats (network/blocks.py:22-23)
    return F.instance_norm(x, weight=sd[p + "weight"], bias=sd[p + "bias"], eps=EPS)


def lrelu(x):
    return F.leaky_relu(x, SLOPE)


class Style:
    """norm / activation choice of a network (network/blocks.py:19-34 get_norm / get_act).  The trainers all use
    ('instance', 'lrelu'); ('batch', 'relu') is the default of the UNet / Encoder / Decoder signatures
    (network/unet.py:14).  BatchNorm keeps its running estimates in `sd` under the nn.BatchNorm2d buffer names."""

    def __init__(self, norm="instance", act="lrelu", training=True, momentum=0.1):
        self.norm_type, self.act_type, self.training, self.momentum = norm, act, training, momentum

    # hooks of the per-layer parity protocol; no-ops here, see RecordingStyle
    def tap(self, key, x):
        return x

    def pool(self, x, key=None):
        return F.max_pool2d(x, 2, 2)

    def norm(self, x, sd, p):
        if self.norm_type == "instance":
            return inorm(x, sd, p)
        # nn.BatchNorm2d(C): batch statistics + running-estimate update in training, running estimates in eval
        return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                            self.training, self.momentum, EPS)

    def act(self, x, key=None):
        return lrelu(x) if self.act_type == "lrelu" else F.relu(x)


class RecordingStyle(Style):
    """Style that (i) records every tapped tensor (block inputs / outputs, retained for .grad), every activation's
    mask (pre-activation > 0) and every max-pool's argmax under the block's key, and (ii) optionally FORCES masks and
    pool indices taken from another run.  LeakyReLU / ReLU / max-pool are the only non-smooth ops of the path: with
    their selections forced, the network is the same piecewise-linear-in-the-selections function on both sides, so a
    gradient comparison measures arithmetic, not which side of a kink a bf16-rounded value fell on.
    forced_masks / forced_pool: dict key -> bool mask (NCHW) / int64 indices as returned by max_pool2d(return_indices),
    or a list consumed in call order when keys are not known."""

    def __init__(self, norm="instance", act="lrelu", training=True, momentum=0.1, forced_masks=None, forced_pool=None):
        super().__init__(norm, act, training, momentum)
        self.taps, self.masks, self.pools = {}, {}, {}
        self.forced_masks, self.forced_pool = forced_masks, forced_pool
        self._n = 0

    def tap(self, key, x):
        if x.requires_grad:
            x.retain_grad()
        self.taps[key] = x
        return x

    def _forced(self, table, key):
        if table is None:
            return None
        if isinstance(table, dict):
            return table.get(key)
        return table.pop(0) if table else None

    def act(self, x, key=None):
        key = key if key is not None else f"act{self._n}"
        self._n += 1
        m = self._forced(self.forced_masks, key)
        if m is None:
            m = x > 0
        self.masks[key] = m.detach()
        neg = SLOPE if self.act_type == "lrelu" else 0.0
        return torch.where(m, x, x * neg)

    def pool(self, x, key=None):
        key = key if key is not None else f"pool{self._n}"
        idx = self._forced(self.forced_pool, key)
        if idx is None:
            y, idx = F.max_pool2d(x, 2, 2, return_indices=True)
            self.pools[key] = idx
            return y
        self.pools[key] = idx
        return x.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)


DEFAULT_STYLE = Style()


def add_bn_buffers(sd):
    """running_mean / running_var / num_batches_tracked entries of nn.BatchNorm2d for every affine norm in `sd`."""
    out = dict(sd)
    for k, v in sd.items():
        if v.dim() == 1 and k.endswith(".bias") and k[:-4] + "weight" in sd and sd[k[:-4] + "weight"].dim() == 1:
            p = k[:-4]
            out[p + "running_mean"] = torch.zeros_like(v)
            out[p + "running_var"] = torch.ones_like(v)
            out[p + "num_batches_tracked"] = torch.zeros((), dtype=torch.int64, device=v.device)
    return out


def basic_block(x, sd, p, st=DEFAULT_STYLE, tag=None):
    # network/blocks.py:66-80.  tag: key prefix of the parity hooks (defaults to the state_dict prefix; enc5 is
    # called once per branch with the same weights and needs distinct keys)
    k = p if tag is None else tag
    x = st.tap(k + "in", x)
    y = st.act(st.norm(F.conv2d(x, sd[p + "conv1.weight"], padding=1), sd, p + "bn1."), k + "act1")
    y = st.norm(F.conv2d(y, sd[p + "conv2.weight"], padding=1), sd, p + "bn2.")
    if p + "shortcut1.weight" in sd:
        x = st.norm(F.conv2d(x, sd[p + "shortcut1.weight"]), sd, p + "shortcut2.")
    return st.tap(k + "out", st.act(y + x, k + "act2"))


def bottle_block(x, sd, p, st=DEFAULT_STYLE):
    # network/blocks.py:99-117 (stride 2): both branches see avg_pool2d
    x = st.tap(p + "in", x)
    ident = F.avg_pool2d(x, 2)
    y = st.act(inorm(F.conv2d(x, sd[p + "conv1.weight"], padding=1), sd, p + "bn1."), p + "act1")
    y = F.avg_pool2d(y, 2)
    y = inorm(F.conv2d(y, sd[p + "conv2.weight"], padding=1), sd, p + "bn2.")
    if p + "downsample.0.weight" in sd:
        ident = inorm(F.conv2d(ident, sd[p + "downsample.0.weight"]), sd, p + "downsample.1.")
    return st.tap(p + "out", st.act(y + ident, p + "act2"))


def unet_forward(sd, x, taps=None, style=DEFAULT_STYLE):
    """network/unet.py:29-32 -> blocks.Encoder.forward (blocks.py:138-153) + blocks.Decoder.forward (:168-174)."""
    t = taps if taps is not None else {}
    st = style
    x = st.tap("encoder.pre.in", x)
    x = st.act(st.norm(F.conv2d(x, sd["encoder.pre_conv.weight"], padding=2), sd, "encoder.pre_bn."), "encoder.pre.act")
    t["encoder.pre"] = st.tap("encoder.pre.out", x)
    skips = []
    for i in range(1, 5):
        x = basic_block(x, sd, f"encoder.layer{i}.", st)
        t[f"encoder.layer{i}"] = x
        skips.append(x)
        x = st.pool(x, f"encoder.pool{i}")
    x = basic_block(x, sd, "encoder.layer5.", st)
    t["encoder.layer5"] = x
    for i in (4, 3, 2, 1):
        up = st.tap(f"decoder.up{i}.out", F.conv_transpose2d(st.tap(f"decoder.up{i}.in", x),
                                                             sd[f"decoder.up{i}.up.weight"], stride=2))
        x = basic_block(torch.cat([up, skips[i - 1]], 1), sd, f"decoder.layer{i}.", st)
        t[f"decoder.layer{i}"] = x
    return st.tap("decoder.fc.out", F.conv2d(st.tap("decoder.fc.in", x), sd["decoder.fc.weight"]))


def ugan_encoder(sd, p, x, t, st=DEFAULT_STYLE):
    # network/ugan.py:39-55
    x = st.tap(p + "pre.in", x)
    x = st.act(inorm(F.conv2d(x, sd[p + "pre.0.weight"], padding=2), sd, p + "pre.1."), p + "pre.act")
    t[p + "pre"] = st.tap(p + "pre.out", x)
    skips = []
    for i in range(1, 5):
        x = basic_block(x, sd, f"{p}enc{i}.", st)
        t[f"{p}enc{i}"] = x
        skips.append(x)
        x = st.pool(x, f"{p}pool{i}")
    skips.reverse()
    return x, skips


def ugan_decoder(sd, p, e5, skips, transposed, use_tanh, t, st=DEFAULT_STYLE):
    # network/ugan.py:76-83, network/blocks.py:37-50
    x = e5
    for k, i in enumerate((4, 3, 2, 1)):
        x = st.tap(f"{p}up{i}.in", x)
        if transposed:
            up = F.conv_transpose2d(x, sd[f"{p}up{i}.up.weight"], stride=2)
        else:
            up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
            up = F.conv2d(up, sd[f"{p}up{i}.up.1.weight"])
        up = st.tap(f"{p}up{i}.out", up)
        x = basic_block(torch.cat([up, skips[k]], 1), sd, f"{p}dec{i}.", st)
        t[f"{p}dec{i}"] = x
    out = F.conv2d(st.tap(p + "fc.in", x), sd[p + "fc.weight"], sd[p + "fc.bias"])
    return st.tap(p + "fc.out", torch.tanh(out) if use_tanh else out)


def l2_normalize(x):
    # network/networks.py:241-242
    return x / (x.pow(2).sum(1, keepdim=True).pow(0.5) + 1e-7)


def patch_sample(sd, feat, ids):
    # network/ugan.py:316-334: NHWC flatten, gather shared ids, 2-layer MLP, L2 normalise
    rows = feat.permute(0, 2, 3, 1).flatten(1, 2)[:, ids, :].flatten(0, 1)
    h = F.relu(F.linear(rows, sd["netF.mlp_0.0.weight"], sd["netF.mlp_0.0.bias"]))
    return l2_normalize(F.linear(h, sd["netF.mlp_0.2.weight"], sd["netF.mlp_0.2.bias"]))


def ugannce_forward(sd, x, m=None, sample_ids=None, val_phase=False, taps=None, style=DEFAULT_STYLE):
    """network/ugan.py:153-195.  `sample_ids` must be given when features are wanted (the reference draws
    torch.randperm(H*W)[:64] itself, ugan.py:321-322; the oracle takes the draw as an input)."""
    t = taps if taps is not None else {}
    n_modal = sd["tsl_encoder.pre.0.weight"].shape[1] - x.shape[1]
    if m is None:
        m = torch.zeros(x.size(0), n_modal, device=x.device)
    planes = m.view(m.size(0), m.size(1), 1, 1).repeat(1, 1, x.size(2), x.size(3))
    st = style
    tsl_out, tsl_skips = ugan_encoder(sd, "tsl_encoder.", torch.cat([x, planes], 1), t, st)
    tsl_e5 = basic_block(tsl_out, sd, "enc5.", st, tag="tsl.enc5.")
    t["tsl.enc5"] = tsl_e5
    tsl = ugan_decoder(sd, "tsl_decoder.", tsl_e5, tsl_skips, False, True, t, st)
    seg_out, seg_skips = ugan_encoder(sd, "seg_encoder.", x, t, st)
    seg_e5 = basic_block(seg_out, sd, "enc5.", st, tag="seg.enc5.")
    t["seg.enc5"] = seg_e5
    seg = ugan_decoder(sd, "seg_decoder.", seg_e5, seg_skips, True, False, t, st)
    if val_phase:
        return seg, tsl
    return seg, tsl, [patch_sample(sd, tsl_e5, sample_ids[0])], sample_ids


def discriminator_forward(sd, x, taps=None, style=DEFAULT_STYLE):
    """network/ugan.py:225-229."""
    t = taps if taps is not None else {}
    st = style
    x = st.tap("main.0.in", x)
    out = st.act(F.conv2d(x, sd["main.0.weight"], sd["main.0.bias"], stride=2, padding=1), "main.0.act")
    t["main.0"] = st.tap("main.0.out", out)
    i = 2
    while f"main.{i}.conv1.weight" in sd:
        out = bottle_block(out, sd, f"main.{i}.", st)
        t[f"main.{i}"] = out
        i += 1
    out = st.tap("heads.in", out)
    out_src = F.conv2d(out, sd["conv_src.weight"], padding=1)
    out_cls = F.conv2d(out, sd["conv_cls.weight"])
    return out_src, out_cls.view(out_cls.size(0), out_cls.size(1))


# --------------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------------
def dice_ce_loss(x, y, weight_ce=0.5, weight_dc=0.5, dice_stats_hook=None):
    """misc/loss.py:16-63 with batch_dice=True (trainer/baseTrainer.py:57), background dropped, smooth 1e-5 and
    the extra 1e-8 in the denominator.  `dice_stats_hook` (tp, fp, fn) -> (tp, fp, fn) lets a data-parallel test
    sum the statistics over ranks, as DataParallel's gathered batch does in the reference."""
    p = F.softmax(x, 1)
    oh = torch.zeros_like(p).scatter_(1, y.unsqueeze(1), 1.0)
    tp = (p * oh).sum((0, 2, 3)); fp = (p * (1 - oh)).sum((0, 2, 3)); fn = ((1 - p) * oh).sum((0, 2, 3))
    if dice_stats_hook is not None:
        tp, fp, fn = dice_stats_hook(tp, fp, fn)
    dc = (2 * tp + 1e-5) / (2 * tp + fp + fn + 1e-5 + 1e-8)
    return weight_dc * (1.0 - dc[1:].mean()) + weight_ce * F.cross_entropy(x, y)


def patchnce_loss(feat_q, feat_k, batch_size, T=0.07):
    """network/patchnce.py:13-51; returns the per-row loss (N,)."""
    n, dim = feat_q.shape
    feat_k = feat_k.detach()
    l_pos = (feat_q * feat_k).sum(1, keepdim=True)
    q = feat_q.view(batch_size, -1, dim); k = feat_k.view(batch_size, -1, dim)
    npatch = q.size(1)
    l_neg = torch.bmm(q, k.transpose(2, 1))
    l_neg = l_neg.masked_fill(torch.eye(npatch, device=q.device, dtype=torch.bool)[None], -10.0).view(-1, npatch)
    out = torch.cat((l_pos, l_neg), 1) / T
    return F.cross_entropy(out, torch.zeros(n, dtype=torch.long, device=q.device), reduction="none")


def gradient_penalty(out_src, x_hat):
    """trainer/uganShp0Trainer.py:127-134."""
    (dydx,) = torch.autograd.grad(out_src, x_hat, torch.ones_like(out_src), retain_graph=True, create_graph=True)
    norm = dydx.view(dydx.size(0), -1).pow(2).sum(1).sqrt()
    return ((norm - 1) ** 2).mean()


def label2onehot(modals, dim):
    # trainer/uganShp0Trainer.py:109-113
    out = torch.zeros(modals.size(0), dim)
    out[torch.arange(modals.size(0)), modals.long()] = 1
    return out


def sigmoid_rampup(current, rampup_length):
    # trainer/baseTrainer.py:64-72
    if rampup_length == 0:
        return 1.0
    current = min(max(current, 0.0), rampup_length)
    phase = 1.0 - current / rampup_length
    return float(math.exp(-5.0 * phase * phase))


def poly_lr(base, it, max_iter, power=0.9):
    # trainer/uganConsisTrainer.py:198
    return base * (1.0 - it / max_iter) ** power


# --------------------------------------------------------------------------------------------------
# optimisers (functional; state lives in dicts keyed like the state_dict)
# --------------------------------------------------------------------------------------------------
def sgd_update(params, grads, state, lr, momentum=0.9, weight_decay=1e-3):
    """torch.optim.SGD(momentum, weight_decay) as used at trainer/uganShp0Trainer.py:72."""
    with torch.no_grad():
        for k, p in params.items():
            d = grads[k] + weight_decay * p
            buf = state.get(k)
            buf = d.clone() if buf is None else buf.mul_(momentum).add_(d)
            state[k] = buf
            p.sub_(lr * buf)


def adam_update(params, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-3):
    """torch.optim.Adam with L2 (not decoupled) weight decay, trainer/uganShp0Trainer.py:74."""
    with torch.no_grad():
        state["step"] = state.get("step", 0) + 1
        t = state["step"]
        for k, p in params.items():
            d = grads[k] + weight_decay * p
            m = state.setdefault("m." + k, torch.zeros_like(p)); v = state.setdefault("v." + k, torch.zeros_like(p))
            m.mul_(beta1).add_(d, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(d, d, value=1 - beta2)
            denom = (v.sqrt() / math.sqrt(1 - beta2 ** t)).add_(eps)
            p.addcdiv_(m, denom, value=-lr / (1 - beta1 ** t))


# --------------------------------------------------------------------------------------------------
# training steps
# --------------------------------------------------------------------------------------------------
def _leaf(sd):
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}


def unet_step(sd, opt_state, x, y, lr, dice_stats_hook=None):
    """trainer/unetTrainer.py:71-80: forward, Dice+CE, backward, SGD.  Updates sd in place; returns loss, grads."""
    leaf = _leaf(sd)
    loss = dice_ce_loss(unet_forward(leaf, x), y, dice_stats_hook=dice_stats_hook)
    grads = dict(zip(leaf, torch.autograd.grad(loss, list(leaf.values()))))
    sgd_update(sd, grads, opt_state, lr)
    return loss.detach(), grads


def _modal_vectors(modal_org, mj, n_modal, dev):
    modal_trg = torch.full_like(modal_org, mj)
    vec_org = label2onehot(modal_org.cpu(), n_modal).to(dev); vec_trg = label2onehot(modal_trg.cpu(), n_modal).to(dev)
    return modal_trg, vec_trg - vec_org, vec_org - vec_trg


def ugan_d_phase(G, D, d_state, x_real, modal_org, mj, alpha, sample_ids, lr, lambdas=(1.0, 10.0, 10.0, 10.0)):
    """D step of UGANConsisTrainer.train_epoch (trainer/uganConsisTrainer.py:129-149) with the random draws
    (target modality L114, alpha L138, patch ids ugan.py:321) injected.  Updates D / d_state in place."""
    lambda_cls, _, lambda_gp, _ = lambdas
    n_modal = D["conv_cls.weight"].shape[0]
    _, vec_ot, _ = _modal_vectors(modal_org, mj, n_modal, x_real.device)
    Dl = _leaf(D)
    out_src, out_cls = discriminator_forward(Dl, x_real)
    d_real = -out_src.mean()
    d_cls = F.cross_entropy(out_cls, modal_org)
    with torch.no_grad():
        _, x_fake0, _, _ = ugannce_forward(G, x_real, vec_ot, sample_ids=sample_ids)
    out_src, _ = discriminator_forward(Dl, x_fake0)
    d_fake = out_src.mean()
    x_hat = (alpha * x_real + (1 - alpha) * x_fake0).requires_grad_(True)
    out_src, _ = discriminator_forward(Dl, x_hat)
    d_gp = gradient_penalty(out_src, x_hat)
    d_loss = d_real + d_fake + lambda_cls * d_cls + lambda_gp * d_gp
    d_grads = dict(zip(Dl, torch.autograd.grad(d_loss, list(Dl.values()))))
    adam_update(D, d_grads, d_state, lr)
    losses = dict(D_real=d_real, D_fake=d_fake, D_cls=d_cls, D_gp=d_gp)
    return {k: float(v.detach()) for k, v in losses.items()}, d_grads


def ugan_g_phase(G, D, g_state, x_real, y_real, modal_org, mj, sample_ids, lr, it, lambda_semi, nce_batch=None,
                 semi_from_iter=1000, lambdas=(1.0, 10.0, 10.0, 10.0)):
    """G step (trainer/uganConsisTrainer.py:151-188) against the already-updated D.  Updates G / g_state in place."""
    lambda_cls, lambda_rec, _, lambda_seg = lambdas
    bs = y_real.shape[0]
    dev = x_real.device
    n_modal = D["conv_cls.weight"].shape[0]
    modal_trg, vec_ot, vec_to = _modal_vectors(modal_org, mj, n_modal, dev)
    nce_batch = nce_batch or bs
    Gl = _leaf(G)
    y_fake, x_fake, feat_x, _ = ugannce_forward(Gl, x_real, vec_ot, sample_ids=sample_ids)
    out_src, out_cls = discriminator_forward(D, x_fake)
    g_fake = -out_src.mean()
    g_cls = F.cross_entropy(out_cls, modal_trg)
    g_seg = dice_ce_loss(y_fake[:bs], y_real)
    y_rec, x_rec, feat_f, _ = ugannce_forward(Gl, x_fake, vec_to, sample_ids=sample_ids)
    g_rec = (x_real - x_rec).abs().mean()
    if it < semi_from_iter:
        g_semi = torch.zeros((), device=dev)
    else:
        g_semi = dice_ce_loss(y_rec, torch.argmax(y_fake, 1))  # consistency_loss, L45-53
    g_nce = patchnce_loss(feat_f[0], feat_x[0], nce_batch).mean()  # nce_loss, L55-64 (one layer)
    g_loss = g_fake + lambda_rec * g_rec + lambda_cls * g_cls + lambda_seg * g_seg + lambda_semi * g_semi + g_nce
    g_grads = dict(zip(Gl, torch.autograd.grad(g_loss, list(Gl.values()))))
    sgd_update(G, g_grads, g_state, lr)
    losses = dict(G_fake=g_fake, G_rec=g_rec, G_cls=g_cls, G_seg=g_seg, G_semi=g_semi, G_nce=g_nce)
    return {k: float(v.detach()) for k, v in losses.items()}, g_grads


def ugan_consis_step(G, D, g_state, d_state, x_real, y_real, modal_org, mj, alpha, sample_ids, lr, it, lambda_semi,
                     nce_batch=None, semi_from_iter=1000, lambdas=(1.0, 10.0, 10.0, 10.0)):
    """One iteration of UGANConsisTrainer.train_epoch (trainer/uganConsisTrainer.py:110-203), n_critic = 1.
    Returns the 10 losses, the D and G gradients; G, D and the optimiser states are updated in place."""
    d_losses, d_grads = ugan_d_phase(G, D, d_state, x_real, modal_org, mj, alpha, sample_ids, lr, lambdas)
    g_losses, g_grads = ugan_g_phase(G, D, g_state, x_real, y_real, modal_org, mj, sample_ids, lr, it, lambda_semi,
                                     nce_batch, semi_from_iter, lambdas)
    return {**d_losses, **g_losses}, d_grads, g_grads


def ema_alpha(it, base=0.99, warm=100):
    # trainer/meanTeacherTrainer.py:63-69 (alpha = min(1 - 1/(iter+1), 0.99); 0 before `warm`)
    return 0.0 if it < warm else min(1.0 - 1.0 / (it + 1), base)


def ema_update(ema, sd, alpha):
    with torch.no_grad():
        for k in ema:
            ema[k].mul_(alpha).add_(sd[k], alpha=1 - alpha)


def mean_teacher_step(sd, ema, opt_state, x, y, noise, lr, it, lambda_semi, warm=100):
    """trainer/meanTeacherTrainer.py:95-153: student on 2*bs slices, EMA teacher (no grad) on the noisy unlabelled half,
    Dice+CE on the labelled half + lambda * mean((softmax_s - softmax_t)^2) once it >= warm, SGD, EMA update."""
    bs = y.shape[0]
    leaf = _leaf(sd)
    out = unet_forward(leaf, x)
    with torch.no_grad():
        t_soft = torch.softmax(unet_forward(ema, x[bs:] + noise), 1)
    seg = dice_ce_loss(out[:bs], y)
    semi = torch.zeros((), device=x.device) if it < warm else ((torch.softmax(out, 1)[bs:] - t_soft) ** 2).mean()
    total = seg + lambda_semi * semi
    grads = dict(zip(leaf, torch.autograd.grad(total, list(leaf.values()))))
    sgd_update(sd, grads, opt_state, lr)
    ema_update(ema, sd, ema_alpha(it, warm=warm))
    return float(seg.detach()), float(semi.detach())


# coraNet (trainer/coraNetTrainer.py).  The shipped config.py pairs n_label = 4 with the 2-class weight vectors of the
# SAML configuration (config.py:82-88: `default_w = [1, 1]`, `w_con = [1, 5]`, `w_rad = [5, 1]`; the CHAOS vectors are
# the commented alternatives) -- nn.CrossEntropyLoss(weight=2 values) raises on 5-class logits, so the trainer runs as
# written only with matching vectors.  These are the CHAOS vectors of those comments.
CORA_W = dict(default=[1.0, 1.0, 1.0, 1.0, 1.0], con=[1.0, 5.0, 5.0, 5.0, 5.0], rad=[5.0, 1.0, 1.0, 1.0, 1.0])


def coranet_heads(out, n_label=4):
    """coraNetTrainer.py:279-286: the (1 + 3 n_label)-channel output -> three (1 + n_label)-class heads that share the
    background channel."""
    back = out[:, :1]
    return [torch.cat([back, out[:, 1 + h * n_label:1 + (h + 1) * n_label]], 1) for h in range(3)]


def soft_dice_loss(x, y, batch_dice):
    """misc/loss.py:39-63"""
    p = F.softmax(x, 1)
    oh = torch.zeros_like(p).scatter_(1, y.unsqueeze(1), 1.0)
    dims = (0, 2, 3) if batch_dice else (2, 3)
    tp = (p * oh).sum(dims); fp = (p * (1 - oh)).sum(dims); fn = ((1 - p) * oh).sum(dims)
    dc = (2 * tp + 1e-5) / (2 * tp + fp + fn + 1e-5 + 1e-8)
    return 1.0 - (dc[1:] if batch_dice else dc[:, 1:]).mean()


def coranet_supervised(heads, y, w=CORA_W, weight_ce=0.5, weight_dc=0.5):
    """coraNetTrainer.py:288-292 (and pre_epoch :485-488): (Dice+CE on head 0 + weighted CE on heads 1, 2) / 4"""
    wt = lambda k: torch.tensor(w[k], dtype=heads[0].dtype, device=heads[0].device)
    cedc = weight_dc * soft_dice_loss(heads[0], y, True) + weight_ce * F.cross_entropy(heads[0], y, weight=wt("default"))
    con = F.cross_entropy(heads[1], y, weight=wt("con"))
    rad = F.cross_entropy(heads[2], y, weight=wt("rad"))
    return (cedc + con + rad) / 4, (cedc, con, rad)


def coranet_unsupervised(heads, ema_heads, plab, mask, consistency_weight, w=CORA_W):
    """coraNetTrainer.py:299-338: certain areas (pseudo label where heads 1 and 2 agreed: per-sample Dice + masked CE,
    / 2) and uncertain areas (masked softmax-MSE against the EMA teacher's three heads, / 3)."""
    wt = torch.tensor(w["default"], dtype=heads[0].dtype, device=heads[0].device)
    dice2 = soft_dice_loss(heads[0], plab, False)
    ce2 = (F.cross_entropy(heads[0], plab, weight=wt, reduction="none") * mask).sum() / (mask.sum() + 1e-16)
    certain = (ce2 + dice2) / 2
    inv = (1 - mask).unsqueeze(1)
    consts = []
    for hs, ht in zip(heads, ema_heads):
        dist = (torch.softmax(hs, 1) - torch.softmax(ht, 1)) ** 2
        consts.append(consistency_weight * ((dist * inv).sum() / (inv.sum() + 1e-16)))
    return certain, sum(consts) / 3


def coranet_pre_step(sd, ema, opt_state, img1, msk, lr, it):
    """one iteration of pre_epoch (coraNetTrainer.py:436-499): only the labelled half reaches the loss"""
    leaf = _leaf(sd)
    heads = coranet_heads(unet_forward(leaf, img1))
    loss, parts = coranet_supervised(heads, msk)
    grads = dict(zip(leaf, torch.autograd.grad(loss, list(leaf.values()))))
    sgd_update(sd, grads, opt_state, lr)
    ema_update(ema, sd, ema_alpha(it))
    return [float(loss.detach())] + [float(v.detach()) for v in parts], grads


def coranet_train_step(sd, ema, opt_state, img1, msk, img2, plab2, mask, lr, it, consistency_weight, unsup_from=1000):
    """one iteration of train_epoch (coraNetTrainer.py:244-352): supervised + certain + 0.1 * uncertain, the latter two
    zeroed while it < 1000; SGD; EMA update."""
    leaf = _leaf(sd)
    sup, _ = coranet_supervised(coranet_heads(unet_forward(leaf, img1)), msk)
    heads2 = coranet_heads(unet_forward(leaf, img2))
    with torch.no_grad():
        ema_heads = coranet_heads(unet_forward(ema, img2))
    certain, uncertain = coranet_unsupervised(heads2, ema_heads, plab2, mask, consistency_weight)
    if it < unsup_from:
        certain, uncertain = torch.zeros((), device=img1.device), torch.zeros((), device=img1.device)
    loss = sup + certain + uncertain * 0.1
    grads = dict(zip(leaf, torch.autograd.grad(loss, list(leaf.values()), allow_unused=True)))
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in grads.items()}
    sgd_update(sd, grads, opt_state, lr)
    ema_update(ema, sd, ema_alpha(it))
    return [float(sup.detach()), float(certain.detach()), float(uncertain.detach())], grads


def coranet_pred_unlabel(sd, img):
    """pred_unlabel (coraNetTrainer.py:177-226) on a batch: pseudo label = argmax of head 0, certainty mask = heads 1
    and 2 agree"""
    with torch.no_grad():
        h = coranet_heads(unet_forward(sd, img))
        p0, p1, p2 = (t.argmax(1) for t in h)
    return p0, (p1 == p2).float()


def cross_pse_step(sd1, sd2, st1, st2, x, y, lr, lambda_semi):
    """trainer/crossPseTrainer.py:96-131 (cross pseudo supervision): two U-Nets on the same 2*bs slices; Dice+CE of
    each on the labelled half, plus lambda * Dice+CE of each net's unlabelled-half logits against the OTHER net's
    argmax; one backward, two SGD steps.  Updates both nets in place; returns the four losses and both grad dicts."""
    bs = y.shape[0]
    l1, l2 = _leaf(sd1), _leaf(sd2)
    out1, out2 = unet_forward(l1, x), unet_forward(l2, x)
    s1, s2 = dice_ce_loss(out1[:bs], y), dice_ce_loss(out2[:bs], y)
    pred1, pred2 = torch.argmax(out1[bs:], 1).detach(), torch.argmax(out2[bs:], 1).detach()
    c1, c2 = dice_ce_loss(out1[bs:], pred2), dice_ce_loss(out2[bs:], pred1)
    total = s1 + s2 + lambda_semi * c1 + lambda_semi * c2
    gs = torch.autograd.grad(total, list(l1.values()) + list(l2.values()))
    g1, g2 = dict(zip(l1, gs[:len(l1)])), dict(zip(l2, gs[len(l1):]))
    sgd_update(sd1, g1, st1, lr)
    sgd_update(sd2, g2, st2, lr)
    return dict(seg1=float(s1.detach()), seg2=float(s2.detach()), semi1=float(c1.detach()), semi2=float(c2.detach())), g1, g2


def ugan_shape_step(G, D, g_state, d_state, x_real, y_real, modal_org, mj, alpha, lr, lambda_shp=None,
                    lambdas=(1.0, 10.0, 10.0, 10.0)):
    """One iteration (n_critic = 1) of UGANTrainer.train_epoch (trainer/uganTrainer.py:159-196: labelled slices only,
    `UGAN` generator without the PatchNCE head, shape loss Dice+CE(y_rec, y_real) weighted lambda_shp) or, with
    lambda_shp=None, of UGANShp0Trainer.train_epoch (trainer/uganShp0Trainer.py:180-217, the same without the
    shape term).  Draws (target modality, alpha) injected.  Updates G, D and both optimiser states in place."""
    lambda_cls, lambda_rec, lambda_gp, lambda_seg = lambdas
    dev = x_real.device
    n_modal = D["conv_cls.weight"].shape[0]
    modal_trg, vec_ot, vec_to = _modal_vectors(modal_org, mj, n_modal, dev)
    Dl = _leaf(D)
    out_src, out_cls = discriminator_forward(Dl, x_real)
    d_real, d_cls = -out_src.mean(), F.cross_entropy(out_cls, modal_org)
    with torch.no_grad():
        _, x_fake0 = ugannce_forward(G, x_real, vec_ot, val_phase=True)
    d_fake = discriminator_forward(Dl, x_fake0)[0].mean()
    x_hat = (alpha * x_real + (1 - alpha) * x_fake0).requires_grad_(True)
    d_gp = gradient_penalty(discriminator_forward(Dl, x_hat)[0], x_hat)
    d_loss = d_real + d_fake + lambda_cls * d_cls + lambda_gp * d_gp
    d_grads = dict(zip(Dl, torch.autograd.grad(d_loss, list(Dl.values()))))
    adam_update(D, d_grads, d_state, lr)

    Gl = _leaf(G)
    y_fake, x_fake = ugannce_forward(Gl, x_real, vec_ot, val_phase=True)
    out_src, out_cls = discriminator_forward(D, x_fake)
    g_fake, g_cls = -out_src.mean(), F.cross_entropy(out_cls, modal_trg)
    g_seg = dice_ce_loss(y_fake, y_real)
    y_rec, x_rec = ugannce_forward(Gl, x_fake, vec_to, val_phase=True)
    g_rec = (x_real - x_rec).abs().mean()
    g_loss = g_fake + lambda_rec * g_rec + lambda_cls * g_cls + lambda_seg * g_seg
    losses = dict(D_real=d_real, D_fake=d_fake, D_cls=d_cls, D_gp=d_gp, G_fake=g_fake, G_rec=g_rec, G_cls=g_cls, G_seg=g_seg)
    if lambda_shp is not None:
        g_shp = dice_ce_loss(y_rec, y_real)
        g_loss = g_loss + lambda_shp * g_shp
        losses["G_shp"] = g_shp
    used = {k: v for k, v in Gl.items() if not k.startswith("netF.")}
    g_grads = dict(zip(used, torch.autograd.grad(g_loss, list(used.values()))))
    sgd_update({k: G[k] for k in used}, g_grads, g_state, lr)
    return {k: float(v.detach()) for k, v in losses.items()}, d_grads, g_grads


def mo_matrix(prd_npys, gt_npys, n_modal=4, n_label=4):
    """Modality-organ Dice matrix of misc/utils.py:180-203 (get_mo_matrix) in numpy.  prd_npys / gt_npys: dicts
    '<modality index>_<patient>' -> (slices, H, W) integer label volumes.  The per-volume score is
    medpy.metric.binary.dc (medpy is not installed here; its published definition: 2 |p & g| / (|p| + |g|) over the
    boolean masks, 0.0 when both are empty), averaged over the volumes of a modality; last row / column = means."""
    import numpy as np
    matrix = np.zeros((n_modal, n_label))
    n = np.zeros((n_modal, 1))
    for k in gt_npys.keys():
        m = int(k.split('_')[0])
        p, g = np.asarray(prd_npys[k]), np.asarray(gt_npys[k])
        for i in range(n_label):
            j = i + 1
            a, b = (p == j), (g == j)
            denom = float(a.sum() + b.sum())
            matrix[m][i] += 2.0 * float((a & b).sum()) / denom if denom > 0 else 0.0
        n[m] += 1
    n[n == 0] += 1e-8
    matrix /= n
    full = np.zeros((n_modal + 1, n_label + 1))
    full[:n_modal, :n_label] = matrix
    full[-1, :] = np.mean(full[0:n_modal], axis=0)
    full[:, -1] = np.mean(full[:, 0:n_label], axis=1)
    return full
