"""Per-layer parity protocol (north_star: "per-layer activations and gradients within 2e-2
relative (bf16)") -- test infrastructure shared by tests/test_host_logic.py (CPU, fp32 test double: validates the
protocol itself) and tests/test_parity_layers_gpu.py (B200, the kernels).

Two experiments, both against oracle/smsut_oracle.py:

(1) TEACHER-FORCED PER LAYER.  The oracle runs the whole network once on seeded inputs (fp32) and a realistic loss
    is back-propagated; every layer's input x_l and output cotangent g_l are recorded.  Each layer of the drop-in
    network is then run ALONE on the bf16-rounded x_l, back-propagated from the bf16-rounded g_l, and compared with
    the oracle's layer evaluated on the same rounded tensors: output, d/dx and every parameter gradient, relative
    L2.  Errors cannot compound across layers, so a wrong kernel shows up in its own layer.
(2) END TO END WITH FORCED SELECTIONS.  The drop-in network runs free; the LeakyReLU masks it used (sign of every
    fused norm+activation output) and its max-pool argmax positions are handed to the oracle (RecordingStyle), which
    then evaluates the same piecewise-linear function in fp32.  With the selections equal on both sides the
    gradients must agree to bf16 accuracy; the mask-flip fraction (oracle free vs drop-in) is reported beside it.
    This is the experiment that separates "bf16 storage flips LeakyReLU masks"  from a
    backward bug: a wrong dgrad / wgrad / norm-backward kernel fails (1) and (2) alike, a mask flip fails neither.
"""
import torch
import torch.nn.functional as F

from oracle import smsut_oracle as O

BF16 = torch.bfloat16


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _r16(t):
    """bf16-rounded copy as fp32 (the values both sides consume)"""
    return t.detach().to(BF16).float()


def _pad16(c):
    return (c + 15) // 16 * 16


def to_ours(x_nchw, act_dtype=BF16):
    """oracle NCHW fp32 -> drop-in NHWC leaf (channels zero-padded to a multiple of 16), requires_grad"""
    n, c, h, w = x_nchw.shape
    v = x_nchw.detach().permute(0, 2, 3, 1)
    if c % 16:
        v = F.pad(v, (0, _pad16(c) - c))
    return v.to(act_dtype).contiguous().requires_grad_(True)


def from_ours(y_nhwc, c):
    """drop-in NHWC -> NCHW fp32 cropped to the c real channels"""
    return y_nhwc.detach()[..., :c].permute(0, 3, 1, 2).float()


def cot_to_ours(g_nchw, like):
    """oracle cotangent (NCHW fp32) -> tensor shaped / typed like the drop-in output `like` (NHWC, padded)"""
    g = g_nchw.detach().permute(0, 2, 3, 1)
    if g.shape[-1] != like.shape[-1]:
        g = F.pad(g, (0, like.shape[-1] - g.shape[-1]))
    return g.to(like.dtype).contiguous()


class Layer:
    """one teacher-forced comparison: `ours(*leaves) -> NHWC tensor`, `oracle(sd, style, *xs) -> NCHW tensor`"""

    def __init__(self, name, ours, oracle, inputs, cot, params, image_input=False, out_c=None):
        self.name, self.ours, self.oracle = name, ours, oracle
        self.inputs, self.cot, self.params = inputs, cot, params      # params: {oracle key: drop-in parameter}
        self.image_input = image_input                                # inputs are fp32 NCHW images, not activations
        self.out_c = out_c


def run_layer(Fn, layer, sd, act_dtype=BF16):
    """returns dict(fwd, dx=[...], params={key: rel}, flips) for one Layer"""
    # ---- the drop-in layer alone, on the rounded inputs
    if layer.image_input:
        xs_r = [x.detach().clone() for x in layer.inputs]                      # images are fp32 on both sides
        leaves = [x.clone().requires_grad_(True) for x in xs_r]
    else:
        xs_r = [_r16(x) if act_dtype == BF16 else x.detach().clone() for x in layer.inputs]
        leaves = [to_ours(x, act_dtype) for x in xs_r]
    for p in layer.params.values():
        p.grad = None
    Fn.ACT_TAPS[0] = []
    try:
        y = layer.ours(*leaves)
    finally:
        taps, Fn.ACT_TAPS[0] = Fn.ACT_TAPS[0], None
    cot = layer.cot
    out_c = layer.out_c or cot.shape[1]
    cot_r = _r16(cot) if (y.dtype == BF16) else cot.detach().clone()
    if y.dim() == 4:
        y.backward(cot_to_ours(cot_r, y))
        y_cmp = from_ours(y, out_c)
    else:
        y.backward(cot_r.to(y.dtype))
        y_cmp = y.detach().float()
    masks = [from_ours(t, t.shape[-1]) > 0 for _, t in taps]
    # ---- the oracle's layer on the same tensors, with the drop-in layer's masks
    leaf_sd = {k: sd[k].detach().clone().requires_grad_(True) for k in layer.params}
    full = dict(sd)
    full.update(leaf_sd)
    xr = [x.clone().requires_grad_(True) for x in xs_r]

    def crop(m, ref_c):
        return m[:, :ref_c]

    class _Forced(O.RecordingStyle):
        def act(self, x, key=None):
            if self.forced_masks:
                self.forced_masks[0] = crop(self.forced_masks[0], x.shape[1])
            return super().act(x, key)

    st = _Forced(forced_masks=list(masks))
    yr = layer.oracle(full, st, *xr)
    grads = torch.autograd.grad(yr, xr + list(leaf_sd.values()), cot_r, allow_unused=True)
    gx, gp = grads[:len(xr)], grads[len(xr):]
    # mask flips of this layer alone (same input on both sides): oracle free vs drop-in
    st_free = O.RecordingStyle()
    with torch.no_grad():
        layer.oracle(full, st_free, *[x.detach() for x in xs_r])
    free = list(st_free.masks.values())
    nflip = sum(int((crop(a, b.shape[1]) != b).sum()) for a, b in zip(masks, free))
    ntot = sum(b.numel() for b in free)
    yr_d = yr.detach().float()
    res = dict(fwd=rel(y_cmp, yr), dx=[], params={}, flips=(nflip / ntot if ntot else 0.0),
               # tail of the forward error: largest element error in units of the reference's RMS (a halo / tile-edge
               # bug would put a few elements far outside what bf16 rounding of large activations explains)
               fwd_max_over_rms=((y_cmp.reshape(yr_d.shape) - yr_d).abs().max() / (yr_d.pow(2).mean().sqrt() + 1e-30)).item(),
               ref_max_over_rms=(yr_d.abs().max() / (yr_d.pow(2).mean().sqrt() + 1e-30)).item())
    for leaf, x, g in zip(leaves, xs_r, gx):
        if g is None:
            continue
        if layer.image_input:
            res["dx"].append(rel(leaf.grad, g))
        else:
            res["dx"].append(rel(from_ours(leaf.grad, x.shape[1]), g))
    for (k, p), g in zip(layer.params.items(), gp):
        if g is not None and p.grad is not None:
            res["params"][k] = rel(p.grad, g)
    return res


# --------------------------------------------------------------------------------------------------
# layer lists of the three networks, built from one recorded oracle run
# --------------------------------------------------------------------------------------------------
def _block_params(net_params, prefix, sd):
    return {k: net_params[k] for k in sd if k.startswith(prefix) and k in net_params}


def _basic_block_layer(name, blk, prefix, st, net_params, sd, split=None, key=None):
    key = key or prefix
    x_in = st.taps[key + "in"]
    ins = [x_in] if split is None else [x_in[:, :split], x_in[:, split:]]

    def ours(*xs):
        return blk.forward_nhwc(list(xs))

    def oracle(full, style, *xs):
        return O.basic_block(torch.cat(xs, 1) if len(xs) > 1 else xs[0], full, prefix, style)
    return Layer(name, ours, oracle, ins, st.taps[key + "out"].grad, _block_params(net_params, prefix, sd))


def unet_layers(Fn, net, sd, st):
    """net: drop-in UNet(instance, lrelu); st: RecordingStyle of the oracle run whose loss was back-propagated"""
    from smsut_b200.network import blocks as B
    P = dict(net.named_parameters())
    enc, dec = net.encoder, net.decoder
    L = []
    L.append(Layer("encoder.pre", lambda x: B._stem(enc.pre_conv, enc.pre_bn, enc.pre_relu, Fn.ImageInputFn.apply(x)),
                   lambda full, style, x: style.act(style.norm(F.conv2d(x, full["encoder.pre_conv.weight"], padding=2),
                                                               full, "encoder.pre_bn."), "a"),
                   [st.taps["encoder.pre.in"]], st.taps["encoder.pre.out"].grad,
                   {k: P[k] for k in sd if k.startswith("encoder.pre_")}, image_input=True))
    for i in range(1, 6):
        L.append(_basic_block_layer(f"encoder.layer{i}", getattr(enc, f"layer{i}"), f"encoder.layer{i}.", st, P, sd))
    for i in (4, 3, 2, 1):
        up = getattr(dec, f"up{i}")
        L.append(Layer(f"decoder.up{i}", lambda x, up=up: up.up.forward_nhwc(x),
                       lambda full, style, x, i=i: F.conv_transpose2d(x, full[f"decoder.up{i}.up.weight"], stride=2),
                       [st.taps[f"decoder.up{i}.in"]], st.taps[f"decoder.up{i}.out"].grad,
                       {f"decoder.up{i}.up.weight": P[f"decoder.up{i}.up.weight"]}))
        c_up = st.taps[f"decoder.up{i}.out"].shape[1]
        L.append(_basic_block_layer(f"decoder.layer{i}", getattr(dec, f"layer{i}"), f"decoder.layer{i}.", st, P, sd,
                                    split=c_up))
    L.append(Layer("decoder.fc", lambda x: dec.fc.forward_nhwc([x]),
                   lambda full, style, x: F.conv2d(x, full["decoder.fc.weight"]),
                   [st.taps["decoder.fc.in"]], st.taps["decoder.fc.out"].grad, {"decoder.fc.weight": P["decoder.fc.weight"]}))
    return L


def ugan_layers(Fn, net, sd, st, m, ids, feat_cot):
    """net: drop-in UGANnce; m: modality difference vectors (B, n_modal); ids: [patch ids]; feat_cot: cotangent of the
    sampled features (R, 256)"""
    from smsut_b200.network import blocks as B
    from smsut_b200.network import ugan as U
    P = dict(net.named_parameters())
    L = []
    planes = m.view(m.size(0), m.size(1), 1, 1)

    def stem(p, encoder, tsl):
        def ours(x):
            xin = U._TslInputFn.apply(x, m) if tsl else Fn.ImageInputFn.apply(x)
            return B._stem(encoder.pre[0], encoder.pre[1], encoder.pre[2], xin)

        def oracle(full, style, x):
            if tsl:
                x = torch.cat([x, planes.repeat(1, 1, x.size(2), x.size(3))], 1)
            return style.act(O.inorm(F.conv2d(x, full[p + "pre.0.weight"], padding=2), full, p + "pre.1."), "a")
        x_in = st.taps[p + "pre.in"][:, :1]
        return Layer(p + "pre", ours, oracle, [x_in], st.taps[p + "pre.out"].grad,
                     {k: P[k] for k in sd if k.startswith(p + "pre.")}, image_input=True)

    for p, encoder, tsl in (("tsl_encoder.", net.tsl_encoder, True), ("seg_encoder.", net.seg_encoder, False)):
        L.append(stem(p, encoder, tsl))
        for i in range(1, 5):
            L.append(_basic_block_layer(f"{p}enc{i}", getattr(encoder, f"enc{i}"), f"{p}enc{i}.", st, P, sd))
    for br in ("tsl", "seg"):
        L.append(_basic_block_layer(f"{br}.enc5", net.enc5, "enc5.", st, P, sd, key=f"{br}.enc5."))
    for p, decoder, transposed in (("tsl_decoder.", net.tsl_decoder, False), ("seg_decoder.", net.seg_decoder, True)):
        for i in (4, 3, 2, 1):
            up = getattr(decoder, f"up{i}")
            if transposed:
                L.append(Layer(f"{p}up{i}", lambda x, up=up: up.up.forward_nhwc(x),
                               lambda full, style, x, k=f"{p}up{i}.up.weight": F.conv_transpose2d(x, full[k], stride=2),
                               [st.taps[f"{p}up{i}.in"]], st.taps[f"{p}up{i}.out"].grad,
                               {f"{p}up{i}.up.weight": P[f"{p}up{i}.up.weight"]}))
            else:
                L.append(Layer(f"{p}up{i}", lambda x, up=up: up.up[1].forward_nhwc([Fn.BilinearFn.apply(x)]),
                               lambda full, style, x, k=f"{p}up{i}.up.1.weight": F.conv2d(
                                   F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), full[k]),
                               [st.taps[f"{p}up{i}.in"]], st.taps[f"{p}up{i}.out"].grad,
                               {f"{p}up{i}.up.1.weight": P[f"{p}up{i}.up.1.weight"]}))
            c_up = st.taps[f"{p}up{i}.out"].shape[1]
            L.append(_basic_block_layer(f"{p}dec{i}", getattr(decoder, f"dec{i}"), f"{p}dec{i}.", st, P, sd, split=c_up))
        tanh = not transposed
        L.append(Layer(p + "fc", lambda x, d=decoder: d.fc.forward_nhwc([x]),
                       lambda full, style, x, p=p, tanh=tanh: (torch.tanh if tanh else (lambda v: v))(
                           F.conv2d(x, full[p + "fc.weight"], full[p + "fc.bias"])),
                       [st.taps[p + "fc.in"]], st.taps[p + "fc.out"].grad,
                       {p + "fc.weight": P[p + "fc.weight"], p + "fc.bias": P[p + "fc.bias"]}))
    # netF: gather 64 shared positions of the bottleneck, Linear-ReLU-Linear, L2 normalise (network/ugan.py:316-334)
    L.append(Layer("netF", lambda x: net.netF([Fn.to_nchw(x)], patch_ids=ids)[0][0],
                   lambda full, style, x: O.patch_sample(full, x, ids[0]),
                   [st.taps["tsl.enc5.out"]], feat_cot, {k: P[k] for k in sd if k.startswith("netF.")}))
    return L


def disc_layers(Fn, D, sd, st, src_cot, cls_cot):
    from smsut_b200.network import blocks as B
    P = dict(D.named_parameters())
    L = []
    L.append(Layer("main.0", lambda x: D.main[0].forward_nhwc([B._image_nhwc(x)]),
                   lambda full, style, x: style.act(F.conv2d(x, full["main.0.weight"], full["main.0.bias"], stride=2,
                                                             padding=1), "a"),
                   [st.taps["main.0.in"]], st.taps["main.0.out"].grad,
                   {"main.0.weight": P["main.0.weight"], "main.0.bias": P["main.0.bias"]}, image_input=True))
    i = 2
    while f"main.{i}.conv1.weight" in sd:
        p = f"main.{i}."
        L.append(Layer(f"main.{i}", lambda x, blk=D.main[i]: blk.forward_nhwc(x),
                       lambda full, style, x, p=p: O.bottle_block(x, full, p, style),
                       [st.taps[p + "in"]], st.taps[p + "out"].grad, _block_params(P, p, sd)))
        i += 1
    L.append(Layer("conv_src", lambda x: D.conv_src.forward_nhwc([x]),
                   lambda full, style, x: F.conv2d(x, full["conv_src.weight"], padding=1),
                   [st.taps["heads.in"]], src_cot, {"conv_src.weight": P["conv_src.weight"]}))
    L.append(Layer("conv_cls", lambda x: D.conv_cls.forward_nhwc([x]),
                   lambda full, style, x: F.conv2d(x, full["conv_cls.weight"]),
                   [st.taps["heads.in"]], cls_cot.view(cls_cot.size(0), cls_cot.size(1), 1, 1),
                   {"conv_cls.weight": P["conv_cls.weight"]}))
    return L


def summarize(results):
    """worst value of each kind over all layers + where"""
    worst = dict(fwd=(0.0, None), dx=(0.0, None), params=(0.0, None), flips=(0.0, None))
    for name, r in results.items():
        if r["fwd"] > worst["fwd"][0]:
            worst["fwd"] = (r["fwd"], name)
        for v in r["dx"]:
            if v > worst["dx"][0]:
                worst["dx"] = (v, name)
        for k, v in r["params"].items():
            if v > worst["params"][0]:
                worst["params"] = (v, k + " @ " + name)
        if r["flips"] > worst["flips"][0]:
            worst["flips"] = (r["flips"], name)
    return worst


# --------------------------------------------------------------------------------------------------
# experiment (2): end to end with forced selections
# --------------------------------------------------------------------------------------------------
def selection_keys(net, kind):
    """id(norm module / fused-activation conv) -> function(occurrence) -> oracle activation key"""
    table = {}
    for name, mod in net.named_modules():
        if kind == "unet":
            if name == "encoder.pre_bn":
                table[id(mod)] = lambda occ: "encoder.pre.act"
        if kind == "ugan" and name.endswith(".pre.1"):
            table[id(mod)] = lambda occ, p=name[:-len("pre.1")]: p + "pre.act"
        if kind == "disc" and name == "main.0":
            table[id(mod)] = lambda occ: "main.0.act"
        if name.endswith(".bn1") or name.endswith(".bn2"):
            blk, which = name[:-4], ("act1" if name.endswith("bn1") else "act2")
            if kind == "ugan" and blk == "enc5":
                # the drop-in generator runs its segmentation half first (network/ugan.py::_branches)
                table[id(mod)] = lambda occ, which=which: ("seg" if occ == 0 else "tsl") + ".enc5." + which
            else:
                table[id(mod)] = lambda occ, blk=blk, which=which: blk + "." + which
    return table


def collect_selections(taps, keys, pool_blocks):
    """taps: Fn.ACT_TAPS list of one free drop-in forward -> (forced_masks, forced_pool) dicts for RecordingStyle.
    pool_blocks: {oracle act2 key of the block feeding a max-pool: oracle pool key}."""
    seen, masks, pools = {}, {}, {}
    for mod, out in taps:
        f = keys.get(id(mod))
        if f is None:
            continue
        occ = seen.get(id(mod), 0)
        seen[id(mod)] = occ + 1
        key = f(occ)
        nchw = out.detach().permute(0, 3, 1, 2).float()
        masks[key] = nchw > 0
        if key in pool_blocks:
            pools[pool_blocks[key]] = F.max_pool2d(nchw, 2, 2, return_indices=True)[1]
    return masks, pools


class ForcedStyle(O.RecordingStyle):
    """RecordingStyle that crops the drop-in network's channel-padded masks / indices to the oracle's channel count"""

    def act(self, x, key=None):
        m = self.forced_masks.get(key) if isinstance(self.forced_masks, dict) else None
        if m is not None and m.shape[1] != x.shape[1]:
            self.forced_masks[key] = m[:, :x.shape[1]]
        return super().act(x, key)

    def pool(self, x, key=None):
        i = self.forced_pool.get(key) if isinstance(self.forced_pool, dict) else None
        if i is not None and i.shape[1] != x.shape[1]:
            self.forced_pool[key] = i[:, :x.shape[1]]
        return super().pool(x, key)
