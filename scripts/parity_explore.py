"""Development: the numbers behind the tightened parity bounds of tests/test_modules_gpu.py (run on a B200).
  (i)  UGANnce forward with the real (undoctored) head weights: seg / tsl / feat relative L2 vs the fp32 oracle
  (ii) argmax agreement as a function of the logit-margin threshold (U-Net, 256x256)
  (iii) 200-step free-running U-Net loss trajectory at 256x256, batch 4, normal and deterministic mode"""
import json
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
from oracle import smsut_oracle as O  # noqa: E402
from smsut_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
out = {}


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def to_dev(sd):
    return {k: v.to(DEV) for k, v in sd.items()}


# (i)
from smsut_b200.network.ugan import UGANnce  # noqa: E402
for seed in (4, 7):
    sd = to_dev(O.make_weights(O.ugan_shapes(), seed))
    net = UGANnce(1, 5, 4, 16).to(DEV)
    net.load_state_dict(sd)
    x, _ = O.synthetic_batch(2, 256, 4, device=DEV)
    m = torch.tensor([[1., 0, -1, 0], [0, 1., -1, 0]], device=DEV)
    ids = [torch.randperm(256, generator=torch.Generator().manual_seed(0))[:64].to(DEV)]
    with torch.no_grad():
        seg, tsl, feats, _ = net(x, m, sample_ids=ids)
        rseg, rtsl, rfeats, _ = O.ugannce_forward(sd, x, m, sample_ids=ids)
        # the tanh head: pre-activation error vs output error
        out[f"ugannce_seed{seed}"] = dict(seg=rel(seg, rseg), tsl=rel(tsl, rtsl), feat=rel(feats[0], rfeats[0]),
                                         tsl_abs_max=(tsl.float() - rtsl).abs().max().item(),
                                         tsl_saturated_frac=(rtsl.abs() > 0.99).float().mean().item(),
                                         tsl_rms=rtsl.pow(2).mean().sqrt().item())
print(json.dumps(out, indent=1), flush=True)

# (ii)
from smsut_b200.network.unet import UNet  # noqa: E402
sd = to_dev(O.make_weights(O.unet_shapes(), 1))
net = UNet(1, 5, 16, 'instance', 'lrelu').to(DEV).eval()
net.load_state_dict(sd)
rows = []
with torch.no_grad():
    for n, seed in ((2, 3), (16, 66)):
        x, _ = O.synthetic_batch(n, 256, seed, device=DEV)
        o, r = net(x).float(), O.unet_forward(sd, x)
        err = (o - r).abs()
        top2 = r.topk(2, dim=1).values
        margin = top2[:, 0] - top2[:, 1]
        scale = r.abs().amax(dim=1)
        agree = o.argmax(1) == r.argmax(1)
        row = dict(n=n, logits_rel=rel(o, r), max_abs_err=err.max().item(), rms_err=err.pow(2).mean().sqrt().item(),
                   rms_logit=r.pow(2).mean().sqrt().item(), disagree_total=int((~agree).sum()),
                   max_margin_of_disagreeing=(margin[~agree].max().item() if (~agree).any() else 0.0),
                   max_margin_over_scale_of_disagreeing=((margin / scale)[~agree].max().item() if (~agree).any() else 0.0))
        for tol in (1e-2, 2e-2, 4e-2, 5e-2, 1e-1):
            mask = margin > tol * scale
            row[f"tol{tol}"] = dict(excluded=1 - mask.float().mean().item(), disagree=int((~agree & mask).sum()))
        rows.append(row)
out["argmax_margin"] = rows
print(json.dumps(rows, indent=1), flush=True)

# (iii)
from smsut_b200.trainer.unetTrainer import UnetTrainer  # noqa: E402
for det in (False, True):
    ops.set_deterministic(det)
    tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=256))
    sd = to_dev(O.make_weights(O.unet_shapes(), 21))
    tr.net.load_state_dict(sd)
    st, traj = {}, []
    for it in range(200):
        x, y = O.synthetic_batch(4, 256, 30 + it % 8, device=DEV)
        loss = tr.train_step(x, y).item()
        ref, _ = O.unet_step(sd, st, x, y, O.poly_lr(1e-2, max(it - 1, 0), 30000))
        traj.append((loss, ref.item()))
    dev = [abs(a - b) / abs(b) for a, b in traj]
    out[f"trajectory256_det{int(det)}"] = dict(mean=sum(dev) / len(dev), worst=max(dev), first40=max(dev[:40]),
                                              first=traj[0], last=traj[-1])
    print(json.dumps(out[f"trajectory256_det{int(det)}"]), flush=True)
    del tr
ops.set_deterministic(False)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/parity_explore.json", "w"), indent=1)
