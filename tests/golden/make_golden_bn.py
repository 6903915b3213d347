"""Generates tests/golden/unet_bn.npz by running the REAL reference UNet with its DEFAULT norm / activation
(norm_type='batch', act_type='relu': network/unet.py:14, network/blocks.py:19-34) from /root/reference (CPU, fp32):
two training-mode forward/backward passes (so the running estimates move twice) and one eval-mode forward.
Run in the build container only; tests/test_oracle.py pins the oracle's batch-norm style against the output."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from misc.loss import DiceAndCrossEntropyLoss  # noqa: E402  (the reference's)
from network.unet import UNet  # noqa: E402

from oracle import smsut_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def npy(t):
    return t.detach().cpu().numpy().astype(np.float32)


net = UNet(1, 5, 16)            # defaults: batch norm + ReLU
net.load_state_dict(O.add_bn_buffers(O.make_weights(O.unet_shapes(), 11)))
crit = DiceAndCrossEntropyLoss(weight_ce=0.5, weight_dc=0.5, batch_dice=True)
out = {}
net.train()
for it, seed in enumerate((21, 22)):
    x, y = O.synthetic_batch(2, 48, seed)
    net.zero_grad()
    logits = net(x)
    loss = crit(logits, y)
    loss.backward()
    out[f"logits{it}"] = npy(logits)
    out[f"loss{it}"] = npy(loss)
    for k, p in net.named_parameters():
        out[f"gn{it}.{k}"] = np.float32(p.grad.norm().item())
out["fc_grad"] = npy(net.decoder.fc.weight.grad)
out["pre_grad"] = npy(net.encoder.pre_conv.weight.grad)
for k, v in net.state_dict().items():
    if "running_" in k:
        out["buf." + k] = npy(v)
net.eval()
x, _ = O.synthetic_batch(2, 48, 23)
out["logits_eval"] = npy(net(x))
np.savez_compressed(os.path.join(HERE, "unet_bn.npz"), **out)
print("unet_bn", sum(a.nbytes for a in out.values()) // 1024, "KiB raw")
