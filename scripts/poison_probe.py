"""Development: run U-Net trainer steps after filling the caching allocator's free blocks with a poison pattern (nan /
zero / 1e30), to expose reads of uninitialised device memory (a fresh box hands out zeroed memory, later processes get
whatever earlier kernels left)."""
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.load_package()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import smsut_oracle as O  # noqa: E402
from smsut_b200.trainer.unetTrainer import UnetTrainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "nan"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
DEV = "cuda"


def poison():
    val = {"nan": float("nan"), "zero": 0.0, "big": 1e30}[mode]
    blocks = [torch.full((64 * 2 ** 20,), val, dtype=torch.float32, device=DEV) for _ in range(24)]   # 6 GiB
    small = [torch.full((n,), val, dtype=torch.float32, device=DEV) for n in (256, 4096, 65536, 2 ** 20) for _ in range(64)]
    del blocks, small
    torch.cuda.synchronize()


poison()
tr = UnetTrainer('train', SimpleNamespace(fold=0, expr_name=None, input_size=128))
sd = {k: v.to(DEV) for k, v in O.make_weights(O.unet_shapes(), 21).items()}
tr.net.load_state_dict(sd)
st, devs = {}, []
for it in range(steps):
    if it % 10 == 0:
        poison()
    x, y = O.synthetic_batch(4, 128, 30 + it % 8, device=DEV)
    loss = tr.train_step(x, y).item()
    ref, _ = O.unet_step(sd, st, x, y, O.poly_lr(1e-2, max(it - 1, 0), 30000))
    devs.append(abs(loss - ref.item()) / abs(ref.item()))
    pairs = globals().setdefault("pairs", [])
    pairs.append((loss, ref.item()))
win = [abs(sum(a for a, _ in pairs[i:i + 8]) - sum(b for _, b in pairs[i:i + 8])) / sum(b for _, b in pairs[i:i + 8])
       for i in range(0, len(pairs) - 7, 8)]
print("8-step window means: mean dev %.4f worst %.4f" % (sum(win) / len(win), max(win)))
print(mode, "mean dev %.4f worst %.4f first5" % (sum(devs) / len(devs), max(devs)), [round(d, 5) for d in devs[:5]], flush=True)
