"""Synthetic stand-in for the reference's PNG loaders (data_loader/inTurnLoader.py:15-97): CHAOS / Synapse data is
not available offline, so batches are abdominal-like slices generated on the fly.  Each
batch holds ONE modality (the InTurn sampler's contract, inTurnLoader.py:36-57) and yields the reference's tuple
(image (B,1,H,W) fp32 in [-1,1], label (B,H,W) int64, modality (B,) int64, names)."""
import torch


def make_slices(n, size, seed, n_label=4, modality=0):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
    img = torch.zeros(n, 1, size, size)
    lab = torch.zeros(n, size, size, dtype=torch.int64)
    organs = [(-0.35, -0.15, 0.30, 0.22), (-0.25, 0.35, 0.10, 0.13), (0.25, 0.35, 0.10, 0.13), (0.42, -0.10, 0.14, 0.18)]
    gain = (1.0, 0.8, 1.2, 0.6)[modality % 4]     # per-modality intensity table
    for i in range(n):
        j = torch.rand(12, generator=g) * 0.1 - 0.05
        body = ((xx / (0.85 + j[0])) ** 2 + (yy / (0.65 + j[1])) ** 2) < 1
        v = torch.full((size, size), 0.05)
        v[body] = 0.35 + j[2].item()
        for c, (cx, cy, rx, ry) in enumerate(organs[:n_label]):
            m = (((xx - cx - j[3 + c]) / rx) ** 2 + ((yy - cy - j[7 + c]) / ry) ** 2) < 1
            m &= body
            v[m] = (0.5 + 0.1 * c + j[11].item()) * gain
            lab[i][m] = c + 1
        v = v + 0.05 * torch.randn(size, size, generator=g)
        u8 = (v.clamp(0, 1) * 255).round()
        img[i, 0] = (u8 / 255 - 0.5) / 0.5
    return img, lab


class SyntheticLoader:
    """Endless iterator of single-modality batches from a pre-generated pinned pool (so the host side costs
    nothing in the timed region, like a warm DataLoader with pin_memory=True)."""

    def __init__(self, batch_size, size=256, n_modal=4, n_label=4, pool_batches=8, seed=2020, pin=True):
        self.batch_size, self.n_modal = batch_size, n_modal
        self.pool = []
        for b in range(pool_batches):
            mod = b % n_modal
            img, lab = make_slices(batch_size, size, seed + b, n_label, mod)
            mdl = torch.full((batch_size,), mod, dtype=torch.int64)
            if pin and torch.cuda.is_available():
                img, lab, mdl = img.pin_memory(), lab.pin_memory(), mdl.pin_memory()
            self.pool.append((img, lab, mdl, [f"{mod}_{b}_{z}" for z in range(batch_size)]))
        self.dataset = self

    def __len__(self):
        return len(self.pool) * self.batch_size

    def __iter__(self):
        for item in self.pool:
            yield item


def get_loader(base_root=None, phase='train', fold=0, batch_size=8, data_aug=None, size=256, seed=None, rank=0, **kw):
    """rank: a data-parallel replica's index -- replicas draw different training slices (and the same test slices)"""
    base = {'train': 2020, 'val': 4040, 'test': 6060}.get(phase, 8080)
    if phase in ('train', 'val'):
        base += 100003 * rank
    return SyntheticLoader(batch_size, size=size, seed=base if seed is None else seed, **kw)
