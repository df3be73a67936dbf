"""Seeded synthetic Gaofen-like multispectral scenes (SURVEY.md section 8d "Synthetic inputs").

The reference ships no imagery (`data/sample.tif` is a missing blob), so every test / benchmark scene is
generated: a smooth low-frequency field + band-correlated mid-frequency texture + small white sensor noise,
quantised to the stated bit depth, CHW uint16.  `make_scene` is the numpy generator used by tests and golden
fixtures (bit-reproducible from the seed); `make_scene_torch` produces a scene of the same character directly
on a CUDA device for the large benchmark shapes (values are not identical to the numpy generator).
"""
import numpy as np

SEED = 19920517


def _upsample(grid, H, W):
    """Bilinear upsample of a coarse (h,w) grid to (H,W) with numpy only (align_corners=True)."""
    h, w = grid.shape
    ys = np.linspace(0, h - 1, H)
    xs = np.linspace(0, w - 1, W)
    y0 = np.floor(ys).astype(np.int64).clip(0, h - 2) if h > 1 else np.zeros(H, np.int64)
    x0 = np.floor(xs).astype(np.int64).clip(0, w - 2) if w > 1 else np.zeros(W, np.int64)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    y1 = np.minimum(y0 + 1, h - 1)
    x1 = np.minimum(x0 + 1, w - 1)
    g = grid
    top = g[y0][:, x0] * (1 - fx) + g[y0][:, x1] * fx
    bot = g[y1][:, x0] * (1 - fx) + g[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def make_scene(C, H, W, bits=12, seed=SEED, peak_frac=0.9):
    """Return a CHW uint16 scene with values in [0, 2**bits)."""
    rng = np.random.default_rng(seed)
    top = float(2 ** bits - 1) * peak_frac
    low = _upsample(rng.random((max(2, H // 96 + 2), max(2, W // 96 + 2))), H, W)
    mid = _upsample(rng.random((max(2, H // 12 + 2), max(2, W // 12 + 2))), H, W)
    fine = _upsample(rng.random((max(2, H // 3 + 2), max(2, W // 3 + 2))), H, W)
    out = np.empty((C, H, W), dtype=np.uint16)
    for c in range(C):
        gain = 0.55 + 0.45 * rng.random()
        own = _upsample(rng.random((max(2, H // 6 + 2), max(2, W // 6 + 2))), H, W)
        v = 0.08 + gain * (0.50 * low + 0.22 * mid + 0.10 * fine + 0.06 * own)
        v = v * top + rng.normal(0.0, top * 0.0025, size=(H, W))
        out[c] = np.clip(np.rint(v), 0, 2 ** bits - 1).astype(np.uint16)
    return out


def make_scene_torch(C, H, W, bits=12, seed=SEED, device="cuda", peak_frac=0.9):
    """Same recipe on a torch device (for 8192^2 / 16384^2 benchmark scenes); CHW uint16 tensor."""
    import torch
    import torch.nn.functional as F

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    top = float(2 ** bits - 1) * peak_frac

    def field(div):
        coarse = torch.rand((1, 1, max(2, H // div + 2), max(2, W // div + 2)), generator=g, device=device)
        return F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True)[0, 0]

    low, mid, fine = field(96), field(12), field(3)
    out = torch.empty((C, H, W), dtype=torch.uint16, device=device)
    for c in range(C):
        gain = 0.55 + 0.45 * float(torch.rand((), generator=g, device=device))
        v = 0.08 + gain * (0.50 * low + 0.22 * mid + 0.10 * fine + 0.06 * field(6))
        v = v * top + torch.randn((H, W), generator=g, device=device) * (top * 0.0025)
        q = v.round_().clamp_(0, 2 ** bits - 1).to(torch.int32)
        out[c] = ((q + 32768) % 65536 - 32768).to(torch.int16).view(torch.uint16)   # int16 bit pattern == uint16 value
    return out
