"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's LBDRN hot path.

This file is the parity oracle: a plain numpy / torch-CPU restatement of what lidq92/LBDRN-MSIC computes on
the path our CUDA library replaces.  It is imported only by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`; no product module imports it and the product path has
no CPU fallback.

PINNING: the reference has no tests or golden vectors of its own.  This restatement is
pinned instead against the UNMODIFIED reference executed in the build container (`oracle/run_reference.py`),
through the fixtures in `tests/golden/` minted by `oracle/make_golden.py` and checked by
`tests/test_oracle_golden.py` (re-mint with `python oracle/make_golden.py` where `/root/reference` is present).  One dependency stays unpinned:
`fpzip` (not installed; value map restated in `oracle/shims/fpzip.py`).

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
import math
from dataclasses import dataclass

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------------------
# feature flags (reference constants.py:3-14)
# ----------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Flags:
    use_coordinates: bool = False
    embedding: bool = False
    sigma: float = 1.4
    n_freq: int = 12
    use_colors: bool = True
    relative: bool = True

    def num_coords(self):
        # LBDRNdataset.py:105
        return (2 * self.n_freq * int(self.embedding) + 1) * 2 * int(self.use_coordinates)

    def num_colors(self, C, D):
        # LBDRNdataset.py:104
        return C * (2 * D + 1) ** 2 * int(self.use_colors)

    def dim_in(self, C, D):
        return self.num_coords() + self.num_colors(C, D)


DEFAULT_FLAGS = Flags()


# ----------------------------------------------------------------------------------------------------------
# a1: MSB / LSB split (LBDRNdataset.py:93-101)
# ----------------------------------------------------------------------------------------------------------
def split_msb_lsb(img, K):
    """img: CHW (or HW) unsigned integer array.  Returns (msb CHW u8|u16, lsb CHW float32 in [0,1])."""
    img = np.asarray(img)
    msb = img >> K
    lsb = (img - (msb << K)).astype(np.float32) / (2 ** K - 1)
    msb = msb.reshape((-1,) + msb.shape[-2:])
    lsb = lsb.reshape((-1,) + lsb.shape[-2:])
    msb = msb.astype(np.uint16) if msb.max() > 255 else msb.astype(np.uint8)
    return msb, lsb


# ----------------------------------------------------------------------------------------------------------
# a2/a3: per-pixel features (LBDRNdataset.py:104-130, duplicated at decode.py:77-102)
# ----------------------------------------------------------------------------------------------------------
def coordinate_block(H, W, flags, row0=0, row1=None):
    """(rows, W, num_coords) float32 coordinate / positional-encoding block (LBDRNdataset.py:108-118) for image rows
    [row0,row1) of an H x W image (every entry depends on its own (y, x) only, so a row range is a slice of the block)."""
    row1 = H if row1 is None else row1
    yy, xx = np.meshgrid(np.arange(row0, row1), np.arange(W), indexing="ij")
    c = np.stack([2 * yy / (H - 1) - 1, 2 * xx / (W - 1) - 1], axis=-1).astype(np.float32)
    if flags.embedding:
        freq = flags.sigma ** np.arange(flags.n_freq) * np.pi          # float64
        arg = freq * c[..., None]                                      # float32 coord promoted to float64
        c = np.concatenate([c[..., None], np.sin(arg), np.cos(arg)], axis=-1)
    return c.reshape(row1 - row0, W, -1).astype(np.float32)


def features(msb, D, flags=DEFAULT_FLAGS, row0=0, row1=None):
    """[n_rows*W, dim_in] float32 feature rows for image rows [row0,row1) of the CHW MSB image.

    Reflect padding (no edge repeat) and the global normaliser `msb.max()` always refer to the WHOLE image,
    so a row range gives exactly the corresponding rows of the full matrix (used to bound memory)."""
    msb = np.asarray(msb)
    C, H, W = msb.shape
    row1 = H if row1 is None else row1
    n = 2 * D + 1
    nco, ncl = flags.num_coords(), flags.num_colors(C, D)
    out = np.zeros((row1 - row0, W, nco + ncl), dtype=np.float32)
    if flags.use_coordinates:
        out[:, :, :nco] = coordinate_block(H, W, flags, row0, row1)
    if flags.use_colors:
        s = msb.astype(np.float32) / msb.max()                                   # LBDRNdataset.py:120
        lo, hi = row0 - D, row1 + D                                              # padded-row window
        ridx = np.arange(lo, hi)
        ridx = np.where(ridx < 0, -ridx, ridx)
        ridx = np.where(ridx > H - 1, 2 * (H - 1) - ridx, ridx)                   # numpy 'reflect'
        cidx = np.arange(-D, W + D)
        cidx = np.where(cidx < 0, -cidx, cidx)
        cidx = np.where(cidx > W - 1, 2 * (W - 1) - cidx, cidx)
        pad = s[:, ridx][:, :, cidx]                                             # C,(rows+2D),(W+2D)
        win = np.lib.stride_tricks.sliding_window_view(pad, (n, n), axis=(1, 2))  # C,rows,W,n,n
        win = win.transpose(1, 2, 0, 3, 4)                                       # rows,W,C,n,n
        if flags.relative and D > 0:                                             # LBDRNdataset.py:126-128
            ctr = pad[:, D:D + (row1 - row0), D:D + W].transpose(1, 2, 0)
            win = win - ctr[..., None, None]
        out[:, :, nco:] = win.reshape(row1 - row0, W, ncl)
    return out.reshape(-1, nco + ncl)


def labels(lsb):
    """[N, C] float32 targets (LBDRNdataset.py:131)."""
    return np.ascontiguousarray(lsb.transpose(1, 2, 0).reshape(-1, lsb.shape[0]))


# ----------------------------------------------------------------------------------------------------------
# a6: the network (LBDRNmodel.py:7-82).  Parameters are a flat list [W1,b1,...,Wo,bo] in state_dict order.
# ----------------------------------------------------------------------------------------------------------
def layer_shapes(dim_in, bc, C, nl):
    shp, d = [], dim_in
    for _ in range(nl):
        shp += [(bc, d), (bc,)]
        d = bc
    return shp + [(C, bc), (C,)]


def n_params(dim_in, bc, C, nl):
    return sum(int(np.prod(s)) for s in layer_shapes(dim_in, bc, C, nl))


def init_params(dim_in, bc, C, nl, w0=30.0, c=6.0):
    """SIREN init consuming torch's default generator exactly like LBDRNModel.__init__:
    per layer an `nn.Linear` constructor (its own default init draws), then uniform_(weight), uniform_(bias)
    with bound 1/dim_in for the first layer, sqrt(c/dim_in)/w0 otherwise (LBDRNmodel.py:32-36,62-77)."""
    params, d = [], dim_in
    for i in range(nl + 1):
        out = bc if i < nl else C
        lin = torch.nn.Linear(d, out)
        bound = (1 / d) if i == 0 else (math.sqrt(c / d) / w0)
        torch.nn.init.uniform_(lin.weight, -bound, bound)
        torch.nn.init.uniform_(lin.bias, -bound, bound)
        params += [lin.weight.detach().clone(), lin.bias.detach().clone()]
        d = bc
    return params


def forward(params, x, w0=30.0, relu=False):
    """y = sigmoid(Wo.sin(w0(...sin(w0(W1 x + b1))...)) + bo)  (LBDRNmodel.py:13,40-41,79-82).
    relu=True: the commented alternative of encode.py:75 / decode.py:108 (`activation=torch.nn.ReLU()`): every hidden
    SirenLayer applies ReLU to its Linear output instead of Sine (LBDRNmodel.py:37,41); init and head are unchanged."""
    h = x
    for i in range(0, len(params) - 2, 2):
        z = torch.nn.functional.linear(h, params[i], params[i + 1])
        h = torch.relu(z) if relu else torch.sin(w0 * z)
    return torch.sigmoid(torch.nn.functional.linear(h, params[-2], params[-1]))


def flatten_params(params):
    """state_dict-order C-order concatenation (encode.py:123-128)."""
    return np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in params]).astype(np.float32)


def unflatten_params(flat, dim_in, bc, C, nl):
    """Slice a flat vector back into tensors in state_dict order (decode.py:114-120)."""
    out, k = [], 0
    for s in layer_shapes(dim_in, bc, C, nl):
        n = int(np.prod(s))
        out.append(torch.from_numpy(np.array(flat[k:k + n], dtype=np.float32).reshape(s)))
        k += n
    return out


def fpzip_value_map(flat, precision):
    """Weights as the decoder will see them after fpzip(precision) (see oracle/shims/fpzip.py; unpinned)."""
    a = np.ascontiguousarray(flat, dtype=np.float32)
    if precision in (0, 32):
        return a.copy()
    mask = np.uint32((0xFFFFFFFF << (32 - precision)) & 0xFFFFFFFF)
    return (a.view(np.uint32) & mask).view(np.float32)


# ----------------------------------------------------------------------------------------------------------
# a14/a15: decode (decode.py:122-134)
# ----------------------------------------------------------------------------------------------------------
def predict(base, params, D, flags=DEFAULT_FLAGS, rows_per_chunk=256, relu=False):
    """Network output y[N, C] float32 for every pixel of the CHW base image (chunked over rows)."""
    C, H, W = base.shape
    outs = []
    with torch.no_grad():
        for r in range(0, H, rows_per_chunk):
            x = torch.from_numpy(features(base, D, flags, r, min(H, r + rows_per_chunk)))
            outs.append(forward(params, x, relu=relu))
    return torch.cat(outs, 0)


def decode_image(base, params, K, D, flags=DEFAULT_FLAGS, rows_per_chunk=256, relu=False):
    """uint16 CHW reconstruction: round-half-even(y*(2^K-1)) added to base<<K (decode.py:131-134)."""
    base = np.asarray(base).astype(np.uint16)
    C, H, W = base.shape
    y = predict(base, params, D, flags, rows_per_chunk, relu)
    residual = torch.round(y * (2 ** K - 1)).numpy().reshape(H, W, C).transpose(2, 0, 1)
    return np.round((base << K).astype(np.float32) + residual).astype(np.uint16)


def eval_mse(msb, lsb, params, D, flags=DEFAULT_FLAGS, rows_per_chunk=256, relu=False):
    """Full-scene MSE used for best-epoch selection (encode.py:105-108, LBDRNperformance.py:18-21),
    accumulated in float64 over row chunks (the reference takes one float32 mean over all N*C)."""
    C, H, W = msb.shape
    t = torch.from_numpy(labels(lsb))
    y = predict(msb, params, D, flags, rows_per_chunk, relu)
    return float(((y.double() - t.double()) ** 2).mean())


# ----------------------------------------------------------------------------------------------------------
# a9/a10: schedule and batch sampling
# ----------------------------------------------------------------------------------------------------------
def lr_at_epoch(lr, epoch, epochs):
    """StepLR(step_size=max(1,int(E/3)), gamma=0.1) stepped at each epoch end (encode.py:85,98); epoch is 1-based."""
    v = lr
    for _ in range((epoch - 1) // max(1, int(epochs / 3))):
        v = v * 0.1                                                    # chained like StepLR's closed loop
    return v


def loader_permutation(n):
    """The index order one `iter(DataLoader(shuffle=True))` yields, consuming torch's DEFAULT generator exactly
    like torch.utils.data does (encode.py:69-70): one int64 `random_()` draw for the loader's base seed, one for
    the RandomSampler seed, then `randperm(n)` on a fresh generator seeded with the latter."""
    torch.empty((), dtype=torch.int64).random_()                       # _BaseDataLoaderIter._base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())    # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def device_permutation(n, seed, rounds=6):
    """CPU restatement of the library's own device sampler `lbdrn_randperm` (include/lbdrn.h; NOT a reference function:
    the reference shuffles with torch's RandomSampler, encode.py:69-70, which `loader_permutation` above restates).
    A keyed Feistel network over the ceil(log2 n) index bits -- each half XORed with a murmur3-finalizer hash of the
    other, round keys from splitmix64(seed) -- cycle-walked into [0, n).  Integer arithmetic: the GPU result must be
    bit-identical (tests/test_gpu_train.py)."""
    m64 = (1 << 64) - 1
    keys, x = [], seed & m64
    for _ in range(rounds):
        x = (x + 0x9E3779B97F4A7C15) & m64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m64
        keys.append(np.uint32(((z ^ (z >> 31)) >> 16) & 0xFFFFFFFF))
    bits = 1
    while (1 << bits) < n:
        bits += 1
    bits = max(bits, 2)
    bl = bits // 2
    mlo, mhi = np.uint32((1 << bl) - 1), np.uint32((1 << (bits - bl)) - 1)

    def mix(h):
        h = h ^ (h >> np.uint32(16))
        h = h * np.uint32(0x85EBCA6B)
        h = h ^ (h >> np.uint32(13))
        h = h * np.uint32(0xC2B2AE35)
        return h ^ (h >> np.uint32(16))

    out = np.empty(n, dtype=np.int64)
    todo, cur = np.arange(n), np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        while len(todo):
            lo = (cur & np.uint64(mlo)).astype(np.uint32)
            hi = ((cur >> np.uint64(bl)) & np.uint64(mhi)).astype(np.uint32)
            for r in range(0, rounds, 2):
                hi = hi ^ (mix(lo ^ keys[r]) & mhi)
                lo = lo ^ (mix(hi ^ keys[r + 1]) & mlo)
            cur = (hi.astype(np.uint64) << np.uint64(bl)) | lo.astype(np.uint64)
            done = cur < n
            out[todo[done]] = cur[done].astype(np.int64)
            todo, cur = todo[~done], cur[~done]
    return out


# ----------------------------------------------------------------------------------------------------------
# a7-a11: the encoder's optimisation loop (encode.py:67-117, modified_ignite_engine.py:18-27,38-43)
# ----------------------------------------------------------------------------------------------------------
def device_sampler_permutation(n):
    """Batch order of the library's `sampler="device"` mode (lbdrn_fused.FusedTrainer): the SAME default-generator draws as
    `loader_permutation` (so every other draw of a run stays aligned with the reference), but the RandomSampler seed keys
    `device_permutation` instead of `torch.randperm`."""
    torch.empty((), dtype=torch.int64).random_()
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    return torch.from_numpy(device_permutation(n, seed & ((1 << 64) - 1)))


def train(msb, lsb, D, bc, nl, lr, bs, epochs, flags=DEFAULT_FLAGS, val_duration=1, max_steps=None, sampler="loader",
          relu=False):
    """Overfit the network to one scene.  Call after `torch.manual_seed(seed)` for a fixed-seed run.
    Returns dict(params=best-epoch tensors, losses=[per-step], mses=[per-epoch], best_epoch).
    sampler="loader": the reference's DataLoader order (encode.py:69-70); "device": the library's own device sampler."""
    draw_perm = loader_permutation if sampler == "loader" else device_sampler_permutation
    C, H, W = msb.shape
    N = H * W
    X = torch.from_numpy(features(msb, D, flags))
    T = torch.from_numpy(labels(lsb))
    params = [p.requires_grad_(True) for p in init_params(X.shape[1], bc, C, nl)]
    opt = torch.optim.Adam(params, lr=lr)
    best, best_epoch, best_params, losses, mses = 1e6, -1, None, [], []
    for epoch in range(1, epochs + 1):
        for g in opt.param_groups:
            g["lr"] = lr_at_epoch(lr, epoch, epochs)
        perm = draw_perm(N)
        for s in range(0, N, bs):                                      # last partial batch kept
            idx = perm[s:s + bs]
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(forward(params, X[idx], relu=relu), T[idx])   # LBDRNloss.py:9
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
            if max_steps is not None and len(losses) >= max_steps:
                return dict(params=[p.detach().clone() for p in params], losses=losses, mses=mses,
                            best_epoch=epoch)
        if epochs == 1:                                                # encode.py:100-103
            best_epoch, best_params = epoch, [p.detach().clone() for p in params]
            break
        if epoch % min(val_duration, epochs) == 0:                     # encode.py:104-117
            loader_permutation(N)                                      # evaluator.run(train_loader) re-iterates
            with torch.no_grad():
                sse = 0.0
                for s in range(0, N, 1 << 18):
                    sse += float(((forward(params, X[s:s + (1 << 18)], relu=relu).double() - T[s:s + (1 << 18)].double()) ** 2).sum())
            mse = sse / (N * C)
            mses.append(mse)
            if mse < best:
                best, best_epoch, best_params = mse, epoch, [p.detach().clone() for p in params]
    return dict(params=best_params, losses=losses, mses=mses, best_epoch=best_epoch)


# ----------------------------------------------------------------------------------------------------------
# container header (encode.py:37-64, decode.py:25-53) and quality read-out (decode.py:210-222)
# ----------------------------------------------------------------------------------------------------------
def pack_header(sr, width, height, K, bc, nl, D, nn_sizes, base_sizes):
    n = 8 + 3 * len(nn_sizes) + 4 * len(base_sizes)
    b = bytes([n, sr]) + width.to_bytes(2, "big") + height.to_bytes(2, "big")
    b += bytes([K * 16 + D, int(np.log2(bc)) * 16 + nl])
    b += b"".join(v.to_bytes(3, "big") for v in nn_sizes)
    b += b"".join(v.to_bytes(4, "big") for v in base_sizes)
    assert len(b) == n
    return b


def unpack_header(b):
    n, sr = b[0], b[1]
    width, height = int.from_bytes(b[2:4], "big"), int.from_bytes(b[4:6], "big")
    K, D, bc, nl = b[6] >> 4, b[6] & 15, 2 ** (b[7] >> 4), b[7] & 15
    p, t = 8, sr * sr
    nn = [int.from_bytes(b[p + 3 * i:p + 3 * i + 3], "big") for i in range(t)]
    p += 3 * t
    base = [int.from_bytes(b[p + 4 * i:p + 4 * i + 4], "big") for i in range(t)]
    return n, sr, width, height, K, bc, nl, D, nn, base


def quality(org, rec, n_bytes=None):
    """MSE in float32, PSNR with the fixed peak 10000, bits per sub-pixel (decode.py:216-222)."""
    mse = np.mean((org.astype(np.float32) - rec.astype(np.float32)) ** 2)
    psnr = 10 * np.log10(10000 ** 2 / mse)
    bpsp = None if n_bytes is None else n_bytes * 8 / np.prod(org.shape)
    return float(mse), float(psnr), bpsp
