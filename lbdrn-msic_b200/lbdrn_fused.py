"""Host side of the fused LBDRN path: keeps the scene resident on the GPU as integer MSB/LSB planes and drives the
C-ABI kernels of liblbdrn_b200 (decode, predict, full-scene MSE, fused training).  Never materialises the
(H*W, dim_in) feature matrix the reference builds on the host (LBDRNdataset.py:104-130, decode.py:77-102).

There is no CPU fallback here: every function needs a CUDA device and the native library.
"""
import ctypes
import math

import numpy as np
import torch

import lbdrn_cabi as cabi


# ----------------------------------------------------------------------------------------------------------------
# feature flags
# ----------------------------------------------------------------------------------------------------------------
class Flags:
    """Snapshot of the codec's feature switches (module globals of constants.py)."""

    def __init__(self, use_coordinates=False, embedding=False, sigma=1.4, n_freq=12, use_colors=True, relative=True):
        self.use_coordinates, self.embedding = bool(use_coordinates), bool(embedding)
        self.sigma, self.n_freq = float(sigma), int(n_freq)
        self.use_colors, self.relative = bool(use_colors), bool(relative)

    @classmethod
    def from_constants(cls):
        import constants as k
        return cls(k.USE_COORDINATES, k.EMBEDDING, k.SIGMA, k.N_FREQ, k.USE_COLORS, k.RELATIVE)

    @property
    def tabw(self):
        return 2 * self.n_freq * int(self.embedding) + 1

    def num_coords(self):
        return self.tabw * 2 * int(self.use_coordinates)

    def dim_in(self, C, D):
        """Feature count (LBDRNdataset.py:104-106)."""
        return self.num_coords() + C * (2 * D + 1) ** 2 * int(self.use_colors)

    def bits(self, relu=False):
        return cabi.flag_bits(self.use_coordinates, self.embedding, self.use_colors, self.relative, relu)


def _flags(flags):
    return Flags.from_constants() if flags is None else flags


def _device(device=None):
    if not torch.cuda.is_available():
        raise cabi.LbdrnError(cabi.E_CUDA, "no CUDA device: the LBDRN fused path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def coord_table(H, W, flags):
    """float32 [(H+W), tabw]: per-row then per-column coordinate / positional-encoding entries.

    The reference evaluates, per pixel, v = float32(2*i/(n-1) - 1) and [v, sin(sigma^k pi v), cos(sigma^k pi v)] in
    float64, stored as float32 (LBDRNdataset.py:108-118).  Each entry depends only on the row or only on the column,
    so an (H+W) x tabw table carries exactly the same values."""
    def axis(n):
        v = (2 * np.arange(n) / (n - 1) - 1).astype(np.float32)
        if not flags.embedding:
            return v[:, None]
        arg = (flags.sigma ** np.arange(flags.n_freq) * np.pi) * v[:, None]      # float64
        return np.concatenate([v[:, None], np.sin(arg), np.cos(arg)], axis=1).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([axis(H), axis(W)], axis=0), dtype=np.float32)


# ----------------------------------------------------------------------------------------------------------------
# scene residency
# ----------------------------------------------------------------------------------------------------------------
class DeviceScene:
    """MSB / LSB planes of one scene on the GPU (a1 of the hot path: LBDRNdataset.py:93-100)."""

    def __init__(self, msb, lsb, K, msb_max):
        self.msb, self.lsb, self.K, self.msb_max = msb, lsb, K, int(msb_max)
        self.C, self.H, self.W = msb.shape

    @property
    def msb_u16(self):
        return self.msb.dtype == torch.uint16

    @classmethod
    def from_image(cls, img, K, device=None):
        """img: CHW (or HW) uint16/uint8 numpy array or torch tensor; the split runs on the device."""
        dev = _device(device)
        lib = cabi.load()
        if isinstance(img, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(img.astype(np.uint16, copy=False)))
        else:
            t = img if img.dtype == torch.uint16 else img.to(torch.int32).to(torch.uint16)
        t = t.reshape((-1,) + tuple(t.shape[-2:])).to(dev).contiguous()
        n = t.numel()
        mx = torch.zeros(1, dtype=torch.int32, device=dev)
        cabi.check(lib.lbdrn_max_shifted(cabi.ptr(t), n, K, cabi.ptr(mx), cabi.stream_ptr()))
        msb_max = int(mx.item())
        u16 = msb_max > 255                                   # LBDRNdataset.py:100
        msb = torch.empty(t.shape, dtype=torch.uint16 if u16 else torch.uint8, device=dev)
        lsb = torch.empty(t.shape, dtype=torch.uint16 if K > 8 else torch.uint8, device=dev)
        cabi.check(lib.lbdrn_split(cabi.ptr(t), n, K, cabi.U16 if u16 else cabi.U8, cabi.ptr(msb), cabi.ptr(lsb),
                                   cabi.stream_ptr()))
        return cls(msb, lsb, K, msb_max)

    def desc(self, D, bc, nl, flags, relu=False, w0=30.0, path=cabi.PATH_AUTO, row0=0, row1=None):
        return cabi.make_desc(self.C, self.H, self.W, self.K, D, bc, nl, flags.bits(relu), self.msb_max, self.msb_u16,
                              row0=row0, row1=row1, w0=w0, n_freq=flags.n_freq, path=path)


def _base_to_device(base, dev, known_max=None):
    """CHW base layer (numpy u8/u16 or tensor) -> device tensor keeping u8 when it is u8, else u16, and its max."""
    if isinstance(base, np.ndarray):
        base = base.reshape((-1,) + base.shape[-2:])
        if base.dtype not in (np.uint8, np.uint16):
            base = base.astype(np.uint16)
        t = torch.from_numpy(np.ascontiguousarray(base))
    else:
        t = base.reshape((-1,) + tuple(base.shape[-2:]))
        if t.dtype not in (torch.uint8, torch.uint16):
            t = t.to(torch.int32).to(torch.uint16)
    t = t.to(dev, non_blocking=True).contiguous()
    if known_max is not None:
        return t, int(known_max)
    if t.dtype == torch.uint8:
        return t, int(t.max().item())
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    cabi.check(cabi.load().lbdrn_max_shifted(cabi.ptr(t), t.numel(), 0, cabi.ptr(mx), cabi.stream_ptr()))
    return t, int(mx.item())


_PATHS = {"auto": cabi.PATH_AUTO, "precise": cabi.PATH_PRECISE, "tensor": cabi.PATH_TENSOR,
          "tensor_fastsin": cabi.PATH_TENSOR_FASTSIN, "tensor_fastsin2": cabi.PATH_TENSOR_FASTSIN2}


def _tab_tensor(H, W, flags, dev):
    if not flags.use_coordinates:
        return None
    return torch.from_numpy(coord_table(H, W, flags)).to(dev)


def decode_image(base, flat_params, K, D, bc, nl, flags=None, relu=False, w0=30.0, path="auto", device=None,
                 return_tensor=False, out_host=None, base_max=None):
    """Fused decode of one scene: uint16 CHW = (base << K) + round_half_even(net(features(base)) * (2^K-1)).
    Replaces decode.py:77-134.  `base` may be a numpy array / CPU tensor (copied to the device; pinned memory makes
    the copy asynchronous) or a CUDA tensor.  `out_host`: optional pinned uint16 CPU tensor receiving the result.
    `base_max`: the global `base.max()` if the caller already knows it (saves a reduction + sync)."""
    flags, dev, lib = _flags(flags), _device(device), cabi.load()
    msb, mx = _base_to_device(base, dev, base_max)
    C, H, W = msb.shape
    params = torch.as_tensor(flat_params, dtype=torch.float32).to(dev).contiguous()
    d = cabi.make_desc(C, H, W, K, D, bc, nl, flags.bits(relu), mx, msb.dtype == torch.uint16, w0=w0,
                       n_freq=flags.n_freq, path=_PATHS[path])
    if params.numel() != lib.lbdrn_param_count(ctypes.byref(d)):
        raise ValueError(f"parameter vector has {params.numel()} values, network needs "
                         f"{lib.lbdrn_param_count(ctypes.byref(d))}")
    tab = _tab_tensor(H, W, flags, dev)
    out = torch.empty((C, H, W), dtype=torch.uint16, device=dev)
    cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(msb), cabi.ptr(params), cabi.ptr(tab), cabi.ptr(out),
                                cabi.stream_ptr()))
    if out_host is not None:
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host
    return out if return_tensor else out.cpu().numpy()


_streams = {}


def msb_upper_bound(msb_dtype, K):
    """A TRUE upper bound of a base layer's maximum, for descriptors that carry the exact value on the device
    (LbdrnDesc.msb_max_dev): the library selects kernels from the host-side bound (the tcgen05 path needs MSB <= 2048 for
    its integer differences to be exact in fp16), so it must never under-state the maximum.  MSB = img >> K with img a
    16-bit word (LBDRNdataset.py:95), hence 65535 >> K for uint16 planes and 255 for uint8 ones."""
    return 255 if msb_dtype == torch.uint8 else (0xFFFF >> K)


def _get_streams(dev):
    key = (dev.type, dev.index)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _streams[key]


def decode_image_streamed(base_host, flat_params, K, D, bc, nl, flags=None, relu=False, w0=30.0, path="auto", device=None,
                          out_host=None, base_max=None, stripe_rows=1024):
    """End-to-end decode from HOST memory to HOST memory with the copies overlapped with the kernel (N1 of SURVEY.md 8f).

    The base layer is uploaded stripe by stripe on a copy stream; each stripe is decoded on a compute stream as soon as
    it and its lower halo are resident, and downloaded on a third stream while the next stripe computes.  The global
    normaliser `base.max()` is a property of the whole image: when the caller does not pass it, the upload is finished
    first and the maximum is reduced on the device (4-byte read-back) before the first stripe is decoded.
    base_host / out_host: CHW CPU tensors (pinned memory makes the copies asynchronous); returns out_host."""
    flags, dev, lib = _flags(flags), _device(device), cabi.load()
    if isinstance(base_host, np.ndarray):
        base_host = torch.from_numpy(np.ascontiguousarray(base_host))
    base_host = base_host.reshape((-1,) + tuple(base_host.shape[-2:]))
    if base_host.dtype not in (torch.uint8, torch.uint16):
        base_host = base_host.to(torch.int32).to(torch.uint16)
    C, H, W = base_host.shape
    if out_host is None:
        out_host = torch.empty((C, H, W), dtype=torch.uint16).pin_memory()
    s_in, s_cmp, s_out = _get_streams(dev)
    cur = torch.cuda.current_stream(dev)
    params = torch.as_tensor(flat_params, dtype=torch.float32).to(dev, non_blocking=True).contiguous()
    msb = torch.empty((C, H, W), dtype=base_host.dtype, device=dev)
    out = torch.empty((C, H, W), dtype=torch.uint16, device=dev)
    tab = _tab_tensor(H, W, flags, dev)
    bounds = list(range(0, H, stripe_rows)) + [H]
    n = len(bounds) - 1
    up = [torch.cuda.Event() for _ in range(n)]
    done = [torch.cuda.Event() for _ in range(n)]
    ready = torch.cuda.Event()
    ready.record(cur)
    for st in (s_in, s_cmp, s_out):
        st.wait_event(ready)
    with torch.cuda.stream(s_in):
        for i in range(n):
            r0, r1 = bounds[i], bounds[i + 1]
            for c in range(C):
                msb[c, r0:r1].copy_(base_host[c, r0:r1], non_blocking=True)
            up[i].record(s_in)
        max_dev = None
        if base_max is None:
            # max is a whole-image property: reduced on the device once the upload is complete and handed to the kernels
            # through LbdrnDesc.msb_max_dev -- no host round trip; the descriptor carries the dtype's upper bound
            max_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            if msb.dtype == torch.uint8:
                max_dev.copy_(msb.max().to(torch.int32).reshape(1))
                base_max = 255
            else:
                # uint16 planes: the exact max decides whether the tensor kernel (MSB <= 2048) applies, so it is read back
                cabi.check(lib.lbdrn_max_shifted(cabi.ptr(msb), msb.numel(), 0, cabi.ptr(max_dev), cabi.stream_ptr()))
                base_max = int(max_dev.item())
        up_all = torch.cuda.Event()
        up_all.record(s_in)
    for i in range(n):
        r0, r1 = bounds[i], bounds[i + 1]
        with torch.cuda.stream(s_cmp):
            # stripe i and the halo rows of stripe i+1 are resident (and, with a device-side max, the whole upload)
            s_cmp.wait_event(up[min(i + 1, n - 1)] if max_dev is None else up_all)
            d = cabi.make_desc(C, H, W, K, D, bc, nl, flags.bits(relu), base_max, msb.dtype == torch.uint16, row0=r0,
                               row1=r1, w0=w0, n_freq=flags.n_freq, path=_PATHS[path], msb_max_dev=max_dev)
            cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(msb), cabi.ptr(params), cabi.ptr(tab), cabi.ptr(out),
                                        cabi.stream_ptr()))
            done[i].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done[i])
            for c in range(C):
                out_host[c, r0:r1].copy_(out[c, r0:r1], non_blocking=True)
    fin = torch.cuda.Event()
    fin.record(s_out)
    fin.synchronize()
    cur.wait_stream(s_cmp)
    for t in (msb, out, params) + ((max_dev,) if max_dev is not None else ()):
        t.record_stream(s_cmp)
    return out_host


class StreamedDecoder:
    """Throughput-oriented host->host decoder for a sequence of same-shaped scenes (base-layer hand-off, N1 of
    SURVEY.md 8f).  `submit()` queues upload, device-side max, stripe kernels and download of one scene on three streams
    and returns immediately; with two buffer slots the upload of scene i+1 overlaps the kernels and the download of scene
    i, so a steady stream of scenes runs at the slowest of (H2D, kernel, D2H) instead of their sum.  `wait()` blocks until
    a submitted scene's reconstruction is complete in its host buffer."""

    def __init__(self, C, H, W, msb_dtype, K, D, bc, nl, flat_params, flags=None, relu=False, w0=30.0, path="auto",
                 device=None, stripe_rows=1024, slots=2):
        self.flags, self.dev, self.lib = _flags(flags), _device(device), cabi.load()
        self.C, self.H, self.W, self.K, self.D, self.bc, self.nl = C, H, W, K, D, bc, nl
        self.relu, self.w0, self.path = relu, w0, _PATHS[path]
        self.u16 = msb_dtype == torch.uint16
        self.params = torch.as_tensor(flat_params, dtype=torch.float32).to(self.dev).contiguous()
        self.tab = _tab_tensor(H, W, self.flags, self.dev)
        self.s_in, self.s_cmp, self.s_out = _get_streams(self.dev)
        self.bounds = list(range(0, H, stripe_rows)) + [H]
        self.slots = [dict(msb=torch.empty((C, H, W), dtype=msb_dtype, device=self.dev),
                           out=torch.empty((C, H, W), dtype=torch.uint16, device=self.dev),
                           mx=torch.zeros(1, dtype=torch.int32, device=self.dev), cmp_done=None, out_done=None)
                      for _ in range(slots)]
        self.n = 0
        torch.cuda.current_stream(self.dev).synchronize()

    def submit(self, base_host, out_host):
        """base_host: [C,H,W] uint8/uint16 CPU tensor (pinned); out_host: [C,H,W] uint16 CPU tensor (pinned)."""
        sl = self.slots[self.n % len(self.slots)]
        self.n += 1
        C, lib = self.C, self.lib
        with torch.cuda.stream(self.s_in):
            if sl["cmp_done"] is not None:
                self.s_in.wait_event(sl["cmp_done"])           # the kernels that read this slot's planes are done
            for c in range(C):
                sl["msb"][c].copy_(base_host[c], non_blocking=True)
            if self.u16:
                sl["mx"].zero_()
                cabi.check(lib.lbdrn_max_shifted(cabi.ptr(sl["msb"]), sl["msb"].numel(), 0, cabi.ptr(sl["mx"]),
                                                 cabi.stream_ptr()))
            else:
                sl["mx"].copy_(sl["msb"].max().to(torch.int32).reshape(1))
            up = torch.cuda.Event()
            up.record(self.s_in)
        bound = msb_upper_bound(sl["msb"].dtype, self.K)       # a true bound; msb_max_dev carries the exact value
        last = None
        for i in range(len(self.bounds) - 1):
            r0, r1 = self.bounds[i], self.bounds[i + 1]
            with torch.cuda.stream(self.s_cmp):
                if i == 0:
                    self.s_cmp.wait_event(up)
                    if sl["out_done"] is not None:
                        self.s_cmp.wait_event(sl["out_done"])  # the previous download from this slot's output is done
                d = cabi.make_desc(C, self.H, self.W, self.K, self.D, self.bc, self.nl, self.flags.bits(self.relu), bound,
                                   self.u16, row0=r0, row1=r1, w0=self.w0, n_freq=self.flags.n_freq, path=self.path,
                                   msb_max_dev=sl["mx"])
                cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(sl["msb"]), cabi.ptr(self.params), cabi.ptr(self.tab),
                                            cabi.ptr(sl["out"]), cabi.stream_ptr()))
                ev = torch.cuda.Event()
                ev.record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev)
                for c in range(C):
                    out_host[c, r0:r1].copy_(sl["out"][c, r0:r1], non_blocking=True)
                last = torch.cuda.Event()
                last.record(self.s_out)
        sl["cmp_done"], sl["out_done"] = ev, last
        return last

    @staticmethod
    def wait(ticket):
        ticket.synchronize()


def image_mse(a, b, device=None):
    """mean((a-b)^2) of two uint16 images on the device (quality read-out of decode.py:216); exact integer sum."""
    dev, lib = _device(device), cabi.load()
    def up(x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x.astype(np.uint16, copy=False)))
        return x.to(dev).contiguous()
    ta, tb = up(a), up(b)
    if ta.numel() != tb.numel():
        raise ValueError("image sizes differ")
    acc = torch.zeros(1, dtype=torch.int64, device=dev)
    cabi.check(lib.lbdrn_sse_u16(cabi.ptr(ta), cabi.ptr(tb), ta.numel(), cabi.ptr(acc), cabi.stream_ptr()))
    return float(acc.item()) / ta.numel()


def predict_image(base, flat_params, D, bc, nl, flags=None, relu=False, w0=30.0, device=None):
    """Network output y [H*W, C] (float32 CUDA tensor) for every pixel of the base layer."""
    flags, dev, lib = _flags(flags), _device(device), cabi.load()
    msb, mx = _base_to_device(base, dev)
    C, H, W = msb.shape
    params = torch.as_tensor(flat_params, dtype=torch.float32).to(dev).contiguous()
    d = cabi.make_desc(C, H, W, 1, D, bc, nl, flags.bits(relu), mx, msb.dtype == torch.uint16, w0=w0,
                       n_freq=flags.n_freq)
    tab = _tab_tensor(H, W, flags, dev)
    y = torch.empty((H * W, C), dtype=torch.float32, device=dev)
    cabi.check(lib.lbdrn_predict(ctypes.byref(d), cabi.ptr(msb), cabi.ptr(params), cabi.ptr(tab), cabi.ptr(y),
                                 cabi.stream_ptr()))
    return y


def eval_mse(scene, flat_params_dev, D, bc, nl, flags=None, relu=False, w0=30.0, tab=None, path="auto"):
    """Full-scene MSE of the network against the LSB labels (encode.py:105-108 / LBDRNperformance.py:18-21)."""
    flags, lib = _flags(flags), cabi.load()
    dev = scene.msb.device
    d = scene.desc(D, bc, nl, flags, relu, w0, path=_PATHS[path])
    if tab is None:
        tab = _tab_tensor(scene.H, scene.W, flags, dev)
    sse = torch.zeros(1, dtype=torch.float64, device=dev)
    cabi.check(lib.lbdrn_eval_sse(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(scene.lsb), cabi.ptr(flat_params_dev),
                                  cabi.ptr(tab), cabi.ptr(sse), cabi.stream_ptr()))
    return float(sse.item()) / (scene.C * scene.H * scene.W)


# ----------------------------------------------------------------------------------------------------------------
# schedule and sampling (encode.py:69-70,85,98)
# ----------------------------------------------------------------------------------------------------------------
def lr_for_epoch(lr, epoch, epochs):
    """StepLR(step_size=max(1,int(E/3)), gamma=0.1) stepped once per finished epoch; `epoch` is 1-based."""
    v = lr
    for _ in range((epoch - 1) // max(1, int(epochs / 3))):
        v = v * 0.1
    return v


def draw_loader_seeds():
    """Consume torch's default generator exactly as one `iter(DataLoader(shuffle=True))` does: an int64 draw for the
    loader's base seed, then one for the RandomSampler's seed.  Returns the sampler seed."""
    torch.empty((), dtype=torch.int64).random_()
    return int(torch.empty((), dtype=torch.int64).random_().item())


def torch_permutation_from_seed(n, seed, out=None):
    """torch.randperm(n) on a CPU generator seeded with `seed`: what RandomSampler yields for the reference's DataLoader."""
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g) if out is None else torch.randperm(n, generator=g, out=out)


def permutation_from_seed(n, seed, out=None, dtype=torch.int64, progress=None):
    """The same permutation, from the library's restatement of torch's CPU randperm (`lbdrn_host_randperm`: forward
    Fisher-Yates on MT19937 outputs with the draws running ahead of the swaps, ~2x faster at 67 M pixels; bit-exactness
    against torch is a CPU test).  dtype=torch.int32 gives the same order as 32-bit indices (half the traffic).  Sizes
    outside the library's range (n >= 2^32 / 20: torch switches algorithm) go to torch."""
    if n < (1 << 32) // 20:
        buf = torch.empty(n, dtype=dtype) if out is None else out
        assert buf.dtype in (torch.int64, torch.int32) and buf.is_contiguous() and buf.numel() == n and buf.device.type == "cpu"
        if progress is not None:
            # `progress` (ctypes.c_int64) counts the leading entries that are final while the call runs (32-bit orders only)
            assert buf.dtype == torch.int32
            if cabi.load().lbdrn_host_randperm32_progress(n, seed & 0xFFFFFFFFFFFFFFFF, buf.data_ptr(), ctypes.addressof(progress)) == 0:
                return buf
        fn = cabi.load().lbdrn_host_randperm if buf.dtype == torch.int64 else cabi.load().lbdrn_host_randperm32
        if fn(n, seed & 0xFFFFFFFFFFFFFFFF, buf.data_ptr()) == 0:
            if progress is not None:
                progress.value = n
            return buf
    perm = torch_permutation_from_seed(n, seed, out if (out is not None and out.dtype == torch.int64) else None)
    if out is not None and out.dtype != torch.int64:
        out.copy_(perm)
        return out
    if progress is not None:
        progress.value = n
    return perm if dtype == torch.int64 else perm.to(dtype)


class HostPermutations:
    """The reference's batch orders, produced ahead of the device.

    `torch.randperm` on the CPU is a sequential Fisher-Yates shuffle (about 3 s for the 67 M pixels of an 8192^2 scene --
    17x the 0.18 s the device needs for the epoch it feeds), but every epoch's seed is known before training starts
    (`FusedTrainer._plan_seeds`), so the permutations of up to `workers` epochs are drawn concurrently in threads (torch
    releases the GIL; the native shuffle `lbdrn_host_randperm32` is called through ctypes, which does too), each into its
    own buffer (32-bit indices when they suffice; pageable unless `pin`).  `get(e)` blocks until epoch e's order is
    ready; `release(e, event)` hands its buffer back once `event` (the upload) has completed.  Results are exactly
    `permutation_from_seed(n, seeds[e-1])`, in epoch order."""

    def __init__(self, seeds, n, workers=None, pin=False, device=None, max_bytes=4 << 30, dtype=None, stream_first=False):
        import concurrent.futures
        import os
        self.seeds, self.n, self.device = list(seeds), n, device
        # 32-bit indices whenever they suffice: half the shuffle's memory traffic and half the upload (widened on the device)
        self.dtype = dtype or (torch.int32 if n < (1 << 31) and n < (1 << 32) // 20 else torch.int64)
        esz = 4 if self.dtype == torch.int32 else 8
        if workers is None:                          # bounded by the epochs, the host cores and `max_bytes` of buffers
            workers = min(len(self.seeds), max(1, (os.cpu_count() or 2) - 2), 10, max_bytes // max(1, esz * n) - 1)
        if os.environ.get("LBDRN_PERM_WORKERS"):      # experiments: fewer concurrent shuffles finish sooner each
            workers = min(workers, int(os.environ["LBDRN_PERM_WORKERS"]))
        self.workers = workers = max(1, workers)
        # pageable buffers by default: pinning 10 x 268 MB costs ~1 s up front (the driver serialises it), more than the staged
        # uploads lose -- those run on the side stream while the previous epoch trains
        self.pin = bool(pin) and torch.cuda.is_available()
        self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=workers)
        self.free, self.n_buffers = [], 0            # idle buffers; buffers allocated so far (at most workers + 1)
        self.busy = {}                               # epoch -> (buffer, upload event or None)
        self.fut = {}
        self._first_fut = None
        self.next_epoch = 1
        # epoch 1's order is the only one nothing can hide: its buffer is allocated here and its progress published, so the
        # trainer can start on the head of the order while the tail is still being shuffled (`first_stream`)
        self.first_progress = ctypes.c_int64(0) if (stream_first and self.dtype == torch.int32 and n < (1 << 32) // 20) else None
        self.first_buf = torch.empty(n, dtype=self.dtype) if self.first_progress is not None else None
        self._fill()

    def _draw(self, e, buf):
        if buf is None:
            if self.pin and self.device is not None:
                torch.cuda.set_device(self.device)   # worker threads start on device 0: pin in this rank's context
            buf = torch.empty(self.n, dtype=self.dtype, pin_memory=self.pin)
        return permutation_from_seed(self.n, self.seeds[e - 1], out=buf,
                                     progress=self.first_progress if (e == 1 and buf is self.first_buf) else None)

    def first_stream(self):
        """(buffer, progress, future) of epoch 1 while it is being drawn: entries [0, progress.value) are final.  The caller
        releases epoch 1 like any other.  None when the order is not streamable (64-bit indices, torch's other algorithm)."""
        if self.first_progress is None or 1 not in self.fut:
            return None
        fut = self.fut.pop(1)
        self.busy[1] = (self.first_buf, None)
        self._fill()
        return self.first_buf, self.first_progress, fut

    def _reclaim(self, block):
        """Move buffers whose upload has finished back to the free list; with `block`, wait for the oldest one."""
        for e in sorted(self.busy):
            buf, ev = self.busy[e]
            if ev is None:
                continue                             # handed out, not released yet
            if block or ev is True or ev.query():
                if ev is not True:
                    ev.synchronize()
                self.free.append(buf)
                del self.busy[e]
                block = False

    def _limit(self):
        """Concurrent shuffles: they share the memory system, so until epoch 1's order exists (the one nothing hides) only
        four run at once -- measured on a 16-core box: 1.87-1.91 s per 8192^2 encode against 2.11-2.15 s with ten at once."""
        first = self._first_fut
        if first is None or first.done():
            return self.workers
        return min(self.workers, 4)

    def _fill(self):
        while self.next_epoch <= len(self.seeds):
            running = len(self.fut) + (1 if (self._first_fut is not None and 1 not in self.fut and not self._first_fut.done()) else 0)
            if running >= self._limit():
                return
            self._reclaim(block=False)
            if self.free:
                buf = self.free.pop()
            elif self.n_buffers < self.workers + 1:
                buf, self.n_buffers = None, self.n_buffers + 1       # allocated inside the worker
                if self.next_epoch == 1 and self.first_buf is not None:
                    buf = self.first_buf
            else:
                return
            e = self.next_epoch
            self.next_epoch += 1
            self.fut[e] = self.pool.submit(self._draw, e, buf)
            if e == 1:
                self._first_fut = self.fut[e]

    def get(self, e):
        if e not in self.fut:                        # every buffer is out: wait for an upload to finish, then draw
            self._reclaim(block=True)
            self._fill()
        if e not in self.fut:
            raise RuntimeError(f"permutation of epoch {e} was not scheduled (epochs must be taken in order and released)")
        perm = self.fut.pop(e).result()
        self.busy[e] = (perm, None)
        self._fill()
        return perm

    def release(self, e, event=None):
        buf, _ = self.busy[e]
        self.busy[e] = (buf, True if event is None else event)
        self._fill()

    def close(self):
        self.pool.shutdown(wait=True)


class FusedTrainer:
    """The encoder's optimisation loop on the device (replaces trainer.run / evaluator.run of encode.py:84-117,157).

    sampler="reference": every epoch's batch order is the permutation the reference's DataLoader would produce from
    the same torch seed (host randperm -- ~3 s per 67 M pixels -- drawn several epochs ahead by `HostPermutations`,
    uploaded as int64 indices): bit-for-bit the reference's batches, host-bound for large scenes.
    sampler="device": permutation written on the GPU by `lbdrn_randperm` (a keyed bijection; every pixel once per epoch,
    not torch's order; no host work) -- the mode for throughput (`encode.py --sampler device`).
    Adam(lr, betas=(0.9,0.999), eps=1e-8), StepLR, per-epoch full-scene MSE and best-epoch selection follow
    encode.py:84-85,96-117 (including the epochs==1 special case, which skips evaluation)."""

    def __init__(self, model, scene, D, lr, batch_size, epochs, val_duration=1, flags=None, sampler="reference",
                 on_epoch=None, betas=(0.9, 0.999), eps=1e-8):
        self.model, self.scene, self.D = model, scene, D
        self.lr, self.bs, self.epochs, self.val_duration = float(lr), int(batch_size), int(epochs), int(val_duration)
        self.flags = _flags(flags)
        self.sampler, self.on_epoch = sampler, on_epoch
        self.bc, self.nl = model.dim_hidden, model.num_layers
        model._require_fused()
        self.relu, self.w0 = model._relu, model.w0
        if model.dim_in != self.flags.dim_in(scene.C, D) or model.dim_out != scene.C:
            raise ValueError("model shape does not match the scene / feature flags")
        self.lib = cabi.load()
        self.dev = scene.msb.device
        self.desc = scene.desc(D, self.bc, self.nl, self.flags, self.relu, self.w0)
        cfg = cabi.LbdrnTrainCfg()
        cfg.batch_size, cfg.world_size, cfg.rank = self.bs, 1, 0
        cfg.beta1, cfg.beta2, cfg.eps = float(betas[0]), float(betas[1]), float(eps)   # copied into the handle at create
        self.cfg = cfg
        self.handle = ctypes.c_void_p()
        cabi.check(self.lib.lbdrn_train_create(ctypes.byref(self.desc), ctypes.byref(cfg), ctypes.byref(self.handle)))
        self.tab = _tab_tensor(scene.H, scene.W, self.flags, self.dev)
        self.losses, self.val_mse, self.best_epoch, self.best_mse = [], [], -1, 1e6
        self.best_params = None

    def close(self):
        if self.handle:
            self.lib.lbdrn_train_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ----------------------------------------------------------------------------------------------
    def _plan_seeds(self):
        """All default-generator draws of the run, in the reference's order: train iterator of epoch e, then (when the
        epoch is evaluated) the evaluator's iterator over the same shuffled loader."""
        seeds = []
        for e in range(1, self.epochs + 1):
            seeds.append(draw_loader_seeds())
            if self.epochs != 1 and e % min(self.val_duration, self.epochs) == 0:
                draw_loader_seeds()
        return seeds

    def begin(self):
        """Upload the model's current parameters into the training handle and reset the step counter."""
        self.params_dev = self.model.flat_params().to(self.dev).contiguous()
        cabi.check(self.lib.lbdrn_train_set_params(self.handle, cabi.ptr(self.params_dev), cabi.stream_ptr()))
        self.adam_t = 0

    def train_epoch(self, perm_dev, lr, losses_out=None):
        """One pass over `perm_dev` (int64 pixel indices on the device) in batches of `bs`; returns the per-step
        losses (device tensor; `losses_out` when given: a slice of an epoch's loss vector).  One persistent kernel launch."""
        n = perm_dev.numel()
        steps = math.ceil(n / self.bs)
        losses = torch.empty(steps, dtype=torch.float32, device=self.dev) if losses_out is None else losses_out
        assert losses.numel() == steps and losses.is_contiguous()
        sc = self.scene
        cabi.check(self.lib.lbdrn_train_steps(self.handle, cabi.ptr(sc.msb), cabi.ptr(sc.lsb), cabi.ptr(self.tab),
                                              cabi.ptr(perm_dev), n, steps, self.adam_t, float(lr), cabi.ptr(losses),
                                              cabi.stream_ptr()))
        self.adam_t += steps
        return losses

    def current_params(self):
        """Flat parameter vector after the steps taken so far (new device tensor)."""
        cur = torch.empty_like(self.params_dev)
        cabi.check(self.lib.lbdrn_train_get_params(self.handle, cabi.ptr(cur), cabi.stream_ptr()))
        return cur

    def scene_mse(self, params_dev):
        return eval_mse(self.scene, params_dev, self.D, self.bc, self.nl, self.flags, self.relu, self.w0, self.tab)

    def run(self):
        """Train; returns dict(params=best flat params (CPU tensor), losses, val_mse, best_epoch).

        Stream plan: the training launches of consecutive epochs run back to back on a high-priority stream (one
        cooperative launch per epoch: 128 of the 148 SMs at bs 8192).  The full-scene evaluation of epoch e and the device
        permutation of epoch e+2 run on a low-priority side stream, i.e. on the idle SMs and in the gaps, while epoch e+1
        trains; the host reads epoch e's MSE only after epoch e+1 has been queued, so no launch waits for a read-back.
        Results are identical to the serial order (same permutations, same snapshots)."""
        N = self.scene.H * self.scene.W
        cur_stream = torch.cuda.current_stream(self.dev)
        main = torch.cuda.Stream(self.dev, priority=-1)
        side = torch.cuda.Stream(self.dev, priority=0)
        main.wait_stream(cur_stream)
        with torch.cuda.stream(main):
            self.begin()
        seeds = self._plan_seeds()
        import os
        import time
        stream_first = os.environ.get("LBDRN_NO_STREAM_FIRST") is None and math.ceil(N / self.bs) >= 64
        host = HostPermutations(seeds, N, device=self.dev, stream_first=stream_first) if self.sampler == "reference" else None

        def device_perm(e):
            """(perm, event): drawn on the side stream"""
            with torch.cuda.stream(side):
                # lbdrn_randperm: a keyed Feistel bijection written in one pass (0.2 ms at 8192^2); torch.randperm sorts
                # 67 M random keys (82 ms alone, several GB of traffic next to the latency-bound training kernel)
                perm = torch.empty(N, dtype=torch.int64, device=self.dev)
                cabi.check(self.lib.lbdrn_randperm(N, seeds[e - 1] & 0xFFFFFFFFFFFFFFFF, cabi.ptr(perm), cabi.stream_ptr()))
                ev = torch.cuda.Event()
                ev.record(side)
            return perm, ev

        pending = []                                                    # evaluations queued on the side stream

        def collect(upto=None):
            """Read back finished evaluations in epoch order: best-epoch selection of encode.py:104-117."""
            while pending and (upto is None or pending[0][0] <= upto):
                e, sse, cur, done = pending.pop(0)
                done.synchronize()                                      # .item() only orders against the CURRENT stream
                mse = float(sse.item()) / (self.scene.C * N)
                self.val_mse.append(mse)
                improved = mse < self.best_mse
                if improved:
                    self.best_mse, self.best_epoch, self.best_params = mse, e, cur
                if self.on_epoch:
                    self.on_epoch(e, mse, improved)

        upl = torch.cuda.Stream(self.dev, priority=0) if host is not None else None

        def host_perm(e):
            """(perm, event): the reference sampler's order of epoch e, uploaded (and widened to int64) on its own stream
            while the previous epoch trains; blocks the host until the order has been drawn"""
            with torch.cuda.stream(upl):
                perm = host.get(e).to(self.dev, non_blocking=True)
                up = torch.cuda.Event()
                up.record(upl)
                host.release(e, up)                                  # its buffer is reused once the upload has completed
                if perm.dtype != torch.int64:
                    perm = perm.to(torch.int64)
                ev = torch.cuda.Event()
                ev.record(upl)
            return perm, ev

        def first_epoch_streamed(first):
            """Epoch 1 of the reference sampler in slices of whole batches, each uploaded and trained as soon as the host
            shuffle (a forward Fisher-Yates: a finished prefix never changes) has got past it -- the one permutation whose
            drawing nothing else can hide.  Same batches, same launch arithmetic as the one-launch epoch."""
            buf, prog, fut = first
            steps = math.ceil(N / self.bs)
            with torch.cuda.stream(main):
                losses = torch.empty(steps, dtype=torch.float32, device=self.dev)
            per = max(1, steps // 8)
            lr1 = lr_for_epoch(self.lr, 1, self.epochs)
            s0, up = 0, None
            while s0 < steps:
                s1 = min(steps, s0 + per)
                need = min(N, s1 * self.bs)
                while prog.value < need:
                    if fut.done():
                        fut.result()                                     # re-raises a failed draw; progress is final by now
                        break
                    time.sleep(0.0002)
                with torch.cuda.stream(upl):
                    sl = buf[s0 * self.bs:need].to(self.dev, non_blocking=True).to(torch.int64)
                    up = torch.cuda.Event()
                    up.record(upl)
                with torch.cuda.stream(main):
                    main.wait_event(up)
                    sl.record_stream(main)
                    self.train_epoch(sl, lr1, losses_out=losses[s0:s1])
                s0 = s1
            fut.result()
            host.release(1, up)
            return losses

        next_perm = host_perm if host is not None else device_perm
        first = host.first_stream() if host is not None else None
        next_dev = next_perm(1) if first is None else None
        for e in range(1, self.epochs + 1):
            if e == 1 and first is not None:
                self.losses.append(first_epoch_streamed(first))
            with torch.cuda.stream(main):
                if not (e == 1 and first is not None):
                    perm, ev = next_dev
                    main.wait_event(ev)
                    perm.record_stream(main)
                    self.losses.append(self.train_epoch(perm, lr_for_epoch(self.lr, e, self.epochs)))
                else:
                    perm = None
                evaluate = self.epochs != 1 and e % min(self.val_duration, self.epochs) == 0
                cur = self.current_params() if (evaluate or self.epochs == 1) else None
                snap = torch.cuda.Event()
                snap.record(main)
            if e < self.epochs:
                next_dev = next_perm(e + 1)                             # overlaps this epoch's training
            if self.epochs == 1:                                        # encode.py:100-103
                self.best_epoch, self.best_params = e, cur
            elif evaluate:
                with torch.cuda.stream(side):
                    side.wait_event(snap)
                    cur.record_stream(side)
                    sse = self._eval_sse_async(cur)
                    done = torch.cuda.Event()
                    done.record(side)
                    pending.append((e, sse, cur, done))
            collect(upto=e - 1)                                         # epoch e-1's result: epoch e is already queued
            del perm
        collect()
        if host is not None:
            host.close()
        cur_stream.wait_stream(main)
        cur_stream.wait_stream(side)
        if upl is not None:
            cur_stream.wait_stream(upl)
        if self.best_params is None:                                    # never evaluated (val_duration > epochs)
            raise RuntimeError("no epoch was evaluated; choose val_duration <= epochs")
        losses = torch.cat(self.losses).cpu()
        best = self.best_params.cpu()
        self.model.load_flat_params(best)
        return dict(params=best, losses=losses.tolist(), val_mse=self.val_mse, best_epoch=self.best_epoch)

    def _eval_sse_async(self, params_dev):
        """Queue the full-scene squared error on the current stream; returns the device double (no host sync)."""
        sse = torch.zeros(1, dtype=torch.float64, device=self.dev)
        cabi.check(self.lib.lbdrn_eval_sse(ctypes.byref(self.desc), cabi.ptr(self.scene.msb), cabi.ptr(self.scene.lsb),
                                           cabi.ptr(params_dev), cabi.ptr(self.tab), cabi.ptr(sse), cabi.stream_ptr()))
        return sse
