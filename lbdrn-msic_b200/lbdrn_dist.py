"""Multi-GPU plumbing (one process per GPU, torch.distributed): row-stripe sharded decode and data-parallel /
scene-parallel encode.  The reference has no distributed code at all; this is the B200-side design.

Decode shards naturally: an output pixel depends on its (2D+1)^2 MSB neighbourhood, the global scalar MSB.max() and the
shared weights.  Each rank owns a contiguous row stripe; the only exchanges are a scalar max all-reduce and a ONE-OFF
D-row halo swap of the static MSB planes with the two neighbouring stripes (<= D*W*C*2 bytes per edge).  Reflect padding
is applied only at the true image border, never at stripe seams, so the result is bit-identical to the 1-GPU decode.

Encode, data-parallel: every rank holds the whole scene and takes a contiguous 1/G slice of every batch; one all-reduce
of the flat gradient vector (P+1 floats) per step, identical Adam update on every rank.  Latency-bound at bs=8192.
Encode, scene-parallel: independent scenes / tiles per rank, no communication (replicas only).
"""
import ctypes
import math

import torch
import torch.distributed as dist

import lbdrn_cabi as cabi


def stripe_bounds(H, world, rank):
    """Rows [r0, r1) of rank `rank` when H rows are cut into `world` contiguous, near-equal stripes."""
    base, extra = divmod(H, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def global_max(local_max, group=None):
    """MSB.max() over all stripes (LBDRNdataset.py:120 takes it over the whole image)."""
    t = local_max.clone().reshape(1) if torch.is_tensor(local_max) else torch.tensor([int(local_max)])
    if t.dtype not in (torch.int32, torch.int64):
        t = t.to(torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def global_max_dev(local_max_dev, group=None):
    """Device-side variant: all-reduce(max) of a 1-element int32 CUDA tensor in place, NO host read-back; pass the
    tensor to the kernels through LbdrnDesc.msb_max_dev."""
    dist.all_reduce(local_max_dev, op=dist.ReduceOp.MAX, group=group)
    return local_max_dev


def _bytes(t):
    """Byte view of a contiguous halo buffer for the wire: NCCL has no 16-bit integer type ("Short" is rejected), so
    uint16 planes (held as int16 views, which torch can slice and copy) travel as uint8."""
    return t if t.dtype == torch.uint8 else t.view(torch.uint8)


def exchange_halos(stripe, D, group=None):
    """stripe: [C, rows, W] tensor holding this rank's own rows.  Returns ([C, rows + top + bottom, W], top) where
    `top` halo rows came from rank-1 and the bottom ones from rank+1 (none at the image border).  Works on CUDA
    (NCCL send/recv over NVLink) and CPU (gloo) tensors; uint16 planes travel as int16 views."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if D == 0 or world == 1:
        return stripe, 0
    if stripe.shape[1] < D:
        raise ValueError("stripe thinner than the halo")
    wire = stripe.view(torch.int16) if stripe.dtype == torch.uint16 else stripe
    C, rows, W = wire.shape
    up = torch.empty((C, D, W), dtype=wire.dtype, device=wire.device) if rank > 0 else None
    down = torch.empty((C, D, W), dtype=wire.dtype, device=wire.device) if rank + 1 < world else None
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, _bytes(wire[:, :D].contiguous()), rank - 1, group),
                dist.P2POp(dist.irecv, _bytes(up), rank - 1, group)]
    if rank + 1 < world:
        ops += [dist.P2POp(dist.isend, _bytes(wire[:, rows - D:].contiguous()), rank + 1, group),
                dist.P2POp(dist.irecv, _bytes(down), rank + 1, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = ([up] if up is not None else []) + [wire] + ([down] if down is not None else [])
    buf = torch.cat(parts, dim=1).contiguous()
    return (buf.view(torch.uint16) if stripe.dtype == torch.uint16 else buf), (D if up is not None else 0)


class StripeBuffer:
    """This rank's stripe of the MSB image resident INSIDE a buffer that already has room for the halo rows, so a halo
    swap moves only 2*D*W*C elements (no re-assembly of the stripe)."""

    def __init__(self, H, W, C, D, dtype, device, group=None):
        self.group, self.D, self.H = group, D, H
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.r0, self.r1 = stripe_bounds(H, self.world, self.rank)
        self.top = D if (self.rank > 0 and D > 0) else 0
        self.bot = D if (self.rank + 1 < self.world and D > 0) else 0
        rows = self.r1 - self.r0
        if self.world > 1 and rows < D:
            raise ValueError("stripe thinner than the halo")
        self.buf = torch.empty((C, self.top + rows + self.bot, W), dtype=dtype, device=device)
        self.out = torch.empty_like(self.buf, dtype=torch.uint16)
        wire_dtype = torch.int16 if dtype == torch.uint16 else dtype
        mk = lambda: torch.empty((C, D, W), dtype=wire_dtype, device=device)
        self._send_up, self._send_dn, self._recv_up, self._recv_dn = (mk() if self.top else None, mk() if self.bot else None,
                                                                      mk() if self.top else None, mk() if self.bot else None)

    def _wire(self, t):
        return t.view(torch.int16) if t.dtype == torch.uint16 else t

    @property
    def own(self):
        """[C, rows, W] view of this rank's own rows (fill it with the stripe, e.g. own.copy_(...))."""
        return self._wire(self.buf)[:, self.top:self.top + (self.r1 - self.r0)]

    def load(self, stripe):
        self.own.copy_(self._wire(stripe), non_blocking=True)

    def exchange(self):
        """One D-row halo swap with each neighbouring stripe (NCCL send/recv over NVLink, or gloo on CPU)."""
        if not (self.top or self.bot):
            return
        w, D, rows = self._wire(self.buf), self.D, self.r1 - self.r0
        ops = []
        if self.top:
            self._send_up.copy_(w[:, self.top:self.top + D])
            ops += [dist.P2POp(dist.isend, _bytes(self._send_up), self.rank - 1, self.group),
                    dist.P2POp(dist.irecv, _bytes(self._recv_up), self.rank - 1, self.group)]
        if self.bot:
            self._send_dn.copy_(w[:, self.top + rows - D:self.top + rows])
            ops += [dist.P2POp(dist.isend, _bytes(self._send_dn), self.rank + 1, self.group),
                    dist.P2POp(dist.irecv, _bytes(self._recv_dn), self.rank + 1, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if self.top:
            w[:, :D].copy_(self._recv_up)
        if self.bot:
            w[:, self.top + rows:].copy_(self._recv_dn)

    def _max_args(self, msb_max, K):
        """msb_max may be an int (host value) or a 1-element int32 CUDA tensor (device value, e.g. fresh from an
        all-reduce): then the descriptor carries a TRUE upper bound (255, or 65535 >> K for uint16 planes: the library
        selects kernels from it) and the device pointer."""
        if torch.is_tensor(msb_max):
            return (255 if self.buf.dtype == torch.uint8 else (0xFFFF >> K)), msb_max
        return int(msb_max), None

    def decode_rows(self, row_a, row_b, flat_params_dev, K, bc, nl, flags, msb_max, relu=False, w0=30.0,
                    path=cabi.PATH_AUTO, tab=None):
        """Decode image rows [row_a, row_b) of this rank's stripe into `self.out` (current stream).  Rows closer than D to
        a stripe seam need current halo rows; the others only the rank's own rows."""
        C, brows, W = self.buf.shape
        if not (self.r0 <= row_a < row_b <= self.r1):
            raise ValueError(f"rows [{row_a},{row_b}) outside this rank's stripe [{self.r0},{self.r1})")
        bound, mdev = self._max_args(msb_max, K)
        d = cabi.make_desc(C, self.H, W, K, self.D, bc, nl, flags.bits(relu), bound, self.buf.dtype == torch.uint16,
                           row0=row_a, row1=row_b, buf_row0=self.r0 - self.top, buf_rows=brows, w0=w0,
                           n_freq=flags.n_freq, path=path, msb_max_dev=mdev)
        cabi.check(cabi.load().lbdrn_decode(ctypes.byref(d), cabi.ptr(self.buf), cabi.ptr(flat_params_dev), cabi.ptr(tab),
                                            cabi.ptr(self.out), cabi.stream_ptr()))

    def decode(self, flat_params_dev, K, bc, nl, flags, msb_max, relu=False, w0=30.0, path=cabi.PATH_AUTO, tab=None):
        """Decode the stripe (halos must be current); returns the [C, buf_rows, W] output buffer and the slice of own rows."""
        self.decode_rows(self.r0, self.r1, flat_params_dev, K, bc, nl, flags, msb_max, relu, w0, path, tab)
        return self.out, slice(self.top, self.top + (self.r1 - self.r0))


    def decode_to_host(self, out_host, flat_params_dev, K, bc, nl, flags, msb_max, sub_rows=1024, relu=False, w0=30.0,
                       path=cabi.PATH_AUTO, tab=None):
        """Decode this rank's rows in sub-stripes and download each one on a side stream while the next computes.
        out_host: [C, rows, W] uint16 CPU tensor (pinned for asynchronous copies) receiving this rank's own rows."""
        C, brows, W = self.buf.shape
        dev = self.buf.device
        side = getattr(self, "_side", None)
        if side is None:
            side = self._side = torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)
        lib = cabi.load()
        last = None
        bound, mdev = self._max_args(msb_max, K)
        for a in range(self.r0, self.r1, sub_rows):
            b = min(self.r1, a + sub_rows)
            d = cabi.make_desc(C, self.H, W, K, self.D, bc, nl, flags.bits(relu), bound, self.buf.dtype == torch.uint16,
                               row0=a, row1=b, buf_row0=self.r0 - self.top, buf_rows=brows, w0=w0, n_freq=flags.n_freq,
                               path=path, msb_max_dev=mdev)
            cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(self.buf), cabi.ptr(flat_params_dev), cabi.ptr(tab),
                                        cabi.ptr(self.out), cabi.stream_ptr()))
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                lo, hi = self.top + (a - self.r0), self.top + (b - self.r0)
                for c in range(C):
                    out_host[c, a - self.r0:b - self.r0].copy_(self.out[c, lo:hi], non_blocking=True)
                last = torch.cuda.Event()
                last.record(side)
        if last is not None:
            last.synchronize()
        return out_host


def plan_pieces(r0, r1, has_top, has_bot, D, sub_rows):
    """Row ranges [(a, b, needs_halo)] that cover stripe [r0, r1) exactly once: interior rows first (their (2D+1)^2 windows
    stay inside the rank's own rows, so they can be decoded before the halo swap has finished), in sub-stripes of at most
    `sub_rows` rows, then the edge bands next to a neighbouring stripe.  Bands are whole tile rows (8) and at least D high."""
    edge = max(8, -(-D // 8) * 8)
    lo, hi = r0 + (edge if has_top else 0), r1 - (edge if has_bot else 0)
    if hi <= lo:                                               # thinner than its bands: everything after the halos
        return [(r0, r1, bool(has_top or has_bot))]
    pieces = [(a, min(hi, a + sub_rows), False) for a in range(lo, hi, sub_rows)]
    if has_top:
        pieces.append((r0, lo, True))
    if has_bot:
        pieces.append((hi, r1, True))
    return pieces


class StreamedStripeDecoder:
    """One rank's share of a STREAM of same-shaped scenes decoded as row stripes (SURVEY.md 8e/8f-N1 on N GPUs).

    Per scene the only exchanges are the scalar `MSB.max()` all-reduce and one D-row halo swap with each neighbouring
    stripe; neither sits in front of the kernels:

      copy stream   : H2D of this rank's stripe (pinned host memory) -> local max (device reduction) -> all-reduce(MAX)
                      -> halo swap (NCCL send/recv over NVLink), all queued while the PREVIOUS scene still computes
      compute stream: when nothing is in flight (first scene, or the previous one already finished) it waits for the max
                      only and decodes the INTERIOR rows (those whose window stays inside the rank's own rows: no halo
                      needed) while the halo swap is in flight, then the two edge bands; in steady state the halos of
                      scene i+1 arrive while scene i computes, so the stripe is decoded by ONE call (no extra launches)
      output stream : D2H of every finished sub-stripe while the next one computes

    Two buffer slots, so scene i+1's upload / collectives overlap scene i's kernels and scene i-1's download; a rank that
    runs ahead only queues work, ranks meet in the collectives of the NEXT scene while this one computes.  The copy stream
    has high priority: its small kernels (reduction, NCCL) take the first SM a finishing decode CTA frees instead of
    waiting behind the persistent grid.  The kernels read the normaliser from the all-reduced device word
    (LbdrnDesc.msb_max_dev): no host round trip anywhere.
    Output is bit-identical to the 1-GPU decode of the whole scene (tests/test_gpu_dist.py, bench.py's stripe check)."""

    def __init__(self, H, W, C, D, msb_dtype, K, bc, nl, flat_params, flags, device, group=None, slots=2, sub_rows=1024,
                 relu=False, w0=30.0, path=cabi.PATH_AUTO, tab=None):
        import lbdrn_fused
        self.group, self.K, self.bc, self.nl, self.flags = group, K, bc, nl, flags
        self.relu, self.w0, self.path, self.tab, self.sub_rows = relu, w0, path, tab, sub_rows
        self.dev = torch.device(device)
        self.params = torch.as_tensor(flat_params, dtype=torch.float32).to(self.dev).contiguous()
        _, self.s_cmp, self.s_out = lbdrn_fused._get_streams(self.dev)
        self.s_in = torch.cuda.Stream(self.dev, priority=-1)
        self.slots = [dict(sb=StripeBuffer(H, W, C, D, msb_dtype, self.dev, group),
                           mx=torch.zeros(1, dtype=torch.int32, device=self.dev),
                           local_mx=torch.zeros(1, dtype=torch.int32, device=self.dev), cmp_done=None, out_done=None,
                           ticket=None)
                      for _ in range(slots)]
        sb = self.slots[0]["sb"]
        self.r0, self.r1, self.D = sb.r0, sb.r1, D
        self.n = 0
        torch.cuda.current_stream(self.dev).synchronize()

    def _local_max(self, sl):
        """This rank's MSB maximum over its own rows -> sl["local_mx"] (current stream)."""
        import lbdrn_cabi as cabi_
        sb = sl["sb"]
        own = sb.buf[:, sb.top:sb.top + (self.r1 - self.r0)]
        if sb.buf.dtype == torch.uint8:
            sl["local_mx"].copy_(own.max().to(torch.int32).reshape(1))
        else:
            sl["local_mx"].zero_()
            for c in range(own.shape[0]):                      # own rows of every plane are contiguous
                cabi_.check(cabi_.load().lbdrn_max_shifted(cabi_.ptr(own[c]), own[c].numel(), 0, cabi_.ptr(sl["local_mx"]),
                                                           cabi_.stream_ptr()))

    def preload(self, stripe_dev):
        """Resident mode: make the scene resident in every slot ONCE -- this rank's rows, their global maximum (local
        reduction + all-reduce) and the halo rows of the neighbouring stripes (the one-off swap of SURVEY.md 8e: halos are
        static input data, like the planes themselves).  `submit()` without a source then only queues kernels."""
        for sl in self.slots:
            sb = sl["sb"]
            sb.load(stripe_dev)
            self._local_max(sl)
            sl["mx"].copy_(sl["local_mx"])
            if sb.world > 1:
                dist.all_reduce(sl["mx"], op=dist.ReduceOp.MAX, group=self.group)
            sb.exchange()
            sl["resident"] = True
        torch.cuda.current_stream(self.dev).synchronize()

    def submit(self, stripe_host=None, out_host=None):
        """Queue one scene and return at once.  stripe_host: [C, rows, W] CPU tensor (pinned) with this rank's OWN rows, or
        None when the slot already holds them (`preload`).  out_host: [C, rows, W] uint16 CPU tensor (pinned) or None.
        Returns an event that completes when the scene's reconstruction is in `out_host` (or in the slot's `out`)."""
        import lbdrn_cabi as cabi_
        sl = self.slots[self.n % len(self.slots)]
        self.n += 1
        sb, lib = sl["sb"], cabi_.load()
        rows = self.r1 - self.r0
        ev_max = ev_halo = None
        if stripe_host is None:
            if not sl.get("resident"):
                raise ValueError("submit() without a source needs preload() first")
        else:
            sl["resident"] = False
            with torch.cuda.stream(self.s_in):
                if sl["cmp_done"] is not None:
                    self.s_in.wait_event(sl["cmp_done"])       # the kernels that read this slot's planes are done
                src = sb._wire(stripe_host)
                for c in range(src.shape[0]):
                    sb.own[c].copy_(src[c], non_blocking=True)
                self._local_max(sl)
                sl["mx"].copy_(sl["local_mx"])
                if sb.world > 1:
                    dist.all_reduce(sl["mx"], op=dist.ReduceOp.MAX, group=self.group)
                ev_max = torch.cuda.Event()
                ev_max.record(self.s_in)
                sb.exchange()
                ev_halo = torch.cuda.Event()
                ev_halo.record(self.s_in)
        prev = self.slots[(self.n - 2) % len(self.slots)]["ticket"] if self.n > 1 else None
        idle = prev is None or prev.query()                    # nothing in flight: overlap the halo swap with interior rows
        if idle and ev_halo is not None:
            pieces = plan_pieces(self.r0, self.r1, bool(sb.top), bool(sb.bot), self.D, self.sub_rows)
        else:                                                  # steady state: the halos land while the previous scene computes
            pieces = [(a, min(self.r1, a + self.sub_rows), True) for a in range(self.r0, self.r1, self.sub_rows)]
        last = None
        waited_halo = False
        for i, (a, b, needs_halo) in enumerate(pieces):
            with torch.cuda.stream(self.s_cmp):
                if i == 0:
                    if ev_max is not None:
                        self.s_cmp.wait_event(ev_max)
                    if sl["out_done"] is not None:
                        self.s_cmp.wait_event(sl["out_done"])  # the previous download from this slot's output is done
                if needs_halo and not waited_halo and ev_halo is not None:
                    self.s_cmp.wait_event(ev_halo)
                    waited_halo = True
                sb.decode_rows(a, b, self.params, self.K, self.bc, self.nl, self.flags, sl["mx"], self.relu, self.w0,
                               self.path, self.tab)
                ev = torch.cuda.Event()
                ev.record(self.s_cmp)
            last = ev
            if out_host is not None:
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev)
                    for c in range(out_host.shape[0]):
                        out_host[c, a - self.r0:b - self.r0].copy_(sb.out[c, sb.top + a - self.r0:sb.top + b - self.r0],
                                                                   non_blocking=True)
                    last = torch.cuda.Event()
                    last.record(self.s_out)
        sl["cmp_done"] = ev
        sl["out_done"] = last if out_host is not None else None
        sl["ticket"] = last
        return last

    def result(self, slot_index):
        """[C, rows, W] view of a slot's reconstruction on the device (own rows)."""
        sb = self.slots[slot_index % len(self.slots)]["sb"]
        return sb.out[:, sb.top:sb.top + (self.r1 - self.r0)]


def decode_stripe(stripe_msb, H, flat_params_dev, K, D, bc, nl, flags, msb_max, group=None, relu=False, w0=30.0,
                  path=cabi.PATH_AUTO, tab=None, halo=None):
    """Decode this rank's stripe of an H-row scene.  stripe_msb: [C, rows, W] CUDA tensor of the rank's OWN rows.
    Returns the [C, rows, W] uint16 reconstruction of those rows.  `halo`: a cached (buffer, top) from
    `exchange_halos` when the same scene is decoded repeatedly."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    r0, r1 = stripe_bounds(H, world, rank)
    buf, top = exchange_halos(stripe_msb, D, group) if halo is None else halo
    C, brows, W = buf.shape
    d = cabi.make_desc(C, H, W, K, D, bc, nl, flags.bits(relu), msb_max, buf.dtype == torch.uint16, row0=r0, row1=r1,
                       buf_row0=r0 - top, buf_rows=brows, w0=w0, n_freq=flags.n_freq, path=path)
    out = torch.empty((C, brows, W), dtype=torch.uint16, device=buf.device)
    cabi.check(cabi.load().lbdrn_decode(ctypes.byref(d), cabi.ptr(buf), cabi.ptr(flat_params_dev), cabi.ptr(tab),
                                        cabi.ptr(out), cabi.stream_ptr()))
    return out.view(torch.int16)[:, top:top + (r1 - r0)].contiguous().view(torch.uint16)


def batch_slice(n, world, rank):
    """Contiguous share [a, b) of an n-pixel batch for `rank` (same rule as stripe_bounds)."""
    return stripe_bounds(n, world, rank)


class DataParallelTrainer:
    """Data-parallel encode: identical replicas, each computing the gradient of its slice of every batch; one
    all-reduce(sum) of P+1 floats per step (NCCL over NVLink/NVSwitch), then the same Adam step everywhere."""

    def __init__(self, fused_trainer, group=None):
        self.t, self.group = fused_trainer, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        n = fused_trainer.model.flat_params().numel()
        self.grad = torch.zeros(n + 1, dtype=torch.float32, device=fused_trainer.dev)

    def train_epoch(self, perm_dev, lr):
        t, lib = self.t, self.t.lib
        n, bs, C = perm_dev.numel(), t.bs, t.scene.C
        losses = []
        for s in range(math.ceil(n / bs)):
            b0, b1 = s * bs, min(n, (s + 1) * bs)
            a, b = batch_slice(b1 - b0, self.world, self.rank)
            if b > a:
                cabi.check(lib.lbdrn_train_grad(t.handle, cabi.ptr(t.scene.msb), cabi.ptr(t.scene.lsb), cabi.ptr(t.tab),
                                                ctypes.c_void_p(perm_dev.data_ptr() + 8 * (b0 + a)), b - a, b1 - b0,
                                                cabi.ptr(self.grad), cabi.stream_ptr()))
            else:
                self.grad.zero_()
            dist.all_reduce(self.grad, group=self.group)
            t.adam_t += 1
            cabi.check(lib.lbdrn_train_apply(t.handle, cabi.ptr(self.grad), t.adam_t, float(lr), cabi.stream_ptr()))
            losses.append(self.grad[-1:] / ((b1 - b0) * C))
        return torch.cat(losses)
