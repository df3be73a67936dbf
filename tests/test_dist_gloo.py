"""CPU, world_size=2 over gloo: host-side logic of the multi-GPU paths (stripe partitioning, halo exchange, global max,
batch slicing, gradient all-reduce).  The kernels themselves need a GPU; what is checked here is that every rank
assembles exactly the rows / batch shares the single-GPU path would see."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG


def _worker(rank, world, port, fn_name, q):
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, globals()[fn_name](rank, world)))
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def _halo_case(rank, world):
    import lbdrn_dist as LD
    H, W, C, D = 37, 11, 3, 2
    full = torch.arange(C * H * W, dtype=torch.int32).reshape(C, H, W).to(torch.int16).view(torch.uint16)
    r0, r1 = LD.stripe_bounds(H, world, rank)
    own = full.view(torch.int16)[:, r0:r1].contiguous().view(torch.uint16)
    buf, top = LD.exchange_halos(own, D)
    b0, b1 = max(0, r0 - D), min(H, r1 + D)
    ok = torch.equal(buf.view(torch.int16), full.view(torch.int16)[:, b0:b1]) and top == r0 - b0
    u8 = (full.view(torch.int16) % 200).to(torch.uint8)
    buf8, top8 = LD.exchange_halos(u8[:, r0:r1].contiguous(), D)
    ok8 = torch.equal(buf8, u8[:, b0:b1]) and top8 == top
    gmax = LD.global_max(int(full.view(torch.int16)[:, r0:r1].max()))
    return bool(ok), bool(ok8), gmax, int(full.view(torch.int16).max())


def _grad_case(rank, world):
    import lbdrn_dist as LD
    n = 1000
    torch.manual_seed(0)
    per_pixel = torch.randn(n, 7)                      # stand-in for per-pixel gradient contributions
    a, b = LD.batch_slice(n, world, rank)
    g = per_pixel[a:b].sum(0)
    dist.all_reduce(g)
    return (a, b), torch.allclose(g, per_pixel.sum(0), atol=1e-4)


def test_stripe_bounds_cover_the_image():
    sys.path.insert(0, PKG)
    import lbdrn_dist as LD
    for H in (1, 7, 8, 37, 8192, 16384):
        for world in (1, 2, 3, 4, 8):
            if world > H:
                continue
            b = [LD.stripe_bounds(H, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == H
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def test_halo_exchange_and_global_max_world2():
    out = _run("_halo_case")
    for rank, (ok, ok8, gmax, ref) in out.items():
        assert ok and ok8, rank
        assert gmax == ref


def test_batch_slices_and_gradient_allreduce_world2():
    out = _run("_grad_case")
    slices = sorted(v[0] for v in out.values())
    assert slices[0][0] == 0 and slices[-1][1] == 1000 and slices[0][1] == slices[1][0]
    assert all(v[1] for v in out.values())


def _sched_case(rank, world):
    import lbdrn_sched as S
    sizes = {f"s{i}.tif": (4, 1000 + 37 * i, 900) for i in range(5)}
    jobs = S.expand_jobs(sizes, [2, 5, 8])
    mine = S.plan(jobs, world)[rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, [tuple(j) for j in mine])
    return sorted(tuple(j) for j in jobs) == sorted(j for g in gathered for j in g), len(mine)


def test_scheduler_ranks_partition_the_job_list_world2():
    out = _run("_sched_case")
    assert all(v[0] for v in out.values())
    assert sum(v[1] for v in out.values()) == 15 and min(v[1] for v in out.values()) >= 6


def test_stripe_piece_plan_covers_every_row_once_and_orders_interior_first():
    """lbdrn_dist.plan_pieces (used by StreamedStripeDecoder): exact cover of the stripe, interior rows (no halo needed)
    before the edge bands, bands at least D high and whole tile rows, thin stripes handled."""
    import lbdrn_dist as LD
    for (r0, r1, top, bot, D, sub) in [(0, 8192, False, True, 2, 1024), (8192, 16384, True, True, 2, 1024),
                                       (100, 117, True, True, 2, 1024), (0, 13, False, False, 2, 4), (7, 4103, True, False, 3, 1000),
                                       (2048, 4096, True, True, 9, 512)]:
        pieces = LD.plan_pieces(r0, r1, top, bot, D, sub)
        rows = sorted(r for a, b, _ in pieces for r in range(a, b))
        assert rows == list(range(r0, r1)), (r0, r1, pieces)
        flags = [h for _, _, h in pieces]
        assert flags == sorted(flags), pieces                               # interior (False) first
        for a, b, needs in pieces:
            if not needs:                                                   # windows stay inside the rank's own rows
                assert (not top or a - D >= r0) and (not bot or b - 1 + D < r1), (a, b, pieces)
                assert b - a <= sub
