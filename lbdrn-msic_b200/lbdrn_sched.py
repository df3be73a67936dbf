"""Scene / K batch scheduler -- row N3 of SURVEY.md 8f.

The reference encodes a data set with a serial shell loop over scenes and K values (run.sh:33-41), and treats the
`-sr` tiles of one scene as independent networks (encode.py:231-238).  Those (scene, K) jobs share nothing, so they are
the unit that gives encode its multi-GPU throughput ("replicas only": no collective on the data path; BASELINE.json
north_star "independent scenes per GPU").  This module

  * expands scenes x K into jobs with a cost estimate (sub-pixels x epochs),
  * assigns them to ranks with the longest-processing-time-first rule, keeping all K values of one scene on one rank
    where that does not unbalance the plan, so a K sweep uploads its scene ONCE (`LBDRNdataset.PRELOADED`: the MSB/LSB
    split for every K runs on the device from the resident uint16 image),
  * runs a rank's share through the unchanged encoder entry point (`encode.main`), so output directories, `.bin`
    layout and log lines are exactly those of the CLI.

    python lbdrn_sched.py -i a.tif b.tif -K 1 2 3 4 5 [encode.py flags ...]            # one GPU
    torchrun --nproc-per-node 8 lbdrn_sched.py -i scenes/*.tif -K 3 5 7 ...             # one rank per GPU
"""
import argparse
import os
import sys
from collections import namedtuple

Job = namedtuple("Job", "path K cost")


def expand_jobs(scene_sizes, Ks, epochs=10):
    """scene_sizes: {path: (bands, height, width)} -> one Job per (scene, K); cost = sub-pixels x epochs (encode time is
    proportional to the number of optimiser steps + evaluation passes, both linear in the pixel count)."""
    return [Job(p, int(K), float(c * h * w * epochs)) for p, (c, h, w) in sorted(scene_sizes.items()) for K in Ks]


def plan(jobs, world):
    """Deterministic longest-processing-time-first assignment of whole scenes (all their K values) to ranks; a scene
    whose K sweep alone exceeds the ideal per-rank load is split K by K.  Returns `world` job lists (same on every rank:
    no communication needed to agree on the plan)."""
    world = max(1, int(world))
    total = sum(j.cost for j in jobs)
    ideal = total / world if world else total
    by_scene = {}
    for j in jobs:
        by_scene.setdefault(j.path, []).append(j)
    units = []                                               # schedulable units: whole sweeps, or single jobs
    for path, js in sorted(by_scene.items()):
        sweep = sum(j.cost for j in js)
        if world > 1 and sweep > 1.25 * ideal and len(js) > 1:
            units.extend([[j] for j in js])
        else:
            units.append(js)
    units.sort(key=lambda u: (-sum(j.cost for j in u), u[0].path, u[0].K))
    load, out = [0.0] * world, [[] for _ in range(world)]
    for u in units:
        r = min(range(world), key=lambda i: (load[i], i))
        out[r].extend(u)
        load[r] += sum(j.cost for j in u)
    for r in range(world):
        out[r].sort(key=lambda j: (j.path, j.K))             # group a rank's jobs by scene: one upload per scene
    return out


def raster_size(path):
    from osgeo import gdal
    ds = gdal.Open(path)
    size = (ds.RasterCount, ds.RasterYSize, ds.RasterXSize)
    ds = None
    return size


def run_rank(jobs, encode_argv, encode_main=None, preload=True, log=print):
    """Encode this rank's jobs.  `encode_argv`: the encode.py flags shared by every job (without -i / -K).  Returns
    [(path, K, status)], status in {"done", "exists"}."""
    import LBDRNdataset
    if encode_main is None:
        import encode
        encode_main = encode.main
    done, resident = [], None
    for job in jobs:
        if preload and resident != job.path:
            LBDRNdataset.PRELOADED.clear()                   # at most one scene resident per rank
            LBDRNdataset.preload(job.path)
            resident = job.path
        log(f"[lbdrn_sched] encode {job.path} K={job.K}")
        try:
            encode_main(["-i", job.path, "-K", str(job.K)] + list(encode_argv))
            done.append((job.path, job.K, "done"))
        except SystemExit:                                   # encode.py exits when the bitstream already exists
            done.append((job.path, job.K, "exists"))
    LBDRNdataset.PRELOADED.clear()
    return done


def main(argv=None):
    ap = argparse.ArgumentParser(description="LBDRN scene x K batch encoder (one rank per GPU)")
    ap.add_argument("-i", "--paths", nargs="+", required=True)
    ap.add_argument("-K", "--Ks", nargs="+", type=int, default=[5])
    ap.add_argument("-e", "--epochs", type=int, default=10)
    args, rest = ap.parse_known_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("lbdrn_sched.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    jobs = expand_jobs({p: raster_size(p) for p in args.paths}, args.Ks, args.epochs)
    mine = plan(jobs, world)[rank]
    res = run_rank(mine, ["-e", str(args.epochs)] + rest)
    print(f"[lbdrn_sched] rank {rank}/{world}: {len(res)} of {len(jobs)} jobs", res)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
