"""TEST INFRASTRUCTURE ONLY -- wall-clock of the UNMODIFIED reference CLI on this container's CPU next to the oracle port.

SURVEY 8d asks for the verbatim reference timed beside the GPU path.  The reference is Python and cannot travel to the
GPU box (nothing there may read /root/reference), so `bench.py` times the oracle PORT on the box's host cores and this
script, run in the build container, measures how the port relates to the verbatim reference on the same scene, same
weights, same thread count.  Result: profiles/r2_cpu_reference_verbatim.json, which bench.py quotes (`cpu_baseline`).

    python oracle/time_reference_cpu.py [side]          (needs /root/reference; about 3 minutes at side 2048)
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(HERE, "shims"), os.path.join(ROOT, "lbdrn-msic_b200")]
import fpzip                          # noqa: E402  (shim)
import lbdrn_oracle as O              # noqa: E402
import run_reference as rr            # noqa: E402
from osgeo import gdal                # noqa: E402  (shim)
from synth_scene import make_scene    # noqa: E402

_DECODE_SNIPPET = r"""
import os, sys, time, types
sys.argv = ['decode.py', '-i', ARGS['bin']]
import torch
torch.set_num_threads(ARGS['threads'])
import decode as D
D.args = types.SimpleNamespace(input=ARGS['bin'], original=None)
blob = open(ARGS['bin'], 'rb').read()
hdr = D.read_image_header(blob)
n, sr, W, H, K, bc, nl, Dd, nn, base = hdr
D.K, D.D, D.bc, D.nl = K, Dd, bc, nl
import logger
logger.create_logger(os.path.dirname(ARGS['bin']), 'decode_timing.txt')
t0 = time.perf_counter()
D.test(blob[n:], os.path.dirname(ARGS['bin']), 'timed', nn[0], base[0])
RESULT = dict(seconds=time.perf_counter() - t0, H=H, W=W)
"""


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    threads = os.cpu_count() or 1
    work = tempfile.mkdtemp(prefix="lbdrn_reftime_")
    img = make_scene(4, side, side, 12, seed=19920517)
    tif = os.path.join(work, "scene.tif")
    gdal._store(tif, img)
    t0 = time.perf_counter()
    d, binp, log = rr.encode(tif, os.path.join(work, "out"), K=5, D=2, bc=64, nl=2, bs=65536, e=1)
    enc_s = time.perf_counter() - t0
    steps = -(-side * side // 65536)
    # (1) the whole decode.py process, as a user runs it (interpreter start, imports, JP2 -> TIFF, features, model, write)
    t0 = time.perf_counter()
    rr.decode(binp)
    cli_s = time.perf_counter() - t0
    # (2) the reference's test() alone (decode.py:56-141: base read, features, model, reconstruction, TIFF write)
    try:
        r = rr.call_snippet(_DECODE_SNIPPET, dict(bin=binp, threads=threads))
        test_s = r["seconds"]
    except Exception as e:                                        # noqa: BLE001
        print("snippet failed:", str(e)[-600:])
        test_s = None
    # (3) the oracle port on the same base layer and weights (what bench.py times on the GPU box)
    torch.set_num_threads(threads)
    blob = open(binp, "rb").read()
    hdr = O.unpack_header(blob)
    nn = blob[hdr[0]:hdr[0] + hdr[8][0]]
    flat = np.asarray(fpzip.decompress(nn)[0][0][0], np.float32)
    msb, _ = O.split_msb_lsb(img, 5)
    t0 = time.perf_counter()
    O.decode_image(msb, O.unflatten_params(flat, 100, 64, 4, 2), 5, 2)
    port_s = time.perf_counter() - t0
    out = dict(side=side, pixels=side * side, threads=threads, host="build container (no GPU)",
               verbatim_decode_cli_s=cli_s, verbatim_decode_test_fn_s=test_s, port_decode_s=port_s,
               verbatim_Mpix_s=side * side / (test_s or cli_s) / 1e6, port_Mpix_s=side * side / port_s / 1e6,
               port_over_verbatim=(test_s or cli_s) / port_s,
               verbatim_encode_1epoch_bs65536_s=enc_s, encode_steps=steps,
               note="verbatim = unmodified /root/reference decode.py under oracle/shims (osgeo/fpzip/ignite stand-ins); "
                    "test_fn = decode.py:test() alone, cli = the whole `python decode.py -i x.bin` process")
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_cpu_reference_verbatim.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
