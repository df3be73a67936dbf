"""Identity fraction of every decode path against the oracle, per K, on a scene large enough to resolve 1e-6
(TEST TOOL: imports the oracle).  usage: sine_parity.py [H] [W] [K,K,...]
Weights: the reference-trained parameters of the k5d2_train fixture (fpzip -prec 16 values)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lbdrn-msic_b200", "oracle", os.path.join("oracle", "shims"), "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import fpzip  # shim
import lbdrn_fused as F
import lbdrn_oracle as O
from conftest import load_case, split_stream
from synth_scene import make_scene

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
Ks = [int(k) for k in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 3, 5, 6, 7, 8]
for case in ("k5d2_train", "k5d2_small"):
    meta, _, blob, _ = load_case(case)
    flat = np.asarray(fpzip.decompress(split_stream(blob)[1][0][0])[0][0][0], dtype=np.float32)
    p = O.unflatten_params(flat, 100, 64, 4, 2)
    img = make_scene(4, H, W, 12, seed=123)
    for K in Ks:
        msb, _ = O.split_msb_lsb(img, K)
        t0 = time.time()
        ref = O.decode_image(msb, p, K, 2)
        row = [f"{case} K={K} (oracle {time.time() - t0:.1f}s)"]
        for path in ("precise", "tensor", "tensor_fastsin", "tensor_fastsin2"):
            out = F.decode_image(msb, flat, K, 2, 64, 2, flags=F.Flags(), path=path)
            diff = np.abs(out.astype(np.int64) - ref.astype(np.int64))
            row.append(f"{path}: {1e6 * (diff != 0).mean():8.2f} ppm max {diff.max()}")
        print(" | ".join(row), flush=True)
