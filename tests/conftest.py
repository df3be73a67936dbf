import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lbdrn-msic_b200")
ORACLE = os.path.join(ROOT, "oracle")
SHIMS = os.path.join(ORACLE, "shims")
GOLD = os.path.join(ROOT, "tests", "golden")
# product modules are flat files (like the reference); the oracle and the shims for packages absent from this
# image (osgeo / fpzip / ignite) are test infrastructure and only ever imported from tests.
for p in (PKG, ORACLE, SHIMS):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ["PATH"] = os.path.join(SHIMS, "bin") + os.pathsep + os.environ.get("PATH", "")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


CODEC_CASES = ["k5d2_small", "k5d2_train", "k3d1_u16", "k4d0_abs", "k5d2_absrel0", "coords_pe", "coords_pe_col",
               "coords_only", "b8_16bit", "d3_bc256", "k1_nl3", "k9_bc128"]


def load_case(name):
    """Golden codec case minted from the unmodified reference: meta, original scene, .bin bytes, reconstruction."""
    from synth_scene import make_scene
    import hashlib
    meta = json.load(open(os.path.join(GOLD, f"{name}.json")))
    img = make_scene(meta["C"], meta["H"], meta["W"], meta["bits"], seed=meta["seed"])
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == meta["scene_sha256"], \
        "synthetic scene generator drifted from the one the fixtures were minted with"
    blob = open(os.path.join(GOLD, f"{name}.bin"), "rb").read()
    recon = np.load(os.path.join(GOLD, f"{name}_recon.npz"))["recon"]
    return meta, img, blob, recon


def case_flags(meta, cls):
    return cls(**meta.get("flags", {}))


def split_stream(blob):
    """(header tuple, [(nn_bytes, base_bytes)] per tile) using the ORACLE's header reader."""
    import lbdrn_oracle as O
    hdr = O.unpack_header(blob)
    n, sr = hdr[0], hdr[1]
    body, tiles = blob[n:], []
    for t in range(sr * sr):
        nn, base = hdr[8][t], hdr[9][t]
        tiles.append((body[:nn], body[nn:nn + base]))
        body = body[nn + base:]
    return hdr, tiles


def read_base(base_bytes):
    """Decode a base sub-stream written through the shim gdal_translate (lossless) to a CHW array."""
    import tempfile
    from osgeo import gdal
    with tempfile.NamedTemporaryFile(suffix=".jp2", delete=False) as f:
        f.write(base_bytes)
    a = gdal._load(f.name)
    os.remove(f.name)
    return a
