"""TEST INFRASTRUCTURE ONLY -- mint the fixtures under tests/golden/ by executing the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Everything written is small; scenes are regenerated from their seed (`synth_scene.make_scene`) and only
their SHA-256 is stored.  The GPU box has no /root/reference: tests there read these files only.
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(HERE, "shims"), os.path.join(ROOT, "lbdrn-msic_b200")]
import run_reference as rr            # noqa: E402
from osgeo import gdal                # noqa: E402  (shim)
from synth_scene import make_scene    # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> scene + codec settings + feature flags.  `bits`/shape chosen so every code path of the hot path has a pin.
CASES = {
    "k5d2_small":   dict(C=4, H=96,  W=80,  bits=12, seed=1, K=5, D=2, bc=64,  nl=2, bs=512,  e=3,  sr=1),
    "k5d2_train":   dict(C=4, H=192, W=160, bits=12, seed=2, K=5, D=2, bc=64,  nl=2, bs=1024, e=10, sr=1),
    "k3d1_u16":     dict(C=3, H=64,  W=72,  bits=14, seed=3, K=3, D=1, bc=32,  nl=1, bs=512,  e=2,  sr=1),
    "k4d0_abs":     dict(C=4, H=64,  W=64,  bits=12, seed=4, K=4, D=0, bc=64,  nl=2, bs=512,  e=2,  sr=1),
    "k5d2_absrel0": dict(C=4, H=64,  W=64,  bits=12, seed=5, K=5, D=2, bc=64,  nl=2, bs=512,  e=2,  sr=1,
                         flags=dict(relative=False)),
    "coords_pe":    dict(C=4, H=64,  W=64,  bits=12, seed=6, K=5, D=2, bc=64,  nl=2, bs=512,  e=2,  sr=1,
                         flags=dict(use_coordinates=True, embedding=True, use_colors=False)),
    "coords_pe_col": dict(C=4, H=64, W=64,  bits=12, seed=7, K=5, D=2, bc=64,  nl=2, bs=512,  e=2,  sr=1,
                          flags=dict(use_coordinates=True, embedding=True, use_colors=True)),
    "coords_only":  dict(C=4, H=48,  W=56,  bits=12, seed=8, K=5, D=2, bc=64,  nl=2, bs=512,  e=2,  sr=1,
                         flags=dict(use_coordinates=True, embedding=False, use_colors=True)),
    "sr2_tiles":    dict(C=4, H=97,  W=90,  bits=12, seed=9, K=5, D=2, bc=64,  nl=2, bs=512,  e=2,  sr=2),
    "b8_16bit":     dict(C=8, H=64,  W=64,  bits=16, seed=10, K=8, D=2, bc=64, nl=2, bs=512,  e=2,  sr=1),
    "d3_bc256":     dict(C=4, H=64,  W=64,  bits=12, seed=11, K=5, D=3, bc=256, nl=2, bs=512, e=2,  sr=1),
    "k1_nl3":       dict(C=4, H=64,  W=64,  bits=12, seed=12, K=1, D=2, bc=64,  nl=3, bs=512,  e=2,  sr=1),
    "k9_bc128":     dict(C=4, H=64,  W=64,  bits=12, seed=13, K=9, D=2, bc=128, nl=1, bs=512,  e=2,  sr=1),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def codec_case(name, cfg, work):
    img = make_scene(cfg["C"], cfg["H"], cfg["W"], cfg["bits"], seed=cfg["seed"])
    tif = os.path.join(work, f"{name}.tif")
    gdal._store(tif, img)
    flags = cfg.get("flags")
    d, binp, enc_log = rr.encode(tif, os.path.join(work, "out"), K=cfg["K"], D=cfg["D"], bc=cfg["bc"],
                                 nl=cfg["nl"], bs=cfg["bs"], e=cfg["e"], sr=cfg["sr"], flags=flags)
    recon, dec_log = rr.decode(binp, tif, flags=flags)
    rec = gdal._load(recon)
    scal = []
    sj = os.path.join(d, "scalars.json")
    if os.path.exists(sj) and cfg["sr"] == 1:
        scal = json.load(open(sj))
    grab = lambda key: [float(l.split(key)[1].split()[0].rstrip(",")) for l in dec_log.splitlines() if key in l]
    meta = dict(cfg, scene_sha256=sha(img), recon_sha256=sha(rec), bin_bytes=os.path.getsize(binp),
                mse=grab("MSE: ")[-1], psnr=grab("PSNR: ")[-1], bpsp=grab("bpsp=")[-1],
                losses=[v for t, v, s in scal if t.startswith("train/loss")],
                val_mse=[v for t, v, s in scal if t.startswith("val/MSE")],
                best_epoch=[int(l.split("best epoch:")[1]) for l in enc_log.splitlines() if "best epoch:" in l])
    shutil.copy(binp, os.path.join(GOLD, f"{name}.bin"))
    np.savez_compressed(os.path.join(GOLD, f"{name}_recon.npz"), recon=rec)
    json.dump(meta, open(os.path.join(GOLD, f"{name}.json"), "w"), indent=1)
    print(f"{name}: {meta['bin_bytes']} B, PSNR {meta['psnr']:.4f}, best {meta['best_epoch']}")


_HEADER_SNIPPET = r"""
import os, tempfile, encode, decode
out = []
for kw in ARGS:
    p = tempfile.mktemp()
    encode.write_image_header(p, **kw)
    b = open(p, 'rb').read(); os.remove(p)
    out.append(dict(args=kw, hex=b.hex(), parsed=list(decode.read_image_header(b))))
RESULT = out
"""

_INIT_SNIPPET = r"""
import hashlib, torch, numpy as np
from LBDRNmodel import LBDRNModel
out = []
for kw in ARGS:
    torch.manual_seed(19920517)
    m = LBDRNModel(**kw)
    flat = np.concatenate([v.numpy().reshape(-1) for v in m.state_dict().values()]).astype(np.float32)
    d1 = int(torch.empty((), dtype=torch.int64).random_()); d2 = int(torch.empty((), dtype=torch.int64).random_())
    out.append(dict(args=kw, n=int(flat.size), sha256=hashlib.sha256(flat.tobytes()).hexdigest(),
                    keys=list(m.state_dict().keys()), head=[float(v) for v in flat[:4]], next_draws=[d1, d2]))
RESULT = out
"""

_FEATURE_SNIPPET = r"""
import os, tempfile, numpy as np
from osgeo import gdal
import LBDRNdataset
tif = ARGS['tif']; out = tempfile.mktemp()
f, l = LBDRNdataset.process(tif, ARGS['K'], ARGS['D'], out)
base = gdal._load(out); os.remove(out)
np.savez_compressed(ARGS['npz'], features=f, labels=l, base=base)
RESULT = dict(shape=list(f.shape))
"""

_FORWARD_SNIPPET = r"""
import numpy as np, torch
from LBDRNmodel import LBDRNModel
torch.manual_seed(7)
m = LBDRNModel(dim_in=ARGS['dim_in'], dim_hidden=ARGS['bc'], dim_out=ARGS['C'], num_layers=ARGS['nl'])
x = (torch.rand(ARGS['n'], ARGS['dim_in']) - 0.5) * 0.2
with torch.no_grad(): y = m(x)
flat = np.concatenate([v.numpy().reshape(-1) for v in m.state_dict().values()]).astype(np.float32)
np.savez_compressed(ARGS['npz'], x=x.numpy(), y=y.numpy(), params=flat)
RESULT = dict(ok=True)
"""


# the commented alternative of encode.py:75 / decode.py:108: hidden activation ReLU.  Forward outputs, the MSE loss against
# random targets and its gradient w.r.t. every parameter, from the reference's own module and torch autograd.
_FORWARD_RELU_SNIPPET = r"""
import numpy as np, torch
from LBDRNmodel import LBDRNModel
from LBDRNloss import LBDRNLoss
torch.manual_seed(11)
m = LBDRNModel(dim_in=ARGS['dim_in'], dim_hidden=ARGS['bc'], dim_out=ARGS['C'], num_layers=ARGS['nl'],
               activation=torch.nn.ReLU())
with torch.no_grad():
    for name, p in m.named_parameters():
        if 'weight' in name: p.mul_(ARGS['gain'])          # spread the pre-activations so that ReLU clips about half
x = (torch.rand(ARGS['n'], ARGS['dim_in']) - 0.5) * 0.2
t = torch.rand(ARGS['n'], ARGS['C'])
y = m(x)
loss = LBDRNLoss()(y, t)
loss.backward()
flat = np.concatenate([v.numpy().reshape(-1) for v in m.state_dict().values()]).astype(np.float32)
grad = np.concatenate([p.grad.numpy().reshape(-1) for p in m.parameters()]).astype(np.float32)
np.savez_compressed(ARGS['npz'], x=x.numpy(), t=t.numpy(), y=y.detach().numpy(), params=flat, loss=np.float32(loss.item()), grad=grad)
RESULT = dict(ok=True)
"""


def main():
    if not rr.available():
        sys.exit("reference not present; fixtures can only be minted in the build container")
    os.makedirs(GOLD, exist_ok=True)
    work = tempfile.mkdtemp(prefix="lbdrn_golden_")
    only = set(sys.argv[1:])

    if not only or "kat" in only:
        hdr = rr.call_snippet(_HEADER_SNIPPET, [
            dict(split_ratio=1, width=2048, height=2048, K=5, bc=64, nl=2, D=2,
                 nn_bytes_list=[19300], base_bytes_list=[1234567]),
            dict(split_ratio=3, width=7340, height=7815, K=11, bc=256, nl=3, D=3,
                 nn_bytes_list=list(range(1000, 10000, 1000)), base_bytes_list=list(range(70000, 700000, 70000))),
            dict(split_ratio=2, width=65535, height=1, K=15, bc=32768, nl=15, D=15,
                 nn_bytes_list=[0, 1, 2 ** 24 - 1, 5], base_bytes_list=[0, 2 ** 32 - 1, 7, 8]),
        ])
        json.dump(hdr, open(os.path.join(GOLD, "header_kat.json"), "w"), indent=1)
        init = rr.call_snippet(_INIT_SNIPPET, [
            dict(dim_in=100, dim_hidden=64, dim_out=4, num_layers=2),
            dict(dim_in=196, dim_hidden=256, dim_out=4, num_layers=2),
            dict(dim_in=200, dim_hidden=64, dim_out=8, num_layers=2),
            dict(dim_in=50, dim_hidden=64, dim_out=4, num_layers=3),
            dict(dim_in=4, dim_hidden=128, dim_out=4, num_layers=1),
        ])
        json.dump(init, open(os.path.join(GOLD, "init_kat.json"), "w"), indent=1)
        # feature layout pins: a 4x9x11 12-bit image under every flag set the README lists
        img = make_scene(4, 9, 11, 12, seed=99)
        tif = os.path.join(work, "feat.tif")
        gdal._store(tif, img)
        combos = {
            "rel_d2": (dict(), 5, 2), "rel_d1": (dict(), 5, 1), "abs_d0": (dict(), 5, 0),
            "abs_d2": (dict(relative=False), 5, 2),
            "coords_rel_d2": (dict(use_coordinates=True), 5, 2),
            "coords": (dict(use_coordinates=True, use_colors=False), 5, 2),
            "coords_pe": (dict(use_coordinates=True, embedding=True, use_colors=False), 5, 2),
            "coords_pe_rel_d2": (dict(use_coordinates=True, embedding=True), 3, 2),
        }
        index = {}
        for nm, (fl, K, D) in combos.items():
            npz = os.path.join(GOLD, f"features_{nm}.npz")
            r = rr.call_snippet(_FEATURE_SNIPPET, dict(tif=tif, K=K, D=D, npz=npz), flags=fl or None)
            index[nm] = dict(flags=fl, K=K, D=D, shape=r["shape"], scene=dict(C=4, H=9, W=11, bits=12, seed=99))
        json.dump(index, open(os.path.join(GOLD, "features_index.json"), "w"), indent=1)
        rr.call_snippet(_FORWARD_SNIPPET, dict(dim_in=100, bc=64, C=4, nl=2, n=257,
                                               npz=os.path.join(GOLD, "forward_d100_bc64.npz")))
        print("KATs written")
    if not only or "kat" in only or "kat_relu" in only:
        rr.call_snippet(_FORWARD_RELU_SNIPPET, dict(dim_in=100, bc=64, C=4, nl=2, n=257, gain=8.0,
                                                    npz=os.path.join(GOLD, "forward_relu_d100_bc64.npz")))
        print("ReLU KAT written")

    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        codec_case(name, cfg, work)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
