"""TEST INFRASTRUCTURE ONLY -- execute the UNMODIFIED reference CLI (`/root/reference/encode.py`, `decode.py`)
on the CPU of this container, as the ground-truth oracle.

The reference cannot travel to the GPU box, so this module is used only (a) by `oracle/make_golden.py` to mint
the fixtures under `tests/golden/`, and (b) by `-m "not gpu"` tests that re-validate the restatement in
`oracle/lbdrn_oracle.py` whenever `/root/reference` is present.  Nothing on the product path imports it.

How the reference is run verbatim:
  * `oracle/shims` is put on `sys.path` ahead of site-packages for the three packages missing from this
    image (`osgeo`, `fpzip`, `ignite`) and `oracle/shims/bin` on `PATH` for `gdal_translate`;
  * an optional directory holding an alternative `constants.py` is put FIRST on `sys.path` (the reference
    switches feature sets by editing that file: README.md "Modify constants.py");
  * `torch.utils.data.DataLoader` is wrapped to force `num_workers=0` (batch contents do not depend on the
    worker count: the sampler runs in the main process) and `SummaryWriter` is replaced by a recorder that
    dumps every scalar to `<log_dir>/scalars.json` (full-precision per-iteration losses / per-epoch MSE);
  * the script is then executed with `runpy.run_path(..., run_name="__main__")` in a fresh interpreter.
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
REFERENCE = os.environ.get("LBDRN_REFERENCE_DIR", "/root/reference")

_BOOT = r"""
import json, os, runpy, sys
cfg = json.loads(sys.argv[1])
sys.argv = [cfg["script"]] + cfg["argv"]
sys.path[:0] = [p for p in (cfg["constants_dir"], cfg["shims"], cfg["reference"]) if p]
import torch, torch.utils.data as tud
torch.set_num_threads(cfg["threads"])
_DL = tud.DataLoader
class DataLoader(_DL):
    def __init__(self, *a, **k):
        k["num_workers"] = 0; k["pin_memory"] = False
        super().__init__(*a, **k)
tud.DataLoader = DataLoader
import torch.utils.tensorboard as tb
class SummaryWriter:
    def __init__(self, log_dir=None, **k):
        self.log_dir, self.rows = log_dir, []
    def add_scalar(self, tag, value, step=None):
        self.rows.append([tag, float(value), int(step)])
    def close(self):
        with open(os.path.join(self.log_dir, "scalars.json"), "w") as f: json.dump(self.rows, f)
tb.SummaryWriter = SummaryWriter
runpy.run_path(os.path.join(cfg["reference"], cfg["script"]), run_name="__main__")
"""


def available():
    return os.path.isfile(os.path.join(REFERENCE, "encode.py"))


def write_constants(dirname, use_coordinates=False, embedding=False, use_colors=True, relative=True,
                    sigma=1.4, n_freq=12):
    """Write a `constants.py` carrying the reference's six feature flags (reference constants.py:3-14)."""
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "constants.py"), "w") as f:
        f.write(f"USE_COORDINATES = {bool(use_coordinates)}\nEMBEDDING = {bool(embedding)}\n"
                f"SIGMA = {sigma!r}\nN_FREQ = {int(n_freq)}\nUSE_COLORS = {bool(use_colors)}\n"
                f"RELATIVE = {bool(relative)}\n")
    return dirname


def _run(script, argv, cwd, flags=None, threads=8):
    constants_dir = None
    if flags:
        constants_dir = write_constants(tempfile.mkdtemp(prefix="lbdrn_flags_"), **flags)
    cfg = dict(script=script, argv=[str(a) for a in argv], shims=SHIMS, reference=REFERENCE,
               constants_dir=constants_dir, threads=threads)
    env = dict(os.environ)
    env["PATH"] = os.path.join(SHIMS, "bin") + os.pathsep + env.get("PATH", "")
    env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, "-c", _BOOT, json.dumps(cfg)], cwd=cwd, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference {script} failed:\n{r.stdout[-4000:]}")
    return r.stdout


_SNIPPET_BOOT = r"""
import json, sys
cfg = json.loads(sys.argv[1])
sys.argv = ["snippet"]
sys.path[:0] = [p for p in (cfg["constants_dir"], cfg["shims"], cfg["reference"]) if p]
ns = {"ARGS": cfg["args"]}
exec(cfg["code"], ns)
print("@@RESULT@@" + json.dumps(ns["RESULT"]))
"""


def call_snippet(code, args=None, flags=None):
    """Run `code` in a fresh interpreter that can `import` the reference's modules (shims on the path);
    the snippet reads its inputs from ARGS and leaves a JSON-serialisable RESULT."""
    constants_dir = None
    if flags:
        constants_dir = write_constants(tempfile.mkdtemp(prefix="lbdrn_flags_"), **flags)
    cfg = dict(code=code, args=args, shims=SHIMS, reference=REFERENCE, constants_dir=constants_dir)
    env = dict(os.environ)
    env["PATH"] = os.path.join(SHIMS, "bin") + os.pathsep + env.get("PATH", "")
    env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, "-c", _SNIPPET_BOOT, json.dumps(cfg)], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference snippet failed:\n{r.stderr[-4000:]}")
    return json.loads(r.stdout.split("@@RESULT@@")[-1])


def outdir_name(out_root, name, sr, K, bc, nl, D, prec, lr, bs, e):
    """Output-directory naming rule of the reference CLI (reference encode.py:210-213)."""
    return f"{out_root}/{name}_r{sr}_K{K}_bc{bc}_nl{nl}_D{D}_prec{prec}_lr{lr}_bs{bs}_e{e}"


def encode(tif_path, out_root, K=5, D=2, bc=64, nl=2, lr=1e-3, bs=8192, e=10, sr=1, prec=16,
           flags=None, cwd=None):
    argv = ["-K", K, "-i", tif_path, "-D", D, "-bc", bc, "-nl", nl, "-lr", lr, "-bs", bs, "-e", e,
            "-sr", sr, "-prec", prec, "-o", out_root]
    log = _run("encode.py", argv, cwd or os.path.dirname(out_root) or ".", flags)
    name = os.path.splitext(os.path.basename(tif_path))[0]
    d = outdir_name(out_root, name, sr, K, bc, nl, D, prec, lr, bs, e)
    return d, f"{d}/{name}.bin", log


def decode(bin_path, org_path=None, flags=None, keep_recon=True):
    """Run reference decode.py; because it deletes the reconstruction when `-org` is given
    (decode.py:223-224) we decode WITHOUT `-org` first to keep `<name>_recon.tif`."""
    d = os.path.dirname(bin_path)
    log = _run("decode.py", ["-i", bin_path], d, flags)
    recon = bin_path[:-4] + "_recon.tif"
    log2 = ""
    if org_path is not None:
        # second pass for the quality read-out lines (MSE / PSNR / bpsp)
        keep = recon + ".keep"
        os.replace(recon, keep)
        os.remove(f"{d}/decode.txt")
        log2 = _run("decode.py", ["-i", bin_path, "-org", org_path], d, flags)
        os.replace(keep, recon)
    return recon, log + log2
