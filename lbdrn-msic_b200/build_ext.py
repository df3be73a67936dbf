"""Build liblbdrn_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python lbdrn-msic_b200/build_ext.py [--force]

Translation units are compiled in parallel (the fp32 inference kernel alone has 48 instantiations) and linked into
one shared object.  The .so is git-ignored but travels to the GPU box with the repo snapshot; nvcc cross-compiles
without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liblbdrn_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                 "-Xptxas", "-v"]
# (object name, source, extra flags)
UNITS = [
    ("api", "lbdrn_api.cu", []),
    ("infer_decode", "lbdrn_infer_fp32.cu", ["-DLBDRN_INFER_MODE=0"]),
    ("infer_predict", "lbdrn_infer_fp32.cu", ["-DLBDRN_INFER_MODE=1"]),
    ("infer_sse", "lbdrn_infer_fp32.cu", ["-DLBDRN_INFER_MODE=2"]),
    ("train", "lbdrn_train_fp32.cu", []),
    ("tc", "lbdrn_tc.cu", []),
    ("tcw", "lbdrn_tcw.cu", []),
    ("fpz", "lbdrn_fpz.cpp", []),          # host-only: the nn sub-stream codec
    ("hostperm", "lbdrn_hostperm.cpp", []),   # host-only: the reference sampler's permutation
]


def _newest_source_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lbdrn.h"),
                                                                   os.path.abspath(__file__)]
    return max(os.path.getmtime(p) for p in paths)


def _compile(unit):
    name, src, extra = unit
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj = os.path.join(OBJ, name + ".o")
    cmd = [nvcc] + CFLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return name, obj, " ".join(cmd), r.returncode, r.stdout


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    log, objs, failed = [], [], False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 2)) as ex:
        for name, obj, cmd, rc, out in ex.map(_compile, UNITS):
            log.append(f"### {name}\n{cmd}\n{out}")
            objs.append(obj)
            if rc != 0:
                failed = True
                sys.stderr.write(f"[build_ext] {name} failed:\n{out[-6000:]}\n")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not failed:
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs      # cudart is linked statically; no libcuda dependency
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append("### link\n" + " ".join(cmd) + "\n" + r.stdout)
        failed = r.returncode != 0
        if failed:
            sys.stderr.write(r.stdout[-4000:])
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed building liblbdrn_b200.so (see lbdrn-msic_b200/build.log)")
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
