"""Builds oak_b200/liboak_b200.so (CUDA, sm_100a only) in-tree with nvcc.

    python -m oak_b200.build [--force] [-v]

nvcc cross-compiles without a GPU.  Objects go to oak_b200/csrc/build/ (git-ignored); the shared
library is git-ignored too but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "liboak_b200.so")
SOURCES = ["api.cu", "obsgrid.cu", "gram.cu", "gram_mma.cu", "eig_simple.cu", "eig_fast.cu", "eig_tridiag.cu", "apply.cu", "apply_mma.cu", "global.cu", "ensemble.cu", "hgen.cu",
           "microbench.cu"]
HEADERS = ["common.cuh", "eig_common.cuh", "tridiag_math.cuh", "tridiag_warp.cuh", "../../include/oak_b200.h", "../../include/oak_b200_math.h"]
# per-file extra flags (none at present: eig_fast.cu spells its approximate fp32 operations in inline PTX)
EXTRA_FLAGS = {}
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"]


def nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: oak_b200 needs the CUDA toolkit to build (no CPU fallback)")


def _stamp():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update((" ".join(NVCC_FLAGS) + repr(sorted(EXTRA_FLAGS.items()))).encode())
    return h.hexdigest()


def build_variant(out, defines):
    """A/B experiments: the same sources with -D switches into another .so (not the product build)."""
    cc = nvcc()
    vdir = os.path.join(BUILD, "variant_" + os.path.basename(out))
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(vdir, src.replace(".cu", ".o"))
        cmd = [cc] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(p.stderr)
        if src == "eig_fast.cu":
            print("\n".join(l for l in p.stderr.splitlines() if "spill" in l or "registers" in l))
        objs.append(obj)
    subprocess.check_call([cc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return out


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    stamp_file = os.path.join(BUILD, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    cc = nvcc()

    def one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [cc] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD, src.replace(".cu", ".ptxas.log"))
        with open(log, "w") as fh:
            fh.write(p.stdout + p.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{p.stdout}\n{p.stderr}")
        if verbose:
            print(p.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(one, SOURCES))
    cmd = [cc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
