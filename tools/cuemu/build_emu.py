"""Builds tools/cuemu/_build/liboak_b200_emu.so: the library's own .cu sources compiled with g++ against the
CPU emulation of the CUDA execution model (cuemu.h).  TEST HARNESS ONLY — see cuemu.h.

    python tools/cuemu/build_emu.py [--force] [--asan] [-D NAME=VALUE ...] [--out path]

--asan builds with AddressSanitizer (shared-memory arrays are globals / heap blocks there, so out-of-bounds
indexing inside a kernel is reported); run the tests with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0.

Source rewriting (the only things g++ cannot parse): `kernel<<<grid, block, smem, stream>>>(args);` becomes
`cuemu::launch(grid, block, smem, [&]() { kernel(args); });` and `extern __shared__ T name[];` becomes a pointer
to the block's dynamic shared memory.  Inline PTX sits behind `#ifdef OAK_CUEMU` alternatives in the sources.
"""
import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "oak_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
SOURCES = ["api.cu", "obsgrid.cu", "gram.cu", "gram_mma.cu", "eig_simple.cu", "eig_fast.cu", "eig_tridiag.cu", "apply.cu", "apply_mma.cu", "global.cu",
           "ensemble.cu", "hgen.cu", "microbench.cu"]
CXXFLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-mfma", "-fno-strict-aliasing",
            "-Wno-attributes", "-Wno-unknown-pragmas", "-DOAK_CUEMU=1", "-I" + os.path.join(HERE, "shim"), "-I" + HERE]

LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^;<>]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)
DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];")


def rewrite(text):
    def launch(mo):
        cfg = [c.strip() for c in split_top(mo.group(2))]
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        return f"cuemu::launch(dim3({grid}), dim3({block}), (size_t)({smem}), [&]() {{ {mo.group(1)}({mo.group(3)}); }});"

    text = LAUNCH.sub(launch, text)
    text = DYN_SMEM.sub(lambda mo: f"{mo.group(1)} *{mo.group(2)} = reinterpret_cast<{mo.group(1)} *>(cuemu::dyn_smem());", text)
    return text


def split_top(s):
    out, depth, cur = [], 0, ""
    s = s.replace("->", "\x00")   # a member access is not a closing bracket
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return [c.replace("\x00", "->") for c in out]


def build(force=False, defines=(), out=None, asan=False):
    os.makedirs(BUILD, exist_ok=True)
    out = out or os.path.join(BUILD, "liboak_b200_emu_asan.so" if asan else "liboak_b200_emu.so")
    cxxflags = CXXFLAGS + (["-O1", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else [])
    defines = list(defines) + (["CUEMU_ASAN=1"] if asan else [])
    tag = hashlib.sha256(repr(sorted(defines)).encode()).hexdigest()[:8]
    h = hashlib.sha256()
    for root in (CSRC, HERE, os.path.join(HERE, "shim"), os.path.join(HERE, "shim", "cub"), os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p) and f.split(".")[-1] in ("cu", "cuh", "h", "cpp", "py"):
                h.update(open(p, "rb").read())
    h.update(repr((CXXFLAGS, sorted(defines))).encode())
    stamp_file = out + ".stamp"
    if not force and os.path.exists(out) and os.path.exists(stamp_file) and open(stamp_file).read() == h.hexdigest():
        return out
    gen = os.path.join(BUILD, "gen_" + tag)
    os.makedirs(gen, exist_ok=True)
    # the rewritten sources keep their relative includes: mirror csrc/ two levels below a fake root
    fake = os.path.join(gen, "oak_b200", "csrc")
    os.makedirs(fake, exist_ok=True)
    inc = os.path.join(gen, "include")
    if not os.path.exists(inc):
        os.symlink(os.path.join(ROOT, "include"), inc)
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            open(os.path.join(fake, f), "w").write(rewrite(open(os.path.join(CSRC, f)).read()))

    def one(src):
        cpp = os.path.join(fake, src.replace(".cu", ".emu.cpp"))
        open(cpp, "w").write(rewrite(open(os.path.join(CSRC, src)).read()))
        obj = cpp.replace(".cpp", ".o")
        cmd = ["g++"] + cxxflags + ["-D" + d for d in defines] + ["-c", cpp, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"g++ failed on {src}:\n{p.stderr[:6000]}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, SOURCES))
    rt = os.path.join(gen, "cuemu_rt.o")
    subprocess.check_call(["g++"] + cxxflags + ["-c", os.path.join(HERE, "cuemu_rt.cpp"), "-o", rt])
    subprocess.check_call(["g++", "-shared", "-o", out] + objs + [rt] + (["-fsanitize=address"] if asan else []))
    open(stamp_file, "w").write(h.hexdigest())
    return out


if __name__ == "__main__":
    defs = [sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "-D"]
    o = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, defines=defs, out=o, asan="--asan" in sys.argv))
