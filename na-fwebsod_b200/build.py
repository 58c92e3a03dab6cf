"""Build libnawsod.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python na-fwebsod_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libnawsod.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", 
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
    # IEEE arithmetic everywhere: the RoI bin maths must match the reference bit for bit
    "--fmad=true", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
]


def _stamp(path: str) -> str:
    h = hashlib.sha1()
    h.update(" ".join(FLAGS).encode())
    for dep in [path] + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "nawsod.h")]:
        with open(dep, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs, dirty = [], False
    procs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        st = o + ".stamp"
        want = _stamp(s)
        have = open(st).read() if os.path.exists(st) else ""
        objs.append(o)
        if force or have != want or not os.path.exists(o):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), s, st, want))
            dirty = True
    for p, s, st, want in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + s)
        if verbose or out.strip():
            sys.stderr.write(out)
        with open(st, "w") as f:
            f.write(want)
    if dirty or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
