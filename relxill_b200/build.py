"""Builds librelxill_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librelxill_b200.so")
SOURCES = ["models.cpp", "tables.cu", "kernels.cu", "line.cu", "xill.cu", "conv.cu", "nthcomp.cu", "peak.cu", "api.cu"]
# translation units compiled with FMA contraction enabled (everything else: -fmad=false)
FMAD_ON = {"line.cu", "xill.cu", "conv.cu", "peak.cu"}
HEADERS = ["common.h", "devutil.cuh", "models.h", "tables.h", "kernels.h", "minifits.h", "../../include/relxill_b200.h"]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    files += [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(f) > t for f in files if os.path.exists(f))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    names = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(name):
        obj = os.path.join(objdir, name + ".o")
        # IEEE mul/add are kept separate by default: the parity budget is spent on libm differences, not
        # on contraction; only the units listed in FMAD_ON opt in
        fmad = "-fmad=true" if name in FMAD_ON else "-fmad=false"
        cmd = [nvcc_path()] + NVCC_FLAGS + [fmad, "-c", os.path.join(CSRC, name), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, " ".join(cmd) + "\n" + r.stdout + r.stderr, r.returncode

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, names))
    log = "\n".join(r[1] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [r[0] for r in results]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log += "\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr
        rc = r.returncode
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if rc != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building librelxill_b200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
