"""Compile the CUDA sources into ``particle_filters_b200/_lib/libsmc.so`` (sm_100a only).

``python -m particle_filters_b200.build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the shared library is git-ignored but travels to
the GPU box with the repo snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libsmc.so")
STAMP = os.path.join(OUT_DIR, "libsmc.stamp")

SOURCES = ["sir_kernels.cu", "ledh_kernels.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsmc.so cannot be built")


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/smc.h", "../../include/smc_ledh.h"]
    for name in files:
        p = os.path.normpath(os.path.join(CSRC, name))
        if os.path.isfile(p):
            with open(p, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == fp:
                return LIB
    # one nvcc process per translation unit (they share no device code), then one link step
    from concurrent.futures import ThreadPoolExecutor

    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    objs = [os.path.join(OUT_DIR, os.path.splitext(os.path.basename(src))[0] + ".o") for src in _sources()]

    def compile_one(pair):
        src, obj = pair
        return subprocess.run([_nvcc()] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(objs)) as pool:
        results = list(pool.map(compile_one, zip(_sources(), objs)))
    for res in results:
        if verbose:
            sys.stderr.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    res = subprocess.run([_nvcc(), "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    for obj in objs:
        os.remove(obj)
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
