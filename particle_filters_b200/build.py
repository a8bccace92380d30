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

SOURCES = ["sir_kernels.cu", "ledh_kernels.cu", "multinomial.cu"]
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


def _hash(names) -> str:
    h = hashlib.sha256()
    for name in names:
        p = os.path.normpath(os.path.join(CSRC, name))
        if os.path.isfile(p):
            with open(p, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _headers():
    return sorted(n for n in os.listdir(CSRC) if not n.endswith(".cu")) + ["../../include/smc.h", "../../include/smc_ledh.h"]


def _deps(name, seen=None):
    """``name`` plus the quoted includes it reaches (paths relative to csrc/)."""
    import re

    seen = [] if seen is None else seen
    if name in seen:
        return seen
    seen.append(name)
    path = os.path.normpath(os.path.join(CSRC, name))
    if os.path.isfile(path):
        with open(path) as f:
            for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', f.read(), flags=re.M):
                _deps(os.path.normpath(os.path.join(os.path.dirname(name), inc)), seen)
    return seen


def _fingerprint() -> str:
    return _hash(sorted(n for n in os.listdir(CSRC) if n.endswith(".cu")) + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == fp:
                return LIB
    # one nvcc process per translation unit (they share no device code), then one link step; an
    # object whose source and headers are unchanged since it was compiled is reused
    from concurrent.futures import ThreadPoolExecutor

    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    jobs = []
    objs = []
    for src in _sources():
        stem = os.path.splitext(os.path.basename(src))[0]
        obj, ostamp = os.path.join(OUT_DIR, stem + ".o"), os.path.join(OUT_DIR, stem + ".ostamp")
        ofp = _hash(_deps(os.path.basename(src)))
        objs.append(obj)
        fresh = False
        if not force and not verbose and os.path.exists(obj) and os.path.exists(ostamp):
            with open(ostamp) as f:
                fresh = f.read().strip() == ofp
        if not fresh:
            jobs.append((src, obj, ostamp, ofp))

    def compile_one(job):
        src, obj, ostamp, ofp = job
        if os.path.exists(ostamp):
            os.remove(ostamp)
        return subprocess.run([_nvcc()] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            results = list(pool.map(compile_one, jobs))
        for job, res in zip(jobs, results):
            if verbose:
                sys.stderr.write(res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            with open(job[2], "w") as f:
                f.write(job[3])
    res = subprocess.run([_nvcc(), "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
