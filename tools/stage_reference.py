"""Stage the UNMODIFIED reference tree under ``baseline/_ref/`` (git-ignored, travels to the GPU box).

The reference is pure Python + NumPy with no packaging metadata (no setup.py / pyproject), so the
"install" of the base contract is a plain copy of the three directories its own test-suite needs:
``models/``, ``simulator/`` and ``tests/`` (about 1 MB; the 18 MB of notebooks stay behind).  Nothing
under ``baseline/_ref`` is product source: it is only ever *run* -- by ``tests/test_reference_suite.py``
(the reference's tests driving the B200 drop-in through the namespace overlay) and by
``bench.py --impl reference`` (the reference's own CPU filter, timed beside the GPU arm).

    python tools/stage_reference.py [--src /root/reference]
"""

from __future__ import annotations

import argparse
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(REPO, "baseline", "_ref")
PARTS = ("models", "simulator", "tests", "requirements.txt", "LICENSE")

MATPLOTLIB_STUB = '''"""Empty stand-in: the reference's tests/conftest.py imports a simulator module that imports
matplotlib at module scope (for plotting helpers the tests never call); matplotlib is not installed."""
'''


def stage(src: str = "/root/reference", dest: str = DEST, quiet: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(src, "models")):
        if not quiet:
            print(f"[stage_reference] no reference tree at {src}; nothing staged")
        return None
    os.makedirs(dest, exist_ok=True)
    for part in PARTS:
        s, d = os.path.join(src, part), os.path.join(dest, part)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache"))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    stub = os.path.join(dest, "_stubs", "matplotlib")
    os.makedirs(stub, exist_ok=True)
    for name in ("__init__.py", "pyplot.py"):
        with open(os.path.join(stub, name), "w") as f:
            f.write(MATPLOTLIB_STUB)
    if not quiet:
        print(f"[stage_reference] staged {src} -> {dest}")
    return dest


def reference_root() -> str | None:
    """Where the unmodified reference can be imported from at run time: ``$SMC_REFERENCE_ROOT``,
    the staged copy, or the build container's read-only tree."""
    for cand in (os.environ.get("SMC_REFERENCE_ROOT"), DEST, "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "models")) and os.path.isdir(os.path.join(cand, "simulator")):
            return cand
    return None


def stub_dir(root: str) -> str | None:
    """Directory holding the matplotlib stand-in (only needed when matplotlib is absent)."""
    try:
        import matplotlib  # noqa: F401

        return None
    except Exception:
        pass
    cand = os.path.join(root, "_stubs")
    if os.path.isdir(cand):
        return cand
    cand = os.path.join(DEST, "_stubs")
    return cand if os.path.isdir(cand) else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    args = ap.parse_args()
    sys.exit(0 if stage(args.src) else 1)
