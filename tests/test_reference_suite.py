"""The reference's OWN test-suite, unmodified, driving the B200 drop-in (north_star: "the simulators in
simulator/ and the tests/ suite drive it unchanged").

Mechanics: ``models/`` is an implicit namespace package both here and in the reference, so
with ``PYTHONPATH=<this repo>:<reference>`` the three hot-path modules (``models.particle_filter``,
``models.LEDH_particle_filter``, ``models.EDH_particle_filter``) resolve to this repository's overlays
while every other ``models.*`` / ``simulator.*`` import -- and the test files themselves -- come from the
unmodified reference tree (``baseline/_ref`` staged by ``tools/stage_reference.py``, or
``$SMC_REFERENCE_ROOT``).  Each file runs in its own pytest subprocess; the combined log goes to
``gpurun_out/r2_reference_suite.log`` when that directory exists (copied to ``profiles/`` for the record).
"""

from __future__ import annotations

import os
import re
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, "tools"))
from stage_reference import reference_root, stub_dir  # noqa: E402

pytestmark = pytest.mark.gpu

# (file relative to <reference>/tests, the number of tests the unmodified reference collects there)
SUITE = [
    ("unit_tests/models/test_pf_shapes_and_api.py", None),
    ("unit_tests/models/test_pf_resampling.py", None),
    ("unit_tests/models/test_ledh_flow_pf.py", None),
    ("unit_tests/models/test_edh_flow_pf.py", None),
    ("unit_tests/models/test_ekf_tracker_wrapper.py", None),
    ("unit_tests/models/test_ukf_tracker_wrapper.py", None),
    ("integration_tests/test_pf_vs_simulator_sv.py", None),
    ("integration_tests/test_filters_sv_simulator.py", None),
    ("integration_tests/test_filters_mat_simulator.py", None),
    ("integration_tests/test_filters_snlg_simulator.py", None),
    ("integration_tests/test_filters_skewt_simulator.py", None),
]

_PROBE = (
    "import models.particle_filter as a, models.LEDH_particle_filter as b, models.EDH_particle_filter as c, "
    "models.kalman_filter as k, simulator.simulator_LGSSM as s; "
    "print('OVERLAY', a.__file__, b.__file__, c.__file__); print('REFERENCE', k.__file__, s.__file__)"
)


def _env(root: str) -> dict:
    env = dict(os.environ)
    paths = [REPO, root]
    stub = stub_dir(root)
    if stub:
        paths.append(stub)
    env["PYTHONPATH"] = os.pathsep.join(paths + [env.get("PYTHONPATH", "")]).rstrip(os.pathsep)
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    return env


def _log(text: str) -> None:
    out = os.path.join(REPO, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "r2_reference_suite.log"), "a") as f:
            f.write(text)


@pytest.fixture(scope="module")
def ref_root():
    root = reference_root()
    if root is None:
        pytest.skip("the reference tree is not staged (run tools/stage_reference.py in the build container)")
    probe = subprocess.run([sys.executable, "-c", _PROBE], env=_env(root), cwd=REPO, capture_output=True, text=True,
                           timeout=600)
    assert probe.returncode == 0, probe.stdout + probe.stderr
    overlay, reference = probe.stdout.strip().splitlines()[-2:]
    # the three hot-path modules come from this repository, everything else from the reference tree
    assert overlay.count(os.path.join(REPO, "models")) == 3, overlay
    assert reference.count(os.path.realpath(root)) == 2 or reference.count(root) == 2, reference
    _log(f"# reference root: {root}\n# {overlay}\n# {reference}\n")
    return root


@pytest.mark.parametrize("rel,_n", SUITE, ids=[s[0].split("/")[-1] for s in SUITE])
def test_reference_test_file_passes_on_the_drop_in(ref_root, rel, _n):
    path = os.path.join(ref_root, "tests", rel)
    assert os.path.isfile(path), path
    cmd = [sys.executable, "-W", "ignore::SyntaxWarning", "-m", "pytest", "-q", "-p", "no:cacheprovider", f"--rootdir={ref_root}",
           "-c", os.devnull, "--durations=3", path]
    res = subprocess.run(cmd, env=_env(ref_root), cwd=REPO, capture_output=True, text=True, timeout=3000)
    tail = "\n".join((res.stdout + res.stderr).strip().splitlines()[-25:])
    _log(f"\n===== {rel} (exit {res.returncode}) =====\n{tail}\n")
    assert res.returncode == 0, tail
    m = re.search(r"(\d+) passed", res.stdout)
    assert m and int(m.group(1)) > 0, tail
    assert "failed" not in res.stdout.splitlines()[-1], tail
