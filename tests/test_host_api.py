"""CPU-only checks: the C-ABI library loads and exports what include/*.h declares,
and the host-side mirror of the reference's class API behaves like the reference
before any GPU work is requested (no compute calls here)."""

from __future__ import annotations

import ctypes
import glob
import os
import re

import numpy as np
import pytest

from conftest import REPO


def _declared_functions():
    names = []
    for hdr in sorted(glob.glob(os.path.join(REPO, "include", "*.h"))):
        src = open(hdr).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(smc_\w+)\s*\(", src, flags=re.M):
            if "static inline" in src[max(0, m.start() - 40):m.start() + 20]:
                continue
            names.append(m.group(1))
    return sorted(set(names))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from particle_filters_b200 import _lib, build

    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 18, declared
    missing = [n for n in declared if n != "smc_stats_stride" and not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert lib.smc_abi_version() == _lib.SMC_ABI_VERSION
    assert lib.smc_sizeof_sir_desc() == ctypes.sizeof(_lib.SirDesc)


def test_workspace_size_is_small_and_monotone():
    from particle_filters_b200 import _lib

    lib = _lib.load()
    a = lib.smc_sir_workspace_bytes(1000, 1, _lib.SMC_F64)
    b = lib.smc_sir_workspace_bytes(1 << 26, 1, _lib.SMC_F32)
    assert 0 < a <= b < (8 << 20)          # a few MiB beside 2^26 particles
    assert lib.smc_partial_stride(1) == 5 and lib.smc_partial_stride(2) == 8 and lib.smc_partial_stride(40) == 3


def test_descriptor_validation_reports_errors_without_a_gpu():
    from particle_filters_b200 import _lib

    lib = _lib.load()
    d = _lib.SirDesc(n=10, n_global=5, ld=12, nx=1, nz=1)
    rc = lib.smc_sir_reset(ctypes.byref(d), None, None)
    assert rc != 0 and b"n_global" in lib.smc_last_error()
    d = _lib.SirDesc(n=10, n_global=10, ld=10, nx=1, nz=1)
    assert lib.smc_sir_reset(ctypes.byref(d), None, None) != 0 and b"ld" in lib.smc_last_error()


def test_constructor_mirrors_reference_attributes():
    from particle_filters_b200 import ParticleFilter

    g, h = (lambda x, u: 0.9 * x), (lambda x: x)
    pf = ParticleFilter(g, h, [[0.1]], [[0.5]])
    assert pf.g is g and pf.h is h
    assert isinstance(pf.Q, np.ndarray) and isinstance(pf.R, np.ndarray)
    assert pf.Np == 1000 and pf.resample_thresh == 0.5 and pf.resample_method == "systematic"
    assert pf.regularize_after_resample is False and pf.state is None
    assert pf.nx == 1 and pf.nz == 1
    np.testing.assert_allclose(pf.LR, np.sqrt(0.5 + 1e-12))
    assert isinstance(pf.rng, np.random.Generator)
    pf = ParticleFilter(g, h, np.eye(3), np.eye(2), Np=50, resample_thresh=0.3, resample_method="multinomial",
                        regularize_after_resample=True, rng=np.random.default_rng(1))
    assert (pf.nx, pf.nz, pf.Np) == (3, 2, 50)


def test_methods_assert_before_initialize_like_the_reference():
    from particle_filters_b200 import ParticleFilter

    pf = ParticleFilter(lambda x, u: x, lambda x: x, np.eye(1), np.eye(1), Np=10)
    for call in (lambda: pf.predict(), lambda: pf.update(np.zeros(1)), lambda: pf.effective_sample_size(),
                 lambda: pf.step(np.zeros(1))):
        with pytest.raises(AssertionError, match="Filter not initialized"):
            call()


def test_pfstate_is_positional_dataclass():
    from particle_filters_b200 import PFState

    s = PFState(np.zeros((3, 1)), np.ones(3) / 3, np.zeros(1), np.eye(1), 4)
    assert s.t == 4 and s.particles.shape == (3, 1)


def test_spec_detection_and_functor_packing():
    from particle_filters_b200 import LinearGaussianSSM, StochVolSSM, _lib
    from particle_filters_b200.ssm import spec_of

    sv = StochVolSSM(0.95, 1.0)
    assert spec_of(sv.g, sv.h) is sv
    assert spec_of(sv.g, lambda x: x) is None
    assert spec_of(lambda x, u: x, lambda x: x) is None
    mid, p = sv.device_functor(np.array([[0.04]]), np.array([[1.0]]))
    assert mid == _lib.SMC_MODEL_STOCH_VOL
    np.testing.assert_allclose(p, [0.95, 0.2, 1.0, np.sqrt(1 + 1e-12)])
    np.testing.assert_allclose(sv.g(np.array([2.0]), None), [1.9])
    np.testing.assert_allclose(sv.h(np.array([0.0])), [1.0])
    lg = LinearGaussianSSM([[0.9, 0.5], [0.0, 0.7]], np.eye(2))
    mid, p = lg.device_functor(np.diag([0.05, 0.02]), 0.1 * np.eye(2))
    assert mid == _lib.SMC_MODEL_LINEAR_GAUSSIAN and p.size == 4 + 4 + 4 + 4
    # LRinv block: inverse of chol(R + 1e-12 I)
    np.testing.assert_allclose(p[12:].reshape(2, 2), np.eye(2) / np.sqrt(0.1 + 1e-12))


def test_models_overlay_exports_reference_symbols():
    import importlib

    m = importlib.import_module("models.particle_filter")
    for name in ("Array", "GFn", "HFn", "PFState", "ParticleFilter"):
        assert hasattr(m, name)


def test_engine_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from particle_filters_b200 import ParticleFilter, _lib

    pf = ParticleFilter(lambda x, u: x, lambda x: x, np.eye(1), np.eye(1), Np=8)
    pf.initialize(np.zeros(1), np.eye(1))            # host draw only
    with pytest.raises(_lib.SmcError, match="no CPU fallback"):
        pf.predict()
