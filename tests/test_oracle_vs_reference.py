"""Live differential check of the oracle against the importable reference.

Only runs where the reference tree exists (the build container); the GPU box
relies on tests/golden instead.  Reference modules are loaded by file path so
they cannot be confused with this repo's ``models/`` overlay.
"""

from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import REFERENCE_ROOT, needs_reference
from oracle import SIROracle, ledh_step, systematic_ancestors
from oracle.build import c_systematic_ancestors
from ssm_zoo import AcousticCVModel, LGSSMModel, SVModel

pytestmark = [needs_reference, pytest.mark.reference]


def _load(relpath, alias):
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


class _U:
    def __init__(self, u):
        self.u = u

    def random(self):
        return self.u


def test_resampler_200_random_vectors_bit_exact():
    ref = _load("models/particle_filter.py", "ref_particle_filter")
    rng = np.random.default_rng(5)
    for k in range(200):
        N = int(rng.integers(1, 3001))
        w = rng.random(N) ** rng.integers(1, 6)
        if k % 3 == 0 and N > 2:
            w[rng.random(N) < 0.3] = 0.0
            if w.sum() == 0:
                w[0] = 1.0
        w = w / w.sum()
        u = float(rng.random())
        pf = ref.ParticleFilter(lambda x, u_: x, lambda x: x, np.eye(1), np.eye(1), Np=N, rng=_U(u))
        want = pf._systematic_resample(w)
        np.testing.assert_array_equal(systematic_ancestors(w, u), want)
        np.testing.assert_array_equal(c_systematic_ancestors(w, u), want)


@pytest.mark.parametrize("model,nx", [(SVModel(0.9, 0.2, 1.0, r=0.1), 1), (LGSSMModel(), 2)])
@pytest.mark.parametrize("regularize", [False, True])
def test_sir_oracle_equals_reference_bitwise_with_per_particle_callbacks(model, nx, regularize):
    ref = _load("models/particle_filter.py", "ref_particle_filter")
    rng = np.random.default_rng(3)
    Y = rng.standard_normal((40, nx))
    pf = ref.ParticleFilter(model.g, model.h, model.Q, model.R, Np=257, resample_thresh=0.6,
                            regularize_after_resample=regularize, rng=np.random.default_rng(9))
    o = SIROracle(SIROracle.per_particle(model.g, True), SIROracle.per_particle(model.h, False),
                  model.Q, model.R, Np=257, resample_thresh=0.6,
                  regularize_after_resample=regularize, rng=np.random.default_rng(9))
    pf.initialize(np.zeros(nx), np.eye(nx))
    o.initialize(np.zeros(nx), np.eye(nx))
    for t in range(40):
        st = pf.step(Y[t])
        o.step(Y[t])
        np.testing.assert_array_equal(o.particles, st.particles)
        np.testing.assert_array_equal(o.weights, st.weights)
        np.testing.assert_array_equal(o.mean, st.mean)
        np.testing.assert_array_equal(o.cov, st.cov)


def test_ledh_oracle_matches_reference_step():
    ref = _load("models/LEDH_particle_filter.py", "ref_ledh_particle_filter")
    m = AcousticCVModel(n_targets=1, grid=(3, 3))
    rng = np.random.default_rng(21)
    N = 30
    P = np.diag([4.0, 4.0, 0.3, 0.3]) + 0.1

    class Tr:
        def predict(self):
            return np.zeros(4), P

        def update(self, z):
            return np.zeros(4), P

        def get_past_mean(self):
            return np.zeros(4)

    cfg = ref.LEDHConfig(n_particles=N, n_lambda_steps=5, resample_ess_ratio=0.5,
                         rng=np.random.default_rng(1))
    f = ref.LEDHFlowPF(Tr(), m.g, m.h, m.jac_h, m.log_trans_pdf, m.log_like_pdf, m.R, cfg)
    st0 = f.init_from_gaussian(np.array([12.0, 14.0, 0.5, -0.3]), np.diag([4.0, 4.0, 0.25, 0.25]))
    z = m.h(np.array([12.5, 13.0, 0.6, -0.2])) + 0.1 * rng.standard_normal(m.nz)
    V = rng.standard_normal((N, 4)) @ np.linalg.cholesky(m.Q).T
    st1 = f.step(st0, z, None, lambda n, d: V)
    rr = np.random.default_rng(1)
    rr.multivariate_normal(np.zeros(4), np.eye(4), size=N)
    x, w, mean, cov, info = ledh_step(
        st0.particles, st0.weights, z, P, m.R, g_batch=m.g_batch, h_batch=m.h_batch,
        jac_batch=m.jac_batch, log_trans_batch=m.log_trans_batch, log_like_batch=m.log_like_batch,
        n_lambda_steps=5, resample_ess_ratio=0.5, v=V, rng=rr)
    np.testing.assert_allclose(x, st1.particles, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(w, st1.weights, rtol=1e-8)
    np.testing.assert_allclose(mean, st1.mean, rtol=1e-10)
    np.testing.assert_allclose(cov, st1.cov, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(info["condition_numbers"], st1.diagnostics["condition_numbers"],
                               rtol=1e-10)
