"""GPU tests of the class-level helpers the reference's own suite pins (tests/unit_tests/models/
test_pf_resampling.py:107-170, :363-383 -- `_resample` both branches, `effective_sample_size`) and of the
spec-path API corners named by the round-1 advisor (control input, `state.particles` assignment, predict-only
Philox draws, device selection)."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import systematic_ancestors

pytestmark = pytest.mark.gpu


def _bridge_pf(**kw):
    from particle_filters_b200 import ParticleFilter

    args = dict(Np=1000, resample_method="systematic", resample_thresh=0.5, rng=np.random.default_rng(123))
    args.update(kw)
    return ParticleFilter(lambda x, u=None: 0.9 * x, lambda x: x, np.array([[0.1]]), np.array([[0.2]]), **args)


def test_resample_is_triggered_by_low_neff_and_matches_the_reference_arithmetic():
    """reference particle_filter.py:188-220 / test_pf_resampling.py:107-139: low Neff -> systematic resampling with
    u = rng.random(), gathered particles, uniform weights."""
    pf = _bridge_pf()
    rng = np.random.default_rng(7)
    particles = rng.standard_normal((1000, 1))
    weights = np.zeros(1000)
    weights[0] = 0.9
    weights[1:] = 0.1 / 999
    assert 1.0 / np.sum(weights ** 2) < 500
    u = np.random.default_rng(123).random()                    # the draw the filter's rng will make
    new_p, new_w = pf._resample(particles.copy(), weights.copy())
    np.testing.assert_allclose(new_w, np.ones(1000) / 1000, rtol=1e-10)
    idx = systematic_ancestors(weights, u)                      # oracle restatement of :146-171
    np.testing.assert_array_equal(new_p, particles[idx])
    assert np.sum(np.abs(new_p[:, 0] - particles[0, 0]) < 1e-12) > 850


def test_resample_leaves_a_healthy_population_alone():
    """test_pf_resampling.py:141-170: Neff >= thresh * Np -> particles and weights returned unchanged and the
    filter's rng is NOT advanced (the reference draws u only when it resamples)."""
    pf = _bridge_pf()
    rng = np.random.default_rng(8)
    particles = rng.standard_normal((1000, 1))
    weights = np.ones(1000) / 1000 + rng.standard_normal(1000) * 0.0001
    weights = weights / np.sum(weights)
    assert 1.0 / np.sum(weights ** 2) > 500
    p0, w0 = particles.copy(), weights.copy()
    new_p, new_w = pf._resample(particles, weights)
    np.testing.assert_array_equal(new_p, p0)
    np.testing.assert_array_equal(new_w, w0)
    assert pf.rng.random() == np.random.default_rng(123).random()


@pytest.mark.parametrize("spec", [False, True])
def test_effective_sample_size_is_one_over_sum_w_squared(spec):
    """test_pf_resampling.py:363-383, bridge and spec (device-resident) filters; then after a weighted, non-resampling
    update, where the weights are no longer uniform."""
    from particle_filters_b200 import ParticleFilter, StochVolSSM

    if spec:
        m = StochVolSSM(0.95, 1.0)
        pf = ParticleFilter(m.g, m.h, np.array([[0.04]]), np.array([[1.0]]), Np=1000, resample_thresh=0.0,
                            rng=np.random.default_rng(3))
    else:
        pf = _bridge_pf(resample_thresh=0.0)
    pf.initialize(np.array([0.0]), np.array([[1.0]]))
    w = pf.state.weights
    np.testing.assert_allclose(pf.effective_sample_size(), 1.0 / np.sum(w ** 2), rtol=1e-12)
    np.testing.assert_allclose(pf.effective_sample_size(), 1000.0, rtol=1e-12)
    st = pf.step(np.array([0.7]))
    w = np.asarray(st.weights)
    assert np.std(w) > 0
    np.testing.assert_allclose(pf.effective_sample_size(), 1.0 / np.sum(w ** 2), rtol=1e-9)


def test_spec_path_api_corners():
    import torch

    from particle_filters_b200 import ParticleFilter, StochVolSSM
    from particle_filters_b200.engine import SirEngine

    m = StochVolSSM(0.95, 1.0)
    pf = ParticleFilter(m.g, m.h, np.array([[0.04]]), np.array([[1.0]]), Np=4096, resample_thresh=0.0, dtype="f64",
                        noise="philox", seed=9)
    pf.initialize(np.array([0.0]), np.array([[0.5]]))
    # a control input must not be dropped silently
    with pytest.raises(ValueError, match="control input"):
        pf.step(np.array([0.3]), u=np.array([1.0]))
    with pytest.raises(ValueError, match="control input"):
        pf.predict(u=np.array([1.0]))
    # assigning state.particles leaves the weights alone (the reference mutates the array only)
    st = pf.step(np.array([0.3]))
    w_before = np.asarray(st.weights).copy()
    assert np.std(w_before) > 0
    st.particles = np.asarray(st.particles) + 1.0
    np.testing.assert_allclose(pf._eng.weights(), w_before, rtol=1e-12)
    np.testing.assert_allclose(pf._eng.get_particles(), st.particles, rtol=0, atol=0)
    # two predict-only calls draw different noise
    x0 = pf._eng.get_particles().copy()
    pf.predict()
    x1 = pf._eng.get_particles().copy()
    pf.predict()
    x2 = pf._eng.get_particles().copy()
    n1, n2 = x1 - 0.95 * x0, x2 - 0.95 * x1
    assert np.std(n1) > 0.1 and np.max(np.abs(n1 - n2)) > 0.1
    # an engine runs on the current device only
    if torch.cuda.device_count() > 1:
        with pytest.raises(ValueError, match="current CUDA device"):
            SirEngine(1024, 1, 1, device="cuda:1")
    with pytest.raises(ValueError, match="current CUDA device"):
        SirEngine(1024, 1, 1, device="cpu")
