"""GPU parity tests of the device multinomial resampler (smc_multinomial_ancestors) against the
oracle restatement of ``rng.choice(N, size, p=w)`` (reference particle_filter.py:173-186).

Bar: ancestors bit-exact.  The device CDF is exact (2^-62 fixed point); the reference's is a
sequential fp64 cumsum, so a uniform that falls within the cumsum's rounding error of a CDF value
may legitimately land one cell away -- such a mismatch has to be explained by that distance.
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle.sir_numpy import multinomial_ancestors_from_uniforms

pytestmark = pytest.mark.gpu


def _weights(n, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "flat":
        w = np.ones(n)
    elif kind == "random":
        w = rng.random(n)
    elif kind == "zeros":
        w = rng.random(n)
        w[rng.random(n) < 0.6] = 0.0
        w[rng.integers(n)] = 1.0
    else:                       # "skewed": log-weights of a degenerate filter
        w = np.exp(-0.5 * (6.0 * rng.standard_normal(n)) ** 2)
        w[rng.integers(n)] = 1.0
    return w / w.sum()


def _assert_tie_explained(w, u, got, want):
    bad = np.flatnonzero(got != want)
    assert bad.size <= max(1, int(1e-5 * len(u))), f"{bad.size} ancestors differ"
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    for i in bad:
        lo, hi = sorted((int(got[i]), int(want[i])))
        assert np.all(np.abs(cdf[lo:hi] - u[i]) <= 64 * np.finfo(float).eps * len(w)), (i, got[i], want[i])


@pytest.mark.parametrize("n", [1, 2, 5, 2047, 2048, 2049, 4097, 100003, 1 << 20])
@pytest.mark.parametrize("kind", ["flat", "random", "zeros", "skewed"])
def test_device_ancestors_equal_the_choice_restatement(n, kind):
    from particle_filters_b200.engine import multinomial_ancestors

    w = _weights(n, kind, n + len(kind))
    u = np.random.default_rng(n).random(n)
    got = multinomial_ancestors(w, u)
    want = multinomial_ancestors_from_uniforms(w, u)
    assert got.dtype == np.int64 and got.shape == (n,)
    assert got.min() >= 0 and got.max() < n
    assert np.all(w[got] > 0), "a zero-weight particle was selected"
    _assert_tie_explained(w, u, got, want)


def test_edge_uniforms_and_output_size_other_than_n():
    from particle_filters_b200.engine import multinomial_ancestors

    w = np.array([0.0, 0.25, 0.0, 0.5, 0.25, 0.0])
    u = np.array([0.0, np.nextafter(0.25, 0), 0.25, 0.5, np.nextafter(0.75, 0), 0.75, np.nextafter(1.0, 0), 5e-324,
                  1e-300, 2.0 ** -53, 0.1])
    got = multinomial_ancestors(w, u)
    np.testing.assert_array_equal(got, multinomial_ancestors_from_uniforms(w, u))
    np.testing.assert_array_equal(got[:7], [1, 1, 3, 3, 3, 4, 4])


def test_private_helper_consumes_the_numpy_stream_like_rng_choice():
    from particle_filters_b200 import ParticleFilter

    n = 3000
    w = _weights(n, "random", 11)
    pf = ParticleFilter(lambda x, u: x, lambda x: x, np.eye(1), np.eye(1), Np=n, resample_method="multinomial",
                        rng=np.random.default_rng(123))
    ref_rng = np.random.default_rng(123)
    want = ref_rng.choice(n, size=n, p=w)                       # what the reference's helper returns
    got = pf._multinomial_resample(w)
    np.testing.assert_array_equal(got, want)
    assert pf.rng.random() == ref_rng.random()                  # both streams advanced identically
    with pytest.raises(ValueError, match="do not sum to 1"):
        pf._multinomial_resample(w * 1.01)
    with pytest.raises(ValueError, match="non-negative"):
        pf._multinomial_resample(np.array([1.5, -0.5]))


def test_on_chip_uniforms_are_reproducible_and_multinomial():
    from particle_filters_b200.engine import multinomial_ancestors

    n = 1 << 22
    w = _weights(n, "skewed", 5)
    a = multinomial_ancestors(w, None, seed=9, t=3)
    b = multinomial_ancestors(w, None, seed=9, t=3)
    c = multinomial_ancestors(w, None, seed=9, t=4)
    np.testing.assert_array_equal(a, b)
    assert np.mean(a != c) > 0.5
    counts = np.bincount(a, minlength=n).astype(float)
    # offspring counts are Multinomial(n, w): check the heavy cells one by one (5 sigma) and the total chi^2
    heavy = np.argsort(w)[-2000:]
    mu, sd = n * w[heavy], np.sqrt(n * w[heavy] * (1 - w[heavy]))
    assert np.all(np.abs(counts[heavy] - mu) <= 5.5 * sd + 1)
    z = (counts[heavy] - mu) / sd
    assert abs(np.mean(z)) < 0.15 and 0.85 < np.std(z) < 1.15
    assert counts[w == 0].sum() == 0
    # unlike systematic resampling the draws are i.i.d.: consecutive ancestors are not sorted
    assert np.mean(np.diff(a) < 0) > 0.3


def test_philox_multinomial_filter_runs_on_device_and_tracks_the_systematic_filter():
    from particle_filters_b200 import ParticleFilter, StochVolSSM

    spec = StochVolSSM(0.95, 1.0)
    Y = np.random.default_rng(4).standard_normal(25)
    means = {}
    for method in ("systematic", "multinomial"):
        pf = ParticleFilter(spec.g, spec.h, np.array([[0.04]]), np.array([[0.5]]), Np=1 << 18, resample_thresh=0.6,
                            resample_method=method, rng=np.random.default_rng(0), noise="philox", dtype="f32",
                            seed=17)
        pf.initialize(np.zeros(1), np.array([[0.5]]))
        means[method] = np.array([pf.step(np.array([y])).mean[0] for y in Y])
        assert np.isfinite(means[method]).all()
    assert np.max(np.abs(means["systematic"] - means["multinomial"])) < 0.02
