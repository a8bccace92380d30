"""GPU parity tests of the systematic resampler (K3) through the C-ABI.

Oracle: oracle/ (pinned to the reference by tests/golden/resample_cases.npz).
Integer outputs must be bit-exact except at documented cumulative-weight ties
(tests/parity.py).
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle import systematic_ancestors as oracle_ancestors
from oracle.build import c_systematic_ancestors
from parity import assert_ancestors, check_ancestor_invariants

pytestmark = pytest.mark.gpu


def gpu_ancestors(w, u):
    from particle_filters_b200.engine import systematic_ancestors

    return systematic_ancestors(w, u)


def test_golden_cases_pf_and_ledh_flavours(golden):
    g = golden("resample_cases.npz")
    n = int(g["n_cases"])
    exact = ties = 0
    for k in range(n):
        w, u = g[f"c{k}_w"], float(g[f"c{k}_u"])
        got = gpu_ancestors(w, u)
        assert got.dtype == np.int64
        check_ancestor_invariants(got, len(w))
        nb = assert_ancestors(got, g[f"c{k}_pf"], w, u, clamp_last=True, msg=f"case {k} {g[f'c{k}_name']}")
        exact += nb == 0
        ties += nb
        if u != 0.0:   # u = 0 cases put positions exactly on CDF values (true ties)
            assert nb == 0, f"case {k} {g[f'c{k}_name']} (u={u}): {nb} mismatches away from ties"
        if bool(g[f"c{k}_ledh_ok"]):
            wn = w / np.sum(w)
            got2 = gpu_ancestors(wn, u)
            assert_ancestors(got2, g[f"c{k}_ledh"], wn, u, clamp_last=False, msg=f"ledh case {k}")
    print(f"resample golden: {exact}/{n} cases bit-exact, {ties} admissible tie mismatches (u = 0 cases)")
    assert exact >= (2 * n) // 3


@pytest.mark.parametrize("N", [1, 2, 5, 255, 256, 257, 2047, 2048, 2049, 4095, 4097, 65536 + 37])
def test_ragged_sizes_bit_exact(N):
    rng = np.random.default_rng(N)
    w = rng.random(N) ** 3
    w /= w.sum()
    u = float(rng.random())
    got = gpu_ancestors(w, u)
    check_ancestor_invariants(got, N)
    np.testing.assert_array_equal(got, oracle_ancestors(w, u))


@pytest.mark.parametrize("logN", [20, 22, 24])
def test_large_n_against_literal_c_oracle(logN):
    N = 1 << logN
    rng = np.random.default_rng(logN)
    w = np.exp(1.5 * rng.standard_normal(N))
    w /= w.sum()
    u = float(rng.random())
    got = gpu_ancestors(w, u)
    check_ancestor_invariants(got, N)
    ref = c_systematic_ancestors(w, u, True)
    # summation-order ties: ~N^2 * 1e-13 / few (0 / 6 / 450 at 2^22 / 2^24 / 2^26)
    assert_ancestors(got, ref, w, u, max_mismatch=max(4, N * N // (1 << 43)), msg=f"N=2^{logN}")
    # offspring counts: every particle gets floor(N w) or ceil(N w) (+-1 from the offset u)
    counts = np.bincount(got, minlength=N)
    assert np.all(np.abs(counts - N * w) < 1.0 + 1e-6 * N * w)


@pytest.mark.parametrize("kind", ["dominant", "first", "last", "zero_runs", "two_spikes"])
@pytest.mark.parametrize("N", [10_000, 1 << 20])
def test_degenerate_weight_shapes(kind, N):
    rng = np.random.default_rng(7)
    if kind == "dominant":
        w = np.full(N, 1e-4 / N)
        w[N // 3] = 0.9
    elif kind == "first":
        w = np.zeros(N)
        w[0] = 1.0
    elif kind == "last":
        w = np.zeros(N)
        w[-1] = 1.0
    elif kind == "zero_runs":
        w = rng.random(N)
        w[(np.arange(N) // 5000) % 2 == 0] = 0.0
    else:
        w = np.zeros(N)
        w[17] = 0.25
        w[N - 5] = 0.75
    w = w / w.sum()
    u = float(rng.random())
    got = gpu_ancestors(w, u)
    check_ancestor_invariants(got, N)
    assert_ancestors(got, c_systematic_ancestors(w, u, True), w, u, max_mismatch=2, msg=kind)


def test_uniform_weights_identity_away_from_ties():
    N = 1 << 18
    w = np.full(N, 1.0 / N)
    got = gpu_ancestors(w, 0.37)
    np.testing.assert_array_equal(got, np.arange(N))


def test_run_to_run_determinism():
    N = (1 << 21) + 3
    rng = np.random.default_rng(1)
    w = rng.random(N)
    w /= w.sum()
    a = gpu_ancestors(w, 0.123456789)
    for _ in range(3):
        np.testing.assert_array_equal(gpu_ancestors(w, 0.123456789), a)


@pytest.mark.slow
def test_full_size_2pow26_tie_rule_and_properties():
    """BASELINE config size: N = 2^26.  Checker = literal C loop; properties = sortedness,
    range, offspring counts within 1 of N w."""
    N = 1 << 26
    rng = np.random.default_rng(26)
    w = np.exp(rng.standard_normal(N))
    w /= w.sum()
    u = float(rng.random())
    got = gpu_ancestors(w, u)
    check_ancestor_invariants(got, N)
    ref = c_systematic_ancestors(w, u, True)
    n_bad = assert_ancestors(got, ref, w, u, max_mismatch=5000, msg="N=2^26")
    print(f"N=2^26: {n_bad} admissible tie mismatches")
    counts = np.bincount(got, minlength=N)
    assert np.all(np.abs(counts - N * w) < 1.0 + 1e-6 * N * w)
