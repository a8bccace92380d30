"""Pin the CPU oracle against the golden vectors generated from the reference.

CPU-only.  The fixtures were produced by tests/golden/make_golden.py from the
unmodified reference (models/particle_filter.py, models/LEDH_particle_filter.py).
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle import (SIROracle, ledh_step, systematic_ancestors, systematic_ancestors_loop)
from oracle.build import c_systematic_ancestors
from ssm_zoo import acoustic_c5_model, AcousticCVModel, LGSSMModel, Linear1DModel, SVModel, Lorenz96Model


def test_golden_rng_stream_is_stable(golden):
    """Fixtures store seeds, not draws: guard against a NumPy stream change."""
    g = golden("rng_stream.npz")
    r = np.random.default_rng(42)
    np.testing.assert_array_equal(r.standard_normal((4, 3)), g["a"])
    assert r.random() == float(g["b"])
    np.testing.assert_array_equal(r.standard_normal(5), g["c"])


def test_systematic_ancestors_match_reference_bit_for_bit(golden):
    g = golden("resample_cases.npz")
    n = int(g["n_cases"])
    assert n >= 100
    for k in range(n):
        w, u = g[f"c{k}_w"], float(g[f"c{k}_u"])
        ref = g[f"c{k}_pf"]
        got = systematic_ancestors(w, u, clamp_last=True)
        assert got.dtype == np.int64
        np.testing.assert_array_equal(got, ref, err_msg=f"case {k} {g[f'c{k}_name']}")
        np.testing.assert_array_equal(c_systematic_ancestors(w, u, True), ref)
        if len(w) <= 300:
            np.testing.assert_array_equal(systematic_ancestors_loop(w, u, clamp_last=True), ref)
        if bool(g[f"c{k}_ledh_ok"]):
            wn = w / np.sum(w)
            np.testing.assert_array_equal(systematic_ancestors(wn, u, clamp_last=False), g[f"c{k}_ledh"])
            np.testing.assert_array_equal(c_systematic_ancestors(wn, u, False), g[f"c{k}_ledh"])


def _run_oracle(g, model, T=None):
    Y = g["Y"]
    T = len(Y) if T is None else T
    o = SIROracle(model.g_batch, model.h_batch, g["Q"], g["R"], Np=int(g["Np"]),
                  resample_thresh=float(g["thresh"]), resample_method=str(g["method"]),
                  regularize_after_resample=bool(g["regularize"]),
                  rng=np.random.default_rng(int(g["seed"])))
    o.initialize(g["mean0"], g["cov0"])
    means, covs, ess, res = [], [], [], []
    anc = {}
    for t in range(T):
        o.step(np.atleast_1d(Y[t]))
        means.append(o.mean)
        covs.append(o.cov)
        ess.append(o.info["ess_pre"])
        res.append(o.info["resampled"])
        if o.info["ancestors"] is not None and str(g["method"]) == "systematic":
            anc[t] = o.info["ancestors"]
    return o, np.array(means), np.array(covs), np.array(ess), np.array(res), anc


@pytest.mark.parametrize("name,model", [
    ("sir_sv_c1.npz", SVModel(0.95, 0.2, 1.0)),
    ("sir_sv_c1_r01.npz", SVModel(0.95, 0.2, 1.0, r=0.1)),
    ("sir_sv_bootstrap.npz", SVModel(0.95, 0.2, 1.0)),
    ("sir_sv_regularized.npz", SVModel(0.95, 0.2, 1.0)),
    ("sir_sv_multinomial.npz", SVModel(0.95, 0.2, 1.0)),
    ("sir_lgssm.npz", LGSSMModel()),
    ("sir_l96_nx8.npz", Lorenz96Model(nx=8, n_sub=5, obs_stride=2)),       # C4-shaped, reduced
    ("sir_l96_nx40.npz", Lorenz96Model(nx=40, n_sub=20, obs_stride=4)),    # C4 dynamics at small N
])
def test_sir_oracle_reproduces_reference_run(golden, name, model):
    """Whole filter runs: identical resampling decisions and ancestors; moments
    to 1e-12 (batched g/h differ from the per-particle callbacks by <= 1 ulp)."""
    g = golden(name)
    o, means, covs, ess, res, anc = _run_oracle(g, model)
    np.testing.assert_array_equal(res, g["resampled"])
    np.testing.assert_allclose(means, g["means"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(covs, g["covs"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ess, g["ess_pre"], rtol=1e-12)
    np.testing.assert_allclose(o.particles, g["final_particles"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(o.weights, g["final_weights"], rtol=1e-12)
    Np = int(g["Np"])
    for t in g["anc_steps"]:
        np.testing.assert_array_equal(anc[int(t)], g[f"anc_t{int(t)}"])
    for t, idx in anc.items():
        h = int(np.sum(idx.astype(np.int64) * (np.arange(Np) % 1009 + 1)) % (2 ** 61 - 1))
        assert h == int(g["anc_hash"][t]), f"ancestor checksum differs at step {t}"


@pytest.mark.parametrize("name,model,tol", [
    ("ledh_linear1d.npz", Linear1DModel(), 1e-9),
    ("ledh_linear1d_resample.npz", Linear1DModel(), 1e-9),
    ("ledh_acoustic1.npz", AcousticCVModel(n_targets=1, grid=(3, 3)), 1e-9),
    # stiff case: cond(S) up to 6e4 and a near-degenerate weight set amplify the
    # batched-vs-looped LAPACK round-off to ~3e-8 relative
    ("ledh_acoustic2.npz", AcousticCVModel(n_targets=2, grid=(4, 4)), 1e-6),
    # BASELINE config 5's shape (16-D state, 25 sensors, R = 0.01 I): the stiffest case
    ("ledh_acoustic4.npz", acoustic_c5_model(), 1e-6),
])
def test_ledh_oracle_reproduces_reference_steps(golden, name, model, tol):
    g = golden(name)
    N, L, T = int(g["N"]), int(g["L"]), int(g["T"])
    x, w = g["particles0"], g["weights0"]
    for k in range(T):
        v = g["V"][k] if bool(g["with_noise"]) else None
        x, w, mean, cov, info = ledh_step(
            x, w, g["Z"][k], g["Ps"][k], g["R"], g_batch=model.g_batch, h_batch=model.h_batch,
            jac_batch=model.jac_batch, log_trans_batch=model.log_trans_batch,
            log_like_batch=model.log_like_batch, n_lambda_steps=L,
            resample_ess_ratio=float(g["ess_ratio"]), v=v, rng=_FixedUFromRef(g, k))
        np.testing.assert_allclose(x, g[f"particles_{k}"], rtol=tol, atol=tol)
        np.testing.assert_allclose(w, g[f"weights_{k}"], rtol=100 * tol, atol=1e-12)
        np.testing.assert_allclose(mean, g[f"mean_{k}"], rtol=tol, atol=tol)
        np.testing.assert_allclose(cov, g[f"cov_{k}"], rtol=10 * tol, atol=10 * tol)
        np.testing.assert_allclose(info["condition_numbers"], g[f"cond_{k}"], rtol=1e-8)
        # continue from the reference's state so one step's round-off cannot compound
        x, w = g[f"particles_{k}"], g[f"weights_{k}"]


class _FixedUFromRef:
    """The resampling uniform is drawn from cfg.rng = default_rng(seed+1) after the
    init draw; replay that stream."""

    _streams = {}

    def __init__(self, g, k):
        key = (int(g["seed"]), int(g["N"]))
        if key not in self._streams or k == 0:
            r = np.random.default_rng(int(g["seed"]) + 1)
            nx = g["particles0"].shape[1]
            r.multivariate_normal(np.zeros(nx), np.eye(nx), size=int(g["N"]))
            self._streams[key] = r
        self.r = self._streams[key]

    def random(self):
        return self.r.random()


def test_multinomial_restatement_equals_rng_choice():
    """oracle.multinomial_ancestors_from_uniforms is what rng.choice(N, size=N, p=w) computes from
    rng.random(N) (reference particle_filter.py:173-186 delegates to it); both generators end in
    the same state."""
    from oracle.sir_numpy import multinomial_ancestors_from_uniforms

    for n in (1, 7, 1000, 4097):
        w = np.random.default_rng(n).random(n)
        if n > 5:
            w[::5] = 0.0
        w /= w.sum()
        r1, r2 = np.random.default_rng(5), np.random.default_rng(5)
        want = r1.choice(n, size=n, p=w)
        got = multinomial_ancestors_from_uniforms(w, r2.random(n))
        np.testing.assert_array_equal(got, want)
        assert got.dtype == want.dtype
        assert r1.random() == r2.random()
