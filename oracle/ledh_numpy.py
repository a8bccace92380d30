"""NumPy restatement of one LEDH particle-flow step -- TEST INFRASTRUCTURE ONLY.

Follows ``models/LEDH_particle_filter.py:93-214`` of the reference.  The
reference runs a Python loop over particles inside a loop over pseudo-time
steps; here every per-particle quantity is a batched array (leading axis N) and
the small solves go through NumPy's batched LAPACK gufuncs, so the arithmetic
is the same fp64 LAPACK/BLAS work in a different batching (agreement with the
reference is therefore ~1e-12 relative, not bit-exact; tested in
tests/test_oracle_vs_reference.py and pinned by tests/golden/ledh_*.npz).
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

Array = np.ndarray


def ledh_systematic_ancestors(weights: Array, u: float) -> Array:
    """Reference ``LEDH_particle_filter.py:25-37``: renormalise, no ``cdf[-1]=1``."""
    from .sir_numpy import systematic_ancestors

    w = weights / np.sum(weights)
    return systematic_ancestors(w, u, clamp_last=False)


def ledh_weighted_stats(x: Array, w: Array) -> Tuple[Array, Array]:
    """Reference ``LEDH_particle_filter.py:217-224``."""
    w = w / np.sum(w)
    mean = np.sum(x * w[:, None], axis=0)
    xc = x - mean[None, :]
    cov = (xc.T * w) @ xc
    return mean, 0.5 * (cov + cov.T)


def ledh_flow(
    eta0: Array,
    P: Array,
    R: Array,
    z: Array,
    h_batch: Callable[[Array], Array],
    jac_batch: Callable[[Array], Array],
    n_lambda_steps: int,
):
    """Pseudo-time integration, reference ``LEDH_particle_filter.py:132-179``.

    Returns ``(eta, theta_log, cond_numbers)``.
    """
    N, nx = eta0.shape
    I = np.eye(nx)
    eta = eta0.copy()
    theta_log = np.zeros(N)
    cond = []
    L = max(1, int(n_lambda_steps))
    dlam = 1.0 / float(L)
    lam = 0.0
    z = np.asarray(z, float)
    for _ in range(L):
        lam = min(1.0, lam + dlam)  # lambda is advanced BEFORE use (:137)
        H = np.asarray(jac_batch(eta), float)                 # (N, nz, nx)   :143
        hx = np.asarray(h_batch(eta), float).reshape(N, -1)   # (N, nz)       :144
        e = hx - np.einsum("nij,nj->ni", H, eta)              # :145
        PHt = np.einsum("ij,nkj->nik", P, H)                  # P H^T (N, nx, nz)
        S = lam * np.einsum("nij,njk->nik", H, PHt) + R       # :149
        try:
            cond.append(float(np.linalg.cond(S[0])))          # :152-157 (particle 0 only)
        except Exception:  # pragma: no cover - mirrors the bare except
            cond.append(float("nan"))
        SinvH = np.linalg.solve(S, H)                         # :158
        A = -0.5 * np.einsum("nik,nkj->nij", PHt, SinvH)      # :159
        Rinv_innov = np.linalg.solve(R, (z - e).T).T          # :163
        v = np.einsum("nik,nk->ni", PHt, Rinv_innov)          # :164
        inner = v + lam * np.einsum("nij,nj->ni", A, v) + np.einsum("nij,nj->ni", A, eta0)
        b = inner + 2.0 * lam * np.einsum("nij,nj->ni", A, inner)   # :165
        eta = eta + dlam * (np.einsum("nij,nj->ni", A, eta) + b)    # :171 (explicit Euler, old eta)
        M = I + dlam * A
        sign, logdet = np.linalg.slogdet(M)                   # :174-179
        bad = sign <= 0
        if np.any(bad):
            s2, l2 = np.linalg.slogdet(M[bad] + 1e-12 * I)
            logdet = logdet.copy()
            logdet[bad] = l2
        theta_log += logdet
    return eta, theta_log, cond


def ledh_step(
    particles: Array,
    weights: Array,
    z: Array,
    P: Array,
    R: Array,
    *,
    g_batch: Callable[[Array, Optional[Array], Array], Array],
    h_batch: Callable[[Array], Array],
    jac_batch: Callable[[Array], Array],
    log_trans_batch: Callable[[Array, Array], Array],
    log_like_batch: Callable[[Array, Array], Array],
    n_lambda_steps: int = 8,
    resample_ess_ratio: float = 0.0,
    u_km1=None,
    v: Optional[Array] = None,
    rng=None,
):
    """One LEDH step given the tracker's predicted covariance ``P``.

    The tracker calls (``tracker.predict()`` before, ``tracker.update(z)`` after
    the weights, reference ``:105,198``) are host-side EKF/UKF work outside the
    hot path; the caller performs them.  ``v`` is the process-noise draw
    (``process_noise_sampler(N, nx)``, ``:109-112``) or ``None`` for zeros.
    """
    N, nx = particles.shape
    P = 0.5 * (P + P.T)                                       # :106
    if v is None:
        v = np.zeros((N, nx))
    eta0 = np.asarray(g_batch(particles, u_km1, v), float).reshape(N, nx)   # :113-115
    eta, theta_log, cond = ledh_flow(eta0, P, np.asarray(R, float), z, h_batch, jac_batch,
                                     n_lambda_steps)
    logw = np.log(weights + 1e-300) + theta_log               # :186
    num = log_trans_batch(eta, particles) + log_like_batch(z, eta)
    den = log_trans_batch(eta0, particles)
    logw = logw + (num - den)                                 # :187-190
    logw = logw - np.max(logw)
    w = np.exp(logw)
    w = w / np.sum(w)                                         # :191-195
    info = {"w_pre": w, "eta0": eta0, "eta": eta, "theta_log": theta_log,
            "ancestors": None, "condition_numbers": cond}
    xk = eta
    if resample_ess_ratio > 0.0:                              # :201-206
        wn = w / np.sum(w)
        ess = 1.0 / float(np.sum(wn * wn))
        info["ess_pre"] = ess
        if ess < resample_ess_ratio * N:
            uu = rng.random()
            idx = ledh_systematic_ancestors(w, uu)
            info["ancestors"], info["u"] = idx, float(uu)
            xk = xk[idx]
            w = np.full_like(w, 1.0 / N)
    mean, cov = ledh_weighted_stats(xk, w)
    return xk, w, mean, cov, info
