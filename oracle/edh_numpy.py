"""NumPy restatement of one EDH particle-flow step -- TEST INFRASTRUCTURE ONLY.

Follows ``models/EDH_particle_filter.py:182-318`` of the reference: the flow field is evaluated at the
tracker's mean path and every particle is advanced by the same explicit Euler (:271-275) or RK4
(:29-35, :277-280) sub-steps -- restated here as the reference runs them (step by step, batched over
particles), NOT as the composed affine map the engine applies, so that the two are independent.
Pinned by tests/golden/edh_*.npz (outputs of the unmodified reference).
"""

from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .ledh_numpy import ledh_systematic_ancestors, ledh_weighted_stats

Array = np.ndarray


def _rk4(x, f, dt):
    k1 = f(x)
    k2 = f(x + 0.5 * dt * k1)
    k3 = f(x + 0.5 * dt * k2)
    k4 = f(x + dt * k3)
    return x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)


def edh_step(particles: Array, weights: Array, z: Array, P: Array, R: Array, etabar0: Array, *,
             g_batch: Callable, h: Callable, jac_h: Callable, log_trans_batch: Callable, log_like_batch: Callable,
             n_lambda_steps: int, integrator: str, ess_ratio: float, u: Optional[float], V: Optional[Array] = None):
    """Returns dict(particles, weights, mean, cov, cond, ess, resampled, ancestors)."""
    N, nx = particles.shape
    P = 0.5 * (P + P.T)
    eta0 = g_batch(particles, None, V)
    eta = eta0.copy()
    etabar = np.asarray(etabar0, float).copy()
    L = max(1, int(n_lambda_steps))
    dlam, lam, I = 1.0 / float(L), 0.0, np.eye(nx)
    cond = []
    for _ in range(L):
        lam = min(1.0, lam + dlam)
        H = jac_h(etabar)
        e = h(etabar) - H @ etabar
        S = lam * H @ P @ H.T + R
        cond.append(float(np.linalg.cond(S)))
        A = -0.5 * P @ H.T @ np.linalg.solve(S, H)
        b = (I + 2.0 * lam * A) @ ((I + lam * A) @ (P @ H.T @ np.linalg.solve(R, z - e)) + A @ etabar)
        if integrator.lower() == "euler":
            eta = eta + dlam * (eta @ A.T + b)
            etabar = etabar + dlam * (A @ etabar + b)
        else:
            eta = _rk4(eta, lambda X: X @ A.T + b, dlam)
            etabar = _rk4(etabar, lambda x: A @ x + b, dlam)
    logw = np.log(weights + 1e-300) + log_trans_batch(eta, particles) + log_like_batch(z, eta) \
        - log_trans_batch(eta0, particles)
    logw -= np.max(logw)
    w = np.exp(logw)
    w /= np.sum(w)
    ess = 1.0 / float(np.sum((w / np.sum(w)) ** 2))
    anc, resampled, xk = None, False, eta
    if ess_ratio > 0.0 and ess < ess_ratio * N:
        anc = ledh_systematic_ancestors(w, u)
        xk, w, resampled = eta[anc], np.full(N, 1.0 / N), True
    mean, cov = ledh_weighted_stats(xk, w)
    return dict(particles=xk, weights=w, mean=mean, cov=cov, cond=cond, ess=ess, resampled=resampled, ancestors=anc)
