"""NumPy restatement of the SIR particle-filter step -- TEST INFRASTRUCTURE ONLY.

Follows ``models/particle_filter.py`` of the reference operation by operation
(fp64, same NumPy calls wherever the reference is already vectorised); the
only changes are that the two per-particle list comprehensions (``:237``,
``:257``) take a batched callable and that the pure-Python two-pointer loop of
``_systematic_resample`` (``:164-170``) is replaced by the equivalent
``searchsorted(..., side='right')`` (equivalence is itself tested bit-for-bit
against the literal loop in ``systematic_ancestors_loop`` and against the
reference).

Parity status: pinned by tests/golden/sir_*.npz (generated from the unmodified
reference) and, in the build container, live against /root/reference.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

Array = np.ndarray


class TapeRNG:
    """Forwarding proxy around ``np.random.Generator`` that tapes every draw.

    The hot path only uses ``standard_normal``, ``random``, ``choice`` and
    ``multivariate_normal`` (reference ``particle_filter.py:128,160,186,217,236``,
    ``LEDH_particle_filter.py:28,87``).  ``tape`` holds ``(method, result)``
    tuples in call order so the same draws can be replayed into the CUDA engine
    ("reference-noise mode").
    """

    def __init__(self, rng: np.random.Generator):
        self._rng = rng
        self.tape: List[Tuple[str, object]] = []

    def standard_normal(self, *a, **k):
        r = self._rng.standard_normal(*a, **k)
        self.tape.append(("standard_normal", np.array(r, copy=True)))
        return r

    def random(self, *a, **k):
        r = self._rng.random(*a, **k)
        self.tape.append(("random", np.array(r, copy=True)))
        return r

    def choice(self, *a, **k):
        r = self._rng.choice(*a, **k)
        self.tape.append(("choice", np.array(r, copy=True)))
        return r

    def multivariate_normal(self, *a, **k):
        r = self._rng.multivariate_normal(*a, **k)
        self.tape.append(("multivariate_normal", np.array(r, copy=True)))
        return r

    def __getattr__(self, name):  # anything else: pass through untaped
        return getattr(self._rng, name)


def chol_with_jitter(M: Array, eps: float) -> Array:
    """``cholesky(M)`` with ONE retry at ``M + eps*I``.

    Reference: ``particle_filter.py:232-235`` (eps=1e-10, predict) and
    ``:213-216`` (eps=1e-12, post-resample jitter).
    """
    try:
        return np.linalg.cholesky(M)
    except np.linalg.LinAlgError:
        return np.linalg.cholesky(M + eps * np.eye(M.shape[0]))


def systematic_ancestors(weights: Array, u: float, *, clamp_last: bool = True) -> Array:
    """Systematic-resampling ancestors, vectorised.

    Reference ``particle_filter.py:146-171``: ``pos_i=(u+i)/N``; ``c=cumsum(w)``;
    ``c[-1]=1``; ``idx_i = #{j : c_j <= pos_i}``.  ``clamp_last=False`` gives the
    LEDH flavour (``LEDH_particle_filter.py:25-37``: no ``c[-1]=1``; the reference
    would raise IndexError when ``c[-1] <= pos[-1]``; we clip to N-1 instead,
    exactly like the device kernel's guard).
    """
    w = np.asarray(weights, dtype=np.float64)
    N = w.shape[0]
    positions = (u + np.arange(N)) / N
    c = np.cumsum(w)
    if clamp_last:
        c[-1] = 1.0
    idx = np.searchsorted(c, positions, side="right")
    np.minimum(idx, N - 1, out=idx)
    return idx.astype(np.int64, copy=False)


def systematic_ancestors_loop(weights: Array, u: float, *, clamp_last: bool = True) -> Array:
    """Literal two-pointer form (small N only; pure Python).

    Same control flow as reference ``particle_filter.py:164-170`` /
    ``LEDH_particle_filter.py:31-36``.
    """
    w = np.asarray(weights, dtype=np.float64)
    N = w.shape[0]
    positions = (u + np.arange(N)) / N
    out = np.zeros(N, dtype=np.int64)
    c = np.cumsum(w)
    if clamp_last:
        c[-1] = 1.0
    i = j = 0
    while i < N:
        if j >= N:  # only reachable with clamp_last=False; reference would IndexError
            out[i] = N - 1
            i += 1
        elif positions[i] < c[j]:
            out[i] = j
            i += 1
        else:
            j += 1
    return out


def multinomial_ancestors(weights: Array, rng) -> Array:
    """Reference ``particle_filter.py:173-186``: ``rng.choice(N, N, p=w)``."""
    n = len(weights)
    return rng.choice(n, size=n, p=weights)


def multinomial_ancestors_from_uniforms(weights: Array, u: Array) -> Array:
    """What ``rng.choice(N, size=len(u), p=w)`` computes from its uniforms ``u = rng.random(len(u))``
    (NumPy 2.3 ``Generator.choice``, replace=True with ``p``: ``cdf = p.cumsum(); cdf /= cdf[-1];
    cdf.searchsorted(uniform_samples, side='right')``); checked against ``rng.choice`` itself in
    tests/test_oracle_golden.py.  This is the checker of the device multinomial resampler."""
    cdf = np.cumsum(np.asarray(weights, float))
    cdf /= cdf[-1]
    return np.searchsorted(cdf, np.asarray(u, float), side="right").astype(np.int64)


def normalize_log_weights(logw: Array) -> Tuple[Array, float]:
    """Stable normalisation, reference ``particle_filter.py:262-263``.

    Returns ``(w, log_norm)`` with ``log_norm = m + log(sum(exp(logw - m)))``.
    """
    m = np.max(logw)
    log_norm = m + np.log(np.sum(np.exp(logw - m)))
    return np.exp(logw - log_norm), float(log_norm)


def sir_predict(particles: Array, g_batch: Callable, u, Q: Array, rng) -> Array:
    """Reference ``particle_filter.py:223-237`` with a batched ``g``."""
    N, nx = particles.shape
    Lq = chol_with_jitter(Q, 1e-10)
    noise = rng.standard_normal((N, nx)) @ Lq.T
    return np.asarray(g_batch(particles, u), dtype=np.float64).reshape(N, -1) + noise


def sir_update(
    particles: Array,
    weights: Array,
    z: Array,
    h_batch: Callable,
    LR: Array,
    *,
    Q: Array,
    resample_thresh: float,
    resample_method: str,
    regularize_after_resample: bool,
    rng,
):
    """Reference ``particle_filter.py:239-269`` + ``_resample`` ``:188-220``.

    Returns ``(particles, w, mean, cov, info)``; ``info`` holds the additive
    diagnostics the reference computes and throws away (pre-resample weights,
    ancestors, ESS, log normaliser).
    """
    N, nx = particles.shape
    z = np.asarray(z, float)
    z_pred = np.asarray(h_batch(particles), dtype=np.float64).reshape(N, -1)
    diffs = (z - z_pred).T
    y = np.linalg.solve(LR, diffs)
    quad = np.sum(y * y, axis=0)
    logw = np.log(weights + 1e-300) - 0.5 * quad
    w, log_norm = normalize_log_weights(logw)

    info = {"w_pre": w, "log_norm": log_norm, "ancestors": None, "u": None}
    Neff = 1.0 / np.sum(w ** 2)
    info["ess_pre"] = float(Neff)
    info["resampled"] = bool(Neff < resample_thresh * N)
    if info["resampled"]:
        if resample_method == "systematic":
            u = rng.random()
            idx = systematic_ancestors(w, u, clamp_last=True)
            info["u"] = float(u)
        else:
            idx = multinomial_ancestors(w, rng)
        info["ancestors"] = idx
        particles = particles[idx]
        w = np.ones_like(w) / len(w)
        if regularize_after_resample:
            Lq = chol_with_jitter(Q, 1e-12)
            particles = particles + rng.standard_normal(particles.shape) @ (0.001 * Lq.T)

    mean = np.average(particles, axis=0, weights=w)
    cov = np.atleast_2d(np.cov(particles.T, aweights=w, bias=True))
    return particles, w, mean, cov, info


class SIROracle:
    """Vectorised twin of the reference ``ParticleFilter`` class.

    ``g_batch(X, u) -> (N, nx)`` and ``h_batch(X) -> (N, nz)`` are batched model
    callables; ``per_particle(fn)`` adapts a reference-style per-particle
    callable with the literal list-comprehension semantics.
    """

    def __init__(self, g_batch, h_batch, Q, R, *, Np=1000, resample_thresh=0.5,
                 resample_method="systematic", regularize_after_resample=False, rng=None):
        self.g_batch, self.h_batch = g_batch, h_batch
        self.Q = np.asarray(Q, float)
        self.R = np.asarray(R, float)
        self.Np = int(Np)
        self.resample_thresh = float(resample_thresh)
        self.resample_method = resample_method
        self.regularize_after_resample = regularize_after_resample
        self.rng = np.random.default_rng() if rng is None else rng
        self.nx, self.nz = self.Q.shape[0], self.R.shape[0]
        self.LR = np.linalg.cholesky(self.R + 1e-12 * np.eye(self.nz))  # :107
        self.particles = self.weights = self.mean = self.cov = None
        self.t = 0
        self.info = {}

    @staticmethod
    def per_particle(fn, with_u: bool):
        if with_u:
            return lambda X, u: np.array([fn(x, u) for x in X])
        return lambda X: np.array([fn(x) for x in X])

    def initialize(self, mean, cov):
        """Reference ``particle_filter.py:110-132``."""
        mean = np.asarray(mean, float)
        cov = np.asarray(cov, float)
        Lc = np.linalg.cholesky(cov + 1e-10 * np.eye(len(mean)))
        self.particles = self.rng.standard_normal((self.Np, len(mean))) @ Lc.T + mean
        self.weights = np.ones(self.Np) / self.Np
        self.mean, self.cov, self.t = mean, np.atleast_2d(cov), 0

    def predict(self, u=None):
        self.particles = sir_predict(self.particles, self.g_batch, u, self.Q, self.rng)

    def update(self, z):
        self.particles, self.weights, self.mean, self.cov, self.info = sir_update(
            self.particles, self.weights, z, self.h_batch, self.LR, Q=self.Q,
            resample_thresh=self.resample_thresh, resample_method=self.resample_method,
            regularize_after_resample=self.regularize_after_resample, rng=self.rng)
        self.t += 1

    def step(self, z, u=None):
        self.predict(u)
        self.update(z)

    def effective_sample_size(self) -> float:
        return 1.0 / np.sum(self.weights ** 2)
