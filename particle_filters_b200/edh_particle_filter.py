"""Drop-in EDH particle-flow particle filter on the B200 engine.

Mirrors the public surface of the reference's ``models/EDH_particle_filter.py``
(``GaussianTracker`` :13-26, ``rk4_step`` :29-35, ``systematic_resample`` :37-51,
``effective_sample_size`` :53-56, ``EDHConfig`` :59-66, ``PFState`` :68-75, ``EKFTracker`` /
``UKFTracker`` :77-132, ``EDHFlowPF`` :135-328).

In the exact-Daum-Huang flow the field ``A(lambda) x + b(lambda)`` is evaluated at the tracker's
mean path, not per particle (:218-258), so the pseudo-time integration of ALL particles is one
affine map: the nx x nx algebra of the L sub-steps stays with the caller's tracker on the host
(as in the reference), is composed into ``x <- M x + c``, and a streaming CUDA kernel applies it
(``smc_affine_map``).  Weights, ESS, systematic resampling and moments run on the same kernels as
the LEDH step.  The user's ``g`` / log-density callables are evaluated on the host exactly as the
reference does (host-callback bridge).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Protocol, Tuple

import numpy as np

from . import _lib
from .ledh_particle_filter import (LEDHFlowPF, PFState, _weights_engine, effective_sample_size,  # noqa: F401
                                   systematic_resample)

Array = np.ndarray


class GaussianTracker(Protocol):
    def predict(self) -> Tuple[Array, Array]: ...
    def update(self, z_k: Array) -> Tuple[Array, Array]: ...
    def get_past_mean(self) -> Array: ...


GFn = Callable[[Array, Optional[Array], Optional[Array]], Array]
HFn = Callable[[Array], Array]
JacobianHFn = Callable[[Array], Array]
LogTransPdf = Callable[[Array, Array], float]
LogLikePdf = Callable[[Array, Array], float]


def rk4_step(x: Array, f: Callable[[Array], Array], dt: float) -> Array:
    """One RK4 step for x' = f(x) (reference :29-35)."""
    k1 = f(x)
    k2 = f(x + 0.5 * dt * k1)
    k3 = f(x + 0.5 * dt * k2)
    k4 = f(x + dt * k3)
    return x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)


@dataclass
class EDHConfig:
    n_particles: int = 512
    n_lambda_steps: int = 8
    resample_ess_ratio: float = 0.5
    flow_integrator: str = "rk4"
    rng: np.random.Generator = np.random.default_rng(0)


class _Tracker:
    """Shared body of the reference's EKF / UKF wrappers (:77-132)."""

    def __init__(self, filt, initial_state):
        self._f = filt
        self.state = initial_state
        self.past_mean = initial_state.mean.copy()

    def predict(self) -> Tuple[Array, Array]:
        self.past_mean = self.state.mean.copy()
        self.state = self._f.predict(self.state, u=None)
        return self.state.mean, self.state.cov

    def update(self, z_k: Array) -> Tuple[Array, Array]:
        self.state = self._f.update(self.state, z_k)
        return self.state.mean, self.state.cov

    def get_past_mean(self) -> Array:
        return self.past_mean


class EKFTracker(_Tracker):
    def __init__(self, ekf, initial_state):
        super().__init__(ekf, initial_state)
        self.ekf = ekf


class UKFTracker(_Tracker):
    def __init__(self, ukf, initial_state):
        super().__init__(ukf, initial_state)
        self.ukf = ukf


def affine_step(A: Array, b: Array, dlam: float, integrator: str) -> Tuple[Array, Array]:
    """One pseudo-time sub-step of x' = A x + b as an affine map x <- Phi x + psi: explicit Euler
    (:271-275) or the classical RK4 of :29-35 applied to an affine field (exact polynomial in dlam A)."""
    n = A.shape[0]
    I = np.eye(n)
    if integrator.lower() == "euler":
        return I + dlam * A, dlam * b
    hA = dlam * A
    hA2 = hA @ hA
    hA3 = hA2 @ hA
    Phi = I + hA + hA2 / 2.0 + hA3 / 6.0 + (hA3 @ hA) / 24.0
    psi = dlam * (I + hA / 2.0 + hA2 / 6.0 + hA3 / 24.0) @ b
    return Phi, psi


class EDHFlowPF:
    """EKF/UKF-assisted EDH particle-flow particle filter (reference :135-328)."""

    def __init__(self, tracker, g, h, jacobian_h, log_trans_pdf, log_like_pdf, R, config: Optional[EDHConfig] = None,
                 *, device=None) -> None:
        self.tracker = tracker
        self.g = g
        self.h = h
        self.Jh = jacobian_h
        self.log_trans_pdf = log_trans_pdf
        self.log_like_pdf = log_like_pdf
        self.R = np.array(R, dtype=float)
        self.cfg = config or EDHConfig()
        self._device = device
        self.last_ess = None
        self.last_resampled = None

    def init_from_gaussian(self, mean0: Array, cov0: Array) -> PFState:
        """Sample initial particles from N(mean0, cov0) with equal weights (:172-180)."""
        n, nx = self.cfg.n_particles, mean0.size
        eps = self.cfg.rng.multivariate_normal(np.zeros(nx), cov0, size=n)
        particles = mean0[None, :] + eps
        weights = np.full(n, 1.0 / n)
        mean, cov = self._weighted_stats(particles, weights)
        return PFState(particles=particles, weights=weights, mean=mean, cov=cov, diagnostics={})

    def flow_map(self, P: Array, etabar: Array, z_k: Array):
        """Host part of the flow (:218-275): per sub-step A, b at the mean path; returns the composed
        affine map (M, c), the flowed mean path and the condition numbers of S."""
        nx = etabar.size
        n_steps = max(1, int(self.cfg.n_lambda_steps))
        dlam = 1.0 / float(n_steps)
        lam = 0.0
        I = np.eye(nx)
        M, c = np.eye(nx), np.zeros(nx)
        cond_numbers = []
        for _ in range(n_steps):
            lam = min(1.0, lam + dlam)
            H = self.Jh(etabar)
            h_bar = self.h(etabar)
            e = h_bar - H @ etabar
            S = lam * H @ P @ H.T + self.R
            try:
                cond_numbers.append(float(np.linalg.cond(S)))
            except Exception:
                cond_numbers.append(np.nan)
            try:
                S_inv_H = np.linalg.solve(S, H)
            except np.linalg.LinAlgError:
                S = S + 1e-8 * np.eye(S.shape[0])
                S_inv_H = np.linalg.solve(S, H)
            A = -0.5 * P @ H.T @ S_inv_H
            try:
                R_inv_innov = np.linalg.solve(self.R, (z_k - e))
            except np.linalg.LinAlgError:
                R_inv_innov = np.linalg.inv(self.R + 1e-8 * np.eye(self.R.shape[0])) @ (z_k - e)
            PHt_Rinv_innov = P @ H.T @ R_inv_innov
            b = (I + 2.0 * lam * A) @ ((I + lam * A) @ PHt_Rinv_innov + A @ etabar)
            Phi, psi = affine_step(A, b, dlam, self.cfg.flow_integrator)
            M, c = Phi @ M, Phi @ c + psi
            if self.cfg.flow_integrator.lower() == "euler":
                etabar = etabar + dlam * (A @ etabar + b)
            else:
                etabar = rk4_step(etabar, lambda vec: A @ vec + b, dlam)
        return M, c, etabar, cond_numbers

    def step(self, state: PFState, z_k: Array, u_km1: Optional[Array] = None,
             process_noise_sampler: Optional[Callable[[int, int], Array]] = None) -> PFState:
        """One EDH-PF step (reference :182-318)."""
        particles = np.asarray(state.particles, float)
        N, nx = particles.shape
        z_k = np.asarray(z_k, float).reshape(-1)
        nz = self.R.shape[0]
        lib = _lib.load()

        m_pred, P = self.tracker.predict()
        P = 0.5 * (P + P.T)
        v = np.zeros((N, nx)) if process_noise_sampler is None else process_noise_sampler(N, nx)
        eta0 = np.empty_like(particles)
        for i in range(N):                                        # user callbacks on the host (:204-206)
            eta0[i] = self.g(particles[i], u_km1, v[i])
        etabar = self.g(self.tracker.get_past_mean(), u_km1, np.zeros(nx))
        M, c, _, cond_numbers = self.flow_map(P, np.asarray(etabar, float), z_k)

        # the flow of all particles: one affine map on the GPU
        eng = _weights_engine(N, nx, nz, self._device)
        torch = eng.torch
        eng.set_particles(eta0, shift=np.asarray(state.mean, float).reshape(-1)[:nx])
        kM, pM = _lib.as_doubles(M)
        kc, pc = _lib.as_doubles(c)
        _lib.check(lib.smc_affine_map(C.c_void_p(eng.x[0].data_ptr()), pM, pc, N, eng.ld, nx, eng.dtype_code, eng.stream))
        xk = eng.get_particles()

        # weights (:287-297): user log-densities on the host, everything else on the GPU
        terms = np.empty(N)
        for i in range(N):
            terms[i] = (self.log_trans_pdf(xk[i], particles[i]) + self.log_like_pdf(z_k, xk[i])
                        - self.log_trans_pdf(eta0[i], particles[i]))
        wprev = torch.from_numpy(np.asarray(state.weights, float).reshape(-1).copy()).to(eng.device)
        theta = torch.zeros((eng.ld,), dtype=eng.tdtype, device=eng.device)
        td = torch.from_numpy(terms).to(eng.device)
        _lib.check(lib.smc_ledh_weights(eng._d, C.c_void_p(wprev.data_ptr()), C.c_void_p(theta.data_ptr()),
                                        C.c_void_p(td.data_ptr()), eng.stream))
        ratio = float(self.cfg.resample_ess_ratio)
        eng.finalize(ratio if ratio > 0.0 else 0.0, u=0.0)
        row = eng.read_row()
        self.last_ess, self.last_resampled = float(row[1]), bool(row[2])

        self.tracker.update(z_k)                                  # :300

        if self.last_resampled:                                   # :303-308
            eng.set_uniform(float(self.cfg.rng.random()))
            eng.resample()
            eng.finish_step()
            xk = eng.get_particles()
            w = np.full(N, 1.0 / N)
        else:
            eng.finish_step()
            w = eng.weights()
        row = eng.read_row()
        H0 = _lib.SMC_STATS_HEADER
        mean = row[H0:H0 + nx].copy()
        cov = row[H0 + nx:].reshape(nx, nx).copy()
        cov = 0.5 * (cov + cov.T)
        return PFState(particles=xk, weights=w, mean=mean, cov=cov, diagnostics={"condition_numbers": cond_numbers})

    _weighted_stats = staticmethod(LEDHFlowPF._weighted_stats)
