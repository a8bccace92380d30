"""State-space-model specs: reference-style callables that also carry a device functor.

The reference's plugin API is a pair of per-particle Python callables
``g(x, u) -> (nx,)`` and ``h(x) -> (nz,)`` (models/particle_filter.py:21-22,
:237, :257).  A spec object offers exactly those callables (so it can be passed
to the reference as well) plus ``device_functor(Q, R)``, which names the
compiled functor of the fused propagate+log-likelihood kernel (K1) and packs its
parameters.  ``ParticleFilter(g=spec.g, h=spec.h, ...)`` detects the spec through
the bound methods and runs fully on the GPU; any other callable goes through
the host-callback bridge.
"""

from __future__ import annotations

import numpy as np

from . import _lib


def _chol_jitter(M, eps):
    try:
        return np.linalg.cholesky(M)
    except np.linalg.LinAlgError:
        return np.linalg.cholesky(M + eps * np.eye(M.shape[0]))


class _Spec:
    smc_model_id = 0

    def device_functor(self, Q, R):
        raise NotImplementedError


class StochVolSSM(_Spec):
    """1-D stochastic volatility in the additive-Gaussian form the reference's tests
    and notebooks filter with: ``g = alpha x``, ``h = beta exp(x / 2)``
    (tests/integration_tests/test_pf_vs_simulator_sv.py:50-56)."""

    smc_model_id = _lib.SMC_MODEL_STOCH_VOL
    nx = nz = 1

    def __init__(self, alpha: float, beta: float):
        self.alpha, self.beta = float(alpha), float(beta)

    def g(self, x, u=None):
        return np.array([self.alpha * x[0]])

    def h(self, x):
        return np.array([self.beta * np.exp(0.5 * x[0])])

    def device_functor(self, Q, R):
        Lq = _chol_jitter(np.asarray(Q, float), 1e-10)                      # particle_filter.py:232-235
        LR = np.linalg.cholesky(np.asarray(R, float) + 1e-12 * np.eye(1))   # :107
        return self.smc_model_id, np.array([self.alpha, Lq[0, 0], self.beta, LR[0, 0]])


class LinearGaussianSSM(_Spec):
    """``g = A x``, ``h = H x`` (README LGSSM, reference README.md:104-113)."""

    smc_model_id = _lib.SMC_MODEL_LINEAR_GAUSSIAN

    def __init__(self, A, H):
        self.A = np.atleast_2d(np.asarray(A, float))
        self.H = np.atleast_2d(np.asarray(H, float))
        self.nx, self.nz = self.A.shape[0], self.H.shape[0]

    def g(self, x, u=None):
        return self.A @ x

    def h(self, x):
        return self.H @ x

    def device_functor(self, Q, R):
        Lq = _chol_jitter(np.asarray(Q, float), 1e-10)
        LR = np.linalg.cholesky(np.asarray(R, float) + 1e-12 * np.eye(self.nz))
        LRinv = np.linalg.solve(LR, np.eye(self.nz))
        return self.smc_model_id, np.concatenate(
            [self.A.ravel(), np.tril(Lq).ravel(), self.H.ravel(), np.tril(LRinv).ravel()])


class Lorenz96SSM(_Spec):
    """Lorenz-96 forecast model of the reference's simulator: ``g`` = ``n_sub`` RK4 steps of
    ``dx_a/dt = (x_{a+1} - x_{a-2}) x_{a-1} - x_a + F`` (simulator/simulator_Lorenz_96.py:
    ``l96_rhs`` :35-58, ``rk4_step`` :61-83, same operation order), ``h(x) = x[::obs_stride]``
    (``ObsModel.H`` with ``H_idx = arange(0, nx, obs_fraction)``, :386).  Q and R must be diagonal."""

    smc_model_id = _lib.SMC_MODEL_LORENZ96

    def __init__(self, nx: int = 40, F: float = 8.0, dt: float = 0.01, n_sub: int = 20, obs_stride: int = 4):
        self.nx, self.F, self.dt, self.n_sub, self.obs_stride = int(nx), float(F), float(dt), int(n_sub), int(obs_stride)
        self.H_idx = np.arange(0, self.nx, self.obs_stride)
        self.nz = self.H_idx.size

    def rhs(self, x):
        x = np.asarray(x)
        return (np.roll(x, -1, axis=-1) - np.roll(x, 2, axis=-1)) * np.roll(x, 1, axis=-1) - x + self.F

    def rk4(self, x):
        dt = self.dt
        k1 = self.rhs(x)
        k2 = self.rhs(x + 0.5 * dt * k1)
        k3 = self.rhs(x + 0.5 * dt * k2)
        k4 = self.rhs(x + dt * k3)
        return x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)

    def g(self, x, u=None):
        x = np.asarray(x, float)
        for _ in range(self.n_sub):
            x = self.rk4(x)
        return x

    def h(self, x):
        return np.asarray(x)[..., self.H_idx]

    def device_functor(self, Q, R):
        Q, R = np.asarray(Q, float), np.asarray(R, float)
        if np.any(Q - np.diag(np.diag(Q))) or np.any(R - np.diag(np.diag(R))):
            raise ValueError("Lorenz96SSM needs diagonal Q and R")
        if Q.shape[0] != self.nx or R.shape[0] != self.nz:
            raise ValueError("Q / R shapes do not match the Lorenz-96 spec")
        Lq = _chol_jitter(Q, 1e-10)
        LR = np.linalg.cholesky(R + 1e-12 * np.eye(self.nz))
        return self.smc_model_id, np.concatenate(
            [[self.F, self.dt, self.n_sub, self.obs_stride, self.nz], np.diag(Lq), np.diag(LR)])


def spec_of(g, h):
    """Return the spec object if ``g`` and ``h`` are bound methods of one spec."""
    sg, sh = getattr(g, "__self__", None), getattr(h, "__self__", None)
    if isinstance(sg, _Spec) and sg is sh and getattr(g, "__func__", None) is type(sg).g \
            and getattr(h, "__func__", None) is type(sg).h:
        return sg
    return None


# ---------------------------------------------------------------------------------------------
# measurement specs for the LEDH flow kernel (h and its Jacobian evaluated on chip)
# ---------------------------------------------------------------------------------------------
class _MeasurementSpec:
    def ledh_functor(self):
        raise NotImplementedError


class LinearMeasurement(_MeasurementSpec):
    """``h(x) = H x``, ``jacobian_h(x) = H``."""

    def __init__(self, H):
        self.H = np.atleast_2d(np.asarray(H, float))

    def h(self, x):
        return self.H @ x

    def jacobian_h(self, x):
        return self.H

    def ledh_functor(self):
        return _lib.SMC_LEDH_H_LINEAR, self.H.ravel()


class AcousticMeasurement(_MeasurementSpec):
    """Amplitude sensors over constant-velocity targets: state = n_targets x [x, y, vx, vy],
    ``z_s = sum_c psi / (|p_c - s|^2 + d0)`` (reference
    simulator/simulator_Multi_acoustic_tracking.py:273-310; single-target form in
    tests/integration_tests/test_filters_mat_simulator.py:44-66)."""

    def __init__(self, sensors, psi: float, d0: float, n_targets: int = 1):
        self.S = np.asarray(sensors, float).reshape(-1, 2)
        self.psi, self.d0, self.n_targets = float(psi), float(d0), int(n_targets)

    def _parts(self, x):
        pos = np.asarray(x, float).reshape(self.n_targets, 4)[:, :2]
        diff = pos[:, None, :] - self.S[None, :, :]
        den = np.sum(diff ** 2, axis=-1) + self.d0
        return diff, den

    def h(self, x):
        _, den = self._parts(x)
        return np.sum(self.psi / den, axis=0)

    def jacobian_h(self, x):
        diff, den = self._parts(x)
        coef = -2.0 * self.psi / den ** 2
        H = np.zeros((self.S.shape[0], 4 * self.n_targets))
        for c in range(self.n_targets):
            H[:, 4 * c] = coef[c] * diff[c, :, 0]
            H[:, 4 * c + 1] = coef[c] * diff[c, :, 1]
        return H

    def ledh_functor(self):
        return _lib.SMC_LEDH_H_ACOUSTIC, np.concatenate([[self.psi, self.d0], self.S.ravel()])


class ExpMeasurement(_MeasurementSpec):
    """``h_i(x) = beta exp(x_i / 2)`` (stochastic volatility, additive-noise form)."""

    def __init__(self, beta: float):
        self.beta = float(beta)

    def h(self, x):
        return self.beta * np.exp(0.5 * np.asarray(x, float))

    def jacobian_h(self, x):
        return np.diag(0.5 * self.beta * np.exp(0.5 * np.asarray(x, float)))

    def ledh_functor(self):
        return _lib.SMC_LEDH_H_EXP, np.array([self.beta])


def measurement_spec_of(h, jacobian_h):
    sh, sj = getattr(h, "__self__", None), getattr(jacobian_h, "__self__", None)
    if isinstance(sh, _MeasurementSpec) and sh is sj and getattr(h, "__func__", None) is type(sh).h \
            and getattr(jacobian_h, "__func__", None) is type(sh).jacobian_h:
        return sh
    return None


class LedhGaussianSSM:
    """Linear-Gaussian dynamics ``x_k = F x_{k-1} + v, v ~ N(0, Q)`` with a measurement spec and
    Gaussian observation noise ``R``: supplies all five LEDH callbacks in the reference's
    signatures (``g(x,u,v)``, ``h``, ``jacobian_h``, ``log_trans_pdf``, ``log_like_pdf``,
    LEDH_particle_filter.py:18-22).  When all five callables handed to ``LEDHFlowPF`` are the
    bound methods of one such object the whole step runs on the GPU."""

    def __init__(self, F, Q, measurement: _MeasurementSpec, R):
        self.F = np.atleast_2d(np.asarray(F, float))
        self.Q = np.atleast_2d(np.asarray(Q, float))
        self.R = np.atleast_2d(np.asarray(R, float))
        self.meas = measurement
        self.nx, self.nz = self.F.shape[0], self.R.shape[0]
        self.Lq = np.linalg.cholesky(self.Q)
        self.Qinv = np.linalg.inv(self.Q)
        self.Rinv = np.linalg.inv(self.R)
        self._cq = np.linalg.slogdet(2 * np.pi * self.Q)[1]
        self._cr = np.linalg.slogdet(2 * np.pi * self.R)[1]

    def g(self, x, u=None, v=None):
        return self.F @ x + (0.0 if v is None else v)

    def h(self, x):
        return self.meas.h(x)

    def jacobian_h(self, x):
        return self.meas.jacobian_h(x)

    def log_trans_pdf(self, xk, xkm1):
        d = xk - self.F @ xkm1
        return float(-0.5 * (d @ self.Qinv @ d + self._cq))

    def log_like_pdf(self, z, x):
        d = z - self.meas.h(x)
        return float(-0.5 * (d @ self.Rinv @ d + self._cr))

    def sampler(self, rng):
        """A ``process_noise_sampler(N, nx)`` drawing N(0, Q) from ``rng`` (host)."""
        return lambda n, nx: rng.standard_normal((n, nx)) @ self.Lq.T


def ledh_spec_of(g, h, jacobian_h, log_trans_pdf, log_like_pdf):
    owners = [getattr(f, "__self__", None) for f in (g, h, jacobian_h, log_trans_pdf, log_like_pdf)]
    s = owners[0]
    if isinstance(s, LedhGaussianSSM) and all(o is s for o in owners):
        names = ("g", "h", "jacobian_h", "log_trans_pdf", "log_like_pdf")
        fns = (g, h, jacobian_h, log_trans_pdf, log_like_pdf)
        if all(getattr(f, "__func__", None) is getattr(type(s), n) for f, n in zip(fns, names)):
            return s
    return None
