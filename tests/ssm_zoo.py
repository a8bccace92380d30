"""State-space models used by the parity tests, in reference-callback form.

Every model offers the reference's per-particle callables (``g(x,u)`` /
``g(x,u,v)``, ``h(x)``, ``jac_h(x)``, ``log_trans_pdf(xk,xkm1)``,
``log_like_pdf(z,x)`` -- SURVEY.md section 8b "operator API") and batched twins
for the NumPy oracle.  Pure NumPy; no engine imports.
"""

from __future__ import annotations

import numpy as np


def sv_data(n, alpha, sigma, beta, seed):
    """1-D stochastic-volatility series (x_t = a x + s v, y_t = b exp(x/2) w).

    Same model as the reference's ``simulate_sv_1d``
    (simulator/simulator_sto_volatility_model.py:51-122), independent code; used
    where the reference simulator is not available (GPU box, bench).
    """
    rng = np.random.default_rng(seed)
    x = np.empty(n)
    x[0] = rng.normal(0.0, np.sqrt(sigma ** 2 / (1 - alpha ** 2)))
    v = rng.standard_normal(n - 1)
    for t in range(1, n):
        x[t] = alpha * x[t - 1] + sigma * v[t - 1]
    y = beta * np.exp(0.5 * x) * rng.standard_normal(n)
    return x, y


class SVModel:
    """Additive-Gaussian SV form the reference tests/notebooks filter with:
    g = alpha x, h = beta exp(x/2) (tests/integration_tests/test_pf_vs_simulator_sv.py:50-56)."""

    nx = nz = 1

    def __init__(self, alpha=0.95, sigma=0.2, beta=1.0, r=None):
        self.alpha, self.sigma, self.beta = alpha, sigma, beta
        self.Q = np.array([[sigma ** 2]])
        self.R = np.array([[beta ** 2 if r is None else r]])

    def g(self, x, u=None):
        return np.array([self.alpha * x[0]])

    def h(self, x):
        return np.array([self.beta * np.exp(0.5 * x[0])])

    def g_batch(self, X, u=None):
        return self.alpha * X

    def h_batch(self, X):
        return self.beta * np.exp(0.5 * X)


class Lorenz96Model:
    """Lorenz-96 forecast model in the reference simulator's formulas
    (simulator/simulator_Lorenz_96.py: l96_rhs :35-58, rk4_step :61-83, ObsModel.H :147):
    g = n_sub RK4 steps, h = x[::obs_stride]; Q = q^2 I, R = r^2 I."""

    def __init__(self, nx=8, F=8.0, dt=0.01, n_sub=5, obs_stride=2, q=0.1, r=1.0):
        self.nx, self.F, self.dt, self.n_sub, self.obs_stride = nx, F, dt, n_sub, obs_stride
        self.H_idx = np.arange(0, nx, obs_stride)
        self.nz = self.H_idx.size
        self.Q = q ** 2 * np.eye(nx)
        self.R = r ** 2 * np.eye(self.nz)

    def rhs(self, x):
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

    g_batch = g
    h_batch = h

    def simulate(self, T, seed):
        """Truth trajectory observed every n_sub RK4 steps (standard initial pattern, short spin-up)."""
        rng = np.random.default_rng(seed)
        x = np.full(self.nx, self.F)
        x[::5] = self.F + 1.0
        for _ in range(200):
            x = self.rk4(x)
        X, Y = np.empty((T, self.nx)), np.empty((T, self.nz))
        r = np.sqrt(self.R[0, 0])
        for t in range(T):
            x = self.g(x)
            X[t] = x
            Y[t] = self.h(x) + r * rng.standard_normal(self.nz)
        return X, Y


class LGSSMModel:
    """README A/B/C/D linear-Gaussian model (reference README.md:104-113)."""

    def __init__(self):
        self.A = np.array([[0.9, 0.5], [0.0, 0.7]])
        self.B = np.diag([np.sqrt(0.05), np.sqrt(0.02)])
        self.C = np.eye(2)
        self.D = np.diag([np.sqrt(0.1), np.sqrt(0.1)])
        self.Q = self.B @ self.B.T
        self.R = self.D @ self.D.T
        self.nx = self.nz = 2

    def simulate(self, T, seed):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(2)
        X, Y = np.empty((T, 2)), np.empty((T, 2))
        for t in range(T):
            Y[t] = self.C @ x + self.D @ rng.standard_normal(2)
            X[t] = x
            x = self.A @ x + self.B @ rng.standard_normal(2)
        return X, Y

    def g(self, x, u=None):
        return self.A @ x

    def h(self, x):
        return self.C @ x

    def g_batch(self, X, u=None):
        return X @ self.A.T

    def h_batch(self, X):
        return X @ self.C.T


class AcousticCVModel:
    """Constant-velocity targets observed by an amplitude sensor grid.

    Joint state = ``n_targets`` x [x, y, vx, vy]; z_s = sum_targets psi / (|p - s|^2 + d0)
    (reference simulator/simulator_Multi_acoustic_tracking.py:273-310 and the
    single-target test model tests/integration_tests/test_filters_mat_simulator.py:44-66).
    """

    def __init__(self, n_targets=1, grid=(3, 3), area=(40.0, 40.0), psi=10.0, d0=0.1,
                 q_scale=1.0, r_var=0.01, Q1=None, sensors=None):
        self.n_targets = n_targets
        self.nx = 4 * n_targets
        gx = np.linspace(0.0, area[0], grid[0])
        gy = np.linspace(0.0, area[1], grid[1])
        self.S = np.array([[a, b] for a in gx for b in gy]) if sensors is None else np.asarray(sensors, float)
        self.nz = self.S.shape[0]
        self.psi, self.d0 = psi, d0
        F1 = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1.0]])
        if Q1 is None:
            Q1 = q_scale * np.array([[3.0, 0, 0.1, 0], [0, 3.0, 0, 0.1], [0.1, 0, 0.03, 0],
                                     [0, 0.1, 0, 0.03]])
        self.F = np.kron(np.eye(n_targets), F1)
        self.Q = np.kron(np.eye(n_targets), Q1)
        self.R = r_var * np.eye(self.nz)
        self._Qinv = np.linalg.inv(self.Q)
        self._Rinv = np.linalg.inv(self.R)
        self._logdet2piQ = np.linalg.slogdet(2 * np.pi * self.Q)[1]
        self._logdet2piR = np.linalg.slogdet(2 * np.pi * self.R)[1]

    # per-particle (reference callback signatures)
    def g(self, x, u=None, v=None):
        return self.F @ x + (0.0 if v is None else v)

    def g2(self, x, u=None):
        return self.F @ x

    def jac_g(self, x, u=None):
        return self.F

    def h(self, x):
        return self.h_batch(x[None, :])[0]

    def jac_h(self, x):
        return self.jac_batch(x[None, :])[0]

    def log_trans_pdf(self, xk, xkm1):
        d = xk - self.F @ xkm1
        return -0.5 * (d @ np.linalg.solve(self.Q, d) + self._logdet2piQ)

    def log_like_pdf(self, z, x):
        d = z - self.h(x)
        return -0.5 * (d @ np.linalg.solve(self.R, d) + self._logdet2piR)

    # batched (oracle)
    def g_batch(self, X, u=None, V=None):
        return X @ self.F.T + (0.0 if V is None else V)

    def h_batch(self, X):
        N = X.shape[0]
        pos = X.reshape(N, self.n_targets, 4)[:, :, :2]            # (N, C, 2)
        diff = pos[:, :, None, :] - self.S[None, None, :, :]       # (N, C, S, 2)
        d2 = np.sum(diff ** 2, axis=-1)
        return np.sum(self.psi / (d2 + self.d0), axis=1)           # (N, S)

    def jac_batch(self, X):
        N = X.shape[0]
        pos = X.reshape(N, self.n_targets, 4)[:, :, :2]
        diff = pos[:, :, None, :] - self.S[None, None, :, :]       # (N, C, S, 2)
        d2 = np.sum(diff ** 2, axis=-1)
        coef = -2.0 * self.psi / (d2 + self.d0) ** 2               # (N, C, S)
        H = np.zeros((N, self.nz, self.nx))
        for c in range(self.n_targets):
            H[:, :, 4 * c + 0] = coef[:, c, :] * diff[:, c, :, 0]
            H[:, :, 4 * c + 1] = coef[:, c, :] * diff[:, c, :, 1]
        return H

    def log_trans_batch(self, Xk, Xkm1):
        d = Xk - Xkm1 @ self.F.T
        return -0.5 * (np.einsum("ni,ij,nj->n", d, self._Qinv, d) + self._logdet2piQ)

    def log_like_batch(self, z, X):
        d = z[None, :] - self.h_batch(X)
        return -0.5 * (np.einsum("ni,ij,nj->n", d, self._Rinv, d) + self._logdet2piR)


def acoustic_c5_model():
    """BASELINE config 5's model shape: 4 targets (16-D joint state), 5 x 5 = 25 sensors, the article's process
    noise V (reference simulator_Multi_acoustic_tracking.py:104), R = 0.01 I."""
    V1 = (1.0 / 20.0) * np.array([[1 / 3, 0, 0.5, 0], [0, 1 / 3, 0, 0.5], [0.5, 0, 1, 0], [0, 0.5, 0, 1.0]])
    return AcousticCVModel(n_targets=4, grid=(5, 5), Q1=V1)


class Linear1DModel:
    """1-D linear system of the LEDH unit tests (tests/unit_tests/models/test_ledh_flow_pf.py:63-113)."""

    nx = nz = 1

    def __init__(self, alpha=0.9, sigma=0.2, r=0.1):
        self.alpha = alpha
        self.Q = np.array([[sigma ** 2]])
        self.R = np.array([[r]])

    def g(self, x, u=None, v=None):
        return np.array([self.alpha * x[0] + (0.0 if v is None else v[0])])

    def g2(self, x, u=None):
        return np.array([self.alpha * x[0]])

    def jac_g(self, x, u=None):
        return np.array([[self.alpha]])

    def h(self, x):
        return np.array([x[0]])

    def jac_h(self, x):
        return np.array([[1.0]])

    def log_trans_pdf(self, xk, xkm1):
        d = xk - self.g2(xkm1)
        return float(-0.5 * (d @ np.linalg.solve(self.Q, d) + np.log(2 * np.pi * self.Q[0, 0])))

    def log_like_pdf(self, z, x):
        d = z - self.h(x)
        return float(-0.5 * (d @ np.linalg.solve(self.R, d) + np.log(2 * np.pi * self.R[0, 0])))

    def g_batch(self, X, u=None, V=None):
        return self.alpha * X + (0.0 if V is None else V)

    def h_batch(self, X):
        return X.copy()

    def jac_batch(self, X):
        return np.ones((X.shape[0], 1, 1))

    def log_trans_batch(self, Xk, Xkm1):
        d = (Xk - self.alpha * Xkm1)[:, 0]
        return -0.5 * (d * d / self.Q[0, 0] + np.log(2 * np.pi * self.Q[0, 0]))

    def log_like_batch(self, z, X):
        d = z[0] - X[:, 0]
        return -0.5 * (d * d / self.R[0, 0] + np.log(2 * np.pi * self.R[0, 0]))
