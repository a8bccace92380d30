"""Drop-in LEDH particle-flow particle filter on the B200 engine.

Mirrors the public surface of the reference's ``models/LEDH_particle_filter.py``
(``GaussianTracker`` :13-16, ``systematic_resample`` :25-37, ``effective_sample_size``
:39-41, ``LEDHConfig`` :44-49, ``PFState`` :51-57, ``LEDHFlowPF`` :60-224).  The
per-particle pseudo-time flow (:132-179) runs in the batched small-matrix CUDA kernel
K4 (one warp per particle); weights, ESS, systematic resampling and the weighted
moments reuse the SIR kernels.  The tracker (EKF/UKF) stays the caller's host object,
exactly as in the reference.

Measurement model on the GPU: if ``h`` and ``jacobian_h`` are the bound methods of a
measurement spec (:mod:`particle_filters_b200.ssm`), all L pseudo-time steps run in one
launch; otherwise the opaque callables are evaluated on the host per pseudo-time step
and the per-particle Jacobians uploaded (host-callback bridge).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Protocol, Tuple

import numpy as np

from . import _lib
from .ssm import ledh_spec_of, measurement_spec_of

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


def _weights_engine(n: int, nx: int, nz: int, device=None):
    from .engine import SirEngine

    key = (n, nx, nz, str(device))
    eng = _weights_engine.cache.get(key)
    if eng is None:
        eng = SirEngine(n, nx, nz, dtype="f64", device=device, write_ancestors=True, generic_moments=True)
        _weights_engine.cache[key] = eng
    return eng


_weights_engine.cache = {}


def _normalised_logw(eng, weights: Array):
    """Load ``log(w + 1e-300)`` of host weights into the engine and reduce (max, sum, sum^2)."""
    n = eng.n
    eng.reset()
    eng.bridge_loglik(np.zeros((n, eng.nz)), np.zeros(eng.nz), np.eye(eng.nz), wprev=np.asarray(weights, float))


def systematic_resample(weights: Array, rng: np.random.Generator) -> Array:
    """Systematic resampling of ``weights / sum(weights)`` (reference :25-37) on the GPU:
    the normalisation is the log-sum-exp kernel, the search the K3 resampler (whose guard
    replaces the reference's IndexError when round-off leaves ``cdf[-1] < pos[-1]``)."""
    w = np.asarray(weights, float).reshape(-1)
    eng = _weights_engine(w.size, 1, 1)
    _normalised_logw(eng, w)
    eng.finalize(2.0, u=float(rng.random()))     # threshold > 1: always resample
    eng.resample()
    eng.sync()
    return eng.ancestors[: w.size].cpu().numpy().astype(np.int64)


def effective_sample_size(weights: Array) -> float:
    """``1 / sum((w / sum w)^2)`` (reference :39-41), reduced on the GPU."""
    w = np.asarray(weights, float).reshape(-1)
    eng = _weights_engine(w.size, 1, 1)
    _normalised_logw(eng, w)
    eng.finalize(0.0, u=0.0)
    return float(eng.read_row()[1])


@dataclass
class LEDHConfig:
    n_particles: int = 512
    n_lambda_steps: int = 8
    resample_ess_ratio: float = 0.0
    rng: np.random.Generator = np.random.default_rng(0)


@dataclass
class PFState:
    particles: Array
    weights: Array
    mean: Array
    cov: Array
    diagnostics: dict = None


class DeviceLedhState(PFState):
    """``PFState`` of a device-resident LEDH run: particles / weights are fetched on first access.

    The arrays live in the owner's engine until then, so a state is only readable while it is the owner's
    LATEST one: once the filter has stepped again, an unread state raises instead of silently returning the
    newer step's particles (the reference returns independent arrays per state -- read ``.particles`` /
    ``.weights`` before the next ``step`` to keep them)."""

    def __init__(self, owner, mean, cov, diagnostics):
        self._owner, self._gen = owner, owner._generation
        self._particles = self._weights = None
        self.mean, self.cov, self.diagnostics = mean, cov, diagnostics

    def _is_current(self) -> bool:
        return self._gen == self._owner._generation

    def _check(self):
        if not self._is_current():
            raise RuntimeError("stale LEDH PFState: the filter has stepped since this state was returned; read "
                               ".particles/.weights before the next step() to keep them")

    @property
    def particles(self):
        if self._particles is None:
            self._check()
            self._particles = self._owner._dev_eng.get_particles()
        return self._particles

    @property
    def weights(self):
        if self._weights is None:
            self._check()
            self._weights = self._owner._dev_eng.weights()
        return self._weights


class _FlowBuffers:
    """Device arrays of one (N, nx, nz) flow problem."""

    def __init__(self, n, nx, nz, L, dtype, device):
        from .engine import _round4, _torch

        torch = _torch()
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.nx, self.nz, self.L = n, nx, nz, L
        self.ld = _round4(n)
        self.tdtype = torch.float32 if dtype == "f32" else torch.float64
        self.npdtype = np.float32 if dtype == "f32" else np.float64
        self.eta0 = torch.zeros((nx, self.ld), dtype=self.tdtype, device=self.device)
        self.eta = torch.zeros((nx, self.ld), dtype=self.tdtype, device=self.device)
        self.theta = torch.zeros((self.ld,), dtype=self.tdtype, device=self.device)
        self.s0 = torch.zeros((L, nz, nz), dtype=torch.float64, device=self.device)
        self.desc = _lib.LedhDesc(
            n=n, ld=self.ld, nx=nx, nz=nz, dtype=_lib.SMC_F32 if dtype == "f32" else _lib.SMC_F64,
            n_lambda_steps=L, eta0_dev=self.eta0.data_ptr(), eta_dev=self.eta.data_ptr(),
            theta_log_dev=self.theta.data_ptr(), s0_dev=self.s0.data_ptr())

    @property
    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def load(self, eta0: Array):
        t = self.torch.from_numpy(np.ascontiguousarray(eta0.T).astype(self.npdtype, copy=False))
        self.eta0[:, : self.n].copy_(t)
        self.eta[:, : self.n].copy_(t)
        self.theta.zero_()

    def eta_host(self) -> Array:
        return self.eta[:, : self.n].to("cpu", dtype=self.torch.float64).numpy().T.copy()


class LEDHFlowPF:
    """EKF/UKF-assisted LEDH particle-flow particle filter (reference Algorithm 1)."""

    def __init__(
        self,
        tracker: GaussianTracker,
        g: GFn,
        h: HFn,
        jacobian_h: JacobianHFn,
        log_trans_pdf: LogTransPdf,
        log_like_pdf: LogLikePdf,
        R: Array,
        config: Optional[LEDHConfig] = None,
        *,
        dtype: str = "f64",
        device=None,
    ) -> None:
        self.tracker = tracker
        self.g = g
        self.h = h
        self.Jh = jacobian_h
        self.log_trans_pdf = log_trans_pdf
        self.log_like_pdf = log_like_pdf
        self.R = np.array(R, dtype=float)
        self.cfg = config or LEDHConfig()
        self._dtype, self._device = dtype, device
        self._meas = measurement_spec_of(h, jacobian_h)
        self._spec = ledh_spec_of(g, h, jacobian_h, log_trans_pdf, log_like_pdf)
        if self._spec is not None:
            self._meas = self._spec.meas
            if max(self._spec.nx, self._spec.nz) > 32:
                # the device-resident linear-Gaussian pieces keep a state vector per thread (<= 32 dimensions); larger
                # models run the flow kernel in its global-workspace mode and evaluate g / log-densities on the host
                self._spec = None
        self._flow = {}
        self._dev_eng = None
        self._generation = 0        # bumped by every device-resident step (see DeviceLedhState)
        self._seed = 0
        self._t = 0

    # ------------------------------------------------------------------ API
    def init_from_gaussian(self, mean0: Array, cov0: Array) -> PFState:
        """Draw particles from the prior, uniform weights, sample moments (reference :84-91)."""
        n, nx = self.cfg.n_particles, mean0.size
        eps = self.cfg.rng.multivariate_normal(np.zeros(nx), cov0, size=n)
        particles = mean0[None, :] + eps
        weights = np.full(n, 1.0 / n)
        mean, cov = self._weighted_stats(particles, weights)
        return PFState(particles=particles, weights=weights, mean=mean, cov=cov, diagnostics={})

    def _buffers(self, n, nx, nz, L) -> _FlowBuffers:
        key = (n, nx, nz, L)
        if key not in self._flow:
            self._flow[key] = _FlowBuffers(n, nx, nz, L, self._dtype, self._device)
        return self._flow[key]

    def step(
        self,
        state: PFState,
        z_k: Array,
        u_km1: Optional[Array] = None,
        process_noise_sampler: Optional[Callable[[int, int], Array]] = None,
    ) -> PFState:
        """One LEDH step (reference :93-214)."""
        if self._spec is not None:
            return self._step_device(state, z_k, process_noise_sampler)
        particles = np.asarray(state.particles, float)
        N, nx = particles.shape
        z_k = np.asarray(z_k, float).reshape(-1)
        nz = self.R.shape[0]
        lib = _lib.load()

        m_pred, P = self.tracker.predict()                        # host EKF/UKF (:105)
        P = 0.5 * (P + P.T)                                       # :106

        if process_noise_sampler is None:
            v = np.zeros((N, nx))
        else:
            v = process_noise_sampler(N, nx)
        eta0 = np.empty_like(particles)
        for i in range(N):                                        # user callbacks on the host (:113-115)
            eta0[i] = self.g(particles[i], u_km1, v[i])

        n_steps = max(1, int(self.cfg.n_lambda_steps))
        buf = self._buffers(N, nx, nz, n_steps)
        buf.load(eta0)
        kP, pP = _lib.as_doubles(P)
        kR, pR = _lib.as_doubles(self.R)
        kz, pz = _lib.as_doubles(z_k)
        if self._meas is not None:                                # all pseudo-time steps in one launch
            kind, hp = self._meas.ledh_functor()
            kh, ph = _lib.as_doubles(hp)
            _lib.check(lib.smc_ledh_flow(C.byref(buf.desc), pP, pR, pz, kind, ph, kh.size, buf.stream))
            eta = buf.eta_host()
        else:                                                     # opaque h / jacobian_h: bridge per step
            torch = buf.torch
            dlam = 1.0 / float(n_steps)
            lam = 0.0
            eta = eta0.copy()
            for j in range(n_steps):
                lam = min(1.0, lam + dlam)                        # :137
                Hs = np.ascontiguousarray(np.array([self.Jh(e) for e in eta], dtype=float).reshape(N, nz, nx))
                hs = np.ascontiguousarray(np.array([self.h(e) for e in eta], dtype=float).reshape(N, nz))
                Hd = torch.from_numpy(Hs).to(buf.device)
                hd = torch.from_numpy(hs).to(buf.device)
                _lib.check(lib.smc_ledh_flow_step_supplied(
                    C.byref(buf.desc), pP, pR, pz, C.c_void_p(Hd.data_ptr()), C.c_void_p(hd.data_ptr()),
                    float(lam), float(dlam), j, buf.stream))
                eta = buf.eta_host()
        s0 = buf.s0.cpu().numpy()
        cond_numbers = []
        for j in range(n_steps):                                  # diagnostic of particle 0 (:152-157)
            try:
                cond_numbers.append(float(np.linalg.cond(s0[j])))
            except Exception:
                cond_numbers.append(float("nan"))

        # weights (:186-195): user log-densities on the host, everything else on the GPU
        terms = np.empty(N)
        for i in range(N):
            num = self.log_trans_pdf(eta[i], particles[i]) + self.log_like_pdf(z_k, eta[i])
            den = self.log_trans_pdf(eta0[i], particles[i])
            terms[i] = num - den
        eng = _weights_engine(N, nx, nz, self._device)
        torch = eng.torch
        eng.x[0][:, :N].copy_(buf.eta[:, :N].to(torch.float64))
        eng.reset(np.asarray(state.mean, float).reshape(-1)[:nx])
        wprev = torch.from_numpy(np.asarray(state.weights, float).reshape(-1)).to(eng.device)
        theta = buf.theta[:N].to(torch.float64).contiguous()
        td = torch.from_numpy(terms).to(eng.device)
        _lib.check(lib.smc_ledh_weights(eng._d, C.c_void_p(wprev.data_ptr()), C.c_void_p(theta.data_ptr()),
                                        C.c_void_p(td.data_ptr()), eng.stream))
        ratio = float(self.cfg.resample_ess_ratio)
        eng.finalize(ratio if ratio > 0.0 else 0.0, u=0.0)
        row = eng.read_row()

        self.tracker.update(z_k)                                  # host EKF/UKF (:198)

        if bool(row[2]):                                          # :201-206
            eng.set_uniform(float(self.cfg.rng.random()))
            eng.resample()
            eng.finish_step()
            xk = eng.get_particles()
            w = np.full(N, 1.0 / N)
        else:
            eng.finish_step()
            xk = eta
            w = eng.weights()
        row = eng.read_row()
        H0 = _lib.SMC_STATS_HEADER
        mean = row[H0:H0 + nx].copy()
        cov = row[H0 + nx:].reshape(nx, nx).copy()
        cov = 0.5 * (cov + cov.T)
        return PFState(particles=xk, weights=w, mean=mean, cov=cov,
                       diagnostics={"condition_numbers": cond_numbers})

    # ------------------------------------------------------------------ device-resident step
    def _step_device(self, state, z_k, process_noise_sampler, philox_seed=None):
        """Whole step on the GPU for a ``LedhGaussianSSM`` spec: propagate, flow, log-density
        terms, weights, ESS-triggered systematic resampling, moments.  The particle set only
        crosses PCIe when the caller reads ``state.particles``.  With
        ``process_noise_sampler=None`` the reference uses v = 0 (:109-110); pass
        ``"philox"`` to draw v ~ N(0, Q) on chip instead."""
        from .engine import SirEngine

        sp = self._spec
        lib = _lib.load()
        nx, nz = sp.nx, sp.nz
        z_k = np.asarray(z_k, float).reshape(-1)
        # the engine still holds exactly this state: it is ours, it is the latest one, and nobody has taken (and
        # possibly modified) its arrays
        resident = isinstance(state, DeviceLedhState) and state._owner is self and state._is_current() \
            and state._particles is None and state._weights is None
        if isinstance(state, DeviceLedhState) and not resident and state._particles is None:
            state._check()          # stale and never read: its arrays are gone
        N = self.cfg.n_particles if resident else np.asarray(state.particles).shape[0]
        if self._dev_eng is None or self._dev_eng.n != N:
            self._dev_eng = SirEngine(N, nx, nz, dtype=self._dtype, device=self._device, write_ancestors=True,
                                      generic_moments=True)
        eng = self._dev_eng
        torch = eng.torch
        if not resident:
            eng.set_particles(np.asarray(state.particles, float), shift=np.asarray(state.mean, float).reshape(-1)[:nx])
            wprev = torch.from_numpy(np.asarray(state.weights, float).reshape(-1).copy()).to(eng.device)
        else:
            wprev = torch.empty((N,), dtype=torch.float64, device=eng.device)
            _lib.check(lib.smc_sir_export_weights(eng._d, C.c_void_p(wprev.data_ptr()), eng.stream))
        cur = eng.current_buffer()
        xprev = eng.x[cur]

        m_pred, P = self.tracker.predict()
        P = 0.5 * (P + P.T)
        n_steps = max(1, int(self.cfg.n_lambda_steps))
        buf = self._buffers(N, nx, nz, n_steps)
        kF, pF = _lib.as_doubles(sp.F)
        kL, pL = _lib.as_doubles(sp.Lq)
        v_dev = None
        if process_noise_sampler is None:
            v_dev = torch.zeros((nx, buf.ld), dtype=buf.tdtype, device=buf.device)
        elif process_noise_sampler != "philox":
            v = np.asarray(process_noise_sampler(N, nx), float)
            v_dev = torch.zeros((nx, buf.ld), dtype=buf.tdtype, device=buf.device)
            v_dev[:, :N].copy_(torch.from_numpy(np.ascontiguousarray(v.T).astype(buf.npdtype, copy=False)))
        _lib.check(lib.smc_ledh_propagate_linear(
            C.byref(buf.desc), C.c_void_p(xprev.data_ptr()), pF, pL,
            C.c_void_p(v_dev.data_ptr()) if v_dev is not None else None, C.c_uint64(self._seed), self._t, buf.stream))
        self._t += 1
        kP, pP = _lib.as_doubles(P)
        kR, pR = _lib.as_doubles(self.R)
        kz, pz = _lib.as_doubles(z_k)
        kind, hp = self._meas.ledh_functor()
        kh, ph = _lib.as_doubles(hp)
        _lib.check(lib.smc_ledh_flow(C.byref(buf.desc), pP, pR, pz, kind, ph, kh.size, buf.stream))
        terms = torch.empty((N,), dtype=torch.float64, device=buf.device)
        kQi, pQi = _lib.as_doubles(sp.Qinv)
        kRi, pRi = _lib.as_doubles(sp.Rinv)
        _lib.check(lib.smc_ledh_gauss_terms(C.byref(buf.desc), C.c_void_p(xprev.data_ptr()), pF, pQi, pRi, pz, kind,
                                            ph, kh.size, C.c_void_p(terms.data_ptr()), buf.stream))
        # weights on the SIR kernels: the flowed particles become the engine's current buffer
        eng.x[cur][:, :N].copy_(buf.eta[:, :N])
        _lib.check(lib.smc_ledh_weights(eng._d, C.c_void_p(wprev.data_ptr()), C.c_void_p(buf.theta.data_ptr()),
                                        C.c_void_p(terms.data_ptr()), eng.stream))
        ratio = float(self.cfg.resample_ess_ratio)
        eng.finalize(ratio if ratio > 0.0 else 0.0, u=0.0)
        row = eng.read_row()
        self.tracker.update(z_k)
        if bool(row[2]):
            eng.set_uniform(float(self.cfg.rng.random()))
            eng.resample()
        eng.finish_step()
        row = eng.read_row()
        H0 = _lib.SMC_STATS_HEADER
        mean = row[H0:H0 + nx].copy()
        cov = row[H0 + nx:].reshape(nx, nx).copy()
        cov = 0.5 * (cov + cov.T)
        s0 = buf.s0.cpu().numpy()
        cond = []
        for j in range(n_steps):                                  # diagnostic of particle 0 (:152-157, bare except there too)
            try:
                cond.append(float(np.linalg.cond(s0[j])))
            except Exception:
                cond.append(float("nan"))
        self.last_ess = float(row[1])
        self.last_resampled = bool(row[2])
        self._generation += 1
        return DeviceLedhState(self, mean, cov, {"condition_numbers": cond})

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _weighted_stats(x: Array, w: Array) -> Tuple[Array, Array]:
        """Weighted mean and symmetrised covariance (reference :217-224), on the GPU."""
        x = np.asarray(x, float)
        n, nx = x.shape
        lib = _lib.load()
        eng = _weights_engine(n, nx, 1)
        eng.set_particles(x, shift=x[0])
        wd = eng.torch.from_numpy(np.asarray(w, float).reshape(-1)).to(eng.device)
        _lib.check(lib.smc_weighted_moments(eng._d, C.c_void_p(wd.data_ptr()), eng._row(0), eng.stream))
        row = eng.read_row()
        H0 = _lib.SMC_STATS_HEADER
        mean = row[H0:H0 + nx].copy()
        cov = row[H0 + nx:].reshape(nx, nx).copy()
        return mean, 0.5 * (cov + cov.T)
