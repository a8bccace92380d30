"""Drop-in SIR particle filter on the B200 engine.

Mirrors the public surface of the reference's ``models/particle_filter.py``
(``PFState`` :27-49, ``ParticleFilter`` :53-287: same constructor, attributes,
methods, private helpers and error behaviour), so the reference's simulators,
notebooks and tests drive it unchanged.  What lies beneath is new: every engine
operation (noise add, weights, log-sum-exp, ESS, CDF scan, search, gather,
moments) is a libsmc CUDA kernel; there is no CPU fallback.

Two ways the state-space model reaches the GPU:

* **spec models** -- ``g`` / ``h`` are the bound methods of a
  :mod:`particle_filters_b200.ssm` object: the fused propagate+likelihood kernel
  runs the compiled functor and the particle set stays resident in HBM;
* **host-callback bridge** -- any other Python callables are evaluated on the
  host exactly as the reference does (``[g(x, u) for x in particles]``) and
  their results uploaded; that is user code, not an engine fallback.

Extra keyword-only constructor arguments (additive): ``dtype`` ("f64" parity,
"f32" throughput), ``noise`` ("reference": draws come from ``rng`` in the
reference's call order and are uploaded; "philox": on-chip Philox4x32-10 keyed
by ``seed``), ``seed``, ``device``, ``record_ancestors``, ``engine`` (adopt an
existing :class:`~particle_filters_b200.engine.SirEngine`).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib
from .ssm import spec_of

Array = np.ndarray
GFn = Callable[[Array, Optional[Array]], Array]
HFn = Callable[[Array], Array]


@dataclass
class PFState:
    """Particle-filter posterior (reference particle_filter.py:27-49)."""

    particles: Array
    weights: Array
    mean: Array
    cov: Array
    t: int


class DevicePFState(PFState):
    """``PFState`` whose particle/weight arrays stay on the GPU until read."""

    def __init__(self, owner: "ParticleFilter", mean: Array, cov: Array, t: int):
        self._owner, self._gen = owner, owner._generation
        self._particles = self._weights = None
        self.mean, self.cov, self.t = mean, cov, t

    def _check(self):
        if self._gen != self._owner._generation:
            raise RuntimeError("stale PFState: the filter has advanced since this state was returned; "
                               "read .particles/.weights before the next predict/update")

    @property
    def particles(self) -> Array:
        if self._particles is None:
            self._check()
            self._particles = self._owner._eng.get_particles()
        return self._particles

    @particles.setter
    def particles(self, value):
        self._check()
        self._particles = np.asarray(value, float)
        self._owner._eng_upload_particles(self._particles)

    @property
    def weights(self) -> Array:
        if self._weights is None:
            self._check()
            self._weights = self._owner._eng.weights()
        return self._weights


def _lower_inverse(L: Array) -> Array:
    return np.linalg.solve(L, np.eye(L.shape[0]))


class ParticleFilter:
    """SIR particle filter with systematic / multinomial resampling."""

    def __init__(
        self,
        g: GFn,
        h: HFn,
        Q: Array,
        R: Array,
        *,
        Np: int = 1000,
        resample_thresh: float = 0.5,
        resample_method: str = "systematic",
        regularize_after_resample: bool = False,
        rng: Optional[np.random.Generator] = None,
        dtype: str = "f64",
        noise: str = "reference",
        seed: Optional[int] = None,
        device=None,
        record_ancestors: bool = False,
        engine=None,
    ) -> None:
        self.g = g
        self.h = h
        self.Q = np.asarray(Q, float)
        self.R = np.asarray(R, float)
        self.Np = int(Np)
        self.resample_thresh = float(resample_thresh)
        self.resample_method = resample_method
        self.regularize_after_resample = regularize_after_resample
        self.rng = np.random.default_rng() if rng is None else rng

        self.nx = self.Q.shape[0]
        self.nz = self.R.shape[0]
        self.state: Optional[PFState] = None

        self.LR = np.linalg.cholesky(self.R + 1e-12 * np.eye(self.nz))   # reference :107

        if dtype not in ("f32", "f64"):
            raise ValueError("dtype must be 'f32' or 'f64'")
        if noise not in ("reference", "philox"):
            raise ValueError("noise must be 'reference' or 'philox'")
        self._dtype, self._noise, self._device = dtype, noise, device
        self._seed = int(self.rng.integers(0, 2 ** 63 - 1)) if (seed is None and noise == "philox") else int(seed or 0)
        self._spec = spec_of(g, h)
        # device-resident engine (spec models).  ``engine=``: adopt an existing SirEngine of the right shape
        # (its particle arrays stay where they are -- e.g. a 2^26-particle population that is already resident)
        self._eng = None
        if engine is not None:
            if self._spec is None:
                raise ValueError("engine= needs a spec model (g/h bound to a particle_filters_b200.ssm object)")
            if (engine.n, engine.nx, engine.nz) != (self.Np, self.nx, self.nz) or \
                    engine.npdtype != (np.float32 if dtype == "f32" else np.float64):
                raise ValueError("engine= does not match this filter's (Np, nx, nz, dtype)")
            self._eng = engine
            self._functor = self._spec.device_functor(self.Q, self.R)
        self._bridge = {}           # bridge engines by particle count
        self._generation = 0
        self._t_phil = 0            # Philox step counter
        self._n_predict = 0         # predict-only calls (own Philox sub-stream)
        # additive diagnostics of the last update (the reference computes and discards them)
        self.last_log_norm = None
        self.last_ess = None
        self.last_resampled = None
        self.last_ancestors = None
        self._record_ancestors = bool(record_ancestors)

    # ------------------------------------------------------------------ engines
    def _bridge_engine(self, n: int):
        from .engine import SirEngine

        eng = self._bridge.get(n)
        if eng is None:
            eng = SirEngine(n, self.nx, self.nz, dtype="f64", device=self._device, write_ancestors=True,
                            generic_moments=True)
            self._bridge[n] = eng
        return eng

    def _device_engine(self):
        from .engine import SirEngine

        if self._eng is None:
            post_generic = self.resample_method != "systematic" or bool(self.regularize_after_resample)
            self._eng = SirEngine(self.Np, self.nx, self.nz, dtype=self._dtype, device=self._device,
                                  write_ancestors=self._record_ancestors, post_moments_generic=post_generic)
            self._functor = self._spec.device_functor(self.Q, self.R)
        return self._eng

    def _eng_upload_particles(self, particles):
        self._eng.upload_particles(particles)        # weights untouched, as in the reference

    @staticmethod
    def _reject_control(u):
        """Spec models compile ``g(x)`` without a control input: a ``u`` the reference would hand to ``g`` for every
        particle (particle_filter.py:237) must not be dropped silently."""
        if u is not None:
            raise ValueError("spec models take no control input u; pass plain callables g(x, u) / h(x) to use one "
                             "(host-callback bridge)")

    def _chol_Q(self, eps: float) -> Array:
        try:
            return np.linalg.cholesky(self.Q)
        except np.linalg.LinAlgError:
            return np.linalg.cholesky(self.Q + eps * np.eye(self.nx))

    # ------------------------------------------------------------------ initialisation
    def initialize(self, mean: Array, cov: Array) -> PFState:
        """Draw particles from N(mean, cov), uniform weights (reference :110-132).

        ``state.mean`` / ``state.cov`` are the *given* prior moments, as in the reference."""
        mean = np.asarray(mean, float)
        cov = np.asarray(cov, float)
        Lc = np.linalg.cholesky(cov + 1e-10 * np.eye(len(mean)))
        cov2 = np.atleast_2d(cov)
        self._generation += 1
        self._t_phil = 0
        if self._spec is not None:
            eng = self._device_engine()
            if self._noise == "reference":
                particles = self.rng.standard_normal((self.Np, len(mean))) @ Lc.T + mean
                eng.set_particles(particles, shift=mean)
            else:
                eng.init_gaussian(mean, Lc, self._seed)
            self.state = DevicePFState(self, mean, cov2, 0)
            return self.state
        particles = self.rng.standard_normal((self.Np, len(mean))) @ Lc.T + mean
        weights = np.ones(self.Np) / self.Np
        self.state = PFState(particles, weights, mean, cov2, 0)
        return self.state

    def effective_sample_size(self) -> float:
        """1 / sum w^2 of the stored weights (reference :134-144), reduced on the GPU."""
        assert self.state is not None, "Filter not initialized."
        w = np.asarray(self.state.weights, float)
        ess_normalised, eng, _ = self._device_ess(w)
        # the kernels reduce (sum w)^2 / sum w^2; the reference reports 1 / sum w^2
        return ess_normalised / float(np.exp(2.0 * eng.read_row()[0]))

    def _device_ess(self, w: Array):
        """(ESS, engine) of host weights through the weight-normalisation kernels."""
        n = len(w)
        eng = self._bridge_engine(n)
        eng.reset()
        eng.bridge_loglik(np.zeros((n, self.nz)), np.zeros(self.nz), np.eye(self.nz), wprev=w)
        eng.finalize(self.resample_thresh * self.Np / n, u=0.0)
        row = eng.read_row()
        return float(row[1]), eng, bool(row[2])

    # ------------------------------------------------------------------ resampling helpers
    def _systematic_resample(self, weights: Array) -> Array:
        """Systematic resampling indices (reference :146-171) from the GPU resampler."""
        from .engine import systematic_ancestors

        return systematic_ancestors(np.asarray(weights, float), self.rng.random(), device=self._device)

    def _multinomial_resample(self, weights: Array) -> Array:
        """Multinomial resampling indices (reference :173-186) from the GPU resampler.  The caller's
        NumPy stream is consumed exactly as ``rng.choice(N, size=N, p=weights)`` consumes it (one
        ``rng.random(N)``), and the same arguments are rejected with the same ``ValueError``."""
        from .engine import multinomial_ancestors

        p = np.asarray(weights, float).reshape(-1)
        if np.isnan(p).any():
            raise ValueError("probabilities contain NaN")
        if (p < 0).any():
            raise ValueError("probabilities are not non-negative")
        if abs(float(np.sum(p)) - 1.0) > np.sqrt(np.finfo(np.float64).eps):
            raise ValueError("probabilities do not sum to 1")
        return multinomial_ancestors(p, self.rng.random(len(p)), device=self._device)

    def _resample(self, particles: Array, weights: Array) -> Tuple[Array, Array]:
        """Resample when Neff < resample_thresh * Np (reference :188-220)."""
        particles = np.asarray(particles)
        ess, eng, do_res = self._device_ess(np.asarray(weights, float))
        if not do_res:
            return particles, weights
        eng.x[0][:, : eng.n].copy_(eng.torch.from_numpy(
            np.ascontiguousarray(np.asarray(particles, float).reshape(eng.n, self.nx).T)))
        new_p, new_w, _, _ = self._finish_resampled(eng, np.asarray(weights, float))
        return new_p, new_w

    def _finish_resampled(self, eng, w_norm: Optional[Array]):
        """Tail of ``_resample`` once the device decided to resample: ancestors,
        gather, optional jitter, uniform weights, moments."""
        n = eng.n
        if self.resample_method == "systematic":
            eng.set_uniform(self.rng.random())
            eng.resample()
        else:
            eng.resample_multinomial(u=self.rng.random(n))
        if self.regularize_after_resample:
            Lq = self._chol_Q(1e-12)
            jitter = self.rng.standard_normal((n, self.nx))
            eng.bridge_predict(None, jitter, Lq, scale=0.001, target_other=True)
        eng.finish_step()
        row = eng.read_row()
        mean = row[_lib.SMC_STATS_HEADER:_lib.SMC_STATS_HEADER + self.nx].copy()
        cov = row[_lib.SMC_STATS_HEADER + self.nx:].reshape(self.nx, self.nx).copy()
        return eng.get_particles(), np.ones(n) / n, mean, cov

    # ------------------------------------------------------------------ filtering
    def predict(self, u: Optional[Array] = None) -> None:
        """Propagate particles through the transition model (reference :223-237)."""
        assert self.state is not None, "Filter not initialized."
        Lq = self._chol_Q(1e-10)
        if self._spec is not None:
            self._reject_control(u)
            eng = self._eng
            self._generation += 1
            mid, params = self._functor
            noise = None
            if self._noise == "reference":
                noise = eng.upload_soa(self.rng.standard_normal((self.Np, self.nx)), self.nx)
            # Philox: predict-only calls draw from their own sub-stream (bit 30 of the step counter + a call count),
            # so two predict() calls without an update do not reuse the noise of step()
            t_noise = self._t_phil if self._noise == "reference" else (1 << 30) | (self._n_predict & ((1 << 30) - 1))
            self._n_predict += 1
            eng.propagate_weight(mid, params, _lib.SMC_MODE_PROPAGATE, noise=noise, seed=self._seed, t=t_noise)
            self.state = DevicePFState(self, self.state.mean, self.state.cov, self.state.t)
            return
        noise = self.rng.standard_normal((self.Np, self.nx))
        gx = np.array([self.g(x, u) for x in self.state.particles])      # user callbacks on the host
        eng = self._bridge_engine(self.Np)
        eng.reset()
        eng.bridge_predict(gx.reshape(self.Np, self.nx), noise, Lq)
        self.state.particles = eng.get_particles()

    def update(self, z: Array) -> PFState:
        """Weight by the likelihood, resample if degenerate, moments (reference :239-269)."""
        assert self.state is not None, "Filter not initialized."
        z = np.asarray(z, float)
        t_next = self.state.t + 1
        if self._spec is not None:
            return self._update_device(z, t_next, _lib.SMC_MODE_WEIGHT)
        particles = self.state.particles
        weights = self.state.weights
        z_pred = np.array([self.h(x) for x in particles])                # user callbacks on the host
        eng = self._bridge_engine(self.Np)
        eng.set_particles(particles, shift=np.asarray(self.state.mean, float).reshape(-1)[: self.nx])
        eng.bridge_loglik(z_pred.reshape(self.Np, self.nz), z.reshape(-1), _lower_inverse(self.LR), wprev=weights)
        eng.finalize(self.resample_thresh, u=0.0)
        row = eng.read_row()
        self.last_log_norm, self.last_ess, self.last_resampled = float(row[0]), float(row[1]), bool(row[2])
        if self.last_resampled:
            particles, w, mean, cov = self._finish_resampled(eng, None)
        else:
            w = eng.weights()
            eng.finish_step()
            row = eng.read_row()
            mean = row[_lib.SMC_STATS_HEADER:_lib.SMC_STATS_HEADER + self.nx].copy()
            cov = row[_lib.SMC_STATS_HEADER + self.nx:].reshape(self.nx, self.nx).copy()
        self.state = PFState(particles, w, mean, np.atleast_2d(cov), t_next)
        return self.state

    def _update_device(self, z: Array, t_next: int, mode: int) -> PFState:
        eng = self._eng
        self._generation += 1
        mid, params = self._functor
        host_rng = self._noise == "reference" or self.resample_method != "systematic"
        if not host_rng and mode == _lib.SMC_MODE_STEP and not self.regularize_after_resample:
            # whole step on the device clock -- K1 (+ finalize), range pre-pass, resampler (+ finish) -- and the step's
            # host traffic (z in, statistics row out) in ONE library call
            if getattr(self, "_functor_c", None) is None or self._functor_c[2] is not params:
                self._functor_c = _lib.as_doubles(params) + (params,)
            row = eng.step_host(mid, self._functor_c[:2], z, self.resample_thresh, seed=self._seed, t=self._t_phil)
            return self._publish_device_state(eng, t_next, row)
        eng.set_observation(z, 0)
        noise = None
        if (mode & _lib.SMC_MODE_PROPAGATE) and self._noise == "reference":
            noise = eng.upload_soa(self.rng.standard_normal((self.Np, self.nx)), self.nx)
        eng.propagate_weight(mid, params, mode, z_index=0, noise=noise, seed=self._seed, t=self._t_phil)
        if host_rng:
            eng.finalize(self.resample_thresh, u=0.0)
            row = eng.read_row()
            if bool(row[2]):
                if self.resample_method == "systematic":
                    eng.set_uniform(self.rng.random())
                    eng.resample()
                else:
                    self._device_multinomial(eng)
                if self.regularize_after_resample:
                    jitter = self.rng.standard_normal((self.Np, self.nx))
                    eng.bridge_predict(None, jitter, self._chol_Q(1e-12), scale=0.001, target_other=True)
        else:
            eng.finalize(self.resample_thresh, u=-1.0, seed=self._seed, t=self._t_phil)
            if self.resample_thresh > 0:
                eng.resample()
                if self.regularize_after_resample:      # on chip, decided by the device's own resampling flag
                    eng.jitter(self._chol_Q(1e-12), 0.001, self._seed, self._t_phil)
        eng.finish_step()
        return self._publish_device_state(eng, t_next)

    def _publish_device_state(self, eng, t_next: int, row=None) -> PFState:
        self._t_phil += 1
        if row is None:
            row = eng.read_row()
        self.last_log_norm, self.last_ess, self.last_resampled = float(row[0]), float(row[1]), bool(row[2])
        mean = row[_lib.SMC_STATS_HEADER:_lib.SMC_STATS_HEADER + self.nx].copy()
        cov = row[_lib.SMC_STATS_HEADER + self.nx:].reshape(self.nx, self.nx).copy()
        if self._record_ancestors:
            self.last_ancestors = (eng.ancestors[: self.Np].cpu().numpy().astype(np.int64)
                                   if self.last_resampled else None)
        self.state = DevicePFState(self, mean, np.atleast_2d(cov), t_next)
        return self.state

    def _device_multinomial(self, eng):
        """Ancestors + gather on the device; only the uniforms come from the host, and only in
        reference-noise mode (so that ``self.rng`` advances as the reference's ``rng.choice`` does)."""
        if self._noise == "reference":
            eng.resample_multinomial(u=self.rng.random(eng.n))
        else:
            eng.resample_multinomial(seed=self._seed, t=self._t_phil)

    def step(self, z: Array, u: Optional[Array] = None) -> PFState:
        """One PF step: predict then update (reference :271-287); fused into a single
        propagate+likelihood launch for spec models."""
        if self._spec is not None:
            assert self.state is not None, "Filter not initialized."
            self._reject_control(u)
            return self._update_device(np.asarray(z, float), self.state.t + 1, _lib.SMC_MODE_STEP)
        self.predict(u)
        return self.update(z)

    # ------------------------------------------------------------------ bulk API (additive)
    def run(self, Z: Array) -> dict:
        """Run ``len(Z)`` steps back to back with no host synchronisation (spec models,
        Philox noise, systematic resampling).  Returns per-step traces."""
        assert self.state is not None, "Filter not initialized."
        if self._spec is None or self._noise != "philox" or self.resample_method != "systematic" \
                or self.regularize_after_resample:
            raise ValueError("run() needs a spec model with noise='philox' and plain systematic resampling")
        Z = np.asarray(Z, float).reshape(len(Z), -1)
        eng = self._eng
        self._generation += 1
        mid, params = self._functor
        rows = eng.run(mid, params, Z, self.resample_thresh, self._seed, t0=self._t_phil)
        self._t_phil += len(Z)
        H, nx = _lib.SMC_STATS_HEADER, self.nx
        out = {"log_norm": rows[:, 0], "ess": rows[:, 1], "resampled": rows[:, 2] > 0.5, "u": rows[:, 3],
               "mean": rows[:, H:H + nx], "cov": rows[:, H + nx:].reshape(-1, nx, nx)}
        self.state = DevicePFState(self, out["mean"][-1].copy(), out["cov"][-1].copy(), self.state.t + len(Z))
        return out
