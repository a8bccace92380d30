#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run from the repo root in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference ships no golden vectors for the hot path,
so these fixtures are outputs of the reference itself: the reference modules
are loaded by file path from /root/reference (nothing is copied), driven on
seeded inputs, and their inputs/outputs are stored.  The GPU box has no
/root/reference; the tests there use only these files.

RNG streams are NOT stored (too large): the fixtures record the seed and a
checksum of the first draws, and tests regenerate the identical draws with
``np.random.default_rng(seed)`` (test_golden_rng_stream_is_stable guards drift).
"""

from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SMC_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(REPO, "tests"))

from ssm_zoo import AcousticCVModel, LGSSMModel, Linear1DModel, Lorenz96Model, SVModel  # noqa: E402


def load_ref(relpath: str, alias: str):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


ref_pf = load_ref("models/particle_filter.py", "ref_particle_filter")
ref_ledh = load_ref("models/LEDH_particle_filter.py", "ref_ledh_particle_filter")
ref_sv = load_ref("simulator/simulator_sto_volatility_model.py", "ref_sim_sv")
ref_lgssm = load_ref("simulator/simulator_LGSSM.py", "ref_sim_lgssm")
ref_l96 = load_ref("simulator/simulator_Lorenz_96.py", "ref_sim_l96")
ref_edh = load_ref("models/EDH_particle_filter.py", "ref_edh_particle_filter")


class FixedU:
    """rng stub: ``random()`` returns a scripted uniform."""

    def __init__(self, u):
        self.u = u

    def random(self):
        return self.u


def weight_cases(rng):
    cases = []
    for N in (1, 2, 3, 7, 64, 257, 1000, 3000):
        w = rng.random(N)
        cases.append(("random", w / w.sum()))
        w = rng.random(N) ** 8
        cases.append(("peaked", w / w.sum()))
        if N >= 3:
            w = rng.random(N)
            w[rng.random(N) < 0.4] = 0.0
            w[0] = 0.0
            w[-1] = 0.0
            if w.sum() == 0:
                w[N // 2] = 1.0
            cases.append(("zeros", w / w.sum()))
            w = np.full(N, 1e-4)
            w[N // 3] = 0.9
            cases.append(("dominant", w / w.sum()))
        cases.append(("uniform", np.full(N, 1.0 / N)))
    return cases


def gen_resample():
    rng = np.random.default_rng(2024)
    out = {}
    k = 0
    for name, w in weight_cases(rng):
        for u in (float(rng.random()), 0.0, 1.0 - 2.0 ** -20):
            pf = ref_pf.ParticleFilter(lambda x, u_: x, lambda x: x, np.eye(1), np.eye(1),
                                       Np=len(w), rng=FixedU(u))
            idx_pf = pf._systematic_resample(w)
            try:
                idx_ledh = ref_ledh.systematic_resample(w, FixedU(u))
                ledh_ok = True
            except IndexError:  # reference runs off the end when cdf[-1] <= pos[-1]
                idx_ledh = np.full(len(w), -1)
                ledh_ok = False
            out[f"c{k}_name"] = np.array(name)
            out[f"c{k}_w"] = w
            out[f"c{k}_u"] = np.array(u)
            out[f"c{k}_pf"] = idx_pf.astype(np.int64)
            out[f"c{k}_ledh"] = idx_ledh.astype(np.int64)
            out[f"c{k}_ledh_ok"] = np.array(ledh_ok)
            k += 1
    out["n_cases"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "resample_cases.npz"), **out)
    print("resample_cases:", k)


def run_ref_pf(model, Y, *, Np, seed, mean0, cov0, thresh=0.5, method="systematic",
               regularize=False, store_particles_every=0):
    """Drive the reference ParticleFilter, spying on _resample's inputs/outputs."""
    pf = ref_pf.ParticleFilter(model.g, model.h, model.Q, model.R, Np=Np, resample_thresh=thresh,
                               resample_method=method, regularize_after_resample=regularize,
                               rng=np.random.default_rng(seed))
    spy = {}
    orig_resample = pf._resample
    orig_sys = pf._systematic_resample

    def spy_sys(weights):
        idx = orig_sys(weights)
        spy["idx"] = idx.copy()
        return idx

    def spy_resample(particles, weights):
        spy["w_pre"] = weights.copy()
        spy["idx"] = None
        return orig_resample(particles, weights)

    pf._systematic_resample = spy_sys
    pf._resample = spy_resample
    pf.initialize(mean0, cov0)
    T = len(Y)
    nx = model.Q.shape[0]
    means = np.empty((T, nx))
    covs = np.empty((T, nx, nx))
    ess_pre = np.empty(T)
    ess_post = np.empty(T)
    lognorm = np.empty(T)
    resampled = np.zeros(T, bool)
    anc_hash = np.zeros(T, np.int64)
    anc_first = {}
    w_prev = pf.state.weights.copy()
    for t in range(T):
        st = pf.step(np.atleast_1d(Y[t]))
        means[t], covs[t] = st.mean, st.cov
        w = spy["w_pre"]
        ess_pre[t] = 1.0 / np.sum(w ** 2)
        ess_post[t] = pf.effective_sample_size()
        resampled[t] = spy["idx"] is not None or (method != "systematic" and ess_pre[t] < thresh * Np)
        if spy["idx"] is not None:
            idx = spy["idx"]
            anc_hash[t] = int(np.sum(idx.astype(np.int64) * (np.arange(Np) % 1009 + 1)) % (2 ** 61 - 1))
            if len(anc_first) < 3:
                anc_first[t] = idx.astype(np.int32)
    out = dict(Y=np.asarray(Y), means=means, covs=covs, ess_pre=ess_pre, ess_post=ess_post,
               resampled=resampled, anc_hash=anc_hash, final_particles=pf.state.particles,
               final_weights=pf.state.weights, Np=np.array(Np), seed=np.array(seed),
               mean0=np.asarray(mean0, float), cov0=np.asarray(cov0, float),
               thresh=np.array(thresh), method=np.array(method), regularize=np.array(regularize),
               Q=model.Q, R=model.R)
    for t, idx in anc_first.items():
        out[f"anc_t{t}"] = idx
    out["anc_steps"] = np.array(sorted(anc_first))
    return out


def gen_sir():
    # C1: SV, alpha=.95 sigma=.2 beta=1, T=1000, seed=42, N=1000 (BASELINE.json configs[0])
    sv = ref_sv.simulate_sv_1d(n=1000, alpha=0.95, sigma=0.2, beta=1.0, seed=42)
    for tag, r in (("c1", None), ("c1_r01", 0.1)):
        m = SVModel(0.95, 0.2, 1.0, r=r)
        out = run_ref_pf(m, sv.Y, Np=1000, seed=42, mean0=np.zeros(1),
                         cov0=np.array([[0.2 ** 2 / (1 - 0.95 ** 2)]]))
        out["X_true"] = sv.X
        out["alpha"], out["sigma"], out["beta"] = map(np.array, (0.95, 0.2, 1.0))
        np.savez_compressed(os.path.join(HERE, f"sir_sv_{tag}.npz"), **out)
        print(f"sir_sv_{tag}: resampled {int(out['resampled'].sum())}/1000 steps")
    # bootstrap (resample every step) + regularisation + multinomial variants, shorter
    m = SVModel(0.95, 0.2, 1.0)
    out = run_ref_pf(m, sv.Y[:200], Np=512, seed=7, mean0=np.zeros(1), cov0=np.array([[0.41]]),
                     thresh=1.1)
    np.savez_compressed(os.path.join(HERE, "sir_sv_bootstrap.npz"), **out)
    out = run_ref_pf(m, sv.Y[:200], Np=300, seed=8, mean0=np.zeros(1), cov0=np.array([[0.41]]),
                     regularize=True)
    np.savez_compressed(os.path.join(HERE, "sir_sv_regularized.npz"), **out)
    out = run_ref_pf(m, sv.Y[:100], Np=300, seed=9, mean0=np.zeros(1), cov0=np.array([[0.41]]),
                     method="multinomial")
    np.savez_compressed(os.path.join(HERE, "sir_sv_multinomial.npz"), **out)
    # C2-shaped: 2-D LGSSM (README A/B/C/D), reference simulator, small N
    lg = LGSSMModel()
    sim = ref_lgssm.simulate_lgssm(lg.A, lg.B, lg.C, lg.D, np.eye(2), 200, seed=42)
    out = run_ref_pf(lg, sim.Y, Np=800, seed=42, mean0=np.zeros(2), cov0=np.eye(2))
    out["X_true"] = sim.X
    np.savez_compressed(os.path.join(HERE, "sir_lgssm.npz"), **out)
    print("sir_lgssm: resampled", int(out["resampled"].sum()), "/ 200")


def gen_l96():
    """C4-shaped (reduced): Lorenz-96 through the reference PF, dynamics = the reference's own
    rk4_step / l96_rhs, observations = ObsModel.H with H_idx = arange(0, nx, obs_fraction)."""
    for tag, nx, n_sub, stride, Np, T in (("nx8", 8, 5, 2, 400, 40), ("nx40", 40, 20, 4, 256, 12)):
        m = Lorenz96Model(nx=nx, n_sub=n_sub, obs_stride=stride, q=0.1, r=1.0)
        obs = ref_l96.ObsModel(H_idx=np.arange(0, nx, stride), R=m.R)

        class RefDyn:
            Q, R = m.Q, m.R

            @staticmethod
            def g(x, u=None):
                for _ in range(n_sub):
                    x = ref_l96.rk4_step(x, m.dt, lambda z: ref_l96.l96_rhs(z, m.F))
                return x

            h = staticmethod(obs.H)

        X, Y = m.simulate(T, seed=42)
        x0 = X[0] + 0.5
        out = run_ref_pf(RefDyn, Y, Np=Np, seed=11, mean0=x0, cov0=0.5 * np.eye(nx))
        out["X_true"] = X
        for k in ("nx", "n_sub", "obs_stride", "F", "dt"):
            out[k] = np.array(getattr(m, k))
        np.savez_compressed(os.path.join(HERE, f"sir_l96_{tag}.npz"), **out)
        print(f"sir_l96_{tag}: resampled", int(out["resampled"].sum()), "/", T)


class ScriptedTracker:
    """GaussianTracker (LEDH_particle_filter.py:13-16) replaying scripted (m, P)."""

    def __init__(self, Ps, nx):
        self.Ps, self.k, self.nx = Ps, 0, nx

    def predict(self):
        P = self.Ps[self.k]
        self.k += 1
        return np.zeros(self.nx), P

    def update(self, z):
        return np.zeros(self.nx), self.Ps[self.k - 1]

    def get_past_mean(self):
        return np.zeros(self.nx)


def gen_ledh(only=None):
    def run(tag, model, *, N, L, T, ess_ratio, seed, mean0, cov0, x_true0, with_noise):
        if only is not None and tag not in only:
            return
        rng = np.random.default_rng(seed)
        nx, nz = model.nx, model.nz
        # scripted covariances: SPD, slowly contracting, slightly asymmetric (exercises :106)
        Ps = []
        Pk = np.array(cov0, float)
        for k in range(T):
            G = rng.standard_normal((nx, nx)) * 0.02
            Pk = 0.8 * Pk + model.Q + G @ G.T
            Ps.append(Pk + 1e-9 * np.triu(np.ones((nx, nx)), 1))
        cfg = ref_ledh.LEDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=ess_ratio,
                                  rng=np.random.default_rng(seed + 1))
        f = ref_ledh.LEDHFlowPF(ScriptedTracker(Ps, nx), model.g, model.h, model.jac_h,
                                model.log_trans_pdf, model.log_like_pdf, model.R, cfg)
        st = f.init_from_gaussian(np.asarray(mean0, float), np.asarray(cov0, float))
        Lq = np.linalg.cholesky(model.Q)
        x = np.asarray(x_true0, float)
        out = dict(N=np.array(N), L=np.array(L), T=np.array(T), ess_ratio=np.array(ess_ratio),
                   seed=np.array(seed), Ps=np.array(Ps), Q=model.Q, R=model.R,
                   particles0=st.particles.copy(), weights0=st.weights.copy())
        Z, V = [], []
        for k in range(T):
            x = model.g(x, None, Lq @ rng.standard_normal(nx))
            z = model.h(x) + np.linalg.cholesky(model.R) @ rng.standard_normal(nz)
            v = (rng.standard_normal((N, nx)) @ Lq.T) if with_noise else None
            sampler = (lambda n_, d_, v=v: v) if with_noise else None
            st = f.step(st, z, None, sampler)
            Z.append(z)
            V.append(v if with_noise else np.zeros((N, nx)))
            out[f"particles_{k}"] = st.particles.copy()
            out[f"weights_{k}"] = st.weights.copy()
            out[f"mean_{k}"] = st.mean.copy()
            out[f"cov_{k}"] = st.cov.copy()
            out[f"cond_{k}"] = np.array(st.diagnostics["condition_numbers"])
        out["Z"], out["V"] = np.array(Z), np.array(V)
        out["with_noise"] = np.array(with_noise)
        np.savez_compressed(os.path.join(HERE, f"ledh_{tag}.npz"), **out)
        print(f"ledh_{tag}: done, final ESS {1.0 / np.sum(st.weights ** 2):.1f}")

    run("linear1d", Linear1DModel(), N=64, L=8, T=6, ess_ratio=0.0, seed=11,
        mean0=np.array([0.5]), cov0=np.array([[0.3]]), x_true0=np.array([0.4]), with_noise=True)
    run("linear1d_resample", Linear1DModel(), N=50, L=4, T=6, ess_ratio=0.9, seed=12,
        mean0=np.array([0.5]), cov0=np.array([[0.3]]), x_true0=np.array([0.4]), with_noise=True)
    ac = AcousticCVModel(n_targets=1, grid=(3, 3))
    run("acoustic1", ac, N=40, L=6, T=4, ess_ratio=0.5, seed=13,
        mean0=np.array([12.0, 14.0, 0.5, -0.3]), cov0=np.diag([4.0, 4.0, 0.25, 0.25]),
        x_true0=np.array([12.5, 13.0, 0.6, -0.2]), with_noise=True)
    ac2 = AcousticCVModel(n_targets=2, grid=(4, 4))
    run("acoustic2", ac2, N=24, L=8, T=3, ess_ratio=0.0, seed=14,
        mean0=np.array([12.0, 14.0, 0.5, -0.3, 30.0, 25.0, -0.4, 0.2]),
        cov0=np.diag([4.0, 4.0, 0.25, 0.25] * 2),
        x_true0=np.array([12.5, 13.0, 0.6, -0.2, 29.0, 26.0, -0.5, 0.3]), with_noise=False)
    # BASELINE config 5's shape: 4 targets, 16-D joint state, 5 x 5 = 25 sensors, the article's process noise
    # (simulator_Multi_acoustic_tracking.py:104), R = 0.01 I, L = 8 -- pins the compile-time 16 x 25 flow kernel
    V1 = (1.0 / 20.0) * np.array([[1 / 3, 0, 0.5, 0], [0, 1 / 3, 0, 0.5], [0.5, 0, 1, 0], [0, 0.5, 0, 1.0]])
    ac4 = AcousticCVModel(n_targets=4, grid=(5, 5), Q1=V1)
    x4 = np.array([12.0, 6.0, 0.001, 0.001, 32.0, 32.0, -0.001, -0.005, 20.0, 13.0, -0.1, 0.01, 15.0, 35.0, 0.002, 0.002])
    run("acoustic4", ac4, N=64, L=8, T=3, ess_ratio=0.5, seed=15,
        mean0=x4 + np.tile([0.8, -0.6, 0.05, -0.05], 4), cov0=np.diag([1.0, 1.0, 0.05, 0.05] * 4),
        x_true0=x4, with_noise=True)


class ScriptedMeanTracker(ScriptedTracker):
    """ScriptedTracker that also replays scripted past means (EDH uses get_past_mean, :213)."""

    def __init__(self, Ps, means, nx):
        super().__init__(Ps, nx)
        self.means = means

    def get_past_mean(self):
        return self.means[self.k - 1]


class RecordingRng:
    """Wraps a Generator and records every ``random()`` (the resampling uniforms)."""

    def __init__(self, rng):
        self._r, self.us = rng, []

    def random(self):
        u = self._r.random()
        self.us.append(u)
        return u

    def __getattr__(self, name):
        return getattr(self._r, name)


def gen_edh():
    """EDHFlowPF steps of the unmodified reference (Euler and RK4 pseudo-time
    integrators, ESS-triggered resampling), scripted tracker covariances and past means."""
    def run(tag, model, *, N, L, T, ess_ratio, integrator, seed, mean0, cov0, x_true0):
        rng = np.random.default_rng(seed)
        nx, nz = model.nx, model.nz
        Ps, means = [], []
        Pk = np.array(cov0, float)
        x = np.asarray(x_true0, float)
        for k in range(T):
            G = rng.standard_normal((nx, nx)) * 0.02
            Pk = 0.8 * Pk + model.Q + G @ G.T
            Ps.append(Pk + 1e-9 * np.triu(np.ones((nx, nx)), 1))
            means.append(x + 0.1 * rng.standard_normal(nx))
            x = model.g(x, None, None)
        rec = RecordingRng(np.random.default_rng(seed + 1))
        cfg = ref_edh.EDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=ess_ratio,
                                flow_integrator=integrator, rng=rec)
        f = ref_edh.EDHFlowPF(ScriptedMeanTracker(Ps, means, nx), model.g, model.h, model.jac_h,
                              model.log_trans_pdf, model.log_like_pdf, model.R, cfg)
        st = f.init_from_gaussian(np.asarray(mean0, float), np.asarray(cov0, float))
        Lq = np.linalg.cholesky(model.Q)
        x = np.asarray(x_true0, float)
        out = dict(N=np.array(N), L=np.array(L), T=np.array(T), ess_ratio=np.array(ess_ratio),
                   integrator=np.array(integrator), seed=np.array(seed), Ps=np.array(Ps), means=np.array(means),
                   Q=model.Q, R=model.R, particles0=st.particles.copy(), weights0=st.weights.copy())
        Z, V, U = [], [], []
        for k in range(T):
            x = model.g(x, None, Lq @ rng.standard_normal(nx))
            z = model.h(x) + np.linalg.cholesky(model.R) @ rng.standard_normal(nz)
            v = rng.standard_normal((N, nx)) @ Lq.T
            n_before = len(rec.us)
            st = f.step(st, z, None, lambda n_, d_, v=v: v)
            U.append(rec.us[-1] if len(rec.us) > n_before else -1.0)
            Z.append(z)
            V.append(v)
            out[f"particles_{k}"] = st.particles.copy()
            out[f"weights_{k}"] = st.weights.copy()
            out[f"mean_{k}"] = st.mean.copy()
            out[f"cov_{k}"] = st.cov.copy()
            out[f"cond_{k}"] = np.array(st.diagnostics["condition_numbers"])
        out["Z"], out["V"], out["U"] = np.array(Z), np.array(V), np.array(U)
        np.savez_compressed(os.path.join(HERE, f"edh_{tag}.npz"), **out)
        print(f"edh_{tag}: resampled {int(np.sum(np.array(U) >= 0))}/{T}, final ESS {1.0 / np.sum(st.weights ** 2):.1f}")

    run("linear1d_euler", Linear1DModel(), N=64, L=8, T=6, ess_ratio=0.5, integrator="euler", seed=21,
        mean0=np.array([0.5]), cov0=np.array([[0.3]]), x_true0=np.array([0.4]))
    run("linear1d_rk4", Linear1DModel(), N=50, L=4, T=6, ess_ratio=0.9, integrator="rk4", seed=22,
        mean0=np.array([0.5]), cov0=np.array([[0.3]]), x_true0=np.array([0.4]))
    ac = AcousticCVModel(n_targets=1, grid=(3, 3))
    run("acoustic1_rk4", ac, N=40, L=6, T=4, ess_ratio=0.5, integrator="rk4", seed=23,
        mean0=np.array([12.0, 14.0, 0.5, -0.3]), cov0=np.diag([4.0, 4.0, 0.25, 0.25]),
        x_true0=np.array([12.5, 13.0, 0.6, -0.2]))


def gen_rng_checksum():
    r = np.random.default_rng(42)
    a = r.standard_normal((4, 3))
    b = r.random()
    c = r.standard_normal(5)
    np.savez(os.path.join(HERE, "rng_stream.npz"), a=a, b=np.array(b), c=c,
             numpy_version=np.array(np.__version__))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--ledh-only":      # regenerate selected LEDH cases only
        gen_ledh(only=set(sys.argv[2].split(",")))
        sys.exit(0)
    gen_resample()
    gen_sir()
    gen_l96()
    gen_ledh()
    gen_edh()
    gen_rng_checksum()
