#!/usr/bin/env python
"""Timing lines for the other BASELINE.json configurations (bench.py measures the headline C3):

    python bench_configs.py [--configs c2,c3ess,c4,c5] [--steps K]

  c2  2-D LGSSM SIR filter vs the exact Kalman filter, N = 2^24 (README A/B/C/D model)
  c3ess  C3 (SV, N = 2^26) with ESS-triggered resampling (resample_thresh = 0.5)
  c4  Lorenz-96 (d = 40, 20 RK4 sub-steps, every 4th variable observed), ESS-triggered resampling, N = 2^22
  c5  LEDH particle flow, 4-target acoustic tracking (16-D state, 25 sensors), N = 2^20

One JSON line per configuration: particle-updates/s over K device-resident steps (CUDA events),
the roofline that binds it, and a correctness figure (error against Kalman / truth).
Synthetic data from the models in tests/ssm_zoo.py (the reference simulators are not on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def peak_hbm():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, torch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def run_c2(K, log2n=24):
    import torch
    from particle_filters_b200 import LinearGaussianSSM, ParticleFilter
    from ssm_zoo import LGSSMModel

    m = LGSSMModel()
    X, Y = m.simulate(K + 5, seed=42)
    N = 1 << log2n
    spec = LinearGaussianSSM(m.A, m.C)
    pf = ParticleFilter(spec.g, spec.h, m.Q, m.R, Np=N, resample_thresh=2.0, dtype="f32", noise="philox", seed=42)
    pf.initialize(np.zeros(2), np.eye(2))
    pf.run(Y[:5])
    out, ms = timed(lambda: pf.run(Y[5:]), torch)
    # exact Kalman filter (predict-then-update, same prior convention as the PF)
    mk, P = np.zeros(2), np.eye(2)
    km = []
    for t in range(K + 5):
        mk, P = m.A @ mk, m.A @ P @ m.A.T + m.Q
        S = m.C @ P @ m.C.T + m.R
        G = P @ m.C.T @ np.linalg.inv(S)
        mk, P = mk + G @ (Y[t] - m.C @ mk), (np.eye(2) - G @ m.C) @ P
        km.append(mk.copy())
    err = float(np.max(np.abs(out["mean"] - np.array(km)[5:])))
    v = N * K / (ms * 1e-3)
    alg = 16 * 2 + 16
    return {"config": "c2", "metric": "particle-updates/s (SIR, 2-D LGSSM, N=2^%d)" % log2n, "value": v, "ms_per_step": ms / K,
            "steps": K, "dtype": "f32", "roofline": {"bound": "hbm", "alg_bytes_per_particle_update": alg,
                                                      "achieved": alg * v / 1e9, "peak": peak_hbm(), "unit": "GB/s",
                                                      "frac": alg * v / 1e9 / peak_hbm()},
            "check": {"max_abs_mean_error_vs_kalman": err, "posterior_std": float(np.sqrt(P[0, 0]))}}


def run_c3ess(K, log2n=26):
    """C3's model and size with the reference's default trigger (resample_thresh = 0.5): the device decides
    per step; B_alg is the mix of 32 B (resampling step) and 16 B."""
    import torch
    from particle_filters_b200 import ParticleFilter, StochVolSSM

    alpha, sigma, beta = 0.95, 0.2, 1.0
    rng = np.random.default_rng(42)
    x, Y = rng.standard_normal() * sigma / np.sqrt(1 - alpha ** 2), np.empty(K + 5)
    for t in range(K + 5):
        x = alpha * x + sigma * rng.standard_normal()
        Y[t] = beta * np.exp(0.5 * x) * rng.standard_normal()
    N = 1 << log2n
    spec = StochVolSSM(alpha, beta)
    pf = ParticleFilter(spec.g, spec.h, np.array([[sigma ** 2]]), np.array([[beta ** 2]]), Np=N, resample_thresh=0.5,
                        dtype="f32", noise="philox", seed=42)
    pf.initialize(np.zeros(1), np.array([[sigma ** 2 / (1 - alpha ** 2)]]))
    pf.run(Y[:5].reshape(-1, 1))
    out, ms = timed(lambda: pf.run(Y[5:].reshape(-1, 1)), torch)
    v = N * K / (ms * 1e-3)
    f = float(np.mean(out["resampled"]))
    alg = 32 * f + 16 * (1 - f)
    return {"config": "c3ess", "metric": "particle-updates/s (SIR, SV model, N=2^%d, resample_thresh=0.5)" % log2n,
            "value": v, "ms_per_step": ms / K, "steps": K, "dtype": "f32",
            "roofline": {"bound": "hbm", "alg_bytes_per_particle_update": alg, "achieved": alg * v / 1e9,
                         "peak": peak_hbm(), "unit": "GB/s", "frac": alg * v / 1e9 / peak_hbm()},
            "check": {"resampled_fraction": f, "ess_over_n_min": float(np.min(out["ess"]) / N),
                      "mean_finite": bool(np.isfinite(out["mean"]).all())}}


def run_c4(K, log2n=22):
    import torch
    from particle_filters_b200 import Lorenz96SSM, ParticleFilter
    from ssm_zoo import Lorenz96Model

    m = Lorenz96Model(nx=40, n_sub=20, obs_stride=4, q=0.1, r=1.0)
    X, Y = m.simulate(K + 3, seed=42)
    N = 1 << log2n
    spec = Lorenz96SSM(nx=40, n_sub=20, obs_stride=4)
    pf = ParticleFilter(spec.g, spec.h, m.Q, m.R, Np=N, resample_thresh=0.5, dtype="f32", noise="philox", seed=42)
    pf.initialize(X[0] + 0.3, 0.5 * np.eye(40))
    pf.run(Y[1:3])
    out, ms = timed(lambda: pf.run(Y[3:]), torch)
    rmse = np.sqrt(np.mean((out["mean"] - X[3:]) ** 2, axis=1))
    v = N * K / (ms * 1e-3)
    flop = 20 * 4 * 40 * 6 + 40 * 30          # RK4 stencils (+ noise, likelihood), per particle-update
    return {"config": "c4", "metric": "particle-updates/s (SIR, Lorenz-96 d=40, N=2^%d)" % log2n, "value": v,
            "ms_per_step": ms / K, "steps": K, "dtype": "f32",
            "roofline": {"bound": "fp32 FMA issue", "flop_per_particle_update": flop, "achieved": flop * v / 1e12,
                         "peak": 148 * 128 * 2 * 1.965e9 / 1e12, "unit": "TFLOP/s",
                         "frac": flop * v / 1e12 / (148 * 128 * 2 * 1.965e9 / 1e12),
                         "hbm_bytes_per_particle_update": 656},
            "check": {"rmse_first": float(rmse[0]), "rmse_median": float(np.median(rmse)),
                      "resampled_steps": int(np.sum(out["resampled"]))}}


class _EkfTracker:
    """Minimal EKF with the reference's GaussianTracker protocol (LEDH_particle_filter.py:13-16):
    predict() -> (m, P), update(z) -> (m, P), get_past_mean()."""

    def __init__(self, model, m0, P0):
        self.am, self.m, self.P = model, np.asarray(m0, float).copy(), np.asarray(P0, float).copy()
        self.m_past = self.m.copy()

    def predict(self):
        am = self.am
        self.m_past = self.m.copy()
        self.m = am.F @ self.m
        self.P = am.F @ self.P @ am.F.T + am.Q
        return self.m.copy(), self.P.copy()

    def update(self, z):
        am = self.am
        H = am.jac_h(self.m)
        S = H @ self.P @ H.T + am.R
        G = np.linalg.solve(S, H @ self.P).T
        self.m = self.m + G @ (np.asarray(z, float) - am.h(self.m))
        self.P = (np.eye(am.nx) - G @ H) @ self.P
        self.P = 0.5 * (self.P + self.P.T)
        return self.m.copy(), self.P.copy()

    def get_past_mean(self):
        return self.m_past.copy()


def run_c5(K, log2n=20, L=8):
    import torch
    from particle_filters_b200 import AcousticMeasurement, LEDHConfig, LEDHFlowPF
    from particle_filters_b200.ssm import LedhGaussianSSM
    from ssm_zoo import AcousticCVModel

    am = AcousticCVModel(n_targets=4, grid=(5, 5), q_scale=0.05)
    rng = np.random.default_rng(56)
    x = np.array([12.0, 6, 0.05, 0.1, 32, 32, -0.1, -0.05, 20, 13, 0.0, 0.1, 15, 35, 0.1, -0.1])
    Lq, r = np.linalg.cholesky(am.Q), np.sqrt(am.R[0, 0])
    X, Z = np.empty((K + 2, 16)), np.empty((K + 2, am.nz))
    for t in range(K + 2):
        x = am.F @ x + Lq @ rng.standard_normal(16)
        X[t], Z[t] = x, am.h(x) + r * rng.standard_normal(am.nz)
    N = 1 << log2n
    P0 = np.kron(np.eye(4), np.diag([1.0, 1.0, 0.01, 0.01]))
    meas = AcousticMeasurement(am.S, am.psi, am.d0, n_targets=4)
    ssm = LedhGaussianSSM(am.F, am.Q, meas, am.R)
    m0 = np.linalg.solve(am.F, X[0])
    pf = LEDHFlowPF(_EkfTracker(am, m0, P0), ssm.g, ssm.h, ssm.jacobian_h, ssm.log_trans_pdf, ssm.log_like_pdf, am.R,
                    LEDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=0.5, rng=np.random.default_rng(0)),
                    dtype="f32")
    st = pf.init_from_gaussian(m0, P0)
    for k in range(2):
        st = pf.step(st, Z[k], process_noise_sampler="philox")

    def loop():
        s = st
        for k in range(2, K + 2):
            s = pf.step(s, Z[k], process_noise_sampler="philox")
        return s

    s, ms = timed(loop, torch)
    v = N * K / (ms * 1e-3)
    flop = 1.07e5 * L
    pos = [0, 1, 4, 5, 8, 9, 12, 13]
    err = float(np.sqrt(np.mean((np.asarray(s.mean)[pos] - X[K + 1][pos]) ** 2)))
    peak = 148 * 128 * 2 * 1.965e9 / 1e12
    return {"config": "c5", "metric": "particle-updates/s (LEDH flow, 4-target acoustic, N=2^%d, L=%d)" % (log2n, L),
            "value": v, "ms_per_step": ms / K, "steps": K, "dtype": "f32 flow, f64 weights",
            "roofline": {"bound": "fp32 FMA issue", "flop_per_particle_update": flop, "achieved": flop * v / 1e12,
                         "peak": peak, "unit": "TFLOP/s", "frac": flop * v / 1e12 / peak,
                         "note": "SURVEY 8a literal reference operation count; includes the host EKF and the "
                                 "per-step host synchronisation of the class API"},
            "check": {"position_rmse_last": err, "ess_last": getattr(pf, "last_ess", None)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c4")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--log2n", type=int, default=0)
    args = ap.parse_args()
    for c in args.configs.split(","):
        fn = {"c2": run_c2, "c3ess": run_c3ess, "c4": run_c4, "c5": run_c5}[c.strip()]
        kw = {"log2n": args.log2n} if args.log2n else {}
        print(json.dumps(fn(args.steps, **kw)), flush=True)


if __name__ == "__main__":
    main()
