#!/usr/bin/env python
"""Benchmark of the SIR particle-filter hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2n 26]

Workload (BASELINE.json configs[2], SURVEY.md section 8d "C3"): bootstrap SIR filter of
the 1-D stochastic-volatility model (alpha=0.95, sigma=0.2, beta=1), N = 2^26
particles, systematic resampling EVERY step, fp32 particles/log-weights, on-chip
Philox noise, synthetic observations.  One "step" = one full filter step
(propagate + log-likelihood + normaliser/ESS + resample + gather + mean/cov).

Prints ONE JSON line (rank 0).  `value` = particle-updates/s with everything
resident in HBM, timed with CUDA events; `e2e` = the same metric through the
public class API (`ParticleFilter.step(z)`, host observation in, host mean out,
every step); `roofline` = dominant kernel vs the measured HBM peak;
`cpu_baseline` = the NumPy oracle port of the reference step on the host.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

ALPHA, SIGMA, BETA = 0.95, 0.2, 1.0
METRIC = "particle-updates/s (SIR, SV model, N=2^26)"
UNIT = "particle-updates/s"


def sv_observations(T, seed=42):
    rng = np.random.default_rng(seed)
    x = np.empty(T)
    x[0] = rng.normal(0.0, SIGMA / np.sqrt(1 - ALPHA ** 2))
    v = rng.standard_normal(T)
    for t in range(1, T):
        x[t] = ALPHA * x[t - 1] + SIGMA * v[t]
    return BETA * np.exp(0.5 * x) * rng.standard_normal(T)


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [s.strip() for s in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle port of the reference step; the reference itself is pure Python + NumPy and
# does not travel to the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_sir_steps(n, steps, threads, seed=1):
    """Vectorised NumPy restatement of the reference SIR step (oracle/sir_numpy.py), chunked over
    `threads` host threads for the element-wise parts.  Returns seconds per step."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle.sir_numpy import systematic_ancestors

    rng = np.random.default_rng(seed)
    Y = sv_observations(steps + 1)
    x = rng.standard_normal(n) * (SIGMA / np.sqrt(1 - ALPHA ** 2))
    w = np.full(n, 1.0 / n)
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    pool = ThreadPoolExecutor(threads) if threads > 1 else None

    def part(k, z, noise, logw):
        a, b = bounds[k], bounds[k + 1]
        xs = ALPHA * x[a:b] + SIGMA * noise[a:b]                  # particle_filter.py:236-237
        d = (z - BETA * np.exp(0.5 * xs)) / np.sqrt(BETA ** 2 + 1e-12)   # :257-259
        x[a:b] = xs
        logw[a:b] = np.log(w[a:b] + 1e-300) - 0.5 * d * d         # :260-261
        return None

    t0 = time.perf_counter()
    for t in range(steps):
        noise = rng.standard_normal(n)
        logw = np.empty(n)
        if pool:
            list(pool.map(lambda k: part(k, Y[t], noise, logw), range(threads)))
        else:
            part(0, Y[t], noise, logw)
        m = logw.max()
        wn = np.exp(logw - (m + np.log(np.sum(np.exp(logw - m)))))    # :262-263
        idx = systematic_ancestors(wn, rng.random())                  # :146-171 (always resample)
        x[:] = x[idx]                                                 # :209
        w[:] = 1.0 / n                                                # :210
        mean = np.average(x, weights=w)                               # :266
        cov = np.cov(x, aweights=w, bias=True)                        # :267
    dt = (time.perf_counter() - t0) / steps
    if pool:
        pool.shutdown()
    return dt, float(mean), float(cov)


def run_reference_arm(args):
    """`--impl reference`: the CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    n_cpu = 1 << 22
    # warm-up + timed steps on a bounded sample of the workload
    cpu_sir_steps(n_cpu, max(1, min(args.warmup, 2)), threads)
    steps = max(1, min(args.steps, 8))
    dt, _, _ = cpu_sir_steps(n_cpu, steps, threads)
    value = n_cpu / dt
    sample = f"oracle NumPy port of the SIR step, N_cpu=2^22 particles x {steps} steps, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SV SIR bootstrap filter N=2^{args.log2n}, systematic resampling every step "
                               "(BASELINE.json configs[2] / SURVEY 8d C3)",
                   "n_particles": 1 << args.log2n, "nx": 1, "nz": 1, "resample": "every step",
                   "noise": "numpy Generator (PCG64)",
                   "sample": f"each step is timed on N_cpu = 2^22 particles with {threads} host threads "
                             "(particle-updates/s is flat in N for this path)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch

    from particle_filters_b200 import ParticleFilter, StochVolSSM, _lib
    from particle_filters_b200.engine import SirEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        from particle_filters_b200.sharded import run_sharded_bench

        return run_sharded_bench(args, sv_observations, measured_peak, ClockSampler, METRIC, UNIT)

    N = 1 << args.log2n
    K, W = args.steps, max(3, args.warmup)
    thresh = 2.0            # resample every step (bootstrap filter; SURVEY 8d headline)
    seed = 42
    Q, R = np.array([[SIGMA ** 2]]), np.array([[BETA ** 2]])
    spec = StochVolSSM(ALPHA, BETA)
    mid, params = spec.device_functor(Q, R)
    Y = sv_observations(W + K + 2)

    eng = SirEngine(N, 1, 1, dtype="f32", trace_len=W + K + 2)
    eng.init_gaussian(np.zeros(1), np.array([[SIGMA / np.sqrt(1 - ALPHA ** 2)]]), seed)
    eng.set_observations(Y.reshape(-1, 1))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def one_step(k):
        eng.step(mid, params, thresh, z_index=k, seed=seed, t=k, row=k)

    # warm-up on the SAME pipeline the timed region uses (allocations, function attributes, first launches)
    eng.enqueue(mid, params, W, thresh, seed, t0=0, z0=0, row0=0, fused=False if args.no_fuse else None)
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    # nvidia-smi needs a moment before its first sample; keep the GPU on the same workload meanwhile so
    # that every sample is taken under load (these steps are outside the timed region)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.4:
        eng.enqueue(mid, params, W, thresh, seed, t0=0, z0=0, row0=0, fused=False if args.no_fuse else None)
        torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    # K full filter steps, enqueued back to back: the engine's run() pipeline (for this bootstrap filter:
    # one weighting launch, K - 1 fused resample->propagate launches with their range pre-passes, one
    # closing resample; --no-fuse times the three-launch step instead)
    fused_used = eng.enqueue(mid, params, K, thresh, seed, t0=W, z0=W, row0=W, fused=False if args.no_fuse else None)
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    ms_step = ms_total / K

    # ---- per-kernel breakdown: the same steps again with CUDA events between the launch groups -------
    NB = 6
    if fused_used:
        phases = ("k3_resample_fused(+finish, +finalize)", "k3a_range_totals")
        acc = {p: 0.0 for p in phases}
        marks = []

        def mark(tag):
            e = ev()
            e.record()
            marks.append((tag, e))

        eng._run_fused(mid, params, NB + 1, thresh, seed, t0=W + K, z0=0, row0=0, events=mark)
        torch.cuda.synchronize()
        for (tag, e_a), (_, e_b) in zip(marks[:-1], marks[1:]):
            if tag == "fused":
                acc[phases[0]] += e_a.elapsed_time(e_b) / NB
            elif tag == "prepass":
                acc[phases[1]] += e_a.elapsed_time(e_b) / NB
        dom, alg_bytes = phases[0], 16.0 * N      # fused kernel: R logw, R x, W x, W logw (fp32, nx = 1)
        # dram__bytes_read + write of this kernel from the committed ncu --set full capture
        # (profiles/r1_ncu_full_summary.txt: 537.2 MB + 500.8 MB per launch at N = 2^26)
        traffic = 1038.0e6 * (N / float(1 << 26))
    else:
        phases = ("k1_propagate_weight+finalize", "k3a_range_totals", "k3_resample+finish")
        acc = {p: 0.0 for p in phases}
        for k in range(W + K, W + K + NB):   # extra steps reuse trace rows / observations via modulo
            evs = [ev() for _ in range(4)]
            kk = k % (W + K)
            evs[0].record()
            eng.step_weight(mid, params, thresh, z_index=kk, seed=seed, t=k, row=kk)
            evs[1].record()
            eng.step_resample(1, row=kk)
            evs[2].record()
            eng.step_resample(2, row=kk)
            evs[3].record()
            torch.cuda.synchronize()
            for i, p in enumerate(phases):
                acc[p] += evs[i].elapsed_time(evs[i + 1]) / NB
        dom = max((phases[0], phases[2]), key=lambda p: acc[p])
        alg_bytes = 12.0 * N          # K1: R x, W x, W logw ; K3: R logw, R x, W x   (fp32, nx = 1)
        # profiles/r1_ncu_full_summary.txt: K3 537 MB + 242 MB, K1 268 MB + 479 MB per launch at N = 2^26
        traffic = (779.0e6 if dom == phases[2] else 747.5e6) * (N / float(1 << 26))
    clk = clocks.stop()
    rows = eng.trace[: W + K].cpu().numpy()
    assert np.all(np.isfinite(rows[:, :5])), "non-finite filter statistics"
    assert np.all(rows[W:, 2] == 1.0), "bootstrap filter must resample every step"

    peak, peak_src = measured_peak()
    achieved = alg_bytes / (acc[dom] * 1e-3) / 1e9
    value = N / (ms_step * 1e-3)
    step_alg = 32.0               # SURVEY 8d closed form 16 nx + 16 bytes per particle-update

    # ---- end-to-end through the public class API: host z in, host mean/cov out, every step ---
    pf = ParticleFilter(spec.g, spec.h, Q, R, Np=N, resample_thresh=thresh, dtype="f32", noise="philox", seed=seed)
    pf._eng = eng                                     # reuse the resident 2^26-particle arrays
    pf._functor = (mid, params)
    pf.initialize(np.zeros(1), np.array([[SIGMA ** 2 / (1 - ALPHA ** 2)]]))
    for k in range(3):
        pf.step(np.array([Y[k]]))
    torch.cuda.synchronize()
    Ke = max(3, min(K, 20))
    t0 = time.perf_counter()
    for k in range(Ke):
        st = pf.step(np.array([Y[3 + k]]))
        _ = float(st.mean[0])
    torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / Ke
    e2e = {"value": N / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": 8 * 1,
           "d2h_bytes_per_step": 8 * int(eng.stride), "ms_per_step": e2e_dt * 1e3,
           "api": "ParticleFilter.step(z) -> PFState.mean (host z in, host moments out, sync every step)"}

    # ---- CPU baseline: oracle port on the host, bounded sample --------------------------------
    n_cpu = 1 << (12 if args.no_cpu else 22)
    cpu_steps = 4
    dt_cpu, _, _ = cpu_sir_steps(n_cpu, cpu_steps, 1)
    cpu = {"value": n_cpu / dt_cpu, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"oracle NumPy port of the SIR step (vectorised; the reference itself is ~1.9e5/s), "
                     f"N_cpu=2^22 x {cpu_steps} steps, 1 thread of {os.cpu_count()}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SV SIR bootstrap filter N=2^{args.log2n}, systematic resampling every step "
                               "(BASELINE.json configs[2] / SURVEY 8d C3)",
                   "n_particles": N, "nx": 1, "nz": 1, "resample": "every step", "noise": "philox4x32-10 on chip",
                   "l2": "inputs (2 x 256 MiB per step) exceed the 126 MB L2; no flush needed"},
        "clocks": clk, "e2e": e2e, "gpu_launches": (2 * K + 1) if fused_used else 3 * K,   # fused: K pre-passes + K-1 fused + K1 + K3; else K1, pre-pass, K3 per step
        "pipeline": ("fused resample->propagate (2 launches per step, 20 B per particle-update)" if fused_used
                     else "three launches per step (28 B per particle-update)"),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_unit": "B per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                     "peak_source": peak_src,
                     "alg_bytes_per_launch": alg_bytes, "kernel_ms": acc[dom]},
        "roofline_step": {"alg_bytes_per_particle_update": step_alg,
                          "achieved": step_alg * value / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": step_alg * value / 1e9 / peak,
                          "note": "SURVEY 8d closed form 16 nx + 16 B; the three-launch step moves 28 B (12 K1 + 4 pre-pass + 12 K3), the fused run() pipeline 20 B"},
        "kernel_ms": acc, "cpu_baseline": cpu,
        "filter": {"mean_last": float(rows[-1, 4]), "ess_last": float(rows[-1, 1])},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log2n", type=int, default=26)
    ap.add_argument("--exchange", default="mailbox", choices=("mailbox", "nccl"),
                    help="multi-GPU record exchange: in-kernel NVLink peer stores (default) or NCCL all-gathers")
    ap.add_argument("--no-fuse", action="store_true", help="time the three-launch step instead of the fused run() pipeline")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (kernel tuning sweeps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
