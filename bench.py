#!/usr/bin/env python
"""Benchmark of the SIR particle-filter hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2n 26]

Workload (BASELINE.json configs[2], SURVEY.md section 8d "C3"): bootstrap SIR filter of
the 1-D stochastic-volatility model (alpha=0.95, sigma=0.2, beta=1), N = 2^26
particles, systematic resampling EVERY step, fp32 particles/log-weights, on-chip
Philox noise; observations = the reference simulator's series (seed 42, fixture
tests/golden/bench_c1_sv.npz).  One "step" = one full filter step
(propagate + log-likelihood + normaliser/ESS + resample + gather + mean/cov).

Prints ONE JSON line (rank 0):
  value         particle-updates/s with everything resident in HBM, CUDA events over K steps
  e2e           the same metric through the public class API (`ParticleFilter.step(z)`: host
                observation in, host mean/cov out, one synchronisation per step)
  roofline      the dominant kernel against the measured HBM peak, on the bytes it moves; `traffic`
                is the ncu DRAM figure of the same kernel read from profiles/ncu_traffic.json
  cpu_baseline  the UNMODIFIED reference (`models/particle_filter.py` from baseline/_ref) on one host
                core, a bounded sample; the vectorised NumPy oracle port beside it
  parity        in-run checks at full size (offspring counts of systematic resampling at N = 2^26)
  configs       the other BASELINE configurations (bench_configs.py), each with its own roofline,
                check and reference CPU leg

`--impl reference` times the reference's own CPU implementation of the path on the host cores.
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
for _p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "particle-updates/s (SIR, SV model, N=2^26)"
UNIT = "particle-updates/s"


def sv_fixture():
    g = np.load(os.path.join(REPO, "tests", "golden", "bench_c1_sv.npz"))
    return float(g["alpha"]), float(g["sigma"]), float(g["beta"]), np.asarray(g["Y"], float)


ALPHA, SIGMA, BETA, _Y_REF = sv_fixture()


def sv_observations(T, seed=42):
    """The reference simulator's observation series (simulate_sv_1d, seed 42), tiled when a run is longer."""
    reps = -(-T // len(_Y_REF))
    return np.tile(_Y_REF, reps)[:T]


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key, n):
    """DRAM bytes per launch of `kernel_key` from the committed ncu capture (profiles/ncu_traffic.json),
    scaled to n particles; None when the file does not list the kernel."""
    p = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    e = t.get(kernel_key)
    if not e:
        return None, None
    return e["dram_bytes_per_launch"] * (n / float(e["n_particles"])), e.get("source")


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
# CPU arms
# ------------------------------------------------------------------------------------------------
def cpu_port_steps(n, steps, threads, seed=1):
    """Vectorised NumPy restatement of the reference SIR step (oracle/sir_numpy.py), the element-wise parts
    chunked over `threads` host threads.  Returns seconds per step (a 'best-effort CPU' figure: the reference
    itself is a per-particle Python loop)."""
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
        np.average(x, weights=w)                                      # :266
        np.cov(x, aweights=w, bias=True)                              # :267
    dt = (time.perf_counter() - t0) / steps
    if pool:
        pool.shutdown()
    return dt


def _reference_worker(args):
    """One replica of the UNMODIFIED reference filter (runs in its own process)."""
    n, steps, seed = args
    import warnings

    warnings.simplefilter("ignore")
    import bench_configs as bc

    ref = bc.reference_module("models/particle_filter.py")
    Y = sv_observations(steps + 2)
    pf = ref.ParticleFilter(lambda x, u=None: np.array([ALPHA * x[0]]), lambda x: np.array([BETA * np.exp(0.5 * x[0])]),
                            np.array([[SIGMA ** 2]]), np.array([[BETA ** 2]]), Np=n, resample_thresh=2.0,
                            rng=np.random.default_rng(seed))
    pf.initialize(np.zeros(1), np.array([[SIGMA ** 2 / (1 - ALPHA ** 2)]]))
    pf.step(np.array([Y[0]]))                     # warm-up
    t0 = time.perf_counter()
    for k in range(steps):
        st = pf.step(np.array([Y[1 + k]]))
    return time.perf_counter() - t0, float(st.mean[0])


def cpu_reference(n, steps, procs):
    """particle-updates/s of `procs` independent replicas of the unmodified reference (one per process; the
    reference itself is single-threaded Python), or None when no reference tree is staged."""
    import bench_configs as bc

    if bc.reference_module("models/particle_filter.py") is None:
        return None
    if procs <= 1:
        dt, _ = _reference_worker((n, steps, 1))
        return n * steps / dt
    import multiprocessing as mp

    # one replica per core: keep each replica's BLAS single-threaded (np.linalg.solve with N right-hand sides
    # would otherwise spawn a thread team per process and the replicas would fight for the cores)
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(procs) as pool:
            res = pool.map(_reference_worker, [(n, steps, 1 + r) for r in range(procs)])
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    wall_inner = max(r[0] for r in res)           # the slowest replica's filter loop (process start-up excluded)
    return procs * n * steps / wall_inner


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    n_cpu = 1 << 14
    steps = max(2, min(args.steps, 200))
    value = cpu_reference(n_cpu, steps, procs)
    kind = "reference"
    if value is None:                              # no reference tree on this box: the NumPy oracle port
        kind, procs = "port", max(1, min(cores, 32))
        n_cpu = 1 << 22
        cpu_port_steps(n_cpu, 1, procs)
        steps = max(1, min(args.steps, 8))
        value = n_cpu / cpu_port_steps(n_cpu, steps, procs)
        sample = f"oracle NumPy port of the SIR step, N_cpu=2^22 particles x {steps} steps, {procs} threads"
    else:
        sample = (f"UNMODIFIED models/particle_filter.py (baseline/_ref), {procs} independent replicas (one process per "
                  f"core; the reference is single-threaded), each N_cpu=2^14 particles x {steps} steps, SV model")
    ms = 1e3 * (1 << args.log2n) / value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SV SIR bootstrap filter N=2^{args.log2n}, systematic resampling every step "
                               "(BASELINE.json configs[2] / SURVEY 8d C3)",
                   "n_particles": 1 << args.log2n, "nx": 1, "nz": 1, "resample": "every step",
                   "noise": "numpy Generator (PCG64)",
                   "sample": sample + "; particle-updates/s of this path is flat in N, ms_per_step is the "
                                      "extrapolation to the full population"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def offspring_parity(N, mid, params, seed):
    """Size-independent property of systematic resampling at the FULL benchmark size, fp32 mode: ancestors are
    monotone and in range, the gather moved the right payload, and every particle's offspring count is within
    one of N w_j (the globally last particle also takes the `cumsum[-1] = 1` remainder).  Checked with torch
    ops on the device (checker only)."""
    import torch

    from particle_filters_b200 import _lib
    from particle_filters_b200.engine import SirEngine

    eng = SirEngine(N, 1, 1, dtype="f32", trace_len=2, write_ancestors=True)
    eng.init_gaussian(np.zeros(1), np.array([[SIGMA / np.sqrt(1 - ALPHA ** 2)]]), seed)
    eng.set_observations(np.array([[0.4]]))
    eng.propagate_weight(mid, params, _lib.SMC_MODE_STEP, z_index=0, seed=seed, t=0)
    eng.finalize(2.0, u=-1.0, seed=seed, t=0, row=0)
    w = torch.empty((N,), dtype=torch.float64, device=eng.device)
    import ctypes as C

    _lib.check(eng.lib.smc_sir_export_weights(eng._d, C.c_void_p(w.data_ptr()), eng.stream))
    x_pre = eng.x[eng.current_buffer()][0, :N].clone()
    eng.resample(row=0)
    eng.finish_step(row=0)
    anc = eng.ancestors[:N].to(torch.int64)
    x_post = eng.x[eng.current_buffer()][0, :N]
    counts = torch.bincount(anc, minlength=N).to(torch.float64)
    dev = (counts - N * w).abs()
    slack = 1.0 + 1e-6 * N * w + 1e-3
    slack[N - 8:] += 4.0
    out = {"n": N, "monotone": bool((anc[1:] >= anc[:-1]).all().item()),
           "in_range": bool((anc[0] >= 0).item() and (anc[-1] < N).item()),
           "payload_matches_ancestors": bool(torch.equal(x_post, x_pre[anc])),
           "max_abs_count_minus_Nw": float(dev.max().item()),
           "counts_within_one_of_Nw": bool((dev <= slack).all().item())}
    out["pass"] = bool(out["monotone"] and out["in_range"] and out["payload_matches_ancestors"]
                       and out["counts_within_one_of_Nw"])
    del eng, w, x_pre, anc, counts, dev, slack
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch

    from particle_filters_b200 import ParticleFilter, StochVolSSM, _lib
    from particle_filters_b200.engine import SirEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        from particle_filters_b200.sharded import run_sharded_bench

        return run_sharded_bench(args, sv_observations, measured_peak, ClockSampler, METRIC, UNIT)

    lib = _lib.load()
    N = 1 << args.log2n
    K, W = args.steps, max(3, args.warmup)
    thresh = 2.0            # resample every step (bootstrap filter; SURVEY 8d headline)
    seed = 42
    Q, R = np.array([[SIGMA ** 2]]), np.array([[BETA ** 2]])
    spec = StochVolSSM(ALPHA, BETA)
    mid, params = spec.device_functor(Q, R)
    Y = sv_observations(W + K + 2)
    fused_arg = False if args.no_fuse else None

    eng = SirEngine(N, 1, 1, dtype="f32", trace_len=W + K + 2)
    eng.init_gaussian(np.zeros(1), np.array([[SIGMA / np.sqrt(1 - ALPHA ** 2)]]), seed)
    eng.set_observations(Y.reshape(-1, 1))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # warm-up on the SAME pipeline the timed region uses (allocations, function attributes, first launches)
    eng.enqueue(mid, params, W, thresh, seed, t0=0, z0=0, row0=0, fused=fused_arg)
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    # nvidia-smi samples every 100 ms: keep the GPU on the same workload for ~1.2 s before the timed region so
    # that the clock record holds >= 10 samples under this load (these steps are outside the timed region)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 1.2:
        eng.enqueue(mid, params, W, thresh, seed, t0=0, z0=0, row0=0, fused=fused_arg)
        torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    launches0 = int(lib.smc_launch_count())
    e0.record()
    # K full filter steps, enqueued back to back: the engine's run() pipeline (for this bootstrap filter:
    # one weighting launch, K - 1 fused resample->propagate launches, each followed by its conditional
    # range pre-pass, one closing resample; --no-fuse times the three-launch step instead)
    fused_used = eng.enqueue(mid, params, K, thresh, seed, t0=W, z0=W, row0=W, fused=fused_arg)
    e1.record()
    torch.cuda.synchronize()
    launches = int(lib.smc_launch_count()) - launches0
    ms_total = e0.elapsed_time(e1)
    ms_step = ms_total / K

    # ---- per-kernel breakdown: the same steps again with CUDA events between the launch groups -------
    NB = 6
    legacy = os.environ.get("SMC_FUSED_LEGACY", "") == "1"
    if fused_used:
        k_fused = ("k3_resample<fused functor> (round-1 pipeline)" if legacy
                   else "k5_resample_propagate (resample t-1 -> propagate + weigh t, +finish, +finalize)")
        phases = (k_fused, "k3a_range_totals (conditional)")
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
        traffic, traffic_src = ncu_traffic("k3_resample_fused" if legacy else "k5_resample_propagate", N)
        step_bytes = 20.0 if legacy else 16.0
    else:
        phases = ("k1_propagate_weight+finalize(+range totals)", "k3a_range_totals (conditional)",
                  "k5_resample_propagate<CopyModel>+finish")
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
        alg_bytes = 12.0 * N          # K1: R x, W x, W logw ; resampler: R logw, R x, W x   (fp32, nx = 1)
        traffic, traffic_src = ncu_traffic("k5_copy" if dom == phases[2] else "k1_propagate_weight", N)
        step_bytes = 24.0
    clk = clocks.stop()
    rows = eng.trace[: W + K].cpu().numpy()
    assert np.all(np.isfinite(rows[:, :5])), "non-finite filter statistics"
    assert np.all(rows[W:, 2] == 1.0), "bootstrap filter must resample every step"

    peak, peak_src = measured_peak()
    achieved = alg_bytes / (acc[dom] * 1e-3) / 1e9
    value = N / (ms_step * 1e-3)

    # ---- end-to-end through the public class API: host z in, host mean/cov out, every step ---
    pf = ParticleFilter(spec.g, spec.h, Q, R, Np=N, resample_thresh=thresh, dtype="f32", noise="philox", seed=seed,
                        engine=eng)                      # adopt the resident 2^26-particle arrays
    pf.initialize(np.zeros(1), np.array([[SIGMA ** 2 / (1 - ALPHA ** 2)]]))
    for k in range(3):
        pf.step(np.array([Y[k]]))
    torch.cuda.synchronize()
    Ke = max(3, min(K, 50))
    t0 = time.perf_counter()
    for k in range(Ke):
        st = pf.step(np.array([Y[3 + k]]))
        _ = float(st.mean[0])
    torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / Ke
    e2e = {"value": N / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": 8 * 1,
           "d2h_bytes_per_step": 8 * int(eng.stride), "ms_per_step": e2e_dt * 1e3, "steps": Ke,
           "api": "ParticleFilter.step(z) -> PFState.mean (host z in, host moments out, sync every step; "
                  "three launches: K1 + range totals, conditional pre-pass, expansion resampler)"}
    del pf

    # ---- in-run parity at full size ---------------------------------------------------------------
    parity = None
    if not args.no_parity:
        del eng
        torch.cuda.empty_cache()
        parity = {"offspring_counts_fp32": offspring_parity(N, mid, params, 5)}

    # ---- CPU baseline: the unmodified reference on one host core (bounded sample) + the NumPy port ----
    cpu = None
    if not args.no_cpu:
        n_ref, s_ref = 1 << 12, 5
        v_ref = cpu_reference(n_ref, s_ref, 1)
        n_port = 1 << 22
        dt_port = cpu_port_steps(n_port, 3, 1)
        port = {"value": n_port / dt_port, "unit": UNIT, "cores": 1,
                "sample": f"vectorised NumPy oracle port, N_cpu=2^22 x 3 steps, 1 thread of {os.cpu_count()}"}
        if v_ref is not None:
            cpu = {"value": v_ref, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"UNMODIFIED models/particle_filter.py (baseline/_ref), N_cpu=2^12 x {s_ref} steps on 1 of "
                             f"{os.cpu_count()} host cores (the reference is single-threaded Python)",
                   "port": port}
        else:
            cpu = dict(port, kind="port", note="no reference tree on this box (baseline/_ref not staged)")

    # ---- the other BASELINE configurations ----------------------------------------------------------
    configs = None
    if args.configs:
        import bench_configs as bc

        torch.cuda.empty_cache()
        configs = bc.run_all([c.strip() for c in args.configs.split(",") if c.strip()], min(K, 20), cpu=not args.no_cpu)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SV SIR bootstrap filter N=2^{args.log2n}, systematic resampling every step "
                               "(BASELINE.json configs[2] / SURVEY 8d C3)",
                   "n_particles": N, "nx": 1, "nz": 1, "resample": "every step", "noise": "philox4x32-10 on chip",
                   "observations": "reference simulate_sv_1d(alpha=0.95, sigma=0.2, beta=1, seed=42), tests/golden/bench_c1_sv.npz",
                   "l2": "inputs (2 x 256 MiB per step) exceed the 126 MB L2; no flush needed"},
        "clocks": clk, "e2e": e2e, "gpu_launches": launches,
        "pipeline": (("round-1 fused pipeline (2 launches, 20 B per particle-update)" if legacy else
                      "fused resample->propagate K5 (1 launch + 1 conditional no-op launch per step, 16 B per particle-update)")
                     if fused_used else "three launches per step (24 B per particle-update)"),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "traffic_unit": "B per launch, ncu dram__bytes_read.sum + dram__bytes_write.sum",
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "alg_bytes_per_launch": alg_bytes, "kernel_ms": acc[dom]},
        "roofline_step": {"bytes_moved_per_particle_update": step_bytes, "achieved": step_bytes * value / 1e9,
                          "peak": peak, "unit": "GB/s", "frac": step_bytes * value / 1e9 / peak,
                          "contract_bytes_per_particle_update": 32.0, "contract_frac": 32.0 * value / 1e9 / peak,
                          "note": "frac is on the bytes the timed pipeline really moves; contract_* uses SURVEY 8d's closed "
                                  "form 16 nx + 16 B (an unfused three-pass step), which fusion beats"},
        "kernel_ms": acc, "cpu_baseline": cpu, "parity": parity,
        "filter": {"mean_last": float(rows[-1, 4]), "ess_last": float(rows[-1, 1])},
        "configs": configs,
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs (kernel tuning sweeps)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run full-size parity checks")
    ap.add_argument("--configs", default="c2,c3ess,c3f64,c4,c5,c5l64",
                    help="other BASELINE configurations to embed (bench_configs.py); '' = none")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
