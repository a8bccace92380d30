#!/usr/bin/env python
"""Per-step wall time, ESS and resampling flag of the C5 LEDH filter (class API) on the reference's acoustic scenario
(tests/golden/bench_c5_acoustic.npz).   python profiles/micro/c5_steps.py [L] [dtype]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import bench_configs as bc


def main():
    import torch
    from particle_filters_b200 import AcousticMeasurement, LEDHConfig, LEDHFlowPF
    from particle_filters_b200.ssm import LedhGaussianSSM
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dtype = sys.argv[2] if len(sys.argv) > 2 else "f32"
    N = 1 << (20 if dtype == "f32" else 17)
    g = bc.fixture("bench_c5_acoustic.npz")
    F, Q, R, S, Z, X = g["F"], g["Q"], g["R"], g["S"], g["Z"], g["X_true"]
    m0, P0 = g["mean0"], g["P0"]
    meas = AcousticMeasurement(S, float(g["psi"]), float(g["d0"]), n_targets=4)
    ssm = LedhGaussianSSM(F, Q, meas, R)
    pf = LEDHFlowPF(bc.EkfTracker(F, Q, R, ssm.h, ssm.jacobian_h, m0, P0), ssm.g, ssm.h, ssm.jacobian_h, ssm.log_trans_pdf,
                    ssm.log_like_pdf, R, LEDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=0.5,
                                                    rng=np.random.default_rng(0)), dtype=dtype)
    st = pf.init_from_gaussian(m0, P0)
    pos = [i for c in range(4) for i in (4 * c, 4 * c + 1)]
    for k in range(int(os.environ.get("C5_STEPS", "12"))):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = pf.step(st, Z[1 + k], process_noise_sampler="philox")
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        err = float(np.sqrt(np.mean((np.asarray(st.mean)[pos] - X[1 + k].reshape(-1)[pos]) ** 2)))
        print(f"step {k:2d} {dt:8.1f} ms  ess {pf.last_ess:12.2f} resampled {pf.last_resampled}  pos rmse {err:7.2f}  "
              f"mean finite {bool(np.all(np.isfinite(st.mean)))}  info-form (jittered, fell back) {counters()}", flush=True)


def counters():
    import ctypes as C
    from particle_filters_b200 import _lib
    a, b = C.c_uint64(0), C.c_uint64(0)
    _lib.load().smc_ledh_info_counters(C.byref(a), C.byref(b))
    return a.value, b.value


if __name__ == "__main__":
    main()
