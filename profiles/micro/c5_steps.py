#!/usr/bin/env python
"""Per-step wall time of the C5 LEDH filter (class API) with SM clock / power samples: is the step time stationary?"""
import os, subprocess, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import bench_configs as bc


def main():
    import torch
    from particle_filters_b200 import AcousticMeasurement, LEDHConfig, LEDHFlowPF
    from particle_filters_b200.ssm import LedhGaussianSSM
    from ssm_zoo import AcousticCVModel
    K, N, L = 24, 1 << 20, 8
    am = AcousticCVModel(n_targets=4, grid=(5, 5), q_scale=0.05)
    rng = np.random.default_rng(56)
    x = np.array([12.0, 6, 0.05, 0.1, 32, 32, -0.1, -0.05, 20, 13, 0.0, 0.1, 15, 35, 0.1, -0.1])
    Lq, r = np.linalg.cholesky(am.Q), np.sqrt(am.R[0, 0])
    Z = np.empty((K, am.nz))
    x0 = None
    for t in range(K):
        x = am.F @ x + Lq @ rng.standard_normal(16)
        if t == 0:
            x0 = x.copy()
        Z[t] = am.h(x) + r * rng.standard_normal(am.nz)
    P0 = np.kron(np.eye(4), np.diag([1.0, 1.0, 0.01, 0.01]))
    meas = AcousticMeasurement(am.S, am.psi, am.d0, n_targets=4)
    ssm = LedhGaussianSSM(am.F, am.Q, meas, am.R)
    m0 = np.linalg.solve(am.F, x0)
    pf = LEDHFlowPF(bc._EkfTracker(am, m0, P0), ssm.g, ssm.h, ssm.jacobian_h, ssm.log_trans_pdf, ssm.log_like_pdf, am.R,
                    LEDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=0.5, rng=np.random.default_rng(0)),
                    dtype="f32")
    st = pf.init_from_gaussian(m0, P0)
    for k in range(K):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = pf.step(st, Z[k], process_noise_sampler="philox")
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                           capture_output=True, text=True).stdout.strip()
        print(f"step {k:2d} {dt:8.1f} ms  ess {pf.last_ess:10.1f} resampled {pf.last_resampled}  | {q}", flush=True)


if __name__ == "__main__":
    main()
