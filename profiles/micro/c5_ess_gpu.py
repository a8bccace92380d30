#!/usr/bin/env python
"""GPU side of profiles/micro/c5_ess_cpu.py: pre-resample ESS / N of the C5 LEDH filter on the same data,
dtype f64 and f32, several N (host N(0,Q) sampler seeded like the CPU script for N <= 2^14, Philox above)."""
import argparse, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "profiles", "micro"))
from c5_ess_cpu import Ekf, scenario


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="look_alike")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--L", type=int, default=8)
    a = ap.parse_args()
    from particle_filters_b200 import AcousticMeasurement, LEDHConfig, LEDHFlowPF
    from particle_filters_b200.ssm import LedhGaussianSSM
    am, X, Z, m0, P0 = scenario(a.scenario, a.steps)
    meas = AcousticMeasurement(am.S, am.psi, am.d0, n_targets=am.n_targets)
    ssm = LedhGaussianSSM(am.F, am.Q, meas, am.R)
    print(f"# scenario {a.scenario}, L={a.L}: pre-resample ESS / N per step (GPU drop-in)")
    for dtype, N in (("f64", 4096), ("f32", 4096), ("f64", 16384), ("f32", 16384), ("f64", 1 << 18), ("f32", 1 << 18),
                     ("f32", 1 << 20)):
        rng = np.random.default_rng(0)
        pf = LEDHFlowPF(Ekf(am, m0, P0), ssm.g, ssm.h, ssm.jacobian_h, ssm.log_trans_pdf, ssm.log_like_pdf, am.R,
                        LEDHConfig(n_particles=N, n_lambda_steps=a.L, resample_ess_ratio=0.5, rng=rng), dtype=dtype)
        st = pf.init_from_gaussian(m0, P0)
        sampler = ssm.sampler(rng) if N <= 16384 else "philox"
        ess = []
        for k in range(a.steps):
            st = pf.step(st, Z[k], process_noise_sampler=sampler)
            ess.append(pf.last_ess / N)
        print(f"GPU {dtype} N={N:8d}                     " + " ".join(f"{e:8.2e}" for e in ess) + f"   | median {np.median(ess):.2e}")


if __name__ == "__main__":
    main()
