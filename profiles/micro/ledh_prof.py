import sys, os, time, cProfile, pstats
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import bench_configs as bc
from particle_filters_b200 import AcousticMeasurement, LEDHConfig, LEDHFlowPF
from particle_filters_b200.ssm import LedhGaussianSSM
from ssm_zoo import AcousticCVModel
am = AcousticCVModel(n_targets=4, grid=(5, 5), q_scale=0.05)
rng = np.random.default_rng(56)
x = np.array([12.0, 6, 0.05, 0.1, 32, 32, -0.1, -0.05, 20, 13, 0.0, 0.1, 15, 35, 0.1, -0.1])
N = 1 << 20
P0 = np.kron(np.eye(4), np.diag([1.0, 1.0, 0.01, 0.01]))
meas = AcousticMeasurement(am.S, am.psi, am.d0, n_targets=4)
ssm = LedhGaussianSSM(am.F, am.Q, meas, am.R)
pf = LEDHFlowPF(bc._EkfTracker(am, x, P0), ssm.g, ssm.h, ssm.jacobian_h, ssm.log_trans_pdf, ssm.log_like_pdf, am.R,
                LEDHConfig(n_particles=N, n_lambda_steps=8, resample_ess_ratio=0.5, rng=np.random.default_rng(0)), dtype="f32")
st = pf.init_from_gaussian(x, P0)
z = am.h(am.F @ x)
st = pf.step(st, z, process_noise_sampler="philox")
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for k in range(3):
    st = pf.step(st, z, process_noise_sampler="philox")
torch.cuda.synchronize()
print("per step", (time.perf_counter() - t0) / 3)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
