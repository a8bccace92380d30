import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from particle_filters_b200 import LinearGaussianSSM
from particle_filters_b200.engine import SirEngine
from ssm_zoo import LGSSMModel
N, seed = (1 << 23) + 4096 * 3 + 1237, 11
m = LGSSMModel()
mid, params = LinearGaussianSSM(m.A, m.C).device_functor(m.Q, m.R)
Z = m.simulate(6, seed=3)[1]
for T in (5, 6):
    res = []
    for fused in (False, True, False, True):
        eng = SirEngine(N, 2, 2, dtype="f32", trace_len=8)
        eng.init_gaussian(np.zeros(2), np.eye(2), seed)
        rows = eng.run(mid, params, Z[:T], 2.0, seed, fused=fused)
        res.append((rows, eng.get_particles()))
    a = res[0]
    for k, b in enumerate(res[1:]):
        d = np.flatnonzero(np.any(a[1] != b[1], axis=1))
        print("T", T, ["fused", "unfused2", "fused2"][k], "lognorm diff", np.abs(b[0][:, 0] - a[0][:, 0]).max(), "ndiff", d.size, d[:6],
              "ess", a[0][:T, 1].round(1))
