import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from particle_filters_b200 import StochVolSSM
from particle_filters_b200.engine import SirEngine
N, seed = (1 << 23) + 4096 * 3 + 1237, 11
mid, params = StochVolSSM(0.95, 1.0).device_functor(np.array([[0.04]]), np.array([[1.0]]))
Z = np.random.default_rng(1).standard_normal((6, 1))
for T in (2, 3, 4, 6):
    res = []
    for fused in (False, True, False, True):
        eng = SirEngine(N, 1, 1, dtype="f32", trace_len=8)
        eng.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
        rows = eng.run(mid, params, Z[:T], 2.0, seed, fused=fused)
        res.append((rows, eng.get_particles()[:, 0]))
    a = res[0]
    for k, b in enumerate(res[1:]):
        d = np.flatnonzero(a[1] != b[1])
        print("T", T, ["fused", "unfused2", "fused2"][k], "lognorm diff", np.abs(b[0][:, 0] - a[0][:, 0]).max(), "ndiff particles", d.size, d[:6],
              "ess diff", np.abs(b[0][:, 1] - a[0][:, 1]).max())
