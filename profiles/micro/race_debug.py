import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from particle_filters_b200 import StochVolSSM
from particle_filters_b200.engine import SirEngine
N, seed = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 17) + 1237, 11
mid, params = StochVolSSM(0.95, 1.0).device_functor(np.array([[0.04]]), np.array([[1.0]]))
Z = np.array([[7.5], [-8.0], [0.3], [8.5]]) if (len(sys.argv) < 3 or sys.argv[2] == "skew") else np.random.default_rng(1).standard_normal((4, 1))
res = []
for rep in range(3):
    eng = SirEngine(N, 1, 1, dtype="f32", trace_len=8)
    eng.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
    rows = eng.run(mid, params, Z, 2.0, seed, fused=False)
    res.append((rows, eng.get_particles()[:, 0]))
for k in (1, 2):
    d = np.flatnonzero(res[0][1] != res[k][1])
    print("rep", k, "differing particles", d.size, d[:10], "lognorm diff", np.abs(res[0][0][:, 0] - res[k][0][:, 0]).max(), "ess", res[0][0][:, 1])
