import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from particle_filters_b200 import StochVolSSM
from particle_filters_b200.engine import SirEngine
N, seed = (1 << 23) + 4096 * 3 + 1237, 11
mid, params = StochVolSSM(0.95, 1.0).device_functor(np.array([[0.04]]), np.array([[1.0]]))
Z = np.array([[7.5], [-8.0], [0.3], [8.5], [-7.0], [9.0]])
for T in (1, 2, 3):
    res = []
    for fused in (False, True, False):
        eng = SirEngine(N, 1, 1, dtype="f32", trace_len=8)
        eng.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
        rows = eng.run(mid, params, Z[:T], 2.0, seed, fused=fused) if T > 1 or not fused else None
        if rows is None:
            continue
        res.append((rows, eng.get_particles()[:, 0]))
    a = res[0]
    for k, b in enumerate(res[1:]):
        print("T", T, "vs", k, "lognorm diff", np.abs(b[0][:, 0] - a[0][:, 0]), "ess", a[0][:, 1], b[0][:, 1],
              "same particles", np.mean(a[1] == b[1]), "sorted-same", np.mean(np.sort(a[1]) == np.sort(b[1])),
              "mean diff", np.abs(b[0][:, 4] - a[0][:, 4]))
