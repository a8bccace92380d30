import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from particle_filters_b200 import StochVolSSM
from particle_filters_b200.engine import SirEngine
mid, params = StochVolSSM(0.95, 1.0).device_functor(np.array([[0.04]]), np.array([[1.0]]))
Z = np.random.default_rng(1).standard_normal((4, 1))
for N in (8401141, 12289, 4097, 8401140):
    ref = None
    bad = 0
    for rep in range(8):
        eng = SirEngine(N, 1, 1, dtype="f32", trace_len=8)
        eng.init_gaussian(np.zeros(1), np.array([[0.6]]), 11)
        snaps = []
        for k in range(4):
            eng.step(mid, params, 2.0, z_index=0, seed=11, t=k, row=k)
            snaps.append(eng.get_particles()[:, 0].copy())
        if ref is None:
            ref = snaps
        else:
            for k in range(4):
                d = np.flatnonzero(ref[k] != snaps[k])
                if d.size:
                    bad += 1
                    print("N", N, "rep", rep, "step", k, "diff idx", d[:8], "ref", ref[k][d[:4]], "got", snaps[k][d[:4]])
                    break
    print("N", N, "bad reps", bad)
