#!/usr/bin/env python
"""SHA-256 of the Lorenz-96 particle set after a few fp32 Philox steps (to compare kernel variants bit for bit):
    python profiles/micro/l96_hash.py ; SMC_LIB_PATH=profiles/micro/variants/libsmc_X.so python profiles/micro/l96_hash.py"""
import hashlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from particle_filters_b200 import Lorenz96SSM
from particle_filters_b200.engine import SirEngine

for nx, n in ((40, 1 << 16), (8, (1 << 14) + 4)):
    spec = Lorenz96SSM(nx=nx, F=8.0, dt=0.01, n_sub=5, obs_stride=4)
    nz = spec.nz
    Q, R = 0.01 * np.eye(nx), np.eye(nz)
    mid, params = spec.device_functor(Q, R)
    eng = SirEngine(n, nx, nz, dtype="f32", trace_len=4)
    eng.init_gaussian(np.full(nx, 8.0), 0.1 * np.eye(nx), 7)
    Y = np.random.default_rng(1).standard_normal((3, nz)) + 8.0
    rows = eng.run(mid, params, Y, 2.0, 11)
    x = eng.get_particles()
    print(nx, n, hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()[:16], float(rows[-1, 4]), float(rows[-1, 0]))
