"""Parity comparators shared by the GPU tests.

Tie rule for resampling ancestors: the reference's CDF is
a sequential fp64 ``np.cumsum``; the device CDF is an exact fixed-point sum.  The
two can only disagree where an output position sits within round-off of a CDF
value.  A mismatch at output ``i`` between ancestors ``a`` and ``b`` is admissible
iff ``|pos_i - c_ref[min(a,b)]| <= tau`` and every CDF value between them lies
within ``tau`` of it too (they differ by 1 or by the length of a ~zero-weight run).
"""

from __future__ import annotations

import numpy as np

TAU = 2.0 ** -40


def ancestor_mismatches(got, ref, weights, u, *, clamp_last=True, tau=TAU):
    """Return (n_mismatch, n_inadmissible)."""
    got = np.asarray(got, dtype=np.int64)
    ref = np.asarray(ref, dtype=np.int64)
    assert got.shape == ref.shape
    bad = np.nonzero(got != ref)[0]
    if bad.size == 0:
        return 0, 0
    N = len(ref)
    c = np.cumsum(np.asarray(weights, dtype=np.float64))
    if clamp_last:
        c[-1] = 1.0
    pos = (u + bad) / N
    lo = np.minimum(got[bad], ref[bad])
    hi = np.maximum(got[bad], ref[bad])
    near = np.abs(pos - c[lo]) <= tau
    flat = (c[np.maximum(hi - 1, lo)] - c[lo]) <= tau
    return int(bad.size), int(np.sum(~(near & flat)))


def assert_ancestors(got, ref, weights, u, *, clamp_last=True, max_mismatch=None, msg=""):
    n_bad, n_inadm = ancestor_mismatches(got, ref, weights, u, clamp_last=clamp_last)
    assert n_inadm == 0, f"{msg}: {n_inadm} inadmissible ancestor mismatches (of {n_bad})"
    if max_mismatch is not None:
        assert n_bad <= max_mismatch, f"{msg}: {n_bad} tie mismatches > {max_mismatch}"
    return n_bad


def check_ancestor_invariants(anc, N):
    anc = np.asarray(anc)
    assert anc.shape == (N,)
    assert anc.min() >= 0 and anc.max() < N
    assert np.all(np.diff(anc) >= 0), "systematic ancestors must be non-decreasing"


# ---- NumPy Philox4x32-10 + Box-Muller, mirroring csrc/smc_common.cuh (tests only) ----
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32).copy() for v in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def philox_process_normals_f32(seed, n, t, dim=0, offset=0):
    """The N(0,1) draws K1 uses for dimension ``dim`` at step ``t`` (fp32 Box-Muller)."""
    nq = (n + 3) // 4
    quad = (np.arange(nq, dtype=np.uint64) + np.uint64(offset // 4))
    c0 = (quad & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (quad >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(nq, t, dtype=np.uint32)
    c3 = np.full(nq, 0 | (dim << 8), dtype=np.uint32)
    r = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    inv = np.float32(2.3283064365386963e-10)
    twopi = np.float32(6.2831853071795865)
    out = np.empty((nq, 4), dtype=np.float64)
    for p in range(2):
        u1 = r[2 * p].astype(np.float32) * inv + np.float32(0.5) * inv
        u2 = r[2 * p + 1].astype(np.float32) * (inv * twopi) + np.float32(0.5) * inv * twopi
        rad = np.sqrt(-2.0 * np.log(u1.astype(np.float64)))
        out[:, 2 * p] = rad * np.cos(u2.astype(np.float64))
        out[:, 2 * p + 1] = rad * np.sin(u2.astype(np.float64))
    return out.reshape(-1)[:n]


def philox_resample_uniform(seed, t):
    r = philox4x32_10([0], [0], [t], [1], seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    v = (np.uint64(r[0][0]) << np.uint64(21)) ^ (np.uint64(r[1][0]) >> np.uint64(11))
    return float(v) * 1.1102230246251565e-16
