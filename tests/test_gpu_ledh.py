"""GPU parity tests of the LEDH particle-flow path (K4 + weight tail) against the
reference runs stored in tests/golden/ledh_*.npz and the NumPy oracle."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import ledh_weighted_stats
from oracle.ledh_numpy import ledh_flow
from parity import assert_ancestors
from ssm_zoo import AcousticCVModel, Linear1DModel, acoustic_c5_model

pytestmark = pytest.mark.gpu


class ScriptedTracker:
    """GaussianTracker protocol (reference LEDH_particle_filter.py:13-16) replaying (m, P)."""

    def __init__(self, Ps, nx):
        self.Ps, self.k, self.nx, self.updates = Ps, 0, nx, 0

    def predict(self):
        P = self.Ps[self.k]
        self.k += 1
        return np.zeros(self.nx), P

    def update(self, z):
        self.updates += 1
        return np.zeros(self.nx), self.Ps[self.k - 1]

    def get_past_mean(self):
        return np.zeros(self.nx)


def _cases():
    return [
        ("ledh_linear1d.npz", Linear1DModel(), 1e-8, "linear"),
        ("ledh_linear1d_resample.npz", Linear1DModel(), 1e-8, "linear"),
        ("ledh_acoustic1.npz", AcousticCVModel(n_targets=1, grid=(3, 3)), 1e-8, "acoustic"),
        ("ledh_acoustic2.npz", AcousticCVModel(n_targets=2, grid=(4, 4)), 2e-5, "acoustic"),
        # BASELINE config 5's shape: pins k4_ledh_flow_fast<16, 25> to the unmodified reference
        ("ledh_acoustic4.npz", acoustic_c5_model(), 2e-5, "acoustic"),
    ]


def _measurement_spec(model, kind):
    from particle_filters_b200 import AcousticMeasurement, LinearMeasurement

    if kind == "linear":
        return LinearMeasurement(np.array([[1.0]]))
    return AcousticMeasurement(model.S, model.psi, model.d0, n_targets=model.n_targets)


@pytest.mark.parametrize("mode", ["bridge", "device_h", "device_full"])
@pytest.mark.parametrize("name,model,tol,kind", _cases())
def test_ledh_step_tracks_reference(golden, name, model, tol, kind, mode):
    from particle_filters_b200.ledh_particle_filter import LEDHConfig, LEDHFlowPF, PFState

    g = golden(name)
    N, L, T = int(g["N"]), int(g["L"]), int(g["T"])
    rng = np.random.default_rng(int(g["seed"]) + 1)
    rng.multivariate_normal(np.zeros(model.nx), np.eye(model.nx), size=N)    # init draw of the reference run
    cfg = LEDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=float(g["ess_ratio"]), rng=rng)
    tracker = ScriptedTracker(g["Ps"], model.nx)
    if mode == "bridge":
        h, jh = (lambda x: model.h(x)), (lambda x: model.jac_h(x))
        f = LEDHFlowPF(tracker, model.g, h, jh, model.log_trans_pdf, model.log_like_pdf, model.R, cfg)
    elif mode == "device_h":
        spec = _measurement_spec(model, kind)
        f = LEDHFlowPF(tracker, model.g, spec.h, spec.jacobian_h, model.log_trans_pdf, model.log_like_pdf, model.R, cfg)
    else:   # every callback from one LedhGaussianSSM spec: the whole step stays on the GPU
        from particle_filters_b200.ssm import LedhGaussianSSM

        F = np.array([[model.alpha]]) if kind == "linear" else model.F
        full = LedhGaussianSSM(F, model.Q, _measurement_spec(model, kind), model.R)
        f = LEDHFlowPF(tracker, full.g, full.h, full.jacobian_h, full.log_trans_pdf, full.log_like_pdf, model.R, cfg)
    x, w = g["particles0"], g["weights0"]
    mean0, cov0 = ledh_weighted_stats(x, w)
    for k in range(T):
        st = PFState(particles=x, weights=w, mean=mean0, cov=cov0, diagnostics={})
        V = g["V"][k]
        sampler = (lambda n_, d_, V=V: V) if bool(g["with_noise"]) else None
        out = f.step(st, g["Z"][k], None, sampler)
        assert out.particles.shape == (N, model.nx) and out.weights.shape == (N,)
        np.testing.assert_allclose(np.sum(out.weights), 1.0, rtol=1e-10)
        np.testing.assert_allclose(out.particles, g[f"particles_{k}"], rtol=tol, atol=tol)
        np.testing.assert_allclose(out.weights, g[f"weights_{k}"], rtol=200 * tol, atol=1e-12)
        np.testing.assert_allclose(out.mean, g[f"mean_{k}"], rtol=tol, atol=tol)
        np.testing.assert_allclose(out.cov, g[f"cov_{k}"], rtol=20 * tol, atol=20 * tol)
        np.testing.assert_allclose(out.cov, out.cov.T, rtol=0, atol=0)
        np.testing.assert_allclose(out.diagnostics["condition_numbers"], g[f"cond_{k}"], rtol=1e-6)
        assert len(out.diagnostics["condition_numbers"]) == L
        # continue from the reference's state so one step's round-off cannot compound
        x, w = g[f"particles_{k}"], g[f"weights_{k}"]
        mean0, cov0 = g[f"mean_{k}"], g[f"cov_{k}"]
    assert tracker.updates == T


def test_flow_kernel_matches_oracle_flow_fp64_and_fp32():
    """K4 alone: eta and log-det sum vs the NumPy oracle flow on a 16-D / 25-sensor problem
    (BASELINE config 5 shape, small N)."""
    import ctypes as C

    from particle_filters_b200 import AcousticMeasurement, _lib
    from particle_filters_b200.ledh_particle_filter import _FlowBuffers

    m = AcousticCVModel(n_targets=4, grid=(5, 5), r_var=0.05)
    rng = np.random.default_rng(4)
    N, L = 257, 8
    centre = np.array([10, 12, .5, -.3, 30, 28, -.4, .2, 12, 30, .3, .3, 28, 10, -.2, .4])
    eta0 = centre + rng.standard_normal((N, 16)) * np.tile([1.5, 1.5, .3, .3], 4)
    P = np.kron(np.eye(4), np.diag([4.0, 4.0, 0.3, 0.3])) + 0.05
    P = 0.5 * (P + P.T)
    z = m.h(centre) + 0.1 * rng.standard_normal(25)
    eta_ref, theta_ref, cond_ref = ledh_flow(eta0, P, m.R, z, m.h_batch, m.jac_batch, L)
    spec = AcousticMeasurement(m.S, m.psi, m.d0, n_targets=4)
    kind, hp = spec.ledh_functor()
    lib = _lib.load()
    for dtype, tol in (("f64", 1e-8), ("f32", 5e-3)):
        buf = _FlowBuffers(N, 16, 25, L, dtype, None)
        buf.load(eta0)
        kP, pP = _lib.as_doubles(P)
        kR, pR = _lib.as_doubles(m.R)
        kz, pz = _lib.as_doubles(z)
        kh, ph = _lib.as_doubles(hp)
        _lib.check(lib.smc_ledh_flow(C.byref(buf.desc), pP, pR, pz, kind, ph, kh.size, buf.stream))
        eta = buf.eta_host()
        theta = buf.theta[:N].cpu().numpy().astype(np.float64)
        scale = np.abs(eta_ref).max()
        err = np.abs(eta - eta_ref)
        if dtype == "f64":
            assert err.max() <= tol * scale, (dtype, err.max())
            np.testing.assert_allclose(theta, theta_ref, rtol=0, atol=1e-8 * max(1.0, np.abs(theta_ref).max()))
        else:
            # fp32 flow: particles that pass close to a sensor see Jacobians ~ psi/d0^2 and a stiff
            # S; their fp32 trajectories are ill-conditioned, so the bar is on the bulk, not the max
            assert np.quantile(err, 0.95) <= tol * scale, (dtype, np.quantile(err, 0.95))
            assert np.median(np.abs(theta - theta_ref)) <= 1e-3 * max(1.0, np.abs(theta_ref).max())
        if dtype == "f64":
            cond = [np.linalg.cond(s) for s in buf.s0.cpu().numpy()]
            np.testing.assert_allclose(cond, cond_ref, rtol=1e-8)


def test_init_from_gaussian_and_weighted_stats():
    from particle_filters_b200.ledh_particle_filter import LEDHConfig, LEDHFlowPF

    m = Linear1DModel()

    def make(seed):
        cfg = LEDHConfig(n_particles=300, rng=np.random.default_rng(seed))
        return LEDHFlowPF(ScriptedTracker([np.eye(1)], 1), m.g, m.h, m.jac_h, m.log_trans_pdf, m.log_like_pdf, m.R, cfg)

    a = make(5).init_from_gaussian(np.array([0.5]), np.array([[0.3]]))
    b = make(5).init_from_gaussian(np.array([0.5]), np.array([[0.3]]))
    assert a.particles.shape == (300, 1) and a.diagnostics == {}
    np.testing.assert_array_equal(a.particles, b.particles)
    np.testing.assert_allclose(a.weights, 1.0 / 300, rtol=0, atol=0)
    mref, cref = ledh_weighted_stats(a.particles, a.weights)
    np.testing.assert_allclose(a.mean, mref, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(a.cov, cref, rtol=1e-10)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((77, 5)) + 3.0
    w = rng.random(77)
    mean, cov = LEDHFlowPF._weighted_stats(x, w)
    mref, cref = ledh_weighted_stats(x, w)
    np.testing.assert_allclose(mean, mref, rtol=1e-12)
    np.testing.assert_allclose(cov, cref, rtol=1e-10, atol=1e-14)


def test_module_level_resample_and_ess(golden):
    from particle_filters_b200.ledh_particle_filter import effective_sample_size, systematic_resample

    g = golden("resample_cases.npz")

    class U:
        def __init__(self, u):
            self.u = u

        def random(self):
            return self.u

    checked = 0
    for k in range(0, int(g["n_cases"]), 4):
        if not bool(g[f"c{k}_ledh_ok"]):
            continue
        w, u = g[f"c{k}_w"] * 3.7, float(g[f"c{k}_u"])          # un-normalised on purpose (:27)
        idx = systematic_resample(w, U(u))
        assert idx.dtype == np.int64 and idx.shape == w.shape
        wn = w / np.sum(w)
        assert_ancestors(idx, g[f"c{k}_ledh"], wn, u, clamp_last=False, msg=f"case {k}")
        np.testing.assert_allclose(effective_sample_size(w), 1.0 / np.sum(wn ** 2), rtol=1e-10)
        checked += 1
    assert checked >= 15


@pytest.mark.parametrize("N,nx,nz", [(1, 1, 1), (5, 3, 2), (64, 9, 4)])
def test_ledh_shapes_small_and_odd_dims(N, nx, nz):
    """Dimensions the reference tests use (nx in {1,3,9}, nz in {1,2,4}); linear-Gaussian model,
    bridge vs on-chip measurement functor must agree."""
    from particle_filters_b200 import LinearMeasurement
    from particle_filters_b200.ledh_particle_filter import LEDHConfig, LEDHFlowPF

    rng = np.random.default_rng(N + nx)
    H = rng.standard_normal((nz, nx))
    F = 0.9 * np.eye(nx)
    Q, R = 0.05 * np.eye(nx), 0.2 * np.eye(nz)
    P = np.eye(nx) * 0.5 + 0.05

    def g(x, u, v):
        return F @ x + (0 if v is None else v)

    def ltp(xk, xp):
        d = xk - F @ xp
        return float(-0.5 * d @ np.linalg.solve(Q, d))

    def llp(z, x):
        d = z - H @ x
        return float(-0.5 * d @ np.linalg.solve(R, d))

    outs = []
    for mode in ("bridge", "device"):
        cfg = LEDHConfig(n_particles=N, n_lambda_steps=4, resample_ess_ratio=0.5, rng=np.random.default_rng(1))
        spec = LinearMeasurement(H)
        h, jh = ((lambda x: H @ x), (lambda x: H)) if mode == "bridge" else (spec.h, spec.jacobian_h)
        f = LEDHFlowPF(ScriptedTracker([P, P], nx), g, h, jh, ltp, llp, R, cfg)
        st = f.init_from_gaussian(np.zeros(nx), np.eye(nx))
        z = rng.standard_normal(nz) if mode == "bridge" else z
        st = f.step(st, z, None, None)
        assert st.particles.shape == (N, nx) and np.all(np.isfinite(st.particles))
        np.testing.assert_allclose(np.sum(st.weights), 1.0, rtol=1e-10)
        assert st.cov.shape == (nx, nx)
        outs.append(st)
    np.testing.assert_allclose(outs[0].particles, outs[1].particles, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(outs[0].weights, outs[1].weights, rtol=1e-7, atol=1e-12)


def _c5_filter(dtype, N, L=8, seed=3):
    from particle_filters_b200 import AcousticMeasurement, LEDHConfig, LEDHFlowPF
    from particle_filters_b200.ssm import LedhGaussianSSM

    am = acoustic_c5_model()
    ssm = LedhGaussianSSM(am.F, am.Q, AcousticMeasurement(am.S, am.psi, am.d0, n_targets=4), am.R)
    P = np.kron(np.eye(4), np.diag([1.0, 1.0, 0.05, 0.05]))
    tracker = ScriptedTracker([P + am.Q] * 8, am.nx)
    f = LEDHFlowPF(tracker, ssm.g, ssm.h, ssm.jacobian_h, ssm.log_trans_pdf, ssm.log_like_pdf, am.R,
                   LEDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=0.0, rng=np.random.default_rng(seed)),
                   dtype=dtype)
    return f, am, ssm


def test_c5_shape_fp32_weights_track_fp64_weights():
    """BASELINE config 5's shape (16-D, 25 sensors, R = 0.01 I): the fp32 flow must not corrupt the WEIGHTS -- the
    Gaussian terms multiply errors of eta by Q^-1 and R^-1 = 100.  One device-resident step from the same
    particles and the same process noise in both precisions: log-weights of every particle that matters agree
    to 0.05 and the ESS to 10 %."""
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ledh_acoustic4.npz"))
    from particle_filters_b200.ledh_particle_filter import PFState

    N = 4096
    rng = np.random.default_rng(8)
    am = acoustic_c5_model()
    x_true = np.array([12.0, 6.0, 0.001, 0.001, 32.0, 32.0, -0.001, -0.005, 20.0, 13.0, -0.1, 0.01, 15.0, 35.0, 0.002, 0.002])
    x0 = x_true + rng.standard_normal((N, 16)) * np.tile([0.7, 0.7, 0.1, 0.1], 4)
    w0 = np.full(N, 1.0 / N)
    z = am.h(am.F @ x_true) + 0.1 * rng.standard_normal(am.nz)
    V = rng.standard_normal((N, 16)) @ np.linalg.cholesky(am.Q).T
    out = {}
    for dtype in ("f64", "f32"):
        f, _, _ = _c5_filter(dtype, N)
        mean0, cov0 = ledh_weighted_stats(x0, w0)
        st = f.step(PFState(particles=x0, weights=w0, mean=mean0, cov=cov0, diagnostics={}), z, None, lambda n_, d_: V)
        out[dtype] = (np.asarray(st.weights).copy(), np.asarray(st.particles).copy(), f.last_ess)
    w64, x64, ess64 = out["f64"]
    w32, x32, ess32 = out["f32"]
    assert abs(w64.sum() - 1) < 1e-9 and abs(w32.sum() - 1) < 1e-4, (w64.sum(), w32.sum())   # fp32 log-weights
    keep = w64 > 1e-6 * w64.max()                               # the particles that carry the estimate
    assert keep.sum() >= 5
    dlog = np.abs(np.log(w32[keep]) - np.log(w64[keep]))
    assert dlog.max() < 0.05, (dlog.max(), keep.sum())
    assert abs(ess32 - ess64) <= 0.1 * ess64 + 0.5, (ess32, ess64)
    scale = np.maximum(1.0, np.abs(x64))
    assert np.percentile(np.abs(x32 - x64) / scale, 99) < 2e-4
    del g


def test_device_ledh_state_is_generation_checked():
    """A device-resident LEDH state is readable only while it is the filter's latest one: an unread
    state of an earlier step must raise instead of returning the newer step's arrays; a state read in time
    keeps its own arrays."""
    N = 512
    f, am, _ = _c5_filter("f64", N)
    rng = np.random.default_rng(1)
    x_true = np.array([12.0, 6.0, 0.0, 0.0, 32.0, 32.0, 0.0, 0.0, 20.0, 13.0, 0.0, 0.0, 15.0, 35.0, 0.0, 0.0])
    st0 = f.init_from_gaussian(x_true, np.kron(np.eye(4), np.diag([0.5, 0.5, 0.01, 0.01])))
    z = am.h(am.F @ x_true)
    s1 = f.step(st0, z, None, "philox")
    p1 = s1.particles.copy()                    # read in time: cached
    s2 = f.step(s1, z, None, "philox")          # s1 was read -> re-uploaded, fine
    s3 = f.step(s2, z, None, "philox")          # s2 unread and current -> stays resident
    np.testing.assert_array_equal(s1.particles, p1)              # s1 still returns ITS arrays
    with pytest.raises(RuntimeError, match="stale"):
        _ = s2.particles                        # s2 was never read and the filter has moved on
    assert s3.particles.shape == (N, 16) and abs(s3.weights.sum() - 1) < 1e-9
    with pytest.raises(RuntimeError, match="stale"):
        f.step(s2, z, None, "philox")           # stepping from a lost state must not silently use newer arrays


@pytest.mark.parametrize("nx,nz,dtype,tol", [(48, 40, "f64", 1e-8), (64, 64, "f64", 1e-8), (144, 144, "f64", 1e-7),
                                             (64, 64, "f32", 5e-3)])
def test_flow_beyond_32_dimensions_uses_the_global_workspace(nx, nz, dtype, tol):
    """The reference runs LEDH on 64-, 144- and 400-dimensional sensor-network models
    (notebooks/PF_PF_results_reproduction_snlg.ipynb, ..._sn_skew.ipynb).  Above 32 dimensions the flow kernel keeps
    its per-particle matrices in a global-memory workspace; same arithmetic, compared with the NumPy oracle flow
    on a linear measurement model with a dense, well-conditioned H."""
    import ctypes as C

    import torch

    from particle_filters_b200 import _lib

    rng = np.random.default_rng(nx + nz)
    N, L = 37, 3
    H = rng.standard_normal((nz, nx)) / np.sqrt(nx)
    G = rng.standard_normal((nx, nx)) * 0.1
    P = 0.5 * np.eye(nx) + G @ G.T
    R = 0.3 * np.eye(nz) + 0.05 * np.diag(rng.random(nz))
    z = rng.standard_normal(nz)
    eta0 = rng.standard_normal((N, nx))
    eta_ref, theta_ref, cond_ref = ledh_flow(eta0, P, R, z, lambda X: X @ H.T, lambda X: np.broadcast_to(H, (len(X), nz, nx)), L)
    lib = _lib.load()
    td = torch.float64 if dtype == "f64" else torch.float32
    ld = (N + 3) // 4 * 4
    e0 = torch.zeros((nx, ld), dtype=td, device="cuda")
    e0[:, :N] = torch.from_numpy(np.ascontiguousarray(eta0.T)).to(td)
    e1 = e0.clone()
    th = torch.zeros((ld,), dtype=td, device="cuda")
    s0 = torch.zeros((L, nz, nz), dtype=torch.float64, device="cuda")
    desc = _lib.LedhDesc(n=N, ld=ld, nx=nx, nz=nz, dtype=_lib.SMC_F64 if dtype == "f64" else _lib.SMC_F32, n_lambda_steps=L,
                         eta0_dev=e0.data_ptr(), eta_dev=e1.data_ptr(), theta_log_dev=th.data_ptr(), s0_dev=s0.data_ptr())
    kP, pP = _lib.as_doubles(P)
    kR, pR = _lib.as_doubles(R)
    kz, pz = _lib.as_doubles(z)
    kh, ph = _lib.as_doubles(H)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.smc_ledh_flow(C.byref(desc), pP, pR, pz, _lib.SMC_LEDH_H_LINEAR, ph, kh.size, st))
    torch.cuda.synchronize()
    eta = e1[:, :N].double().cpu().numpy().T
    theta = th[:N].double().cpu().numpy()
    scale = np.maximum(1.0, np.abs(eta_ref))
    assert np.max(np.abs(eta - eta_ref) / scale) < tol
    assert np.max(np.abs(theta - theta_ref) / np.maximum(1.0, np.abs(theta_ref))) < 10 * tol
    cond = [float(np.linalg.cond(s0[j].cpu().numpy())) for j in range(L)]
    np.testing.assert_allclose(cond, cond_ref, rtol=1e-6 if dtype == "f64" else 1e-3)


def test_ledh_filter_runs_at_64_dimensions_through_the_class_api():
    """Drop-in check at a notebook-like size (nx = nz = 64, N = 200, host callbacks as the reference's notebooks pass
    them): one step against the NumPy oracle step."""
    from oracle.ledh_numpy import ledh_step
    from particle_filters_b200.ledh_particle_filter import LEDHConfig, LEDHFlowPF, PFState

    nx = nz = 64
    N, L = 200, 4
    rng = np.random.default_rng(5)
    A = 0.9 * np.eye(nx)
    Q, R = 0.2 * np.eye(nx), 0.5 * np.eye(nz)
    Hm = np.eye(nz) + 0.05 * rng.standard_normal((nz, nx))
    Qi, Ri = np.linalg.inv(Q), np.linalg.inv(R)

    def logn(d, Pi):
        return -0.5 * float(d @ Pi @ d)

    P = 0.6 * np.eye(nx)
    x0 = rng.standard_normal((N, nx))
    w0 = np.full(N, 1.0 / N)
    z = Hm @ (A @ x0[0]) + 0.3 * rng.standard_normal(nz)
    V = rng.standard_normal((N, nx)) * np.sqrt(0.2)
    f = LEDHFlowPF(ScriptedTracker([P], nx), lambda x, u=None, v=None: A @ x + (0.0 if v is None else v), lambda x: Hm @ x,
                   lambda x: Hm, lambda xk, xp: logn(xk - A @ xp, Qi), lambda zz, x: logn(zz - Hm @ x, Ri), R,
                   LEDHConfig(n_particles=N, n_lambda_steps=L, resample_ess_ratio=0.0, rng=np.random.default_rng(0)))
    mean0, cov0 = ledh_weighted_stats(x0, w0)
    out = f.step(PFState(particles=x0, weights=w0, mean=mean0, cov=cov0, diagnostics={}), z, None, lambda n_, d_: V)
    xo, wo, mo, co, info = ledh_step(
        x0, w0, z, P, R, g_batch=lambda X, u, Vv: X @ A.T + Vv, h_batch=lambda X: X @ Hm.T,
        jac_batch=lambda X: np.broadcast_to(Hm, (len(X), nz, nx)),
        log_trans_batch=lambda Xk, Xp: -0.5 * np.einsum("ni,ij,nj->n", Xk - Xp @ A.T, Qi, Xk - Xp @ A.T),
        log_like_batch=lambda zz, X: -0.5 * np.einsum("ni,ij,nj->n", zz - X @ Hm.T, Ri, zz - X @ Hm.T),
        n_lambda_steps=L, resample_ess_ratio=0.0, v=V, rng=np.random.default_rng(0))
    np.testing.assert_allclose(out.particles, xo, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(out.weights, wo, rtol=1e-6, atol=1e-14)
    np.testing.assert_allclose(out.mean, mo, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(out.cov, co, rtol=1e-7, atol=1e-8)
