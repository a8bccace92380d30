#!/usr/bin/env python
"""Is the C5 LEDH filter (16-D, 25 sensors, R = 0.01 I) really as degenerate as profiles/r1_c5_steps.txt
shows (pre-resample ESS 1..357 of 2^20), or does the fp32 flow corrupt the weights?  CPU answer: the same
data through (a) the UNMODIFIED reference LEDHFlowPF (Python loops, fp64) at small N and (b) the batched
fp64 oracle at larger N.  ESS/N of a correct filter is roughly independent of N.

    python profiles/micro/c5_ess_cpu.py [--scenario look_alike|article] [--steps K] [--n-ref 200] [--n-oracle 4096]
"""
import argparse, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "tools"))


def scenario(name, K):
    from ssm_zoo import AcousticCVModel
    if name == "look_alike":      # bench_configs.run_c5 of round 1
        am = AcousticCVModel(n_targets=4, grid=(5, 5), q_scale=0.05)
        rng = np.random.default_rng(56)
        x = np.array([12.0, 6, 0.05, 0.1, 32, 32, -0.1, -0.05, 20, 13, 0.0, 0.1, 15, 35, 0.1, -0.1])
        Lq, r = np.linalg.cholesky(am.Q), np.sqrt(am.R[0, 0])
        X, Z = np.empty((K, 16)), np.empty((K, am.nz))
        for t in range(K):
            x = am.F @ x + Lq @ rng.standard_normal(16)
            X[t], Z[t] = x, am.h(x) + r * rng.standard_normal(am.nz)
        P0 = np.kron(np.eye(4), np.diag([1.0, 1.0, 0.01, 0.01]))
        m0 = np.linalg.solve(am.F, X[0])
        return am, X, Z, m0, P0
    if name == "article":        # tests/golden/bench_c5_acoustic.npz: the reference simulator + notebook model
        g = np.load(os.path.join(REPO, "tests", "golden", "bench_c5_acoustic.npz"))
        am = AcousticCVModel(n_targets=4, psi=float(g["psi"]), d0=float(g["d0"]), Q1=g["Q"][:4, :4], sensors=g["S"],
                             r_var=float(g["R"][0, 0]))
        K = min(K, g["Z"].shape[0] - 1)
        return am, g["X_true"][1:K + 1].reshape(K, 16), g["Z"][1:K + 1], g["mean0"], g["P0"]
    raise SystemExit("unknown scenario")


class Ekf:
    def __init__(self, am, m0, P0):
        self.am, self.m, self.P = am, np.array(m0, float), np.array(P0, float)
    def predict(self):
        am = self.am
        self.m_past = self.m.copy()
        self.m = am.F @ self.m; self.P = am.F @ self.P @ am.F.T + am.Q
        return self.m.copy(), self.P.copy()
    def update(self, z):
        am = self.am
        H = am.jac_h(self.m); S = H @ self.P @ H.T + am.R
        G = np.linalg.solve(S, H @ self.P).T
        self.m = self.m + G @ (z - am.h(self.m)); self.P = (np.eye(am.nx) - G @ H) @ self.P
        self.P = 0.5 * (self.P + self.P.T)
        return self.m.copy(), self.P.copy()
    def get_past_mean(self):
        return self.m_past.copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="look_alike")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--n-ref", type=int, default=200)
    ap.add_argument("--n-oracle", type=int, default=4096)
    ap.add_argument("--L", type=int, default=8)
    a = ap.parse_args()
    am, X, Z, m0, P0 = scenario(a.scenario, a.steps)
    Lq = np.linalg.cholesky(am.Q)
    from oracle.ledh_numpy import ledh_step
    from stage_reference import reference_root
    root = reference_root()
    rows = []
    # (b) oracle, fp64, batched
    for N in (a.n_oracle,):
        rng = np.random.default_rng(0)
        trk = Ekf(am, m0, P0)
        x = m0 + rng.multivariate_normal(np.zeros(16), P0, size=N)
        w = np.full(N, 1.0 / N)
        ess = []
        for k in range(a.steps):
            _, P = trk.predict()
            v = rng.standard_normal((N, 16)) @ Lq.T
            x, w, mean, cov, info = ledh_step(x, w, Z[k], P, am.R, g_batch=am.g_batch, h_batch=am.h_batch,
                                              jac_batch=am.jac_batch, log_trans_batch=am.log_trans_batch,
                                              log_like_batch=am.log_like_batch, n_lambda_steps=a.L,
                                              resample_ess_ratio=0.5, v=v, rng=rng)
            trk.update(Z[k])
            ess.append(info["ess_pre"] / N)
        rows.append((f"oracle fp64 N={N}", ess))
    # (a) the unmodified reference
    if root:
        sys.path.insert(0, root)
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_ledh", os.path.join(root, "models", "LEDH_particle_filter.py"))
        ref = importlib.util.module_from_spec(spec); sys.modules["ref_ledh"] = ref; spec.loader.exec_module(ref)
        N = a.n_ref
        rng = np.random.default_rng(0)
        trk = Ekf(am, m0, P0)
        pf = ref.LEDHFlowPF(trk, am.g, am.h, am.jac_h, am.log_trans_pdf, am.log_like_pdf, am.R,
                            ref.LEDHConfig(n_particles=N, n_lambda_steps=a.L, resample_ess_ratio=0.0, rng=rng))
        st = pf.init_from_gaussian(m0, P0)
        ess = []
        t0 = time.time()
        for k in range(a.steps):
            st = pf.step(st, Z[k], process_noise_sampler=lambda n, nx: rng.standard_normal((n, nx)) @ Lq.T)
            wn = st.weights / st.weights.sum()
            e = 1.0 / np.sum(wn ** 2)
            ess.append(e / N)
            if e < 0.5 * N:     # resample as the reference does (:201-206), outside it so that the ESS is observable
                idx = ref.systematic_resample(wn, rng)
                st.particles = st.particles[idx]; st.weights = np.full(N, 1.0 / N)
        rows.append((f"reference (unmodified) N={N}, {time.time() - t0:.0f} s", ess))
    print(f"# scenario {a.scenario}, L={a.L}: pre-resample ESS / N per step")
    for name, ess in rows:
        print(f"{name:42s} " + " ".join(f"{e:8.2e}" for e in ess) + f"   | median {np.median(ess):.2e}")


if __name__ == "__main__":
    main()
