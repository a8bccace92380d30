// K4 information form: the LEDH particle flow in nx x nx algebra (fp32 throughput mode).
//
// The reference (models/LEDH_particle_filter.py:149-165) builds, per particle and pseudo-time step, the
// nz x nz matrix S = lam H P H^T + R and A = -1/2 P H^T S^-1 H.  With P = L L^T (one Cholesky per step for ALL
// particles, on the host) and whitened coordinates xi = L^-1 eta, Hh = H L, the push-through identity
//   (lam H P H^T + R)^-1 H = R^-1 H (I + lam P G)^-1,   G = H^T R^-1 H
// turns the same flow into SYMMETRIC nx x nx algebra:
//   Gh = Hh^T R^-1 Hh,  wh = Hh^T R^-1 (z - e),  M1 = I + lam Gh  (SPD),
//   Ah = L^-1 A L = -1/2 Gh M1^-1,
//   d xi / d lam = Ah xi + (I + 2 lam Ah)[(I + lam Ah) wh + Ah xi0] = p + Ah (xi + 2 lam p),  p = wh + Ah (lam wh + xi0),
//   log|det(I + dlam A)| = log det(I + (lam - dlam/2) Gh) - log det(I + lam Gh)     (both SPD: no pivoting, no sign),
// i.e. two Cholesky factorisations of nx x nx matrices and two products with Ah per step instead of an nz x nz
// Cholesky, nx triangular solves of length nz and a pivoted LU: ~11 K multiply-adds per particle-substep at nx = 16,
// nz = 25 against ~53 K.  Same flow as the literal kernels (k4_ledh_flow / k4_ledh_flow_fast, which keep
// the reference's operation order and serve the fp64 parity mode); results agree to fp32 round-off of either route.
// Requirements: R diagonal, nx <= 16, one of the compiled measurement functors.
//
// Layout: LP = 4 / 8 / 16 lanes per particle (32 / LP particles per warp), lane j = row j of the symmetric
// matrices and entry j of the whitened vectors; rows other lanes need travel by width-LP shuffles or through a
// small per-particle shared scratch read as broadcasts.
#pragma once

namespace smc {

constexpr int kLedhInfoWarps = 4;
constexpr int kLedhInfoMaxNz = 64;

template <int NX> struct LedhInfoLanes { static constexpr int value = NX <= 4 ? 4 : (NX <= 8 ? 8 : 16); };

// floats of shared memory: CTA constants, then per-particle scratch
template <int NX> struct LedhInfoSmem {
    static constexpr int LP = LedhInfoLanes<NX>::value;
    __host__ __device__ static constexpr int pad4(int v) { return (v + 3) / 4 * 4; }
    // every array starts on a 16-byte boundary (rows are read as float4 broadcasts)
    __host__ __device__ static constexpr int kPerParticle(int nz) {
        return NX /*eta*/ + nz * NX /*H rows*/ + pad4(nz) /*innov / r*/ + nz * NX /*Hh rows*/ + pad4(NX * (NX + 1)) /*Cholesky factor*/
               + 2 * NX;
    }
    static size_t bytes(int nz, int n_hp) {
        const int n_const = pad4(NX * (NX + 1)) /*L rows, padded*/ + 2 * pad4(nz) /*rinv, z*/ + pad4(n_hp);
        return sizeof(float) * (size_t)(n_const + kLedhInfoWarps * (32 / LP) * kPerParticle(nz));
    }
};

struct LedhInfoArgs {
    const float* eta0;
    float* eta;
    float* theta;
    double* s0;                 // condition-number diagnostic: S of particle 0 per step (nullable)
    const double* Lc;           // chol(P), lower, nx x nx
    const double* R;            // nz x nz (diagonal used; the full matrix only for the diagnostic)
    const double* z;
    const double* hp;
    int64_t n, ld;
    int nz, n_hp, L, step0;
};

template <int NX, int KIND>
__global__ void __launch_bounds__(32 * kLedhInfoWarps) k4_ledh_flow_info(LedhInfoArgs a) {
    constexpr int LP = LedhInfoLanes<NX>::value, PPW = 32 / LP;       // lanes per particle, particles per warp
    constexpr int LS = NX + 1;                                        // padded row stride (column reads by consecutive lanes)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw);
    const int nz = a.nz;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = lane % LP, sub = lane / LP;                         // my row, my particle within the warp
    const unsigned gmask = 0xffffffffu;
    // ---- CTA constants ------------------------------------------------------------------------------------
    using SM = LedhInfoSmem<NX>;
    static_assert(NX % 4 == 0, "rows are read as float4");
    float* sL = sm;            sm += SM::pad4(NX * LS);   // L[k][c] at k * LS + c
    float* srinv = sm;         sm += SM::pad4(nz);
    float* sz = sm;            sm += SM::pad4(nz);
    float* shp = sm;           sm += SM::pad4(a.n_hp);
    for (int i = tid; i < NX * NX; i += blockDim.x) sL[(i / NX) * LS + i % NX] = (float)a.Lc[i];
    for (int i = tid; i < nz; i += blockDim.x) { srinv[i] = (float)(1.0 / a.R[(size_t)i * nz + i]); sz[i] = (float)a.z[i]; }
    for (int i = tid; i < a.n_hp; i += blockDim.x) shp[i] = (float)a.hp[i];
    __syncthreads();
    // ---- per-particle scratch -----------------------------------------------------------------------------
    float* w = sm + (warp * PPW + sub) * LedhInfoSmem<NX>::kPerParticle(nz);
    float* seta = w;           w += NX;
    float* sH = w;             w += nz * NX;              // Jacobian rows
    float* sinn = w;           w += SM::pad4(nz);         // (z - e)_s / r_s
    float* sHh = w;            w += nz * NX;              // whitened Jacobian rows Hh = H L
    float* sC = w;             w += SM::pad4(NX * LS);    // Cholesky factor of M1 (rows), for the backward solve

    const int64_t i = ((int64_t)blockIdx.x * kLedhInfoWarps + warp) * PPW + sub;
    const bool live = i < a.n && j < NX;
    const int jx = j < NX ? j : NX - 1;
    // (particles beyond n and lanes beyond nx mirror valid ones: every lane runs every shuffle)
    const int64_t ic = i < a.n ? i : a.n - 1;

    // column j of L (for Hh = H L) in registers
    float Lcol[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) Lcol[k] = sL[k * LS + jx];

    auto bcast = [&](float v, int src) -> float { return __shfl_sync(gmask, v, src, LP); };
    // xi = L^-1 v (forward substitution, lane j ends with entry j)
    auto solve_L = [&](float v) -> float {
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            const float tk = bcast(v, k) / sL[k * LS + k];
            if (j == k) v = tk;
            else if (j > k) v = fmaf(-sL[jx * LS + k], tk, v);
        }
        return v;
    };
    float xi = solve_L(a.eta[(int64_t)jx * a.ld + ic]);
    const float xi0 = solve_L(a.eta0[(int64_t)jx * a.ld + ic]);
    float theta = a.theta[ic];

    for (int step = 0; step < a.L; ++step) {
        const float dlam = 1.0f / (float)a.L;
        float lam;
        {   // lam = min(1, lam + dlam) accumulated exactly as the reference does (:136-137)
            double l = 0.0; const double dl = 1.0 / (double)a.L;
            for (int q = 0; q <= step; ++q) l = fmin(1.0, l + dl);
            lam = (float)l;
        }
        // ---- 1. eta = L xi ----------------------------------------------------------------------------------
        float eta_j = 0.f;
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            const float xc = bcast(xi, c);
            if (c <= jx) eta_j = fmaf(sL[jx * LS + c], xc, eta_j);
        }
        if (j < NX) seta[j] = eta_j;
        __syncwarp();
        // ---- 2. sensors: Jacobian row, residual (:143-145); lanes share the sensors ---------------------------
        for (int s = j; s < nz; s += LP) {
            float hx = 0.f, he = 0.f;                      // h_s(eta), H_s . eta
            float* Hr = sH + s * NX;
            if (KIND == SMC_LEDH_H_LINEAR) {
#pragma unroll
                for (int t = 0; t < NX; ++t) { const float hv = shp[s * NX + t]; Hr[t] = hv; hx = fmaf(hv, seta[t], hx); }
                he = hx;
            } else if (KIND == SMC_LEDH_H_ACOUSTIC) {
                const float psi = shp[0], d0 = shp[1], sx = shp[2 + 2 * s], sy = shp[3 + 2 * s];
#pragma unroll
                for (int c = 0; c < NX / 4; ++c) {
                    const float dx = seta[4 * c] - sx, dy = seta[4 * c + 1] - sy;
                    const float den = dx * dx + dy * dy + d0;
                    const float inv = 1.0f / den;
                    hx = fmaf(psi, inv, hx);
                    const float coef = -2.0f * psi * inv * inv;
                    const float hxc = coef * dx, hyc = coef * dy;
                    Hr[4 * c] = hxc; Hr[4 * c + 1] = hyc; Hr[4 * c + 2] = 0.f; Hr[4 * c + 3] = 0.f;
                    he = fmaf(hxc, seta[4 * c], fmaf(hyc, seta[4 * c + 1], he));
                }
            } else {   // SMC_LEDH_H_EXP: h_s = beta exp(x_s / 2), nz == nx
                const float v = shp[0] * __expf(0.5f * seta[s < NX ? s : NX - 1]);
                hx = v;
#pragma unroll
                for (int t = 0; t < NX; ++t) Hr[t] = (t == s) ? 0.5f * v : 0.f;
                he = 0.5f * v * seta[s < NX ? s : NX - 1];
            }
            sinn[s] = (sz[s] - hx + he) * srinv[s];       // R^-1 (z - e), e = h - H eta
        }
        __syncwarp();
        // ---- 3. Hh = H L (row s, entry j), wh = Hh^T R^-1 (z - e) ---------------------------------------------
        float wh = 0.f;
        for (int s = 0; s < nz; ++s) {
            const float* Hr = sH + s * NX;
            float hh = 0.f;
            if (KIND == SMC_LEDH_H_ACOUSTIC) {
#pragma unroll
                for (int c = 0; c < NX / 4; ++c) hh = fmaf(Hr[4 * c], Lcol[4 * c], fmaf(Hr[4 * c + 1], Lcol[4 * c + 1], hh));
            } else {
#pragma unroll
                for (int k = 0; k < NX; ++k) hh = fmaf(Hr[k], Lcol[k], hh);
            }
            if (j < NX) sHh[s * NX + j] = hh;
            wh = fmaf(hh, sinn[s], wh);
        }
        __syncwarp();
        // ---- 4. Gh = Hh^T R^-1 Hh: row j -----------------------------------------------------------------------
        float G[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) G[c] = 0.f;
        for (int s = 0; s < nz; ++s) {
            float row[NX];
            ld_row<NX>(sHh + s * NX, row);                 // 16-byte broadcasts
            const float mine = sHh[s * NX + jx] * srinv[s];
#pragma unroll
            for (int c = 0; c < NX; ++c) G[c] = fmaf(mine, row[c], G[c]);
        }
        if (a.s0 && i == 0) {   // diagnostic of particle 0 (:152-157): S = lam H P H^T + R = lam Hh Hh^T + R
            double* out = a.s0 + (int64_t)(a.step0 + step) * nz * nz;
            for (int idx = j; idx < nz * nz; idx += LP) {
                const int ra = idx / nz, rb = idx - ra * nz;
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < NX; ++c) acc = fmaf(sHh[ra * NX + c], sHh[rb * NX + c], acc);
                out[idx] = (double)lam * (double)acc + a.R[idx];
            }
        }
        // ---- 5. Cholesky of M1 = I + lam Gh and of M2 = I + (lam - dlam/2) Gh, interleaved ---------------------
        float C1[NX], C2[NX];
        const float lam2 = lam - 0.5f * dlam;
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            C1[c] = fmaf(lam, G[c], c == j ? 1.f : 0.f);
            C2[c] = fmaf(lam2, G[c], c == j ? 1.f : 0.f);
        }
        float ld1 = 0.f, ld2 = 0.f, d1 = 1.f;
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            // pivots.  M = I + (positive) Gh with Gh positive semi-definite, so every Schur complement is >= 1 in exact
            // arithmetic; fp32 round-off can push the pivot of a direction the measurement says nothing about below
            // that (even below zero) when another direction is extremely stiff (a particle on top of a sensor under a
            // wide prior: cond(M) ~ 1e9): clamp to the exact lower bound instead of propagating NaNs
            const float q1 = bcast(C1[c], c), q2 = bcast(C2[c], c);
            const bool bad1 = !(q1 >= 1.0f), bad2 = !(q2 >= 1.0f);
            const float p1 = bad1 ? 1.0f : q1, p2 = bad2 ? 1.0f : q2;
            // (a direction whose pivot is lost is decoupled: its column below the diagonal is dropped too, so that the
            // factor stays consistent and the triangular solves bounded)
            if (bad1 && j > c) C1[c] = 0.f;
            if (bad2 && j > c) C2[c] = 0.f;
            const float r1 = rsqrtf(p1), r2 = rsqrtf(p2);
            ld1 += __logf(p1); ld2 += __logf(p2);                          // log det = sum log pivots of L L^T... (pivot^2 = p)
            if (j == c) { d1 = p1 * r1; C1[c] = d1; C2[c] = p2 * r2; }
            else if (j > c) { C1[c] *= r1; C2[c] *= r2; }
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                if (k <= c) continue;
                const float l1 = bcast(C1[c], k), l2 = bcast(C2[c], k);
                C1[k] = fmaf(-C1[c], l1, C1[k]);                           // rows above the diagonal are never read
                C2[k] = fmaf(-C2[c], l2, C2[k]);
            }
        }
        theta += ld2 - ld1;                                                // log det(I + dlam A) (:174-179)
        if (j < NX) {
#pragma unroll
            for (int c = 0; c < NX; ++c) sC[j * LS + c] = C1[c];
        }
        __syncwarp();
        const float invd = 1.0f / d1;
        // y = M1^-1 x (forward with the rows in registers, backward with the columns read from shared memory), then Gh y
        auto apply_Ah = [&](float x) -> float {
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                const float tk = bcast(x, k) * bcast(invd, k);
                if (j == k) x = tk;
                else if (j > k) x = fmaf(-C1[k], tk, x);
            }
#pragma unroll
            for (int k = NX - 1; k >= 0; --k) {
                const float tk = bcast(x, k) * bcast(invd, k);
                if (j == k) x = tk;
                else if (j < k) x = fmaf(-sC[k * LS + jx], tk, x);
            }
            float r = 0.f;
#pragma unroll
            for (int c = 0; c < NX; ++c) r = fmaf(G[c], bcast(x, c), r);
            return -0.5f * r;
        };
        // ---- 6. the flow step (:163-171), two products with Ah ---------------------------------------------------
        const float p = wh + apply_Ah(fmaf(lam, wh, xi0));
        const float rate = p + apply_Ah(fmaf(2.0f * lam, p, xi));
        xi = fmaf(dlam, rate, xi);
        __syncwarp();
    }
    // eta = L xi
    float eta_j = 0.f;
#pragma unroll
    for (int c = 0; c < NX; ++c) {
        const float xc = bcast(xi, c);
        if (c <= jx) eta_j = fmaf(sL[jx * LS + c], xc, eta_j);
    }
    if (live) a.eta[(int64_t)j * a.ld + i] = eta_j;
    if (live && j == 0) a.theta[i] = theta;
}

}  // namespace smc
