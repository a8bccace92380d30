// K4 fast path: LEDH particle flow with compile-time dimensions, one warp per particle, lane = row.
//
// Same arithmetic as k4_ledh_flow (reference models/LEDH_particle_filter.py:132-179), re-laid-out so
// that the per-particle matrices live in REGISTERS, one row per lane, instead of shared memory:
//   lane r < NZ  owns row r of H (Jacobian), of S = lam H P H^T + R and of its Cholesky factor L,
//                and row r of the triangular-solve right-hand sides (Y, X = S^-1 H);
//   lane r < NX  owns row r of P, P H^T, A and of M = I + dlam A.
// Rows other lanes need are broadcast: the Cholesky and LU updates by warp shuffles, the dense
// products through a small per-warp shared scratch read as 16-byte broadcasts.  Every loop is fully
// unrolled (static register indexing); no tensor cores: 16 x 16 / 25 x 25 factorisations per particle
// are latency chains, not a dense contraction.
#pragma once

namespace smc {

template <int N> struct Pad4 { static constexpr int value = (N + 3) / 4 * 4; };

template <typename real> __device__ __forceinline__ real shfl_real(real v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ float log_abs(float v) { return __logf(fabsf(v)); }
__device__ __forceinline__ double log_abs(double v) { return log(fabs(v)); }
__device__ __forceinline__ float sqrt_real(float v) { return sqrtf(v); }
__device__ __forceinline__ double sqrt_real(double v) { return sqrt(v); }

// lane holding the largest |v| among the lanes not yet used as a pivot row (lowest lane on ties)
__device__ __forceinline__ int pivot_lane(float v, bool used) {
    // non-negative floats order like their bit patterns: one redux.sync instead of a shuffle tree
    const unsigned key = used ? 0u : __float_as_uint(fabsf(v));
    const unsigned top = __reduce_max_sync(0xffffffffu, key);
    return __ffs(__ballot_sync(0xffffffffu, !used && key == top)) - 1;
}
__device__ __forceinline__ int pivot_lane(double v, bool used) {
    double best = used ? -1.0 : fabs(v);
    int arg = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    return arg;
}

// 16-byte broadcast loads of one padded row from the per-warp scratch
template <int NP> __device__ __forceinline__ void ld_row(const float* p, float (&out)[NP]) {
#pragma unroll
    for (int q = 0; q < NP / 4; ++q) {
        const float4 v = reinterpret_cast<const float4*>(p)[q];
        out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
    }
}
// 16-byte stores of the first N entries of a register row into a padded scratch row (pad lanes get 0)
template <int NP, int N> __device__ __forceinline__ void st_row(float* p, const float (&v)[N]) {
#pragma unroll
    for (int q = 0; q < NP / 4; ++q)
        reinterpret_cast<float4*>(p)[q] = make_float4(4 * q < N ? v[4 * q < N ? 4 * q : 0] : 0.f,
                                                      4 * q + 1 < N ? v[4 * q + 1 < N ? 4 * q + 1 : 0] : 0.f,
                                                      4 * q + 2 < N ? v[4 * q + 2 < N ? 4 * q + 2 : 0] : 0.f,
                                                      4 * q + 3 < N ? v[4 * q + 3 < N ? 4 * q + 3 : 0] : 0.f);
}
template <int NP, int N> __device__ __forceinline__ void st_row(double* p, const double (&v)[N]) {
#pragma unroll
    for (int q = 0; q < NP / 2; ++q)
        reinterpret_cast<double2*>(p)[q] = make_double2(2 * q < N ? v[2 * q < N ? 2 * q : 0] : 0.0,
                                                        2 * q + 1 < N ? v[2 * q + 1 < N ? 2 * q + 1 : 0] : 0.0);
}
template <int NP> __device__ __forceinline__ void ld_row(const double* p, double (&out)[NP]) {
#pragma unroll
    for (int q = 0; q < NP / 2; ++q) {
        const double2 v = reinterpret_cast<const double2*>(p)[q];
        out[2 * q] = v.x; out[2 * q + 1] = v.y;
    }
}
__device__ __forceinline__ float log_abs_acc(float v) { return logf(fabsf(v)); }
__device__ __forceinline__ double log_abs_acc(double v) { return log(fabs(v)); }

#ifndef SMC_LEDH_FAST_WARPS
#define SMC_LEDH_FAST_WARPS 4
#endif
constexpr int kLedhFastWarps = SMC_LEDH_FAST_WARPS;
#ifndef SMC_LEDH_FAST_MINB
#define SMC_LEDH_FAST_MINB 4        // fp32: resident CTAs per SM the register allocation is held to.  Measured C5 step
                                   // (N=2^20, L=8): 1 (254 regs, 2 CTAs) 299 ms, 3 (168 regs) 163 ms, 4 (128) 165 ms,
                                   // 5 (96) 296 ms, 6 (80) 174 ms -- the kernel is latency bound, not issue bound
#endif

template <typename real, int NX, int NZ> struct LedhFastSmem {
    static constexpr int NXP = Pad4<NX>::value, NZP = Pad4<NZ>::value;
    // per-CTA constants, then per-warp scratch
    static constexpr int kConst = NX * NXP + 2 * NZ * NZP + NZP + 64;
    static constexpr int kPerWarp = 3 * NZ * NXP + NZ * NZP + 2 * NZP + 4 * NXP;
    static constexpr size_t bytes(int n_hp) {
        return sizeof(real) * (size_t)(kConst + (n_hp > 64 ? (n_hp + 3) / 4 * 4 - 64 : 0) + kLedhFastWarps * kPerWarp);
    }
};

template <typename real, int NX, int NZ>
__global__ void __launch_bounds__(32 * kLedhFastWarps, sizeof(real) == 4 ? SMC_LEDH_FAST_MINB : 1) k4_ledh_flow_fast(LedhArgs<real> a) {
    static_assert(NX <= 32 && NZ <= 32, "lane = row layout needs nx, nz <= 32");
    constexpr int NXP = Pad4<NX>::value, NZP = Pad4<NZ>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* sm = reinterpret_cast<real*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- CTA constants (row-padded so that rows are 16-byte aligned) -----------------------------
    real* sP = sm;            sm += NX * NXP;
    real* sR = sm;            sm += NZ * NZP;
    real* sRinv = sm;         sm += NZ * NZP;
    real* sz = sm;            sm += NZP;
    real* shp = sm;           sm += (a.n_hp > 64 ? (a.n_hp + 3) / 4 * 4 : 64);
    for (int i = tid; i < NX * NX; i += blockDim.x) sP[(i / NX) * NXP + i % NX] = (real)a.P[i];
    for (int i = tid; i < NZ * NZ; i += blockDim.x) {
        sR[(i / NZ) * NZP + i % NZ] = (real)a.R[i];
        sRinv[(i / NZ) * NZP + i % NZ] = (real)a.Rinv[i];
    }
    for (int i = tid; i < NZ; i += blockDim.x) sz[i] = (real)a.z[i];
    for (int i = tid; i < a.n_hp; i += blockDim.x) shp[i] = (real)a.hp[i];
    __syncthreads();
    // ---- per-warp scratch ---------------------------------------------------------------------------
    real* w = sm + warp * LedhFastSmem<real, NX, NZ>::kPerWarp;
    real* Hs = w;             w += NZ * NXP;      // H rows (broadcast source)
    real* Xs = w;             w += NZ * NXP;      // rows of Y / X as they are finalised
    real* HPs = w;            w += NZ * NXP;      // rows of H P (P H^T is read from it by columns)
    real* Ls = w;             w += NZ * NZP;      // Cholesky factor (column reads in the backward solve)
    real* vin = w;            w += NZP;           // innov, then R^-1 innov
    real* vz = w;             w += NZP;
    real* veta = w;           w += NXP;           // eta (current)
    real* veta0 = w;          w += NXP;
    real* vv = w;             w += NXP;           // v = P H^T R^-1 innov
    real* vinner = w;         w += NXP;

    const int64_t i = (int64_t)blockIdx.x * kLedhFastWarps + warp;
    if (i >= a.n) return;                         // whole warp leaves together (no block sync below)
    const bool rz = lane < NZ, rx = lane < NX;
    const int lz = rz ? lane : NZ - 1, lx = rx ? lane : NX - 1;    // clamped row indices (idle lanes mirror the last row)

    if (rx) { veta[lane] = a.eta[lane * a.ld + i]; veta0[lane] = a.eta0[lane * a.ld + i]; }
    real Pr[NX], Rr[NZ], Rinvr[NZ];
#pragma unroll
    for (int k = 0; k < NX; ++k) Pr[k] = sP[lx * NXP + k];
#pragma unroll
    for (int c = 0; c < NZ; ++c) { Rr[c] = sR[lz * NZP + c]; Rinvr[c] = sRinv[lz * NZP + c]; }
    __syncwarp();
    real theta = a.theta[i];

    for (int j = 0; j < a.L; ++j) {
        const real dlam = real(1) / (real)a.L;
        real lam;
        {   // lam = min(1, lam + dlam) accumulated exactly as the reference does (:136-137)
            double l = 0.0; const double dl = 1.0 / (double)a.L;
            for (int q = 0; q <= j; ++q) l = fmin(1.0, l + dl);
            lam = (real)l;
        }
        real eta[NX];
#pragma unroll
        for (int t = 0; t < NX; ++t) eta[t] = veta[t];
        // ---- 1. row r of the Jacobian and h_r at the particle (:143-145) ----------------------------
        real Hr[NX], hx = 0;
#pragma unroll
        for (int t = 0; t < NX; ++t) Hr[t] = 0;
        if (a.kind == SMC_LEDH_H_LINEAR) {
#pragma unroll
            for (int t = 0; t < NX; ++t) { Hr[t] = shp[lz * NX + t]; hx = fma(Hr[t], eta[t], hx); }
        } else if (a.kind == SMC_LEDH_H_ACOUSTIC) {
            const real psi = shp[0], d0 = shp[1];
            const real sx = shp[2 + 2 * lz], sy = shp[3 + 2 * lz];
#pragma unroll
            for (int c = 0; c < NX / 4; ++c) {
                const real dx = eta[4 * c] - sx, dy = eta[4 * c + 1] - sy;
                const real den = dx * dx + dy * dy + d0;
                hx += psi / den;
                const real coef = real(-2) * psi / (den * den);
                Hr[4 * c] = coef * dx;
                Hr[4 * c + 1] = coef * dy;
            }
        } else {   // SMC_LEDH_H_EXP: h_i = beta exp(x_i / 2)
            const real v = shp[0] * exp(real(0.5) * veta[lz < NX ? lz : NX - 1]);
            hx = v;
#pragma unroll
            for (int t = 0; t < NX; ++t) Hr[t] = (t == lz) ? real(0.5) * v : real(0);
        }
        if (rz) st_row<NXP>(Hs + lane * NXP, Hr);
        real innov = sz[lz] - hx;                 // z - e = z - (h - H eta)
#pragma unroll
        for (int t = 0; t < NX; ++t) innov = fma(Hr[t], eta[t], innov);
        if (rz) vin[lane] = innov;
        __syncwarp();
        // ---- 2. row r of H P, then of S = lam H P H^T + R (:149) --------------------------------------
        real HPr[NX];
#pragma unroll
        for (int k = 0; k < NX; ++k) HPr[k] = 0;
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            real row[NXP];
            ld_row<NXP>(sP + t * NXP, row);
#pragma unroll
            for (int k = 0; k < NX; ++k) HPr[k] = fma(Hr[t], row[k], HPr[k]);
        }
        if (rz) st_row<NXP>(HPs + lane * NXP, HPr);
        real Sr[NZ];
#pragma unroll
        for (int c = 0; c < NZ; ++c) {
            real acc = 0, row[NXP];
            ld_row<NXP>(Hs + c * NXP, row);
#pragma unroll
            for (int k = 0; k < NX; ++k) acc = fma(HPr[k], row[k], acc);
            Sr[c] = fma(lam, acc, Rr[c]);
        }
        if (a.s0 && i == 0 && rz) {
            double* out = a.s0 + (int64_t)(a.step0 + j) * NZ * NZ + lane * NZ;
#pragma unroll
            for (int c = 0; c < NZ; ++c) out[c] = (double)Sr[c];
        }
        // ---- 3. Cholesky S = L L^T in registers, row per lane, broadcasts by shuffle -------------------
        real diag = 1;                                           // L[lane][lane]
#pragma unroll
        for (int c = 0; c < NZ; ++c) {
            const real d = sqrt_real(shfl_real(Sr[c], c));
            if (lane == c) { Sr[c] = d; diag = d; }
            else if (lane > c) Sr[c] = Sr[c] / d;
            // constant trip count (the triangular bound is a compile-time predicate after unrolling):
            // with `k = c + 1` as the lower bound Sr[] was left in local memory (ncu: 930 LDL/STL
            // per particle-substep, profiles/r1_ledh_notes.txt)
#pragma unroll
            for (int k = 0; k < NZ; ++k) {
                if (k <= c) continue;
                const real lkc = shfl_real(Sr[c], k);
                Sr[k] = fma(-Sr[c], lkc, Sr[k]);        // rows above the diagonal (lane < k) are never read
            }
        }
        if (rz) st_row<NZP>(Ls + lane * NZP, Sr);
        // ---- 4. X = S^-1 H: forward and backward substitution, one row per lane ------------------------
        // Row k is broadcast UNSCALED and the other rows fold 1/L_kk into their own coefficient
        // (L_rk / L_kk) w_k; every lane scales its own row once at the end of each sweep.  (Scaling
        // the broadcast row first costs 16 single-lane multiplies and a division per step.)
        const real invd = real(1) / diag;
        real Yr[NX];
#pragma unroll
        for (int t = 0; t < NX; ++t) Yr[t] = Hr[t];
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
            if (lane == k) st_row<NXP>(Xs + k * NXP, Yr);
            const real ik = shfl_real(invd, k);
            __syncwarp();
            if (lane > k && rz) {
                const real coef = -Sr[k] * ik;
                real row[NXP];
                ld_row<NXP>(Xs + k * NXP, row);
#pragma unroll
                for (int t = 0; t < NX; ++t) Yr[t] = fma(coef, row[t], Yr[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < NX; ++t) Yr[t] *= invd;
        __syncwarp();
#pragma unroll
        for (int k = NZ - 1; k >= 0; --k) {
            if (lane == k) st_row<NXP>(Xs + k * NXP, Yr);
            const real ik = shfl_real(invd, k);
            __syncwarp();
            if (lane < k) {
                const real coef = -Ls[k * NZP + lane] * ik;     // L[k][r] / L[k][k]
                real row[NXP];
                ld_row<NXP>(Xs + k * NXP, row);
#pragma unroll
                for (int t = 0; t < NX; ++t) Yr[t] = fma(coef, row[t], Yr[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < NX; ++t) Yr[t] *= invd;
        __syncwarp();
        if (rz) st_row<NXP>(Xs + lane * NXP, Yr);
        __syncwarp();                                           // Xs = S^-1 H, complete
        // ---- 5. rows of P H^T and A = -1/2 P H^T S^-1 H (:158-159) -------------------------------------
        // (P H^T)[r][s] = (H P)[s][r] bit for bit (P is symmetrised by the host, same summation order):
        // one conflict-free column read per entry instead of a second nx-long dot product
        real PHtr[NZ];
#pragma unroll
        for (int s = 0; s < NZ; ++s) PHtr[s] = HPs[s * NXP + lx];
        real Ar[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) Ar[c] = 0;
#pragma unroll
        for (int s = 0; s < NZ; ++s) {
            real row[NXP];
            ld_row<NXP>(Xs + s * NXP, row);
#pragma unroll
            for (int c = 0; c < NX; ++c) Ar[c] = fma(PHtr[s], row[c], Ar[c]);
        }
#pragma unroll
        for (int c = 0; c < NX; ++c) Ar[c] *= real(-0.5);
        // ---- 6. b and the Euler step (:163-171) ------------------------------------------------------------
        real rinn = 0;
#pragma unroll
        for (int c = 0; c < NZ; ++c) rinn = fma(Rinvr[c], vin[c], rinn);
        if (rz) vz[lane] = rinn;
        __syncwarp();
        real v = 0;
#pragma unroll
        for (int s = 0; s < NZ; ++s) v = fma(PHtr[s], vz[s], v);
        if (rx) vv[lane] = v;
        __syncwarp();
        real av = 0, a0 = 0, ae = 0;
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            av = fma(Ar[t], vv[t], av);
            a0 = fma(Ar[t], veta0[t], a0);
            ae = fma(Ar[t], eta[t], ae);
        }
        const real inner = v + lam * av + a0;
        if (rx) vinner[lane] = inner;
        __syncwarp();
        real ai = 0;
#pragma unroll
        for (int t = 0; t < NX; ++t) ai = fma(Ar[t], vinner[t], ai);
        const real eta_new = veta[lx] + dlam * (ae + inner + real(2) * lam * ai);
        __syncwarp();
        if (rx) veta[lane] = eta_new;
        // ---- 7. theta += log|det(I + dlam A)| (:174-179): LU with partial pivoting, rows stay put ------
        real ld = 0;
        int sg = 1;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            real Mr[NX];
#pragma unroll
            for (int c = 0; c < NX; ++c)
                Mr[c] = dlam * Ar[c] + ((c == lane) ? real(1) + (attempt ? real(1e-12) : real(0)) : real(0));
            bool used = !rx;
            int step_of = 0;
            ld = 0; sg = 1;
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                const int arg = pivot_lane(Mr[c], used);
                const real piv = shfl_real(Mr[c], arg);
                if (piv < 0) sg = -sg;
                if (piv == 0) sg = 0;
                ld += log_abs_acc(piv);
                if (lane == arg) { used = true; step_of = c; }
                const real f = used ? real(0) : Mr[c] / piv;
#pragma unroll
                for (int k = 0; k < NX; ++k) {
                    if (k <= c) continue;
                    const real pk = shfl_real(Mr[k], arg);
                    Mr[k] = fma(-f, pk, Mr[k]);
                }
            }
            // parity of the row permutation (lane -> elimination step)
            int inv = 0;
#pragma unroll
            for (int l = 0; l < NX; ++l) {
                const int so = __shfl_sync(0xffffffffu, step_of, l);
                if (rx && l < lane && so > step_of) ++inv;
            }
            inv = warp_sum(inv);
            if (inv & 1) sg = -sg;
            if (sg > 0) break;                                    // else: jitter retry (:176-178)
        }
        theta += ld;
        __syncwarp();
    }
    if (rx) a.eta[lane * a.ld + i] = veta[lane];
    if (lane == 0) a.theta[i] = theta;
}

}  // namespace smc
