// LEDH particle-flow hot path (K4) for sm_100a: one warp per particle, all per-particle
// matrices in shared memory, no tensor cores (batched 16x16 / 25x25 factorisations are not a
// dense contraction).
//
// Reference semantics: models/LEDH_particle_filter.py:132-179 (flow), :186-195 (weights),
// :217-224 (_weighted_stats) of liyiyang-amber/Particle_filters.
#include "smc_common.cuh"
#include "../../include/smc_ledh.h"

#include <stdlib.h>
#include <vector>

namespace smc {

constexpr int kLedhMaxDim = 32;      // shared-memory / register kernels (and the device-resident linear-Gaussian pieces)
constexpr int kLedhBigMaxDim = 512;  // flow kernel with its matrices in a global-memory workspace
constexpr int kLedhWarps = 4;        // particles (warps) per CTA

// per-CTA constants staged once into shared memory
template <typename real> struct LedhShared {
    real* P;      // nx*nx
    real* R;      // nz*nz
    real* Rinv;   // nz*nz
    real* z;      // nz
    real* hp;     // functor parameters
};

// ---- warp-cooperative small dense kernels on shared-memory matrices (row-major) --------------
// C(m x n) = alpha * A(m x k) * B(k x n) [+ D]
template <typename real>
__device__ inline void w_matmul(real* C, const real* A, const real* B, int m, int k, int n, real alpha,
                                const real* D, int lane) {
    for (int idx = lane; idx < m * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        real acc = 0;
        for (int t = 0; t < k; ++t) acc = fma(A[i * k + t], B[t * n + j], acc);
        C[idx] = alpha * acc + (D ? D[idx] : real(0));
    }
    __syncwarp();
}
// C(m x n) = A(m x k) * B^T  with B (n x k)
template <typename real>
__device__ inline void w_matmul_bt(real* C, const real* A, const real* B, int m, int k, int n, int lane) {
    for (int idx = lane; idx < m * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        real acc = 0;
        for (int t = 0; t < k; ++t) acc = fma(A[i * k + t], B[j * k + t], acc);
        C[idx] = acc;
    }
    __syncwarp();
}
// y(m) = A(m x k) x(k)
template <typename real>
__device__ inline void w_matvec(real* y, const real* A, const real* x, int m, int k, int lane) {
    for (int i = lane; i < m; i += 32) {
        real acc = 0;
        for (int t = 0; t < k; ++t) acc = fma(A[i * k + t], x[t], acc);
        y[i] = acc;
    }
    __syncwarp();
}

// In-place Cholesky S = L L^T (lower triangle holds L).  S is SPD: lam H P H^T + R (:149).
template <typename real> __device__ inline void w_cholesky(real* S, int n, int lane) {
    for (int j = 0; j < n; ++j) {
        const real d = sqrt(S[j * n + j]);
        __syncwarp();
        if (lane == 0) S[j * n + j] = d;
        for (int i = j + 1 + lane; i < n; i += 32) S[i * n + j] /= d;
        __syncwarp();
        const int rem = n - j - 1;
        for (int idx = lane; idx < rem * rem; idx += 32) {
            const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
            if (k <= i) S[i * n + k] = fma(-S[i * n + j], S[k * n + j], S[i * n + k]);
        }
        __syncwarp();
    }
}
// Solve (L L^T) X = B in place, B (n x m): one lane per right-hand-side column.
template <typename real> __device__ inline void w_chol_solve(const real* L, real* B, int n, int m, int lane) {
    for (int c = lane; c < m; c += 32) {
        for (int r = 0; r < n; ++r) {
            real acc = B[r * m + c];
            for (int k = 0; k < r; ++k) acc = fma(-L[r * n + k], B[k * m + c], acc);
            B[r * m + c] = acc / L[r * n + r];
        }
        for (int r = n - 1; r >= 0; --r) {
            real acc = B[r * m + c];
            for (int k = r + 1; k < n; ++k) acc = fma(-L[k * n + r], B[k * m + c], acc);
            B[r * m + c] = acc / L[r * n + r];
        }
    }
    __syncwarp();
}
// log|det M| and sign via LU with partial pivoting, in place (np.linalg.slogdet, :174).
template <typename real> __device__ inline void w_lu_logdet(real* M, int n, int lane, real* logdet, int* sign) {
    real ld = 0;
    int sg = 1;
    for (int j = 0; j < n; ++j) {
        real best = -1;
        int arg = j;
        for (int i = j + lane; i < n; i += 32) {
            const real v = fabs(M[i * n + j]);
            if (v > best) { best = v; arg = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const real ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (arg != j) {
            for (int k = lane; k < n; k += 32) {
                const real t = M[j * n + k];
                M[j * n + k] = M[arg * n + k];
                M[arg * n + k] = t;
            }
            sg = -sg;
        }
        __syncwarp();
        const real piv = M[j * n + j];
        if (piv < 0) sg = -sg;
        if (piv == 0) sg = 0;
        ld += log(fabs(piv));
        for (int i = j + 1 + lane; i < n; i += 32) M[i * n + j] /= piv;
        __syncwarp();
        const int rem = n - j - 1;
        for (int idx = lane; idx < rem * rem; idx += 32) {
            const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
            M[i * n + k] = fma(-M[i * n + j], M[j * n + k], M[i * n + k]);
        }
        __syncwarp();
    }
    *logdet = ld;
    *sign = sg;
}

// ---- measurement functors: fill hx (nz) and H (nz x nx) at eta --------------------------------
template <typename real>
__device__ inline void eval_h(int kind, const real* hp, const real* eta, real* hx, real* H, int nx, int nz, int lane) {
    if (kind == SMC_LEDH_H_LINEAR) {
        for (int idx = lane; idx < nz * nx; idx += 32) H[idx] = hp[idx];
        __syncwarp();
        w_matvec(hx, H, eta, nz, nx, lane);
    } else if (kind == SMC_LEDH_H_ACOUSTIC) {
        const real psi = hp[0], d0 = hp[1];
        const int nt = nx / 4;
        for (int idx = lane; idx < nz * nx; idx += 32) H[idx] = 0;
        __syncwarp();
        for (int s = lane; s < nz; s += 32) {
            const real sx = hp[2 + 2 * s], sy = hp[3 + 2 * s];
            real acc = 0;
            for (int c = 0; c < nt; ++c) {
                const real dx = eta[4 * c] - sx, dy = eta[4 * c + 1] - sy;
                const real den = dx * dx + dy * dy + d0;
                acc += psi / den;
                const real coef = real(-2) * psi / (den * den);
                H[s * nx + 4 * c] = coef * dx;
                H[s * nx + 4 * c + 1] = coef * dy;
            }
            hx[s] = acc;
        }
        __syncwarp();
    } else {   // SMC_LEDH_H_EXP: h_i = beta exp(x_i / 2)
        const real beta = hp[0];
        for (int idx = lane; idx < nz * nx; idx += 32) H[idx] = 0;
        __syncwarp();
        for (int i = lane; i < nz; i += 32) {
            const real v = beta * exp(real(0.5) * eta[i]);
            hx[i] = v;
            H[i * nx + i] = real(0.5) * v;
        }
        __syncwarp();
    }
}

template <typename real> struct LedhArgs {
    const real* eta0;
    real* eta;
    real* theta;
    double* s0;
    const double* P; const double* R; const double* Rinv; const double* z; const double* hp;
    const double* Hsup; const double* hsup;     // supplied-Jacobian mode (nullable)
    int64_t n, ld;
    int nx, nz, n_hp, kind, L, step0;
    double lam0, dlam;                          // supplied mode: one step at lam0 (L = 1)
    real* gws;                                  // global-memory workspace (dimensions beyond shared memory), else null
};

// GWS = false: constants and per-warp matrices in shared memory, one particle per warp, grid = particles / 4.
// GWS = true : the same code with every matrix in a global-memory workspace (nx, nz up to 512: the reference's
//              64-, 144- and 400-dimensional sensor-network notebooks) and a persistent grid -- each warp owns
//              one workspace slot and loops over particles.  A warp only ever reads what it wrote itself (or the
//              per-launch constants), so __syncwarp()'s memory ordering is all the coherence it needs.
template <typename real, bool GWS>
__global__ void __launch_bounds__(32 * kLedhWarps) k4_ledh_flow(LedhArgs<real> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* sm = GWS ? a.gws : reinterpret_cast<real*>(smem_raw);
    const int nx = a.nx, nz = a.nz;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // block-shared constants (GWS: one global copy, written by CTA 0's launch-mate kernel below)
    real* sP = sm;            sm += nx * nx;
    real* sR = sm;            sm += nz * nz;
    real* sRinv = sm;         sm += nz * nz;
    real* sz = sm;            sm += nz;
    real* shp = sm;           sm += a.n_hp;
    if (!GWS) {
        for (int i = tid; i < nx * nx; i += blockDim.x) sP[i] = (real)a.P[i];
        for (int i = tid; i < nz * nz; i += blockDim.x) { sR[i] = (real)a.R[i]; sRinv[i] = (real)a.Rinv[i]; }
        for (int i = tid; i < nz; i += blockDim.x) sz[i] = (real)a.z[i];
        for (int i = tid; i < a.n_hp; i += blockDim.x) shp[i] = (real)a.hp[i];
        __syncthreads();
    }
    // per-warp workspace
    const int per_warp = 3 * nz * nx + nz * nz + 2 * nx * nx + 5 * nx + 3 * nz;
    const int64_t slot = GWS ? (int64_t)blockIdx.x * kLedhWarps + warp : warp;
    real* w = sm + slot * per_warp;
    real* H = w;              w += nz * nx;     // Jacobian, later S^-1 H
    real* PHt = w;            w += nx * nz;
    real* Hc = w;             w += nz * nx;     // copy of H kept for the solve right-hand side
    real* S = w;              w += nz * nz;
    real* A = w;              w += nx * nx;
    real* M = w;              w += nx * nx;
    real* eta = w;            w += nx;
    real* eta0 = w;           w += nx;
    real* v = w;              w += nx;
    real* inner = w;          w += nx;
    real* tx = w;             w += nx;
    real* hx = w;             w += nz;
    real* innov = w;          w += nz;
    real* rinn = w;           w += nz;

    const int64_t i_stride = GWS ? (int64_t)gridDim.x * kLedhWarps : a.n;
    for (int64_t i = (int64_t)blockIdx.x * kLedhWarps + warp; i < a.n; i += i_stride) {   // whole warps (no block sync below)
    for (int d = lane; d < nx; d += 32) { eta[d] = a.eta[d * a.ld + i]; eta0[d] = a.eta0[d * a.ld + i]; }
    __syncwarp();
    real theta = a.theta[i];

    for (int j = 0; j < a.L; ++j) {
        real lam, dlam;
        if (a.Hsup) { lam = (real)a.lam0; dlam = (real)a.dlam; }
        else {
            dlam = real(1) / (real)a.L;
            // lam = min(1, lam + dlam) accumulated exactly as the reference does (:136-137)
            double l = 0.0; const double dl = 1.0 / (double)a.L;
            for (int q = 0; q <= j; ++q) l = fmin(1.0, l + dl);
            lam = (real)l;
        }
        // 1. linearise at the particle (:143-145)
        if (a.Hsup) {
            for (int idx = lane; idx < nz * nx; idx += 32) H[idx] = (real)a.Hsup[i * nz * nx + idx];
            for (int s = lane; s < nz; s += 32) hx[s] = (real)a.hsup[i * nz + s];
            __syncwarp();
        } else {
            eval_h(a.kind, shp, eta, hx, H, nx, nz, lane);
        }
        // innov = z - e = z - (h - H eta)
        for (int s = lane; s < nz; s += 32) {
            real acc = sz[s] - hx[s];
            for (int t = 0; t < nx; ++t) acc = fma(H[s * nx + t], eta[t], acc);
            innov[s] = acc;
        }
        __syncwarp();
        // 2. P H^T and S = lam H P H^T + R (:149)
        w_matmul_bt(PHt, sP, H, nx, nx, nz, lane);
        w_matmul(S, H, PHt, nz, nx, nz, lam, sR, lane);
        if (a.s0 && i == 0) {
            double* out = a.s0 + (int64_t)(a.step0 + j) * nz * nz;
            for (int idx = lane; idx < nz * nz; idx += 32) out[idx] = (double)S[idx];
        }
        // 3. A = -1/2 P H^T S^-1 H (:158-159)
        for (int idx = lane; idx < nz * nx; idx += 32) Hc[idx] = H[idx];
        __syncwarp();
        w_cholesky(S, nz, lane);
        w_chol_solve(S, Hc, nz, nx, lane);
        w_matmul(A, PHt, Hc, nx, nz, nx, real(-0.5), (const real*)nullptr, lane);
        // 4. b = (I + 2 lam A)[(I + lam A) P H^T R^-1 (z - e) + A eta0] (:163-165)
        w_matvec(rinn, sRinv, innov, nz, nz, lane);
        w_matvec(v, PHt, rinn, nx, nz, lane);
        for (int r = lane; r < nx; r += 32) {
            real av = 0, a0 = 0;
            for (int t = 0; t < nx; ++t) { av = fma(A[r * nx + t], v[t], av); a0 = fma(A[r * nx + t], eta0[t], a0); }
            inner[r] = v[r] + lam * av + a0;
        }
        __syncwarp();
        // eta <- eta + dlam (A eta + b), b = inner + 2 lam A inner (:171, explicit Euler on the old eta)
        for (int r = lane; r < nx; r += 32) {
            real ai = 0, ae = 0;
            for (int t = 0; t < nx; ++t) { ai = fma(A[r * nx + t], inner[t], ai); ae = fma(A[r * nx + t], eta[t], ae); }
            tx[r] = eta[r] + dlam * (ae + inner[r] + real(2) * lam * ai);
        }
        __syncwarp();
        for (int r = lane; r < nx; r += 32) eta[r] = tx[r];
        // 5. theta += log|det(I + dlam A)| (:174-179)
        for (int idx = lane; idx < nx * nx; idx += 32) M[idx] = dlam * A[idx] + ((idx / nx == idx % nx) ? real(1) : real(0));
        __syncwarp();
        real ld; int sg;
        w_lu_logdet(M, nx, lane, &ld, &sg);
        if (sg <= 0) {   // jitter retry, :176-178
            for (int idx = lane; idx < nx * nx; idx += 32)
                M[idx] = dlam * A[idx] + ((idx / nx == idx % nx) ? real(1) + real(1e-12) : real(0));
            __syncwarp();
            w_lu_logdet(M, nx, lane, &ld, &sg);
        }
        theta += ld;
        __syncwarp();
    }
    for (int d = lane; d < nx; d += 32) a.eta[d * a.ld + i] = eta[d];
    if (lane == 0) a.theta[i] = theta;
    __syncwarp();
    }
}
// constants of the global-workspace mode, converted to `real` once per launch
template <typename real> __global__ void k4_ledh_stage_constants(LedhArgs<real> a) {
    real* sm = a.gws;
    const int nx = a.nx, nz = a.nz;
    real* sP = sm;            sm += nx * nx;
    real* sR = sm;            sm += nz * nz;
    real* sRinv = sm;         sm += nz * nz;
    real* sz = sm;            sm += nz;
    real* shp = sm;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < nx * nx; i += nt) sP[i] = (real)a.P[i];
    for (int i = tid; i < nz * nz; i += nt) { sR[i] = (real)a.R[i]; sRinv[i] = (real)a.Rinv[i]; }
    for (int i = tid; i < nz; i += nt) sz[i] = (real)a.z[i];
    for (int i = tid; i < a.n_hp; i += nt) shp[i] = (real)a.hp[i];
}


// ---- EDH flow: the pseudo-time integration is ONE affine map for every particle -------------------
// (models/EDH_particle_filter.py:218-275: A(lambda), b(lambda) are evaluated at the tracker's mean
// path, not per particle, so the L Euler / RK4 sub-steps compose to x <- M x + c on the host and the
// particle update is a streaming kernel: thread per particle, M and c in shared memory, fp64 math)
template <typename real>
__global__ void __launch_bounds__(kThreads) k_affine_map(real* x, const double* M, const double* c, int64_t n,
                                                         int64_t ld, int nx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sM = reinterpret_cast<double*>(smem_raw);
    double* sc = sM + nx * nx;
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) sM[i] = M[i];
    for (int i = threadIdx.x; i < nx; i += blockDim.x) sc[i] = c[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    double xi[kLedhMaxDim], yi[kLedhMaxDim];
    for (int d = 0; d < nx; ++d) xi[d] = (double)x[d * ld + i];
    for (int r = 0; r < nx; ++r) {
        double acc = sc[r];
        for (int d = 0; d < nx; ++d) acc = fma(sM[r * nx + d], xi[d], acc);
        yi[r] = acc;
    }
    for (int d = 0; d < nx; ++d) x[d * ld + i] = (real)yi[d];
}

}  // namespace smc
#include "ledh_fast.cuh"
#include "ledh_info.cuh"
namespace smc {

// ---- device-resident linear-Gaussian pieces (thread per particle, constants in shared memory) ----
template <typename real>
__global__ void __launch_bounds__(kThreads) k_ledh_propagate_linear(const real* x, real* eta0, real* eta, real* theta,
                                                                    const real* v, const double* F, const double* Lq,
                                                                    int64_t n, int64_t ld, int nx, uint64_t seed,
                                                                    uint32_t t) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sF = reinterpret_cast<double*>(smem_raw);
    double* sL = sF + nx * nx;
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) { sF[i] = F[i]; sL[i] = Lq[i]; }
    __syncthreads();
    const int64_t quad = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t i0 = quad * 4;
    if (i0 >= n) return;
    // one thread handles 4 consecutive particles so that one Philox block serves them (per dimension)
    double xs[kLedhMaxDim][4];
    for (int d = 0; d < nx; ++d) {
        real tmp[4];
        load4<real>(x + d * ld + i0, tmp);
#pragma unroll
        for (int j = 0; j < 4; ++j) xs[d][j] = (double)tmp[j];
    }
    double nz[kLedhMaxDim][4];
    if (!v) {
        for (int d = 0; d < nx; ++d) {
            uint32_t r[4];
            double z2[2];
            philox_quad(seed, (uint64_t)i0 >> 2, t, kStreamProcess | (uint32_t(d) << 8), r);
            normals2_f64(r, z2); nz[d][0] = z2[0]; nz[d][1] = z2[1];
            philox_quad(seed, (uint64_t)i0 >> 2, t, kStreamProcess | (uint32_t(d) << 8) | (1u << 16), r);
            normals2_f64(r, z2); nz[d][2] = z2[0]; nz[d][3] = z2[1];
        }
    }
    for (int d = 0; d < nx; ++d) {
        double acc[4] = {0, 0, 0, 0};
        for (int k = 0; k < nx; ++k) {
            const double f = sF[d * nx + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fma(f, xs[k][j], acc[j]);
        }
        if (v) {
            real tmp[4];
            load4<real>(v + d * ld + i0, tmp);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += (double)tmp[j];
        } else {
            for (int k = 0; k <= d; ++k) {
                const double l = sL[d * nx + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fma(l, nz[k][j], acc[j]);
            }
        }
        real out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = (real)acc[j];
        store4<real>(eta0 + d * ld + i0, out);
        store4<real>(eta + d * ld + i0, out);
    }
    real zero[4] = {0, 0, 0, 0};
    store4<real>(theta + i0, zero);
}

template <typename real>
__global__ void __launch_bounds__(128) k_ledh_gauss_terms(const real* x, const real* eta0, const real* eta,
                                                          const double* F, const double* Qinv, const double* Rinv,
                                                          const double* z, const double* hp, int n_hp, int kind,
                                                          int64_t n, int64_t ld, int nx, int nz, double* terms) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sF = reinterpret_cast<double*>(smem_raw);
    double* sQi = sF + nx * nx;
    double* sRi = sQi + nx * nx;
    double* sz = sRi + nz * nz;
    double* shp = sz + nz;
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) { sF[i] = F[i]; sQi[i] = Qinv[i]; }
    for (int i = threadIdx.x; i < nz * nz; i += blockDim.x) sRi[i] = Rinv[i];
    for (int i = threadIdx.x; i < nz; i += blockDim.x) sz[i] = z[i];
    for (int i = threadIdx.x; i < n_hp; i += blockDim.x) shp[i] = hp[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xv[kLedhMaxDim], e1[kLedhMaxDim], e0[kLedhMaxDim], et[kLedhMaxDim];
    for (int d = 0; d < nx; ++d) { xv[d] = (double)x[d * ld + i]; et[d] = (double)eta[d * ld + i]; }
    for (int d = 0; d < nx; ++d) {
        double fx = 0;
        for (int k = 0; k < nx; ++k) fx = fma(sF[d * nx + k], xv[k], fx);
        e1[d] = et[d] - fx;
        e0[d] = (double)eta0[d * ld + i] - fx;
    }
    double q1 = 0, q0 = 0;
    for (int d = 0; d < nx; ++d) {
        double a1 = 0, a0 = 0;
        for (int k = 0; k < nx; ++k) { a1 = fma(sQi[d * nx + k], e1[k], a1); a0 = fma(sQi[d * nx + k], e0[k], a0); }
        q1 = fma(e1[d], a1, q1);
        q0 = fma(e0[d], a0, q0);
    }
    // innovation z - h(eta)
    double dz[kLedhMaxDim];
    if (kind == SMC_LEDH_H_LINEAR) {
        for (int s = 0; s < nz; ++s) {
            double acc = sz[s];
            for (int k = 0; k < nx; ++k) acc = fma(-shp[s * nx + k], et[k], acc);
            dz[s] = acc;
        }
    } else if (kind == SMC_LEDH_H_ACOUSTIC) {
        const double psi = shp[0], d0 = shp[1];
        for (int s = 0; s < nz; ++s) {
            double acc = 0;
            for (int c = 0; c < nx / 4; ++c) {
                const double dx = et[4 * c] - shp[2 + 2 * s], dy = et[4 * c + 1] - shp[3 + 2 * s];
                acc += psi / (dx * dx + dy * dy + d0);
            }
            dz[s] = sz[s] - acc;
        }
    } else {
        for (int s = 0; s < nz; ++s) dz[s] = sz[s] - shp[0] * exp(0.5 * et[s]);
    }
    double qz = 0;
    for (int s = 0; s < nz; ++s) {
        double acc = 0;
        for (int k = 0; k < nz; ++k) acc = fma(sRi[s * nz + k], dz[k], acc);
        qz = fma(dz[s], acc, qz);
    }
    terms[i] = -0.5 * (q1 - q0) - 0.5 * qz;
}

// logw = log(w_prev + 1e-300) + theta_log + (log-density terms) (:186-190), with block partials
template <typename real>
__global__ void __launch_bounds__(kThreads) k_ledh_weights(real* logw, const double* wprev, const real* theta,
                                                           const double* terms, int64_t n, double* parts) {
    __shared__ double s_tmp[kWarps];
    double m = -INFINITY, s = 0, s2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double v = log(wprev[i] + 1e-300) + (double)theta[i] + terms[i];
        // NaN (a particle whose flow blew up) counts as zero weight, and is STORED as -inf: the resampler's
        // quantisation assumes NaN-free log-weights, as K1 guarantees for its own
        logw[i] = (v == v) ? (real)v : real(-INFINITY);
        if (!(v == v)) continue;
        if (v > m) { const double sc = exp(m - v); s *= sc; s2 *= sc * sc; m = v; }
        if (m != -INFINITY) { const double e = exp(v - m); s += e; s2 = fma(e, e, s2); }
    }
    const double bmax = block_max_256(m, s_tmp);
    const double sc = (m == -INFINITY) ? 0.0 : exp(m - bmax);
    const double bs = block_sum_256(s * sc, s_tmp);
    const double bs2 = block_sum_256(s2 * sc * sc, s_tmp);
    if (threadIdx.x == 0) {
        double* out = parts + (int64_t)blockIdx.x * 3;
        out[0] = bmax; out[1] = bs; out[2] = bs2;
    }
}

// weighted mean / covariance with explicit linear weights (or uniform), any nx (:217-224)
template <typename real>
__global__ void __launch_bounds__(kThreads) k_weighted_moments(const Ctrl* ctrl, const real* x0, const real* x1,
                                                               const double* w, int64_t n, int64_t ld, int nx,
                                                               double* stats) {
    __shared__ double s_tmp[kWarps];
    const real* x = ctrl->cur ? x1 : x0;
    const int b = blockIdx.x;
    int d, d2;
    if (b < nx) { d = b; d2 = -1; }
    else {
        int t = b - nx; d = 0;
        while (t >= d + 1) { t -= d + 1; ++d; }
        d2 = t;
    }
    const double c1 = d < kMaxNx ? ctrl->shift[d] : 0.0;
    const double c2 = (d2 >= 0 && d2 < kMaxNx) ? ctrl->shift[d2] : 0.0;
    double acc = 0, accw = 0;
    for (int64_t i = threadIdx.x; i < n; i += kThreads) {
        const double wi = w ? w[i] : 1.0;
        const double av = (double)x[d * ld + i] - c1;
        const double v = d2 >= 0 ? av * ((double)x[d2 * ld + i] - c2) : av;
        acc = fma(wi, v, acc);
        accw += wi;
    }
    const double tot = block_sum_256(acc, s_tmp);
    const double totw = block_sum_256(accw, s_tmp);
    if (threadIdx.x == 0) {
        double* raw = stats + SMC_STATS_HEADER;
        if (d2 < 0) raw[d] = tot / totw;
        else raw[nx + d * nx + d2] = tot / totw;
    }
}
__global__ void k_weighted_moments_fix(const Ctrl* ctrl, int nx, double* stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double* mean = stats + SMC_STATS_HEADER;
    double* cov = mean + nx;
    for (int d = 0; d < nx; ++d)
        for (int d2 = 0; d2 <= d; ++d2) {
            const double v = cov[d * nx + d2] - mean[d] * mean[d2];
            cov[d * nx + d2] = v;
            cov[d2 * nx + d] = v;
        }
    for (int d = 0; d < nx; ++d) mean[d] += (d < kMaxNx ? ctrl->shift[d] : 0.0);
}

}  // namespace smc

using namespace smc;

namespace {

// dense inverse by Gauss-Jordan with partial pivoting (host, nz <= 32)
bool invert(const double* A, int n, std::vector<double>& inv) {
    std::vector<double> M(A, A + n * n);
    inv.assign(n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[i * n + i] = 1.0;
    for (int j = 0; j < n; ++j) {
        int p = j;
        for (int i = j + 1; i < n; ++i) if (fabs(M[i * n + j]) > fabs(M[p * n + j])) p = i;
        if (M[p * n + j] == 0.0) return false;
        if (p != j)
            for (int k = 0; k < n; ++k) { std::swap(M[j * n + k], M[p * n + k]); std::swap(inv[j * n + k], inv[p * n + k]); }
        const double d = M[j * n + j];
        for (int k = 0; k < n; ++k) { M[j * n + k] /= d; inv[j * n + k] /= d; }
        for (int i = 0; i < n; ++i) {
            if (i == j) continue;
            const double f = M[i * n + j];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) { M[i * n + k] -= f * M[j * n + k]; inv[i * n + k] -= f * inv[j * n + k]; }
        }
    }
    return true;
}

// host Cholesky P = L L^T (lower); false when P is not positive definite
unsigned long long g_info_jitter = 0, g_info_fallback = 0;      // smc_ledh_info_counters
bool cholesky_lower(const double* P, int n, std::vector<double>& Lc) {
    Lc.assign((size_t)n * n, 0.0);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c <= r; ++c) {
            double acc = 0.5 * (P[r * n + c] + P[c * n + r]);
            for (int k = 0; k < c; ++k) acc -= Lc[r * n + k] * Lc[c * n + k];
            if (r == c) {
                if (!(acc > 0.0)) return false;
                Lc[r * n + r] = sqrt(acc);
            } else {
                Lc[r * n + c] = acc / Lc[c * n + c];
            }
        }
    return true;
}

// information-form flow (ledh_info.cuh): fp32, R diagonal, nx in {4, 8, 16}.  Returns -1 when it does not apply.
int launch_flow_info(const smc_ledh_desc* d, const double* P, const double* R, const double* z, int kind, const double* hp,
                     int n_hp, int step0, int L, cudaStream_t st) {
    const int nx = d->nx, nz = d->nz;
    if (d->dtype != SMC_F32 || !(nx == 4 || nx == 8 || nx == 16) || nz > kLedhInfoMaxNz || getenv("SMC_LEDH_LITERAL")) return -1;
    for (int r = 0; r < nz; ++r)
        for (int c = 0; c < nz; ++c)
            if (r != c && R[r * nz + c] != 0.0) return -1;
            else if (r == c && !(R[r * nz + c] > 0.0)) return -1;
    // P = L L^T.  A tracker's covariance can lose definiteness in the last bits ((I - K H) P is not symmetrised by the
    // reference's EKF); the literal kernels do not care, the information form needs the factor: retry with a relative
    // jitter far below the fp32 resolution of this mode (<= 1e-7 of the mean diagonal) before giving the step to the
    // literal kernel (five times slower at 16 x 25)
    std::vector<double> Lc;
    bool have = cholesky_lower(P, nx, Lc);
    if (!have) {
        double tr = 0.0;
        for (int i = 0; i < nx; ++i) tr += fabs(P[i * nx + i]);
        std::vector<double> Pj(P, P + (size_t)nx * nx);
        for (double eps = 1e-13; eps <= 1e-7 && !have; eps *= 100.0) {
            for (int i = 0; i < nx; ++i) Pj[i * nx + i] = P[i * nx + i] + eps * tr / nx;
            have = cholesky_lower(Pj.data(), nx, Lc);
        }
        __atomic_add_fetch(&g_info_jitter, 1ull, __ATOMIC_RELAXED);
    }
    if (!have) { __atomic_add_fetch(&g_info_fallback, 1ull, __ATOMIC_RELAXED); return -1; }
    const size_t cnt = (size_t)nx * nx + (size_t)nz * nz + nz + (size_t)(n_hp > 0 ? n_hp : 0);
    std::vector<double> host(cnt);
    double* q = host.data();
    memcpy(q, Lc.data(), sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, R, sizeof(double) * nz * nz); q += nz * nz;
    memcpy(q, z, sizeof(double) * nz); q += nz;
    if (n_hp > 0) memcpy(q, hp, sizeof(double) * n_hp);
    double* dev = nullptr;
    SMC_CHECK_CUDA(cudaMallocAsync(&dev, cnt * sizeof(double), st));
    SMC_CHECK_CUDA(cudaMemcpyAsync(dev, host.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    LedhInfoArgs a;
    a.eta0 = static_cast<const float*>(d->eta0_dev);
    a.eta = static_cast<float*>(d->eta_dev);
    a.theta = static_cast<float*>(d->theta_log_dev);
    a.s0 = d->s0_dev;
    a.Lc = dev; a.R = dev + nx * nx; a.z = a.R + nz * nz; a.hp = a.z + nz;
    a.n = d->n; a.ld = d->ld; a.nz = nz; a.n_hp = n_hp > 0 ? n_hp : 0; a.L = L; a.step0 = step0;
    int rc = -1;
#define SMC_LEDH_INFO(NX_, KIND_)                                                                           \
    if (nx == NX_ && kind == KIND_) {                                                                       \
        auto kf = k4_ledh_flow_info<NX_, KIND_>;                                                            \
        const size_t sm = LedhInfoSmem<NX_>::bytes(nz, a.n_hp);                                             \
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
        const int ppc = kLedhInfoWarps * (32 / LedhInfoLanes<NX_>::value);                                  \
        kf<<<(unsigned)cdiv64(d->n, ppc), 32 * kLedhInfoWarps, sm, st>>>(a);                                \
        SMC_CHECK_CUDA(cudaGetLastError());                                                                 \
        rc = 0;                                                                                             \
    }
    SMC_LEDH_INFO(16, SMC_LEDH_H_ACOUSTIC) SMC_LEDH_INFO(8, SMC_LEDH_H_ACOUSTIC) SMC_LEDH_INFO(4, SMC_LEDH_H_ACOUSTIC)
    SMC_LEDH_INFO(16, SMC_LEDH_H_LINEAR) SMC_LEDH_INFO(8, SMC_LEDH_H_LINEAR) SMC_LEDH_INFO(4, SMC_LEDH_H_LINEAR)
    SMC_LEDH_INFO(16, SMC_LEDH_H_EXP) SMC_LEDH_INFO(8, SMC_LEDH_H_EXP) SMC_LEDH_INFO(4, SMC_LEDH_H_EXP)
#undef SMC_LEDH_INFO
    SMC_CHECK_CUDA(cudaFreeAsync(dev, st));
    return rc;
}

template <typename real>
int launch_flow(const smc_ledh_desc* d, const double* P, const double* R, const double* z, int kind, const double* hp,
                int n_hp, const double* Hsup, const double* hsup, double lam0, double dlam, int step0, int L,
                cudaStream_t st) {
    const int nx = d->nx, nz = d->nz;
    if (sizeof(real) == 4 && !Hsup) {   // fp32 throughput mode: nx x nx information form where it applies
        const int rc = launch_flow_info(d, P, R, z, kind, hp, n_hp, step0, L, st);
        if (rc >= 0) return rc;
    }
    std::vector<double> Rinv;
    SMC_REQUIRE(invert(R, nz, Rinv), "smc_ledh: R is singular");
    // stream-ordered staging buffer: P | R | Rinv | z | hp
    const size_t cnt = (size_t)nx * nx + 2 * (size_t)nz * nz + nz + (size_t)(n_hp > 0 ? n_hp : 0);
    std::vector<double> host(cnt);
    double* q = host.data();
    memcpy(q, P, sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, R, sizeof(double) * nz * nz); q += nz * nz;
    memcpy(q, Rinv.data(), sizeof(double) * nz * nz); q += nz * nz;
    memcpy(q, z, sizeof(double) * nz); q += nz;
    if (n_hp > 0) memcpy(q, hp, sizeof(double) * n_hp);
    double* dev = nullptr;
    SMC_CHECK_CUDA(cudaMallocAsync(&dev, cnt * sizeof(double), st));
    SMC_CHECK_CUDA(cudaMemcpyAsync(dev, host.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    LedhArgs<real> a;
    a.eta0 = static_cast<const real*>(d->eta0_dev);
    a.eta = static_cast<real*>(d->eta_dev);
    a.theta = static_cast<real*>(d->theta_log_dev);
    a.s0 = d->s0_dev;
    a.P = dev; a.R = dev + nx * nx; a.Rinv = a.R + nz * nz; a.z = a.Rinv + nz * nz; a.hp = a.z + nz;
    a.Hsup = Hsup; a.hsup = hsup;
    a.n = d->n; a.ld = d->ld; a.nx = nx; a.nz = nz; a.n_hp = n_hp > 0 ? n_hp : 0; a.kind = kind;
    a.L = L; a.step0 = step0; a.lam0 = lam0; a.dlam = dlam;
    // compile-time-dimension fast path (lane = row, matrices in registers) for the benchmarked shapes
    if (!Hsup) {
#define SMC_LEDH_FAST(NX_, NZ_)                                                                          \
        if (nx == NX_ && nz == NZ_) {                                                                    \
            auto kf = k4_ledh_flow_fast<real, NX_, NZ_>;                                                 \
            const size_t sm = LedhFastSmem<real, NX_, NZ_>::bytes(a.n_hp);                               \
            SMC_CHECK_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
            kf<<<(unsigned)cdiv64(d->n, kLedhFastWarps), 32 * kLedhFastWarps, sm, st>>>(a);              \
            SMC_CHECK_CUDA(cudaGetLastError());                                                          \
            SMC_CHECK_CUDA(cudaFreeAsync(dev, st));                                                      \
            return 0;                                                                                    \
        }
        SMC_LEDH_FAST(16, 25) SMC_LEDH_FAST(8, 16) SMC_LEDH_FAST(4, 9)
#undef SMC_LEDH_FAST
    }
    const size_t per_warp = 3 * (size_t)nz * nx + (size_t)nz * nz + 2 * (size_t)nx * nx + 5 * (size_t)nx + 3 * (size_t)nz;
    const size_t n_const = (size_t)nx * nx + 2 * (size_t)nz * nz + nz + a.n_hp;
    const size_t smem = sizeof(real) * (n_const + (size_t)kLedhWarps * per_warp);
    a.gws = nullptr;
    if (smem <= 227 * 1024 && nx <= kLedhMaxDim && nz <= kLedhMaxDim) {
        auto kern = k4_ledh_flow<real, false>;
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)cdiv64(d->n, kLedhWarps);
        kern<<<grid, 32 * kLedhWarps, smem, st>>>(a);
        SMC_CHECK_CUDA(cudaGetLastError());
    } else {
        // dimensions beyond shared memory: per-warp matrices in a stream-ordered global workspace, persistent grid
        int sms = 148, dev_id = 0;
        if (cudaGetDevice(&dev_id) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
        int64_t grid = cdiv64(d->n, kLedhWarps);
        if (grid > (int64_t)sms * 4) grid = (int64_t)sms * 4;
        const size_t ws_elems = n_const + (size_t)grid * kLedhWarps * per_warp;
        real* gws = nullptr;
        SMC_CHECK_CUDA(cudaMallocAsync(&gws, ws_elems * sizeof(real), st));
        a.gws = gws;
        k4_ledh_stage_constants<real><<<32, 256, 0, st>>>(a);
        SMC_CHECK_CUDA(cudaGetLastError());
        k4_ledh_flow<real, true><<<(unsigned)grid, 32 * kLedhWarps, 0, st>>>(a);
        SMC_CHECK_CUDA(cudaGetLastError());
        SMC_CHECK_CUDA(cudaFreeAsync(gws, st));
    }
    SMC_CHECK_CUDA(cudaFreeAsync(dev, st));
    return 0;
}

int check_ledh(const smc_ledh_desc* d) {
    SMC_REQUIRE(d != nullptr, "smc_ledh: null descriptor");
    SMC_REQUIRE(d->n > 0 && d->ld >= d->n, "smc_ledh: bad n / ld");
    SMC_REQUIRE(d->nx >= 1 && d->nx <= kLedhBigMaxDim && d->nz >= 1 && d->nz <= kLedhBigMaxDim,
                "smc_ledh: dimensions must satisfy 1 <= nx, nz <= %d (got %d, %d)", kLedhBigMaxDim, d->nx, d->nz);
    SMC_REQUIRE(d->dtype == SMC_F32 || d->dtype == SMC_F64, "smc_ledh: bad dtype");
    SMC_REQUIRE(d->eta0_dev && d->eta_dev && d->theta_log_dev, "smc_ledh: null buffer");
    return 0;
}

}  // namespace

extern "C" {

uint64_t smc_sizeof_ledh_desc(void) { return sizeof(smc_ledh_desc); }
void smc_ledh_info_counters(uint64_t* jittered, uint64_t* fell_back) {
    if (jittered) *jittered = __atomic_load_n(&g_info_jitter, __ATOMIC_RELAXED);
    if (fell_back) *fell_back = __atomic_load_n(&g_info_fallback, __ATOMIC_RELAXED);
}

int smc_ledh_flow(const smc_ledh_desc* d, const double* P_host, const double* R_host, const double* z_host,
                  int32_t h_kind, const double* h_params_host, int32_t n_h_params, void* stream) {
    if (int rc = check_ledh(d)) return rc;
    SMC_REQUIRE(P_host && R_host && z_host, "smc_ledh: null parameter array");
    const int nx = d->nx, nz = d->nz;
    if (h_kind == SMC_LEDH_H_LINEAR) SMC_REQUIRE(n_h_params == nz * nx, "smc_ledh: LINEAR needs nz*nx parameters");
    else if (h_kind == SMC_LEDH_H_ACOUSTIC)
        SMC_REQUIRE(n_h_params == 2 + 2 * nz && nx % 4 == 0, "smc_ledh: ACOUSTIC needs 2 + 2 nz parameters and nx %% 4 == 0");
    else if (h_kind == SMC_LEDH_H_EXP) SMC_REQUIRE(n_h_params == 1 && nx == nz, "smc_ledh: EXP needs 1 parameter and nx == nz");
    else SMC_REQUIRE(false, "smc_ledh: unknown measurement functor %d", h_kind);
    const int L = d->n_lambda_steps < 1 ? 1 : d->n_lambda_steps;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return d->dtype == SMC_F32
               ? launch_flow<float>(d, P_host, R_host, z_host, h_kind, h_params_host, n_h_params, nullptr, nullptr, 0, 0, 0, L, st)
               : launch_flow<double>(d, P_host, R_host, z_host, h_kind, h_params_host, n_h_params, nullptr, nullptr, 0, 0, 0, L, st);
}

int smc_ledh_flow_step_supplied(const smc_ledh_desc* d, const double* P_host, const double* R_host, const double* z_host,
                                const double* H_dev, const double* h_dev, double lam, double dlam, int32_t step_index,
                                void* stream) {
    if (int rc = check_ledh(d)) return rc;
    SMC_REQUIRE(P_host && R_host && z_host && H_dev && h_dev, "smc_ledh: null parameter array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return d->dtype == SMC_F32
               ? launch_flow<float>(d, P_host, R_host, z_host, 0, nullptr, 0, H_dev, h_dev, lam, dlam, step_index, 1, st)
               : launch_flow<double>(d, P_host, R_host, z_host, 0, nullptr, 0, H_dev, h_dev, lam, dlam, step_index, 1, st);
}

int smc_affine_map(void* x_dev, const double* M_host, const double* c_host, int64_t n, int64_t ld, int32_t nx,
                   int32_t dtype, void* stream) {
    SMC_REQUIRE(x_dev && M_host && c_host && n > 0 && ld >= n, "smc_affine_map: bad arguments");
    SMC_REQUIRE(nx >= 1 && nx <= kLedhMaxDim, "smc_affine_map: nx must be in [1, %d]", kLedhMaxDim);
    SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_affine_map: bad dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t cnt = (size_t)nx * nx + nx;
    std::vector<double> host(cnt);
    memcpy(host.data(), M_host, sizeof(double) * nx * nx);
    memcpy(host.data() + nx * nx, c_host, sizeof(double) * nx);
    double* dev = nullptr;
    SMC_CHECK_CUDA(cudaMallocAsync(&dev, cnt * sizeof(double), st));
    SMC_CHECK_CUDA(cudaMemcpyAsync(dev, host.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    SMC_CHECK_CUDA(cudaStreamSynchronize(st));   // `host` leaves scope; the copy is tiny
    const unsigned grid = (unsigned)cdiv64(n, kThreads);
    const size_t smem = cnt * sizeof(double);
    if (dtype == SMC_F32)
        k_affine_map<float><<<grid, kThreads, smem, st>>>((float*)x_dev, dev, dev + nx * nx, n, ld, nx);
    else
        k_affine_map<double><<<grid, kThreads, smem, st>>>((double*)x_dev, dev, dev + nx * nx, n, ld, nx);
    SMC_CHECK_CUDA(cudaGetLastError());
    SMC_CHECK_CUDA(cudaFreeAsync(dev, st));
    return 0;
}

int smc_ledh_propagate_linear(const smc_ledh_desc* d, const void* x_dev, const double* F_host, const double* Lq_host,
                              const void* v_dev, uint64_t seed, int64_t t, void* stream) {
    if (int rc = check_ledh(d)) return rc;
    SMC_REQUIRE(x_dev && F_host && Lq_host, "smc_ledh_propagate_linear: null argument");
    SMC_REQUIRE(d->nx <= kLedhMaxDim, "smc_ledh_propagate_linear: nx <= %d (per-thread state vectors)", kLedhMaxDim);
    SMC_REQUIRE(d->ld % 4 == 0, "smc_ledh_propagate_linear: ld must be a multiple of 4");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nx = d->nx;
    std::vector<double> host(2 * (size_t)nx * nx);
    memcpy(host.data(), F_host, sizeof(double) * nx * nx);
    memcpy(host.data() + nx * nx, Lq_host, sizeof(double) * nx * nx);
    double* dev = nullptr;
    SMC_CHECK_CUDA(cudaMallocAsync(&dev, host.size() * sizeof(double), st));
    SMC_CHECK_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned)cdiv64((d->n + 3) / 4, kThreads);
    const size_t smem = 2 * (size_t)nx * nx * sizeof(double);
    if (d->dtype == SMC_F32)
        k_ledh_propagate_linear<float><<<grid, kThreads, smem, st>>>((const float*)x_dev, (float*)d->eta0_dev, (float*)d->eta_dev,
                                                                    (float*)d->theta_log_dev, (const float*)v_dev, dev,
                                                                    dev + nx * nx, d->n, d->ld, nx, seed, (uint32_t)t);
    else
        k_ledh_propagate_linear<double><<<grid, kThreads, smem, st>>>((const double*)x_dev, (double*)d->eta0_dev,
                                                                     (double*)d->eta_dev, (double*)d->theta_log_dev,
                                                                     (const double*)v_dev, dev, dev + nx * nx, d->n, d->ld, nx,
                                                                     seed, (uint32_t)t);
    SMC_CHECK_CUDA(cudaGetLastError());
    SMC_CHECK_CUDA(cudaFreeAsync(dev, st));
    return 0;
}

int smc_ledh_gauss_terms(const smc_ledh_desc* d, const void* x_dev, const double* F_host, const double* Qinv_host,
                         const double* Rinv_host, const double* z_host, int32_t h_kind, const double* h_params_host,
                         int32_t n_h_params, double* terms_dev, void* stream) {
    if (int rc = check_ledh(d)) return rc;
    SMC_REQUIRE(x_dev && F_host && Qinv_host && Rinv_host && z_host && terms_dev && h_params_host && n_h_params > 0,
                "smc_ledh_gauss_terms: null argument");
    SMC_REQUIRE(d->nx <= kLedhMaxDim && d->nz <= kLedhMaxDim, "smc_ledh_gauss_terms: nx, nz <= %d (per-thread state vectors)",
                kLedhMaxDim);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nx = d->nx, nz = d->nz;
    const size_t cnt = 2 * (size_t)nx * nx + (size_t)nz * nz + nz + n_h_params;
    std::vector<double> host(cnt);
    double* q = host.data();
    memcpy(q, F_host, sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, Qinv_host, sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, Rinv_host, sizeof(double) * nz * nz); q += nz * nz;
    memcpy(q, z_host, sizeof(double) * nz); q += nz;
    memcpy(q, h_params_host, sizeof(double) * n_h_params);
    double* dev = nullptr;
    SMC_CHECK_CUDA(cudaMallocAsync(&dev, cnt * sizeof(double), st));
    SMC_CHECK_CUDA(cudaMemcpyAsync(dev, host.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    const double *dF = dev, *dQ = dF + nx * nx, *dR = dQ + nx * nx, *dz = dR + nz * nz, *dh = dz + nz;
    const unsigned grid = (unsigned)cdiv64(d->n, 128);
    const size_t smem = cnt * sizeof(double);
    if (d->dtype == SMC_F32)
        k_ledh_gauss_terms<float><<<grid, 128, smem, st>>>((const float*)x_dev, (const float*)d->eta0_dev,
                                                          (const float*)d->eta_dev, dF, dQ, dR, dz, dh, n_h_params, h_kind,
                                                          d->n, d->ld, nx, nz, terms_dev);
    else
        k_ledh_gauss_terms<double><<<grid, 128, smem, st>>>((const double*)x_dev, (const double*)d->eta0_dev,
                                                           (const double*)d->eta_dev, dF, dQ, dR, dz, dh, n_h_params,
                                                           h_kind, d->n, d->ld, nx, nz, terms_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    SMC_CHECK_CUDA(cudaFreeAsync(dev, st));
    return 0;
}

int smc_ledh_weights(const smc_sir_desc* d, const double* wprev_dev, const void* theta_log_dev, const double* terms_dev,
                     void* stream) {
    SMC_REQUIRE(d && d->n > 0 && d->ws_dev && d->logw_dev, "smc_ledh_weights: bad descriptor");
    SMC_REQUIRE(d->flags & SMC_FLAG_GENERIC_MOMENTS, "smc_ledh_weights: needs SMC_FLAG_GENERIC_MOMENTS (3-wide partials)");
    SMC_REQUIRE(wprev_dev && theta_log_dev && terms_dev, "smc_ledh_weights: null buffer");
    const WsLayout L = ws_layout(d->n, d->nx, d->dtype);
    SMC_REQUIRE(d->ws_bytes >= L.total, "smc_ledh_weights: workspace too small");
    double* parts = reinterpret_cast<double*>(static_cast<unsigned char*>(d->ws_dev) + L.k1_parts);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)L.n_k1_blocks;
    if (d->dtype == SMC_F32)
        k_ledh_weights<float><<<grid, kThreads, 0, st>>>((float*)d->logw_dev, wprev_dev, (const float*)theta_log_dev,
                                                        terms_dev, d->n, parts);
    else
        k_ledh_weights<double><<<grid, kThreads, 0, st>>>((double*)d->logw_dev, wprev_dev, (const double*)theta_log_dev,
                                                         terms_dev, d->n, parts);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_weighted_moments(const smc_sir_desc* d, const double* w_dev, double* stats_row_dev, void* stream) {
    SMC_REQUIRE(d && d->n > 0 && d->ws_dev && stats_row_dev, "smc_weighted_moments: bad arguments");
    const WsLayout L = ws_layout(d->n, d->nx, d->dtype);
    SMC_REQUIRE(d->ws_bytes >= L.total, "smc_weighted_moments: workspace too small");
    const Ctrl* ctrl = reinterpret_cast<const Ctrl*>(static_cast<unsigned char*>(d->ws_dev) + L.ctrl);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nx = d->nx;
    const unsigned grid = (unsigned)(nx + nx * (nx + 1) / 2);
    if (d->dtype == SMC_F32)
        k_weighted_moments<float><<<grid, kThreads, 0, st>>>(ctrl, (const float*)d->x0_dev, (const float*)d->x1_dev, w_dev,
                                                            d->n, d->ld, nx, stats_row_dev);
    else
        k_weighted_moments<double><<<grid, kThreads, 0, st>>>(ctrl, (const double*)d->x0_dev, (const double*)d->x1_dev, w_dev,
                                                             d->n, d->ld, nx, stats_row_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    k_weighted_moments_fix<<<1, 32, 0, st>>>(ctrl, nx, stats_row_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
