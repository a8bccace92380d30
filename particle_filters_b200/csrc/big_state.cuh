// High-dimensional state path of the SIR step (nx > 4), included by sir_kernels.cu:
//   * k1_lorenz96: fused propagate + log-likelihood for the Lorenz-96 model of the reference's
//     simulator (simulator/simulator_Lorenz_96.py: l96_rhs :35-58, rk4_step :61-83), one particle per
//     thread with the whole state in registers (the RK4 stages are fully unrolled stencils);
//   * k_moments_syrk: weighted mean / covariance of an (n, nx) particle set as a register-blocked
//     X^T W X reduction (particle_filter.py:266-267) -- one pass over the particles instead of one
//     pass per covariance entry.
// BASELINE.json configs[3] ("C4"): nx = 40, every 4th variable observed, 20 RK4 sub-steps per
// observation: ~1.9e4 FLOP per particle-step against 40 x 8 B of traffic -> bound by fp32 FMA issue.
#pragma once

namespace smc {

constexpr int kL96MaxNx = 64;
constexpr int kL96Threads = 128;

template <typename real> struct L96Params {
    real F, dt;
    int32_t n_sub, obs_stride, nz;
    real lq[kL96MaxNx];      // diagonal of chol(Q)
    real rinv[kL96MaxNx];    // 1 / diagonal of chol(R + 1e-12 I)
};

// products / sums that must not be contracted in the fp64 parity mode (NumPy evaluates
// (xp1 - xm2) * xm1 - x + F with one rounding per operation)
__device__ __forceinline__ float l96_term(float xp1, float xm2, float xm1, float x, float F) {
    return fmaf(xp1 - xm2, xm1, F - x);
}
__device__ __forceinline__ double l96_term(double xp1, double xm2, double xm1, double x, double F) {
    return __dadd_rn(__dsub_rn(__dmul_rn(__dsub_rn(xp1, xm2), xm1), x), F);
}
__device__ __forceinline__ float axpy_nc(float a, float k, float x) { return fmaf(a, k, x); }
__device__ __forceinline__ double axpy_nc(double a, double k, double x) { return __dadd_rn(x, __dmul_rn(a, k)); }

template <typename real, int NX>
__device__ __forceinline__ void l96_rhs_reg(const real (&x)[NX], real F, real (&k)[NX]) {
#pragma unroll
    for (int i = 0; i < NX; ++i)
        k[i] = l96_term(x[(i + 1) % NX], x[(i + NX - 2) % NX], x[(i + NX - 1) % NX], x[i], F);
}

// one RK4 step, x <- x + dt/6 (k1 + 2 k2 + 2 k3 + k4)   (rk4_step :78-83, same association).
// Stages 2 and 3 update the stage state IN PLACE while the stencil walks over it (the two trailing old
// values ride in scalars, the old first entry is kept for the wrap-around), so only x, the stage state
// and the k-sum are live: 3 NX registers instead of 4 NX (NX = 40: 128 registers, 4 CTAs per SM; measured
// time-neutral against 168 registers / 3 CTAs -- the kernel sits on the fp32 pipe, 55 % of its issue rate).
// Every element sees exactly the operations of the array form, in the same order.
template <typename real, int NX, int STAGE>
__device__ __forceinline__ void l96_stage(real (&x)[NX], real (&xt)[NX], real (&acc)[NX], real F, real c) {
    const real (&src)[NX] = (STAGE == 1) ? x : xt;
    const real old0 = src[0];
    real d2 = src[NX - 2], d1 = src[NX - 1];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const real cur = src[i];
        const real nxt = (i + 1 < NX) ? src[(i + 1) % NX] : old0;
        const real k = l96_term(nxt, d2, d1, cur, F);
        if (STAGE == 1) { acc[i] = k; xt[i] = axpy_nc(c, k, x[i]); }
        else if (STAGE < 4) { acc[i] = axpy_nc(real(2), k, acc[i]); xt[i] = axpy_nc(c, k, x[i]); }
        else x[i] = axpy_nc(c, acc[i] + k, x[i]);
        d2 = d1; d1 = cur;
    }
}
template <typename real, int NX>
__device__ __forceinline__ void l96_rk4(real (&x)[NX], real F, real dt) {
    real xt[NX], acc[NX];
    const real hdt = real(0.5) * dt;
    l96_stage<real, NX, 1>(x, xt, acc, F, hdt);
    l96_stage<real, NX, 2>(x, xt, acc, F, hdt);
    l96_stage<real, NX, 3>(x, xt, acc, F, dt);
    l96_stage<real, NX, 4>(x, xt, acc, F, dt / real(6));
}

// fp32 throughput form of the same step: ONE stage body in a loop over the four stages.  The fully unrolled form
// above is 760 floating-point instructions (12 KB of code) per RK4 step, and ncu showed the kernel's top stall to be
// "no instruction" (1.9 stalled warps per issue, 57 % issue utilisation): four warps per scheduler at four places of
// a loop that does not fit the instruction cache.  Here the stage state is walked in place by all four stages
// (stage 1 starts from a copy of x), acc collects k1 + 2 k2 + 2 k3 + k4 by one FMA per stage (1 k + 0 and acc + 1 k are
// exact), and the stage-4 update x += dt/6 acc follows the loop: every element sees the same operations with the same
// roundings as in the unrolled form (bit-identical results) at 1.2 x the instructions and a quarter of the code.
template <int NX>
__device__ __forceinline__ void l96_rk4_compact(float (&x)[NX], float F, float dt) {
    float xt[NX], acc[NX];
    const float hdt = 0.5f * dt;
#pragma unroll
    for (int i = 0; i < NX; ++i) { xt[i] = x[i]; acc[i] = 0.f; }
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const float c = s < 2 ? hdt : dt;                       // (stage 4's stage state is not used)
        const float ca = (s == 0 || s == 3) ? 1.f : 2.f;
        const float old0 = xt[0];
        float d2 = xt[NX - 2], d1 = xt[NX - 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const float cur = xt[i];
            const float nxt = (i + 1 < NX) ? xt[(i + 1) % NX] : old0;
            const float k = l96_term(nxt, d2, d1, cur, F);
            acc[i] = fmaf(ca, k, acc[i]);
            xt[i] = fmaf(c, k, x[i]);
            d2 = d1; d1 = cur;
        }
    }
    const float c6 = dt / 6.f;
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = fmaf(c6, acc[i], x[i]);
}
// What ships: stages 1 and 4 as in the unrolled form (no copy of x, no wasted FMA), stages 2 and 3 share one body --
// the unrolled form's 760 floating-point instructions per step in 9 KB of code.  Measured at C4 (N = 2^22, nx = 40, 20
// sub-steps; ms per filter step): unrolled 5.67 (kernel 4.11, 57 % issue), compact 5.27 (kernel 3.70, 75 % issue, "no
// instruction" stall 0.06), peeled 4.73; all three bit-identical (profiles/micro/l96_hash.py).  -DSMC_L96_COMPACT /
// -DSMC_L96_UNROLLED select the other two.
template <int NX>
__device__ __forceinline__ void l96_rk4_peeled(float (&x)[NX], float F, float dt) {
    float xt[NX], acc[NX];
    const float hdt = 0.5f * dt;
    l96_stage<float, NX, 1>(x, xt, acc, F, hdt);
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const float c = s == 0 ? hdt : dt;
        const float old0 = xt[0];
        float d2 = xt[NX - 2], d1 = xt[NX - 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const float cur = xt[i];
            const float nxt = (i + 1 < NX) ? xt[(i + 1) % NX] : old0;
            const float k = l96_term(nxt, d2, d1, cur, F);
            acc[i] = fmaf(2.f, k, acc[i]);
            xt[i] = fmaf(c, k, x[i]);
            d2 = d1; d1 = cur;
        }
    }
    l96_stage<float, NX, 4>(x, xt, acc, F, dt / 6.f);
}
template <typename real, int NX> struct L96Step {
    static __device__ __forceinline__ void run(real (&x)[NX], real F, real dt) { l96_rk4<real, NX>(x, F, dt); }
};
#ifndef SMC_L96_UNROLLED
template <int NX> struct L96Step<float, NX> {
#ifdef SMC_L96_COMPACT
    static __device__ __forceinline__ void run(float (&x)[NX], float F, float dt) { l96_rk4_compact<NX>(x, F, dt); }
#else
    static __device__ __forceinline__ void run(float (&x)[NX], float F, float dt) { l96_rk4_peeled<NX>(x, F, dt); }
#endif
};
#endif

// the standard normal that draw_normals() hands to particle `gidx`, dimension d (same Philox block,
// same Box-Muller lane), so results do not depend on which kernel propagates the particle
__device__ inline float normal_one(const PhiloxKeys& keys, uint64_t gidx, uint32_t t, uint32_t d, float,
                                   uint32_t stream = kStreamProcess) {
    uint32_t r[4];
    philox_quad_keys(keys, gidx >> 2, t, stream | (d << 8), r);
    const int j = (int)(gidx & 3), p = j >> 1;
    const float k2pow32inv = 2.3283064365386963e-10f, k2pi = 6.2831853071795865f;
    const float u1 = fmaf((float)(p ? r[2] : r[0]), k2pow32inv, 0.5f * k2pow32inv);
    const float u2 = fmaf((float)(p ? r[3] : r[1]), k2pow32inv * k2pi, 0.5f * k2pow32inv * k2pi);
    const float rad = sqrt_approx(-1.3862943611198906f * lg2_approx(u1));
    float s, c;
    __sincosf(u2, &s, &c);
    return (j & 1) ? rad * s : rad * c;
}
__device__ inline double normal_one(const PhiloxKeys& keys, uint64_t gidx, uint32_t t, uint32_t d, double,
                                    uint32_t stream = kStreamProcess) {
    uint32_t r[4];
    const int j = (int)(gidx & 3), h = j >> 1;
    philox_quad_keys(keys, gidx >> 2, t, stream | (d << 8) | ((uint32_t)h << 16), r);
    double z2[2];
    normals2_f64(r, z2);
    return z2[j & 1];
}

#ifndef SMC_L96_MINB
#define SMC_L96_MINB 4      // fp32: CTAs per SM the register allocation is held to (3 NX live registers per thread)
#endif
template <typename real, int NX>
__global__ void __launch_bounds__(kL96Threads, sizeof(real) == 4 ? SMC_L96_MINB : 1) k1_lorenz96(K1Args<real> a, L96Params<real> p) {
    constexpr int kW = kL96Threads / 32;
    __shared__ real s_z[kL96MaxNx];
    __shared__ real s_max[kW];
    __shared__ double s_red[kW][2];
    pdl_wait();
    pdl_trigger();
    const Ctrl* ctrl = a.ctrl;
    real* x_all = ctrl->cur ? a.x1 : a.x0;
    const bool uniform = ctrl->uniform != 0;
    const real log_norm_prev = (real)ctrl->log_norm;
    const real lw_uniform = (real)(-log((double)a.n_global));
    const bool do_prop = a.mode & SMC_MODE_PROPAGATE, do_w = a.mode & SMC_MODE_WEIGHT;
    const int tid = threadIdx.x;
    if (do_w && tid < p.nz) s_z[tid] = (real)a.z[tid];
    __syncthreads();

    const real NEG_INF = -INFINITY;
    real tm = NEG_INF, acc0 = 0, acc1 = 0;     // online (max, sum e, sum e^2)
    // (warp-uniform trip count: the four lanes of a particle quad share their Philox blocks by shuffles)
    const int lane = tid & 31;
    const bool share_philox = sizeof(real) == 4 && a.noise == nullptr && (a.offset & 3) == 0 && NX % 4 == 0;
    for (int64_t i0 = (int64_t)blockIdx.x * kL96Threads + (tid & ~31); i0 < a.n; i0 += (int64_t)gridDim.x * kL96Threads) {
        const int64_t i = i0 + lane;
        const bool live = i < a.n;
        real x[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] = live ? x_all[d * a.ld + i] : real(0);
        if (do_prop) {
            for (int s = 0; s < p.n_sub; ++s) L96Step<real, NX>::run(x, p.F, p.dt);
            if (share_philox) {
                // draw_normals' layout: particle quad q = gidx >> 2, dimension d -> ONE Philox block (q, t, stream | d << 8)
                // whose words (0, 1) / (2, 3) make the normals of lanes (0, 1) / (2, 3) of the quad.  normal_one() computes
                // that block in each of the four lanes (40 blocks per particle, 15 % of this kernel's instructions); here
                // lane m of the quad computes the blocks of the dimensions d = m (mod 4) and hands the words over.
                const uint64_t gidx = (uint64_t)(a.offset + i);
                const int j = lane & 3, pr = j >> 1;
                const float k2pow32inv = 2.3283064365386963e-10f, k2pi = 6.2831853071795865f;
#pragma unroll
                for (int g = 0; g < NX / 4; ++g) {
                    uint32_t r[4];
                    philox_quad_keys(a.keys, gidx >> 2, a.t, kStreamProcess | ((uint32_t)(4 * g + j) << 8), r);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int src = (lane & ~3) | m;
                        const uint32_t w0 = __shfl_sync(0xffffffffu, r[0], src), w1 = __shfl_sync(0xffffffffu, r[1], src);
                        const uint32_t w2 = __shfl_sync(0xffffffffu, r[2], src), w3 = __shfl_sync(0xffffffffu, r[3], src);
                        const float u1 = fmaf((float)(pr ? w2 : w0), k2pow32inv, 0.5f * k2pow32inv);
                        const float u2 = fmaf((float)(pr ? w3 : w1), k2pow32inv * k2pi, 0.5f * k2pow32inv * k2pi);
                        const float rad = sqrt_approx(-1.3862943611198906f * lg2_approx(u1));
                        float sn, cs;
                        __sincosf(u2, &sn, &cs);
                        const real nrm = (real)((j & 1) ? rad * sn : rad * cs);
                        x[4 * g + m] = fma(p.lq[4 * g + m], nrm, x[4 * g + m]);
                    }
                }
#pragma unroll
                for (int d = 0; d < NX; ++d) if (live) x_all[d * a.ld + i] = x[d];
            } else {
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                    const real nrm = a.noise ? (live ? a.noise[d * a.ld + i] : real(0))
                                             : normal_one(a.keys, (uint64_t)(a.offset + i), a.t, (uint32_t)d, real(0));
                    x[d] = fma(p.lq[d], nrm, x[d]);
                    if (live) x_all[d * a.ld + i] = x[d];
                }
            }
        }
        if (do_w && live) {
            // h(x) = x[0 :: obs_stride]; y = (z - h(x)) / chol(R)_kk  (particle_filter.py:257-260, R diagonal)
            real quad = 0;
            int kobs = 0;
#pragma unroll
            for (int d = 0; d < NX; ++d) {
                if (kobs < p.nz && d == kobs * p.obs_stride) {
                    const real y = (s_z[kobs] - x[d]) * p.rinv[kobs];
                    quad = fma(y, y, quad);
                    ++kobs;
                }
            }
            const real prev = uniform ? lw_uniform : prev_logweight(a.logw[i], log_norm_prev);
            real v = prev - real(0.5) * quad;
            v = fmax(v, NEG_INF);                       // NaN -> -inf (zero weight)
            a.logw[i] = v;
            if (v > tm) {
                const real sc = fast_exp(tm - v);       // tm = -inf -> 0
                acc0 *= sc; acc1 *= sc * sc; tm = v;
            }
            if (tm != NEG_INF) {
                const real e = fast_exp(v - tm);
                acc0 += e; acc1 = fma(e, e, acc1);
            }
        }
    }
    if (!do_w) return;
    // ---- block combine (same record layout as k1_propagate_weight with 3 fields) -----------------
    const int warp = tid >> 5;
    const real wmax = warp_max(tm);
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    real bmax = s_max[0];
#pragma unroll
    for (int w = 1; w < kW; ++w) bmax = s_max[w] > bmax ? s_max[w] : bmax;
    const real sc = (tm == NEG_INF) ? real(0) : fast_exp(tm - bmax);
    const real v0 = warp_sum(acc0 * sc), v1 = warp_sum(acc1 * sc * sc);
    if (lane == 0) { s_red[warp][0] = (double)v0; s_red[warp][1] = (double)v1; }
    __syncthreads();
    if (tid == 0) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int w = 0; w < kW; ++w) { s0 += s_red[w][0]; s1 += s_red[w][1]; }
        double* out = a.parts + (int64_t)blockIdx.x * 3;
        out[0] = (double)bmax; out[1] = s0; out[2] = s1;
    }
    // (the K2 finalize runs as its own launch here: a 10 us kernel next to milliseconds of RK4)
}

// ==============================================================================================
// Weighted mean / covariance as one register-blocked X^T W X pass.
//   raw[0] = sum w, raw[1 + a] = sum w y_a, raw[1 + nx + a nx + b] = sum w y_a y_b,  y = x - shift
// CTA = 4 groups of 8 x 8 threads; a group takes every 4th particle of the staged tile and each of its
// threads owns a BS x BS block of the covariance (BS = 5 covers nx <= 40, BS = 8 covers nx <= 64).
// ==============================================================================================
constexpr int kSyrkTile = 128;     // particles staged per iteration
#ifndef SMC_SYRK_UNROLL
#define SMC_SYRK_UNROLL 4         // C4 step: 4.735 (1) / 4.613 (2) / 4.572 ms (4)
#endif
constexpr int kSyrkUnroll = SMC_SYRK_UNROLL;   // unroll factor of the per-particle rank-1 update loop (tuning knob)

template <typename real, int BS>
__global__ void __launch_bounds__(256, BS <= 5 ? 2 : 1) k_moments_syrk(const Ctrl* ctrl, const real* x0, const real* x1, const real* logw,
                                                       int64_t n, int64_t ld, int nx, int post_flip, double* parts) {
    constexpr int NP = 8 * BS;            // padded state dimension
    constexpr int LDS = NP + 1;           // row stride (odd: conflict-free column walks)
    extern __shared__ __align__(16) unsigned char syrk_smem[];
    real* ys = reinterpret_cast<real*>(syrk_smem);          // [kSyrkTile][LDS]  y
    real* yw = ys + kSyrkTile * LDS;                         // [kSyrkTile][LDS]  w * y
    real* ws = yw + kSyrkTile * LDS;                         // [kSyrkTile]
    const int cur = ctrl->cur ^ ((ctrl->do_resample && !post_flip) ? 1 : 0);
    const real* x = cur ? x1 : x0;
    const bool weighted = !ctrl->do_resample;
    const double log_norm = ctrl->log_norm;
    const int tid = threadIdx.x, grp = tid >> 6, ti = (tid >> 3) & 7, tj = tid & 7;
    __shared__ real s_shift[NP];
    if (tid < NP) s_shift[tid] = (real)(tid < kMaxNx ? ctrl->shift[tid] : 0.0);
    double acc[BS][BS], accm[BS], accw = 0;
#pragma unroll
    for (int a = 0; a < BS; ++a) {
        accm[a] = 0;
#pragma unroll
        for (int b = 0; b < BS; ++b) acc[a][b] = 0;
    }
    for (int64_t base = (int64_t)blockIdx.x * kSyrkTile; base < n; base += (int64_t)gridDim.x * kSyrkTile) {
        const int cnt = (int)((n - base) < kSyrkTile ? (n - base) : kSyrkTile);
        __syncthreads();
        // stage: thread handles particle (tid % 128), dimensions d = (tid / 128), +2, ... (coalesced over particles)
        {
            const int pl = tid & (kSyrkTile - 1);
            real w = 0;
            if (pl < cnt) w = weighted ? (real)exp((double)logw[base + pl] - log_norm) : real(1);
            if (tid < kSyrkTile) ws[pl] = w;
            // the thread's loads are issued in batches before the first is consumed (the tile is not
            // double-buffered, so the staging phase has to cover the HBM latency by itself: with one
            // load in flight per thread this phase was 85 % of the kernel)
            constexpr int ND = NP / 2, NB = ND % 10 == 0 ? 10 : 8;
            const bool live = pl < cnt;
            const int d0 = tid >> 7;
#pragma unroll
            for (int q0 = 0; q0 < ND; q0 += NB) {
                real v[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int d = d0 + 2 * (q0 + q);
                    v[q] = (live && d < nx) ? __ldg(x + (int64_t)d * ld + base + pl) : real(0);
                }
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int d = d0 + 2 * (q0 + q);
                    const real y = (live && d < nx) ? v[q] - s_shift[d] : real(0);
                    ys[pl * LDS + d] = y;
                    yw[pl * LDS + d] = w * y;
                }
            }
        }
        __syncthreads();
        real t_acc[BS][BS], t_m[BS], t_w = 0;
#pragma unroll
        for (int a = 0; a < BS; ++a) {
            t_m[a] = 0;
#pragma unroll
            for (int b = 0; b < BS; ++b) t_acc[a][b] = 0;
        }
#pragma unroll kSyrkUnroll
        for (int pl = grp; pl < cnt; pl += 4) {
            real ra[BS], rb[BS];
#pragma unroll
            for (int a = 0; a < BS; ++a) ra[a] = yw[pl * LDS + ti * BS + a];
#pragma unroll
            for (int b = 0; b < BS; ++b) rb[b] = ys[pl * LDS + tj * BS + b];
#pragma unroll
            for (int a = 0; a < BS; ++a)
#pragma unroll
                for (int b = 0; b < BS; ++b) t_acc[a][b] = fma(ra[a], rb[b], t_acc[a][b]);
            if (tj == 0) {
#pragma unroll
                for (int a = 0; a < BS; ++a) t_m[a] += ra[a];
                if (ti == 0) t_w += ws[pl];
            }
        }
#pragma unroll
        for (int a = 0; a < BS; ++a) {
            accm[a] += (double)t_m[a];
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[a][b] += (double)t_acc[a][b];
        }
        accw += (double)t_w;
    }
    // ---- fold the 4 groups through shared memory, then one record per CTA ------------------------
    __syncthreads();
    double* red = reinterpret_cast<double*>(syrk_smem);     // [4][1 + NP + NP * NP] doubles
    constexpr int REC = 1 + NP + NP * NP;
    double* mine = red + (size_t)grp * REC;
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) mine[1 + NP + (ti * BS + a) * NP + tj * BS + b] = acc[a][b];
    if (tj == 0) {
#pragma unroll
        for (int a = 0; a < BS; ++a) mine[1 + ti * BS + a] = accm[a];
        if (ti == 0) mine[0] = accw;
    }
    __syncthreads();
    const int rec_out = 1 + nx + nx * nx;
    double* out = parts + (size_t)blockIdx.x * rec_out;
    for (int e = tid; e < rec_out; e += 256) {
        int src;
        if (e == 0) src = 0;
        else if (e < 1 + nx) src = e;                                  // mean sums (unpadded index = padded index)
        else { const int q = e - 1 - nx; src = 1 + NP + (q / nx) * NP + (q % nx); }
        out[e] = red[src] + red[REC + src] + red[2 * REC + src] + red[3 * REC + src];
    }
}

// combine the per-CTA records (fixed order) into the stats row: mean, covariance.  Each CTA owns a slice of
// the covariance entries and re-derives the weight total and the mean sums it needs (same order everywhere).
__global__ void __launch_bounds__(256) k_moments_syrk_finish(const Ctrl* ctrl, const double* parts, int n_parts, int nx,
                                                              double* stats) {
    extern __shared__ __align__(16) unsigned char syrk_smem[];
    double* tot = reinterpret_cast<double*>(syrk_smem);     // [1 + nx] head, then this CTA's second moments
    const int rec = 1 + nx + nx * nx, head = 1 + nx;
    const int per = (nx * nx + (int)gridDim.x - 1) / (int)gridDim.x;
    const int e0 = (int)blockIdx.x * per, e1 = min(nx * nx, e0 + per);
    for (int e = threadIdx.x; e < head + (e1 - e0); e += 256) {
        const int src = e < head ? e : e0 + e;               // head + e0 + (e - head)
        double s = 0;
        for (int p = 0; p < n_parts; ++p) s += parts[(size_t)p * rec + src];
        tot[e] = s;
    }
    __syncthreads();
    double* mean = stats + SMC_STATS_HEADER;
    double* cov = mean + nx;
    const double W = tot[0];
    for (int e = e0 + threadIdx.x; e < e1; e += 256) {
        const int a = e / nx, b = e % nx;
        cov[e] = tot[head + e - e0] / W - (tot[1 + a] / W) * (tot[1 + b] / W);
    }
    if (blockIdx.x == 0)
        for (int a = threadIdx.x; a < nx; a += 256) mean[a] = (a < kMaxNx ? ctrl->shift[a] : 0.0) + tot[1 + a] / W;
}

// ==============================================================================================
// Post-resample regularisation on chip (particle_filter.py:212-218): x += n (scale Lq)^T with
// n ~ N(0, I) from the Philox jitter stream, applied to the resampling destination iff the control
// block says the step resampled.  Thread per particle; nx <= 64.
// ==============================================================================================
template <typename real>
__global__ void __launch_bounds__(kThreads) k_jitter(const Ctrl* ctrl, real* x0, real* x1, const double* Lq, double scale,
                                                     int64_t n, int64_t offset, int64_t ld, int nx, PhiloxKeys keys,
                                                     uint32_t t) {
    if (!ctrl->do_resample) return;
    real* x = ctrl->cur ? x0 : x1;                      // the buffer K3 has just filled
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    real z[kMaxNx];
    for (int d = 0; d < nx; ++d) z[d] = normal_one(keys, (uint64_t)(offset + i), t, (uint32_t)d, real(0), kStreamJitter);
    for (int d = nx - 1; d >= 0; --d) {
        double acc = 0;
        for (int k = 0; k <= d; ++k) acc = fma((double)z[k], Lq[d * nx + k] * scale, acc);
        x[d * ld + i] = (real)((double)x[d * ld + i] + acc);
    }
}

}  // namespace smc
