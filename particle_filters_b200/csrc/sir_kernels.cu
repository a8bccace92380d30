// SIR particle-filter hot path for sm_100a: K1 fused propagate+log-likelihood,
// K2 normaliser/ESS/moment reduction, K3 single-pass systematic resampler.
//
// Reference semantics: models/particle_filter.py of liyiyang-amber/Particle_filters
// (predict :223-237, update :239-269, _resample :188-220, _systematic_resample :146-171).
#include "smc_common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace smc {

// ----------------------------------------------------------------------------------------------
// error string (per thread)
// ----------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ==============================================================================================
// Model functors.  x is one particle (NX registers); propagate() applies g and adds Lq * nrm.
// ==============================================================================================
template <typename real> struct StochVolModel {
    static constexpr int NX = 1, NZ = 1;
    real alpha, lq, beta, rinv;
    __device__ inline void propagate(real (&x)[1], const real (&nrm)[1]) const {
        x[0] = fma(lq, nrm[0], alpha * x[0]);
    }
    __device__ inline real loglik(const real (&x)[1], const real (&z)[1]) const {
        real d = (z[0] - beta * half_exp(x[0])) * rinv;
        return real(-0.5) * d * d;
    }
    static __device__ inline float half_exp(float x) { return ex2_approx(x * 0.7213475204444817f); }   // exp(x/2)
    static __device__ inline double half_exp(double x) { return exp(0.5 * x); }
};

template <typename real, int NX_, int NZ_> struct LinGaussModel {
    static constexpr int NX = NX_, NZ = NZ_;
    real A[NX_ * NX_], Lq[NX_ * NX_], H[NZ_ * NX_], LRinv[NZ_ * NZ_];
    __device__ inline void propagate(real (&x)[NX_], const real (&nrm)[NX_]) const {
        real y[NX_];
#pragma unroll
        for (int i = 0; i < NX_; ++i) {
            real acc = 0;
#pragma unroll
            for (int j = 0; j < NX_; ++j) acc = fma(A[i * NX_ + j], x[j], acc);
#pragma unroll
            for (int j = 0; j <= i; ++j) acc = fma(Lq[i * NX_ + j], nrm[j], acc);
            y[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < NX_; ++i) x[i] = y[i];
    }
    __device__ inline real loglik(const real (&x)[NX_], const real (&z)[NZ_]) const {
        real d[NZ_];
#pragma unroll
        for (int i = 0; i < NZ_; ++i) {
            real acc = z[i];
#pragma unroll
            for (int j = 0; j < NX_; ++j) acc = fma(-H[i * NX_ + j], x[j], acc);
            d[i] = acc;
        }
        real quad = 0;
#pragma unroll
        for (int i = 0; i < NZ_; ++i) {   // y = LR^-1 d (lower triangular inverse)
            real acc = 0;
#pragma unroll
            for (int j = 0; j <= i; ++j) acc = fma(LRinv[i * NZ_ + j], d[j], acc);
            quad = fma(acc, acc, quad);
        }
        return real(-0.5) * quad;
    }
};

// ==============================================================================================
// K1: fused propagate + log-likelihood + block partials
// ==============================================================================================
// ==============================================================================================
// Multi-GPU mailboxes.  Every rank owns one Mailbox in peer-mapped device memory (CUDA IPC); a rank
// publishes its per-step records by storing them straight into the mailbox of every rank over
// NVLink and then raising a stamped flag there.  Consumers spin on the flags in their OWN memory,
// so the three per-step exchanges of the sharded filter (partial record, weight total, moment
// record) need no collective call and no extra kernel: the producer is the last CTA of the kernel
// that computes the record, the consumer the prologue of the kernel that needs it.
// Stamps increase by one per step; a slot can be reused every step because a rank can publish
// step t + 1 of an exchange only after every rank has consumed step t of it.
// ==============================================================================================
constexpr int kMaxShards = 16;
struct Mailbox {
    unsigned long long flag_rec[kMaxShards];
    unsigned long long flag_tot[kMaxShards];
    unsigned long long flag_mom[kMaxShards];
    unsigned long long tot[kMaxShards];
    double rec[kMaxShards][32];
    double mom[kMaxShards][16];
    unsigned long long error;    // != 0: a wait on this rank timed out (bit 63 | stamp << 8 | exchange << 4 | peer rank)
    // single-exchange fused step (k5_resample_propagate, sharded form): everything launch t publishes -- the record
    // of step t's log-weights, the post-resample moments of step t-1, the fixed-point weight this rank produced
    // for every rank's index range -- under ONE flag.  Two slots (stamp & 1): a peer may finish launch t before this
    // rank has read the slot of launch t-1, but cannot pass the prologue of launch t+1 without this rank's flag t.
    struct FusedSlot {
        unsigned long long flag[kMaxShards];
        unsigned long long contrib[kMaxShards][kMaxShards];    // [producing rank][rank that owns the indices]
        double rec[kMaxShards][4];                             // (max, sum e, sum e^2) of the producing rank
        double mom[kMaxShards][16];
    } fs[2];
};
using FusedSlot = Mailbox::FusedSlot;
struct ShardCtx {
    Mailbox* const* mail;        // device array of n_shards mailbox pointers (own one included)
    int32_t rank, n_shards;
    unsigned long long stamp;    // this step's stamp (> 0)
};
// Bounded spin: a rank that raised an error, ran a different number of steps or died must not hang its peers
// inside a kernel for ever.  After kMailboxTimeoutNs without the expected stamp the waiting thread records the
// failure in the own mailbox's `error` word (host: smc_mailbox_error / ShardedSir.check) and the kernel goes on
// with whatever the slots hold -- its results are then garbage, but every rank reaches its next host call.
constexpr unsigned long long kMailboxTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;     // 20 s
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ inline void mailbox_wait(const volatile unsigned long long* flags, int n, unsigned long long stamp,
                                    unsigned long long* error_word = nullptr, unsigned long long what = 0) {
    // threads 0 .. n-1 of the CTA each watch one flag; CTA-wide barrier afterwards
    if ((int)threadIdx.x < n) {
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        auto poll = [&]() {     // acquire load: every later read of this thread -- and, through the barrier, of the CTA --
            unsigned long long v;   // sees what the publisher stored before it released the flag
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(const_cast<const unsigned long long*>(flags) + threadIdx.x) : "memory");
            return v;
        };
        while (poll() != stamp) {
            __nanosleep(20);
            if ((++spins & 0x3ffu) == 0u && globaltimer_ns() - t0 > kMailboxTimeoutNs) {
                if (error_word) atomicMax(error_word, (stamp << 8) | (what << 4) | (unsigned long long)threadIdx.x | 0x8000000000000000ull);
                break;
            }
        }
        // (no fence: a system-wide fence after the poll measured 4 us per launch in all 256 threads, ~2 us in G)
    }
    __syncthreads();
}

// Fractional bits Fa of the fp32 mode's fixed-point weights: 2^Fa N must stay below 2^62
__host__ __device__ inline int fa_bits_for(int64_t n_global) {
    int lg = 0;
    while ((1ll << lg) < n_global) ++lg;
    const int F = 62 - lg;
    return F > 36 ? 36 : F;          // n_global <= 2^30 -> F >= 32
}
constexpr float kLog2eF = 1.4426950408889634f;
// a-priori bound of a bootstrap step's log-weights (the fp32 value K1 / the fused functor add the
// log-likelihood to) and the quantisation shift that goes with it
inline void quant_bound_for(int64_t n_global, double* lw_bound, float* kshift_bound) {
    const float lw_uniform = (float)(-log((double)n_global));
    *lw_bound = (double)lw_uniform;
    *kshift_bound = (float)((double)fa_bits_for(n_global) - (double)lw_uniform * (double)kLog2eF);
}

// first tile of pre-pass range / resampler CTA b of `grid` (see K3a)
__host__ __device__ inline int64_t k3_range_begin(int64_t b, int64_t n_tiles, int64_t grid) {
    return (b * n_tiles) / grid;
}
// fp32 quantisation of one log-weight: a = rni(p), p = 2^(lw log2e + kshift) (see "Per-item quantisation" below)
__device__ __forceinline__ float quant_p(float lw, float kshift) {
    return ex2_approx(fminf(fmaf(lw, kLog2eF, kshift), 62.5f));       // NaN-free input (see K1)
}
__device__ __forceinline__ unsigned long long quant_a(float p) {
    unsigned long long q;
    asm("cvt.rni.u64.f32 %0, %1;" : "=l"(q) : "f"(p));
    return q;
}
struct K2Args {
    Ctrl* ctrl;
    const double* parts;
    double* stats;     // nullable
    int64_t n_global;
    int32_t n_parts, nx, inline_moments, f32;
    int32_t rec_stride;   // doubles between records (0 = packed: the field count)
    double thresh, u_host;
    uint64_t seed;
    uint32_t t;
    // fp32 CDF (see below): the a-priori bound on the log-weights of a bootstrap step and its quantisation shift,
    // computed once on the host so that every kernel quantises with the very same constant
    double lw_bound;
    float kshift_bound;
    int32_t producer_rsum;   // the calling kernel accumulated the range totals against the bound
};

__device__ inline void finalize_body(const K2Args& a, double* s_out, double* s_tmp);   // defined with K2
__device__ inline void combine_partials(const double* parts, int n_parts, int ps, double* s_out, double* s_tmp,
                                        int stride);

template <typename real> struct K1Args {
    Ctrl* ctrl;
    real* x0;
    real* x1;
    real* logw;
    const real* noise;   // nullable: SoA standard normals, same ld
    const double* z;     // device, NZ doubles
    double* parts;       // block partials
    int64_t n, n_global, offset, ld;
    uint64_t seed;
    uint32_t t;
    int32_t mode;
    PhiloxKeys keys;            // the ten round keys of `seed`, expanded once on the host
    int32_t fuse_finalize;      // the last CTA to finish also runs the K2 finalize (fin)
    unsigned int* ticket;       // arrival counter for fuse_finalize (zero between launches)
    K2Args fin;
    ShardCtx shard;             // mail != nullptr: the last CTA publishes this shard's partial record instead
    // fp32 bootstrap steps (uniform prior weights): K1 also delivers the range totals of the resampler's CDF,
    // quantised against the a-priori bound (see finalize_body / fused_step.cuh); CTAs then own contiguous chunks
    unsigned long long* rsum;   // two buffers of kRsumStride totals (nullptr: not wanted)
    int32_t n_ranges;
    float kshift_bound;
};

template <typename real, int NX>
__device__ inline void draw_normals(uint64_t seed, uint64_t gquad, uint32_t t, real (&nrm)[NX][4]);

template <int NX> __device__ inline void draw_normals_f32(const PhiloxKeys& keys, uint64_t gquad, uint32_t t,
                                                          float (&nrm)[NX][4]) {
#pragma unroll
    for (int d = 0; d < NX; ++d) {
        uint32_t r[4];
        philox_quad_keys(keys, gquad, t, kStreamProcess | (uint32_t(d) << 8), r);
        normals4_f32(r, nrm[d]);
    }
}
template <int NX> __device__ inline void draw_normals_f64(const PhiloxKeys& keys, uint64_t gquad, uint32_t t,
                                                          double (&nrm)[NX][4]) {
#pragma unroll
    for (int d = 0; d < NX; ++d) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t r[4];
            philox_quad_keys(keys, gquad, t, kStreamProcess | (uint32_t(d) << 8) | (uint32_t(h) << 16), r);
            double z2[2];
            normals2_f64(r, z2);
            nrm[d][2 * h] = z2[0];
            nrm[d][2 * h + 1] = z2[1];
        }
    }
}
template <int NX> struct Draw {
    static __device__ inline void run(const PhiloxKeys& s, uint64_t q, uint32_t t, float (&n)[NX][4]) {
        draw_normals_f32<NX>(s, q, t, n);
    }
    static __device__ inline void run(const PhiloxKeys& s, uint64_t q, uint32_t t, double (&n)[NX][4]) {
        draw_normals_f64<NX>(s, q, t, n);
    }
};

__device__ inline float prev_logweight(float lw_old, float log_norm) { return lw_old - log_norm; }
// fp64 parity mode follows the reference literally: log(w + 1e-300), particle_filter.py:261
__device__ inline double prev_logweight(double lw_old, double log_norm) {
    return log(exp(lw_old - log_norm) + 1e-300);
}

// Programmatic dependent launch: the step's kernels are launched with the "programmatic stream
// serialization" attribute, so a kernel's CTAs are scheduled while its predecessor drains; nothing
// the predecessor produced may be touched before pdl_wait() (a no-op for ordinary launches).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Particle loads of one 4096-particle chunk of K1 (CHECK = the chunk may run past n: last chunk only).
template <typename real, int NX, bool CHECK>
__device__ __forceinline__ void k1_load(const K1Args<real>& a, const real* x, int64_t blk_base, int tid,
                                        real (&xs)[kK1Quads][NX][4]) {
#pragma unroll
    for (int q = 0; q < kK1Quads; ++q) {
        const int64_t i0 = blk_base + (int64_t)(q * kThreads + tid) * 4;
        if (!CHECK || i0 < a.n) {
#pragma unroll
            for (int d = 0; d < NX; ++d) load4<real>(x + d * a.ld + i0, xs[q][d]);
        } else {
#pragma unroll
            for (int d = 0; d < NX; ++d)
#pragma unroll
                for (int j = 0; j < 4; ++j) xs[q][d][j] = 0;
        }
    }
}

// One chunk of K1 on particles already in registers.  MOM = accumulate the weighted moments (not
// needed when the step is known to resample, because then the post-resample moments come from K3).
template <typename real, class Model, bool CHECK, bool MOM, int NR>
__device__ __forceinline__ void k1_chunk(const K1Args<real>& a, const Model& m, real* x, int64_t blk_base, int tid,
                                         bool uniform, real log_norm_prev, real lw_uniform, bool do_prop, bool do_w,
                                         const real (&z)[Model::NZ], const real (&shift)[Model::NX], real& tm,
                                         real (&acc)[NR], real (&xs)[kK1Quads][Model::NX][4],
                                         const bool do_rsum = false, unsigned long long* racc = nullptr) {
    constexpr int NX = Model::NX;
    const real NEG_INF = -INFINITY;
    real lw[kK1Quads][4];
#pragma unroll
    for (int q = 0; q < kK1Quads; ++q) {
        const int64_t i0 = blk_base + (int64_t)(q * kThreads + tid) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) lw[q][j] = NEG_INF;
        if (!CHECK || i0 < a.n) {
            if (do_prop) {
                real nrm[NX][4];
                if (a.noise) {
#pragma unroll
                    for (int d = 0; d < NX; ++d) load4<real>(a.noise + d * a.ld + i0, nrm[d]);
                } else {
                    Draw<NX>::run(a.keys, (uint64_t)(a.offset + i0) >> 2, a.t, nrm);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    real xv[NX], nv[NX];
#pragma unroll
                    for (int d = 0; d < NX; ++d) { xv[d] = xs[q][d][j]; nv[d] = nrm[d][j]; }
                    m.propagate(xv, nv);
#pragma unroll
                    for (int d = 0; d < NX; ++d) xs[q][d][j] = xv[d];
                }
#pragma unroll
                for (int d = 0; d < NX; ++d) store4<real>(x + d * a.ld + i0, xs[q][d]);
            }
            if (do_w) {
                real prev[4];
                if (uniform) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) prev[j] = lw_uniform;
                } else {
                    load4<real>(a.logw + i0, prev);
#pragma unroll
                    for (int j = 0; j < 4; ++j) prev[j] = prev_logweight(prev[j], log_norm_prev);
                }
                real out[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    real xv[NX];
#pragma unroll
                    for (int d = 0; d < NX; ++d) xv[d] = xs[q][d][j];
                    real v = prev[j] + m.loglik(xv, z);
                    if (CHECK && (i0 + j) >= a.n) v = NEG_INF;
                    v = fmax(v, NEG_INF);                 // NaN -> -inf (counts as zero weight)
                    out[j] = v;
                    lw[q][j] = v;
                }
                store4<real>(a.logw + i0, out);
                if constexpr (sizeof(real) == 4) {
                    if (do_rsum) {      // fixed-point weight against the a-priori bound (padding: -inf -> 0)
#pragma unroll
                        for (int j = 0; j < 4; ++j) *racc += quant_a(quant_p(out[j], a.kshift_bound));
                    }
                }
            }
        }
    }
    if (!do_w) return;
    // ---- fold this chunk into the thread's running (max, sums) ------------------------------
    real cmax = NEG_INF;
#pragma unroll
    for (int q = 0; q < kK1Quads; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) cmax = fmax(cmax, lw[q][j]);
    if (cmax > tm) {
        const real sc = fast_exp(tm - cmax);          // tm = -inf -> 0
        acc[0] *= sc;
        acc[1] *= sc * sc;
#pragma unroll
        for (int k = 2; k < NR; ++k) acc[k] *= sc;
        tm = cmax;
    }
    if (tm == NEG_INF) return;                         // nothing but zero weights so far
#pragma unroll
    for (int q = 0; q < kK1Quads; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const real e = fast_exp(lw[q][j] - tm);    // exp(-inf) = 0 for padding / zero weights
            acc[0] += e;
            acc[1] = fma(e, e, acc[1]);
            if (MOM) {
                real xc[NX];   // padding lanes hold finite values (zero-filled) and e = 0
#pragma unroll
                for (int d = 0; d < NX; ++d) xc[d] = xs[q][d][j] - shift[d];
                int k = 2;
#pragma unroll
                for (int d = 0; d < NX; ++d) { acc[k] = fma(e, xc[d], acc[k]); ++k; }
#pragma unroll
                for (int d = 0; d < NX; ++d)
#pragma unroll
                    for (int d2 = 0; d2 <= d; ++d2) { acc[k] = fma(e * xc[d], xc[d2], acc[k]); ++k; }
            }
        }
}

// Resident CTAs per SM of K1 (= persistent grid per SM)
template <typename real, class Model, bool MOM> struct K1Occ {
    static constexpr int value = (sizeof(real) == 4 && Model::NX <= 2) ? kK1Occ : 2;
};

template <typename real, class Model, bool MOM>
__global__ void __launch_bounds__(kThreads, K1Occ<real, Model, MOM>::value)
k1_propagate_weight(K1Args<real> a, Model m) {
    constexpr int NX = Model::NX, NZ = Model::NZ;
    constexpr int NM = NX + NX * (NX + 1) / 2;   // moment accumulators
    constexpr int NR = 2 + NM;                   // sum e, sum e^2, moments
    // software pipelining: the next chunk's particles are loaded before the current chunk is
    // computed (16 registers per dimension in fp32) -- only where the register budget allows
    constexpr bool PREFETCH = SMC_K1_PREFETCH && sizeof(real) == 4 && NX <= 1;
    __shared__ real s_max[kWarps];
    __shared__ double s_red[kWarps][NR];

    pdl_wait();
    pdl_trigger();
    const Ctrl* ctrl = a.ctrl;
    real* x = ctrl->cur ? a.x1 : a.x0;
    const bool uniform = ctrl->uniform != 0;
    const real log_norm_prev = (real)ctrl->log_norm;
    const real lw_uniform = (real)(-log((double)a.n_global));
    const bool do_prop = a.mode & SMC_MODE_PROPAGATE, do_w = a.mode & SMC_MODE_WEIGHT;

    real z[NZ], shift[NX];
#pragma unroll
    for (int i = 0; i < NZ; ++i) z[i] = do_w ? (real)a.z[i] : real(0);
#pragma unroll
    for (int d = 0; d < NX; ++d) shift[d] = (real)ctrl->shift[d];

    const int tid = threadIdx.x;
    const real NEG_INF = -INFINITY;
    // thread-local online accumulators: sums are kept relative to the running thread max `tm`
    real tm = NEG_INF;
    real acc[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[k] = 0;

    const int64_t n_chunks = cdiv64(a.n, kK1Block);
    const int64_t n_full = a.n / kK1Block;       // chunks [0, n_full) need no bounds checks
    // Range totals (fp32 bootstrap steps with the fused finalize): CTA b then owns the contiguous chunks of the
    // pre-pass ranges [b R / G, (b + 1) R / G) -- a chunk is a resampling tile -- and flushes one exact 64-bit sum
    // per range into the buffer the control block does NOT name (all zero by invariant); the last CTA flips it.
    const bool do_rsum = sizeof(real) == 4 && a.rsum != nullptr && uniform && do_w && (a.fuse_finalize || a.shard.mail != nullptr);
    const int rsel = ctrl->rsum_sel;
    unsigned long long racc = 0ull;
    int64_t c_first = blockIdx.x, c_end = n_chunks, c_stride = gridDim.x;
    int r_cur = 0;
    int64_t r_bnd = 0;                           // first chunk of the next range
    if (do_rsum) {
        const int rpc = a.n_ranges / (int)gridDim.x;
        r_cur = (int)blockIdx.x * rpc;
        c_first = k3_range_begin(r_cur, n_chunks, a.n_ranges);
        c_end = k3_range_begin(r_cur + rpc, n_chunks, a.n_ranges);
        c_stride = 1;
        r_bnd = k3_range_begin(r_cur + 1, n_chunks, a.n_ranges);
    }
    auto flush_range = [&]() {                   // CTA-uniform: one 64-bit atomic per warp
        const unsigned long long v = warp_sum(racc);
        if ((tid & 31) == 0 && v) atomicAdd(a.rsum + (size_t)(rsel ^ 1) * kRsumStride + r_cur, v);
        racc = 0ull;
    };
    if (PREFETCH) {
        real cur[kK1Quads][NX][4], nxt[kK1Quads][NX][4];
        int64_t chunk = c_first;
        if (chunk < c_end) {
            if (chunk < n_full) k1_load<real, NX, false>(a, x, chunk * kK1Block, tid, cur);
            else k1_load<real, NX, true>(a, x, chunk * kK1Block, tid, cur);
        }
        for (; chunk < c_end; chunk += c_stride) {
            const int64_t next = chunk + c_stride;
            if (next < c_end) {
                if (next < n_full) k1_load<real, NX, false>(a, x, next * kK1Block, tid, nxt);
                else k1_load<real, NX, true>(a, x, next * kK1Block, tid, nxt);
            }
            if (do_rsum && chunk >= r_bnd) {
                flush_range();
                ++r_cur;
                r_bnd = k3_range_begin(r_cur + 1, n_chunks, a.n_ranges);
            }
            if (chunk < n_full)
                k1_chunk<real, Model, false, MOM, NR>(a, m, x, chunk * kK1Block, tid, uniform, log_norm_prev, lw_uniform,
                                                      do_prop, do_w, z, shift, tm, acc, cur, do_rsum, &racc);
            else
                k1_chunk<real, Model, true, MOM, NR>(a, m, x, chunk * kK1Block, tid, uniform, log_norm_prev, lw_uniform,
                                                     do_prop, do_w, z, shift, tm, acc, cur, do_rsum, &racc);
#pragma unroll
            for (int q = 0; q < kK1Quads; ++q)
#pragma unroll
                for (int d = 0; d < NX; ++d)
#pragma unroll
                    for (int j = 0; j < 4; ++j) cur[q][d][j] = nxt[q][d][j];
        }
    } else {
        for (int64_t chunk = c_first; chunk < c_end; chunk += c_stride) {
            const int64_t blk_base = chunk * kK1Block;
            real cur[kK1Quads][NX][4];
            if (do_rsum && chunk >= r_bnd) {
                flush_range();
                ++r_cur;
                r_bnd = k3_range_begin(r_cur + 1, n_chunks, a.n_ranges);
            }
            if (chunk < n_full) {
                k1_load<real, NX, false>(a, x, blk_base, tid, cur);
                k1_chunk<real, Model, false, MOM, NR>(a, m, x, blk_base, tid, uniform, log_norm_prev, lw_uniform, do_prop,
                                                      do_w, z, shift, tm, acc, cur, do_rsum, &racc);
            } else {
                k1_load<real, NX, true>(a, x, blk_base, tid, cur);
                k1_chunk<real, Model, true, MOM, NR>(a, m, x, blk_base, tid, uniform, log_norm_prev, lw_uniform, do_prop,
                                                     do_w, z, shift, tm, acc, cur, do_rsum, &racc);
            }
        }
    }
    if (do_rsum) flush_range();
    if (!do_w) return;

    // ---- block combine: common max, rescale, fixed-tree sums -----------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    real wmax = warp_max(tm);
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    real bmax = s_max[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) bmax = s_max[w] > bmax ? s_max[w] : bmax;
    const real sc = (tm == NEG_INF) ? real(0) : fast_exp(tm - bmax);
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        real v = acc[k] * (k == 1 ? sc * sc : sc);
        v = warp_sum(v);
        if (lane == 0) s_red[warp][k] = (double)v;
    }
    __syncthreads();
    if (tid < NR + 1) {
        double* out = a.parts + (int64_t)blockIdx.x * partial_stride(NX);
        if (tid == 0) {
            out[0] = (double)bmax;
        } else {
            double s = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += s_red[w][tid - 1];
            out[tid] = s;
        }
    }
    if (a.fuse_finalize) {
        // the last CTA to arrive reduces all block partials (fixed order) and finalises the step
        __shared__ int s_last;
        __shared__ double s_out[32];
        __shared__ double s_tmp[kWarps];
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            K2Args f = a.fin;
            f.producer_rsum = do_rsum ? 1 : 0;
            if (do_rsum) {   // hand the range totals over: clear the old buffer, make the new one current
                for (int i = tid; i < kRsumStride; i += kThreads) a.rsum[(size_t)rsel * kRsumStride + i] = 0ull;
                if (tid == 0) const_cast<Ctrl*>(ctrl)->rsum_sel = rsel ^ 1;
                __syncthreads();
            }
            finalize_body(f, s_out, s_tmp);
            if (tid == 0) *a.ticket = 0u;
        }
    } else if (a.shard.mail) {
        // sharded: the last CTA folds the block partials into ONE record and stores it into every
        // rank's mailbox (NVLink peer stores), then raises this rank's flag there
        __shared__ int s_last;
        __shared__ double s_out[32];
        __shared__ double s_tmp[kWarps];
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            const int ps = partial_stride(NX);
            combine_partials(a.parts, (int)gridDim.x, ps, s_out, s_tmp, 0);
            if (tid < ps)
                for (int g = 0; g < a.shard.n_shards; ++g) a.shard.mail[g]->rec[a.shard.rank][tid] = s_out[tid];
            if (do_rsum) {
                // bootstrap step: the shard's fixed-point weight total travels with the partial record, so the
                // resampler needs no second exchange (and no range pre-pass); hand the range totals over
                __shared__ unsigned long long s_rt[kWarps];
                unsigned long long t = 0;
                for (int i = tid; i < a.n_ranges; i += kThreads) t += __ldcg(a.rsum + (size_t)(rsel ^ 1) * kRsumStride + i);
                t = warp_sum(t);
                __syncthreads();
                if ((tid & 31) == 0) s_rt[tid >> 5] = t;
                __syncthreads();
                if (tid == 0) {
                    unsigned long long tot = 0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) tot += s_rt[w];
                    for (int g = 0; g < a.shard.n_shards; ++g) a.shard.mail[g]->tot[a.shard.rank] = tot;
                    const_cast<Ctrl*>(ctrl)->rsum_sel = rsel ^ 1;
                }
                for (int i = tid; i < kRsumStride; i += kThreads) a.rsum[(size_t)rsel * kRsumStride + i] = 0ull;
            }
            __threadfence_system();
            __syncthreads();
            if (tid < a.shard.n_shards)
                *(volatile unsigned long long*)&a.shard.mail[tid]->flag_rec[a.shard.rank] = a.shard.stamp;
            if (tid == 0) *a.ticket = 0u;
        }
    }
}

// ==============================================================================================
// K2: finalize.  One CTA reduces the partial records in a fixed order.
// ==============================================================================================
__device__ inline void ctrl_prepare_search(Ctrl* c, int64_t n_global, bool f32);   // defined with the resampler


// Combine partial records (max-shifted sums) into one record `out` of the same layout
// (out[0] = global max, sums relative to it).  All threads of the CTA return the record in smem.
__device__ inline void combine_partials(const double* parts, int n_parts, int ps, double* s_out,
                                        double* s_tmp, int stride) {
    if (stride <= 0) stride = ps;
    // One pass over the records (thread-local online rescaling), one CTA-wide max, one CTA-wide sum
    // of all fields: 4 barriers and a single round of global loads on the critical path of a step.
    __shared__ double s_x[kWarps][20];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double m = -INFINITY;
    double acc[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) acc[k] = 0;
    for (int p = tid; p < n_parts; p += kThreads) {     // fixed record -> thread assignment
        const double* r = parts + (int64_t)p * stride;
        const double r0 = __ldcg(r);
        double v[20];
#pragma unroll
        for (int k = 1; k < 20; ++k) v[k] = (k < ps) ? __ldcg(r + k) : 0.0;
        if (r0 == -INFINITY) continue;                  // an all-zero-weight block contributes nothing
        if (r0 > m) {                                   // rescale what this thread has so far
            const double sc = exp(m - r0);              // m = -inf -> 0
#pragma unroll
            for (int k = 1; k < 20; ++k) if (k < ps) acc[k] *= (k == 2 ? sc * sc : sc);
            m = r0;
        }
        const double sc = exp(r0 - m);
#pragma unroll
        for (int k = 1; k < 20; ++k)
            if (k < ps) acc[k] = fma(v[k], k == 2 ? sc * sc : sc, acc[k]);   // sum e^2 scales with the square
    }
    const double M = block_max_256(m, s_tmp);
    const double sc = (m == -INFINITY) ? 0.0 : exp(m - M);
#pragma unroll
    for (int k = 1; k < 20; ++k) {
        if (k < ps) {
            const double t = warp_sum(acc[k] * (k == 2 ? sc * sc : sc));
            if (lane == 0) s_x[warp][k] = t;
        }
    }
    __syncthreads();
    if (tid < ps && tid >= 1) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_x[w][tid];
        s_out[tid] = t;
    }
    if (tid == 0) s_out[0] = M;
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) k2_reduce_partials(const double* parts, int n_parts, int ps,
                                                               double* out) {
    __shared__ double s_out[32];
    __shared__ double s_tmp[kWarps];
    combine_partials(parts, n_parts, ps, s_out, s_tmp, 0);
    if (threadIdx.x < ps) out[threadIdx.x] = s_out[threadIdx.x];
}

// K2 body: combine the partial records, publish normaliser / ESS / decision / u (+ weighted moments).
// Runs in one 256-thread CTA: the stand-alone kernel below, or the last CTA of K1 to finish.
__device__ inline void finalize_tail(const K2Args& a, const double* s_out);
__device__ inline void finalize_body(const K2Args& a, double* s_out, double* s_tmp) {
    const int ps = a.inline_moments ? partial_stride(a.nx) : 3;
    combine_partials(a.parts, a.n_parts, ps, s_out, s_tmp, a.rec_stride);
    if (threadIdx.x != 0) return;
    finalize_tail(a, s_out);
}
// One-warp form for a handful of records (<= 32, one per lane): the sharded resampler runs it in every CTA's
// prologue, where a CTA-wide fp64 reduction in 4 x 148 CTAs at once would cost more than the launch it replaces.
__device__ inline void finalize_warp(const K2Args& a, double* s_out) {
    const int ps = a.inline_moments ? partial_stride(a.nx) : 3;
    const int stride = a.rec_stride > 0 ? a.rec_stride : ps;
    const int lane = threadIdx.x & 31;
    const bool have = lane < a.n_parts;
    const double r0 = have ? __ldcg(a.parts + (int64_t)lane * stride) : -INFINITY;
    const double M = warp_max(r0);
    const double sc = (have && r0 != -INFINITY) ? exp(r0 - M) : 0.0;
    for (int k = 1; k < ps; ++k) {
        const double v = have ? __ldcg(a.parts + (int64_t)lane * stride + k) : 0.0;
        const double t = warp_sum(v * (k == 2 ? sc * sc : sc));
        if (lane == 0) s_out[k] = t;
    }
    if (lane == 0) { s_out[0] = M; finalize_tail(a, s_out); }
    __syncwarp();
}
__device__ inline void finalize_tail(const K2Args& a, const double* s_out) {
    Ctrl* c = a.ctrl;
    const double M = s_out[0], S = s_out[1], Q = s_out[2];
    const double log_norm = M + log(S);
    const double ess = (S * S) / Q;
    const int do_res = ess < a.thresh * (double)a.n_global;   // strict <, particle_filter.py:204
    double u = a.u_host;
    if (u < 0.0) {
        uint32_t r[4] = {0u, 0u, a.t, kStreamResample};
        philox4x32_10(r, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        u = uniform53(r[0], r[1]);
    }
    c->log_norm = log_norm;
    c->ess = ess;
    c->u = u;
    c->do_resample = do_res;
    c->uniform = 0;
    ctrl_prepare_search(c, a.n_global, a.f32 != 0);
    if (a.f32) {
        // fp32 CDF: fixed-point weights a_i = rni(2^Fa exp(lw_i - lw_ref)).  Any reference lw_ref >= max lw_i
        // that leaves the weights that matter enough resolution will do, because the stratified positions are laid
        // over the ACTUAL total (see RhoMap).  A bootstrap step knows such a reference before it has seen a
        // single weight -- lw_i = -log N - quad_i / 2 <= -log N -- which is what lets the kernel that WRITES
        // the log-weights also deliver the range totals of the resampler's CDF (no pre-pass over them).
        const int Fa = fa_bits_for(a.n_global);
        const double log_mean = log_norm - a.lw_bound - log((double)a.n_global);   // log of the mean of exp(lw - bound)
        const bool bound_ok = M <= a.lw_bound && log_mean >= -(double)(Fa - 20) * 0.6931471805599453;
        if (bound_ok) {
            c->lw_ref = a.lw_bound;
            c->kshift = a.kshift_bound;
            c->rsum_valid = a.producer_rsum ? 1 : 0;
        } else {
            // otherwise the largest log-weight of the step: exact whatever the order of the reductions (unlike the
            // normaliser, whose fp32 block sums differ in the last bits between launch shapes and shard counts), so
            // every pipeline quantises alike; the heaviest particle gets 2^Fa, the total stays below 2^Fa N <= 2^62
            c->lw_ref = M;
            c->kshift = M == -INFINITY ? 0.f : (float)((double)Fa - M * (double)kLog2eF);
            c->rsum_valid = 0;
        }
    }
    if (a.stats) {
        a.stats[0] = log_norm;
        a.stats[1] = ess;
        a.stats[2] = (double)do_res;
        a.stats[3] = u;
        if (a.inline_moments) {
            const int nx = a.nx;
            double* mean = a.stats + SMC_STATS_HEADER;
            double* cov = mean + nx;
            double m1[kFusedNxMax];
            for (int d = 0; d < nx; ++d) {
                m1[d] = s_out[3 + d] / S;
                mean[d] = c->shift[d] + m1[d];
            }
            int k = 3 + nx;
            for (int d = 0; d < nx; ++d)
                for (int d2 = 0; d2 <= d; ++d2) {
                    double v = s_out[k++] / S - m1[d] * m1[d2];
                    cov[d * nx + d2] = v;
                    cov[d2 * nx + d] = v;
                }
        }
    }
}

__global__ void __launch_bounds__(kThreads) k2_finalize(K2Args a) {
    __shared__ double s_out[32];
    __shared__ double s_tmp[kWarps];
    finalize_body(a, s_out, s_tmp);
}

// sharded: wait until every rank's partial record of this step sits in the own mailbox, then
// reduce the G records in rank order (identical arithmetic, hence identical results, on all ranks)
__global__ void __launch_bounds__(kThreads) k2_finalize_sharded(K2Args a, ShardCtx sc) {
    __shared__ double s_out[32];
    __shared__ double s_tmp[kWarps];
    Mailbox* own = sc.mail[sc.rank];
    mailbox_wait(own->flag_rec, sc.n_shards, sc.stamp, &own->error, 1);
    a.parts = &own->rec[0][0];
    a.n_parts = sc.n_shards;
    a.rec_stride = 32;
    // producer_rsum (host: K1 was launched with range totals) holds iff the step was a bootstrap step, which every
    // rank decides alike from its replicated control block
    a.producer_rsum = a.producer_rsum && a.ctrl->uniform != 0;
    finalize_body(a, s_out, s_tmp);
}

// ==============================================================================================
// K3: single-pass systematic resampler
// ==============================================================================================
// Exact predicate "pos_i < c" with pos_i = fl(fl(u + i) / N) (particle_filter.py:160) and the
// CDF value c = C / (N 2^F) given in fixed point.  Integer arithmetic only after the fp64 position.
__device__ inline bool pos_less(int64_t i, unsigned long long C, double u, int64_t N, int F) {
    const double p = (u + (double)i) / (double)N;       // IEEE fp64, same two roundings as NumPy
    if (!(p > 0.0)) return C > 0ull;                     // p == 0  (u = 0, i = 0)
    int ex;
    const double fr = frexp(p, &ex);                     // p = fr * 2^ex, fr in [0.5, 1)
    const unsigned long long mant = (unsigned long long)ldexp(fr, 53);   // 53-bit integer
    const int e2 = ex - 53;                              // p = mant * 2^e2
    // p * N * 2^F < C  <=>  floor(mant * N / 2^s) < C  with s = -(e2 + F)
    const int s = -(e2 + F);
    unsigned __int128 prod = (unsigned __int128)mant * (unsigned __int128)(unsigned long long)N;
    if (s >= 128) return C > 0ull;
    if (s > 0) prod >>= s; else prod <<= (-s);
    return prod < (unsigned __int128)C;
}

struct SearchParams {
    double u;
    unsigned long long uq;        // floor(u * 2^F)
    unsigned long long margin;    // |frac - uq| <= margin -> exact path
    unsigned long long fmask;     // 2^F - 1
    int64_t n_global;
    int F;
};

// Fractional bits of the fixed-point CDF ("index units" N*w): as many as fit below the status bit.
// fp32 populations use F = 32 throughout: the output count below a CDF value is then simply the
// high word of the biased value, and -- fp32 weights being what they are -- positions within fp64
// round-off of a CDF value need no special treatment (the exact predicate serves the fp64 parity mode).
__host__ __device__ inline int frac_bits_for(int64_t n_global, bool f32 = false) {
    if (f32) return 32;
    int lg = 0;
    while ((1ll << lg) < n_global) ++lg;
    int F = 62 - lg;
    return F < 32 ? 32 : (F > 60 ? 60 : F);
}

__device__ inline SearchParams make_search(double u, int64_t n_global, bool f32) {
    SearchParams sp;
    sp.F = frac_bits_for(n_global, f32);
    sp.u = u;
    sp.uq = __double2ull_rd(ldexp(u, sp.F));
    sp.fmask = (1ull << sp.F) - 1ull;
    // fp64 round-off of (u + i) / N is <= N 2^-52 index units = N 2^(F-52) fraction units
    const int sh = sp.F - 50;
    sp.margin = 8ull + (sh >= 0 ? ((unsigned long long)n_global << sh) : ((unsigned long long)n_global >> (-sh)));
    sp.n_global = n_global;
    return sp;
}

// Once per step (finalize / set_uniform): cache the resampler constants in the control block so
// that the 16 K resampling CTAs do not each redo the fp64 log / ldexp work.
__device__ inline void ctrl_prepare_search(Ctrl* c, int64_t n_global, bool f32) {
    const SearchParams sp = make_search(c->u, n_global, f32);
    c->frac_bits = sp.F;
    c->uq = sp.uq;
    c->margin = sp.margin;
    c->lw_ref = c->log_norm - log((double)n_global);
}
__device__ inline SearchParams load_search(const Ctrl* c, int64_t n_global) {
    SearchParams sp;
    sp.F = c->frac_bits;
    sp.u = c->u;
    sp.uq = c->uq;
    sp.fmask = (1ull << sp.F) - 1ull;
    sp.margin = c->margin;
    sp.n_global = n_global;
    return sp;
}

// e(C) = #{ i in [0, N) : pos_i < C }  (first output index NOT owned by CDF value C)
__device__ inline int64_t outputs_below(unsigned long long C, const SearchParams& sp) {
    const int64_t I = (int64_t)(C >> sp.F);
    const unsigned long long f = C & sp.fmask;
    int64_t e = I + (f > sp.uq ? 1 : 0);
    const unsigned long long diff = f > sp.uq ? f - sp.uq : sp.uq - f;
    if (e > sp.n_global) e = sp.n_global;
    if (diff <= sp.margin || diff >= sp.fmask - sp.margin) {
        while (e > 0 && !pos_less(e - 1, C, sp.u, sp.n_global, sp.F)) --e;
        while (e < sp.n_global && pos_less(e, C, sp.u, sp.n_global, sp.F)) ++e;
    }
    return e;
}

struct FinishArgs {
    double* stats;              // stats row (nullable)
    const double* mom_parts;    // n_mom_parts records of moment_stride(nx) doubles
    int32_t n_mom_parts, nx, inline_moments, no_moments;
    int32_t mom_stride;         // doubles between records (0 = packed)
    int64_t n_global;
};

// Fused "resample step t-1, then propagate + weigh step t" (run() pipeline, bootstrap filters): the
// copy-out of the resampler hands every gathered particle straight to the model functor, so the
// resampled set is never written to and re-read from HBM (20 B instead of 28 B per particle-update).
struct NoFuse { static constexpr int NX = 0, NZ = 1; };
template <typename real> struct FuseK1 {
    real* lw_out;            // log-weights of step t (a second array: other CTAs still read step t-1's)
    const double* z;         // observation of step t (device)
    PhiloxKeys keys;
    uint32_t t;
    double* parts;           // per-CTA (max, sum e, sum e^2) records of step t
    K2Args fin;              // finalize of step t, run by the last CTA after the finish of step t-1
};

template <typename real> struct K3Args {
    Ctrl* ctrl;
    const real* wsrc;           // log-weights (or linear normalised weights)
    const real* x0;
    const real* x1;
    real* dst0;                 // destination when ctrl->cur == 1  (i.e. buffer 0)
    real* dst1;                 // destination when ctrl->cur == 0  (i.e. buffer 1)
    real* const* peer0;         // nullable: per-shard destination pointers (cur == 1 / cur == 0)
    real* const* peer1;
    int32_t* ancestors;         // nullable
    const unsigned long long* rtot;   // per-range weight totals from the pre-pass (one per resampler CTA)
    double* mom_parts;          // nullable
    unsigned long long cdf_offset;
    const unsigned long long* totals;   // nullable: per-shard weight totals (offset = sum over lower ranks)
    int32_t rank, n_totals;
    int64_t n, n_global, offset, ld, shard_n;
    int32_t force;              // ignore ctrl->do_resample (stand-alone use)
    int64_t n_tiles;
    int32_t fuse_finish;        // the last CTA to finish also runs finish_body(fin)
    unsigned int* ticket;       // arrival counter for fuse_finish (zero between launches)
    FinishArgs fin;
    ShardCtx shard;             // mail != nullptr: weight totals come from / moment record goes to the mailboxes
    FuseK1<real> fk;            // used by the fused instantiations only
    int32_t ranges_per_cta;     // pre-pass ranges owned by one CTA of this launch
    // sharded bootstrap steps: the finalize runs in EVERY CTA's prologue (on a shared-memory copy of the control
    // block, once all ranks' partial records have arrived) instead of in a launch of its own
    int32_t prologue_finalize;
    K2Args pfin;
    unsigned int* gbar;         // grid-barrier counter of the in-kernel range pre-pass (fallback of such steps)
};

// Tile geometry: THREADS threads own IPT consecutive items each (fp32: 256 x 16 = 4096-particle
// tiles, fp64: 256 x 8); every shared/global access moves 16-byte units of EPU items.
// Measured on B200 at N = 2^26 (resampler + pre-pass, ms): 128 x 16 -> 0.347, 256 x 16 -> 0.336,
// 512 x 16 -> 0.368, 512 x 8 (3 CTAs/SM, spilling) -> 0.423.
template <typename real> struct K3Geom;
template <> struct K3Geom<float> {
    static constexpr int THREADS = 256, IPT = 16, EPU = 4;
    using unit = float4;
};
template <> struct K3Geom<double> {
    static constexpr int THREADS = 256, IPT = 8, EPU = 2;
    using unit = double2;
};

__device__ inline void unit_to_array(const float4& u, float* v) { v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; }
__device__ inline void unit_to_array(const double2& u, double* v) { v[0] = u.x; v[1] = u.y; }


// Per-item quantisation of the weights.
//  fp32 mode: every weight is converted on its own, a = rni(p), p = 2^(lw log2e + kshift) (one FFMA, FMNMX,
//  MUFU.EX2 and F2I); the same three instructions run wherever a log-weight is quantised -- the kernel that
//  writes it, the range pre-pass, the resampler -- so all of them see the same integers and every sum of
//  them is exact and order-independent;
//  fp64 mode (parity): the thread's fp64 running sum is quantised instead, so the round-off of
//  2^-F index units does not accumulate over a run of tiny weights.
template <typename real> struct Quant;
template <> struct Quant<float> {
    float fscale, kshift;
    __device__ inline void init(const Ctrl* c, int64_t n_global, bool) {
        fscale = __uint_as_float((127u + (uint32_t)fa_bits_for(n_global)) << 23);   // 2^Fa (linear weights)
        kshift = c->kshift;
    }
    __device__ inline float p_logw(float lw) const { return quant_p(lw, kshift); }
    __device__ inline float p_lin(float v) const { return v * fscale; }           // v = N w, clamped to [0, N]
    __device__ inline unsigned long long add(unsigned long long run, float v, double&) const {
        return run + quant_a(p_lin(v));
    }
    __device__ inline unsigned long long add_logw(unsigned long long run, float lw, double&) const {
        return run + quant_a(p_logw(lw));
    }
};
template <> struct Quant<double> {
    double fscale, lw_ref, nmax;
    __device__ inline void init(const Ctrl* c, int64_t, bool linear_w) {
        const int F = c->frac_bits;
        fscale = ldexp(1.0, F); lw_ref = linear_w ? 0.0 : c->lw_ref; nmax = ldexp(1.0, 62 - F);
    }
    __device__ inline unsigned long long add(unsigned long long, double v, double& runf) const {
        runf += v;
        return __double2ull_rn(runf * fscale);
    }
    __device__ inline unsigned long long add_logw(unsigned long long run, double lw, double& runf) const {
        double v = exp(lw - lw_ref);
        v = fmin(fmax(v, 0.0), nmax);                                        // NaN / negative -> 0
        return add(run, v, runf);
    }
};

// fp32 CDF: the stratified positions are laid over the ACTUAL fixed-point total A of the weights instead of
// over a nominal one: the output boundary of a CDF value C (in the weights' own units) is taken at
// C rho, rho = N 2^32 / A, in 2^-32 index units.  Whatever reference the weights were quantised against,
// nothing is then left over for, or missing from, the last particles, and neither the round-off of a
// normaliser nor the order of any floating-point sum can move a boundary: the three-launch step, the
// fused pipeline and any shard count select the same ancestors.  rho = R 2^-sh with R normalised to 64 bits.
struct RhoMap {
    unsigned long long R, bias;    // bias = 2^32 - 1 - floor(u 2^32): e(C) = (floor(C rho) + bias) >> 32
    int sh;                        // C rho = mulhi64(C << sh, R): the total, shifted, fills 63 bits
    float rho_f;                   // rho 2^-32: thread-local weight sums (fp32) -> index units
};
__device__ inline RhoMap make_rho(unsigned long long total, int64_t n_global, unsigned long long bias) {
    RhoMap m;
    m.bias = bias; m.R = 0ull; m.sh = 0; m.rho_f = 0.f;
    if (total > 0ull && total < (1ull << 63)) {
        m.sh = __clzll((long long)total) - 1;                                // total << sh in [2^62, 2^63)
        const double r = (double)n_global / (double)(total << m.sh);         // correctly rounded: same on every CTA
        m.R = __double2ull_rz(fmin(ldexp(r, 96), 18446744073709549568.0));   // rho 2^(64 - sh) <= N 2^34 <= 2^64
        m.rho_f = (float)((double)n_global / (double)total);
    }
    return m;
}
// floor(C rho) + bias, monotone in C (C <= total, so the result stays below N 2^32 + 2^32)
__device__ __forceinline__ unsigned long long map_pos(unsigned long long C, const RhoMap& m) {
    return __umul64hi(C << m.sh, m.R) + m.bias;
}
// Thread-local boundaries of IPT consecutive items from their fp32 weights p[] (the values the fixed-point
// weights were rounded from) between the thread's exact first and last boundary: e[k] = e_start +
// trunc(frac(start) + rho sum_{j<=k} p_j), clamped to e_end; the item that carries the thread's last weight
// takes e_end (so that an item of zero weight never receives an output).  Used where no thread has more
// than a few dozen outputs (error < 2^-17 index units); skewed tiles take the exact form below.
template <int IPT>
__device__ __forceinline__ void f32_boundaries(const float (&p)[IPT], unsigned long long Ps, uint32_t e_end, float rho_f,
                                               uint32_t (&e)[IPT]) {
    const uint32_t e_start = (uint32_t)(Ps >> 32);
    const float fr = __uint2float_rz((uint32_t)Ps) * 2.3283064365386963e-10f;    // < 1
    float ftot = 0.f;
#pragma unroll
    for (int k = 0; k < IPT; ++k) ftot = __fadd_rn(ftot, p[k]);
    float fk = 0.f;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        fk = __fadd_rn(fk, p[k]);
        const uint32_t ek = umin(e_start + __float2uint_rz(__fmaf_rn(fk, rho_f, fr)), e_end);
        e[k] = fk < ftot ? ek : e_end;
    }
}
template <int IPT>
__device__ __forceinline__ void exact_boundaries(const float (&p)[IPT], unsigned long long base_raw, const RhoMap& rm,
                                                 uint32_t (&e)[IPT]) {
    unsigned long long c = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        unsigned long long q;   // (volatile: recomputed here rather than kept live across the hot path)
        asm volatile("cvt.rni.u64.f32 %0, %1;" : "=l"(q) : "f"(p[k]));
        c += q;
        e[k] = (uint32_t)(map_pos(base_raw + c, rm) >> 32);
    }
}

// Fast boundary search in 32-bit arithmetic.  With C' = C + (2^F - 1 - uq):
//   e(C) = C' >> F           (number of stratified positions below the CDF value C)
// and the position is within fp64 round-off of the CDF value iff the top 32 fraction bits of C'
// lie within m32 of 0 (mod 2^32); only then is the exact predicate evaluated.
struct FastSearch {
    unsigned long long bias;   // 2^F - 1 - uq
    uint32_t sh;               // F - 32
    uint32_t m32;              // margin in units of 2^(F-32), rounded up, + 1
    uint32_t n_global;
};
// rare path, kept out of line so that it costs the hot loop neither registers nor I-cache
__device__ __noinline__ uint32_t exact_outputs_below(unsigned long long C, uint32_t e0, double u, int64_t n_global,
                                                     int F) {
    long long ee = e0;
    while (ee > 0 && !pos_less(ee - 1, C, u, n_global, F)) --ee;
    while (ee < n_global && pos_less(ee, C, u, n_global, F)) ++ee;
    return (uint32_t)ee;
}
__device__ inline uint32_t fast_outputs_below(unsigned long long Cb /* = C + bias */, const FastSearch& fs,
                                              const SearchParams& sp) {
    const uint32_t hi = (uint32_t)(Cb >> 32), lo = (uint32_t)Cb;
    uint32_t e = hi >> fs.sh;
    const uint32_t frac_hi = __funnelshift_r(lo, hi, fs.sh);    // bits [F-32, F) of Cb
    if (frac_hi + fs.m32 <= 2u * fs.m32)                         // near 0 or near 2^32 (wraps)
        e = exact_outputs_below(Cb - fs.bias, e, sp.u, sp.n_global, sp.F);
    return e;
}

// ---- K3a: fixed-point weight totals of the resampler's tile ranges --------------------------------
// The resampler is a persistent kernel: CTA b owns the contiguous tiles [b T / G, (b + 1) T / G).  It
// needs the exclusive CDF prefix of its first tile only; the prefixes inside its range follow from
// its own tile totals.  So the pre-pass is a pure streaming reduction over the same ranges (one extra
// read of the log-weights), with no inter-CTA dependency, no per-tile records and no scan kernel:
// CTA b of the resampler sums the totals of the ranges below it (<= G values) in its prologue.
// The quantisation is the one k3_resample applies item by item, so both kernels agree to the bit.
// The pre-pass cuts the tiles into R ranges with R a multiple of every resampler grid (148 x 12 ranges
// for grids of 148 x {4, 3, 2} CTAs): CTA b of a grid of G owns ranges [b R / G, (b + 1) R / G), whose
// first tile is floor(b T / G) -- the plain and the fused resampler can consume the same pre-pass.

// (1776 ranges = exactly three waves of 4 resident CTAs on 148 SMs; measured at N = 2^26: 2 -> 0.139 ms,
// 3 -> 0.091, 4 -> 0.061, 5 -> 0.067, 6 -> 0.069)
#ifndef SMC_K3A_OCC
#define SMC_K3A_OCC 4
#endif
template <typename real, bool LINEAR_W>
__global__ void __launch_bounds__(K3Geom<real>::THREADS, sizeof(real) == 4 ? SMC_K3A_OCC : 3) k3a_range_totals(const Ctrl* ctrl, const real* wsrc, int64_t n,
                                                                         int64_t n_global, int64_t n_tiles,
                                                                         unsigned long long* rtot, int force,
                                                                         ShardCtx shard, unsigned int* ticket) {
    using G = K3Geom<real>;
    using unit = typename G::unit;
    constexpr int kK3Threads = G::THREADS, kK3Warps = kK3Threads / 32;
    constexpr int IPT = G::IPT, EPU = G::EPU, UPT = IPT / EPU, TILE = kK3Threads * IPT;
    __shared__ unsigned long long s_tot[kK3Warps];
    pdl_wait();
    pdl_trigger();
    if (!force && !ctrl->do_resample) return;
    // fp32 bootstrap steps: the kernel that wrote the log-weights has already delivered these totals
    if (!force && sizeof(real) == 4 && !LINEAR_W && ctrl->rsum_valid) return;
    rtot += (size_t)ctrl->rsum_sel * kRsumStride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Quant<real> quant;
    quant.init(ctrl, n_global, LINEAR_W);
    const real n_real = (real)n_global;
    const real fill = LINEAR_W ? real(0) : real(-INFINITY);
    const int64_t t0 = k3_range_begin(blockIdx.x, n_tiles, gridDim.x), t1 = k3_range_begin(blockIdx.x + 1, n_tiles, gridDim.x);
    unsigned long long total = 0;
    // one tile ahead: the next tile's loads are in flight while the current one is quantised
    auto load_tile = [&](int64_t tile, real (&tmp)[UPT][EPU]) {
        const int64_t tile_base = tile * TILE;
        const int n_in_tile = tile_base + TILE <= n ? TILE : (int)(n - tile_base);
        const unit* wtile = reinterpret_cast<const unit*>(wsrc + tile_base);
        // fp32: weights are quantised one by one, so any (coalesced) assignment gives the same total;
        // fp64: the thread's running sum is quantised, so the thread must own the same IPT consecutive
        //       items as in k3_resample (blocked assignment)
#pragma unroll
        for (int r = 0; r < UPT; ++r) {
            const int f = (sizeof(real) == 4) ? (r * kK3Threads + tid) : (UPT * tid + r);
#pragma unroll
            for (int j = 0; j < EPU; ++j) tmp[r][j] = fill;
            if (f * EPU < n_in_tile) {
                unit_to_array(wtile[f], tmp[r]);
                if (f * EPU + EPU > n_in_tile) {
#pragma unroll
                    for (int j = 0; j < EPU; ++j) if (f * EPU + j >= n_in_tile) tmp[r][j] = fill;
                }
            }
        }
    };
    real cur[UPT][EPU], nxt[UPT][EPU];
    if (t0 < t1) load_tile(t0, cur);
    for (int64_t tile = t0; tile < t1; ++tile) {
        if (tile + 1 < t1) load_tile(tile + 1, nxt);
        unsigned long long run = 0;
        double runf = 0.0;
#pragma unroll
        for (int r = 0; r < UPT; ++r)
#pragma unroll
            for (int j = 0; j < EPU; ++j) {
                if (LINEAR_W) {
                    real v = fmin(fmax(cur[r][j] * n_real, real(0)), n_real);
                    run = quant.add(run, v, runf);
                } else {
                    run = quant.add_logw(run, cur[r][j], runf);
                }
            }
        total += run;
#pragma unroll
        for (int r = 0; r < UPT; ++r)
#pragma unroll
            for (int j = 0; j < EPU; ++j) cur[r][j] = nxt[r][j];
    }
    total = warp_sum(total);
    if (lane == 0) s_tot[warp] = total;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kK3Warps; ++w) t += s_tot[w];
        rtot[blockIdx.x] = t;
    }
    if (shard.mail) {
        // sharded: the last CTA sums the range totals (integers: any order) and stores the shard total
        // into every rank's mailbox, then raises this rank's flag there
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            unsigned long long t = 0;
            for (int i = tid; i < (int)gridDim.x; i += kK3Threads) t += __ldcg(rtot + i);
            t = warp_sum(t);
            __syncthreads();
            if (lane == 0) s_tot[warp] = t;
            __syncthreads();
            if (tid == 0) {
                unsigned long long tot = 0;
#pragma unroll
                for (int w = 0; w < kK3Warps; ++w) tot += s_tot[w];
                for (int g = 0; g < shard.n_shards; ++g) shard.mail[g]->tot[shard.rank] = tot;
                __threadfence_system();
                for (int g = 0; g < shard.n_shards; ++g)
                    *(volatile unsigned long long*)&shard.mail[g]->flag_tot[shard.rank] = shard.stamp;
                *ticket = 0u;
            }
        }
    }
}

// shard total = sum of the range totals (multi-GPU: input of the cross-shard exclusive scan)
__global__ void __launch_bounds__(kThreads) k3a_sum_ranges(const Ctrl* ctrl, const unsigned long long* rtot, int n_ranges,
                                                          unsigned long long* total_out, int force) {
    __shared__ unsigned long long s_tot[kWarps];
    unsigned long long t = 0;
    rtot += (size_t)ctrl->rsum_sel * kRsumStride;
    if (force || ctrl->do_resample)
        for (int i = threadIdx.x; i < n_ranges; i += kThreads) t += rtot[i];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tot += s_tot[w];
        *total_out = tot;
    }
}

// ---- fused "finish step" (also a stand-alone kernel below): post-resample moments from the
// per-CTA partial records, buffer flip, moment-shift refresh.  Runs in ONE CTA (any size that is
// a multiple of 32, <= 1024).

__device__ inline void finish_body(Ctrl* ctrl, const FinishArgs& f, double (*s_w)[16], double* s_m) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int do_res = ctrl->do_resample;
    if (do_res && f.inline_moments && !f.no_moments && f.stats) {
        const int nm = moment_stride(f.nx);
        double acc[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) acc[k] = 0;
#pragma unroll 2
        const int mstride = f.mom_stride > 0 ? f.mom_stride : nm;
        for (int p = tid; p < f.n_mom_parts; p += nthreads) {    // fixed record -> thread assignment
            const double* r = f.mom_parts + (int64_t)p * mstride;
#pragma unroll
            for (int k = 0; k < 14; ++k) if (k < nm) acc[k] += __ldcg(r + k);
        }
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            if (k < nm) {
                const double v = warp_sum(acc[k]);
                if (lane == 0) s_w[warp][k] = v;
            }
        }
        __syncthreads();
        if (tid < nm) {
            double t = 0;
            for (int w = 0; w < nwarps; ++w) t += s_w[w][tid];
            s_m[tid] = t;
        }
        __syncthreads();
        if (tid == 0) {
            double* mean = f.stats + SMC_STATS_HEADER;
            double* cov = mean + f.nx;
            const double inv = 1.0 / (double)f.n_global;
            double m1[kFusedNxMax];
            for (int d = 0; d < f.nx; ++d) { m1[d] = s_m[d] * inv; mean[d] = ctrl->shift[d] + m1[d]; }
            int k = f.nx;
            for (int d = 0; d < f.nx; ++d)
                for (int d2 = 0; d2 <= d; ++d2) {
                    double v = s_m[k++] * inv - m1[d] * m1[d2];
                    cov[d * f.nx + d2] = v;
                    cov[d2 * f.nx + d] = v;
                }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (do_res) { ctrl->cur ^= 1; ctrl->uniform = 1; }
        if (f.stats && !f.no_moments)
            for (int d = 0; d < f.nx && d < kMaxNx; ++d) ctrl->shift[d] = f.stats[SMC_STATS_HEADER + d];
    }
}

// explicit shared-memory loads by 32-bit address (volatile + memory clobber: the window is rewritten
// every tile, so the loads must neither be cached in registers nor moved across the barriers)
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr) : "memory");
    return (uint32_t)v;
}
template <typename real> __device__ __forceinline__ real lds_real(uint32_t saddr);
template <> __device__ __forceinline__ float lds_real<float>(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
template <> __device__ __forceinline__ double lds_real<double>(uint32_t saddr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr) : "memory");
    return v;
}

// Resident CTAs per SM of the resampler (shared memory: (1 + NXF) tiles + the id window)
template <typename real, int NXF, bool FUSED = false> struct K3Occ {
    static constexpr int plain = (sizeof(real) == 4) ? (NXF <= 1 ? 4 : (NXF == 2 ? 3 : 2)) : (NXF <= 1 ? 3 : 2);
#ifndef SMC_FUSED_OCC_DROP
#define SMC_FUSED_OCC_DROP 1
#endif
    static constexpr int value = FUSED ? (plain > 2 ? plain - SMC_FUSED_OCC_DROP : 2) : plain;   // the fused functor needs the registers
};

// K3 proper.  Persistent grid: CTA b owns the contiguous tiles [b T / G, (b + 1) T / G) (static
// assignment, so the per-CTA moment partials are reproducible) and carries the CDF prefix from tile
// to tile.  Per tile:
//   A  the tile's log-weights arrive in shared memory by cp.async (16-byte swizzle), issued while the
//      previous tile was still being expanded                                                  | S1
//      cp.async of the tile's particles -> shared (awaited at S3); the id window is cleared
//   B  blocked read-back, fixed-point quantisation, thread / warp / CTA inclusive scan          | S2
//      (then the cp.async of the NEXT tile's log-weights is issued)
//   C  output boundaries e[k] = #{positions below the CDF value after item k} (32-bit, branch-free;
//      the exact fp64 predicate only when a position lies within round-off of a CDF value)      | S3
//   D  every item with offspring stores (local id + 1) as a u16 marker at its first output slot
//      (bank-swizzled window)                                                                   | S4
//   E  copy-out: slot p holds its ancestor's marker, or 0 = "same ancestor as slot p - 1" (probed
//      backwards; one step covers two offspring, a short loop the rest) -> particle payload from
//      shared -> coalesced global stores, moments.
//   Tiles with a very heavy item or more than W outputs take the marker + max-scan route instead,
//   which costs O(window) whatever the skew.
template <typename real, int NXF, bool LINEAR_W, class FM = NoFuse>
__global__ void __launch_bounds__(K3Geom<real>::THREADS, K3Occ<real, NXF, (FM::NX > 0)>::value)
k3_resample(K3Args<real> a, FM fm) {
    constexpr bool FUSED = FM::NX > 0;
    static_assert(!FUSED || (FM::NX == NXF && !LINEAR_W), "the fused functor moves the resampler's own payload");
    using G = K3Geom<real>;
    using unit = typename G::unit;
    constexpr int kK3Threads = G::THREADS, kK3Warps = kK3Threads / 32;
    constexpr int IPT = G::IPT, EPU = G::EPU;
    constexpr int TILE = kK3Threads * IPT;
    constexpr int UPT = IPT / EPU;                    // 16-byte units per thread (= 4)
    constexpr int CH = IPT + 2;                       // window slots per thread (even)
    constexpr int W = kK3Threads * CH;                // output window (tile + 12.5 % / 25 % slack)
    constexpr int NM = NXF + NXF * (NXF + 1) / 2;
    constexpr int NMS = NM > 0 ? NM : 1;
    constexpr uint32_t MAXD = 4 * IPT;                // marker + back-probe while no thread has more outputs than this
    constexpr int PADROW = kK3Threads + 2 * (kK3Threads / IPT);   // swizzled slots per row of kK3Threads positions
    constexpr int GUARD = 8;                          // zeroed slots in front of the window
    static_assert(UPT == 2 || UPT == 4, "swizzle assumes 2 or 4 16-byte units per thread");
    static_assert(CH % 2 == 0, "fill chunk must be word aligned");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unit* lw_s = reinterpret_cast<unit*>(smem_raw);                       // TILE reals, swizzled
    real* x_s = reinterpret_cast<real*>(smem_raw + sizeof(real) * TILE);  // NXF * TILE reals
    unsigned short* ids_raw = reinterpret_cast<unsigned short*>(smem_raw + sizeof(real) * TILE * (1 + NXF));
    unsigned short* ids = ids_raw + GUARD;
    __shared__ unsigned long long s_warp_tot[kK3Warps];
    __shared__ unsigned long long s_warp_all[kK3Warps];
    __shared__ uint32_t s_E[4];
    __shared__ int s_wmax[kK3Warps];
    __shared__ double s_mom[kK3Warps][16];
    __shared__ double s_m[16];
    __shared__ int s_last;

    Ctrl* const gctrl = a.ctrl;
    Ctrl* ctrl = gctrl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();
    pdl_trigger();
    __shared__ Ctrl s_ctrl;
    if (a.prologue_finalize) {
        // Sharded bootstrap step: every CTA waits for the G partial records in its own mailbox and finalises the step
        // on a private copy of the control block (identical arithmetic, identical result in every CTA of every
        // rank); CTA 0 also writes the statistics row, the last CTA to finish copies the block back (nobody else
        // touches the global one meanwhile).
        __shared__ double s_pout[32];
        for (int i = tid; i < (int)(sizeof(Ctrl) / 4); i += kK3Threads)
            reinterpret_cast<uint32_t*>(&s_ctrl)[i] = reinterpret_cast<const uint32_t*>(gctrl)[i];
        Mailbox* own = a.shard.mail[a.shard.rank];
        mailbox_wait(own->flag_rec, a.shard.n_shards, a.shard.stamp, &own->error, 1);
        K2Args f = a.pfin;
        f.ctrl = &s_ctrl; f.parts = &own->rec[0][0]; f.n_parts = a.shard.n_shards; f.rec_stride = 32;
        f.producer_rsum = f.producer_rsum && s_ctrl.uniform != 0;
        if (blockIdx.x != 0) f.stats = nullptr;
        if (warp == 0) finalize_warp(f, s_pout);
        __syncthreads();
        ctrl = &s_ctrl;
    }
    if (!a.force && !ctrl->do_resample) {
        if (a.prologue_finalize && blockIdx.x == 0 && tid == 0) *gctrl = s_ctrl;
        // nothing to resample: only the moment-shift refresh of the fused finish remains
        if (a.fuse_finish && blockIdx.x == 0 && tid == 0 && a.fin.stats && !a.fin.no_moments)
            for (int d = 0; d < a.fin.nx && d < kMaxNx; ++d) ctrl->shift[d] = a.fin.stats[SMC_STATS_HEADER + d];
        // sharded: the finish kernels of all ranks wait for this rank's flag (no data this step)
        if (a.shard.mail && blockIdx.x == 0 && tid < a.shard.n_shards)
            *(volatile unsigned long long*)&a.shard.mail[tid]->flag_mom[a.shard.rank] = a.shard.stamp;
        return;
    }
    const int cur = ctrl->cur;
    const real* xsrc = cur ? a.x1 : a.x0;
    real* xdst = cur ? a.dst0 : a.dst1;
    real* const* peer = cur ? a.peer0 : a.peer1;

    const SearchParams sp = load_search(ctrl, a.n_global);
    FastSearch fs;
    fs.bias = sp.fmask - sp.uq;
    fs.sh = (uint32_t)(sp.F - 32);
    fs.m32 = (uint32_t)(sp.margin >> fs.sh) + 2u;
    fs.n_global = (uint32_t)a.n_global;
    Quant<real> quant;
    quant.init(ctrl, a.n_global, LINEAR_W);
    const real n_real = (real)a.n_global;
    const real fill = LINEAR_W ? real(0) : real(-INFINITY);

    const int64_t t_begin = k3_range_begin(blockIdx.x, a.n_tiles, gridDim.x);
    const int64_t t_end = k3_range_begin(blockIdx.x + 1, a.n_tiles, gridDim.x);

    // log-weights of one tile -> lw_s: cp.async for interior tiles, a guarded copy for the ragged last one
    auto fetch_lw = [&](int64_t tile) {
        const int64_t tile_base = tile * TILE;
        if (tile_base + TILE <= a.n) {
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int f = r * kK3Threads + tid;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&lw_s[f ^ ((f >> 3) & (UPT - 1))]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst),
                             "l"(a.wsrc + tile_base + (int64_t)f * EPU) : "memory");
            }
        } else {
            const int n_in_tile = (int)(a.n - tile_base);
            const unit* wtile = reinterpret_cast<const unit*>(a.wsrc + tile_base);
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int f = r * kK3Threads + tid;
                real tmp[EPU];
#pragma unroll
                for (int j = 0; j < EPU; ++j) tmp[j] = fill;
                if (f * EPU < n_in_tile) {
                    unit_to_array(wtile[f], tmp);
#pragma unroll
                    for (int j = 0; j < EPU; ++j) if (f * EPU + j >= n_in_tile) tmp[j] = fill;
                }
                real* dstp = reinterpret_cast<real*>(&lw_s[f ^ ((f >> 3) & (UPT - 1))]);
#pragma unroll
                for (int j = 0; j < EPU; ++j) dstp[j] = tmp[j];
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if (t_begin < t_end) fetch_lw(t_begin);

    // exclusive CDF prefix of the first tile: the lower shards' totals (multi-GPU) and the totals of
    // the ranges below this CTA's
    unsigned long long prefix, cdf_total;
    {
        const unsigned long long* totals = a.totals;
        if (a.shard.mail) {   // every rank's weight total of this step must have arrived in the own mailbox
            Mailbox* own = a.shard.mail[a.shard.rank];
            if (a.prologue_finalize && !(sizeof(real) == 4 && ctrl->rsum_valid)) {
                // Rare: the a-priori bound left this step's weights too little resolution and no range pre-pass was
                // launched (the step was enqueued as a bootstrap step).  Run it here: every CTA sums the ranges it
                // owns against the normaliser, a grid barrier (all CTAs of this persistent grid are resident), CTA 0
                // publishes the shard total to every rank.
                unsigned long long* rt = const_cast<unsigned long long*>(a.rtot) + (size_t)ctrl->rsum_sel * kRsumStride;
                for (int r = (int)blockIdx.x * a.ranges_per_cta; r < ((int)blockIdx.x + 1) * a.ranges_per_cta; ++r) {
                    const int64_t tr0 = k3_range_begin(r, a.n_tiles, (int64_t)gridDim.x * a.ranges_per_cta);
                    const int64_t tr1 = k3_range_begin(r + 1, a.n_tiles, (int64_t)gridDim.x * a.ranges_per_cta);
                    unsigned long long acc = 0;
                    double runf = 0.0;
                    for (int64_t i = tr0 * TILE + tid; i < tr1 * TILE && i < a.n; i += kK3Threads)
                        acc = quant.add_logw(acc, a.wsrc[i], runf);          // fp32: items are quantised one by one
                    acc = warp_sum(acc);
                    __syncthreads();
                    if (lane == 0) s_warp_tot[warp] = acc;
                    __syncthreads();
                    if (tid == 0) {
                        unsigned long long t = 0;
#pragma unroll
                        for (int w = 0; w < kK3Warps; ++w) t += s_warp_tot[w];
                        rt[r] = t;
                    }
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    atomicAdd(a.gbar, 1u);
                    while (*(volatile unsigned int*)a.gbar < gridDim.x) __nanosleep(40);
                }
                __syncthreads();
                __threadfence();
                if (blockIdx.x == 0) {
                    unsigned long long t = 0;
                    const int nr = (int)gridDim.x * a.ranges_per_cta;
                    for (int i = tid; i < nr; i += kK3Threads) t += __ldcg(rt + i);
                    t = warp_sum(t);
                    __syncthreads();
                    if (lane == 0) s_warp_tot[warp] = t;
                    __syncthreads();
                    if (tid == 0) {
                        unsigned long long tot = 0;
#pragma unroll
                        for (int w = 0; w < kK3Warps; ++w) tot += s_warp_tot[w];
                        for (int g = 0; g < a.shard.n_shards; ++g) a.shard.mail[g]->tot[a.shard.rank] = tot;
                        __threadfence_system();
                        for (int g = 0; g < a.shard.n_shards; ++g)
                            *(volatile unsigned long long*)&a.shard.mail[g]->flag_tot[a.shard.rank] = a.shard.stamp;
                    }
                    __syncthreads();
                }
                mailbox_wait(own->flag_tot, a.shard.n_shards, a.shard.stamp, &own->error, 2);
            } else if (!(sizeof(real) == 4 && ctrl->rsum_valid)) {   // bootstrap steps: the totals came with the partial records
                mailbox_wait(own->flag_tot, a.shard.n_shards, a.shard.stamp, &own->error, 2);
            }
            totals = own->tot;
        }
        unsigned long long part = 0, all = 0;
        const int n_ranges = (int)gridDim.x * a.ranges_per_cta;
        const unsigned long long* rtot = a.rtot + (size_t)ctrl->rsum_sel * kRsumStride;
        for (int i = tid; i < n_ranges; i += kK3Threads) {
            const unsigned long long v = __ldcg(rtot + i);
            if (i < (int)blockIdx.x * a.ranges_per_cta) part += v;
            all += v;
        }
        if (totals) {
            all = 0;                                             // the global total is the sum over ALL shards
            const int ns = a.shard.mail ? a.shard.n_shards : a.n_totals;
            for (int r = tid; r < ns; r += kK3Threads) {
                const unsigned long long v = __ldcg(totals + r);
                if (r < a.rank) part += v;
                all += v;
            }
        }
        part = warp_sum(part);
        all = warp_sum(all);
        if (lane == 0) { s_warp_tot[warp] = part; s_warp_all[warp] = all; }
        __syncthreads();
        prefix = a.cdf_offset;
        cdf_total = 0;
#pragma unroll
        for (int w = 0; w < kK3Warps; ++w) { prefix += s_warp_tot[w]; cdf_total += s_warp_all[w]; }
        __syncthreads();
    }
    // fp32 mode: the stratified positions are laid over the actual fixed-point total (see RhoMap)
    const RhoMap rm = make_rho(cdf_total, a.n_global, fs.bias);

    real shift[NXF > 0 ? NXF : 1];
#pragma unroll
    for (int d = 0; d < NXF; ++d) shift[d] = (real)ctrl->shift[d];
    double macc_d[NMS];
#pragma unroll
    for (int k = 0; k < NMS; ++k) macc_d[k] = 0.0;

    // ---- fused functor state: observation of step t, online (max, sum e, sum e^2) of its log-weights ---
    constexpr int FNX = FUSED ? FM::NX : 1, FNZ = FUSED ? FM::NZ : 1;
    real fz[FNZ];
    real f_tm = -INFINITY, f_acc0 = 0, f_acc1 = 0;
    const real f_lw_uniform = (real)(-log((double)a.n_global));
    if constexpr (FUSED) {
#pragma unroll
        for (int k = 0; k < FNZ; ++k) fz[k] = (real)a.fk.z[k];
    }
    // one quad of resampled particles (global indices gi0 .. gi0 + 3, gi0 % 4 == 0; `mask` = lanes that
    // belong to this tile): Philox normals of step t, propagate, log-likelihood, stores, online sums --
    // exactly k1_chunk's arithmetic on exactly its Philox blocks, so the particles are bit-identical
    // to the unfused pipeline's
    auto process_quad = [&](real (&xt)[FNX][4], const int64_t gi0, const unsigned mask, real* xout) {
        if constexpr (FUSED) {
            real nrm[FNX][4];
            Draw<FNX>::run(a.fk.keys, (uint64_t)gi0 >> 2, a.fk.t, nrm);
            real lwq[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                real xv[FNX], nv[FNX];
#pragma unroll
                for (int d = 0; d < FNX; ++d) { xv[d] = xt[d][j]; nv[d] = nrm[d][j]; }
                fm.propagate(xv, nv);
#pragma unroll
                for (int d = 0; d < FNX; ++d) xt[d][j] = xv[d];
                real v = f_lw_uniform + fm.loglik(xv, fz);
                v = fmax(v, real(-INFINITY));                     // NaN -> -inf (zero weight)
                lwq[j] = ((mask >> j) & 1u) ? v : real(-INFINITY);
            }
            const int64_t li = gi0 - a.offset;
            if (mask == 15u) {
#pragma unroll
                for (int d = 0; d < FNX; ++d) store4<real>(xout + d * a.ld + li, xt[d]);
                store4<real>(a.fk.lw_out + li, lwq);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((mask >> j) & 1u) {
#pragma unroll
                        for (int d = 0; d < FNX; ++d) xout[d * a.ld + li + j] = xt[d][j];
                        a.fk.lw_out[li + j] = lwq[j];
                    }
            }
            real cmax = fmax(fmax(lwq[0], lwq[1]), fmax(lwq[2], lwq[3]));
            if (cmax > f_tm) {
                const real sc = fast_exp(f_tm - cmax);            // f_tm = -inf -> 0
                f_acc0 *= sc; f_acc1 *= sc * sc; f_tm = cmax;
            }
            if (f_tm != real(-INFINITY)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const real e = fast_exp(lwq[j] - f_tm);       // masked lanes: exp(-inf) = 0
                    f_acc0 += e; f_acc1 = fma(e, e, f_acc1);
                }
            }
        }
    };

    // window slot of output position s: two pad slots per IPT positions, so that the CTA-wide marker
    // stores of phase D (lane stride ~ IPT positions) spread over the banks
    auto slot = [](uint32_t s) -> uint32_t { return s + ((s / (uint32_t)IPT) << 1); };
    const int slot_t = tid + 2 * (tid / IPT);                 // slot(it * kK3Threads + tid) = it * PADROW + slot_t
    const int slot_b = slot_t - ((tid % IPT) ? 1 : 3);        // slot of the position before (tid >= 1 or it >= 1)

    for (int64_t tile = t_begin; tile < t_end; ++tile) {
        const int64_t tile_base = tile * TILE;
        const unsigned long long tile_prefix = prefix;
        const int n_in_tile = tile_base + TILE <= a.n ? TILE : (int)(a.n - tile_base);

        // ---- phase A ---------------------------------------------------------------------------
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();   // S1: lw_s complete; every thread has finished the previous tile's copy-out
        // particle payload: asynchronous global->shared copies (LDGSTS, no register staging), issued
        // only now because x_s was still being gathered from until S1; awaited at S3
        if (NXF > 0) {
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int f = r * kK3Threads + tid;
                if (f * EPU < n_in_tile) {
#pragma unroll
                    for (int d = 0; d < NXF; ++d) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(x_s + d * TILE + f * EPU);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst),
                                     "l"(xsrc + d * a.ld + tile_base + f * EPU) : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        {   // clear the marker window (guard slots included)
            uint4* o4 = reinterpret_cast<uint4*>(ids_raw);
            constexpr int N16 = (GUARD + CH * PADROW + 8) * 2 / 16;
            for (int p = tid; p < N16; p += kK3Threads) o4[p] = make_uint4(0, 0, 0, 0);
        }

        // ---- phase B: fixed-point weights, thread / warp / CTA scan -------------------------------
        unsigned long long c[IPT];                               // fp64: inclusive fixed-point CDF within the thread
        float pf[IPT];                                           // fp32: the weights the fixed-point values are rounded from
        unsigned long long run = 0;
        {
            double runf = 0.0;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                const int g = UPT * tid + k;
                real tmp[EPU];
                unit_to_array(lw_s[g ^ ((g >> 3) & (UPT - 1))], tmp);
#pragma unroll
                for (int j = 0; j < EPU; ++j) {
                    if constexpr (sizeof(real) == 4) {
                        const float pv = LINEAR_W ? quant.p_lin(fminf(fmaxf(tmp[j] * n_real, 0.f), n_real))   // NaN / negative -> 0
                                                  : quant.p_logw(tmp[j]);
                        pf[k * EPU + j] = pv;
                        run += quant_a(pv);
                    } else {
                        if (LINEAR_W) {
                            real v = fmin(fmax(tmp[j] * n_real, real(0)), n_real);   // NaN / negative -> 0, cap at N
                            run = quant.add(run, v, runf);
                        } else {
                            run = quant.add_logw(run, tmp[j], runf);
                        }
                        c[k * EPU + j] = run;                    // inclusive within thread
                    }
                }
            }
        }
        unsigned long long incl_w = run;                         // warp inclusive scan of thread totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl_w, o);
            if (lane >= o) incl_w += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl_w;
        __syncthreads();   // S2 (lw_s is dead from here on, except for the rare exact path below)
        unsigned long long warp_excl = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kK3Warps; ++w) {
            unsigned long long t = s_warp_tot[w];
            if (w < warp) warp_excl += t;
            tile_total += t;
        }
        prefix = tile_prefix + tile_total;

        // ---- phase C: output boundaries (32-bit) -------------------------------------------------
        // e[k] = number of outputs whose stratified position lies below the CDF value after item k
        const unsigned long long base_raw = tile_prefix + warp_excl + (incl_w - run);
        const unsigned long long base_b = base_raw + fs.bias;                                    // + bias folded in
        uint32_t e_start, e[IPT];
        if constexpr (sizeof(real) == 4) {
            // The thread's first and last boundary are exact (64 x 64 -> 128 bit products of the global CDF
            // values with rho), so consecutive threads, tiles, CTAs and shards tile the outputs exactly; the
            // items in between follow the thread-local fp32 sums (skewed tiles: exact, after S3).
            const unsigned long long Ps = map_pos(base_raw, rm), Pe = map_pos(base_raw + run, rm);
            e_start = (uint32_t)(Ps >> 32);
            f32_boundaries<IPT>(pf, Ps, (uint32_t)(Pe >> 32), rm.rho_f, e);
        } else
        {
            uint32_t near = 0;
            auto fast_e = [&](unsigned long long Cb) -> uint32_t {
                const uint32_t hi = (uint32_t)(Cb >> 32), lo = (uint32_t)Cb;
                const uint32_t frac_hi = __funnelshift_r(lo, hi, fs.sh);    // bits [F-32, F) of Cb
                near |= (frac_hi + fs.m32 <= 2u * fs.m32) ? 1u : 0u;        // near 0 or near 2^32 (wraps)
                return hi >> fs.sh;
            };
            e_start = fast_e(base_b);
#pragma unroll
            for (int k = 0; k < IPT; ++k) e[k] = fast_e(base_b + c[k]);
            if (near) {
                // rare (~1e-6 per thread): redo with the exact predicate where needed; the running CDF is
                // recomputed from shared memory so that c[] need not stay live
                e_start = fast_outputs_below(base_b, fs, sp);
                unsigned long long run2 = 0;
                double runf2 = 0.0;
#pragma unroll
                for (int k = 0; k < UPT; ++k) {
                    const int g = UPT * tid + k;
                    real tmp[EPU];
                    unit_to_array(lw_s[g ^ ((g >> 3) & (UPT - 1))], tmp);
#pragma unroll
                    for (int j = 0; j < EPU; ++j) {
                        if (LINEAR_W) {
                            real v = fmin(fmax(tmp[j] * n_real, real(0)), n_real);
                            run2 = quant.add(run2, v, runf2);
                        } else {
                            run2 = quant.add_logw(run2, tmp[j], runf2);
                        }
                        e[k * EPU + j] = fast_outputs_below(base_b + run2, fs, sp);
                    }
                }
            }
        }
        auto tail_clamps = [&]() {   // cumsum[-1] = 1.0 clamp (particle_filter.py:163): the globally last particle owns the tail
            const int64_t j0 = a.offset + tile_base + (int64_t)tid * IPT;
            if (j0 + IPT >= a.n_global) {
#pragma unroll
                for (int k = 0; k < IPT; ++k) if (j0 + k >= a.n_global - 1) e[k] = fs.n_global;
                // a thread that holds padding only starts where the last particle's range ends -- otherwise a
                // quantised total slightly below N 2^F would give its first padding item the tail outputs too
                if (j0 >= a.n_global) e_start = fs.n_global;
            }
            // the quantised total may exceed N 2^F by round-off: no boundary may pass N
            // (boundaries are monotone, so one test per thread suffices)
            if (e[IPT - 1] > fs.n_global) {
                e_start = e_start < fs.n_global ? e_start : fs.n_global;
#pragma unroll
                for (int k = 0; k < IPT; ++k) e[k] = e[k] < fs.n_global ? e[k] : fs.n_global;
            }
        };
        tail_clamps();
        // (padding items beyond the shard carry zero weight, hence no outputs of their own)
        if (tid == 0) {
            s_E[0] = e_start;
            if (peer) s_E[2] = (uint32_t)((int64_t)e_start / a.shard_n);   // shard that owns the first output
        }
        if (tid == kK3Threads - 1) {
            s_E[1] = e[IPT - 1];
            if (peer) s_E[3] = (uint32_t)((int64_t)(e[IPT - 1] > 0u ? e[IPT - 1] - 1u : 0u) / a.shard_n);   // ... and the last
        }
        if (NXF > 0) asm volatile("cp.async.wait_all;\n" ::: "memory");
        const int heavy = __syncthreads_or(e[IPT - 1] - e_start > MAXD);   // S3: publishes s_E, x_s and the cleared window;
                                                           //     nobody reads lw_s any more
        if constexpr (sizeof(real) == 4) {
            if (heavy) {   // skewed tile: a thread-local fp32 sum no longer resolves single outputs -> exact boundaries
                exact_boundaries<IPT>(pf, base_raw, rm, e);
                tail_clamps();
            }
        }
        if (tile + 1 < t_end) fetch_lw(tile + 1);          // next tile's log-weights travel during D / E
        const uint32_t E0 = s_E[0], E1 = s_E[1];
        // Memory safety on inconsistent input (weights that do not sum to one can wrap the 64-bit
        // CDF): a tile whose output range is not inside [0, N] writes nothing.
        const bool sane = E0 <= E1 && E1 <= fs.n_global;
        const uint32_t count = sane ? E1 - E0 : 0u;
        // fused mode lays the window out on global quads: position r sits at r + (E0 & 3)
        const uint32_t s0 = FUSED ? (E0 & 3u) : 0u;
        const bool direct = !heavy && count + s0 <= (uint32_t)W;

        real macc[NMS];
#pragma unroll
        for (int k = 0; k < NMS; ++k) macc[k] = 0;

        // payload of local ancestor (li1 - 1) -> output slot `o` of this thread's column; moments
        auto emit = [&](real* dcol, int32_t* acol, const int li1, const int32_t abase) {
            if (acol) *acol = abase + li1;
            if (NXF > 0) {
                real xv[NXF > 0 ? NXF : 1];
#pragma unroll
                for (int d = 0; d < NXF; ++d) {
                    xv[d] = x_s[d * TILE + li1 - 1];
                    dcol[d * a.ld] = xv[d];
                }
                int k = 0;
#pragma unroll
                for (int d = 0; d < NXF; ++d) { xv[d] -= shift[d]; macc[k] += xv[d]; ++k; }
#pragma unroll
                for (int d = 0; d < NXF; ++d)
#pragma unroll
                    for (int d2 = 0; d2 <= d; ++d2) { macc[k] = fma(xv[d], xv[d2], macc[k]); ++k; }
            }
        };

        if (direct) {
            // ---- phase D (common case): one marker per item with offspring -----------------------
            {
                uint32_t r0 = e_start - E0;
#pragma unroll
                for (int k = 0; k < IPT; ++k) {
                    const uint32_t r1 = e[k] - E0;
                    if (r0 < r1 && r1 <= count) ids[slot(r0 + s0)] = (unsigned short)(tid * IPT + k + 1);
                    r0 = r1;
                }
            }
            __syncthreads();   // S4
            if constexpr (FUSED) {
                // ---- phase E (fused): gather a quad of ancestors, propagate + weigh it, store step t ------
                const uint32_t wend = s0 + count;                         // valid window positions [s0, wend)
                const uint32_t nquad = (wend + 3u) >> 2;
                const int64_t gw0 = (int64_t)(E0 - s0);                   // global index of window position 0
                for (uint32_t q = tid; q < nquad; q += kK3Threads) {
                    const uint32_t w0 = 4u * q, sl = slot(w0);            // four consecutive slots of one 16-group
                    const uint32_t p01 = *reinterpret_cast<const uint32_t*>(ids + sl);
                    const uint32_t p23 = *reinterpret_cast<const uint32_t*>(ids + sl + 2);
                    uint32_t m[4] = {p01 & 0xffffu, p01 >> 16, p23 & 0xffffu, p23 >> 16};
                    if (m[0] == 0u && w0 > s0) {                          // ancestor of the slots before this quad
                        m[0] = ids[sl - ((w0 & (uint32_t)(IPT - 1)) ? 1u : 3u)];   // the slot just before (one step covers most)
                        if (m[0] == 0u) {
                            uint32_t qq = w0 - 1u;
                            do { --qq; m[0] = ids[slot(qq)]; } while (m[0] == 0u && qq > 0u);
                        }
                    }
#pragma unroll
                    for (int j = 1; j < 4; ++j) m[j] = m[j] ? m[j] : m[j - 1];
                    unsigned mask = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (w0 + j >= s0 && w0 + j < wend) mask |= 1u << j;
                        if (m[j] == 0u) m[j] = 1u;                        // lanes in front of the tile's first output
                    }
                    real xt[FNX][4];
#pragma unroll
                    for (int d = 0; d < FNX; ++d)
#pragma unroll
                        for (int j = 0; j < 4; ++j) xt[d][j] = x_s[d * TILE + m[j] - 1];
                    // post-resample moments of step t-1 (the gathered, not yet propagated, particles)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((mask >> j) & 1u) {
                            real xc[FNX];
                            int k = 0;
#pragma unroll
                            for (int d = 0; d < FNX; ++d) { xc[d] = xt[d][j] - shift[d]; macc[k] += xc[d]; ++k; }
#pragma unroll
                            for (int d = 0; d < FNX; ++d)
#pragma unroll
                                for (int d2 = 0; d2 <= d; ++d2) { macc[k] = fma(xc[d], xc[d2], macc[k]); ++k; }
                        }
                    process_quad(xt, gw0 + w0, mask, xdst);
                }
            } else {
            // ---- phase E: coalesced copy-out -----------------------------------------------------
            const int64_t g0 = (int64_t)E0 - a.offset;
            // cross-shard migration: almost every tile's outputs [E0, E1) belong to ONE shard (the own
            // one or a neighbour); such a tile takes the same copy-out as the single-GPU case, only
            // with a destination in that shard's buffer (NVLink peer stores when it is not the own)
            const bool one_shard = !peer || s_E[2] == s_E[3];
            real* dwin = (!peer ? xdst + g0 : peer[s_E[2]] + ((int64_t)E0 - (int64_t)s_E[2] * a.shard_n)) + tid;
            int32_t* awin = a.ancestors ? a.ancestors + g0 + tid : nullptr;
            const int32_t abase = (int32_t)(a.offset + tile_base) - 1;   // markers are index + 1
            const int nfull = (int)(count / kK3Threads);                  // rows in which every thread has an output
            // cross-shard migration: outputs [E0, E1) start in shard s_E[2]; `thr` = first window
            // position owned by the next shard
            int64_t pr = 0, thr = 0;
            if (peer) { pr = (int64_t)s_E[2]; thr = (pr + 1) * a.shard_n - (int64_t)E0; }
            auto row = [&](const int it) {
                int li1 = (int)ids[it * PADROW + slot_t];
                if (li1 == 0) {                                           // same ancestor as the slot before
                    li1 = (int)ids[it * PADROW + slot_b];
                    if (li1 == 0) {                                       // three or more offspring: walk back
                        uint32_t q = (uint32_t)(it * kK3Threads + tid) - 1u;
                        do { --q; li1 = (int)ids[slot(q)]; } while (li1 == 0 && q > 0u);
                        if (li1 == 0) li1 = 1;                            // (only on inconsistent input)
                    }
                }
                if (!peer) {
                    emit(dwin + it * kK3Threads, awin ? awin + it * kK3Threads : nullptr, li1, abase);
                } else {
                    const int64_t p = (int64_t)it * kK3Threads + tid;
                    while (p >= thr) { ++pr; thr += a.shard_n; }          // rows advance, so pr only grows
                    emit(peer[pr] + ((int64_t)E0 + p - pr * a.shard_n), nullptr, li1, abase);
                }
            };
            const bool tail = (int)(count - nfull * kK3Threads) > tid;
            if (NXF > 0 && one_shard && !awin) {
                // the hot form: no ancestor trace, one destination shard; all addresses are immediates
                // Row `it` = output position it * 256 + tid.  A slot that holds no marker (0) shares the
                // ancestor of the slot before; a run of three or more offspring is rare and is finished
                // after the unrolled rows (bit mask `pending`), so that the rows themselves are branch-free.
                const uint32_t a_t = (uint32_t)__cvta_generic_to_shared(ids + slot_t);
                const uint32_t a_b = (uint32_t)__cvta_generic_to_shared(ids + slot_b);
                const uint32_t a_x = (uint32_t)__cvta_generic_to_shared(x_s) - (uint32_t)sizeof(real);   // markers are index + 1
                uint32_t pending = 0;
                auto payload = [&](const int it, const uint32_t li1) {
                    real xv[NXF > 0 ? NXF : 1];
#pragma unroll
                    for (int d = 0; d < NXF; ++d) {
                        xv[d] = lds_real<real>(a_x + (uint32_t)(d * TILE * sizeof(real)) + li1 * (uint32_t)sizeof(real));
                        dwin[d * a.ld + it * kK3Threads] = xv[d];
                    }
                    int k = 0;
#pragma unroll
                    for (int d = 0; d < NXF; ++d) { xv[d] -= shift[d]; macc[k] += xv[d]; ++k; }
#pragma unroll
                    for (int d = 0; d < NXF; ++d)
#pragma unroll
                        for (int d2 = 0; d2 <= d; ++d2) { macc[k] = fma(xv[d], xv[d2], macc[k]); ++k; }
                };
                auto hot = [&](const int it) {
                    const uint32_t m0 = lds_u16(a_t + (uint32_t)(it * PADROW * 2));
                    const uint32_t m1 = lds_u16(a_b + (uint32_t)(it * PADROW * 2));
                    const uint32_t li1 = m0 ? m0 : m1;
                    if (li1) payload(it, li1);
                    else pending |= 1u << it;
                };
                constexpr int NB = IPT - 2;          // rows virtually every tile has in full
                if (nfull >= NB) {
#pragma unroll
                    for (int it = 0; it < NB; ++it) hot(it);
#pragma unroll
                    for (int it = NB; it < CH; ++it)
                        if (it < nfull) hot(it);
                } else {
#pragma unroll
                    for (int it = 0; it < NB; ++it)
                        if (it < nfull) hot(it);
                }
                if (tail) hot(nfull);
                while (pending) {                     // three or more offspring: walk back to the marker
                    const int it = __ffs(pending) - 1;
                    pending &= pending - 1u;
                    uint32_t q = (uint32_t)(it * kK3Threads + tid) - 1u, li1;
                    do { --q; li1 = ids[slot(q)]; } while (li1 == 0u && q > 0u);
                    if (li1 == 0u) li1 = 1u;          // (only on inconsistent input)
                    payload(it, li1);
                }
            } else {
                for (int it = 0; it < nfull; ++it) row(it);
                if (tail) row(nfull);
            }
            }   // !FUSED
        } else {
            // ---- skewed tile / cross-shard migration: head markers + max-scan fill, window by window
            for (uint32_t wbase = 0; wbase < count; wbase += W) {
                if (wbase > 0) {
                    uint4* o4 = reinterpret_cast<uint4*>(ids);
                    for (int p = tid; p < W * 2 / 16; p += kK3Threads) o4[p] = make_uint4(0, 0, 0, 0);
                    __syncthreads();
                }
                {   // head markers (local index + 1) at the first output of every item with offspring
                    uint32_t r0 = e_start - E0;
#pragma unroll
                    for (int k = 0; k < IPT; ++k) {
                        const uint32_t r1 = e[k] - E0;
                        if (r1 > r0 && r1 > wbase && r0 < wbase + W) {
                            const uint32_t pos = (r0 > wbase ? r0 : wbase) - wbase;
                            ids[pos] = (unsigned short)(tid * IPT + k + 1);
                        }
                        r0 = r1;
                    }
                }
                __syncthreads();
                {   // forward fill = inclusive max-scan (markers increase with position)
                    uint32_t* o32 = reinterpret_cast<uint32_t*>(ids) + tid * (CH / 2);
                    uint32_t wv[CH / 2];
                    int tmax = 0;
#pragma unroll
                    for (int k = 0; k < CH / 2; ++k) {
                        wv[k] = o32[k];
                        int lo = wv[k] & 0xffff, hi = wv[k] >> 16;
                        tmax = max(tmax, max(lo, hi));
                    }
                    int inc = tmax;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        int t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc = max(inc, t);
                    }
                    int excl = __shfl_up_sync(0xffffffffu, inc, 1);
                    if (lane == 0) excl = 0;
                    if (lane == 31) s_wmax[warp] = inc;
                    __syncthreads();
                    int carry = excl;
#pragma unroll
                    for (int w = 0; w < kK3Warps; ++w) if (w < warp) carry = max(carry, s_wmax[w]);
#pragma unroll
                    for (int k = 0; k < CH / 2; ++k) {
                        int lo = wv[k] & 0xffff, hi = wv[k] >> 16;
                        lo = max(lo, carry); carry = lo;
                        hi = max(hi, carry); carry = hi;
                        o32[k] = (uint32_t)lo | ((uint32_t)hi << 16);
                    }
                }
                __syncthreads();
                {
                    const uint32_t rem = count - wbase;
                    const int cnt = rem < (uint32_t)W ? (int)rem : W;
                    const int64_t g0 = (int64_t)E0 + wbase;                  // global index of window slot 0
                    const int32_t abase = (int32_t)(a.offset + tile_base) - 1;
                    for (int p = tid; p < cnt; p += kK3Threads) {
                        const int li1 = (int)ids[p];
                        const int64_t gi = g0 + p;
                        if (!peer) {
                            emit(xdst + (gi - a.offset), a.ancestors ? a.ancestors + (gi - a.offset) : nullptr, li1, abase);
                        } else {                                             // cross-shard migration
                            const int64_t r = gi / a.shard_n;
                            emit(peer[r] + (gi - r * a.shard_n), nullptr, li1, abase);
                        }
                    }
                }
                __syncthreads();
            }
            if constexpr (FUSED) {
                // skewed tile: the window loop has written the resampled particles of this tile to the
                // destination buffer; run the functor over them in place (block-visible after the barrier)
                if (count > 0u) {
                    const int64_t gq0 = (int64_t)(E0 & ~3u);
                    const uint32_t nq = (uint32_t)((((int64_t)E1 + 3) >> 2) - (gq0 >> 2));
                    for (uint32_t q = tid; q < nq; q += kK3Threads) {
                        const int64_t gi0 = gq0 + 4 * (int64_t)q;
                        unsigned mask = 0;
                        real xt[FNX][4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int64_t gi = gi0 + j;
                            const bool valid = gi >= (int64_t)E0 && gi < (int64_t)E1;
                            if (valid) mask |= 1u << j;
#pragma unroll
                            for (int d = 0; d < FNX; ++d) xt[d][j] = valid ? xdst[d * a.ld + gi - a.offset] : real(0);
                        }
                        process_quad(xt, gi0, mask, xdst);
                    }
                }
                __syncthreads();
            }
        }
        if (NM > 0) {
#pragma unroll
            for (int k = 0; k < NMS; ++k) macc_d[k] += (double)macc[k];
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    if constexpr (FUSED) {
        // per-CTA (max, sum e, sum e^2) record of step t's log-weights (k1_propagate_weight's layout)
        __shared__ real s_fmax[kK3Warps];
        __shared__ double s_fred[kK3Warps][2];
        const real wmax = warp_max(f_tm);
        if (lane == 0) s_fmax[warp] = wmax;
        __syncthreads();
        real bmax = s_fmax[0];
#pragma unroll
        for (int w = 1; w < kK3Warps; ++w) bmax = s_fmax[w] > bmax ? s_fmax[w] : bmax;
        const real sc = (f_tm == real(-INFINITY)) ? real(0) : fast_exp(f_tm - bmax);
        const real v0 = warp_sum(f_acc0 * sc), v1 = warp_sum(f_acc1 * sc * sc);
        if (lane == 0) { s_fred[warp][0] = (double)v0; s_fred[warp][1] = (double)v1; }
        __syncthreads();
        if (tid == 0) {
            double t0 = 0, t1 = 0;
#pragma unroll
            for (int w = 0; w < kK3Warps; ++w) { t0 += s_fred[w][0]; t1 += s_fred[w][1]; }
            double* out = a.fk.parts + (int64_t)blockIdx.x * 3;
            out[0] = (double)bmax; out[1] = t0; out[2] = t1;
        }
    }

    // ---- per-CTA post-resample moment partials (+ the fused finish in the last CTA to arrive) ----
    if (NM > 0 && a.mom_parts) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NM; ++k) {
            double v = warp_sum(macc_d[k]);
            if (lane == 0) s_mom[warp][k] = v;
        }
        __syncthreads();
        if (tid < NM) {
            double s = 0;
#pragma unroll
            for (int w = 0; w < kK3Warps; ++w) s += s_mom[w][tid];
            a.mom_parts[(int64_t)blockIdx.x * NM + tid] = s;
        }
    }
    if (a.fuse_finish) {
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            finish_body(ctrl, a.fin, s_mom, s_m);
            if constexpr (FUSED) {
                // ... then the finalize of step t from the records every CTA has just written
                __shared__ double s_fout[32];
                __shared__ double s_ftmp[kWarps];
                __syncthreads();
                K2Args f2 = a.fk.fin;
                f2.ctrl = ctrl; f2.parts = a.fk.parts; f2.n_parts = (int)gridDim.x;
                finalize_body(f2, s_fout, s_ftmp);
            }
            if (tid == 0) *a.ticket = 0u;
        }
    } else if (a.shard.mail) {
        // sharded: every CTA's peer stores are made visible system-wide before it takes a ticket; the
        // last CTA folds the per-CTA moment partials into one record, stores it into every rank's
        // mailbox and raises this rank's flag there -- which also tells the other ranks that all
        // particles this rank migrated into their buffers have landed
        __threadfence_system();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (a.prologue_finalize && tid == 0) { *gctrl = s_ctrl; *a.gbar = 0u; }
            if (NM > 0 && a.mom_parts) {
                double acc[NMS];
#pragma unroll
                for (int k = 0; k < NMS; ++k) acc[k] = 0.0;
                for (int p = tid; p < (int)gridDim.x; p += kK3Threads)        // fixed record -> thread assignment
#pragma unroll
                    for (int k = 0; k < NM; ++k) acc[k] += __ldcg(a.mom_parts + (int64_t)p * NM + k);
                __syncthreads();
#pragma unroll
                for (int k = 0; k < NM; ++k) {
                    const double v = warp_sum(acc[k]);
                    if (lane == 0) s_mom[warp][k] = v;
                }
                __syncthreads();
                if (tid < NM) {
                    double t = 0;
#pragma unroll
                    for (int w = 0; w < kK3Warps; ++w) t += s_mom[w][tid];
                    for (int g = 0; g < a.shard.n_shards; ++g) a.shard.mail[g]->mom[a.shard.rank][tid] = t;
                }
            }
            __threadfence_system();
            __syncthreads();
            if (tid < a.shard.n_shards)
                *(volatile unsigned long long*)&a.shard.mail[tid]->flag_mom[a.shard.rank] = a.shard.stamp;
            if (tid == 0) *a.ticket = 0u;
        }
    }
}

// ==============================================================================================
// finish step: post-resample moments, buffer flip, shift refresh
// ==============================================================================================
constexpr int kFinishThreads = 1024;
__global__ void __launch_bounds__(kFinishThreads) k_finish_step(Ctrl* ctrl, FinishArgs f) {
    __shared__ double s_w[kFinishThreads / 32][16];
    __shared__ double s_m[16];
    finish_body(ctrl, f, s_w, s_m);
}

// sharded: wait until every rank has raised its moment flag for this step (= all cross-shard particle
// stores into this rank's buffers have landed), then combine the G moment records in rank order
__global__ void __launch_bounds__(kThreads) k_finish_step_sharded(Ctrl* ctrl, FinishArgs f, ShardCtx sc) {
    __shared__ double s_w[kWarps][16];
    __shared__ double s_m[16];
    Mailbox* own = sc.mail[sc.rank];
    mailbox_wait(own->flag_mom, sc.n_shards, sc.stamp, &own->error, 3);
    f.mom_parts = &own->mom[0][0];
    f.n_mom_parts = sc.n_shards;
    f.mom_stride = 16;
    finish_body(ctrl, f, s_w, s_m);
}

__global__ void __launch_bounds__(kThreads) k_reduce_moment_partials(const Ctrl* ctrl, const double* mom_parts,
                                                                     int n_parts, int nm, double* out) {
    __shared__ double s_tmp[kWarps];
    const bool live = ctrl->do_resample != 0;
    for (int k = 0; k < nm; ++k) {
        double acc = 0;
        if (live)
            for (int p = threadIdx.x; p < n_parts; p += kThreads) acc += mom_parts[(int64_t)p * nm + k];
        const double tot = block_sum_256(acc, s_tmp);
        if (threadIdx.x == 0) out[k] = tot;
    }
}

// ==============================================================================================
// generic (any nx) moments: one CTA per mean entry / per lower-triangular covariance entry.
// Reads ctrl: weighted by exp(logw - log_norm) when the step did not resample, uniform otherwise.
// ==============================================================================================
template <typename real>
__global__ void __launch_bounds__(kThreads) k_moments_generic(const Ctrl* ctrl, const real* x0, const real* x1,
                                                              const real* logw, int64_t n, int64_t n_global,
                                                              int64_t ld, int nx, double* stats, int post_flip) {
    // post_flip: 1 = called after finish_step flipped `cur` (use cur as is); 0 = before
    __shared__ double s_tmp[kWarps];
    const int cur = ctrl->cur ^ ((ctrl->do_resample && !post_flip) ? 1 : 0);
    const real* x = cur ? x1 : x0;
    const bool weighted = !ctrl->do_resample;
    const double log_norm = ctrl->log_norm;
    const int b = blockIdx.x;
    int d, d2;
    if (b < nx) { d = b; d2 = -1; }
    else {
        int t = b - nx; d = 0;
        while (t >= d + 1) { t -= d + 1; ++d; }
        d2 = t;
    }
    const double c1 = ctrl->shift[d < kMaxNx ? d : 0] * (d < kMaxNx);
    const double c2 = d2 >= 0 ? ctrl->shift[d2 < kMaxNx ? d2 : 0] * (d2 < kMaxNx) : 0.0;
    double acc = 0, accw = 0;
    for (int64_t i = threadIdx.x; i < n; i += kThreads) {
        double w = weighted ? exp((double)logw[i] - log_norm) : 1.0;
        double a = (double)x[d * ld + i] - c1;
        double v = d2 >= 0 ? a * ((double)x[d2 * ld + i] - c2) : a;
        acc = fma(w, v, acc);
        accw += w;
    }
    double tot = block_sum_256(acc, s_tmp);
    double totw = block_sum_256(accw, s_tmp);
    if (threadIdx.x == 0) {
        // raw shifted moments; k_moments_generic_fix turns them into mean / cov
        double* raw = stats + SMC_STATS_HEADER;
        if (d2 < 0) raw[d] = tot / totw;
        else raw[nx + d * nx + d2] = tot / totw;
    }
    (void)n_global;
}
__global__ void k_moments_generic_fix(const Ctrl* ctrl, int nx, double* stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double* mean = stats + SMC_STATS_HEADER;
    double* cov = mean + nx;
    for (int d = 0; d < nx; ++d)
        for (int d2 = 0; d2 <= d; ++d2) {
            double v = cov[d * nx + d2] - mean[d] * mean[d2];
            cov[d * nx + d2] = v;
            cov[d2 * nx + d] = v;
        }
    for (int d = 0; d < nx; ++d) mean[d] += (d < kMaxNx ? ctrl->shift[d] : 0.0);
}

// ==============================================================================================
// small utility kernels
// ==============================================================================================
__global__ void k_reset_ctrl(Ctrl* c, const double* shift, int nx) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        c->log_norm = 0.0; c->ess = 0.0; c->u = 0.0; c->cur = 0; c->uniform = 1; c->do_resample = 0;
        c->frac_bits = 32; c->lw_ref = 0.0; c->uq = 0ull; c->margin = 0ull;
        c->kshift = 0.f; c->rsum_valid = 0;      // (rsum_sel stays: the buffer it does NOT name is all zero, always)
        for (int d = 0; d < kMaxNx; ++d) c->shift[d] = (shift && d < nx) ? shift[d] : 0.0;
    }
}
__global__ void k_set_uniform(Ctrl* c, double u, int64_t n_global, int f32) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { c->u = u; ctrl_prepare_search(c, n_global, f32 != 0); }
}
__global__ void k_set_resample(Ctrl* c, double u, int64_t n_global, int f32) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        c->u = u; c->do_resample = 1; c->cur = 0; c->log_norm = 0.0;
        ctrl_prepare_search(c, n_global, f32 != 0);
    }
}

template <typename real>
__global__ void __launch_bounds__(kThreads) k_init_gaussian(Ctrl* ctrl, real* x0, real* x1, int64_t n, int64_t offset,
                                                            int64_t ld, int nx, const double* mean, const double* Lc,
                                                            uint64_t seed) {
    real* x = ctrl->cur ? x1 : x0;
    const int64_t quad = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t i0 = quad * 4;
    if (i0 >= n) return;
    // nx can be large here: draw dimension by dimension, accumulate the triangular product
    for (int d = 0; d < nx; ++d) {
        double acc[4] = {mean[d], mean[d], mean[d], mean[d]};
        for (int k = 0; k <= d; ++k) {
            double z4[4];
            uint32_t r[4];
            philox_quad(seed, (uint64_t)(offset + i0) >> 2, 0u, kStreamInit | (uint32_t(k) << 8), r);
            double z2[2];
            normals2_f64(r, z2); z4[0] = z2[0]; z4[1] = z2[1];
            philox_quad(seed, (uint64_t)(offset + i0) >> 2, 0u, kStreamInit | (uint32_t(k) << 8) | (1u << 16), r);
            normals2_f64(r, z2); z4[2] = z2[0]; z4[3] = z2[1];
            const double l = Lc[d * nx + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fma(l, z4[j], acc[j]);
        }
        real out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = (real)acc[j];
        store4<real>(x + d * ld + i0, out);
    }
}

template <typename real>
__global__ void __launch_bounds__(kThreads) k_bridge_predict(Ctrl* ctrl, real* x0, real* x1, const real* gx,
                                                             const real* noise, const double* Lq, double scale,
                                                             int64_t n, int64_t ld, int nx, int target_other) {
    real* x = (ctrl->cur ^ target_other) ? x1 : x0;
    if (!gx) gx = x;   // in-place jitter: each thread reads then writes its own elements only
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    for (int d = nx - 1; d >= 0; --d) {
        // (noise @ Lq.T)[i, d] = sum_k noise[i, k] Lq[d, k]   (particle_filter.py:236)
        double acc = 0;
        for (int k = 0; k <= d; ++k) acc = fma((double)noise[k * ld + i], Lq[d * nx + k] * scale, acc);
        x[d * ld + i] = (real)((double)gx[d * ld + i] + acc);
    }
}

// log-likelihood from host-evaluated h(x): grid-stride over particles, online (max, sum, sum^2)
template <typename real>
__global__ void __launch_bounds__(kThreads) k_bridge_loglik(Ctrl* ctrl, real* logw, const real* zpred, const double* z,
                                                            const double* LRinv, const real* wprev, int64_t n,
                                                            int64_t n_global, int64_t ld, int nz, double* parts) {
    __shared__ double s_tmp[kWarps];
    const double lw_uniform = log(1.0 / (double)n_global + 1e-300);
    double m = -INFINITY, s = 0, s2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        double quad = 0;
        for (int r = 0; r < nz; ++r) {        // y = LR^-1 (z - zpred), particle_filter.py:258-260
            double acc = 0;
            for (int k = 0; k <= r; ++k) acc = fma(LRinv[r * nz + k], z[k] - (double)zpred[k * ld + i], acc);
            quad = fma(acc, acc, quad);
        }
        const double prev = wprev ? log((double)wprev[i] + 1e-300) : lw_uniform;   // :261
        double v = prev - 0.5 * quad;
        logw[i] = (v == v) ? (real)v : real(-INFINITY);   // NaN counts as zero weight (stored as -inf, see K1)
        if (!(v == v)) continue;
        if (v > m) {
            const double sc = exp(m - v);     // m = -inf -> 0
            s *= sc; s2 *= sc * sc; m = v;
        }
        if (m != -INFINITY) {
            const double e = exp(v - m);
            s += e; s2 = fma(e, e, s2);
        }
    }
    const double bmax = block_max_256(m, s_tmp);
    const double sc = (m == -INFINITY) ? 0.0 : exp(m - bmax);
    const double bs = block_sum_256(s * sc, s_tmp);
    const double bs2 = block_sum_256(s2 * sc * sc, s_tmp);
    if (threadIdx.x == 0) {
        double* out = parts + (int64_t)blockIdx.x * 3;
        out[0] = bmax; out[1] = bs; out[2] = bs2;
    }
    (void)ctrl;
}

template <typename real>
__global__ void __launch_bounds__(kThreads) k_export_weights(const Ctrl* ctrl, const real* logw, int64_t n,
                                                             int64_t n_global, double* w) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    w[i] = ctrl->uniform ? 1.0 / (double)n_global : exp((double)logw[i] - ctrl->log_norm);
}

template <typename real>
__global__ void __launch_bounds__(kThreads) k_gather(const real* src, real* dst, const int32_t* idx, int64_t n,
                                                     int64_t ld, int nx) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const int64_t j = idx[i];
    for (int d = 0; d < nx; ++d) dst[d * ld + i] = src[d * ld + j];
}

// gather for the non-fused (nx > kFusedNxMax) resampling path; honours ctrl->do_resample / cur
template <typename real>
__global__ void __launch_bounds__(kThreads) k_gather_step(const Ctrl* ctrl, const real* x0, const real* x1,
                                                          real* d0, real* d1, const int32_t* idx, int64_t n,
                                                          int64_t ld, int nx) {
    if (!ctrl->do_resample) return;
    const real* src = ctrl->cur ? x1 : x0;
    real* dst = ctrl->cur ? d0 : d1;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const int64_t j = idx[i];
    for (int d = 0; d < nx; ++d) dst[d * ld + i] = src[d * ld + j];
}

}  // namespace smc

#include "fused_step.cuh"
#include "big_state.cuh"

// ================================================================================================
// C-ABI
// ================================================================================================
using namespace smc;

namespace {

struct Launch {
    const smc_sir_desc* d;
    WsLayout L;
    unsigned char* ws;
    cudaStream_t st;
    Ctrl* ctrl() const { return reinterpret_cast<Ctrl*>(ws + L.ctrl); }
    double* slot(int k) const { return reinterpret_cast<double*>(ws + L.params) + (size_t)k * kParamSlotDoubles; }
    double* k1_parts() const { return reinterpret_cast<double*>(ws + L.k1_parts); }
    unsigned long long* agg() const { return reinterpret_cast<unsigned long long*>(ws + L.tile_agg); }
    unsigned int* ticket() const { return reinterpret_cast<unsigned int*>(ws + L.ticket); }
    double* mom_parts() const { return reinterpret_cast<double*>(ws + L.mom_parts); }
    double* big_parts() const { return reinterpret_cast<double*>(ws + L.big_parts); }
    unsigned int* board() const { return reinterpret_cast<unsigned int*>(ws + L.board); }
};

// Launch with programmatic stream serialization (see pdl_wait / pdl_trigger).  Measured on B200 at
// N = 2^26: 0.432 ms per step with the attribute against 0.421 ms without (the early-scheduled CTAs
// of the next kernel take residency from the persistent kernel that is still draining), so it is off.
#ifndef SMC_USE_PDL
#define SMC_USE_PDL 0
#endif
static unsigned long long g_launch_count = 0;     // (host threads of one process share the engine's stream in practice)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
    __atomic_add_fetch(&g_launch_count, 1ull, __ATOMIC_RELAXED);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = SMC_USE_PDL;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

int check_desc(const smc_sir_desc* d, Launch* out, void* stream) {
    SMC_REQUIRE(d != nullptr, "smc: null descriptor");
    SMC_REQUIRE(d->n > 0 && d->n_global >= d->n, "smc: need 0 < n <= n_global (n=%lld n_global=%lld)",
                (long long)d->n, (long long)d->n_global);
    SMC_REQUIRE(d->n_global <= (1ll << 30), "smc: n_global=%lld exceeds 2^30", (long long)d->n_global);
    SMC_REQUIRE(d->dtype == SMC_F32 || d->dtype == SMC_F64, "smc: bad dtype %d", d->dtype);
    SMC_REQUIRE(d->ld % 4 == 0 && d->ld >= ((d->n + 3) / 4) * 4,
                "smc: ld=%lld must be a multiple of 4 and >= round_up(n,4)", (long long)d->ld);
    SMC_REQUIRE(d->offset % 4 == 0 && d->offset >= 0, "smc: shard offset must be a non-negative multiple of 4");
    SMC_REQUIRE(d->nx >= 1, "smc: nx must be >= 1");
    SMC_REQUIRE(d->ws_dev != nullptr, "smc: workspace missing");
    out->d = d;
    out->L = ws_layout(d->n, d->nx, d->dtype);
    SMC_REQUIRE(d->ws_bytes >= out->L.total, "smc: workspace too small (%llu < %llu)",
                (unsigned long long)d->ws_bytes, (unsigned long long)out->L.total);
    out->ws = static_cast<unsigned char*>(d->ws_dev);
    out->st = static_cast<cudaStream_t>(stream);
    return 0;
}

// moments accumulated inside K1 / K3 (true) or by the any-nx kernel (false)
inline bool inline_moments(const smc_sir_desc* d) {
    return d->nx <= kFusedNxMax && !(d->flags & SMC_FLAG_GENERIC_MOMENTS);
}

// Small host arrays (matrices, observation) are staged into a parameter slot of the workspace.
// cudaMemcpyAsync from pageable memory copies the source before returning, so the host buffer may
// be reused at once; stream order protects the slot between consecutive calls.
int stage(const Launch& l, int slot, const double* host, int count, double** dev) {
    SMC_REQUIRE(count <= kParamSlotDoubles, "smc: parameter array too large (%d doubles)", count);
    *dev = l.slot(slot);
    SMC_CHECK_CUDA(cudaMemcpyAsync(*dev, host, sizeof(double) * count, cudaMemcpyHostToDevice, l.st));
    return 0;
}

inline int sm_count();
inline int sm_count_early() { return sm_count(); }
// whether K1 may deliver the range totals of a bootstrap step (fp32, enough tiles for the 148 x 12 pre-pass ranges,
// which every K1 CTA then owns a whole number of)
inline bool k1_rsum_ok(const Launch& l) {
    const smc_sir_desc* d = l.d;
    const int64_t ranges = (int64_t)sm_count_early() * 12;
    return d->dtype == SMC_F32 && l.L.n_tiles >= ranges && ranges <= kRsumStride && ranges % l.L.n_k1_blocks == 0 &&
           d->nx <= kFusedNxMax && !getenv("SMC_NO_K1_RSUM");
}

template <typename real, class Model>
int launch_k1(const Launch& l, const Model& m, int mode, const double* z_dev, const void* noise_dev, uint64_t seed,
              int64_t t, const K2Args* fin, const ShardCtx* shard) {
    const smc_sir_desc* d = l.d;
    K1Args<real> a;
    a.ctrl = l.ctrl();
    a.x0 = static_cast<real*>(d->x0_dev);
    a.x1 = static_cast<real*>(d->x1_dev);
    a.logw = static_cast<real*>(d->logw_dev);
    a.noise = static_cast<const real*>(noise_dev);
    a.z = z_dev;
    a.parts = l.k1_parts();
    a.n = d->n; a.n_global = d->n_global; a.offset = d->offset; a.ld = d->ld;
    a.seed = seed; a.t = (uint32_t)t; a.mode = mode & 3;
    a.keys = philox_expand(seed);
    a.fuse_finalize = 0; a.ticket = l.ticket();
    memset(&a.fin, 0, sizeof(a.fin));
    memset(&a.shard, 0, sizeof(a.shard));
    if (shard && (mode & SMC_MODE_WEIGHT)) a.shard = *shard;
    a.rsum = nullptr; a.n_ranges = 0; a.kshift_bound = 0.f;
    if (fin && (mode & SMC_MODE_WEIGHT)) {
        a.fuse_finalize = 1;
        a.fin = *fin;
        a.fin.ctrl = l.ctrl(); a.fin.parts = l.k1_parts(); a.fin.n_parts = (int)l.L.n_k1_blocks;
    }
    // fp32, large populations, finalize fused into K1 or sharded step: K1 also delivers the range totals of the
    // resampler's CDF when the step turns out to be a bootstrap step (uniform prior weights; decided on the device)
    if ((fin || shard) && (mode & SMC_MODE_WEIGHT) && k1_rsum_ok(l)) {
        double lw_bound;
        a.rsum = l.agg(); a.n_ranges = (int32_t)((int64_t)sm_count_early() * 12);
        quant_bound_for(d->n_global, &lw_bound, &a.kshift_bound);
    }
    // mode bit 4 (internal): the caller knows the step resamples -> skip the pre-resample moments
    if (mode & 4)
        SMC_CHECK_CUDA(launch_pdl(k1_propagate_weight<real, Model, false>, (unsigned)l.L.n_k1_blocks, kThreads, 0, l.st, a, m));
    else
        SMC_CHECK_CUDA(launch_pdl(k1_propagate_weight<real, Model, true>, (unsigned)l.L.n_k1_blocks, kThreads, 0, l.st, a, m));
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <typename real> int dispatch_k1(const Launch& l, int model_id, const double* p, int np, int mode,
                                         const double* z_dev, const void* noise_dev, uint64_t seed, int64_t t,
                                         const K2Args* fin, const ShardCtx* shard = nullptr) {
    const smc_sir_desc* d = l.d;
    if (model_id == SMC_MODEL_STOCH_VOL) {
        // params: alpha, lq (= chol(Q)), beta, lr (= chol(R + 1e-12))
        SMC_REQUIRE(np == 4 && d->nx == 1 && d->nz == 1, "smc: STOCH_VOL needs 4 params and nx = nz = 1");
        StochVolModel<real> m;
        m.alpha = (real)p[0]; m.lq = (real)p[1]; m.beta = (real)p[2]; m.rinv = (real)(1.0 / p[3]);
        return launch_k1<real>(l, m, mode, z_dev, noise_dev, seed, t, fin, shard);
    }
    if (model_id == SMC_MODEL_LINEAR_GAUSSIAN) {
        // params: A (nx*nx), Lq (nx*nx lower), H (nz*nx), LRinv (nz*nz lower), all row-major
        const int nx = d->nx, nz = d->nz;
        SMC_REQUIRE(np == nx * nx * 2 + nz * nx + nz * nz, "smc: LINEAR_GAUSSIAN parameter count mismatch");
#define SMC_LG_CASE(NX_, NZ_)                                                                     \
        if (nx == NX_ && nz == NZ_) {                                                             \
            LinGaussModel<real, NX_, NZ_> m;                                                      \
            const double* q = p;                                                                  \
            for (int i = 0; i < NX_ * NX_; ++i) m.A[i] = (real)*q++;                              \
            for (int i = 0; i < NX_ * NX_; ++i) m.Lq[i] = (real)*q++;                             \
            for (int i = 0; i < NZ_ * NX_; ++i) m.H[i] = (real)*q++;                              \
            for (int i = 0; i < NZ_ * NZ_; ++i) m.LRinv[i] = (real)*q++;                          \
            return launch_k1<real>(l, m, mode, z_dev, noise_dev, seed, t, fin, shard);                     \
        }
        SMC_LG_CASE(1, 1) SMC_LG_CASE(2, 1) SMC_LG_CASE(2, 2) SMC_LG_CASE(4, 1) SMC_LG_CASE(4, 2) SMC_LG_CASE(4, 4)
#undef SMC_LG_CASE
        SMC_REQUIRE(false, "smc: LINEAR_GAUSSIAN supports (nx,nz) in {(1,1),(2,1),(2,2),(4,1),(4,2),(4,4)}, got (%d,%d)",
                    nx, nz);
    }
    if (model_id == SMC_MODEL_LORENZ96) {
        // params: F, dt, n_sub, obs_stride, nz, then diag chol(Q) (nx), then diag chol(R + 1e-12 I) (nz)
        const int nx = d->nx, nz = d->nz;
        SMC_REQUIRE(np == 5 + nx + nz && nx <= kL96MaxNx, "smc: LORENZ96 parameter count mismatch");
        L96Params<real> lp;
        lp.F = (real)p[0]; lp.dt = (real)p[1]; lp.n_sub = (int)p[2]; lp.obs_stride = (int)p[3]; lp.nz = (int)p[4];
        SMC_REQUIRE(lp.nz == nz && lp.n_sub >= 0 && lp.obs_stride >= 1 && (nz - 1) * lp.obs_stride < nx,
                    "smc: LORENZ96 observation layout does not fit nx = %d", nx);
        for (int i = 0; i < kL96MaxNx; ++i) { lp.lq[i] = 0; lp.rinv[i] = 0; }
        for (int i = 0; i < nx; ++i) lp.lq[i] = (real)p[5 + i];
        for (int i = 0; i < nz; ++i) lp.rinv[i] = (real)(1.0 / p[5 + nx + i]);
        K1Args<real> a;
        memset(&a, 0, sizeof(a));
        a.ctrl = l.ctrl();
        a.x0 = static_cast<real*>(d->x0_dev); a.x1 = static_cast<real*>(d->x1_dev);
        a.logw = static_cast<real*>(d->logw_dev);
        a.noise = static_cast<const real*>(noise_dev);
        a.z = z_dev; a.parts = l.k1_parts();
        a.n = d->n; a.n_global = d->n_global; a.offset = d->offset; a.ld = d->ld;
        a.seed = seed; a.t = (uint32_t)t; a.mode = mode & 3;
        a.keys = philox_expand(seed);
        const unsigned grid = (unsigned)l.L.n_k1_blocks;
        if (nx == 40) k1_lorenz96<real, 40><<<grid, kL96Threads, 0, l.st>>>(a, lp);
        else if (nx == 8) k1_lorenz96<real, 8><<<grid, kL96Threads, 0, l.st>>>(a, lp);
        else SMC_REQUIRE(false, "smc: LORENZ96 is compiled for nx in {8, 40}, got %d", nx);
        SMC_CHECK_CUDA(cudaGetLastError());
        if (fin && (mode & SMC_MODE_WEIGHT)) {   // finalize as its own launch (see big_state.cuh)
            K2Args f2 = *fin;
            f2.ctrl = l.ctrl(); f2.parts = l.k1_parts(); f2.n_parts = (int)grid;
            k2_finalize<<<1, kThreads, 0, l.st>>>(f2);
            SMC_CHECK_CUDA(cudaGetLastError());
        }
        return 0;
    }
    SMC_REQUIRE(false, "smc: unknown model id %d", model_id);
}

inline int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// persistent resampler grid (also the number of post-resample moment records it writes)
inline int k3_occupancy(int dtype, int nxf) {
    if (dtype == SMC_F32) return nxf <= 1 ? 4 : (nxf == 2 ? 3 : 2);
    return nxf <= 1 ? 3 : 2;
}
inline int64_t k3_grid(int64_t n_tiles, int dtype, int nxf, bool fused = false) {
    int occ = k3_occupancy(dtype, nxf);
    if (fused) occ = occ > 2 ? occ - SMC_FUSED_OCC_DROP : 2;
    const int64_t g = (int64_t)sm_count() * occ;
    return n_tiles < g ? n_tiles : g;
}
// number of pre-pass ranges: a common multiple of all resampler grids when the problem is large enough
// (then plain and fused resamplers share a pre-pass), else one range per CTA of the plain grid
inline int64_t k3_ranges(int64_t n_tiles, int dtype, int nxf) {
    const int64_t big = (int64_t)sm_count() * 12;
    return n_tiles >= big ? big : k3_grid(n_tiles, dtype, nxf);
}

template <typename real, int NXF, bool LIN, class FM = NoFuse>
int launch_k3_t(const Launch& l, K3Args<real>& a, int64_t n_tiles, FM fm = FM()) {
    using G = K3Geom<real>;

    constexpr int PADROW = G::THREADS + 2 * (G::THREADS / G::IPT);
    constexpr size_t ids_bytes = (size_t)(8 + (G::IPT + 2) * PADROW + 8) * 2;   // guard + window + tail, u16 slots
    static_assert(ids_bytes % 16 == 0, "marker window must be a whole number of 16-byte units");
    const size_t smem = sizeof(real) * G::THREADS * G::IPT * (1 + NXF) + ids_bytes;
    auto kern = k3_resample<real, NXF, LIN, FM>;
    static bool attr_set = false;   // one flag per template instantiation
    if (!attr_set) {
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    a.n_tiles = n_tiles;
    constexpr bool kFused = FM::NX > 0;
    const int dt = sizeof(real) == 4 ? SMC_F32 : SMC_F64;
    const int64_t grid = k3_grid(n_tiles, dt, NXF, kFused);
    const int64_t ranges = k3_ranges(n_tiles, dt, NXF);
    SMC_REQUIRE(ranges % grid == 0, "smc: resampler grid %lld does not divide the %lld pre-pass ranges (population too small "
                "for this variant)", (long long)grid, (long long)ranges);
    a.ranges_per_cta = (int32_t)(ranges / grid);
    if (a.fuse_finish) a.fin.n_mom_parts = (int)grid;
    SMC_CHECK_CUDA(launch_pdl(kern, (unsigned)grid, G::THREADS, smem, l.st, a, fm));
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <typename real>
int launch_prepass(const Launch& l, bool linear_w, const void* wsrc, int nxf, int force, unsigned long long* total_out,
                   const ShardCtx* shard = nullptr) {
    const smc_sir_desc* d = l.d;
    unsigned long long* rtot = l.agg();
    const unsigned grid = (unsigned)k3_ranges(l.L.n_tiles, d->dtype, nxf);   // refines every resampler partition
    const Ctrl* ctrl = l.ctrl();
    const real* w = static_cast<const real*>(wsrc);
    ShardCtx sc;
    memset(&sc, 0, sizeof(sc));
    if (shard) sc = *shard;
    unsigned int* ticket = l.ticket() + 2;
    if (linear_w)
        SMC_CHECK_CUDA(launch_pdl(k3a_range_totals<real, true>, grid, K3Geom<real>::THREADS, 0, l.st, ctrl, w, d->n,
                                  d->n_global, l.L.n_tiles, rtot, force, sc, ticket));
    else
        SMC_CHECK_CUDA(launch_pdl(k3a_range_totals<real, false>, grid, K3Geom<real>::THREADS, 0, l.st, ctrl, w, d->n,
                                  d->n_global, l.L.n_tiles, rtot, force, sc, ticket));
    SMC_CHECK_CUDA(cudaGetLastError());
    if (total_out) {
        k3a_sum_ranges<<<1, kThreads, 0, l.st>>>(l.ctrl(), rtot, (int)grid, total_out, force);
        SMC_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

template <typename real>
void fill_k3_args(const Launch& l, K3Args<real>& a, const void* wsrc, int nxf, unsigned long long cdf_offset, int force,
                  void* const* peer, int n_shards, bool want_moments, const unsigned long long* totals, int rank,
                  const FinishArgs* fin, const ShardCtx* shard) {
    const smc_sir_desc* d = l.d;
    a.ctrl = l.ctrl();
    a.wsrc = static_cast<const real*>(wsrc);
    a.x0 = static_cast<const real*>(d->x0_dev);
    a.x1 = static_cast<const real*>(d->x1_dev);
    a.dst0 = static_cast<real*>(d->x0_dev);
    a.dst1 = static_cast<real*>(d->x1_dev);
    a.peer0 = nullptr; a.peer1 = nullptr;
    if (peer) {   // device array: [n_shards pointers to buffer 0][n_shards pointers to buffer 1]
        a.peer0 = reinterpret_cast<real* const*>(peer);
        a.peer1 = reinterpret_cast<real* const*>(peer) + n_shards;
    }
    a.ancestors = ((d->flags & SMC_FLAG_WRITE_ANCESTORS) || nxf == 0) ? d->ancestors_dev : nullptr;
    a.rtot = l.agg();
    a.mom_parts = want_moments ? l.mom_parts() : nullptr;
    a.cdf_offset = cdf_offset;
    a.totals = totals; a.rank = rank; a.n_totals = n_shards;
    a.n = d->n; a.n_global = d->n_global; a.offset = d->offset; a.ld = d->ld;
    a.shard_n = n_shards > 0 ? d->n_global / n_shards : d->n_global;
    a.force = force;
    a.fuse_finish = 0; a.ticket = l.ticket() + 1;
    memset(&a.fin, 0, sizeof(a.fin));
    if (fin) {
        a.fuse_finish = 1;
        a.fin = *fin;
        a.fin.mom_parts = l.mom_parts();
    }
    memset(&a.shard, 0, sizeof(a.shard));
    if (shard) a.shard = *shard;
    memset(&a.fk, 0, sizeof(a.fk));
    a.prologue_finalize = 0;
    memset(&a.pfin, 0, sizeof(a.pfin));
    a.gbar = l.ticket() + 3;
}

template <typename real>
int launch_k3(const Launch& l, bool linear_w, const void* wsrc, int nxf, unsigned long long cdf_offset, int force,
              void* const* peer, int n_shards, bool want_moments, const unsigned long long* totals = nullptr,
              int rank = 0, bool run_prepass = true, const FinishArgs* fin = nullptr, const ShardCtx* shard = nullptr,
              const K2Args* prologue_fin = nullptr) {
    if (run_prepass) {
        if (int rc = launch_prepass<real>(l, linear_w, wsrc, nxf, force, nullptr)) return rc;
    }
    K3Args<real> a;
    fill_k3_args<real>(l, a, wsrc, nxf, cdf_offset, force, peer, n_shards, want_moments, totals, rank, fin, shard);
    if (prologue_fin) { a.prologue_finalize = 1; a.pfin = *prologue_fin; }
    const int64_t nt = l.L.n_tiles;
    if (linear_w) {
        SMC_REQUIRE(nxf == 0, "smc: linear-weight resampling is ancestors-only");
        return launch_k3_t<real, 0, true>(l, a, nt);
    }
    switch (nxf) {
        case 0: return launch_k3_t<real, 0, false>(l, a, nt);
        case 1: return launch_k3_t<real, 1, false>(l, a, nt);
        case 2: return launch_k3_t<real, 2, false>(l, a, nt);
        case 3: return launch_k3_t<real, 3, false>(l, a, nt);
        case 4: return launch_k3_t<real, 4, false>(l, a, nt);
    }
    SMC_REQUIRE(false, "smc: bad fused nx %d", nxf);
}

template <typename real, int BS>
int launch_syrk_t(const Launch& l, double* stats_row_dev) {
    const smc_sir_desc* d = l.d;
    constexpr int NP = 8 * BS;
    const size_t stage = sizeof(real) * (2 * (size_t)kSyrkTile * (NP + 1) + kSyrkTile);
    const size_t fold = sizeof(double) * 4 * (1 + NP + (size_t)NP * NP);
    const size_t smem = stage > fold ? stage : fold;
    auto kern = k_moments_syrk<real, BS>;
    static bool attr_set = false;
    if (!attr_set) {
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int64_t grid = cdiv64(d->n, kSyrkTile);
    if (grid > kSyrkMaxBlocks) grid = kSyrkMaxBlocks;
    kern<<<(unsigned)grid, 256, smem, l.st>>>(l.ctrl(), (const real*)d->x0_dev, (const real*)d->x1_dev,
                                              (const real*)d->logw_dev, d->n, d->ld, d->nx, 0, l.big_parts());
    SMC_CHECK_CUDA(cudaGetLastError());
    const size_t fsm = sizeof(double) * (1 + d->nx + (size_t)d->nx * d->nx);
    k_moments_syrk_finish<<<d->nx > 8 ? 16 : 1, 256, fsm, l.st>>>(l.ctrl(), l.big_parts(), (int)grid, d->nx, stats_row_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template <typename real> int launch_syrk(const Launch& l, double* stats_row_dev) {
    return l.d->nx <= 40 ? launch_syrk_t<real, 5>(l, stats_row_dev) : launch_syrk_t<real, 8>(l, stats_row_dev);
}

// ---- fused resample(t-1) -> propagate + weigh(t) (run() pipeline of bootstrap filters, fp32) ------------
// K5 (fused_step.cuh): one launch per step plus the conditional range pre-pass (a no-op unless the last CTA
// of K5 found the a-priori bound too loose for this step's weights).  SMC_FUSED_LEGACY=1 in the environment
// selects round 1's pipeline (k3_resample with the functor in its copy-out + an unconditional pre-pass) for
// A/B measurements.
inline bool fused_legacy() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SMC_FUSED_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
// Shares of a slot's ranges for the CTAs of one SM, by their rank on the SM (K5Args::share).  Environment override
// for measurements: SMC_K5_SHARES / SMC_K5S_SHARES = "5,4,3"; "0" = equal shares, static block -> ranges mapping.
inline void k5_set_shares(K5Args& a, const Launch& l, int64_t grid, int64_t ranges, const char* env_name, const int* dflt) {
    a.slots = sm_count();
    a.n_waves = (int32_t)(grid / a.slots);
    a.board = nullptr;
    const int per_slot = (int)(ranges / a.slots);
    int sh[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sum = 0;
    bool ok = false;
    const char* e = getenv(env_name);
    if (e && e[0] == '0' && e[1] == 0) return;
    if (e) {
        int k = 0;
        for (const char* q = e; *q && k < 8; ++k) { sh[k] = atoi(q); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
        for (int i = 0; i < a.n_waves; ++i) sum += sh[i];
        ok = k == a.n_waves && sum == per_slot;
        for (int i = 0; i < a.n_waves; ++i) ok = ok && sh[i] >= 1;
    }
    if (!ok && dflt) {
        sum = 0;
        for (int i = 0; i < a.n_waves; ++i) { sh[i] = dflt[i]; sum += sh[i]; }
        ok = sum == per_slot;
    }
    if (!ok || grid != (int64_t)a.slots * a.n_waves || grid > kBoardPieces || a.n_waves > 6) return;
    for (int i = 0; i < 8; ++i) a.share[i] = (int8_t)sh[i];
    a.board = l.board();
    const char* w = getenv("SMC_K5_BOARD_WAIT_NS");
    a.board_wait_ns = w ? (uint32_t)atoi(w) : 4000u;
}
inline int64_t k5_grid(int nx) { return (int64_t)sm_count() * (nx <= 1 ? SMC_K5_OCC : 2); }

// K5 launch: `fin2` != nullptr for the fused instantiations (functor of step t + its finalize), nullptr for the
// plain fp32 resampler (CopyModel)
template <int NX, class Model>
int launch_k5(const Launch& l, const Model& m, const void* lw_in, void* lw_out, const double* z_dev, uint64_t seed,
              int64_t t, const FinishArgs& f, const K2Args* fin2) {
    const smc_sir_desc* d = l.d;
    K5Args a;
    memset(&a, 0, sizeof(a));
    a.ctrl = l.ctrl();
    a.lw_in = static_cast<const float*>(lw_in);
    a.lw_out = static_cast<float*>(lw_out);
    a.x0 = static_cast<const float*>(d->x0_dev); a.x1 = static_cast<const float*>(d->x1_dev);
    a.dst0 = static_cast<float*>(d->x0_dev); a.dst1 = static_cast<float*>(d->x1_dev);
    a.rsum = l.agg();
    a.mom_parts = (d->flags & SMC_FLAG_NO_MOMENTS) ? nullptr : l.mom_parts();
    a.parts = l.k1_parts();
    a.z = z_dev;
    a.keys = philox_expand(seed);
    a.t = (uint32_t)t;
    a.n = d->n; a.ld = d->ld; a.n_tiles = l.L.n_tiles;
    const int64_t grid = k5_grid(NX);
    const int64_t ranges = k3_ranges(l.L.n_tiles, d->dtype, NX);
    SMC_REQUIRE(ranges % grid == 0 && l.L.n_tiles >= ranges, "smc: K5 needs at least %lld tiles", (long long)ranges);
    a.n_ranges = (int32_t)ranges; a.ranges_per_cta = (int32_t)(ranges / grid);
    a.ticket = l.ticket() + 1;
    // four CTAs per SM, 12 ranges per SM: measured 0.3462 ms against 0.3625 ms with equal shares at N = 2^26 (and 5,3,2,2:
    // 0.3647, 4,4,3,1: 0.3530, 4,4,2,2: 0.3525; profiles/r2_k5_shares.txt); short launches keep the static mapping
    static const int kShares4[4] = {4, 3, 3, 2};
    const bool long_launch = l.L.n_tiles >= 16 * grid;
    k5_set_shares(a, l, grid, ranges, "SMC_K5_SHARES", (NX == 1 && SMC_K5_OCC == 4 && long_launch) ? kShares4 : nullptr);
    a.fin = f; a.fin.mom_parts = l.mom_parts(); a.fin.n_mom_parts = (int)grid;
    if (fin2) { a.fin2 = *fin2; a.kshift_bound = fin2->kshift_bound; }
    auto kern = k5_resample_propagate<NX, Model>;
    const size_t smem = k5_smem_bytes<NX>();
    static bool attr_set = false;   // one flag per template instantiation
    if (!attr_set) {
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    SMC_CHECK_CUDA(launch_pdl(kern, (unsigned)grid, 256u, smem, l.st, a, m));
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}
// ---- sharded fused step: K5 on one shard, one launch and one exchange per step ---------------------------
// flush of a sharded fused run: finalize of the last step and finish of the one before it from the slot the last
// launch published (what the next launch's prologue would have done), into the GLOBAL control block
__global__ void __launch_bounds__(kThreads) k5s_flush(Ctrl* ctrl, K2Args pfin, FinishArgs pfinish, ShardCtx sc,
                                                      unsigned long long* rsum_area) {
    __shared__ double s_pout[32];
    __shared__ double s_m[16];
    Mailbox* own = sc.mail[sc.rank];
    const unsigned long long prev = sc.stamp - 1ull;
    const FusedSlot* slot = &own->fs[prev & 1ull];
    mailbox_wait(slot->flag, sc.n_shards, prev, &own->error, 4);
    const int warp = threadIdx.x >> 5;
    pfin.ctrl = ctrl; pfin.parts = &slot->rec[0][0]; pfin.rec_stride = 4; pfin.n_parts = sc.n_shards;
    if (warp == 0) finalize_warp(pfin, s_pout);
    if (warp == 1 && pfinish.stats && !pfinish.no_moments)
        finish_warp(ctrl, pfinish, &slot->mom[0][0], sc.n_shards, 16, pfinish.stats, s_m);
    // every rank has finished its last fused launch: nobody accumulates into this rank's range totals any more
    for (int i = threadIdx.x; i < 2 * kRsumStride + kMaxShards; i += kThreads) rsum_area[i] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) ctrl->rsum_valid = 0;      // the resampling that follows takes its range totals from a pre-pass
}

inline int64_t k5s_grid() { return (int64_t)sm_count() * SMC_K5S_OCC; }
inline bool k5_sharded_ok(const Launch& l) {
    const smc_sir_desc* d = l.d;
    if (d->dtype != SMC_F32 || d->nx != 1 || d->n % 4096 != 0) return false;
    if (d->flags & (SMC_FLAG_WRITE_ANCESTORS | SMC_FLAG_POST_MOMENTS_GENERIC | SMC_FLAG_GENERIC_MOMENTS)) return false;
    const int64_t ranges = k3_ranges(l.L.n_tiles, d->dtype, d->nx);
    return l.L.n_tiles >= ranges && ranges % k5s_grid() == 0 && ranges <= kRsumStride;
}

template <class Model>
int launch_k5_sharded(const Launch& l, const Model& m, const smc_shard_ctx* ctx, int lw_out_sel, int first,
                      const double* z_dev, double resample_thresh, uint64_t seed, int64_t t, double* stats_prev2_dev,
                      double* stats_prev_dev) {
    constexpr int NX = Model::NX;
    const smc_sir_desc* d = l.d;
    const int G = ctx->n_shards;
    K5Args a;
    memset(&a, 0, sizeof(a));
    a.ctrl = l.ctrl();
    a.lw_in = static_cast<const float*>(d->logw_dev);
    a.x0 = static_cast<const float*>(d->x0_dev); a.x1 = static_cast<const float*>(d->x1_dev);
    a.dst0 = static_cast<float*>(d->x0_dev); a.dst1 = static_cast<float*>(d->x1_dev);
    a.rsum = l.agg();
    a.mom_parts = (d->flags & SMC_FLAG_NO_MOMENTS) ? nullptr : l.mom_parts();
    a.parts = l.k1_parts();
    a.z = z_dev;
    a.keys = philox_expand(seed);
    a.t = (uint32_t)t;
    a.n = d->n; a.ld = d->ld; a.n_tiles = l.L.n_tiles;
    const int64_t grid = k5s_grid();
    const int64_t ranges = k3_ranges(l.L.n_tiles, d->dtype, NX);
    a.n_ranges = (int32_t)ranges; a.ranges_per_cta = (int32_t)(ranges / grid);
    a.ticket = l.ticket() + 1;
    // three CTAs per SM: 5,4,3 measured 365.8 against 372.9 us per step on a 2^26 shard, no gain at 2^25 and a loss on
    // short launches (the claim costs ~3 us in the prologue): only the largest shards take the board
    static const int kShares3[3] = {5, 4, 3};
    k5_set_shares(a, l, grid, ranges, "SMC_K5S_SHARES",
                  (SMC_K5S_OCC == 3 && l.L.n_tiles >= 32 * grid) ? kShares3 : nullptr);
    quant_bound_for(d->n_global, &a.sh.pfin.lw_bound, &a.sh.pfin.kshift_bound);
    a.kshift_bound = a.sh.pfin.kshift_bound;
    K5Shard& sh = a.sh;
    sh.ctx.mail = reinterpret_cast<Mailbox* const*>(ctx->mail_dev);
    sh.ctx.rank = ctx->rank; sh.ctx.n_shards = G; sh.ctx.stamp = ctx->stamp;
    sh.n_global = d->n_global; sh.offset = d->offset; sh.shard_n = d->n;
    sh.first = first;
    float* const* px = reinterpret_cast<float* const*>(ctx->peer_x_dev);
    void* const* aux = ctx->peer_aux_dev;
    sh.peer_x0 = px; sh.peer_x1 = px + G;
    sh.peer_lw_out = reinterpret_cast<float* const*>(aux + (size_t)lw_out_sel * G);
    sh.peer_rsum = reinterpret_cast<unsigned long long* const*>(aux + 2 * (size_t)G);
    sh.own = static_cast<Mailbox*>(ctx->mailbox_dev);
    sh.own_rsum = static_cast<unsigned long long*>(ctx->rsum_area_dev);
    sh.rsel_in = (int32_t)((ctx->stamp - 1) & 1); sh.rsel_out = (int32_t)(ctx->stamp & 1);
    sh.gbar = l.ticket() + 3;
    sh.pfin.stats = stats_prev_dev; sh.pfin.n_global = d->n_global; sh.pfin.nx = d->nx;
    sh.pfin.inline_moments = first ? 1 : 0; sh.pfin.f32 = 1;
    sh.pfin.thresh = resample_thresh; sh.pfin.u_host = -1.0; sh.pfin.seed = seed; sh.pfin.t = (uint32_t)(t - 1);
    sh.pfin.producer_rsum = first ? (k1_rsum_ok(l) ? 1 : 0) : 1;
    sh.pfinish.stats = first ? nullptr : stats_prev2_dev; sh.pfinish.nx = d->nx; sh.pfinish.inline_moments = 1;
    sh.pfinish.no_moments = (d->flags & SMC_FLAG_NO_MOMENTS) ? 1 : 0; sh.pfinish.n_global = d->n_global;
    auto kern = k5_resample_propagate<NX, Model, true>;
    const size_t smem = k5_smem_bytes<NX>();
    static bool attr_set = false;   // one flag per template instantiation
    if (!attr_set) {
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    SMC_CHECK_CUDA(launch_pdl(kern, (unsigned)grid, 256u, smem, l.st, a, m));
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// whether the plain fp32 resampler can be K5<CopyModel> (else k3_resample)
inline bool k5_copy_ok(const Launch& l) {
    const smc_sir_desc* d = l.d;
    if (d->dtype != SMC_F32 || d->nx > 2 || d->n != d->n_global || d->offset != 0) return false;
    if (d->flags & (SMC_FLAG_WRITE_ANCESTORS | SMC_FLAG_POST_MOMENTS_GENERIC | SMC_FLAG_GENERIC_MOMENTS)) return false;
    const int64_t ranges = k3_ranges(l.L.n_tiles, d->dtype, d->nx);
    return l.L.n_tiles >= (int64_t)sm_count() * 12 && ranges % k5_grid(d->nx) == 0 && !fused_legacy();
}

template <class Model>
int launch_fused(const Launch& l, const Model& m, void* logw_out_dev, const double* z_dev, double resample_thresh,
                 uint64_t seed, int64_t t, double* stats_prev_dev, double* stats_row_dev, int phases) {
    const smc_sir_desc* d = l.d;
    // nx = 2: K5 holds two tiles + two windows per CTA (94 KB, 2 CTAs per SM, 116 registers) and measured 0.201 ms
    // per 2^24 step against 0.152 ms of round 1's pipeline, which therefore still serves nx = 2
    const bool legacy = fused_legacy() || Model::NX >= 2;
    if (phases & 1) {
        FinishArgs f;
        memset(&f, 0, sizeof(f));
        f.stats = stats_prev_dev; f.nx = d->nx; f.inline_moments = 1;
        f.no_moments = (d->flags & SMC_FLAG_NO_MOMENTS) ? 1 : 0; f.n_global = d->n_global;
        K2Args fin2;
        memset(&fin2, 0, sizeof(fin2));
        fin2.stats = stats_row_dev; fin2.n_global = d->n_global; fin2.nx = d->nx;
        fin2.inline_moments = 0; fin2.f32 = 1;
        fin2.thresh = resample_thresh; fin2.u_host = -1.0; fin2.seed = seed; fin2.t = (uint32_t)t;
        quant_bound_for(d->n_global, &fin2.lw_bound, &fin2.kshift_bound);
        if (legacy) {
            K3Args<float> a;
            fill_k3_args<float>(l, a, d->logw_dev, d->nx, 0ull, 0, nullptr, 0, !(d->flags & SMC_FLAG_NO_MOMENTS), nullptr, 0,
                                &f, nullptr);
            a.fk.lw_out = static_cast<float*>(logw_out_dev);
            a.fk.z = z_dev;
            a.fk.keys = philox_expand(seed);
            a.fk.t = (uint32_t)t;
            a.fk.parts = l.k1_parts();            // <= 592 records of 3 doubles: fits the K1 partial area
            a.fk.fin = fin2;
            if (int rc = launch_k3_t<float, Model::NX, false, Model>(l, a, l.L.n_tiles, m)) return rc;
        } else {
            fin2.producer_rsum = 1;
            if (int rc = launch_k5<Model::NX, Model>(l, m, d->logw_dev, logw_out_dev, z_dev, seed, t, f, &fin2)) return rc;
        }
    }
    if (!(phases & 2)) return 0;
    // range pre-pass of step t over the weights just written: returns at once unless ctrl says it is needed
    return launch_prepass<float>(l, false, logw_out_dev, d->nx, 0, nullptr);
}

}  // namespace

extern "C" {

int smc_abi_version(void) { return SMC_ABI_VERSION; }
uint64_t smc_launch_count(void) { return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED); }
const char* smc_last_error(void) { return smc::g_err; }
uint64_t smc_sizeof_sir_desc(void) { return sizeof(smc_sir_desc); }
uint64_t smc_sizeof_shard_ctx(void) { return sizeof(smc_shard_ctx); }
int64_t smc_partial_stride(int32_t nx) { return partial_stride(nx); }

uint64_t smc_sir_workspace_bytes(int64_t n, int32_t nx, int32_t dtype) {
    if (n < 1) n = 1;
    return ws_layout(n, nx, dtype).total;
}

int smc_sir_reset(const smc_sir_desc* d, const double* shift_host, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    double* dev_shift = nullptr;
    if (shift_host)
        if (int rc = stage(l, 0, shift_host, d->nx < kMaxNx ? d->nx : kMaxNx, &dev_shift)) return rc;
    k_reset_ctrl<<<1, 32, 0, l.st>>>(l.ctrl(), dev_shift, d->nx);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_init_gaussian(const smc_sir_desc* d, const double* mean_host, const double* Lc_host, uint64_t seed,
                          void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    const int nx = d->nx;
    double *mean_dev, *lc_dev;
    if (int rc = stage(l, 0, mean_host, nx, &mean_dev)) return rc;
    if (int rc = stage(l, 1, Lc_host, nx * nx, &lc_dev)) return rc;
    const int64_t quads = (d->n + 3) / 4;
    const unsigned grid = (unsigned)cdiv64(quads, kThreads);
    if (d->dtype == SMC_F32)
        k_init_gaussian<float><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (float*)d->x0_dev, (float*)d->x1_dev, d->n, d->offset,
                                                           d->ld, nx, mean_dev, lc_dev, seed);
    else
        k_init_gaussian<double><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (double*)d->x0_dev, (double*)d->x1_dev, d->n,
                                                            d->offset, d->ld, nx, mean_dev, lc_dev, seed);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int sir_propagate_weight_impl(const smc_sir_desc* d, int32_t model_id, const double* params_host,
                                    int32_t n_params, int32_t mode, const double* z_dev, const void* noise_dev,
                                    uint64_t seed, int64_t t, void* stream, const K2Args* fin) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE((mode & 3) >= 1 && mode <= 7, "smc: bad mode %d", mode);
    SMC_REQUIRE(!(mode & SMC_MODE_WEIGHT) || z_dev, "smc: z_dev required when weighting");
    SMC_REQUIRE(params_host != nullptr, "smc: params missing");
    SMC_REQUIRE(model_id == SMC_MODEL_LORENZ96 || !(d->flags & SMC_FLAG_GENERIC_MOMENTS),
                "smc: fused K1 carries inline moments; clear GENERIC_MOMENTS");
    return d->dtype == SMC_F32
               ? dispatch_k1<float>(l, model_id, params_host, n_params, mode, z_dev, noise_dev, seed, t, fin)
               : dispatch_k1<double>(l, model_id, params_host, n_params, mode, z_dev, noise_dev, seed, t, fin);
}

int smc_sir_propagate_weight(const smc_sir_desc* d, int32_t model_id, const double* params_host, int32_t n_params,
                             int32_t mode, const double* z_dev, const void* noise_dev, uint64_t seed, int64_t t,
                             void* stream) {
    return sir_propagate_weight_impl(d, model_id, params_host, n_params, mode, z_dev, noise_dev, seed, t, stream,
                                     nullptr);
}

int smc_bridge_predict(const smc_sir_desc* d, const void* gx_dev, const void* noise_dev, const double* Lq_host,
                       double scale, int32_t target_other, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    const int nx = d->nx;
    SMC_REQUIRE(gx_dev || target_other, "smc: gx_dev may be NULL only with target_other = 1");
    SMC_REQUIRE(noise_dev && Lq_host, "smc: noise / Lq missing");
    double* lq_dev;
    if (int rc = stage(l, 1, Lq_host, nx * nx, &lq_dev)) return rc;
    const unsigned grid = (unsigned)cdiv64(d->n, kThreads);
    if (d->dtype == SMC_F32)
        k_bridge_predict<float><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (float*)d->x0_dev, (float*)d->x1_dev,
                                                            (const float*)gx_dev, (const float*)noise_dev, lq_dev, scale,
                                                            d->n, d->ld, nx, target_other ? 1 : 0);
    else
        k_bridge_predict<double><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (double*)d->x0_dev, (double*)d->x1_dev,
                                                             (const double*)gx_dev, (const double*)noise_dev, lq_dev, scale,
                                                             d->n, d->ld, nx, target_other ? 1 : 0);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_jitter(const smc_sir_desc* d, const double* Lq_host, double scale, uint64_t seed, int64_t t, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    const int nx = d->nx;
    SMC_REQUIRE(Lq_host && nx <= kMaxNx, "smc: jitter needs Lq and nx <= %d", kMaxNx);
    double* lq_dev;
    if (int rc = stage(l, 1, Lq_host, nx * nx, &lq_dev)) return rc;
    const unsigned grid = (unsigned)cdiv64(d->n, kThreads);
    const PhiloxKeys keys = philox_expand(seed);
    if (d->dtype == SMC_F32)
        k_jitter<float><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (float*)d->x0_dev, (float*)d->x1_dev, lq_dev, scale, d->n,
                                                    d->offset, d->ld, nx, keys, (uint32_t)t);
    else
        k_jitter<double><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (double*)d->x0_dev, (double*)d->x1_dev, lq_dev, scale, d->n,
                                                     d->offset, d->ld, nx, keys, (uint32_t)t);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_bridge_loglik(const smc_sir_desc* d, const void* zpred_dev, const double* z_host, const double* LRinv_host,
                      const void* wprev_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(d->flags & SMC_FLAG_GENERIC_MOMENTS, "smc: bridge mode needs SMC_FLAG_GENERIC_MOMENTS");
    const int nz = d->nz;
    double *z_dev, *lr_dev;
    if (int rc = stage(l, 2, z_host, nz, &z_dev)) return rc;
    if (int rc = stage(l, 3, LRinv_host, nz * nz, &lr_dev)) return rc;
    const unsigned grid = (unsigned)l.L.n_k1_blocks;
    if (d->dtype == SMC_F32)
        k_bridge_loglik<float><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (float*)d->logw_dev, (const float*)zpred_dev, z_dev,
                                                           lr_dev, (const float*)wprev_dev, d->n, d->n_global, d->ld, nz,
                                                           l.k1_parts());
    else
        k_bridge_loglik<double><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (double*)d->logw_dev, (const double*)zpred_dev, z_dev,
                                                            lr_dev, (const double*)wprev_dev, d->n, d->n_global, d->ld, nz,
                                                            l.k1_parts());
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_reduce_partials(const smc_sir_desc* d, double* out_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    const int ps = inline_moments(d) ? partial_stride(d->nx) : 3;
    k2_reduce_partials<<<1, kThreads, 0, l.st>>>(l.k1_parts(), (int)l.L.n_k1_blocks, ps, out_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_finalize(const smc_sir_desc* d, double resample_thresh, double u_host, uint64_t seed, int64_t t,
                     double* stats_row_dev, const double* parts_dev, int32_t n_parts_override, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    K2Args a;
    memset(&a, 0, sizeof(a));
    a.ctrl = l.ctrl();
    a.parts = n_parts_override > 0 ? parts_dev : l.k1_parts();
    a.n_parts = n_parts_override > 0 ? n_parts_override : (int)l.L.n_k1_blocks;
    a.stats = stats_row_dev;
    a.n_global = d->n_global;
    a.nx = d->nx;
    a.inline_moments = inline_moments(d) ? 1 : 0;
    a.f32 = d->dtype == SMC_F32;
    quant_bound_for(d->n_global, &a.lw_bound, &a.kshift_bound);
    a.thresh = resample_thresh; a.u_host = u_host; a.seed = seed; a.t = (uint32_t)t;
    k2_finalize<<<1, kThreads, 0, l.st>>>(a);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_set_uniform(const smc_sir_desc* d, double u, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(u >= 0.0 && u < 1.0, "smc: u must lie in [0, 1)");
    k_set_uniform<<<1, 32, 0, l.st>>>(l.ctrl(), u, d->n_global, d->dtype == SMC_F32);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_shard_weight_total(const smc_sir_desc* d, uint64_t* total_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(total_dev, "smc: total_dev missing");
    // the tile pre-pass of the resampler: leaves the tile prefixes in the workspace for
    // smc_sir_resample_sharded and the shard total (0 when the step does not resample) in total_dev
    SMC_REQUIRE(d->nx <= kFusedNxMax, "smc: sharded resampling requires nx <= %d", kFusedNxMax);
    return d->dtype == SMC_F32
               ? launch_prepass<float>(l, false, d->logw_dev, d->nx, 0, (unsigned long long*)total_dev)
               : launch_prepass<double>(l, false, d->logw_dev, d->nx, 0, (unsigned long long*)total_dev);
}

static int sir_resample_impl(const smc_sir_desc* d, uint64_t cdf_offset_fix, void* const* peer_x_dev,
                            int32_t n_shards, void* stream, const FinishArgs* fin) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    const bool fused = d->nx <= kFusedNxMax;
    const int nxf = fused ? d->nx : 0;
    SMC_REQUIRE(fused || d->ancestors_dev, "smc: nx > %d needs an ancestors buffer for the two-pass gather", kFusedNxMax);
    SMC_REQUIRE(!peer_x_dev || fused, "smc: cross-shard migration requires nx <= %d", kFusedNxMax);
    SMC_REQUIRE(!peer_x_dev || (n_shards > 0 && d->n_global % n_shards == 0), "smc: bad shard count");
    const bool want_m = inline_moments(d) && !(d->flags & SMC_FLAG_NO_MOMENTS);
    SMC_REQUIRE(!fin || fused, "smc: fused finish needs nx <= %d", kFusedNxMax);
    int rc = d->dtype == SMC_F32
                 ? launch_k3<float>(l, false, d->logw_dev, nxf, cdf_offset_fix, 0, peer_x_dev, n_shards, want_m, nullptr, 0,
                                    true, fin)
                 : launch_k3<double>(l, false, d->logw_dev, nxf, cdf_offset_fix, 0, peer_x_dev, n_shards, want_m, nullptr,
                                     0, true, fin);
    if (rc) return rc;
    if (!fused) {
        const unsigned grid = (unsigned)cdiv64(d->n, kThreads);
        if (d->dtype == SMC_F32)
            k_gather_step<float><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (const float*)d->x0_dev, (const float*)d->x1_dev,
                                                             (float*)d->x0_dev, (float*)d->x1_dev, d->ancestors_dev, d->n,
                                                             d->ld, d->nx);
        else
            k_gather_step<double><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (const double*)d->x0_dev, (const double*)d->x1_dev,
                                                              (double*)d->x0_dev, (double*)d->x1_dev, d->ancestors_dev,
                                                              d->n, d->ld, d->nx);
        SMC_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

int smc_sir_resample(const smc_sir_desc* d, uint64_t cdf_offset_fix, double* stats_row_dev, void* const* peer_x_dev,
                     int32_t n_shards, void* stream) {
    (void)stats_row_dev;
    return sir_resample_impl(d, cdf_offset_fix, peer_x_dev, n_shards, stream, nullptr);
}

int smc_sir_resample_sharded(const smc_sir_desc* d, const uint64_t* totals_dev, int32_t rank, int32_t n_shards,
                             void* const* peer_x_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(d->nx <= kFusedNxMax, "smc: sharded resampling requires nx <= %d", kFusedNxMax);
    SMC_REQUIRE(n_shards > 0 && rank >= 0 && rank < n_shards && d->n_global % n_shards == 0 &&
                    d->n == d->n_global / n_shards && d->offset == (int64_t)rank * d->n,
                "smc: inconsistent shard geometry (rank %d of %d)", rank, n_shards);
    SMC_REQUIRE(totals_dev && peer_x_dev, "smc: totals / peer table missing");
    const bool want_m = inline_moments(d) && !(d->flags & SMC_FLAG_NO_MOMENTS);
    return d->dtype == SMC_F32
               ? launch_k3<float>(l, false, d->logw_dev, d->nx, 0ull, 0, peer_x_dev, n_shards, want_m,
                                  (const unsigned long long*)totals_dev, rank, /*run_prepass=*/false)
               : launch_k3<double>(l, false, d->logw_dev, d->nx, 0ull, 0, peer_x_dev, n_shards, want_m,
                                   (const unsigned long long*)totals_dev, rank, /*run_prepass=*/false);
}

int smc_sir_reduce_moment_partials(const smc_sir_desc* d, double* out_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(inline_moments(d), "smc: moment partials exist only for nx <= %d", kFusedNxMax);
    k_reduce_moment_partials<<<1, kThreads, 0, l.st>>>(l.ctrl(), l.mom_parts(),
                                                      (int)k3_grid(l.L.n_tiles, d->dtype, d->nx), moment_stride(d->nx),
                                                      out_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_dev_alloc(uint64_t bytes, void** ptr_out) {
    SMC_REQUIRE(ptr_out, "smc_dev_alloc: null out pointer");
    SMC_CHECK_CUDA(cudaMalloc(ptr_out, bytes));
    SMC_CHECK_CUDA(cudaMemset(*ptr_out, 0, bytes));
    return 0;
}
int smc_dev_free(void* ptr) {
    SMC_CHECK_CUDA(cudaFree(ptr));
    return 0;
}
int smc_ipc_export(void* ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    SMC_CHECK_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, 64);
    return 0;
}
int smc_ipc_import(const unsigned char handle[64], void** ptr_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SMC_CHECK_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int smc_ipc_close(void* ptr) {
    SMC_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

int smc_sir_finish_step(const smc_sir_desc* d, double* stats_row_dev, const double* mom_parts_dev, int32_t n_mom_parts,
                        void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    const int no_m = (d->flags & SMC_FLAG_NO_MOMENTS) ? 1 : 0;
    // POST_MOMENTS_GENERIC: the particle set was modified after K3 (regularisation jitter) or was
    // not produced by K3 at all (multinomial gather), so its moments are recomputed from memory
    const bool inl = inline_moments(d) && !(d->flags & SMC_FLAG_POST_MOMENTS_GENERIC);
    if (!inl && !no_m && stats_row_dev && d->nx > kFusedNxMax && d->nx <= 64) {
        // high-dimensional states: one register-blocked X^T W X pass (weighted or uniform, decided on device);
        // beyond 64 dimensions (LEDH on 144- / 400-dimensional sensor networks) the any-nx kernel below
        SMC_REQUIRE(d->n == d->n_global, "smc: high-dimensional moments are single-shard only");
        if (int rc = (d->dtype == SMC_F32 ? launch_syrk<float>(l, stats_row_dev) : launch_syrk<double>(l, stats_row_dev)))
            return rc;
    } else if (!inl && !no_m && stats_row_dev) {
        // any-nx moments straight into the stats row (weighted or uniform, decided on device)
        SMC_REQUIRE(d->n == d->n_global, "smc: generic moments are single-shard only");
        const int nx = d->nx;
        const unsigned grid = (unsigned)(nx + nx * (nx + 1) / 2);
        if (d->dtype == SMC_F32)
            k_moments_generic<float><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (const float*)d->x0_dev, (const float*)d->x1_dev,
                                                                 (const float*)d->logw_dev, d->n, d->n_global, d->ld, nx,
                                                                 stats_row_dev, 0);
        else
            k_moments_generic<double><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (const double*)d->x0_dev,
                                                                  (const double*)d->x1_dev, (const double*)d->logw_dev, d->n,
                                                                  d->n_global, d->ld, nx, stats_row_dev, 0);
        SMC_CHECK_CUDA(cudaGetLastError());
        k_moments_generic_fix<<<1, 32, 0, l.st>>>(l.ctrl(), nx, stats_row_dev);
        SMC_CHECK_CUDA(cudaGetLastError());
    }
    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.stats = stats_row_dev;
    f.mom_parts = mom_parts_dev ? mom_parts_dev : l.mom_parts();
    f.n_mom_parts = n_mom_parts > 0 ? n_mom_parts : (inl ? (int)k3_grid(l.L.n_tiles, d->dtype, d->nx) : 0);
    f.nx = d->nx; f.inline_moments = inl ? 1 : 0; f.no_moments = no_m; f.n_global = d->n_global;
    k_finish_step<<<1, kFinishThreads, 0, l.st>>>(l.ctrl(), f);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_step_weight(const smc_sir_desc* d, int32_t model_id, const double* params_host, int32_t n_params,
                        const double* z_dev, const void* noise_dev, double resample_thresh, double u_host, uint64_t seed,
                        int64_t t, double* stats_row_dev, void* stream) {
    SMC_REQUIRE(d != nullptr, "smc: null descriptor");
    // resample_thresh > 1 means ESS <= N < thresh N: the step certainly resamples, so the weighted
    // (pre-resample) moments would be overwritten by the post-resample ones and are not accumulated
    const int mode = SMC_MODE_STEP | (resample_thresh > 1.0 ? 4 : 0);
    K2Args fin;
    memset(&fin, 0, sizeof(fin));
    fin.stats = stats_row_dev; fin.n_global = d->n_global; fin.nx = d->nx;
    fin.inline_moments = inline_moments(d) ? 1 : 0;
    fin.f32 = d->dtype == SMC_F32;
    quant_bound_for(d->n_global, &fin.lw_bound, &fin.kshift_bound);
    fin.thresh = resample_thresh; fin.u_host = u_host; fin.seed = seed; fin.t = (uint32_t)t;
    return sir_propagate_weight_impl(d, model_id, params_host, n_params, mode, z_dev, noise_dev, seed, t, stream, &fin);
}

int smc_sir_step_resample(const smc_sir_desc* d, int32_t phases, double* stats_row_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(phases >= 1 && phases <= 3, "smc: phases must be 1 (pre-pass), 2 (resample + finish) or 3");
    const bool fuse_finish = d->nx <= kFusedNxMax && inline_moments(d) && !(d->flags & SMC_FLAG_POST_MOMENTS_GENERIC);
    if (!fuse_finish) {
        SMC_REQUIRE(phases == 3, "smc: the split form needs the fused resampler (nx <= %d)", kFusedNxMax);
        if (int rc = smc_sir_resample(d, 0, stats_row_dev, nullptr, 0, stream)) return rc;
        return smc_sir_finish_step(d, stats_row_dev, nullptr, 0, stream);
    }
    if (phases & 1) {
        int rc = d->dtype == SMC_F32 ? launch_prepass<float>(l, false, d->logw_dev, d->nx, 0, nullptr)
                                     : launch_prepass<double>(l, false, d->logw_dev, d->nx, 0, nullptr);
        if (rc) return rc;
    }
    if (phases & 2) {
        FinishArgs f;
        memset(&f, 0, sizeof(f));
        f.stats = stats_row_dev; f.nx = d->nx; f.inline_moments = 1;
        f.no_moments = (d->flags & SMC_FLAG_NO_MOMENTS) ? 1 : 0; f.n_global = d->n_global;
        const bool want_m = !(d->flags & SMC_FLAG_NO_MOMENTS);
        if (k5_copy_ok(l)) {     // fp32, one shard, no ancestor trace: the expansion resampler
            if (d->nx == 1) return launch_k5<1, CopyModel<1>>(l, CopyModel<1>(), d->logw_dev, nullptr, nullptr, 0, 0, f, nullptr);
            return launch_k5<2, CopyModel<2>>(l, CopyModel<2>(), d->logw_dev, nullptr, nullptr, 0, 0, f, nullptr);
        }
        return d->dtype == SMC_F32
                   ? launch_k3<float>(l, false, d->logw_dev, d->nx, 0ull, 0, nullptr, 0, want_m, nullptr, 0, false, &f)
                   : launch_k3<double>(l, false, d->logw_dev, d->nx, 0ull, 0, nullptr, 0, want_m, nullptr, 0, false, &f);
    }
    return 0;
}

int smc_sir_step_fused(const smc_sir_desc* d, void* logw_out_dev, int32_t model_id, const double* params_host,
                       int32_t n_params, const double* z_dev, double resample_thresh, uint64_t seed, int64_t t,
                       double* stats_prev_dev, double* stats_row_dev, int32_t phases, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(phases >= 1 && phases <= 3, "smc: phases must be 1 (fused kernel), 2 (pre-pass of step t) or 3");
    SMC_REQUIRE(d->dtype == SMC_F32 && d->n == d->n_global && d->offset == 0,
                "smc: the fused step is the single-GPU fp32 throughput path");
    SMC_REQUIRE(resample_thresh > 1.0, "smc: the fused step is for filters that resample every step (thresh > 1)");
    SMC_REQUIRE(inline_moments(d) && !(d->flags & (SMC_FLAG_POST_MOMENTS_GENERIC | SMC_FLAG_WRITE_ANCESTORS)),
                "smc: the fused step needs inline moments and no ancestor trace");
    SMC_REQUIRE(logw_out_dev && logw_out_dev != d->logw_dev && z_dev && params_host, "smc: fused step: bad buffers");
    SMC_REQUIRE(k3_grid(l.L.n_tiles, d->dtype, d->nx, true) <= kK1MaxBlocks, "smc: partial area too small");
    SMC_REQUIRE(l.L.n_tiles >= (int64_t)sm_count() * 12, "smc: the fused step needs at least %d tiles", sm_count() * 12);
    const double* p = params_host;
    if (model_id == SMC_MODEL_STOCH_VOL) {
        SMC_REQUIRE(n_params == 4 && d->nx == 1 && d->nz == 1, "smc: STOCH_VOL needs 4 params and nx = nz = 1");
        StochVolModel<float> m;
        m.alpha = (float)p[0]; m.lq = (float)p[1]; m.beta = (float)p[2]; m.rinv = (float)(1.0 / p[3]);
        return launch_fused(l, m, logw_out_dev, z_dev, resample_thresh, seed, t, stats_prev_dev, stats_row_dev, phases);
    }
    if (model_id == SMC_MODEL_LINEAR_GAUSSIAN && d->nx == 2 && d->nz == 2) {
        SMC_REQUIRE(n_params == 16, "smc: LINEAR_GAUSSIAN (2,2) needs 16 params");
        LinGaussModel<float, 2, 2> m;
        const double* q = p;
        for (int i = 0; i < 4; ++i) m.A[i] = (float)*q++;
        for (int i = 0; i < 4; ++i) m.Lq[i] = (float)*q++;
        for (int i = 0; i < 4; ++i) m.H[i] = (float)*q++;
        for (int i = 0; i < 4; ++i) m.LRinv[i] = (float)*q++;
        return launch_fused(l, m, logw_out_dev, z_dev, resample_thresh, seed, t, stats_prev_dev, stats_row_dev, phases);
    }
    SMC_REQUIRE(false, "smc: the fused step is compiled for STOCH_VOL and LINEAR_GAUSSIAN (2,2)");
}

// ---- multi-GPU: whole sharded step with mailbox exchanges (no collective calls) ----------------
uint64_t smc_mailbox_bytes(void) { return sizeof(Mailbox); }
uint64_t smc_mailbox_error_offset(void) { return offsetof(Mailbox, error); }
#ifdef SMC_K5_TRACE
int smc_k5_trace_read(unsigned long long* host_out) {      // tuning builds only: 2 x 1024 x 16 stamps
    SMC_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, smc::g_k5_trace, sizeof(unsigned long long) * 2 * 1024 * 16));
    return 0;
}
#endif
uint64_t smc_shard_rsum_bytes(void) { return sizeof(unsigned long long) * (2 * kRsumStride + kMaxShards); }

int smc_sir_step_sharded(const smc_sir_desc* d, const smc_shard_ctx* ctx, int32_t model_id, const double* params_host,
                         int32_t n_params, const double* z_dev, double resample_thresh, uint64_t seed, int64_t t,
                         double* stats_row_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    SMC_REQUIRE(ctx && ctx->mail_dev && ctx->peer_x_dev, "smc: shard context incomplete");
    const int G = ctx->n_shards, rank = ctx->rank;
    SMC_REQUIRE(G >= 1 && G <= kMaxShards && rank >= 0 && rank < G, "smc: bad shard geometry (rank %d of %d)", rank, G);
    SMC_REQUIRE(d->n_global % G == 0 && d->n == d->n_global / G && d->offset == (int64_t)rank * d->n,
                "smc: inconsistent shard geometry (rank %d of %d)", rank, G);
    SMC_REQUIRE(d->nx <= kFusedNxMax && inline_moments(d), "smc: sharded step requires nx <= %d", kFusedNxMax);
    SMC_REQUIRE(ctx->stamp > 0, "smc: stamp must be positive and increase by one per step");
    SMC_REQUIRE(partial_stride(d->nx) <= 32 && moment_stride(d->nx) <= 16, "smc: record too wide for the mailbox");
    ShardCtx sc;
    sc.mail = reinterpret_cast<Mailbox* const*>(ctx->mail_dev);
    sc.rank = rank; sc.n_shards = G; sc.stamp = ctx->stamp;
    // 1. K1; its last CTA publishes the shard's partial record
    const int mode = SMC_MODE_STEP | (resample_thresh > 1.0 ? 4 : 0);
    SMC_REQUIRE(params_host != nullptr && z_dev != nullptr, "smc: params / observation missing");
    int rc = d->dtype == SMC_F32
                 ? dispatch_k1<float>(l, model_id, params_host, n_params, mode, z_dev, nullptr, seed, t, nullptr, &sc)
                 : dispatch_k1<double>(l, model_id, params_host, n_params, mode, z_dev, nullptr, seed, t, nullptr, &sc);
    if (rc) return rc;
    // 2. finalize once all G records have arrived
    K2Args fin;
    memset(&fin, 0, sizeof(fin));
    fin.ctrl = l.ctrl(); fin.stats = stats_row_dev; fin.n_global = d->n_global; fin.nx = d->nx;
    fin.inline_moments = 1; fin.f32 = d->dtype == SMC_F32;
    quant_bound_for(d->n_global, &fin.lw_bound, &fin.kshift_bound);
    fin.thresh = resample_thresh; fin.u_host = -1.0; fin.seed = seed; fin.t = (uint32_t)t;
    fin.producer_rsum = k1_rsum_ok(l) ? 1 : 0;       // (and the step is a bootstrap step: decided on the device)
    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.stats = stats_row_dev; f.nx = d->nx; f.inline_moments = 1;
    f.no_moments = (d->flags & SMC_FLAG_NO_MOMENTS) ? 1 : 0; f.n_global = d->n_global;
    if (resample_thresh > 1.0 && fin.producer_rsum && d->dtype == SMC_F32 && !getenv("SMC_SHARDED_5LAUNCH")) {
        // Bootstrap filter on large shards: THREE launches, two exchanges.  K1 has published (partial record, weight
        // total); the resampler finalises the step in every CTA's prologue once the G records are there, migrates by
        // peer stores and publishes the moment record; the finish waits for the G moment flags.
        const bool want_m = !(d->flags & SMC_FLAG_NO_MOMENTS);
        rc = launch_k3<float>(l, false, d->logw_dev, d->nx, 0ull, 0, ctx->peer_x_dev, G, want_m, nullptr, rank, false, nullptr,
                              &sc, &fin);
        if (rc) return rc;
        __atomic_add_fetch(&g_launch_count, 1ull, __ATOMIC_RELAXED);
        k_finish_step_sharded<<<1, kThreads, 0, l.st>>>(l.ctrl(), f, sc);
        SMC_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    __atomic_add_fetch(&g_launch_count, 2ull, __ATOMIC_RELAXED);     // finalize + finish (the others count themselves)
    k2_finalize_sharded<<<1, kThreads, 0, l.st>>>(fin, sc);
    SMC_CHECK_CUDA(cudaGetLastError());
    if (resample_thresh > 0.0) {
        // 3. range pre-pass; its last CTA publishes the shard's weight total
        rc = d->dtype == SMC_F32 ? launch_prepass<float>(l, false, d->logw_dev, d->nx, 0, nullptr, &sc)
                                 : launch_prepass<double>(l, false, d->logw_dev, d->nx, 0, nullptr, &sc);
        if (rc) return rc;
        // 4. resampling with migration by peer stores; its last CTA publishes the moment record
        const bool want_m = !(d->flags & SMC_FLAG_NO_MOMENTS);
        rc = d->dtype == SMC_F32
                 ? launch_k3<float>(l, false, d->logw_dev, d->nx, 0ull, 0, ctx->peer_x_dev, G, want_m, nullptr, rank, false,
                                    nullptr, &sc)
                 : launch_k3<double>(l, false, d->logw_dev, d->nx, 0ull, 0, ctx->peer_x_dev, G, want_m, nullptr, rank, false,
                                     nullptr, &sc);
        if (rc) return rc;
        // 5. finish once all G moment flags are up
        k_finish_step_sharded<<<1, kThreads, 0, l.st>>>(l.ctrl(), f, sc);
    } else {
        f.n_mom_parts = 0;
        k_finish_step<<<1, kFinishThreads, 0, l.st>>>(l.ctrl(), f);
    }
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---- sharded fused run: K1 of the first step, then one k5_resample_propagate launch per step, then flush ----
static int shard_ctx_check(const smc_sir_desc* d, const smc_shard_ctx* ctx, ShardCtx* sc) {
    SMC_REQUIRE(ctx && ctx->mail_dev && ctx->peer_x_dev, "smc: shard context incomplete");
    const int G = ctx->n_shards, rank = ctx->rank;
    SMC_REQUIRE(G >= 1 && G <= kMaxShards && rank >= 0 && rank < G, "smc: bad shard geometry (rank %d of %d)", rank, G);
    SMC_REQUIRE(d->n_global % G == 0 && d->n == d->n_global / G && d->offset == (int64_t)rank * d->n,
                "smc: inconsistent shard geometry (rank %d of %d)", rank, G);
    SMC_REQUIRE(ctx->stamp > 0, "smc: stamp must be positive and increase by one per launch group");
    sc->mail = reinterpret_cast<Mailbox* const*>(ctx->mail_dev);
    sc->rank = rank; sc->n_shards = G; sc->stamp = ctx->stamp;
    return 0;
}

int smc_sir_weight_sharded(const smc_sir_desc* d, const smc_shard_ctx* ctx, int32_t model_id, const double* params_host,
                           int32_t n_params, const double* z_dev, double resample_thresh, uint64_t seed, int64_t t,
                           void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    ShardCtx sc;
    if (int rc = shard_ctx_check(d, ctx, &sc)) return rc;
    SMC_REQUIRE(d->nx <= kFusedNxMax && inline_moments(d), "smc: sharded step requires nx <= %d", kFusedNxMax);
    SMC_REQUIRE(params_host != nullptr && z_dev != nullptr, "smc: params / observation missing");
    const int mode = SMC_MODE_STEP | (resample_thresh > 1.0 ? 4 : 0);
    return d->dtype == SMC_F32
               ? dispatch_k1<float>(l, model_id, params_host, n_params, mode, z_dev, nullptr, seed, t, nullptr, &sc)
               : dispatch_k1<double>(l, model_id, params_host, n_params, mode, z_dev, nullptr, seed, t, nullptr, &sc);
}

int smc_sir_step_fused_sharded(const smc_sir_desc* d, const smc_shard_ctx* ctx, int32_t lw_out_sel, int32_t first,
                               int32_t model_id, const double* params_host, int32_t n_params, const double* z_dev,
                               double resample_thresh, uint64_t seed, int64_t t, double* stats_prev2_dev,
                               double* stats_prev_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    ShardCtx sc;
    if (int rc = shard_ctx_check(d, ctx, &sc)) return rc;
    SMC_REQUIRE(ctx->peer_aux_dev && ctx->mailbox_dev && ctx->rsum_area_dev,
                "smc: the fused sharded step needs the peer tables of the log-weight arrays and range totals");
    SMC_REQUIRE(ctx->stamp > 1, "smc: the fused sharded step follows smc_sir_weight_sharded (stamp - 1) or another fused step");
    SMC_REQUIRE(k5_sharded_ok(l), "smc: the fused sharded step is the fp32, nx = 1 throughput path for shards of at least "
                                  "%d tiles of 4096 particles", sm_count() * 12);
    SMC_REQUIRE(resample_thresh > 1.0, "smc: the fused step is for filters that resample every step (thresh > 1)");
    SMC_REQUIRE(lw_out_sel == 0 || lw_out_sel == 1, "smc: lw_out_sel selects one of the two log-weight arrays");
    SMC_REQUIRE(model_id == SMC_MODEL_STOCH_VOL && n_params == 4 && d->nz == 1 && params_host && z_dev,
                "smc: the fused sharded step is compiled for STOCH_VOL");
    StochVolModel<float> m;
    m.alpha = (float)params_host[0]; m.lq = (float)params_host[1]; m.beta = (float)params_host[2];
    m.rinv = (float)(1.0 / params_host[3]);
    return launch_k5_sharded(l, m, ctx, lw_out_sel, first, z_dev, resample_thresh, seed, t, stats_prev2_dev, stats_prev_dev);
}

int smc_sir_fused_sharded_flush(const smc_sir_desc* d, const smc_shard_ctx* ctx, double resample_thresh, uint64_t seed,
                                int64_t t_last, double* stats_prev2_dev, double* stats_prev_dev, int32_t resample,
                                void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    ShardCtx sc;
    if (int rc = shard_ctx_check(d, ctx, &sc)) return rc;
    SMC_REQUIRE(ctx->peer_aux_dev && ctx->rsum_area_dev && ctx->stamp > 1, "smc: flush follows a fused sharded step");
    K2Args fin;
    memset(&fin, 0, sizeof(fin));
    fin.stats = stats_prev_dev; fin.n_global = d->n_global; fin.nx = d->nx; fin.inline_moments = 0; fin.f32 = 1;
    quant_bound_for(d->n_global, &fin.lw_bound, &fin.kshift_bound);
    fin.thresh = resample_thresh; fin.u_host = -1.0; fin.seed = seed; fin.t = (uint32_t)t_last; fin.producer_rsum = 1;
    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.stats = stats_prev2_dev; f.nx = d->nx; f.inline_moments = 1;
    f.no_moments = (d->flags & SMC_FLAG_NO_MOMENTS) ? 1 : 0; f.n_global = d->n_global;
    __atomic_add_fetch(&g_launch_count, 1ull, __ATOMIC_RELAXED);
    k5s_flush<<<1, kThreads, 0, l.st>>>(l.ctrl(), fin, f, sc, static_cast<unsigned long long*>(ctx->rsum_area_dev));
    SMC_CHECK_CUDA(cudaGetLastError());
    if (!resample) return 0;
    // resampling of the last step (stamp of this call): range pre-pass, resampler with migration, finish
    int rc = launch_prepass<float>(l, false, d->logw_dev, d->nx, 0, nullptr, &sc);
    if (rc) return rc;
    const bool want_m = !(d->flags & SMC_FLAG_NO_MOMENTS);
    rc = launch_k3<float>(l, false, d->logw_dev, d->nx, 0ull, 0, ctx->peer_x_dev, ctx->n_shards, want_m, nullptr, ctx->rank,
                          false, nullptr, &sc);
    if (rc) return rc;
    f.stats = stats_prev_dev;
    __atomic_add_fetch(&g_launch_count, 1ull, __ATOMIC_RELAXED);
    k_finish_step_sharded<<<1, kThreads, 0, l.st>>>(l.ctrl(), f, sc);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_step(const smc_sir_desc* d, int32_t model_id, const double* params_host, int32_t n_params,
                 const double* z_dev, const void* noise_dev, double resample_thresh, double u_host, uint64_t seed,
                 int64_t t, double* stats_row_dev, void* stream) {
    // Four launches per step: K1 (+ finalize in its last CTA), tile pre-pass, tile scan, K3 (+ finish
    // in its last CTA).
    if (int rc = smc_sir_step_weight(d, model_id, params_host, n_params, z_dev, noise_dev, resample_thresh, u_host, seed,
                                     t, stats_row_dev, stream))
        return rc;
    if (resample_thresh > 0.0) return smc_sir_step_resample(d, 3, stats_row_dev, stream);
    return smc_sir_finish_step(d, stats_row_dev, nullptr, 0, stream);
}

int smc_sir_step_host(const smc_sir_desc* d, int32_t model_id, const double* params_host, int32_t n_params,
                      double* z_dev, const double* z_host, double resample_thresh, double u_host, uint64_t seed,
                      int64_t t, double* stats_row_dev, double* row_host, int32_t n_row, void* stream) {
    // ParticleFilter.step(z) in one call: observation in, statistics row out, one stream synchronisation
    SMC_REQUIRE(d && z_dev && z_host && stats_row_dev && row_host && n_row > 0, "smc_sir_step_host: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SMC_CHECK_CUDA(cudaMemcpyAsync(z_dev, z_host, sizeof(double) * (size_t)d->nz, cudaMemcpyHostToDevice, st));
    if (int rc = smc_sir_step(d, model_id, params_host, n_params, z_dev, nullptr, resample_thresh, u_host, seed, t,
                              stats_row_dev, stream))
        return rc;
    SMC_CHECK_CUDA(cudaMemcpyAsync(row_host, stats_row_dev, sizeof(double) * (size_t)n_row, cudaMemcpyDeviceToHost, st));
    SMC_CHECK_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int smc_systematic_ancestors(const double* w_dev, int64_t n, double u, int32_t* ancestors_dev, void* ws_dev,
                             uint64_t ws_bytes, void* stream) {
    smc_sir_desc d;
    memset(&d, 0, sizeof(d));
    d.n = n; d.n_global = n; d.offset = 0; d.ld = ((n + 3) / 4) * 4; d.nx = 1; d.nz = 1; d.dtype = SMC_F64;
    d.flags = SMC_FLAG_WRITE_ANCESTORS | SMC_FLAG_NO_MOMENTS;
    d.ancestors_dev = ancestors_dev; d.ws_dev = ws_dev; d.ws_bytes = ws_bytes;
    Launch l;
    if (int rc = check_desc(&d, &l, stream)) return rc;
    SMC_REQUIRE(u >= 0.0 && u < 1.0, "smc: u must lie in [0, 1)");
    SMC_REQUIRE(w_dev && ancestors_dev, "smc: null buffer");
    k_set_resample<<<1, 32, 0, l.st>>>(l.ctrl(), u, d.n_global, 0);
    SMC_CHECK_CUDA(cudaGetLastError());
    return launch_k3<double>(l, true, w_dev, 0, 0ull, 1, nullptr, 0, false);
}

int smc_gather(const void* src_dev, void* dst_dev, const int32_t* idx_dev, int64_t n, int64_t ld, int32_t nx,
               int32_t dtype, void* stream) {
    SMC_REQUIRE(n > 0 && nx > 0 && ld >= n, "smc_gather: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)cdiv64(n, kThreads);
    if (dtype == SMC_F32)
        k_gather<float><<<grid, kThreads, 0, st>>>((const float*)src_dev, (float*)dst_dev, idx_dev, n, ld, nx);
    else if (dtype == SMC_F64)
        k_gather<double><<<grid, kThreads, 0, st>>>((const double*)src_dev, (double*)dst_dev, idx_dev, n, ld, nx);
    else
        SMC_REQUIRE(false, "smc_gather: bad dtype");
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_export_weights(const smc_sir_desc* d, double* w_dev, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    const unsigned grid = (unsigned)cdiv64(d->n, kThreads);
    if (d->dtype == SMC_F32)
        k_export_weights<float><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (const float*)d->logw_dev, d->n, d->n_global, w_dev);
    else
        k_export_weights<double><<<grid, kThreads, 0, l.st>>>(l.ctrl(), (const double*)d->logw_dev, d->n, d->n_global,
                                                             w_dev);
    SMC_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int smc_sir_current_buffer(const smc_sir_desc* d, int32_t* cur_host, void* stream) {
    Launch l;
    if (int rc = check_desc(d, &l, stream)) return rc;
    Ctrl h;
    SMC_CHECK_CUDA(cudaMemcpyAsync(&h, l.ctrl(), sizeof(Ctrl), cudaMemcpyDeviceToHost, l.st));
    SMC_CHECK_CUDA(cudaStreamSynchronize(l.st));
    *cur_host = h.cur;
    return 0;
}

}  // extern "C"
