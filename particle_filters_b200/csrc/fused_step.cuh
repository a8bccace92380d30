// K5: the fused step of the run() pipeline -- "resample step t-1, propagate + weigh step t" in ONE launch,
// 16 B per particle-update (R logw, R x, W x, W logw), for filters that resample every step (bootstrap
// filters; fp32).  Included by sir_kernels.cu (namespace smc).
//
// Reference semantics: models/particle_filter.py:146-171 (systematic resampling), :209-210 (gather, uniform
// weights), :266-268 (post-resample moments) of step t-1, then :223-237 (predict) and :253-263 (log-weights,
// normaliser, ESS) of step t.
//
// Per 4096-particle tile (four barriers):
//   B  the tile's log-weights (cp.async, issued one tile ahead) are quantised to fixed point, a = rni(p),
//      p = 2^(lw log2e + kshift); thread / warp / CTA scan of the exact 64-bit sums
//   C  the thread's first and last output boundary from the exact CDF (128-bit products with rho, RhoMap)
//   D  EXPANSION: the thread that owns an item stores the item's PAYLOAD once per offspring into a linear
//      output window in shared memory (boundaries inside the thread from the fp32 sums of p) -- no markers,
//      no back-probing, no gather: the window is the resampled tile
//   E  every thread takes quads of four consecutive outputs (one LDS.128 per dimension, laid out on global
//      quads), accumulates the post-resample moments of step t-1, draws the quad's Philox block of step t,
//      runs the model functor and stores x and log-weights of step t with 128-bit stores.
//   Skewed tiles (a thread with more than 64 outputs, or more outputs than the window holds) take the
//   marker + max-scan route of k3_resample and run the functor over the written particles in place.
//
// No range pre-pass: a bootstrap step's log-weights are bounded a priori (lw <= -log N, the Gaussian
// constant being omitted as in the reference), so phase E quantises the weights it writes against that
// bound -- with the very instructions phase B of the NEXT launch will use -- and adds them, exactly, into the
// per-range totals the next launch's CTAs take their CDF prefixes from (64-bit integer atomics: the result
// does not depend on the order).  The last CTA decides (finalize_body) whether the bound left the weights
// enough resolution; if not -- an observation that costs the whole cloud more than ~11 nats -- it says so in
// the control block and the (otherwise empty) pre-pass launch that follows re-quantises against the
// largest log-weight of the step.
//
// Two more things live in this file:
//  * the SHARDED form (template flag; K5Shard below): the same kernel on one shard of a G-way split population, one
//    launch and one exchange per step -- finalize / finish in every CTA's prologue from the slot all ranks published at
//    the end of the launch before, outputs stored into the memory of the rank that owns their index, in-kernel range
//    pre-pass on outlier steps;
//  * the placement board (K5Args::board): the CTAs that share an SM rank themselves by block index and take unequal
//    shares of the SM's ranges, because they do not advance at the same pace (profiles/r2_k5_shares.txt).
#pragma once

namespace smc {

// Resident CTAs per SM: shared memory = (1 + NX) tiles + NX windows (51 KB at NX = 1)
#ifndef SMC_K5_OCC
#define SMC_K5_OCC 4
#endif
#ifndef SMC_K5S_OCC
#define SMC_K5S_OCC 3            // resident CTAs per SM of the sharded form (80 registers, no spills worth mentioning)
#endif
#ifndef SMC_K5_UNROLL
#define SMC_K5_UNROLL 1          // unroll factor of phase E's quad loop (tuning knob)
#endif
template <int NX> struct K5Occ { static constexpr int value = NX <= 1 ? SMC_K5_OCC : 2; };
constexpr int kK5Tile = 4096;
constexpr int kK5Window = 256 * 18;          // u16 marker slots of the skewed route (aliases the output window)
constexpr int kK5WinStride = 4096 + 1024 + 512;   // output window: carried outputs (< 1027) + one tile's new ones
template <int NX> constexpr size_t k5_smem_bytes() {
    return sizeof(float) * ((size_t)kK5Tile * (1 + NX) + (size_t)kK5WinStride * NX);
}

template <bool B> struct K5Flag { static constexpr bool value = B; };

// "No functor": K5 then is the plain fp32 resampler -- gather, uniform weights, post-resample moments, finish
// (models/particle_filter.py:209-210, :266-268) -- with the same boundaries, hence the same ancestors, as the
// fused instantiations and as k3_resample (which keeps the fp64 parity mode, the ancestor trace and the shards).
template <int NX_> struct CopyModel { static constexpr int NX = NX_, NZ = 1; };
template <class M> struct K5IsCopy { static constexpr bool value = false; };
template <int NX_> struct K5IsCopy<CopyModel<NX_>> { static constexpr bool value = true; };

// Sharded form (SHARDED = true): one shard of a G-way split population, ONE launch and ONE exchange per step.
//  * every CTA's prologue waits (in the rank's own mailbox) for the slot the previous launch of every rank
//    published, finalises step t-1 and finishes step t-2 on a shared-memory copy of the control block (one warp
//    each, identical arithmetic on every CTA of every rank) and takes its CDF prefix from the G x G table of
//    weight contributions plus the rank's own range totals;
//  * a resampled quad goes through the functor on the rank that holds its ancestor and is then stored -- particle,
//    log-weight, and its fixed-point weight into the range totals -- into the memory of the rank that owns its index
//    (NVLink peer stores / 64-bit peer atomics for the few that cross a shard boundary);
//  * the last CTA publishes the rank's slot of this launch to every mailbox; its flag also certifies that all this
//    rank's peer stores have landed.
//  A step whose weights the a-priori bound leaves without resolution runs the range pre-pass inside the kernel
//  (every CTA its own ranges, a grid barrier, one extra exchange of the shard totals).
struct K5Shard {
    ShardCtx ctx;              // stamp = this launch's; the slot it consumes carries stamp - 1
    int64_t n_global, offset, shard_n;
    int32_t first;             // step t-1's log-weights were written by K1: records in flag_rec / rec / tot
    float* const* peer_x0;     // per-rank particle buffers (device tables of G pointers)
    float* const* peer_x1;
    float* const* peer_lw_out; // per-rank log-weight arrays of step t
    unsigned long long* const* peer_rsum;   // per-rank range-total areas: two buffers of kRsumStride, then G accumulators
    Mailbox* own;              // this rank's mailbox and range-total area (= the tables' own entries: saves two
    unsigned long long* own_rsum;           //  dependent loads on the critical path of every launch)
    int32_t rsel_in, rsel_out; // which buffer of the area this launch reads (first: the workspace's) / accumulates into
    unsigned int* gbar;        // grid-barrier counter of the in-kernel pre-pass
    K2Args pfin;               // finalize of step t-1
    FinishArgs pfinish;        // finish of step t-2
};

struct K5Args {
    Ctrl* ctrl;
    const float* lw_in;        // log-weights of step t-1
    float* lw_out;             // log-weights of step t (a second array: other CTAs still read step t-1's)
    const float* x0;
    const float* x1;
    float* dst0;               // destination when ctrl->cur == 1 (buffer 0)
    float* dst1;
    unsigned long long* rsum;  // two buffers of kRsumStride range totals (ctrl->rsum_sel: the one to read)
    double* mom_parts;         // per-CTA post-resample moment records of step t-1 (nullable)
    double* parts;             // per-CTA (max, sum e, sum e^2) records of step t
    const double* z;           // observation of step t (device)
    PhiloxKeys keys;
    uint32_t t;
    int64_t n, ld, n_tiles;
    int32_t n_ranges, ranges_per_cta;
    float kshift_bound;        // quantisation shift of the a-priori bound (K2Args::kshift_bound)
    unsigned int* ticket;
    // Placement board (nullable = static mapping, CTA b takes the b-th run of ranges_per_cta ranges).  The CTAs that
    // share an SM do not advance at the same pace: with equal work the one with the lowest block index leaves the tile
    // loop after 0.74 of the time the third one needs, and the SM ends the launch on a single CTA's eight warps
    // (profiles/r2_k5_shares.txt).  With a board the work is cut into slots x n_waves PIECES, piece (s, k) = share[k] of
    // the 12 consecutive ranges of slot s; the CTAs of an SM find one another on the SM's row of the board, rank
    // themselves by block index and take the pieces of one slot, the lowest block index the largest.  Correctness never
    // depends on the placement: a piece is claimed by compare-and-swap, a CTA that finds its piece taken (or its SM's
    // row incomplete after 4 us) takes the next free one, and the per-CTA records are indexed by piece.
    unsigned int* board;
    uint32_t board_wait_ns;    // how long a CTA waits for the other CTAs of its SM before it takes any free piece (4000;
                               // SMC_K5_BOARD_WAIT_NS: the tests run with 0 and 150 as well -- all / some CTAs fall back)
    int32_t n_waves, slots;
    int8_t share[8];
    FinishArgs fin;            // finish of step t-1
    K2Args fin2;               // finalize of step t
    K5Shard sh;                // sharded form only
};

// one-warp finish of a resampled step from G moment records (finish_body's arithmetic on a handful of records):
// posterior mean / covariance into the statistics row (`stats` nullable: the shift is refreshed all the same)
__device__ inline void finish_warp(Ctrl* c, const FinishArgs& f, const double* mom, int n_parts, int stride,
                                   double* stats, double* s_m) {
    const int lane = threadIdx.x & 31, nm = moment_stride(f.nx);
    if (lane < nm) {
        double t = 0;
        for (int p = 0; p < n_parts; ++p) t += __ldcg(mom + (int64_t)p * stride + lane);     // rank order
        s_m[lane] = t;
    }
    __syncwarp();
    if (lane == 0) {
        const double inv = 1.0 / (double)f.n_global;
        double m1[kFusedNxMax], mean[kFusedNxMax];
        for (int d = 0; d < f.nx; ++d) { m1[d] = s_m[d] * inv; mean[d] = c->shift[d] + m1[d]; }
        if (stats) {
            double* cov = stats + SMC_STATS_HEADER + f.nx;
            int k = f.nx;
            for (int d = 0; d < f.nx; ++d) {
                stats[SMC_STATS_HEADER + d] = mean[d];
                for (int d2 = 0; d2 <= d; ++d2) {
                    const double v = s_m[k++] * inv - m1[d] * m1[d2];
                    cov[d * f.nx + d2] = v;
                    cov[d2 * f.nx + d] = v;
                }
            }
        }
        for (int d = 0; d < f.nx && d < kMaxNx; ++d) c->shift[d] = mean[d];
    }
    __syncwarp();
}

// -DSMC_K5_TRACE (tuning builds only, profiles/micro/k5s_trace.py): per-CTA globaltimer stamps of the sharded form --
// [launch parity][CTA][entry, after wait, after finalize, after prefix, after tiles, ticket, last-CTA done]
#ifdef SMC_K5_TRACE
__device__ unsigned long long g_k5_trace[2 * 1024 * 16];
#define K5_TRACE(slot) do { if (SHARDED && threadIdx.x == 0) g_k5_trace[(((a.sh.ctx.stamp & 1ull) * 1024 + blockIdx.x) << 4) + (slot)] = globaltimer_ns(); } while (0)
#else
#define K5_TRACE(slot) do { } while (0)
#endif

template <bool SHARDED> struct K5CtrlCopy {
    Ctrl c; unsigned long long contrib[kMaxShards]; double pout[32]; double m[16];
    Mailbox* mail[kMaxShards];          // the ranks' mailboxes (fetched once, while the CTA waits for the flags)
};
template <> struct K5CtrlCopy<false> { int unused; };

template <int NX, class Model, bool SHARDED = false>
__global__ void __launch_bounds__(256, SHARDED ? SMC_K5S_OCC : K5Occ<NX>::value) k5_resample_propagate(K5Args a, Model fm) {
    static_assert(Model::NX == NX, "functor / payload mismatch");
    static_assert(!SHARDED || !K5IsCopy<Model>::value, "the sharded form always carries a functor");
    constexpr int THREADS = 256, WARPS = 8, IPT = 16, UPT = 4, TILE = kK5Tile;
    constexpr int W = kK5Window, WIN = kK5WinStride, CH = IPT + 2;
    constexpr uint32_t MAXD = 4 * IPT;
    constexpr int NZ = Model::NZ;
    constexpr bool COPY = K5IsCopy<Model>::value;      // resample only (no step t functor)
    constexpr int NM = NX + NX * (NX + 1) / 2;
    static_assert(W == THREADS * CH && WIN % 4 == 0 && 2 * W <= 4 * WIN, "window geometry");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* lw_s = reinterpret_cast<float4*>(smem_raw);                                   // TILE floats, units swizzled
    float* x_s = reinterpret_cast<float*>(smem_raw + sizeof(float) * TILE);               // NX x TILE, same swizzle
    float* win = reinterpret_cast<float*>(smem_raw + sizeof(float) * TILE * (1 + NX));    // NX x WIN
    unsigned short* ids = reinterpret_cast<unsigned short*>(win);                         // skewed route only (aliases win)
    __shared__ unsigned long long s_warp_tot[WARPS];
    __shared__ unsigned long long s_warp_all[WARPS];
    __shared__ uint32_t s_E[2];
    __shared__ int s_wmax[WARPS];
    __shared__ int s_last;
    // scratch of the kernel's tail (after the tile loop): aliases the tile buffers
    double (*s_mom)[16] = reinterpret_cast<double (*)[16]>(smem_raw);             // [WARPS][16]
    double* s_m = reinterpret_cast<double*>(smem_raw) + WARPS * 16;               // [16]
    double* s_fout = s_m + 16;                                                    // [32]
    double* s_ftmp = s_fout + 32;                                                 // [WARPS]
    double (*s_fred)[2] = reinterpret_cast<double (*)[2]>(s_ftmp + WARPS);        // [WARPS][2]
    float* s_fmax = reinterpret_cast<float*>(s_ftmp + WARPS + 2 * WARPS);         // [WARPS]

    Ctrl* const gctrl = a.ctrl;
    Ctrl* ctrl = gctrl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();
    pdl_trigger();
    K5_TRACE(0);
#ifdef SMC_K5_TRACE
    if (SHARDED && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_k5_trace[(((a.sh.ctx.stamp & 1ull) * 1024 + blockIdx.x) << 4) + 7] = smid;
    }
#endif
    __shared__ K5CtrlCopy<SHARDED> s_sh;
    [[maybe_unused]] Mailbox* own = nullptr;
    [[maybe_unused]] const FusedSlot* slot_in = nullptr;
    [[maybe_unused]] const int G = SHARDED ? a.sh.ctx.n_shards : 1;
    // range totals of the log-weights to resample (eight per thread, all loads in flight at once) and, sharded, one
    // entry of the G x G table of weight contributions per thread: thread (g, src) = (tid / 16, tid % 16)
    static_assert(kRsumStride == 8 * THREADS, "range totals: eight per thread");
    static_assert(kMaxShards == 16 && THREADS == 256, "one thread per (rank, producing rank) pair");
    unsigned long long tv[8];
    [[maybe_unused]] unsigned long long tc = 0ull;
    auto load_totals = [&](const unsigned long long* r, [[maybe_unused]] const bool own_totals) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = tid + k * THREADS;
            tv[k] = i < a.n_ranges ? __ldcg(r + i) : 0ull;
        }
        if constexpr (SHARDED) {
            const int g = tid >> 4, src = tid & 15;
            tc = 0ull;
            if (g < G && src < G) {
                if (!own_totals) tc = __ldcg(&slot_in->contrib[src][g]);     // (Mailbox::tot: K1 / in-kernel pre-pass)
                else if (src == 0) tc = __ldcg(&own->tot[g]);
            }
        }
    };
    if constexpr (SHARDED) {
        // every CTA: wait for the G slots of the previous launch, then finalize(t-1) and finish(t-2) on a private copy
        // of the control block (CTA 0 also writes the statistics rows; the last CTA copies the block back)
        for (int i = tid; i < (int)(sizeof(Ctrl) / 4); i += THREADS)
            reinterpret_cast<uint32_t*>(&s_sh.c)[i] = reinterpret_cast<const uint32_t*>(gctrl)[i];
        if (tid < kMaxShards) s_sh.contrib[tid] = 0ull;
        if (tid >= 32 && tid < 32 + G) s_sh.mail[tid - 32] = a.sh.ctx.mail[tid - 32];
        own = a.sh.own;
        const unsigned long long prev = a.sh.ctx.stamp - 1ull;
        slot_in = &own->fs[prev & 1ull];
        mailbox_wait(a.sh.first ? own->flag_rec : slot_in->flag, G, prev, &own->error, 4);
        K5_TRACE(1);
        // the range totals and the weight contributions travel while two warps finalise the step
        load_totals(a.sh.first ? a.rsum + (size_t)s_sh.c.rsum_sel * kRsumStride
                               : a.sh.own_rsum + (size_t)a.sh.rsel_in * kRsumStride, a.sh.first != 0);
        K2Args f = a.sh.pfin;
        f.ctrl = &s_sh.c; f.n_parts = G;
        if (a.sh.first) {
            f.parts = &own->rec[0][0]; f.rec_stride = 32;
            f.producer_rsum = f.producer_rsum && s_sh.c.uniform != 0;      // K1 delivered range totals iff the step was a bootstrap step
        } else {
            f.parts = &slot_in->rec[0][0]; f.rec_stride = 4;
        }
        if (blockIdx.x != 0) f.stats = nullptr;
        if (warp == 0) finalize_warp(f, s_sh.pout);
        if (warp == 1 && !a.sh.first && a.sh.pfinish.stats && !a.sh.pfinish.no_moments)
            finish_warp(&s_sh.c, a.sh.pfinish, &slot_in->mom[0][0], G, 16, blockIdx.x == 0 ? a.sh.pfinish.stats : nullptr, s_sh.m);
        __syncthreads();
        ctrl = &s_sh.c;
        K5_TRACE(2);
    }
    if (!ctrl->do_resample) {
        // (only a filter whose weights have all vanished gets here) nothing to resample: moment-shift refresh
        if constexpr (SHARDED) {
            // ... and an empty slot, so that no peer waits for this rank
            if (blockIdx.x == 0) {
                if (tid < G) {
                    FusedSlot* so = &a.sh.ctx.mail[tid]->fs[a.sh.ctx.stamp & 1ull];
                    so->rec[a.sh.ctx.rank][0] = -INFINITY; so->rec[a.sh.ctx.rank][1] = 0.0; so->rec[a.sh.ctx.rank][2] = 0.0;
                    for (int g = 0; g < G; ++g) so->contrib[a.sh.ctx.rank][g] = 0ull;
                    for (int k = 0; k < 16; ++k) so->mom[a.sh.ctx.rank][k] = 0.0;
                }
                if (tid == 0) *gctrl = s_sh.c;
                __threadfence_system();
                __syncthreads();
                if (tid < G) *(volatile unsigned long long*)&a.sh.ctx.mail[tid]->fs[a.sh.ctx.stamp & 1ull].flag[a.sh.ctx.rank] = a.sh.ctx.stamp;
            }
            return;
        }
        if (blockIdx.x == 0 && tid == 0 && a.fin.stats && !a.fin.no_moments)
            for (int d = 0; d < a.fin.nx && d < kMaxNx; ++d) ctrl->shift[d] = a.fin.stats[SMC_STATS_HEADER + d];
        return;
    }
    const int cur = ctrl->cur;
    const float* xsrc = cur ? a.x1 : a.x0;
    float* xdst = cur ? a.dst0 : a.dst1;
    const int sel = ctrl->rsum_sel;
    // range totals: single GPU = the workspace's two buffers (ctrl->rsum_sel: the one to read); sharded = the rank's
    // peer-mapped area (the first fused launch of a run still reads what K1 left in the workspace)
    const unsigned long long* rin = a.rsum + (size_t)sel * kRsumStride;
    unsigned long long* rout = a.rsum + (size_t)(sel ^ 1) * kRsumStride;
    [[maybe_unused]] float* const* px = nullptr;                // sharded: per-rank destination tables
    [[maybe_unused]] float* const* plw = nullptr;
    [[maybe_unused]] const uint32_t shard_n = SHARDED ? (uint32_t)a.sh.shard_n : 0u;
    if constexpr (SHARDED) {
        px = cur ? a.sh.peer_x0 : a.sh.peer_x1;
        plw = a.sh.peer_lw_out;
        if (!a.sh.first) rin = a.sh.own_rsum + (size_t)a.sh.rsel_in * kRsumStride;
        rout = a.sh.own_rsum + (size_t)a.sh.rsel_out * kRsumStride;
    }
    const float kshift_in = ctrl->kshift;
    const int64_t n_glob = SHARDED ? a.sh.n_global : a.n;
    const int64_t g_off = SHARDED ? a.sh.offset : 0;            // global index of this shard's first particle
    const uint32_t Nn = (uint32_t)n_glob;

    // this CTA's ranges (measured and dropped: handing the CTAs of one SM unequal shares by arrival order -- the three
    // co-resident CTAs finish 1 : 1.17 : 1.36 apart -- made every launch slower, profiles/r2_k5_shares.txt)
    int r_begin = (int)blockIdx.x * a.ranges_per_cta, r_end = r_begin + a.ranges_per_cta;
    int piece = (int)blockIdx.x;
    if (a.board) {
        __shared__ int s_piece;
        if (tid == 0) {
            const int occ = a.n_waves;
            unsigned int* taken = a.board + kBoardSms * 8;
            unsigned int* slot_ctr = taken + kBoardPieces;
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            volatile unsigned int* row = a.board + (size_t)(smid % (unsigned)kBoardSms) * 8;
            const unsigned arr = atomicAdd(const_cast<unsigned int*>(row), 1u);       // row[0]: arrivals on this SM
            int want = -1;
            if (arr < (unsigned)occ) {
                row[1 + arr] = blockIdx.x + 1u;                                       // row[1 .. occ]: their block indices
                const unsigned long long t0 = globaltimer_ns();
                bool all = false;
                int rank = 0;
                while (!all && globaltimer_ns() - t0 < (unsigned long long)a.board_wait_ns) {
                    all = true; rank = 0;
                    for (int j = 0; j < occ; ++j) {
                        const unsigned v = row[1 + j];
                        if (v == 0u) all = false;
                        else if (v - 1u < blockIdx.x) ++rank;
                    }
                }
                if (all) {
                    // the lowest block index of the SM draws the slot and posts it in row[7]
                    unsigned slot1 = 0u;
                    if (rank == 0) { slot1 = atomicAdd(slot_ctr, 1u) + 1u; row[7] = slot1; }
                    else while ((slot1 = row[7]) == 0u && globaltimer_ns() - t0 < 2ull * a.board_wait_ns) { }
                    if (slot1 != 0u && slot1 <= (unsigned)a.slots) want = (int)(slot1 - 1u) * occ + rank;
                }
            }
            // claim: the wanted piece, else the next free one (pieces and CTAs are equal in number)
            int pc = want >= 0 ? want : (int)blockIdx.x;
            const int n_pieces = a.slots * occ;
            while (atomicCAS(taken + pc, 0u, 1u) != 0u) pc = pc + 1 == n_pieces ? 0 : pc + 1;
            s_piece = pc;
        }
        __syncthreads();
        piece = s_piece;
        const int sl = piece / a.n_waves, w = piece % a.n_waves;
        int before = 0;
        for (int k = 0; k < w; ++k) before += a.share[k];
        r_begin = sl * (a.n_ranges / a.slots) + before;
        r_end = r_begin + a.share[w];
    }
    const int64_t t_begin = k3_range_begin(r_begin, a.n_tiles, a.n_ranges);
    const int64_t t_end = k3_range_begin(r_end, a.n_tiles, a.n_ranges);
    auto swz = [](int f) -> int { return f ^ ((f >> 3) & (UPT - 1)); };

    // log-weights / particles of one tile -> shared memory (cp.async; the ragged last tile is guarded).  Two
    // commit groups per tile, both issued after phase D of the tile before (issuing the log-weights right after
    // phase B measured 2.5 % slower); the particles are awaited only at S3.  `live` = false commits an empty group,
    // so that "all but the newest group" (cp.async.wait_group 1) always names the right data.
    auto fetch_lw = [&](int64_t tile, bool live) {
        const int64_t tb = tile * TILE;
        if (live && tb + TILE <= a.n) {
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int f = r * THREADS + tid;
                const unsigned dl = (unsigned)__cvta_generic_to_shared(&lw_s[swz(f)]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dl), "l"(a.lw_in + tb + (int64_t)f * 4) : "memory");
            }
        } else if (live) {
            const int n_in = (int)(a.n - tb);
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int f = r * THREADS + tid;
                float tmp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                if (f * 4 < n_in) {
                    load4<float>(a.lw_in + tb + f * 4, tmp);              // rows are padded to a multiple of 4
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (f * 4 + j >= n_in) tmp[j] = -INFINITY;
                }
                lw_s[swz(f)] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    auto fetch_x = [&](int64_t tile, bool live) {
        const int64_t tb = tile * TILE;
        if (live) {
            const int n_in = tb + TILE <= a.n ? TILE : (int)(a.n - tb);
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int f = r * THREADS + tid;
                if (f * 4 < n_in) {
#pragma unroll
                    for (int d = 0; d < NX; ++d) {
                        const unsigned dx = (unsigned)__cvta_generic_to_shared(x_s + d * TILE + swz(f) * 4);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dx),
                                     "l"(xsrc + d * a.ld + tb + (int64_t)f * 4) : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    fetch_lw(t_begin, t_begin < t_end);
    fetch_x(t_begin, t_begin < t_end);

    // exclusive CDF prefix of this CTA's first tile and the CDF total, from the range totals
    unsigned long long prefix, cdf_total;
    {
        if constexpr (!SHARDED) load_totals(rin, false);
        if constexpr (SHARDED) {
            if (!ctrl->rsum_valid) {
                // The a-priori bound left this step's weights without resolution (or K1 delivered no totals): range
                // pre-pass inside the kernel against the reference the finalize chose -- every CTA its own ranges, a
                // grid barrier (the grid is resident: 3 CTAs per SM), CTA 0 publishes the shard total
                unsigned long long* rt = const_cast<unsigned long long*>(rin);
                for (int r = r_begin; r < r_end; ++r) {
                    const int64_t tr0 = k3_range_begin(r, a.n_tiles, a.n_ranges), tr1 = k3_range_begin(r + 1, a.n_tiles, a.n_ranges);
                    unsigned long long acc = 0;
                    for (int64_t i = tr0 * TILE + tid; i < tr1 * TILE && i < a.n; i += THREADS)
                        acc += quant_a(quant_p(a.lw_in[i], kshift_in));
                    acc = warp_sum(acc);
                    __syncthreads();
                    if (lane == 0) s_warp_tot[warp] = acc;
                    __syncthreads();
                    if (tid == 0) {
                        unsigned long long t = 0;
#pragma unroll
                        for (int w = 0; w < WARPS; ++w) t += s_warp_tot[w];
                        rt[r] = t;
                    }
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    atomicAdd(a.sh.gbar, 1u);
                    while (*(volatile unsigned int*)a.sh.gbar < gridDim.x) __nanosleep(40);
                }
                __syncthreads();
                __threadfence();
                if (blockIdx.x == 0) {
                    unsigned long long t = 0;
                    for (int i = tid; i < a.n_ranges; i += THREADS) t += __ldcg(rt + i);
                    t = warp_sum(t);
                    __syncthreads();
                    if (lane == 0) s_warp_tot[warp] = t;
                    __syncthreads();
                    if (tid == 0) {
                        unsigned long long tot = 0;
#pragma unroll
                        for (int w = 0; w < WARPS; ++w) tot += s_warp_tot[w];
                        for (int g = 0; g < G; ++g) a.sh.ctx.mail[g]->tot[a.sh.ctx.rank] = tot;
                        __threadfence_system();
                        for (int g = 0; g < G; ++g)
                            *(volatile unsigned long long*)&a.sh.ctx.mail[g]->flag_tot[a.sh.ctx.rank] = a.sh.ctx.stamp;
                    }
                    __syncthreads();
                }
                mailbox_wait(own->flag_tot, G, a.sh.ctx.stamp, &own->error, 2);
                load_totals(rin, true);
            }
        }
        unsigned long long part = 0, all = 0;
        const int mine = r_begin;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (tid + k * THREADS < mine) part += tv[k];
            all += tv[k];
        }
        if constexpr (SHARDED) {
            // the CDF runs over ALL shards: offset = the total weight of the lower ranks; 16-lane sums of the
            // contribution table give the G shard totals
            all = 0;
            const int g = tid >> 4, src = tid & 15;
            unsigned long long v = tc;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
            if (src == 0) {
                if (g < a.sh.ctx.rank) part += v;
                all = v;
            }
        }
        part = warp_sum(part);
        all = warp_sum(all);
        if (lane == 0) { s_warp_tot[warp] = part; s_warp_all[warp] = all; }
        __syncthreads();
        prefix = 0; cdf_total = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { prefix += s_warp_tot[w]; cdf_total += s_warp_all[w]; }
        __syncthreads();
    }
    const RhoMap rm = make_rho(cdf_total, n_glob, 0xffffffffull - ctrl->uq);
    K5_TRACE(3);

    float shift[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) shift[d] = (float)ctrl->shift[d];
    double macc_d[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) macc_d[k] = 0.0;

    // ---- step t: observation, online (max, sum e, sum e^2) of its log-weights, range totals of its CDF ----
    float fz[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k) fz[k] = COPY ? 0.f : (float)a.z[k];
    float f_tm = -INFINITY, f_acc0 = 0.f, f_acc1 = 0.f;
    const float f_lw_uniform = (float)(-log((double)n_glob));
    // range cursor: the pre-pass range (k3_range_begin over n_ranges) that holds the thread's current quad
    const uint32_t cR = (uint32_t)a.n_ranges, cTd = (uint32_t)(a.n_tiles / a.n_ranges), cTm = (uint32_t)(a.n_tiles % a.n_ranges);
    uint32_t cur_r = 0xffffffffu, cur_bnd = 0u, cur_rem = 0u;
    unsigned long long cur_acc = 0ull;
    // sharded: destination cursor -- the rank whose index range [d_lo, d_hi) holds the thread's current quad (a
    // thread's quads have increasing global indices), and that rank's buffers
    [[maybe_unused]] uint32_t d_lo = 0u, d_hi = 0u;
    [[maybe_unused]] int d_rank = -1;
    float* xq = xdst;                    // destination particle buffer, log-weight array and range totals of the quad
    float* lwq_out = a.lw_out;
    unsigned long long* rq = rout;
    [[maybe_unused]] unsigned long long my_contrib = 0ull;    // sharded: weight this thread produced for rank d_rank
    [[maybe_unused]] bool wrote_peer = false;
    auto flush_acc = [&]() {
        if (cur_acc) {
            atomicAdd(rq + cur_r, cur_acc);
            if constexpr (SHARDED) my_contrib += cur_acc;
        }
        cur_acc = 0ull;
    };
    [[maybe_unused]] auto retarget = [&](const uint32_t gi0) {
        if (d_rank >= 0 && cur_r != 0xffffffffu) flush_acc();
        if constexpr (SHARDED) {
            if (d_rank >= 0 && my_contrib) atomicAdd(&s_sh.contrib[d_rank], my_contrib);     // (only threads that cross a shard boundary)
            my_contrib = 0ull;
        }
        d_rank = (int)(gi0 / shard_n);
        if constexpr (SHARDED) wrote_peer = wrote_peer || d_rank != a.sh.ctx.rank;
        d_lo = (uint32_t)d_rank * shard_n; d_hi = d_lo + shard_n;
        if constexpr (SHARDED) {
            xq = px[d_rank]; lwq_out = plw[d_rank];
            rq = a.sh.peer_rsum[d_rank] + (size_t)a.sh.rsel_out * kRsumStride;
        }
        cur_r = 0xffffffffu; cur_bnd = 0u; cur_acc = 0ull;
    };
    auto cursor_slow = [&](const uint32_t gi0) {                     // gi0: index within the destination shard
        if (cur_r == 0xffffffffu) {                                  // first quad of this thread: locate its range
            const unsigned long long tile = gi0 >> 12;
            const unsigned long long r = ((tile + 1ull) * cR - 1ull) / (unsigned long long)a.n_tiles;
            const unsigned long long base = r * (unsigned long long)a.n_tiles;
            const uint32_t beg = (uint32_t)(base / cR), rem = (uint32_t)(base % cR);
            const uint32_t s = rem + cTm, c = s >= cR ? 1u : 0u;
            cur_r = (uint32_t)r; cur_rem = s - (c ? cR : 0u); cur_bnd = (beg + cTd + c) << 12;
        } else {
            do {
                flush_acc();
                ++cur_r;
                const uint32_t s = cur_rem + cTm, c = s >= cR ? 1u : 0u;
                cur_rem = s - (c ? cR : 0u); cur_bnd += (cTd + c) << 12;
            } while (gi0 >= cur_bnd);
        }
    };
    // one quad of resampled particles (global indices gi0 .. gi0 + 3, gi0 % 4 == 0; `mask` = the lanes that
    // belong to this tile): exactly k1_chunk's arithmetic on exactly its Philox blocks, so the particles are
    // bit-identical to the three-launch step's
    auto process_quad = [&](auto full_tag, float (&xt)[NX][4], const uint32_t gi0, const unsigned mask) {
        constexpr bool FULL = decltype(full_tag)::value;           // all four lanes belong to this tile
        if constexpr (COPY) {
            if (FULL) {
#pragma unroll
                for (int d = 0; d < NX; ++d) store4<float>(xdst + d * a.ld + gi0, xt[d]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((mask >> j) & 1u) {
#pragma unroll
                        for (int d = 0; d < NX; ++d) xdst[d * a.ld + gi0 + j] = xt[d][j];
                    }
            }
        } else {
        uint32_t li0 = gi0;                                          // index within the destination shard
        if constexpr (SHARDED) {
            if (gi0 >= d_hi) retarget(gi0);
            li0 = gi0 - d_lo;
        }
        if (li0 >= cur_bnd) cursor_slow(li0);
        float nrm[NX][4];
        Draw<NX>::run(a.keys, (uint64_t)(gi0 >> 2), a.t, nrm);
        float lwq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float xv[NX], nv[NX];
#pragma unroll
            for (int d = 0; d < NX; ++d) { xv[d] = xt[d][j]; nv[d] = nrm[d][j]; }
            fm.propagate(xv, nv);
#pragma unroll
            for (int d = 0; d < NX; ++d) xt[d][j] = xv[d];
            float v = f_lw_uniform + fm.loglik(xv, fz);
            v = fmaxf(v, -INFINITY);                                 // NaN -> -inf (zero weight)
            lwq[j] = (FULL || ((mask >> j) & 1u)) ? v : -INFINITY;
        }
        if (FULL) {
#pragma unroll
            for (int d = 0; d < NX; ++d) store4<float>(xq + d * a.ld + li0, xt[d]);
            store4<float>(lwq_out + li0, lwq);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((mask >> j) & 1u) {
#pragma unroll
                    for (int d = 0; d < NX; ++d) xq[d * a.ld + li0 + j] = xt[d][j];
                    lwq_out[li0 + j] = lwq[j];
                }
        }
        const float cmax = fmaxf(fmaxf(lwq[0], lwq[1]), fmaxf(lwq[2], lwq[3]));
        if (cmax > f_tm) {
            const float sc = fast_exp(f_tm - cmax);                  // f_tm = -inf -> 0
            f_acc0 *= sc; f_acc1 *= sc * sc; f_tm = cmax;
        }
        if (f_tm != -INFINITY) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = fast_exp(lwq[j] - f_tm);             // masked lanes: exp(-inf) = 0
                f_acc0 += e; f_acc1 = fmaf(e, e, f_acc1);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cur_acc += quant_a(quant_p(lwq[j], a.kshift_bound));   // masked lanes: 0
        }
    };
    using FullQuad = K5Flag<true>;
    using PartQuad = K5Flag<false>;

    // ---- output window state (uniform over the CTA): window position i holds the output with global index
    // g0 + i (g0 % 4 == 0, so window quads are global quads); positions [lead, fill) are valid.  Phase E only
    // ever runs whole rounds of 256 quads -- every lane of every warp busy -- and leaves what is left over
    // (< 1024 + 3 outputs) in the window; it is moved to the front at the start of the next tile.
    uint32_t g0 = 0u, fill = 0u, lead = 0u, done4 = 0u;     // done4 = 4 x quads consumed by the last round (to be compacted)
    bool restart = true;                                    // the next tile (re)starts the window at its own E0
    const uint32_t win_sa = (uint32_t)__cvta_generic_to_shared(win);
    float macc[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) macc[k] = 0.f;
    auto add_moments = [&](const float (&xv)[NX]) {
        float xc[NX];
        int k = 0;
#pragma unroll
        for (int d = 0; d < NX; ++d) { xc[d] = xv[d] - shift[d]; macc[k] += xc[d]; ++k; }
#pragma unroll
        for (int d = 0; d < NX; ++d)
#pragma unroll
            for (int d2 = 0; d2 <= d; ++d2) { macc[k] = fmaf(xc[d], xc[d2], macc[k]); ++k; }
    };
    // one round of phase E over window quads [0, nq): moments of step t-1, functor of step t, stores.
    // `ragged`: quad 0 may start with `lead` foreign lanes and the last quad may be partial (masked path).
    constexpr int kConsumeUnroll = SMC_K5_UNROLL;
    auto consume = [&](const uint32_t nq, const bool ragged) {
#pragma unroll kConsumeUnroll
        for (uint32_t q = tid; q < nq; q += THREADS) {
            float xt[NX][4];
#pragma unroll
            for (int d = 0; d < NX; ++d)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(xt[d][0]), "=f"(xt[d][1]), "=f"(xt[d][2]), "=f"(xt[d][3])
                             : "r"(win_sa + (uint32_t)(d * WIN * 4) + 16u * q) : "memory");
            const uint32_t w0 = 4u * q;
            if (!((q == 0u && lead) || (ragged && w0 + 4u > fill))) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {                  // post-resample moments of step t-1
                    float xv[NX];
#pragma unroll
                    for (int d = 0; d < NX; ++d) xv[d] = xt[d][j];
                    add_moments(xv);
                }
                process_quad(FullQuad(), xt, g0 + w0, 15u);
            } else {
                unsigned mask = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool valid = w0 + j >= lead && w0 + j < fill;
                    if (valid) mask |= 1u << j;
#pragma unroll
                    for (int d = 0; d < NX; ++d) xt[d][j] = valid ? xt[d][j] : 0.f;
                    if (valid) {
                        float xv[NX];
#pragma unroll
                        for (int d = 0; d < NX; ++d) xv[d] = xt[d][j];
                        add_moments(xv);
                    }
                }
                if (mask) process_quad(PartQuad(), xt, g0 + w0, mask);
            }
        }
    };

    for (int64_t tile = t_begin; tile < t_end; ++tile) {
        const int64_t tile_base = tile * TILE;
        const unsigned long long tile_prefix = prefix;

        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();   // S1: lw_s complete; every thread has left the previous round of phase E

        // ---- compaction: what the last round left over moves to the front of the window ------------------
        if (done4) {
            const uint32_t nl4 = (fill - done4 + 3u) >> 2;     // float4 units to move (< 258)
            for (uint32_t i = tid; i < nl4; i += THREADS) {
                // (source and destination overlap by at most two floats, which the same thread reads, then writes)
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                    float4* wq = reinterpret_cast<float4*>(win + d * WIN);
                    const float4 v = wq[(done4 >> 2) + i];
                    wq[i] = v;
                }
            }
            g0 += done4; fill -= done4; lead = 0u; done4 = 0u;
        }

        // ---- phase B: fixed-point weights, thread / warp / CTA scan ------------------------------------
        float p[IPT];
        unsigned long long run = 0;
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
            const float4 u = lw_s[swz(UPT * tid + k)];
            p[4 * k + 0] = quant_p(u.x, kshift_in); p[4 * k + 1] = quant_p(u.y, kshift_in);
            p[4 * k + 2] = quant_p(u.z, kshift_in); p[4 * k + 3] = quant_p(u.w, kshift_in);
#pragma unroll
            for (int j = 0; j < 4; ++j) run += quant_a(p[4 * k + j]);
        }
        unsigned long long incl_w = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl_w, o);
            if (lane >= o) incl_w += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl_w;
        __syncthreads();   // S2: lw_s is dead
        const bool last = tile + 1 == t_end;
        unsigned long long warp_excl = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const unsigned long long t = s_warp_tot[w];
            if (w < warp) warp_excl += t;
            tile_total += t;
        }
        prefix = tile_prefix + tile_total;

        // ---- phase C: the thread's exact first / last boundary ---------------------------------------
        const unsigned long long base_raw = tile_prefix + warp_excl + (incl_w - run);
        const unsigned long long Ps = map_pos(base_raw, rm), Pe = map_pos(base_raw + run, rm);
        uint32_t e_start = umin((uint32_t)(Ps >> 32), Nn), e_end = umin((uint32_t)(Pe >> 32), Nn);
        int klast = IPT;               // first local item that owns the tail: cumsum[-1] = 1 (particle_filter.py:163)
        if (tile_base + TILE >= a.n && g_off + a.n >= n_glob) { // the last tile of the (global) population
            const int64_t j0 = tile_base + (int64_t)tid * IPT;
            if (j0 + IPT >= a.n) {
                e_end = Nn;
                klast = j0 >= a.n - 1 ? 0 : (int)(a.n - 1 - j0);
                if (j0 >= a.n) e_start = Nn;   // a thread that holds padding only starts where the last particle's range ends
            }
        }
        if (tid == 0) s_E[0] = e_start;
        if (tid == THREADS - 1) s_E[1] = e_end;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");        // this tile's particles
        const int heavy = __syncthreads_or(e_end - e_start > MAXD);   // S3: x_s complete
        const uint32_t E0 = s_E[0], E1 = s_E[1];
        const bool sane = E0 <= E1 && E1 <= Nn;
        const uint32_t count = sane ? E1 - E0 : 0u;
        if (restart) { g0 = E0 & ~3u; lead = E0 & 3u; fill = lead; restart = false; }
        const bool direct = !heavy && fill + count <= (uint32_t)WIN && E0 == g0 + fill;

        if (direct) {
            // ---- phase D: expansion -- one payload store per offspring ------------------------------
            if (count > 0u) {
                uint32_t e[IPT];
                f32_boundaries<IPT>(p, Ps, e_end, rm.rho_f, e);
                if (klast < IPT) {
#pragma unroll
                    for (int k = 0; k < IPT; ++k) if (k >= klast) e[k] = Nn;
                }
                // byte address of output index 0 in the window (boundaries are monotone and lie in [E0, E1], so
                // every store below falls into [fill, fill + count))
                const uint32_t wb = win_sa - 4u * g0;
                uint32_t a0 = wb + 4u * e_start, most = 0u;
#pragma unroll
                for (int k4 = 0; k4 < UPT; ++k4) {
                    float xk[NX][4];
#pragma unroll
                    for (int d = 0; d < NX; ++d) {
                        const float4 u = reinterpret_cast<const float4*>(x_s + d * TILE)[swz(UPT * tid + k4)];
                        xk[d][0] = u.x; xk[d][1] = u.y; xk[d][2] = u.z; xk[d][3] = u.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t a1 = wb + 4u * e[4 * k4 + j];
                        const uint32_t nb = a1 - a0;                  // 4 x offspring count of this item
                        most = umax(most, nb);
#pragma unroll
                        for (int d = 0; d < NX; ++d)                   // first and second offspring: predicated stores
                            asm volatile("{\n\t.reg .pred p1, p2;\n\t"
                                         "setp.ge.u32 p1, %0, 4;\n\tsetp.ge.u32 p2, %0, 8;\n\t"
                                         "@p1 st.shared.f32 [%1], %2;\n\t@p2 st.shared.f32 [%1+4], %2;\n\t}"
                                         :: "r"(nb), "r"(a0 + (uint32_t)(d * WIN * 4)), "f"(xk[d][j]) : "memory");
                        a0 = a1;
                    }
                }
                if (most >= 12u) {                                     // some item has three or more offspring (rare)
                    uint32_t b0 = wb + 4u * e_start;
#pragma unroll
                    for (int k = 0; k < IPT; ++k) {
                        const uint32_t b1 = wb + 4u * e[k];
                        if (b1 - b0 >= 12u) {
#pragma unroll
                            for (int d = 0; d < NX; ++d) {
                                const int li = IPT * tid + k;
                                const float xv = x_s[d * TILE + swz(li >> 2) * 4 + (li & 3)];
#pragma unroll 1
                                for (uint32_t bb = b0 + 8u; bb < b1; bb += 4u)
                                    asm volatile("st.shared.f32 [%0], %1;" :: "r"(bb + (uint32_t)(d * WIN * 4)), "f"(xv) : "memory");
                            }
                        }
                        b0 = b1;
                    }
                }
            }
            fill += count;
            __syncthreads();   // S4: the window holds the resampled outputs; x_s is dead
            fetch_lw(tile + 1, !last);                         // travels during phase E
            fetch_x(tile + 1, !last);                          // travels during phase E and the next tile's B, C

            // ---- phase E: whole rounds of 256 quads (the CTA's last tile: everything that is left) --------
            if (last) {
                consume((fill + 3u) >> 2, true);
            } else {
                const uint32_t nq = (fill >> 2) & ~255u;
                if (lead) consume(nq, true); else consume(nq, false);
                done4 = 4u * nq;
            }
        } else {
            // ---- skewed tile: first everything the window still holds (a ragged round), then exact boundaries,
            // head markers + max-scan fill window by window, straight to the destination buffer --------------
            if (fill > lead) consume((fill + 3u) >> 2, true);
            __syncthreads();
            restart = true; done4 = 0u; fill = 0u; lead = 0u;
            // (which boundaries a tile gets depends on the tile alone -- exact iff some thread has more than MAXD
            // outputs -- never on the window, so that this kernel and k3_resample always agree)
            uint32_t e[IPT];
            if (heavy) exact_boundaries<IPT>(p, base_raw, rm, e);
            else f32_boundaries<IPT>(p, Ps, e_end, rm.rho_f, e);
#pragma unroll
            for (int k = 0; k < IPT; ++k) e[k] = umin(k >= klast ? Nn : e[k], e_end);
            auto x_of = [&](const int d, const int li) -> float {      // item li of the (swizzled) particle tile
                return x_s[d * TILE + swz(li >> 2) * 4 + (li & 3)];
            };
            for (uint32_t wbase = 0; wbase < count; wbase += W) {
                {
                    uint4* o4 = reinterpret_cast<uint4*>(ids);
                    for (int i = tid; i < W * 2 / 16; i += THREADS) o4[i] = make_uint4(0, 0, 0, 0);
                }
                __syncthreads();
                {   // head markers (local index + 1) at the first output of every item with offspring
                    uint32_t r0 = umax(e_start, E0) - E0;
#pragma unroll
                    for (int k = 0; k < IPT; ++k) {
                        const uint32_t r1 = umin(umax(e[k], E0) - E0, count);
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
                        const int lo = wv[k] & 0xffff, hi = wv[k] >> 16;
                        tmax = max(tmax, max(lo, hi));
                    }
                    int inc = tmax;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc = max(inc, t);
                    }
                    int excl = __shfl_up_sync(0xffffffffu, inc, 1);
                    if (lane == 0) excl = 0;
                    if (lane == 31) s_wmax[warp] = inc;
                    __syncthreads();
                    int carry = excl;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) if (w < warp) carry = max(carry, s_wmax[w]);
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
                    const uint32_t gb = E0 + wbase;                          // global index of window slot 0
                    for (int i = tid; i < cnt; i += THREADS) {
                        int li1 = (int)ids[i];
                        if (li1 == 0) li1 = 1;                               // (only on inconsistent input)
                        float xv[NX];
                        float* xo = xdst + gb + i;
                        if constexpr (SHARDED) {
                            const uint32_t gi = gb + (uint32_t)i, r = gi / shard_n;
                            xo = px[r] + (gi - r * shard_n);
                            wrote_peer = wrote_peer || (int)r != a.sh.ctx.rank;
                        }
#pragma unroll
                        for (int d = 0; d < NX; ++d) {
                            xv[d] = x_of(d, li1 - 1);
                            xo[d * a.ld] = xv[d];
                        }
                        add_moments(xv);
                    }
                }
                __syncthreads();
            }
            // the window loop has written the resampled particles of this tile to the destination buffer (block-
            // visible after the barrier): run the functor over them in place
            if (!COPY && count > 0u) {
                const uint32_t gq0 = E0 & ~3u;
                const uint32_t nq = ((E1 + 3u) >> 2) - (gq0 >> 2);
                for (uint32_t q = tid; q < nq; q += THREADS) {
                    const uint32_t gi0 = gq0 + 4u * q;
                    unsigned mask = 0;
                    float xt[NX][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t gi = gi0 + j;
                        const bool valid = gi >= E0 && gi < E1;
                        if (valid) mask |= 1u << j;
                        const float* xi = xdst + gi;
                        if constexpr (SHARDED) {
                            const uint32_t r = gi0 / shard_n;             // (a quad never straddles two shards)
                            xi = px[r] + (gi - r * shard_n);
                        }
#pragma unroll
                        for (int d = 0; d < NX; ++d) xt[d][j] = valid ? (SHARDED ? __ldcg(xi + d * a.ld) : xi[d * a.ld]) : 0.f;
                    }
                    process_quad(PartQuad(), xt, gi0, mask);
                }
            }
            __syncthreads();
            fetch_lw(tile + 1, !last);
            fetch_x(tile + 1, !last);
        }
#pragma unroll
        for (int k = 0; k < NM; ++k) { macc_d[k] += (double)macc[k]; macc[k] = 0.f; }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    K5_TRACE(4);
#ifdef SMC_K5_TRACE
    if (SHARDED && threadIdx.x == 0) {     // outputs this CTA produced: CDF positions of its first and last tile
        const unsigned long long base = (((a.sh.ctx.stamp & 1ull) * 1024 + blockIdx.x) << 4);
        g_k5_trace[base + 8] = (unsigned long long)(t_end - t_begin);
        g_k5_trace[base + 9] = map_pos(prefix, rm) >> 32;          // (prefix = CDF at the end of the CTA's last tile)
    }
#endif
    if (!COPY && cur_r != 0xffffffffu) flush_acc();

    __syncthreads();                                         // the tile buffers are dead: reuse them as scratch
    if constexpr (SHARDED) {
        // The same per-CTA records as below in three barriers instead of eight (this tail is on the critical path of
        // every step of a strong-scaled filter): warp partials of everything, one barrier, the CTA sums, the ticket.
        const float wmax = warp_max(f_tm);
        if (lane == 0) s_fmax[warp] = wmax;
        if (a.mom_parts) {
#pragma unroll
            for (int k = 0; k < NM; ++k) {
                const double v = warp_sum(macc_d[k]);
                if (lane == 0) s_mom[warp][k] = v;
            }
        }
        // the CTA's share of the weight this rank produced for every rank's index range
        for (int g = 0; g < G; ++g) {
            const unsigned long long v = warp_sum(d_rank == g ? my_contrib : 0ull);
            if (lane == 0 && v) atomicAdd(&s_sh.contrib[g], v);
        }
        // every peer store / peer atomic of this CTA is made visible system-wide before the CTA takes its ticket (by
        // the threads that crossed a shard boundary; the rest of the CTA's stores are for the next launch on this GPU)
        if (wrote_peer) __threadfence_system();
        __syncthreads();
        float bmax = s_fmax[0];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) bmax = s_fmax[w] > bmax ? s_fmax[w] : bmax;
        const float sc = (f_tm == -INFINITY) ? 0.f : fast_exp(f_tm - bmax);
        const float v0 = warp_sum(f_acc0 * sc), v1 = warp_sum(f_acc1 * sc * sc);
        if (lane == 0) { s_fred[warp][0] = (double)v0; s_fred[warp][1] = (double)v1; }
        if (a.mom_parts && tid < NM) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += s_mom[w][tid];
            a.mom_parts[(int64_t)piece * NM + tid] = t;
            __threadfence();
        }
        if (tid >= 32 && tid < 32 + G && s_sh.contrib[tid - 32]) {
            atomicAdd(a.sh.own_rsum + 2 * kRsumStride + (tid - 32), s_sh.contrib[tid - 32]);
            __threadfence();
        }
        __syncthreads();
        K5_TRACE(5);
        if (tid == 0) {
            double t0 = 0, t1 = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { t0 += s_fred[w][0]; t1 += s_fred[w][1]; }
            double* out = a.parts + (int64_t)piece * 3;
            out[0] = (double)bmax; out[1] = t0; out[2] = t1;
            __threadfence();
            s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
            __threadfence();
        }
        __syncthreads();
        if (s_last) {
            // the last CTA publishes this launch's slot to every rank: record of step t's log-weights, post-resample
            // moments of step t-1, weight contributions; then the flag.  All loads first, then the reductions.
            const int rank = a.sh.ctx.rank;
            const unsigned long long slot = a.sh.ctx.stamp & 1ull;
            double acc[NM];
#pragma unroll
            for (int k = 0; k < NM; ++k) acc[k] = 0.0;
            if (a.mom_parts)
                for (int p = tid; p < (int)gridDim.x; p += THREADS)          // fixed record -> thread assignment
#pragma unroll
                    for (int k = 0; k < NM; ++k) acc[k] += __ldcg(a.mom_parts + (int64_t)p * NM + k);
            unsigned long long* acc_g = a.sh.own_rsum + 2 * kRsumStride;
            const unsigned long long cg = tid < G ? __ldcg(acc_g + tid) : 0ull;
            K5_TRACE(10);
            combine_partials(a.parts, (int)gridDim.x, 3, s_fout, s_ftmp, 0);
            K5_TRACE(11);
            if (tid < 3)
                for (int g = 0; g < G; ++g) s_sh.mail[g]->fs[slot].rec[rank][tid] = s_fout[tid];
            if (a.mom_parts) {
#pragma unroll
                for (int k = 0; k < NM; ++k) {
                    const double v = warp_sum(acc[k]);
                    if (lane == 0) s_mom[warp][k] = v;
                }
                __syncthreads();
                if (tid < NM) {
                    double t = 0;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) t += s_mom[w][tid];
                    for (int g = 0; g < G; ++g) s_sh.mail[g]->fs[slot].mom[rank][tid] = t;
                }
            }
            if (tid < G) {
                for (int g = 0; g < G; ++g) s_sh.mail[g]->fs[slot].contrib[rank][tid] = cg;
                acc_g[tid] = 0ull;
            }
            K5_TRACE(12);
            // every CTA of this rank has read the totals of step t-1: clear that buffer for the launch after the next
            // (the peers accumulate into it only once they have seen this launch's flag)
            if (!a.sh.first)
                for (int i = tid; i < kRsumStride; i += THREADS) const_cast<unsigned long long*>(rin)[i] = 0ull;
            if (a.board) for (int i = tid; i < kBoardSms * 8 + kBoardPieces + 8; i += THREADS) a.board[i] = 0u;
            if (tid == 0) {
                s_sh.c.cur ^= 1;                                  // the resampled, propagated set is the current one
                *a.sh.gbar = 0u;
                *a.ticket = 0u;
            }
            K5_TRACE(13);
            if (tid < 32) __threadfence_system();            // (every store into a mailbox came from warp 0)
            K5_TRACE(14);
            __syncthreads();
            if (tid < G) *(volatile unsigned long long*)&s_sh.mail[tid]->fs[slot].flag[rank] = a.sh.ctx.stamp;
            // (the control block is this GPU's own business: after the flags)
            for (int i = tid; i < (int)(sizeof(Ctrl) / 4); i += THREADS)
                reinterpret_cast<uint32_t*>(gctrl)[i] = reinterpret_cast<const uint32_t*>(&s_sh.c)[i];
            K5_TRACE(6);
        }
        return;
    }
    if (!COPY) {   // per-CTA (max, sum e, sum e^2) record of step t's log-weights (k1_propagate_weight's layout)
        const float wmax = warp_max(f_tm);
        if (lane == 0) s_fmax[warp] = wmax;
        __syncthreads();
        float bmax = s_fmax[0];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) bmax = s_fmax[w] > bmax ? s_fmax[w] : bmax;
        const float sc = (f_tm == -INFINITY) ? 0.f : fast_exp(f_tm - bmax);
        const float v0 = warp_sum(f_acc0 * sc), v1 = warp_sum(f_acc1 * sc * sc);
        if (lane == 0) { s_fred[warp][0] = (double)v0; s_fred[warp][1] = (double)v1; }
        __syncthreads();
        if (tid == 0) {
            double t0 = 0, t1 = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { t0 += s_fred[w][0]; t1 += s_fred[w][1]; }
            double* out = a.parts + (int64_t)piece * 3;
            out[0] = (double)bmax; out[1] = t0; out[2] = t1;
        }
    }
    // per-CTA post-resample moment partials of step t-1
    if (a.mom_parts) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NM; ++k) {
            const double v = warp_sum(macc_d[k]);
            if (lane == 0) s_mom[warp][k] = v;
        }
        __syncthreads();
        if (tid < NM) {
            double s = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) s += s_mom[w][tid];
            a.mom_parts[(int64_t)piece * NM + tid] = s;
        }
    }
    // the last CTA to arrive: finish of step t-1 (moments, buffer flip, shift refresh), hand-over of the
    // range-total buffers, finalize of step t (normaliser, ESS, decision, u, quantisation reference)
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        finish_body(ctrl, a.fin, s_mom, s_m);
        if (!COPY) {
            // every CTA has read the totals of step t-1: clear them for the launch after the next, and make the
            // buffer this launch has accumulated into the current one
            for (int i = tid; i < kRsumStride; i += THREADS) a.rsum[(size_t)sel * kRsumStride + i] = 0ull;
            if (tid == 0) ctrl->rsum_sel = sel ^ 1;
            __syncthreads();
            K2Args f2 = a.fin2;
            f2.ctrl = ctrl; f2.parts = a.parts; f2.n_parts = (int)gridDim.x;
            finalize_body(f2, s_fout, s_ftmp);
        }
        if (a.board) for (int i = tid; i < kBoardSms * 8 + kBoardPieces + 8; i += THREADS) a.board[i] = 0u;
        if (tid == 0) *a.ticket = 0u;
    }
}

}  // namespace smc
