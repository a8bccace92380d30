/*
 * libsmc -- C-ABI of the B200-native Sequential-Monte-Carlo engine.
 *
 * This is the drop-in boundary beneath the reference's Python filter classes
 * (liyiyang-amber/Particle_filters).  The reference has no FFI of its own: the
 * public interface is the class surface of models/particle_filter.py and
 * models/LEDH_particle_filter.py.  Every entry point below names the reference
 * method (file:line) whose arithmetic it replaces; INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a CUDA device pointer,
 *     everything else is host memory;
 *   - every call enqueues work on the caller's stream (cudaStream_t passed as
 *     void*; NULL = legacy default stream) and returns without synchronising;
 *   - return value 0 = ok, non-zero = error, text via smc_last_error();
 *   - particles are stored SoA: x[d * ld + i], d < nx, i < n, ld % 4 == 0 and
 *     ld >= round_up(n, 4); log-weight arrays hold round_up(n, 4) elements;
 *   - dtype selects the storage/arithmetic type of particles and log-weights
 *     (SMC_F32 throughput mode, SMC_F64 parity mode).  Normalisers, ESS, moments
 *     are always accumulated in fp64 and the resampling CDF in 64-bit fixed
 *     point, so results do not depend on launch geometry or shard count.
 */
#ifndef SMC_H_
#define SMC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMC_ABI_VERSION 2

enum { SMC_F32 = 0, SMC_F64 = 1 };

/* State-space-model functors compiled into the fused kernel (K1). */
enum {
    SMC_MODEL_STOCH_VOL = 1,       /* g = alpha x, h = beta exp(x/2); nx = nz = 1           */
    SMC_MODEL_LINEAR_GAUSSIAN = 2, /* g = A x, h = H x; nx, nz in {1, 2, 4}                 */
    SMC_MODEL_LORENZ96 = 3         /* g = n_sub RK4 steps of Lorenz-96, h = x[obs_idx]      */
};

/* smc_sir_propagate_weight mode bits */
enum { SMC_MODE_PROPAGATE = 1, SMC_MODE_WEIGHT = 2, SMC_MODE_STEP = 3 };

/* smc_sir_desc.flags */
enum {
    SMC_FLAG_WRITE_ANCESTORS = 1, /* store int32 ancestor indices (additive trace output)   */
    SMC_FLAG_NO_MOMENTS = 2,      /* skip mean/cov (not used by the reference-equivalent path) */
    SMC_FLAG_GENERIC_MOMENTS = 4, /* block partials carry no moments; mean/cov come from the
                                     any-nx moment kernel (bridge mode, nx > 4)               */
    SMC_FLAG_POST_MOMENTS_GENERIC = 8 /* post-resample moments are recomputed from memory by
                                     smc_sir_finish_step (set when the resampled set is
                                     jittered or gathered outside K3)                         */
};

/* One shard of a particle population (the whole population when n == n_global). */
typedef struct smc_sir_desc {
    int64_t n;          /* particles held by this shard                                    */
    int64_t n_global;   /* particles over all shards (<= 2^30)                             */
    int64_t offset;     /* global index of local particle 0 (multiple of 4)                */
    int64_t ld;         /* row stride of x0/x1 in elements                                 */
    int32_t nx, nz;
    int32_t dtype;      /* SMC_F32 | SMC_F64                                               */
    int32_t flags;
    void *x0_dev, *x1_dev; /* ping-pong particle buffers (ctrl.cur says which is current)  */
    void *logw_dev;     /* un-normalised log-weights of the current step                   */
    int32_t *ancestors_dev; /* n int32 or NULL                                             */
    void *ws_dev;       /* workspace of smc_sir_workspace_bytes()                          */
    uint64_t ws_bytes;
} smc_sir_desc;

/* Per-step statistics row written by the device (doubles):
 *   [0] log normaliser  m + log sum exp(logw - m)     (particle_filter.py:262-263)
 *   [1] ESS before resampling  1 / sum w^2            (particle_filter.py:203)
 *   [2] 1.0 if the step resampled else 0.0            (particle_filter.py:204)
 *   [3] the systematic-resampling uniform u           (particle_filter.py:160)
 *   [4 .. 4+nx)            posterior mean             (particle_filter.py:266)
 *   [4+nx .. 4+nx+nx*nx)   posterior covariance       (particle_filter.py:267)
 */
#define SMC_STATS_HEADER 4
static inline int64_t smc_stats_stride(int nx) { return SMC_STATS_HEADER + nx + (int64_t)nx * nx; }

int smc_abi_version(void);
/* Number of kernel launches of the SIR step kernels (K1, range pre-pass, K3, K5) this process has issued
 * so far; bench.py reports the difference over its timed region as `gpu_launches`. */
uint64_t smc_launch_count(void);
const char *smc_last_error(void);
uint64_t smc_sizeof_sir_desc(void);
uint64_t smc_sizeof_shard_ctx(void);     /* sizeof(smc_shard_ctx), for binding self-checks */

/* Workspace size for a shard (control block, block partials, pre-pass range totals, moment partials). */
uint64_t smc_sir_workspace_bytes(int64_t n, int32_t nx, int32_t dtype);

/* Reset the control block: current buffer = x0, weights uniform, moment shift =
 * prior mean (host pointer, nx doubles, may be NULL = 0). */
int smc_sir_reset(const smc_sir_desc *d, const double *shift_host, void *stream);

/* ParticleFilter.initialize (particle_filter.py:110-132): x = mean + Lc n with
 * n ~ N(0, I) from Philox stream (seed, t = 0); weights uniform.  Lc_host is
 * the nx*nx row-major lower Cholesky factor, mean_host nx doubles. */
int smc_sir_init_gaussian(const smc_sir_desc *d, const double *mean_host, const double *Lc_host,
                          uint64_t seed, void *stream);

/* K1: fused ParticleFilter.predict + the likelihood half of update
 * (particle_filter.py:223-237 and :253-261).  params_host holds the functor's
 * parameters (layout per model in DESIGN.md), z_dev the nz-vector observation
 * (doubles, device).  noise_dev: NULL = on-chip Philox4x32-10 keyed by
 * (seed, t, global particle id); otherwise SoA standard normals (same dtype,
 * same ld) -- the "reference-noise mode" used for parity.  Also emits per-block
 * (max, sum exp, sum exp^2, weighted moment) partials into the workspace. */
int smc_sir_propagate_weight(const smc_sir_desc *d, int32_t model_id, const double *params_host,
                             int32_t n_params, int32_t mode, const double *z_dev,
                             const void *noise_dev, uint64_t seed, int64_t t, void *stream);

/* Host-callback bridge (opaque Python g/h evaluated by the caller exactly as the
 * reference does; the engine's own arithmetic still runs here).
 *  predict: x = gx + noise @ Lq^T                      (particle_filter.py:236-237)
 *  loglik : logw = log(w_prev + 1e-300) - 0.5 |LR^-1 (z - zpred)|^2   (:258-261)
 * gx_dev/noise_dev/zpred_dev are SoA arrays in the desc dtype (zpred: nz rows
 * of ld); Lq_host / LRinv_host are row-major lower-triangular factors;
 * wprev_dev = linear weights of the previous step or NULL for uniform.
 * target_other = 0 writes the current buffer; 1 writes the other ping-pong buffer
 * (the resampling destination) and then gx_dev may be NULL = "that buffer itself",
 * which is the post-resample regularisation x += noise @ (scale Lq)^T (:212-218). */
int smc_bridge_predict(const smc_sir_desc *d, const void *gx_dev, const void *noise_dev,
                       const double *Lq_host, double scale, int32_t target_other, void *stream);
int smc_bridge_loglik(const smc_sir_desc *d, const void *zpred_dev, const double *z_host,
                      const double *LRinv_host, const void *wprev_dev, void *stream);

/* Post-resample regularisation on chip (particle_filter.py:212-218): x += n (scale Lq)^T with
 * n ~ N(0, I) from the Philox jitter stream (seed, t), applied to the buffer K3 has just filled iff
 * the control block says the step resampled.  Call between smc_sir_resample and smc_sir_finish_step
 * on a descriptor with SMC_FLAG_POST_MOMENTS_GENERIC (the moments are then taken after the jitter). */
int smc_sir_jitter(const smc_sir_desc *d, const double *Lq_host, double scale, uint64_t seed, int64_t t,
                   void *stream);

/* K2: combine block partials -> log normaliser, ESS, weighted mean/cov; decide
 * resampling (ESS < resample_thresh * n_global, particle_filter.py:203-204);
 * fix the systematic uniform u (u_host in [0,1) = caller's draw for
 * reference-noise mode, negative = Philox(seed, t)).  stats_row_dev receives
 * smc_stats_stride(nx) doubles.  n_parts_override > 0: reduce that many
 * partials found at parts_dev instead of this shard's own (multi-GPU: the
 * all-gathered per-rank partials, see smc_sir_reduce_partials). */
int smc_sir_finalize(const smc_sir_desc *d, double resample_thresh, double u_host, uint64_t seed,
                     int64_t t, double *stats_row_dev, const double *parts_dev,
                     int32_t n_parts_override, void *stream);

/* Replace the systematic uniform chosen by smc_sir_finalize (reference-noise mode:
 * the caller draws rng.random() only after learning that the step resamples, to
 * keep the reference's RNG call order, particle_filter.py:204-206 then :160). */
int smc_sir_set_uniform(const smc_sir_desc *d, double u, void *stream);

/* Multi-GPU helper: fold this shard's block partials into ONE partial record of
 * smc_partial_stride(nx) doubles at out_dev (to be all-gathered). */
int64_t smc_partial_stride(int32_t nx);
int smc_sir_reduce_partials(const smc_sir_desc *d, double *out_dev, void *stream);

/* K3: systematic resampling (particle_filter.py:146-171 + gather :209-210;
 * LEDH_particle_filter.py:25-37).  A range pre-pass sums the 64-bit fixed-point weights of contiguous
 * tile ranges; the persistent resampler then carries the exact CDF prefix from tile to tile, finds the
 * output boundaries in closed form (fp64 mode: exact tie handling against NumPy's rounding of
 * (u + i) / N; fp32 mode: positions laid over the actual fixed-point total), expands them through a
 * shared-memory marker window and gathers the payload into the other ping-pong buffer, accumulating
 * the post-resample moments.  A no-op when the control
 * block says "do not resample".  cdf_offset_fix / total are used by sharded
 * runs (exclusive prefix of lower ranks' weight totals in 2^-32 index units,
 * see smc_sir_shard_weight_total); pass 0 for a single shard.  peer_x_dev:
 * NULL, or n_shards device pointers (each shard's destination buffer for this
 * step, peer-mapped) for fused cross-shard migration. */
int smc_sir_resample(const smc_sir_desc *d, uint64_t cdf_offset_fix, double *stats_row_dev,
                     void *const *peer_x_dev, int32_t n_shards, void *stream);

/* Sharded form of smc_sir_resample: the CDF offset of this shard is the sum of the lower
 * ranks' entries of totals_dev (n_shards uint64 values, the all-gathered results of
 * smc_sir_shard_weight_total), read on the device so that no host synchronisation is needed.
 * peer_x_dev: device array of 2*n_shards pointers [buffer 0 of every rank][buffer 1 of every
 * rank] (peer-mapped, see smc_ipc_*); every output is stored straight into the buffer of the
 * rank that owns its index (NVLink P2P stores) -- the gather and the cross-shard migration are
 * one kernel. */
int smc_sir_resample_sharded(const smc_sir_desc *d, const uint64_t *totals_dev, int32_t rank,
                             int32_t n_shards, void *const *peer_x_dev, void *stream);
/* Fold this shard's per-tile post-resample moment partials into one record
 * (nx + nx(nx+1)/2 doubles) at out_dev, to be all-gathered and passed to smc_sir_finish_step. */
int smc_sir_reduce_moment_partials(const smc_sir_desc *d, double *out_dev, void *stream);

/* Raw device allocations that can be shared between the ranks of one node (CUDA IPC). */
int smc_dev_alloc(uint64_t bytes, void **ptr_out);
int smc_dev_free(void *ptr);
int smc_ipc_export(void *ptr, unsigned char handle_out[64]);
int smc_ipc_import(const unsigned char handle[64], void **ptr_out);
int smc_ipc_close(void *ptr);

/* Fixed-point total of this shard's normalised weights (index units * 2^32),
 * written to total_dev (one uint64); input to the cross-shard exclusive scan. */
int smc_sir_shard_weight_total(const smc_sir_desc *d, uint64_t *total_dev, void *stream);

/* Finish a step: publish post-resample moments into stats_row_dev, flip the
 * current buffer if the step resampled, refresh the moment shift.
 * mom_parts_dev/n_mom_parts: NULL/0 = this shard's own tile partials. */
int smc_sir_finish_step(const smc_sir_desc *d, double *stats_row_dev, const double *mom_parts_dev,
                        int32_t n_mom_parts, void *stream);

/* Whole SIR step on one shard = propagate_weight + finalize + resample +
 * finish_step (ParticleFilter.step, particle_filter.py:271-287). */
int smc_sir_step(const smc_sir_desc *d, int32_t model_id, const double *params_host,
                 int32_t n_params, const double *z_dev, const void *noise_dev,
                 double resample_thresh, double u_host, uint64_t seed, int64_t t,
                 double *stats_row_dev, void *stream);

/* The two halves of smc_sir_step, for callers that want to time or interleave them:
 *  _weight  : K1 with the K2 finalize folded into its last CTA (one launch);
 *  _resample: phases bit 1 = range pre-pass, bit 2 = K3 with the finish step folded
 *             into its last CTA; a no-op on the data when the control block says "do not resample". */
int smc_sir_step_weight(const smc_sir_desc *d, int32_t model_id, const double *params_host,
                        int32_t n_params, const double *z_dev, const void *noise_dev,
                        double resample_thresh, double u_host, uint64_t seed, int64_t t,
                        double *stats_row_dev, void *stream);
int smc_sir_step_resample(const smc_sir_desc *d, int32_t phases, double *stats_row_dev, void *stream);

/* Fused "resample step t-1, then propagate + weigh step t" for back-to-back runs of filters that
 * resample every step (resample_thresh > 1), fp32, single GPU, nx <= 2 spec models (kernel K5,
 * csrc/fused_step.cuh; reference particle_filter.py:146-171, :209-210, :266-268 of step t-1 and
 * :223-237, :253-263 of step t): the resampler expands the tile's offspring into shared memory and hands
 * every resampled particle straight to the model functor, so the resampled set never travels through HBM
 * -- 16 B per particle-update (R logw, R x, W x, W logw) in ONE launch.  There is no range pre-pass: the
 * kernel that writes the log-weights of step t also adds their fixed-point values (quantised against the
 * a-priori bound lw <= -log N of a bootstrap step) into the per-range totals the next launch takes its CDF
 * prefixes from.  Should the bound leave a step's weights too little resolution (mean likelihood below
 * 2^-16), the last CTA says so in the control block and the pre-pass of `phases` bit 2 -- which otherwise
 * returns at once -- recomputes the totals against the normaliser.
 * Precondition: step t-1 has been weighed and finalized (smc_sir_step_weight or a previous fused call)
 * and its range totals exist (delivered by those calls, or by smc_sir_step_resample phase 1).  Reads
 * d->logw_dev, writes the log-weights of step t to logw_out_dev (a second array; swap them for the next
 * call), the moments of step t-1 to stats_prev_dev and the header of step t to stats_row_dev.  Finish
 * the run with smc_sir_step_resample(d', 2, last_row) on the descriptor holding the last weights.
 * phases: bit 1 = the fused kernel, bit 2 = the conditional range pre-pass of step t (3 = both).
 * SMC_FUSED_LEGACY=1 in the environment selects round 1's pipeline (k3_resample with the functor in its
 * copy-out + an unconditional pre-pass, 20 B per particle-update) for A/B measurements. */
int smc_sir_step_fused(const smc_sir_desc *d, void *logw_out_dev, int32_t model_id,
                       const double *params_host, int32_t n_params, const double *z_dev,
                       double resample_thresh, uint64_t seed, int64_t t, double *stats_prev_dev,
                       double *stats_row_dev, int32_t phases, void *stream);

/* Multi-GPU step without collective calls.  Every rank owns a mailbox (smc_mailbox_bytes() of
 * zeroed, peer-mapped device memory, see smc_ipc_*).  The last CTA of K1 / the pre-pass / K3 stores
 * this rank's partial record / weight total / moment record straight into the mailbox of EVERY rank
 * (NVLink peer stores) and raises a stamped flag there; the consuming kernels (finalize, K3's
 * prologue, finish) spin on the flags in their own memory.  mail_dev: device array of n_shards
 * mailbox pointers (index = rank, own mailbox included); peer_x_dev: as for
 * smc_sir_resample_sharded; stamp: positive, increased by one per step, identical on all ranks.
 * Results are bit-identical to the single-GPU step (same reduction orders, integer CDF). */
typedef struct smc_shard_ctx {
    void *const *mail_dev;
    void *const *peer_x_dev;
    int32_t rank, n_shards;
    uint64_t stamp;
    /* fused sharded step only (NULL otherwise): device table [log-weight array 0 of every rank][log-weight
     * array 1 of every rank][range-total area of every rank]; a range-total area is smc_shard_rsum_bytes() of
     * zeroed peer-mapped memory; rsum_area_dev / mailbox_dev = this rank's own area / mailbox */
    void *const *peer_aux_dev;
    void *rsum_area_dev;
    void *mailbox_dev;            /* this rank's own mailbox (= mail_dev[rank]) */
} smc_shard_ctx;
uint64_t smc_mailbox_bytes(void);
/* Byte offset of the mailbox's error word: non-zero after a flag wait on this rank timed out (20 s) because a
 * peer never published -- bit 63 | stamp << 8 | exchange (1 record, 2 total, 3 moments) << 4 | peer rank.  The
 * kernels then run on with stale slots instead of hanging; the host checks the word (ShardedSir.check). */
uint64_t smc_mailbox_error_offset(void);
int smc_sir_step_sharded(const smc_sir_desc *d, const smc_shard_ctx *ctx, int32_t model_id,
                         const double *params_host, int32_t n_params, const double *z_dev,
                         double resample_thresh, uint64_t seed, int64_t t, double *stats_row_dev,
                         void *stream);

/* Sharded form of smc_sir_step_fused: ONE launch and ONE exchange per step of a bootstrap filter (fp32,
 * STOCH_VOL, shards of >= 12 x SM-count tiles), 16 B per particle-update.  A run is
 *   smc_sir_weight_sharded            stamp s      K1 of the first step (publishes record + weight total)
 *   smc_sir_step_fused_sharded x k    stamp s+1..  resample(t-1) -> propagate + weigh(t); first = 1 for the
 *                                                  launch that follows K1; d->logw_dev = the array holding step
 *                                                  t-1's log-weights, lw_out_sel = the OTHER array's index
 *   smc_sir_fused_sharded_flush       stamp s+k+1  finalize of the last step, finish of the one before; with
 *                                                  resample != 0 also the last step's resampling (pre-pass,
 *                                                  resampler with migration, finish)
 * Every launch's prologue waits (bounded) for the slot all ranks published at the end of the launch before,
 * finalises step t-1 and finishes step t-2 in every CTA (identical arithmetic everywhere; CTA 0 writes
 * stats_prev_dev = row t-1 and the moments of stats_prev2_dev = row t-2); resampled quads go through the
 * functor on the rank that holds their ancestor and are stored -- particle, log-weight, fixed-point weight
 * into the range totals -- into the memory of the rank that owns their index.  The particle sets are
 * bit-identical to the single-GPU engine's (tests/sharded_check.py). */
uint64_t smc_shard_rsum_bytes(void);
int smc_sir_weight_sharded(const smc_sir_desc *d, const smc_shard_ctx *ctx, int32_t model_id,
                           const double *params_host, int32_t n_params, const double *z_dev,
                           double resample_thresh, uint64_t seed, int64_t t, void *stream);
int smc_sir_step_fused_sharded(const smc_sir_desc *d, const smc_shard_ctx *ctx, int32_t lw_out_sel,
                               int32_t first, int32_t model_id, const double *params_host,
                               int32_t n_params, const double *z_dev, double resample_thresh,
                               uint64_t seed, int64_t t, double *stats_prev2_dev,
                               double *stats_prev_dev, void *stream);
int smc_sir_fused_sharded_flush(const smc_sir_desc *d, const smc_shard_ctx *ctx, double resample_thresh,
                                uint64_t seed, int64_t t_last, double *stats_prev2_dev,
                                double *stats_prev_dev, int32_t resample, void *stream);

/* ParticleFilter.step(z) (particle_filter.py:271-287) in ONE call for spec models with on-chip noise: copies the
 * observation z_host (nz doubles, pinned or pageable) to z_dev, enqueues smc_sir_step, copies the statistics row
 * (n_row doubles) to row_host and synchronises the stream -- the class API's per-step host <-> device traffic
 * without three round trips through the Python binding. */
int smc_sir_step_host(const smc_sir_desc *d, int32_t model_id, const double *params_host,
                      int32_t n_params, double *z_dev, const double *z_host, double resample_thresh,
                      double u_host, uint64_t seed, int64_t t, double *stats_row_dev,
                      double *row_host, int32_t n_row, void *stream);

/* Stand-alone pieces used by the drop-in classes' private helpers. */

/* ParticleFilter._systematic_resample(weights) (particle_filter.py:146-171):
 * w_dev = n normalised fp64 weights, ancestors_dev = n int32 out. */
int smc_systematic_ancestors(const double *w_dev, int64_t n, double u, int32_t *ancestors_dev,
                             void *ws_dev, uint64_t ws_bytes, void *stream);
/* ParticleFilter._multinomial_resample(weights) (particle_filter.py:173-186), i.e. NumPy's
 * rng.choice(n, size=n_out, p=w) = searchsorted(cumsum(w)/sum(w), u, side='right'):
 * ancestors[i] = #{ j : cdf_j <= u_i }.  w_dev = n fp64 probabilities (16-byte aligned); u_dev =
 * n_out fp64 uniforms in [0,1) (the host's rng.random(n_out): reference-noise mode) or NULL for
 * on-chip Philox uniforms keyed by (seed, t, i).  The CDF is an exact 2^-62 fixed-point prefix
 * sum and the comparison is exact (128-bit), so results do not depend on the launch shape.
 * ws_dev: smc_multinomial_workspace_bytes(n) bytes, 256-byte aligned. */
uint64_t smc_multinomial_workspace_bytes(int64_t n);
int smc_multinomial_ancestors(const double *w_dev, int64_t n, int64_t n_out, const double *u_dev,
                              uint64_t seed, int64_t t, int32_t *ancestors_dev, void *ws_dev,
                              uint64_t ws_bytes, void *stream);
/* particles[idx] (particle_filter.py:209): dst[d*ld+i] = src[d*ld+idx[i]]. */
int smc_gather(const void *src_dev, void *dst_dev, const int32_t *idx_dev, int64_t n, int64_t ld,
               int32_t nx, int32_t dtype, void *stream);
/* w = exp(logw - log_norm) as fp64 (particle_filter.py:263); reads the control
 * block's log normaliser; if the current weights are uniform writes 1/n_global. */
int smc_sir_export_weights(const smc_sir_desc *d, double *w_dev, void *stream);
/* Which ping-pong buffer is current (synchronises the stream). */
int smc_sir_current_buffer(const smc_sir_desc *d, int32_t *cur_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SMC_H_ */
