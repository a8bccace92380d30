/*
 * libsmc -- LEDH particle-flow entry points (path 2 of the hot path).
 *
 * Replaces the per-particle pseudo-time loop of LEDHFlowPF.step,
 * models/LEDH_particle_filter.py:132-179 of liyiyang-amber/Particle_filters, and the
 * weight update :186-195.  Conventions as in smc.h (SoA particle arrays, device
 * pointers named *_dev, host arrays row-major doubles, calls enqueue on `stream`).
 */
#ifndef SMC_LEDH_H_
#define SMC_LEDH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* measurement functors h(x), dh/dx evaluated on chip */
enum {
    SMC_LEDH_H_LINEAR = 1,   /* h = H x;          params: H (nz*nx row-major)                          */
    SMC_LEDH_H_ACOUSTIC = 2, /* z_s = sum_c psi / (|p_c - s|^2 + d0), state = n_targets x [x,y,vx,vy];
                                params: psi, d0, then nz sensor (x, y) pairs
                                (simulator_Multi_acoustic_tracking.py:273-310)                    */
    SMC_LEDH_H_EXP = 3       /* h_i = beta exp(x_i / 2), nz = nx; params: beta (stochastic volatility) */
};

typedef struct smc_ledh_desc {
    int64_t n;             /* particles                                                            */
    int64_t ld;            /* row stride of the SoA arrays (multiple of 4, >= n)                   */
    int32_t nx, nz;        /* 1 <= nx, nz <= 512: up to 32 the per-particle matrices live in registers /
                            * shared memory, beyond that in a global-memory workspace (smc_ledh_flow*)   */
    int32_t dtype;         /* SMC_F32 | SMC_F64                                                    */
    int32_t n_lambda_steps;
    void *eta0_dev;        /* in : nx rows, propagated particles eta_0            (:113-115)       */
    void *eta_dev;         /* in/out: nx rows, flowed particles eta (start = eta_0) (:119, :171)    */
    void *theta_log_dev;   /* in/out: n, running sum of log|det(I + dlam A)|       (:174-179)       */
    double *s0_dev;        /* out (nullable): n_lambda_steps * nz * nz, the matrix S of particle 0
                              at every pseudo-time step, from which the host forms the
                              condition-number diagnostic (:152-157)                               */
} smc_ledh_desc;

uint64_t smc_sizeof_ledh_desc(void);
/* Diagnostics of the fp32 information-form flow: how many flow calls needed a jittered Cholesky factor of P
 * (relative jitter <= 1e-7) and how many fell back to the literal kernel because P had no factor at all. */
void smc_ledh_info_counters(uint64_t *jittered, uint64_t *fell_back);

/* K4: all pseudo-time steps lambda_j = j / L, j = 1..L, for every particle, on chip.
 * P_host (nx*nx, already symmetrised, :106), R_host (nz*nz), z_host (nz). */
int smc_ledh_flow(const smc_ledh_desc *d, const double *P_host, const double *R_host,
                  const double *z_host, int32_t h_kind, const double *h_params_host,
                  int32_t n_h_params, void *stream);

/* One pseudo-time step with per-particle Jacobians H_dev (n * nz * nx, particle-major,
 * row-major inside) and h(eta) values h_dev (n * nz) supplied by the host-callback bridge
 * (opaque jacobian_h / h, :143-144).  step_index selects the s0_dev slot. */
int smc_ledh_flow_step_supplied(const smc_ledh_desc *d, const double *P_host, const double *R_host,
                                const double *z_host, const double *H_dev, const double *h_dev,
                                double lam, double dlam, int32_t step_index, void *stream);

/* Device-resident pieces for linear-Gaussian dynamics (spec model LedhGaussianSSM): the
 * reference evaluates these through N Python callbacks per step (:113-115, :187-189).
 *  propagate: eta0 = F x + v, v = Lq n with n ~ Philox(seed, t) when v_dev is NULL, else the
 *             caller's coloured noise (SoA, desc dtype); eta is initialised to eta0, theta_log to 0;
 *  terms    : log N(eta; F x, Q) + log N(z; h(eta), R) - log N(eta0; F x, Q) per particle
 *             (the additive constants of the two transition densities cancel, the likelihood
 *             constant is common to all particles and drops out in the normalisation).
 * x_dev: previous particles, SoA like eta (nx rows of ld).  Qinv/Rinv: inverse covariances. */
int smc_ledh_propagate_linear(const smc_ledh_desc *d, const void *x_dev, const double *F_host,
                              const double *Lq_host, const void *v_dev, uint64_t seed, int64_t t,
                              void *stream);
int smc_ledh_gauss_terms(const smc_ledh_desc *d, const void *x_dev, const double *F_host,
                         const double *Qinv_host, const double *Rinv_host, const double *z_host,
                         int32_t h_kind, const double *h_params_host, int32_t n_h_params,
                         double *terms_dev, void *stream);

/* Weight update of LEDHFlowPF.step (:186-190):
 *   logw = log(w_prev + 1e-300) + theta_log + terms,
 * terms = log p(x_k|x_{k-1}) + log p(z|x_k) - log p(eta_0|x_{k-1}) per particle (fp64, device).
 * Fills d->logw_dev and the block partials; continue with smc_sir_finalize /
 * smc_sir_resample / smc_sir_finish_step on the same descriptor (which must carry
 * SMC_FLAG_GENERIC_MOMENTS).  wprev_dev: n linear fp64 weights; theta_log_dev: desc dtype. */
struct smc_sir_desc;
int smc_ledh_weights(const struct smc_sir_desc *d, const double *wprev_dev, const void *theta_log_dev,
                     const double *terms_dev, void *stream);

/* LEDHFlowPF._weighted_stats (:217-224): mean and covariance of the current particle buffer
 * under explicit linear weights w_dev (fp64, need not be normalised; NULL = uniform), written
 * to stats_row_dev in the smc.h row layout. */
int smc_weighted_moments(const struct smc_sir_desc *d, const double *w_dev, double *stats_row_dev,
                         void *stream);

/* EDH flow (models/EDH_particle_filter.py:218-275).  The EDH flow field A(lambda) x + b(lambda) is
 * evaluated at the tracker's mean path, not per particle, so the caller composes the L Euler / RK4
 * pseudo-time sub-steps into one affine map (M, c) on the host (nx x nx, tiny) and this call applies
 * x <- M x + c to all n particles in place (SoA x[d * ld + i], dtype as in smc.h, fp64 arithmetic).
 * Weights, ESS, resampling and moments of the EDH step run on smc_ledh_weights (theta = 0) and the
 * SIR kernels. */
int smc_affine_map(void *x_dev, const double *M_host, const double *c_host, int64_t n, int64_t ld,
                   int32_t nx, int32_t dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SMC_LEDH_H_ */
