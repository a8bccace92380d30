/*
 * Plain-C restatement of the reference's systematic resampler -- TEST
 * INFRASTRUCTURE ONLY (checker at large N and CPU-baseline timing; never linked
 * into the product library).
 *
 * Follows models/particle_filter.py:146-171 (ParticleFilter._systematic_resample)
 * and models/LEDH_particle_filter.py:25-37 (systematic_resample) of
 * liyiyang-amber/Particle_filters literally: sequential fp64 cumulative sum,
 * optional cumsum[-1] = 1.0, positions (u + i) / N, two-pointer merge.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/libsir_oracle.so oracle/csrc/sir_oracle.c
 * (-ffp-contract=off: no FMA contraction, so the arithmetic is the same IEEE
 * fp64 sequence NumPy performs.)
 */
#include <stdint.h>
#include <stdlib.h>

/* returns 0 on success, 1 on allocation failure */
int oracle_systematic_ancestors(const double *w, int64_t n, double u, int clamp_last,
                                int64_t *out)
{
    double *c = (double *)malloc((size_t)n * sizeof(double));
    if (!c) return 1;
    double acc = 0.0;
    for (int64_t k = 0; k < n; ++k) { acc += w[k]; c[k] = acc; }      /* np.cumsum, :162 */
    if (clamp_last && n > 0) c[n - 1] = 1.0;                           /* :163 */
    int64_t i = 0, j = 0;
    const double nd = (double)n;
    while (i < n) {                                                    /* :165-170 */
        double pos = (u + (double)i) / nd;                             /* :160 */
        if (j >= n) { out[i++] = n - 1; }                              /* guard (LEDH flavour) */
        else if (pos < c[j]) { out[i++] = j; }
        else { ++j; }
    }
    free(c);
    return 0;
}

/* Sequential fp64 cumulative sum (np.cumsum order), exposed for scan checks. */
void oracle_cumsum(const double *w, int64_t n, double *c)
{
    double acc = 0.0;
    for (int64_t k = 0; k < n; ++k) { acc += w[k]; c[k] = acc; }
}
