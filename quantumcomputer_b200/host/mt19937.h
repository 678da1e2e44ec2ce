/* mt19937.h -- the host-side random stream of the Shor driver.
 *
 * The reference draws its measurement variate with gsl_rng_uniform on
 * gsl_rng_mt19937 (qc_shor.c:281, 1296-1299).  GSL is not a dependency of this
 * repository; this is the published MT19937 recurrence with GSL's seeding
 * (seed 0 is replaced by 4357) and GSL's scaling (u32 / 2^32), so a given seed
 * yields the same r as the reference (KAT-3 of SURVEY.md).
 */
#ifndef QCS_HOST_MT19937_H
#define QCS_HOST_MT19937_H

#include <stdint.h>

typedef struct {
    uint32_t word[624];
    int next;
} qcsh_rng;

void qcsh_rng_seed(qcsh_rng *g, unsigned long seed);
uint32_t qcsh_rng_u32(qcsh_rng *g);
double qcsh_rng_uniform(qcsh_rng *g);   /* [0, 1), granularity 2^-32 */

#endif
