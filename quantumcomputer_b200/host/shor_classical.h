/* shor_classical.h -- the classical half of Shor's algorithm as the reference
 * runs it around the gate path (qc_shor.c:756-1134), host C, no device work
 * except through the qcs_* calls made by qcsh_find_period.
 *
 * Two behaviours are offered:
 *   QCSH_VERBATIM  reproduces the reference's arithmetic, including INT_POW's
 *                  wrap-around (qc_shor.c:158-159) and the continued-fraction
 *                  casts (qc_shor.c:822-828), with the out-of-range conversions
 *                  resolved as on x86-64.  `period_found` starts false (the
 *                  reference leaves it uninitialised).
 *   QCSH_ROBUST    the same algorithm without the undefined behaviour: modular
 *                  powers, expansion stops when the fraction terminates, the
 *                  a^(p/2) = -1 test uses the trial integer.
 */
#ifndef QCS_HOST_SHOR_CLASSICAL_H
#define QCS_HOST_SHOR_CLASSICAL_H

#include <stdint.h>

#include "mt19937.h"

struct qcs_register;

enum { QCSH_VERBATIM = 0, QCSH_ROBUST = 1 };

#define QCSH_NUM_CONTINUED_FRACTIONS 15   /* qc_shor.c:121 */
#define QCSH_TRIALS_PER_DENOMINATOR 10    /* qc_shor.c:122 */

unsigned qcsh_gcd(unsigned a, unsigned b);                                    /* qc_shor.c:756-779 */
double qcsh_read_omega(unsigned long long state_num, int L_size, int M_size); /* qc_shor.c:868-883 */
/* returns how many denominators were produced (always n in verbatim mode) */
unsigned qcsh_continued_fraction_denominators(double omega, unsigned n, unsigned *out, int mode); /* qc_shor.c:806-846 */
unsigned long long qcsh_modpow(unsigned base, unsigned long long exponent, unsigned modulus);
/* does a^period == 1 (mod C) hold, the way `mode` evaluates it (qc_shor.c:946) */
int qcsh_period_is_valid(unsigned a, unsigned period, unsigned C, int mode);

typedef struct {
    int mode;            /* QCSH_VERBATIM / QCSH_ROBUST */
    int verbose;         /* -v  */
    int very_verbose;    /* -V  */
    unsigned long long last_measured;
    double last_omega;
} qcsh_options;

/* qc_shor.c:912-964; returns a qcs error code (QCS_PERIOD_NOT_FOUND = 3) */
int qcsh_find_period(unsigned *period, unsigned C, unsigned a, struct qcs_register *reg,
                     qcsh_rng *rng, qcsh_options *opt);
/* qc_shor.c:1003-1134 */
int qcsh_shors_algorithm(unsigned factors[2], unsigned C, unsigned forced_trial_int,
                         struct qcs_register *reg, qcsh_rng *rng, qcsh_options *opt);

#endif
