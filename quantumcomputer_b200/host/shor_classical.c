#include "shor_classical.h"

#include <math.h>
#include <stdio.h>
#include <time.h>

#include "qcs.h"

/* double -> unsigned as gcc/x86-64 compiles the reference's casts */
static unsigned to_u32_x86(double d)
{
    if (!(d < 9223372036854775808.0) || !(d > -9223372036854775808.0)) return 0u;
    return (unsigned) (uint64_t) (int64_t) d;
}

unsigned qcsh_gcd(unsigned a, unsigned b)
{
    if (a == 0 || b == 0) return a + b;
    while (b != 0) {
        unsigned r = a % b;
        a = b;
        b = r;
    }
    return a;
}

double qcsh_read_omega(unsigned long long state_num, int L_size, int M_size)
{
    /* the x register is read most-significant-qubit first into the LEAST
     * significant result bit: the inverse QFT of the reference leaves its
     * output bit-reversed and applies no swaps (qc_shor.c:876-880) */
    unsigned x_tilde = 0;
    for (int k = 0; k < L_size; k++) {
        const int qubit = M_size + L_size - 1 - k;
        x_tilde |= (unsigned) ((state_num >> qubit) & 1ull) << k;
    }
    return (double) x_tilde / (double) qcs_int_pow(2, (unsigned) L_size);
}

unsigned qcsh_continued_fraction_denominators(double omega, unsigned n, unsigned *out, int mode)
{
    unsigned coeff[64];
    if (n > 64) n = 64;
    unsigned made = 0;
    for (unsigned i = 0; i < n; i++) {
        /* robust mode: once the fraction has terminated (omega reached 0, or a
         * rounding residue whose reciprocal no longer fits the coefficient
         * type) emit the convergent of the coefficients found so far and stop */
        const int last = mode == QCSH_ROBUST && (!(omega > 0.0) || !(1.0 / omega < 4294967296.0));
        if (!last) {
            const double inverse = 1.0 / omega;
            omega = inverse - (double) to_u32_x86(inverse);
            coeff[i] = to_u32_x86(inverse - omega);
        }
        /* convergent built from coefficients i-1 .. 0 (qc_shor.c:834-842) */
        unsigned den = 1, num = 0;
        for (unsigned k = i; k-- > 0;) {
            const unsigned keep = den;
            den = num + den * coeff[k];
            num = keep;
        }
        out[made++] = den;
        if (last) break;
    }
    return made;
}

unsigned long long qcsh_modpow(unsigned base, unsigned long long exponent, unsigned modulus)
{
    if (modulus == 0) return 0;
    unsigned long long result = 1 % modulus, b = base % modulus;
    while (exponent) {
        if (exponent & 1ull) result = (result * b) % modulus;
        b = (b * b) % modulus;
        exponent >>= 1;
    }
    return result;
}

int qcsh_period_is_valid(unsigned a, unsigned period, unsigned C, int mode)
{
    if (mode == QCSH_ROBUST) return period != 0 && qcsh_modpow(a, period, C) == 1 % C;
    return qcs_int_pow(a, period) % C == 1;
}

static unsigned half_power(unsigned a, unsigned period, unsigned C, int mode)
{
    /* a^(p/2), reduced in robust mode, wrapped like INT_POW in verbatim mode */
    if (mode == QCSH_ROBUST) return (unsigned) qcsh_modpow(a, period / 2, C);
    return qcs_int_pow(a, period / 2);
}

int qcsh_find_period(unsigned *period, unsigned C, unsigned a, struct qcs_register *reg,
                     qcsh_rng *rng, qcsh_options *opt)
{
    unsigned den[QCSH_NUM_CONTINUED_FRACTIONS];
    unsigned long long measured = 0;
    int rc;

    if (opt->very_verbose) {
        /* the three stages are one library call here; the reference names them as it goes (qc_shor.c:717-735) */
        printf("      - Performing quantum computation...\n");
        printf("         - Applying Hadamard matrices.\n");
        printf("         - Applying a^x mod (C) gates.\n");
        printf("         - Performing inverse quantum Fourier transform.\n");
    }
    if ((rc = qcs_reset_register(reg)) != QCS_NO_ERROR) return rc;
    rc = qcs_quantum_computation(reg, C, a, opt->mode == QCSH_ROBUST ? QCS_POW_MODULAR : QCS_POW_VERBATIM);
    if (rc != QCS_NO_ERROR) return rc;

    if (opt->very_verbose) printf("      - Measuring state...\n");
    const double r = qcsh_rng_uniform(rng);            /* qc_shor.c:281 */
    if ((rc = qcs_measure_state(reg, r, &measured)) != QCS_NO_ERROR) return rc;
    const double omega = qcsh_read_omega(measured, qcs_L_size(reg), qcs_M_size(reg));
    opt->last_measured = measured;
    opt->last_omega = omega;

    if (opt->very_verbose) printf("      - Using continued fractions to guess period...\n");
    const unsigned n_den = qcsh_continued_fraction_denominators(omega, QCSH_NUM_CONTINUED_FRACTIONS, den, opt->mode);
    for (unsigned d = 0; d < n_den; d++) {
        for (unsigned m = 1; m <= QCSH_TRIALS_PER_DENOMINATOR; m++) {
            *period = m * den[d];
            if (qcsh_period_is_valid(a, *period, C, opt->mode)) return QCS_NO_ERROR;
        }
    }
    return QCS_PERIOD_NOT_FOUND;
}

static double seconds_since(const struct timespec *t0)
{
    struct timespec t1;
    clock_gettime(CLOCK_REALTIME, &t1);
    return (double) (t1.tv_sec - t0->tv_sec) + (double) (t1.tv_nsec - t0->tv_nsec) / 1e9;
}

int qcsh_shors_algorithm(unsigned factors[2], unsigned C, unsigned forced_trial_int,
                         struct qcs_register *reg, qcsh_rng *rng, qcsh_options *opt)
{
    struct timespec t0;
    unsigned period = 0;
    printf("\n --- Finding factors...\n\n");
    clock_gettime(CLOCK_REALTIME, &t0);

    const unsigned first = forced_trial_int ? forced_trial_int : 2;
    const unsigned last = forced_trial_int ? forced_trial_int : (C >= 2 ? C - 2 : 0);
    for (unsigned a = first; a <= last; a++) {
        if (opt->verbose)
            printf(forced_trial_int ? " --- Forced trial integer a = %d, finding period ...\n"
                                    : " --- Trial integer a = %d, finding period ...\n", a);
        const int rc = qcsh_find_period(&period, C, a, reg, rng, opt);
        if (rc != QCS_NO_ERROR && rc != QCS_PERIOD_NOT_FOUND) return rc;
        int rejected = 0;
        if (rc == QCS_PERIOD_NOT_FOUND) {
            if (opt->verbose && !forced_trial_int) printf(" --- A valid period could not be found for a = %d.\n\n", a);
            rejected = 1;
        } else {
            /* the reference tests forced_trial_int in both branches (qc_shor.c:1037,
             * 1091); in the loop branch that is 0, kept in verbatim mode */
            const unsigned tested = (opt->mode == QCSH_ROBUST) ? a : forced_trial_int;
            const unsigned hp = (opt->mode == QCSH_ROBUST) ? half_power(tested, period, C, opt->mode)
                                                           : qcs_int_pow(tested, period / 2) % C;
            if (period % 2 != 0 || hp == C - 1) {
                if (opt->verbose)
                    printf(" --- Period was found to be %d, but it did not pass the validity requirements.\n%s",
                           period, forced_trial_int ? "" : "\n");
                rejected = 1;
            }
        }
        if (rejected) {
            if (forced_trial_int) break;
            continue;
        }
        if (opt->verbose)
            printf(" --- A valid period = %d has been found so the factors of C = %d have been found quantum mechanically.\n\n",
                   period, C);
        const unsigned hp = half_power(a, period, C, opt->mode);
        factors[0] = qcsh_gcd(hp + 1, C);
        factors[1] = qcsh_gcd(hp - 1, C);
        if (factors[0] == 1 || factors[1] == 1) {
            if (forced_trial_int) {
                printf(" --- The factors found are trivial, consider trying a different trial integer.\n");
            } else {
                printf(" --- Factors found are trivial. Continuing to find non-trivial factors.\n");
                continue;
            }
        }
        if (opt->verbose) printf(" --- Time to run Shor's Algorithm: %.6fs.\n", seconds_since(&t0));
        return QCS_NO_ERROR;
    }
    printf(" --- A valid period was not found and hence C = %d could not be factorised.\n", C);
    if (opt->verbose && !forced_trial_int) printf(" --- Time to run Shor's Algorithm: %.6fs.\n", seconds_since(&t0));
    return QCS_PERIOD_NOT_FOUND;
}
