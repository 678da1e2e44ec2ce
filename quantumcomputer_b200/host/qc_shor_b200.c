/* qc_shor_b200.c -- command-line Shor driver on top of libqcs.so.
 *
 * Same interface as the reference program (qc_shor.c:1173-1348):
 *     qc_shor_b200 -C num -L L_reg_size -M M_reg_size [-a trial_int] [-v] [-V]
 * plus  -f trial_int  (the spelling the reference documents but does not
 * parse, qc_shor.c:26,1177,1185), -s seed (the reference seeds with time(NULL),
 * qc_shor.c:1299), -r (robust classical post-processing), -x (gate-by-gate
 * reference-order kernels instead of fused sweeps), -d device, -g n_gpus (the
 * register sharded over the first n_gpus devices of the box: qcs_register_create_multi).
 * All state lives on the GPU; this file only makes the three calls the
 * reference's find_period makes into the gate path (qc_shor.c:922-928).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>

#include "qcs.h"
#include "mt19937.h"
#include "shor_classical.h"

static const char *USAGE =
    "Usage: ./qc_shor_b200 -C num -L L_reg_size -M M_reg_size [-a|-f trial_int] [-v] [-V] [-s seed] [-r] [-x] [-d device] [-g n_gpus]\n";

static void issue_warnings(unsigned C, int L, int M)
{
    /* qc_shor.c:340-351, text and arithmetic as there: INT_POW(2, M) < C, and INT_POW(2, L) < C*C with
     * the product taken in unsigned int */
    if (qcs_int_pow(2, (unsigned) M) < C)
        printf(" --- *WARNING* The M register is not large enough for reliable results. Ensure 2^M >= C. Minimum: M = %d.\n",
               ((int) (log2(C) + 0.5)) + 1);
    if (qcs_int_pow(2, (unsigned) L) < C * C)
        printf(" --- *WARNING* The L register is not large enough for full confidence in finding the period. Ensure 2^L >= C^2 for confidence. Suggested: L = %d.\n",
               (int) (log2(C * C) + 0.5));
}

int main(int argc, char *argv[])
{
    unsigned C = 0, trial = 0;
    int L = 0, M = 0, have_C = 0, have_L = 0, have_M = 0, device = -1, exact = 0, n_gpus = 1;
    unsigned long seed = (unsigned long) time(NULL);
    qcsh_options opt = {QCSH_VERBATIM, 0, 0, 0, 0.0};
    int arg;

    while ((arg = getopt(argc, argv, "C:L:M:a:f:vVs:rxd:g:")) != -1) {
        switch (arg) {
            case 'C': C = (unsigned) atoi(optarg); have_C = 1; break;
            case 'L': L = atoi(optarg); have_L = 1; break;
            case 'M': M = atoi(optarg); have_M = 1; break;
            case 'a':
            case 'f': trial = (unsigned) atoi(optarg); break;
            case 'v': opt.verbose = 1; break;
            case 'V': opt.verbose = 1; opt.very_verbose = 1; break;
            case 's': seed = strtoul(optarg, NULL, 10); break;
            case 'r': opt.mode = QCSH_ROBUST; break;
            case 'x': exact = 1; break;
            case 'd': device = atoi(optarg); break;
            case 'g': n_gpus = atoi(optarg); break;
            default: fputs(USAGE, stdout); return QCS_BAD_ARGUMENTS;
        }
    }
    if (!have_C || !have_L || !have_M) {
        fprintf(stderr, "Error: %s not given.\n", !have_C ? "Number to be factorised 'C'"
                                                 : !have_L ? "Size of L register" : "Size of M register");
        fputs(USAGE, stdout);
        return QCS_BAD_ARGUMENTS;
    }
    if (C < 3 || L <= 0 || M <= 0) {
        fprintf(stderr, "Error: C, L and M must be positive (C >= 3).\n");
        fputs(USAGE, stdout);
        return QCS_BAD_ARGUMENTS;
    }
    issue_warnings(C, L, M);

    qcsh_rng rng;
    qcsh_rng_seed(&rng, seed);

    qcs_register *reg = NULL;
    int rc = n_gpus > 1 ? qcs_register_create_multi(&reg, L, M, n_gpus) : qcs_register_create(&reg, L, M, device);
    if (rc != QCS_NO_ERROR) {
        fprintf(stderr, "Error: %s.\n", rc == QCS_INSUFFICIENT_MEMORY ? "Insufficient memory" : qcs_error_string(rc));
        return rc;
    }
    if (exact) qcs_set_option(reg, QCS_OPT_FUSION, 0);

    unsigned factors[2] = {0, 0};
    rc = qcsh_shors_algorithm(factors, C, trial, reg, &rng, &opt);
    qcs_register_destroy(reg);

    if (rc == QCS_NO_ERROR) {
        printf(" --- Factors of %d found: (%d, %d).\n", C, factors[0], factors[1]);
        if (factors[0] == 0 || C / factors[0] != factors[1])
            printf(" --- These factors are incorrect. Consider increasing register sizes as per the warnings.\n");
        return QCS_NO_ERROR;
    }
    return rc == QCS_PERIOD_NOT_FOUND ? QCS_PERIOD_NOT_FOUND : QCS_UNKNOWN_ERROR;
}
