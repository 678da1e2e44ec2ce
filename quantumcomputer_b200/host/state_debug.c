#include "state_debug.h"

#include <stdio.h>
#include <stdlib.h>

#include "qcs.h"

int qcsh_display_state(struct qcs_register *reg)
{
    unsigned long long count = 0;
    int rc = qcs_nonzero_states(reg, 0, NULL, NULL, &count);
    if (rc != QCS_NO_ERROR) return rc;
    if (count == 0) return QCS_NO_ERROR;
    unsigned long long *index = (unsigned long long *) malloc(count * sizeof *index);
    double *modulus = (double *) malloc(count * sizeof *modulus);
    if (!index || !modulus) {
        free(index);
        free(modulus);
        return QCS_INSUFFICIENT_MEMORY;
    }
    rc = qcs_nonzero_states(reg, count, index, modulus, &count);
    if (rc == QCS_NO_ERROR) {
        const int n = (int) qcs_num_qubits(reg);
        for (unsigned long long k = 0; k < count; k++) {
            printf("|");
            for (int b = n - 1; b >= 0; b--) printf("%d", (int) ((index[k] >> b) & 1ull));
            printf("> ");
            printf("%.2f\n", modulus[k]);      /* the modulus, as the reference prints it (T:13,22) */
        }
    }
    free(index);
    free(modulus);
    return rc;
}

int qcsh_check_normalisation(struct qcs_register *reg)
{
    double sum_of_sq = 0.0;
    const int rc = qcs_norm2(reg, &sum_of_sq);
    if (rc != QCS_NO_ERROR) return rc;
    printf("Total Probability: %.16f\n", sum_of_sq);
    return QCS_NO_ERROR;
}
