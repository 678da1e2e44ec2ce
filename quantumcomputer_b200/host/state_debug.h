/* state_debug.h -- the reference's two development helpers (testing_and_debug.c), which a developer
 * inserts by hand around the gate calls, on top of the C ABI: same names' worth of behaviour, same
 * text on stdout.  Host C; the device work is qcs_nonzero_states / qcs_norm2. */
#ifndef QCS_HOST_STATE_DEBUG_H
#define QCS_HOST_STATE_DEBUG_H

struct qcs_register;

/* display_state, testing_and_debug.c:7-26: one line "|<bits, most significant first>> <|amp|, %.2f>" per
 * basis state whose amplitude has non-zero modulus, in index order.  Returns a qcs error code. */
int qcsh_display_state(struct qcs_register *reg);

/* check_normalisation, testing_and_debug.c:28-37: "Total Probability: %.16f".  The sum is the device's
 * deterministic parallel reduction (qcs_norm2); the reference adds in index order, so the sixteenth
 * decimal may differ. */
int qcsh_check_normalisation(struct qcs_register *reg);

#endif
