/*
 * mock_qcs.c -- TEST INFRASTRUCTURE ONLY (tests/test_host_stdout.py, tests/test_dropin.py; CPU).
 *
 * The handful of libqcs.so entry points the C host driver calls (include/qcs.h; qc_shor.c:922-928,
 * 1316-1333), answered by the CPU oracle (oracle/qcs_oracle.c), so that the classical half of the
 * host driver -- argument parsing, RNG, read_omega, continued fractions, candidate search, every
 * line it prints -- can be compared with the unmodified reference's output in a container without
 * a GPU.  It is injected with LD_PRELOAD by that one test and is never built into, linked by or
 * shipped with the product: libqcs.so itself has no CPU path.
 */
#include <math.h>
#include <stdlib.h>

#include "qcs.h"
#include "qcs_oracle.h"

struct qcs_register {
    orc_register *o;
    int L_size, M_size;
};

int qcs_register_create(qcs_register **out, int L_size, int M_size, int device)
{
    (void) device;
    if (!out || L_size <= 0 || M_size < 0 || L_size + M_size > 24) return QCS_BAD_ARGUMENTS;
    qcs_register *reg = (qcs_register *) calloc(1, sizeof *reg);
    if (!reg) return QCS_INSUFFICIENT_MEMORY;
    reg->o = orc_create(L_size, M_size);
    if (!reg->o) { free(reg); return QCS_INSUFFICIENT_MEMORY; }
    reg->L_size = L_size;
    reg->M_size = M_size;
    *out = reg;
    return QCS_NO_ERROR;
}

int qcs_register_create_multi(qcs_register **out, int L_size, int M_size, int n_gpus)
{
    (void) n_gpus;
    return qcs_register_create(out, L_size, M_size, 0);
}

void qcs_register_destroy(qcs_register *reg)
{
    if (!reg) return;
    orc_destroy(reg->o);
    free(reg);
}

int qcs_L_size(const qcs_register *reg) { return reg->L_size; }
int qcs_M_size(const qcs_register *reg) { return reg->M_size; }
int qcs_set_option(qcs_register *reg, int option, long long value) { (void) reg; (void) option; (void) value; return QCS_NO_ERROR; }

int qcs_reset_register(qcs_register *reg)
{
    orc_reset_register(reg->o);
    return QCS_NO_ERROR;
}

int qcs_quantum_computation(qcs_register *reg, unsigned C, unsigned a, int pow_mode)
{
    orc_quantum_computation(reg->o, C, a, pow_mode == QCS_POW_MODULAR ? 1 : 0);
    return QCS_NO_ERROR;
}

int qcs_measure_state(qcs_register *reg, double r, unsigned long long *state_num)
{
    *state_num = (unsigned long long) orc_measure_state(reg->o, r);
    return QCS_NO_ERROR;
}

/* the primitive gates (oracle/make_dropin.py: the reference's own quantum_computation on top of them) */
int qcs_hadamard_gate(qcs_register *reg, unsigned qubit_num)
{
    orc_hadamard_gate(reg->o, qubit_num);
    return QCS_NO_ERROR;
}

int qcs_c_phase_shift_gate(qcs_register *reg, unsigned c_qubit_num, unsigned qubit_num, double theta)
{
    orc_c_phase_shift_gate(reg->o, c_qubit_num, qubit_num, theta);
    return QCS_NO_ERROR;
}

int qcs_c_amodc_gate(qcs_register *reg, unsigned C, unsigned long long atox, unsigned c_qubit_num)
{
    orc_c_amodc_gate(reg->o, C, atox, c_qubit_num);
    return QCS_NO_ERROR;
}

int qcs_inverse_QFT(qcs_register *reg)
{
    orc_inverse_QFT(reg->o);
    return QCS_NO_ERROR;
}

/* read-back used by the host-side debug helpers (quantumcomputer_b200/host/state_debug.c) */
unsigned qcs_num_qubits(const qcs_register *reg) { return (unsigned) (reg->L_size + reg->M_size); }

int qcs_norm2(qcs_register *reg, double *sum_of_sq)
{
    *sum_of_sq = orc_norm2(reg->o);
    return QCS_NO_ERROR;
}

int qcs_nonzero_states(qcs_register *reg, unsigned long long capacity, unsigned long long *indices,
                       double *abs_values, unsigned long long *count)
{
    const unsigned long long N = (unsigned long long) orc_num_states(reg->o);
    double *amp = (double *) malloc(2 * N * sizeof *amp);
    if (!amp) return QCS_INSUFFICIENT_MEMORY;
    orc_get_state(reg->o, amp);
    unsigned long long total = 0;
    for (unsigned long long i = 0; i < N; i++) {
        const double m = hypot(amp[2 * i], amp[2 * i + 1]);      /* gsl_complex_abs */
        if (m != 0.0) {
            if (total < capacity) {
                if (indices) indices[total] = i;
                if (abs_values) abs_values[total] = m;
            }
            total++;
        }
    }
    free(amp);
    *count = total;
    return QCS_NO_ERROR;
}
