/*
 * ref_bridge.c -- TEST INFRASTRUCTURE ONLY (oracle side, never shipped, never on
 * the product path).
 *
 * Compiles the UNMODIFIED reference program in place -- the translation unit
 * below textually includes the file named by QC_REF_SOURCE (set by
 * oracle/Makefile to /root/reference/qc_shor.c; nothing is copied into this
 * repository) against the GSL stand-in in oracle/gsl_shim -- and exports a few
 * plain-C entry points that call the reference's own `static` functions so that
 * tests and bench.py's cpu_baseline / --impl reference legs can run them through
 * ctypes.  Output: oracle/_ref/libqcref.so (git-ignored).
 *
 * Every amplitude operation executed through this file is the reference's code:
 * operate_matrix (qc_shor.c:370-420) and the three matrix builders
 * (qc_shor.c:442-484, 513-565, 595-660).
 */
#ifndef QC_REF_SOURCE
#error "QC_REF_SOURCE must name the reference qc_shor.c (see oracle/Makefile)"
#endif

#define main qc_ref_main
#include QC_REF_SOURCE
#undef main
/* the two development helpers the reference keeps in a file of their own, "not linked to by qc_shor.c so
 * should be inserted ... by hand" (testing_and_debug.c:1-5): inserted here, unmodified */
#ifdef QC_REF_DEBUG_SOURCE
#include QC_REF_DEBUG_SOURCE
#endif

#include <string.h>

typedef struct {
    Register reg;
    gsl_spmatrix_complex *matrix;
    gsl_rng *rng;
} qcref_handle;

/* mirrors the allocation block of main(), qc_shor.c:1316-1324 and the size
 * computation of parse_command_line_args(), qc_shor.c:1255-1261 */
void *qcref_create(int L_size, int M_size)
{
    qcref_handle *h = (qcref_handle *) calloc(1, sizeof *h);
    if (!h) return NULL;
    h->reg.L_size = L_size;
    h->reg.M_size = M_size;
    h->reg.num_qubits = (unsigned int) (L_size + M_size);
    h->reg.num_states = 1;
    for (unsigned int k = 0; k < h->reg.num_qubits; k++) h->reg.num_states *= 2;
    h->reg.state_a = gsl_vector_complex_alloc(h->reg.num_states);
    h->reg.state_b = gsl_vector_complex_alloc(h->reg.num_states);
    h->matrix = gsl_spmatrix_complex_alloc_nzmax(h->reg.num_states, h->reg.num_states,
                                                 2 * h->reg.num_states, GSL_SPMATRIX_COO);
    h->rng = gsl_rng_alloc(gsl_rng_mt19937);
    if (!h->reg.state_a || !h->reg.state_b || !h->matrix || !h->rng) return NULL;
    h->reg.current_state = &h->reg.state_a;
    h->reg.new_state = &h->reg.state_b;
    return h;
}

void qcref_destroy(void *hv)
{
    qcref_handle *h = (qcref_handle *) hv;
    if (!h) return;
    gsl_vector_complex_free(h->reg.state_a);
    gsl_vector_complex_free(h->reg.state_b);
    gsl_spmatrix_complex_free(h->matrix);
    gsl_rng_free(h->rng);
    free(h);
}

unsigned long long qcref_num_states(void *hv)
{
    return (unsigned long long) ((qcref_handle *) hv)->reg.num_states;
}

void qcref_set_verbosity(int v, int vv) { verbose = v != 0; very_verbose = vv != 0; }

void qcref_seed(void *hv, unsigned long seed) { gsl_rng_set(((qcref_handle *) hv)->rng, seed); }

double qcref_rng_uniform(void *hv) { return gsl_rng_uniform(((qcref_handle *) hv)->rng); }

void qcref_get_state(void *hv, double *out_interleaved)
{
    qcref_handle *h = (qcref_handle *) hv;
    memcpy(out_interleaved, (*h->reg.current_state)->data, 2 * h->reg.num_states * sizeof(double));
}

void qcref_set_state(void *hv, const double *in_interleaved)
{
    qcref_handle *h = (qcref_handle *) hv;
    memcpy((*h->reg.current_state)->data, in_interleaved, 2 * h->reg.num_states * sizeof(double));
}

void qcref_reset_register(void *hv) { reset_register(((qcref_handle *) hv)->reg); }

void qcref_hadamard_gate(void *hv, unsigned int qubit_num)
{
    qcref_handle *h = (qcref_handle *) hv;
    hadamard_gate(qubit_num, &h->reg, h->matrix);
}

void qcref_c_phase_shift_gate(void *hv, unsigned int c_qubit_num, unsigned int qubit_num, double theta)
{
    qcref_handle *h = (qcref_handle *) hv;
    c_phase_shift_gate(c_qubit_num, qubit_num, theta, &h->reg, h->matrix);
}

void qcref_c_amodc_gate(void *hv, unsigned int C, unsigned long long atox, unsigned int c_qubit_num)
{
    qcref_handle *h = (qcref_handle *) hv;
    c_amodc_gate(C, atox, c_qubit_num, &h->reg, h->matrix);
}

void qcref_inverse_QFT(void *hv)
{
    qcref_handle *h = (qcref_handle *) hv;
    inverse_QFT(&h->reg, h->matrix);
}

void qcref_quantum_computation(void *hv, unsigned int C, unsigned int a)
{
    qcref_handle *h = (qcref_handle *) hv;
    quantum_computation(C, a, &h->reg, h->matrix);
}

/* the value INT_POW (qc_shor.c:158-159) yields on this platform, for the host
 * "reference-verbatim" power mode tests */
unsigned int qcref_int_pow(unsigned int base, unsigned int power) { return INT_POW(base, power); }

/* measure with the handle's MT19937 stream (qc_shor.c:272-306) */
unsigned long long qcref_measure_state(void *hv)
{
    qcref_handle *h = (qcref_handle *) hv;
    return (unsigned long long) measure_state(h->reg, h->rng);
}

/* measure with an explicit r (stand-in extension: forced uniform) */
unsigned long long qcref_measure_state_r(void *hv, double r)
{
    qcref_handle *h = (qcref_handle *) hv;
    h->rng->forced[0] = r;
    h->rng->n_forced = 1;
    return (unsigned long long) measure_state(h->reg, h->rng);
}

double qcref_read_omega(void *hv, unsigned long long state_num)
{
    return read_omega((unsigned long int) state_num, ((qcref_handle *) hv)->reg);
}

void qcref_continued_fraction_denominators(double omega, unsigned int n, unsigned int *out)
{
    get_continued_fractions_denominators(omega, n, out);
}

/* issue_warnings, qc_shor.c:340-351: prints to stdout */
void qcref_issue_warnings(unsigned int C, int L_size, int M_size)
{
    Register reg;
    memset(&reg, 0, sizeof reg);
    reg.L_size = L_size;
    reg.M_size = M_size;
    issue_warnings(C, reg);
    fflush(stdout);
}

unsigned int qcref_gcd(unsigned int a, unsigned int b) { return greatest_common_divisor(a, b); }

/* find_period (qc_shor.c:912-964): returns the ErrorCode, period in *period */
int qcref_find_period(void *hv, unsigned int C, unsigned int a, unsigned int *period)
{
    qcref_handle *h = (qcref_handle *) hv;
    return (int) find_period(period, C, a, &h->reg, h->matrix, h->rng);
}

/* shors_algorithm (qc_shor.c:1003-1134) */
int qcref_shors_algorithm(void *hv, unsigned int C, unsigned int forced_a, unsigned int *factors2)
{
    qcref_handle *h = (qcref_handle *) hv;
    return (int) shors_algorithm(factors2, C, forced_a, &h->reg, h->matrix, h->rng);
}

/* sum of |amp|^2 in index order, the loop of testing_and_debug.c:28-37 */
double qcref_norm2(void *hv)
{
    qcref_handle *h = (qcref_handle *) hv;
    double s = 0.0;
    for (unsigned long int i = 0; i < h->reg.num_states; i++)
        s += gsl_complex_abs2(gsl_vector_complex_get(*h->reg.current_state, i));
    return s;
}

#ifdef QC_REF_DEBUG_SOURCE
/* display_state / check_normalisation, testing_and_debug.c:7-37: print to stdout */
void qcref_display_state(void *hv)
{
    display_state(((qcref_handle *) hv)->reg);
    fflush(stdout);
}

void qcref_check_normalisation(void *hv)
{
    check_normalisation(((qcref_handle *) hv)->reg);
    fflush(stdout);
}
#endif
