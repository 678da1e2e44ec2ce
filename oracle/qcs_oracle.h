/*
 * qcs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's gate-application path (qc_shor.c:242-324,
 * 370-737) and of the classical driver around it (qc_shor.c:756-1134), written
 * matrix-free (O(2^n) per gate instead of the reference's O(4^n) matrix build)
 * but with the SAME floating-point operations in the SAME order as
 * operate_matrix (qc_shor.c:396-413) performs them, so results are identical to
 * the reference bit for bit (sign of zero included).
 *
 * Parity status: PINNED -- tests/test_oracle.py checks this file against
 *   (a) oracle/_ref/libqcref.so, the unmodified reference compiled in place, and
 *   (b) the JSON files under tests/golden, vectors generated from (a) by oracle/make_golden.py,
 * and against KAT-1..4 of SURVEY.md section 4.  The reference itself ships no tests.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libqcs.so) never
 * does: it has no CPU fallback.
 */
#ifndef QCS_ORACLE_H
#define QCS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_register orc_register;

/* register life-cycle: qc_shor.c:1255-1261, 1316-1324, 1330-1333 */
orc_register *orc_create(int L_size, int M_size);
void orc_destroy(orc_register *r);
uint64_t orc_num_states(const orc_register *r);
void orc_get_state(const orc_register *r, double *out_interleaved);
void orc_set_state(orc_register *r, const double *in_interleaved);

/* gate path */
void orc_reset_register(orc_register *r);                                   /* qc_shor.c:318-324 */
void orc_hadamard_gate(orc_register *r, unsigned qubit_num);                /* qc_shor.c:442-484 */
void orc_c_phase_shift_gate(orc_register *r, unsigned c_qubit_num,
                            unsigned qubit_num, double theta);              /* qc_shor.c:513-565 */
void orc_c_amodc_gate(orc_register *r, unsigned C, unsigned long long atox,
                      unsigned c_qubit_num);                                /* qc_shor.c:595-660 */
void orc_inverse_QFT(orc_register *r);                                      /* qc_shor.c:678-690 */
/* pow_mode 0: atox = INT_POW(a, x) exactly as qc_shor.c:158-159,729 behaves on
 * x86-64; pow_mode 1: atox = a^(2^k) mod C by modular squaring (intended). */
void orc_quantum_computation(orc_register *r, unsigned C, unsigned a, int pow_mode); /* qc_shor.c:712-737 */
uint64_t orc_measure_state(orc_register *r, double rnd);                    /* qc_shor.c:272-306 */
double orc_norm2(const orc_register *r);                                    /* testing_and_debug.c:28-37 */

/* scalar helpers of the classical driver */
unsigned orc_int_pow(unsigned base, unsigned power);                        /* qc_shor.c:158-159 */
unsigned orc_gcd(unsigned a, unsigned b);                                   /* qc_shor.c:756-779 */
double orc_read_omega(uint64_t state_num, int L_size, int M_size);          /* qc_shor.c:868-883 */
void orc_continued_fraction_denominators(double omega, unsigned n, unsigned *out); /* qc_shor.c:806-846 */

/* MT19937 with gsl_rng_mt19937 seeding and gsl_rng_uniform scaling (KAT-3) */
typedef struct { uint32_t mt[624]; int idx; } orc_mt19937;
void orc_mt_seed(orc_mt19937 *g, unsigned long seed);
uint32_t orc_mt_next(orc_mt19937 *g);
double orc_mt_uniform(orc_mt19937 *g);

/* find_period / shors_algorithm with an explicit RNG; returns the reference's
 * ErrorCode (0 ok, 3 PERIOD_NOT_FOUND).  period_found starts false (the
 * reference leaves it uninitialised). */
int orc_find_period(orc_register *r, unsigned C, unsigned a, int pow_mode,
                    orc_mt19937 *g, unsigned *period, uint64_t *measured);
int orc_shors_algorithm(orc_register *r, unsigned C, unsigned forced_a, int pow_mode,
                        orc_mt19937 *g, unsigned factors[2]);

/* synthetic benchmark state of SURVEY section 8(d): counter-based splitmix64 */
double orc_synthetic_u(uint64_t seed, uint64_t k);
void orc_fill_synthetic(orc_register *r, uint64_t seed);   /* un-normalised */
void orc_scale(orc_register *r, double s);

#ifdef __cplusplus
}
#endif
#endif
