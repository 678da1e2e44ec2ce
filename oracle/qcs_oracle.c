/*
 * qcs_oracle.c -- TEST INFRASTRUCTURE ONLY.  See qcs_oracle.h for scope and
 * parity status.  Compile with -ffp-contract=off: the reference arithmetic has
 * no fused multiply-adds (qc_shor.c:409,412 are written as separate products
 * and sums and the documented build line has no -march flag).
 */
#include "qcs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
#ifndef M_SQRT1_2
#define M_SQRT1_2 0.70710678118654752440084436210
#endif

struct orc_register {
    int L_size, M_size;
    unsigned num_qubits;
    uint64_t num_states;
    double *cur;   /* interleaved (re, im): gsl_vector_complex.data layout */
    double *nxt;
};

/* One COO entry applied the way operate_matrix does it, qc_shor.c:399-412:
 * new[row] += m * cur[col] as two products and one sum/difference per part. */
static inline void accumulate(double *dst, double m_re, double m_im, const double *src)
{
    double c_re = src[0], c_im = src[1];
    dst[0] += (m_re * c_re) - (m_im * c_im);
    dst[1] += (m_re * c_im) + (m_im * c_re);
}

/* The all-cores CPU baseline of bench.py (the "fair" matrix-free port) is this same
 * file compiled with -fopenmp: the rows of a Hadamard / controlled-phase gate are independent, so
 * the row loop is split over the threads and each thread zeroes its own rows (operate_matrix's
 * zeroing, qc_shor.c:393) -- the arithmetic per row, and therefore every result bit, is unchanged. */
#ifdef _OPENMP
#define ORC_ZERO(nxt, N) ((void) 0)
#define ORC_ROW_ZERO(p) do { (p)[0] = 0.0; (p)[1] = 0.0; } while (0)
#define ORC_PARALLEL_FOR _Pragma("omp parallel for schedule(static)")
#else
#define ORC_ZERO(nxt, N) memset(nxt, 0, 2 * (N) * sizeof(double))
#define ORC_ROW_ZERO(p) ((void) 0)
#define ORC_PARALLEL_FOR
#endif

static void finish_gate(orc_register *r)
{
    /* swap_states, qc_shor.c:242-249 */
    double *t = r->cur; r->cur = r->nxt; r->nxt = t;
}

orc_register *orc_create(int L_size, int M_size)
{
    orc_register *r = (orc_register *) calloc(1, sizeof *r);
    if (!r) return NULL;
    r->L_size = L_size;
    r->M_size = M_size;
    r->num_qubits = (unsigned) (L_size + M_size);
    r->num_states = (uint64_t) 1 << r->num_qubits;
    r->cur = (double *) calloc(2 * r->num_states, sizeof(double));
    r->nxt = (double *) calloc(2 * r->num_states, sizeof(double));
    if (!r->cur || !r->nxt) { orc_destroy(r); return NULL; }
    return r;
}

void orc_destroy(orc_register *r)
{
    if (!r) return;
    free(r->cur); free(r->nxt); free(r);
}

uint64_t orc_num_states(const orc_register *r) { return r->num_states; }

void orc_get_state(const orc_register *r, double *out)
{
    memcpy(out, r->cur, 2 * r->num_states * sizeof(double));
}

void orc_set_state(orc_register *r, const double *in)
{
    memcpy(r->cur, in, 2 * r->num_states * sizeof(double));
}

/* qc_shor.c:318-324: zero, then amp[1] = polar(1, 0) = (1*cos 0, 1*sin 0) */
void orc_reset_register(orc_register *r)
{
    memset(r->cur, 0, 2 * r->num_states * sizeof(double));
    r->cur[2] = 1.0 * cos(0.0);
    r->cur[3] = 1.0 * sin(0.0);
}

/* qc_shor.c:442-484.  The builder emits, for every row i, the two columns that
 * differ from i at most in bit q, in ascending column order, with value
 * HADAMARD_BASE_MATRIX[bit_q(i)][bit_q(j)] + 0.0i (qc_shor.c:453,476). */
void orc_hadamard_gate(orc_register *r, unsigned q)
{
    const uint64_t N = r->num_states, bit = (uint64_t) 1 << q;
    const double h = M_SQRT1_2;
    const double *cur = r->cur;
    double *nxt = r->nxt;
    ORC_ZERO(nxt, N);                                      /* qc_shor.c:393 */
    ORC_PARALLEL_FOR
    for (uint64_t i = 0; i < N; i++) {
        uint64_t j0 = i & ~bit, j1 = i | bit;
        ORC_ROW_ZERO(nxt + 2 * i);
        double m0 = h;                                      /* H[b][0] */
        double m1 = (i & bit) ? -h : h;                     /* H[b][1] */
        accumulate(nxt + 2 * i, m0, 0.0, cur + 2 * j0);
        accumulate(nxt + 2 * i, m1, 0.0, cur + 2 * j1);
    }
    finish_gate(r);
}

/* qc_shor.c:513-565.  For every row i the builder emits all four columns that
 * agree with i outside bits {c, q}, ascending, with C_PHASE_SHIFT_BASE_MATRIX
 * [2 c_i + q_i][2 c_j + q_j]; the three off-diagonal zeros are stored
 * explicitly (qc_shor.c:549-559) and therefore take part in the accumulation. */
void orc_c_phase_shift_gate(orc_register *r, unsigned c, unsigned q, double theta)
{
    const uint64_t N = r->num_states;
    const uint64_t cb = (uint64_t) 1 << c, qb = (uint64_t) 1 << q;
    const double e_re = 1.0 * cos(theta), e_im = 1.0 * sin(theta);   /* gsl_complex_polar, qc_shor.c:526 */
    const double *cur = r->cur;
    double *nxt = r->nxt;
    if (c == q) {
        memset(nxt, 0, 2 * N * sizeof(double));
        /* degenerate call (never made by the reference): the delta test leaves
         * rows/cols free in one bit only, base index is 3*bit */
        for (uint64_t i = 0; i < N; i++) {
            for (int s = 0; s < 2; s++) {
                uint64_t j = s ? (i | cb) : (i & ~cb);
                int bi = (i & cb) ? 3 : 0, bj = (j & cb) ? 3 : 0;
                double m_re = 0.0, m_im = 0.0;
                if (bi == bj) { if (bi == 3) { m_re = e_re; m_im = e_im; } else m_re = 1.0; }
                accumulate(nxt + 2 * i, m_re, m_im, cur + 2 * j);
            }
        }
        finish_gate(r);
        return;
    }
    const uint64_t lo = cb < qb ? cb : qb, hi = cb < qb ? qb : cb;
    ORC_ZERO(nxt, N);
    ORC_PARALLEL_FOR
    for (uint64_t i = 0; i < N; i++) {
        uint64_t base = i & ~(cb | qb);
        ORC_ROW_ZERO(nxt + 2 * i);
        int bi = ((i & cb) ? 2 : 0) + ((i & qb) ? 1 : 0);
        uint64_t cols[4] = { base, base | lo, base | hi, base | lo | hi };
        for (int s = 0; s < 4; s++) {
            uint64_t j = cols[s];
            int bj = ((j & cb) ? 2 : 0) + ((j & qb) ? 1 : 0);
            double m_re = 0.0, m_im = 0.0;
            if (bi == bj) {
                if (bi == 3) { m_re = e_re; m_im = e_im; }  /* COMPLEX_ELEMENT -> e^{i theta} */
                else m_re = 1.0;
            }
            accumulate(nxt + 2 * i, m_re, m_im, cur + 2 * j);
        }
    }
    finish_gate(r);
}

/* qc_shor.c:595-660.  One entry (row j(k), column k) = 1 + 0i per column k,
 * emitted for k ascending, so colliding rows sum their sources in ascending k. */
void orc_c_amodc_gate(orc_register *r, unsigned C, unsigned long long atox, unsigned c)
{
    const uint64_t N = r->num_states;
    const unsigned M = (unsigned) r->M_size;
    const uint64_t low_mask = M ? (((uint64_t) 1 << M) - 1) : 0;
    const unsigned A = (unsigned) (atox % C);               /* qc_shor.c:605 */
    const double *cur = r->cur;
    double *nxt = r->nxt;
    memset(nxt, 0, 2 * N * sizeof(double));
    for (uint64_t k = 0; k < N; k++) {
        uint64_t row = k;
        if ((k >> c) & 1) {
            unsigned f = (unsigned) (k & low_mask);         /* qc_shor.c:619-622 */
            if (f < C) {
                f = (A * f) % C;                            /* 32-bit unsigned, qc_shor.c:639 */
                row = (k & ~low_mask) | (uint64_t) (f & (unsigned) low_mask);   /* qc_shor.c:642-652 */
            }
        }
        accumulate(nxt + 2 * row, 1.0, 0.0, cur + 2 * k);
    }
    finish_gate(r);
}

/* INT_POW, qc_shor.c:158-159: (unsigned int)(pow(base, power) + 0.5).  The cast
 * is undefined when out of range; gcc on x86-64 converts through a 64-bit
 * truncation (cvttsd2si) and keeps the low 32 bits, which is what the
 * reference binary does on this platform (INT_POW row). */
unsigned orc_int_pow(unsigned base, unsigned power)
{
    double d = pow((double) base, (double) power) + 0.5;
    if (!(d < 9223372036854775808.0) || !(d > -9223372036854775808.0)) return 0u;
    return (unsigned) (uint64_t) (int64_t) d;
}

static unsigned long long modpow2k(unsigned a, unsigned k, unsigned C)
{
    unsigned long long v = a % C;
    for (unsigned s = 0; s < k; s++) v = (v * v) % C;
    return v;
}

/* qc_shor.c:678-690 */
void orc_inverse_QFT(orc_register *r)
{
    const int M = r->M_size, top = r->L_size + r->M_size - 1;
    for (int l = top; l >= M; l--) {
        orc_hadamard_gate(r, (unsigned) l);
        for (int k = l - 1; k >= M; k--) {
            double theta = M_PI / (double) orc_int_pow(2, (unsigned) (l - k));
            orc_c_phase_shift_gate(r, (unsigned) l, (unsigned) k, theta);
        }
    }
}

/* qc_shor.c:712-737 */
void orc_quantum_computation(orc_register *r, unsigned C, unsigned a, int pow_mode)
{
    const unsigned first = r->num_qubits - (unsigned) r->L_size;
    unsigned x = 1;
    for (unsigned l = first; l < r->num_qubits; l++) orc_hadamard_gate(r, l);
    for (unsigned l = first; l < r->num_qubits; l++) {
        unsigned long long atox = pow_mode ? modpow2k(a, l - first, C)
                                           : (unsigned long long) orc_int_pow(a, x);
        orc_c_amodc_gate(r, C, atox, l);
        x *= 2;
    }
    orc_inverse_QFT(r);
}

/* qc_shor.c:272-306 with the random draw made by the caller */
uint64_t orc_measure_state(orc_register *r, double rnd)
{
    const uint64_t N = r->num_states;
    double cum = 0.0;
    uint64_t s;
    for (s = 0; s < N - 1; s++) {
        double x = r->cur[2 * s], y = r->cur[2 * s + 1];
        cum += x * x + y * y;                               /* gsl_complex_abs2 */
        if (cum >= rnd) break;
    }
    memset(r->cur, 0, 2 * N * sizeof(double));
    r->cur[2 * s] = 1.0;
    r->cur[2 * s + 1] = 0.0;
    return s;
}

double orc_norm2(const orc_register *r)
{
    double sum = 0.0;
    for (uint64_t i = 0; i < r->num_states; i++) {
        double x = r->cur[2 * i], y = r->cur[2 * i + 1];
        sum += x * x + y * y;
    }
    return sum;
}

/* ---- classical driver ------------------------------------------------ */

unsigned orc_gcd(unsigned a, unsigned b)
{
    if (a == 0) return b;
    if (b == 0) return a;
    if (a == b) return a;
    while ((a % b) > 0) { unsigned t = a % b; a = b; b = t; }
    return b;
}

/* qc_shor.c:868-883: bits n-1 .. M are read into powers 0 .. L-1 */
double orc_read_omega(uint64_t state_num, int L_size, int M_size)
{
    unsigned x_tilde = 0, power = 0;
    for (int i = L_size + M_size - 1; i >= M_size; i--) {
        x_tilde += (unsigned) ((state_num >> i) & 1) << power;
        power++;
    }
    return (double) x_tilde / (double) orc_int_pow(2, (unsigned) L_size);
}

/* double -> unsigned int the way gcc/x86-64 does it, including out of range */
static unsigned cast_u32(double d)
{
    if (!(d < 9223372036854775808.0) || !(d > -9223372036854775808.0)) return 0u;
    return (unsigned) (uint64_t) (int64_t) d;
}

/* qc_shor.c:806-846 */
void orc_continued_fraction_denominators(double omega, unsigned n, unsigned *out)
{
    unsigned *coeffs = (unsigned *) malloc((n ? n : 1) * sizeof(unsigned));
    for (unsigned i = 0; i < n; i++) {
        double omega_inv = 1.0 / omega;
        omega = omega_inv - (double) cast_u32(omega_inv);
        coeffs[i] = cast_u32(omega_inv - omega);
        unsigned denominator = 1, numerator = 0;
        for (int c = (int) i - 1; c >= 0; c--) {
            unsigned t = denominator;
            denominator = numerator + (denominator * coeffs[c]);
            numerator = t;
        }
        out[i] = denominator;
    }
    free(coeffs);
}

void orc_mt_seed(orc_mt19937 *g, unsigned long seed)
{
    if (seed == 0) seed = 4357;
    g->mt[0] = (uint32_t) (seed & 0xffffffffUL);
    for (int k = 1; k < 624; k++) {
        uint32_t p = g->mt[k - 1];
        g->mt[k] = 1812433253u * (p ^ (p >> 30)) + (uint32_t) k;
    }
    g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937 *g)
{
    if (g->idx >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
            uint32_t v = g->mt[(k + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            g->mt[k] = v;
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

double orc_mt_uniform(orc_mt19937 *g) { return (double) orc_mt_next(g) / 4294967296.0; }

#define ORC_NUM_CF 15          /* NUM_CONTINUED_FRACTIONS, qc_shor.c:121 */
#define ORC_TRIALS 10          /* TRIALS_PER_DENOMINATOR, qc_shor.c:122 */

/* qc_shor.c:912-964 */
int orc_find_period(orc_register *r, unsigned C, unsigned a, int pow_mode,
                    orc_mt19937 *g, unsigned *period, uint64_t *measured)
{
    unsigned den[ORC_NUM_CF];
    int found = 0;
    orc_reset_register(r);
    orc_quantum_computation(r, C, a, pow_mode);
    uint64_t s = orc_measure_state(r, orc_mt_uniform(g));
    if (measured) *measured = s;
    double omega = orc_read_omega(s, r->L_size, r->M_size);
    orc_continued_fraction_denominators(omega, ORC_NUM_CF, den);
    for (unsigned d = 0; d < ORC_NUM_CF && !found; d++) {
        for (unsigned m = 1; m < ORC_TRIALS + 1; m++) {
            *period = m * den[d];
            if (orc_int_pow(a, *period) % C == 1) { found = 1; break; }
        }
    }
    return found ? 0 : 3;
}

/* qc_shor.c:1003-1134 without the printing and timing */
int orc_shors_algorithm(orc_register *r, unsigned C, unsigned forced_a, int pow_mode,
                        orc_mt19937 *g, unsigned factors[2])
{
    unsigned period = 0;
    if (forced_a != 0) {
        if (orc_find_period(r, C, forced_a, pow_mode, g, &period, NULL) == 3) return 3;
        if (period % 2 != 0) return 3;
        if (orc_int_pow(forced_a, period / 2) % C == C - 1) return 3;
        factors[0] = orc_gcd(orc_int_pow(forced_a, period / 2) + 1, C);
        factors[1] = orc_gcd(orc_int_pow(forced_a, period / 2) - 1, C);
        return 0;
    }
    for (unsigned a = 2; a < C - 1; a++) {
        if (orc_find_period(r, C, a, pow_mode, g, &period, NULL) == 3) continue;
        if (period % 2 != 0) continue;
        /* qc_shor.c:1091 tests forced_trial_int (== 0 here), kept verbatim */
        if (orc_int_pow(forced_a, period / 2) % C == C - 1) continue;
        factors[0] = orc_gcd(orc_int_pow(a, period / 2) + 1, C);
        factors[1] = orc_gcd(orc_int_pow(a, period / 2) - 1, C);
        if (factors[0] == 1 || factors[1] == 1) continue;
        return 0;
    }
    return 3;
}

/* ---- synthetic state ---------------------- */

double orc_synthetic_u(uint64_t seed, uint64_t k)
{
    uint64_t z = seed + k + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double) (z >> 11) * 0x1.0p-53 - 0.5;
}

void orc_fill_synthetic(orc_register *r, uint64_t seed)
{
    for (uint64_t i = 0; i < r->num_states; i++) {
        r->cur[2 * i] = orc_synthetic_u(seed, 2 * i);
        r->cur[2 * i + 1] = orc_synthetic_u(seed, 2 * i + 1);
    }
}

/* threads the row loops run on: 1 unless built with -fopenmp */
int orc_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_scale(orc_register *r, double s)
{
    for (uint64_t i = 0; i < 2 * r->num_states; i++) r->cur[i] *= s;
}
