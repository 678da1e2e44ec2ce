/*
 * gsl_standin.h -- TEST INFRASTRUCTURE ONLY (oracle side).
 *
 * Header-only stand-in for the 17 GSL 2.6 symbols that the reference program
 * uses.  GSL itself is not installed in this image and is
 * not vendored by the reference, so the unmodified reference source is compiled
 * against this file to obtain oracle/_ref/.  Nothing here performs arithmetic on
 * amplitudes other than cos/sin (polar), x*x+y*y (abs2) and MT19937.
 *
 * Behaviour that matters for parity:
 *   - spmatrix "set" appends triplets in call order and grows when full (the
 *     controlled-phase builder emits 4N entries into a matrix sized 2N);
 *   - polar(r, t) = (r cos t, r sin t);
 *   - MT19937: Knuth-style seeding, seed 0 -> 4357, uniform = u32 / 2^32.
 *
 * Extension (not in GSL): gsl_rng carries an optional queue of forced uniform
 * values so the bridge can drive measure_state() with an explicit r.
 */
#ifndef QCS_GSL_STANDIN_H
#define QCS_GSL_STANDIN_H

#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
#ifndef M_SQRT1_2
#define M_SQRT1_2 0.70710678118654752440084436210
#endif

/* ---- complex -------------------------------------------------------- */
typedef struct { double dat[2]; } gsl_complex;

#define GSL_REAL(z) ((z).dat[0])
#define GSL_IMAG(z) ((z).dat[1])
#define GSL_SET_COMPLEX(zp, x, y) do { (zp)->dat[0] = (x); (zp)->dat[1] = (y); } while (0)
#define GSL_SET_REAL(zp, x) do { (zp)->dat[0] = (x); } while (0)
#define GSL_SET_IMAG(zp, y) do { (zp)->dat[1] = (y); } while (0)

static inline gsl_complex gsl_complex_rect(double x, double y)
{
    gsl_complex z; z.dat[0] = x; z.dat[1] = y; return z;
}
static inline gsl_complex gsl_complex_polar(double r, double theta)
{
    gsl_complex z; z.dat[0] = r * cos(theta); z.dat[1] = r * sin(theta); return z;
}
static inline double gsl_complex_abs2(gsl_complex z)
{
    double x = z.dat[0], y = z.dat[1];
    return x * x + y * y;
}
static inline double gsl_complex_abs(gsl_complex z)
{
    return hypot(z.dat[0], z.dat[1]);
}

/* ---- dense complex vector ------------------------------------------- */
typedef struct {
    size_t size;
    size_t stride;
    double *data;     /* interleaved (re, im) */
    void *block;
    int owner;
} gsl_vector_complex;

static inline gsl_vector_complex *gsl_vector_complex_alloc(size_t n)
{
    gsl_vector_complex *v = (gsl_vector_complex *) malloc(sizeof *v);
    if (!v) return NULL;
    v->data = (double *) calloc(2 * n, sizeof(double));
    if (!v->data) { free(v); return NULL; }
    v->size = n; v->stride = 1; v->block = NULL; v->owner = 1;
    return v;
}
static inline void gsl_vector_complex_free(gsl_vector_complex *v)
{
    if (v) { free(v->data); free(v); }
}
static inline gsl_complex gsl_vector_complex_get(const gsl_vector_complex *v, size_t i)
{
    gsl_complex z; z.dat[0] = v->data[2 * i]; z.dat[1] = v->data[2 * i + 1]; return z;
}
static inline void gsl_vector_complex_set(gsl_vector_complex *v, size_t i, gsl_complex z)
{
    v->data[2 * i] = z.dat[0]; v->data[2 * i + 1] = z.dat[1];
}
static inline void gsl_vector_complex_set_zero(gsl_vector_complex *v)
{
    for (size_t k = 0; k < 2 * v->size; k++) v->data[k] = 0.0;
}

/* ---- sparse complex matrix, COO only -------------------------------- */
#define GSL_SPMATRIX_COO 0
#define GSL_SPMATRIX_TRIPLET 0

typedef struct {
    size_t size1, size2;
    int *i;          /* COO row of each stored triplet   */
    int *p;          /* COO column of each stored triplet */
    double *data;    /* interleaved values                */
    size_t nzmax;
    size_t nz;
    int sptype;
} gsl_spmatrix_complex;

static inline gsl_spmatrix_complex *
gsl_spmatrix_complex_alloc_nzmax(size_t n1, size_t n2, size_t nzmax, int sptype)
{
    gsl_spmatrix_complex *m = (gsl_spmatrix_complex *) malloc(sizeof *m);
    if (!m) return NULL;
    if (nzmax == 0) nzmax = 1;
    m->size1 = n1; m->size2 = n2; m->nzmax = nzmax; m->nz = 0; m->sptype = sptype;
    m->i = (int *) malloc(nzmax * sizeof(int));
    m->p = (int *) malloc(nzmax * sizeof(int));
    m->data = (double *) malloc(2 * nzmax * sizeof(double));
    if (!m->i || !m->p || !m->data) { free(m->i); free(m->p); free(m->data); free(m); return NULL; }
    return m;
}
static inline void gsl_spmatrix_complex_free(gsl_spmatrix_complex *m)
{
    if (m) { free(m->i); free(m->p); free(m->data); free(m); }
}
static inline int gsl_spmatrix_complex_set_zero(gsl_spmatrix_complex *m)
{
    m->nz = 0; return 0;
}
static inline int
gsl_spmatrix_complex_set(gsl_spmatrix_complex *m, size_t row, size_t col, gsl_complex x)
{
    if (m->nz >= m->nzmax) {
        size_t cap = 2 * m->nzmax;
        int *ni = (int *) realloc(m->i, cap * sizeof(int));
        int *np = (int *) realloc(m->p, cap * sizeof(int));
        double *nd = (double *) realloc(m->data, 2 * cap * sizeof(double));
        if (ni) m->i = ni;
        if (np) m->p = np;
        if (nd) m->data = nd;
        if (!ni || !np || !nd) abort();   /* GSL's default handler aborts too */
        m->nzmax = cap;
    }
    m->i[m->nz] = (int) row;
    m->p[m->nz] = (int) col;
    m->data[2 * m->nz] = x.dat[0];
    m->data[2 * m->nz + 1] = x.dat[1];
    m->nz++;
    return 0;
}

/* ---- RNG: MT19937 only ---------------------------------------------- */
typedef struct { const char *name; } gsl_rng_type;
static const gsl_rng_type qcs_standin_mt19937_type = { "mt19937" };
static const gsl_rng_type *const gsl_rng_mt19937 = &qcs_standin_mt19937_type;

typedef struct {
    const gsl_rng_type *type;
    unsigned long mt[624];
    int mti;
    /* stand-in extension: forced outputs, consumed before the generator */
    double forced[8];
    int n_forced;
} gsl_rng;

static inline void gsl_rng_set(gsl_rng *r, unsigned long s)
{
    if (s == 0) s = 4357;
    r->mt[0] = s & 0xffffffffUL;
    for (int k = 1; k < 624; k++) {
        unsigned long prev = r->mt[k - 1];
        r->mt[k] = (1812433253UL * (prev ^ (prev >> 30)) + (unsigned long) k) & 0xffffffffUL;
    }
    r->mti = 624;
}
static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    gsl_rng *r = (gsl_rng *) calloc(1, sizeof *r);
    if (!r) return NULL;
    r->type = T;
    gsl_rng_set(r, 0);
    return r;
}
static inline void gsl_rng_free(gsl_rng *r) { free(r); }

static inline unsigned long qcs_standin_mt_next(gsl_rng *r)
{
    unsigned long *mt = r->mt;
    if (r->mti >= 624) {
        int k;
        for (k = 0; k < 624; k++) {
            unsigned long y = (mt[k] & 0x80000000UL) | (mt[(k + 1) % 624] & 0x7fffffffUL);
            unsigned long v = mt[(k + 397) % 624] ^ (y >> 1);
            if (y & 1UL) v ^= 0x9908b0dfUL;
            mt[k] = v;
        }
        r->mti = 0;
    }
    unsigned long y = mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680UL;
    y ^= (y << 15) & 0xefc60000UL;
    y ^= (y >> 18);
    return y & 0xffffffffUL;
}
static inline double gsl_rng_uniform(gsl_rng *r)
{
    if (r->n_forced > 0) {
        double v = r->forced[0];
        for (int k = 1; k < r->n_forced; k++) r->forced[k - 1] = r->forced[k];
        r->n_forced--;
        return v;
    }
    return (double) qcs_standin_mt_next(r) / 4294967296.0;
}

#endif /* QCS_GSL_STANDIN_H */
