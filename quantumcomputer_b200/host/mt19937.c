#include "mt19937.h"

enum { MT_N = 624, MT_M = 397 };

void qcsh_rng_seed(qcsh_rng *g, unsigned long seed)
{
    uint32_t s = (uint32_t) (seed & 0xffffffffUL);
    if (seed == 0) s = 4357u;                 /* gsl mt.c convention */
    g->word[0] = s;
    for (int i = 1; i < MT_N; i++) {
        s = 1812433253u * (s ^ (s >> 30)) + (uint32_t) i;
        g->word[i] = s;
    }
    g->next = MT_N;
}

static void refill(qcsh_rng *g)
{
    uint32_t *w = g->word;
    for (int i = 0; i < MT_N; i++) {
        uint32_t upper = w[i] & 0x80000000u;
        uint32_t lower = w[(i + 1) % MT_N] & 0x7fffffffu;
        uint32_t mix = (upper | lower) >> 1;
        if (lower & 1u) mix ^= 0x9908b0dfu;
        w[i] = w[(i + MT_M) % MT_N] ^ mix;
    }
    g->next = 0;
}

uint32_t qcsh_rng_u32(qcsh_rng *g)
{
    if (g->next >= MT_N) refill(g);
    uint32_t y = g->word[g->next++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

double qcsh_rng_uniform(qcsh_rng *g)
{
    return (double) qcsh_rng_u32(g) / 4294967296.0;
}
