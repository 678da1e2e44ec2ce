// qft_fused.cu -- the (inverse) QFT as a few fused tile sweeps.
//
// The reference applies inverse_QFT (qc_shor.c:678-690) as L Hadamards and
// L(L-1)/2 controlled phase gates, each a full pass over the state.  Written
// out, stage l of that circuit is "butterfly on bit l, then multiply every
// amplitude whose bit l is 1 by exp(i pi (x mod 2^j) / 2^j)" with x the value
// of the L register and j = l - M: a radix-2 decimation-in-frequency FFT with
// positive exponent and no final bit reversal.  Here the stages are grouped:
//
//   sweep  = one pass over HBM.  A tile of 2^t amplitudes (the low `a` index
//            bits, for coalescing, plus a run of stage bits [g_lo, g_hi)) is
//            brought on chip, all stages whose bit lies in the tile are applied,
//            and the tile is written back in place: 32 B per amplitude per sweep
//            instead of 32 B (H) or 8 B (C-phase) per gate.
//   step   = r <= 4 consecutive stages inside a sweep, done in registers as one
//            radix-2^r butterfly network (constant internal twiddles) followed
//            by one "external" twiddle w^k per element, w = exp(i pi y / 2^j),
//            y = the register bits below the step.  y splits into a part fixed
//            per tile (one sincospi per step per tile) and a part fixed per
//            thread column (one sincospi per column per kernel); the powers
//            w^2..w^(R-1) are formed by squaring / one extra multiply.
//
// Between steps the tile lives in shared memory (XOR-swizzled at 16-byte
// granularity so every step's access is bank-conflict free); the first step of
// a sweep loads straight from global memory and the last stores straight back,
// so a sweep with k steps makes k-1 shared-memory round trips.
//
// The forward QFT (the adjoint circuit) runs the same schedule backwards with
// conjugated twiddles: external twiddle first, then decimation-in-time
// butterflies.
//
// Amplitudes agree with the gate-by-gate reference order to ~1e-15 relative
// (tests/test_fused_gpu.py); they are not bit-identical, which is why
// QCS_OPT_FUSION = 0 keeps the gate-by-gate kernels available.
#include "qcs_internal.h"

#include <algorithm>
#include <math.h>

namespace {

constexpr int kMaxSteps = 5;

struct sweep_step {
    int s;          // tile-local position of the lowest bit of the step
    int r;          // number of stages in the step (radix 2^r)
    int j;          // stage index of the step's top bit (physical bit - lo)
    int low_phys;   // physical position of the step's lowest bit
    int col_off;    // offset of the step's column-twiddle table (in double2)
};

struct sweep_desc {
    int a, g_lo, g_hi, t;   // tile = physical bits [0,a) U [g_lo,g_hi); t = a + g_hi - g_lo
    int lo;                 // lowest qubit of the transform (the reference's M_size)
    int n_steps;
    int sw;                 // swizzle: phys(e) = e ^ ((e >> sw) & 7)
    int inverse;            // 1: inverse_QFT of the reference, 0: its adjoint
    int wcol_total;         // total column-twiddle entries
    double scale;           // (1/sqrt 2)^(stages in this sweep), applied in the last step
    sweep_step step[kMaxSteps];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 csqr(double2 a)
{
    return make_double2(a.x * a.x - a.y * a.y, 2.0 * a.x * a.y);
}

// v * exp(+i pi q / 8) (INV) or v * exp(-i pi q / 8) (!INV); q is a compile-time
// constant after unrolling
template <bool INV>
__device__ __forceinline__ double2 rot8(double2 v, int q)
{
    const double H = 0.70710678118654752440, C1 = 0.92387953251128673848, S1 = 0.38268343236508977173;
    double c, s;
    switch (q) {
        case 0: return v;
        case 4: return INV ? make_double2(-v.y, v.x) : make_double2(v.y, -v.x);
        case 2: return INV ? make_double2(H * (v.x - v.y), H * (v.x + v.y))
                           : make_double2(H * (v.x + v.y), H * (v.y - v.x));
        case 6: return INV ? make_double2(-H * (v.x + v.y), H * (v.x - v.y))
                           : make_double2(H * (v.y - v.x), -H * (v.x + v.y));
        case 1: c = C1; s = S1; break;
        case 3: c = S1; s = C1; break;
        case 5: c = -S1; s = C1; break;
        default: c = -C1; s = S1; break;
    }
    if (!INV) s = -s;
    return make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
}

// r stages of the reference circuit on R = 2^r register-resident amplitudes,
// x[d] = amplitude whose step digit is d (top stage bit = MSB of d).
// Decimation in frequency, positive exponent, unscaled: afterwards x[d] holds
// frequency bitrev(d).
template <int R>
__device__ __forceinline__ void dif_inverse(double2 (&x)[R])
{
#pragma unroll
    for (int span = R / 2; span >= 1; span >>= 1) {
#pragma unroll
        for (int start = 0; start < R; start += 2 * span) {
#pragma unroll
            for (int m = 0; m < span; m++) {
                const double2 u = x[start + m], v = x[start + m + span];
                x[start + m] = make_double2(u.x + v.x, u.y + v.y);
                x[start + m + span] = rot8<true>(make_double2(u.x - v.x, u.y - v.y), m * (8 / span));
            }
        }
    }
}

// the adjoint network: decimation in time, negative exponent
template <int R>
__device__ __forceinline__ void dit_forward(double2 (&x)[R])
{
#pragma unroll
    for (int span = 1; span <= R / 2; span <<= 1) {
#pragma unroll
        for (int start = 0; start < R; start += 2 * span) {
#pragma unroll
            for (int m = 0; m < span; m++) {
                const double2 u = x[start + m];
                const double2 v = rot8<false>(x[start + m + span], m * (8 / span));
                x[start + m] = make_double2(u.x + v.x, u.y + v.y);
                x[start + m + span] = make_double2(u.x - v.x, u.y - v.y);
            }
        }
    }
}

template <int R>
__device__ __forceinline__ constexpr int bitrev(int k)
{
    int out = 0;
    for (int b = 1, o = R >> 1; b < R; b <<= 1, o >>= 1)
        if (k & b) out |= o;
    return out;
}

// x[d] *= w^bitrev(d)
template <int R>
__device__ __forceinline__ void external_twiddle(double2 (&x)[R], double2 w)
{
    double2 p[R];
    p[1] = w;
#pragma unroll
    for (int k = 2; k < R; k++) p[k] = (k & 1) ? cmul(p[k - 1], w) : csqr(p[k >> 1]);
#pragma unroll
    for (int k = 1; k < R; k++) x[bitrev<R>(k)] = cmul(x[bitrev<R>(k)], p[k]);
}

__device__ __forceinline__ double2 ld256_lo(const double2 *p, double2 &hi)
{
    double a, b, c, d;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
    hi = make_double2(c, d);
    return make_double2(a, b);
}
__device__ __forceinline__ void st256(double2 *p, double2 lo, double2 hi)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(lo.x), "d"(lo.y), "d"(hi.x), "d"(hi.y) : "memory");
}

struct tile_geom {
    int a, g_lo, sw;
    __device__ __forceinline__ uint64_t spread(unsigned e) const
    {
        return (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
    }
    __device__ __forceinline__ int phys(int s) const { return s < a ? s : g_lo + (s - a); }
    __device__ __forceinline__ unsigned swz(unsigned e) const { return e ^ ((e >> sw) & 7u); }
};

template <int R, bool INV>
__device__ __forceinline__ void run_step(double2 *__restrict__ amp, double2 *__restrict__ tile,
                                         const double2 *__restrict__ wcol, double2 wb, const tile_geom G,
                                         const sweep_step S, int t, uint64_t base, bool first, bool last,
                                         double scale)
{
    const unsigned n_cols = 1u << (t - S.r);
    const unsigned low_mask = (1u << S.s) - 1u;
    for (unsigned c = threadIdx.x; c < n_cols; c += blockDim.x) {
        const unsigned e_base = ((c >> S.s) << (S.s + S.r)) | (c & low_mask);
        double2 *g = amp + base + G.spread(e_base);
        const uint64_t g_stride = 1ull << G.phys(S.s);
        double2 x[R];
        if (first) {
            if (S.s == 0 && R >= 2) {
#pragma unroll
                for (int d = 0; d < R; d += 2) x[d] = ld256_lo(g + d, x[d + 1]);
            } else {
#pragma unroll
                for (int d = 0; d < R; d++) x[d] = g[(uint64_t) d * g_stride];
            }
        } else {
#pragma unroll
            for (int d = 0; d < R; d++) x[d] = tile[G.swz(e_base + ((unsigned) d << S.s))];
        }
        const double2 w = cmul(wb, wcol[c]);
        if (INV) {
            dif_inverse<R>(x);
            external_twiddle<R>(x, w);
        } else {
            external_twiddle<R>(x, w);
            dit_forward<R>(x);
        }
        if (last) {
#pragma unroll
            for (int d = 0; d < R; d++) { x[d].x *= scale; x[d].y *= scale; }
            if (S.s == 0 && R >= 2) {
#pragma unroll
                for (int d = 0; d < R; d += 2) st256(g + d, x[d], x[d + 1]);
            } else {
#pragma unroll
                for (int d = 0; d < R; d++) g[(uint64_t) d * g_stride] = x[d];
            }
        } else {
#pragma unroll
            for (int d = 0; d < R; d++) tile[G.swz(e_base + ((unsigned) d << S.s))] = x[d];
        }
    }
}

template <bool INV>
__device__ __forceinline__ void dispatch_step(double2 *amp, double2 *tile, const double2 *wcol, double2 wb,
                                              const tile_geom G, const sweep_step S, int t, uint64_t base,
                                              bool first, bool last, double scale)
{
    switch (S.r) {
        case 4: run_step<16, INV>(amp, tile, wcol, wb, G, S, t, base, first, last, scale); break;
        case 3: run_step<8, INV>(amp, tile, wcol, wb, G, S, t, base, first, last, scale); break;
        case 2: run_step<4, INV>(amp, tile, wcol, wb, G, S, t, base, first, last, scale); break;
        default: run_step<2, INV>(amp, tile, wcol, wb, G, S, t, base, first, last, scale); break;
    }
}

// exp(sign * i pi y / 2^j)
__device__ __forceinline__ double2 unit_phase(uint64_t y, int j, bool positive)
{
    double s, c;
    sincospi(ldexp((double) y, -j), &s, &c);
    return make_double2(c, positive ? s : -s);
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_qft_sweep(double2 *__restrict__ amp, uint64_t n_tiles, const sweep_desc P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *tile = (double2 *) smem_raw;
    double2 *wcol = tile + (1u << P.t);
    double2 *wbase = wcol + P.wcol_total;        // [2][kMaxSteps]
    tile_geom G;
    G.a = P.a;
    G.g_lo = P.g_lo;
    G.sw = P.sw;
    const bool inv = P.inverse != 0;

    // per-column part of the external twiddle: fixed for the whole kernel
    for (int k = 0; k < P.n_steps; k++) {
        const sweep_step S = P.step[k];
        const unsigned n_cols = 1u << (P.t - S.r);
        for (unsigned c = threadIdx.x; c < n_cols; c += NT) {
            const unsigned e_base = ((c >> S.s) << (S.s + S.r)) | (c & ((1u << S.s) - 1u));
            uint64_t y = 0;
            if (S.low_phys > P.lo) y = (G.spread(e_base) & ((1ull << S.low_phys) - 1ull)) >> P.lo;
            wcol[S.col_off + c] = unit_phase(y, S.j, inv);
        }
    }

    unsigned parity = 0;
    const int lo_gap = P.g_lo - P.a;             // number of index bits between the two tile runs
    for (uint64_t tix = blockIdx.x; tix < n_tiles; tix += gridDim.x, parity ^= 1u) {
        // deposit the tile number into the index bits that are not in the tile
        const uint64_t base = lo_gap > 0 ? (((tix >> lo_gap) << P.g_hi) | ((tix & ((1ull << lo_gap) - 1ull)) << P.a))
                                         : (tix << P.t);
        if (threadIdx.x < (unsigned) P.n_steps) {
            const sweep_step S = P.step[threadIdx.x];
            uint64_t y = 0;
            if (S.low_phys > P.lo) y = (base & ((1ull << S.low_phys) - 1ull)) >> P.lo;
            wbase[parity * kMaxSteps + threadIdx.x] = unit_phase(y, S.j, inv);
        }
        __syncthreads();
        for (int k = 0; k < P.n_steps; k++) {
            const sweep_step S = P.step[k];
            const bool first = k == 0, last = k == P.n_steps - 1;
            const double2 wb = wbase[parity * kMaxSteps + k];
            if (inv) dispatch_step<true>(amp, tile, wcol + S.col_off, wb, G, S, P.t, base, first, last, P.scale);
            else dispatch_step<false>(amp, tile, wcol + S.col_off, wb, G, S, P.t, base, first, last, P.scale);
            if (!last) __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// host-side planning
// ---------------------------------------------------------------------------
struct sweep_plan {
    sweep_desc d;
    uint64_t n_tiles;
    int stages;
};

// pick the XOR-swizzle shift that minimises shared-memory bank conflicts over
// every shared-memory access pattern of the sweep (quarter-warp = 8 lanes must
// hit 8 distinct 16-byte bank groups)
int choose_swizzle(const sweep_desc &d)
{
    int best_sw = 4;
    long best_cost = -1;
    for (int sw = 3; sw <= 9; sw++) {
        long cost = 0;
        for (int k = 0; k < d.n_steps; k++) {
            const sweep_step &S = d.step[k];
            const bool reads = k > 0, writes = k < d.n_steps - 1;
            if (!reads && !writes) continue;
            const unsigned n_cols = 1u << (d.t - S.r);
            const unsigned lanes = n_cols < 8 ? n_cols : 8;
            for (unsigned c0 = 0; c0 < n_cols && c0 < 64; c0 += 8) {
                for (int dd = 0; dd < (1 << S.r); dd++) {
                    int seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    int worst = 0;
                    for (unsigned ln = 0; ln < lanes; ln++) {
                        const unsigned c = c0 + ln;
                        const unsigned e = (((c >> S.s) << (S.s + S.r)) | (c & ((1u << S.s) - 1u))) + ((unsigned) dd << S.s);
                        const unsigned ph = e ^ ((e >> sw) & 7u);
                        worst = std::max(worst, ++seen[ph & 7u]);
                    }
                    cost += (worst - 1) * ((reads ? 1 : 0) + (writes ? 1 : 0));
                }
            }
        }
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_sw = sw; }
    }
    return best_sw;
}

void split_even(int total, int cap, std::vector<int> &out)
{
    out.clear();
    if (total <= 0) return;
    const int m = (total + cap - 1) / cap;
    for (int i = 0; i < m; i++) out.push_back(total / m + (i < total % m ? 1 : 0));
}

// sweeps of the inverse transform on qubits [lo, hi), top stages first
void plan_inverse(unsigned n_local, unsigned lo, unsigned hi, int T, int a_req, std::vector<sweep_plan> &plans)
{
    plans.clear();
    const int t = std::min<int>(T, (int) n_local);
    const int a = (int) n_local <= t ? t : std::min(a_req, t - 1);
    int top = (int) hi;
    std::vector<int> sizes;
    const int first_hi = std::max<int>((int) lo, t);
    if ((int) n_local > t) split_even(top - first_hi, t - a, sizes);
    struct raw { int a, g_lo, g_hi, s_lo, s_hi; };
    std::vector<raw> raws;
    for (int g : sizes) {
        raws.push_back({a, top - g, top, top - g, top});
        top -= g;
    }
    if (top > (int) lo) raws.push_back({t, t, t, (int) lo, top});
    for (const raw &rw : raws) {
        sweep_plan p;
        sweep_desc &d = p.d;
        d.a = rw.a;
        d.g_lo = rw.g_lo;
        d.g_hi = rw.g_hi;
        d.t = rw.a + (rw.g_hi - rw.g_lo);
        d.lo = (int) lo;
        d.inverse = 1;
        p.stages = rw.s_hi - rw.s_lo;
        d.scale = pow(0.70710678118654752440, (double) p.stages);
        if (p.stages % 2 == 0) d.scale = ldexp(1.0, -p.stages / 2);      // exact power of two
        std::vector<int> rs;
        split_even(p.stages, 4, rs);
        d.n_steps = (int) rs.size();
        int l = rw.s_hi, off = 0;
        for (int k = 0; k < d.n_steps; k++) {
            sweep_step &S = d.step[k];
            S.r = rs[(size_t) k];
            S.low_phys = l - S.r;
            S.s = S.low_phys < d.a ? S.low_phys : d.a + (S.low_phys - d.g_lo);
            S.j = (l - 1) - (int) lo;
            S.col_off = off;
            off += 1 << (d.t - S.r);
            l -= S.r;
        }
        d.wcol_total = off;
        d.sw = choose_swizzle(d);
        p.n_tiles = 1ull << (n_local - (unsigned) d.t);
        plans.push_back(p);
    }
}

template <int NT, int MINB>
int launch_sweep(qcs_register *reg, const sweep_plan &p, size_t smem)
{
    auto kern = k_qft_sweep<NT, MINB>;
    QCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0;
    QCS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) return QCS_UNKNOWN_ERROR;
    uint64_t grid = (uint64_t) reg->sm_count * (uint64_t) per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    qcs_launch_begin(reg, QCS_K_TILE_SWEEP, 32.0 * (double) reg->N_local);
    kern<<<(unsigned) grid, NT, smem, reg->stream>>>(reg->amp, p.n_tiles, p.d);
    return qcs_launch_end(reg, QCS_K_TILE_SWEEP, "k_qft_sweep");
}

}  // namespace

int qcs_fused_qft(qcs_register *reg, unsigned lo, unsigned hi, bool inverse)
{
    if (hi > reg->n_local) {
        fprintf(stderr, "qcs: fused QFT over global qubits is handled by the distributed schedule\n");
        return QCS_BAD_ARGUMENTS;
    }
    const int T = reg->opt_tile_bits ? reg->opt_tile_bits : 12;
    std::vector<sweep_plan> plans;
    plan_inverse(reg->n_local, lo, hi, T, 4, plans);
    if (!inverse) {
        // adjoint: sweeps and steps in reverse order, conjugated twiddles
        std::reverse(plans.begin(), plans.end());
        for (sweep_plan &p : plans) {
            p.d.inverse = 0;
            std::reverse(p.d.step, p.d.step + p.d.n_steps);
            int off = 0;
            for (int k = 0; k < p.d.n_steps; k++) {
                p.d.step[k].col_off = off;
                off += 1 << (p.d.t - p.d.step[k].r);
            }
            p.d.sw = choose_swizzle(p.d);
        }
    }
    for (const sweep_plan &p : plans) {
        const size_t smem = ((size_t) 16 << p.d.t) + (size_t) 16 * (size_t) p.d.wcol_total + 16 * 2 * kMaxSteps;
        if (smem > reg->smem_optin) return QCS_BAD_ARGUMENTS;
        int rc;
        if (p.d.t >= 13) rc = launch_sweep<512, 1>(reg, p, smem);
        else if (p.d.t == 12) rc = launch_sweep<256, 2>(reg, p, smem);
        else rc = launch_sweep<128, 4>(reg, p, smem);
        if (rc != QCS_NO_ERROR) return rc;
    }
    return QCS_NO_ERROR;
}

// placeholder: the modular-exponentiation half of quantum_computation still runs
// gate by gate (modexp_fused.cu replaces this)
int qcs_fused_modexp(qcs_register *reg, unsigned C, const unsigned *A_per_gate, unsigned n_gates)
{
    const unsigned first = reg->n - n_gates;
    for (unsigned l = first; l < reg->n; l++) {
        if (l >= reg->n_local) return QCS_BAD_ARGUMENTS;
        QCS_TRY(qcs_k_hadamard_local(reg, l));
    }
    for (unsigned k = 0; k < n_gates; k++) QCS_TRY(qcs_k_amodc(reg, C, A_per_gate[k], (int) (first + k), false));
    return QCS_NO_ERROR;
}
