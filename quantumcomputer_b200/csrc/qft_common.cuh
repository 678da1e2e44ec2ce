// qft_common.cuh -- pieces shared by the two fused-QFT sweep kernels
// (qft_fused.cu: direct global<->register kernel; qft_pipeline.cu: TMA +
// mbarrier pipelined kernel): the sweep descriptor, the register butterfly
// networks, the external-twiddle power tree and the host-side sweep planner.
#pragma once

#include "qcs_internal.h"

#include <algorithm>
#include <math.h>
#include <vector>

namespace qft {


constexpr int kMaxSteps = 5;

struct sweep_step {
    int s;          // tile-local position of the lowest bit of the step
    int r;          // number of stages in the step (radix 2^r)
    int j;          // stage index of the step's top bit (physical bit - lo)
    int low_phys;   // physical position of the step's lowest bit
    int col_off;    // offset of the step's column-twiddle table (in double2; 2^s entries: the twiddle
                    // depends on the tile-local bits below the step only)
    int notw;       // 1: no register bit lies below the step (y == 0): the external twiddle is 1
};

// a diagonal two-qubit phase gate (c_phase_shift_gate) in fused form: multiply by
// (c + i s) every amplitude whose GLOBAL index has all bits of `mask` set
struct diag_gate {
    unsigned long long mask;
    double c, s;
};

struct sweep_desc {
    int a, g_lo, g_hi, t;   // tile = physical bits [0,a) U [g_lo,g_hi); t = a + g_hi - g_lo
    int lo;                 // lowest qubit of the transform (the reference's M_size)
    int n_steps;
    int sw;                 // swizzle: phys(e) = e ^ ((e >> sw) & 7)
    int inverse;            // 1: inverse_QFT of the reference, 0: its adjoint
    int wcol_total;         // total column-twiddle entries
    int hadamard_only;      // 1: the stages are bare Hadamards (no phase gates): Walsh-Hadamard sweep
    unsigned long long y_const;   // added to the per-tile y: register bits held by the rank (sharded layouts)
    unsigned long long tile_first; // first tile of this launch (sub-range launches: a rank's share, a slice)
    int slice_pos, slice_bits;     // a launch may cover one SLICE of the tiles: the tile numbers whose bits
    unsigned slice_val;            // [slice_pos, slice_pos + slice_bits) equal slice_val (slice_bits = 0: all)
    unsigned long long index_or;   // global index bits held by the rank (for the diagonal masks)
    int n_diag;                    // diagonal gates applied at the end of the sweep
    const diag_gate *diag;         // device pointer
    double scale;           // applied in the last step (1.0: nothing to do); the planner puts the
                            // whole (1/sqrt 2)^stages of a transform on its last sweep, where it
                            // is an exact power of two whenever the stage count is even
    sweep_step step[kMaxSteps];
};

// the k-th tile of a launch -> its tile number: tile_first + k, with the slice value deposited
__host__ __device__ __forceinline__ uint64_t tile_number(const sweep_desc &d, uint64_t k)
{
    const uint64_t t = d.tile_first + k;
    if (d.slice_bits == 0) return t;
    const uint64_t low = t & ((1ull << d.slice_pos) - 1ull);
    return ((t >> d.slice_pos) << (d.slice_pos + d.slice_bits)) | ((uint64_t) d.slice_val << d.slice_pos) | low;
}

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
template <int R, bool TW>
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
                const double2 dif = make_double2(u.x - v.x, u.y - v.y);
                x[start + m + span] = TW ? rot8<true>(dif, m * (8 / span)) : dif;
            }
        }
    }
}

// the adjoint network: decimation in time, negative exponent
template <int R, bool TW>
__device__ __forceinline__ void dit_forward(double2 (&x)[R])
{
#pragma unroll
    for (int span = 1; span <= R / 2; span <<= 1) {
#pragma unroll
        for (int start = 0; start < R; start += 2 * span) {
#pragma unroll
            for (int m = 0; m < span; m++) {
                const double2 u = x[start + m];
                const double2 v = TW ? rot8<false>(x[start + m + span], m * (8 / span)) : x[start + m + span];
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

// Shared-memory position (in amplitudes) of tile-local element e: e ^ ((e >> sw) & 7); sw = 28: linear
__host__ __device__ __forceinline__ unsigned tile_phys(unsigned e, int sw) { return e ^ ((e >> sw) & 7u); }

// The SPLIT-3 layout of a contiguous 2^t tile: TMA brings it in as a 3-D box {8 amplitudes} x {e >> 4} x
// {bit 3 of e} with the 128-byte hardware swizzle, so bit 3 of the index moves to the top of the tile
// address and the 16-byte chunk is XORed with bits 4..6 of e.  Every radix-16 step -- also the one on bits
// 0..3, which the plain swizzle e ^ ((e >> 3) & 7) serves with a 2-way bank conflict -- then touches 8
// distinct bank groups per quarter warp.  The map is GF(2)-linear: phys(e_base | d << s) = phys(e_base) ^ phys(d << s).
constexpr int kSwizzleSplit3 = 64;          // sweep_desc::sw value that selects it
__host__ __device__ __forceinline__ unsigned split3_phys(unsigned e, int t)
{
    return ((e & 8u) << (t - 4)) | ((e >> 4) << 3) | ((e & 7u) ^ ((e >> 4) & 7u));
}
// offset (amplitudes, t = 12) of element d of a radix-16 column behind tile + (phys(e_base) ^ (d & 7)):
//   digit at bits 8..11: phys(d << 8) = d << 7                       (no XOR part)
//   digit at bits 4..7 : phys(d << 4) = (d << 3) | (d & 7)           -> XOR part d & 7, rest d << 3
//   digit at bits 0..3 : phys(d)      = ((d & 8) << 8) | (d & 7)     -> XOR part d & 7, rest (d & 8) << 8
template <int SPLIT>
__host__ __device__ __forceinline__ constexpr unsigned split3_offset(int d)
{
    return SPLIT == 8 ? ((unsigned) d << 7) : SPLIT == 4 ? ((unsigned) d << 3) : (((unsigned) d & 8u) << 8);
}

struct tile_geom {
    int a, g_lo, sw;
    __device__ __forceinline__ uint64_t spread(unsigned e) const
    {
        return (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
    }
    __device__ __forceinline__ int phys(int s) const { return s < a ? s : g_lo + (s - a); }
    __device__ __forceinline__ unsigned swz(unsigned e) const { return tile_phys(e, sw); }
};

// LIN: the tile is laid out linearly in shared memory (strided sweeps of the pipelined kernel): the R
// elements of a column are `stride` apart, one multiply-add per address instead of the swizzle arithmetic
// SPLIT >= 0: radix-16 step with its digit at the compile-time position SPLIT (0, 4 or 8) of a
// contiguous 2^12 tile held in the split-3 layout (split3_phys): every address is the column's base
// XOR a 3-bit constant plus a compile-time offset -- eight XORs per column instead of arithmetic per element
template <int R, bool INV, bool TW, bool LIN = false, int SPLIT = -1>
__device__ __forceinline__ void run_step(double2 *__restrict__ amp, double2 *__restrict__ tile,
                                         const double2 *__restrict__ wcol, double2 wb, const tile_geom G,
                                         const sweep_step S, int t, uint64_t base, bool from_global,
                                         bool to_global, bool apply_scale, double scale,
                                         unsigned tid, unsigned nthreads, const diag_gate *diag = nullptr,
                                         int n_diag = 0, uint64_t index_or = 0)
{
    const unsigned n_cols = 1u << (t - S.r);
    const unsigned low_mask = (1u << S.s) - 1u;
    const unsigned lin_stride = 1u << S.s;
    // XOR swizzle e ^ ((e >> sw) & 7): a step whose digit lies above bit sw + 3 does not change the XOR
    // term, so its elements are lin_stride apart behind the swizzled column base
    const bool swz_linear = !LIN && S.s >= G.sw + 3;
    for (unsigned c = tid; c < n_cols; c += nthreads) {
        const unsigned e_base = ((c >> S.s) << (S.s + S.r)) | (c & low_mask);
        double2 *g = amp + base + G.spread(e_base);
        const uint64_t g_stride = 1ull << G.phys(S.s);
        double2 x[R];
        // split-3 layout: elements j = d & 7 sit behind tile + (phys(e_base) ^ j) at compile-time offsets
        unsigned sp_base = 0;
        if (SPLIT >= 0) sp_base = split3_phys(e_base, 12);
        if (SPLIT >= 0) {
#pragma unroll
            for (int d = 0; d < R; d++)
                x[d] = tile[(sp_base ^ (unsigned) (SPLIT == 8 ? 0 : (d & 7))) + split3_offset<SPLIT>(d)];
        } else if (from_global) {
            if (S.s == 0 && R >= 2) {
#pragma unroll
                for (int d = 0; d < R; d += 2) x[d] = ld256_lo(g + d, x[d + 1]);
            } else {
#pragma unroll
                for (int d = 0; d < R; d++) x[d] = g[(uint64_t) d * g_stride];
            }
        } else if (LIN) {
            // a running pointer: one add per element
            const double2 *col = tile + e_base;
#pragma unroll
            for (int d = 0; d < R; d++) {
                x[d] = *col;
                col += lin_stride;
            }
        } else if (swz_linear) {
            const double2 *col = tile + G.swz(e_base);
#pragma unroll
            for (int d = 0; d < R; d++) {
                x[d] = *col;
                col += lin_stride;
            }
        } else {
#pragma unroll
            for (int d = 0; d < R; d++) x[d] = tile[G.swz(e_base + ((unsigned) d << S.s))];
        }
        if (TW) {
            if (S.notw) {
                if (INV) dif_inverse<R, true>(x);
                else dit_forward<R, true>(x);
            } else {
                const double2 w = cmul(wb, wcol[c & low_mask]);      // the table is indexed by the tile bits below the step
                if (INV) {
                    dif_inverse<R, true>(x);
                    external_twiddle<R>(x, w);
                } else {
                    external_twiddle<R>(x, w);
                    dit_forward<R, true>(x);
                }
            }
        } else {
            dif_inverse<R, false>(x);       // H on r qubits in any order: a Walsh-Hadamard butterfly
        }
        if (apply_scale) {
            if (scale != 1.0) {
#pragma unroll
                for (int d = 0; d < R; d++) { x[d].x *= scale; x[d].y *= scale; }
            }
            // diagonal gates riding along (circuit.cu): the elements are in registers and their
            // full basis-state index is known.  The mask bits outside the step's digit are the
            // same for all R elements of the thread and are tested once per gate.
            if (n_diag > 0) {
                const uint64_t i0 = index_or | base | G.spread(e_base);
                const int dshift = G.phys(S.s);
                const uint64_t digit_bits = (uint64_t) (R - 1) << dshift;
                for (int gi = 0; gi < n_diag; gi++) {
                    const diag_gate dg = diag[gi];
                    const uint64_t rest = dg.mask & ~digit_bits;
                    if ((i0 & rest) != rest) continue;
                    const unsigned dm = (unsigned) ((dg.mask & digit_bits) >> dshift);
                    const double2 ph = make_double2(dg.c, dg.s);
#pragma unroll
                    for (int d = 0; d < R; d++)
                        if (((unsigned) d & dm) == dm) x[d] = cmul(x[d], ph);
                }
            }
        }
        if (SPLIT >= 0) {
#pragma unroll
            for (int d = 0; d < R; d++)
                tile[(sp_base ^ (unsigned) (SPLIT == 8 ? 0 : (d & 7))) + split3_offset<SPLIT>(d)] = x[d];
        } else if (to_global) {
            if (S.s == 0 && R >= 2) {
#pragma unroll
                for (int d = 0; d < R; d += 2) st256(g + d, x[d], x[d + 1]);
            } else {
#pragma unroll
                for (int d = 0; d < R; d++) g[(uint64_t) d * g_stride] = x[d];
            }
        } else if (LIN) {
            double2 *col = tile + e_base;
#pragma unroll
            for (int d = 0; d < R; d++) {
                *col = x[d];
                col += lin_stride;
            }
        } else if (swz_linear) {
            double2 *col = tile + G.swz(e_base);
#pragma unroll
            for (int d = 0; d < R; d++) {
                *col = x[d];
                col += lin_stride;
            }
        } else {
#pragma unroll
            for (int d = 0; d < R; d++) tile[G.swz(e_base + ((unsigned) d << S.s))] = x[d];
        }
    }
}

// radix-16 step of a split-3 tile (t = 12), digit at bit 8, 4 or 0
template <bool INV, bool TW>
__device__ __forceinline__ void dispatch_split3(double2 *tile, const double2 *wcol, double2 wb, const tile_geom G,
                                                const sweep_step S, uint64_t base, bool apply_scale, double scale,
                                                unsigned tid, unsigned nthreads, const diag_gate *diag = nullptr,
                                                int n_diag = 0, uint64_t index_or = 0)
{
    switch (S.s) {
        case 8: run_step<16, INV, TW, false, 8>(nullptr, tile, wcol, wb, G, S, 12, base, false, false, apply_scale, scale, tid, nthreads, diag, n_diag, index_or); break;
        case 4: run_step<16, INV, TW, false, 4>(nullptr, tile, wcol, wb, G, S, 12, base, false, false, apply_scale, scale, tid, nthreads, diag, n_diag, index_or); break;
        default: run_step<16, INV, TW, false, 0>(nullptr, tile, wcol, wb, G, S, 12, base, false, false, apply_scale, scale, tid, nthreads, diag, n_diag, index_or); break;
    }
}

template <bool INV, bool TW = true, bool LIN = false>
__device__ __forceinline__ void dispatch_step(double2 *amp, double2 *tile, const double2 *wcol, double2 wb,
                                              const tile_geom G, const sweep_step S, int t, uint64_t base,
                                              bool from_global, bool to_global, bool apply_scale, double scale,
                                              unsigned tid, unsigned nthreads, const diag_gate *diag = nullptr,
                                              int n_diag = 0, uint64_t index_or = 0)
{
    switch (S.r) {
        case 4: run_step<16, INV, TW, LIN>(amp, tile, wcol, wb, G, S, t, base, from_global, to_global, apply_scale, scale, tid, nthreads, diag, n_diag, index_or); break;
        case 3: run_step<8, INV, TW, LIN>(amp, tile, wcol, wb, G, S, t, base, from_global, to_global, apply_scale, scale, tid, nthreads, diag, n_diag, index_or); break;
        case 2: run_step<4, INV, TW, LIN>(amp, tile, wcol, wb, G, S, t, base, from_global, to_global, apply_scale, scale, tid, nthreads, diag, n_diag, index_or); break;
        default: run_step<2, INV, TW, LIN>(amp, tile, wcol, wb, G, S, t, base, from_global, to_global, apply_scale, scale, tid, nthreads, diag, n_diag, index_or); break;
    }
}

// exp(sign * i pi y / 2^j)
__device__ __forceinline__ double2 unit_phase(uint64_t y, int j, bool positive)
{
    double s, c;
    sincospi(ldexp((double) y, -j), &s, &c);
    return make_double2(c, positive ? s : -s);
}

// ---------------------------------------------------------------------------
// host-side planning
// ---------------------------------------------------------------------------
// where a sweep runs: the register's shard, or a staging buffer holding a slice of it
struct sweep_target {
    double2 *amp;
    unsigned n_bits;        // log2 of the number of amplitudes addressed
    cudaStream_t stream;
    int kind = QCS_K_TILE_SWEEP;    // kernel class the launch is accounted to
    double bytes = 0.0;             // algorithmic bytes of the launch (0: 32 B per amplitude of the tiles)
    int max_ctas = 0;               // > 0: use at most this many SMs (a concurrent launch takes the others)
};

struct sweep_plan {
    sweep_desc d;
    uint64_t n_tiles;
    int stages;
};

// pick the XOR-swizzle shift that minimises shared-memory bank conflicts over
// every shared-memory access pattern of the sweep (quarter-warp = 8 lanes must
// hit 8 distinct 16-byte bank groups)
// worst-case serialisation (extra wavefronts) of the shared-memory accesses of a
// sweep under the XOR swizzle `sw`; all_steps: every step reads and writes smem
inline long conflict_cost(const sweep_desc &d, int sw, bool all_steps)
{
    long cost = 0;
    for (int k = 0; k < d.n_steps; k++) {
        const sweep_step &S = d.step[k];
        const bool reads = all_steps || k > 0, writes = all_steps || k < d.n_steps - 1;
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
                    const unsigned ph = sw >= kSwizzleSplit3 ? split3_phys(e, sw - kSwizzleSplit3) : tile_phys(e, sw);
                    worst = std::max(worst, ++seen[ph & 7u]);
                }
                cost += (worst - 1) * ((reads ? 1 : 0) + (writes ? 1 : 0));
            }
        }
    }
    return cost;
}

inline int choose_swizzle(const sweep_desc &d)
{
    int best_sw = 4;
    long best_cost = -1;
    for (int sw = 3; sw <= 9; sw++) {
        const long cost = conflict_cost(d, sw, false);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_sw = sw; }
    }
    return best_sw;
}

inline void split_even(int total, int cap, std::vector<int> &out)
{
    out.clear();
    if (total <= 0) return;
    const int m = (total + cap - 1) / cap;
    for (int i = 0; i < m; i++) out.push_back(total / m + (i < total % m ? 1 : 0));
}

// sweeps of the inverse transform on qubits [lo, hi), top stages first.  Strided tiles keep
// the low `a` index bits (a contiguous run of 2^a amplitudes) for coalescing and spend the
// other t - a bits on stages: `a` is the largest value in [a_min, a_pref] ... that still gives
// the fewest sweeps (a longer run never costs a sweep).
// stage_lo >= lo (default lo): plan only the stages [stage_lo, hi); the twiddles still refer to
// the register's lowest qubit `lo`.
inline void plan_inverse(unsigned n_local, unsigned lo, unsigned hi, int T, int a_min, std::vector<sweep_plan> &plans,
                         int stage_lo = -1)
{
    plans.clear();
    if (stage_lo < (int) lo) stage_lo = (int) lo;
    const int t = std::min<int>(T, (int) n_local);
    const int first_hi = std::max<int>(stage_lo, t);
    int a = t;
    if ((int) n_local > t) {
        a_min = std::max(1, std::min(a_min, t - 1));
        const int rest = std::max(0, (int) hi - first_hi);
        auto sweeps_with = [&](int aa) { return (rest + (t - aa) - 1) / (t - aa); };
        a = a_min;
        while (a + 1 <= std::min(7, t - 1) && sweeps_with(a + 1) == sweeps_with(a_min)) a++;
    }
    int top = (int) hi;
    std::vector<int> sizes;
    if ((int) n_local > t) split_even(top - first_hi, t - a, sizes);
    struct raw { int a, g_lo, g_hi, s_lo, s_hi; };
    std::vector<raw> raws;
    for (int g : sizes) {
        raws.push_back({t - g, top - g, top, top - g, top});     // low run widened so the tile has exactly t bits
        top -= g;
    }
    if (top > stage_lo) raws.push_back({t, t, t, stage_lo, top});
    for (const raw &rw : raws) {
        sweep_plan p = {};
        sweep_desc &d = p.d;
        d.a = rw.a;
        d.g_lo = rw.g_lo;
        d.g_hi = rw.g_hi;
        d.t = rw.a + (rw.g_hi - rw.g_lo);
        d.lo = (int) lo;
        d.inverse = 1;
        d.hadamard_only = 0;
        d.y_const = 0;
        d.tile_first = 0;
        d.index_or = 0;
        d.n_diag = 0;
        d.diag = nullptr;
        p.stages = rw.s_hi - rw.s_lo;
        d.scale = 1.0;
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
            S.notw = S.low_phys <= (int) lo ? 1 : 0;
            S.col_off = off;
            off += S.notw ? 0 : (1 << S.s);
            l -= S.r;
        }
        d.wcol_total = off;
        d.sw = choose_swizzle(d);
        p.n_tiles = 1ull << (n_local - (unsigned) d.t);
        plans.push_back(p);
    }
    // one scaling for the whole transform, on its last sweep
    if (!plans.empty()) {
        const int total = (int) hi - stage_lo;
        plans.back().d.scale = total % 2 == 0 ? ldexp(1.0, -total / 2) : ldexp(0.70710678118654752440, -(total - 1) / 2);
    }
}


// turn an inverse plan into the plan of the adjoint circuit: sweeps and steps in
// reverse order, conjugated twiddles
inline void make_forward(std::vector<sweep_plan> &plans)
{
    std::reverse(plans.begin(), plans.end());
    for (sweep_plan &p : plans) {
        p.d.inverse = 0;
        std::reverse(p.d.step, p.d.step + p.d.n_steps);
        int off = 0;
        for (int k = 0; k < p.d.n_steps; k++) {
            p.d.step[k].col_off = off;
            off += p.d.step[k].notw ? 0 : (1 << p.d.step[k].s);
        }
        p.d.wcol_total = off;
        p.d.wcol_total = off;
        p.d.sw = choose_swizzle(p.d);
    }
}

}  // namespace qft

// ---- host entry points of qft_fused.cu used by the gate-stream scheduler (circuit.cu)
// Walsh-Hadamard tile sweeps for H on every qubit of [lo, hi) (hi <= n_local), top run first
int qcs_plan_hadamard_sweeps(const qcs_register *reg, unsigned lo, unsigned hi, std::vector<qft::sweep_plan> &plans);
// launch one planned sweep on the register's shard and stream
int qcs_launch_sweep_plan(qcs_register *reg, const qft::sweep_plan &plan);
// two consecutive sweeps of the shard as one L2-paired launch if they qualify (*paired), else nothing happens
int qcs_launch_sweep_pair(qcs_register *reg, const qft::sweep_plan &a, const qft::sweep_plan &b, bool *paired);
