// measure.cu -- probabilities: sum of |amp|^2 and the measurement scan.
//
//  * norm^2 (check_normalisation, testing_and_debug.c:28-37): warp-shuffle +
//    block reduction into per-block partials, then one block adds the partials
//    in a fixed order -- deterministic for a given grid, HBM-bound (16 B/amp).
//  * measure_state (qc_shor.c:272-306): the reference adds |amp_i|^2 in index
//    order into one double and stops at the first i with cum >= r.  Which i
//    that is depends on the rounding of that particular summation order, so
//    the scan kernel reproduces it exactly: all threads of one CTA stage
//    p_i = re*re + im*im (separate products and sum, like gsl_complex_abs2)
//    in shared memory, double-buffered, and one thread performs the
//    sequential additions and the `>=` test.
#include "qcs_internal.h"

#include <stdlib.h>
#include <vector>

namespace {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double abs2_ref(double2 a)
{
    // gsl_complex_abs2: x*x + y*y, no contraction
    return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kRedThreads)
k_norm2_partial(const double2 *__restrict__ amp, uint64_t n_amps, double *__restrict__ partials)
{
    __shared__ double warp_part[kRedThreads / 32];
    const uint64_t stride = (uint64_t) gridDim.x * kRedThreads;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    uint64_t i = (uint64_t) blockIdx.x * kRedThreads + threadIdx.x;
    for (; i + 3 * stride < n_amps; i += 4 * stride) {
        const double2 a = amp[i], b = amp[i + stride], c = amp[i + 2 * stride], d = amp[i + 3 * stride];
        s0 += abs2_ref(a); s1 += abs2_ref(b); s2 += abs2_ref(c); s3 += abs2_ref(d);
    }
    for (; i < n_amps; i += stride) s0 += abs2_ref(amp[i]);
    double s = warp_sum((s0 + s1) + (s2 + s3));
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kRedThreads / 32 ? warp_part[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) partials[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(1024)
k_sum_partials(const double *__restrict__ partials, unsigned n, double *__restrict__ out)
{
    __shared__ double warp_part[32];
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < n; i += 1024) s += partials[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = warp_sum(warp_part[threadIdx.x]);
        if (threadIdx.x == 0) *out = v;
    }
}

// ---------------------------------------------------------------------------
// sequential-semantics scan (single CTA)
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanPerThread = 2;   // 2 x 2048 doubles x 2 buffers = 32 KiB static
constexpr int kScanChunk = kScanThreads * kScanPerThread;

struct scan_result {
    double cum;
    unsigned long long index;
    int found;
    int pad;
};

__global__ void __launch_bounds__(kScanThreads)
k_measure_scan(const double2 *__restrict__ amp, uint64_t limit, double cum_in, double r,
               scan_result *__restrict__ out)
{
    __shared__ double p[2][kScanChunk];
    __shared__ int s_found;
    if (threadIdx.x == 0) s_found = 0;
    double cum = cum_in;
    uint64_t hit = 0;
    const uint64_t n_chunks = (limit + kScanChunk - 1) / kScanChunk;

    auto stage = [&](uint64_t chunk, int buf) {
        const uint64_t base = chunk * kScanChunk;
#pragma unroll
        for (int u = 0; u < kScanPerThread; u++) {
            const uint64_t i = base + (uint64_t) u * kScanThreads + threadIdx.x;
            p[buf][u * kScanThreads + threadIdx.x] = i < limit ? abs2_ref(amp[i]) : 0.0;
        }
    };

    if (n_chunks > 0) stage(0, 0);
    __syncthreads();
    for (uint64_t c = 0; c < n_chunks; c++) {
        const int buf = (int) (c & 1);
        if (threadIdx.x == 0) {
            const uint64_t base = c * kScanChunk;
            const uint64_t len = limit - base < (uint64_t) kScanChunk ? limit - base : (uint64_t) kScanChunk;
            const double *pc = p[buf];
            for (uint64_t j = 0; j < len; j++) {
                cum = __dadd_rn(cum, pc[j]);            // cumulative_prob += abs2, qc_shor.c:286
                if (cum >= r) { hit = base + j; s_found = 1; break; }   // qc_shor.c:289
            }
        }
        // the other threads fetch the next chunk while thread 0 walks this one
        if (c + 1 < n_chunks) stage(c + 1, buf ^ 1);
        __syncthreads();
        if (s_found) break;
    }
    if (threadIdx.x == 0) {
        out->cum = cum;
        out->index = hit;
        out->found = s_found;
    }
}


// ---------------------------------------------------------------------------
// Exact parallel emulation of the sequential scan (large registers).
//
// s_{i+1} = RN(s_i + p_i) with p_i >= 0.  While s stays inside one binade
// [2^e, 2^(e+1)) it is an integer multiple a of ulp = 2^(e-52), and
// RN(s + p) = (a + RNint(p / ulp)) ulp, where the round-to-nearest-even tie
// break depends only on the parity of a.  One element is therefore the map
//     a -> a + (a even ? de : do)
// and such maps compose associatively, so whole chunks (and runs of chunks)
// collapse into one (de, do) pair computed in parallel.
//
//  pass 1  per-chunk approximate sums (any order; only used for bounds)
//  pass 2  prefix of those sums; a chunk is CLEAN(e) when rigorous error
//          bounds (relative margin delta >= 4 N 2^-53) prove that the exact
//          sequential sum stays inside binade e throughout the chunk AND stays
//          below r; ZERO when all its p are 0 -- or all below half an ulp of
//          the smallest possible running sum, so that none of them moves it
//          (pass 1 also keeps each chunk's largest p); DUAL(e) when it stays
//          below r and inside binades e and e + 1 (both maps are computed,
//          the walk picks); otherwise SEQ
//  pass 3  (de, do) of every CLEAN chunk, then of every uniform super-chunk
//  pass 4  one CTA walks the summaries in index order carrying the exact s;
//          a SEQ chunk (binade crossing, the neighbourhood of r) is refined on
//          the spot into 128 sub-chunks of 32 elements, classified the same
//          way from the exact s at its start; only the SEQ sub-chunks are
//          added element by element with the `>=` test of qc_shor.c:289
//          (a Shor state after the inverse QFT has ~200 SEQ chunks: 4096
//          dependent additions each made the walk 22-37 ms at n = 30)
//
// The result is the index the reference's loop (qc_shor.c:283-292) returns.
// ---------------------------------------------------------------------------
constexpr int kChunkBits = 12;
constexpr int kChunk = 1 << kChunkBits;
constexpr int kSuperBits = 8;
constexpr int kSuper = 1 << kSuperBits;
constexpr int kWalkBlock = 512;                  // summaries staged per round of the walk (>= 2 kSuper: both maps of DUAL chunks)
static_assert(kWalkBlock >= 2 * kSuper, "the chunk-level staging keeps two maps per chunk");
constexpr int kCodeSeq = -1, kCodeZero = -2;     // otherwise: binade exponent + 2000
// DUAL(e) = kCodeDual + e + 2000: the chunk stays below r and inside binades e and e + 1, but the bounds cannot
// tell on which side of 2^(e+1) it runs.  Pass 3 computes its map for BOTH binades; the walk, which knows the
// exact running sum, picks one (or refines the chunk when the sum really crosses inside it).
constexpr int kCodeDual = -100000;
__host__ __device__ __forceinline__ bool is_dual(int cd) { return cd <= kCodeDual + 4000; }
constexpr double kSubDelta = 0x1p-36;            // relative bound on a running sum over <= 4096 additions, with margin

struct pair64 { long long de, od; };

__device__ __forceinline__ pair64 compose(pair64 f1, pair64 f2)
{
    pair64 r;
    r.de = f1.de + ((f1.de & 1) ? f2.od : f2.de);
    r.od = f1.od + (((1 + f1.od) & 1) ? f2.od : f2.de);
    return r;
}

// floor(log2 x) of a positive normal x (-1023 for a subnormal one: never classified) and 2^e (0 below the
// normal range, +inf at e = 1024) -- bit patterns instead of frexp / ldexp, which are library calls
__device__ __forceinline__ int binade_of(double x)
{
    return (int) (((unsigned long long) __double_as_longlong(x) >> 52) & 0x7ff) - 1023;
}
__device__ __forceinline__ double pow2(int e)
{
    return e < -1022 ? 0.0 : __longlong_as_double((long long) (e + 1023) << 52);
}

// Class of a (sub-)chunk from rigorous bounds lo <= (exact sequential sum at its start), (the same at its
// end) <= hi, lo > 0, and its largest addend:
//   ZERO (skipped)  every addend is below half an ulp of ANY running sum >= 2^e, e = binade of lo:
//                   RN(s + p) = s all the way through, whichever binade s is in.  This keeps the plateaus
//                   of a symmetric state cheap -- a Shor state after the inverse QFT parks the running
//                   sum within rounding of 2^-k for runs of chunks, where no bound can name the binade;
//   CLEAN(e)        the sum stays inside binade e and below r;
//   DUAL(e)         (chunks only) below r, inside binades e and e + 1: one boundary possibly crossed;
//   SEQ             otherwise.
template <bool DUAL_OK>
__device__ __forceinline__ int classify(double lo, double hi, double biggest, double r)
{
    const int e = binade_of(lo);
    if (biggest < pow2(e - 53)) return kCodeZero;
    if (hi < r && e > -960) {
        if (hi < pow2(e + 1)) return e + 2000;
        if (DUAL_OK && hi < pow2(e + 2)) return kCodeDual + e + 2000;
    }
    return kCodeSeq;
}

// 2^(52 - e): p * scale = p / ulp of binade e (exact: a power of two; -960 < e keeps it finite)
__device__ __forceinline__ double binade_scale(int e)
{
    return __longlong_as_double((long long) (1023 + 52 - e) << 52);
}

// the map of one addend p inside the binade whose scale is given.  x = p / ulp, exactly, and x < 2^52: a
// p >= 2^e would carry the sum out of binade e.  RN(a + x) - a = RNint(x) unless x is an exact tie k + 1/2,
// where the parity of a decides: an even a rounds like x itself (half to even), an odd a takes the other
// neighbour.  RNint by the 2^52 trick, read back from the bit pattern (which also covers y = 2^53): no
// branch, no 64-bit float <-> integer conversion -- this runs once per amplitude of the scanned range.
__device__ __forceinline__ pair64 element_map(double p, double scale)
{
    const double x = p * scale;
    pair64 r;
    const double y = __dadd_rn(x, 0x1p52);
    const double t = __dadd_rn(x, -__dadd_rn(y, -0x1p52));       // exact
    r.de = __double_as_longlong(y) - __double_as_longlong(0x1p52);
    r.od = r.de + (t == 0.5 ? 1 : 0) - (t == -0.5 ? 1 : 0);
    return r;
}

// Pass 1 runs in up to kScanSegments launches over consecutive ranges of chunks.  A launch first adds up what
// the launches before it found (one atomicAdd per CTA and segment): once that provably exceeds r -- by the
// margin delta that covers the rounding of the sequential sum and of these tree sums -- the variate is
// reached before this segment, and this launch and the ones behind it do nothing but record where the
// summaries end (*n_valid).  On average half of the pass is saved; r = huge never stops.
constexpr int kScanSegments = 16;

__global__ void k_scan_init(double *acc, unsigned long long *n_valid, unsigned long long n_chunks)
{
    if (threadIdx.x < kScanSegments) acc[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) *n_valid = n_chunks;
}

__global__ void __launch_bounds__(256)
k_chunk_sums(const double2 *__restrict__ amp, uint64_t limit, uint64_t c_begin, uint64_t c_end, double *__restrict__ csum,
             double *__restrict__ cmax, int segment, double *__restrict__ acc, unsigned long long *__restrict__ n_valid,
             double approx_cum_in, double r, double delta)
{
    __shared__ double warp_part[8], warp_big[8];
    if (segment > 0) {
        double before = approx_cum_in;
        for (int t = 0; t < segment; t++) before += acc[t];
        if (before * (1.0 - delta) > r) {
            if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(n_valid, (unsigned long long) c_begin);
            return;
        }
    }
    double mine = 0.0;                                          // thread 0: this CTA's share of the segment
    for (uint64_t c = c_begin + blockIdx.x; c < c_end; c += gridDim.x) {
        const uint64_t base = c << kChunkBits;
        double s = 0.0, big = 0.0;
#pragma unroll 4
        for (int u = 0; u < kChunk / 256; u++) {
            const uint64_t i = base + (uint64_t) u * 256 + threadIdx.x;
            if (i < limit) {
                const double p = abs2_ref(amp[i]);
                s += p;
                big = fmax(big, p);
            }
        }
        s = warp_sum(s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) big = fmax(big, __shfl_xor_sync(0xffffffffu, big, o));
        if ((threadIdx.x & 31) == 0) { warp_part[threadIdx.x >> 5] = s; warp_big[threadIdx.x >> 5] = big; }
        __syncthreads();
        if (threadIdx.x < 32) {
            double v = threadIdx.x < 8 ? warp_part[threadIdx.x] : 0.0;
            double m = threadIdx.x < 8 ? warp_big[threadIdx.x] : 0.0;
            v = warp_sum(v);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (threadIdx.x == 0) { csum[c] = v; cmax[c] = m; mine += v; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && mine != 0.0) atomicAdd(acc + segment, mine);
}

// single CTA: exclusive prefix over chunk sums + classification
__global__ void __launch_bounds__(1024)
k_classify(const double *__restrict__ csum, const double *__restrict__ cmax, uint64_t n_all, const unsigned long long *__restrict__ n_valid,
           double cum_in, double r, double delta, int *__restrict__ code, int *__restrict__ super_code)
{
    // chunks from *n_valid on (a multiple of kSuper, or all of them) were not summarised: r is reached before them
    const uint64_t n_chunks = *n_valid < n_all ? (uint64_t) *n_valid : n_all;
    __shared__ double warp_tot[32], warp_off[32];
    __shared__ double carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = cum_in;
    __syncthreads();
    for (uint64_t seg = 0; seg < n_chunks; seg += 1024) {
        const uint64_t c = seg + threadIdx.x;
        const double v = c < n_chunks ? csum[c] : 0.0;
        const double big = c < n_chunks ? cmax[c] : 0.0;
        double x = v;                                // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        double excl = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) excl = 0.0;
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {                             // scan of the 32 warp totals on top of the carry
            double t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            double before = __shfl_up_sync(0xffffffffu, t, 1);
            if (lane == 0) before = 0.0;
            const double carry = carry_s;
            warp_off[lane] = carry + before;
            __syncwarp();
            if (lane == 31) carry_s = carry + t;
        }
        __syncthreads();
        const double P = warp_off[warp] + excl;      // approximate sum before the chunk (additions only)
        if (c < n_chunks) {
            const double lo = P * (1.0 - delta), hi = (P + v) * (1.0 + delta);
            code[c] = v == 0.0 ? kCodeZero : lo > 0.0 ? classify<true>(lo, hi, big, r) : kCodeSeq;
        }
    }
    __syncthreads();
    // super-chunks: uniform when every chunk is ZERO or CLEAN with one common binade (a warp each)
    const uint64_t n_super = (n_all + kSuper - 1) >> kSuperBits;
    for (uint64_t sc = warp; sc < n_super; sc += 32) {
        if ((sc << kSuperBits) >= n_chunks) {                    // never walked; SEQ keeps the walk exact regardless
            if (lane == 0) super_code[sc] = kCodeSeq;
            continue;
        }
        int lo_cd = 0x7fffffff, hi_cd = -1;
        bool seq = false;
#pragma unroll
        for (int k = 0; k < kSuper / 32; k++) {
            const uint64_t c = (sc << kSuperBits) + k * 32 + lane;
            const int cd = c < n_chunks ? code[c] : kCodeZero;
            seq |= cd == kCodeSeq || is_dual(cd);
            if (cd >= 0) { lo_cd = min(lo_cd, cd); hi_cd = max(hi_cd, cd); }
        }
        seq = __any_sync(0xffffffffu, seq);
        lo_cd = __reduce_min_sync(0xffffffffu, lo_cd);
        hi_cd = __reduce_max_sync(0xffffffffu, hi_cd);
        if (lane == 0) super_code[sc] = seq || (hi_cd >= 0 && lo_cd != hi_cd) ? kCodeSeq : (hi_cd >= 0 ? hi_cd : kCodeZero);
    }
}

// (de, do) of every CLEAN chunk.  256 threads in two roles over a double-buffered chunk of probabilities:
// warps 0-3 fetch the NEXT clean chunk (|amp|^2 into shared memory, 16 loads in flight per thread) while
// warps 4-7 compose the current one -- 32 consecutive elements per thread, an ordered tree over the lanes,
// a short chain over the 4 warps.  With one role per CTA (load, then compose) the kernel sat at 4.8 TB/s:
// 24 warps per SM, idle memory during every compose phase (ncu: long_scoreboard 11.5 per issue).
constexpr int kMapThreads = 256;
constexpr int kMapRun = 32;                                      // consecutive elements per composing thread
constexpr int kMapBuf = kChunk + kChunk / kMapRun;               // one pad per run: conflict-free both ways
constexpr size_t kMapSmem = 2 * kMapBuf * sizeof(double);

__global__ void __launch_bounds__(kMapThreads, 3)
k_chunk_maps(const double2 *__restrict__ amp, uint64_t limit, uint64_t n_all, const unsigned long long *__restrict__ n_valid,
             const int *__restrict__ code, pair64 *__restrict__ maps, pair64 *__restrict__ maps_upper)
{
    extern __shared__ double map_smem[];
    const uint64_t n_chunks = *n_valid < n_all ? (uint64_t) *n_valid : n_all;
    __shared__ pair64 warp_map[4];
    const bool loader = threadIdx.x < 128;
    const int t = threadIdx.x & 127;
    auto next_clean = [&](uint64_t c) {
        while (c < n_chunks && code[c] < 0 && !is_dual(code[c])) c += gridDim.x;
        return c;
    };
    auto fetch = [&](uint64_t c, int buf) {
        double *p = map_smem + buf * kMapBuf;
        const uint64_t base = c << kChunkBits;
#pragma unroll 1
        for (int h = 0; h < kChunk / 128; h += 16) {             // 16 loads issued before the first use
            double2 v[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const uint64_t i = base + (uint64_t) ((h + u) * 128 + t);
                v[u] = i < limit ? amp[i] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int i = (h + u) * 128 + t;
                p[i + i / kMapRun] = abs2_ref(v[u]);
            }
        }
    };
    uint64_t c = next_clean(blockIdx.x);
    int buf = 0;
    if (loader && c < n_chunks) fetch(c, 0);
    __syncthreads();
    while (c < n_chunks) {
        const uint64_t c_next = next_clean(c + gridDim.x);
        if (loader) {
            if (c_next < n_chunks) fetch(c_next, buf ^ 1);
        } else {
            const double *p = map_smem + buf * kMapBuf;
            const int cd = code[c];
            const bool dual = is_dual(cd);
            const int e = dual ? cd - kCodeDual - 2000 : cd - 2000;
            for (int half = 0; half < (dual ? 2 : 1); half++) {          // DUAL: binade e into maps, e + 1 into maps_upper
                const double scale = binade_scale(e + half);
                pair64 f = {0, 0};
#pragma unroll 4
                for (int k = 0; k < kMapRun; k++) f = compose(f, element_map(p[t * (kMapRun + 1) + k], scale));
                // ordered reduction: lane 0 ends with the composition of lanes 0..31 in order
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    pair64 g;
                    g.de = __shfl_down_sync(0xffffffffu, f.de, o);
                    g.od = __shfl_down_sync(0xffffffffu, f.od, o);
                    if ((t & 31) + o < 32 && ((t & 31) & (2 * o - 1)) == 0) f = compose(f, g);
                }
                if ((t & 31) == 0) warp_map[t >> 5] = f;
                asm volatile("bar.sync 1, 128;" ::: "memory");      // the 4 composing warps only
                if (t == 0)
                    (half ? maps_upper : maps)[c] = compose(compose(warp_map[0], warp_map[1]), compose(warp_map[2], warp_map[3]));
                if (dual) asm volatile("bar.sync 1, 128;" ::: "memory");   // warp_map is written again
            }
        }
        __syncthreads();
        c = c_next;
        buf ^= 1;
    }
}

// (de, do) of every uniform super-chunk: a warp each, 8 consecutive chunk maps per lane, ordered tree
__global__ void __launch_bounds__(256)
k_super_maps(uint64_t n_chunks, uint64_t n_super, const int *__restrict__ code, const int *__restrict__ super_code,
             const pair64 *__restrict__ maps, pair64 *__restrict__ super_maps)
{
    const uint64_t sc = ((uint64_t) blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (sc >= n_super || super_code[sc] < 0) return;             // the whole warp
    pair64 f = {0, 0};
#pragma unroll
    for (int k = 0; k < kSuper / 32; k++) {
        const uint64_t c = (sc << kSuperBits) + lane * (kSuper / 32) + k;
        if (c < n_chunks && code[c] >= 0) f = compose(f, maps[c]);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        pair64 g;
        g.de = __shfl_down_sync(0xffffffffu, f.de, o);
        g.od = __shfl_down_sync(0xffffffffu, f.od, o);
        if (lane + o < 32 && (lane & (2 * o - 1)) == 0) f = compose(f, g);
    }
    if (lane == 0) super_maps[sc] = f;
}

struct walk_result {
    double cum;
    unsigned long long index;
    int found;
    int bad;           // an invariant failed: caller falls back to the plain sequential scan
    unsigned seq_chunks;        // chunks refined into sub-chunks              (QCS_MEASURE_DEBUG prints both)
    unsigned mixed_supers;      // super-chunks walked chunk by chunk
};

// apply a (de, do) map valid in binade e to the exact running sum.  Integer arithmetic on the bit
// pattern (s >= 0 is normal with exponent field e + 1023, a = 2^52 + mantissa): the serial walk applies
// thousands of these back to back, and two ldexp calls each were most of its time.
__device__ __forceinline__ bool apply_map(double &s, pair64 f, int e)
{
    const unsigned long long bits = (unsigned long long) __double_as_longlong(s);
    if ((long long) (bits >> 52) != (long long) e + 1023) return false;
    const long long a = (long long) ((bits & 0xfffffffffffffull) | (1ull << 52));
    const long long a2 = a + ((a & 1) ? f.od : f.de);
    if (a2 < (1ll << 52) || a2 >= (1ll << 53)) return false;
    s = __longlong_as_double((long long) (((unsigned long long) (e + 1023) << 52) | ((unsigned long long) a2 & 0xfffffffffffffull)));
    return true;
}

// RECORD = false: stop at the first index whose running sum reaches r (measure_state).
// RECORD = true : never stop; write the exact running sum at the start of every super-chunk and,
//                 last, after the final element (bnd[0 .. n_super]) -- the variate-independent part
//                 of the scan, computed once for any number of samples (qcs_sample_states).
template <bool RECORD>
__global__ void __launch_bounds__(1024)
k_exact_walk(const double2 *__restrict__ amp, uint64_t limit, uint64_t n_chunks, double cum_in, double r,
             const int *__restrict__ code, const int *__restrict__ super_code, const pair64 *__restrict__ maps,
             const pair64 *__restrict__ maps_upper, const double *__restrict__ csum,
             const pair64 *__restrict__ super_maps, walk_result *__restrict__ out, double *__restrict__ bnd)
{
    __shared__ double p[kChunk];
    __shared__ int s_code[kWalkBlock];
    __shared__ pair64 s_map[kWalkBlock];
    __shared__ int sub_code[kChunk / 32];
    __shared__ pair64 sub_map[kChunk / 32];
    __shared__ double warp_tot[32];
    __shared__ int warp_code[32];
    __shared__ pair64 warp_map[32];
    __shared__ double s_run;           // exact running sum at the start of the chunk being refined
    __shared__ long long req;          // index requested by thread 0 (-1: none)
    __shared__ int s_found, s_bad;
    const int lane = threadIdx.x & 31;
    const uint64_t n_super = (n_chunks + kSuper - 1) >> kSuperBits;
    double s = cum_in;
    uint64_t hit = 0;
    unsigned n_seq = 0, n_mixed = 0;
    if (threadIdx.x == 0) { s_found = 0; s_bad = 0; req = -1; }
    __syncthreads();

    for (uint64_t sb = 0; sb < n_super && !s_found && !s_bad; sb += kWalkBlock) {
        const uint64_t sb_end = sb + kWalkBlock < n_super ? sb + kWalkBlock : n_super;
        uint64_t sc_next = sb;
        while (sc_next < sb_end && !s_found && !s_bad) {
            // stage the super-chunk summaries [sc_next, sb_end)
            __syncthreads();
            if (threadIdx.x < kWalkBlock && sc_next + threadIdx.x < sb_end) {
                s_code[threadIdx.x] = super_code[sc_next + threadIdx.x];
                s_map[threadIdx.x] = s_code[threadIdx.x] >= 0 ? super_maps[sc_next + threadIdx.x] : pair64{0, 0};
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                req = -1;
                uint64_t sc = sc_next;
                for (; sc < sb_end; sc++) {
                    const int cd = s_code[sc - sc_next];
                    if (RECORD) bnd[sc] = s;
                    if (cd == kCodeZero) continue;
                    if (cd == kCodeSeq) { req = (long long) sc; break; }
                    if (!apply_map(s, s_map[sc - sc_next], cd - 2000)) { s_bad = 1; break; }
                }
            }
            __syncthreads();
            if (s_bad) break;
            if (req < 0) { sc_next = sb_end; break; }
            const uint64_t sc = (uint64_t) req;
            n_mixed++;
            // walk the chunks of the non-uniform super-chunk sc
            const uint64_t c_begin = sc << kSuperBits;
            const uint64_t c_end = c_begin + kSuper < n_chunks ? c_begin + kSuper : n_chunks;
            __syncthreads();
            if (threadIdx.x < kSuper && c_begin + threadIdx.x < c_end) {
                const int cd = code[c_begin + threadIdx.x];
                s_code[threadIdx.x] = cd;
                s_map[threadIdx.x] = (cd >= 0 || is_dual(cd)) ? maps[c_begin + threadIdx.x] : pair64{0, 0};
                if (is_dual(cd)) s_map[kSuper + threadIdx.x] = maps_upper[c_begin + threadIdx.x];
            }
            __syncthreads();
            uint64_t c_next = c_begin;
            while (c_next < c_end && !s_found && !s_bad) {
                if (threadIdx.x == 0) {
                    req = -1;
                    uint64_t c = c_next;
                    for (; c < c_end; c++) {
                        const int cd = s_code[c - c_begin];
                        if (cd == kCodeZero) continue;
                        if (cd == kCodeSeq) { req = (long long) c; break; }
                        if (is_dual(cd)) {
                            // the exact sum decides the side of 2^(e+1): at or above it the whole chunk runs in binade
                            // e + 1 (it ends below 2^(e+2)); provably below it to the end, in binade e (the margin covers
                            // 4096 additions and the tree sum); else the sum crosses inside this chunk: refine it
                            const int e = cd - kCodeDual - 2000;
                            const double edge = pow2(e + 1);
                            if (s >= edge) {
                                if (!apply_map(s, s_map[kSuper + (c - c_begin)], e + 1)) { s_bad = 1; break; }
                            } else if ((s + csum[c]) * (1.0 + kSubDelta) < edge) {
                                if (!apply_map(s, s_map[c - c_begin], e)) { s_bad = 1; break; }
                            } else {
                                req = (long long) c;
                                break;
                            }
                            continue;
                        }
                        if (!apply_map(s, s_map[c - c_begin], cd - 2000)) { s_bad = 1; break; }
                    }
                }
                __syncthreads();
                if (s_bad || req < 0) break;
                const uint64_t c = (uint64_t) req;
                const uint64_t base = c << kChunkBits;
                n_seq++;
                // Refine the chunk: 128 sub-chunks of 32 elements, 8 threads each.  The same three
                // classes one level down, now measured from the EXACT running sum at the chunk's start, so
                // the bound only has to cover <= 4096 additions (2 * gamma_4096 < 2^-39 < kSubDelta).
                // Thread 0 then adds element by element only inside the sub-chunks that cross a binade
                // boundary or reach r -- a handful of 32-element runs instead of 4096 dependent additions.
                if (threadIdx.x == 0) s_run = s;
                double q[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t i = base + 4u * threadIdx.x + k;
                    q[k] = i < limit ? abs2_ref(amp[i]) : 0.0;
                    p[4 * threadIdx.x + k] = q[k];
                }
                const double mine = (q[0] + q[1]) + (q[2] + q[3]);
                double x = mine;                                        // inclusive scan over the CTA
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                double excl = __shfl_up_sync(0xffffffffu, x, 1);
                if (lane == 0) excl = 0.0;
                if (lane == 31) warp_tot[threadIdx.x >> 5] = x;
                __syncthreads();
                const double before = warp_sum(lane < (int) (threadIdx.x >> 5) ? warp_tot[lane] : 0.0);
                const int g0 = lane & ~7;
                const double p_start = s_run + (before + __shfl_sync(0xffffffffu, excl, g0));
                const double p_end = s_run + (before + __shfl_sync(0xffffffffu, x, g0 | 7));
                double gsum = mine;                                     // exactly 0 iff all 32 addends are 0
                gsum += __shfl_xor_sync(0xffffffffu, gsum, 1);
                gsum += __shfl_xor_sync(0xffffffffu, gsum, 2);
                gsum += __shfl_xor_sync(0xffffffffu, gsum, 4);
                double gmax = fmax(fmax(q[0], q[1]), fmax(q[2], q[3]));
                gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 1));
                gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 2));
                gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 4));
                const double lo = p_start * (1.0 - kSubDelta), hi = p_end * (1.0 + kSubDelta);
                const int cd = gsum == 0.0 ? kCodeZero : lo > 0.0 ? classify<false>(lo, hi, gmax, r) : kCodeSeq;
                pair64 f = {0, 0};
                if (cd >= 0) {
                    const double scale = binade_scale(cd - 2000);
#pragma unroll
                    for (int k = 0; k < 4; k++) f = compose(f, element_map(q[k], scale));
                }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {                       // ordered: lane g0 ends with lanes g0..g0+7
                    pair64 g;
                    g.de = __shfl_down_sync(0xffffffffu, f.de, o);
                    g.od = __shfl_down_sync(0xffffffffu, f.od, o);
                    if (((lane & 7) & (2 * o - 1)) == 0) f = compose(f, g);
                }
                if ((lane & 7) == 0) { sub_code[threadIdx.x >> 3] = cd; sub_map[threadIdx.x >> 3] = f; }
                // one level up: the warp's 4 sub-chunks collapse into one map when none is SEQ and the
                // CLEAN ones agree on the binade (a ZERO sub-chunk carries the identity map {0, 0})
                const int lo_cd = __reduce_min_sync(0xffffffffu, cd >= 0 ? cd : 0x7fffffff);
                const int hi_cd = __reduce_max_sync(0xffffffffu, cd >= 0 ? cd : -1);
                const bool any_seq = __any_sync(0xffffffffu, cd == kCodeSeq);
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) {
                    pair64 g;
                    g.de = __shfl_down_sync(0xffffffffu, f.de, o);
                    g.od = __shfl_down_sync(0xffffffffu, f.od, o);
                    if ((lane & (2 * o - 1)) == 0) f = compose(f, g);
                }
                if (lane == 0) {
                    const int w = threadIdx.x >> 5;
                    warp_code[w] = any_seq || (hi_cd >= 0 && lo_cd != hi_cd) ? kCodeSeq : (hi_cd >= 0 ? hi_cd : kCodeZero);
                    warp_map[w] = f;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    const uint64_t len = limit - base < (uint64_t) kChunk ? limit - base : (uint64_t) kChunk;
                    for (int w = 0; w < 32 && !s_found && !s_bad; w++) {
                        const int wd = warp_code[w];
                        if (wd == kCodeZero) continue;
                        if (wd >= 0) {
                            if (!apply_map(s, warp_map[w], wd - 2000)) s_bad = 1;
                            continue;
                        }
                        for (int j = 4 * w; j < 4 * w + 4 && !s_found && !s_bad; j++) {
                            const int sd = sub_code[j];
                            if (sd == kCodeZero) continue;
                            if (sd >= 0) {
                                if (!apply_map(s, sub_map[j], sd - 2000)) s_bad = 1;
                                continue;
                            }
                            double v[32];                                // fetched ahead of the dependent chain
#pragma unroll
                            for (int k = 0; k < 32; k++) v[k] = p[32 * j + k];
#pragma unroll
                            for (int k = 0; k < 32; k++) {
                                if ((uint64_t) (32 * j + k) >= len) break;
                                s = __dadd_rn(s, v[k]);                 // qc_shor.c:286
                                if (!RECORD && s >= r) { hit = base + 32 * j + k; s_found = 1; break; }   // qc_shor.c:289
                            }
                        }
                    }
                }
                __syncthreads();
                c_next = c + 1;
            }
            sc_next = sc + 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (RECORD) bnd[n_super] = s;
        out->cum = s;
        out->index = hit;
        out->found = s_found;
        out->bad = s_bad;
        out->seq_chunks = n_seq;
        out->mixed_supers = n_mixed;
    }
}

}  // namespace

int qcs_k_norm2_local(qcs_register *reg, double *out_host)
{
    uint64_t want = (reg->N_local + (uint64_t) kRedThreads * 8 - 1) / ((uint64_t) kRedThreads * 8);
    uint64_t cap = (uint64_t) reg->sm_count * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    if (want > reg->partials_cap) want = reg->partials_cap;
    const unsigned grid = (unsigned) want;
    qcs_launch_begin(reg, QCS_K_REDUCE, 16.0 * (double) reg->N_local);
    k_norm2_partial<<<grid, kRedThreads, 0, reg->stream>>>(reg->amp, reg->N_local, reg->d_partials);
    QCS_TRY(qcs_launch_end(reg, QCS_K_REDUCE, "k_norm2_partial"));
    qcs_launch_begin(reg, QCS_K_REDUCE, 8.0 * grid);
    k_sum_partials<<<1, 1024, 0, reg->stream>>>(reg->d_partials, grid, (double *) reg->d_small);
    QCS_TRY(qcs_launch_end(reg, QCS_K_REDUCE, "k_sum_partials"));
    QCS_CUDA(cudaMemcpyAsync(reg->h_small, reg->d_small, sizeof(double), cudaMemcpyDeviceToHost, reg->stream));
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    *out_host = *(double *) reg->h_small;
    return QCS_NO_ERROR;
}

static int measure_scan_sequential(qcs_register *reg, double cum_in, double r, uint64_t limit,
                                   int *found, uint64_t *index, double *cum_out)
{
    scan_result *d_res = (scan_result *) reg->d_small;
    qcs_launch_begin(reg, QCS_K_REDUCE, 16.0 * (double) limit);
    k_measure_scan<<<1, kScanThreads, 0, reg->stream>>>(reg->amp, limit, cum_in, r, d_res);
    QCS_TRY(qcs_launch_end(reg, QCS_K_REDUCE, "k_measure_scan"));
    QCS_CUDA(cudaMemcpyAsync(reg->h_small, d_res, sizeof(scan_result), cudaMemcpyDeviceToHost, reg->stream));
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    const scan_result *h = (const scan_result *) reg->h_small;
    *found = h->found;
    *index = h->index;
    *cum_out = h->cum;
    return QCS_NO_ERROR;
}

// scratch of the parallel scan (lazily allocated, sized for the whole shard)
struct scan_buffers {
    pair64 *maps = nullptr, *maps_upper = nullptr, *super_maps = nullptr;
    double *csum = nullptr, *cmax = nullptr, *acc = nullptr;
    unsigned long long *n_valid = nullptr;
    int *code = nullptr, *super_code = nullptr;
};

static int scan_scratch(qcs_register *reg, scan_buffers &b)
{
    const uint64_t cap_chunks = (reg->N_local + kChunk - 1) >> kChunkBits;
    const uint64_t cap_super = (cap_chunks + kSuper - 1) >> kSuperBits;
    if (!reg->d_meas) {
        const size_t bytes = cap_chunks * (2 * sizeof(double) + sizeof(int) + 2 * sizeof(pair64)) +
                             cap_super * (sizeof(int) + sizeof(pair64)) + 64 + (kScanSegments + 1) * sizeof(double);
        QCS_CUDA(cudaMalloc(&reg->d_meas, bytes));
    }
    unsigned char *at = (unsigned char *) reg->d_meas;
    b.maps = (pair64 *) at;            at += cap_chunks * sizeof(pair64);
    b.maps_upper = (pair64 *) at;      at += cap_chunks * sizeof(pair64);
    b.super_maps = (pair64 *) at;      at += cap_super * sizeof(pair64);
    b.csum = (double *) at;            at += cap_chunks * sizeof(double);
    b.cmax = (double *) at;            at += cap_chunks * sizeof(double);
    b.acc = (double *) at;             at += kScanSegments * sizeof(double);
    b.n_valid = (unsigned long long *) at; at += sizeof(double);
    b.code = (int *) at;               at += cap_chunks * sizeof(int);
    b.super_code = (int *) at;
    return QCS_NO_ERROR;
}

// grid-stride over the chunks with exactly one wave of CTAs: a partial second wave would run at a fraction
// of the occupancy for as long as the first (k_chunk_maps fits 6 CTAs per SM, not 8: it cost 1.5x)
static unsigned scan_grid(const qcs_register *reg, uint64_t n_chunks, int ctas_per_sm)
{
    uint64_t grid = n_chunks;
    const uint64_t cap = (uint64_t) reg->sm_count * (uint64_t) (ctas_per_sm > 0 ? ctas_per_sm : 1);
    if (grid > cap) grid = cap;
    return (unsigned) (grid < 1 ? 1 : grid);
}

static int resident_sums()
{
    static int n = 0;
    if (!n && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_chunk_sums, 256, 0) != cudaSuccess) n = 4;
    return n;
}

static int resident_maps()
{
    static int n = 0;
    if (!n) {
        if (cudaFuncSetAttribute(k_chunk_maps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMapSmem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_chunk_maps, kMapThreads, kMapSmem) != cudaSuccess || n < 1)
            n = -1;
    }
    return n;
}

static double scan_delta(const qcs_register *reg)
{
    // rigorous relative margin: |sequential - exact| <= (N-1) u and the same for
    // the tree sums, u = 2^-53; 2^(n+3-53) covers both with slack (n: ALL qubits of the register,
    // so the margin also covers an approximate running sum handed over from the shards before)
    return ldexp(1.0, (int) reg->n + 3 - 53);
}

// pass 1: per-chunk approximate sums (and largest addends) of amp[first .. first + limit), in segments that
// stop once the variate r is provably behind them (r = 1e300: everything is summarised)
static int scan_sums(qcs_register *reg, uint64_t first, uint64_t limit, double approx_cum_in, double r)
{
    scan_buffers b;
    QCS_TRY(scan_scratch(reg, b));
    const uint64_t n_chunks = (limit + kChunk - 1) >> kChunkBits;
    k_scan_init<<<1, 32, 0, reg->stream>>>(b.acc, b.n_valid, (unsigned long long) n_chunks);
    const int n_seg = (n_chunks >= (1ull << 14) && r < 1e299) ? kScanSegments : 1;
    uint64_t per_seg = (n_chunks + (uint64_t) n_seg - 1) / (uint64_t) n_seg;
    per_seg = (per_seg + kSuper - 1) & ~(uint64_t) (kSuper - 1);             // whole super-chunks
    qcs_launch_begin(reg, QCS_K_REDUCE, 16.0 * (double) limit);
    for (int sg = 0; sg < n_seg; sg++) {
        const uint64_t c_begin = (uint64_t) sg * per_seg;
        if (c_begin >= n_chunks) break;
        const uint64_t c_end = c_begin + per_seg < n_chunks ? c_begin + per_seg : n_chunks;
        k_chunk_sums<<<scan_grid(reg, c_end - c_begin, resident_sums()), 256, 0, reg->stream>>>(
            reg->amp + first, limit, c_begin, c_end, b.csum, b.cmax, sg, b.acc, b.n_valid, approx_cum_in, r, scan_delta(reg));
    }
    return qcs_launch_end(reg, QCS_K_REDUCE, "k_chunk_sums");
}

// passes 2 and 3: classification against r given the (approximate) running sum before `first`, then
// the (de, do) maps of the clean chunks and uniform super-chunks
static int scan_maps(qcs_register *reg, uint64_t first, double approx_cum_in, double r, uint64_t limit)
{
    scan_buffers b;
    QCS_TRY(scan_scratch(reg, b));
    const double2 *amp = reg->amp + first;
    const uint64_t n_chunks = (limit + kChunk - 1) >> kChunkBits;
    const uint64_t n_super = (n_chunks + kSuper - 1) >> kSuperBits;
    const double delta = scan_delta(reg);
    qcs_launch_begin(reg, QCS_K_REDUCE, 12.0 * (double) n_chunks);
    k_classify<<<1, 1024, 0, reg->stream>>>(b.csum, b.cmax, n_chunks, b.n_valid, approx_cum_in, r, delta, b.code, b.super_code);
    QCS_TRY(qcs_launch_end(reg, QCS_K_REDUCE, "k_classify"));
    qcs_launch_begin(reg, QCS_K_REDUCE, 16.0 * (double) limit * (r < 1.0 ? (r > 0.0 ? r : 0.0) : 1.0));
    if (resident_maps() < 1) return QCS_UNKNOWN_ERROR;
    k_chunk_maps<<<scan_grid(reg, n_chunks, resident_maps()), kMapThreads, kMapSmem, reg->stream>>>(amp, limit, n_chunks, b.n_valid, b.code, b.maps, b.maps_upper);
    QCS_TRY(qcs_launch_end(reg, QCS_K_REDUCE, "k_chunk_maps"));
    qcs_launch_begin(reg, QCS_K_REDUCE, 20.0 * (double) n_chunks);
    k_super_maps<<<(unsigned) ((n_super + 7) / 8), 256, 0, reg->stream>>>(n_chunks, n_super, b.code, b.super_code,
                                                                              b.maps, b.super_maps);
    return qcs_launch_end(reg, QCS_K_REDUCE, "k_super_maps");
}

// pass 4: the exact walk from the exact running sum before `first`.  d_bnd != nullptr: record the
// running sum at the super-chunk boundaries instead of searching for r.
static int scan_walk(qcs_register *reg, uint64_t first, double cum_in, double r, uint64_t limit,
                     int *found, uint64_t *index, double *cum_out, double *d_bnd, int *bad)
{
    scan_buffers b;
    QCS_TRY(scan_scratch(reg, b));
    const double2 *amp = reg->amp + first;
    const uint64_t n_chunks = (limit + kChunk - 1) >> kChunkBits;
    const uint64_t n_super = (n_chunks + kSuper - 1) >> kSuperBits;
    walk_result *d_res = (walk_result *) reg->d_small;
    qcs_launch_begin(reg, QCS_K_REDUCE, 20.0 * (double) n_super);
    if (d_bnd)
        k_exact_walk<true><<<1, 1024, 0, reg->stream>>>(amp, limit, n_chunks, cum_in, r, b.code, b.super_code, b.maps,
                                                        b.maps_upper, b.csum, b.super_maps, d_res, d_bnd);
    else
        k_exact_walk<false><<<1, 1024, 0, reg->stream>>>(amp, limit, n_chunks, cum_in, r, b.code, b.super_code, b.maps,
                                                         b.maps_upper, b.csum, b.super_maps, d_res, nullptr);
    QCS_TRY(qcs_launch_end(reg, QCS_K_REDUCE, "k_exact_walk"));
    QCS_CUDA(cudaMemcpyAsync(reg->h_small, d_res, sizeof(walk_result), cudaMemcpyDeviceToHost, reg->stream));
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    const walk_result *h = (const walk_result *) reg->h_small;
    static const bool debug = getenv("QCS_MEASURE_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr, "qcs measure walk: %llu chunks, %llu super-chunks: %u chunks refined, %u super-chunks chunk by chunk\n",
                (unsigned long long) n_chunks, (unsigned long long) n_super, h->seq_chunks, h->mixed_supers);
    if (debug && getenv("QCS_MEASURE_DEBUG")[0] == '2') {      // which chunks, and why
        std::vector<int> code(n_chunks);
        std::vector<double> csum(n_chunks), cmax(n_chunks);
        cudaMemcpy(code.data(), b.code, n_chunks * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(csum.data(), b.csum, n_chunks * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(cmax.data(), b.cmax, n_chunks * sizeof(double), cudaMemcpyDeviceToHost);
        double run = cum_in;
        int shown = 0;
        for (uint64_t c = 0; c < n_chunks && run < r && shown < 60; c++) {
            if (code[c] == -1) {
                fprintf(stderr, "  chunk %llu: prefix %.17g sum %.6g max %.6g\n", (unsigned long long) c, run, csum[c], cmax[c]);
                shown++;
            }
            run += csum[c];
        }
    }
    *bad = h->bad;
    *found = h->found;
    *index = first + h->index;
    *cum_out = h->cum;
    return QCS_NO_ERROR;
}

// chunk / super-chunk summaries of amp[first .. first + limit) for the variate r (pass a huge r
// for variate-independent summaries), then the exact walk
static int parallel_scan(qcs_register *reg, uint64_t first, double cum_in, double r, uint64_t limit,
                         int *found, uint64_t *index, double *cum_out, double *d_bnd, int *bad)
{
    QCS_TRY(scan_sums(reg, first, limit, cum_in, d_bnd ? 1e300 : r));
    QCS_TRY(scan_maps(reg, first, cum_in, r, limit));
    return scan_walk(reg, first, cum_in, r, limit, found, index, cum_out, d_bnd, bad);
}

// ---- the three parts on their own (sharded registers: api.cu runs parts 1 and 2 on all shards at once)
bool qcs_k_scan_parallel_ok(const qcs_register *reg, uint64_t limit)
{
    return limit >= (1ull << 17) && !reg->opt_measure_sequential;
}

int qcs_k_scan_sums(qcs_register *reg, uint64_t limit, double *approx_total)
{
    QCS_TRY(scan_sums(reg, 0, limit, 0.0, 1e300));
    scan_buffers b;
    QCS_TRY(scan_scratch(reg, b));
    const uint64_t n_chunks = (limit + kChunk - 1) >> kChunkBits;
    if (n_chunks > 0xffffffffull) return QCS_BAD_ARGUMENTS;
    qcs_launch_begin(reg, QCS_K_REDUCE, 8.0 * (double) n_chunks);
    k_sum_partials<<<1, 1024, 0, reg->stream>>>(b.csum, (unsigned) n_chunks, (double *) reg->d_small);
    QCS_TRY(qcs_launch_end(reg, QCS_K_REDUCE, "k_sum_partials"));
    QCS_CUDA(cudaMemcpyAsync(reg->h_small, reg->d_small, sizeof(double), cudaMemcpyDeviceToHost, reg->stream));
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    *approx_total = *(double *) reg->h_small;
    return QCS_NO_ERROR;
}

int qcs_k_scan_maps(qcs_register *reg, double approx_cum_in, double r, uint64_t limit)
{
    return scan_maps(reg, 0, approx_cum_in, r, limit);
}

int qcs_k_scan_walk(qcs_register *reg, double cum_in, double r, uint64_t limit, int *found, uint64_t *index,
                    double *cum_out, int *bad)
{
    // qc_shor.c:286-289: a running sum that already reaches r stops at the first index
    if (limit > 0 && cum_in >= r) {
        *bad = 0;
        return measure_scan_sequential(reg, cum_in, r, 1, found, index, cum_out);
    }
    return scan_walk(reg, 0, cum_in, r, limit, found, index, cum_out, nullptr, bad);
}

int qcs_k_measure_scan(qcs_register *reg, double cum_in, double r, uint64_t limit,
                       int *found, uint64_t *index, double *cum_out)
{
    // qc_shor.c:286-289: the sum only grows, so a running sum that already reaches r (r <= 0 on the
    // first shard; gsl_rng_uniform can return exactly 0.0) stops at the first index whatever its
    // amplitude -- the chunk classification below would skip all-zero chunks without that test
    if (limit > 0 && cum_in >= r) return measure_scan_sequential(reg, cum_in, r, 1, found, index, cum_out);
    // small registers: the plain sequential scan is already fast
    if (limit < (1ull << 17) || reg->opt_measure_sequential)
        return measure_scan_sequential(reg, cum_in, r, limit, found, index, cum_out);
    int bad = 0;
    QCS_TRY(parallel_scan(reg, 0, cum_in, r, limit, found, index, cum_out, nullptr, &bad));
    if (bad) {
        fprintf(stderr, "qcs: measure_state: binade invariant failed, falling back to the sequential GPU scan\n");
        return measure_scan_sequential(reg, cum_in, r, limit, found, index, cum_out);
    }
    return QCS_NO_ERROR;
}

// Many variates against one state (qcs_sample_states, single-GPU registers): the exact sequential
// running sum at every boundary of 2^20 amplitudes is computed ONCE (two passes over the state,
// independent of the variates); a variate is then located by a binary search over the boundaries --
// the running sums are monotone -- and one exact scan of the 2^20 amplitudes it falls into.
// indices[k] is what measure_state would return for r[k] (qc_shor.c:283-292).
int qcs_k_sample_many(qcs_register *reg, uint64_t n_shots, const double *r, unsigned long long *indices,
                      bool *handled)
{
    *handled = false;
    // qc_shor.c:283: index N-1 is the fall-through, so the last shard stops one element early
    const uint64_t limit = reg->rank == reg->world - 1 ? reg->N_local - 1 : reg->N_local;
    if (reg->N_local < (1ull << 22) || reg->opt_measure_sequential || n_shots < 3) return QCS_NO_ERROR;
    const uint64_t n_chunks = (limit + kChunk - 1) >> kChunkBits;
    const uint64_t n_super = (n_chunks + kSuper - 1) >> kSuperBits;
    double *d_bnd = nullptr;
    QCS_CUDA(cudaMalloc((void **) &d_bnd, (n_super + 1) * sizeof(double)));
    std::vector<double> bnd((size_t) n_super + 1);
    std::vector<double> rank_end((size_t) reg->world, 0.0);   // exact running sum after each shard
    int found = 0, bad = 0, rc = QCS_NO_ERROR;
    uint64_t index = 0;
    double total = 0.0, carry = 0.0;
    // sharded: chunk sums and maps on all shards at once (an all-gather of the approximate totals gives
    // every shard the running sum before it, to the accuracy the classification bounds need); only the
    // walk that records the exact boundary sums is handed from rank to rank in index order, once
    if (reg->world > 1) {
        double mine = 0.0;
        rc = qcs_k_scan_sums(reg, limit, &mine);
        std::vector<double> all((size_t) reg->world * 2);
        const double pack[2] = {mine, (double) rc};
        const int rc2 = qcs_dist_allgather_doubles(reg, pack, 2, all.data());
        if (rc == QCS_NO_ERROR) rc = rc2;
        double before = 0.0;
        for (int s = 0; s < reg->world && rc == QCS_NO_ERROR; s++) {
            if (all[(size_t) s * 2 + 1] != 0.0) rc = (int) all[(size_t) s * 2 + 1];
            if (s < reg->rank) before += all[(size_t) s * 2];
        }
        if (rc == QCS_NO_ERROR) rc = scan_maps(reg, 0, before, 1e300, limit);
    }
    for (int turn = 0; turn < reg->world; turn++) {
        double mine = 0.0;
        if (turn == reg->rank && rc == QCS_NO_ERROR) {
            rc = reg->world > 1 ? scan_walk(reg, 0, carry, 1e300, limit, &found, &index, &total, d_bnd, &bad)
                                : parallel_scan(reg, 0, carry, 1e300, limit, &found, &index, &total, d_bnd, &bad);
            if (rc == QCS_NO_ERROR && !bad &&
                (cudaMemcpyAsync(bnd.data(), d_bnd, bnd.size() * sizeof(double), cudaMemcpyDeviceToHost, reg->stream) != cudaSuccess ||
                 cudaStreamSynchronize(reg->stream) != cudaSuccess))
                rc = QCS_UNKNOWN_ERROR;
            mine = bad ? -1.0 : total;                   // a negative sum tells every rank to give up
        }
        if (reg->world > 1) {
            // an error on any shard is seen by all of them in the same all-gather
            std::vector<double> all((size_t) reg->world * 2);
            const double pack[2] = {rc == QCS_NO_ERROR ? mine : -1.0, (double) rc};
            const int rc2 = qcs_dist_allgather_doubles(reg, pack, 2, all.data());
            if (rc == QCS_NO_ERROR) rc = rc2;
            for (int s = 0; s < reg->world && rc == QCS_NO_ERROR; s++)
                if (all[(size_t) s * 2 + 1] != 0.0) rc = (int) all[(size_t) s * 2 + 1];
            mine = all[(size_t) turn * 2];
        }
        if (rc != QCS_NO_ERROR) break;
        rank_end[(size_t) turn] = mine;
        carry = mine;
        if (mine < 0.0) bad = 1;
        if (bad) break;
    }
    cudaFree(d_bnd);
    if (rc != QCS_NO_ERROR) return rc;
    if (bad) return QCS_NO_ERROR;                       // caller takes the one-scan-per-variate path
    const uint64_t super_len = (uint64_t) kSuper << kChunkBits;
    for (uint64_t k = 0; k < n_shots; k++) {
        // the shard, then the super-chunk, whose closing running sum first reaches r[k]
        int owner = 0;
        while (owner < reg->world && !(rank_end[(size_t) owner] >= r[k])) owner++;
        double answer = (double) (reg->N - 1);          // indices < 2^53 are exact in a double
        int failed = 0;
        if (owner == reg->rank) {
            uint64_t lo = 0, hi = n_super;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) / 2;
                if (bnd[(size_t) mid + 1] >= r[k]) hi = mid; else lo = mid + 1;
            }
            if (lo == n_super) failed = 1;               // cannot happen: rank_end == bnd[n_super]
            else if (bnd[(size_t) lo] >= r[k]) {
                // the running sum reaches r before the first addend (only lo == 0 with r <= the
                // carried-in sum): the scan stops at its first index (qc_shor.c:289)
                answer = (double) ((uint64_t) reg->rank * reg->N_local + lo * super_len);
            } else {
                const uint64_t first = lo * super_len;
                const uint64_t len = limit - first < super_len ? limit - first : super_len;
                double cum = 0.0;
                rc = parallel_scan(reg, first, bnd[(size_t) lo], r[k], len, &found, &index, &cum, nullptr, &bad);
                if (rc != QCS_NO_ERROR || bad || !found) failed = 1;
                else answer = (double) ((uint64_t) reg->rank * reg->N_local + index);
            }
        }
        if (reg->world > 1 && owner < reg->world) {
            std::vector<double> all((size_t) reg->world);
            QCS_TRY(qcs_dist_allgather_double(reg, failed ? -1.0 : answer, all.data()));
            answer = all[(size_t) owner];
            failed = answer < 0.0;
        }
        if (failed) return rc;                           // never observed: redo everything the slow way
        indices[k] = (unsigned long long) answer;
    }
    *handled = true;
    return QCS_NO_ERROR;
}
