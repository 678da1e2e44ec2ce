// gates_exact.cu -- one kernel per reference gate, applied in place.
//
// These kernels perform exactly the floating-point operations that the
// reference's operate_matrix (qc_shor.c:396-413) performs for the matrix the
// corresponding *_gate builder would have produced, in the same order, with
// no fused multiply-add (the __d*_rn intrinsics are never contracted).  The
// amplitudes they produce are therefore value-identical to the reference's
// (the only representational difference is the sign of an exact zero, which
// the reference normalises to +0 by accumulating into a zeroed vector).
//
// All of them are HBM-bound streaming kernels: 128-bit (double2) accesses,
// a warp touching 512 contiguous bytes per access whenever the target qubit
// is >= 5, 64-bit indices, several independent loads in flight per thread.
#include "qcs_internal.h"

#include <math.h>

namespace {

constexpr int kThreads = 256;

// one COO entry the way qc_shor.c:409,412 applies it: m * cur as two products
// and a difference / a sum
__device__ __forceinline__ double2 ref_term(double m_re, double m_im, double2 c)
{
    double2 t;
    t.x = __dsub_rn(__dmul_rn(m_re, c.x), __dmul_rn(m_im, c.y));
    t.y = __dadd_rn(__dmul_rn(m_re, c.y), __dmul_rn(m_im, c.x));
    return t;
}
__device__ __forceinline__ double2 ref_acc(double2 acc, double2 t)
{
    return make_double2(__dadd_rn(acc.x, t.x), __dadd_rn(acc.y, t.y));
}

// ---------------------------------------------------------------------------
// fill kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_basis_state(double2 *__restrict__ amp, uint64_t n_amps, uint64_t one_at, int has_one)
{
    // reset_register (qc_shor.c:318-324) and the collapse of measure_state
    // (qc_shor.c:302-303): all zero except one amplitude = 1 + 0i
    const uint64_t stride = (uint64_t) gridDim.x * kThreads;
    for (uint64_t i = (uint64_t) blockIdx.x * kThreads + threadIdx.x; i < n_amps; i += stride) {
        double2 v = make_double2(0.0, 0.0);
        if (has_one && i == one_at) v.x = 1.0;
        amp[i] = v;
    }
}

__device__ __forceinline__ double synthetic_u(uint64_t seed, uint64_t k)
{
    uint64_t z = seed + k + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double) (z >> 11) * 0x1.0p-53 - 0.5;
}

__global__ void __launch_bounds__(kThreads)
k_fill_synthetic(double2 *__restrict__ amp, uint64_t n_amps, uint64_t first_global, uint64_t seed)
{
    const uint64_t stride = (uint64_t) gridDim.x * kThreads;
    for (uint64_t i = (uint64_t) blockIdx.x * kThreads + threadIdx.x; i < n_amps; i += stride) {
        const uint64_t g = first_global + i;
        amp[i] = make_double2(synthetic_u(seed, 2 * g), synthetic_u(seed, 2 * g + 1));
    }
}

__global__ void __launch_bounds__(kThreads)
k_scale(double2 *__restrict__ amp, uint64_t n_amps, double s)
{
    const uint64_t stride = (uint64_t) gridDim.x * kThreads;
    for (uint64_t i = (uint64_t) blockIdx.x * kThreads + threadIdx.x; i < n_amps; i += stride) {
        double2 v = amp[i];
        v.x = __dmul_rn(v.x, s);
        v.y = __dmul_rn(v.y, s);
        amp[i] = v;
    }
}

// ---------------------------------------------------------------------------
// Hadamard, qc_shor.c:442-484.  Row i of the built matrix holds two entries in
// ascending column order: (i & ~bit) with H[b][0] = h and (i | bit) with
// H[b][1] = +-h, imaginary parts 0.0 (qc_shor.c:453,476), accumulated into a
// zeroed row (qc_shor.c:393).
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(kThreads)
k_hadamard_exact(double2 *__restrict__ amp, uint64_t first_pair, uint64_t n_pairs, unsigned q)
{
    const double h = 0.70710678118654752440;   // M_SQRT1_2
    const uint64_t bit = 1ull << q;
    const uint64_t first = (uint64_t) blockIdx.x * (kThreads * U) + threadIdx.x;
    uint64_t i0[U];
    double2 a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t p = first + (uint64_t) u * kThreads;
        i0[u] = qcs_insert_zero_bit(first_pair + p, q);
        if (p < n_pairs) {
            a0[u] = amp[i0[u]];
            a1[u] = amp[i0[u] | bit];
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t p = first + (uint64_t) u * kThreads;
        if (p < n_pairs) {
            const double2 zero = make_double2(0.0, 0.0);
            const double2 t0 = ref_term(h, 0.0, a0[u]);
            const double2 r0 = ref_acc(ref_acc(zero, t0), ref_term(h, 0.0, a1[u]));
            const double2 r1 = ref_acc(ref_acc(zero, t0), ref_term(-h, 0.0, a1[u]));
            amp[i0[u]] = r0;
            amp[i0[u] | bit] = r1;
        }
    }
}

// ---------------------------------------------------------------------------
// Controlled phase, qc_shor.c:513-565.  The matrix is diagonal: 1 except
// e^{i theta} on rows whose control and target bits are both 1; the explicit
// off-diagonal zeros (qc_shor.c:549-559) only add signed zeros.  Only the
// |11> quarter is read and written: row = 0 + (c*re - s*im, c*im + s*re).
// NB = number of index bits that must be 1 (2 in the single-GPU case; 1 or 0
// when the other qubit(s) are global and this rank's bit is set).
// ---------------------------------------------------------------------------
template <int NB, int U>
__global__ void __launch_bounds__(kThreads)
k_phase_masked(double2 *__restrict__ amp, uint64_t n_sel, unsigned b_lo, unsigned b_hi,
               double c, double s)
{
    const uint64_t first = (uint64_t) blockIdx.x * (kThreads * U) + threadIdx.x;
    uint64_t idx[U];
    double2 a[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t p = first + (uint64_t) u * kThreads;
        uint64_t i = p;
        if (NB >= 1) i = qcs_insert_zero_bit(i, b_lo) | (1ull << b_lo);
        if (NB >= 2) i = qcs_insert_zero_bit(i, b_hi) | (1ull << b_hi);
        idx[u] = i;
        if (p < n_sel) a[u] = amp[i];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t p = first + (uint64_t) u * kThreads;
        if (p < n_sel) amp[idx[u]] = ref_acc(make_double2(0.0, 0.0), ref_term(c, s, a[u]));
    }
}

// ---------------------------------------------------------------------------
// Controlled a^x mod C, qc_shor.c:595-660.  The matrix has one entry
// (row j(k), column k) = 1 + 0i per column k, emitted for ascending k:
//   j = k                                   if control bit of k is 0 or f >= C
//   j = (k with its low M bits := A f % C)  otherwise,  f = k mod 2^M.
// The map is local to blocks of 2^M consecutive amplitudes.  A chunk of
// 2^T >= 2^M amplitudes is staged in shared memory and every output row
// gathers its sources in ascending source order:
//   rows f' with g = gcd(A, C) not dividing f' receive nothing (-> 0),
//   otherwise the g sources f0 + t * (C / g), t = 0..g-1 (bit-exact integer
//   arithmetic; for a bijective map g = 1 and this is a pure move).
// ---------------------------------------------------------------------------
struct amodc_params {
    unsigned T;            // log2 chunk size
    unsigned M;
    unsigned C, A;
    unsigned g, step, inv; // gcd(A,C), C/g, (A/g)^-1 mod C/g
    int ctrl_in_chunk;     // control bit position if < T, else -1 (chunk pre-selected)
    int ctrl_chunk_bit;    // control bit position - T if >= T, else -1 (no selection)
    int generic;           // 1: scan-all-sources path (C > 2^M or control below M)
};

__device__ __forceinline__ uint64_t amodc_row_of(uint64_t k, const amodc_params &P, uint64_t maskM)
{
    // the reference's forward map for a column k whose control bit is 1
    unsigned f = (unsigned) (k & maskM);
    if (f >= P.C) return k;
    f = (P.A * f) % P.C;
    return (k & ~maskM) | (uint64_t) (f & (unsigned) maskM);
}

__global__ void __launch_bounds__(1024)
k_amodc(double2 *__restrict__ amp, uint64_t n_chunks_sel, amodc_params P)
{
    extern __shared__ double2 chunk[];
    const uint64_t maskM = P.M ? ((1ull << P.M) - 1ull) : 0ull;
    const unsigned chunk_len = 1u << P.T;
    for (uint64_t cs = blockIdx.x; cs < n_chunks_sel; cs += gridDim.x) {
        uint64_t ci = cs;
        if (P.ctrl_chunk_bit >= 0) ci = qcs_insert_zero_bit(cs, (unsigned) P.ctrl_chunk_bit) | (1ull << P.ctrl_chunk_bit);
        double2 *g_chunk = amp + (ci << P.T);
        for (unsigned e = threadIdx.x; e < chunk_len; e += blockDim.x) chunk[e] = g_chunk[e];
        __syncthreads();
        for (unsigned j = threadIdx.x; j < chunk_len; j += blockDim.x) {
            const unsigned fp = (unsigned) (j & maskM);
            const unsigned hi = j & ~(unsigned) maskM;
            double2 acc = make_double2(0.0, 0.0);
            if (P.generic) {
                // every source of the block in ascending order
                const unsigned blk = 1u << P.M;
                for (unsigned f = 0; f < blk; f++) {
                    const unsigned k = hi | f;
                    const bool on = P.ctrl_in_chunk < 0 || ((k >> P.ctrl_in_chunk) & 1u);
                    const uint64_t row = on ? amodc_row_of(k, P, maskM) : (uint64_t) k;
                    if (row == (uint64_t) j) acc = ref_acc(acc, ref_term(1.0, 0.0, chunk[k]));
                }
                g_chunk[j] = acc;
                continue;
            }
            const bool on = P.ctrl_in_chunk < 0 || ((j >> P.ctrl_in_chunk) & 1u);
            if (!on || fp >= P.C) continue;                 // identity row
            if (fp % P.g == 0) {
                const unsigned f0 = (unsigned) (((uint64_t) (fp / P.g) * P.inv) % P.step);
                for (unsigned t = 0; t < P.g; t++)
                    acc = ref_acc(acc, ref_term(1.0, 0.0, chunk[hi | (f0 + t * P.step)]));
            }
            g_chunk[j] = acc;
        }
        __syncthreads();
    }
}

unsigned host_gcd(unsigned a, unsigned b)
{
    while (b) { unsigned t = a % b; a = b; b = t; }
    return a;
}

// modular inverse of a mod m (gcd(a, m) == 1), m >= 1
unsigned host_modinv(unsigned a, unsigned m)
{
    if (m == 1) return 0;
    long long t = 0, nt = 1, r = m, nr = a % m;
    while (nr != 0) {
        long long q = r / nr;
        long long tmp = t - q * nt; t = nt; nt = tmp;
        tmp = r - q * nr; r = nr; nr = tmp;
    }
    if (t < 0) t += m;
    return (unsigned) t;
}

inline unsigned grid_for(uint64_t items, unsigned per_block)
{
    uint64_t b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return (unsigned) b;
}

inline unsigned streaming_grid(const qcs_register *reg, uint64_t items)
{
    // grid-stride fill kernels: a multiple of the SM count, 8 resident CTAs each
    uint64_t want = (items + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t) reg->sm_count * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (unsigned) want;
}

}  // namespace

int qcs_k_reset(qcs_register *reg)
{
    // index 1 lives on rank 0 (qc_shor.c:323)
    qcs_launch_begin(reg, QCS_K_FILL, 16.0 * (double) reg->N_local);
    k_basis_state<<<streaming_grid(reg, reg->N_local), kThreads, 0, reg->stream>>>(
        reg->amp, reg->N_local, 1ull, reg->rank == 0 ? 1 : 0);
    return qcs_launch_end(reg, QCS_K_FILL, "k_basis_state");
}

int qcs_k_collapse(qcs_register *reg, uint64_t local_index, bool owner)
{
    qcs_launch_begin(reg, QCS_K_FILL, 16.0 * (double) reg->N_local);
    k_basis_state<<<streaming_grid(reg, reg->N_local), kThreads, 0, reg->stream>>>(
        reg->amp, reg->N_local, local_index, owner ? 1 : 0);
    return qcs_launch_end(reg, QCS_K_FILL, "k_basis_state");
}

int qcs_k_fill_synthetic(qcs_register *reg, uint64_t seed)
{
    qcs_launch_begin(reg, QCS_K_FILL, 16.0 * (double) reg->N_local);
    k_fill_synthetic<<<streaming_grid(reg, reg->N_local), kThreads, 0, reg->stream>>>(
        reg->amp, reg->N_local, (uint64_t) reg->rank * reg->N_local, seed);
    return qcs_launch_end(reg, QCS_K_FILL, "k_fill_synthetic");
}

int qcs_k_scale(qcs_register *reg, double s)
{
    qcs_launch_begin(reg, QCS_K_SCALE, 32.0 * (double) reg->N_local);
    k_scale<<<streaming_grid(reg, reg->N_local), kThreads, 0, reg->stream>>>(reg->amp, reg->N_local, s);
    return qcs_launch_end(reg, QCS_K_SCALE, "k_scale");
}

int qcs_k_hadamard_local(qcs_register *reg, unsigned q)
{
    constexpr int U = 4;
    const uint64_t n_pairs = reg->N_local >> 1;
    qcs_launch_begin(reg, QCS_K_HADAMARD, 32.0 * (double) reg->N_local);
    k_hadamard_exact<U><<<grid_for(n_pairs, kThreads * U), kThreads, 0, reg->stream>>>(reg->amp, 0, n_pairs, q);
    return qcs_launch_end(reg, QCS_K_HADAMARD, "k_hadamard_exact");
}

// Hadamard on a global qubit of a register with peer memory: the same kernel (the same
// reference-order arithmetic) on the stitched array, this rank taking its share of the pairs of
// the whole register; half of each pair it touches lives on the partner rank (NVLink).
int qcs_k_hadamard_peer(qcs_register *reg, unsigned q)
{
    constexpr int U = 4;
    if (!reg->peer || q < reg->n_local || q >= reg->n) return QCS_BAD_ARGUMENTS;
    const uint64_t share = (reg->N >> 1) / (uint64_t) reg->world;
    QCS_TRY(qcs_dist_stream_barrier(reg));
    qcs_launch_begin(reg, QCS_K_HADAMARD, 32.0 * (double) reg->N_local);
    k_hadamard_exact<U><<<grid_for(share, kThreads * U), kThreads, 0, reg->stream>>>(
        reg->amp_all, (uint64_t) reg->rank * share, share, q);
    QCS_TRY(qcs_launch_end(reg, QCS_K_HADAMARD, "k_hadamard_exact"));
    return qcs_dist_stream_barrier(reg);
}

int qcs_k_phase_masked(qcs_register *reg, int nbits, unsigned b0, unsigned b1, double c, double s)
{
    constexpr int U = 4;
    const uint64_t n_sel = reg->N_local >> nbits;
    if (n_sel == 0) return QCS_NO_ERROR;
    unsigned lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
    const unsigned grid = grid_for(n_sel, kThreads * U);
    qcs_launch_begin(reg, QCS_K_CPHASE, 32.0 * (double) n_sel);
    if (nbits == 2)
        k_phase_masked<2, U><<<grid, kThreads, 0, reg->stream>>>(reg->amp, n_sel, lo, hi, c, s);
    else if (nbits == 1)
        k_phase_masked<1, U><<<grid, kThreads, 0, reg->stream>>>(reg->amp, n_sel, b0, 0, c, s);
    else
        k_phase_masked<0, U><<<grid, kThreads, 0, reg->stream>>>(reg->amp, n_sel, 0, 0, c, s);
    return qcs_launch_end(reg, QCS_K_CPHASE, "k_phase_masked");
}

int qcs_k_amodc(qcs_register *reg, unsigned C, unsigned A, int ctrl_local, bool ctrl_off)
{
    if (ctrl_off) return QCS_NO_ERROR;   // global control bit is 0 on this rank: identity
    const unsigned M = (unsigned) reg->M_size;
    amodc_params P;
    P.M = M;
    P.C = C;
    P.A = A;
    P.g = host_gcd(A, C);                 // gcd(0, C) = C
    P.step = C / P.g;
    P.inv = host_modinv((A / P.g) % (P.step ? P.step : 1), P.step);
    unsigned T = M < 10 ? 10 : M;
    if (T > reg->n_local) T = reg->n_local;
    if (T < M) return QCS_BAD_ARGUMENTS;
    const size_t smem = (size_t) 16 << T;
    if (smem > reg->smem_optin) {
        fprintf(stderr, "qcs: c_amodc_gate needs a 2^%u-amplitude block in shared memory (M too large)\n", M);
        return QCS_BAD_ARGUMENTS;
    }
    P.T = T;
    P.generic = (M < 32 && ((uint64_t) C > (1ull << M))) || (ctrl_local >= 0 && (unsigned) ctrl_local < M);
    P.ctrl_in_chunk = -1;
    P.ctrl_chunk_bit = -1;
    uint64_t n_chunks = reg->N_local >> T;
    if (ctrl_local >= 0) {
        if ((unsigned) ctrl_local < T) P.ctrl_in_chunk = ctrl_local;
        else { P.ctrl_chunk_bit = ctrl_local - (int) T; n_chunks >>= 1; }
    }
    if (n_chunks == 0) return QCS_NO_ERROR;
    QCS_CUDA(cudaFuncSetAttribute(k_amodc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    unsigned threads = (1u << T) < 1024u ? (1u << T) : 1024u;
    if (threads < 32) threads = 32;
    uint64_t grid = n_chunks;
    const uint64_t cap = (uint64_t) reg->sm_count * 8;
    if (grid > cap) grid = cap;
    // algorithmic bytes: read+write of the rows that move (control on, f < C)
    double frac = (double) (C < (1u << M) ? C : (1u << M)) / (double) (1u << M);
    qcs_launch_begin(reg, QCS_K_AMODC, 32.0 * (double) (n_chunks << T) * frac *
                                           (P.ctrl_in_chunk >= 0 ? 0.5 : 1.0));
    k_amodc<<<(unsigned) grid, threads, smem, reg->stream>>>(reg->amp, n_chunks, P);
    return qcs_launch_end(reg, QCS_K_AMODC, "k_amodc");
}
