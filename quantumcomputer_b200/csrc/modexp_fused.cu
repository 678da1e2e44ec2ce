// modexp_fused.cu -- the first half of quantum_computation (qc_shor.c:720-731)
// as fused sweeps: Hadamards on the L register (Walsh-Hadamard tile sweeps, see
// qft_common.cuh) and then ALL L controlled a^(2^k) mod C gates in one
// block-local pass.
//
// Gate k (control qubit first + k, multiplier A_k = atox_k % C) maps, inside
// every block of 2^M consecutive amplitudes whose control bit is 1,
//     f -> A_k f mod C   (f < C;  rows f >= C stay put)          qc_shor.c:619-652
// The blocks never mix, so a chunk of 2^T >= 2^M amplitudes is staged in
// shared memory once and all gates are applied to it there:
//   * every A_k coprime to C (the normal case): the L maps commute and compose
//     into one bijection f -> B(x) f mod C with B(x) = prod_{k: x_k = 1} A_k,
//     applied as a single gather with B(x)^-1 -- pure moves, bit-exact;
//   * otherwise (gcd(A_k, C) != 1, A_k = 0 from INT_POW overflow): the gates
//     are applied one after another between two shared-memory buffers, each
//     row summing its sources in ascending order exactly like
//     gates_exact.cu / operate_matrix do.
// Either way the result is value-identical to applying the L gates one by one.
#include "qcs_internal.h"

namespace {

constexpr int kMaxGates = 62;

struct modexp_params {
    unsigned T, M, C, L;
    unsigned first;            // control qubit of gate 0
    unsigned n_local;
    unsigned rank;             // supplies control bits >= n_local
    int bijective;
    unsigned A[kMaxGates];
    unsigned inv[kMaxGates];   // bijective: A^-1 mod C;  general: (A/g)^-1 mod C/g
    unsigned g[kMaxGates];     // gcd(A, C)
};

__device__ __forceinline__ double2 ref_term1(double2 c)
{
    // the entry 1 + 0i applied as qc_shor.c:409,412 does
    return make_double2(__dsub_rn(__dmul_rn(1.0, c.x), __dmul_rn(0.0, c.y)),
                        __dadd_rn(__dmul_rn(1.0, c.y), __dmul_rn(0.0, c.x)));
}

// control bit of gate k for an element whose shard-local index is `local`
__device__ __forceinline__ unsigned control_bit(const modexp_params &P, unsigned k, uint64_t local)
{
    const unsigned pos = P.first + k;
    if (pos >= P.n_local) return (P.rank >> (pos - P.n_local)) & 1u;
    return (unsigned) (local >> pos) & 1u;
}

__global__ void __launch_bounds__(1024)
k_modexp_sweep(double2 *__restrict__ amp, uint64_t n_chunks, const modexp_params P)
{
    extern __shared__ double2 smem[];
    double2 *buf0 = smem;
    const unsigned chunk_len = 1u << P.T;
    double2 *buf1 = smem + chunk_len;                       // general path only
    unsigned *mult = (unsigned *) (P.bijective ? (smem + chunk_len) : (smem + 2 * chunk_len));
    const unsigned maskM = (1u << P.M) - 1u;
    const unsigned n_blocks = 1u << (P.T - P.M);

    for (uint64_t ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const uint64_t chunk_base = ci << P.T;
        double2 *g_chunk = amp + chunk_base;
        for (unsigned e = threadIdx.x; e < chunk_len; e += blockDim.x) buf0[e] = g_chunk[e];
        if (P.bijective) {
            // per-block inverse multiplier B(x)^-1 = prod A_k^-1 over the set control bits
            for (unsigned b = threadIdx.x; b < n_blocks; b += blockDim.x) {
                const uint64_t local = chunk_base | ((uint64_t) b << P.M);
                unsigned long long m = 1 % P.C;
                for (unsigned k = 0; k < P.L; k++)
                    if (control_bit(P, k, local)) m = (m * P.inv[k]) % P.C;
                mult[b] = (unsigned) m;
            }
            __syncthreads();
            for (unsigned j = threadIdx.x; j < chunk_len; j += blockDim.x) {
                const unsigned fp = j & maskM;
                const unsigned mb = mult[j >> P.M];
                if (fp >= P.C || mb == 1u) continue;          // identity row
                const unsigned f = (unsigned) (((unsigned long long) mb * fp) % P.C);
                g_chunk[j] = buf0[(j & ~maskM) | f];
            }
            __syncthreads();
            continue;
        }
        __syncthreads();
        double2 *src = buf0, *dst = buf1;
        bool moved = false;
        for (unsigned k = 0; k < P.L; k++) {
            const unsigned pos = P.first + k;
            if (pos >= P.T && !control_bit(P, k, chunk_base)) continue;     // off for the whole chunk
            const unsigned gk = P.g[k], stepk = P.C / gk;
            for (unsigned j = threadIdx.x; j < chunk_len; j += blockDim.x) {
                const unsigned fp = j & maskM;
                const bool on = pos >= P.T || ((j >> pos) & 1u);
                double2 v;
                if (!on || fp >= P.C) {
                    v = src[j];
                } else {
                    v = make_double2(0.0, 0.0);
                    if (fp % gk == 0) {
                        const unsigned f0 = (unsigned) (((unsigned long long) (fp / gk) * P.inv[k]) % stepk);
                        for (unsigned t = 0; t < gk; t++) {
                            const double2 term = ref_term1(src[(j & ~maskM) | (f0 + t * stepk)]);
                            v.x = __dadd_rn(v.x, term.x);
                            v.y = __dadd_rn(v.y, term.y);
                        }
                    }
                }
                dst[j] = v;
            }
            __syncthreads();
            double2 *tmp = src; src = dst; dst = tmp;
            moved = true;
        }
        if (moved)
            for (unsigned e = threadIdx.x; e < chunk_len; e += blockDim.x) g_chunk[e] = src[e];
        __syncthreads();
    }
}

// quantum_computation's first two loops (qc_shor.c:720-731) applied to the reset state |0...01>, in closed
// form.  H on every qubit of the L register turns |x = 0, f = 1> into 2^(-L/2) sum_x |x, 1>; gate k then moves,
// in every block x whose bit k is set, the block's single non-zero amplitude from f to (A_k f) % C when
// f < C (32-bit unsigned product, row index masked to the M register: qc_shor.c:619-652) and leaves it
// otherwise.  A single source per block never collides, so this holds for non-bijective multipliers and for
// A_k = 0 (INT_POW overflow) as well.  One write pass; the value is the Walsh-Hadamard sweep's own scale.
struct closed_params {
    unsigned M, L, C, n_local, rank;
    unsigned A[kMaxGates];
    double value;
};

__global__ void __launch_bounds__(256)
k_shor_state_from_reset(double2 *__restrict__ amp, uint64_t n_windows, int window_bits, const closed_params P)
{
    // a WINDOW = 2^window_bits <= 2^M consecutive amplitudes of one block (one x, one f).  Windows of >= 32
    // amplitudes are written by a whole warp, lane t taking elements t, t + 32, ... (coalesced 512-byte
    // stores; f is computed by every lane); smaller windows (M < 5) by one thread each.
    const unsigned maskM = (1u << P.M) - 1u;
    const bool by_warp = window_bits >= 5;
    const uint64_t unit = by_warp ? ((uint64_t) blockIdx.x * 256 + threadIdx.x) >> 5 : (uint64_t) blockIdx.x * 256 + threadIdx.x;
    const uint64_t stride = by_warp ? ((uint64_t) gridDim.x * 256) >> 5 : (uint64_t) gridDim.x * 256;
    const unsigned lane = threadIdx.x & 31u;
    for (uint64_t w = unit; w < n_windows; w += stride) {
        const uint64_t first = w << window_bits;                  // shard-local index of the window
        uint64_t x = first >> P.M;
        if (P.n_local < P.M + P.L) x |= (uint64_t) P.rank << (P.n_local - P.M);   // global qubits: the rank's bits
        unsigned f = 1u;
        for (unsigned k = 0; k < P.L; k++)
            if (((x >> k) & 1ull) && f < P.C) f = ((P.A[k] * f) % P.C) & maskM;
        const unsigned e0 = (unsigned) first & maskM;
        if (by_warp) {
            for (unsigned e = lane; e < (1u << window_bits); e += 32u)
                amp[first + e] = make_double2(e0 + e == f ? P.value : 0.0, 0.0);
        } else {
            for (unsigned e = 0; e < (1u << window_bits); e++)
                amp[first + e] = make_double2(e0 + e == f ? P.value : 0.0, 0.0);
        }
    }
}

// f(x) for every block x of the L register (the same gate-by-gate rule), for the generating sweep
__global__ void __launch_bounds__(256)
k_shor_f_table(unsigned *__restrict__ table, uint64_t n_blocks, const closed_params P)
{
    const unsigned maskM = (1u << P.M) - 1u;
    for (uint64_t x = (uint64_t) blockIdx.x * 256 + threadIdx.x; x < n_blocks; x += (uint64_t) gridDim.x * 256) {
        unsigned f = 1u;
        for (unsigned k = 0; k < P.L; k++)
            if (((x >> k) & 1ull) && f < P.C) f = ((P.A[k] * f) % P.C) & maskM;
        table[x] = f;
    }
}

unsigned h_gcd(unsigned a, unsigned b)
{
    while (b) { const unsigned t = a % b; a = b; b = t; }
    return a;
}

unsigned h_modinv(unsigned a, unsigned m)
{
    if (m == 1) return 0;
    long long t = 0, nt = 1, r = m, nr = a % m;
    while (nr != 0) {
        const long long q = r / nr;
        long long tmp = t - q * nt; t = nt; nt = tmp;
        tmp = r - q * nr; r = nr; nr = tmp;
    }
    if (t < 0) t += m;
    return (unsigned) t;
}

}  // namespace

int qcs_fused_modexp(qcs_register *reg, unsigned C, const unsigned *A_per_gate, unsigned n_gates)
{
    const unsigned first = reg->n - n_gates;                  // qc_shor.c:720,728
    const unsigned M = (unsigned) reg->M_size;
    // Hadamards: local qubits by Walsh-Hadamard sweeps; global ones inside the sweep over the
    // stitched array when the register has peer memory, else by exchange
    const unsigned h_hi = reg->n < reg->n_local ? reg->n : reg->n_local;
    const bool sharded_sweeps = reg->world > 1 && qcs_sharded_sweeps_supported(reg, first, reg->n);
    if (sharded_sweeps) {
        QCS_TRY(qcs_fused_sweeps_sharded(reg, first, reg->n, true, true));
    } else if (first < h_hi) {
        QCS_TRY(qcs_fused_hadamards(reg, first, h_hi));
    }
    if (reg->world > 1 && !sharded_sweeps) {
        if (first <= reg->n_local && reg->n_local >= 2u * (unsigned) reg->p_global)
            QCS_TRY(qcs_dist_top_stages(reg, 0, true, true));       // H on every global qubit at once
        else
            for (unsigned q = h_hi > first ? h_hi : first; q < reg->n; q++) QCS_TRY(qcs_dist_hadamard_global(reg, q));
    }

    if (M == 0) return QCS_NO_ERROR;                          // f = f' = 0: every gate is the identity
    const bool weird = C > 65536u || (M < 32 && (uint64_t) C > (1ull << M)) || first < M || M > reg->n_local ||
                       n_gates > (unsigned) kMaxGates;
    unsigned T = M < 10 ? 10 : M;
    if (T > reg->n_local) T = reg->n_local;
    modexp_params P;
    P.bijective = 1;
    for (unsigned k = 0; k < n_gates && !weird; k++) {
        P.A[k] = A_per_gate[k] % C;
        P.g[k] = h_gcd(P.A[k], C);
        if (P.g[k] != 1) P.bijective = 0;
    }
    const size_t smem = weird ? 0 : (((size_t) 16 << T) * (P.bijective ? 1 : 2) + 4 * ((size_t) 1 << (T - M)) + 16);
    if (weird || smem > reg->smem_optin) {
        // shapes the block-local sweep does not cover: the per-gate kernels handle them
        for (unsigned k = 0; k < n_gates; k++) {
            const unsigned c = first + k;
            if (c >= reg->n_local) {
                const bool on = ((unsigned) reg->rank >> (c - reg->n_local)) & 1u;
                QCS_TRY(qcs_k_amodc(reg, C, A_per_gate[k] % C, -1, !on));
            } else {
                QCS_TRY(qcs_k_amodc(reg, C, A_per_gate[k] % C, (int) c, false));
            }
        }
        return QCS_NO_ERROR;
    }
    for (unsigned k = 0; k < n_gates; k++) {
        const unsigned stepk = C / P.g[k];
        P.inv[k] = P.bijective ? h_modinv(P.A[k], C) : h_modinv((P.A[k] / P.g[k]) % (stepk ? stepk : 1), stepk);
    }
    P.T = T;
    P.M = M;
    P.C = C;
    P.L = n_gates;
    P.first = first;
    P.n_local = reg->n_local;
    P.rank = (unsigned) reg->rank;
    const uint64_t n_chunks = reg->N_local >> T;
    QCS_CUDA(cudaFuncSetAttribute(k_modexp_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    unsigned threads = (1u << T) < 1024u ? (1u << T) : 1024u;
    if (threads < 32) threads = 32;
    int per_sm = 1;
    QCS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_modexp_sweep, (int) threads, smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t) reg->sm_count * (uint64_t) per_sm;
    if (grid > n_chunks) grid = n_chunks;
    const double frac = (double) C / (double) (1u << M);
    qcs_launch_begin(reg, QCS_K_MODEXP_SWEEP, 32.0 * (double) reg->N_local * (frac < 1.0 ? frac : 1.0));
    k_modexp_sweep<<<(unsigned) grid, threads, smem, reg->stream>>>(reg->amp, n_chunks, P);
    return qcs_launch_end(reg, QCS_K_MODEXP_SWEEP, "k_modexp_sweep");
}


int qcs_shor_state_from_reset(qcs_register *reg, unsigned C, const unsigned *A_per_gate, unsigned n_gates, bool *done)
{
    *done = false;
    const unsigned M = (unsigned) reg->M_size, L = n_gates;
    // M >= 1 (index 1 must belong to the M register, else the Hadamards see |x = 1>), controls = the whole L
    // register, products below 2^32 as in c_amodc_gate, the M register inside the shard
    if (M < 1 || M > 30 || L != (unsigned) reg->L_size || L > (unsigned) kMaxGates || C == 0 || C > 65536u || M > reg->n_local ||
        reg->n != M + L)
        return QCS_NO_ERROR;
    closed_params P;
    P.M = M;
    P.L = L;
    P.C = C;
    P.n_local = reg->n_local;
    P.rank = (unsigned) reg->rank;
    for (unsigned k = 0; k < L; k++) P.A[k] = A_per_gate[k] % C;
    // the scale the fused Walsh-Hadamard sweeps apply (qft_common.cuh): an exact power of two for even L
    P.value = L % 2 == 0 ? ldexp(1.0, -(int) L / 2) : ldexp(0.70710678118654752440, -((int) L - 1) / 2);
    const int window_bits = M < 8 ? (int) M : 8;              // 256 amplitudes = 4 KiB per warp iteration
    const uint64_t n_windows = reg->N_local >> window_bits;
    const uint64_t units_per_cta = window_bits >= 5 ? 8 : 256;
    uint64_t grid = (n_windows + units_per_cta - 1) / units_per_cta;
    const uint64_t cap = (uint64_t) reg->sm_count * 16;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    qcs_launch_begin(reg, QCS_K_FILL, 16.0 * (double) reg->N_local);
    k_shor_state_from_reset<<<(unsigned) grid, 256, 0, reg->stream>>>(reg->amp, n_windows, window_bits, P);
    QCS_TRY(qcs_launch_end(reg, QCS_K_FILL, "k_shor_state_from_reset"));
    *done = true;
    return QCS_NO_ERROR;
}


int qcs_shor_state_generated(qcs_register *reg, unsigned C, const unsigned *A_per_gate, unsigned n_gates, bool *armed)
{
    *armed = false;
    const unsigned M = (unsigned) reg->M_size, L = n_gates;
    if (reg->world != 1 || !reg->opt_gen_sweep || M < 1 || M > 30 || L != (unsigned) reg->L_size || L > (unsigned) kMaxGates || L > 26 ||
        C == 0 || C > 65536u || reg->n != M + L || reg->n != reg->n_local || !qcs_fused_gen_supported(reg, M, reg->n))
        return QCS_NO_ERROR;
    const uint64_t n_blocks = 1ull << L;
    if (reg->gen_table_cap < n_blocks) {
        if (reg->d_gen_table) cudaFree(reg->d_gen_table);
        reg->d_gen_table = nullptr;
        reg->gen_table_cap = 0;
        QCS_CUDA(cudaMalloc((void **) &reg->d_gen_table, n_blocks * sizeof(unsigned)));
        reg->gen_table_cap = n_blocks;
    }
    closed_params P;
    P.M = M;
    P.L = L;
    P.C = C;
    P.n_local = reg->n_local;
    P.rank = 0;
    for (unsigned k = 0; k < L; k++) P.A[k] = A_per_gate[k] % C;
    P.value = L % 2 == 0 ? ldexp(1.0, -(int) L / 2) : ldexp(0.70710678118654752440, -((int) L - 1) / 2);
    uint64_t grid = (n_blocks + 255) / 256;
    const uint64_t cap = (uint64_t) reg->sm_count * 8;
    if (grid > cap) grid = cap;
    qcs_launch_begin(reg, QCS_K_FILL, 4.0 * (double) n_blocks);
    k_shor_f_table<<<(unsigned) grid, 256, 0, reg->stream>>>(reg->d_gen_table, n_blocks, P);
    QCS_TRY(qcs_launch_end(reg, QCS_K_FILL, "k_shor_f_table"));
    reg->gen.f = reg->d_gen_table;
    reg->gen.M = M;
    reg->gen.value = P.value;
    reg->gen.armed = 1;
    *armed = true;
    return QCS_NO_ERROR;
}
