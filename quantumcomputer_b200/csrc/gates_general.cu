// gates_general.cu -- arbitrary single-qubit and controlled single-qubit gates.
//
// The reference hard-codes two base matrices, HADAMARD_BASE_MATRIX and
// C_PHASE_SHIFT_BASE_MATRIX (qc_shor.c:210-225), and expands each into a
// 2^n x 2^n COO matrix per call.  This is the generalisation its report asks for
// ("future flexibility"): any 2x2 complex matrix U on a target
// qubit, optionally controlled by another qubit, applied with the same
// pair-stride access pattern as the Hadamard kernel of gates_exact.cu:
//   amp'[i0] = U00 amp[i0] + U01 amp[i1],  amp'[i1] = U10 amp[i0] + U11 amp[i1],
// i0 / i1 = the basis states that differ only in the target bit (and, for a
// controlled gate, only where the control bit is 1).  HBM-bound: 32 B per
// amplitude touched (all of them, or the control = 1 half).
//
// There is no reference-order arithmetic to mirror here, so products may be
// contracted to FMA.  A global target qubit runs on the stitched peer-memory
// array (every rank takes its share of the pairs) or, without peer memory,
// through a pairwise shard exchange.
#include "qcs_internal.h"

namespace {

constexpr int kThreads = 256;

struct mat2 {
    double2 u00, u01, u10, u11;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// HAS_CTRL: index space is the N/4 quads; the control bit is forced to 1
template <bool HAS_CTRL, int U>
__global__ void __launch_bounds__(kThreads)
k_gate_1q(double2 *__restrict__ amp, uint64_t first_item, uint64_t n_items, unsigned q, unsigned c, mat2 M)
{
    const uint64_t bit = 1ull << q;
    const uint64_t first = (uint64_t) blockIdx.x * (kThreads * U) + threadIdx.x;
    uint64_t i0[U];
    double2 a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t p = first + (uint64_t) u * kThreads;
        uint64_t i = first_item + p;
        if (HAS_CTRL) {
            const unsigned lo = q < c ? q : c, hi = q < c ? c : q;
            i = qcs_insert_zero_bit(qcs_insert_zero_bit(i, lo), hi) | (1ull << c);
        } else {
            i = qcs_insert_zero_bit(i, q);
        }
        i0[u] = i;
        if (p < n_items) {
            a0[u] = amp[i];
            a1[u] = amp[i | bit];
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t p = first + (uint64_t) u * kThreads;
        if (p < n_items) {
            amp[i0[u]] = cadd(cmul(M.u00, a0[u]), cmul(M.u01, a1[u]));
            amp[i0[u] | bit] = cadd(cmul(M.u10, a0[u]), cmul(M.u11, a1[u]));
        }
    }
}

mat2 load_matrix(const double *u)
{
    mat2 M;
    M.u00 = make_double2(u[0], u[1]);
    M.u01 = make_double2(u[2], u[3]);
    M.u10 = make_double2(u[4], u[5]);
    M.u11 = make_double2(u[6], u[7]);
    return M;
}

int launch_pairs(qcs_register *reg, double2 *base, uint64_t first_item, uint64_t n_items, unsigned q, int c, const mat2 &M)
{
    constexpr int U = 4;
    if (n_items == 0) return QCS_NO_ERROR;
    const uint64_t grid = (n_items + (uint64_t) kThreads * U - 1) / ((uint64_t) kThreads * U);
    qcs_launch_begin(reg, QCS_K_GATE_1Q, 64.0 * (double) n_items);
    if (c >= 0)
        k_gate_1q<true, U><<<(unsigned) grid, kThreads, 0, reg->stream>>>(base, first_item, n_items, q, (unsigned) c, M);
    else
        k_gate_1q<false, U><<<(unsigned) grid, kThreads, 0, reg->stream>>>(base, first_item, n_items, q, 0, M);
    return qcs_launch_end(reg, QCS_K_GATE_1Q, "k_gate_1q");
}

}  // namespace

// c < 0: no control
int qcs_k_gate_1q(qcs_register *reg, unsigned q, int c, const double *u_interleaved)
{
    if (!u_interleaved || q >= reg->n || (c >= 0 && ((unsigned) c >= reg->n || (unsigned) c == q))) return QCS_BAD_ARGUMENTS;
    const mat2 M = load_matrix(u_interleaved);
    const bool q_global = q >= reg->n_local;
    const bool c_global = c >= 0 && (unsigned) c >= reg->n_local;
    if (!q_global) {
        // target local: a global control is this rank's constant bit
        if (c_global) {
            if (!(((unsigned) reg->rank >> ((unsigned) c - reg->n_local)) & 1u)) return QCS_NO_ERROR;
            return launch_pairs(reg, reg->amp, 0, reg->N_local >> 1, q, -1, M);
        }
        return launch_pairs(reg, reg->amp, 0, c >= 0 ? reg->N_local >> 2 : reg->N_local >> 1, q, c, M);
    }
    // target global
    if (reg->peer) {
        // the pairs of the WHOLE register, this rank's share of them, on the stitched array
        const uint64_t items = c >= 0 ? reg->N >> 2 : reg->N >> 1;
        const uint64_t share = items / (uint64_t) reg->world;
        QCS_TRY(qcs_dist_stream_barrier(reg));
        QCS_TRY(launch_pairs(reg, reg->amp_all, (uint64_t) reg->rank * share, share, q, c, M));
        return qcs_dist_stream_barrier(reg);
    }
    return qcs_dist_gate_global(reg, q, c, u_interleaved);
}

extern "C" int qcs_apply_gate(qcs_register *reg, unsigned qubit_num, const double *u_interleaved)
{
    QCS_GROUP_FORWARD(reg, qcs_apply_gate(m, qubit_num, u_interleaved));
    if (!reg) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(reg->device));
    int rc = QCS_NO_ERROR;
    if (qcs_fuse_record_dense(reg, qubit_num, -1, u_interleaved, &rc)) return rc;
    QCS_TRY(qcs_fuse_flush(reg));
    return qcs_k_gate_1q(reg, qubit_num, -1, u_interleaved);
}

extern "C" int qcs_apply_controlled_gate(qcs_register *reg, unsigned c_qubit_num, unsigned qubit_num,
                                         const double *u_interleaved)
{
    QCS_GROUP_FORWARD(reg, qcs_apply_controlled_gate(m, c_qubit_num, qubit_num, u_interleaved));
    if (!reg) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(reg->device));
    int rc = QCS_NO_ERROR;
    if (qcs_fuse_record_dense(reg, qubit_num, (int) c_qubit_num, u_interleaved, &rc)) return rc;
    QCS_TRY(qcs_fuse_flush(reg));
    return qcs_k_gate_1q(reg, qubit_num, (int) c_qubit_num, u_interleaved);
}
