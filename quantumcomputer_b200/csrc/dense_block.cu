// dense_block.cu -- a run of k low qubits fused into ONE dense 2^k x 2^k
// complex block, applied with the FP64 tensor-core path (DMMA).
//
// new[g * 2^k + i] = sum_j U[i][j] * old[g * 2^k + j]   for every group g of 2^k
// consecutive amplitudes, i.e. an arbitrary unitary on qubits 0..k-1 -- what a
// run of single-/two-qubit gates on those qubits multiplies out to (the
// reference's HADAMARD_BASE_MATRIX / C_PHASE_SHIFT_BASE_MATRIX products,
// qc_shor.c:210-225, generalised).  This is the one place of the path that is a
// dense contraction: a [2R x 2R] real matrix (R = 2^k, complex arithmetic
// unrolled as [Ur -Ui; Ui Ur]) times a [2R x (N/R)] matrix of amplitudes.
//
// Blackwell's tcgen05.mma has no f64 kind; the FP64 tensor path on sm_100a is
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA).  One warp owns 8 groups (columns)
// per tile:
//   * B fragment (K x 8): lane (g = lane/4, t = lane%4) supplies, for k-step s,
//     row t of the slab.  The K order is permuted so that this is memory double
//     8t + s (k = 4) of group g: each lane loads one contiguous 64-byte piece.
//   * A fragment (8 x 4 per m-tile, k-step): the constant matrix, permuted the
//     same way, lives in registers for the whole kernel (2R*2R/32 doubles/lane).
//   * D rows are ordered so lane (g, t) ends up with 4 (k = 4) consecutive
//     doubles of groups 2t and 2t+1: 32-byte vector stores.
// HBM traffic is one read + one write of the state (32 B per amplitude);
// arithmetic is 8 * 2^k flops per amplitude.
#include "qcs_internal.h"

namespace {

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// K2 = 2R real components per group; MT = K2/8 m-tiles; KS = K2/4 k-steps;
// lane piece = K2/4 consecutive doubles
template <int K2>
__global__ void __launch_bounds__(256)
k_dense_block(double *__restrict__ amp, uint64_t n_groups, const double *__restrict__ a_frag)
{
    constexpr int MT = K2 / 8, KS = K2 / 4, PIECE = K2 / 4, OUT = K2 / 8;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    // constant operand: a_frag[(mt * KS + s) * 32 + lane]
    double a[MT][KS];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int s = 0; s < KS; s++) a[mt][s] = a_frag[(mt * KS + s) * 32 + lane];

    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t warp_id = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_tiles = n_groups >> 3;               // 8 groups per warp tile
    for (uint64_t tile = warp_id; tile < n_tiles; tile += warps) {
        double *base = amp + (tile << 3) * K2;            // 8 groups * K2 doubles
        // B: this lane's contiguous piece of group g
        double b[PIECE];
        const double *src = base + (uint64_t) g * K2 + t * PIECE;
#pragma unroll
        for (int s = 0; s < PIECE; s += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(src + s);
            b[s] = v.x;
            b[s + 1] = v.y;
        }
        double d[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) d[mt][0] = d[mt][1] = 0.0;
#pragma unroll
        for (int s = 0; s < KS; s++)
#pragma unroll
            for (int mt = 0; mt < MT; mt++) dmma_8x8x4(d[mt][0], d[mt][1], a[mt][s], b[s]);
        // D: lane (g, t) holds rows (mt, g) = components OUT*g + mt of groups 2t, 2t+1
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; c++) {
            double *dst = base + (uint64_t) (2 * t + c) * K2 + g * OUT;
#pragma unroll
            for (int mt = 0; mt < MT; mt += 2)
                *reinterpret_cast<double2 *>(dst + mt) = make_double2(d[mt][c], d[mt + 1][c]);
        }
    }
}

}  // namespace

extern "C" int qcs_apply_dense_block(qcs_register *reg, unsigned k, const double *u_interleaved)
{
    QCS_GROUP_FORWARD(reg, qcs_apply_dense_block(m, k, u_interleaved));
    if (!reg || !u_interleaved) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(reg->device));
    QCS_TRY(qcs_fuse_flush(reg));
    return qcs_k_dense_block(reg, k, u_interleaved);
}

// the launch itself (no flush): also what the gate stream emits for a run of general gates on the low qubits
int qcs_k_dense_block(qcs_register *reg, unsigned k, const double *u_interleaved)
{
    if ((k != 3 && k != 4) || k + 3 > reg->n_local) return QCS_BAD_ARGUMENTS;
    const int R = 1 << k, K2 = 2 * R;
    const int MT = K2 / 8, KS = K2 / 4, PIECE = K2 / 4, OUT = K2 / 8;
    // real form: component index = 2*j + c (c = 0 re, 1 im) = position in memory
    //   D[2i]   = sum_j Ur[i][j] X[2j] - Ui[i][j] X[2j+1]
    //   D[2i+1] = sum_j Ui[i][j] X[2j] + Ur[i][j] X[2j+1]
    std::vector<double> real((size_t) K2 * K2);
    for (int i = 0; i < R; i++)
        for (int j = 0; j < R; j++) {
            const double ur = u_interleaved[2 * (i * R + j)], ui = u_interleaved[2 * (i * R + j) + 1];
            real[(size_t) (2 * i) * K2 + 2 * j] = ur;
            real[(size_t) (2 * i) * K2 + 2 * j + 1] = -ui;
            real[(size_t) (2 * i + 1) * K2 + 2 * j] = ui;
            real[(size_t) (2 * i + 1) * K2 + 2 * j + 1] = ur;
        }
    // A fragment for (m-tile mt, k-step s), lane (g, t): row = OUT*g + mt, column = PIECE*t + s
    std::vector<double> frag((size_t) MT * KS * 32);
    for (int mt = 0; mt < MT; mt++)
        for (int s = 0; s < KS; s++)
            for (int lane = 0; lane < 32; lane++) {
                const int g = lane >> 2, t = lane & 3;
                frag[(size_t) (mt * KS + s) * 32 + lane] = real[(size_t) (OUT * g + mt) * K2 + PIECE * t + s];
            }
    // the fragment buffer lives in the handle (8 KiB covers k = 4): no allocation, no host
    // synchronisation per call.  The copy is stream-ordered behind the previous call's kernel, and a
    // pageable source is staged by the runtime before cudaMemcpyAsync returns, so `frag` may go.
    if (!reg->d_dense) QCS_CUDA(cudaMalloc(&reg->d_dense, 1024 * sizeof(double)));
    double *d_frag = (double *) reg->d_dense;
    QCS_CUDA(cudaMemcpyAsync(d_frag, frag.data(), frag.size() * sizeof(double), cudaMemcpyHostToDevice, reg->stream));
    const uint64_t n_groups = reg->N_local >> k;
    uint64_t grid = (n_groups / 8 + 7) / 8;
    const uint64_t cap = (uint64_t) reg->sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    qcs_launch_begin(reg, QCS_K_DENSE_BLOCK, 32.0 * (double) reg->N_local);
    if (k == 4) k_dense_block<32><<<(unsigned) grid, 256, 0, reg->stream>>>((double *) reg->amp, n_groups, d_frag);
    else k_dense_block<16><<<(unsigned) grid, 256, 0, reg->stream>>>((double *) reg->amp, n_groups, d_frag);
    return qcs_launch_end(reg, QCS_K_DENSE_BLOCK, "k_dense_block");
}
