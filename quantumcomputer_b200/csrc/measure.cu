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

int qcs_k_measure_scan(qcs_register *reg, double cum_in, double r, uint64_t limit,
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
