// qft_pipeline.cu -- the fused-QFT sweep as a TMA + mbarrier pipeline.
//
// Same mathematics as qft_fused.cu (steps of radix-2^r register butterflies,
// one external twiddle per element), but the HBM traffic is taken out of the
// compute warps' hands:
//
//   warp 0        producer: one elected lane issues cp.async.bulk.tensor loads
//                 (TMA) of whole tiles into a ring of kStages shared-memory
//                 stages, each guarded by a `full` mbarrier (complete_tx bytes)
//   warp 1        store issuer: waits for a stage to be `computed`, issues the
//                 TMA store of the tile back to its place, and when the store
//                 has finished reading shared memory frees the stage (`empty`)
//   warps 2..     kGroups consumer groups of 128 threads; group g owns every
//                 kGroups-th tile of the CTA and runs all steps of the sweep on
//                 it in shared memory (named barrier per group between steps)
//
// so at any time one SM has tiles loading, tiles being transformed and a tile
// draining, and the compute warps never wait on a global-memory round trip.
// One persistent CTA per SM; tiles are dealt round-robin.
//
// Tile shapes (physical index bits [0,a) U [g_lo,g_hi), 16-byte amplitudes):
//   strided sweep : 3-D tensor map  {2^(a+1) doubles = 256 B} x {2^g rows,
//                   stride 2^g_lo * 16 B} x {1 of the remaining index};  every
//                   step has s >= a so shared-memory accesses are conflict-free
//                   without swizzling
//   final sweep   : contiguous 2^t amplitudes viewed as rows of 128 B with
//                   CU_TENSOR_MAP_SWIZZLE_128B, which is exactly the XOR pattern
//                   e ^ ((e >> 3) & 7) on 16-byte units that makes the s = 0,
//                   r = 3 step conflict-free.
#include "qft_common.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace {

using namespace qft;

// pipeline shapes (template parameters of the kernel): log2 tile size, ring depth,
// consumer groups and threads per group
struct pipe_params {
    sweep_desc d;
    uint64_t n_tiles;
    int lo_gap;             // g_lo - a (strided) or -1 (contiguous final sweep)
    int prefetch;           // tiles ahead of the load that are prefetched into L2 (0: off)
    int n_boxes;            // TMA boxes per tile (a box has at most 256 rows)
    int box_rows;
    uint32_t box_bytes;
    unsigned long long *timing;   // -DQCS_PIPE_TIMING builds only: per-CTA cycle counters
};

// Role timing (where do the producer, the store issuer and the consumer groups spend their
// cycles) is compiled in only with -DQCS_PIPE_TIMING and switched on by the environment
// variable of the same name; production builds carry none of it.
#ifdef QCS_PIPE_TIMING
#define QCS_TICK(P) ((P).timing ? clock64() : 0)
#define QCS_TIMING_ADD(P, slot, cycles)                                                      \
    do {                                                                                     \
        if ((P).timing) (P).timing[blockIdx.x * 16 + (slot)] += (unsigned long long) (cycles); \
    } while (0)
#define QCS_TIMING_SET(P, slot, value)                                              \
    do {                                                                            \
        if ((P).timing) (P).timing[blockIdx.x * 16 + (slot)] = (unsigned long long) (value); \
    } while (0)
#else
#define QCS_TICK(P) 0ll
#define QCS_TIMING_ADD(P, slot, cycles) do { (void) (cycles); } while (0)
#define QCS_TIMING_SET(P, slot, value) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void group_barrier(int group, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(threads) : "memory");
}

template <int TB>
__device__ __forceinline__ void tile_coords(const pipe_params &P, uint64_t tix, int &c0, int &c1, int &c2)
{
    if (P.lo_gap >= 0) {
        c0 = (int) ((tix & ((1ull << P.lo_gap) - 1ull)) << (P.d.a + 1));   // doubles
        c1 = 0;
        c2 = (int) (tix >> P.lo_gap);
    } else {
        c0 = 0;
        c1 = (int) (tix << (TB - 3));                                       // rows of 8 amplitudes
        c2 = 0;
    }
}

template <int TB, int STAGES, int GROUPS, int GT>
__global__ void __launch_bounds__(64 + GROUPS * GT, 1)
k_qft_sweep_tma(const __grid_constant__ CUtensorMap tmap, const pipe_params P)
{
    static_assert(STAGES % GROUPS == 0, "a ring stage must always be consumed by the same group");
    constexpr int kThreads = 64 + GROUPS * GT;
    constexpr uint32_t kTileBytes = 16u << TB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // STAGES tiles; the 128-byte swizzle of the final sweep needs 1024-byte alignment
    double2 *stage_buf = (double2 *) (smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    double2 *wcol = stage_buf + (size_t) STAGES * (1u << TB);
    double2 *wbase = wcol + P.d.wcol_total;                                 // [GROUPS][kMaxSteps]
    uint64_t *bars = (uint64_t *) (wbase + GROUPS * kMaxSteps);
    uint64_t *full = bars, *computed = bars + STAGES, *empty = bars + 2 * STAGES;
    diag_gate *sdiag = (diag_gate *) (bars + 3 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tile_geom G;
    G.a = P.d.a;
    G.g_lo = P.d.g_lo;
    G.sw = P.d.sw;
    const bool inv = P.d.inverse != 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&computed[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    // per-column part of the external twiddle: fixed for the whole kernel
    if (!P.d.hadamard_only) {
        for (int k = 0; k < P.d.n_steps; k++) {
            const sweep_step S = P.d.step[k];
            const unsigned n_cols = 1u << (TB - S.r);
            for (unsigned c = threadIdx.x; c < n_cols; c += kThreads) {
                const unsigned e_base = ((c >> S.s) << (S.s + S.r)) | (c & ((1u << S.s) - 1u));
                uint64_t y = 0;
                if (S.low_phys > P.d.lo) y = (G.spread(e_base) & ((1ull << S.low_phys) - 1ull)) >> P.d.lo;
                wcol[S.col_off + c] = unit_phase(y, S.j, inv);
            }
        }
    }
    for (int i = threadIdx.x; i < P.d.n_diag; i += kThreads) sdiag[i] = P.d.diag[i];
    __syncthreads();

    // this CTA's tiles: tix = blockIdx.x + k * gridDim.x, k = 0 .. my_tiles-1
    const uint64_t my_tiles = P.n_tiles > blockIdx.x ? (P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int c0, c1, c2;
            // optional L2 prefetch a few tiles ahead of the shared-memory ring
            for (uint64_t k = 0; k < (uint64_t) P.prefetch && k < my_tiles; k++) {
                tile_coords<TB>(P, tile_number(P.d, blockIdx.x + k * gridDim.x), c0, c1, c2);
                for (int b = 0; b < P.n_boxes; b++) tma_prefetch_3d(&tmap, c0, c1 + b * P.box_rows, c2);
            }
            for (uint64_t k = 0; k < my_tiles; k++) {
                const int s = (int) (k % STAGES);
                const uint32_t round = (uint32_t) (k / STAGES);
                if (P.prefetch && k + P.prefetch < my_tiles) {
                    tile_coords<TB>(P, tile_number(P.d, blockIdx.x + (k + P.prefetch) * gridDim.x), c0, c1, c2);
                    for (int b = 0; b < P.n_boxes; b++) tma_prefetch_3d(&tmap, c0, c1 + b * P.box_rows, c2);
                }
                const long long t0 = QCS_TICK(P);
                mbar_wait(&empty[s], (round & 1u) ^ 1u);
                QCS_TIMING_ADD(P, 0, QCS_TICK(P) - t0);
                tile_coords<TB>(P, tile_number(P.d, blockIdx.x + k * gridDim.x), c0, c1, c2);
                mbar_expect_tx(&full[s], kTileBytes);
                unsigned char *dst = (unsigned char *) (stage_buf + (size_t) s * (1u << TB));
                for (int b = 0; b < P.n_boxes; b++)
                    tma_load_3d(dst + (size_t) b * P.box_bytes, &tmap, &full[s], c0, c1 + b * P.box_rows, c2);
            }
        }
    } else if (warp == 1) {
        // ---------------- store issuer ----------------
        if (lane == 0) {
            for (uint64_t k = 0; k < my_tiles; k++) {
                const int s = (int) (k % STAGES);
                const uint32_t round = (uint32_t) (k / STAGES);
                const long long t0 = QCS_TICK(P);
                mbar_wait(&computed[s], round & 1u);
                const long long t1 = QCS_TICK(P);
                int c0, c1, c2;
                tile_coords<TB>(P, tile_number(P.d, blockIdx.x + k * gridDim.x), c0, c1, c2);
                const unsigned char *src = (const unsigned char *) (stage_buf + (size_t) s * (1u << TB));
                for (int b = 0; b < P.n_boxes; b++)
                    tma_store_3d(&tmap, src + (size_t) b * P.box_bytes, c0, c1 + b * P.box_rows, c2);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                QCS_TIMING_ADD(P, 1, t1 - t0);               // waiting for a computed tile
                QCS_TIMING_ADD(P, 2, QCS_TICK(P) - t1);      // the store reading shared memory
                mbar_arrive(&empty[s]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // ---------------- consumers ----------------
        const int group = (warp - 2) / (GT / 32);
        const unsigned tig = threadIdx.x - 64 - group * GT;
        double2 *my_wbase = wbase + group * kMaxSteps;
        const int lo_gap = P.lo_gap;
        for (uint64_t k = group; k < my_tiles; k += GROUPS) {
            const int s = (int) (k % STAGES);
            const uint32_t round = (uint32_t) (k / STAGES);
            const uint64_t tix = tile_number(P.d, blockIdx.x + k * gridDim.x);
            const uint64_t base = lo_gap >= 0 ? (((tix >> lo_gap) << P.d.g_hi) | ((tix & ((1ull << lo_gap) - 1ull)) << P.d.a))
                                              : (tix << TB);
            if (!P.d.hadamard_only && tig < (unsigned) P.d.n_steps) {
                const sweep_step S = P.d.step[tig];
                uint64_t y = 0;
                if (S.low_phys > P.d.lo) y = (base & ((1ull << S.low_phys) - 1ull)) >> P.d.lo;
                my_wbase[tig] = unit_phase(y + P.d.y_const, S.j, inv);
            }
            group_barrier(group, GT);
            double2 *tile = stage_buf + (size_t) s * (1u << TB);
            const long long t0 = QCS_TICK(P);
            mbar_wait(&full[s], round & 1u);
            const long long t1 = QCS_TICK(P);
            for (int st = 0; st < P.d.n_steps; st++) {
                const sweep_step S = P.d.step[st];
                const bool last = st == P.d.n_steps - 1;
                const double2 wb = my_wbase[st];
                if (P.d.hadamard_only) dispatch_step<true, false>(nullptr, tile, wcol + S.col_off, wb, G, S, TB, base, false, false, last, P.d.scale, tig, GT, sdiag, P.d.n_diag, P.d.index_or);
                else if (inv) dispatch_step<true>(nullptr, tile, wcol + S.col_off, wb, G, S, TB, base, false, false, last, P.d.scale, tig, GT);
                else dispatch_step<false>(nullptr, tile, wcol + S.col_off, wb, G, S, TB, base, false, false, last, P.d.scale, tig, GT);
                if (last) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to TMA
                group_barrier(group, GT);
            }
            if (tig == 0) {
                mbar_arrive(&computed[s]);
                QCS_TIMING_ADD(P, 4 + 2 * group, t1 - t0);              // waiting for the load
                QCS_TIMING_ADD(P, 5 + 2 * group, QCS_TICK(P) - t1);     // the steps
            }
        }
        if (tig == 0 && group == 0) QCS_TIMING_SET(P, 3, my_tiles);
    }
}

typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn_t get_encode()
{
    static encode_fn_t fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (encode_fn_t) p;
    return fn;
}

}  // namespace

// ---------------------------------------------------------------------------
// host side: the kernel shapes that are instantiated and how a plan maps to one
// ---------------------------------------------------------------------------
namespace {

struct pipe_shape { int tb, stages, groups, gt; };
// A stage of the ring must always be consumed by the same group (STAGES % GROUPS == 0): a
// group is in order, so it can never wait on a `full` barrier two phases ahead of the barrier's
// current phase -- which a parity wait cannot tell from "already complete".
constexpr pipe_shape kShapes[] = {
    {12, 3, 3, 128},    // 0: 64 KiB tiles, one group per stage: one sweep fewer at n = 30
    {11, 6, 3, 128},    // 1: 32 KiB tiles, three groups of 128, each owning two stages
    {11, 6, 6, 64},     // 2: 32 KiB tiles, six groups of 64
    {12, 3, 1, 256},    // 3
    {12, 3, 3, 64},     // 4
    {11, 6, 2, 128},    // 5
};
constexpr int kNumShapes = (int) (sizeof kShapes / sizeof kShapes[0]);

size_t pipe_smem(const pipe_shape &sh, const qft::sweep_desc &d)
{
    return (size_t) sh.stages * ((size_t) 16 << sh.tb) + 16 * (size_t) d.wcol_total + 16 * (size_t) sh.groups * kMaxSteps +
           8 * 3 * (size_t) sh.stages + sizeof(diag_gate) * (size_t) d.n_diag + 1024;
}

template <int TB, int STAGES, int GROUPS, int GT>
int launch_shape(qcs_register *reg, const CUtensorMap &tmap, const pipe_params &P, size_t smem, const qft::sweep_target &tg)
{
    auto kern = k_qft_sweep_tma<TB, STAGES, GROUPS, GT>;
    QCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    uint64_t grid = (uint64_t) reg->sm_count;
    if (tg.max_ctas > 0 && grid > (uint64_t) tg.max_ctas) grid = (uint64_t) tg.max_ctas;
    if (grid > P.n_tiles) grid = P.n_tiles;
    qcs_launch_begin(reg, tg.kind, tg.bytes > 0.0 ? tg.bytes : 32.0 * (double) (P.n_tiles << TB));
    kern<<<(unsigned) grid, 64 + GROUPS * GT, smem, tg.stream>>>(tmap, P);
    return qcs_launch_end(reg, tg.kind, "k_qft_sweep_tma");
}

}  // namespace

static int launch_by_shape(qcs_register *reg, int shape_id, const CUtensorMap &tmap, const pipe_params &P, size_t smem,
                           const qft::sweep_target &tg);

int qcs_pipeline_tile_bits(const qcs_register *reg)
{
    const int v = reg->opt_pipe_shape >= 0 && reg->opt_pipe_shape < kNumShapes ? reg->opt_pipe_shape : 0;
    return kShapes[v].tb;
}

// true when the pipelined kernel can run this sweep
bool qcs_pipeline_supports(const qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &p)
{
    const pipe_shape sh = kShapes[reg->opt_pipe_shape >= 0 && reg->opt_pipe_shape < kNumShapes ? reg->opt_pipe_shape : 0];
    if (p.d.t != sh.tb || tg.n_bits < (unsigned) sh.tb + 3) return false;
    const bool strided = p.d.g_lo > p.d.a;
    // a TMA box row is 2^(a+1) doubles (32 B .. 2 KiB); coordinates are 32-bit
    if (strided && (p.d.a < 1 || p.d.a > 7 || p.d.g_lo + 1 > 31)) return false;
    // the linear layout TMA writes for a strided tile must make every step conflict-free; the
    // contiguous sweep takes the 128-byte hardware swizzle (conflict-free for every step except a
    // radix-16 step on bits 0..3, which is 2-way conflicted)
    if (strided && conflict_cost(p.d, 28, true) != 0) return false;
    return pipe_smem(sh, p.d) <= reg->smem_optin;
}

int qcs_pipeline_launch(qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &plan)
{
    encode_fn_t encode = get_encode();
    if (!encode) {
        fprintf(stderr, "qcs: cuTensorMapEncodeTiled is unavailable\n");
        return QCS_UNKNOWN_ERROR;
    }
    const int shape_id = reg->opt_pipe_shape >= 0 && reg->opt_pipe_shape < kNumShapes ? reg->opt_pipe_shape : 0;
    const pipe_shape sh = kShapes[shape_id];
    pipe_params P;
    P.d = plan.d;
    P.n_tiles = plan.n_tiles;
    P.prefetch = reg->opt_prefetch_tiles;
    CUtensorMap tmap;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    CUtensorMapSwizzle swz;
    const bool strided = plan.d.g_lo > plan.d.a;
    unsigned rows;
    if (strided) {
        const int g = plan.d.g_hi - plan.d.g_lo;
        rows = 1u << g;
        dims[0] = 2ull << plan.d.g_lo;                       // doubles below g_lo
        dims[1] = rows;
        dims[2] = 1ull << (tg.n_bits - (unsigned) plan.d.g_hi);
        strides[0] = 16ull << plan.d.g_lo;
        strides[1] = 16ull << plan.d.g_hi;
        box[0] = 2u << plan.d.a;
        swz = CU_TENSOR_MAP_SWIZZLE_NONE;
        P.lo_gap = plan.d.g_lo - plan.d.a;
        P.d.sw = 28;                                         // no XOR
    } else {
        rows = 1u << (sh.tb - 3);
        dims[0] = 16;                                        // 128 B rows
        dims[1] = (1ull << tg.n_bits) >> 3;
        dims[2] = 1;
        strides[0] = 128;
        strides[1] = 128ull * dims[1];
        box[0] = 16;
        swz = CU_TENSOR_MAP_SWIZZLE_128B;
        P.lo_gap = -1;
        P.d.sw = 3;
    }
    P.box_rows = rows < 256u ? (int) rows : 256;             // a box dimension is at most 256
    P.n_boxes = (int) (rows / (unsigned) P.box_rows);
    P.box_bytes = (uint32_t) (((size_t) 16 << sh.tb) / (size_t) P.n_boxes);
    box[1] = (cuuint32_t) P.box_rows;
    box[2] = 1;
    // 128 B rows: promoting the requests to 256 B would fetch a neighbour's half-line with every row
    // (measured at n = 30: 21.2-21.4 ms for every promotion setting -- not a lever)
    const CUtensorMapL2promotion promo = (strided && plan.d.a <= 3) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                    : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, tg.amp, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        fprintf(stderr, "qcs: cuTensorMapEncodeTiled failed (%d)\n", (int) cr);
        return QCS_UNKNOWN_ERROR;
    }
    const size_t smem = pipe_smem(sh, P.d);
    // -DQCS_PIPE_TIMING builds: QCS_PIPE_TIMING=1 prints where the roles of the pipeline spend their cycles
#ifdef QCS_PIPE_TIMING
    static const bool timing_on = getenv("QCS_PIPE_TIMING") != nullptr;
#else
    const bool timing_on = false;
#endif
    P.timing = nullptr;
    if (timing_on) {
        QCS_CUDA(cudaMalloc((void **) &P.timing, 16 * 8 * (size_t) reg->sm_count));
        QCS_CUDA(cudaMemsetAsync(P.timing, 0, 16 * 8 * (size_t) reg->sm_count, tg.stream));
    }
    const int rc = launch_by_shape(reg, shape_id, tmap, P, smem, tg);
    if (timing_on && rc == QCS_NO_ERROR) {
        std::vector<unsigned long long> h(16 * (size_t) reg->sm_count);
        QCS_CUDA(cudaMemcpyAsync(h.data(), P.timing, h.size() * 8, cudaMemcpyDeviceToHost, tg.stream));
        QCS_CUDA(cudaStreamSynchronize(tg.stream));
        double sum[16] = {0};
        for (size_t i = 0; i < h.size(); i++) sum[i % 16] += (double) h[i];
        const double tiles = sum[3] > 0 ? sum[3] : 1;
        fprintf(stderr, "qcs pipe timing (cycles per tile, avg over CTAs) shape %d a=%d g=[%d,%d) steps=%d: producer wait empty %.0f | "
                "store wait computed %.0f, store read %.0f |", shape_id, plan.d.a, plan.d.g_lo, plan.d.g_hi, plan.d.n_steps,
                sum[0] / tiles, sum[1] / tiles, sum[2] / tiles);
        for (int g = 0; g < sh.groups && g < 6; g++)
            fprintf(stderr, " g%d wait load %.0f steps %.0f;", g, sum[4 + 2 * g] / tiles * sh.groups, sum[5 + 2 * g] / tiles * sh.groups);
        fprintf(stderr, "\n");
        cudaFree(P.timing);
    }
    return rc;
}

static int launch_by_shape(qcs_register *reg, int shape_id, const CUtensorMap &tmap, const pipe_params &P, size_t smem,
                           const qft::sweep_target &tg)
{
    switch (shape_id) {
        case 1: return launch_shape<11, 6, 3, 128>(reg, tmap, P, smem, tg);
        case 2: return launch_shape<11, 6, 6, 64>(reg, tmap, P, smem, tg);
        case 3: return launch_shape<12, 3, 1, 256>(reg, tmap, P, smem, tg);
        case 4: return launch_shape<12, 3, 3, 64>(reg, tmap, P, smem, tg);
        case 5: return launch_shape<11, 6, 2, 128>(reg, tmap, P, smem, tg);
        default: return launch_shape<12, 3, 3, 128>(reg, tmap, P, smem, tg);
    }
}
