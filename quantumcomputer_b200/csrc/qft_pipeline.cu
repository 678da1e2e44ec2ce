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

namespace {

using namespace qft;

constexpr int kStages = 6;
constexpr int kGroups = 3;
constexpr int kGroupThreads = 128;
constexpr int kThreads = 64 + kGroups * kGroupThreads;     // 448
constexpr int kTileBits = 11;                               // 2048 amplitudes = 32 KiB per stage
constexpr uint32_t kTileBytes = 16u << kTileBits;

struct pipe_params {
    sweep_desc d;
    uint64_t n_tiles;
    int lo_gap;             // g_lo - a (strided) or -1 (contiguous final sweep)
    int prefetch;           // tiles ahead of the load that are prefetched into L2 (0: off)
};

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
__device__ __forceinline__ void group_barrier(int group)
{
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kGroupThreads) : "memory");
}

__device__ __forceinline__ void tile_coords(const pipe_params &P, uint64_t tix, int &c0, int &c1, int &c2)
{
    if (P.lo_gap >= 0) {
        c0 = (int) ((tix & ((1ull << P.lo_gap) - 1ull)) << (P.d.a + 1));   // doubles
        c1 = 0;
        c2 = (int) (tix >> P.lo_gap);
    } else {
        c0 = 0;
        c1 = (int) (tix << (kTileBits - 3));                                // rows of 8 amplitudes
        c2 = 0;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
k_qft_sweep_tma(const __grid_constant__ CUtensorMap tmap, const pipe_params P)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // kStages tiles; the 128-byte swizzle of the final sweep needs 1024-byte alignment
    double2 *stage_buf = (double2 *) (smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    double2 *wcol = stage_buf + (size_t) kStages * (1u << kTileBits);
    double2 *wbase = wcol + P.d.wcol_total;                                 // [kGroups][kMaxSteps]
    uint64_t *bars = (uint64_t *) (wbase + kGroups * kMaxSteps);
    uint64_t *full = bars, *computed = bars + kStages, *empty = bars + 2 * kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tile_geom G;
    G.a = P.d.a;
    G.g_lo = P.d.g_lo;
    G.sw = P.d.sw;
    const bool inv = P.d.inverse != 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&computed[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    // per-column part of the external twiddle: fixed for the whole kernel
    for (int k = 0; k < P.d.n_steps; k++) {
        const sweep_step S = P.d.step[k];
        const unsigned n_cols = 1u << (kTileBits - S.r);
        for (unsigned c = threadIdx.x; c < n_cols; c += kThreads) {
            const unsigned e_base = ((c >> S.s) << (S.s + S.r)) | (c & ((1u << S.s) - 1u));
            uint64_t y = 0;
            if (S.low_phys > P.d.lo) y = (G.spread(e_base) & ((1ull << S.low_phys) - 1ull)) >> P.d.lo;
            wcol[S.col_off + c] = unit_phase(y, S.j, inv);
        }
    }
    __syncthreads();

    // this CTA's tiles: tix = blockIdx.x + k * gridDim.x, k = 0 .. my_tiles-1
    const uint64_t my_tiles = P.n_tiles > blockIdx.x ? (P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int c0, c1, c2;
            // the smem ring holds ~2 tiles in flight per SM; an L2 prefetch a few tiles
            // ahead deepens the HBM pipeline without costing shared memory
            for (uint64_t k = 0; k < (uint64_t) P.prefetch && k < my_tiles; k++) {
                tile_coords(P, P.d.tile_first + blockIdx.x + k * gridDim.x, c0, c1, c2);
                tma_prefetch_3d(&tmap, c0, c1, c2);
            }
            for (uint64_t k = 0; k < my_tiles; k++) {
                const int s = (int) (k % kStages);
                const uint32_t round = (uint32_t) (k / kStages);
                if (P.prefetch && k + P.prefetch < my_tiles) {
                    tile_coords(P, P.d.tile_first + blockIdx.x + (k + P.prefetch) * gridDim.x, c0, c1, c2);
                    tma_prefetch_3d(&tmap, c0, c1, c2);
                }
                mbar_wait(&empty[s], (round & 1u) ^ 1u);
                tile_coords(P, P.d.tile_first + blockIdx.x + k * gridDim.x, c0, c1, c2);
                mbar_expect_tx(&full[s], kTileBytes);
                tma_load_3d(stage_buf + (size_t) s * (1u << kTileBits), &tmap, &full[s], c0, c1, c2);
            }
        }
    } else if (warp == 1) {
        // ---------------- store issuer ----------------
        if (lane == 0) {
            for (uint64_t k = 0; k < my_tiles; k++) {
                const int s = (int) (k % kStages);
                const uint32_t round = (uint32_t) (k / kStages);
                mbar_wait(&computed[s], round & 1u);
                int c0, c1, c2;
                tile_coords(P, P.d.tile_first + blockIdx.x + k * gridDim.x, c0, c1, c2);
                tma_store_3d(&tmap, stage_buf + (size_t) s * (1u << kTileBits), c0, c1, c2);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(&empty[s]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // ---------------- consumers ----------------
        const int group = (warp - 2) / (kGroupThreads / 32);
        const unsigned tig = threadIdx.x - 64 - group * kGroupThreads;
        double2 *my_wbase = wbase + group * kMaxSteps;
        const int lo_gap = P.lo_gap;
        for (uint64_t k = group; k < my_tiles; k += kGroups) {
            const int s = (int) (k % kStages);
            const uint32_t round = (uint32_t) (k / kStages);
            const uint64_t tix = P.d.tile_first + blockIdx.x + k * gridDim.x;
            const uint64_t base = lo_gap >= 0 ? (((tix >> lo_gap) << P.d.g_hi) | ((tix & ((1ull << lo_gap) - 1ull)) << P.d.a))
                                              : (tix << kTileBits);
            if (tig < (unsigned) P.d.n_steps) {
                const sweep_step S = P.d.step[tig];
                uint64_t y = 0;
                if (S.low_phys > P.d.lo) y = (base & ((1ull << S.low_phys) - 1ull)) >> P.d.lo;
                my_wbase[tig] = unit_phase(y + P.d.y_const, S.j, inv);
            }
            group_barrier(group);
            double2 *tile = stage_buf + (size_t) s * (1u << kTileBits);
            mbar_wait(&full[s], round & 1u);
            for (int st = 0; st < P.d.n_steps; st++) {
                const sweep_step S = P.d.step[st];
                const bool last = st == P.d.n_steps - 1;
                const double2 wb = my_wbase[st];
                if (P.d.hadamard_only) dispatch_step<true, false>(nullptr, tile, wcol + S.col_off, wb, G, S, kTileBits, base, false, false, last, P.d.scale, tig, kGroupThreads, P.d.diag, P.d.n_diag, P.d.index_or);
                else if (inv) dispatch_step<true>(nullptr, tile, wcol + S.col_off, wb, G, S, kTileBits, 0, false, false, last, P.d.scale, tig, kGroupThreads);
                else dispatch_step<false>(nullptr, tile, wcol + S.col_off, wb, G, S, kTileBits, 0, false, false, last, P.d.scale, tig, kGroupThreads);
                if (last) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to TMA
                group_barrier(group);
            }
            if (tig == 0) mbar_arrive(&computed[s]);
        }
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

// true when the pipelined kernel can run this sweep
bool qcs_pipeline_supports(const qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &p)
{
    if (p.d.t != kTileBits || tg.n_bits < 14) return false;
    const bool strided = p.d.g_lo > p.d.a;
    if (strided && (p.d.a < 1 || p.d.a > 7 || p.d.g_hi - p.d.g_lo > 8 || p.d.g_lo + 1 > 31)) return false;
    // the layout TMA writes (linear, or 128-byte swizzle for the contiguous sweep)
    // must make every step's shared-memory access conflict-free
    if (conflict_cost(p.d, strided ? 28 : 3, true) != 0) return false;
    const size_t smem = (size_t) kStages * kTileBytes + 16 * (size_t) p.d.wcol_total + 16 * kGroups * kMaxSteps +
                        8 * 3 * kStages + 1024;
    return smem <= reg->smem_optin;
}

int qcs_pipeline_launch(qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &plan)
{
    encode_fn_t encode = get_encode();
    if (!encode) {
        fprintf(stderr, "qcs: cuTensorMapEncodeTiled is unavailable\n");
        return QCS_UNKNOWN_ERROR;
    }
    pipe_params P;
    P.d = plan.d;
    P.n_tiles = plan.n_tiles;
    P.prefetch = reg->opt_prefetch_tiles;
    CUtensorMap tmap;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    CUtensorMapSwizzle swz;
    const bool strided = plan.d.g_lo > plan.d.a;
    if (strided) {
        const int g = plan.d.g_hi - plan.d.g_lo;
        dims[0] = 2ull << plan.d.g_lo;                       // doubles below g_lo
        dims[1] = 1ull << g;
        dims[2] = 1ull << (tg.n_bits - (unsigned) plan.d.g_hi);
        strides[0] = 16ull << plan.d.g_lo;
        strides[1] = 16ull << plan.d.g_hi;
        box[0] = 2u << plan.d.a;
        box[1] = 1u << g;
        box[2] = 1;
        swz = CU_TENSOR_MAP_SWIZZLE_NONE;
        P.lo_gap = plan.d.g_lo - plan.d.a;
        P.d.sw = 28;                                         // no XOR
    } else {
        dims[0] = 16;                                        // 128 B rows
        dims[1] = (1ull << tg.n_bits) >> 3;
        dims[2] = 1;
        strides[0] = 128;
        strides[1] = 128ull * dims[1];
        box[0] = 16;
        box[1] = 1u << (kTileBits - 3);
        box[2] = 1;
        swz = CU_TENSOR_MAP_SWIZZLE_128B;
        P.lo_gap = -1;
        P.d.sw = 3;
    }
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, tg.amp, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        fprintf(stderr, "qcs: cuTensorMapEncodeTiled failed (%d)\n", (int) cr);
        return QCS_UNKNOWN_ERROR;
    }
    const size_t smem = (size_t) kStages * kTileBytes + 16 * (size_t) P.d.wcol_total + 16 * kGroups * kMaxSteps +
                        8 * 3 * kStages + 1024;
    QCS_CUDA(cudaFuncSetAttribute(k_qft_sweep_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    uint64_t grid = (uint64_t) reg->sm_count;
    if (grid > plan.n_tiles) grid = plan.n_tiles;
    qcs_launch_begin(reg, QCS_K_TILE_SWEEP, 32.0 * (double) (plan.n_tiles << kTileBits));
    k_qft_sweep_tma<<<(unsigned) grid, kThreads, smem, tg.stream>>>(tmap, P);
    return qcs_launch_end(reg, QCS_K_TILE_SWEEP, "k_qft_sweep_tma");
}
