// ubench_tma.cu -- what can the TMA unit of one SM move?
//
// The pipelined sweep kernel brings every tile in and out with TMA.  This probe runs the same
// transfer skeleton WITHOUT any computation: one persistent CTA per SM, a ring of 3 x 64 KiB stages,
// one thread issuing loads (cp.async.bulk.tensor -> mbarrier), another issuing the stores of the
// loaded tiles back to where they came from.  Per tile shape (the boxes the sweeps use, plus plain
// 1-D bulk copies) it reports GB/s (read + write) and bytes per clock per SM, i.e. the ceiling that
// the transfers alone put on a sweep.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/ubench_tma tools/ubench_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int kStages = 3;
constexpr uint32_t kTileBytes = 65536;

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

struct params {
    uint64_t n_tiles;
    int mode;          // 0: tensor boxes, 1: 1-D bulk copies of `piece` bytes
    int lo_gap, a, n_boxes, box_rows;
    uint32_t box_bytes;
    uint32_t piece;
    double2 *amp;
    int stores;        // 0: loads only
};

__global__ void __launch_bounds__(128, 1) k_tma_copy(const __grid_constant__ CUtensorMap tmap, const params P)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *stage = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = (uint64_t *) (stage + (size_t) kStages * kTileBytes);
    uint64_t *full = bars, *empty = bars + kStages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t my = P.n_tiles > blockIdx.x ? (P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (uint64_t k = 0; k < my; k++) {
            const int s = (int) (k % kStages);
            const uint32_t round = (uint32_t) (k / kStages);
            mbar_wait(&empty[s], (round & 1u) ^ 1u);
            const uint64_t tix = blockIdx.x + k * gridDim.x;
            mbar_expect_tx(&full[s], kTileBytes);
            unsigned char *dst = stage + (size_t) s * kTileBytes;
            if (P.mode == 0) {
                int c0, c1 = 0, c2;
                if (P.lo_gap >= 0) { c0 = (int) ((tix & ((1ull << P.lo_gap) - 1ull)) << (P.a + 1)); c2 = (int) (tix >> P.lo_gap); }
                else { c0 = 0; c1 = (int) (tix << 9); c2 = 0; }
                for (int b = 0; b < P.n_boxes; b++)
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                                     "r"(smem_u32(dst + (size_t) b * P.box_bytes)), "l"(&tmap), "r"(smem_u32(&full[s])), "r"(c0), "r"(c1 + b * P.box_rows), "r"(c2)
                                 : "memory");
            } else {
                const unsigned char *src = (const unsigned char *) (P.amp + (tix << 12));
                for (uint32_t off = 0; off < kTileBytes; off += P.piece)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                                     "r"(smem_u32(dst + off)), "l"(src + off), "r"(P.piece), "r"(smem_u32(&full[s]))
                                 : "memory");
            }
        }
    } else if (threadIdx.x == 32) {
        for (uint64_t k = 0; k < my; k++) {
            const int s = (int) (k % kStages);
            const uint32_t round = (uint32_t) (k / kStages);
            mbar_wait(&full[s], round & 1u);
            const uint64_t tix = blockIdx.x + k * gridDim.x;
            const unsigned char *src = stage + (size_t) s * kTileBytes;
            if (P.stores) {
                if (P.mode == 0) {
                    int c0, c1 = 0, c2;
                    if (P.lo_gap >= 0) { c0 = (int) ((tix & ((1ull << P.lo_gap) - 1ull)) << (P.a + 1)); c2 = (int) (tix >> P.lo_gap); }
                    else { c0 = 0; c1 = (int) (tix << 9); c2 = 0; }
                    for (int b = 0; b < P.n_boxes; b++)
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap),
                                     "r"(smem_u32(src + (size_t) b * P.box_bytes)), "r"(c0), "r"(c1 + b * P.box_rows), "r"(c2)
                                     : "memory");
                } else {
                    unsigned char *dst = (unsigned char *) (P.amp + (tix << 12));
                    for (uint32_t off = 0; off < kTileBytes; off += P.piece)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_u32(src + off)), "r"(P.piece)
                                     : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            mbar_arrive(&empty[s]);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const int n = 30;
    double2 *buf;
    CK(cudaMalloc(&buf, 16ull << n));
    CK(cudaMemset(buf, 0, 16ull << n));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
    encode_fn_t encode = (encode_fn_t) fp;
    const size_t smem = (size_t) kStages * kTileBytes + 64 + 1024;
    CK(cudaFuncSetAttribute(k_tma_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    printf("device %s, %d SMs, max clock %.0f MHz; ring of %d x 64 KiB stages, no computation\n", prop.name, sms, clock_khz / 1e3, kStages);

    struct { int a, g_lo; const char *what; } shapes[] = {
        {-1, 0, "contiguous tile, 512 rows x 128 B, 128B swizzle (final sweep)"},
        {3, 21, "strided a=3 g=[21,30): 512 rows x 128 B, stride 32 MiB"},
        {3, 12, "strided a=3 g=[12,21): 512 rows x 128 B, stride 64 KiB"},
        {4, 22, "strided a=4 g=[22,30): 256 rows x 256 B"},
        {5, 23, "strided a=5 g=[23,30): 128 rows x 512 B"},
        {7, 25, "strided a=7 g=[25,30): 32 rows x 2 KiB"},
    };
    for (auto &sh : shapes) {
        for (int stores = 1; stores >= 0; stores--) {
            params P = {};
            P.n_tiles = 1ull << (n - 12);
            P.mode = 0;
            P.stores = stores;
            P.amp = buf;
            CUtensorMap tmap;
            cuuint64_t dims[3], strides[2];
            cuuint32_t box[3], estr[3] = {1, 1, 1};
            CUtensorMapSwizzle swz;
            unsigned rows;
            if (sh.a >= 0) {
                const int g = 12 - sh.a, g_hi = sh.g_lo + g;
                rows = 1u << g;
                dims[0] = 2ull << sh.g_lo; dims[1] = rows; dims[2] = 1ull << (n - g_hi);
                strides[0] = 16ull << sh.g_lo; strides[1] = 16ull << g_hi;
                box[0] = 2u << sh.a;
                swz = CU_TENSOR_MAP_SWIZZLE_NONE;
                P.lo_gap = sh.g_lo - sh.a; P.a = sh.a;
            } else {
                rows = 512;
                dims[0] = 16; dims[1] = (1ull << n) >> 3; dims[2] = 1;
                strides[0] = 128; strides[1] = 128ull * dims[1];
                box[0] = 16;
                swz = CU_TENSOR_MAP_SWIZZLE_128B;
                P.lo_gap = -1;
            }
            P.box_rows = rows < 256u ? (int) rows : 256;
            P.n_boxes = (int) (rows / (unsigned) P.box_rows);
            P.box_bytes = kTileBytes / (uint32_t) P.n_boxes;
            box[1] = (cuuint32_t) P.box_rows; box[2] = 1;
            if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("encode failed for %s\n", sh.what);
                continue;
            }
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaEventRecord(e0));
                k_tma_copy<<<sms, 128, smem>>>(tmap, P);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const double bytes = (double) (16ull << n) * (stores ? 2.0 : 1.0);
            printf("%-66s %s: %7.3f ms = %7.1f GB/s = %5.1f B/clk/SM at max clock\n", sh.what, stores ? "load+store" : "load only ", best,
                   bytes / best / 1e6, bytes / (best * 1e-3) / sms / (clock_khz * 1e3));
        }
    }
    for (uint32_t piece : {65536u, 16384u, 2048u, 256u}) {
        for (int stores = 1; stores >= 0; stores--) {
            params P = {};
            P.n_tiles = 1ull << (n - 12);
            P.mode = 1; P.stores = stores; P.amp = buf; P.piece = piece;
            CUtensorMap dummy = {};
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaEventRecord(e0));
                k_tma_copy<<<sms, 128, smem>>>(dummy, P);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const double bytes = (double) (16ull << n) * (stores ? 2.0 : 1.0);
            printf("1-D bulk copies, contiguous 64 KiB tile in pieces of %6u B %26s %s: %7.3f ms = %7.1f GB/s = %5.1f B/clk/SM at max clock\n", piece, "",
                   stores ? "load+store" : "load only ", best, bytes / best / 1e6, bytes / (best * 1e-3) / sms / (clock_khz * 1e3));
        }
    }
    return 0;
}
