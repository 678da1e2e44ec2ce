// ubench_l2.cu -- how much of a two-phase sweep can live in the B200's L2?
//
// Decision data for the "L2-resident sweep pairing" plan of the fused QFT:
//   1. read bandwidth of a working set of W MiB that is re-read R times (L2 hits when it fits)
//   2. in-place read-modify-write of the same working set, R times
//   3. the blocked schedule itself on a 16 GiB array: for every block of W MiB, pass A
//      (read + write), grid barrier, pass B (read + write) with the chunks dealt to different
//      CTAs than in pass A -- against one and two plain streaming passes over the array.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/ubench_l2 tools/ubench_l2.cu
// Run under ncu with --metrics dram__bytes_read.sum,dram__bytes_write.sum to see what reaches HBM.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ double2 ldcg(const double2 *p)
{
    double a, b;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
    return make_double2(a, b);
}
__device__ __forceinline__ void stcg(double2 *p, double2 v)
{
    asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// chunk = 256 threads x 4 x 16 B = 16 KiB
constexpr int kChunkElems = 1024;

template <bool WRITE>
__device__ __forceinline__ double touch_chunk(double2 *base, uint64_t chunk, double acc)
{
    double2 *p = base + chunk * kChunkElems + threadIdx.x;
    double2 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = ldcg(p + k * 256);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (WRITE) {
            stcg(p + k * 256, make_double2(v[k].y, -v[k].x));
        } else {
            acc += v[k].x + v[k].y;
        }
    }
    return acc;
}

// R passes over a working set of n_chunks chunks; pass r deals chunk c to CTA (c + 37 r) mod grid
template <bool WRITE>
__global__ void __launch_bounds__(256) k_ws(double2 *buf, uint64_t n_chunks, int passes, double *sink)
{
    cg::grid_group grid = cg::this_grid();
    double acc = 0.0;
    for (int r = 0; r < passes; r++) {
        const uint64_t shift = (uint64_t) (37 * r) % gridDim.x;
        for (uint64_t c = (blockIdx.x + gridDim.x - shift) % gridDim.x; c < n_chunks; c += gridDim.x)
            acc = touch_chunk<WRITE>(buf, c, acc);
        grid.sync();
    }
    if (acc == 1.2345e-300) sink[0] = acc;
}

// the blocked two-phase schedule: every block of blk_chunks chunks gets pass A, barrier, pass B
__global__ void __launch_bounds__(256) k_blocked(double2 *buf, uint64_t n_blocks, uint64_t blk_chunks, int sync_each, double *sink)
{
    cg::grid_group grid = cg::this_grid();
    double acc = 0.0;
    for (uint64_t b = 0; b < n_blocks; b++) {
        double2 *base = buf + b * blk_chunks * kChunkElems;
        for (uint64_t c = blockIdx.x; c < blk_chunks; c += gridDim.x) acc = touch_chunk<true>(base, c, acc);
        if (sync_each) grid.sync();
        // pass B: strided deal, so a chunk is revisited by another SM
        for (uint64_t c = (blockIdx.x + 61) % gridDim.x; c < blk_chunks; c += gridDim.x) acc = touch_chunk<true>(base, c, acc);
        if (sync_each) grid.sync();
    }
    if (acc == 1.2345e-300) sink[0] = acc;
}

// software-pipelined variant without grid barriers: work items in the order A0, A1, B0, A2, B1, ...;
// a B item of block b waits until all A items of block b are done (a counter per block)
__global__ void __launch_bounds__(256) k_queue(double2 *buf, uint64_t n_blocks, uint64_t blk_chunks, unsigned long long *ticket,
                                               unsigned *done, double *sink)
{
    __shared__ unsigned long long s_item;
    double acc = 0.0;
    const uint64_t total = 2 * n_blocks * blk_chunks;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1ull);
        __syncthreads();
        const uint64_t item = s_item;
        if (item >= total) break;
        // segments of blk_chunks items: seg 0 = A0, seg 1 = A1, seg 2 = B0, seg 3 = A2, seg 4 = B1, ...
        const uint64_t seg = item / blk_chunks, c = item % blk_chunks;
        bool is_b;
        uint64_t b;
        if (seg == 0) { is_b = false; b = 0; }
        else if (seg == 2 * n_blocks - 1) { is_b = true; b = n_blocks - 1; }
        else if (seg & 1) { is_b = false; b = (seg + 1) / 2; }
        else { is_b = true; b = seg / 2 - 1; }
        double2 *base = buf + b * blk_chunks * kChunkElems;
        if (is_b) {
            if (threadIdx.x == 0) {
                while (atomicAdd(&done[b], 0u) < (unsigned) blk_chunks) __nanosleep(64);
            }
            __syncthreads();
            acc = touch_chunk<true>(base, (c * 61) % blk_chunks, acc);     // blk_chunks is a power of two: a permutation
        } else {
            acc = touch_chunk<true>(base, c, acc);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(&done[b], 1u);
        }
    }
    if (acc == 1.2345e-300) sink[0] = acc;
}


// ---- access-pattern probe: what does the memory system give a cheap kernel that moves the
// tiles of a strided sweep (2^rows_log2 rows of 2^a amplitudes, row stride 2^g_lo amplitudes)?
// One CTA = 256 threads holds a whole 2^12-amplitude tile in registers (16 x 16 B per thread).
__global__ void __launch_bounds__(256, 2) k_tile_rmw(double2 *buf, int n_bits, int a, int g_lo, int write)
{
    const int g = 12 - a;                       // stage bits (rows = 2^g)
    const int lo_gap = g_lo - a;
    const uint64_t n_tiles = 1ull << (n_bits - 12);
    double acc = 0.0;
    for (uint64_t tix = blockIdx.x; tix < n_tiles; tix += gridDim.x) {
        const uint64_t base = lo_gap > 0 ? (((tix >> lo_gap) << (g_lo + g)) | ((tix & ((1ull << lo_gap) - 1ull)) << a)) : (tix << 12);
        double2 v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const unsigned e = (unsigned) k * 256u + threadIdx.x;            // tile-local element
            const uint64_t off = (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
            v[k] = write == 2 ? make_double2((double) e, acc) : ldcg(buf + base + off);
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const unsigned e = (unsigned) k * 256u + threadIdx.x;
            const uint64_t off = (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
            if (write) stcg(buf + base + off, make_double2(v[k].y, -v[k].x));
            else acc += v[k].x + v[k].y;
        }
    }
    if (acc == 1.2345e-300) buf[0].x = acc;
}


// the same with a DRAM-friendly L2 prefetch one WAVE (2^wave_bits consecutive tiles) ahead: the
// tiles of a wave cover, per row, one contiguous run of 2^(wave_bits + a) amplitudes, which is
// bulk-prefetched into L2 (cp.async.bulk.prefetch.L2) by elected threads while the previous wave
// is being processed -- the tile loads themselves then hit L2.
__global__ void __launch_bounds__(256, 2) k_tile_rmw_pf(double2 *buf, int n_bits, int a, int g_lo, int wave_bits, int piece_log2)
{
    const int g = 12 - a;
    const int lo_gap = g_lo - a;
    const uint64_t n_tiles = 1ull << (n_bits - 12);
    const uint64_t wave = 1ull << wave_bits;
    const uint64_t n_waves = n_tiles >> wave_bits;
    const unsigned rows = 1u << g;
    // a wave's run per row: 2^(wave_bits + a) amplitudes (wave_bits <= lo_gap), cut into pieces
    const uint64_t run_bytes = 16ull << (wave_bits + a);
    const uint64_t piece_bytes = (1ull << piece_log2) < run_bytes ? (1ull << piece_log2) : run_bytes;
    const uint64_t pieces_per_row = run_bytes / piece_bytes;
    const uint64_t n_pieces = pieces_per_row * rows;
    for (uint64_t w = 0; w < n_waves; w++) {
        if (threadIdx.x == 0 && w + 1 < n_waves) {
            const uint64_t t0 = (w + 1) << wave_bits;     // first tile of the next wave
            const uint64_t base0 = ((t0 >> lo_gap) << (g_lo + g)) | ((t0 & ((1ull << lo_gap) - 1ull)) << a);
            for (uint64_t p = blockIdx.x; p < n_pieces; p += gridDim.x) {
                const uint64_t r = p / pieces_per_row, q = p % pieces_per_row;
                const char *ptr = (const char *) (buf + base0 + ((uint64_t) r << g_lo)) + q * piece_bytes;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"((unsigned) piece_bytes) : "memory");
            }
        }
        for (uint64_t tix = (w << wave_bits) + blockIdx.x; tix < ((w + 1) << wave_bits); tix += gridDim.x) {
            const uint64_t base = ((tix >> lo_gap) << (g_lo + g)) | ((tix & ((1ull << lo_gap) - 1ull)) << a);
            double2 v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const unsigned e = (unsigned) k * 256u + threadIdx.x;
                const uint64_t off = (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
                v[k] = ldcg(buf + base + off);
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const unsigned e = (unsigned) k * 256u + threadIdx.x;
                const uint64_t off = (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
                stcg(buf + base + off, make_double2(v[k].y, -v[k].x));
            }
        }
    }
}


// tile-order probe: which tiles a CTA takes one after the other.  run_log2 = r: a CTA takes 2^r
// tiles with consecutive numbers (adjacent 2^a-amplitude columns) back to back; hashed: the tile
// number is multiplied by an odd constant (spreads concurrently processed tiles over the address bits)
__global__ void __launch_bounds__(256, 2) k_tile_order(double2 *buf, int n_bits, int a, int g_lo, int mode, int run_log2, int hashed)
{
    const int g = 12 - a;
    const int lo_gap = g_lo - a;
    const uint64_t n_tiles = 1ull << (n_bits - 12);
    double acc = 0.0;
    for (uint64_t k = blockIdx.x; k < n_tiles; k += gridDim.x) {
        // k-th work item -> tile number
        uint64_t tix = ((k >> run_log2) / 1) ;
        tix = ((k / ((uint64_t) gridDim.x << run_log2)) * ((uint64_t) gridDim.x << run_log2)) +
              (((k % ((uint64_t) gridDim.x << run_log2)) % gridDim.x) << run_log2) + ((k % ((uint64_t) gridDim.x << run_log2)) / gridDim.x);
        if (tix >= n_tiles) continue;
        if (hashed) tix = (tix * 0x9E3779B1ull) & (n_tiles - 1);
        const uint64_t base = lo_gap > 0 ? (((tix >> lo_gap) << (g_lo + g)) | ((tix & ((1ull << lo_gap) - 1ull)) << a)) : (tix << 12);
        double2 v[16];
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            const unsigned e = (unsigned) kk * 256u + threadIdx.x;
            const uint64_t off = (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
            v[kk] = mode == 2 ? make_double2((double) e, acc) : ldcg(buf + base + off);
        }
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            const unsigned e = (unsigned) kk * 256u + threadIdx.x;
            const uint64_t off = (uint64_t) (e & ((1u << a) - 1u)) | ((uint64_t) (e >> a) << g_lo);
            if (mode) stcg(buf + base + off, make_double2(v[kk].y, -v[kk].x));
            else acc += v[kk].x + v[kk].y;
        }
    }
    if (acc == 1.2345e-300) buf[0].x = acc;
}

static float run_coop(const void *fn, int grid, void **args)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, 0));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main(int argc, char **argv)
{
    const int quick = argc > 1 ? atoi(argv[1]) : 0;      // 1: only the blocked runs (for ncu)
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int per_sm = 4;
    const int grid = sms * per_sm;
    const uint64_t total_bytes = 16ull << 30;
    double2 *buf;
    double *sink;
    unsigned long long *ticket;
    unsigned *done;
    CK(cudaMalloc(&buf, total_bytes));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMalloc(&ticket, 8));
    CK(cudaMalloc(&done, 4 * 65536));
    CK(cudaMemset(buf, 0, total_bytes));
    printf("device %s, %d SMs, L2 %d MiB, grid %d x 256\n", prop.name, sms, prop.l2CacheSize >> 20, grid);

    if (!quick) {
        const int sizes_mib[] = {8, 16, 24, 32, 40, 48, 64, 80, 96, 128, 256, 1024};
        for (int w : sizes_mib) {
            uint64_t n_chunks = ((uint64_t) w << 20) / (16 * kChunkElems);
            int passes = w <= 128 ? 200 : 20;
            for (int write = 0; write < 2; write++) {
                void *args[] = {&buf, &n_chunks, &passes, &sink};
                const void *fn = write ? (const void *) k_ws<true> : (const void *) k_ws<false>;
                run_coop(fn, grid, args);
                const float ms = run_coop(fn, grid, args);
                const double bytes = (double) w * 1048576.0 * passes * (write ? 2.0 : 1.0);
                printf("ws %4d MiB %s: %8.3f ms for %d passes = %8.1f GB/s (read%s)\n", w, write ? "rmw " : "read", ms, passes,
                       bytes / ms / 1e6, write ? "+write" : "");
            }
        }
        // plain streaming passes over 16 GiB
        for (int passes = 1; passes <= 2; passes++) {
            uint64_t n_chunks = total_bytes / (16 * kChunkElems);
            void *args[] = {&buf, &n_chunks, &passes, &sink};
            run_coop((const void *) k_ws<true>, grid, args);
            const float ms = run_coop((const void *) k_ws<true>, grid, args);
            printf("stream 16 GiB rmw x%d: %8.3f ms = %8.1f GB/s\n", passes, ms, 2.0 * total_bytes * passes / ms / 1e6);
        }
    }




    if (quick == 6) {
        // which property of the first sweep's tiles is slow: the number of rows (pages touched per
        // tile), the stride, or the 128 B row itself?  k_tile_rmw always moves 2^12-amplitude tiles:
        // rows = 2^(12 - a)
        struct { int a, g_lo; } pats[] = {{3, 12}, {3, 16}, {3, 17}, {3, 18}, {3, 19}, {3, 20}, {3, 21}, {4, 17}, {4, 18}, {4, 20},
                                          {4, 21}, {4, 22}, {5, 23}, {2, 20}, {2, 12}};
        const char *modes[] = {"read ", "rmw  ", "write"};
        for (auto &pt : pats) {
            for (int mode = 0; mode < 3; mode++) {
                cudaEvent_t a, b;
                CK(cudaEventCreate(&a));
                CK(cudaEventCreate(&b));
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(a));
                    k_tile_rmw<<<sms * 2, 256>>>(buf, 30, pt.a, pt.g_lo, mode);
                    CK(cudaEventRecord(b));
                    CK(cudaEventSynchronize(b));
                    CK(cudaGetLastError());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, a, b));
                    if (ms < best) best = ms;
                }
                const double bytes = (double) total_bytes * (mode == 1 ? 2.0 : 1.0);
                printf("rows %4d x %4d B, stride %8.0f KiB (%6.1f pages of 2 MiB) %s: %8.3f ms = %7.1f GB/s\n", 1 << (12 - pt.a), 16 << pt.a,
                       (double) (16ull << pt.g_lo) / 1024.0, (double) (16ull << pt.g_lo) / 2097152.0, modes[mode], best, bytes / best / 1e6);
            }
        }
        return 0;
    }
    if (quick == 5) {
        struct { int a, g_lo; const char *what; } pats[] = {
            {3, 21, "128 B rows, stride 32 MiB"},
            {3, 12, "128 B rows, stride 64 KiB"},
        };
        const char *modes[] = {"read ", "rmw  ", "write"};
        for (auto &pt : pats)
            for (int mode = 0; mode < 3; mode++)
                for (int hashed = 0; hashed < 2; hashed++)
                    for (int run = 0; run <= (hashed ? 0 : 3); run++) {
                        cudaEvent_t a, b;
                        CK(cudaEventCreate(&a));
                        CK(cudaEventCreate(&b));
                        float best = 1e30f;
                        for (int rep = 0; rep < 3; rep++) {
                            CK(cudaEventRecord(a));
                            k_tile_order<<<sms * 2, 256>>>(buf, 30, pt.a, pt.g_lo, mode, run, hashed);
                            CK(cudaEventRecord(b));
                            CK(cudaEventSynchronize(b));
                            CK(cudaGetLastError());
                            float ms;
                            CK(cudaEventElapsedTime(&ms, a, b));
                            if (ms < best) best = ms;
                        }
                        const double bytes = (double) total_bytes * (mode == 1 ? 2.0 : 1.0);
                        printf("order %s %-28s %s run 2^%d: %8.3f ms = %7.1f GB/s\n", modes[mode], pt.what, hashed ? "hashed" : "linear", run, best,
                               bytes / best / 1e6);
                    }
        return 0;
    }
    if (quick == 4) {
        struct { int n, a, g_lo; const char *what; } pats[] = {
            {30, 12, 12, "16 GiB, contiguous 64 KiB tiles"},
            {30, 3, 21, "16 GiB, 128 B rows, stride 32 MiB"},
            {30, 3, 12, "16 GiB, 128 B rows, stride 64 KiB"},
            {30, 4, 22, "16 GiB, 256 B rows, stride 64 MiB"},
            {20, 12, 12, "16 MiB (L2), contiguous 64 KiB tiles"},
            {20, 4, 12, "16 MiB (L2), 256 B rows, stride 64 KiB"},
            {20, 3, 11, "16 MiB (L2), 128 B rows, stride 32 KiB"},
            {20, 2, 10, "16 MiB (L2), 64 B rows, stride 16 KiB"},
            {21, 3, 12, "32 MiB (L2), 128 B rows, stride 64 KiB"},
        };
        const char *modes[] = {"read ", "rmw  ", "write"};
        for (auto &pt : pats) {
            for (int mode = 0; mode < 3; mode++) {
                const int reps = pt.n == 30 ? 1 : 200;
                cudaEvent_t a, b;
                CK(cudaEventCreate(&a));
                CK(cudaEventCreate(&b));
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(a));
                    for (int i = 0; i < reps; i++) k_tile_rmw<<<sms * 2, 256>>>(buf, pt.n, pt.a, pt.g_lo, mode);
                    CK(cudaEventRecord(b));
                    CK(cudaEventSynchronize(b));
                    CK(cudaGetLastError());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, a, b));
                    if (ms < best) best = ms;
                }
                const double bytes = (double) (16ull << pt.n) * reps * (mode == 1 ? 2.0 : 1.0);
                printf("tile %s %-44s: %8.3f ms = %7.1f GB/s\n", modes[mode], pt.what, best, bytes / best / 1e6);
            }
        }
        return 0;
    }
    if (quick == 2 || !quick) {
        struct { int a, g_lo; const char *what; } pats[] = {
            {12, 12, "contiguous 64 KiB tiles"},
            {3, 21, "128 B rows, stride 32 MiB (sweep 1 of n=30)"},
            {3, 12, "128 B rows, stride 64 KiB (sweep 2 of n=30)"},
            {4, 22, "256 B rows, stride 64 MiB"},
            {4, 14, "256 B rows, stride 256 KiB"},
            {5, 23, "512 B rows, stride 128 MiB"},
            {7, 25, "2 KiB rows, stride 512 MiB"},
        };
        for (auto &pt : pats) {
            for (int per = 1; per <= 2; per++) {
                cudaEvent_t a, b;
                CK(cudaEventCreate(&a));
                CK(cudaEventCreate(&b));
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(a));
                    k_tile_rmw<<<sms * per, 256>>>(buf, 30, pt.a, pt.g_lo, 1);
                    CK(cudaEventRecord(b));
                    CK(cudaEventSynchronize(b));
                    CK(cudaGetLastError());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, a, b));
                    if (ms < best) best = ms;
                }
                printf("tile rmw, %-46s %d CTA/SM: %7.3f ms = %7.1f GB/s\n", pt.what, per, best, 2.0 * total_bytes / best / 1e6);
            }
        }
    }
    if (quick == 2) return 0;

    if (quick == 3 || !quick) {
        struct { int a, g_lo; const char *what; } pats[] = {
            {3, 21, "128 B rows, stride 32 MiB (sweep 1 of n=30)"},
            {3, 12, "128 B rows, stride 64 KiB (sweep 2 of n=30)"},
            {2, 20, "64 B rows, stride 16 MiB"},
        };
        for (auto &pt : pats) {
            for (int wave_bits = 7; wave_bits <= 9; wave_bits++) {
                for (int piece_log2 = 11; piece_log2 <= 16; piece_log2 += 5) {
                    for (int per = 1; per <= 2; per++) {
                        cudaEvent_t a, b;
                        CK(cudaEventCreate(&a));
                        CK(cudaEventCreate(&b));
                        float best = 1e30f;
                        for (int rep = 0; rep < 3; rep++) {
                            CK(cudaEventRecord(a));
                            k_tile_rmw_pf<<<sms * per, 256>>>(buf, 30, pt.a, pt.g_lo, wave_bits, piece_log2);
                            CK(cudaEventRecord(b));
                            CK(cudaEventSynchronize(b));
                            CK(cudaGetLastError());
                            float ms;
                            CK(cudaEventElapsedTime(&ms, a, b));
                            if (ms < best) best = ms;
                        }
                        printf("tile rmw + L2 prefetch, %-44s wave 2^%d tiles (%3d MiB), pieces <= %2d KiB, %d CTA/SM: %7.3f ms = %7.1f GB/s\n",
                               pt.what, wave_bits, (int) ((65536ull << wave_bits) >> 20), (1 << piece_log2) >> 10, per, best,
                               2.0 * total_bytes / best / 1e6);
                    }
                }
            }
        }
    }
    if (quick == 3) return 0;
    const int blocks_mib[] = {8, 16, 32, 64};
    for (int w : blocks_mib) {
        uint64_t blk_chunks = ((uint64_t) w << 20) / (16 * kChunkElems);
        uint64_t n_blocks = total_bytes / ((uint64_t) w << 20);
        for (int sync_each = 1; sync_each >= 1; sync_each--) {
            void *args[] = {&buf, &n_blocks, &blk_chunks, &sync_each, &sink};
            run_coop((const void *) k_blocked, grid, args);
            const float ms = run_coop((const void *) k_blocked, grid, args);
            printf("blocked A|B, block %3d MiB, grid.sync: %8.3f ms for 2 logical passes over 16 GiB (one streaming pass at 6.5 TB/s = 5.29 ms)\n",
                   w, ms);
        }
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(ticket, 0, 8));
            CK(cudaMemset(done, 0, 4 * 65536));
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            CK(cudaEventRecord(a));
            k_queue<<<grid, 256>>>(buf, n_blocks, blk_chunks, ticket, done, sink);
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (rep) printf("queued  A|B, block %3d MiB, no barrier: %8.3f ms for 2 logical passes over 16 GiB\n", w, ms);
        }
    }
    CK(cudaFree(buf));
    return 0;
}
