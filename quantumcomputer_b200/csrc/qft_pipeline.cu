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

// One sweep of a launch: its descriptor and how its tiles map to TMA boxes.
struct phase_params {
    sweep_desc d;
    int lo_gap;             // g_lo - a (strided) or -1 (contiguous final sweep)
    int n_boxes;            // TMA boxes per tile (a box has at most 256 rows)
    int box_rows;
    uint32_t box_bytes;
    int wcol_base;          // where this phase's column-twiddle tables start in the kernel's table area
    int inb_pos, inb_bits;  // paired launch: the tile-number bits [inb_pos, inb_pos + inb_bits) count the tiles
                            // inside one block, the others (in order) number the block
};

// A launch runs one sweep (tiles dealt round robin over the CTAs) or an L2-PAIRED couple of
// consecutive sweeps A -> B whose tiles have the same low run: the amplitudes fall into BLOCKS that
// are closed under both sweeps (strided sweep + contiguous final sweep: 2^g_hi consecutive
// amplitudes; two strided sweeps with stage bits [l, m) and [m, h): the low run x bits [l, h)), a few
// MiB each, with as many tiles of A as of B.  The tiles of both
// sweeps are handed out through one ticket queue in the order
//     A[0 .. lag),  then alternating  A[lag + i], B[i],  then the last B's,
// (lag >= one block of tiles) so that B trails A by a little more than one block: what A wrote is
// still in the 126 MB L2 when B reads it, and A's lines are overwritten by B's before they are ever
// evicted -- the intermediate state never travels to HBM.  A B tile may only be loaded when every A
// tile of its block has been stored: a counter per block, incremented by the store warp once its
// bulk store has completed, polled by the producer.  A waiting tile depends only on tiles with
// smaller tickets, which are held by running CTAs: no deadlock whatever the residency.
struct pipe_params {
    phase_params ph[2];
    int n_phases;
    uint64_t n_tiles;               // tiles per phase
    int blk_tile_bits;              // pair: log2(tiles per block)
    uint64_t lag;                   // pair: B trails A by this many tiles
    int l2_hints;                   // pair: eviction-priority hints on the TMA loads and stores
    unsigned long long *ticket;     // pair: the queue head
    unsigned *done;                 // pair: stored A tiles per block
    unsigned long long *timing;     // -DQCS_PIPE_TIMING builds only: per-CTA cycle counters
    // GEN kernels (quantum_computation from the reset state): the tiles are not loaded but built from
    // f(x) -- amplitude gen_value at (x, f(x)) of every block x, zero elsewhere (modexp_fused.cu)
    const unsigned *gen_f;
    unsigned gen_M;
    double gen_value;
};

constexpr uint64_t kNoItem = ~0ull;
constexpr uint64_t kPhaseB = 1ull << 62;

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
// the same with an L2 eviction-priority hint (paired launches: what the second sweep will read is kept,
// what nobody reads again leaves first)
__device__ __forceinline__ uint64_t l2_policy(bool keep)
{
    uint64_t pol;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap *map, const void *src, int c0, int c1, int c2, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
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
// the same barrier, OR-reducing a predicate over the group
__device__ __forceinline__ bool group_barrier_or(int group, int threads, bool pred)
{
    unsigned r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %3, 0;\n\tbar.red.or.pred q, %1, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r)
                 : "r"(group + 1), "r"(threads), "r"((unsigned) pred)
                 : "memory");
    return r != 0;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int TB>
__device__ __forceinline__ void tile_coords(const phase_params &Q, uint64_t tix, int &c0, int &c1, int &c2)
{
    if (Q.lo_gap >= 0) {
        c0 = (int) ((tix & ((1ull << Q.lo_gap) - 1ull)) << (Q.d.a + 1));   // doubles
        c1 = 0;
        c2 = (int) (tix >> Q.lo_gap);
    } else {
        c0 = 0;
        c1 = (int) (tix << (Q.d.sw == kSwizzleSplit3 ? TB - 4 : TB - 3));     // rows of 8 amplitudes (split-3 view: pairs of rows)
        c2 = 0;
    }
}

// first amplitude of tile tix
template <int TB>
__host__ __device__ __forceinline__ uint64_t tile_base(const phase_params &Q, uint64_t tix)
{
    return Q.lo_gap >= 0 ? (((tix >> Q.lo_gap) << Q.d.g_hi) | ((tix & ((1ull << Q.lo_gap) - 1ull)) << Q.d.a)) : (tix << TB);
}

// item of a paired launch (block-major: block * tiles_per_block + tile in block) -> tile number
__host__ __device__ __forceinline__ uint64_t pair_tile(const pipe_params &P, const phase_params &Q, uint64_t idx)
{
    const uint64_t inb = idx & ((1ull << P.blk_tile_bits) - 1ull), blk = idx >> P.blk_tile_bits;
    return ((blk >> Q.inb_pos) << (Q.inb_pos + Q.inb_bits)) | (inb << Q.inb_pos) | (blk & ((1ull << Q.inb_pos) - 1ull));
}

// the t-th ticket of a paired launch -> phase bit | item within the phase
__host__ __device__ __forceinline__ uint64_t pair_item(const pipe_params &P, uint64_t t)
{
    const uint64_t na = P.n_tiles, lag = P.lag;
    if (t >= 2 * na) return kNoItem;
    if (t < lag) return t;
    const uint64_t v = t - lag, pairs = na - lag;
    if (v < 2 * pairs) return (v & 1ull) ? (kPhaseB | (v >> 1)) : (lag + (v >> 1));
    return kPhaseB | (pairs + (v - 2 * pairs));
}

// MODE: what the stages are -- kWalsh: bare Hadamards (+ diagonal gates riding along), kInverse: the
// reference's inverse_QFT stages, kForward: their adjoint.  One instantiation per mode keeps the code of
// a launch small (the instruction cache holds about 40 KiB).
enum { kWalsh = 0, kInverse = 1, kForward = 2 };

// GEN: the (first) sweep's input is the state |x, f(x)> quantum_computation builds from the reset state.  Nothing
// is loaded for it: the consumer group zeroes the stage, sets the amplitudes of the blocks x whose f(x) falls into
// the tile's rows, and skips the steps of a tile that holds none (its output is zero as well).
template <int TB, int STAGES, int GROUPS, int GT, int MODE, bool PAIRED, bool GEN = false>
__global__ void __launch_bounds__(64 + GROUPS * GT, 1)
k_qft_sweep_tma(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1, const pipe_params P)
{
    static_assert(STAGES % GROUPS == 0, "a ring stage must always be consumed by the same group");
    constexpr int kThreads = 64 + GROUPS * GT;
    constexpr uint32_t kTileBytes = 16u << TB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // STAGES tiles; the 128-byte swizzle of the final sweep needs 1024-byte alignment
    double2 *stage_buf = (double2 *) (smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    double2 *wcol = stage_buf + (size_t) STAGES * (1u << TB);
    const int wcol_all = P.ph[0].d.wcol_total + (PAIRED ? P.ph[1].d.wcol_total : 0);
    double2 *wbase = wcol + wcol_all;                                       // [STAGES][kMaxSteps]: per-tile twiddle bases
    uint64_t *bars = (uint64_t *) (wbase + STAGES * kMaxSteps);
    uint64_t *full = bars, *computed = bars + STAGES, *empty = bars + 2 * STAGES;
    uint64_t *s_item = bars + 3 * STAGES;                                   // [STAGES]: what the stage holds
    diag_gate *sdiag = (diag_gate *) (s_item + STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool paired = PAIRED;       // compile time: a single sweep's descriptor then sits at fixed constant-bank offsets

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 2);          // the producer's expect_tx and its arrive once the tile's bases are written
            mbar_init(&computed[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap0) : "memory");
        if (paired) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap1) : "memory");
    }
    // per-column part of the external twiddle: fixed for the whole kernel; one entry per value of
    // the tile-local bits below the step
    for (int ph = 0; ph < (PAIRED ? 2 : 1); ph++) {
        const phase_params &Q = P.ph[ph];
        if (Q.d.hadamard_only) continue;
        tile_geom G;
        G.a = Q.d.a;
        G.g_lo = Q.d.g_lo;
        G.sw = Q.d.sw;
        for (int k = 0; k < Q.d.n_steps; k++) {
            const sweep_step S = Q.d.step[k];
            if (S.notw) continue;
            for (unsigned c = threadIdx.x; c < (1u << S.s); c += kThreads) {
                const uint64_t y = (G.spread(c) & ((1ull << S.low_phys) - 1ull)) >> Q.d.lo;
                wcol[Q.wcol_base + S.col_off + c] = unit_phase(y, S.j, Q.d.inverse != 0);
            }
        }
    }
    // diagonal gates riding along (Walsh-Hadamard sweeps of the gate stream): phase 0's list, then phase 1's
    for (int i = threadIdx.x; i < P.ph[0].d.n_diag; i += kThreads) sdiag[i] = P.ph[0].d.diag[i];
    if (PAIRED)
        for (int i = threadIdx.x; i < P.ph[1].d.n_diag; i += kThreads) sdiag[P.ph[0].d.n_diag + i] = P.ph[1].d.diag[i];
    __syncthreads();

    if (warp == 0) {
        // ---------------- producer: the whole warp walks the items; lane 0 owns barriers and TMA,
        // lanes < n_steps compute the tile's twiddle bases ----------------
        uint64_t next = kNoItem;
        if (paired && lane == 0) next = pair_item(P, atomicAdd(P.ticket, 1ull));
        const uint64_t pol_stream = l2_policy(false);
        int sentinels = 0;
        for (uint64_t k = 0;; k++) {
            const int s = (int) (k % STAGES);
            const uint32_t round = (uint32_t) (k / STAGES);
            uint64_t item = kNoItem;
            if (lane == 0) {
                const long long t0 = QCS_TICK(P);
                mbar_wait(&empty[s], (round & 1u) ^ 1u);
                QCS_TIMING_ADD(P, 0, QCS_TICK(P) - t0);
                if (paired) {
                    item = next;
                } else {
                    const uint64_t idx = blockIdx.x + k * gridDim.x;
                    item = idx < P.n_tiles ? idx : kNoItem;
                }
            }
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item == kNoItem) {
                // every group finds one sentinel at its next turn and leaves
                if (lane == 0) {
                    s_item[s] = kNoItem;
                    mbar_arrive(&full[s]);
                    mbar_arrive(&full[s]);
                }
                if (++sentinels == GROUPS) break;
                continue;
            }
            const int ph = (PAIRED && (item & kPhaseB)) ? 1 : 0;
            const phase_params &Q = PAIRED ? P.ph[ph] : P.ph[0];
            const uint64_t idx = item & (kPhaseB - 1ull);
            const uint64_t tix = paired ? pair_tile(P, Q, idx) : tile_number(Q.d, idx);
            // the loads go out first; the tile's twiddle bases are computed while they travel
            if (lane == 0) {
                if (ph == 1) {
                    // all A tiles of the block must be in L2 / memory
                    const unsigned *flag = P.done + (idx >> P.blk_tile_bits);
                    const unsigned need = 1u << P.blk_tile_bits;
                    const long long tw = QCS_TICK(P);
                    while (ld_acquire_gpu(flag) < need) __nanosleep(200);
                    QCS_TIMING_ADD(P, 10, QCS_TICK(P) - tw);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                int c0, c1, c2;
                tile_coords<TB>(Q, tix, c0, c1, c2);
                const bool generated = GEN && ph == 0;
                if (generated) mbar_arrive(&full[s]);      // nothing to wait for: the stage is free, the group fills it
                else mbar_expect_tx(&full[s], kTileBytes);
                unsigned char *dst = (unsigned char *) (stage_buf + (size_t) s * (1u << TB));
                const CUtensorMap *map = ph ? &tmap1 : &tmap0;
                if (generated) {
                } else if (P.l2_hints) {
                    // A's input and B's input (A's output, read for the last time) may leave the L2 first
                    for (int b = 0; b < Q.n_boxes; b++)
                        tma_load_3d_hint(dst + (size_t) b * Q.box_bytes, map, &full[s], c0, c1 + b * Q.box_rows, c2, pol_stream);
                } else {
                    for (int b = 0; b < Q.n_boxes; b++)
                        tma_load_3d(dst + (size_t) b * Q.box_bytes, map, &full[s], c0, c1 + b * Q.box_rows, c2);
                }
                // the next ticket travels while this tile loads
                if (paired) next = pair_item(P, atomicAdd(P.ticket, 1ull));
            }
            if (!Q.d.hadamard_only && lane < Q.d.n_steps) {
                const sweep_step S = Q.d.step[lane];
                const uint64_t base = tile_base<TB>(Q, tix);
                uint64_t y = 0;
                if (S.low_phys > Q.d.lo) y = (base & ((1ull << S.low_phys) - 1ull)) >> Q.d.lo;
                wbase[s * kMaxSteps + lane] = unit_phase(y + Q.d.y_const, S.j, Q.d.inverse != 0);
            }
            __syncwarp();
            if (lane == 0) {
                s_item[s] = item;
                mbar_arrive(&full[s]);           // second arrival: what the consumers read besides the tile is in place
            }
        }
    } else if (warp == 1) {
        // ---------------- store issuer ----------------
        if (lane == 0) {
            const uint64_t pol_keep = l2_policy(true), pol_stream = l2_policy(false);
            for (uint64_t k = 0;; k++) {
                const int s = (int) (k % STAGES);
                const uint32_t round = (uint32_t) (k / STAGES);
                const long long t0 = QCS_TICK(P);
                mbar_wait(&computed[s], round & 1u);
                const long long t1 = QCS_TICK(P);
                const uint64_t item = s_item[s];
                if (item == kNoItem) break;
                const int ph = (PAIRED && (item & kPhaseB)) ? 1 : 0;
                const phase_params &Q = PAIRED ? P.ph[ph] : P.ph[0];
                const uint64_t idx = item & (kPhaseB - 1ull);
                int c0, c1, c2;
                tile_coords<TB>(Q, paired ? pair_tile(P, Q, idx) : tile_number(Q.d, idx), c0, c1, c2);
                const unsigned char *src = (const unsigned char *) (stage_buf + (size_t) s * (1u << TB));
                const CUtensorMap *map = ph ? &tmap1 : &tmap0;
                if (P.l2_hints) {
                    // A's output stays in the L2 until B has read it; B's output (and a single sweep's) is final
                    const uint64_t pol = (PAIRED && ph == 0) ? pol_keep : pol_stream;
                    for (int b = 0; b < Q.n_boxes; b++)
                        tma_store_3d_hint(map, src + (size_t) b * Q.box_bytes, c0, c1 + b * Q.box_rows, c2, pol);
                } else {
                    for (int b = 0; b < Q.n_boxes; b++)
                        tma_store_3d(map, src + (size_t) b * Q.box_bytes, c0, c1 + b * Q.box_rows, c2);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                QCS_TIMING_ADD(P, 1, t1 - t0);               // waiting for a computed tile
                QCS_TIMING_ADD(P, 2, QCS_TICK(P) - t1);      // the store reading shared memory
                mbar_arrive(&empty[s]);
                if (paired && ph == 0) {
                    // publish the tile to the B tiles of its block as soon as it has been written.  (Not
                    // deferred to the next store: the next tile of this CTA may be a B tile waiting for it.)
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                    atomicAdd(P.done + (idx >> P.blk_tile_bits), 1u);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // ---------------- consumers ----------------
        const int group = (warp - 2) / (GT / 32);
        const unsigned tig = threadIdx.x - 64 - group * GT;
        uint64_t n_done = 0;
        for (uint64_t k = group;; k += GROUPS) {
            const int s = (int) (k % STAGES);
            const uint32_t round = (uint32_t) (k / STAGES);
            double2 *tile = stage_buf + (size_t) s * (1u << TB);
            const long long t0 = QCS_TICK(P);
            // every thread waits on the barrier itself: electing one lane per warp to poll (the others at
            // __syncwarp) was measured 10 % slower (24.05 against 21.81 ms at n = 30)
            mbar_wait(&full[s], round & 1u);
            const long long t1 = QCS_TICK(P);
            const uint64_t item = s_item[s];
            if (item == kNoItem) {
                if (tig == 0) mbar_arrive(&computed[s]);     // lets the store issuer see the sentinel
                break;
            }
            const int ph = (PAIRED && (item & kPhaseB)) ? 1 : 0;
            const phase_params &Q = PAIRED ? P.ph[ph] : P.ph[0];
            const uint64_t tix = paired ? pair_tile(P, Q, item & (kPhaseB - 1ull)) : tile_number(Q.d, item & (kPhaseB - 1ull));
            const uint64_t base = tile_base<TB>(Q, tix);
            tile_geom G;
            G.a = Q.d.a;
            G.g_lo = Q.d.g_lo;
            G.sw = Q.d.sw;
            const double2 *my_wcol = wcol + Q.wcol_base;
            const diag_gate *my_diag = sdiag + (ph ? P.ph[0].d.n_diag : 0);
            bool live = true;
            if (GEN && ph == 0) {
                // rows j of the tile = blocks x_base | j << (g_lo - M); columns = the f values f_hi << a | [0, 2^a)
                const unsigned a = (unsigned) Q.d.a, M = P.gen_M, rows = 1u << (TB - Q.d.a);
                const unsigned f_hi = ((unsigned) base & ((1u << M) - 1u)) >> a;
                const uint64_t x_base = base >> M;
                const int x_shift = Q.d.g_lo - (int) M;
                unsigned fx[4];                              // the table reads travel while the stage is zeroed
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned j = tig + (unsigned) u * GT;
                    fx[u] = j < rows ? P.gen_f[x_base | ((uint64_t) j << x_shift)] : 0xffffffffu;
                }
                for (unsigned e = tig; e < (1u << TB); e += GT) tile[e] = make_double2(0.0, 0.0);
                group_barrier(group, GT);
                bool hit = false;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned j = tig + (unsigned) u * GT;
                    if (j < rows && (fx[u] >> a) == f_hi) { tile[(j << a) | (fx[u] & ((1u << a) - 1u))] = make_double2(P.gen_value, 0.0); hit = true; }
                }
                for (unsigned j = tig + 4u * GT; j < rows; j += GT) {
                    const unsigned f = P.gen_f[x_base | ((uint64_t) j << x_shift)];
                    if ((f >> a) == f_hi) { tile[(j << a) | (f & ((1u << a) - 1u))] = make_double2(P.gen_value, 0.0); hit = true; }
                }
                live = group_barrier_or(group, GT, hit);
                if (!live) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zeroes -> visible to the TMA store
                    group_barrier(group, GT);
                }
            }
            for (int st = 0; live && st < Q.d.n_steps; st++) {
                const sweep_step S = Q.d.step[st];
                const bool last = st == Q.d.n_steps - 1;
                const double2 wb = wbase[s * kMaxSteps + st];
                // strided tiles lie linearly in shared memory, the contiguous one carries the 128-byte swizzle
                if (Q.lo_gap >= 0) {
                    if (MODE == kWalsh) dispatch_step<true, false, true>(nullptr, tile, my_wcol + S.col_off, wb, G, S, TB, base, false, false, last, Q.d.scale, tig, GT, my_diag, Q.d.n_diag, Q.d.index_or);
                    else dispatch_step<MODE == kInverse, true, true>(nullptr, tile, my_wcol + S.col_off, wb, G, S, TB, base, false, false, last, Q.d.scale, tig, GT);
                } else if (TB == 12 && Q.d.sw == kSwizzleSplit3) {
                    // contiguous tile in the split-3 layout: radix-16 steps at bits 8, 4, 0
                    if (MODE == kWalsh) dispatch_split3<true, false>(tile, my_wcol + S.col_off, wb, G, S, base, last, Q.d.scale, tig, GT, my_diag, Q.d.n_diag, Q.d.index_or);
                    else dispatch_split3<MODE == kInverse, true>(tile, my_wcol + S.col_off, wb, G, S, base, last, Q.d.scale, tig, GT);
                } else {
                    if (MODE == kWalsh) dispatch_step<true, false, false>(nullptr, tile, my_wcol + S.col_off, wb, G, S, TB, base, false, false, last, Q.d.scale, tig, GT, my_diag, Q.d.n_diag, Q.d.index_or);
                    else dispatch_step<MODE == kInverse, true, false>(nullptr, tile, my_wcol + S.col_off, wb, G, S, TB, base, false, false, last, Q.d.scale, tig, GT);
                }
                if (last) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to TMA
                group_barrier(group, GT);
            }
            n_done++;
            if (tig == 0) {
                mbar_arrive(&computed[s]);
                QCS_TIMING_ADD(P, 4 + 2 * group, t1 - t0);              // waiting for the load
                QCS_TIMING_ADD(P, 5 + 2 * group, QCS_TICK(P) - t1);     // the steps
            }
        }
        if (tig == 0) QCS_TIMING_ADD(P, 3, n_done);
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

const pipe_shape &shape_of(const qcs_register *reg)
{
    return kShapes[reg->opt_pipe_shape >= 0 && reg->opt_pipe_shape < kNumShapes ? reg->opt_pipe_shape : 0];
}

size_t pipe_smem(const pipe_shape &sh, int wcol_entries, int n_diag)
{
    return (size_t) sh.stages * ((size_t) 16 << sh.tb) + 16 * (size_t) wcol_entries + 16 * (size_t) sh.stages * kMaxSteps +
           8 * 4 * (size_t) sh.stages + sizeof(diag_gate) * (size_t) n_diag + 1024;
}

template <int TB, int STAGES, int GROUPS, int GT, int MODE>
int launch_mode(qcs_register *reg, const CUtensorMap &tmap0, const CUtensorMap &tmap1, const pipe_params &P, size_t smem,
                const qft::sweep_target &tg)
{
    auto kern = P.n_phases > 1 ? k_qft_sweep_tma<TB, STAGES, GROUPS, GT, MODE, true> : k_qft_sweep_tma<TB, STAGES, GROUPS, GT, MODE, false>;
    if (P.gen_f) {
        if (MODE != kInverse) return QCS_BAD_ARGUMENTS;
        kern = P.n_phases > 1 ? k_qft_sweep_tma<TB, STAGES, GROUPS, GT, kInverse, true, true>
                              : k_qft_sweep_tma<TB, STAGES, GROUPS, GT, kInverse, false, true>;
    }
    QCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    uint64_t grid = (uint64_t) reg->sm_count;
    if (tg.max_ctas > 0 && grid > (uint64_t) tg.max_ctas) grid = (uint64_t) tg.max_ctas;
    if (grid > P.n_tiles) grid = P.n_tiles;
    // algorithmic bytes = sweeps made x 32 B per amplitude; a paired launch makes two sweeps
    // (its DRAM traffic is lower: the intermediate state stays in the L2)
    // (a generating sweep only writes)
    qcs_launch_begin(reg, tg.kind, tg.bytes > 0.0 ? tg.bytes : (32.0 * (double) P.n_phases - (P.gen_f ? 16.0 : 0.0)) * (double) (P.n_tiles << TB));
    kern<<<(unsigned) grid, 64 + GROUPS * GT, smem, tg.stream>>>(tmap0, tmap1, P);
    return qcs_launch_end(reg, tg.kind, "k_qft_sweep_tma");
}

template <int TB, int STAGES, int GROUPS, int GT>
int launch_shape(qcs_register *reg, const CUtensorMap &tmap0, const CUtensorMap &tmap1, const pipe_params &P, size_t smem,
                 const qft::sweep_target &tg)
{
    // both sweeps of a paired launch belong to the same transform: one mode
    if (P.ph[0].d.hadamard_only) return launch_mode<TB, STAGES, GROUPS, GT, kWalsh>(reg, tmap0, tmap1, P, smem, tg);
    if (P.ph[0].d.inverse) return launch_mode<TB, STAGES, GROUPS, GT, kInverse>(reg, tmap0, tmap1, P, smem, tg);
    return launch_mode<TB, STAGES, GROUPS, GT, kForward>(reg, tmap0, tmap1, P, smem, tg);
}

int launch_by_shape(qcs_register *reg, int shape_id, const CUtensorMap &tmap0, const CUtensorMap &tmap1, const pipe_params &P,
                    size_t smem, const qft::sweep_target &tg)
{
    switch (shape_id) {
        case 1: return launch_shape<11, 6, 3, 128>(reg, tmap0, tmap1, P, smem, tg);
        case 2: return launch_shape<11, 6, 6, 64>(reg, tmap0, tmap1, P, smem, tg);
        case 3: return launch_shape<12, 3, 1, 256>(reg, tmap0, tmap1, P, smem, tg);
        case 4: return launch_shape<12, 3, 3, 64>(reg, tmap0, tmap1, P, smem, tg);
        case 5: return launch_shape<11, 6, 2, 128>(reg, tmap0, tmap1, P, smem, tg);
        default: return launch_shape<12, 3, 3, 128>(reg, tmap0, tmap1, P, smem, tg);
    }
}

// tensor map + box geometry of one sweep
int encode_phase(const pipe_shape &sh, const qft::sweep_target &tg, const qft::sweep_plan &plan, phase_params &Q, CUtensorMap &tmap,
                 bool allow_split3)
{
    encode_fn_t encode = get_encode();
    if (!encode) {
        fprintf(stderr, "qcs: cuTensorMapEncodeTiled is unavailable\n");
        return QCS_UNKNOWN_ERROR;
    }
    Q.d = plan.d;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    CUtensorMapSwizzle swz;
    const bool strided = plan.d.g_lo > plan.d.a;
    unsigned rows = 0;
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
        Q.lo_gap = plan.d.g_lo - plan.d.a;
        Q.d.sw = 28;                                         // no XOR
    } else {
        // contiguous tile, 128 B rows, 128-byte hardware swizzle.  Two views of the same 2^tb amplitudes:
        //   plain   : rows e >> 3                              -> phys(e) = e ^ ((e >> 3) & 7)
        //   split-3 : {8 amplitudes} x {e >> 4} x {bit 3 of e} -> split3_phys(e): conflict-free for every
        //             radix-16 step; taken when the sweep consists of radix-16 steps at bits 8, 4, 0 of a 2^12
        //             tile (the kernel has compile-time addressing for exactly those)
        Q.lo_gap = -1;
        swz = CU_TENSOR_MAP_SWIZZLE_128B;
        dims[0] = 16;
        box[0] = 16;
        bool split = allow_split3 && sh.tb == 12 && plan.d.n_steps >= 1;
        for (int k = 0; k < plan.d.n_steps && split; k++)
            split = plan.d.step[k].r == 4 && (plan.d.step[k].s == 8 || plan.d.step[k].s == 4 || plan.d.step[k].s == 0);
        if (split) {
            dims[1] = (1ull << tg.n_bits) >> 4;
            dims[2] = 2;
            strides[0] = 256;
            strides[1] = 128;
            Q.d.sw = kSwizzleSplit3;
            Q.box_rows = 1 << (sh.tb - 4);
            Q.n_boxes = 1;
            Q.box_bytes = 16u << sh.tb;
            box[1] = (cuuint32_t) Q.box_rows;
            box[2] = 2;
        } else {
            rows = 1u << (sh.tb - 3);
            dims[1] = (1ull << tg.n_bits) >> 3;
            dims[2] = 1;
            strides[0] = 128;
            strides[1] = 128ull * dims[1];
            Q.d.sw = 3;
        }
    }
    if (Q.lo_gap >= 0 || Q.d.sw == 3) {
        Q.box_rows = rows < 256u ? (int) rows : 256;         // a box dimension is at most 256
        Q.n_boxes = (int) (rows / (unsigned) Q.box_rows);
        Q.box_bytes = (uint32_t) (((size_t) 16 << sh.tb) / (size_t) Q.n_boxes);
        box[1] = (cuuint32_t) Q.box_rows;
        box[2] = 1;
    }
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
    return QCS_NO_ERROR;
}

// -DQCS_PIPE_TIMING builds: QCS_PIPE_TIMING=1 prints where the roles of the pipeline spend their cycles
int run_launch(qcs_register *reg, const pipe_shape &sh, int shape_id, const CUtensorMap &tmap0, const CUtensorMap &tmap1,
               pipe_params &P, size_t smem, const qft::sweep_target &tg)
{
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
    const int rc = launch_by_shape(reg, shape_id, tmap0, tmap1, P, smem, tg);
    if (timing_on && rc == QCS_NO_ERROR) {
        std::vector<unsigned long long> h(16 * (size_t) reg->sm_count);
        QCS_CUDA(cudaMemcpyAsync(h.data(), P.timing, h.size() * 8, cudaMemcpyDeviceToHost, tg.stream));
        QCS_CUDA(cudaStreamSynchronize(tg.stream));
        double sum[16] = {0};
        for (size_t i = 0; i < h.size(); i++) sum[i % 16] += (double) h[i];
        const double tiles = sum[3] > 0 ? sum[3] : 1;
        const sweep_desc &d = P.ph[0].d;
        fprintf(stderr, "qcs pipe timing (cycles per tile, avg over CTAs) shape %d phases %d a=%d g=[%d,%d) steps=%d: producer wait empty %.0f, "
                "wait dependency %.0f | store wait computed %.0f, store read %.0f |", shape_id, P.n_phases, d.a, d.g_lo, d.g_hi,
                d.n_steps, sum[0] / tiles, sum[10] / tiles, sum[1] / tiles, sum[2] / tiles);
        for (int g = 0; g < sh.groups && g < 3; g++)
            fprintf(stderr, " g%d wait load %.0f steps %.0f;", g, sum[4 + 2 * g] / tiles * sh.groups, sum[5 + 2 * g] / tiles * sh.groups);
        fprintf(stderr, "\n");
        cudaFree(P.timing);
    }
    return rc;
}

}  // namespace

int qcs_pipeline_tile_bits(const qcs_register *reg) { return shape_of(reg).tb; }

// true when the pipelined kernel can run this sweep
bool qcs_pipeline_supports(const qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &p)
{
    const pipe_shape sh = shape_of(reg);
    if (p.d.t != sh.tb || tg.n_bits < (unsigned) sh.tb + 3) return false;
    const bool strided = p.d.g_lo > p.d.a;
    // a TMA box row is 2^(a+1) doubles (32 B .. 2 KiB); coordinates are 32-bit
    if (strided && (p.d.a < 1 || p.d.a > 7 || p.d.g_lo + 1 > 31)) return false;
    // the linear layout TMA writes for a strided tile must make every step conflict-free; the
    // contiguous sweep takes the 128-byte hardware swizzle (conflict-free for every step except a
    // radix-16 step on bits 0..3, which is 2-way conflicted)
    if (strided && conflict_cost(p.d, 28, true) != 0) return false;
    return pipe_smem(sh, p.d.wcol_total, p.d.n_diag) <= reg->smem_optin;
}

// quantum_computation from the reset state (qcs_shor_state_generated) has armed the generator: this launch's
// first sweep is the first sweep of its inverse QFT; qcs_fused_gen_supported has checked the geometry
static int take_generator(qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &first, pipe_params &P)
{
    if (!reg->gen.armed) return QCS_NO_ERROR;
    const bool strided = first.d.g_lo > first.d.a;
    if (!strided || !first.d.inverse || first.d.hadamard_only || first.d.a > (int) reg->gen.M || first.d.g_lo < (int) reg->gen.M ||
        tg.amp != reg->amp)
        return QCS_BAD_ARGUMENTS;
    P.gen_f = reg->gen.f;
    P.gen_M = reg->gen.M;
    P.gen_value = reg->gen.value;
    reg->gen.armed = 0;
    return QCS_NO_ERROR;
}

int qcs_pipeline_launch(qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &plan)
{
    const pipe_shape sh = shape_of(reg);
    const int shape_id = reg->opt_pipe_shape >= 0 && reg->opt_pipe_shape < kNumShapes ? reg->opt_pipe_shape : 0;
    pipe_params P = {};
    CUtensorMap tmap;
    QCS_TRY(encode_phase(sh, tg, plan, P.ph[0], tmap, reg->opt_split3 != 0));
    P.ph[0].wcol_base = 0;
    P.n_phases = 1;
    P.n_tiles = plan.n_tiles;
    P.l2_hints = 0;              // evict_first on a single sweep's traffic was measured: no effect (21.6 ms either way)
    QCS_TRY(take_generator(reg, tg, plan, P));
    return run_launch(reg, sh, shape_id, tmap, tmap, P, pipe_smem(sh, plan.d.wcol_total, plan.d.n_diag), tg);
}

// block geometry of the pair a -> b (either order of the two sweeps); false: not a pair
//   * one contiguous sweep (tile = bits [0, tb)) and the strided sweep whose stage bits start at tb:
//     blocks of 2^g_hi consecutive amplitudes, tile-number bits [0, g) count inside the block
//   * two strided sweeps with the same low run a and g stage bits each, [l, m) and [m, h):
//     blocks = low run x bits [l, h); in both tile numbers the bits [l - a, l - a + g) count inside the block
static bool pair_geometry(const pipe_shape &sh, const qft::sweep_plan &a, const qft::sweep_plan &b, int &inb_pos, int &inb_bits,
                          uint64_t &block_bytes)
{
    const bool a_strided = a.d.g_lo > a.d.a, b_strided = b.d.g_lo > b.d.a;
    if (a.n_tiles != b.n_tiles) return false;
    if (a_strided != b_strided) {
        const qft::sweep_plan &st = a_strided ? a : b;
        if (st.d.g_lo != sh.tb) return false;
        inb_pos = 0;
        inb_bits = st.d.g_hi - st.d.g_lo;
        block_bytes = 16ull << st.d.g_hi;
        return true;
    }
    if (!a_strided) return false;
    const qft::sweep_plan &hi = a.d.g_lo > b.d.g_lo ? a : b, &lo = a.d.g_lo > b.d.g_lo ? b : a;
    if (hi.d.g_lo != lo.d.g_hi || hi.d.a != lo.d.a) return false;
    if (hi.d.g_hi - hi.d.g_lo != lo.d.g_hi - lo.d.g_lo) return false;
    inb_pos = lo.d.g_lo - lo.d.a;
    inb_bits = lo.d.g_hi - lo.d.g_lo;
    block_bytes = 16ull << (lo.d.a + 2 * inb_bits);
    return true;
}

// Can the consecutive sweeps a -> b run as one L2-paired launch?
bool qcs_pipeline_pair_supported(const qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &a,
                                 const qft::sweep_plan &b)
{
    const pipe_shape sh = shape_of(reg);
    if (!reg->opt_l2_pair || !qcs_pipeline_supports(reg, tg, a) || !qcs_pipeline_supports(reg, tg, b)) return false;
    int inb_pos, inb_bits;
    uint64_t block_bytes;
    if (!pair_geometry(sh, a, b, inb_pos, inb_bits, block_bytes)) return false;
    if (a.d.slice_bits || b.d.slice_bits || a.d.tile_first || b.d.tile_first) return false;
    if (a.d.hadamard_only != b.d.hadamard_only || a.d.inverse != b.d.inverse) return false;    // one kernel mode per launch
    if (inb_bits > 20 || (uint64_t) a.n_tiles < (2ull << inb_bits)) return false;          // at least two blocks
    if (block_bytes > (uint64_t) reg->opt_l2_pair_max_block) return false;                  // a few blocks must fit the L2
    return pipe_smem(sh, a.d.wcol_total + b.d.wcol_total, a.d.n_diag + b.d.n_diag) <= reg->smem_optin;
}

int qcs_pipeline_launch_pair(qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &a, const qft::sweep_plan &b)
{
    const pipe_shape sh = shape_of(reg);
    const int shape_id = reg->opt_pipe_shape >= 0 && reg->opt_pipe_shape < kNumShapes ? reg->opt_pipe_shape : 0;
    pipe_params P = {};
    CUtensorMap tmap0, tmap1;
    const bool split_ok = reg->opt_split3 != 0;
    QCS_TRY(encode_phase(sh, tg, a, P.ph[0], tmap0, split_ok));
    QCS_TRY(encode_phase(sh, tg, b, P.ph[1], tmap1, split_ok));
    P.ph[0].wcol_base = 0;
    P.ph[1].wcol_base = a.d.wcol_total;
    P.n_phases = 2;
    P.n_tiles = a.n_tiles;
    int inb_pos = 0, inb_bits = 0;
    uint64_t block_bytes = 0;
    if (!pair_geometry(sh, a, b, inb_pos, inb_bits, block_bytes)) return QCS_BAD_ARGUMENTS;
    P.ph[0].inb_pos = P.ph[1].inb_pos = inb_pos;
    P.ph[0].inb_bits = P.ph[1].inb_bits = inb_bits;
    P.blk_tile_bits = inb_bits;
    const uint64_t per_block = 1ull << P.blk_tile_bits;
    const uint64_t n_blocks = a.n_tiles >> P.blk_tile_bits;
    // B trails A by one block plus what the chip has in flight (every CTA up to `stages` tiles)
    uint64_t lag = per_block + (uint64_t) reg->opt_l2_pair_lag;
    if (lag > a.n_tiles) lag = a.n_tiles;
    P.lag = lag;
    P.l2_hints = reg->opt_l2_pair_hints;
    QCS_TRY(take_generator(reg, tg, a, P));
    // queue head + one counter per block, zeroed in stream order
    const size_t need = 8 + 4 * (size_t) n_blocks;
    if (need > reg->d_pair_cap) {
        QCS_CUDA(cudaStreamSynchronize(tg.stream));
        if (reg->d_pair) QCS_CUDA(cudaFree(reg->d_pair));
        reg->d_pair = nullptr;
        reg->d_pair_cap = 0;
        QCS_CUDA(cudaMalloc(&reg->d_pair, need));
        reg->d_pair_cap = need;
    }
    QCS_CUDA(cudaMemsetAsync(reg->d_pair, 0, need, tg.stream));
    P.ticket = (unsigned long long *) reg->d_pair;
    P.done = (unsigned *) ((unsigned char *) reg->d_pair + 8);
    return run_launch(reg, sh, shape_id, tmap0, tmap1, P, pipe_smem(sh, a.d.wcol_total + b.d.wcol_total, a.d.n_diag + b.d.n_diag), tg);
}


// Host-only self-check of the paired-launch bookkeeping for tests (no device work): plans the inverse
// transform on qubits [lo, hi) of an n_qubits shard with the default shape, pairs the sweeps the way
// launch_plans does, and verifies for every pair that
//   * the ticket order hands out every tile of both sweeps exactly once, and a tile of the second sweep
//     only after `lag` >= one block of tiles of the first sweep were handed out before the first tile of
//     its block's successor ... precisely: after every first-sweep tile of its own block;
//   * in sampled blocks, the tiles of the first sweep and the tiles of the second sweep cover exactly the
//     same set of amplitudes (the block is closed under both sweeps).
// Returns the number of pairs checked (>= 0) or a negative number describing the first violation.
extern "C" long long qcs_pair_selfcheck(unsigned n_qubits, unsigned lo, unsigned hi, int lag_tiles)
{
    if (n_qubits < 15 || n_qubits > 40 || lo >= hi || hi > n_qubits) return -1;
    qcs_register fake = {};
    fake.n = fake.n_local = n_qubits;
    fake.N = fake.N_local = 1ull << n_qubits;
    fake.world = 1;
    fake.opt_pipeline = 1;
    fake.opt_pipe_shape = -1;
    fake.opt_min_run_bits = 3;
    fake.opt_l2_pair = 1;
    fake.opt_l2_pair_lag = lag_tiles;
    fake.opt_l2_pair_max_block = 16ll << 20;
    fake.smem_optin = 232448;
    const pipe_shape sh = shape_of(&fake);
    std::vector<qft::sweep_plan> plans;
    qft::plan_inverse(n_qubits, lo, hi, sh.tb, fake.opt_min_run_bits, plans);
    const qft::sweep_target tg = {nullptr, n_qubits, nullptr};
    long long pairs = 0;
    for (size_t k = 0; k + 1 < plans.size(); k++) {
        if (!qcs_pipeline_pair_supported(&fake, tg, plans[k], plans[k + 1])) continue;
        const qft::sweep_plan &a = plans[k], &b = plans[k + 1];
        pipe_params P = {};
        int inb_pos = 0, inb_bits = 0;
        uint64_t block_bytes = 0;
        if (!pair_geometry(sh, a, b, inb_pos, inb_bits, block_bytes)) return -2;
        const qft::sweep_plan *pl[2] = {&a, &b};
        for (int ph = 0; ph < 2; ph++) {
            P.ph[ph].d = pl[ph]->d;
            P.ph[ph].lo_gap = pl[ph]->d.g_lo > pl[ph]->d.a ? pl[ph]->d.g_lo - pl[ph]->d.a : -1;
            P.ph[ph].inb_pos = inb_pos;
            P.ph[ph].inb_bits = inb_bits;
        }
        P.n_phases = 2;
        P.n_tiles = a.n_tiles;
        P.blk_tile_bits = inb_bits;
        const uint64_t per_block = 1ull << inb_bits;
        P.lag = per_block + (uint64_t) lag_tiles;
        if (P.lag > P.n_tiles) P.lag = P.n_tiles;
        // --- ticket order (on a truncated problem so the check stays cheap: the order only depends on n_tiles and lag)
        const uint64_t n_tiles_checked = P.n_tiles > (1ull << 16) ? (1ull << 16) : P.n_tiles;
        pipe_params Pq = P;
        Pq.n_tiles = n_tiles_checked;
        if (Pq.lag > Pq.n_tiles) Pq.lag = Pq.n_tiles;
        std::vector<unsigned char> seen_a((size_t) n_tiles_checked, 0), seen_b((size_t) n_tiles_checked, 0);
        std::vector<uint64_t> done_a((size_t) (n_tiles_checked >> inb_bits) + 1, 0);
        for (uint64_t t = 0; t < 2 * n_tiles_checked; t++) {
            const uint64_t item = pair_item(Pq, t);
            if (item == kNoItem) return -3;
            const uint64_t idx = item & (kPhaseB - 1ull);
            if (idx >= n_tiles_checked) return -4;
            if (item & kPhaseB) {
                if (seen_b[(size_t) idx]++) return -5;
                if (done_a[(size_t) (idx >> inb_bits)] != per_block) return -6;      // a B tile before its block's A tiles
            } else {
                if (seen_a[(size_t) idx]++) return -7;
                done_a[(size_t) (idx >> inb_bits)]++;
            }
        }
        if (pair_item(Pq, 2 * n_tiles_checked) != kNoItem) return -8;
        // --- block closure in a few blocks: XOR / sum fingerprints of the amplitude sets of both sweeps
        const uint64_t n_blocks = P.n_tiles >> inb_bits;
        const uint64_t sample[4] = {0, 1 % n_blocks, n_blocks / 2, n_blocks - 1};
        for (uint64_t blk : sample) {
            uint64_t fp[2][2] = {{0, 0}, {0, 0}};
            for (int ph = 0; ph < 2; ph++) {
                const phase_params &Q = P.ph[ph];
                for (uint64_t inb = 0; inb < per_block; inb++) {
                    const uint64_t tix = pair_tile(P, Q, (blk << inb_bits) | inb);
                    const uint64_t base = sh.tb == 12 ? tile_base<12>(Q, tix) : tile_base<11>(Q, tix);
                    for (unsigned e = 0; e < (1u << sh.tb); e++) {
                        const uint64_t i = base + ((uint64_t) (e & ((1u << Q.d.a) - 1u)) | ((uint64_t) (e >> Q.d.a) << Q.d.g_lo));
                        fp[ph][0] ^= i * 0x9E3779B97F4A7C15ull;
                        fp[ph][1] += i * i + 1;
                    }
                }
            }
            if (fp[0][0] != fp[1][0] || fp[0][1] != fp[1][1]) return -9;
        }
        pairs++;
        k++;
    }
    return pairs;
}
