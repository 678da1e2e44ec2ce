// qft_fused.cu -- the (inverse) QFT as a few fused tile sweeps.
//
// The reference applies inverse_QFT (qc_shor.c:678-690) as L Hadamards and
// L(L-1)/2 controlled phase gates, each a full pass over the state.  Written
// out, stage l of that circuit is "butterfly on bit l, then multiply every
// amplitude whose bit l is 1 by exp(i pi (x mod 2^j) / 2^j)" with x the value
// of the L register and j = l - M: a radix-2 decimation-in-frequency FFT with
// positive exponent and no final bit reversal.  Here the stages are grouped:
//
//   sweep  = one pass over HBM.  A tile of 2^t amplitudes (the low `a` index
//            bits, for coalescing, plus a run of stage bits [g_lo, g_hi)) is
//            brought on chip, all stages whose bit lies in the tile are applied,
//            and the tile is written back in place: 32 B per amplitude per sweep
//            instead of 32 B (H) or 8 B (C-phase) per gate.
//   step   = r <= 4 consecutive stages inside a sweep, done in registers as one
//            radix-2^r butterfly network (constant internal twiddles) followed
//            by one "external" twiddle w^k per element, w = exp(i pi y / 2^j),
//            y = the register bits below the step.  y splits into a part fixed
//            per tile (one sincospi per step per tile) and a part fixed per
//            thread column (one sincospi per column per kernel); the powers
//            w^2..w^(R-1) are formed by squaring / one extra multiply.
//
// Between steps the tile lives in shared memory (XOR-swizzled at 16-byte
// granularity so every step's access is bank-conflict free); the first step of
// a sweep loads straight from global memory and the last stores straight back,
// so a sweep with k steps makes k-1 shared-memory round trips.
//
// The forward QFT (the adjoint circuit) runs the same schedule backwards with
// conjugated twiddles: external twiddle first, then decimation-in-time
// butterflies.
//
// Amplitudes agree with the gate-by-gate reference order to ~1e-15 relative
// (tests/test_fused_gpu.py); they are not bit-identical, which is why
// QCS_OPT_FUSION = 0 keeps the gate-by-gate kernels available.
#include "qft_common.cuh"

#include <string.h>
#include <string>


namespace {

using namespace qft;

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_qft_sweep(double2 *__restrict__ amp, uint64_t n_tiles, const sweep_desc P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *tile = (double2 *) smem_raw;
    double2 *wcol = tile + (1u << P.t);
    double2 *wbase = wcol + P.wcol_total;        // [2][kMaxSteps]
    tile_geom G;
    G.a = P.a;
    G.g_lo = P.g_lo;
    G.sw = P.sw;
    const bool inv = P.inverse != 0;

    // per-column part of the external twiddle: fixed for the whole kernel
    for (int k = 0; k < (P.hadamard_only ? 0 : P.n_steps); k++) {
        const sweep_step S = P.step[k];
        if (S.notw) continue;
        // one entry per value of the tile-local bits below the step (e_base = c for c < 2^s)
        for (unsigned c = threadIdx.x; c < (1u << S.s); c += NT) {
            const uint64_t y = (G.spread(c) & ((1ull << S.low_phys) - 1ull)) >> P.lo;
            wcol[S.col_off + c] = unit_phase(y, S.j, inv);
        }
    }

    unsigned parity = 0;
    const int lo_gap = P.g_lo - P.a;             // number of index bits between the two tile runs
    for (uint64_t kt = blockIdx.x; kt < n_tiles; kt += gridDim.x, parity ^= 1u) {
        const uint64_t tix = tile_number(P, kt);
        // deposit the tile number into the index bits that are not in the tile
        const uint64_t base = lo_gap > 0 ? (((tix >> lo_gap) << P.g_hi) | ((tix & ((1ull << lo_gap) - 1ull)) << P.a))
                                         : (tix << P.t);
        if (threadIdx.x < (unsigned) P.n_steps) {
            const sweep_step S = P.step[threadIdx.x];
            uint64_t y = 0;
            if (S.low_phys > P.lo) y = (base & ((1ull << S.low_phys) - 1ull)) >> P.lo;
            wbase[parity * kMaxSteps + threadIdx.x] = unit_phase(y + P.y_const, S.j, inv);
        }
        __syncthreads();
        for (int k = 0; k < P.n_steps; k++) {
            const sweep_step S = P.step[k];
            const bool first = k == 0, last = k == P.n_steps - 1;
            const double2 wb = wbase[parity * kMaxSteps + k];
            if (P.hadamard_only) dispatch_step<true, false>(amp, tile, wcol + S.col_off, wb, G, S, P.t, base, first, last, last, P.scale, threadIdx.x, NT, P.diag, P.n_diag, P.index_or);
            else if (inv) dispatch_step<true>(amp, tile, wcol + S.col_off, wb, G, S, P.t, base, first, last, last, P.scale, threadIdx.x, NT);
            else dispatch_step<false>(amp, tile, wcol + S.col_off, wb, G, S, P.t, base, first, last, last, P.scale, threadIdx.x, NT);
            if (!last) __syncthreads();
        }
    }
}

template <int NT, int MINB>
int launch_sweep(qcs_register *reg, const sweep_target &tg, const sweep_plan &p, size_t smem)
{
    auto kern = k_qft_sweep<NT, MINB>;
    QCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0;
    QCS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) return QCS_UNKNOWN_ERROR;
    uint64_t grid = (uint64_t) reg->sm_count * (uint64_t) per_sm;
    if (tg.max_ctas > 0 && grid > (uint64_t) tg.max_ctas * (uint64_t) per_sm) grid = (uint64_t) tg.max_ctas * (uint64_t) per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    qcs_launch_begin(reg, tg.kind, tg.bytes > 0.0 ? tg.bytes : 32.0 * (double) (p.n_tiles << p.d.t));
    kern<<<(unsigned) grid, NT, smem, tg.stream>>>(tg.amp, p.n_tiles, p.d);
    return qcs_launch_end(reg, tg.kind, "k_qft_sweep");
}

}  // namespace

int qcs_pipeline_tile_bits(const qcs_register *reg);
bool qcs_pipeline_supports(const qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &p);
int qcs_pipeline_launch(qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &plan);
bool qcs_pipeline_pair_supported(const qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &a,
                                 const qft::sweep_plan &b);
int qcs_pipeline_launch_pair(qcs_register *reg, const qft::sweep_target &tg, const qft::sweep_plan &a, const qft::sweep_plan &b);

static int launch_plan_inner(qcs_register *reg, const sweep_target &tg, const sweep_plan &p);

static int default_tile_bits(const qcs_register *reg)
{
    if (reg->opt_tile_bits) return reg->opt_tile_bits;
    return reg->opt_pipeline ? qcs_pipeline_tile_bits(reg) : 11;
}

static int launch_plan(qcs_register *reg, const sweep_target &tg, const sweep_plan &p)
{
    reg->launch_stream = tg.stream;
    const int rc = launch_plan_inner(reg, tg, p);
    reg->launch_stream = nullptr;
    return rc;
}

static int launch_plan_inner(qcs_register *reg, const sweep_target &tg, const sweep_plan &p)
{
    const size_t smem = ((size_t) 16 << p.d.t) + (size_t) 16 * (size_t) p.d.wcol_total + 16 * 2 * kMaxSteps;
    if (smem > reg->smem_optin) return QCS_BAD_ARGUMENTS;
    if (reg->opt_pipeline && qcs_pipeline_supports(reg, tg, p)) return qcs_pipeline_launch(reg, tg, p);
    if (p.d.t >= 13) return launch_sweep<512, 1>(reg, tg, p, smem);
    if (p.d.t == 12) return launch_sweep<256, 2>(reg, tg, p, smem);
    return launch_sweep<128, 4>(reg, tg, p, smem);
}

// a list of consecutive sweeps on one target: neighbours that can share an L2-paired launch do.
// Pairs are formed from the contiguous sweep outwards (it is the last sweep of an inverse transform
// and the first of a forward one), so that it always has a partner.
static std::vector<int> pair_flags(const qcs_register *reg, const sweep_target &tg, const std::vector<sweep_plan> &plans)
{
    const size_t n = plans.size();
    std::vector<int> with_next(n, 0);
    if (reg->opt_pipeline && n >= 2) {
        const bool contiguous_last = !(plans[n - 1].d.g_lo > plans[n - 1].d.a);
        if (contiguous_last) {
            for (size_t k = n - 1; k >= 1; ) {
                if (qcs_pipeline_pair_supported(reg, tg, plans[k - 1], plans[k])) { with_next[k - 1] = 1; if (k < 2) break; k -= 2; }
                else { k -= 1; }
            }
        } else {
            for (size_t k = 0; k + 1 < n; ) {
                if (qcs_pipeline_pair_supported(reg, tg, plans[k], plans[k + 1])) { with_next[k] = 1; k += 2; }
                else { k += 1; }
            }
        }
    }
    return with_next;
}

static int launch_plans(qcs_register *reg, const sweep_target &tg, const std::vector<sweep_plan> &plans)
{
    const size_t n = plans.size();
    const std::vector<int> with_next = pair_flags(reg, tg, plans);
    for (size_t k = 0; k < n; k++) {
        if (with_next[k]) {
            reg->launch_stream = tg.stream;
            const int rc = qcs_pipeline_launch_pair(reg, tg, plans[k], plans[k + 1]);
            reg->launch_stream = nullptr;
            QCS_TRY(rc);
            k++;
            continue;
        }
        QCS_TRY(launch_plan(reg, tg, plans[k]));
    }
    return QCS_NO_ERROR;
}

static int run_sweeps(qcs_register *reg, unsigned lo, unsigned hi, bool inverse, bool hadamard_only)
{
    if (hi > reg->n_local) {
        fprintf(stderr, "qcs: fused sweeps over global qubits are handled by the distributed schedule\n");
        return QCS_BAD_ARGUMENTS;
    }
    const int T = default_tile_bits(reg);
    std::vector<sweep_plan> plans;
    plan_inverse(reg->n_local, lo, hi, T, reg->opt_min_run_bits, plans);
    if (!inverse) make_forward(plans);
    const sweep_target tg = {reg->amp, reg->n_local, reg->stream};
    for (sweep_plan &p : plans) {
        p.d.hadamard_only = hadamard_only ? 1 : 0;
        if (hadamard_only) p.d.wcol_total = 0;
    }
    return launch_plans(reg, tg, plans);
}

// Can the first sweep of the inverse transform on [lo, hi) GENERATE its tiles (quantum_computation from the
// reset state, modexp_fused.cu)?  It must be the first sweep launched, a strided tile of the pipelined
// kernel (laid out linearly), and its rows must lie inside one block: low run inside the M register (a <= lo),
// stage bits inside the L register (lo <= g_lo), nothing sliced.
bool qcs_fused_gen_supported(const qcs_register *reg, unsigned lo, unsigned hi)
{
    if (!reg->opt_pipeline || reg->world != 1 || lo >= hi || hi > reg->n_local) return false;
    std::vector<sweep_plan> plans;
    plan_inverse(reg->n_local, lo, hi, default_tile_bits(reg), reg->opt_min_run_bits, plans);
    if (plans.empty()) return false;
    const sweep_target tg = {reg->amp, reg->n_local, reg->stream};
    const sweep_desc &d = plans[0].d;
    if (!(d.g_lo > d.a) || d.a > (int) lo || d.g_lo < (int) lo || d.slice_bits != 0 || d.tile_first != 0 ||
        d.t != qcs_pipeline_tile_bits(reg) || !qcs_pipeline_supports(reg, tg, plans[0]))
        return false;
    return true;                 // alone or as the first sweep of an L2-paired launch
}

int qcs_plan_hadamard_sweeps(const qcs_register *reg, unsigned lo, unsigned hi, std::vector<sweep_plan> &plans)
{
    if (lo >= hi || hi > reg->n_local) return QCS_BAD_ARGUMENTS;
    plan_inverse(reg->n_local, lo, hi, default_tile_bits(reg), reg->opt_min_run_bits, plans);
    for (sweep_plan &p : plans) {
        p.d.hadamard_only = 1;
        p.d.wcol_total = 0;             // no twiddle tables
    }
    return QCS_NO_ERROR;
}

// two consecutive sweeps as one L2-paired launch when they qualify (*paired tells), else nothing is launched
int qcs_launch_sweep_pair(qcs_register *reg, const sweep_plan &a, const sweep_plan &b, bool *paired)
{
    const sweep_target tg = {reg->amp, reg->n_local, reg->stream};
    *paired = reg->opt_pipeline && qcs_pipeline_pair_supported(reg, tg, a, b);
    if (!*paired) return QCS_NO_ERROR;
    reg->launch_stream = tg.stream;
    const int rc = qcs_pipeline_launch_pair(reg, tg, a, b);
    reg->launch_stream = nullptr;
    return rc;
}

int qcs_launch_sweep_plan(qcs_register *reg, const sweep_plan &plan)
{
    const sweep_target tg = {reg->amp, reg->n_local, reg->stream};
    return launch_plan(reg, tg, plan);
}

// One sweep over the top `p` index bits of `buf` (2^(c+p) amplitudes laid out
// [slot][2^c]) that stand for the register's global qubits [n-p, n) after the
// global<->local exchange (dist.cu).  y_const carries the register bits that
// are not index bits of `buf` (the slice offset and the bits held by the rank).
int qcs_fused_top_sweep(qcs_register *reg, double2 *buf, unsigned c, unsigned p, unsigned lo,
                        unsigned long long y_const, bool inverse, bool hadamard_only, cudaStream_t stream)
{
    if (p < 1 || p > 4) return QCS_BAD_ARGUMENTS;
    const unsigned nb = c + p;
    int T = default_tile_bits(reg);
    if ((unsigned) T > nb) T = (int) nb;
    sweep_plan pl = {};              // no diagonal gates, no rank bits
    sweep_desc &d = pl.d;
    d.t = T;
    d.a = T - (int) p;
    if (d.a < 0) return QCS_BAD_ARGUMENTS;
    d.g_lo = (int) c;
    d.g_hi = (int) (c + p);
    if (d.g_lo < d.a) return QCS_BAD_ARGUMENTS;
    d.lo = (int) lo;
    d.inverse = inverse ? 1 : 0;
    d.hadamard_only = hadamard_only ? 1 : 0;
    d.y_const = y_const;
    d.tile_first = 0;
    d.n_steps = 1;
    d.scale = (p % 2 == 0) ? ldexp(1.0, -(int) p / 2) : pow(0.70710678118654752440, (double) p);
    d.step[0].r = (int) p;
    d.step[0].s = d.a;
    d.step[0].low_phys = (int) c;
    d.step[0].j = (int) (reg->n - 1) - (int) lo;
    d.step[0].col_off = 0;
    d.step[0].notw = 0;
    d.wcol_total = 1 << d.a;
    d.sw = 28;
    pl.n_tiles = 1ull << (nb - (unsigned) T);
    pl.stages = (int) p;
    const sweep_target tg = {buf, nb, stream};
    return launch_plan(reg, tg, pl);
}

int qcs_fused_qft(qcs_register *reg, unsigned lo, unsigned hi, bool inverse)
{
    return run_sweeps(reg, lo, hi, inverse, false);
}

// Sharded register with peer memory (peer.cu), transform on qubits [lo, hi) reaching the
// global qubits (hi > n_local):
//   * the top stages -- the global qubits and as many local ones as fit beside a long
//     contiguous run -- are ONE tile sweep on the stitched array amp_all.  Every rank runs it on
//     its 1/world share of the tiles; the rows of a tile that live on other ranks are fetched
//     (and written back) over NVLink by the TMA engine while the consumer groups work on tiles
//     that already arrived: the exchange is part of the sweep.  A stream-ordered barrier
//     separates it from the work around it;
//   * the remaining stages are ordinary sweeps on the rank's own shard: no communication.
// NVLink traffic: each rank reads and writes (world-1)/world of a shard, once.
// can qcs_fused_sweeps_sharded run the transform on [lo, hi)?  (callers fall back to the
// exchange schedules of dist.cu otherwise)
bool qcs_sharded_sweeps_supported(const qcs_register *reg, unsigned lo, unsigned hi)
{
    if (!reg->peer || !reg->amp_all || lo >= hi || hi > reg->n || hi <= reg->n_local) return false;
    const int T = default_tile_bits(reg);
    if ((int) reg->n_local < T + 3 || (int) lo + T > (int) reg->n_local) return false;
    int a_glob = reg->opt_global_run_bits;
    if (a_glob > T - reg->p_global) a_glob = T - reg->p_global;
    if (a_glob < 1) return false;
    int split = (int) hi - (T - a_glob);
    if (split < (int) lo) split = (int) lo;
    return split <= (int) reg->n_local;
}

int qcs_fused_sweeps_sharded(qcs_register *reg, unsigned lo, unsigned hi, bool inverse, bool hadamard_only)
{
    if (!qcs_sharded_sweeps_supported(reg, lo, hi)) return QCS_BAD_ARGUMENTS;
    const int T = default_tile_bits(reg);
    // The sweep over the global qubits is bound by NVLink, not by HBM: it keeps long contiguous
    // runs (2^a amplitudes per TMA row) and takes along only as many local stage bits as fit.
    int a_glob = reg->opt_global_run_bits;
    if (a_glob > T - reg->p_global) a_glob = T - reg->p_global;
    if (a_glob < 1) return QCS_BAD_ARGUMENTS;
    int split = (int) hi - (T - a_glob);                    // global sweep: stages [split, hi)
    if (split < (int) lo) split = (int) lo;
    if (split > (int) reg->n_local) return QCS_BAD_ARGUMENTS;
    std::vector<sweep_plan> global_plans, local_plans;
    plan_inverse(reg->n, lo, hi, T, a_glob, global_plans, split);
    if (split > (int) lo) plan_inverse(reg->n_local, lo, (unsigned) split, T, reg->opt_min_run_bits, local_plans);
    if (!inverse) {
        make_forward(global_plans);
        make_forward(local_plans);
    }
    auto run_global = [&]() -> int {
        for (sweep_plan &p : global_plans) {
            p.d.hadamard_only = hadamard_only ? 1 : 0;
            if (hadamard_only) p.d.wcol_total = 0;
            const uint64_t share = p.n_tiles >> reg->p_global;
            if (share == 0) return QCS_BAD_ARGUMENTS;
            // peers read what this rank wrote before the sweep and write what it reads after it
            QCS_TRY(qcs_dist_stream_barrier(reg));
            p.d.tile_first = (uint64_t) reg->rank * share;
            p.n_tiles = share;
            // accounted as NVLink traffic per direction: this rank's remote reads plus the peers'
            // reads of its shard one way, its remote writes plus theirs the other way
            sweep_target tg = {reg->amp_all, reg->n, reg->stream};
            tg.kind = QCS_K_GLOBAL_SWEEP;
            tg.bytes = 2.0 * 16.0 * (double) (share << p.d.t) * (double) (reg->world - 1) / (double) reg->world;
            QCS_TRY(launch_plan(reg, tg, p));
            QCS_TRY(qcs_dist_stream_barrier(reg));
        }
        return QCS_NO_ERROR;
    };
    auto run_local = [&]() -> int {
        const sweep_target tg = {reg->amp, reg->n_local, reg->stream};
        for (sweep_plan &p : local_plans) {
            p.d.hadamard_only = hadamard_only ? 1 : 0;
            if (hadamard_only) p.d.wcol_total = 0;
            QCS_TRY(launch_plan(reg, tg, p));
        }
        return QCS_NO_ERROR;
    };
    // ---- overlapped schedule -----------------------------------------------------------------
    // The global sweep is bound by NVLink and needs only a fraction of the SMs; the local sweeps
    // are bound by HBM.  The tiles are cut into K slices by the index bits just above the global
    // sweep's contiguous run, bits [a_glob, a_glob + log2 K): they are "gap" bits (neither in the
    // tile nor stage bits) of the global sweep and of every strided local sweep whose own run is
    // not longer, so slice j of those local sweeps touches exactly the amplitudes slice j of the
    // global sweep produced.  Stream G: global slice 0, barrier, global slice 1, barrier, ...;
    // stream L: after barrier j, the strided local sweeps of slice j on the other SMs.
    int K = reg->opt_overlap_slices;
    int sb = 0;
    while ((1 << (sb + 1)) <= K) sb++;
    // the strided local sweeps next to the global sweep (first in the inverse order, last in the
    // forward order) whose tiles leave the slice bits alone
    size_t n_sliceable = 0;
    if (K >= 2 && global_plans.size() == 1 && global_plans[0].d.g_lo > global_plans[0].d.a) {
        for (size_t k = 0; k < local_plans.size(); k++) {
            const sweep_plan &p = local_plans[inverse ? k : local_plans.size() - 1 - k];
            const bool strided = p.d.g_lo > p.d.a;
            if (!strided || p.d.a > a_glob || p.d.g_lo < a_glob + sb || (p.n_tiles >> sb) == 0) break;
            n_sliceable++;
        }
        const sweep_plan &g = global_plans[0];
        if (((g.n_tiles >> reg->p_global) >> sb) == 0 || g.d.g_lo < a_glob + sb) n_sliceable = 0;
    }
    // one slice of the global sweep / of a strided local sweep
    auto global_slice = [&](int j, cudaStream_t st, int max_ctas) -> int {
        sweep_plan p = global_plans[0];
        p.d.hadamard_only = hadamard_only ? 1 : 0;
        if (hadamard_only) p.d.wcol_total = 0;
        const uint64_t share = (p.n_tiles >> reg->p_global) >> sb;
        p.d.tile_first = (uint64_t) reg->rank * share;
        p.n_tiles = share;
        p.d.slice_pos = 0;
        p.d.slice_bits = sb;
        p.d.slice_val = (unsigned) j;
        sweep_target tg = {reg->amp_all, reg->n, st};
        tg.kind = QCS_K_GLOBAL_SWEEP;
        tg.bytes = 2.0 * 16.0 * (double) (share << p.d.t) * (double) (reg->world - 1) / (double) reg->world;
        tg.max_ctas = max_ctas;
        return launch_plan(reg, tg, p);
    };
    auto local_slice = [&](size_t k, int j, cudaStream_t st, int max_ctas) -> int {
        sweep_plan q = local_plans[k];
        q.d.hadamard_only = hadamard_only ? 1 : 0;
        if (hadamard_only) q.d.wcol_total = 0;
        q.n_tiles >>= sb;
        q.d.tile_first = 0;
        q.d.slice_pos = a_glob - q.d.a;
        q.d.slice_bits = sb;
        q.d.slice_val = (unsigned) j;
        sweep_target tl = {reg->amp, reg->n_local, st};
        tl.max_ctas = max_ctas;
        return launch_plan(reg, tl, q);
    };
    auto local_full = [&](size_t k, cudaStream_t st) -> int {
        sweep_plan q = local_plans[k];
        q.d.hadamard_only = hadamard_only ? 1 : 0;
        if (hadamard_only) q.d.wcol_total = 0;
        const sweep_target tl = {reg->amp, reg->n_local, st};
        return launch_plan(reg, tl, q);
    };
    if (n_sliceable > 0 && !inverse) {
        // forward transform: the mirror image.  L: the unsliced local sweeps, then slice by slice the
        // strided ones; G: after a barrier, the global sweep of each finished slice.
        cudaStream_t G = qcs_dist_side_stream(reg), L = reg->stream;
        cudaEvent_t ev;
        int g_sms = reg->opt_global_sms;
        if (g_sms >= reg->sm_count) g_sms = reg->sm_count / 2;
        const size_t first_sliced = local_plans.size() - n_sliceable;
        for (size_t k = 0; k < first_sliced; k++) QCS_TRY(local_full(k, L));
        for (int j = 0; j < (1 << sb); j++) {
            // global slices of earlier j run beside these on their own SMs
            for (size_t k = first_sliced; k < local_plans.size(); k++)
                QCS_TRY(local_slice(k, j, L, j > 0 ? reg->sm_count - g_sms : 0));
            QCS_TRY(qcs_dist_slice_event(reg, j, &ev));
            QCS_CUDA(cudaEventRecord(ev, L));
            QCS_CUDA(cudaStreamWaitEvent(G, ev, 0));
            QCS_TRY(qcs_dist_barrier_on(reg, G));          // slice j is finished on every rank
            QCS_TRY(global_slice(j, G, j + 1 < (1 << sb) ? g_sms : 0));
        }
        QCS_TRY(qcs_dist_barrier_on(reg, G));
        QCS_TRY(qcs_dist_slice_event(reg, 15, &ev));
        QCS_CUDA(cudaEventRecord(ev, G));
        QCS_CUDA(cudaStreamWaitEvent(L, ev, 0));
        return QCS_NO_ERROR;
    }
    if (n_sliceable > 0) {
        cudaStream_t G = qcs_dist_side_stream(reg), L = reg->stream;
        cudaEvent_t ev;
        QCS_TRY(qcs_dist_slice_event(reg, 15, &ev));
        QCS_CUDA(cudaEventRecord(ev, L));
        QCS_CUDA(cudaStreamWaitEvent(G, ev, 0));
        QCS_TRY(qcs_dist_barrier_on(reg, G));
        int g_sms = reg->opt_global_sms;
        if (g_sms >= reg->sm_count) g_sms = reg->sm_count / 2;
        for (int j = 0; j < (1 << sb); j++) {
            QCS_TRY(global_slice(j, G, g_sms));
            QCS_TRY(qcs_dist_barrier_on(reg, G));
            QCS_TRY(qcs_dist_slice_event(reg, j, &ev));
            QCS_CUDA(cudaEventRecord(ev, G));
            QCS_CUDA(cudaStreamWaitEvent(L, ev, 0));
            // while global slices are still to come they keep their SMs
            for (size_t k = 0; k < n_sliceable; k++)
                QCS_TRY(local_slice(k, j, L, j + 1 < (1 << sb) ? reg->sm_count - g_sms : 0));
        }
        for (size_t k = n_sliceable; k < local_plans.size(); k++) QCS_TRY(local_full(k, L));
        return QCS_NO_ERROR;
    }
    if (inverse) {
        QCS_TRY(run_global());
        return run_local();
    }
    QCS_TRY(run_local());
    return run_global();
}

// Hadamard on every qubit of [lo, hi) (the first loop of quantum_computation,
// qc_shor.c:720-722) as Walsh-Hadamard tile sweeps
int qcs_fused_hadamards(qcs_register *reg, unsigned lo, unsigned hi)
{
    return run_sweeps(reg, lo, hi, true, true);
}


// Host-only view of the sweep planner for tests and tooling (no device work): the sweeps an
// inverse transform on qubits [lo, hi) of a 2^n_qubits-amplitude array is run as, one line each:
//   "sweep a=<run bits> g=[<g_lo>,<g_hi>) tiles=<count> steps=<low_phys>+<r>,... scale=<factor>"
extern "C" int qcs_plan_describe(unsigned n_qubits, unsigned lo, unsigned hi, int tile_bits, int min_run_bits,
                                 char *out, unsigned long long out_cap)
{
    if (!out || out_cap == 0 || n_qubits < 1 || n_qubits > 62 || lo >= hi || hi > n_qubits) return QCS_BAD_ARGUMENTS;
    if (tile_bits == 0) tile_bits = 12;
    if (min_run_bits == 0) min_run_bits = 3;
    if (tile_bits < 4 || tile_bits > 13 || min_run_bits < 1 || min_run_bits > 7) return QCS_BAD_ARGUMENTS;
    std::vector<sweep_plan> plans;
    plan_inverse(n_qubits, lo, hi, tile_bits, min_run_bits, plans);
    std::string text;
    for (const sweep_plan &p : plans) {
        char buf[160];
        snprintf(buf, sizeof buf, "sweep a=%d g=[%d,%d) tiles=%llu steps=", p.d.a, p.d.g_lo, p.d.g_hi,
                 (unsigned long long) p.n_tiles);
        text += buf;
        for (int k = 0; k < p.d.n_steps; k++) {
            snprintf(buf, sizeof buf, "%s%d+%d", k ? "," : "", p.d.step[k].low_phys, p.d.step[k].r);
            text += buf;
        }
        snprintf(buf, sizeof buf, " scale=%.17g\n", p.d.scale);
        text += buf;
    }
    if (text.size() + 1 > out_cap) return QCS_INSUFFICIENT_MEMORY;
    memcpy(out, text.c_str(), text.size() + 1);
    return QCS_NO_ERROR;
}
