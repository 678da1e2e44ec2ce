// circuit.cu -- deferred gate stream: hadamard_gate / c_phase_shift_gate calls
// recorded between qcs_fuse_begin and qcs_fuse_end are scheduled into as few
// passes over HBM as their commutation rules allow, then launched.
//
// The reference applies every gate as one full pass over the state
// (operate_matrix, qc_shor.c:370-420).  Two facts let a run of gates share a
// pass:
//   * Hadamards on different qubits commute, so a set of them is a
//     Walsh-Hadamard transform on those index bits: contiguous runs of qubits
//     become the tile sweeps of qft_fused.cu / qft_pipeline.cu with the
//     twiddles compiled out;
//   * controlled phase gates are diagonal (qc_shor.c:220-225): they commute
//     with each other and with a Hadamard on any qubit they do not touch.  A
//     diagonal gate may therefore be applied at the end of ANY pass between the
//     last earlier Hadamard on one of its qubits and the next later one; it is
//     attached to the least loaded tile sweep in that window, where the
//     amplitudes are in registers and their full index is known.
//
// The stream is cut into groups  { set of H qubits } -> { diagonal gates }  by a
// single scan in program order; nothing is reordered across a gate it does not
// commute with.  Amplitudes agree with gate-by-gate application to ~1e-15
// relative (bar 1e-12); with QCS_OPT_FUSION = 0 qcs_fuse_begin records nothing
// and every gate runs immediately through the reference-order kernels.
#include "qft_common.cuh"

#include <string.h>
#include <string>

namespace {

using qft::diag_gate;
using qft::sweep_plan;

constexpr int kThreads = 256;
constexpr size_t kDiagScratchSlots = 1024;   // head of d_diag: staging of the standalone diagonal passes
constexpr int kMaxDiagPerSweep = 48;   // beyond this a sweep turns FP64-bound: spill to a diagonal pass

// one pass over the shard applying a list of diagonal gates: 32 B per amplitude
// for any number of gates (each gate alone would cost 8 B per amplitude)
__global__ void __launch_bounds__(kThreads)
k_diag_multi(double2 *__restrict__ amp, uint64_t n_pairs, const diag_gate *__restrict__ gates, int n_gates)
{
    extern __shared__ diag_gate sg[];
    for (int i = threadIdx.x; i < n_gates; i += kThreads) sg[i] = gates[i];
    __syncthreads();
    const uint64_t stride = (uint64_t) gridDim.x * kThreads;
    for (uint64_t p = (uint64_t) blockIdx.x * kThreads + threadIdx.x; p < n_pairs; p += stride) {
        double2 x0, x1;
        x0 = qft::ld256_lo(amp + 2 * p, x1);
        const uint64_t i0 = 2 * p, i1 = 2 * p + 1;
        for (int g = 0; g < n_gates; g++) {
            const diag_gate dg = sg[g];
            const double2 ph = make_double2(dg.c, dg.s);
            if ((i0 & dg.mask) == dg.mask) x0 = qft::cmul(x0, ph);
            if ((i1 & dg.mask) == dg.mask) x1 = qft::cmul(x1, ph);
        }
        qft::st256(amp + 2 * p, x0, x1);
    }
}

struct pass {
    int type;                   // 0: tile sweep, 1: single local Hadamard kernel, 2: global Hadamard(s),
                                // 3: run of H qubits reaching the global ones, as sharded sweeps on peer memory
    unsigned run_lo = 0, run_hi = 0;    // type 3
    int group = 0;
    uint64_t hmask = 0;         // qubits this pass applies H to
    sweep_plan plan = {};       // type 0
    unsigned q = 0;             // type 1
    bool top_stages = false;    // type 2: all global qubits at once through the qubit-swap pipeline
    std::vector<diag_gate> in_sweep;    // applied in the last step of the sweep (type 0)
    std::vector<int> in_sweep_idx;      // ... and their positions in the recorded stream
    std::vector<int> after;             // recorded gates applied by standalone kernels after the pass
};

int ensure_diag_capacity(qcs_register *reg, size_t gates)
{
    if (gates <= reg->d_diag_cap) return QCS_NO_ERROR;
    // the old array may still be read by sweeps in flight on the stream
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    if (reg->d_diag) QCS_CUDA(cudaFree(reg->d_diag));
    reg->d_diag = nullptr;
    reg->d_diag_cap = 0;
    size_t cap = 256;
    while (cap < gates) cap *= 2;
    QCS_CUDA(cudaMalloc(&reg->d_diag, cap * sizeof(diag_gate)));
    reg->d_diag_cap = cap;
    return QCS_NO_ERROR;
}

// the part of a recorded diagonal gate that lives in this shard: global qubits
// contribute the rank's (constant) bit.  false: the gate is the identity here.
bool localise(const qcs_register *reg, const qcs_pending_gate &g, diag_gate &out)
{
    uint64_t mask = 0;
    const unsigned both[2] = {g.q0, g.q1};
    for (int k = 0; k < 2; k++) {
        const unsigned b = both[k];
        if (b < reg->n_local) mask |= 1ull << b;
        else if (!(((unsigned) reg->rank >> (b - reg->n_local)) & 1u)) return false;
    }
    out.mask = mask;
    out.c = g.c;
    out.s = g.s;
    return true;
}

int launch_diag_list(qcs_register *reg, const std::vector<qcs_pending_gate> &queue, const std::vector<int> &list,
                     diag_gate *d_slot, std::vector<diag_gate> &host_stage)
{
    if (list.empty()) return QCS_NO_ERROR;
    if (list.size() <= 3) {
        // 8 B per amplitude each: cheaper than one 32 B pass, and reference-order arithmetic
        for (int gi : list) {
            const qcs_pending_gate &g = queue[(size_t) gi];
            diag_gate dg;
            if (!localise(reg, g, dg)) continue;
            unsigned bits[2];
            int nb = 0;
            for (unsigned b = 0; b < reg->n_local; b++)
                if ((dg.mask >> b) & 1ull) bits[nb++] = b;
            QCS_TRY(qcs_k_phase_masked(reg, nb, nb > 0 ? bits[0] : 0, nb > 1 ? bits[1] : 0, dg.c, dg.s));
        }
        return QCS_NO_ERROR;
    }
    host_stage.clear();
    for (int gi : list) {
        diag_gate dg;
        if (localise(reg, queue[(size_t) gi], dg)) host_stage.push_back(dg);
    }
    if (host_stage.empty()) return QCS_NO_ERROR;
    // every chunk of 1024 gates goes through the SAME 1024-slot scratch region: the slots after it
    // hold the in-sweep tables of later sweeps.  Reuse is safe in stream order (chunk k's kernel has
    // read the region before chunk k+1's copy overwrites it).
    for (size_t at = 0; at < host_stage.size(); at += kDiagScratchSlots) {
        const int n = (int) std::min<size_t>(kDiagScratchSlots, host_stage.size() - at);
        QCS_CUDA(cudaMemcpyAsync(d_slot, host_stage.data() + at, (size_t) n * sizeof(diag_gate),
                                 cudaMemcpyHostToDevice, reg->stream));
        const uint64_t n_pairs = reg->N_local >> 1;
        uint64_t grid = (n_pairs + kThreads - 1) / kThreads;
        const uint64_t cap = (uint64_t) reg->sm_count * 8;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        qcs_launch_begin(reg, QCS_K_DIAG, 32.0 * (double) reg->N_local);
        k_diag_multi<<<(unsigned) grid, kThreads, (size_t) n * sizeof(diag_gate), reg->stream>>>(
            reg->amp, n_pairs, d_slot, n);
        QCS_TRY(qcs_launch_end(reg, QCS_K_DIAG, "k_diag_multi"));
    }
    return QCS_NO_ERROR;
}

}  // namespace

// The scheduler proper: host-only (it reads the register's shape and options, nothing on the
// device), so tests can drive it without a GPU through qcs_schedule_describe.
static int schedule_stream(const qcs_register *reg, const std::vector<qcs_pending_gate> &queue,
                           std::vector<pass> &passes, std::vector<int> &before_all)
{
    passes.clear();
    before_all.clear();

    // ---- 1. groups: {H set} then {diagonal gates}, cut in program order
    struct group { uint64_t hset = 0, dmask = 0; std::vector<int> diag; };
    std::vector<group> groups(1);
    std::vector<int> group_of(queue.size(), 0);
    for (size_t i = 0; i < queue.size(); i++) {
        const qcs_pending_gate &g = queue[i];
        if (g.kind == 0) {
            const uint64_t bit = 1ull << g.q0;
            if ((groups.back().hset | groups.back().dmask) & bit) groups.emplace_back();
            groups.back().hset |= bit;
        } else {
            groups.back().dmask |= (1ull << g.q0) | (1ull << g.q1);
            groups.back().diag.push_back((int) i);
        }
        group_of[i] = (int) groups.size() - 1;
    }

    // ---- 2. passes
    const uint64_t local_mask = reg->n_local >= 64 ? ~0ull : ((1ull << reg->n_local) - 1ull);
    for (size_t gi = 0; gi < groups.size(); gi++) {
        uint64_t hl = groups[gi].hset & local_mask, hg = groups[gi].hset & ~local_mask;
        if (hg && reg->peer) {
            // peer memory: the run of H qubits that ends at the top global qubit in the set goes
            // through the sharded sweeps (global sweep on the stitched array + local sweeps)
            unsigned hi = reg->n;
            while (hi > 0 && !((groups[gi].hset >> (hi - 1)) & 1ull)) hi--;
            unsigned lo = hi;
            while (lo > 0 && ((groups[gi].hset >> (lo - 1)) & 1ull)) lo--;
            if (qcs_sharded_sweeps_supported(reg, lo, hi)) {
                pass ps;
                ps.type = 3;
                ps.group = (int) gi;
                ps.run_lo = lo;
                ps.run_hi = hi;
                ps.hmask = ((hi >= 64 ? ~0ull : ((1ull << hi) - 1ull))) & ~((1ull << lo) - 1ull);
                passes.push_back(ps);
                hl &= ~ps.hmask;
                hg &= ~ps.hmask;
            }
        }
        if (hg) {
            // sharded register, Hadamards on global qubits: all p of them at once through the
            // qubit-swap pipeline when that is the whole global set, else pairwise exchanges
            pass ps;
            ps.type = 2;
            ps.group = (int) gi;
            ps.hmask = hg;
            const uint64_t all_global = ((reg->n >= 64 ? ~0ull : ((1ull << reg->n) - 1ull))) & ~local_mask;
            ps.top_stages = hg == all_global && reg->n_local >= 2u * (unsigned) reg->p_global;
            passes.push_back(ps);
        }
        unsigned b = reg->n_local;
        while (b > 0) {
            // contiguous runs of H qubits, top run first (the order the sweep planner works in)
            if (!((hl >> (b - 1)) & 1ull)) { b--; continue; }
            unsigned hi = b;
            while (b > 0 && ((hl >> (b - 1)) & 1ull)) b--;
            const unsigned lo = b;
            if (hi - lo == 1) {
                pass ps;
                ps.type = 1;
                ps.group = (int) gi;
                ps.hmask = 1ull << lo;
                ps.q = lo;
                passes.push_back(ps);
                continue;
            }
            std::vector<sweep_plan> plans;
            QCS_TRY(qcs_plan_hadamard_sweeps(reg, lo, hi, plans));
            for (const sweep_plan &pl : plans) {
                pass ps;
                ps.type = 0;
                ps.group = (int) gi;
                ps.plan = pl;
                ps.hmask = 0;
                for (int k = 0; k < pl.d.n_steps; k++)
                    for (int r = 0; r < pl.d.step[k].r; r++) ps.hmask |= 1ull << (pl.d.step[k].low_phys + r);
                passes.push_back(ps);
            }
        }
    }

    // ---- 3. place every diagonal gate
    for (size_t i = 0; i < queue.size(); i++) {
        if (queue[i].kind != 1) continue;
        const uint64_t bits = (1ull << queue[i].q0) | (1ull << queue[i].q1);
        const int g = group_of[i];
        int earliest = -1, latest = (int) passes.size() - 1;
        for (int p = 0; p < (int) passes.size(); p++) {
            if (!(passes[(size_t) p].hmask & bits)) continue;
            if (passes[(size_t) p].group <= g) earliest = p;
            else { latest = p - 1; break; }
        }
        int best = -1;
        for (int p = earliest < 0 ? 0 : earliest; p <= latest; p++) {
            const pass &ps = passes[(size_t) p];
            if (ps.type != 0 || (int) ps.in_sweep.size() >= kMaxDiagPerSweep) continue;
            if (best < 0 || ps.in_sweep.size() < passes[(size_t) best].in_sweep.size()) best = p;
        }
        if (best >= 0) {
            diag_gate dg;
            if (localise(reg, queue[i], dg)) {
                passes[(size_t) best].in_sweep.push_back(dg);
                passes[(size_t) best].in_sweep_idx.push_back((int) i);
            }
        } else if (earliest < 0) {
            before_all.push_back((int) i);
        } else {
            passes[(size_t) earliest].after.push_back((int) i);
        }
    }
    return QCS_NO_ERROR;
}

// A run of general (controlled) single-qubit gates on qubits 0..3 inside the window -- the generalisation
// of HADAMARD_BASE_MATRIX / C_PHASE_SHIFT_BASE_MATRIX (qc_shor.c:210-225) to arbitrary 2x2 matrices -- is
// multiplied up on the host into one 16 x 16 unitary and applied as ONE dense block on the FP64 tensor cores
// (dense_block.cu) when the run ends: the "k low qubits as one 2^k x 2^k dense contraction" of north_star.
bool qcs_fuse_record_dense(qcs_register *reg, unsigned q, int c, const double *u, int *rc)
{
    *rc = QCS_NO_ERROR;
    if (!reg->fusing || !reg->opt_fusion || !u || q >= 4 || c >= 4 || c == (int) q || reg->n_local < 7) return false;
    // program order: gates recorded before this run are launched first
    if (!reg->queue.empty()) {
        *rc = qcs_fuse_flush(reg);
        if (*rc != QCS_NO_ERROR) return true;
    }
    if (!reg->dense_pending) {
        for (int i = 0; i < 512; i++) reg->dense_acc[i] = 0.0;
        for (int i = 0; i < 16; i++) reg->dense_acc[2 * (i * 16 + i)] = 1.0;
        reg->dense_pending = 1;
        reg->dense_gates = 0;
    }
    // acc <- G acc, G = the gate embedded on index bits 0..3: row i mixes rows i with bit q cleared / set
    double next[512];
    for (int i = 0; i < 16; i++) {
        const bool on = c < 0 || ((i >> c) & 1);
        const int i0 = i & ~(1 << q), i1 = i | (1 << q), b = (i >> q) & 1;
        for (int j = 0; j < 16; j++) {
            double re, im;
            if (!on) {
                re = reg->dense_acc[2 * (i * 16 + j)];
                im = reg->dense_acc[2 * (i * 16 + j) + 1];
            } else {
                const double a_re = u[2 * (2 * b)], a_im = u[2 * (2 * b) + 1];            // U[b][0]
                const double b_re = u[2 * (2 * b + 1)], b_im = u[2 * (2 * b + 1) + 1];    // U[b][1]
                const double x_re = reg->dense_acc[2 * (i0 * 16 + j)], x_im = reg->dense_acc[2 * (i0 * 16 + j) + 1];
                const double y_re = reg->dense_acc[2 * (i1 * 16 + j)], y_im = reg->dense_acc[2 * (i1 * 16 + j) + 1];
                re = a_re * x_re - a_im * x_im + b_re * y_re - b_im * y_im;
                im = a_re * x_im + a_im * x_re + b_re * y_im + b_im * y_re;
            }
            next[2 * (i * 16 + j)] = re;
            next[2 * (i * 16 + j) + 1] = im;
        }
    }
    for (int i = 0; i < 512; i++) reg->dense_acc[i] = next[i];
    reg->dense_gates++;
    return true;
}

int qcs_fuse_flush(qcs_register *reg)
{
    if (reg->lazy_reset) QCS_TRY(qcs_materialise_reset(reg));        // a deferred reset_register comes first
    if (reg->dense_pending) {
        // never both pending: recording into one launches the other first
        reg->dense_pending = 0;
        return qcs_k_dense_block(reg, 4, reg->dense_acc);
    }
    if (reg->queue.empty()) return QCS_NO_ERROR;
    std::vector<qcs_pending_gate> queue;
    queue.swap(reg->queue);             // the per-gate calls below must not see a pending queue
    std::vector<pass> passes;
    std::vector<int> before_all;        // diagonal gates no pass may precede
    QCS_TRY(schedule_stream(reg, queue, passes, before_all));

    // ---- 4. launch
    size_t need = kDiagScratchSlots;
    for (const pass &ps : passes) need += ps.in_sweep.size();
    QCS_TRY(ensure_diag_capacity(reg, need));
    diag_gate *d_all = (diag_gate *) reg->d_diag;
    // the array may still be read by the sweeps of the previous flush: reuse is stream-ordered
    // (the copies below are issued on the same stream as those sweeps)
    std::vector<diag_gate> stage;
    diag_gate *d_scratch = d_all;       // first kDiagScratchSlots slots: standalone diagonal passes
    size_t at = kDiagScratchSlots;
    {
        std::vector<diag_gate> all;
        for (const pass &ps : passes) all.insert(all.end(), ps.in_sweep.begin(), ps.in_sweep.end());
        if (!all.empty())
            QCS_CUDA(cudaMemcpyAsync(d_all + at, all.data(), all.size() * sizeof(diag_gate), cudaMemcpyHostToDevice,
                                     reg->stream));
    }
    QCS_TRY(launch_diag_list(reg, queue, before_all, d_scratch, stage));
    for (pass &ps : passes) {
        if (ps.type == 0) {
            ps.plan.d.n_diag = (int) ps.in_sweep.size();
            ps.plan.d.diag = ps.in_sweep.empty() ? nullptr : d_all + at;
            at += ps.in_sweep.size();
        }
    }
    for (size_t k = 0; k < passes.size(); k++) {
        pass &ps = passes[k];
        if (ps.type == 0) {
            // two neighbouring sweeps with nothing in between may share an L2-paired launch
            if (k + 1 < passes.size() && passes[k + 1].type == 0 && ps.after.empty()) {
                bool paired = false;
                QCS_TRY(qcs_launch_sweep_pair(reg, ps.plan, passes[k + 1].plan, &paired));
                if (paired) {
                    k++;
                    QCS_TRY(launch_diag_list(reg, queue, passes[k].after, d_scratch, stage));
                    continue;
                }
            }
            QCS_TRY(qcs_launch_sweep_plan(reg, ps.plan));
        } else if (ps.type == 1) {
            QCS_TRY(qcs_k_hadamard_local(reg, ps.q));
        } else if (ps.type == 3) {
            QCS_TRY(qcs_fused_sweeps_sharded(reg, ps.run_lo, ps.run_hi, true, true));
        } else if (ps.top_stages) {
            QCS_TRY(qcs_dist_top_stages(reg, 0, true, true));
        } else {
            for (unsigned q = reg->n_local; q < reg->n; q++)
                if ((ps.hmask >> q) & 1ull)
                    QCS_TRY(reg->peer ? qcs_k_hadamard_peer(reg, q) : qcs_dist_hadamard_global(reg, q));
        }
        QCS_TRY(launch_diag_list(reg, queue, ps.after, d_scratch, stage));
    }
    return QCS_NO_ERROR;
}

// Host-only view of the scheduler for tests and tooling: the passes a recorded stream would be
// launched as, one line each, for a register of n_qubits sharded over world_size ranks
// (library-default options).  "sweep h=<hex mask of H qubits> diag=<stream positions>",
// "hadamard q=<qubit>", "global h=<mask>", each optionally followed by " after=<positions>" (diagonal
// gates run by standalone kernels after the pass); a leading "before=<positions>" line lists
// diagonal gates that precede every pass.  Gates that are the identity on this rank are omitted.
extern "C" int qcs_schedule_describe(unsigned n_qubits, int world_size, int rank, unsigned long long n_gates,
                                     const int *kinds, const unsigned *q0, const unsigned *q1, char *out,
                                     unsigned long long out_cap)
{
    if (!out || out_cap == 0 || (n_gates && (!kinds || !q0 || !q1))) return QCS_BAD_ARGUMENTS;
    if (n_qubits < 1 || n_qubits > 62 || world_size < 1 || (world_size & (world_size - 1)) || rank < 0 || rank >= world_size)
        return QCS_BAD_ARGUMENTS;
    int p = 0;
    while ((1 << p) < world_size) p++;
    if ((unsigned) p >= n_qubits) return QCS_BAD_ARGUMENTS;
    qcs_register fake = {};
    fake.n = n_qubits;
    fake.n_local = n_qubits - (unsigned) p;
    fake.N = 1ull << fake.n;
    fake.N_local = 1ull << fake.n_local;
    fake.rank = rank;
    fake.world = world_size;
    fake.p_global = p;
    fake.opt_fusion = 1;
    fake.opt_pipeline = 1;
    fake.opt_pipe_shape = -1;
    fake.opt_min_run_bits = 3;
    std::vector<qcs_pending_gate> queue;
    for (unsigned long long i = 0; i < n_gates; i++) {
        if (q0[i] >= n_qubits || (kinds[i] == 1 && q1[i] >= n_qubits) || (kinds[i] != 0 && kinds[i] != 1)) return QCS_BAD_ARGUMENTS;
        queue.push_back({kinds[i], q0[i], kinds[i] == 1 ? q1[i] : q0[i], 1.0, 0.0});
    }
    std::vector<pass> passes;
    std::vector<int> before_all;
    QCS_TRY(schedule_stream(&fake, queue, passes, before_all));
    std::string text;
    auto list = [&](const char *key, const std::vector<int> &v) {
        if (v.empty()) return;
        text += key;
        for (size_t k = 0; k < v.size(); k++) text += (k ? "," : "") + std::to_string(v[k]);
    };
    if (!before_all.empty()) { list("before=", before_all); text += "\n"; }
    for (const pass &ps : passes) {
        char buf[64];
        if (ps.type == 0) snprintf(buf, sizeof buf, "sweep h=%llx", (unsigned long long) ps.hmask);
        else if (ps.type == 1) snprintf(buf, sizeof buf, "hadamard q=%u", ps.q);
        else snprintf(buf, sizeof buf, "global h=%llx", (unsigned long long) ps.hmask);
        text += buf;
        list(" diag=", ps.in_sweep_idx);
        list(" after=", ps.after);
        text += "\n";
    }
    if (text.size() + 1 > out_cap) return QCS_INSUFFICIENT_MEMORY;
    memcpy(out, text.c_str(), text.size() + 1);
    return QCS_NO_ERROR;
}

extern "C" int qcs_fuse_begin(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_fuse_begin(m));
    if (!reg) return QCS_BAD_ARGUMENTS;
    if (reg->fusing) return QCS_BAD_ARGUMENTS;
    reg->fusing = 1;
    return QCS_NO_ERROR;
}

extern "C" int qcs_fuse_end(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_fuse_end(m));
    if (!reg) return QCS_BAD_ARGUMENTS;
    if (!reg->fusing) return QCS_BAD_ARGUMENTS;
    reg->fusing = 0;
    QCS_CUDA(cudaSetDevice(reg->device));
    return qcs_fuse_flush(reg);
}

extern "C" unsigned long long qcs_fuse_pending(const qcs_register *reg)
{
    if (reg && reg->group) return qcs_fuse_pending(qcs_group_member(reg, 0));
    return reg ? (unsigned long long) reg->queue.size() + (reg->dense_pending ? reg->dense_gates : 0u) : 0ull;
}
