// qcs_internal.h -- shared declarations of libqcs.so (not installed).
//
// The register is one in-place array of 2^n_local complex doubles in HBM
// (double2 = 16 B, interleaved re/im: the gsl_vector_complex.data layout of
// the reference, qc_shor.c:385-386).  All indices are 64-bit.
#pragma once

#include <cuda_runtime.h>
#include <functional>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../../include/qcs.h"

struct qcs_dist;   // multi-GPU state (dist.cu)
struct qcs_group;  // single-caller facade over the shards of all GPUs of the box (group.cu)
struct qcs_peer;   // stitched peer-memory mapping of all shards (peer.cu)

struct qcs_profile_slot {
    cudaEvent_t begin, end;
    int kind;
};

// a gate recorded between qcs_fuse_begin and qcs_fuse_end (circuit.cu)
struct qcs_pending_gate {
    int kind;               // 0: hadamard_gate(q0), 1: c_phase_shift_gate(q0, q1, theta)
    unsigned q0, q1;
    double c, s;            // cos(theta), sin(theta) as the reference forms them (qc_shor.c:526)
};

struct qcs_register {
    int L_size, M_size;
    unsigned n;             // total qubits
    unsigned n_local;       // qubits addressed inside this shard
    uint64_t N;             // 2^n
    uint64_t N_local;       // 2^n_local
    int device;
    cudaStream_t stream;
    cudaStream_t launch_stream;   // stream the next accounted launch goes to (nullptr: `stream`)
    double2 *amp;           // this shard, in place

    // small device/host scratch
    double *d_partials;     // reduction partials
    size_t partials_cap;    // in doubles
    void *d_small;          // 4 KiB device scratch (measurement result etc.)
    void *h_small;          // 4 KiB pinned host mirror
    void *d_meas;           // chunk summaries of the exact parallel measurement scan (lazy)
    void *d_dense;          // A-fragment buffer of qcs_apply_dense_block (lazy, 8 KiB)
    // quantum_computation from the reset state: f(x) of every block x (built gate by gate), and the order for
    // the NEXT pipelined sweep launch to generate its tiles from it instead of loading them (qft_pipeline.cu)
    unsigned *d_gen_table;
    size_t gen_table_cap;   // entries
    struct { const unsigned *f; unsigned M; double value; int armed; } gen;
    void *d_pair;           // ticket + per-block counters of an L2-paired sweep launch (lazy)
    size_t d_pair_cap;      // bytes

    // options
    int opt_fusion;
    int opt_profile;
    int opt_tile_bits;
    int opt_prefetch_tiles;       // accepted, ignored (the L2 prefetch was measured slower and removed)
    int opt_pipeline;             // 1: TMA/mbarrier pipelined sweep kernel where it applies
    int opt_pipe_shape;           // which instantiated pipeline shape (qft_pipeline.cu kShapes)
    int opt_overlap_slices;       // sharded QFT: the global sweep runs in this many slices, the local strided
                                  // sweeps of a finished slice overlapping the next one (0 / 1: no overlap)
    int opt_global_sms;           // SMs given to the (NVLink-bound) global sweep while local sweeps run beside it
    int opt_global_run_bits;      // log2 of the contiguous run of the sweep over the global qubits (peer memory)
    int opt_min_run_bits;         // log2 of the shortest contiguous run (amplitudes) a strided tile may use
    int opt_measure_sequential;   // 1: always use the single-CTA sequential scan
    int opt_l2_pair;              // 1: the last strided sweep and the contiguous sweep of a transform share one launch
                                  // whose intermediate state stays in L2 (qft_pipeline.cu)
    int opt_gen_sweep;            // 1: quantum_computation from reset lets the first inverse-QFT sweep generate its tiles
    int opt_split3;               // 1: contiguous 2^12 tiles of radix-16 steps use the conflict-free split-3 layout
    int opt_l2_pair_hints;        // 1: L2 eviction-priority hints on the TMA traffic of a paired launch
    int opt_l2_pair_lag;          // tiles the second sweep of a pair trails the first by, beyond one block
    long long opt_l2_pair_max_block;   // largest block (bytes) a paired launch may use

    // qcs_reset_register is deferred (fused mode): the all-zero-but-one state is only written when something
    // needs it, so that quantum_computation from the reset state (qc_shor.c:922-923, every find_period) can write
    // the state after the Hadamards and the modular exponentiation in closed form instead (modexp_fused.cu)
    int lazy_reset;
    uint64_t lazy_index;          // the pending basis state (1 after reset_register; the measured index after measure_state)

    // deferred gate stream (qcs_fuse_begin .. qcs_fuse_end)
    int fusing;
    std::vector<qcs_pending_gate> queue;
    // a run of general (controlled) single-qubit gates on qubits 0..3 recorded in the window is kept as ONE
    // 16 x 16 matrix (row-major, interleaved) and leaves as one DMMA dense block (dense_block.cu)
    int dense_pending;
    unsigned dense_gates;         // gates folded into it
    double dense_acc[512];
    void *d_diag;                 // device array of qft::diag_gate for the sweeps in flight
    size_t d_diag_cap;            // in gates

    // accounting
    unsigned long long launches_total;
    unsigned long long launches[QCS_K_COUNT];
    double alg_bytes[QCS_K_COUNT];
    double ms[QCS_K_COUNT];
    std::vector<qcs_profile_slot> pending;     // event pairs not yet resolved
    std::vector<qcs_profile_slot> free_slots;
    cudaEvent_t timer_begin, timer_end;

    int sm_count;
    size_t smem_optin;      // max dynamic shared memory per block

    // sharding: rank holds amplitudes whose top log2(world) index bits == rank
    int rank, world, p_global;
    qcs_dist *dist;
    // peer memory: when non-null, amp_all[i] addresses basis state i of the WHOLE register on
    // every rank (shard k is mapped at k * N_local; remote shards travel over NVLink) and
    // amp == amp_all + rank * N_local
    qcs_peer *peer;
    double2 *amp_all;

    // non-null: this handle is the facade of qcs_register_create_multi; it owns no device memory,
    // every call is forwarded to the per-device shard registers (group.cu)
    qcs_group *group;
};

// ---- single-caller multi-GPU facade: group.cu -------------------------------
// run `job` on every shard register of the facade at once (one worker thread per device);
// returns the first non-zero code in rank order
int qcs_group_run(qcs_register *facade, const std::function<int(qcs_register *)> &job);
int qcs_group_world(const qcs_register *facade);
qcs_register *qcs_group_member(const qcs_register *facade, int rank);
void qcs_group_destroy(qcs_register *facade);
// first line of every entry point that works shard by shard
#define QCS_GROUP_FORWARD(reg, call)                                                              \
    do {                                                                                          \
        if ((reg) && (reg)->group)                                                                \
            return qcs_group_run((reg), [&](qcs_register *m) -> int { return (call); });          \
    } while (0)

// ---- error plumbing -------------------------------------------------------
int qcs_map_cuda_error(cudaError_t e, const char *what, const char *file, int line);
#define QCS_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t qcs_e_ = (call);                                                \
        if (qcs_e_ != cudaSuccess) return qcs_map_cuda_error(qcs_e_, #call, __FILE__, __LINE__); \
    } while (0)
#define QCS_TRY(call)                                \
    do {                                             \
        int qcs_rc_ = (call);                        \
        if (qcs_rc_ != QCS_NO_ERROR) return qcs_rc_; \
    } while (0)

// ---- launch accounting ----------------------------------------------------
// Every kernel launch of the library goes through these two calls so that
// launches are counted and, with QCS_OPT_PROFILE, timed on the launching stream.
void qcs_launch_begin(qcs_register *reg, int kind, double algorithmic_bytes);
int qcs_launch_end(qcs_register *reg, int kind, const char *name);
int qcs_profile_resolve(qcs_register *reg);

// ---- device-side index helpers --------------------------------------------
__host__ __device__ __forceinline__ uint64_t qcs_insert_zero_bit(uint64_t x, unsigned pos)
{
    const uint64_t low = (1ull << pos) - 1ull;
    return ((x & ~low) << 1) | (x & low);
}

// ---- per-gate (reference-order) kernels: gates_exact.cu -------------------
int qcs_k_reset(qcs_register *reg);
// write a deferred reset_register now (no-op when none is pending)
int qcs_materialise_reset(qcs_register *reg);
int qcs_k_collapse(qcs_register *reg, uint64_t local_index, bool owner);
int qcs_k_fill_synthetic(qcs_register *reg, uint64_t seed);
int qcs_k_scale(qcs_register *reg, double s);
int qcs_k_hadamard_local(qcs_register *reg, unsigned q);
int qcs_k_hadamard_peer(qcs_register *reg, unsigned q);     // global qubit, register with peer memory
// multiply by (c + i s) every local amplitude whose bits in `mask_bits` are all 1
int qcs_k_phase_masked(qcs_register *reg, int nbits, unsigned b0, unsigned b1, double c, double s);
int qcs_k_amodc(qcs_register *reg, unsigned C, unsigned A, int ctrl_local /* -1: always on */,
                bool ctrl_off);

// ---- reductions / measurement: measure.cu ---------------------------------
int qcs_k_norm2_local(qcs_register *reg, double *out_host);
// sequential-semantics scan of this shard starting from `cum_in`; *found / *index
// as in measure_state (qc_shor.c:283-292) restricted to [0, limit)
int qcs_k_measure_scan(qcs_register *reg, double cum_in, double r, uint64_t limit,
                       int *found, uint64_t *index, double *cum_out);
// the same scan in three parts, so that the shards of a sharded register can do everything except
// the walk at the same time: (1) chunk sums and their total (approximate, for bounds only),
// (2) classification + chunk / super-chunk maps given an APPROXIMATE running sum before the shard,
// (3) the exact walk from the EXACT running sum before the shard.  *bad: an invariant failed, the
// caller falls back to qcs_k_measure_scan.  false from qcs_k_scan_parallel_ok: use the plain scan.
bool qcs_k_scan_parallel_ok(const qcs_register *reg, uint64_t limit);
int qcs_k_scan_sums(qcs_register *reg, uint64_t limit, double *approx_total);
int qcs_k_scan_maps(qcs_register *reg, double approx_cum_in, double r, uint64_t limit);
int qcs_k_scan_walk(qcs_register *reg, double cum_in, double r, uint64_t limit, int *found, uint64_t *index,
                    double *cum_out, int *bad);

// many variates against one state; *handled = false: not applicable, use one scan per variate
int qcs_k_sample_many(qcs_register *reg, uint64_t n_shots, const double *r, unsigned long long *indices, bool *handled);

// ---- fused sweeps: qft_fused.cu / modexp_fused.cu ---------------------------
int qcs_fused_qft(qcs_register *reg, unsigned lo, unsigned hi, bool inverse);
// the same on a sharded register with peer memory: qubits up to n, sweeps whose tile holds
// global qubits run on the stitched array, every rank taking its share of the tiles
int qcs_fused_sweeps_sharded(qcs_register *reg, unsigned lo, unsigned hi, bool inverse, bool hadamard_only);
bool qcs_sharded_sweeps_supported(const qcs_register *reg, unsigned lo, unsigned hi);
int qcs_fused_hadamards(qcs_register *reg, unsigned lo, unsigned hi);
int qcs_fused_top_sweep(qcs_register *reg, double2 *buf, unsigned c, unsigned p, unsigned lo,
                        unsigned long long y_const, bool inverse, bool hadamard_only, cudaStream_t stream);
// H on the L register, then all L controlled a^(2^k) mod C gates in one block-local sweep
int qcs_fused_modexp(qcs_register *reg, unsigned C, const unsigned *A_per_gate, unsigned n_gates);
// the same two loops applied to the reset state |0...01> in closed form: one write pass.  *done = false: the
// shape is not covered (caller writes the reset state and takes the general path)
int qcs_shor_state_from_reset(qcs_register *reg, unsigned C, const unsigned *A_per_gate, unsigned n_gates, bool *done);
// the same state, not written: arms the generating first sweep of the inverse QFT on [M, n) when the launch
// plan allows it (*armed tells; qcs_fused_gen_supported, qft_fused.cu)
int qcs_shor_state_generated(qcs_register *reg, unsigned C, const unsigned *A_per_gate, unsigned n_gates, bool *armed);
bool qcs_fused_gen_supported(const qcs_register *reg, unsigned lo, unsigned hi);

// ---- general gates: gates_general.cu ------------------------------------------
int qcs_k_gate_1q(qcs_register *reg, unsigned q, int c /* < 0: no control */, const double *u_interleaved);

// ---- dense block on the low qubits: dense_block.cu
int qcs_k_dense_block(qcs_register *reg, unsigned k, const double *u_interleaved);

// ---- deferred gate stream: circuit.cu ------------------------------------------
// schedule and launch every recorded gate (no-op when the queue is empty)
int qcs_fuse_flush(qcs_register *reg);
// record u (2x2, row-major interleaved) on qubit q < 4, control c < 4 or < 0, into the pending dense block;
// false: not recordable here (not fusing, qubits too high, register too small): apply it the ordinary way
bool qcs_fuse_record_dense(qcs_register *reg, unsigned q, int c, const double *u, int *rc);

// ---- multi-GPU: dist.cu ---------------------------------------------------
int qcs_dist_init(qcs_register *reg, const void *comm_id);
void qcs_dist_destroy(qcs_register *reg);
int qcs_dist_hadamard_global(qcs_register *reg, unsigned q);
// arbitrary 2x2 gate u (row-major, interleaved) on a global target qubit, control c (< 0: none)
int qcs_dist_gate_global(qcs_register *reg, unsigned q, int c, const double *u);
// stages of the (inverse) QFT / Hadamards on the global qubits [n_local, n): exchange in,
// one sweep, exchange back, pipelined over slices of the shard
int qcs_dist_top_stages(qcs_register *reg, unsigned lo, bool inverse, bool hadamard_only);
int qcs_dist_allgather_double(qcs_register *reg, double mine, double *all_host);
// `count` <= 4 doubles per rank: all_host[r * count + i]
int qcs_dist_allgather_doubles(qcs_register *reg, const double *mine, int count, double *all_host);
int qcs_dist_barrier(qcs_register *reg);
// cross-rank barrier ordered on the register's stream (no host synchronisation): work queued
// after it on any rank starts only when the work queued before it has finished on every rank
int qcs_dist_stream_barrier(qcs_register *reg);
int qcs_dist_barrier_on(qcs_register *reg, cudaStream_t stream);
cudaStream_t qcs_dist_side_stream(qcs_register *reg);       // second stream of a sharded register
int qcs_dist_slice_event(qcs_register *reg, int j, cudaEvent_t *ev);

// ---- peer memory: peer.cu ----------------------------------------------------
bool qcs_peer_can(const qcs_register *reg);
bool qcs_peer_try_alloc(qcs_register *reg, const void *comm_id);
void qcs_peer_free(qcs_register *reg);
