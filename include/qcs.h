/*
 * qcs.h -- C ABI of libqcs.so, the B200 (sm_100a) state-vector engine that
 * replaces the gate-application path of adamalderton/QuantumComputer's
 * qc_shor.c.  "Q:a-b" below cites /root/reference/qc_shor.c lines a-b and
 * "T:a-b" cites testing_and_debug.c.
 *
 * The reference has no header and no FFI: every function is `static` in one
 * translation unit.  The boundary is therefore the set of calls the classical
 * driver makes into the gate path (find_period, Q:922-928; main, Q:1316-1333)
 * plus the individual gate operators, with the reference's argument order.
 * The `gsl_spmatrix_complex *matrix` scratch argument of the reference is
 * dropped: no gate matrix is ever materialised here.
 *
 * Conventions (identical to the reference):
 *   - amplitudes are complex double, interleaved (re, im), 16 bytes each, the
 *     layout of gsl_vector_complex.data (Q:385-386, 405-412);
 *   - qubit b is bit b of the basis-state index, least significant = qubit 0
 *     (GET_BIT, Q:150-151);
 *   - the M ("f") register is qubits 0..M-1, the L ("x") register is qubits
 *     M..L+M-1 (Q:620, 650, 720).
 *
 * Every entry point returns one of the reference's ErrorCode values (Q:164-170)
 * as an int.  Gate calls are asynchronous on the register's CUDA stream;
 * qcs_measure_state, qcs_norm2, qcs_get_state and qcs_synchronize block.
 * One host thread per register.  There is no CPU fallback: without a usable
 * CUDA device qcs_register_create fails with QCS_UNKNOWN_ERROR.
 */
#ifndef QCS_H
#define QCS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ErrorCode, Q:164-170 */
enum {
    QCS_NO_ERROR = 0,
    QCS_INSUFFICIENT_MEMORY = 1,
    QCS_BAD_ARGUMENTS = 2,
    QCS_PERIOD_NOT_FOUND = 3,
    QCS_UNKNOWN_ERROR = 4
};

/* Opaque device register: replaces `Register` (Q:194-203).  One in-place
 * amplitude array per GPU; the state_a/state_b ping-pong and swap_states
 * (Q:242-249) do not exist. */
typedef struct qcs_register qcs_register;

/* how quantum_computation derives `atox` for gate k of the modular
 * exponentiation (Q:728-731) */
enum {
    QCS_POW_VERBATIM = 0,   /* INT_POW(a, 2^k) exactly as Q:158-159 behaves on x86-64 */
    QCS_POW_MODULAR = 1     /* a^(2^k) mod C by modular squaring (the intended value) */
};

/* options for qcs_set_option */
enum {
    /* 1 (default): composite operators (inverse_QFT, quantum_computation) run
     * as fused tile sweeps -- amplitudes agree with the reference to <= 1e-12
     * relative L2.  0: they run gate by gate in the reference's order with the
     * reference's floating-point operations -- amplitudes are value-identical. */
    QCS_OPT_FUSION = 1,
    /* 1: record a CUDA-event pair around every kernel launch so that
     * qcs_profile_get can report per-kernel device time.  Default 0. */
    QCS_OPT_PROFILE = 2,
    /* log2 of the shared-memory tile (amplitudes) used by the fused sweeps;
     * 0 = library default. */
    QCS_OPT_TILE_BITS = 3,
    /* 1: measure_state always uses the single-CTA sequential scan; 0 (default):
     * registers of >= 2^17 amplitudes use the parallel scan that reproduces the
     * sequential rounding exactly (same index, see csrc/measure.cu). */
    QCS_OPT_MEASURE_SEQUENTIAL = 4,
    /* 1 (default): fused sweeps use the TMA + mbarrier pipelined kernel where it
     * applies (tile = 2^12 amplitudes by default); 0: the direct global<->register kernel. */
    QCS_OPT_PIPELINE = 5,
    /* accepted and ignored: the L2 prefetch ahead of the TMA ring was measured slower on B200 in both
     * rounds (profiles/README.md) and its code is gone */
    QCS_OPT_PREFETCH_TILES = 6,
    /* which instantiated shape of the pipelined sweep runs (tile size, ring depth, consumer
     * groups; csrc/qft_pipeline.cu kShapes).  Tuning knob; -1 = library default. */
    QCS_OPT_PIPE_SHAPE = 7,
    /* log2 of the shortest contiguous run of amplitudes a strided tile may use (3 = 128 B,
     * the default; 4 = 256 B ...): shorter runs leave more tile bits for stages, i.e. fewer
     * sweeps. */
    QCS_OPT_MIN_RUN_BITS = 8,
    /* sharded registers with peer memory: log2 of the contiguous run (amplitudes per TMA row) of
     * the one sweep that covers the global qubits.  That sweep is NVLink-bound, so it prefers
     * long rows to many stage bits. */
    QCS_OPT_GLOBAL_RUN_BITS = 9,
    /* sharded inverse QFT with peer memory: the sweep over the global qubits runs in this many
     * slices (a power of two <= 8, default 8) on a second stream and a subset of the SMs
     * (QCS_OPT_GLOBAL_SMS, default 48) while the strided local sweeps of the slices that are
     * already complete run on the other SMs: the NVLink exchange overlaps the HBM-bound local
     * work.  0 or 1: one after the other. */
    QCS_OPT_OVERLAP_SLICES = 10,
    QCS_OPT_GLOBAL_SMS = 11,
    /* 1 (default): the last strided sweep and the contiguous sweep of a fused transform run as ONE
     * launch over blocks of <= QCS_OPT_L2_PAIR_MAX_BLOCK bytes (default 16 MiB) whose intermediate
     * state stays in the 126 MB L2: the pair reads and writes HBM once.  The second sweep trails the
     * first by one block plus QCS_OPT_L2_PAIR_LAG tiles (default 444 = 3 tiles per SM). */
    QCS_OPT_L2_PAIR = 12,
    QCS_OPT_L2_PAIR_LAG = 13,
    QCS_OPT_L2_PAIR_MAX_BLOCK = 14,
    /* 1 (default): the TMA loads and stores of a paired launch carry L2 eviction-priority hints
     * (evict_last for what the second sweep reads, evict_first for the rest) */
    QCS_OPT_L2_PAIR_HINTS = 15,
    /* 16: reserved (direct stores from registers were measured slower and removed) */
    /* 1 (default): a contiguous 2^12 tile whose steps are radix-16 at bits 8, 4, 0 is held in the
     * split-3 shared-memory layout (bank-conflict free; qft_common.cuh); 0: plain 128-byte swizzle */
    QCS_OPT_SPLIT3 = 17,
    /* 1 (default): quantum_computation right after reset_register does not write the state |x, a^x mod C>
     * at all -- the first sweep of the inverse QFT builds each tile in shared memory from a table of
     * f(x) and skips the tiles that hold no non-zero amplitude (single GPU, when the transform starts
     * with a strided sweep of the pipelined kernel; otherwise, and with 0, the state is written in one
     * pass first) */
    QCS_OPT_GEN_SWEEP = 18
};

/* kernel classes reported by qcs_profile_get */
enum {
    QCS_K_HADAMARD = 0,      /* pair-stride H                       32*2^n B/launch */
    QCS_K_CPHASE = 1,        /* controlled phase, |11> quarter       8*2^n B/launch */
    QCS_K_AMODC = 2,         /* controlled a^x mod C permutation                     */
    QCS_K_FILL = 3,          /* reset / synthetic fill / collapse    16*2^n B/launch */
    QCS_K_REDUCE = 4,        /* norm^2, measurement scan             16*2^n B/launch */
    QCS_K_TILE_SWEEP = 5,    /* fused QFT tile sweep                 32*2^n B/launch */
    QCS_K_MODEXP_SWEEP = 6,  /* fused H^L + all L a^x mod C gates                    */
    QCS_K_EXCHANGE = 7,      /* global-qubit exchange (multi-GPU)                    */
    QCS_K_SCALE = 8,         /* in-place scaling                     32*2^n B/launch */
    QCS_K_DENSE_BLOCK = 9,   /* dense 2^k x 2^k block on the k low qubits (DMMA) 32*2^n B/launch */
    QCS_K_DIAG = 10,         /* several diagonal gates in one pass (gate stream) 32*2^n B/launch */
    QCS_K_GLOBAL_SWEEP = 11, /* tile sweep over the global qubits on peer memory (multi-GPU); bytes =
                              * NVLink traffic per direction, 2*(P-1)/P * 16*2^n_local B/launch */
    QCS_K_GATE_1Q = 12,      /* arbitrary (controlled) single-qubit gate  32*2^n B/launch (16*2^n controlled) */
    QCS_K_COUNT = 13
};

const char *qcs_version(void);
const char *qcs_error_string(int code);
/* number of CUDA devices visible to this process (0 if none) */
int qcs_device_count(void);

/* ---- register life-cycle ------------------------------------------------
 * replaces the allocation block of main(), Q:1316-1324 (two state vectors and
 * a COO matrix with nzmax = 2N) and the frees at Q:1330-1333.  num_qubits =
 * L_size + M_size and num_states = 2^num_qubits as in Q:1255-1261.
 * `device` is a CUDA device ordinal, or -1 for the current device.          */
int qcs_register_create(qcs_register **out, int L_size, int M_size, int device);
void qcs_register_destroy(qcs_register *reg);

/* Sharded register: this process holds shard `rank` of `world_size` = 2^p
 * equal shards; the top p qubits are global.  `comm_id` is the
 * 128-byte id produced by qcs_comm_unique_id on rank 0 and distributed by the
 * launcher (torch.distributed, MPI, a file ...).  One id per sharded register:
 * an id cannot be reused for a second communicator.                        */
#define QCS_COMM_ID_BYTES 128
int qcs_comm_unique_id(void *id_out);
int qcs_register_create_sharded(qcs_register **out, int L_size, int M_size, int device,
                                int rank, int world_size, const void *comm_id);

/* The same sharded register for ONE caller: the allocation block of main(), Q:1316-1324, for a
 * state spread over n_gpus = 2^p devices of the box (devices 0 .. n_gpus-1), no launcher and no
 * communicator id.  The handle is used exactly like a single-GPU register -- find_period's
 * reset_register / quantum_computation / measure_state (Q:922-928) run sharded unchanged; every
 * call is executed by one worker thread per GPU inside the library.  Bulk calls (qcs_get_state,
 * qcs_set_state, qcs_nonzero_states) address the whole register; qcs_local_states == qcs_num_states,
 * qcs_world_size == 1, qcs_num_gpus == n_gpus; profile counters are per GPU (shard 0's), the
 * stopwatch is the maximum over the GPUs.  n_gpus == 1 is qcs_register_create on the current device. */
int qcs_register_create_multi(qcs_register **out, int L_size, int M_size, int n_gpus);
int qcs_num_gpus(const qcs_register *reg);

int qcs_L_size(const qcs_register *reg);
int qcs_M_size(const qcs_register *reg);
unsigned qcs_num_qubits(const qcs_register *reg);
unsigned long long qcs_num_states(const qcs_register *reg);        /* 2^n, all shards */
unsigned long long qcs_local_states(const qcs_register *reg);      /* this shard      */
int qcs_rank(const qcs_register *reg);
int qcs_world_size(const qcs_register *reg);
/* 1 when the shards of a sharded register are mapped into one virtual address range on every
 * rank (CUDA virtual memory management + NVLink peer access, csrc/peer.cu): sweeps over global
 * qubits then read and write the peers' shards directly instead of exchanging through staging
 * buffers.  0: private shards, NCCL send/recv exchange (csrc/dist.cu). */
int qcs_peer_memory(const qcs_register *reg);

int qcs_set_option(qcs_register *reg, int option, long long value);
long long qcs_get_option(const qcs_register *reg, int option);
int qcs_synchronize(qcs_register *reg);

/* ---- gate path ---------------------------------------------------------- */

/* void reset_register(Register), Q:318-324: state <- |0...01> (index 1).  With QCS_OPT_FUSION = 1 the
 * write is deferred until something needs the state; qcs_quantum_computation right after it (the
 * sequence of find_period, Q:922-923) never needs it: it writes the state after the Hadamards and the
 * controlled multiplications in closed form.  Not observable through the ABI. */
int qcs_reset_register(qcs_register *reg);

/* void hadamard_gate(unsigned qubit_num, Register*, matrix*), Q:442-484. */
int qcs_hadamard_gate(qcs_register *reg, unsigned qubit_num);

/* void c_phase_shift_gate(unsigned c_qubit_num, unsigned qubit_num,
 *                         double theta, Register*, matrix*), Q:513-565. */
int qcs_c_phase_shift_gate(qcs_register *reg, unsigned c_qubit_num, unsigned qubit_num,
                           double theta);

/* void c_amodc_gate(unsigned C, unsigned long long atox, unsigned c_qubit_num,
 *                   Register*, matrix*), Q:595-660.  Bit-exact index map;
 * non-bijective maps (gcd(atox % C, C) != 1, atox % C == 0) sum colliding
 * sources in ascending source order like operate_matrix does. */
int qcs_c_amodc_gate(qcs_register *reg, unsigned C, unsigned long long atox,
                     unsigned c_qubit_num);

/* void inverse_QFT(Register*, matrix*), Q:678-690, on the L register. */
int qcs_inverse_QFT(qcs_register *reg);
/* the adjoint circuit (not in the reference): same gates, reverse order, -theta */
int qcs_QFT(qcs_register *reg);
/* the same circuits on qubits lo..hi-1 (lo plays the role of M_size) */
int qcs_inverse_QFT_range(qcs_register *reg, unsigned lo, unsigned hi);
int qcs_QFT_range(qcs_register *reg, unsigned lo, unsigned hi);

/* void quantum_computation(unsigned C, unsigned a, Register*, matrix*),
 * Q:712-737: H on the L register, L controlled a^(2^k) mod C gates, inverse QFT. */
int qcs_quantum_computation(qcs_register *reg, unsigned C, unsigned a, int pow_mode);

/* unsigned long measure_state(Register, gsl_rng*), Q:272-306.  The caller
 * draws r = gsl_rng_uniform(rng) on the host (Q:281); the scan, the `>=`
 * comparison, the N-1 fall-through and the collapse are done on the device
 * with the reference's sequential summation semantics (bit-exact index). */
int qcs_measure_state(qcs_register *reg, double r, unsigned long long *state_num);

/* Sampling without collapse (not in the reference, which reruns the whole computation per
 * sample; Q:294-301 mentions the option): state_nums[k] = the index qcs_measure_state would
 * return for r[k], the state is left untouched. */
int qcs_sample_states(qcs_register *reg, unsigned long long n_shots, const double *r,
                      unsigned long long *state_nums);

/* check_normalisation, T:28-37: sum of |amp|^2 (deterministic parallel order) */
int qcs_norm2(qcs_register *reg, double *sum_of_sq);

/* display_state, T:7-26: indices and |amp| of the non-zero amplitudes, in
 * index order; at most `capacity` are written, *count receives the total. */
int qcs_nonzero_states(qcs_register *reg, unsigned long long capacity,
                       unsigned long long *indices, double *abs_values,
                       unsigned long long *count);

/* bulk access to this shard's amplitudes [first, first+count), interleaved
 * doubles, host memory (pinned memory makes the copies asynchronous-capable) */
int qcs_get_state(qcs_register *reg, unsigned long long first, unsigned long long count,
                  double *interleaved_out);
int qcs_set_state(qcs_register *reg, unsigned long long first, unsigned long long count,
                  const double *interleaved_in);
/* qcs_set_state returns when the copy is done.  The _async form is stream-ordered: the source
 * (pinned, qcs_host_alloc) must stay valid and unchanged until qcs_synchronize / any synchronising
 * call returns. */
int qcs_set_state_async(qcs_register *reg, unsigned long long first, unsigned long long count,
                        const double *interleaved_in);

/* Arbitrary single-qubit gate: the 2x2 complex matrix u (row-major, interleaved re/im, 8
 * doubles) on qubit_num -- what HADAMARD_BASE_MATRIX (Q:210-213) is one instance of -- and its
 * controlled form (u applied where bit c_qubit_num is 1), the generalisation of
 * C_PHASE_SHIFT_BASE_MATRIX (Q:220-225).  Same pair-stride pass as qcs_hadamard_gate.  Between
 * qcs_fuse_begin and qcs_fuse_end a run of such gates whose qubits are all below 4 is multiplied up on
 * the host and applied as ONE 16 x 16 dense block on the FP64 tensor cores when the run ends. */
int qcs_apply_gate(qcs_register *reg, unsigned qubit_num, const double *u_interleaved);
int qcs_apply_controlled_gate(qcs_register *reg, unsigned c_qubit_num, unsigned qubit_num,
                              const double *u_interleaved);

/* State dump / load: 64-byte header + this shard's amplitudes as raw little-endian interleaved
 * doubles, the layout of gsl_vector_complex.data (Q:385-386).  A sharded register writes /
 * reads "<path>.rank<k>" on every rank.  The file must match the register's L, M and sharding. */
int qcs_save_state(qcs_register *reg, const char *path);
int qcs_load_state(qcs_register *reg, const char *path);

/* Deferred gate stream.  Between qcs_fuse_begin and qcs_fuse_end,
 * qcs_hadamard_gate and qcs_c_phase_shift_gate record their gate instead of
 * launching it; qcs_fuse_end schedules the recorded run into as few passes over
 * the state as the gates' commutation rules allow (Hadamards on a run of qubits
 * = one Walsh-Hadamard tile sweep; diagonal gates ride along in the sweeps'
 * registers) and launches them.  This is what replaces "one operate_matrix
 * pass per gate" (Q:370-420) for arbitrary H / C-phase circuits such as the
 * layered circuit of BASELINE configs[3].  Any other call on the register
 * flushes the recorded gates first, so results never depend on when the flush
 * happens.  With QCS_OPT_FUSION = 0 nothing is recorded: every gate runs at
 * once through the reference-order kernels.  Amplitudes agree with gate-by-gate
 * application to <= 1e-12 relative L2. */
int qcs_fuse_begin(qcs_register *reg);
int qcs_fuse_end(qcs_register *reg);
/* gates recorded and not yet launched */
unsigned long long qcs_fuse_pending(const qcs_register *reg);

/* Host-only view of the gate-stream scheduler (no device work, usable without a GPU): writes
 * the passes the recorded stream (kinds[i] = 0: hadamard_gate(q0[i]); 1:
 * c_phase_shift_gate(q0[i], q1[i], .)) would be launched as on rank `rank` of a register of
 * n_qubits sharded over world_size ranks, one text line per pass:
 *   "sweep h=<hex mask of H qubits> [diag=<stream positions>] [after=<stream positions>]"
 *   "hadamard q=<qubit> [after=...]"      "global h=<mask> [after=...]"      "before=<positions>"
 * diag: diagonal gates applied inside the sweep; after / before: by standalone kernels after the
 * pass / before every pass.  QCS_INSUFFICIENT_MEMORY when out_cap is too small. */
int qcs_schedule_describe(unsigned n_qubits, int world_size, int rank, unsigned long long n_gates,
                          const int *kinds, const unsigned *q0, const unsigned *q1, char *out,
                          unsigned long long out_cap);

/* Host-only view of the sweep planner (no device work): the fused sweeps an inverse transform on
 * qubits [lo, hi) of an n_qubits register is run as (tile_bits / min_run_bits = 0: library
 * defaults), one text line per sweep:
 *   "sweep a=<run bits> g=[<g_lo>,<g_hi>) tiles=<count> steps=<lowest bit>+<radix bits>,... scale=<factor>" */
int qcs_plan_describe(unsigned n_qubits, unsigned lo, unsigned hi, int tile_bits, int min_run_bits,
                      char *out, unsigned long long out_cap);

/* Fused dense block: an arbitrary 2^k x 2^k complex matrix U (row-major,
 * interleaved re/im, k = 3 or 4) applied to qubits 0..k-1 of every basis state,
 * i.e. what a run of gates on those qubits multiplies out to (generalises
 * HADAMARD_BASE_MATRIX / C_PHASE_SHIFT_BASE_MATRIX, Q:210-225).  FP64 tensor
 * cores (DMMA, mma.sync.m8n8k4.f64). */
int qcs_apply_dense_block(qcs_register *reg, unsigned k, const double *u_interleaved);

/* synthetic benchmark state: amp[i] = (u(2i), u(2i+1)),
 * u(k) = (mix64(seed + k) >> 11) * 2^-53 - 0.5, generated on the device */
int qcs_fill_synthetic(qcs_register *reg, unsigned long long seed);
int qcs_scale(qcs_register *reg, double factor);

/* ---- host-side scalar helpers (no device work) -------------------------- */
/* INT_POW, Q:158-159, with the out-of-range cast resolved as on x86-64 */
unsigned qcs_int_pow(unsigned base, unsigned power);
/* a^(2^k) mod C */
unsigned long long qcs_modpow2k(unsigned a, unsigned k, unsigned C);

/* Host-only self-check of the L2-paired launch bookkeeping (ticket order, block closure) for the sweeps
 * of an inverse transform on qubits [lo, hi) of a 2^n_qubits shard: pairs checked, or < 0 on a violation */
long long qcs_pair_selfcheck(unsigned n_qubits, unsigned lo, unsigned hi, int lag_tiles);

/* pinned host memory for qcs_get_state / qcs_set_state */
int qcs_host_alloc(void **ptr, size_t bytes);
/* the same, with the pages placed on the NUMA node next to `device` (host <-> device copies of
 * several GPUs then do not share one memory controller / socket link) */
int qcs_host_alloc_near(void **ptr, size_t bytes, int device);
int qcs_host_free(void *ptr);

/* ---- measurement of the engine itself ----------------------------------- */
/* CUDA-event stopwatch on the register's stream */
int qcs_timer_start(qcs_register *reg);
int qcs_timer_stop(qcs_register *reg, double *milliseconds);
/* kernels launched by this register since creation / since the last reset */
unsigned long long qcs_launch_count(const qcs_register *reg);
int qcs_profile_reset(qcs_register *reg);
/* per kernel class: launches, summed device milliseconds (QCS_OPT_PROFILE=1
 * only, else 0) and summed algorithmic bytes of those launches */
int qcs_profile_get(qcs_register *reg, int kernel_class, unsigned long long *launches,
                    double *milliseconds, double *algorithmic_bytes);
const char *qcs_kernel_class_name(int kernel_class);

#ifdef __cplusplus
}
#endif
#endif /* QCS_H */
