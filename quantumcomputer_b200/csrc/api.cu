// api.cu -- the C ABI of include/qcs.h: register life-cycle, gate dispatch
// (local vs. global qubits), composite operators and bookkeeping.
#include "qcs_internal.h"

#include <ctype.h>
#include <math.h>
#include <new>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif

// ---------------------------------------------------------------------------
// error plumbing and launch accounting
// ---------------------------------------------------------------------------
int qcs_map_cuda_error(cudaError_t e, const char *what, const char *file, int line)
{
    fprintf(stderr, "qcs: CUDA error %d (%s) at %s:%d in %s\n", (int) e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return QCS_INSUFFICIENT_MEMORY;
    if (e == cudaErrorInvalidValue || e == cudaErrorInvalidDevice) return QCS_BAD_ARGUMENTS;
    return QCS_UNKNOWN_ERROR;
}

void qcs_launch_begin(qcs_register *reg, int kind, double algorithmic_bytes)
{
    reg->launches_total++;
    reg->launches[kind]++;
    reg->alg_bytes[kind] += algorithmic_bytes;
    if (reg->opt_profile) {
        qcs_profile_slot s;
        if (!reg->free_slots.empty()) {
            s = reg->free_slots.back();
            reg->free_slots.pop_back();
        } else {
            cudaEventCreate(&s.begin);
            cudaEventCreate(&s.end);
        }
        s.kind = kind;
        cudaEventRecord(s.begin, reg->launch_stream ? reg->launch_stream : reg->stream);
        reg->pending.push_back(s);
    }
}

int qcs_launch_end(qcs_register *reg, int kind, const char *name)
{
    (void) kind;
    if (reg->opt_profile && !reg->pending.empty())
        cudaEventRecord(reg->pending.back().end, reg->launch_stream ? reg->launch_stream : reg->stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return qcs_map_cuda_error(e, name, __FILE__, __LINE__);
    if (reg->pending.size() > 4096) return qcs_profile_resolve(reg);
    return QCS_NO_ERROR;
}

int qcs_profile_resolve(qcs_register *reg)
{
    if (reg->pending.empty()) return QCS_NO_ERROR;
    QCS_CUDA(reg->dist ? cudaDeviceSynchronize() : cudaStreamSynchronize(reg->stream));
    for (auto &s : reg->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.begin, s.end) == cudaSuccess) reg->ms[s.kind] += (double) ms;
        reg->free_slots.push_back(s);
    }
    reg->pending.clear();
    return QCS_NO_ERROR;
}

// ---------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------
extern "C" const char *qcs_version(void) { return "qcs 0.2.0 (sm_100a)"; }

extern "C" const char *qcs_error_string(int code)
{
    switch (code) {
        case QCS_NO_ERROR: return "NO_ERROR";
        case QCS_INSUFFICIENT_MEMORY: return "INSUFFICIENT_MEMORY";
        case QCS_BAD_ARGUMENTS: return "BAD_ARGUMENTS";
        case QCS_PERIOD_NOT_FOUND: return "PERIOD_NOT_FOUND";
        default: return "UNKNOWN_ERROR";
    }
}

extern "C" const char *qcs_kernel_class_name(int k)
{
    static const char *names[QCS_K_COUNT] = {"hadamard", "cphase", "amodc", "fill", "reduce",
                                             "tile_sweep", "modexp_sweep", "exchange", "scale", "dense_block",
                                             "diag_multi", "global_sweep", "gate_1q"};
    return (k >= 0 && k < QCS_K_COUNT) ? names[k] : "?";
}

extern "C" int qcs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// INT_POW, qc_shor.c:158-159: (unsigned int)(pow(base, power) + 0.5); the
// out-of-range cast is resolved the way gcc/x86-64 compiles it (64-bit
// truncating conversion, low 32 bits kept; >= 2^63 or NaN -> 0).
extern "C" unsigned qcs_int_pow(unsigned base, unsigned power)
{
    const double d = pow((double) base, (double) power) + 0.5;
    if (!(d < 9223372036854775808.0) || !(d > -9223372036854775808.0)) return 0u;
    return (unsigned) (uint64_t) (int64_t) d;
}

extern "C" unsigned long long qcs_modpow2k(unsigned a, unsigned k, unsigned C)
{
    if (C == 0) return 0;
    unsigned long long v = a % C;
    for (unsigned s = 0; s < k; s++) v = (v * v) % C;
    return v;
}

extern "C" int qcs_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return QCS_NO_ERROR;
}

// Pinned memory close to a device: pages are placed on the NUMA node of the CPU that allocates
// them, so the calling thread is moved onto the CPUs next to the device's PCIe root (sysfs) for the
// duration of the allocation.  Falls back to qcs_host_alloc wherever the topology cannot be read.
extern "C" int qcs_host_alloc_near(void **ptr, size_t bytes, int device)
{
    if (!ptr) return QCS_BAD_ARGUMENTS;
    cpu_set_t old_set, near_set;
    bool moved = false;
    char bdf[32] = {0};
    if (device >= 0 && cudaDeviceGetPCIBusId(bdf, (int) sizeof bdf, device) == cudaSuccess &&
        sched_getaffinity(0, sizeof old_set, &old_set) == 0) {
        for (char *c = bdf; *c; c++) *c = (char) tolower(*c);
        char path[128];
        snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bdf);
        FILE *f = fopen(path, "r");
        if (f) {
            char list[4096] = {0};
            if (fgets(list, (int) sizeof list, f)) {
                CPU_ZERO(&near_set);
                int n_set = 0;
                for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
                    int a = 0, b = 0;
                    const int k = sscanf(tok, "%d-%d", &a, &b);
                    if (k == 1) b = a;
                    if (k >= 1)
                        for (int cpu = a; cpu <= b && cpu < CPU_SETSIZE; cpu++)
                            if (CPU_ISSET(cpu, &old_set)) { CPU_SET(cpu, &near_set); n_set++; }
                }
                if (n_set > 0 && sched_setaffinity(0, sizeof near_set, &near_set) == 0) moved = true;
            }
            fclose(f);
        }
    }
    // pinning populates the pages now, on the node of the CPU this thread runs on
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    if (e != cudaSuccess) return qcs_map_cuda_error(e, "cudaHostAlloc", __FILE__, __LINE__);
    return QCS_NO_ERROR;
}

extern "C" int qcs_host_free(void *ptr)
{
    QCS_CUDA(cudaFreeHost(ptr));
    return QCS_NO_ERROR;
}

// every entry point that is not a recordable gate first launches what the gate
// stream has recorded (circuit.cu), so deferral is never observable
#define QCS_ENTER(reg)                                              \
    do {                                                            \
        if (!(reg)) return QCS_BAD_ARGUMENTS;                       \
        QCS_CUDA(cudaSetDevice((reg)->device));                     \
        if (!(reg)->queue.empty() || (reg)->dense_pending) QCS_TRY(qcs_fuse_flush(reg));    \
        if ((reg)->lazy_reset) QCS_TRY(qcs_materialise_reset(reg)); \
    } while (0)
#define QCS_ENTER_GATE(reg)                         \
    do {                                            \
        if (!(reg)) return QCS_BAD_ARGUMENTS;       \
        QCS_CUDA(cudaSetDevice((reg)->device));     \
        if ((reg)->lazy_reset && !((reg)->fusing && (reg)->opt_fusion)) QCS_TRY(qcs_materialise_reset(reg)); \
    } while (0)

// write the pending basis state: |0...01> of a deferred reset_register (qc_shor.c:318-324) or the collapsed
// state of a measure_state (qc_shor.c:302-303)
int qcs_materialise_reset(qcs_register *reg)
{
    if (!reg->lazy_reset) return QCS_NO_ERROR;
    reg->lazy_reset = 0;
    const uint64_t owner = reg->lazy_index >> reg->n_local;
    return qcs_k_collapse(reg, reg->lazy_index & (reg->N_local - 1), owner == (uint64_t) reg->rank);
}

// ---------------------------------------------------------------------------
// register life-cycle
// ---------------------------------------------------------------------------
static int create_common(qcs_register **out, int L_size, int M_size, int device, int rank, int world,
                         const void *comm_id)
{
    if (!out) return QCS_BAD_ARGUMENTS;
    *out = nullptr;
    if (L_size < 0 || M_size < 0 || L_size + M_size < 1 || L_size + M_size > 62) return QCS_BAD_ARGUMENTS;
    if (world < 1 || (world & (world - 1)) != 0 || rank < 0 || rank >= world) return QCS_BAD_ARGUMENTS;
    int p = 0;
    while ((1 << p) < world) p++;
    if (p >= L_size + M_size) return QCS_BAD_ARGUMENTS;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fprintf(stderr, "qcs: no usable CUDA device (%s); this library has no CPU path\n",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return QCS_UNKNOWN_ERROR;
    }
    if (device < 0) QCS_CUDA(cudaGetDevice(&device));
    if (device >= ndev) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(device));

    qcs_register *reg = new (std::nothrow) qcs_register();
    if (!reg) return QCS_INSUFFICIENT_MEMORY;
    reg->L_size = L_size;
    reg->M_size = M_size;
    reg->n = (unsigned) (L_size + M_size);
    reg->n_local = reg->n - (unsigned) p;
    reg->N = 1ull << reg->n;
    reg->N_local = 1ull << reg->n_local;
    reg->device = device;
    reg->rank = rank;
    reg->world = world;
    reg->p_global = p;
    reg->dist = nullptr;
    reg->peer = nullptr;
    reg->group = nullptr;
    reg->amp_all = nullptr;
    reg->opt_fusion = 1;
    reg->opt_profile = 0;
    reg->opt_tile_bits = 0;
    reg->opt_measure_sequential = 0;
    reg->opt_pipeline = 1;
    reg->opt_pipe_shape = -1;
    reg->opt_min_run_bits = 3;
    reg->opt_global_run_bits = 7;
    reg->opt_overlap_slices = 8;
    reg->opt_global_sms = 48;
    reg->opt_prefetch_tiles = 0;     // measured: L2 prefetch slows the sweep down (profiles/README.md)
    reg->d_meas = nullptr;
    reg->d_pair = nullptr;
    reg->d_dense = nullptr;
    reg->d_gen_table = nullptr;
    reg->gen_table_cap = 0;
    reg->gen.f = nullptr;
    reg->gen.armed = 0;
    reg->d_pair_cap = 0;
    reg->opt_l2_pair = 1;
    reg->opt_l2_pair_hints = 1;
    reg->opt_split3 = 1;
    reg->opt_gen_sweep = 1;
    reg->opt_l2_pair_lag = 3 * 148;
    reg->opt_l2_pair_max_block = 16ll << 20;     // measured: 32 MiB blocks (n = 30) no longer stay in the L2 (profiles/README.md)
    reg->fusing = 0;
    reg->lazy_reset = 0;
    reg->lazy_index = 1;
    reg->dense_pending = 0;
    reg->dense_gates = 0;
    reg->d_diag = nullptr;
    reg->d_diag_cap = 0;
    reg->launches_total = 0;
    memset(reg->launches, 0, sizeof reg->launches);
    memset(reg->alg_bytes, 0, sizeof reg->alg_bytes);
    memset(reg->ms, 0, sizeof reg->ms);
    reg->amp = nullptr;
    reg->d_partials = nullptr;
    reg->d_small = nullptr;
    reg->h_small = nullptr;
    reg->stream = nullptr;
    reg->launch_stream = nullptr;
    reg->timer_begin = reg->timer_end = nullptr;

    int rc = QCS_NO_ERROR;
    cudaDeviceProp prop;
    do {
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) break;
        reg->sm_count = prop.multiProcessorCount;
        reg->smem_optin = prop.sharedMemPerBlockOptin;
        if ((e = cudaStreamCreateWithFlags(&reg->stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        reg->partials_cap = (size_t) reg->sm_count * 16;
        if ((e = cudaMalloc((void **) &reg->d_partials, reg->partials_cap * sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMalloc(&reg->d_small, 4096)) != cudaSuccess) break;
        if ((e = cudaHostAlloc(&reg->h_small, 4096, cudaHostAllocDefault)) != cudaSuccess) break;
        if ((e = cudaEventCreate(&reg->timer_begin)) != cudaSuccess) break;
        if ((e = cudaEventCreate(&reg->timer_end)) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) {
        rc = qcs_map_cuda_error(e, "register allocation", __FILE__, __LINE__);
        qcs_register_destroy(reg);
        return rc;
    }
    if (world > 1) {
        rc = qcs_dist_init(reg, comm_id);
        if (rc != QCS_NO_ERROR) { qcs_register_destroy(reg); return rc; }
        // one address range over all shards (peer.cu); every rank must have it or none.  First agree
        // that every rank can try at all (a rank that opted out would never answer the exchange) ...
        std::vector<double> flags((size_t) world);
        rc = qcs_dist_allgather_double(reg, qcs_peer_can(reg) ? 1.0 : 0.0, flags.data());
        if (rc != QCS_NO_ERROR) { qcs_register_destroy(reg); return rc; }
        bool all = true;
        for (double f : flags) all = all && f != 0.0;
        if (all) {
            // ... then that every rank succeeded
            const bool mine = qcs_peer_try_alloc(reg, comm_id);
            rc = qcs_dist_allgather_double(reg, mine ? 1.0 : 0.0, flags.data());
            if (rc != QCS_NO_ERROR) { qcs_register_destroy(reg); return rc; }
            for (double f : flags) all = all && f != 0.0;
            if (!all && mine) qcs_peer_free(reg);
        }
    }
    if (!reg->amp) {
        e = cudaMalloc((void **) &reg->amp, reg->N_local * sizeof(double2));
        if (e != cudaSuccess) {
            rc = qcs_map_cuda_error(e, "register allocation", __FILE__, __LINE__);
            qcs_register_destroy(reg);
            return rc;
        }
    }
    *out = reg;
    return QCS_NO_ERROR;
}

extern "C" int qcs_register_create(qcs_register **out, int L_size, int M_size, int device)
{
    return create_common(out, L_size, M_size, device, 0, 1, nullptr);
}

extern "C" int qcs_register_create_sharded(qcs_register **out, int L_size, int M_size, int device,
                                           int rank, int world_size, const void *comm_id)
{
    if (world_size > 1 && !comm_id) return QCS_BAD_ARGUMENTS;
    return create_common(out, L_size, M_size, device, rank, world_size, comm_id);
}

extern "C" void qcs_register_destroy(qcs_register *reg)
{
    if (!reg) return;
    if (reg->group) { qcs_group_destroy(reg); return; }
    cudaSetDevice(reg->device);
    if (reg->stream) cudaStreamSynchronize(reg->stream);
    // no collective here: a peer that still touches this shard does so through its own mapping of
    // the allocation, which stays alive until that peer releases its imported handle
    if (reg->dist) qcs_dist_destroy(reg);
    for (auto &s : reg->pending) { cudaEventDestroy(s.begin); cudaEventDestroy(s.end); }
    for (auto &s : reg->free_slots) { cudaEventDestroy(s.begin); cudaEventDestroy(s.end); }
    if (reg->timer_begin) cudaEventDestroy(reg->timer_begin);
    if (reg->timer_end) cudaEventDestroy(reg->timer_end);
    if (reg->peer) qcs_peer_free(reg);
    else if (reg->amp) cudaFree(reg->amp);
    if (reg->d_partials) cudaFree(reg->d_partials);
    if (reg->d_small) cudaFree(reg->d_small);
    if (reg->d_meas) cudaFree(reg->d_meas);
    if (reg->d_pair) cudaFree(reg->d_pair);
    if (reg->d_dense) cudaFree(reg->d_dense);
    if (reg->d_gen_table) cudaFree(reg->d_gen_table);
    if (reg->d_diag) cudaFree(reg->d_diag);
    if (reg->h_small) cudaFreeHost(reg->h_small);
    if (reg->stream) cudaStreamDestroy(reg->stream);
    delete reg;
}

extern "C" int qcs_L_size(const qcs_register *reg) { return reg ? reg->L_size : -1; }
extern "C" int qcs_M_size(const qcs_register *reg) { return reg ? reg->M_size : -1; }
extern "C" unsigned qcs_num_qubits(const qcs_register *reg) { return reg ? reg->n : 0; }
extern "C" unsigned long long qcs_num_states(const qcs_register *reg) { return reg ? reg->N : 0; }
extern "C" unsigned long long qcs_local_states(const qcs_register *reg) { return reg ? reg->N_local : 0; }
extern "C" int qcs_rank(const qcs_register *reg) { return reg ? reg->rank : -1; }
extern "C" int qcs_world_size(const qcs_register *reg) { return reg ? reg->world : 0; }
extern "C" int qcs_peer_memory(const qcs_register *reg)
{
    if (reg && reg->group) return qcs_peer_memory(qcs_group_member(reg, 0));
    return reg && reg->peer ? 1 : 0;
}

extern "C" int qcs_set_option(qcs_register *reg, int option, long long value)
{
    QCS_GROUP_FORWARD(reg, qcs_set_option(m, option, value));
    QCS_ENTER(reg);      // recorded gates are launched under the options they were recorded with
    switch (option) {
        case QCS_OPT_FUSION: reg->opt_fusion = value != 0; return QCS_NO_ERROR;
        case QCS_OPT_PROFILE:
            if (!value) QCS_TRY(qcs_profile_resolve(reg));
            reg->opt_profile = value != 0;
            return QCS_NO_ERROR;
        case QCS_OPT_TILE_BITS:
            if (value != 0 && (value < 8 || value > 13)) return QCS_BAD_ARGUMENTS;
            reg->opt_tile_bits = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_MEASURE_SEQUENTIAL: reg->opt_measure_sequential = value != 0; return QCS_NO_ERROR;
        case QCS_OPT_PIPELINE: reg->opt_pipeline = value != 0; return QCS_NO_ERROR;
        case QCS_OPT_PIPE_SHAPE:
            if (value < -1 || value > 16) return QCS_BAD_ARGUMENTS;
            reg->opt_pipe_shape = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_MIN_RUN_BITS:
            if (value < 1 || value > 7) return QCS_BAD_ARGUMENTS;
            reg->opt_min_run_bits = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_OVERLAP_SLICES:
            if (value < 0 || value > 8 || (value & (value - 1)) != 0) return QCS_BAD_ARGUMENTS;
            reg->opt_overlap_slices = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_GLOBAL_SMS:
            if (value < 1 || value > 1024) return QCS_BAD_ARGUMENTS;
            reg->opt_global_sms = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_GLOBAL_RUN_BITS:
            if (value < 1 || value > 7) return QCS_BAD_ARGUMENTS;
            reg->opt_global_run_bits = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_PREFETCH_TILES:
            if (value < 0 || value > 64) return QCS_BAD_ARGUMENTS;
            reg->opt_prefetch_tiles = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_L2_PAIR: reg->opt_l2_pair = value != 0; return QCS_NO_ERROR;
        case QCS_OPT_L2_PAIR_LAG:
            if (value < 0 || value > (1 << 20)) return QCS_BAD_ARGUMENTS;
            reg->opt_l2_pair_lag = (int) value;
            return QCS_NO_ERROR;
        case QCS_OPT_L2_PAIR_HINTS: reg->opt_l2_pair_hints = value != 0; return QCS_NO_ERROR;
        case QCS_OPT_SPLIT3: reg->opt_split3 = value != 0; return QCS_NO_ERROR;
        case QCS_OPT_GEN_SWEEP: reg->opt_gen_sweep = value != 0; return QCS_NO_ERROR;
        case QCS_OPT_L2_PAIR_MAX_BLOCK:
            if (value < (1 << 20)) return QCS_BAD_ARGUMENTS;
            reg->opt_l2_pair_max_block = value;
            return QCS_NO_ERROR;
        default: return QCS_BAD_ARGUMENTS;
    }
}

extern "C" long long qcs_get_option(const qcs_register *reg, int option)
{
    if (!reg) return -1;
    if (reg->group) return qcs_get_option(qcs_group_member(reg, 0), option);
    switch (option) {
        case QCS_OPT_FUSION: return reg->opt_fusion;
        case QCS_OPT_PROFILE: return reg->opt_profile;
        case QCS_OPT_TILE_BITS: return reg->opt_tile_bits;
        case QCS_OPT_MEASURE_SEQUENTIAL: return reg->opt_measure_sequential;
        case QCS_OPT_PIPELINE: return reg->opt_pipeline;
        case QCS_OPT_PIPE_SHAPE: return reg->opt_pipe_shape;
        case QCS_OPT_MIN_RUN_BITS: return reg->opt_min_run_bits;
        case QCS_OPT_GLOBAL_RUN_BITS: return reg->opt_global_run_bits;
        case QCS_OPT_OVERLAP_SLICES: return reg->opt_overlap_slices;
        case QCS_OPT_GLOBAL_SMS: return reg->opt_global_sms;
        case QCS_OPT_PREFETCH_TILES: return reg->opt_prefetch_tiles;
        case QCS_OPT_L2_PAIR: return reg->opt_l2_pair;
        case QCS_OPT_L2_PAIR_LAG: return reg->opt_l2_pair_lag;
        case QCS_OPT_L2_PAIR_MAX_BLOCK: return reg->opt_l2_pair_max_block;
        case QCS_OPT_L2_PAIR_HINTS: return reg->opt_l2_pair_hints;
        case QCS_OPT_SPLIT3: return reg->opt_split3;
        case QCS_OPT_GEN_SWEEP: return reg->opt_gen_sweep;
        default: return -1;
    }
}

extern "C" int qcs_synchronize(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_synchronize(m));
    QCS_ENTER(reg);
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    return QCS_NO_ERROR;
}

// ---------------------------------------------------------------------------
// single gates
// ---------------------------------------------------------------------------
extern "C" int qcs_reset_register(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_reset_register(m));
    if (!reg) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(reg->device));
    // whatever was recorded or pending is overwritten by the reset: drop it
    reg->queue.clear();
    reg->dense_pending = 0;
    if (reg->opt_fusion) {
        // fused mode: written when first needed -- or never, when quantum_computation follows (qc_shor.c:922-923)
        reg->lazy_reset = 1;
        reg->lazy_index = 1;
        return QCS_NO_ERROR;
    }
    reg->lazy_reset = 0;
    return qcs_k_reset(reg);
}

static int hadamard_any(qcs_register *reg, unsigned q)
{
    if (q >= reg->n) return QCS_BAD_ARGUMENTS;
    if (q < reg->n_local) return qcs_k_hadamard_local(reg, q);
    if (reg->peer) return qcs_k_hadamard_peer(reg, q);
    return qcs_dist_hadamard_global(reg, q);
}

extern "C" int qcs_hadamard_gate(qcs_register *reg, unsigned qubit_num)
{
    QCS_GROUP_FORWARD(reg, qcs_hadamard_gate(m, qubit_num));
    QCS_ENTER_GATE(reg);
    if (reg->fusing && reg->opt_fusion) {
        if (qubit_num >= reg->n) return QCS_BAD_ARGUMENTS;
        if (reg->dense_pending) QCS_TRY(qcs_fuse_flush(reg));         // program order: the dense block first
        reg->queue.push_back({0, qubit_num, qubit_num, 0.0, 0.0});
        return QCS_NO_ERROR;
    }
    return hadamard_any(reg, qubit_num);
}

// diagonal gate: a global qubit only contributes this rank's (constant) bit
static int cphase_any(qcs_register *reg, unsigned c, unsigned q, double co, double si)
{
    if (c >= reg->n || q >= reg->n) return QCS_BAD_ARGUMENTS;
    unsigned bits[2];
    int nb = 0;
    const unsigned both[2] = {c, q};
    for (int k = 0; k < (c == q ? 1 : 2); k++) {
        const unsigned b = both[k];
        if (b < reg->n_local) bits[nb++] = b;
        else if (!(((unsigned) reg->rank >> (b - reg->n_local)) & 1u)) return QCS_NO_ERROR;
    }
    return qcs_k_phase_masked(reg, nb, nb > 0 ? bits[0] : 0, nb > 1 ? bits[1] : 0, co, si);
}

extern "C" int qcs_c_phase_shift_gate(qcs_register *reg, unsigned c_qubit_num, unsigned qubit_num, double theta)
{
    QCS_GROUP_FORWARD(reg, qcs_c_phase_shift_gate(m, c_qubit_num, qubit_num, theta));
    QCS_ENTER_GATE(reg);
    // gsl_complex_polar(1.0, theta), qc_shor.c:526: host libm like the reference
    if (reg->fusing && reg->opt_fusion) {
        if (c_qubit_num >= reg->n || qubit_num >= reg->n) return QCS_BAD_ARGUMENTS;
        if (reg->dense_pending) QCS_TRY(qcs_fuse_flush(reg));
        reg->queue.push_back({1, c_qubit_num, qubit_num, 1.0 * cos(theta), 1.0 * sin(theta)});
        return QCS_NO_ERROR;
    }
    return cphase_any(reg, c_qubit_num, qubit_num, 1.0 * cos(theta), 1.0 * sin(theta));
}

static int amodc_any(qcs_register *reg, unsigned C, unsigned long long atox, unsigned c)
{
    if (C == 0 || c >= reg->n) return QCS_BAD_ARGUMENTS;
    if (C > 65536u) {
        fprintf(stderr, "qcs: c_amodc_gate: C > 65536 would wrap the reference's 32-bit product (qc_shor.c:639)\n");
        return QCS_BAD_ARGUMENTS;
    }
    if ((unsigned) reg->M_size > reg->n_local) return QCS_BAD_ARGUMENTS;
    if (reg->M_size == 0) return QCS_NO_ERROR;               // f = f' = 0: identity
    const unsigned A = (unsigned) (atox % C);                // qc_shor.c:605
    if (c >= reg->n_local) {
        const bool on = ((unsigned) reg->rank >> (c - reg->n_local)) & 1u;
        return qcs_k_amodc(reg, C, A, -1, !on);
    }
    return qcs_k_amodc(reg, C, A, (int) c, false);
}

extern "C" int qcs_c_amodc_gate(qcs_register *reg, unsigned C, unsigned long long atox, unsigned c_qubit_num)
{
    QCS_GROUP_FORWARD(reg, qcs_c_amodc_gate(m, C, atox, c_qubit_num));
    QCS_ENTER(reg);
    return amodc_any(reg, C, atox, c_qubit_num);
}

// ---------------------------------------------------------------------------
// composite operators
// ---------------------------------------------------------------------------
// theta = M_PI / INT_POW(2, d) (qc_shor.c:686); exact power of two for d <= 31,
// continued as ldexp beyond the reference's overflow point
static double qft_theta(unsigned d)
{
    return d <= 31 ? M_PI / (double) qcs_int_pow(2, d) : ldexp(M_PI, -(int) d);
}

static int qft_gate_by_gate(qcs_register *reg, unsigned lo, unsigned hi, bool inverse)
{
    if (inverse) {
        // qc_shor.c:682-689
        for (int l = (int) hi - 1; l >= (int) lo; l--) {
            QCS_TRY(hadamard_any(reg, (unsigned) l));
            for (int k = l - 1; k >= (int) lo; k--) {
                const double th = qft_theta((unsigned) (l - k));
                QCS_TRY(cphase_any(reg, (unsigned) l, (unsigned) k, 1.0 * cos(th), 1.0 * sin(th)));
            }
        }
    } else {
        for (int l = (int) lo; l < (int) hi; l++) {
            for (int k = (int) lo; k < l; k++) {
                const double th = -qft_theta((unsigned) (l - k));
                QCS_TRY(cphase_any(reg, (unsigned) l, (unsigned) k, 1.0 * cos(th), 1.0 * sin(th)));
            }
            QCS_TRY(hadamard_any(reg, (unsigned) l));
        }
    }
    return QCS_NO_ERROR;
}

static int qft_any(qcs_register *reg, unsigned lo, unsigned hi, bool inverse)
{
    if (lo > hi || hi > reg->n) return QCS_BAD_ARGUMENTS;
    if (lo == hi) return QCS_NO_ERROR;
    if (!reg->opt_fusion) return qft_gate_by_gate(reg, lo, hi, inverse);
    if (hi <= reg->n_local) return qcs_fused_qft(reg, lo, hi, inverse);
    // sharded register, transform reaches the global qubits.  With peer memory the sweeps whose
    // tile holds global qubits run on the stitched array (qft_fused.cu); without it their stages
    // run as exchange + sweep + exchange (dist.cu), the local ones as ordinary sweeps
    if (qcs_sharded_sweeps_supported(reg, lo, hi)) return qcs_fused_sweeps_sharded(reg, lo, hi, inverse, false);
    const unsigned q = reg->n_local - (unsigned) reg->p_global;
    if (hi != reg->n || lo > q || reg->n_local < 2u * (unsigned) reg->p_global)
        return qft_gate_by_gate(reg, lo, hi, inverse);       // odd shapes: pairwise exchanges, gate by gate
    if (inverse) {
        QCS_TRY(qcs_dist_top_stages(reg, lo, true, false));
        return qcs_fused_qft(reg, lo, reg->n_local, true);
    }
    QCS_TRY(qcs_fused_qft(reg, lo, reg->n_local, false));
    return qcs_dist_top_stages(reg, lo, false, false);
}

extern "C" int qcs_inverse_QFT(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_inverse_QFT(m));
    QCS_ENTER(reg);
    return qft_any(reg, (unsigned) reg->M_size, reg->n, true);
}

extern "C" int qcs_QFT(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_QFT(m));
    QCS_ENTER(reg);
    return qft_any(reg, (unsigned) reg->M_size, reg->n, false);
}

extern "C" int qcs_inverse_QFT_range(qcs_register *reg, unsigned lo, unsigned hi)
{
    QCS_GROUP_FORWARD(reg, qcs_inverse_QFT_range(m, lo, hi));
    QCS_ENTER(reg);
    return qft_any(reg, lo, hi, true);
}

extern "C" int qcs_QFT_range(qcs_register *reg, unsigned lo, unsigned hi)
{
    QCS_GROUP_FORWARD(reg, qcs_QFT_range(m, lo, hi));
    QCS_ENTER(reg);
    return qft_any(reg, lo, hi, false);
}

extern "C" int qcs_quantum_computation(qcs_register *reg, unsigned C, unsigned a, int pow_mode)
{
    QCS_GROUP_FORWARD(reg, qcs_quantum_computation(m, C, a, pow_mode));
    if (!reg) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(reg->device));
    // find_period's sequence reset_register -> quantum_computation (qc_shor.c:922-923): the state after the
    // Hadamards and the controlled multiplications of |0...01> is known in closed form -- 2^(-L/2) at
    // (x, a^x mod C built gate by gate exactly as c_amodc_gate would) -- and is written in ONE pass instead of
    // a reset pass, the Walsh-Hadamard sweeps and the modular-exponentiation sweep
    const bool from_reset = reg->lazy_reset && reg->lazy_index == 1 && reg->queue.empty() && !reg->dense_pending && reg->opt_fusion && reg->L_size > 0;
    if (!from_reset) QCS_ENTER(reg);
    if (C == 0 || (pow_mode != QCS_POW_VERBATIM && pow_mode != QCS_POW_MODULAR)) return QCS_BAD_ARGUMENTS;
    const unsigned first = reg->n - (unsigned) reg->L_size;          // qc_shor.c:720
    // atox per gate, qc_shor.c:728-731 (x doubles as an unsigned int)
    std::vector<unsigned long long> atox((size_t) reg->L_size);
    unsigned x = 1;
    for (int k = 0; k < reg->L_size; k++) {
        atox[(size_t) k] = pow_mode == QCS_POW_MODULAR ? qcs_modpow2k(a, (unsigned) k, C)
                                                       : (unsigned long long) qcs_int_pow(a, x);
        x *= 2;
    }
    if (reg->opt_fusion && reg->L_size > 0) {
        std::vector<unsigned> A((size_t) reg->L_size);
        for (int k = 0; k < reg->L_size; k++) A[(size_t) k] = (unsigned) (atox[(size_t) k] % C);
        if (from_reset) {
            // ... or not written at all: the first sweep of the inverse QFT builds its tiles from f(x)
            bool armed = false;
            QCS_TRY(qcs_shor_state_generated(reg, C, A.data(), (unsigned) reg->L_size, &armed));
            if (armed) {
                reg->lazy_reset = 0;
                const int rc = qft_any(reg, (unsigned) reg->M_size, reg->n, true);
                if (rc == QCS_NO_ERROR && reg->gen.armed) {          // the plan check and the launch path disagree
                    reg->gen.armed = 0;
                    fprintf(stderr, "qcs: the generating sweep was armed but never launched\n");
                    return QCS_UNKNOWN_ERROR;
                }
                reg->gen.armed = 0;
                return rc;
            }
            bool done = false;
            QCS_TRY(qcs_shor_state_from_reset(reg, C, A.data(), (unsigned) reg->L_size, &done));
            if (done) {
                reg->lazy_reset = 0;
                return qft_any(reg, (unsigned) reg->M_size, reg->n, true);
            }
            QCS_TRY(qcs_materialise_reset(reg));
        }
        int rc = qcs_fused_modexp(reg, C, A.data(), (unsigned) reg->L_size);
        if (rc != QCS_NO_ERROR) return rc;
        return qft_any(reg, (unsigned) reg->M_size, reg->n, true);
    }
    for (unsigned l = first; l < reg->n; l++) QCS_TRY(hadamard_any(reg, l));
    for (unsigned l = first; l < reg->n; l++) QCS_TRY(amodc_any(reg, C, atox[l - first], l));
    return qft_any(reg, (unsigned) reg->M_size, reg->n, true);
}

// ---------------------------------------------------------------------------
// measurement and read-back
// ---------------------------------------------------------------------------
extern "C" int qcs_norm2(qcs_register *reg, double *sum_of_sq)
{
    if (reg && reg->group) {
        if (!sum_of_sq) return QCS_BAD_ARGUMENTS;
        std::vector<double> v((size_t) qcs_group_world(reg), 0.0);
        const int rc = qcs_group_run(reg, [&](qcs_register *m) { return qcs_norm2(m, &v[(size_t) m->rank]); });
        *sum_of_sq = v[0];                        // the rank-ordered sum, identical on every shard
        return rc;
    }
    QCS_ENTER(reg);
    if (!sum_of_sq) return QCS_BAD_ARGUMENTS;
    double mine = 0.0;
    QCS_TRY(qcs_k_norm2_local(reg, &mine));
    if (reg->world == 1) { *sum_of_sq = mine; return QCS_NO_ERROR; }
    std::vector<double> all((size_t) reg->world);
    QCS_TRY(qcs_dist_allgather_double(reg, mine, all.data()));
    double s = 0.0;
    for (int r = 0; r < reg->world; r++) s += all[(size_t) r];   // rank order: deterministic
    *sum_of_sq = s;
    return QCS_NO_ERROR;
}

// the index measure_state would return for this r (qc_shor.c:283-292), without the collapse
static int locate_state(qcs_register *reg, double r, uint64_t *out_index)
{
    // qc_shor.c:283: the scan covers indices 0 .. N-2; N-1 is the fall-through
    int found = 0;
    uint64_t index = 0;
    double cum = 0.0;
    uint64_t global_index = reg->N - 1;
    if (reg->world == 1) {
        QCS_TRY(qcs_k_measure_scan(reg, 0.0, r, reg->N_local - 1, &found, &index, &cum));
        if (found) global_index = index;
    } else {
        // Everything except the walk runs on all shards at once: chunk sums -> one all-gather of the
        // shards' approximate totals (enough for the rigorous classification bounds) -> chunk and
        // super-chunk maps.  Only the exact running sum is handed from shard to shard in index order,
        // one packed all-gather {sum, found, index, error code} per shard.
        const uint64_t limit = reg->rank == reg->world - 1 ? reg->N_local - 1 : reg->N_local;
        const bool parallel = qcs_k_scan_parallel_ok(reg, reg->N_local);     // the same decision on every shard (limit differs on the last one)
        std::vector<double> all((size_t) reg->world * 4);
        int rc = QCS_NO_ERROR;
        if (parallel) {
            double mine = 0.0;
            rc = qcs_k_scan_sums(reg, limit, &mine);
            const double pack[2] = {mine, (double) rc};
            QCS_TRY(qcs_dist_allgather_doubles(reg, pack, 2, all.data()));
            double before = 0.0;
            for (int s = 0; s < reg->world; s++) {
                if (all[(size_t) s * 2 + 1] != 0.0) rc = (int) all[(size_t) s * 2 + 1];    // every shard returns the same error
                if (s < reg->rank) before += all[(size_t) s * 2];
            }
            if (rc != QCS_NO_ERROR) return rc;
            rc = qcs_k_scan_maps(reg, before, r, limit);
        }
        for (int turn = 0; turn < reg->world; turn++) {
            if (turn == reg->rank && rc == QCS_NO_ERROR && !found) {
                int bad = 0;
                if (parallel) rc = qcs_k_scan_walk(reg, cum, r, limit, &found, &index, &cum, &bad);
                if (rc == QCS_NO_ERROR && (!parallel || bad)) {
                    if (bad) fprintf(stderr, "qcs: measure_state: binade invariant failed, falling back to the sequential GPU scan\n");
                    rc = qcs_k_measure_scan(reg, cum, r, limit, &found, &index, &cum);
                }
                if (found) global_index = (uint64_t) reg->rank * reg->N_local + index;
            }
            // {running sum, found, index (< 2^53: exact in a double), error code} of the shard whose turn it was
            const double pack[4] = {cum, found ? 1.0 : 0.0, (double) global_index, (double) rc};
            QCS_TRY(qcs_dist_allgather_doubles(reg, pack, 4, all.data()));
            // an error on any shard ends the scan on all of them
            for (int s = 0; s < reg->world; s++)
                if (all[(size_t) s * 4 + 3] != 0.0) return (int) all[(size_t) s * 4 + 3];
            cum = all[(size_t) turn * 4];
            found = all[(size_t) turn * 4 + 1] != 0.0;
            global_index = (uint64_t) all[(size_t) turn * 4 + 2];
            if (found) break;
        }
    }
    *out_index = global_index;
    return QCS_NO_ERROR;
}

extern "C" int qcs_measure_state(qcs_register *reg, double r, unsigned long long *state_num)
{
    if (reg && reg->group) {
        if (!state_num) return QCS_BAD_ARGUMENTS;
        std::vector<unsigned long long> v((size_t) qcs_group_world(reg), 0ull);
        const int rc = qcs_group_run(reg, [&](qcs_register *m) { return qcs_measure_state(m, r, &v[(size_t) m->rank]); });
        *state_num = v[0];
        return rc;
    }
    QCS_ENTER(reg);
    if (!state_num) return QCS_BAD_ARGUMENTS;
    uint64_t global_index = 0;
    QCS_TRY(locate_state(reg, r, &global_index));
    // collapse, qc_shor.c:302-303.  Fused mode: deferred like reset_register -- in find_period the next call
    // is the next trial's reset_register, which overwrites it unseen
    if (reg->opt_fusion) {
        reg->lazy_reset = 1;
        reg->lazy_index = global_index;
    } else {
        const uint64_t owner = global_index >> reg->n_local;
        QCS_TRY(qcs_k_collapse(reg, global_index & (reg->N_local - 1), owner == (uint64_t) reg->rank));
    }
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    *state_num = global_index;
    return QCS_NO_ERROR;
}

// Sampling without collapse (qc_shor.c:294-301 notes the option): for every
// variate r[k] the index measure_state would have returned, the state left untouched, so the
// reference's "rerun the whole computation per sample" loop becomes one state preparation.
extern "C" int qcs_sample_states(qcs_register *reg, unsigned long long n_shots, const double *r,
                                 unsigned long long *state_nums)
{
    if (reg && reg->group) {
        if (n_shots && (!r || !state_nums)) return QCS_BAD_ARGUMENTS;
        // every shard fills a complete copy of the answer; the caller gets shard 0's
        std::vector<std::vector<unsigned long long>> v((size_t) qcs_group_world(reg));
        const int rc = qcs_group_run(reg, [&](qcs_register *m) {
            v[(size_t) m->rank].assign((size_t) n_shots, 0ull);
            return qcs_sample_states(m, n_shots, r, v[(size_t) m->rank].data());
        });
        for (unsigned long long k = 0; k < n_shots; k++) state_nums[k] = v[0][(size_t) k];
        return rc;
    }
    QCS_ENTER(reg);
    if (n_shots && (!r || !state_nums)) return QCS_BAD_ARGUMENTS;
    bool handled = false;
    QCS_TRY(qcs_k_sample_many(reg, n_shots, r, state_nums, &handled));
    if (handled) return QCS_NO_ERROR;
    for (unsigned long long k = 0; k < n_shots; k++) {
        uint64_t idx = 0;
        QCS_TRY(locate_state(reg, r[k], &idx));
        state_nums[k] = idx;
    }
    return QCS_NO_ERROR;
}

extern "C" int qcs_get_state(qcs_register *reg, unsigned long long first, unsigned long long count,
                             double *interleaved_out)
{
    if (reg && reg->group) {
        if (first > reg->N || count > reg->N - first || (!interleaved_out && count)) return QCS_BAD_ARGUMENTS;
        return qcs_group_run(reg, [&](qcs_register *m) -> int {
            const unsigned long long lo = (unsigned long long) m->rank * m->N_local, hi = lo + m->N_local;
            const unsigned long long a = first > lo ? first : lo, b = first + count < hi ? first + count : hi;
            if (a >= b) return QCS_NO_ERROR;
            return qcs_get_state(m, a - lo, b - a, interleaved_out + 2 * (a - first));
        });
    }
    QCS_ENTER(reg);
    if (first > reg->N_local || count > reg->N_local - first || (!interleaved_out && count)) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaMemcpyAsync(interleaved_out, reg->amp + first, count * sizeof(double2),
                             cudaMemcpyDeviceToHost, reg->stream));
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    return QCS_NO_ERROR;
}

extern "C" int qcs_set_state(qcs_register *reg, unsigned long long first, unsigned long long count,
                             const double *interleaved_in)
{
    if (reg && reg->group) {
        if (first > reg->N || count > reg->N - first || (!interleaved_in && count)) return QCS_BAD_ARGUMENTS;
        return qcs_group_run(reg, [&](qcs_register *m) -> int {
            const unsigned long long lo = (unsigned long long) m->rank * m->N_local, hi = lo + m->N_local;
            const unsigned long long a = first > lo ? first : lo, b = first + count < hi ? first + count : hi;
            if (a >= b) return QCS_NO_ERROR;
            return qcs_set_state(m, a - lo, b - a, interleaved_in + 2 * (a - first));
        });
    }
    // the whole shard is overwritten: whatever is pending (a deferred reset / collapse, recorded gates) is dead
    if (reg && !reg->group && first == 0 && count == reg->N_local && interleaved_in) {
        reg->lazy_reset = 0;
        reg->queue.clear();
        reg->dense_pending = 0;
    }
    QCS_ENTER(reg);
    if (first > reg->N_local || count > reg->N_local - first || (!interleaved_in && count)) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaMemcpyAsync(reg->amp + first, interleaved_in, count * sizeof(double2),
                             cudaMemcpyHostToDevice, reg->stream));
    // the caller may reuse or free the buffer as soon as this returns (qcs_get_state synchronises too)
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    return QCS_NO_ERROR;
}

// the same, stream-ordered: the source (pinned memory from qcs_host_alloc for a real overlap) must
// stay valid and unchanged until qcs_synchronize or any synchronising call returns
extern "C" int qcs_set_state_async(qcs_register *reg, unsigned long long first, unsigned long long count,
                                   const double *interleaved_in)
{
    if (reg && reg->group) {
        if (first > reg->N || count > reg->N - first || (!interleaved_in && count)) return QCS_BAD_ARGUMENTS;
        return qcs_group_run(reg, [&](qcs_register *m) -> int {
            const unsigned long long lo = (unsigned long long) m->rank * m->N_local, hi = lo + m->N_local;
            const unsigned long long a = first > lo ? first : lo, b = first + count < hi ? first + count : hi;
            if (a >= b) return QCS_NO_ERROR;
            return qcs_set_state_async(m, a - lo, b - a, interleaved_in + 2 * (a - first));
        });
    }
    // the whole shard is overwritten: whatever is pending (a deferred reset / collapse, recorded gates) is dead
    if (reg && !reg->group && first == 0 && count == reg->N_local && interleaved_in) {
        reg->lazy_reset = 0;
        reg->queue.clear();
        reg->dense_pending = 0;
    }
    QCS_ENTER(reg);
    if (first > reg->N_local || count > reg->N_local - first || (!interleaved_in && count)) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaMemcpyAsync(reg->amp + first, interleaved_in, count * sizeof(double2),
                             cudaMemcpyHostToDevice, reg->stream));
    return QCS_NO_ERROR;
}

extern "C" int qcs_nonzero_states(qcs_register *reg, unsigned long long capacity,
                                  unsigned long long *indices, double *abs_values,
                                  unsigned long long *count)
{
    if (reg && reg->group) {
        if (!count) return QCS_BAD_ARGUMENTS;
        const int world = qcs_group_world(reg);
        std::vector<std::vector<unsigned long long>> idx((size_t) world);
        std::vector<std::vector<double>> mag((size_t) world);
        std::vector<unsigned long long> cnt((size_t) world, 0ull);
        const int rc = qcs_group_run(reg, [&](qcs_register *m) {
            idx[(size_t) m->rank].assign((size_t) capacity, 0ull);
            mag[(size_t) m->rank].assign((size_t) capacity, 0.0);
            return qcs_nonzero_states(m, capacity, idx[(size_t) m->rank].data(), mag[(size_t) m->rank].data(),
                                      &cnt[(size_t) m->rank]);
        });
        unsigned long long total = 0;
        for (int r = 0; r < world; r++) {
            const unsigned long long have = cnt[(size_t) r] < capacity ? cnt[(size_t) r] : capacity;
            for (unsigned long long i = 0; i < have && total + i < capacity; i++) {
                if (indices) indices[total + i] = idx[(size_t) r][(size_t) i];
                if (abs_values) abs_values[total + i] = mag[(size_t) r][(size_t) i];
            }
            total += cnt[(size_t) r];
        }
        *count = total;
        return rc;
    }
    QCS_ENTER(reg);
    if (!count) return QCS_BAD_ARGUMENTS;
    // display_state, testing_and_debug.c:7-26: a console listing, so the
    // amplitudes are streamed to the host in bounded pieces and filtered there
    const uint64_t piece = 1ull << 20;
    double *buf = nullptr;
    QCS_CUDA(cudaHostAlloc((void **) &buf, piece * sizeof(double2), cudaHostAllocDefault));
    unsigned long long total = 0;
    int rc = QCS_NO_ERROR;
    for (uint64_t at = 0; at < reg->N_local && rc == QCS_NO_ERROR; at += piece) {
        const uint64_t len = reg->N_local - at < piece ? reg->N_local - at : piece;
        rc = qcs_get_state(reg, at, len, buf);
        for (uint64_t i = 0; i < len && rc == QCS_NO_ERROR; i++) {
            const double mag = hypot(buf[2 * i], buf[2 * i + 1]);    // gsl_complex_abs
            if (mag != 0.0) {
                if (total < capacity) {
                    if (indices) indices[total] = (uint64_t) reg->rank * reg->N_local + at + i;
                    if (abs_values) abs_values[total] = mag;
                }
                total++;
            }
        }
    }
    cudaFreeHost(buf);
    *count = total;
    return rc;
}

extern "C" int qcs_fill_synthetic(qcs_register *reg, unsigned long long seed)
{
    QCS_GROUP_FORWARD(reg, qcs_fill_synthetic(m, seed));
    // the whole shard is overwritten: whatever is pending (a deferred reset / collapse, recorded gates) is dead
    if (reg && !reg->group && true) {
        reg->lazy_reset = 0;
        reg->queue.clear();
        reg->dense_pending = 0;
    }
    QCS_ENTER(reg);
    return qcs_k_fill_synthetic(reg, seed);
}

extern "C" int qcs_scale(qcs_register *reg, double factor)
{
    QCS_GROUP_FORWARD(reg, qcs_scale(m, factor));
    QCS_ENTER(reg);
    return qcs_k_scale(reg, factor);
}

// ---------------------------------------------------------------------------
// stopwatch and profile
// ---------------------------------------------------------------------------
extern "C" int qcs_timer_start(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_timer_start(m));
    QCS_ENTER(reg);
    QCS_CUDA(cudaEventRecord(reg->timer_begin, reg->stream));
    return QCS_NO_ERROR;
}

extern "C" int qcs_timer_stop(qcs_register *reg, double *milliseconds)
{
    if (reg && reg->group) {
        if (!milliseconds) return QCS_BAD_ARGUMENTS;
        std::vector<double> v((size_t) qcs_group_world(reg), 0.0);
        const int rc = qcs_group_run(reg, [&](qcs_register *m) { return qcs_timer_stop(m, &v[(size_t) m->rank]); });
        *milliseconds = 0.0;
        for (double x : v) *milliseconds = x > *milliseconds ? x : *milliseconds;     // max over the GPUs
        return rc;
    }
    QCS_ENTER(reg);
    if (!milliseconds) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaEventRecord(reg->timer_end, reg->stream));
    QCS_CUDA(cudaEventSynchronize(reg->timer_end));
    float ms = 0.f;
    QCS_CUDA(cudaEventElapsedTime(&ms, reg->timer_begin, reg->timer_end));
    *milliseconds = (double) ms;
    return QCS_NO_ERROR;
}

extern "C" unsigned long long qcs_launch_count(const qcs_register *reg)
{
    if (reg && reg->group) return qcs_launch_count(qcs_group_member(reg, 0));     // per GPU
    return reg ? reg->launches_total : 0;
}

extern "C" int qcs_profile_reset(qcs_register *reg)
{
    QCS_GROUP_FORWARD(reg, qcs_profile_reset(m));
    QCS_ENTER(reg);
    QCS_TRY(qcs_profile_resolve(reg));
    reg->launches_total = 0;
    memset(reg->launches, 0, sizeof reg->launches);
    memset(reg->alg_bytes, 0, sizeof reg->alg_bytes);
    memset(reg->ms, 0, sizeof reg->ms);
    return QCS_NO_ERROR;
}

extern "C" int qcs_profile_get(qcs_register *reg, int k, unsigned long long *launches, double *milliseconds,
                               double *algorithmic_bytes)
{
    if (reg && reg->group) {                     // per GPU: shard 0's counters
        std::vector<unsigned long long> l((size_t) qcs_group_world(reg), 0ull);
        std::vector<double> t((size_t) qcs_group_world(reg), 0.0), b((size_t) qcs_group_world(reg), 0.0);
        const int rc = qcs_group_run(reg, [&](qcs_register *m) {
            return qcs_profile_get(m, k, &l[(size_t) m->rank], &t[(size_t) m->rank], &b[(size_t) m->rank]);
        });
        if (launches) *launches = l[0];
        if (milliseconds) *milliseconds = t[0];
        if (algorithmic_bytes) *algorithmic_bytes = b[0];
        return rc;
    }
    QCS_ENTER(reg);
    if (k < 0 || k >= QCS_K_COUNT) return QCS_BAD_ARGUMENTS;
    QCS_TRY(qcs_profile_resolve(reg));
    if (launches) *launches = reg->launches[k];
    if (milliseconds) *milliseconds = reg->ms[k];
    if (algorithmic_bytes) *algorithmic_bytes = reg->alg_bytes[k];
    return QCS_NO_ERROR;
}
