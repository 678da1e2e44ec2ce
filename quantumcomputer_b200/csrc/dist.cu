// dist.cu -- sharded registers: one process per GPU, the top log2(P) qubits
// are global.  Gates on local qubits and all diagonal gates run
// without communication (api.cu); the only gate of the reference that needs
// an exchange is a Hadamard on a global qubit.
//
// NCCL is resolved with dlopen("libnccl.so.2") at the first sharded create, so
// libqcs.so has no link-time dependency on it and, inside a torch process,
// binds to the NCCL instance torch already loaded.
#include "qcs_internal.h"

#include <dlfcn.h>
#include <mutex>
#include <nccl.h>     // types only; every entry point is looked up at run time
#include <string.h>

namespace {

struct nccl_api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

nccl_api g_nccl;

int load_nccl_once();
int load_nccl()
{
    // threads of one process may create shards at the same time (group.cu)
    static std::once_flag once;
    static int rc = QCS_UNKNOWN_ERROR;
    std::call_once(once, [] { rc = load_nccl_once(); });
    return rc;
}

int load_nccl_once()
{
    if (g_nccl.handle) return QCS_NO_ERROR;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        fprintf(stderr, "qcs: cannot load NCCL (%s); sharded registers are unavailable\n", dlerror());
        return QCS_UNKNOWN_ERROR;
    }
#define QCS_SYM(field, name)                                              \
    *(void **) (&g_nccl.field) = dlsym(h, name);                          \
    if (!g_nccl.field) { fprintf(stderr, "qcs: NCCL symbol %s missing\n", name); dlclose(h); return QCS_UNKNOWN_ERROR; }
    QCS_SYM(GetUniqueId, "ncclGetUniqueId")
    QCS_SYM(CommInitRank, "ncclCommInitRank")
    QCS_SYM(CommDestroy, "ncclCommDestroy")
    QCS_SYM(Send, "ncclSend")
    QCS_SYM(Recv, "ncclRecv")
    QCS_SYM(AllGather, "ncclAllGather")
    QCS_SYM(AllReduce, "ncclAllReduce")
    QCS_SYM(GroupStart, "ncclGroupStart")
    QCS_SYM(GroupEnd, "ncclGroupEnd")
    QCS_SYM(GetErrorString, "ncclGetErrorString")
#undef QCS_SYM
    g_nccl.handle = h;
    return QCS_NO_ERROR;
}

int nccl_fail(ncclResult_t r, const char *what)
{
    fprintf(stderr, "qcs: NCCL error in %s: %s\n", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return QCS_UNKNOWN_ERROR;
}
#define QCS_NCCL(call)                                              \
    do {                                                            \
        ncclResult_t qcs_nr_ = (call);                              \
        if (qcs_nr_ != ncclSuccess) return nccl_fail(qcs_nr_, #call); \
    } while (0)

// combine step of a Hadamard whose target qubit is global.  `mine` holds the
// amplitudes of this rank, `theirs` the partner's copy of the same local
// indices.  Arithmetic and order are those of gates_exact.cu / qc_shor.c:409-412:
//   bit 0 rank:  new = (0 + h*mine)   + h*theirs
//   bit 1 rank:  new = (0 + h*theirs) + (-h)*mine
__global__ void __launch_bounds__(256)
k_hadamard_global_combine(double2 *__restrict__ mine, const double2 *__restrict__ theirs, uint64_t count, int my_bit)
{
    const double h = 0.70710678118654752440;
    const uint64_t stride = (uint64_t) gridDim.x * 256;
    for (uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x; i < count; i += stride) {
        const double2 a = mine[i], b = theirs[i];
        const double2 lo = my_bit ? b : a;        // amplitude with target bit 0
        const double2 hi = my_bit ? a : b;        // amplitude with target bit 1
        const double m1 = my_bit ? -h : h;
        double2 r;
        // h * lo + 0.0 * ... with the imaginary part of H being literally 0.0
        const double t0x = __dsub_rn(__dmul_rn(h, lo.x), __dmul_rn(0.0, lo.y));
        const double t0y = __dadd_rn(__dmul_rn(h, lo.y), __dmul_rn(0.0, lo.x));
        const double t1x = __dsub_rn(__dmul_rn(m1, hi.x), __dmul_rn(0.0, hi.y));
        const double t1y = __dadd_rn(__dmul_rn(m1, hi.y), __dmul_rn(0.0, hi.x));
        r.x = __dadd_rn(__dadd_rn(0.0, t0x), t1x);
        r.y = __dadd_rn(__dadd_rn(0.0, t0y), t1y);
        mine[i] = r;
    }
}

// the same for an arbitrary 2x2 gate (gates_general.cu): new = A * mine + B * theirs with
// (A, B) = (U00, U01) on the rank whose target bit is 0 and (U11, U10) on the other one;
// ctrl_mask != 0 restricts the update to local indices with that bit set
__global__ void __launch_bounds__(256)
k_gate_global_combine(double2 *__restrict__ mine, const double2 *__restrict__ theirs, uint64_t count, uint64_t first_index,
                      double2 A, double2 B, uint64_t ctrl_mask)
{
    const uint64_t stride = (uint64_t) gridDim.x * 256;
    for (uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x; i < count; i += stride) {
        if (ctrl_mask && !((first_index + i) & ctrl_mask)) continue;
        const double2 a = mine[i], b = theirs[i];
        mine[i] = make_double2(A.x * a.x - A.y * a.y + B.x * b.x - B.y * b.y,
                               A.x * a.y + A.y * a.x + B.x * b.y + B.y * b.x);
    }
}

}  // namespace

constexpr int kGatherMax = 4;      // doubles per rank in one small all-gather

struct qcs_dist {
    ncclComm_t comm = nullptr;
    cudaStream_t comm_stream = nullptr;
    double2 *staging[2] = {nullptr, nullptr};
    uint64_t staging_amps = 0;
    cudaEvent_t recv_done[2] = {nullptr, nullptr};
    cudaEvent_t buf_free[2] = {nullptr, nullptr};
    cudaEvent_t ready = nullptr;
    double *d_gather = nullptr;     // (world + 1) * kGatherMax doubles: receive area + send slot
    // global<->local exchange slices: kSlices buffers of world * 2^slice_bits amplitudes
    double2 *slice[3] = {nullptr, nullptr, nullptr};
    unsigned slice_bits = 0;
    cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr};     // slice received
    cudaEvent_t ev_done[3] = {nullptr, nullptr, nullptr};   // slice transformed
    cudaEvent_t ev_tail = nullptr;
    double *h_gather = nullptr;     // pinned
    cudaEvent_t ev_slice[16] = {};  // global-sweep slice finished on every rank (overlapped schedule)
};

extern "C" int qcs_comm_unique_id(void *id_out)
{
    if (!id_out) return QCS_BAD_ARGUMENTS;
    QCS_TRY(load_nccl());
    ncclUniqueId id;
    QCS_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == QCS_COMM_ID_BYTES, "id size");
    memcpy(id_out, &id, sizeof id);
    return QCS_NO_ERROR;
}

int qcs_dist_init(qcs_register *reg, const void *comm_id)
{
    QCS_TRY(load_nccl());
    qcs_dist *d = new qcs_dist();
    reg->dist = d;
    ncclUniqueId id;
    memcpy(&id, comm_id, sizeof id);
    QCS_NCCL(g_nccl.CommInitRank(&d->comm, reg->world, id, reg->rank));
    QCS_CUDA(cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
    // bounded staging: two buffers of at most 2^24 amplitudes (256 MiB) each
    d->staging_amps = reg->N_local < (1ull << 24) ? reg->N_local : (1ull << 24);
    for (int b = 0; b < 2; b++) {
        QCS_CUDA(cudaMalloc((void **) &d->staging[b], d->staging_amps * sizeof(double2)));
        QCS_CUDA(cudaEventCreateWithFlags(&d->recv_done[b], cudaEventDisableTiming));
        QCS_CUDA(cudaEventCreateWithFlags(&d->buf_free[b], cudaEventDisableTiming));
    }
    QCS_CUDA(cudaEventCreateWithFlags(&d->ready, cudaEventDisableTiming));
    QCS_CUDA(cudaMalloc((void **) &d->d_gather, (size_t) (reg->world + 1) * kGatherMax * sizeof(double)));
    QCS_CUDA(cudaMemset(d->d_gather, 0, (size_t) (reg->world + 1) * kGatherMax * sizeof(double)));
    QCS_CUDA(cudaHostAlloc((void **) &d->h_gather, (size_t) (reg->world + 1) * kGatherMax * sizeof(double), cudaHostAllocDefault));
    return QCS_NO_ERROR;
}

void qcs_dist_destroy(qcs_register *reg)
{
    qcs_dist *d = reg->dist;
    if (!d) return;
    if (d->comm_stream) cudaStreamSynchronize(d->comm_stream);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    for (int b = 0; b < 3; b++) {
        if (d->slice[b]) cudaFree(d->slice[b]);
        if (d->ev_in[b]) cudaEventDestroy(d->ev_in[b]);
        if (d->ev_done[b]) cudaEventDestroy(d->ev_done[b]);
    }
    if (d->ev_tail) cudaEventDestroy(d->ev_tail);
    for (int j = 0; j < 16; j++)
        if (d->ev_slice[j]) cudaEventDestroy(d->ev_slice[j]);
    for (int b = 0; b < 2; b++) {
        if (d->staging[b]) cudaFree(d->staging[b]);
        if (d->recv_done[b]) cudaEventDestroy(d->recv_done[b]);
        if (d->buf_free[b]) cudaEventDestroy(d->buf_free[b]);
    }
    if (d->ready) cudaEventDestroy(d->ready);
    if (d->d_gather) cudaFree(d->d_gather);
    if (d->h_gather) cudaFreeHost(d->h_gather);
    if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
    delete d;
    reg->dist = nullptr;
}

// every rank contributes `count` <= kGatherMax doubles; all_host[r * count + i] = rank r's i-th
int qcs_dist_allgather_doubles(qcs_register *reg, const double *mine, int count, double *all_host)
{
    qcs_dist *d = reg->dist;
    if (!d || count < 1 || count > kGatherMax) return QCS_BAD_ARGUMENTS;
    const size_t w = (size_t) reg->world * kGatherMax;                    // send slot behind the receive area
    for (int i = 0; i < count; i++) d->h_gather[w + (size_t) i] = mine[i];
    QCS_CUDA(cudaMemcpyAsync(d->d_gather + w, d->h_gather + w, (size_t) count * sizeof(double),
                             cudaMemcpyHostToDevice, reg->stream));
    QCS_NCCL(g_nccl.AllGather(d->d_gather + w, d->d_gather, (size_t) count, ncclDouble, d->comm, reg->stream));
    QCS_CUDA(cudaMemcpyAsync(d->h_gather, d->d_gather, (size_t) reg->world * (size_t) count * sizeof(double),
                             cudaMemcpyDeviceToHost, reg->stream));
    QCS_CUDA(cudaStreamSynchronize(reg->stream));
    memcpy(all_host, d->h_gather, (size_t) reg->world * (size_t) count * sizeof(double));
    return QCS_NO_ERROR;
}

int qcs_dist_allgather_double(qcs_register *reg, double mine, double *all_host)
{
    return qcs_dist_allgather_doubles(reg, &mine, 1, all_host);
}

int qcs_dist_barrier(qcs_register *reg)
{
    if (reg->world == 1) return QCS_NO_ERROR;
    std::vector<double> all((size_t) reg->world);
    return qcs_dist_allgather_double(reg, 0.0, all.data());
}

int qcs_dist_barrier_on(qcs_register *reg, cudaStream_t stream)
{
    qcs_dist *d = reg->dist;
    if (!d) return QCS_NO_ERROR;
    // a one-element all-reduce: its kernel on a rank completes only after every rank has
    // reached it, i.e. after everything queued before it on every rank's stream
    reg->launches_total++;
    reg->launches[QCS_K_EXCHANGE]++;
    QCS_NCCL(g_nccl.AllReduce(d->d_gather, d->d_gather, 1, ncclDouble, ncclSum, d->comm, stream));
    return QCS_NO_ERROR;
}

int qcs_dist_stream_barrier(qcs_register *reg) { return qcs_dist_barrier_on(reg, reg->stream); }

cudaStream_t qcs_dist_side_stream(qcs_register *reg) { return reg->dist ? reg->dist->comm_stream : nullptr; }

int qcs_dist_slice_event(qcs_register *reg, int j, cudaEvent_t *ev)
{
    qcs_dist *d = reg->dist;
    if (!d || j < 0 || j >= 16) return QCS_BAD_ARGUMENTS;
    if (!d->ev_slice[j]) QCS_CUDA(cudaEventCreateWithFlags(&d->ev_slice[j], cudaEventDisableTiming));
    *ev = d->ev_slice[j];
    return QCS_NO_ERROR;
}

// Hadamard on a global qubit: pairwise exchange with rank ^ 2^(q - n_local),
// chunked through two staging buffers so that the combine kernel of chunk c
// overlaps the NVLink transfer of chunk c+1.
int qcs_dist_hadamard_global(qcs_register *reg, unsigned q)
{
    qcs_dist *d = reg->dist;
    if (!d || q < reg->n_local || q >= reg->n) return QCS_BAD_ARGUMENTS;
    const int bitpos = (int) (q - reg->n_local);
    const int partner = reg->rank ^ (1 << bitpos);
    const int my_bit = (reg->rank >> bitpos) & 1;
    const uint64_t chunk = d->staging_amps;
    const uint64_t n_chunks = reg->N_local / chunk;

    // the exchange reads amplitudes produced by earlier kernels on the compute stream
    QCS_CUDA(cudaEventRecord(d->ready, reg->stream));
    QCS_CUDA(cudaStreamWaitEvent(d->comm_stream, d->ready, 0));
    for (uint64_t c = 0; c < n_chunks; c++) {
        const int b = (int) (c & 1);
        double2 *mine = reg->amp + c * chunk;
        if (c >= 2) QCS_CUDA(cudaStreamWaitEvent(d->comm_stream, d->buf_free[b], 0));
        reg->launches_total++;
        reg->launches[QCS_K_EXCHANGE]++;
        reg->alg_bytes[QCS_K_EXCHANGE] += 16.0 * (double) chunk;
        QCS_NCCL(g_nccl.GroupStart());
        QCS_NCCL(g_nccl.Send(mine, chunk * 2, ncclDouble, partner, d->comm, d->comm_stream));
        QCS_NCCL(g_nccl.Recv(d->staging[b], chunk * 2, ncclDouble, partner, d->comm, d->comm_stream));
        QCS_NCCL(g_nccl.GroupEnd());
        QCS_CUDA(cudaEventRecord(d->recv_done[b], d->comm_stream));
        QCS_CUDA(cudaStreamWaitEvent(reg->stream, d->recv_done[b], 0));
        uint64_t grid = (chunk + 255) / 256;
        const uint64_t cap = (uint64_t) reg->sm_count * 8;
        if (grid > cap) grid = cap;
        qcs_launch_begin(reg, QCS_K_HADAMARD, 48.0 * (double) chunk);
        k_hadamard_global_combine<<<(unsigned) grid, 256, 0, reg->stream>>>(mine, d->staging[b], chunk, my_bit);
        QCS_TRY(qcs_launch_end(reg, QCS_K_HADAMARD, "k_hadamard_global_combine"));
        QCS_CUDA(cudaEventRecord(d->buf_free[b], reg->stream));
    }
    return QCS_NO_ERROR;
}


// Arbitrary 2x2 gate on a global target qubit, optionally controlled (c < 0: no control), for
// registers without peer memory: the same chunked pairwise exchange as the Hadamard above.
int qcs_dist_gate_global(qcs_register *reg, unsigned q, int c, const double *u)
{
    qcs_dist *d = reg->dist;
    if (!d || q < reg->n_local || q >= reg->n) return QCS_BAD_ARGUMENTS;
    const int bitpos = (int) (q - reg->n_local);
    const int partner = reg->rank ^ (1 << bitpos);
    const int my_bit = (reg->rank >> bitpos) & 1;
    uint64_t ctrl_mask = 0;
    if (c >= 0) {
        if ((unsigned) c >= reg->n_local) {
            // a global control is the same constant bit on both partners
            if (!(((unsigned) reg->rank >> ((unsigned) c - reg->n_local)) & 1u)) return QCS_NO_ERROR;
        } else {
            ctrl_mask = 1ull << c;
        }
    }
    const double2 A = my_bit ? make_double2(u[6], u[7]) : make_double2(u[0], u[1]);
    const double2 B = my_bit ? make_double2(u[4], u[5]) : make_double2(u[2], u[3]);
    const uint64_t chunk = d->staging_amps;
    const uint64_t n_chunks = reg->N_local / chunk;
    QCS_CUDA(cudaEventRecord(d->ready, reg->stream));
    QCS_CUDA(cudaStreamWaitEvent(d->comm_stream, d->ready, 0));
    for (uint64_t k = 0; k < n_chunks; k++) {
        const int b = (int) (k & 1);
        double2 *mine = reg->amp + k * chunk;
        if (k >= 2) QCS_CUDA(cudaStreamWaitEvent(d->comm_stream, d->buf_free[b], 0));
        reg->launches_total++;
        reg->launches[QCS_K_EXCHANGE]++;
        reg->alg_bytes[QCS_K_EXCHANGE] += 16.0 * (double) chunk;
        QCS_NCCL(g_nccl.GroupStart());
        QCS_NCCL(g_nccl.Send(mine, chunk * 2, ncclDouble, partner, d->comm, d->comm_stream));
        QCS_NCCL(g_nccl.Recv(d->staging[b], chunk * 2, ncclDouble, partner, d->comm, d->comm_stream));
        QCS_NCCL(g_nccl.GroupEnd());
        QCS_CUDA(cudaEventRecord(d->recv_done[b], d->comm_stream));
        QCS_CUDA(cudaStreamWaitEvent(reg->stream, d->recv_done[b], 0));
        uint64_t grid = (chunk + 255) / 256;
        const uint64_t cap = (uint64_t) reg->sm_count * 8;
        if (grid > cap) grid = cap;
        qcs_launch_begin(reg, QCS_K_GATE_1Q, 48.0 * (double) chunk);
        k_gate_global_combine<<<(unsigned) grid, 256, 0, reg->stream>>>(mine, d->staging[b], chunk, k * chunk, A, B, ctrl_mask);
        QCS_TRY(qcs_launch_end(reg, QCS_K_GATE_1Q, "k_gate_global_combine"));
        QCS_CUDA(cudaEventRecord(d->buf_free[b], reg->stream));
    }
    // the partner must not overwrite amplitudes this rank is still sending: the sends above read
    // `mine` on the comm stream while the combine of the same chunk waits for recv_done, which
    // follows the send in the same NCCL group
    return QCS_NO_ERROR;
}

// ---------------------------------------------------------------------------
// Stages on the global qubits [n_local, n) -- the first p stages of the
// inverse QFT (qc_shor.c:682-689), the last p of the forward one, or the bare
// Hadamards of quantum_computation -- by qubit-swap relabelling:
//
//   rank r, local index (s, m)  [s = top p local bits, m = the rest]
//   exchange in : piece (r; s, m-slice) goes to rank s, so rank r ends up with
//                 the amplitudes (s'; r, m-slice) of every rank s' : for its
//                 value r of the swapped-out bits, all values of the global ones
//   sweep       : one tile sweep over those p bits (qcs_fused_top_sweep); the
//                 phases see the swapped-out bits through y_const
//   exchange out: the transformed pieces return to where they came from.
//
// Each rank sends (P-1)/P of its shard twice (vs. p full-shard pairwise
// exchanges gate by gate).  The three phases are pipelined over slices of m
// through three staging buffers: while slice u is transformed, slice u+1 is
// arriving and slice u-1 is leaving over NVLink; nothing is staged through
// host memory and the shard is updated in place.
// ---------------------------------------------------------------------------
int qcs_dist_top_stages(qcs_register *reg, unsigned lo, bool inverse, bool hadamard_only)
{
    qcs_dist *d = reg->dist;
    if (!d) return QCS_BAD_ARGUMENTS;
    const unsigned p = (unsigned) reg->p_global;
    const int P = reg->world, r = reg->rank;
    if (reg->n_local < 2 * p) return QCS_BAD_ARGUMENTS;
    const unsigned q = reg->n_local - p;                 // bits of m
    if (lo > q) {
        fprintf(stderr, "qcs: sharded fused QFT needs M_size <= n_local - log2(world)\n");
        return QCS_BAD_ARGUMENTS;
    }
    if (!d->slice[0]) {
        unsigned c = q < 26u - p ? q : 26u - p;          // at most 1 GiB per buffer
        if (q >= 2 && c > q - 2) c = q - 2;              // at least 4 slices so the phases overlap
        d->slice_bits = c;
        for (int b = 0; b < 3; b++) {
            QCS_CUDA(cudaMalloc((void **) &d->slice[b], ((size_t) P << c) * sizeof(double2)));
            QCS_CUDA(cudaEventCreateWithFlags(&d->ev_in[b], cudaEventDisableTiming));
            QCS_CUDA(cudaEventCreateWithFlags(&d->ev_done[b], cudaEventDisableTiming));
        }
        QCS_CUDA(cudaEventCreateWithFlags(&d->ev_tail, cudaEventDisableTiming));
    }
    const unsigned c = d->slice_bits;
    const uint64_t Sc = 1ull << c, B = 1ull << q;
    const uint64_t n_slices = B >> c;

    // the exchange reads amplitudes produced by earlier kernels on the compute stream
    QCS_CUDA(cudaEventRecord(d->ready, reg->stream));
    QCS_CUDA(cudaStreamWaitEvent(d->comm_stream, d->ready, 0));

    auto exchange = [&](uint64_t u, bool inbound) -> int {
        const int b = (int) (u % 3);
        reg->launches_total++;
        reg->launches[QCS_K_EXCHANGE]++;
        reg->alg_bytes[QCS_K_EXCHANGE] += 16.0 * (double) Sc * (double) (P - 1);
        QCS_NCCL(g_nccl.GroupStart());
        for (int s = 0; s < P; s++) {
            if (s == r) continue;
            double2 *mine = reg->amp + (uint64_t) s * B + u * Sc;       // piece (r; s, slice u)
            double2 *stg = d->slice[b] + (uint64_t) s * Sc;             // slot of rank s in the staging slice
            if (inbound) {
                QCS_NCCL(g_nccl.Send(mine, Sc * 2, ncclDouble, s, d->comm, d->comm_stream));
                QCS_NCCL(g_nccl.Recv(stg, Sc * 2, ncclDouble, s, d->comm, d->comm_stream));
            } else {
                QCS_NCCL(g_nccl.Send(stg, Sc * 2, ncclDouble, s, d->comm, d->comm_stream));
                QCS_NCCL(g_nccl.Recv(mine, Sc * 2, ncclDouble, s, d->comm, d->comm_stream));
            }
        }
        QCS_NCCL(g_nccl.GroupEnd());
        return QCS_NO_ERROR;
    };

    for (uint64_t u = 0; u <= n_slices; u++) {
        if (u < n_slices) {
            const int b = (int) (u % 3);
            // slice u in: the staging buffer was last sent from by exchange-out(u-3), earlier on this stream
            QCS_TRY(exchange(u, true));
            QCS_CUDA(cudaEventRecord(d->ev_in[b], d->comm_stream));
            // compute stream: own piece into the slice, transform, own piece back
            double2 *own = reg->amp + (uint64_t) r * B + u * Sc;
            QCS_CUDA(cudaMemcpyAsync(d->slice[b] + (uint64_t) r * Sc, own, Sc * sizeof(double2),
                                     cudaMemcpyDeviceToDevice, reg->stream));
            QCS_CUDA(cudaStreamWaitEvent(reg->stream, d->ev_in[b], 0));
            const unsigned long long y_const = (((unsigned long long) u << c) >> lo) +
                                               ((unsigned long long) r << (q - lo));
            QCS_TRY(qcs_fused_top_sweep(reg, d->slice[b], c, p, lo, y_const, inverse, hadamard_only, reg->stream));
            QCS_CUDA(cudaMemcpyAsync(own, d->slice[b] + (uint64_t) r * Sc, Sc * sizeof(double2),
                                     cudaMemcpyDeviceToDevice, reg->stream));
            QCS_CUDA(cudaEventRecord(d->ev_done[b], reg->stream));
        }
        if (u >= 1) {
            const int b = (int) ((u - 1) % 3);
            QCS_CUDA(cudaStreamWaitEvent(d->comm_stream, d->ev_done[b], 0));
            QCS_TRY(exchange(u - 1, false));
        }
    }
    // later kernels on the compute stream see the returned pieces
    QCS_CUDA(cudaEventRecord(d->ev_tail, d->comm_stream));
    QCS_CUDA(cudaStreamWaitEvent(reg->stream, d->ev_tail, 0));
    return QCS_NO_ERROR;
}
