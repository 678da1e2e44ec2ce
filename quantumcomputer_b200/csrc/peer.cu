// peer.cu -- one virtual address range over all shards of a sharded register.
//
// Each rank owns 2^n_local amplitudes of physical HBM (cuMemCreate).  Every rank
// reserves a virtual range of world * shard bytes and maps shard k -- its own
// allocation or the one imported from rank k -- at offset k * shard, so on every
// GPU the whole 2^n register is ONE array whose index is the reference's basis
// state index (qc_shor.c:150-151): amp_all[i], i < 2^n.  Loads and stores to
// another rank's part travel over NVLink / NVSwitch peer memory.  A tile sweep
// whose tile contains global qubits is then the same kernel as on one GPU
// (csrc/qft_pipeline.cu), run by every rank on its share of the tiles: the TMA
// engine gathers the rows that live on the peers while the consumer groups
// transform the rows that already arrived, i.e. the exchange is fused into the
// sweep instead of preceding it.
//
// The allocation handles travel between the ranks' processes as POSIX file
// descriptors over abstract unix sockets (SCM_RIGHTS); the socket names are
// derived from the register's communicator id, so concurrent registers do not
// collide.  Single node only (the reference path shards over the GPUs of one
// box).  If anything here is unavailable (no P2P path, shard smaller
// than the allocation granularity) the register falls back to a private
// cudaMalloc shard and the NCCL exchange schedule of dist.cu.
#include "qcs_internal.h"

#include <cuda.h>
#include <errno.h>
#include <mutex>
#include <poll.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

struct qcs_peer {
    CUdeviceptr base = 0;
    size_t shard_bytes = 0;
    int world = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    std::vector<char> have;
    std::vector<char> mapped;
};

namespace {

struct driver_api {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

driver_api g_drv;

// the worker threads of a single-caller register (group.cu) arrive here at the same time: resolve once
bool load_driver_once();
bool load_driver()
{
    static std::once_flag once;
    static bool ok = false;
    std::call_once(once, [] { ok = load_driver_once(); });
    return ok;
}

bool load_driver_once()
{
    if (g_drv.ok) return true;
    struct { const char *name; void **slot; } syms[] = {
        {"cuMemCreate", (void **) &g_drv.MemCreate},
        {"cuMemRelease", (void **) &g_drv.MemRelease},
        {"cuMemAddressReserve", (void **) &g_drv.MemAddressReserve},
        {"cuMemAddressFree", (void **) &g_drv.MemAddressFree},
        {"cuMemMap", (void **) &g_drv.MemMap},
        {"cuMemUnmap", (void **) &g_drv.MemUnmap},
        {"cuMemSetAccess", (void **) &g_drv.MemSetAccess},
        {"cuMemExportToShareableHandle", (void **) &g_drv.MemExport},
        {"cuMemImportFromShareableHandle", (void **) &g_drv.MemImport},
        {"cuMemGetAllocationGranularity", (void **) &g_drv.MemGranularity},
    };
    for (auto &s : syms) {
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !*s.slot) {
            cudaGetLastError();
            return false;
        }
    }
    g_drv.ok = true;
    return true;
}

// ---- file descriptors between the ranks' processes -------------------------
void socket_name(const void *comm_id, int rank, sockaddr_un &addr, socklen_t &len)
{
    uint64_t h = 1469598103934665603ull;                 // FNV-1a over the communicator id
    for (int i = 0; i < QCS_COMM_ID_BYTES; i++) h = (h ^ ((const unsigned char *) comm_id)[i]) * 1099511628211ull;
    memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    // abstract namespace: sun_path[0] == 0, nothing is left in the file system
    const int n = snprintf(addr.sun_path + 1, sizeof addr.sun_path - 1, "qcs-peer-%016llx-%d", (unsigned long long) h, rank);
    len = (socklen_t) (offsetof(sockaddr_un, sun_path) + 1 + n);
}

bool send_fd(int sock, int fd)
{
    char byte = 'q';
    iovec iov = {&byte, 1};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof ctrl);
    msghdr msg = {};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof ctrl;
    cmsghdr *c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return sendmsg(sock, &msg, 0) == 1;
}

int recv_fd(int sock)
{
    char byte = 0;
    iovec iov = {&byte, 1};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    msghdr msg = {};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof ctrl;
    pollfd pf = {sock, POLLIN, 0};
    if (poll(&pf, 1, 30000) <= 0) return -1;
    if (recvmsg(sock, &msg, 0) != 1) return -1;
    cmsghdr *c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
    int fd = -1;
    memcpy(&fd, CMSG_DATA(c), sizeof(int));
    return fd;
}

struct server_args {
    int listen_fd;
    int my_fd;
    int clients;
};

void *serve_fd(void *p)
{
    server_args *a = (server_args *) p;
    for (int k = 0; k < a->clients; k++) {
        pollfd pf = {a->listen_fd, POLLIN, 0};
        if (poll(&pf, 1, 30000) <= 0) break;             // a peer gave up: stop waiting for the rest
        const int c = accept(a->listen_fd, nullptr, nullptr);
        if (c < 0) break;
        // the descriptor gives read/write access to the shard: only a process of the same user gets it
        ucred cred = {};
        socklen_t cl = sizeof cred;
        if (getsockopt(c, SOL_SOCKET, SO_PEERCRED, &cred, &cl) == 0 && cred.uid == geteuid()) {
            send_fd(c, a->my_fd);
        } else {
            k--;                                         // not one of the peers: keep waiting for them
        }
        close(c);
    }
    return nullptr;
}

// every rank hands its descriptor to every other rank; fds[k] = descriptor of rank k's shard
bool exchange_fds(const void *comm_id, int rank, int world, int my_fd, std::vector<int> &fds)
{
    fds.assign((size_t) world, -1);
    sockaddr_un addr;
    socklen_t len;
    const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    if (ls < 0) return false;
    socket_name(comm_id, rank, addr, len);
    if (bind(ls, (sockaddr *) &addr, len) != 0 || listen(ls, world) != 0) { close(ls); return false; }
    server_args args = {ls, my_fd, world - 1};
    pthread_t th;
    if (pthread_create(&th, nullptr, serve_fd, &args) != 0) { close(ls); return false; }
    bool ok = true;
    for (int k = 0; k < world && ok; k++) {
        if (k == rank) continue;
        socket_name(comm_id, k, addr, len);
        int s = -1;
        for (int attempt = 0; attempt < 30000; attempt++) {      // the peer may not be listening yet
            s = socket(AF_UNIX, SOCK_STREAM, 0);
            if (s < 0) break;
            if (connect(s, (sockaddr *) &addr, len) == 0) break;
            close(s);
            s = -1;
            timespec ts = {0, 1000000};
            nanosleep(&ts, nullptr);
        }
        if (s < 0) { ok = false; break; }
        fds[(size_t) k] = recv_fd(s);
        close(s);
        if (fds[(size_t) k] < 0) ok = false;
    }
    pthread_join(th, nullptr);
    close(ls);
    return ok;
}

void release(qcs_peer *p)
{
    if (!p) return;
    for (int k = 0; k < p->world; k++) {
        if (p->mapped[(size_t) k]) g_drv.MemUnmap(p->base + (size_t) k * p->shard_bytes, p->shard_bytes);
        if (p->have[(size_t) k]) g_drv.MemRelease(p->handles[(size_t) k]);
    }
    if (p->base) g_drv.MemAddressFree(p->base, p->shard_bytes * (size_t) p->world);
    delete p;
}

}  // namespace

// Can this rank take part in a stitched mapping at all?  No side effects; the ranks agree on the
// answer (api.cu) BEFORE the descriptor exchange, so that a rank that cannot never leaves the
// others waiting on its socket.
bool qcs_peer_can(const qcs_register *reg)
{
    if (getenv("QCS_NO_PEER_MEMORY")) return false;
    if (!load_driver()) return false;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = reg->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (g_drv.MemGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return false;
    return ((size_t) reg->N_local * sizeof(double2)) % gran == 0;
}

// Tries to build the stitched mapping.  Returns true and sets reg->peer / amp / amp_all on
// success; false (nothing allocated) otherwise.  Collective over the ranks only through the
// descriptor exchange: the caller must still agree on the outcome across ranks.
bool qcs_peer_try_alloc(qcs_register *reg, const void *comm_id)
{
    if (getenv("QCS_NO_PEER_MEMORY")) return false;
    if (!load_driver()) return false;
    const size_t shard_bytes = (size_t) reg->N_local * sizeof(double2);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = reg->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (g_drv.MemGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return false;
    const bool size_ok = shard_bytes % gran == 0;

    qcs_peer *p = new qcs_peer();
    p->world = reg->world;
    p->shard_bytes = shard_bytes;
    p->handles.assign((size_t) reg->world, 0);
    p->have.assign((size_t) reg->world, 0);
    p->mapped.assign((size_t) reg->world, 0);
    int my_fd = -1;
    bool ok = size_ok;
    if (ok) ok = g_drv.MemCreate(&p->handles[(size_t) reg->rank], shard_bytes, &prop, 0) == CUDA_SUCCESS;
    if (ok) p->have[(size_t) reg->rank] = 1;
    if (ok) ok = g_drv.MemExport(&my_fd, p->handles[(size_t) reg->rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
    // the exchange runs even after a local failure (with an invalid descriptor) so that no peer
    // waits for this rank longer than it has to
    int placeholder = -1;
    if (!ok) placeholder = dup(0);
    std::vector<int> fds;
    const bool got_all = exchange_fds(comm_id, reg->rank, reg->world, ok ? my_fd : placeholder, fds);
    if (placeholder >= 0) close(placeholder);
    ok = ok && got_all;
    if (ok) ok = g_drv.MemAddressReserve(&p->base, shard_bytes * (size_t) reg->world, gran, 0, 0) == CUDA_SUCCESS;
    for (int k = 0; k < reg->world && ok; k++) {
        if (k != reg->rank) {
            ok = g_drv.MemImport(&p->handles[(size_t) k], (void *) (uintptr_t) fds[(size_t) k],
                                 CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
            if (!ok) break;
            p->have[(size_t) k] = 1;
        }
        ok = g_drv.MemMap(p->base + (size_t) k * shard_bytes, shard_bytes, 0, p->handles[(size_t) k], 0) == CUDA_SUCCESS;
        if (ok) p->mapped[(size_t) k] = 1;
    }
    if (ok) {
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = reg->device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        ok = g_drv.MemSetAccess(p->base, shard_bytes * (size_t) reg->world, &acc, 1) == CUDA_SUCCESS;
    }
    if (my_fd >= 0) close(my_fd);
    for (int fd : fds)
        if (fd >= 0) close(fd);
    if (!ok) {
        release(p);
        return false;
    }
    reg->peer = p;
    reg->amp_all = (double2 *) p->base;
    reg->amp = reg->amp_all + (size_t) reg->rank * reg->N_local;
    return true;
}

void qcs_peer_free(qcs_register *reg)
{
    if (!reg->peer) return;
    release(reg->peer);
    reg->peer = nullptr;
    reg->amp_all = nullptr;
    reg->amp = nullptr;
}
