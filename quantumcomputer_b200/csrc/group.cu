// group.cu -- a sharded register driven from ONE host thread.
//
// The reference's main() allocates its register once (qc_shor.c:1316-1324) and find_period calls
// reset_register / quantum_computation / measure_state on it (qc_shor.c:922-928) from a single
// thread.  qcs_register_create_multi(&reg, L, M, n_gpus) gives that caller a register sharded over
// n_gpus devices of the box without any launcher: the handle it returns is a FACADE over one
// ordinary per-device shard register per GPU (exactly what qcs_register_create_sharded builds in
// the one-process-per-GPU mode), each owned by a worker thread of this library.  Every C-ABI call
// on the facade is handed to all workers at once -- the shards' collectives (stream barriers,
// all-gathers, the peer-memory sweeps) therefore meet the way they do across processes -- and
// returns when all of them have: the caller sees the calling convention of a single-GPU register.
// Amplitude indices of bulk calls (qcs_get_state / qcs_set_state) address the WHOLE register.
#include "qcs_internal.h"

#include <condition_variable>
#include <mutex>
#include <thread>

struct qcs_group {
    int world = 0;
    std::vector<qcs_register *> member;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    unsigned long long generation = 0;
    const std::function<int(qcs_register *, int)> *job = nullptr;
    std::vector<int> rc;
    int pending = 0;
    bool quit = false;
};

namespace {

void worker_main(qcs_group *g, int rank)
{
    cudaSetDevice(rank);
    unsigned long long seen = 0;
    for (;;) {
        const std::function<int(qcs_register *, int)> *job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_job.wait(lk, [&] { return g->quit || g->generation != seen; });
            if (g->quit) return;
            seen = g->generation;
            job = g->job;
        }
        const int rc = (*job)(g->member[(size_t) rank], rank);
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->rc[(size_t) rank] = rc;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

int run_all(qcs_group *g, const std::function<int(qcs_register *, int)> &job)
{
    std::unique_lock<std::mutex> lk(g->mu);
    g->job = &job;
    g->pending = g->world;
    g->generation++;
    g->cv_job.notify_all();
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
    for (int r = 0; r < g->world; r++)
        if (g->rc[(size_t) r] != QCS_NO_ERROR) return g->rc[(size_t) r];
    return QCS_NO_ERROR;
}

}  // namespace

int qcs_group_run(qcs_register *facade, const std::function<int(qcs_register *)> &job)
{
    return run_all(facade->group, [&](qcs_register *m, int) { return job(m); });
}

int qcs_group_world(const qcs_register *facade) { return facade->group->world; }
qcs_register *qcs_group_member(const qcs_register *facade, int rank) { return facade->group->member[(size_t) rank]; }

void qcs_group_destroy(qcs_register *facade)
{
    qcs_group *g = facade->group;
    // every shard is destroyed on the thread (device) that owns it
    run_all(g, [&](qcs_register *m, int rank) {
        if (m) qcs_register_destroy(m);
        g->member[(size_t) rank] = nullptr;
        return QCS_NO_ERROR;
    });
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
    }
    g->cv_job.notify_all();
    for (std::thread &t : g->workers) t.join();
    delete g;
    delete facade;
}

// main(), qc_shor.c:1316-1324, for a register spread over the GPUs of the box
extern "C" int qcs_register_create_multi(qcs_register **out, int L_size, int M_size, int n_gpus)
{
    if (!out) return QCS_BAD_ARGUMENTS;
    *out = nullptr;
    if (n_gpus < 1 || (n_gpus & (n_gpus - 1)) != 0) return QCS_BAD_ARGUMENTS;
    if (n_gpus == 1) return qcs_register_create(out, L_size, M_size, -1);
    if (L_size < 0 || M_size < 0 || L_size + M_size < 1 || L_size + M_size > 62) return QCS_BAD_ARGUMENTS;
    int p = 0;
    while ((1 << p) < n_gpus) p++;
    if (p >= L_size + M_size) return QCS_BAD_ARGUMENTS;
    if (qcs_device_count() < n_gpus) {
        fprintf(stderr, "qcs: %d GPUs requested, %d visible; this library has no CPU path\n", n_gpus, qcs_device_count());
        return QCS_UNKNOWN_ERROR;
    }
    unsigned char id[QCS_COMM_ID_BYTES];
    QCS_TRY(qcs_comm_unique_id(id));

    qcs_register *facade = new (std::nothrow) qcs_register();
    qcs_group *g = new (std::nothrow) qcs_group();
    if (!facade || !g) { delete facade; delete g; return QCS_INSUFFICIENT_MEMORY; }
    facade->group = g;
    facade->L_size = L_size;
    facade->M_size = M_size;
    facade->n = (unsigned) (L_size + M_size);
    facade->N = 1ull << facade->n;
    // to its one caller the whole register is "local": bulk indices run over all 2^n amplitudes
    facade->n_local = facade->n;
    facade->N_local = facade->N;
    facade->rank = 0;
    facade->world = 1;
    facade->p_global = 0;
    facade->device = 0;
    g->world = n_gpus;
    g->member.assign((size_t) n_gpus, nullptr);
    g->rc.assign((size_t) n_gpus, QCS_NO_ERROR);
    for (int r = 0; r < n_gpus; r++) g->workers.emplace_back(worker_main, g, r);
    const int rc = run_all(g, [&](qcs_register *, int rank) {
        qcs_register *m = nullptr;
        const int e = qcs_register_create_sharded(&m, L_size, M_size, rank, rank, n_gpus, id);
        g->member[(size_t) rank] = m;
        return e;
    });
    if (rc != QCS_NO_ERROR) {
        qcs_group_destroy(facade);
        return rc;
    }
    *out = facade;
    return QCS_NO_ERROR;
}

extern "C" int qcs_num_gpus(const qcs_register *reg)
{
    if (!reg) return 0;
    return reg->group ? reg->group->world : reg->world;
}
