// state_io.cu -- state dump / load (checkpoint).
//
// File layout: a 64-byte little-endian header followed by this shard's amplitudes
// as raw interleaved (re, im) doubles -- the layout of gsl_vector_complex.data
// that the reference keeps on the host (qc_shor.c:385-386), so a dump of a
// single-GPU register is byte-for-byte what the reference's state vector would
// be.  A sharded register writes one file per rank ("<path>.rank<k>").  The
// amplitudes stream through two pinned 64 MiB buffers so the device->host copy
// of one piece overlaps the file I/O of the previous one.
#include "qcs_internal.h"

#include <string.h>
#include <string>

namespace {

struct file_header {
    char magic[8];              // "QCSSTATE"
    uint32_t version;           // 1
    int32_t L_size, M_size;
    int32_t world, rank;
    uint32_t n_local;
    uint64_t n_amplitudes;      // in this file
    unsigned char pad[24];
};
static_assert(sizeof(file_header) == 64, "header is 64 bytes");

constexpr uint64_t kPiece = 1ull << 22;     // amplitudes per piece: 64 MiB

std::string shard_path(const qcs_register *reg, const char *path)
{
    std::string p(path);
    if (reg->world > 1) p += ".rank" + std::to_string(reg->rank);
    return p;
}

struct pinned_pair {
    double2 *buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    ~pinned_pair()
    {
        for (int b = 0; b < 2; b++) {
            if (buf[b]) cudaFreeHost(buf[b]);
            if (done[b]) cudaEventDestroy(done[b]);
        }
    }
    int init(uint64_t piece)
    {
        for (int b = 0; b < 2; b++) {
            QCS_CUDA(cudaHostAlloc((void **) &buf[b], piece * sizeof(double2), cudaHostAllocDefault));
            QCS_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
        }
        return QCS_NO_ERROR;
    }
};

}  // namespace

extern "C" int qcs_save_state(qcs_register *reg, const char *path)
{
    QCS_GROUP_FORWARD(reg, qcs_save_state(m, path));
    if (!reg || !path) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(reg->device));
    QCS_TRY(qcs_fuse_flush(reg));
    FILE *f = fopen(shard_path(reg, path).c_str(), "wb");
    if (!f) return QCS_BAD_ARGUMENTS;
    file_header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "QCSSTATE", 8);
    h.version = 1;
    h.L_size = reg->L_size;
    h.M_size = reg->M_size;
    h.world = reg->world;
    h.rank = reg->rank;
    h.n_local = reg->n_local;
    h.n_amplitudes = reg->N_local;
    int rc = fwrite(&h, sizeof h, 1, f) == 1 ? QCS_NO_ERROR : QCS_UNKNOWN_ERROR;
    const uint64_t piece = reg->N_local < kPiece ? reg->N_local : kPiece;
    pinned_pair pp;
    if (rc == QCS_NO_ERROR) rc = pp.init(piece);
    const uint64_t n_pieces = (reg->N_local + piece - 1) / piece;
    // piece k+1 is copied while piece k is written
    for (uint64_t k = 0; k <= n_pieces && rc == QCS_NO_ERROR; k++) {
        if (k < n_pieces) {
            const int b = (int) (k & 1);
            const uint64_t len = std::min<uint64_t>(piece, reg->N_local - k * piece);
            if (cudaMemcpyAsync(pp.buf[b], reg->amp + k * piece, len * sizeof(double2), cudaMemcpyDeviceToHost, reg->stream) != cudaSuccess ||
                cudaEventRecord(pp.done[b], reg->stream) != cudaSuccess)
                rc = QCS_UNKNOWN_ERROR;
        }
        if (k >= 1 && rc == QCS_NO_ERROR) {
            const int b = (int) ((k - 1) & 1);
            const uint64_t len = std::min<uint64_t>(piece, reg->N_local - (k - 1) * piece);
            if (cudaEventSynchronize(pp.done[b]) != cudaSuccess || fwrite(pp.buf[b], sizeof(double2), len, f) != len)
                rc = QCS_UNKNOWN_ERROR;
        }
    }
    cudaStreamSynchronize(reg->stream);
    if (fclose(f) != 0 && rc == QCS_NO_ERROR) rc = QCS_UNKNOWN_ERROR;
    return rc;
}

extern "C" int qcs_load_state(qcs_register *reg, const char *path)
{
    QCS_GROUP_FORWARD(reg, qcs_load_state(m, path));
    if (!reg || !path) return QCS_BAD_ARGUMENTS;
    QCS_CUDA(cudaSetDevice(reg->device));
    QCS_TRY(qcs_fuse_flush(reg));
    FILE *f = fopen(shard_path(reg, path).c_str(), "rb");
    if (!f) return QCS_BAD_ARGUMENTS;
    file_header h;
    int rc = QCS_NO_ERROR;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "QCSSTATE", 8) != 0 || h.version != 1) rc = QCS_BAD_ARGUMENTS;
    // the file must describe exactly this register and shard
    if (rc == QCS_NO_ERROR && (h.L_size != reg->L_size || h.M_size != reg->M_size || h.world != reg->world ||
                               h.rank != reg->rank || h.n_amplitudes != reg->N_local))
        rc = QCS_BAD_ARGUMENTS;
    const uint64_t piece = reg->N_local < kPiece ? reg->N_local : kPiece;
    pinned_pair pp;
    if (rc == QCS_NO_ERROR) rc = pp.init(piece);
    const uint64_t n_pieces = (reg->N_local + piece - 1) / piece;
    for (uint64_t k = 0; k < n_pieces && rc == QCS_NO_ERROR; k++) {
        const int b = (int) (k & 1);
        const uint64_t len = std::min<uint64_t>(piece, reg->N_local - k * piece);
        // the buffer was last used by the copy of piece k-2
        if (k >= 2 && cudaEventSynchronize(pp.done[b]) != cudaSuccess) rc = QCS_UNKNOWN_ERROR;
        if (rc == QCS_NO_ERROR && fread(pp.buf[b], sizeof(double2), len, f) != len) rc = QCS_BAD_ARGUMENTS;
        if (rc == QCS_NO_ERROR &&
            (cudaMemcpyAsync(reg->amp + k * piece, pp.buf[b], len * sizeof(double2), cudaMemcpyHostToDevice, reg->stream) != cudaSuccess ||
             cudaEventRecord(pp.done[b], reg->stream) != cudaSuccess))
            rc = QCS_UNKNOWN_ERROR;
    }
    cudaStreamSynchronize(reg->stream);
    fclose(f);
    return rc;
}
