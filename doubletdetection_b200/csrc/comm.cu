// comm.cu -- cell-block sharding of ONE fit over the GPUs of a box (SURVEY 8e level 2, BASELINE config c5).
//
// Every rank holds the raw CSR and the parents, builds the dense rows of its block of cells, and the PCA's
// reductions over cells (column sums, D^T Y, the Gram matrix of the tall panel) become NCCL all-reduces of
// G x L / L x L float64 blocks; the low-dimensional embedding and the kNN lists are all-gathered so that
// every rank sees the whole graph.  NCCL is taken from the process at run time (dlopen of libnccl.so.2 --
// the copy torch already loaded when the host side uses torch.distributed for its rendezvous): the library
// does not link against it, and a single-GPU user never needs it.
#include "dd_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

template <typename F>
bool sym(void *lib, const char *name, F &out) {
    out = reinterpret_cast<F>(dlsym(lib, name));
    return out != nullptr;
}

void load_nccl() {
    const char *env = getenv("DD_NCCL_LIB");
    void *lib = nullptr;
    if (env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy already in the process (torch's)
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        g_nccl.err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return;
    }
    NcclApi &a = g_nccl;
    if (!(sym(lib, "ncclGetUniqueId", a.GetUniqueId) && sym(lib, "ncclCommInitRank", a.CommInitRank) &&
          sym(lib, "ncclCommDestroy", a.CommDestroy) && sym(lib, "ncclAllReduce", a.AllReduce) &&
          sym(lib, "ncclBroadcast", a.Broadcast) && sym(lib, "ncclGroupStart", a.GroupStart) &&
          sym(lib, "ncclGroupEnd", a.GroupEnd) && sym(lib, "ncclGetErrorString", a.GetErrorString))) {
        a.err = "libnccl.so.2 lacks an expected symbol";
        return;
    }
    a.lib = lib;
}

int nccl_ready(dd_handle *h) {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.lib) return dd_fail(h, DD_ERR_UNSUPPORTED, "cell-block sharding: " + g_nccl.err);
    return DD_OK;
}

#define DD_NCCL(h, expr)                                                                                  \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess)                                                                            \
            return dd_fail((h), DD_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));      \
    } while (0)

}  // namespace

void dd_set_block(dd_handle *h) {
    if (dd_sharded(h)) {
        h->blk_n0 = h->N * h->rank / h->world;
        h->blk_n = h->N * (h->rank + 1) / h->world - h->blk_n0;
        h->blk_m0 = h->M * h->rank / h->world;
        h->blk_m = h->M * (h->rank + 1) / h->world - h->blk_m0;
    } else {
        h->blk_n0 = 0; h->blk_n = h->N; h->blk_m0 = 0; h->blk_m = h->M;
    }
    h->A = h->blk_n + h->blk_m;
    h->A_glob = h->N + h->M;
}

extern "C" int dd_comm_unique_id(void *id_out, int64_t id_bytes) {
    if (!id_out || id_bytes < (int64_t)sizeof(ncclUniqueId))
        return dd_fail(nullptr, DD_ERR_ARG, "dd_comm_unique_id: buffer smaller than DD_COMM_ID_BYTES");
    DD_TRY(nccl_ready(nullptr));
    ncclUniqueId id;
    DD_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memset(id_out, 0, (size_t)id_bytes);
    memcpy(id_out, &id, sizeof(id));
    return DD_OK;
}

extern "C" int dd_comm_init(dd_handle *h, int32_t rank, int32_t world, const void *id, int64_t id_bytes) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_comm_init: null handle");
    if (world < 1 || rank < 0 || rank >= world) return dd_fail(h, DD_ERR_ARG, "dd_comm_init: bad rank / world size");
    if (h->nccl_comm) return dd_fail(h, DD_ERR_ARG, "dd_comm_init: the handle already has a communicator");
    if (world == 1) {
        h->rank = 0; h->world = 1;
        return DD_OK;
    }
    if (!id || id_bytes < (int64_t)sizeof(ncclUniqueId)) return dd_fail(h, DD_ERR_ARG, "dd_comm_init: bad unique id");
    DD_TRY(nccl_ready(h));
    DD_CUDA(h, cudaSetDevice(h->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    DD_NCCL(h, g_nccl.CommInitRank(&comm, world, uid, rank));
    h->nccl_comm = comm;
    h->rank = rank;
    h->world = world;
    return DD_OK;
}

extern "C" int dd_comm_shard_cells(dd_handle *h, int32_t on) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_comm_shard_cells: null handle");
    if (on && h->world > 1 && !h->nccl_comm) return dd_fail(h, DD_ERR_ARG, "dd_comm_shard_cells: call dd_comm_init first");
    h->shard_cells = on != 0;
    h->dense_valid = false;
    h->emb_valid = false;
    if (h->d_indptr) dd_set_block(h);
    return DD_OK;
}

extern "C" int dd_comm_info(const dd_handle *h, int32_t *rank_out, int32_t *world_out, int64_t *block_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_comm_info: null handle");
    if (rank_out) *rank_out = h->rank;
    if (world_out) *world_out = h->world;
    if (block_out) {
        block_out[0] = h->blk_n0; block_out[1] = h->blk_n; block_out[2] = h->blk_m0; block_out[3] = h->blk_m;
    }
    return DD_OK;
}

void dd_comm_destroy(dd_handle *h) {
    if (h->nccl_comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr;
}

// The collectives are booked in the per-kernel timing report (dd_set_kernel_timing) under "nccl_*" -- CUDA events around the
// NCCL call on the handle's stream, i.e. including the wait for the slowest rank -- but not in the launch counter (they are
// not this library's kernels).
int dd_comm_allreduce_f64(dd_handle *h, double *buf, int64_t count) {
    if (!dd_sharded(h) || count <= 0) return DD_OK;
    dd_launch_begin(h);
    DD_NCCL(h, g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    DD_TRY(dd_launch_end(h, "nccl_allreduce"));
    h->launches--;
    return DD_OK;
}

// sum of int32 words: used to merge buffers in which every word is written by exactly one rank and zero elsewhere (exact for
// any bit pattern, float32 included)
int dd_comm_allreduce_i32(dd_handle *h, int32_t *buf, int64_t count) {
    if (!dd_sharded(h) || count <= 0) return DD_OK;
    dd_launch_begin(h);
    DD_NCCL(h, g_nccl.AllReduce(buf, buf, (size_t)count, ncclInt32, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    DD_TRY(dd_launch_end(h, "nccl_allreduce"));
    h->launches--;
    return DD_OK;
}

int dd_comm_bcast(dd_handle *h, void *buf, int64_t bytes, int root) {
    if (!dd_sharded(h) || bytes <= 0) return DD_OK;
    dd_launch_begin(h);
    DD_NCCL(h, g_nccl.Broadcast(buf, buf, (size_t)bytes, ncclUint8, root, (ncclComm_t)h->nccl_comm, h->stream));
    DD_TRY(dd_launch_end(h, "nccl_bcast"));
    h->launches--;
    return DD_OK;
}

int dd_comm_gather_ranges(dd_handle *h, void *base, int64_t row_bytes, int n_ranges, const int64_t *begin,
                          const int64_t *count, const int *owner) {
    if (!dd_sharded(h)) return DD_OK;
    dd_launch_begin(h);
    DD_NCCL(h, g_nccl.GroupStart());
    for (int r = 0; r < n_ranges; r++) {
        if (count[r] <= 0) continue;
        uint8_t *p = static_cast<uint8_t *>(base) + begin[r] * row_bytes;
        ncclResult_t rc = g_nccl.Broadcast(p, p, (size_t)(count[r] * row_bytes), ncclUint8, owner[r],
                                           (ncclComm_t)h->nccl_comm, h->stream);
        if (rc != ncclSuccess) {
            g_nccl.GroupEnd();
            return dd_fail(h, DD_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(rc));
        }
    }
    DD_NCCL(h, g_nccl.GroupEnd());
    DD_TRY(dd_launch_end(h, "nccl_allgather"));
    h->launches--;
    return DD_OK;
}
