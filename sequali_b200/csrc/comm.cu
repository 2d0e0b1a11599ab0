// comm.cu -- collectives between the ranks of a sharded run (one process per GPU), NCCL over
// NVLink / NVSwitch, enqueued on the context's launch stream: a merge is ordered behind the
// collectors' kernels and in front of the read-out without any host synchronisation in between.
//
// The reference has no counterpart (it is single process); what is merged and why
// the result equals one sequential pass is laid out in sequali_b200/sharded.py.  NCCL is loaded
// with dlopen on first use, so a single-GPU run does not depend on it.
#include <dlfcn.h>

#include "modules.cuh"

// the few NCCL declarations used here (nccl.h 2.x; the ABI of these calls is stable across 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
enum { NC_SUM = 0, NC_MAX = 2, NC_MIN = 3 };                    // ncclRedOp_t
enum { NC_UINT8 = 1, NC_INT32 = 2, NC_UINT32 = 3, NC_UINT64 = 5 };  // ncclDataType_t

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return SQ_OK;
    const char *names[] = {getenv("SQ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        sq_set_error("cannot load libnccl.so.2 (%s): multi-GPU runs need NCCL", dlerror());
        return SQ_E_ARG;
    }
#define NC_SYM(field, name)                                             \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                       \
    if (!g_nccl.field) {                                                \
        sq_set_error("libnccl lacks %s", name);                         \
        return SQ_E_ARG;                                                \
    }
    NC_SYM(GetUniqueId, "ncclGetUniqueId");
    NC_SYM(CommInitRank, "ncclCommInitRank");
    NC_SYM(CommDestroy, "ncclCommDestroy");
    NC_SYM(AllReduce, "ncclAllReduce");
    NC_SYM(Broadcast, "ncclBroadcast");
    NC_SYM(AllGather, "ncclAllGather");
    NC_SYM(Send, "ncclSend");
    NC_SYM(Recv, "ncclRecv");
    NC_SYM(GroupStart, "ncclGroupStart");
    NC_SYM(GroupEnd, "ncclGroupEnd");
    NC_SYM(GetErrorString, "ncclGetErrorString");
#undef NC_SYM
    g_nccl.lib = lib;
    return SQ_OK;
}

#define NCCL_TRY(expr)                                                                     \
    do {                                                                                   \
        ncclResult_t _r = (expr);                                                          \
        if (_r != ncclSuccess) {                                                           \
            sq_set_error("NCCL error %d (%s) in %s", (int)_r, g_nccl.GetErrorString(_r), #expr); \
            return SQ_E_CUDA;                                                              \
        }                                                                                  \
    } while (0)

struct sq_comm {
    sq_ctx *ctx = nullptr;
    ncclComm_t nc = nullptr;
    int rank = 0, world = 1;
    uint8_t *d_stage = nullptr;  // device staging for host payloads (grow-only)
    size_t d_stage_cap = 0;
    uint8_t *h_stage = nullptr;  // pinned
    size_t h_stage_cap = 0;
};

static int comm_stage(sq_comm *c, size_t nbytes) {
    if (nbytes <= c->d_stage_cap) return SQ_OK;
    CUDA_TRY(cudaStreamSynchronize(c->ctx->stream));
    if (c->d_stage) CUDA_TRY(cudaFree(c->d_stage));
    if (c->h_stage) CUDA_TRY(cudaFreeHost(c->h_stage));
    c->d_stage = c->h_stage = nullptr;
    c->d_stage_cap = c->h_stage_cap = 0;
    const size_t cap = nbytes + nbytes / 4 + 4096;
    CUDA_TRY(cudaMalloc(&c->d_stage, cap));
    CUDA_TRY(cudaMallocHost(&c->h_stage, cap));
    c->d_stage_cap = c->h_stage_cap = cap;
    return SQ_OK;
}

extern "C" int sq_comm_unique_id(uint8_t *out128) {
    SQ_TRY(nccl_load());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return SQ_OK;
}

extern "C" int sq_comm_create(sq_ctx *ctx, const uint8_t *id128, int rank, int world, sq_comm **out) {
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) {
        sq_set_error("rank %d of %d", rank, world);
        return SQ_E_ARG;
    }
    SQ_TRY(nccl_load());
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_comm *c = new sq_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->nc, world, id, rank);
    if (r != ncclSuccess) {
        sq_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return SQ_E_CUDA;
    }
    *out = c;
    return SQ_OK;
}

extern "C" void sq_comm_destroy(sq_comm *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->nc) g_nccl.CommDestroy(c->nc);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    delete c;
}

extern "C" int sq_comm_rank(const sq_comm *c) { return c->rank; }
extern "C" int sq_comm_world(const sq_comm *c) { return c->world; }

static int red_op(int op) { return op == 1 ? NC_MAX : op == 2 ? NC_MIN : NC_SUM; }

// ---- device buffers, on the launch stream ----------------------------------------------------------
extern "C" int sq_comm_allreduce_u64(sq_comm *c, uint64_t *dev, uint64_t n, int op) {
    if (n == 0) return SQ_OK;
    NCCL_TRY(g_nccl.AllReduce(dev, dev, n, NC_UINT64, red_op(op), c->nc, c->ctx->stream));
    return SQ_OK;
}
extern "C" int sq_comm_allreduce_u32(sq_comm *c, uint32_t *dev, uint64_t n, int op) {
    if (n == 0) return SQ_OK;
    NCCL_TRY(g_nccl.AllReduce(dev, dev, n, NC_UINT32, red_op(op), c->nc, c->ctx->stream));
    return SQ_OK;
}
extern "C" int sq_comm_bcast(sq_comm *c, void *dev, uint64_t nbytes, int root) {
    if (nbytes == 0) return SQ_OK;
    NCCL_TRY(g_nccl.Broadcast(dev, dev, nbytes, NC_UINT8, root, c->nc, c->ctx->stream));
    return SQ_OK;
}
extern "C" int sq_comm_send(sq_comm *c, const void *dev, uint64_t nbytes, int dst) {
    if (nbytes == 0) return SQ_OK;
    NCCL_TRY(g_nccl.Send(dev, nbytes, NC_UINT8, dst, c->nc, c->ctx->stream));
    return SQ_OK;
}
extern "C" int sq_comm_recv(sq_comm *c, void *dev, uint64_t nbytes, int src) {
    if (nbytes == 0) return SQ_OK;
    NCCL_TRY(g_nccl.Recv(dev, nbytes, NC_UINT8, src, c->nc, c->ctx->stream));
    return SQ_OK;
}
// several sends / receives as ONE NCCL operation (no serialisation between the pairs)
extern "C" int sq_comm_group_start(sq_comm *c) {
    (void)c;
    NCCL_TRY(g_nccl.GroupStart());
    return SQ_OK;
}
extern "C" int sq_comm_group_end(sq_comm *c) {
    (void)c;
    NCCL_TRY(g_nccl.GroupEnd());
    return SQ_OK;
}

// ---- small host payloads (counters, sizes, tile tables): staged through pinned memory ------------------
extern "C" int sq_comm_allreduce_host_u64(sq_comm *c, uint64_t *host, uint64_t n, int op) {
    if (n == 0) return SQ_OK;
    sq_ctx *ctx = c->ctx;
    SQ_TRY(comm_stage(c, n * 8));
    memcpy(c->h_stage, host, n * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_stage, c->h_stage, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(g_nccl.AllReduce(c->d_stage, c->d_stage, n, NC_UINT64, red_op(op), c->nc, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(c->h_stage, c->d_stage, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(host, c->h_stage, n * 8);
    return SQ_OK;
}
extern "C" int sq_comm_bcast_host(sq_comm *c, void *host, uint64_t nbytes, int root) {
    if (nbytes == 0) return SQ_OK;
    sq_ctx *ctx = c->ctx;
    SQ_TRY(comm_stage(c, nbytes));
    if (c->rank == root) {
        memcpy(c->h_stage, host, nbytes);
        CUDA_TRY(cudaMemcpyAsync(c->d_stage, c->h_stage, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    NCCL_TRY(g_nccl.Broadcast(c->d_stage, c->d_stage, nbytes, NC_UINT8, root, c->nc, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(c->h_stage, c->d_stage, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (c->rank != root) memcpy(host, c->h_stage, nbytes);
    return SQ_OK;
}
// every rank contributes `nbytes` (the same on all ranks); out = world * nbytes in rank order
extern "C" int sq_comm_allgather_host(sq_comm *c, const void *host_in, void *host_out, uint64_t nbytes) {
    if (nbytes == 0) return SQ_OK;
    sq_ctx *ctx = c->ctx;
    const size_t total = (size_t)nbytes * c->world;
    SQ_TRY(comm_stage(c, total));
    memcpy(c->h_stage + (size_t)c->rank * nbytes, host_in, nbytes);
    CUDA_TRY(cudaMemcpyAsync(c->d_stage + (size_t)c->rank * nbytes, c->h_stage + (size_t)c->rank * nbytes, nbytes,
                             cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(g_nccl.AllGather(c->d_stage + (size_t)c->rank * nbytes, c->d_stage, nbytes, NC_UINT8, c->nc, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(c->h_stage, c->d_stage, total, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(host_out, c->h_stage, total);
    return SQ_OK;
}
extern "C" int sq_comm_barrier(sq_comm *c) {
    uint64_t one = 1;
    return sq_comm_allreduce_host_u64(c, &one, 1, 0);
}

// ---- collectors whose tables are sums: all-reduce in place on the device --------------
// QCMetrics: every rank grows its per-position tables to the longest read of any rank, then
// ncclAllReduce(sum, u64) on each table; afterwards every rank's collector answers the usual getters
// with the merged tables.
extern "C" int sq_qc_allreduce(sq_qc *m, sq_comm *c) {
    sq_ctx *ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t v[2] = {m->max_len, 0};
    SQ_TRY(sq_comm_allreduce_host_u64(c, v, 1, 1));
    const uint64_t max_len = v[0];
    SQ_TRY(qc_grow(m, max_len));
    uint64_t n_reads = m->n_reads;
    SQ_TRY(sq_comm_allreduce_host_u64(c, &n_reads, 1, 0));
    NCCL_TRY(g_nccl.GroupStart());
    if (max_len) {
        NCCL_TRY(g_nccl.AllReduce(m->base, m->base, max_len * 5, NC_UINT64, NC_SUM, c->nc, ctx->stream));
        NCCL_TRY(g_nccl.AllReduce(m->phred, m->phred, max_len * 12, NC_UINT64, NC_SUM, c->nc, ctx->stream));
    }
    if (m->ea_len) {
        NCCL_TRY(g_nccl.AllReduce(m->ea_base, m->ea_base, m->ea_len * 5, NC_UINT64, NC_SUM, c->nc, ctx->stream));
        NCCL_TRY(g_nccl.AllReduce(m->ea_phred, m->ea_phred, m->ea_len * 12, NC_UINT64, NC_SUM, c->nc, ctx->stream));
    }
    NCCL_TRY(g_nccl.AllReduce(m->gc, m->gc, 101, NC_UINT64, NC_SUM, c->nc, ctx->stream));
    NCCL_TRY(g_nccl.AllReduce(m->mean_phred, m->mean_phred, 94, NC_UINT64, NC_SUM, c->nc, ctx->stream));
    // the first invalid phred byte of the whole stream: smallest (global record << 8 | byte)
    NCCL_TRY(g_nccl.AllReduce(m->err_key, m->err_key, 1, NC_UINT64, NC_MIN, c->nc, ctx->stream));
    NCCL_TRY(g_nccl.GroupEnd());
    m->max_len = max_len;
    m->n_reads = n_reads;
    return SQ_OK;
}

// AdapterCounter: the same for the per-adapter forward / reverse position counts
extern "C" int sq_adapters_allreduce(sq_adapters *a, sq_comm *c) {
    sq_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t v[1] = {a->max_len};
    SQ_TRY(sq_comm_allreduce_host_u64(c, v, 1, 1));
    SQ_TRY(adapters_grow(a, v[0]));
    uint64_t n_seqs = a->n_seqs;
    SQ_TRY(sq_comm_allreduce_host_u64(c, &n_seqs, 1, 0));
    if (v[0]) {
        // the capacity (row stride) may differ between ranks: one reduction per (adapter, direction) row
        NCCL_TRY(g_nccl.GroupStart());
        for (uint64_t r = 0; r < (uint64_t)a->n_adapters * 2; r++) {
            uint64_t *row = a->counts + r * a->cap_len;
            NCCL_TRY(g_nccl.AllReduce(row, row, v[0], NC_UINT64, NC_SUM, c->nc, ctx->stream));
        }
        NCCL_TRY(g_nccl.GroupEnd());
    }
    a->max_len = v[0];
    a->n_seqs = n_seqs;
    return SQ_OK;
}
