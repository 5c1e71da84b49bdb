// fries_comm: peer-mapped inboxes for the in-kernel cross-rank reductions (comm.cuh).
#include "comm.cuh"

#define FR_COMM_BYTES 8192
// layout of one rank's allocation: [0, 4096) inbox doubles, [4096, 6144) flags, [6144] epoch, [6152] error
#define FR_COMM_OFF_FLAGS 4096
#define FR_COMM_OFF_EPOCH 6144
#define FR_COMM_OFF_ERROR 6152

static void bind_view(fries_comm *cm) {
    CommView &v = cm->view;
    v.n_ranks = cm->n_ranks;
    v.rank = cm->rank;
    for (int p = 0; p < FR_MAX_RANKS; p++) {
        char *base = (char *)(p < cm->n_ranks ? cm->peer[p] : nullptr);
        v.inbox[p] = (double *)base;
        v.flags[p] = (unsigned long long *)(base ? base + FR_COMM_OFF_FLAGS : nullptr);
    }
    v.epoch = (unsigned long long *)((char *)cm->local + FR_COMM_OFF_EPOCH);
    v.error = (unsigned long long *)((char *)cm->local + FR_COMM_OFF_ERROR);
}

CommView fries_comm_view(const fries_comm *cm) {
    if (cm) return cm->view;
    CommView v;
    memset(&v, 0, sizeof(v));
    v.n_ranks = 1;
    return v;
}

extern "C" int fries_comm_create(fries_ctx *c, int n_ranks, int rank, fries_comm **out, void *h_ipc_handle64) {
    FRIES_REQUIRE(c && out && h_ipc_handle64, "fries_comm_create: NULL argument");
    FRIES_REQUIRE(n_ranks >= 1 && n_ranks <= FR_MAX_RANKS && rank >= 0 && rank < n_ranks,
                  "fries_comm_create: bad rank %d of %d (max %d ranks)", rank, n_ranks, FR_MAX_RANKS);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(c->device));
    fries_comm *cm = new fries_comm();
    cm->ctx = c;
    cm->n_ranks = n_ranks;
    cm->rank = rank;
    CUDA_TRY(cudaMalloc(&cm->local, FR_COMM_BYTES));
    CUDA_TRY(cudaMemset(cm->local, 0, FR_COMM_BYTES));
    cm->peer[rank] = cm->local;
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, cm->local));
    memcpy(h_ipc_handle64, &h, 64);
    bind_view(cm);
    *out = cm;
    return FRIES_OK;
}

// h_all_handles: n_ranks x 64 bytes, gathered by the host (torch.distributed all_gather); call on every rank,
// then synchronise the ranks once (barrier) before the first kernel that communicates.
extern "C" int fries_comm_connect(fries_comm *cm, const void *h_all_handles) {
    FRIES_REQUIRE(cm && h_all_handles, "fries_comm_connect: NULL argument");
    CUDA_TRY(cudaSetDevice(cm->ctx->device));
    for (int p = 0; p < cm->n_ranks; p++) {
        if (p == cm->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)h_all_handles + 64 * p, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(&cm->peer[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    bind_view(cm);
    return FRIES_OK;
}

extern "C" int fries_comm_destroy(fries_comm *cm) {
    if (!cm) return FRIES_OK;
    cudaSetDevice(cm->ctx->device);
    for (int p = 0; p < cm->n_ranks; p++)
        if (p != cm->rank && cm->peer[p]) cudaIpcCloseMemHandle(cm->peer[p]);
    if (cm->local) cudaFree(cm->local);
    delete cm;
    return FRIES_OK;
}

// nonzero when an in-kernel poll timed out (a peer did not arrive)
extern "C" int fries_comm_error(fries_comm *cm, uint64_t *epoch_of_failure) {
    FRIES_REQUIRE(cm && epoch_of_failure, "fries_comm_error: NULL argument");
    CUDA_TRY(cudaSetDevice(cm->ctx->device));
    unsigned long long e = 0;
    CUDA_TRY(cudaMemcpyAsync(&e, (char *)cm->local + FR_COMM_OFF_ERROR, 8, cudaMemcpyDeviceToHost, cm->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(cm->ctx->stream));
    *epoch_of_failure = e;
    return FRIES_OK;
}

// the resident compression entry points (fries_find_preserve_dev, fries_sys_comp_dev) of this context become
// collective over the ranks of `comm` (every rank must call them in the same order); NULL detaches
extern "C" int fries_ctx_set_comm(fries_ctx *c, fries_comm *comm) {
    FRIES_REQUIRE(c, "fries_ctx_set_comm: NULL context");
    c->comm = comm;
    return FRIES_OK;
}
