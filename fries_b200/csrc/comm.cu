// fries_comm: peer-mapped inboxes for the in-kernel cross-rank reductions (comm.cuh).
#include "comm.cuh"

#define FR_COMM_HEAD 8192
// layout of one rank's allocation: [0, 4096) inbox words, [4096, 6144) unused, [6144] epoch, [6152] error,
// [8192, ...) candidate windows: x[n_ranks][FR_COMM_XCAP] doubles, then mult[n_ranks][FR_COMM_XCAP] u32
#define FR_COMM_BYTES(n_ranks) (FR_COMM_HEAD + (size_t)(n_ranks) * FR_COMM_XCAP * 12)
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
        v.cand_x[p] = (double *)(base ? base + FR_COMM_HEAD : nullptr);
        v.cand_m[p] = (uint32_t *)(base ? base + FR_COMM_HEAD + (size_t)cm->n_ranks * FR_COMM_XCAP * 8 : nullptr);
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
    CUDA_TRY(cudaMalloc(&cm->local, FR_COMM_BYTES(n_ranks)));
    CUDA_TRY(cudaMemset(cm->local, 0, FR_COMM_BYTES(n_ranks)));
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

// ---- spawn route window --------------------------------------------------------------------------------------------
static size_t route_bytes(int n_ranks, size_t seg_cap) { return (size_t)n_ranks * 2 * seg_cap * 8 + 2 * FR_MAX_RANKS * 8; }
static void bind_route(fries_comm *cm) {
    RouteView &r = cm->route;
    r.n_ranks = cm->n_ranks;
    r.rank = cm->rank;
    r.seg_cap = cm->seg_cap;
    size_t data = (size_t)cm->n_ranks * 2 * cm->seg_cap * 8;
    for (int p = 0; p < FR_MAX_RANKS; p++) {
        char *base = (char *)(p < cm->n_ranks ? cm->win_peer[p] : nullptr);
        r.win[p] = (uint64_t *)base;
        r.counts[p] = (unsigned long long *)(base ? base + data : nullptr);
        r.flags[p] = (unsigned long long *)(base ? base + data + FR_MAX_RANKS * 8 : nullptr);
    }
}

extern "C" int fries_comm_route_create(fries_comm *cm, size_t seg_cap, void *h_ipc_handle64) {
    FRIES_REQUIRE(cm && h_ipc_handle64 && seg_cap > 0, "fries_comm_route_create: bad argument");
    FRIES_REQUIRE(cm->win_local == nullptr, "fries_comm_route_create: the window exists already");
    CUDA_TRY(cudaSetDevice(cm->ctx->device));
    cm->seg_cap = seg_cap;
    size_t bytes = route_bytes(cm->n_ranks, seg_cap);
    CUDA_TRY(cudaMalloc(&cm->win_local, bytes));
    CUDA_TRY(cudaMemset(cm->win_local, 0, bytes));
    cm->win_peer[cm->rank] = cm->win_local;
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, cm->win_local));
    memcpy(h_ipc_handle64, &h, 64);
    bind_route(cm);
    return FRIES_OK;
}

extern "C" int fries_comm_route_connect(fries_comm *cm, const void *h_all_handles) {
    FRIES_REQUIRE(cm && h_all_handles && cm->win_local, "fries_comm_route_connect: create the window first");
    CUDA_TRY(cudaSetDevice(cm->ctx->device));
    for (int p = 0; p < cm->n_ranks; p++) {
        if (p == cm->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)h_all_handles + 64 * p, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(&cm->win_peer[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    bind_route(cm);
    return FRIES_OK;
}

// after the spawn kernel: publish this source's per-destination counts, then the epoch flag, to every destination
__global__ void route_publish_kernel(RouteView r, const unsigned long long *send_counts, unsigned long long epoch) {
    int p = threadIdx.x;
    if (p < r.n_ranks) {
        unsigned long long c = send_counts[p];
        if (c > r.seg_cap) c = r.seg_cap;
        __threadfence_system();  // the spawn kernel's stores (previous kernel on this stream) before the flag
        *((volatile unsigned long long *)(r.counts[p] + r.rank)) = c;
        __threadfence_system();
        *((volatile unsigned long long *)(r.flags[p] + r.rank)) = epoch;
    }
}
// before the merge: wait until every source has published `epoch`; copy the counts next to the merge's other inputs
__global__ void route_wait_kernel(RouteView r, unsigned long long epoch, unsigned long long *recv_counts,
                                  unsigned long long *error) {
    int q = threadIdx.x;
    if (q < r.n_ranks) {
        volatile unsigned long long *f = r.flags[r.rank] + q;
        long long t0 = clock64();
        while (*f < epoch) {
            if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer died; do not hang the GPU
                *error = epoch;
                break;
            }
        }
        __threadfence_system();
        recv_counts[q] = *((volatile unsigned long long *)(r.counts[r.rank] + q));
    }
}
int fries_comm_route_publish(fries_comm *cm, const unsigned long long *d_send_counts) {
    cm->route_epoch++;
    route_publish_kernel<<<1, 32, 0, cm->ctx->stream>>>(cm->route, d_send_counts, cm->route_epoch);
    cm->ctx->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}
int fries_comm_route_wait(fries_comm *cm, unsigned long long *d_recv_counts) {
    route_wait_kernel<<<1, 32, 0, cm->ctx->stream>>>(cm->route, cm->route_epoch, d_recv_counts, cm->view.error);
    cm->ctx->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

extern "C" int fries_comm_destroy(fries_comm *cm) {
    if (!cm) return FRIES_OK;
    cudaSetDevice(cm->ctx->device);
    for (int p = 0; p < cm->n_ranks; p++)
        if (p != cm->rank && cm->win_peer[p]) cudaIpcCloseMemHandle(cm->win_peer[p]);
    if (cm->win_local) cudaFree(cm->win_local);
    for (int p = 0; p < cm->n_ranks; p++)
        if (p != cm->rank && cm->peer[p]) cudaIpcCloseMemHandle(cm->peer[p]);
    if (cm->local) cudaFree(cm->local);
    delete cm;
    return FRIES_OK;
}

// Diagnostics: cost of one in-kernel exchange.  `iters` all-gathers of n doubles, a grid barrier between two of them,
// in a cooperative grid of `ctas` CTAs of 512 threads (0: the compression kernels' launch shape).
__global__ void __launch_bounds__(512) comm_pingpong_kernel(CommView cm, int iters, int n, double *out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_x[FR_COMM_PAYLOAD][FR_MAX_RANKS];
    CommCursor cur = comm_begin(cm);
    double acc = 0;
    for (int it = 0; it < iters; it++) {
        double vals[FR_COMM_PAYLOAD];
        for (int k = 0; k < FR_COMM_PAYLOAD; k++) vals[k] = it + k + cm.rank;
        comm_allgather_v(cm, cur, vals, n, sh_x);
        acc += sh_x[0][0];
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = acc;
    grid.sync();
    comm_end(cm, cur);
}
extern "C" int fries_comm_pingpong(fries_comm *cm, int ctas, int iters, int n, double *us_per_exchange) {
    FRIES_REQUIRE(cm && us_per_exchange && iters > 0 && n >= 1 && n <= FR_COMM_PAYLOAD, "fries_comm_pingpong: bad argument");
    fries_ctx *c = cm->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    if (ctas <= 0) ctas = c->coop_grid((const void *)comm_pingpong_kernel, 512, 0);
    DevBuf<double> out;
    FRIES_TRY(out.alloc(1));
    CommView v = cm->view;
    double *po = out.p;
    void *args[] = {(void *)&v, (void *)&iters, (void *)&n, (void *)&po};
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {  // the first launch absorbs the skew between the ranks
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        CUDA_TRY(cudaLaunchCooperativeKernel((const void *)comm_pingpong_kernel, dim3(ctas), dim3(512), args, 0, c->stream));
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
    }
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *us_per_exchange = 1e3 * ms / iters;
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
