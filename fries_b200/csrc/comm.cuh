// Cross-rank scalar exchange INSIDE persistent kernels, over peer-mapped memory (NVLink / NVSwitch).
//
// The reference synchronises its ranks with MPI_Allgather of one double / one int per global reduction
// (sum_mpi, compress_utils.hpp:179-231; ~50 per iteration).  Here every rank (one process per GPU) owns a
// small inbox in its HBM; the inboxes are mapped into every peer with CUDA IPC.  After the CTAs of a
// cooperative kernel have agreed on their GPU's partial, CTA 0 stores it into slot [my rank] of every
// peer's inbox and publishes it with a system-scope flag; every CTA then polls ITS OWN GPU's inbox and sums the
// partials in rank order -- the same order as sum_mpi, so all ranks obtain bit-identical totals.  No kernel
// boundary, no host round trip, no NCCL launch on the critical path.
//
// Slots are double buffered by epoch parity.  A rank can only publish epoch e+2 after all ranks published
// e+1, and a rank publishes e+1 only after all its CTAs passed the grid.sync that follows their reads of
// epoch e, so a slot is never overwritten while it is being read: TWO EXCHANGES OF ONE RANK MUST BE SEPARATED BY A
// GRID BARRIER (every caller's rounds have one: the reduction that produces the next payload).
#pragma once
#include "common.cuh"

#define FR_MAX_RANKS 8
#define FR_COMM_PAYLOAD 12  // doubles per slot
#define FR_COMM_XCAP 262144 // candidates one rank may contribute to a bracketed threshold solve (compress.cuh)

struct CommView {
    int n_ranks, rank;
    double *inbox[FR_MAX_RANKS];              // inbox[p]: [2 parities][n_ranks][FR_COMM_PAYLOAD] doubles on rank p
    unsigned long long *flags[FR_MAX_RANKS];  // flags[p]: [2][n_ranks] epochs on rank p
    unsigned long long *epoch;                // local: next epoch to use (persists across kernels)
    unsigned long long *error;                // local: set when a poll times out
    // candidate windows of the bracketed threshold solve: cand_x[p] / cand_m[p] = rank p's window,
    // [n_ranks sources][FR_COMM_XCAP]; every rank stores its candidates into ALL windows, so that after one exchange
    // every rank holds the complete list and solves it without further communication
    double *cand_x[FR_MAX_RANKS];
    uint32_t *cand_m[FR_MAX_RANKS];
};

struct CommCursor {
    unsigned long long e;  // current epoch (uniform over the grid and, by construction, over the ranks)
};

__device__ __forceinline__ CommCursor comm_begin(const CommView &cm) {
    CommCursor c;
    c.e = cm.n_ranks > 1 ? *((volatile unsigned long long *)cm.epoch) : 0;
    return c;
}
// call once at the end of the kernel, after a grid.sync that follows the last exchange
__device__ __forceinline__ void comm_end(const CommView &cm, const CommCursor &c) {
    if (cm.n_ranks > 1 && blockIdx.x == 0 && threadIdx.x == 0) *cm.epoch = c.e;
}

// Exchange protocol ("LL": data and flag travel in the same 8-byte store, as in NCCL's low-latency protocol).  A
// double is sent as two words (32 data bits | 32-bit epoch tag); a word is valid when its tag equals the epoch, so an
// exchange costs ONE NVLink traversal: no __threadfence_system between payload and flag on the writer, none between
// flag and payload on the reader.  Layout of rank p's inbox: [2 parities][n_ranks sources][2 * FR_COMM_PAYLOAD] words.
__device__ __forceinline__ void comm_exchange_ll(const CommView &cm, CommCursor &cur, const double *vals, int n,
                                                 double (*out)[FR_MAX_RANKS]) {
    __shared__ unsigned sh_stage[FR_MAX_RANKS][2 * FR_COMM_PAYLOAD];
    const unsigned long long e = cur.e + 1;  // epochs start at 1; the inbox starts zeroed
    const unsigned tag = (unsigned)e;
    const int par = (int)(e & 1), nw = 2 * n, total = cm.n_ranks * nw;
    if (blockIdx.x == 0) {
        for (int t = threadIdx.x; t < total; t += blockDim.x) {
            int p = t / nw, w = t - p * nw;
            unsigned long long bits = (unsigned long long)__double_as_longlong(vals[w >> 1]);
            unsigned data = (w & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
            volatile unsigned long long *slot = (volatile unsigned long long *)cm.inbox[p] +
                                                ((size_t)par * cm.n_ranks + cm.rank) * (2 * FR_COMM_PAYLOAD) + w;
            *slot = (unsigned long long)data | ((unsigned long long)tag << 32);
        }
    }
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        int q = t / nw, w = t - q * nw;
        volatile unsigned long long *slot = (volatile unsigned long long *)cm.inbox[cm.rank] +
                                            ((size_t)par * cm.n_ranks + q) * (2 * FR_COMM_PAYLOAD) + w;
        unsigned long long v = *slot;
        long long t0 = clock64();
        while ((unsigned)(v >> 32) != tag) {
            if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer died; do not hang the GPU
                *cm.error = e;
                break;
            }
            v = *slot;
        }
        sh_stage[q][w] = (unsigned)v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cm.n_ranks * n; t += blockDim.x) {
        int q = t / n, k = t - q * n;
        unsigned long long bits = (unsigned long long)sh_stage[q][2 * k] | ((unsigned long long)sh_stage[q][2 * k + 1] << 32);
        out[k][q] = __longlong_as_double((long long)bits);
    }
    __syncthreads();
    cur.e = e;
}

// All-gather of n <= FR_COMM_PAYLOAD doubles per rank.  Must be called by every thread of every CTA with grid-uniform
// arguments, and two exchanges of one rank must be separated by a grid barrier (see the header comment).
// out is shared memory [n][FR_MAX_RANKS].
__device__ __forceinline__ void comm_allgather_v(const CommView &cm, CommCursor &cur, const double *vals, int n,
                                                 double (*out)[FR_MAX_RANKS]) {
    if (cm.n_ranks <= 1) {
        if (threadIdx.x == 0)
            for (int k = 0; k < n; k++) out[k][0] = vals[k];
        __syncthreads();
        return;
    }
    comm_exchange_ll(cm, cur, vals, n, out);
}

// All-gather of (d0, d1, c) across ranks; out_* (shared memory, >= FR_MAX_RANKS entries) in rank order.
__device__ __forceinline__ void comm_allgather(const CommView &cm, CommCursor &cur, double d0, double d1,
                                               unsigned long long c, double *out_d0, double *out_d1,
                                               unsigned long long *out_c) {
    if (cm.n_ranks <= 1) {
        if (threadIdx.x == 0) {
            out_d0[0] = d0;
            out_d1[0] = d1;
            out_c[0] = c;
        }
        __syncthreads();
        return;
    }
    __shared__ double sh_o[3][FR_MAX_RANKS];
    double vals[3] = {d0, d1, __longlong_as_double((long long)c)};
    comm_exchange_ll(cm, cur, vals, 3, sh_o);
    if (threadIdx.x < cm.n_ranks) {
        out_d0[threadIdx.x] = sh_o[0][threadIdx.x];
        out_d1[threadIdx.x] = sh_o[1][threadIdx.x];
        out_c[threadIdx.x] = (unsigned long long)__double_as_longlong(sh_o[2][threadIdx.x]);
    }
    __syncthreads();
}

// rank-ordered sums (sum_mpi): total over all ranks and the prefix over the ranks before this one
__device__ __forceinline__ void comm_sum(const CommView &cm, const double *v, double &total, double &before) {
    double t = 0, b = 0;
    for (int p = 0; p < cm.n_ranks; p++) {
        if (p == cm.rank) b = t;
        t += v[p];
    }
    total = t;
    before = b;
}
__device__ __forceinline__ unsigned long long comm_sum_u64(const CommView &cm, const unsigned long long *v) {
    unsigned long long t = 0;
    for (int p = 0; p < cm.n_ranks; p++) t += v[p];
    return t;
}

// Receive window of the spawn route (Adder::perform_add vec_utils.hpp:991-1019 without a collective call): every
// rank owns [n_ranks sources][keys[seg_cap] | value bits[seg_cap]] + counts[n_ranks] + flags[n_ranks] in its HBM,
// mapped into every peer.  The spawn kernel of source r stores an element owned by rank p straight into
// win[p] segment r over NVLink (slots from a LOCAL per-destination counter: a (source, destination) pair has its own
// segment, so no remote atomics); a one-warp kernel then publishes the counts and an epoch flag to the peers, and the
// receiving side waits on its own flags before it merges.
struct RouteView {
    uint64_t *win[FR_MAX_RANKS];               // win[p]: rank p's window
    unsigned long long *counts[FR_MAX_RANKS];  // counts[p]: [n_ranks] on rank p (entry r written by source r)
    unsigned long long *flags[FR_MAX_RANKS];   // flags[p]:  [n_ranks] epochs on rank p
    unsigned long long seg_cap;
    int n_ranks, rank;
};

struct fries_comm {
    fries_ctx *ctx = nullptr;
    int n_ranks = 1, rank = 0;
    void *local = nullptr;                 // this rank's inbox + flags + epoch + error (one allocation)
    void *peer[FR_MAX_RANKS] = {nullptr};  // mapped peers (peer[rank] == local)
    CommView view;
    // spawn route window (fries_comm_route_create / _connect); win_local == nullptr: no window
    void *win_local = nullptr;
    void *win_peer[FR_MAX_RANKS] = {nullptr};
    size_t seg_cap = 0;
    unsigned long long route_epoch = 0;    // host-side: one per exchange, identical on all ranks
    RouteView route;
};
CommView fries_comm_view(const fries_comm *cm);  // n_ranks = 1 view when cm == nullptr
