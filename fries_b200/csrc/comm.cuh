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
// epoch e, so a slot is never overwritten while it is being read.
#pragma once
#include "common.cuh"

#define FR_MAX_RANKS 8
#define FR_COMM_PAYLOAD 12  // doubles per slot

struct CommView {
    int n_ranks, rank;
    double *inbox[FR_MAX_RANKS];              // inbox[p]: [2 parities][n_ranks][FR_COMM_PAYLOAD] doubles on rank p
    unsigned long long *flags[FR_MAX_RANKS];  // flags[p]: [2][n_ranks] epochs on rank p
    unsigned long long *epoch;                // local: next epoch to use (persists across kernels)
    unsigned long long *error;                // local: set when a poll times out
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

// All-gather of (d0, d1, c) across ranks.  Must be called by every thread of every CTA with grid-uniform
// arguments.  out_* (shared memory, >= FR_MAX_RANKS entries) receive the per-rank values in rank order.
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
    const unsigned long long e = cur.e + 1;  // epochs start at 1; flags start at 0
    const int par = (int)(e & 1);
    if (blockIdx.x == 0 && threadIdx.x < cm.n_ranks) {
        int p = threadIdx.x;
        double *slot = cm.inbox[p] + ((size_t)par * cm.n_ranks + cm.rank) * FR_COMM_PAYLOAD;
        slot[0] = d0;
        slot[1] = d1;
        slot[2] = __longlong_as_double((long long)c);
        __threadfence_system();
        *((volatile unsigned long long *)(cm.flags[p] + (size_t)par * cm.n_ranks + cm.rank)) = e;
    }
    if (threadIdx.x < cm.n_ranks) {
        int q = threadIdx.x;
        volatile unsigned long long *f = cm.flags[cm.rank] + (size_t)par * cm.n_ranks + q;
        long long t0 = clock64();
        while (*f < e) {
            if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer died; do not hang the GPU
                *cm.error = e;
                break;
            }
        }
        __threadfence_system();
        const volatile double *slot = cm.inbox[cm.rank] + ((size_t)par * cm.n_ranks + q) * FR_COMM_PAYLOAD;
        out_d0[q] = slot[0];
        out_d1[q] = slot[1];
        out_c[q] = (unsigned long long)__double_as_longlong(slot[2]);
    }
    __syncthreads();
    cur.e = e;
}

// All-gather of n <= FR_COMM_PAYLOAD doubles per rank.  out is shared memory [n][FR_MAX_RANKS].
__device__ __forceinline__ void comm_allgather_v(const CommView &cm, CommCursor &cur, const double *vals, int n,
                                                 double (*out)[FR_MAX_RANKS]) {
    if (cm.n_ranks <= 1) {
        if (threadIdx.x == 0)
            for (int k = 0; k < n; k++) out[k][0] = vals[k];
        __syncthreads();
        return;
    }
    const unsigned long long e = cur.e + 1;
    const int par = (int)(e & 1);
    if (blockIdx.x == 0 && threadIdx.x < cm.n_ranks) {
        int p = threadIdx.x;
        double *slot = cm.inbox[p] + ((size_t)par * cm.n_ranks + cm.rank) * FR_COMM_PAYLOAD;
        for (int k = 0; k < n; k++) slot[k] = vals[k];
        __threadfence_system();
        *((volatile unsigned long long *)(cm.flags[p] + (size_t)par * cm.n_ranks + cm.rank)) = e;
    }
    if (threadIdx.x < cm.n_ranks) {
        int q = threadIdx.x;
        volatile unsigned long long *f = cm.flags[cm.rank] + (size_t)par * cm.n_ranks + q;
        long long t0 = clock64();
        while (*f < e) {
            if (clock64() - t0 > 20000000000ll) {
                *cm.error = e;
                break;
            }
        }
        __threadfence_system();
        const volatile double *slot = cm.inbox[cm.rank] + ((size_t)par * cm.n_ranks + q) * FR_COMM_PAYLOAD;
        for (int k = 0; k < n; k++) out[k][q] = slot[k];
    }
    __syncthreads();
    cur.e = e;
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
