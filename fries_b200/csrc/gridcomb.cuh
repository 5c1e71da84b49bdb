// Grid-wide all-reduce (+ exclusive prefix over CTAs) for persistent cooperative kernels: one barrier with payload.
//
// What it replaces.  The engines of compress.cuh reduce across the CTAs as "every CTA stores its partial, cg grid.sync(),
// every CTA re-reads ALL partials and sums them in a fixed order".  Measured on a B200 (round 2, 148 CTAs, in-kernel
// timeline of compress2.cuh): the re-read alone costs ~7 us -- 148 CTAs x 4 words hit the same 74 cache lines at the same
// moment and an L2 slice serves same-line requests one after the other (~23 cycles each) --, cg's barrier has the same shape
// (148 CTAs poll ONE word while the arrivals queue behind the polls), and a first version of this file with a ticket
// counter and __threadfence() between data and flag still took ~6 us per call (an atomic, three fences and five dependent
// L2 round trips).  Three reductions were a third of a 60 us stage.
//
// Protocol ("LL", as csrc/comm.cuh uses across GPUs: data and flag travel in the same 8-byte store, so no fence separates
// payload and flag and none follows the poll).  A double or u64 is sent as two words, 32 data bits | 32-bit epoch tag.
//   1. every CTA stores its partial record into its own 64-byte slot;
//   2. CTA 0 polls the records (thread t polls record t: distinct lines, all polls in flight at once), reduces them with a
//      fixed tree -- warp scan, then the warp totals in order --, computes every CTA's exclusive prefix of the first pair
//      on the way, and stores one 256-byte result line PER CTA;
//   3. every CTA polls its own result line.
// No line is read by more than one CTA, no atomics, L2 traffic O(nb) instead of O(nb^2); latency = two store -> poll hops
// plus one block tree.  `fence` adds the memory ordering of cg::grid_group::sync() (global writes of all threads before
// the call are visible to all threads after it: one __threadfence() by the storing / polling threads on either side of the
// protocol); without it the call only orders the payload -- enough when the CTAs exchange nothing else.
//
// State in global memory (zeroed once at allocation): the epoch persists across launches, so any kernel may use the state
// as long as launches do not overlap and every kernel makes at least one call (the epoch is written back by CTA 0 when the
// kernel ends, after which no CTA of that launch reads it).
#pragma once
#include "common.cuh"

#define GC_MAX_CTAS 1024
#define GC_REC_WORDS 8    // 64-byte partial record: up to 4 values (K doubles, K u64) x 2 LL words
#define GC_RES_WORDS 32   // 256-byte result line: 2K totals + prefix pair + up to 8 extras, x 2 LL words
#define GC_STATE_WORDS (16 + GC_MAX_CTAS * GC_REC_WORDS + GC_MAX_CTAS * GC_RES_WORDS)

struct GridComb {
    unsigned long long *state;  // [GC_STATE_WORDS]; state[0] = epoch of the last completed call
    __device__ __forceinline__ volatile unsigned long long *rec(int cta) const { return state + 16 + (size_t)cta * GC_REC_WORDS; }
    __device__ __forceinline__ volatile unsigned long long *res(int cta) const {
        return state + 16 + (size_t)GC_MAX_CTAS * GC_REC_WORDS + (size_t)cta * GC_RES_WORDS;
    }
};

struct GridCombShared {
    unsigned bc[GC_RES_WORDS];
    double wd[2][33];
    unsigned long long wc[2][33];
};

struct GridCombCursor {
    unsigned epoch;  // tag of the last call (uniform over the grid)
};
__device__ __forceinline__ GridCombCursor grid_comb_begin(const GridComb &g) {
    GridCombCursor c;
    c.epoch = (unsigned)*(volatile unsigned long long *)g.state;
    return c;
}
// once per kernel, after its last grid_comb call
__device__ __forceinline__ void grid_comb_end(const GridComb &g, const GridCombCursor &c) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *(volatile unsigned long long *)g.state = c.epoch;
}

__device__ __forceinline__ void gc_put(volatile unsigned long long *slot, int v, unsigned long long bits, unsigned tag) {
    slot[2 * v] = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
    slot[2 * v + 1] = (bits >> 32) | ((unsigned long long)tag << 32);
}
// spin until both words of value v carry the tag
__device__ __forceinline__ unsigned long long gc_get(volatile unsigned long long *slot, int v, unsigned tag) {
    unsigned long long a = slot[2 * v], b = slot[2 * v + 1];
    while ((unsigned)(a >> 32) != tag) a = slot[2 * v];
    while ((unsigned)(b >> 32) != tag) b = slot[2 * v + 1];
    return (a & 0xffffffffull) | (b << 32);
}

// The protocol in four pieces (grid_comb below composes them; a kernel may run its own code on CTA 0 between gc_reduce
// and gc_publish and send up to GC_MAX_EXTRA more words with the totals -- e.g. a threshold solve whose inputs only CTA 0
// then has to read):
//   gc_post     every CTA     store the CTA's record                       (leading __syncthreads; fence: release)
//   gc_reduce   CTA 0         poll + reduce the records, store every CTA's prefix   (fence: acquire for CTA 0)
//   gc_publish  CTA 0         store totals + extras into every CTA's result line
//   gc_wait     every CTA     poll the CTA's own line                      (fence: acquire)
#define GC_MAX_EXTRA 8
template <int K>
__device__ __forceinline__ void gc_post(const GridComb &g, unsigned tag, const double (&d)[K], const unsigned long long (&c)[K],
                                        bool fence) {
    static_assert(4 * K <= GC_REC_WORDS, "record too small");
    __syncthreads();  // the CTA's partials are final; with `fence`: its global writes precede thread 0's fence
    if (threadIdx.x == 0) {
        if (fence) __threadfence();
        volatile unsigned long long *r = g.rec(blockIdx.x);
#pragma unroll
        for (int k = 0; k < K; k++) {
            gc_put(r, k, (unsigned long long)__double_as_longlong(d[k]), tag);
            gc_put(r, K + k, c[k], tag);
        }
    }
}
// CTA 0 only (all its threads).  td / tc: grid totals, uniform over the CTA.
template <int K>
__device__ __forceinline__ void gc_reduce(const GridComb &g, GridCombShared &sh, unsigned tag, double (&td)[K],
                                          unsigned long long (&tc)[K], bool fence) {
    const int tid = threadIdx.x, nb = gridDim.x, lane = tid & 31, w = tid >> 5, nw = (blockDim.x + 31) >> 5;
    double run_d = 0, tot_d1 = 0;
    unsigned long long run_c = 0, tot_c1 = 0;
    for (int base = 0; base < nb; base += blockDim.x) {  // thread t owns record t; larger grids take several passes
        const int i = base + tid;
        double xd[K];
        unsigned long long xc[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            xd[k] = 0;
            xc[k] = 0;
        }
        if (i < nb) {
            volatile unsigned long long *r = g.rec(i);
#pragma unroll
            for (int k = 0; k < K; k++) {
                xd[k] = __longlong_as_double((long long)gc_get(r, k, tag));
                xc[k] = gc_get(r, K + k, tag);
            }
            if (fence) __threadfence();
        }
        __syncwarp();
        // inclusive warp scan of the first pair (for the prefixes), plain warp sum of the second
        double id = xd[0];
        unsigned long long ic = xc[0];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double a = __shfl_up_sync(0xffffffffu, id, o);
            unsigned long long b = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) {
                id += a;
                ic += b;
            }
        }
        const double up_d = __shfl_up_sync(0xffffffffu, id, 1);
        const unsigned long long up_c = __shfl_up_sync(0xffffffffu, ic, 1);
        double sd1 = 0;
        unsigned long long sc1 = 0;
        if (K > 1) {
            sd1 = xd[K - 1];
            sc1 = xc[K - 1];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sd1 += __shfl_xor_sync(0xffffffffu, sd1, o);
                sc1 += __shfl_xor_sync(0xffffffffu, sc1, o);
            }
        }
        __syncthreads();  // the previous pass has read sh.wd / sh.wc
        if (lane == 31) {
            sh.wd[0][w] = id;
            sh.wc[0][w] = ic;
            sh.wd[1][w] = sd1;
            sh.wc[1][w] = sc1;
        }
        __syncthreads();
        double pre = run_d, tot0 = run_d;
        unsigned long long prc = run_c, totc0 = run_c;
        for (int q = 0; q < nw; q++) {
            const double a = sh.wd[0][q];
            const unsigned long long b = sh.wc[0][q];
            if (q < w) {
                pre += a;
                prc += b;
            }
            tot0 += a;
            totc0 += b;
            tot_d1 += sh.wd[1][q];
            tot_c1 += sh.wc[1][q];
        }
        run_d = tot0;
        run_c = totc0;
        if (i < nb) {  // the prefix goes out now, the totals with gc_publish (they carry the same tag)
            volatile unsigned long long *r = g.res(i);
            gc_put(r, 2 * K, (unsigned long long)__double_as_longlong(pre + (lane ? up_d : 0.0)), tag);
            gc_put(r, 2 * K + 1, prc + (lane ? up_c : 0ull), tag);
        }
    }
    td[0] = run_d;
    tc[0] = run_c;
    if (K > 1) {
        td[K - 1] = tot_d1;
        tc[K - 1] = tot_c1;
    }
    __syncthreads();
}
template <int K, int E>
__device__ __forceinline__ void gc_publish(const GridComb &g, unsigned tag, const double (&td)[K], const unsigned long long (&tc)[K],
                                           const unsigned long long *extra) {
    static_assert(E <= GC_MAX_EXTRA && 2 * (2 * K + 2 + E) <= GC_RES_WORDS, "result line too small");
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        volatile unsigned long long *r = g.res(i);
#pragma unroll
        for (int k = 0; k < K; k++) {
            gc_put(r, k, (unsigned long long)__double_as_longlong(td[k]), tag);
            gc_put(r, K + k, tc[k], tag);
        }
#pragma unroll
        for (int e = 0; e < E; e++) gc_put(r, 2 * K + 2 + e, extra[e], tag);
    }
}
template <int K, int E>
__device__ __forceinline__ void gc_wait(const GridComb &g, GridCombShared &sh, unsigned tag, double (&d)[K],
                                        unsigned long long (&c)[K], double &pre_d, unsigned long long &pre_c,
                                        unsigned long long *extra, bool fence) {
    const int tid = threadIdx.x;
    if (tid < 2 * (2 * K + 2 + E)) {  // one word per thread of the CTA's own result line
        volatile unsigned long long *r = g.res(blockIdx.x);
        unsigned long long a = r[tid];
        while ((unsigned)(a >> 32) != tag) a = r[tid];
        sh.bc[tid] = (unsigned)a;
        if (fence) __threadfence();
    }
    __syncthreads();
    auto val = [&](int v) { return (unsigned long long)sh.bc[2 * v] | ((unsigned long long)sh.bc[2 * v + 1] << 32); };
#pragma unroll
    for (int k = 0; k < K; k++) {
        d[k] = __longlong_as_double((long long)val(k));
        c[k] = val(K + k);
    }
    pre_d = __longlong_as_double((long long)val(2 * K));
    pre_c = val(2 * K + 1);
#pragma unroll
    for (int e = 0; e < E; e++) extra[e] = val(2 * K + 2 + e);
    __syncthreads();  // sh.bc may be rewritten by the next call
}

// d, c: this CTA's partials (uniform over the CTA; K <= 2 pairs).  On return: grid totals in d, c (bit-identical in every
// CTA) and, with `prefix`, the exclusive prefix of (d[0], c[0]) over the CTAs before this one.  Every thread of every CTA
// calls it; 32 <= blockDim.x, a multiple of 32.
// CTA 0's part of a plain call, out of line: ~300 instructions that only one CTA of the grid executes, and a kernel makes
// five to ten calls (the instruction cache of the other CTAs' SMs should not have to step over them)
template <int K>
__device__ __noinline__ void gc_cta0(const GridComb g, GridCombShared *sh, unsigned tag, bool fence) {
    double td[K];
    unsigned long long tc[K];
    gc_reduce<K>(g, *sh, tag, td, tc, fence);
    gc_publish<K, 0>(g, tag, td, tc, nullptr);
}
template <int K>
__device__ __forceinline__ void grid_comb(const GridComb &g, GridCombShared &sh, GridCombCursor &cur, double (&d)[K],
                                          unsigned long long (&c)[K], bool prefix, bool fence, double &pre_d,
                                          unsigned long long &pre_c) {
    const unsigned tag = cur.epoch + 1 ? cur.epoch + 1 : 1;  // tags start at 1 (the state starts zeroed) and skip 0 on wrap
    gc_post<K>(g, tag, d, c, fence);
    if (blockIdx.x == 0) gc_cta0<K>(g, &sh, tag, fence);
    gc_wait<K, 0>(g, sh, tag, d, c, pre_d, pre_c, nullptr, fence);
    if (!prefix) {
        pre_d = 0;
        pre_c = 0;
    }
    cur.epoch = tag;
}
__device__ __forceinline__ unsigned grid_comb_next_tag(const GridCombCursor &cur) { return cur.epoch + 1 ? cur.epoch + 1 : 1; }
