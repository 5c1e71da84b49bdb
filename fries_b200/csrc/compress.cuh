// Device-side stochastic compression engine (a4, a5, a6): persistent cooperative kernels.
//
// Every compression in the reference is "find the exactly-preserved set by a fixed-point iteration
// on a global threshold, then systematically resample the rest with one uniform".  The reference
// runs this per MPI rank with an Allgather per round (compress_utils.cpp:52-92, :153-265); here one
// cooperative kernel (a CTA or two per SM, grid.sync between rounds) plays the role of the ranks:
// each CTA owns a contiguous chunk of the input, per-round partial sums go through a
// double-buffered array that every CTA re-reduces in a fixed order, so the result is deterministic
// and identical in all CTAs.  The resampling step is a chunked device-wide prefix sum of the
// residual weights; an element draws the grid points rn0 + k*unit that fall in its interval.
#pragma once
#include "common.cuh"
#include "comm.cuh"

#define FR_COMP_BLOCK 512

struct CompState {
    double loc_norm;             // residual one-norm after the keep phase (sum of wt_remain)
    double glob_norm;            // one-norm before (find_preserve's *global_norm)
    double new_norm;             // one-norm after resampling (sys_comp's loc_norms[rank] output)
    unsigned n_samp_left;        // budget left after the keep phase
    unsigned rounds;             // fixed-point rounds executed
    unsigned long long n_kept;   // budget consumed by exact preservation
    unsigned long long n_out;    // outputs written (comp_sub) / samples drawn (sys_comp)
    unsigned long long n_in;     // inputs seen
    unsigned long long overflow; // outputs that did not fit out_cap (dropped)
    unsigned long long anomalies;// FP-tie events (two grid points in one sub-element, clamped sub index)
    unsigned long long n_cand;   // candidates collected inside the threshold bracket (bracket_solve)
    unsigned long long fast;     // 1: the bracketed solve decided the preserved set, 0: plain rounds
    unsigned long long ts[8];    // %globaltimer (ns) of CTA 0 at the phase boundaries of comp_sub_engine
    unsigned long long gacc[40]; // grid-wide integer accumulators of the distributed candidate rounds (zeroed by the host)
    unsigned long long rts[16];  // %globaltimer of CTA 0 inside distributed rounds 0-3: start, before / after the grid
                                 // barrier, after the cross-rank exchange
    long long tl[48];            // timeline of thread 0 of CTA 0 through comp_sub_engine2 (clock64, SM cycles): FR_TL marks
};
// (compiled in with -DFRIES_TIMELINE_BUILD only: 37 marks are ~300 instructions of a stage kernel's hot path)
#ifdef FRIES_TIMELINE_BUILD
#define FR_TL(st, k)                                                        \
    do {                                                                    \
        if (blockIdx.x == 0 && threadIdx.x == 0) (st)->tl[k] = clock64();   \
    } while (0)
#else
#define FR_TL(st, k) ((void)0)
#endif

__device__ __forceinline__ unsigned long long fr_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FR_STAMP(st, k)                                                  \
    do {                                                                 \
        if (blockIdx.x == 0 && threadIdx.x == 0) (st)->ts[k] = fr_globaltimer(); \
    } while (0)

// Layout of the partial arrays: [2 parities][nb CTAs][FR_RED_STRIDE]; the stride does not depend on how many values a
// reduction carries, so that consecutive reductions of different widths never overlap (a CTA may already write the
// partials of the next reduction while another one still reads those of the current one).
#define FR_RED_STRIDE 8
#define FR_RED_PART_LEN (2 * 1024 * FR_RED_STRIDE)  // entries per partial array: cooperative grids of <= 1024 CTAs
struct GridRed {
    double *pd;               // [2][nb][FR_RED_STRIDE]
    unsigned long long *pc;   // [2][nb][FR_RED_STRIDE]
    int parity;
    int nb;
    double *shd;              // 33 doubles of shared memory
    unsigned long long *shc;  // 33 u64
};

// K fused (double, u64) block sums with 3 barriers; results in every thread.  shd / shc: K x 33 entries.
template <int K>
__device__ __forceinline__ void block_sum_vec(double (&d)[K], unsigned long long (&c)[K], double *shd,
                                              unsigned long long *shc) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d[k] += __shfl_down_sync(0xffffffffu, d[k], o);
            c[k] += __shfl_down_sync(0xffffffffu, c[k], o);
        }
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            shd[k * 33 + w] = d[k];
            shc[k * 33 + w] = c[k];
        }
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double t = lane < nw ? shd[k * 33 + lane] : 0.0;
            unsigned long long u = lane < nw ? shc[k * 33 + lane] : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t += __shfl_down_sync(0xffffffffu, t, o);
                u += __shfl_down_sync(0xffffffffu, u, o);
            }
            if (lane == 0) {
                shd[k * 33 + 32] = t;
                shc[k * 33 + 32] = u;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        d[k] = shd[k * 33 + 32];
        c[k] = shc[k * 33 + 32];
    }
}

// all-CTA reduction of K (double, u64) pairs with ONE grid barrier; result in every thread of every CTA,
// bit-identical (every CTA re-reduces the per-CTA partials in the same fixed order: one partial per thread, all
// loads in flight at once, then the block tree).  The partial arrays hold 2 x nb x K entries.
template <int K>
__device__ __forceinline__ void grid_reduce_vec(cg::grid_group &grid, GridRed &r, double (&d)[K], unsigned long long (&c)[K],
                                                double *shd, unsigned long long *shc, bool block_reduced = false) {
    if (!block_reduced) block_sum_vec<K>(d, c, shd, shc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            __stcg(&r.pd[((size_t)r.parity * r.nb + blockIdx.x) * FR_RED_STRIDE + k], d[k]);
            __stcg(&r.pc[((size_t)r.parity * r.nb + blockIdx.x) * FR_RED_STRIDE + k], c[k]);
        }
    }
    grid.sync();
#pragma unroll
    for (int k = 0; k < K; k++) {
        d[k] = 0;
        c[k] = 0;
    }
    for (int i = threadIdx.x; i < r.nb; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            d[k] += __ldcg(&r.pd[((size_t)r.parity * r.nb + i) * FR_RED_STRIDE + k]);
            c[k] += __ldcg(&r.pc[((size_t)r.parity * r.nb + i) * FR_RED_STRIDE + k]);
        }
    }
    block_sum_vec<K>(d, c, shd, shc);
    r.parity ^= 1;
}
__device__ __forceinline__ void grid_reduce(cg::grid_group &grid, GridRed &r, double &d, unsigned long long &c) {
    double dd[1] = {d};
    unsigned long long cc[1] = {c};
    grid_reduce_vec<1>(grid, r, dd, cc, r.shd, r.shc);
    d = dd[0];
    c = cc[0];
}
// same, for values that are already reduced over the CTA (uniform in the CTA)
__device__ __forceinline__ void grid_reduce_blk(cg::grid_group &grid, GridRed &r, double &d, unsigned long long &c) {
    double dd[1] = {d};
    unsigned long long cc[1] = {c};
    grid_reduce_vec<1>(grid, r, dd, cc, r.shd, r.shc, true);
    d = dd[0];
    c = cc[0];
}

// Exclusive prefix over CTAs of per-CTA values (d, c) and their totals, identical arithmetic in every CTA.
// sh_scan_d / sh_scan_c: 2 x 33 entries.
__device__ __forceinline__ void grid_excl_scan(cg::grid_group &grid, GridRed &r, double d, unsigned long long c,
                                               double &ex_d, unsigned long long &ex_c, double &tot_d,
                                               unsigned long long &tot_c, double *sh_scan_d,
                                               unsigned long long *sh_scan_c) {
    if (threadIdx.x == 0) {
        __stcg(&r.pd[((size_t)r.parity * r.nb + blockIdx.x) * FR_RED_STRIDE], d);
        __stcg(&r.pc[((size_t)r.parity * r.nb + blockIdx.x) * FR_RED_STRIDE], c);
    }
    grid.sync();
    double dv[2] = {0, 0};
    unsigned long long cv[2] = {0, 0};
    for (int i = threadIdx.x; i < r.nb; i += blockDim.x) {
        double a = __ldcg(&r.pd[((size_t)r.parity * r.nb + i) * FR_RED_STRIDE]);
        unsigned long long b = __ldcg(&r.pc[((size_t)r.parity * r.nb + i) * FR_RED_STRIDE]);
        dv[1] += a;
        cv[1] += b;
        if (i < (int)blockIdx.x) {
            dv[0] += a;
            cv[0] += b;
        }
    }
    block_sum_vec<2>(dv, cv, sh_scan_d, sh_scan_c);
    ex_d = dv[0];
    tot_d = dv[1];
    ex_c = cv[0];
    tot_c = cv[1];
    __syncthreads();
    r.parity ^= 1;
}

// ---------------------------------------------------------------------------------------------------
// Bracketed solve of the preservation threshold.
//
// The keep rounds of find_preserve / find_keep_sub are Newton's iteration from above on the threshold
// t -> R(t) / n(t) (R = one-norm of what is not preserved, n = budget left); thresholds only decrease, the
// preserved set is always "every element >= t", and the rounds may start from ANY such set that is inside the
// final one.  Lemma used below: the set H = {x >= t_hi} is inside the final set whenever
// t_hi >= R(H) / n(H) (downward induction over H in sorted order).  Given a bracket [t_lo, t_hi) around the
// expected fixed point (the fixed point of the previous iteration of the same compression, widened by a
// relative half-width h), ONE pass over the data yields the exact (count, sum) of H and the short list of
// candidates inside the bracket; the Newton rounds then run on that list alone, redundantly in every CTA
// (registers + block reductions, no grid barrier), with candidate sums accumulated as exact integers
// (every x >= t_lo is a multiple of ulp(t_lo)), so the result does not depend on the order in which the
// candidates were appended.  The bracket is valid iff H passes the lemma's test and t_lo is below the
// final threshold; otherwise the caller falls back to the plain rounds over the full data.
// ---------------------------------------------------------------------------------------------------
#define FR_CAND_PER_THREAD 8
#define FR_CAND_CAP (FR_COMP_BLOCK * FR_CAND_PER_THREAD)  // candidates one CTA can hold in registers
#define FR_CAND_GCAP (1u << 20)                            // capacity of the list (grid-distributed rounds beyond one CTA)

struct KeepPred {
    double t;       // centre of the next bracket: the last fixed point extrapolated by the last ratio (0: none)
    double h;       // relative half-width of the bracket
    double t_last;  // fixed point of the previous run
    double e;       // relative prediction error, largest of the recent runs (decays by `decay` per run)
    double decay;   // 0: the width follows the last error only
    double factor;  // half-width = factor x that error (0: 6)
};

struct CandList {
    double *x;                  // [FR_CAND_GCAP] candidate magnitudes
    uint32_t *mult;             // [FR_CAND_GCAP] multiplicities (uniform division: n_div pieces of equal size)
    unsigned long long *count;  // appended so far (may exceed the capacity: then the bracket is invalid)
};

// Single rank: the list lives in cl.x / cl.mult.  Several ranks: the slot comes from the same local counter, and the
// candidate is stored into segment `rank` of EVERY rank's window (comm.cuh).  The caller issues ONE system-scope fence
// per thread after its collection loop (cand_flush), then a grid barrier; its first cross-rank exchange after that
// publishes the counts, and every rank solves the merged list on its own.
__device__ __forceinline__ void cand_append(const CandList &cl, const CommView &cm, double x, uint32_t mult) {
    // once the list has overflowed the bracket is invalid anyway: stop hammering the counter
    const unsigned long long cap = cm.n_ranks > 1 ? FR_COMM_XCAP : FR_CAND_GCAP;
    if (*(volatile unsigned long long *)cl.count > cap) return;
    unsigned long long k = atomicAdd(cl.count, 1ull);
    if (k >= cap) return;
    if (cm.n_ranks > 1) {
        for (int p = 0; p < cm.n_ranks; p++) {
            cm.cand_x[p][(size_t)cm.rank * FR_COMM_XCAP + k] = x;
            cm.cand_m[p][(size_t)cm.rank * FR_COMM_XCAP + k] = mult;
        }
    } else {
        cl.x[k] = x;
        cl.mult[k] = mult;
    }
}

__device__ __forceinline__ void cand_flush(const CommView &cm, bool appended) {
    if (cm.n_ranks > 1 && appended) __threadfence_system();
}

struct BracketResult {
    bool valid;
    unsigned long long n_cand;      // length of the candidate list
    double x_cut;                   // preserved <=> x >= x_cut
    double R;                       // one-norm of what is not preserved (by subtraction, as the reference's rounds)
    unsigned nrem;                  // budget left
    unsigned long long kept_cand;   // pieces preserved from the candidate list
    unsigned rounds;
};

#ifdef FR_BRACKET_TIMING
__device__ long long fr_bt[16];
#define FR_BT(k)                                                    \
    do {                                                            \
        if (blockIdx.x == 0 && threadIdx.x == 0 && (k) < 16) fr_bt[k] = clock64(); \
    } while (0)
#else
#define FR_BT(k)
#endif
// shc: >= 32 u64 of shared memory (shd unused, kept for symmetry).  Uniform result in every thread of every CTA.
// All sums are integers (counts; candidate magnitudes in units of ulp(t_lo), as eight 16-bit limbs so that a warp's
// total fits the 32-bit redux instruction), reduced with redux.sync + shared-memory atomics: order-independent,
// and far cheaper than 64-bit shuffle trees (SHFL issues once per cycle per SM).
//
// Lists of up to FR_CAND_CAP candidates are solved redundantly by every CTA (no grid barrier).  Longer lists (large
// vectors: the list grows like the square root of the vector length) are split across the CTAs, FR_CAND_CAP each;
// the per-round totals then go through integer atomics on gacc (CompState) and one grid barrier per round.
//
// Several ranks (cm.n_ranks > 1): R0 / nrem0 are the global values and the list is the concatenation, in rank order, of
// the ranks' segments in this rank's candidate window (cand_append stored them there; `seg_len` = their lengths from the
// caller's first all-gather, which also orders the remote stores).  Every rank solves the identical list with identical
// arithmetic: no communication inside the solve.  `peers_ok`: every rank's list fitted its segment.
__device__ __forceinline__ unsigned long long bracket_list_len(const CandList &cl, unsigned long long *shc) {
    if (threadIdx.x == 0) shc[20] = __ldcg(cl.count);
    __syncthreads();
    unsigned long long ncand = shc[20];
    __syncthreads();
    return ncand;
}
__device__ __forceinline__ BracketResult bracket_solve(cg::grid_group &grid, const CandList &cl,
                                                       unsigned long long *gacc, double R0, long long nrem0,
                                                       double t_lo, double t_hi, double *shd, unsigned long long *shc,
                                                       const CommView &cm, const unsigned long long *seg_len,
                                                       bool peers_ok) {
    (void)shd;
    const bool multi = cm.n_ranks > 1;
    BracketResult res;
    res.valid = false;
    res.x_cut = t_hi;
    res.R = R0;
    res.nrem = 0;
    res.kept_cand = 0;
    res.rounds = 0;
    res.n_cand = 0;
    FR_BT(0);
    unsigned long long ncand;
    unsigned long long seg_end[FR_MAX_RANKS];  // multi: exclusive end of each rank's segment in the concatenation
    if (multi) {
        unsigned long long t = 0;
        for (int p = 0; p < FR_MAX_RANKS; p++) {
            if (p < cm.n_ranks) t += seg_len[p];
            seg_end[p] = t;
        }
        ncand = t;
    } else {
        ncand = bracket_list_len(cl, shc);
    }
    res.n_cand = ncand;
    unsigned long long gcap = (unsigned long long)gridDim.x * FR_CAND_CAP;
    if (gcap > FR_CAND_GCAP) gcap = FR_CAND_GCAP;
    if (ncand > gcap || !peers_ok || nrem0 <= 0 || nrem0 > 0xffffffffll) return res;
    const bool dist = ncand > FR_CAND_CAP;
    const unsigned long long slice0 = dist ? (unsigned long long)blockIdx.x * FR_CAND_CAP : 0ull;
    if (!(t_hi * (double)nrem0 >= R0)) return res;  // H is not certainly preserved
    // ulp(t_lo) = 2^(E_lo - 1075) with E_lo the biased exponent; every candidate is an integer multiple of it
    const int E_lo = (int)((__double_as_longlong(t_lo) >> 52) & 0x7ff);
    if (E_lo < 64 || E_lo > 1900) return res;  // (sub)normal extremes: leave to the plain rounds
    const double ulp_lo = __longlong_as_double((long long)(E_lo - 52) << 52);
    double x[FR_CAND_PER_THREAD];
    uint32_t mu[FR_CAND_PER_THREAD];
    unsigned state = 0;  // bit k: candidate k of this thread not yet preserved
#pragma unroll
    for (int k = 0; k < FR_CAND_PER_THREAD; k++) {
        unsigned long long idx = slice0 + threadIdx.x + k * FR_COMP_BLOCK;
        x[k] = 0;
        mu[k] = 0;
        if (idx < ncand) {
            if (multi) {
                int q = 0;
                while (idx >= seg_end[q]) q++;
                size_t off = (size_t)q * FR_COMM_XCAP + (size_t)(idx - (q ? seg_end[q - 1] : 0ull));
                x[k] = __ldcg(cm.cand_x[cm.rank] + off);
                mu[k] = __ldcg(cm.cand_m[cm.rank] + off);
            } else {
                x[k] = __ldcg(cl.x + idx);
                mu[k] = __ldcg(cl.mult + idx);
            }
            state |= 1u << k;
        }
    }
    // shared accumulators (32-bit: native shared-memory atomics): [buf][0..2] count limbs, [buf][3..10] sum limbs,
    // 16 bits each (a warp's total < 2^21, a CTA's < 2^25).  THREE buffers rotate so that one barrier per round
    // suffices: round r accumulates into buffer r % 3 and clears buffer (r + 1) % 3, which was last read in round
    // r - 2 -- every thread has passed the barrier of round r - 1 after those reads.  (With two buffers a warp that
    // runs ahead could clear the totals of round r - 1 while a slower warp is still reading them.)
    unsigned *acc = reinterpret_cast<unsigned *>(shc);
    unsigned long long *acc_min = shc + 18;
    if (threadIdx.x < 36) acc[threadIdx.x] = 0;
    if (threadIdx.x == 36) *acc_min = 0x7ff0000000000000ull;  // +inf
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned long long cnt_tot = 0;
    unsigned __int128 sum_tot = 0;
    double R = R0;
    unsigned long long nrem = (unsigned long long)nrem0;
    double xmin = INFINITY, dist_xmin = INFINITY;
    FR_BT(1);
    for (unsigned round = 0; round < 4096; round++) {
        if (dist && round < 4 && blockIdx.x == 0 && threadIdx.x == 0) gacc[40 + 4 * round] = fr_globaltimer();
        unsigned *a = acc + 12 * (round % 3);
        unsigned long long c = 0;
        unsigned __int128 s = 0;
        const double fac = (double)nrem;
#pragma unroll
        for (int k = 0; k < FR_CAND_PER_THREAD; k++) {
            if (((state >> k) & 1u) && x[k] * fac >= R) {
                state &= ~(1u << k);
                c += mu[k];
                // x / ulp(t_lo), exact: mantissa shifted by the exponent difference (0 or 1: x < t_hi < 2 t_lo)
                const long long xb = __double_as_longlong(x[k]);
                const unsigned long long ix = ((unsigned long long)(xb & 0xfffffffffffffll) | (1ull << 52))
                                              << ((int)((xb >> 52) & 0x7ff) - E_lo);
                s += (unsigned __int128)ix * mu[k];
                xmin = fmin(xmin, x[k]);
            }
        }
        if (__any_sync(0xffffffffu, c != 0)) {
            const unsigned long long slo = (unsigned long long)s, shi = (unsigned long long)(s >> 64);
#pragma unroll
            for (int q = 0; q < 11; q++) {
                unsigned v = q < 3 ? (unsigned)((c >> (16 * q)) & 0xffffu)
                                   : q < 7 ? (unsigned)((slo >> (16 * (q - 3))) & 0xffffu)
                                           : (unsigned)((shi >> (16 * (q - 7))) & 0xffffu);
                if (__any_sync(0xffffffffu, v != 0)) {
                    unsigned w = __reduce_add_sync(0xffffffffu, v);
                    if (lane == 0) atomicAdd(&a[q], w);
                }
            }
            if (dist) {  // running minimum of the preserved candidates (positive doubles order like their bit patterns)
                double wm = xmin;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wm = fmin(wm, __shfl_xor_sync(0xffffffffu, wm, o));
                if (lane == 0 && wm < INFINITY) atomicMin(acc_min, (unsigned long long)__double_as_longlong(wm));
            }
        }
        // clear the next round's buffer (its last readers passed the previous barrier)
        if (threadIdx.x < 12) acc[12 * ((round + 1) % 3) + threadIdx.x] = 0;
        __syncthreads();
        res.rounds = round + 1;
        unsigned long long lim[11];
        if (!dist) {
#pragma unroll
            for (int q = 0; q < 11; q++) lim[q] = a[q];
        } else {
            // CTA totals -> grid totals.  Three global buffers rotate: buffer (round + 1) % 3 was last read two rounds
            // ago (every CTA has passed the barrier after those reads), so CTA 0 may clear it before this barrier.
            unsigned long long *g = gacc + 12 * (round % 3);
            unsigned long long *rts = gacc + 40;  // CompState::rts follows gacc
            if (round < 4 && blockIdx.x == 0 && threadIdx.x == 0) rts[4 * round + 1] = fr_globaltimer();
            if (round == 0) FR_BT(8);
            if (threadIdx.x < 11 && a[threadIdx.x]) atomicAdd(&g[threadIdx.x], (unsigned long long)a[threadIdx.x]);
            // gacc[36] (zeroed by the host) collects max(inf_bits - bits) = the smallest preserved magnitude so far
            if (threadIdx.x == 11 && *acc_min < 0x7ff0000000000000ull) atomicMax(&gacc[36], 0x7ff0000000000000ull - *acc_min);
            if (blockIdx.x == 0 && threadIdx.x < 12) gacc[12 * ((round + 1) % 3) + threadIdx.x] = 0;
            if (round == 0) FR_BT(9);
            grid.sync();
            if (round == 0) FR_BT(10);
            if (round < 4 && blockIdx.x == 0 && threadIdx.x == 0) rts[4 * round + 2] = fr_globaltimer();
            // one load per CTA and value, then a shared-memory broadcast: 150 000 threads reading the same eleven
            // words serialise in L2 (measured: 11 us)
            unsigned long long *bc = shc + 20;
            if (threadIdx.x < 11) bc[threadIdx.x] = __ldcg(&g[threadIdx.x]);
            if (threadIdx.x == 11) bc[11] = __ldcg(&gacc[36]);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 11; q++) lim[q] = bc[q];
            dist_xmin = __longlong_as_double((long long)(0x7ff0000000000000ull - bc[11]));
            __syncthreads();
            if (round == 0) FR_BT(11);
            if (round < 4 && blockIdx.x == 0 && threadIdx.x == 0) rts[4 * round + 3] = fr_globaltimer();
        }
        const unsigned long long c_round = lim[0] + (lim[1] << 16) + (lim[2] << 32);
        if (c_round == 0) break;
        cnt_tot += c_round;
        unsigned __int128 s_round = 0;
#pragma unroll
        for (int q = 7; q >= 0; q--) s_round = (s_round << 16) + lim[3 + q];
        sum_tot += s_round;
        if (cnt_tot >= (unsigned long long)nrem0) return res;  // budget exhausted inside the bracket
        nrem = (unsigned long long)nrem0 - cnt_tot;
        double kept_sum = (double)(unsigned long long)(sum_tot >> 64) * 18446744073709551616.0 +
                          (double)(unsigned long long)sum_tot;
        R = R0 - kept_sum * ulp_lo;
        FR_BT(2 + round);
    }
    FR_BT(6);
    // smallest preserved candidate: positive doubles order like their bit patterns
    if (!dist) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        if (lane == 0 && xmin < INFINITY) atomicMin(acc_min, (unsigned long long)__double_as_longlong(xmin));
        __syncthreads();
        xmin = __longlong_as_double((long long)*acc_min);
        __syncthreads();
    } else {
        xmin = dist_xmin;  // grid (and rank) minimum, read with the totals of the last round
    }
    FR_BT(7);
    res.x_cut = xmin < t_hi ? xmin : t_hi;
    res.R = R;
    res.nrem = (unsigned)nrem;
    res.kept_cand = cnt_tot;
    res.valid = t_lo * (double)nrem < R;  // nothing below the bracket can be preserved
    return res;
}

// Next bracket from this run's fixed point.  Centre: geometric extrapolation t_fin * (t_fin / t_last) -- the one-norm of
// an FRI iterate drifts smoothly (death/cloning, shift updates), and so do the thresholds.  Half-width: 6 x the error of
// the centre that was used for this run, otherwise steered towards a list of ~2500 candidates (one CTA's registers hold
// 4096: no grid barrier in the rounds).
__device__ __forceinline__ void keep_pred_update(KeepPred *p, double t_used, double h_prev, double t_fin,
                                                 unsigned long long ncand) {
    const double h_min = 1e-4, h_max = 0.05;
    const double t_last = p->t_last;
    double h = 0.01, ratio = 1.0;
    double e_keep = 0;
    if (t_used > 0 && t_fin > 0) {
        double err = fabs(t_fin / t_used - 1.0);
        // the width follows the LARGEST recent error (decay 0.7 per run, fries_hbpp_alloc), not the last one: the threshold of a 2.4e5-element vector moves by
        // ~1e-3 between iterations with occasional larger steps, and a bracket sized by one quiet iteration missed the
        // next one often enough to matter (a miss = 10-14 plain rounds, ~70 us; a wider bracket = more candidates, which
        // cost microseconds).  Measured round 2, Ne-sized run.
        e_keep = fmax(err, p->decay * p->e);
        double steer = h_prev * 2500.0 / (double)(ncand > 0 ? ncand : 1);
        steer = fmin(fmax(steer, 0.5 * h_prev), 1.5 * h_prev);
        h = fmax((p->factor > 0 ? p->factor : 6.0) * e_keep, steer);
    }
    if (t_last > 0 && t_fin > 0) ratio = fmin(fmax(t_fin / t_last, 0.8), 1.25);
    h = fmin(fmax(h, h_min), h_max);
    const bool ok = t_fin > 0 && isfinite(t_fin);
    p->t = ok ? t_fin * ratio : 0.0;
    p->t_last = ok ? t_fin : 0.0;
    p->h = h;
    p->e = e_keep;
}

// seed_sys compress_utils.cpp:107-127 for a rank whose lower ranks hold `lbound` of the `glob` norm
__device__ __forceinline__ double seed_sys_dev(double lbound, double glob, double rn, unsigned n_samp) {
    rn *= glob / n_samp;
    rn += glob / n_samp * (int)(lbound * n_samp / glob);
    if (rn < lbound) rn += glob / n_samp;
    return rn;
}

struct CompSubBufs {
    // per-input state (capacity >= number of inputs)
    double *veff;
    double *wt_remain;
    double *lb;          // exclusive prefix of wt_remain (resampling lower bound)
    double *rinv;        // 1 / (row norm): the row generator's normalisation, computed once in prep
    uint32_t *ndiv;
    uint32_t *keep;      // bit j = sub-element j preserved exactly (bit 0 for uniform rows)
    uint32_t *kcnt;      // outputs this input emits
    uint8_t *nsub;
    // outputs
    double *out_val;
    uint32_t *out_widx;
    uint32_t *out_sub;
    unsigned long long out_cap;
    // reduction scratch
    double *part_d;               // [2][grid]
    unsigned long long *part_c;   // [2][grid]
    CompState *st;
    CommView cm;                  // n_ranks == 1: no cross-rank exchange
    // bracketed threshold solve (optional): prediction carried between runs + candidate list
    KeepPred *pred;
    CandList cand;
};

// The hierarchical compression engine.  Provider P supplies
//   size_t count();                                            number of inputs
//   void prep(size_t i, double &v, uint32_t &ndiv, uint32_t &nsub, double &rinv, double &wmax);  effective weight, row shape,
//                                                              normalisation factor of the row (stored per input)
//   void visit(size_t i, double rinv, F f);                    calls f(j, w_j) for the nsub sub-weights in order --
//                                                              rows are streamed, never materialised
// Restates comp_sub (compress_utils.cpp:797-820): find_keep_sub :130-276 then sys_sub :702-794.
template <class P>
__device__ void comp_sub_engine(P &prov, const CompSubBufs &b, unsigned n_samp_in, double rn_uniform) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[6 * 33];
    __shared__ unsigned long long sh_sc[6 * 33];
    GridRed red{b.part_d, b.part_c, 0, (int)gridDim.x, sh_d, sh_c};

    // scalars every thread needs: loaded once per CTA and broadcast (all-thread loads of one address serialise in L2)
    __shared__ unsigned long long sh_bc[4];
    if (threadIdx.x == 0) {
        sh_bc[0] = (unsigned long long)prov.count();
        sh_bc[1] = b.pred ? (unsigned long long)__double_as_longlong(__ldcg(&b.pred->t)) : 0ull;
        sh_bc[2] = b.pred ? (unsigned long long)__double_as_longlong(__ldcg(&b.pred->h)) : 0ull;
    }
    __syncthreads();
    const size_t n = (size_t)sh_bc[0];
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;

    __shared__ double sh_x0[FR_MAX_RANKS], sh_x1[FR_MAX_RANKS];
    __shared__ double sh_xv[12][FR_MAX_RANKS];
    __shared__ unsigned long long sh_xc[FR_MAX_RANKS];
    const CommView &cm = b.cm;
    CommCursor cur = comm_begin(cm);
    const bool multi = cm.n_ranks > 1;

    FR_STAMP(b.st, 0);
    // bracket around the expected fixed point (see bracket_solve); uniform over the grid
    const double t_pred = __longlong_as_double((long long)sh_bc[1]), h_pred = __longlong_as_double((long long)sh_bc[2]);
    const bool try_fast = t_pred > 0 && h_pred > 0 && h_pred < 0.25;
    const double t_lo = t_pred * (1.0 - h_pred), t_hi = t_pred * (1.0 + h_pred);

    // ---- phase 0: effective weights (find_keep_sub :134-137); with a bracket also the exact (count, sum) of
    // everything at or above it and the list of candidates inside it ----
    double s = 0, s_hi = 0;
    unsigned long long c_hi = 0;
    bool appended = false;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double v, rinv = 1.0, wmax = 1.0;
        uint32_t nd, ns;
        prov.prep(i, v, nd, ns, rinv, wmax);
        b.rinv[i] = rinv;
        b.veff[i] = v;
        b.wt_remain[i] = v;
        b.ndiv[i] = nd;
        b.nsub[i] = (uint8_t)ns;
        b.keep[i] = 0;
        s += v;
        // largest piece of this input (an upper bound for rows): parked in lb[], which is free until the count pass;
        // an input whose largest piece is below the bracket needs no row here and none in the cut pass
        const double xmax = nd > 0 ? v / nd : v * wmax;
        b.lb[i] = xmax;
        if (try_fast && xmax >= t_lo) {
            if (nd > 0) {
                double x = v / nd;
                if (x >= t_hi) {
                    c_hi += nd;
                    s_hi += v;
                } else if (x >= t_lo) {
                    cand_append(b.cand, cm, x, nd);
                    appended = true;
                }
            } else {
                prov.visit(i, rinv, [&](uint32_t j, double wj) {
                    if (j < ns) {
                        double x = v * wj;
                        if (x >= t_hi) {
                            c_hi++;
                            s_hi += x;
                        } else if (x >= t_lo) {
                            cand_append(b.cand, cm, x, 1u);
                            appended = true;
                        }
                    }
                });
            }
        }
    }
    cand_flush(cm, appended);
    unsigned long long dummy = 0;
    {
        double dd[2] = {s, s_hi};
        unsigned long long cc[2] = {c_hi, 0ull};
        grid_reduce_vec<2>(grid, red, dd, cc, sh_sd, sh_sc);
        s = dd[0];
        s_hi = dd[1];
        c_hi = cc[0];
    }
    double loc = s;          // this rank's loc_one_norm
    double R_next = s;       // sum_mpi(loc_one_norm) for the coming round
    __shared__ unsigned long long sh_seg[FR_MAX_RANKS];  // multi: length of every rank's candidate segment
    const unsigned long long my_cand = try_fast ? bracket_list_len(b.cand, sh_sc) : 0ull;
    bool peers_ok = try_fast && my_cand <= (multi ? (unsigned long long)FR_COMM_XCAP : (unsigned long long)FR_CAND_GCAP);
    if (multi) {
        // one exchange: norm, bracket statistics, list length (the remote candidate stores were fenced by their
        // writers before the grid barrier of the reduction above, so they are ordered before this exchange)
        double pay[5] = {s, s_hi, (double)c_hi, peers_ok ? 1.0 : 0.0, (double)my_cand};
        comm_allgather_v(cm, cur, pay, 5, sh_xv);
        double before;
        comm_sum(cm, sh_xv[0], R_next, before);
        s = R_next;          // global one-norm (reported)
        double gs = 0, gc = 0;
        for (int p = 0; p < cm.n_ranks; p++) {
            gs += sh_xv[1][p];
            gc += sh_xv[2][p];
            if (sh_xv[3][p] == 0.0) peers_ok = false;
        }
        if (threadIdx.x < cm.n_ranks) sh_seg[threadIdx.x] = (unsigned long long)sh_xv[4][threadIdx.x];
        s_hi = gs;           // global (count, sum) of everything at or above the bracket
        c_hi = (unsigned long long)gc;
        if (try_fast) __threadfence_system();  // acquire side: the peers' candidates in this rank's window
        __syncthreads();
    }

    FR_STAMP(b.st, 1);  // phase 0 done
    // ---- keep rounds (find_keep_sub :153-265) ----
    unsigned nrem = n_samp_in;
    unsigned long long glob_sampled = 1;
    int last_pass = 0;
    double R = 0;
    unsigned rounds = 0;
    unsigned long long kept_total = 0;
    // the residual norm of the last exact recomputation stays valid while no round preserves anything
    bool fresh = false;
    double fresh_cs = 0, fresh_loc = 0, fresh_G = 0, fresh_lb0 = 0;
    unsigned long long n_cand = 0, fast_kc = 0;
    bool fast_done = false;
    if (try_fast) {
        // Newton rounds on the candidate list only (every CTA, redundantly), then ONE pass that applies the cut
        BracketResult br = bracket_solve(grid, b.cand, b.st->gacc, s - s_hi, (long long)n_samp_in - (long long)c_hi, t_lo, t_hi,
                                         sh_sd, sh_sc, cm, sh_seg, peers_ok);
        n_cand = br.n_cand;
        FR_STAMP(b.st, 6);  // candidate rounds done
        if (br.valid) {
            const double x_cut = br.x_cut;
            double t = 0;
            unsigned long long kc = 0;
            for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                double v = b.veff[i];
                double wr = v;
                if (b.lb[i] >= x_cut) {  // largest piece (phase 0); below the cut nothing of this input is preserved
                    uint32_t nd = b.ndiv[i];
                    if (nd > 0) {
                        if (v / nd >= x_cut) {
                            b.keep[i] = 1;
                            wr = 0;
                            kc += nd;
                        }
                    } else {
                        uint32_t ns = b.nsub[i], kb = 0;
                        double sub_remain = 0;
                        prov.visit(i, b.rinv[i], [&](uint32_t j, double wj) {
                            if (j < ns) {
                                double x = v * wj;
                                if (x >= x_cut) {
                                    kb |= 1u << j;
                                    kc++;
                                } else {
                                    sub_remain += x;
                                }
                            }
                        });
                        b.keep[i] = kb;
                        wr = sub_remain;
                    }
                    b.wt_remain[i] = wr;
                }
                t += wr;
            }
            block_sum_pair(t, kc, sh_d, sh_c);
            fresh_cs = t;      // this CTA's share of the resampling line
            fast_kc = kc;      // pieces this CTA preserved
            FR_STAMP(b.st, 7);  // cut applied (CTA 0)
            // the grid totals (residual norm, preserved pieces) come out of the line scan below: no reduction here
            kept_total = c_hi + br.kept_cand;
            nrem = br.nrem;
            R = br.R;
            rounds = br.rounds;
            fast_done = true;
            glob_sampled = 0;  // skip the plain rounds
            if (blockIdx.x == 0 && threadIdx.x == 0) b.st->fast = 1;
        }
    }
    while (glob_sampled > 0 && rounds < 100000) {  // the bound only guards against a corrupted reduction
        R = R_next;
        if (R < 0) break;
        const double wt_factor = (double)nrem;
        double rem = 0;
        unsigned long long cnt = 0;
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            double wr = b.wt_remain[i];
            if (wr > 0) {
                double v = b.veff[i];
                uint32_t nd = b.ndiv[i];
                double cw = v * wt_factor;
                if (nd > 0) cw /= nd;
                if (cw >= R) {
                    if (nd > 0) {
                        b.keep[i] = 1;
                        b.wt_remain[i] = 0;
                        cnt += nd;
                        rem += v;
                    } else {
                        uint32_t ns = b.nsub[i], kb = b.keep[i];
                        uint32_t full = (ns / 8) * 8;
                        double sub_remain = 0;
                        prov.visit(i, b.rinv[i], [&](uint32_t j, double wj) {
                            if (j < ns && !((kb >> j) & 1u)) {
                                double sm = cw * wj;
                                double eps = j < full ? 1e-12 : 1e-10;
                                if (sm >= R && fabs(sm) > eps) {
                                    kb |= 1u << j;
                                    cnt++;
                                } else {
                                    sub_remain += sm;
                                }
                            }
                        });
                        b.keep[i] = kb;
                        sub_remain /= wt_factor;
                        double change = wr - sub_remain;
                        b.wt_remain[i] = sub_remain;
                        rem += change;
                    }
                }
            }
        }
        grid_reduce(grid, red, rem, cnt);
        loc -= rem;
        R_next = loc;
        if (multi) {  // glob_sampled = sum_mpi(loc_sampled); next round's norm = sum_mpi(loc_one_norm)
            double before;
            comm_allgather(cm, cur, loc, 0.0, cnt, sh_x0, sh_x1, sh_xc);
            comm_sum(cm, sh_x0, R_next, before);
            cnt = comm_sum_u64(cm, sh_xc);
        }
        glob_sampled = cnt;
        nrem -= (unsigned)cnt;
        kept_total += cnt;
        rounds++;
        if (cnt) fresh = false;
        if (last_pass && glob_sampled) last_pass = 0;
        if (glob_sampled == 0 && !last_pass) {
            last_pass = 1;
            glob_sampled = 1;
            double t = 0;
            for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) t += b.wt_remain[i];
            block_sum_pair(t, dummy, sh_d, sh_c);
            fresh_cs = t;  // this CTA's chunk of the residual weights = its share of the resampling line
            grid_reduce_blk(grid, red, t, dummy);
            loc = t;
            R_next = t;
            fresh_G = t;
            fresh_lb0 = 0;
            if (multi) {
                comm_allgather(cm, cur, t, 0.0, 0ull, sh_x0, sh_x1, sh_xc);
                comm_sum(cm, sh_x0, R_next, fresh_lb0);
                fresh_G = R_next;
            }
            fresh = true;
            fresh_loc = t;
        }
    }
    // ---- residual norm (find_keep_sub :267-275) and this rank's place on the resampling line (comp_sub :818,
    // seed_sys :107-127): reuse the last exact recomputation when nothing was preserved after it ----
    double loc_final = 0, cs = 0, G = 0, lbound0 = 0;
    if (b.pred && blockIdx.x == 0 && threadIdx.x == 0)
        keep_pred_update(b.pred, try_fast ? t_pred : 0.0, h_pred, nrem > 0 ? R / nrem : 0.0, n_cand);
    bool defer = false;  // bracketed solve: residual norm and line position come out of the line scan
    if (R / nrem < 1e-8) {
        nrem = 0;
    } else if (fast_done) {
        cs = fresh_cs;
        defer = true;
    } else if (fresh) {
        loc_final = fresh_loc;
        cs = fresh_cs;
        G = fresh_G;
        lbound0 = fresh_lb0;
    } else {
        double t = 0;
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) t += b.wt_remain[i];
        block_sum_pair(t, dummy, sh_d, sh_c);
        cs = t;
        grid_reduce_blk(grid, red, t, dummy);
        loc_final = t;
        G = t;
        if (multi) {
            comm_allgather(cm, cur, loc_final, 0.0, 0ull, sh_x0, sh_x1, sh_xc);
            comm_sum(cm, sh_x0, G, lbound0);
        }
    }
    FR_STAMP(b.st, 2);  // preserved set decided
    // CTA boundaries on the resampling line from the chunk sums of the residual weights
    double blk_lb, tot_lb;
    unsigned long long e0, e1;
    grid_excl_scan(grid, red, cs, defer ? fast_kc : 0ull, blk_lb, e0, tot_lb, e1, sh_sd, sh_sc);
    if (defer) {
        loc_final = tot_lb;
        G = tot_lb;
        lbound0 = 0;
        if (multi) {  // this rank's place on the global resampling line (seed_sys :107-127)
            comm_allgather(cm, cur, loc_final, 0.0, 0ull, sh_x0, sh_x1, sh_xc);
            comm_sum(cm, sh_x0, G, lbound0);
        } else if (e1 != kept_total && blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(&b.st->anomalies, 1ull << 32);  // the cut pass and the candidate rounds disagree: cannot happen
        }
    }
    SysGrid sg;
    if (nrem > 0) {
        sg.unit = G / nrem;
        long long j0 = (long long)(int)(lbound0 * nrem / G);
        double r = rn_uniform * sg.unit;
        r += sg.unit * (int)(lbound0 * nrem / G);
        if (r < lbound0) {
            r += sg.unit;
            j0++;
        }
        sg.rn0 = r;
        sg.inv = 1.0 / sg.unit;
        sg.n = (long long)nrem - j0;
    } else {
        sg.rn0 = INFINITY;
        sg.unit = INFINITY;
        sg.inv = 0;
        sg.n = 0;
    }

    FR_STAMP(b.st, 3);
    // pass 2: per-input lower bound + number of outputs
    double carry = lbound0 + blk_lb;
    unsigned long long my_out = 0;
    unsigned long long anomalies = 0;
    for (size_t base = lo; base < hi; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        bool act = i < hi;
        double v = act ? b.veff[i] : 0.0;
        double wr = (act && v != 0) ? b.wt_remain[i] : 0.0;
        double ex, tot;
        unsigned long long ec, tc;
        block_excl_scan(wr, 0ull, ex, ec, tot, tc, sh_sd, sh_sc);
        if (act) {
            double start = carry + ex;
            b.lb[i] = start;
            uint32_t k = 0, code = 0;
            if (v != 0) {
                uint32_t nd = b.ndiv[i];
                double lbound = start + wr;
                if (nd > 0) {
                    if (b.keep[i]) {
                        k = nd;
                    } else {
                        long long k0 = sg.count_below(start), k1 = sg.count_below(lbound);
                        k = (uint32_t)(k1 > k0 ? k1 - k0 : 0);
                    }
                } else {
                    long long k0 = sg.count_below(start);
                    double g = sg.point(k0);
                    if (wr < v || g < lbound) {
                        uint32_t ns = b.nsub[i], kb = b.keep[i];
                        double sub_lb = lbound - wr;
                        uint32_t n_kept_out = 0, s1 = 0, s2 = 0;
                        prov.visit(i, b.rinv[i], [&](uint32_t j, double wj) -> bool {
                            if (j >= ns) return false;
                            if (((kb >> j) & 1u) && wj != 0) {
                                k++;
                                n_kept_out++;
                            } else {
                                sub_lb += v * wj;
                                if (g < sub_lb && wj != 0) {
                                    if (k == 0) s1 = j;
                                    if (k == 1) s2 = j;
                                    k++;
                                    k0++;
                                    g = sg.point(k0);
                                    if (g < sub_lb) anomalies++;
                                }
                            }
                            // the rest of the row matters only while a grid point or a preserved piece is left in it
                            return g < lbound || ((unsigned long long)kb >> (j + 1)) != 0;
                        });
                        // one or two resampled outputs and nothing preserved: remember the sub-indices so that the
                        // emit pass does not have to regenerate the row (kcnt bits 30-31 = count, 16-20 / 21-25 = subs)
                        if (n_kept_out == 0 && k >= 1 && k <= 2) code = (k << 30) | (s1 << 16) | (s2 << 21);
                    }
                }
            }
            b.kcnt[i] = code ? code : k;
            my_out += k;
        }
        carry += tot;
    }
    my_out = block_sum_u64(my_out, sh_c);
    double d0, d1;
    unsigned long long blk_off, tot_out;
    grid_excl_scan(grid, red, 0.0, my_out, d0, blk_off, d1, tot_out, sh_sd, sh_sc);

    FR_STAMP(b.st, 4);  // count pass + offset scan done
    // pass 3: emit
    unsigned long long ocarry = blk_off;
    unsigned long long overflow = 0;
    const double samp_val = G / nrem;  // tmp_glob_norm / n_samp
    for (size_t base = lo; base < hi; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        bool act = i < hi;
        const uint32_t code = act ? b.kcnt[i] : 0;
        const uint32_t mode = code >> 30;
        uint32_t k = mode ? mode : code;
        double ex, tot;
        unsigned long long ec, tc;
        block_excl_scan(0.0, (unsigned long long)k, ex, ec, tot, tc, sh_sd, sh_sc);
#define FR_EMIT(VAL, SUB)                              \
    do {                                               \
        if (o < b.out_cap) {                           \
            b.out_val[o] = (VAL);                      \
            b.out_widx[o] = (uint32_t)i;               \
            b.out_sub[o] = (uint32_t)(SUB);             \
        } else {                                       \
            overflow++;                                \
        }                                              \
        o++;                                           \
    } while (0)
        if (act && mode) {  // one or two resampled outputs recorded by the count pass
            unsigned long long o = ocarry + ec;
            FR_EMIT(samp_val, (code >> 16) & 31u);
            if (mode == 2) FR_EMIT(samp_val, (code >> 21) & 31u);
        } else if (act && k > 0) {
            unsigned long long o = ocarry + ec;
            double v = b.veff[i];
            uint32_t nd = b.ndiv[i];
            double start = b.lb[i];
            double wr = b.wt_remain[i];
            double lbound = start + wr;
            if (nd > 0) {
                if (b.keep[i]) {
                    double each = v / nd;
                    for (uint32_t j = 0; j < nd; j++) FR_EMIT(each, j);
                } else {
                    long long k0 = sg.count_below(start);
                    for (uint32_t t = 0; t < k; t++) {
                        double g = sg.point(k0 + t);
                        unsigned long long sub = (unsigned long long)((lbound - g) * nd / v);
                        if (sub >= nd) {
                            sub = nd - 1;
                            anomalies++;
                        }
                        FR_EMIT(samp_val, sub);
                    }
                }
            } else {
                uint32_t ns = b.nsub[i], kb = b.keep[i];
                long long k0 = sg.count_below(start);
                double g = sg.point(k0);
                double sub_lb = lbound - wr;
                const unsigned long long o_end = o + k;
                prov.visit(i, b.rinv[i], [&](uint32_t j, double wj) -> bool {
                    if (j >= ns) return false;
                    if (((kb >> j) & 1u) && wj != 0) {
                        FR_EMIT(v * wj, j);
                    } else {
                        sub_lb += v * wj;
                        if (g < sub_lb && wj != 0) {
                            FR_EMIT(samp_val, j);
                            k0++;
                            g = sg.point(k0);
                        }
                    }
                    return o < o_end;  // all outputs of this input are written
                });
            }
#undef FR_EMIT
        }
        ocarry += tc;
    }
    overflow = block_sum_u64(overflow, sh_c);
    anomalies = block_sum_u64(anomalies, sh_c);
    if (threadIdx.x == 0) {
        if (overflow) atomicAdd(&b.st->overflow, overflow);
        if (anomalies) atomicAdd(&b.st->anomalies, anomalies);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        b.st->loc_norm = loc_final;
        b.st->glob_norm = s;
        b.st->n_samp_left = nrem;
        b.st->rounds = rounds;
        b.st->n_kept = kept_total;
        b.st->n_out = tot_out < b.out_cap ? tot_out : b.out_cap;
        b.st->n_in = n;
    }
    if (cm.n_ranks > 1) grid.sync();  // the epoch is stored once every CTA is past its last exchange
    FR_STAMP(b.st, 5);  // emit done (CTA 0)
    comm_end(cm, cur);
}

// host launchers (compress.cu).  d_st must be zeroed by the caller; pred / cand_x / cand_m == nullptr: plain rounds.
int fries_find_preserve_launch(fries_ctx *c, const double *d_values, size_t count, const unsigned long long *d_n,
                               unsigned n_samp, uint8_t *d_keep, CompState *d_st, double *pd, unsigned long long *pc,
                               int grid, const fries_comm *comm, KeepPred *pred = nullptr, double *cand_x = nullptr,
                               uint32_t *cand_m = nullptr);
int fries_sys_comp_launch(fries_ctx *c, double *d_values, size_t count, const unsigned long long *d_n, uint8_t *d_keep,
                          const double *d_in, double lbound0, double glob, long long n_samp, double rn, CompState *d_out,
                          double *pd, unsigned long long *pc, int grid, const fries_comm *comm);
