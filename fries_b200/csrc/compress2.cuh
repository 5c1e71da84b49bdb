// Second-generation hierarchical compression engine (a6: comp_sub = find_keep_sub + sys_sub,
// compress_utils.cpp:130-276,702-820), same semantics and same provider interface as comp_sub_engine
// (compress.cuh), rebuilt around what the round-1 profiles showed:
//
//   * at 2.6e5 inputs every thread of the old engine executed ~4000 instructions and 81 block barriers of pure
//     control skeleton (per-tile scans of ONE item per thread, 3-barrier reductions, a separate cut pass);
//   * at 1.25e7 inputs every 512-item tile cost ~3.5 us of barrier latency with one load in flight per thread.
//
// Here a thread owns FR2_ITEMS consecutive inputs of a 2048-input tile (vector loads issued before use, ONE block
// scan per tile), the preserved set is written by the pass that computes the effective weights anyway (everything at or
// above the bracket is preserved for certain; the few inputs with a piece INSIDE the bracket are revisited after the
// solve through the candidate list, which now carries the input index), and the inputs that need their sub-weight row
// in the count / emit passes are compacted per CTA first, so that the row loops run with full warps.
//
// Passes over the inputs: A prep + classification, B residual sums, C count, D emit (the old engine: prep, cut,
// [sum], count, emit); grid barriers on the fast path: 3 (after A, after B, after C).
// When no valid bracket exists (first call of a site, jump of the vector) the plain rounds of the reference run on the
// per-input state exactly as in the old engine.
#pragma once
#include "compress.cuh"
#include "gridcomb.cuh"

#define FR2_NT 512
#define FR2_ITEMS 4
#define FR2_TILE (FR2_NT * FR2_ITEMS)
#define FR2_NW (FR2_NT / 32)

#define FR2_CAND_MASK ((1ull << 40) - 1)
#define FR2_CAND_STAGE 1024  // candidates a CTA stages in shared memory before it reserves their place in the global list
struct StageShared {
    double cx[FR2_CAND_STAGE];             // staged candidates: magnitude, multiplicity, input index
    uint32_t cm[FR2_CAND_STAGE], ci[FR2_CAND_STAGE];
    unsigned n_stage;                      // staged so far (may run past the capacity: the excess went to the global list directly)
    unsigned long long stage_base;
};
struct Comp2Shared {
    StageShared stg;
    double wsum[2][FR2_NW + 1];            // warp partials of the block scans / sums (double buffered)
    unsigned long long wcnt[2][FR2_NW + 1];
    double start[FR2_TILE];                // line position of every input of the current tile
    unsigned short list[FR2_TILE];         // compacted slots of the inputs that need their row
    unsigned n_list;
    unsigned long long bc[8];              // broadcasts
};

// ---- lean block primitives: one barrier each; `buf` alternates so that no trailing barrier is needed ----
// exclusive prefix over the threads of the CTA + total, fixed association (warp tree, then the warp totals in order)
__device__ __forceinline__ void fr2_scan_d(double x, double &ex, double &tot, double *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    double prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = 0;
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    double pre = 0, all = 0;
#pragma unroll
    for (int q = 0; q < FR2_NW; q++) {
        double t = sh[q];
        if (q < w) pre += t;
        all += t;
    }
    ex = pre + prev;
    tot = all;
}
__device__ __forceinline__ void fr2_scan_u(unsigned x, unsigned &ex, unsigned &tot, unsigned long long *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    unsigned pre = 0, all = 0;
#pragma unroll
    for (int q = 0; q < FR2_NW; q++) {
        unsigned t = (unsigned)sh[q];
        if (q < w) pre += t;
        all += t;
    }
    ex = pre + inc - x;
    tot = all;
}
// K fused (double, u64) sums over the CTA; result in every thread; one barrier
template <int K>
__device__ __forceinline__ void fr2_sum(double (&d)[K], unsigned long long (&c)[K], double *shd, unsigned long long *shc,
                                        CompState *tl_st = nullptr, int tl_k = 0) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
            c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
        }
    }
    if (tl_st) FR_TL(tl_st, tl_k);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            shd[k * (FR2_NW + 1) + w] = d[k];
            shc[k * (FR2_NW + 1) + w] = c[k];
        }
    }
    __syncthreads();
    if (tl_st) FR_TL(tl_st, tl_k + 1);
#pragma unroll
    for (int k = 0; k < K; k++) {
        double t = 0;
        unsigned long long u = 0;
#pragma unroll
        for (int q = 0; q < FR2_NW; q++) {
            t += shd[k * (FR2_NW + 1) + q];
            u += shc[k * (FR2_NW + 1) + q];
        }
        d[k] = t;
        c[k] = u;
    }
}

// all-CTA sum of K (double, u64) pairs with one grid barrier; every CTA re-reduces the per-CTA partials in the same
// order (thread t loads partial t: ONE load latency, then the fixed block tree of fr2_sum), so the totals are
// bit-identical everywhere.  With `prefix` the first pair's exclusive prefix over the CTAs before this one comes back in
// (pre_d, pre_c).  Grids of at most FR2_NT CTAs.
template <int K>
__device__ __forceinline__ void fr2_grid_sum(cg::grid_group &grid, GridRed &r, double (&d)[K], unsigned long long (&c)[K],
                                             double *shd, unsigned long long *shc, bool prefix, double &pre_d,
                                             unsigned long long &pre_c, CompState *tl_st = nullptr, int tl_k = 0) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            __stcg(&r.pd[((size_t)r.parity * r.nb + blockIdx.x) * FR_RED_STRIDE + k], d[k]);
            __stcg(&r.pc[((size_t)r.parity * r.nb + blockIdx.x) * FR_RED_STRIDE + k], c[k]);
        }
    }
    if (tl_st) FR_TL(tl_st, tl_k);
    grid.sync();
    if (tl_st) FR_TL(tl_st, tl_k + 1);
    double t[K + 1];
    unsigned long long u[K + 1];
    const int i = (int)threadIdx.x;
#pragma unroll
    for (int k = 0; k < K; k++) {
        t[k] = i < r.nb ? __ldcg(&r.pd[((size_t)r.parity * r.nb + i) * FR_RED_STRIDE + k]) : 0.0;
        u[k] = i < r.nb ? __ldcg(&r.pc[((size_t)r.parity * r.nb + i) * FR_RED_STRIDE + k]) : 0ull;
    }
    t[K] = (prefix && i < (int)blockIdx.x) ? t[0] : 0.0;
    u[K] = (prefix && i < (int)blockIdx.x) ? u[0] : 0ull;
    if (tl_st) FR_TL(tl_st, tl_k + 2);
    fr2_sum<K + 1>(t, u, shd, shc, tl_st, tl_k + 4);
    if (tl_st) FR_TL(tl_st, tl_k + 3);
#pragma unroll
    for (int k = 0; k < K; k++) {
        d[k] = t[k];
        c[k] = u[k];
    }
    pre_d = t[K];
    pre_c = u[K];
    __syncthreads();  // shd / shc may be rewritten by the caller's next reduction
    r.parity ^= 1;
}

// candidate list with the input index of every candidate (local; the index never crosses ranks)
struct CandList2 {
    CandList base;
    uint32_t *idx;  // [FR_CAND_GCAP]
};
__device__ __forceinline__ void cand_append2(const CandList2 &cl, const CommView &cm, double x, uint32_t mult, uint32_t i) {
    const unsigned long long cap = cm.n_ranks > 1 ? FR_COMM_XCAP : FR_CAND_GCAP;
    if (*(volatile unsigned long long *)cl.base.count > cap) return;
    unsigned long long k = atomicAdd(cl.base.count, 1ull);
    if (k >= cap) return;
    cl.idx[k] = i;
    if (cm.n_ranks > 1) {
        for (int p = 0; p < cm.n_ranks; p++) {
            cm.cand_x[p][(size_t)cm.rank * FR_COMM_XCAP + k] = x;
            cm.cand_m[p][(size_t)cm.rank * FR_COMM_XCAP + k] = mult;
        }
    } else {
        cl.base.x[k] = x;
        cl.base.mult[k] = mult;
    }
}

// Lean variant of bracket_solve (compress.cuh) for lists that one CTA holds in registers (<= FR_CAND_CAP candidates, the
// normal case: the bracket is steered to ~2500).  Same arithmetic and the same result: Newton rounds on the candidates
// alone, sums as exact integers in units of ulp(t_lo).  What changed is the cost per round: a candidate's integer
// weight ix * mult fits 60 bits (mult <= 64, the HB-PP rows have at most 32 pieces; larger multiplicities leave the
// bracket to the plain rounds), so a round
// reduces TWO 64-bit words per thread (low halves | high halves + count) with one shuffle tree and two shared-memory
// atomics per warp instead of eleven 16-bit limbs with a vote each, and no 128-bit arithmetic per candidate.  The old
// solve cost ~1000 instructions per thread (12 us of a 60 us stage at 2.6e5 inputs, measured round 2).
// CTA-local: uses block barriers only, so ONE CTA may run it on behalf of the grid (comp_sub_engine2 does, on CTA 0, and
// sends the result to the others with the grid reduction's payload).  seg_end: multi-rank list layout (nullptr: one rank).
__device__ __forceinline__ BracketResult bracket_solve2_local(const CandList &cl, double R0, long long nrem0, double t_lo,
                                                              double t_hi, unsigned long long *shc, const CommView &cm,
                                                              const unsigned long long *seg_end, bool peers_ok,
                                                              unsigned long long ncand) {
    const bool multi = seg_end != nullptr;
    BracketResult res;
    res.valid = false;
    res.x_cut = t_hi;
    res.R = R0;
    res.nrem = 0;
    res.kept_cand = 0;
    res.rounds = 0;
    res.n_cand = ncand;
    if (!peers_ok || nrem0 <= 0 || nrem0 > 0xffffffffll) return res;
    if (!(t_hi * (double)nrem0 >= R0)) return res;  // H is not certainly preserved
    const int E_lo = (int)((__double_as_longlong(t_lo) >> 52) & 0x7ff);
    if (E_lo < 64 || E_lo > 1900) return res;
    const double ulp_lo = __longlong_as_double((long long)(E_lo - 52) << 52);
    double x[FR_CAND_PER_THREAD];
    uint32_t mu[FR_CAND_PER_THREAD];
    unsigned state = 0;
    bool big = false;
#pragma unroll
    for (int k = 0; k < FR_CAND_PER_THREAD; k++) {
        const unsigned long long idx = threadIdx.x + (unsigned long long)k * FR2_NT;
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
            if (mu[k] > 64u) big = true;
        }
    }
    // shared accumulators: [3 rotating buffers][2] u64 (see bracket_solve for the rotation argument), + the minimum
    unsigned long long *acc = shc;
    unsigned long long *acc_min = shc + 8;
    if (threadIdx.x < 6) acc[threadIdx.x] = 0;
    if (threadIdx.x == 6) *acc_min = 0x7ff0000000000000ull;
    if (__syncthreads_or(big)) return res;  // a multiplicity that does not fit: plain rounds
    const int lane = threadIdx.x & 31;
    unsigned long long cnt_tot = 0;
    unsigned __int128 sum_tot = 0;
    double R = R0;
    unsigned long long nrem = (unsigned long long)nrem0;
    double xmin = INFINITY;
    for (unsigned round = 0; round < 4096; round++) {
        unsigned long long *a = acc + 2 * (round % 3);
        unsigned long long w_lo = 0, w_hc = 0;  // low halves; high halves + (count << 40)
        const double fac = (double)nrem;
#pragma unroll
        for (int k = 0; k < FR_CAND_PER_THREAD; k++) {
            if (((state >> k) & 1u) && x[k] * fac >= R) {
                state &= ~(1u << k);
                const long long xb = __double_as_longlong(x[k]);
                const unsigned long long ix = ((unsigned long long)(xb & 0xfffffffffffffll) | (1ull << 52))
                                              << ((int)((xb >> 52) & 0x7ff) - E_lo);
                const unsigned long long p = ix * mu[k];  // < 2^54 * 2^6: high half < 2^28, a CTA's 4096 of them < 2^40
                w_lo += p & 0xffffffffull;
                w_hc += (p >> 32) + ((unsigned long long)mu[k] << 40);
                xmin = fmin(xmin, x[k]);
            }
        }
        if (__any_sync(0xffffffffu, w_hc != 0)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                w_lo += __shfl_xor_sync(0xffffffffu, w_lo, o);
                w_hc += __shfl_xor_sync(0xffffffffu, w_hc, o);
            }
            if (lane == 0) {
                atomicAdd(&a[0], w_lo);
                atomicAdd(&a[1], w_hc);
            }
        }
        if (threadIdx.x < 2) acc[2 * ((round + 1) % 3) + threadIdx.x] = 0;
        __syncthreads();
        res.rounds = round + 1;
        const unsigned long long t_lo64 = a[0], t_hc = a[1];
        const unsigned long long c_round = t_hc >> 40, hi_round = t_hc & ((1ull << 40) - 1);
        if (c_round == 0) break;
        cnt_tot += c_round;
        sum_tot += ((unsigned __int128)hi_round << 32) + t_lo64;
        if (cnt_tot >= (unsigned long long)nrem0) return res;  // budget exhausted inside the bracket
        nrem = (unsigned long long)nrem0 - cnt_tot;
        double kept_sum = (double)(unsigned long long)(sum_tot >> 64) * 18446744073709551616.0 +
                          (double)(unsigned long long)sum_tot;
        R = R0 - kept_sum * ulp_lo;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    if (lane == 0 && xmin < INFINITY) atomicMin(acc_min, (unsigned long long)__double_as_longlong(xmin));
    __syncthreads();
    xmin = __longlong_as_double((long long)*acc_min);
    __syncthreads();
    res.x_cut = xmin < t_hi ? xmin : t_hi;
    res.R = R;
    res.nrem = (unsigned)nrem;
    res.kept_cand = cnt_tot;
    res.valid = t_lo * (double)nrem < R;
    return res;
}

// every CTA solves (multi-rank, or lists beyond one CTA's registers: the distributed rounds of bracket_solve)
__device__ __forceinline__ BracketResult bracket_solve2(cg::grid_group &grid, const CandList &cl, unsigned long long *gacc,
                                                        double R0, long long nrem0, double t_lo, double t_hi, double *shd,
                                                        unsigned long long *shc, const CommView &cm,
                                                        const unsigned long long *seg_len, bool peers_ok,
                                                        unsigned long long n_local) {
    const bool multi = cm.n_ranks > 1;
    unsigned long long ncand = n_local;
    unsigned long long seg_end[FR_MAX_RANKS];
    if (multi) {
        unsigned long long t = 0;
        for (int p = 0; p < FR_MAX_RANKS; p++) {
            if (p < cm.n_ranks) t += seg_len[p];
            seg_end[p] = t;
        }
        ncand = t;
    }
    if (ncand > FR_CAND_CAP) return bracket_solve(grid, cl, gacc, R0, nrem0, t_lo, t_hi, shd, shc, cm, seg_len, peers_ok);
    return bracket_solve2_local(cl, R0, nrem0, t_lo, t_hi, shc, cm, multi ? seg_end : nullptr, peers_ok, ncand);
}

// Candidates are staged per CTA: the round-1 engine appended every candidate with one atomicAdd on ONE global counter and a
// volatile read of it -- ~2500 same-address L2 operations per stage, which an L2 slice serves one at a time (~23 cycles
// each, measured round 2): the appends of a stage queued for tens of microseconds and produced the long tail of pass A.
// Now a candidate costs a shared-memory atomic; after the pass the CTA reserves its range with one global atomicAdd.
__device__ __forceinline__ void cand_stage(StageShared &sm, const CandList2 &cl, const CommView &cm, double x, uint32_t mult,
                                           uint32_t i) {
    const unsigned k = atomicAdd(&sm.n_stage, 1u);
    if (k < FR2_CAND_STAGE) {
        sm.cx[k] = x;
        sm.cm[k] = mult;
        sm.ci[k] = i;
    } else {
        cand_append2(cl, cm, x, mult, i);  // staging area full (large chunks): straight to the global list
    }
}
// after the pass: every thread of the CTA calls this
__device__ __forceinline__ void cand_stage_flush(StageShared &sm, const CandList2 &cl, const CommView &cm) {
    __syncthreads();
    const unsigned ns = sm.n_stage < FR2_CAND_STAGE ? sm.n_stage : FR2_CAND_STAGE;
    if (threadIdx.x == 0) sm.stage_base = ns ? atomicAdd(cl.base.count, (unsigned long long)ns) : 0ull;
    __syncthreads();
    const unsigned long long base = sm.stage_base, cap = cm.n_ranks > 1 ? FR_COMM_XCAP : FR_CAND_GCAP;
    for (unsigned e = threadIdx.x; e < ns; e += blockDim.x) {
        const unsigned long long k = base + e;
        if (k >= cap) continue;  // the list overflowed: the bracket is invalid (the count says so)
        cl.idx[k] = sm.ci[e];
        if (cm.n_ranks > 1) {
            for (int p = 0; p < cm.n_ranks; p++) {
                cm.cand_x[p][(size_t)cm.rank * FR_COMM_XCAP + k] = sm.cx[e];
                cm.cand_m[p][(size_t)cm.rank * FR_COMM_XCAP + k] = sm.cm[e];
            }
        } else {
            cl.base.x[k] = sm.cx[e];
            cl.base.mult[k] = sm.cm[e];
        }
    }
}

// The solve for lists beyond one CTA's registers, single rank: every CTA runs the Newton rounds on the candidates it staged
// itself (shared memory; no CTA's staging area may have overflowed), one barrier with payload per round for the exact
// integer sums.  CTA 0 streaming the whole list from L2 took 75 us for 6e4 candidates, cg's grid.sync rounds of
// bracket_solve ~30 us (measured round 2, H2O-sized run); this is ~5 us per round.  Same arithmetic as
// bracket_solve2_local; the cut is the smallest double that passes a round's test (x -> x * fac is monotone), which
// selects the same candidates as the smallest kept candidate does.  Every thread of every CTA calls it.
__device__ __forceinline__ BracketResult bracket_solve_staged(const StageShared &stg, const GridComb &gcb, GridCombShared &gsh,
                                                              GridCombCursor &gcur, double R0, long long nrem0, double t_lo,
                                                              double t_hi, double *sh_sd, unsigned long long *sh_sc,
                                                              unsigned long long ncand) {
    BracketResult br;
    br.n_cand = ncand;
    br.valid = false;
    br.x_cut = t_hi;
    br.R = R0;
    br.nrem = 0;
    br.kept_cand = 0;
    br.rounds = 0;
    const int tid = threadIdx.x;
    const int E_lo = (int)((__double_as_longlong(t_lo) >> 52) & 0x7ff);
    if (!(nrem0 > 0 && nrem0 <= 0xffffffffll && t_hi * (double)nrem0 >= R0 && E_lo >= 64 && E_lo <= 1900)) return br;
    const double ulp_lo = __longlong_as_double((long long)(E_lo - 52) << 52);
    const unsigned nst = stg.n_stage;  // <= FR2_CAND_STAGE on every CTA
    unsigned live = 0;                 // bit q: staged candidate tid + q * FR2_NT not counted yet
    bool big = false;
    for (unsigned q = 0; q * FR2_NT < FR2_CAND_STAGE; q++)
        if (tid + q * FR2_NT < nst) {
            live |= 1u << q;
            if (stg.cm[tid + q * FR2_NT] > 64u) big = true;  // its integer weight would not fit 60 bits: plain rounds
        }
    unsigned long long cnt_tot = 0, nrem_r = (unsigned long long)nrem0;
    unsigned __int128 sum_tot = 0;
    double R_r = R0, cut = t_hi, pre_d;
    unsigned long long pre_c;
    for (unsigned round = 0; round < 4096; round++) {
        const double fac = (double)nrem_r;
        unsigned long long w_lo = 0, w_hi = 0, w_c = 0;
        for (unsigned q = 0; q * FR2_NT < FR2_CAND_STAGE; q++) {
            if (!((live >> q) & 1u)) continue;
            const double x = stg.cx[tid + q * FR2_NT];
            if (x * fac >= R_r) {
                live &= ~(1u << q);
                const unsigned mu = stg.cm[tid + q * FR2_NT];
                const long long xb = __double_as_longlong(x);
                const unsigned long long ix = ((unsigned long long)(xb & 0xfffffffffffffll) | (1ull << 52))
                                              << ((int)((xb >> 52) & 0x7ff) - E_lo);
                const unsigned long long p = ix * mu;
                w_lo += p & 0xffffffffull;
                w_hi += p >> 32;
                w_c += mu;
            }
        }
        double rd[2] = {(double)w_c, (round == 0 && big) ? 1.0 : 0.0};
        unsigned long long rc[2] = {w_lo, w_hi};
        fr2_sum<2>(rd, rc, sh_sd, sh_sc);
        grid_comb<2>(gcb, gsh, gcur, rd, rc, false, false, pre_d, pre_c);
        br.rounds = round + 1;
        if (round == 0 && rd[1] != 0.0) return br;  // some CTA holds a multiplicity that does not fit
        const unsigned long long c_round = (unsigned long long)rd[0];
        if (c_round == 0) break;
        double bnd = R_r / fac;  // smallest x with x * fac >= R_r
        while (__longlong_as_double(__double_as_longlong(bnd) - 1) * fac >= R_r) bnd = __longlong_as_double(__double_as_longlong(bnd) - 1);
        while (bnd * fac < R_r) bnd = __longlong_as_double(__double_as_longlong(bnd) + 1);
        cut = fmin(cut, bnd);
        cnt_tot += c_round;
        sum_tot += ((unsigned __int128)rc[1] << 32) + rc[0];
        if (cnt_tot >= (unsigned long long)nrem0) return br;  // budget exhausted inside the bracket
        nrem_r = (unsigned long long)nrem0 - cnt_tot;
        const double kept_sum = (double)(unsigned long long)(sum_tot >> 64) * 18446744073709551616.0 +
                                (double)(unsigned long long)sum_tot;
        R_r = R0 - kept_sum * ulp_lo;
    }
    br.x_cut = cut < t_hi ? cut : t_hi;
    br.R = R_r;
    br.nrem = (unsigned)nrem_r;
    br.kept_cand = cnt_tot;
    br.valid = t_lo * (double)nrem_r < R_r && cut > t_lo;
    return br;
}

// single rank, out of line (rare at the sizes where code size matters): the grid-distributed rounds of compress.cuh on the
// global candidate list
static __device__ __noinline__ BracketResult bracket_solve_global(const CandList cl, unsigned long long *gacc, double R0,
                                                                  long long nrem0, double t_lo, double t_hi, double *shd,
                                                                  unsigned long long *shc, const CommView cm,
                                                                  unsigned long long n_local) {
    cg::grid_group grid = cg::this_grid();
    return bracket_solve(grid, cl, gacc, R0, nrem0, t_lo, t_hi, shd, shc, cm, nullptr, n_local <= FR_CAND_GCAP);
}

struct CompSubBufs2 {
    CompSubBufs b;      // per-input state, outputs, reduction scratch, prediction (compress.cuh)
    uint32_t *cand_idx; // [FR_CAND_GCAP]
    unsigned long long *gcomb;      // GridComb state (gridcomb.cuh), [GC_STATE_WORDS], zeroed at allocation
    unsigned long long *cta_marks;  // diagnostics (may be nullptr): [8][gridDim.x] %globaltimer of every CTA at the phase ends
};
#define FR2_CTA_MARK(b2, k)                                                                              \
    do {                                                                                                 \
        if ((b2).cta_marks && threadIdx.x == 0) (b2).cta_marks[(size_t)(k) * gridDim.x + blockIdx.x] = fr_globaltimer(); \
    } while (0)


// four consecutive per-input values of a thread (i0 is a multiple of 4 and the state arrays are 256-byte aligned: 128-bit
// accesses on full groups, element-wise at the ragged end of a chunk)
__device__ __forceinline__ void fr2_ld4(const double *p, size_t i0, size_t hi, double (&o)[FR2_ITEMS]) {
    if (i0 + FR2_ITEMS <= hi) {
        const double2 a = *reinterpret_cast<const double2 *>(p + i0), c = *reinterpret_cast<const double2 *>(p + i0 + 2);
        o[0] = a.x; o[1] = a.y; o[2] = c.x; o[3] = c.y;
    } else {
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) o[k] = i0 + k < hi ? p[i0 + k] : 0.0;
    }
}
__device__ __forceinline__ void fr2_ld4(const uint32_t *p, size_t i0, size_t hi, uint32_t (&o)[FR2_ITEMS], uint32_t fill) {
    if (i0 + FR2_ITEMS <= hi) {
        const uint4 a = *reinterpret_cast<const uint4 *>(p + i0);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    } else {
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) o[k] = i0 + k < hi ? p[i0 + k] : fill;
    }
}
__device__ __forceinline__ void fr2_st4(double *p, size_t i0, size_t hi, const double (&v)[FR2_ITEMS]) {
    if (i0 + FR2_ITEMS <= hi) {
        *reinterpret_cast<double2 *>(p + i0) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2 *>(p + i0 + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++)
            if (i0 + k < hi) p[i0 + k] = v[k];
    }
}

// preserved pieces and residual weight of input i for the cut `x_cut` (x >= x_cut <=> preserved); the arithmetic of
// the old engine's cut pass: the residual of a partially preserved row is the sum of its other pieces in row order
template <class P>
__device__ __forceinline__ void fr2_apply_cut(P &prov, const CompSubBufs &b, size_t i, double x_cut, unsigned long long &kc) {
    const double v = b.veff[i];
    const uint32_t nd = b.ndiv[i];
    if (nd > 0) {
        const bool kp = v / nd >= x_cut;
        b.keep[i] = kp ? 1u : 0u;
        b.wt_remain[i] = kp ? 0.0 : v;
        if (kp) kc += nd;
    } else {
        const uint32_t ns = b.nsub[i];
        uint32_t kb = 0;
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
        b.wt_remain[i] = kb ? sub_remain : v;
    }
}

// MULTI = false: a build without the cross-rank code (the exchanges, the remote candidate stores, the redundant solve) --
// the single-rank kernels are ~30 % smaller, which the instruction cache notices (no-instruction stalls were 18 % of the
// stall samples of a 450 kB stage kernel with two CTAs per SM in different phases, ncu round 2).
template <bool MULTI, class P>
__device__ void comp_sub_engine2(P &prov, const CompSubBufs2 &b2, unsigned n_samp_in, double rn_uniform) {
    const CompSubBufs &b = b2.b;
    cg::grid_group grid = cg::this_grid();
    __shared__ Comp2Shared sm;
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[6 * 33];
    __shared__ unsigned long long sh_sc[6 * 33];
    GridRed red{b.part_d, b.part_c, 0, (int)gridDim.x, sh_d, sh_c};
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ GridCombShared gsh;
    const GridComb gcb{b2.gcomb};
    (void)red;

    if (tid == 0) {
        sm.bc[0] = (unsigned long long)prov.count();
        sm.bc[1] = b.pred ? (unsigned long long)__double_as_longlong(__ldcg(&b.pred->t)) : 0ull;
        sm.bc[2] = b.pred ? (unsigned long long)__double_as_longlong(__ldcg(&b.pred->h)) : 0ull;
        sm.bc[3] = (unsigned long long)grid_comb_begin(gcb).epoch;
        sm.stg.n_stage = 0;
    }
    __syncthreads();
    const size_t n = (size_t)sm.bc[0];
    GridCombCursor gcur;
    gcur.epoch = (unsigned)sm.bc[3];
    // contiguous chunk per CTA, a multiple of 128 inputs; processed in tiles of FR2_TILE (a thread owns FR2_ITEMS
    // consecutive inputs of a tile)
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 127) & ~(size_t)127;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;

    __shared__ double sh_x0[FR_MAX_RANKS], sh_x1[FR_MAX_RANKS];
    __shared__ double sh_xv[12][FR_MAX_RANKS];
    __shared__ unsigned long long sh_xc[FR_MAX_RANKS];
    __shared__ unsigned long long sh_seg[FR_MAX_RANKS];
    const CommView &cm = b.cm;
    CommCursor cur = comm_begin(cm);
    const bool multi = MULTI && cm.n_ranks > 1;
    CandList2 cand{b.cand, b2.cand_idx};

    FR_STAMP(b.st, 0);
    FR_TL(b.st, 0);
    FR2_CTA_MARK(b2, 0);
    const double t_pred = __longlong_as_double((long long)sm.bc[1]), h_pred = __longlong_as_double((long long)sm.bc[2]);
    const bool try_fast = t_pred > 0 && h_pred > 0 && h_pred < 0.25;
    const double t_lo = t_pred * (1.0 - h_pred), t_hi = t_pred * (1.0 + h_pred);

    // ---- pass A: effective weights (find_keep_sub :134-137) and, with a bracket, the preserved set above it ----
    double s = 0, s_hi = 0;
    unsigned long long c_hi = 0;
    bool appended = false;
    unsigned long long n_app = 0;
    for (size_t base = lo; base < hi; base += FR2_TILE) {
        // the thread's inputs of this tile, fetched level by level (independent loads in flight) before any row is generated
        typename P::Pre pre[FR2_ITEMS];
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) {
            const size_t i = base + (size_t)k * FR2_NT + tid;  // striped: this pass has no scan, the accesses coalesce
            if (i < hi) prov.fetch1(i, pre[k]);
        }
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++)
            if (base + (size_t)k * FR2_NT + tid < hi) prov.fetch2(pre[k]);
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++)
            if (base + (size_t)k * FR2_NT + tid < hi) prov.fetch3(pre[k]);
#pragma unroll 1
        for (int k = 0; k < FR2_ITEMS; k++) {
            const size_t i = base + (size_t)k * FR2_NT + tid;
            if (i >= hi) break;
            double v, rinv = 1.0, wmax = 1.0;
            uint32_t nd, ns;
            prov.prep_core(i, pre[k], v, nd, ns, rinv, wmax);
            b.rinv[i] = rinv;
            b.veff[i] = v;
            b.ndiv[i] = nd;
            b.nsub[i] = (uint8_t)ns;
            s += v;
            double wr = v;
            uint32_t kb = 0;
            const double xmax = nd > 0 ? v / nd : v * wmax;
            if (__builtin_expect(try_fast && xmax >= t_lo, 0)) {  // a few thousand of 1e6 inputs
                if (nd > 0) {
                    if (xmax >= t_hi) {
                        c_hi += nd;
                        s_hi += v;
                        kb = 1;
                        wr = 0;
                    } else {
                        cand_stage(sm.stg, cand, cm, xmax, nd, (uint32_t)i);
                        appended = true;
                        n_app++;
                    }
                } else {
                    double sub_remain = 0;
                    prov.visit(i, rinv, [&](uint32_t j, double wj) {
                        if (j < ns) {
                            double x = v * wj;
                            if (x >= t_hi) {
                                c_hi++;
                                s_hi += x;
                                kb |= 1u << j;
                            } else {
                                sub_remain += x;
                                if (x >= t_lo) {
                                    cand_stage(sm.stg, cand, cm, x, 1u, (uint32_t)i);
                                    appended = true;
                                    n_app++;
                                }
                            }
                        }
                    });
                    if (kb) wr = sub_remain;
                }
            }
            b.keep[i] = kb;
            b.wt_remain[i] = wr;
        }
    }
    FR_TL(b.st, 1);  // pass A loop done (this thread)
    if (try_fast) cand_stage_flush(sm.stg, cand, cm);
    cand_flush(cm, try_fast);
    double pre_d;
    unsigned long long pre_c;
    unsigned long long my_cand = 0, stage_overflows = 0;
    BracketResult br;
    br.valid = false;
    br.n_cand = 0;
    bool have_br = false;  // the solve ran on CTA 0 and its result came with the reduction
    {
        double dd[2] = {s, s_hi};
        unsigned long long cc[2] = {c_hi, n_app};  // above bit 40 of the count: CTAs whose staging area overflowed
        fr2_sum<2>(dd, cc, sh_sd, sh_sc);
        if (sm.stg.n_stage > FR2_CAND_STAGE) cc[1] |= 1ull << 40;
        FR_TL(b.st, 2);  // CTA sum done
        FR2_CTA_MARK(b2, 1);
        // one barrier with payload: the sums, and -- single rank, list within one CTA's registers -- the threshold solve
        // itself, run by CTA 0 alone: the other CTAs never read the candidate list (148 CTAs loading the same 30 kB
        // queue on the same L2 lines: ~2 us of skew per stage, measured round 2)
        const unsigned tag = grid_comb_next_tag(gcur);
        gc_post<2>(gcb, tag, dd, cc, true);  // fence: the candidate lists cross CTAs
        unsigned long long ex[6] = {0, 0, 0, 0, 0, 0};
        if (blockIdx.x == 0) {
            double td[2];
            unsigned long long tc[2];
            FR_TL(b.st, 25);  // posted
            gc_reduce<2>(gcb, gsh, tag, td, tc, true);
            FR_TL(b.st, 26);  // reduced
            if (!multi && try_fast && (tc[1] & FR2_CAND_MASK) <= FR_CAND_CAP && (tc[1] >> 40) == 0) {
                BracketResult r = bracket_solve2_local(b.cand, td[0] - td[1], (long long)n_samp_in - (long long)tc[0], t_lo, t_hi,
                                                       sh_sc, cm, nullptr, true, tc[1]);
                ex[0] = (unsigned long long)__double_as_longlong(r.x_cut);
                ex[1] = (unsigned long long)__double_as_longlong(r.R);
                ex[2] = (unsigned long long)r.nrem | ((unsigned long long)r.rounds << 32);
                ex[3] = r.kept_cand;
                ex[4] = (r.valid ? 1ull : 0ull) | (1ull << 8);
                ex[5] = r.n_cand;
            }
            FR_TL(b.st, 27);  // solved
            gc_publish<2, 6>(gcb, tag, td, tc, ex);
            FR_TL(b.st, 28);  // published
        }
        gc_wait<2, 6>(gcb, gsh, tag, dd, cc, pre_d, pre_c, ex, true);
        gcur.epoch = tag;
        FR_TL(b.st, 3);  // grid sum (+ solve) done
        s = dd[0];
        s_hi = dd[1];
        c_hi = cc[0];
        my_cand = cc[1] & FR2_CAND_MASK;  // candidates this rank appended (may exceed the capacity: then the bracket is invalid)
        stage_overflows = cc[1] >> 40;
        if ((ex[4] >> 8) == 1) {
            have_br = true;
            br.x_cut = __longlong_as_double((long long)ex[0]);
            br.R = __longlong_as_double((long long)ex[1]);
            br.nrem = (unsigned)ex[2];
            br.rounds = (unsigned)(ex[2] >> 32);
            br.kept_cand = ex[3];
            br.valid = (ex[4] & 1ull) != 0;
            br.n_cand = ex[5];
        }
    }
    double loc = s, R_next = s;
    bool peers_ok = try_fast && my_cand <= (multi ? (unsigned long long)FR_COMM_XCAP : (unsigned long long)FR_CAND_GCAP);
    if (multi) {
        double pay[5] = {s, s_hi, (double)c_hi, peers_ok ? 1.0 : 0.0, (double)my_cand};
        comm_allgather_v(cm, cur, pay, 5, sh_xv);
        double before;
        comm_sum(cm, sh_xv[0], R_next, before);
        s = R_next;
        double gs = 0, gc = 0;
        for (int p = 0; p < cm.n_ranks; p++) {
            gs += sh_xv[1][p];
            gc += sh_xv[2][p];
            if (sh_xv[3][p] == 0.0) peers_ok = false;
        }
        if (tid < cm.n_ranks) sh_seg[tid] = (unsigned long long)sh_xv[4][tid];
        s_hi = gs;
        c_hi = (unsigned long long)gc;
        if (try_fast) __threadfence_system();
        __syncthreads();
    }
    FR_STAMP(b.st, 1);

    // ---- the preserved set ----
    unsigned nrem = n_samp_in;
    double R = 0;
    unsigned rounds = 0;
    unsigned long long kept_total = 0, n_cand = 0;
    bool fast_done = false;
    if (try_fast) {
        if (!have_br && multi)
            br = bracket_solve2(grid, b.cand, b.st->gacc, s - s_hi, (long long)n_samp_in - (long long)c_hi, t_lo, t_hi, sh_sd, sh_sc,
                                cm, sh_seg, peers_ok, my_cand);
        else if (!have_br && stage_overflows == 0)  // single rank, a list beyond one CTA's registers
            br = bracket_solve_staged(sm.stg, gcb, gsh, gcur, s - s_hi, (long long)n_samp_in - (long long)c_hi, t_lo, t_hi, sh_sd,
                                      sh_sc, my_cand);
        else if (!have_br)  // some CTA staged more than its area holds (1e7 inputs: 3e5 - 5e5 candidates): the global list
            br = bracket_solve_global(b.cand, b.st->gacc, s - s_hi, (long long)n_samp_in - (long long)c_hi, t_lo, t_hi, sh_sd, sh_sc,
                                      cm, my_cand);
        n_cand = br.n_cand;
        FR_STAMP(b.st, 6);
        FR_TL(b.st, 4);  // solve done
        if (br.valid) {
            // the inputs with a piece inside the bracket: this CTA's staged candidates (an input listed twice is rewritten
            // with the same result); if the staging area overflowed, the CTA's share of the global list as well
            unsigned long long kc_dummy = 0;
            const unsigned nst = sm.stg.n_stage < FR2_CAND_STAGE ? sm.stg.n_stage : FR2_CAND_STAGE;
            for (unsigned e = tid; e < nst; e += FR2_NT) fr2_apply_cut(prov, b, (size_t)sm.stg.ci[e], br.x_cut, kc_dummy);
            if (sm.stg.n_stage > FR2_CAND_STAGE) {
                for (unsigned long long k = tid; k < my_cand; k += FR2_NT) {
                    const size_t i = (size_t)__ldcg(&b2.cand_idx[k]);
                    if (i >= lo && i < hi) fr2_apply_cut(prov, b, i, br.x_cut, kc_dummy);
                }
            }
            kept_total = c_hi + br.kept_cand;
            nrem = br.nrem;
            R = br.R;
            rounds = br.rounds;
            fast_done = true;
            if (blockIdx.x == 0 && tid == 0) b.st->fast = 1;
            FR_TL(b.st, 5);  // fix-up loop done (this thread)
            __syncthreads();
            FR_STAMP(b.st, 7);
        } else {
            // no valid bracket: forget the classification and run the plain rounds
            for (size_t i = lo + tid; i < hi; i += FR2_NT) {
                b.keep[i] = 0;
                b.wt_remain[i] = b.veff[i];
            }
            __syncthreads();
        }
    }
    if (!fast_done) {
        // plain rounds (find_keep_sub :153-265), as in comp_sub_engine
        unsigned long long glob_sampled = 1, dummy = 0;
        int last_pass = 0;
        while (glob_sampled > 0 && rounds < 100000) {
            R = R_next;
            if (R < 0) break;
            const double wt_factor = (double)nrem;
            double rem = 0;
            unsigned long long cnt = 0;
            for (size_t i = lo + tid; i < hi; i += FR2_NT) {
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
                                    double smg = cw * wj;
                                    double eps = j < full ? 1e-12 : 1e-10;
                                    if (smg >= R && fabs(smg) > eps) {
                                        kb |= 1u << j;
                                        cnt++;
                                    } else {
                                        sub_remain += smg;
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
            {
                double dd[1] = {rem};
                unsigned long long cc[1] = {cnt};
                fr2_sum<1>(dd, cc, sh_sd, sh_sc);
                grid_comb<1>(gcb, gsh, gcur, dd, cc, false, false, pre_d, pre_c);
                rem = dd[0];
                cnt = cc[0];
            }
            loc -= rem;
            R_next = loc;
            if (multi) {
                double before;
                comm_allgather(cm, cur, loc, 0.0, cnt, sh_x0, sh_x1, sh_xc);
                comm_sum(cm, sh_x0, R_next, before);
                cnt = comm_sum_u64(cm, sh_xc);
            }
            glob_sampled = cnt;
            nrem -= (unsigned)cnt;
            kept_total += cnt;
            rounds++;
            if (last_pass && glob_sampled) last_pass = 0;
            if (glob_sampled == 0 && !last_pass) {
                last_pass = 1;
                glob_sampled = 1;
                double t = 0;
                for (size_t i = lo + tid; i < hi; i += FR2_NT) t += b.wt_remain[i];
                double dd[1] = {t};
                unsigned long long cc[1] = {dummy};
                fr2_sum<1>(dd, cc, sh_sd, sh_sc);
                grid_comb<1>(gcb, gsh, gcur, dd, cc, false, false, pre_d, pre_c);
                loc = dd[0];
                R_next = dd[0];
                if (multi) {
                    double before;
                    comm_allgather(cm, cur, loc, 0.0, 0ull, sh_x0, sh_x1, sh_xc);
                    comm_sum(cm, sh_x0, R_next, before);
                }
            }
        }
    }
    if (b.pred && blockIdx.x == 0 && tid == 0)
        keep_pred_update(b.pred, try_fast ? t_pred : 0.0, h_pred, nrem > 0 ? R / nrem : 0.0, n_cand);
    if (R / nrem < 1e-8) nrem = 0;
    FR_STAMP(b.st, 2);
    FR_TL(b.st, 6);
    FR2_CTA_MARK(b2, 2);

    // ---- pass B: this CTA's share of the resampling line (sum of the residual weights, fixed order) ----
    double cs = 0;
    for (size_t base = lo; base < hi; base += FR2_TILE) {
        const size_t i0 = base + (size_t)tid * FR2_ITEMS;
        double w4[FR2_ITEMS];
        fr2_ld4(b.wt_remain, i0, hi, w4);
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) cs += w4[k];
    }
    double blk_lb, loc_final, G, lbound0 = 0;
    {
        double dd[1] = {cs};
        unsigned long long cc[1] = {0ull};
        FR_TL(b.st, 7);  // pass B loads + thread sum
        fr2_sum<1>(dd, cc, sh_sd, sh_sc);
        FR_TL(b.st, 8);
        FR2_CTA_MARK(b2, 3);
        grid_comb<1>(gcb, gsh, gcur, dd, cc, true, false, blk_lb, pre_c);
        FR_TL(b.st, 9);
        loc_final = dd[0];
    }
    if (nrem == 0) loc_final = 0;  // find_keep_sub :267-275: nothing left to resample
    G = loc_final;
    if (multi) {
        comm_allgather(cm, cur, loc_final, 0.0, 0ull, sh_x0, sh_x1, sh_xc);
        comm_sum(cm, sh_x0, G, lbound0);
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
    FR_TL(b.st, 10);

    // ---- pass C: line position and number of outputs of every input ----
    double carry = lbound0 + blk_lb;
    unsigned long long my_out = 0, anomalies = 0;
    int buf = 0;
    for (size_t base = lo; base < hi; base += FR2_TILE) {
        const size_t i0 = base + (size_t)tid * FR2_ITEMS;
        double v4[FR2_ITEMS], w4[FR2_ITEMS], st4[FR2_ITEMS];
        uint32_t nd4[FR2_ITEMS];
        fr2_ld4(b.veff, i0, hi, v4);
        fr2_ld4(b.wt_remain, i0, hi, w4);
        fr2_ld4(b.ndiv, i0, hi, nd4, 1u);
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++)
            if (v4[k] == 0) w4[k] = 0;
        double tsum = 0;
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) tsum += w4[k];
        if (base == lo) FR_TL(b.st, 11);  // first tile: loads done
        double ex, tot;
        fr2_scan_d(tsum, ex, tot, sm.wsum[buf]);
        if (base == lo) FR_TL(b.st, 12);  // scan done
        double st0 = carry + ex;
        unsigned need = 0;
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) {
            const size_t i = i0 + k;
            const double start = st0;
            st0 += w4[k];
            sm.start[tid * FR2_ITEMS + k] = start;
            st4[k] = start;
            if (i < hi) {
                uint32_t kk = 0;
                bool row = false;
                if (v4[k] != 0) {
                    const double lbound = start + w4[k];
                    if (nd4[k] > 0) {
                        if (w4[k] == 0) {  // preserved (a uniform input with v != 0 has a residual unless it is preserved)
                            kk = nd4[k];
                        } else {
                            const double k0 = sg.count_below_d(start), k1 = sg.count_below_d(lbound);
                            kk = (uint32_t)(k1 > k0 ? k1 - k0 : 0.0);
                        }
                    } else {
                        const double g = sg.point_d(sg.count_below_d(start));
                        row = w4[k] < v4[k] || g < lbound;
                    }
                }
                if (row)
                    need |= 1u << k;
                else {
                    b.kcnt[i] = kk;
                    my_out += kk;
                }
            }
        }
        if (base == lo) FR_TL(b.st, 13);  // grid tests done
        fr2_st4(b.lb, i0, hi, st4);
        // compact the inputs that need their row
        unsigned nneed = __popc(need), nex, ntot;
        fr2_scan_u(nneed, nex, ntot, sm.wcnt[buf]);
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++)
            if ((need >> k) & 1u) sm.list[nex++] = (unsigned short)(tid * FR2_ITEMS + k);
        __syncthreads();
        if (base == lo) FR_TL(b.st, 14);  // compaction done
        for (unsigned e = tid; e < ntot; e += FR2_NT) {
            const unsigned slot = sm.list[e];
            const size_t i = base + slot;
            const double v = b.veff[i], wr = b.wt_remain[i], start = sm.start[slot];
            const double lbound = start + wr;
            const uint32_t ns = b.nsub[i], kb = b.keep[i];
            double k0 = sg.count_below_d(start);
            double g = sg.point_d(k0);
            double sub_lb = lbound - wr;
            uint32_t k = 0, n_kept_out = 0, s1 = 0, s2 = 0;
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
                        k0 += 1.0;
                        g = sg.point_d(k0);
                        if (g < sub_lb) anomalies++;
                    }
                }
                return g < lbound || ((unsigned long long)kb >> (j + 1)) != 0;
            });
            uint32_t code = 0;
            if (n_kept_out == 0 && k >= 1 && k <= 2) code = (k << 30) | (s1 << 16) | (s2 << 21);
            b.kcnt[i] = code ? code : k;
            my_out += k;
        }
        if (base == lo) FR_TL(b.st, 15);  // row loop done (this thread)
        carry += tot;
        buf ^= 1;
        __syncthreads();  // sm.start / sm.list are rewritten by the next tile
        if (base == lo) FR_TL(b.st, 16);
    }
    unsigned long long blk_off, tot_out;
    {
        double dd[1] = {0.0};
        unsigned long long cc[1] = {my_out};
        FR_TL(b.st, 17);  // pass C done
        fr2_sum<1>(dd, cc, sh_sd, sh_sc);
        FR2_CTA_MARK(b2, 4);
        double d0;
        grid_comb<1>(gcb, gsh, gcur, dd, cc, true, false, d0, blk_off);
        FR_TL(b.st, 18);
        tot_out = cc[0];
    }
    FR_STAMP(b.st, 4);

    // ---- pass D: emit ----
    unsigned long long ocarry = blk_off, overflow = 0;
    const double samp_val = G / nrem;  // tmp_glob_norm / n_samp
#define FR2_EMIT(VAL, SUB)                             \
    do {                                               \
        if (o < b.out_cap) {                           \
            b.out_val[o] = (VAL);                      \
            b.out_widx[o] = (uint32_t)i;               \
            b.out_sub[o] = (uint32_t)(SUB);            \
        } else {                                       \
            overflow++;                                \
        }                                              \
        o++;                                           \
    } while (0)
    unsigned long long *sm_off = reinterpret_cast<unsigned long long *>(sm.start);  // output offset of every slot
    for (size_t base = lo; base < hi; base += FR2_TILE) {
        const size_t i0 = base + (size_t)tid * FR2_ITEMS;
        uint32_t c4[FR2_ITEMS];
        fr2_ld4(b.kcnt, i0, hi, c4, 0u);
        unsigned tk = 0;
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) tk += (c4[k] >> 30) ? (c4[k] >> 30) : c4[k];
        if (base == lo) FR_TL(b.st, 19);  // emit: loads done
        unsigned ex, tot;
        fr2_scan_u(tk, ex, tot, sm.wcnt[buf]);
        if (base == lo) FR_TL(b.st, 20);
        unsigned long long o = ocarry + ex;
        unsigned need = 0;
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++) {
            const size_t i = i0 + k;
            const uint32_t code = c4[k], mode = code >> 30;
            if (mode) {  // one or two resampled outputs recorded by the count pass
                FR2_EMIT(samp_val, (code >> 16) & 31u);
                if (mode == 2) FR2_EMIT(samp_val, (code >> 21) & 31u);
            } else if (code) {
                sm_off[tid * FR2_ITEMS + k] = o;
                need |= 1u << k;
                o += code;
            }
        }
        if (base == lo) FR_TL(b.st, 21);  // coded outputs written
        unsigned nneed = __popc(need), nex, ntot;
        fr2_scan_u(nneed, nex, ntot, sm.wcnt[buf ^ 1]);
#pragma unroll
        for (int k = 0; k < FR2_ITEMS; k++)
            if ((need >> k) & 1u) sm.list[nex++] = (unsigned short)(tid * FR2_ITEMS + k);
        __syncthreads();
        for (unsigned e = tid; e < ntot; e += FR2_NT) {
            const unsigned slot = sm.list[e];
            const size_t i = base + slot;
            unsigned long long o = sm_off[slot];
            const uint32_t k = b.kcnt[i];
            const double v = b.veff[i];
            const uint32_t nd = b.ndiv[i];
            const double start = b.lb[i], wr = b.wt_remain[i];
            const double lbound = start + wr;
            if (nd > 0) {
                if (b.keep[i]) {
                    const double each = v / nd;
                    for (uint32_t j = 0; j < nd; j++) FR2_EMIT(each, j);
                } else {
                    const double k0 = sg.count_below_d(start);
                    for (uint32_t t = 0; t < k; t++) {
                        double g = sg.point_d(k0 + (double)t);
                        unsigned long long sub = (unsigned long long)((lbound - g) * nd / v);
                        if (sub >= nd) {
                            sub = nd - 1;
                            anomalies++;
                        }
                        FR2_EMIT(samp_val, sub);
                    }
                }
            } else {
                const uint32_t ns = b.nsub[i], kb = b.keep[i];
                double k0 = sg.count_below_d(start);
                double g = sg.point_d(k0);
                double sub_lb = lbound - wr;
                const unsigned long long o_end = o + k;
                prov.visit(i, b.rinv[i], [&](uint32_t j, double wj) -> bool {
                    if (j >= ns) return false;
                    if (((kb >> j) & 1u) && wj != 0) {
                        FR2_EMIT(v * wj, j);
                    } else {
                        sub_lb += v * wj;
                        if (g < sub_lb && wj != 0) {
                            FR2_EMIT(samp_val, j);
                            k0 += 1.0;
                            g = sg.point_d(k0);
                        }
                    }
                    return o < o_end;
                });
            }
        }
        if (base == lo) FR_TL(b.st, 22);  // heavy outputs written (this thread)
        ocarry += tot;
        __syncthreads();
        if (base == lo) FR_TL(b.st, 23);
    }
#undef FR2_EMIT
    {
        double dd[1] = {0.0};
        unsigned long long cc[1] = {overflow + (anomalies << 32)};
        // overflow and anomaly counts are rare: one atomic per CTA that has any
        fr2_sum<1>(dd, cc, sh_sd, sh_sc);
        if (tid == 0 && cc[0]) {
            if (cc[0] & 0xffffffffull) atomicAdd(&b.st->overflow, cc[0] & 0xffffffffull);
            if (cc[0] >> 32) atomicAdd(&b.st->anomalies, cc[0] >> 32);
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        b.st->loc_norm = loc_final;
        b.st->glob_norm = s;
        b.st->n_samp_left = nrem;
        b.st->rounds = rounds;
        b.st->n_kept = kept_total;
        b.st->n_out = tot_out < b.out_cap ? tot_out : b.out_cap;
        b.st->n_in = n;
    }
    FR2_CTA_MARK(b2, 5);
    grid_comb_end(gcb, gcur);
    if (multi) grid.sync();
    FR_STAMP(b.st, 5);
    FR_TL(b.st, 24);
    comm_end(cm, cur);
    (void)wid;
    (void)lane;
}
