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
};

struct GridRed {
    double *pd;               // [2][nb]
    unsigned long long *pc;   // [2][nb]
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
            __stcg(&r.pd[((size_t)r.parity * r.nb + blockIdx.x) * K + k], d[k]);
            __stcg(&r.pc[((size_t)r.parity * r.nb + blockIdx.x) * K + k], c[k]);
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
            d[k] += __ldcg(&r.pd[((size_t)r.parity * r.nb + i) * K + k]);
            c[k] += __ldcg(&r.pc[((size_t)r.parity * r.nb + i) * K + k]);
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
        __stcg(&r.pd[r.parity * r.nb + blockIdx.x], d);
        __stcg(&r.pc[r.parity * r.nb + blockIdx.x], c);
    }
    grid.sync();
    double dv[2] = {0, 0};
    unsigned long long cv[2] = {0, 0};
    for (int i = threadIdx.x; i < r.nb; i += blockDim.x) {
        double a = __ldcg(&r.pd[r.parity * r.nb + i]);
        unsigned long long b = __ldcg(&r.pc[r.parity * r.nb + i]);
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
};

// The hierarchical compression engine.  Provider P supplies
//   size_t count();                                            number of inputs
//   void prep(size_t i, double &v, uint32_t &ndiv, uint32_t &nsub, double &rinv);  effective weight, row shape,
//                                                              normalisation factor of the row (stored per input)
//   void visit(size_t i, double rinv, F f);                    calls f(j, w_j) for the nsub sub-weights in order --
//                                                              rows are streamed, never materialised
// Restates comp_sub (compress_utils.cpp:797-820): find_keep_sub :130-276 then sys_sub :702-794.
template <class P>
__device__ void comp_sub_engine(P &prov, const CompSubBufs &b, unsigned n_samp_in, double rn_uniform) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[68];
    __shared__ unsigned long long sh_sc[68];
    GridRed red{b.part_d, b.part_c, 0, (int)gridDim.x, sh_d, sh_c};

    const size_t n = prov.count();
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;

    __shared__ double sh_x0[FR_MAX_RANKS], sh_x1[FR_MAX_RANKS];
    __shared__ unsigned long long sh_xc[FR_MAX_RANKS];
    const CommView &cm = b.cm;
    CommCursor cur = comm_begin(cm);
    const bool multi = cm.n_ranks > 1;

    // ---- phase 0: effective weights (find_keep_sub :134-137) ----
    double s = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double v, rinv = 1.0;
        uint32_t nd, ns;
        prov.prep(i, v, nd, ns, rinv);
        b.rinv[i] = rinv;
        b.veff[i] = v;
        b.wt_remain[i] = v;
        b.ndiv[i] = nd;
        b.nsub[i] = (uint8_t)ns;
        b.keep[i] = 0;
        s += v;
    }
    unsigned long long dummy = 0;
    grid_reduce(grid, red, s, dummy);
    double loc = s;          // this rank's loc_one_norm
    double R_next = s;       // sum_mpi(loc_one_norm) for the coming round
    if (multi) {
        double before;
        comm_allgather(cm, cur, s, 0.0, 0ull, sh_x0, sh_x1, sh_xc);
        comm_sum(cm, sh_x0, R_next, before);
        s = R_next;          // global one-norm (reported)
    }

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
    if (R / nrem < 1e-8) {
        nrem = 0;
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
        sg.n = (long long)nrem - j0;
    } else {
        sg.rn0 = INFINITY;
        sg.unit = INFINITY;
        sg.n = 0;
    }

    // CTA boundaries on the resampling line from the chunk sums of the residual weights
    double blk_lb, tot_lb;
    unsigned long long e0, e1;
    grid_excl_scan(grid, red, cs, 0ull, blk_lb, e0, tot_lb, e1, sh_sd, sh_sc);

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
            uint32_t k = 0;
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
                        prov.visit(i, b.rinv[i], [&](uint32_t j, double wj) {
                            if (j >= ns) return;
                            if (((kb >> j) & 1u) && wj != 0) {
                                k++;
                            } else {
                                sub_lb += v * wj;
                                if (g < sub_lb && wj != 0) {
                                    k++;
                                    k0++;
                                    g = sg.point(k0);
                                    if (g < sub_lb) anomalies++;
                                }
                            }
                        });
                    }
                }
            }
            b.kcnt[i] = k;
            my_out += k;
        }
        carry += tot;
    }
    my_out = block_sum_u64(my_out, sh_c);
    double d0, d1;
    unsigned long long blk_off, tot_out;
    grid_excl_scan(grid, red, 0.0, my_out, d0, blk_off, d1, tot_out, sh_sd, sh_sc);

    // pass 3: emit
    unsigned long long ocarry = blk_off;
    unsigned long long overflow = 0;
    const double samp_val = G / nrem;  // tmp_glob_norm / n_samp
    for (size_t base = lo; base < hi; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        bool act = i < hi;
        uint32_t k = act ? b.kcnt[i] : 0;
        double ex, tot;
        unsigned long long ec, tc;
        block_excl_scan(0.0, (unsigned long long)k, ex, ec, tot, tc, sh_sd, sh_sc);
        if (act && k > 0) {
            unsigned long long o = ocarry + ec;
            double v = b.veff[i];
            uint32_t nd = b.ndiv[i];
            double start = b.lb[i];
            double wr = b.wt_remain[i];
            double lbound = start + wr;
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
                prov.visit(i, b.rinv[i], [&](uint32_t j, double wj) {
                    if (j >= ns) return;
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
    grid.sync();
    comm_end(cm, cur);
}
