// a4/a5: find_preserve + sys_comp on a resident vector; a6: comp_sub with explicit sub-weights.
#include "compress.cuh"

// ---------------------------------------------------------------------------------------------------
// find_preserve (compress_utils.cpp:29-105).  The reference pops a max-heap per rank: an element is preserved
// when |v| >= R / n_left (R = one-norm of what is not yet preserved), and R, n_left are re-synchronised across
// ranks every round.  That is Newton's iteration on the concave function g(t) = R(t) - t n(t) from above; its
// state is a single threshold (preserved = {|v| >= thr}), so the rounds here are read-only passes.  Each pass
// probes FP_PROBES thresholds t_k = (R / n_left) 2^-k at once; the lowest probe that is still at or above its
// own Newton image (t_k >= R_k / n_k, i.e. not below the fixed point) becomes the new threshold, which skips
// several Newton steps per grid barrier.  Keep flags are written by the exact recomputation pass that the
// reference also performs when a round preserves nothing (:78-90).
// Traffic per round: 8 B/element (values only).
// ---------------------------------------------------------------------------------------------------
#define FP_PROBES 4
extern int fr_debug_repeat, fr_last_fast, fr_bracket_on;  // hbpp.cu
int fr_debug_upload_vals(fries_ctx *c, double *d_vals, const double *h_vals, size_t n, bool warmup);
__global__ void __launch_bounds__(FR_COMP_BLOCK)
find_preserve_kernel(const double *__restrict__ vals, size_t n, const unsigned long long *__restrict__ n_ptr,
                     unsigned n_samp_in, uint8_t *__restrict__ keep, double *part_d, unsigned long long *part_c,
                     CompState *st, CommView cm, KeepPred *pred, CandList cand) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_x[12][FR_MAX_RANKS];
    __shared__ double sh_d[6 * 33];
    __shared__ unsigned long long sh_c[6 * 33];
    CommCursor cur = comm_begin(cm);
    const bool multi = cm.n_ranks > 1;
    if (n_ptr) {  // resident pipeline: the element count lives on the device
        unsigned long long dn = *n_ptr;
        if (dn < n) n = (size_t)dn;
    }
    GridRed red{part_d, part_c, 0, (int)gridDim.x, sh_d, sh_c};
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;

    // bracket around the expected fixed point (compress.cuh, bracket_solve)
    __shared__ unsigned long long sh_bc[2];  // one load per CTA, then broadcast
    if (threadIdx.x == 0) {
        sh_bc[0] = pred ? (unsigned long long)__double_as_longlong(__ldcg(&pred->t)) : 0ull;
        sh_bc[1] = pred ? (unsigned long long)__double_as_longlong(__ldcg(&pred->h)) : 0ull;
    }
    __syncthreads();
    const double t_pred = __longlong_as_double((long long)sh_bc[0]), h_pred = __longlong_as_double((long long)sh_bc[1]);
    const bool try_fast = t_pred > 0 && h_pred > 0 && h_pred < 0.25;
    const double t_lo = t_pred * (1.0 - h_pred), t_hi = t_pred * (1.0 + h_pred);
    double s = 0, s_hi = 0;
    unsigned long long c_hi = 0;
    bool appended = false;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double m = fabs(vals[i]);
        s += m;
        if (try_fast && m >= t_lo) {
            if (m >= t_hi) {
                c_hi++;
                s_hi += m;
            } else {
                cand_append(cand, cm, m, 1u);
                appended = true;
            }
        }
    }
    cand_flush(cm, appended);
    unsigned long long dummy = 0;
    {
        double dd[2] = {s, s_hi};
        unsigned long long cc[2] = {c_hi, 0ull};
        grid_reduce_vec<2>(grid, red, dd, cc, sh_d, sh_c);
        s = dd[0];
        s_hi = dd[1];
        c_hi = cc[0];
    }
    double loc = s, R_next = s;
    __shared__ unsigned long long sh_seg[FR_MAX_RANKS];
    const unsigned long long my_cand = try_fast ? bracket_list_len(cand, sh_c) : 0ull;
    bool peers_ok = try_fast && my_cand <= (multi ? (unsigned long long)FR_COMM_XCAP : (unsigned long long)FR_CAND_GCAP);
    if (multi) {  // *global_norm = sum_mpi(loc_one_norm) (:50); the bracket statistics and list lengths ride along
        double pay[5] = {s, s_hi, (double)c_hi, peers_ok ? 1.0 : 0.0, (double)my_cand};
        comm_allgather_v(cm, cur, pay, 5, sh_x);
        double before;
        comm_sum(cm, sh_x[0], R_next, before);
        double gs = 0, gc = 0;
        for (int p = 0; p < cm.n_ranks; p++) {
            gs += sh_x[1][p];
            gc += sh_x[2][p];
            if (sh_x[3][p] == 0.0) peers_ok = false;
        }
        if (threadIdx.x < cm.n_ranks) sh_seg[threadIdx.x] = (unsigned long long)sh_x[4][threadIdx.x];
        s_hi = gs;
        c_hi = (unsigned long long)gc;
        if (try_fast) __threadfence_system();  // acquire side: the peers' candidates in this rank's window
        __syncthreads();
    }
    const double glob_total = R_next;
    unsigned nrem = n_samp_in;
    unsigned long long glob_sampled = 1, kept_total = 0;
    bool recalc = false;
    double R = 0, thr = INFINITY, fresh_loc = 0;
    unsigned rounds = 0;
    unsigned long long n_cand = 0;
    if (try_fast) {
        BracketResult br = bracket_solve(grid, cand, st->gacc, glob_total - s_hi, (long long)n_samp_in - (long long)c_hi, t_lo,
                                         t_hi, sh_d, sh_c, cm, sh_seg, peers_ok);
        n_cand = br.n_cand;
        if (br.valid) {
            // one pass: keep flags of the cut + exact residual norm (:78-90)
            double t = 0;
            unsigned long long kc = 0;
            for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                double m = fabs(vals[i]);
                bool kp = m >= br.x_cut;
                keep[i] = kp ? 1 : 0;
                if (kp)
                    kc++;
                else
                    t += m;
            }
            grid_reduce(grid, red, t, kc);
            kept_total = c_hi + br.kept_cand;
            if (!multi && kc != kept_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->anomalies, 1ull << 32);
            fresh_loc = t;
            thr = br.x_cut;
            nrem = br.nrem;
            R = br.R;
            rounds = br.rounds;
            glob_sampled = 0;  // skip the plain rounds
            if (blockIdx.x == 0 && threadIdx.x == 0) st->fast = 1;
        }
    }
    while (glob_sampled > 0 && rounds < 100000) {  // the bound only guards against a corrupted reduction
        R = R_next;
        double sk[FP_PROBES];
        unsigned long long ck[FP_PROBES];
        double tk[FP_PROBES];
#pragma unroll
        for (int k = 0; k < FP_PROBES; k++) {
            sk[k] = 0;
            ck[k] = 0;
        }
        tk[0] = R / nrem;  // glob_one_norm / (*n_samp - loc_sampled), loc_sampled stale (:59)
#pragma unroll
        for (int k = 1; k < FP_PROBES; k++) tk[k] = tk[k - 1] * 0.5;
        if (R >= 0) {
            for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                double m = fabs(vals[i]);
                if (m < thr && m >= tk[FP_PROBES - 1]) {
                    int kk = FP_PROBES - 1;
#pragma unroll
                    for (int k = FP_PROBES - 2; k >= 0; k--)
                        if (m >= tk[k]) kk = k;
                    // bucket kk = the highest probe the element reaches
#pragma unroll
                    for (int k = 0; k < FP_PROBES; k++)
                        if (k == kk) {
                            sk[k] += m;
                            ck[k]++;
                        }
                }
            }
        }
        grid_reduce_vec<FP_PROBES>(grid, red, sk, ck, sh_d, sh_c);
        // cumulative over the probes: everything at or above t_k
#pragma unroll
        for (int k = 1; k < FP_PROBES; k++) {
            sk[k] += sk[k - 1];
            ck[k] += ck[k - 1];
        }
        double gS[FP_PROBES], locs_after[FP_PROBES];
        unsigned long long gC[FP_PROBES];
        if (multi) {
            double pay[2 * FP_PROBES + 1];
#pragma unroll
            for (int k = 0; k < FP_PROBES; k++) {
                pay[k] = sk[k];
                pay[FP_PROBES + k] = (double)ck[k];
            }
            pay[2 * FP_PROBES] = loc;
            comm_allgather_v(cm, cur, pay, 2 * FP_PROBES + 1, sh_x);
#pragma unroll
            for (int k = 0; k < FP_PROBES; k++) {
                double a = 0, c = 0, l = 0;
                for (int p = 0; p < cm.n_ranks; p++) {
                    a += sh_x[k][p];
                    c += sh_x[FP_PROBES + k][p];
                    l += sh_x[2 * FP_PROBES][p] - sh_x[k][p];  // sum_mpi(loc_one_norm) after this probe's subtraction
                }
                gS[k] = a;
                gC[k] = (unsigned long long)c;
                locs_after[k] = l;
            }
        } else {
#pragma unroll
            for (int k = 0; k < FP_PROBES; k++) {
                gS[k] = sk[k];
                gC[k] = ck[k];
                locs_after[k] = loc - sk[k];
            }
        }
        int ks = 0;  // probe 0 is the reference's own test
#pragma unroll
        for (int k = 1; k < FP_PROBES; k++) {
            if (gC[k] < nrem) {
                double Rk = R - gS[k], nk = (double)(nrem - (unsigned)gC[k]);
                if (tk[k] >= Rk / nk) ks = k;
            }
        }
        if (gC[ks] > 0) thr = tk[ks];
        loc -= sk[ks];
        R_next = locs_after[ks];
        glob_sampled = gC[ks];
        nrem -= (unsigned)gC[ks];
        kept_total += gC[ks];
        rounds++;
        if (glob_sampled == 0 && !recalc) {
            // exact recomputation of the residual norm (:78-90) + keep flags of the current threshold
            double t = 0;
            for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                double m = fabs(vals[i]);
                bool kp = m >= thr;
                keep[i] = kp ? 1 : 0;
                if (!kp) t += m;
            }
            grid_reduce(grid, red, t, dummy);
            loc = t;
            R_next = t;
            fresh_loc = t;
            if (multi) {
                comm_allgather_v(cm, cur, &t, 1, sh_x);
                double before;
                comm_sum(cm, sh_x[0], R_next, before);
            }
            glob_sampled = 1;
            recalc = true;
        } else {
            recalc = false;
        }
    }
    // the loop ends with a round that preserved nothing right after an exact recomputation, so the flags and
    // the residual norm of that recomputation are final (the reference sums once more, :98-102: same quantity)
    double loc_final = 0;
    if (pred && blockIdx.x == 0 && threadIdx.x == 0)
        keep_pred_update(pred, try_fast ? t_pred : 0.0, h_pred, nrem > 0 ? R / nrem : 0.0, n_cand);
    if (R < 1e-9) {
        nrem = 0;
    } else {
        loc_final = fresh_loc;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->loc_norm = loc_final;
        st->glob_norm = glob_total;
        st->n_samp_left = nrem;
        st->rounds = rounds;
        st->n_kept = kept_total;
        st->n_in = n;
    }
    if (cm.n_ranks > 1) grid.sync();  // the epoch is stored once every CTA is past its last exchange; a single rank needs no barrier at the end
    comm_end(cm, cur);
}

// ---------------------------------------------------------------------------------------------------
// sys_comp (compress_utils.cpp:278-327): chunked device-wide exclusive scan of |v| over the
// non-preserved elements in storage order; element i keeps +-G/n iff a grid point rn0 + k*unit lies in
// (lb_i, lb_i + |v_i|).  lbound0 / glob / n_samp describe this shard's place in the global
// systematic grid (seed_sys :107-127).  Traffic: 2 reads + 1 write of 8 B per element.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FR_COMP_BLOCK)
sys_comp_kernel(double *__restrict__ vals, size_t n, const unsigned long long *__restrict__ n_ptr,
                uint8_t *__restrict__ keep, const double *__restrict__ in_r4, double lbound0_host, double glob_host,
                long long n_samp_host, double rn_uniform, double *part_d, unsigned long long *part_c,
                CompState *out_st, CommView cm) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_x0[FR_MAX_RANKS], sh_x1[FR_MAX_RANKS];
    __shared__ unsigned long long sh_xc[FR_MAX_RANKS];
    CommCursor cur = comm_begin(cm);
    if (n_ptr) {
        unsigned long long dn = *n_ptr;
        if (dn < n) n = (size_t)dn;
    }
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[68];
    __shared__ unsigned long long sh_sc[68];
    GridRed red{part_d, part_c, 0, (int)gridDim.x, sh_d, sh_c};
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;

    // single shard: take norm and budget from the preceding find_preserve on the device
    double G = glob_host, lbound0 = lbound0_host;
    unsigned n_samp = (unsigned)n_samp_host;
    if (n_samp_host < 0) {
        // resident pipeline: this rank's residual norm and the global budget come from find_preserve; the
        // loc_norms are all-gathered here (frisys_mol.cpp:532) and summed in rank order (:292-295, :114-121)
        G = in_r4[0];
        n_samp = (unsigned)in_r4[2];
        lbound0 = 0;
        if (cm.n_ranks > 1) {
            comm_allgather(cm, cur, in_r4[0], 0.0, 0ull, sh_x0, sh_x1, sh_xc);
            comm_sum(cm, sh_x0, G, lbound0);
        }
    }
    SysGrid sg;
    if (n_samp > 0) {
        // seed_sys :107-127: this shard's first grid point is global grid index j0 (or j0 + 1)
        sg.unit = G / n_samp;
        long long j0 = (long long)(int)(lbound0 * n_samp / G);
        double r = rn_uniform * sg.unit;
        r += sg.unit * (int)(lbound0 * n_samp / G);
        if (r < lbound0) {
            r += sg.unit;
            j0++;
        }
        sg.rn0 = r;
        sg.inv = 1.0 / sg.unit;
        sg.n = (long long)n_samp - j0;
    } else {
        sg.rn0 = INFINITY;
        sg.unit = INFINITY;
        sg.inv = 0;
        sg.n = 0;
    }
    double cs = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
        if (!keep[i]) cs += fabs(vals[i]);
    cs = block_sum(cs, sh_d);
    double blk_lb, tot_lb;
    unsigned long long e0, e1;
    grid_excl_scan(grid, red, cs, 0ull, blk_lb, e0, tot_lb, e1, sh_sd, sh_sc);

    double carry = lbound0 + blk_lb;
    double new_norm = 0;
    unsigned long long n_samples = 0;
    for (size_t base = lo; base < hi; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        bool act = i < hi;
        double v = act ? vals[i] : 0.0;
        bool kp = act ? keep[i] != 0 : true;
        double m = kp ? 0.0 : fabs(v);
        double ex, tot;
        unsigned long long ec, tc;
        block_excl_scan(m, 0ull, ex, ec, tot, tc, sh_sd, sh_sc);
        if (act) {
            if (kp) {
                new_norm += fabs(v);
                keep[i] = 0;
            } else if (v != 0) {
                double start = carry + ex;
                double lbound = start + m;
                double g = sg.point(sg.count_below(start));
                if (g < lbound) {
                    double nv = G / n_samp;
                    vals[i] = v > 0 ? nv : -nv;
                    new_norm += nv;
                    n_samples++;
                } else {
                    vals[i] = 0;
                    keep[i] = 1;
                }
            }
        }
        carry += tot;
    }
    grid_reduce(grid, red, new_norm, n_samples);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out_st->new_norm = new_norm;
        out_st->n_out = n_samples;
    }
    if (cm.n_ranks > 1) grid.sync();  // the epoch is stored once every CTA is past its last exchange; a single rank needs no barrier at the end
    comm_end(cm, cur);
}

// ---------------------------------------------------------------------------------------------------
// comp_sub with materialised sub-weights (the reference's calling convention)
// ---------------------------------------------------------------------------------------------------
struct MatProvider {
    const double *values;
    const uint32_t *ndiv;
    const double *subw;
    const uint16_t *sub_sizes;
    size_t n, n_sub;
    __device__ size_t count() const { return n; }
    __device__ void prep(size_t i, double &v, uint32_t &nd, uint32_t &ns, double &rinv, double &wmax) const {
        v = values[i];
        nd = ndiv[i];
        ns = sub_sizes ? sub_sizes[i] : (uint32_t)n_sub;
        rinv = 1.0;
        wmax = 0;
        if (nd == 0)
            for (uint32_t j = 0; j < ns && j < n_sub; j++) wmax = fmax(wmax, subw[i * n_sub + j]);
    }
    template <class F>
    __device__ void visit(size_t i, double, F &&f) const {
        for (uint32_t j = 0; j < n_sub; j++)
            if (!fr_emit(f, j, subw[i * n_sub + j])) return;
    }
};

__global__ void __launch_bounds__(FR_COMP_BLOCK)
comp_sub_mat_kernel(MatProvider prov, CompSubBufs bufs, unsigned n_samp, double rn) {
    comp_sub_engine(prov, bufs, n_samp, rn);
}

// ---------------------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------------------
static int coop_launch(fries_ctx *c, const void *kernel, int grid, void **args) {
    CUDA_TRY(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(FR_COMP_BLOCK), args, 0, c->stream));
    c->launch_count++;
    return FRIES_OK;
}

struct RedScratch {
    double *pd;
    unsigned long long *pc;
    CompState *st;
};
// carve [2][grid][FR_RED_STRIDE] doubles + the same in u64 + 2 CompState out of the context scratch, after `skip` bytes
static int red_scratch(fries_ctx *c, int grid, size_t skip, RedScratch &r) {
    size_t per = (size_t)grid * 2 * FR_RED_STRIDE * 8;
    size_t need = skip + 2 * per + 2 * sizeof(CompState) + 512;
    FRIES_TRY(c->ensure_scratch(need));
    char *p = (char *)c->d_scratch + ((skip + 255) & ~(size_t)255);
    r.pd = (double *)p;
    r.pc = (unsigned long long *)(p + per);
    r.st = (CompState *)(p + 2 * per);
    return FRIES_OK;
}

extern "C" int fries_find_preserve_dev(fries_ctx *c, const double *d_values, size_t count, unsigned n_samp,
                                       uint8_t *d_keep, double *d_result4);
extern "C" int fries_sys_comp_dev(fries_ctx *c, double *d_values, size_t count, const double *d_result4,
                                  uint8_t *d_keep, double rand_num, double *d_new_norm);

__global__ void state_to_result4(const CompState *st, double *r4) {
    r4[0] = st->loc_norm;
    r4[1] = st->glob_norm;
    r4[2] = (double)st->n_samp_left;
    r4[3] = (double)st->n_kept;
}

int fries_find_preserve_launch(fries_ctx *c, const double *d_values, size_t count, const unsigned long long *d_n,
                               unsigned n_samp, uint8_t *d_keep, CompState *d_st, double *pd, unsigned long long *pc,
                               int grid, const fries_comm *comm, KeepPred *pred, double *cand_x, uint32_t *cand_m) {
    CommView cmv = fries_comm_view(comm);
    if (grid <= 0) grid = c->coop_grid((const void *)find_preserve_kernel, FR_COMP_BLOCK, 0);
    CandList cand{cand_x, cand_m, &d_st->n_cand};
    if (!fr_bracket_on || !cand_x || !cand_m) pred = nullptr;
    void *args[] = {(void *)&d_values, (void *)&count, (void *)&d_n, (void *)&n_samp, (void *)&d_keep, (void *)&pd,
                    (void *)&pc, (void *)&d_st, (void *)&cmv, (void *)&pred, (void *)&cand};
    ProfScope ps(c, "find_preserve");
    return coop_launch(c, (const void *)find_preserve_kernel, grid, args);
}

int fries_sys_comp_launch(fries_ctx *c, double *d_values, size_t count, const unsigned long long *d_n, uint8_t *d_keep,
                          const double *d_in, double lbound0, double glob, long long n_samp, double rn, CompState *d_out,
                          double *pd, unsigned long long *pc, int grid, const fries_comm *comm) {
    CommView cmv = fries_comm_view(comm);
    if (grid <= 0) grid = c->coop_grid((const void *)sys_comp_kernel, FR_COMP_BLOCK, 0);
    void *args[] = {(void *)&d_values, (void *)&count, (void *)&d_n,  (void *)&d_keep, (void *)&d_in, (void *)&lbound0,
                    (void *)&glob,     (void *)&n_samp, (void *)&rn,  (void *)&pd,     (void *)&pc,   (void *)&d_out,
                    (void *)&cmv};
    ProfScope ps(c, "sys_comp");
    return coop_launch(c, (const void *)sys_comp_kernel, grid, args);
}

extern "C" int fries_find_preserve(fries_ctx *c, const double *h_values, size_t count, unsigned *n_samp,
                                   double *glob_norm, uint8_t *h_keep, double *loc_norm) {
    FRIES_REQUIRE(c && n_samp && glob_norm && loc_norm && (count == 0 || (h_values && h_keep)),
                  "fries_find_preserve: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = c->coop_grid((const void *)find_preserve_kernel, FR_COMP_BLOCK, 0);
    DevBuf<double> vals;
    DevBuf<uint8_t> keep;
    FRIES_TRY(vals.alloc(count));
    FRIES_TRY(keep.alloc(count));
    RedScratch r;
    FRIES_TRY(red_scratch(c, grid, 0, r));
    CUDA_TRY(cudaMemcpyAsync(vals.p, h_values, count * 8, cudaMemcpyHostToDevice, c->stream));
    // the bracketed threshold solve needs the fixed point of a previous run: only with fries_debug_set_repeat(> 1)
    DevBuf<double> cand_x;
    DevBuf<uint32_t> cand_m;
    DevBuf<KeepPred> pred;
    if (fr_debug_repeat > 1) {
        FRIES_TRY(cand_x.alloc(FR_CAND_GCAP));
        FRIES_TRY(cand_m.alloc(FR_CAND_GCAP));
        FRIES_TRY(pred.alloc(1));
        CUDA_TRY(cudaMemsetAsync(pred.p, 0, sizeof(KeepPred), c->stream));
    }
    for (int rep = 0; rep < fr_debug_repeat; rep++) {
        FRIES_TRY(fr_debug_upload_vals(c, vals.p, h_values, count, rep + 1 < fr_debug_repeat));
        CUDA_TRY(cudaMemsetAsync(r.st, 0, sizeof(CompState), c->stream));
        FRIES_TRY(fries_find_preserve_launch(c, vals.p, count, nullptr, *n_samp, keep.p, r.st, r.pd, r.pc, grid, nullptr,
                                             pred.p, cand_x.p, cand_m.p));
    }
    CompState st;
    CUDA_TRY(cudaMemcpyAsync(&st, r.st, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    if (count) CUDA_TRY(cudaMemcpyAsync(h_keep, keep.p, count, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    fr_last_fast = (int)st.fast;
    *n_samp = st.n_samp_left;
    *glob_norm = st.glob_norm;
    *loc_norm = st.loc_norm;
    return FRIES_OK;
}

extern "C" int fries_sys_comp(fries_ctx *c, double *h_values, size_t count, double *loc_norms, int n_ranks, int rank,
                              unsigned n_samp, uint8_t *h_keep, double rand_num) {
    FRIES_REQUIRE(c && loc_norms && (count == 0 || (h_values && h_keep)), "fries_sys_comp: NULL argument");
    FRIES_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "fries_sys_comp: bad rank %d of %d", rank, n_ranks);
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = c->coop_grid((const void *)sys_comp_kernel, FR_COMP_BLOCK, 0);
    DevBuf<double> vals;
    DevBuf<uint8_t> keep;
    FRIES_TRY(vals.alloc(count));
    FRIES_TRY(keep.alloc(count));
    RedScratch r;
    FRIES_TRY(red_scratch(c, grid, 0, r));
    // seed_sys inputs, summed in rank order as the reference does (compress_utils.cpp:292-295,114-121)
    double glob = 0, lbound0 = 0;
    for (int p = 0; p < n_ranks; p++) glob += loc_norms[p];
    for (int p = 0; p < rank; p++) lbound0 += loc_norms[p];
    CUDA_TRY(cudaMemcpyAsync(vals.p, h_values, count * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(keep.p, h_keep, count, cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(fries_sys_comp_launch(c, vals.p, count, nullptr, keep.p, nullptr, lbound0, glob, (long long)n_samp, rand_num,
                                    r.st + 1, r.pd, r.pc, grid, nullptr));
    CompState st;
    CUDA_TRY(cudaMemcpyAsync(&st, r.st + 1, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    if (count) {
        CUDA_TRY(cudaMemcpyAsync(h_values, vals.p, count * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(h_keep, keep.p, count, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    loc_norms[rank] = st.new_norm;
    return FRIES_OK;
}

extern "C" int fries_find_preserve_dev(fries_ctx *c, const double *d_values, size_t count, unsigned n_samp,
                                       uint8_t *d_keep, double *d_result4) {
    FRIES_REQUIRE(c && d_result4 && (count == 0 || (d_values && d_keep)), "fries_find_preserve_dev: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = c->coop_grid((const void *)find_preserve_kernel, FR_COMP_BLOCK, 0);
    RedScratch r;
    FRIES_TRY(red_scratch(c, grid, 0, r));
    CUDA_TRY(cudaMemsetAsync(r.st, 0, sizeof(CompState), c->stream));
    FRIES_TRY(fries_find_preserve_launch(c, d_values, count, nullptr, n_samp, d_keep, r.st, r.pd, r.pc, grid, c->comm));
    state_to_result4<<<1, 1, 0, c->stream>>>(r.st, d_result4);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

extern "C" int fries_sys_comp_dev(fries_ctx *c, double *d_values, size_t count, const double *d_result4,
                                  uint8_t *d_keep, double rand_num, double *d_new_norm) {
    FRIES_REQUIRE(c && d_result4 && (count == 0 || (d_values && d_keep)), "fries_sys_comp_dev: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = c->coop_grid((const void *)sys_comp_kernel, FR_COMP_BLOCK, 0);
    RedScratch r;
    FRIES_TRY(red_scratch(c, grid, 0, r));
    FRIES_TRY(fries_sys_comp_launch(c, d_values, count, nullptr, d_keep, d_result4, 0.0, 0.0, -1LL, rand_num, r.st + 1, r.pd,
                                    r.pc, grid, c->comm));
    if (d_new_norm) {
        CUDA_TRY(cudaMemcpyAsync(d_new_norm, &r.st[1].new_norm, 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    return FRIES_OK;
}

extern "C" int fries_comp_sub(fries_ctx *c, const double *h_values, size_t count, const uint32_t *h_ndiv,
                              const double *h_sub_weights, size_t n_sub, const uint16_t *h_sub_sizes, unsigned n_samp,
                              double rand_num, double *h_new_vals, uint64_t *h_new_idx, size_t out_cap, size_t *n_out,
                              unsigned *n_samp_left, double *loc_norm) {
    FRIES_REQUIRE(c && n_out && (count == 0 || (h_values && h_ndiv && h_sub_weights)), "fries_comp_sub: NULL argument");
    FRIES_REQUIRE(n_sub >= 1 && n_sub <= FRIES_MAX_SUB, "fries_comp_sub: n_sub %zu not in 1..%d", n_sub, FRIES_MAX_SUB);
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = c->coop_grid((const void *)comp_sub_mat_kernel, FR_COMP_BLOCK, 0);
    size_t cn = count ? count : 1;
    DevBuf<double> values, subw, veff, wtr, lb, oval, rinv;
    DevBuf<uint32_t> ndiv, keep, kcnt, owidx, osub;
    DevBuf<uint16_t> ssz;
    DevBuf<uint8_t> nsub;
    FRIES_TRY(values.alloc(cn));
    FRIES_TRY(subw.alloc(cn * n_sub));
    FRIES_TRY(veff.alloc(cn));
    FRIES_TRY(wtr.alloc(cn));
    FRIES_TRY(lb.alloc(cn));
    FRIES_TRY(rinv.alloc(cn));
    FRIES_TRY(ndiv.alloc(cn));
    FRIES_TRY(keep.alloc(cn));
    FRIES_TRY(kcnt.alloc(cn));
    FRIES_TRY(nsub.alloc(cn));
    FRIES_TRY(oval.alloc(out_cap));
    FRIES_TRY(owidx.alloc(out_cap));
    FRIES_TRY(osub.alloc(out_cap));
    if (h_sub_sizes) FRIES_TRY(ssz.alloc(cn));
    RedScratch r;
    FRIES_TRY(red_scratch(c, grid, 0, r));
    if (count) {
        CUDA_TRY(cudaMemcpyAsync(values.p, h_values, count * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(ndiv.p, h_ndiv, count * 4, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(subw.p, h_sub_weights, count * n_sub * 8, cudaMemcpyHostToDevice, c->stream));
        if (h_sub_sizes) CUDA_TRY(cudaMemcpyAsync(ssz.p, h_sub_sizes, count * 2, cudaMemcpyHostToDevice, c->stream));
    }
    CUDA_TRY(cudaMemsetAsync(r.st, 0, sizeof(CompState), c->stream));
    MatProvider prov{values.p, ndiv.p, subw.p, h_sub_sizes ? ssz.p : nullptr, count, n_sub};
    // the bracketed threshold solve needs the fixed point of a previous run: only with fries_debug_set_repeat(> 1)
    DevBuf<double> cand_x;
    DevBuf<uint32_t> cand_m;
    DevBuf<KeepPred> pred;
    if (fr_debug_repeat > 1) {
        FRIES_TRY(cand_x.alloc(FR_CAND_GCAP));
        FRIES_TRY(cand_m.alloc(FR_CAND_GCAP));
        FRIES_TRY(pred.alloc(1));
        CUDA_TRY(cudaMemsetAsync(pred.p, 0, sizeof(KeepPred), c->stream));
    }
    CompSubBufs bufs{veff.p, wtr.p, lb.p, rinv.p, ndiv.p, keep.p, kcnt.p, nsub.p, oval.p, owidx.p, osub.p,
                     (unsigned long long)out_cap, r.pd, r.pc, r.st, fries_comm_view(nullptr),
                     pred.p, CandList{cand_x.p, cand_m.p, &r.st->n_cand}};
    // ndiv is both provider input and engine state: give the engine its own copy
    DevBuf<uint32_t> ndiv_state;
    FRIES_TRY(ndiv_state.alloc(cn));
    bufs.ndiv = ndiv_state.p;
    void *args[] = {(void *)&prov, (void *)&bufs, (void *)&n_samp, (void *)&rand_num};
    for (int rep = 0; rep < fr_debug_repeat; rep++) {
        if (count) FRIES_TRY(fr_debug_upload_vals(c, values.p, h_values, count, rep + 1 < fr_debug_repeat));
        if (rep) CUDA_TRY(cudaMemsetAsync(r.st, 0, sizeof(CompState), c->stream));
        ProfScope ps(c, "comp_sub");
        FRIES_TRY(coop_launch(c, (const void *)comp_sub_mat_kernel, grid, args));
    }
    CompState st;
    CUDA_TRY(cudaMemcpyAsync(&st, r.st, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    fr_last_fast = (int)st.fast;
    if (st.overflow) {
        fries_set_error("fries_comp_sub: %llu outputs did not fit out_cap %zu", st.overflow, out_cap);
        return FRIES_ERR_CAPACITY;
    }
    size_t no = (size_t)st.n_out;
    std::vector<uint32_t> widx(no ? no : 1);
    std::vector<uint32_t> sub(no ? no : 1);
    if (no) {
        CUDA_TRY(cudaMemcpy(h_new_vals, oval.p, no * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(widx.data(), owidx.p, no * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(sub.data(), osub.p, no * 4, cudaMemcpyDeviceToHost));
    }
    for (size_t i = 0; i < no; i++) {
        h_new_idx[2 * i] = widx[i];
        h_new_idx[2 * i + 1] = sub[i];
    }
    *n_out = no;
    if (n_samp_left) *n_samp_left = st.n_samp_left;
    if (loc_norm) *loc_norm = st.loc_norm;
    return FRIES_OK;
}
