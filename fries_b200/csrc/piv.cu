// Pivotal compression family on the device (FRIES/compress_utils.cpp:354-681): see piv.cuh for
// the parallel restatement.  Persistent cooperative kernels in the shape of the other compressions (compress.cu):
// each CTA owns a contiguous chunk, a chunked device-wide prefix sum of the non-preserved magnitudes, grid barriers
// between the phases.
#include "piv.cuh"
#include "compress.cuh"
#include "vec.cuh"

struct PivResult {
    double new_norm;               // one-norm after sampling (preserved + drawn)
    double unit;                   // seg_norm / n_samp
    unsigned long long n_drawn;    // elements that carry a sample
    unsigned long long anomalies;  // units whose sample could not be placed (FP ties; 0 in practice)
    unsigned n_units;              // sampling units processed = half the draws consumed
    unsigned n_crossed;
    unsigned n_loc;                // adjust_probs: budget after the adjustment
    unsigned adjusted;             // adjust_probs: 1 if an element was too big and the walk ran
    double adj_norm;               // adjust_probs: norm to hand to the sampler
};

struct PivBufs {
    double *E;          // [n] inclusive prefix sums of the non-preserved magnitudes
    uint32_t *cross;    // [n_samp] straddling element of each unit
    uint32_t *sample;   // [n_samp]
    uint32_t *carry;    // [n_samp]
};

// d_par (optional, resident pipeline): {seg_norm, n_samp} produced on the device by the preceding kernel
__global__ void __launch_bounds__(FR_COMP_BLOCK)
piv_samp_kernel(double *__restrict__ vals, size_t n, uint8_t *__restrict__ keep, double seg_norm, unsigned n_samp,
                const double *__restrict__ d_par, const uint32_t *__restrict__ draws, PivBufs b, double *part_d,
                unsigned long long *part_c, PivResult *res) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[68];
    __shared__ unsigned long long sh_sc[68];
    GridRed red{part_d, part_c, 0, (int)gridDim.x, sh_d, sh_c};
    if (d_par) {
        seg_norm = d_par[0];
        n_samp = (unsigned)d_par[1];
    }
    const PivGrid g = piv_grid(seg_norm, n_samp);
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (size_t)gridDim.x * blockDim.x;
    unsigned n_units = 0, n_crossed = 0;
    unsigned long long anomalies = 0;

    if (n_samp > 0 && n > 0) {
        // ---- prefix sums ------------------------------------------------------------------------------------------
        for (size_t k = gtid; k < n_samp; k += gsize) b.cross[k] = PIV_NONE;
        double cs = 0;
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
            if (!keep[i]) cs += fabs(vals[i]);
        cs = block_sum(cs, sh_d);
        double blk_lb, tot;
        unsigned long long e0, e1;
        grid_excl_scan(grid, red, cs, 0ull, blk_lb, e0, tot, e1, sh_sd, sh_sc);
        double carry = blk_lb;
        for (size_t base = lo; base < hi; base += blockDim.x) {
            size_t i = base + threadIdx.x;
            bool act = i < hi;
            double m = (act && !keep[i]) ? fabs(vals[i]) : 0.0;
            double ex, t;
            unsigned long long ec, tc;
            block_excl_scan(m, 0ull, ex, ec, t, tc, sh_sd, sh_sc);
            if (act) b.E[i] = carry + ex + m;
            carry += t;
        }
        grid.sync();
        // ---- straddling elements ------------------------------------------------------------------------------------
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            double s = i ? b.E[i - 1] : 0.0, e = b.E[i];
            piv_mark_cross(g, s, e, (uint32_t)i, [&](uint32_t k, uint32_t el) { b.cross[k] = el; });
        }
        grid.sync();
        n_crossed = g.borders_le(__ldcg(&b.E[n - 1]));
        auto loadE = [&](uint32_t j) { return __ldcg(&b.E[j]); };
        // a border that no interval (E[i-1], E[i]] claimed (prefix sums that are not monotone to the last bit): search
        auto loadC = [&](uint32_t k) {
            uint32_t c = __ldcg(&b.cross[k]);
            if (c == PIV_NONE) c = piv_search(loadE, 0u, (uint32_t)(n - 1), g.border((uint64_t)k + 1));
            return c;
        };
        uint32_t last_cross = n_crossed ? loadC(n_crossed - 1) : 0;
        n_units = piv_n_units(g, n_crossed, last_cross, n);
        // ---- one thread per unit: candidate, border decision ------------------------------------------------------------
        for (size_t k = gtid; k < n_units; k += gsize) {
            double r1 = draws[2 * k] * (1.0 / 4294967296.0), r2 = draws[2 * k + 1] * (1.0 / 4294967296.0);
            uint32_t s, c;
            piv_unit(g, (uint32_t)k, n_crossed, n, loadE, loadC, r1, r2, s, c);
            b.sample[k] = s;
            b.carry[k] = c;
        }
        grid.sync();
        // ---- resolve "the carried element", flag the samples ----------------------------------------------------------
        for (size_t k = gtid; k < n_units; k += gsize) {
            uint32_t s = b.sample[k];
            if (s == PIV_CARRIED) s = piv_resolve((uint32_t)k, [&](uint32_t j) { return __ldcg(&b.carry[j]); });
            if (s < n && keep[s] != 1) keep[s] = 2;
            else anomalies++;
        }
        grid.sync();
    }
    // ---- final values and flags ---------------------------------------------------------------------------------------
    double new_norm = 0;
    unsigned long long n_drawn = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double v = vals[i];
        uint8_t f = keep[i];
        if (f == 2) n_drawn++;
        piv_finish(g.unit, n_samp, v, f);
        vals[i] = v;
        keep[i] = f;
        new_norm += fabs(v);
    }
    double dd[2] = {new_norm, 0.0};
    unsigned long long cc[2] = {n_drawn, anomalies};
    grid_reduce_vec<2>(grid, red, dd, cc, sh_sd, sh_sc);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        res->new_norm = dd[0];
        res->unit = g.unit;
        res->n_drawn = cc[0];
        res->anomalies = cc[1];
        res->n_units = n_units;
        res->n_crossed = n_crossed;
    }
}

// adjust_probs (compress_utils.cpp:617-681).  d_in (optional): {n_loc, exp_loc, n_tot, tot_norm} from the device.
// Writes res->adj_norm / n_loc / adjusted and, when d_par_out is given, {adj_norm, n_loc} for the sampler.
__global__ void __launch_bounds__(FR_COMP_BLOCK)
piv_adjust_kernel(double *__restrict__ vals, size_t n, uint8_t *__restrict__ keep, unsigned n_loc, double exp_loc,
                  unsigned n_tot, double tot_norm, double *part_d, unsigned long long *part_c, PivResult *res,
                  double *d_par_out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[68];
    __shared__ unsigned long long sh_sc[68];
    GridRed red{part_d, part_c, 0, (int)gridDim.x, sh_d, sh_c};
    const PivAdjust a = piv_adjust_setup(n_loc, exp_loc, n_tot, tot_norm);
    const double thresh = a.loc_norm / ceil(exp_loc);
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;

    double sg = 0;
    unsigned long long big = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        if (keep[i]) continue;
        double m = fabs(vals[i]), dg;
        unsigned long long dk;
        if (m >= thresh) big++;
        a.delta(m, dg, dk);
        sg += dg;
    }
    double zero = 0;
    grid_reduce(grid, red, zero, big);
    if (big == 0) {  // uniform over the grid
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            res->adj_norm = a.loc_norm;
            res->n_loc = n_loc;
            res->adjusted = 0;
            if (d_par_out) {
                d_par_out[0] = a.loc_norm;
                d_par_out[1] = (double)n_loc;
            }
        }
        return;
    }
    sg = block_sum(sg, sh_d);
    double blk_g, tot_g;
    unsigned long long e0, e1;
    grid_excl_scan(grid, red, sg, 0ull, blk_g, e0, tot_g, e1, sh_sd, sh_sc);
    double carry = a.g0() + blk_g;
    unsigned long long made_exact = 0;
    for (size_t base = lo; base < hi; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        bool act = i < hi && !keep[i];
        double v = act ? vals[i] : 0.0, dg = 0;
        unsigned long long dk = 0;
        if (act) a.delta(fabs(v), dg, dk);
        double ex, t;
        unsigned long long ec, tc;
        block_excl_scan(dg, 0ull, ex, ec, t, tc, sh_sd, sh_sc);
        if (act) {
            double g_before = carry + ex;
            if (a.reached(g_before)) {
                bool exact;
                double nv = a.apply(v, exact);
                double g_after = g_before + dg;
                if (a.last(g_after)) nv = fma((v > 0 ? 1.0 : -1.0) * a.unit, -g_after, nv);
                vals[i] = nv;
                if (exact) {
                    keep[i] = 1;
                    made_exact++;
                }
            }
        }
        carry += t;
    }
    zero = 0;
    grid_reduce(grid, red, zero, made_exact);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned left = n_loc - (unsigned)made_exact;
        double nn = left * a.loc_norm / exp_loc;
        res->adj_norm = nn;
        res->n_loc = left;
        res->adjusted = 1;
        if (d_par_out) {
            d_par_out[0] = nn;
            d_par_out[1] = (double)left;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct PivScratch {
    double *pd;
    unsigned long long *pc;
    PivResult *res;
    CompState *st;
    double *par;  // 2 doubles handed from adjust to the sampler
};
static int piv_scratch(fries_ctx *c, int, PivScratch &s) {
    size_t per = (size_t)(2 * c->sm_count) * 2 * FR_RED_STRIDE * 8;  // any cooperative grid here has <= 2 CTAs per SM
    size_t need = 2 * per + sizeof(PivResult) + sizeof(CompState) + 1024;
    FRIES_TRY(c->ensure_scratch(need));
    char *p = (char *)c->d_scratch;
    s.pd = (double *)p;
    s.pc = (unsigned long long *)(p + per);
    p += 2 * per;
    s.res = (PivResult *)p;
    p += (sizeof(PivResult) + 255) & ~(size_t)255;
    s.st = (CompState *)p;
    p += (sizeof(CompState) + 255) & ~(size_t)255;
    s.par = (double *)p;
    return FRIES_OK;
}
static int piv_coop(fries_ctx *c, const void *kernel, int grid, void **args, const char *name) {
    ProfScope ps(c, name);
    CUDA_TRY(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(FR_COMP_BLOCK), args, 0, c->stream));
    c->launch_count++;
    return FRIES_OK;
}
static int piv_grid_size(fries_ctx *c) {
    int g1 = c->coop_grid((const void *)piv_samp_kernel, FR_COMP_BLOCK, 0);
    int g2 = c->coop_grid((const void *)piv_adjust_kernel, FR_COMP_BLOCK, 0);
    return g1 < g2 ? g1 : g2;
}
__global__ void fr_piv_result4(const PivResult *r, double *out) {
    out[0] = r->new_norm;
    out[1] = (double)r->n_drawn;
    out[2] = (double)r->n_units;
    out[3] = (double)r->anomalies;
}

struct PivWork {
    DevBuf<double> E;
    DevBuf<uint32_t> cross, sample, carry;
    int alloc(size_t n, size_t n_samp) {
        FRIES_TRY(E.alloc(n));
        FRIES_TRY(cross.alloc(n_samp));
        FRIES_TRY(sample.alloc(n_samp));
        FRIES_TRY(carry.alloc(n_samp));
        return FRIES_OK;
    }
    PivBufs view() { return PivBufs{E.p, cross.p, sample.p, carry.p}; }
};

static int piv_samp_launch(fries_ctx *c, double *d_vals, size_t n, uint8_t *d_keep, double seg_norm, unsigned n_samp,
                           const double *d_par, const uint32_t *d_draws, PivBufs b, PivScratch &s, int grid) {
    void *args[] = {(void *)&d_vals, (void *)&n,       (void *)&d_keep, (void *)&seg_norm, (void *)&n_samp, (void *)&d_par,
                    (void *)&d_draws, (void *)&b,      (void *)&s.pd,   (void *)&s.pc,     (void *)&s.res};
    return piv_coop(c, (const void *)piv_samp_kernel, grid, args, "piv_samp");
}
static int piv_adjust_launch(fries_ctx *c, double *d_vals, size_t n, uint8_t *d_keep, unsigned n_loc, double exp_loc,
                             unsigned n_tot, double tot_norm, PivScratch &s, int grid, double *d_par_out) {
    void *args[] = {(void *)&d_vals, (void *)&n,    (void *)&d_keep, (void *)&n_loc, (void *)&exp_loc, (void *)&n_tot,
                    (void *)&tot_norm, (void *)&s.pd, (void *)&s.pc, (void *)&s.res, (void *)&d_par_out};
    return piv_coop(c, (const void *)piv_adjust_kernel, grid, args, "piv_adjust");
}

extern "C" int fries_piv_samp_dev(fries_ctx *c, double *d_values, size_t count, double seg_norm, uint32_t n_samp,
                                  uint8_t *d_keep, const uint32_t *d_draws, void *d_work, size_t work_bytes,
                                  double *d_result4) {
    FRIES_REQUIRE(c && (count == 0 || (d_values && d_keep)) && (n_samp == 0 || d_draws),
                  "fries_piv_samp_dev: NULL argument");
    FRIES_REQUIRE(count < 0xfffffff0ull, "fries_piv_samp_dev: at most 2^32 - 16 elements");
    size_t need = ((count * 8 + 255) & ~(size_t)255) + 3 * (((size_t)n_samp * 4 + 255) & ~(size_t)255);
    FRIES_REQUIRE(d_work && work_bytes >= need, "fries_piv_samp_dev: work area of %zu bytes needed (8 per element + 12 per sample)", need);
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = piv_grid_size(c);
    PivScratch s;
    FRIES_TRY(piv_scratch(c, grid, s));
    char *w = (char *)d_work;
    PivBufs b;
    b.E = (double *)w;
    w += (count * 8 + 255) & ~(size_t)255;
    b.cross = (uint32_t *)w;
    w += ((size_t)n_samp * 4 + 255) & ~(size_t)255;
    b.sample = (uint32_t *)w;
    w += ((size_t)n_samp * 4 + 255) & ~(size_t)255;
    b.carry = (uint32_t *)w;
    FRIES_TRY(piv_samp_launch(c, d_values, count, d_keep, seg_norm, n_samp, nullptr, d_draws, b, s, grid));
    if (d_result4) {
        PivResult *r = s.res;
        fr_piv_result4<<<1, 1, 0, c->stream>>>(r, d_result4);
        c->launch_count++;
        CUDA_TRY(cudaGetLastError());
    }
    return FRIES_OK;
}

extern "C" int fries_piv_samp_serial(fries_ctx *c, double *h_values, size_t count, double seg_norm, uint32_t n_samp,
                                     uint8_t *h_keep, const uint32_t *h_draws, size_t *n_draws_used) {
    FRIES_REQUIRE(c && (count == 0 || (h_values && h_keep)) && (n_samp == 0 || h_draws),
                  "fries_piv_samp_serial: NULL argument");
    FRIES_REQUIRE(count < 0xfffffff0ull, "fries_piv_samp_serial: at most 2^32 - 16 elements");
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = piv_grid_size(c);
    DevBuf<double> vals;
    DevBuf<uint8_t> keep;
    DevBuf<uint32_t> draws;
    PivWork w;
    FRIES_TRY(vals.alloc(count));
    FRIES_TRY(keep.alloc(count));
    FRIES_TRY(draws.alloc(2 * (size_t)n_samp));
    FRIES_TRY(w.alloc(count, n_samp));
    PivScratch s;
    FRIES_TRY(piv_scratch(c, grid, s));
    CUDA_TRY(cudaMemcpyAsync(vals.p, h_values, count * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(keep.p, h_keep, count, cudaMemcpyHostToDevice, c->stream));
    if (n_samp) CUDA_TRY(cudaMemcpyAsync(draws.p, h_draws, 8 * (size_t)n_samp, cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(piv_samp_launch(c, vals.p, count, keep.p, seg_norm, n_samp, nullptr, draws.p, w.view(), s, grid));
    PivResult r;
    CUDA_TRY(cudaMemcpyAsync(&r, s.res, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
    if (count) {
        CUDA_TRY(cudaMemcpyAsync(h_values, vals.p, count * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(h_keep, keep.p, count, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (n_draws_used) *n_draws_used = 2 * (size_t)r.n_units;
    return FRIES_OK;
}

extern "C" int fries_adjust_probs(fries_ctx *c, double *h_values, size_t count, uint32_t *n_samp_loc,
                                  double exp_nsamp_loc, uint32_t n_samp_tot, double tot_norm, uint8_t *h_keep,
                                  double *new_norm) {
    FRIES_REQUIRE(c && n_samp_loc && new_norm && (count == 0 || (h_values && h_keep)), "fries_adjust_probs: NULL argument");
    FRIES_REQUIRE(n_samp_tot > 0 && exp_nsamp_loc > 0, "fries_adjust_probs: empty budget");
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = piv_grid_size(c);
    DevBuf<double> vals;
    DevBuf<uint8_t> keep;
    FRIES_TRY(vals.alloc(count));
    FRIES_TRY(keep.alloc(count));
    PivScratch s;
    FRIES_TRY(piv_scratch(c, grid, s));
    CUDA_TRY(cudaMemcpyAsync(vals.p, h_values, count * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(keep.p, h_keep, count, cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(piv_adjust_launch(c, vals.p, count, keep.p, *n_samp_loc, exp_nsamp_loc, n_samp_tot, tot_norm, s, grid, nullptr));
    PivResult r;
    CUDA_TRY(cudaMemcpyAsync(&r, s.res, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
    if (count) {
        CUDA_TRY(cudaMemcpyAsync(h_values, vals.p, count * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(h_keep, keep.p, count, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *n_samp_loc = r.n_loc;
    *new_norm = r.adj_norm;
    return FRIES_OK;
}

// piv_budget (compress_utils.cpp:560-608) is rank 0's host arithmetic on n_ranks numbers in the reference too; the
// fractional parts are settled by the same per-unit rules the kernel uses (piv.cuh), evaluated here on the host.
static void piv_samp_host(std::vector<double> &w, double seg_norm, uint32_t n_samp, const uint32_t *draws, size_t &used) {
    const size_t n = w.size();
    const PivGrid g = piv_grid(seg_norm, n_samp);
    std::vector<double> E(n);
    double run = 0;
    for (size_t i = 0; i < n; i++) E[i] = run += fabs(w[i]);
    std::vector<uint32_t> cross(n_samp, PIV_NONE), sample(n_samp, PIV_NONE), carry(n_samp, PIV_NONE);
    for (size_t i = 0; i < n; i++)
        piv_mark_cross(g, i ? E[i - 1] : 0.0, E[i], (uint32_t)i, [&](uint32_t k, uint32_t el) { cross[k] = el; });
    uint32_t n_crossed = n ? g.borders_le(E[n - 1]) : 0;
    auto loadE = [&](uint32_t j) { return E[j]; };
    auto loadC = [&](uint32_t k) { return cross[k]; };
    uint32_t n_units = piv_n_units(g, n_crossed, n_crossed ? cross[n_crossed - 1] : 0, n);
    std::vector<uint8_t> drawn(n, 0);
    for (uint32_t k = 0; k < n_units; k++) {
        double r1 = draws[used++] / 4294967296.0, r2 = draws[used++] / 4294967296.0;
        piv_unit(g, k, n_crossed, n, loadE, loadC, r1, r2, sample[k], carry[k]);
    }
    for (uint32_t k = 0; k < n_units; k++) {
        uint32_t s = sample[k];
        if (s == PIV_CARRIED) s = piv_resolve(k, [&](uint32_t j) { return carry[j]; });
        if (s < n) drawn[s] = 1;
    }
    for (size_t i = 0; i < n; i++) w[i] = drawn[i] ? g.unit : 0.0;
}

extern "C" int fries_piv_budget(const double *loc_norms, int n_ranks, uint32_t n_samp, const uint32_t *h_draws,
                                size_t *n_draws_used, uint32_t *budgets) {
    FRIES_REQUIRE(loc_norms && budgets && n_ranks >= 1, "fries_piv_budget: bad argument");
    double glob = 0;
    for (int p = 0; p < n_ranks; p++) glob += loc_norms[p];
    std::vector<double> wt(n_ranks);
    uint32_t tot = 0, n_frac = 0;
    for (int p = 0; p < n_ranks; p++) {
        budgets[p] = glob > 0 ? (uint32_t)(loc_norms[p] / glob * n_samp) : 0;
        tot += budgets[p];
        wt[p] = loc_norms[p] - budgets[p] * glob / n_samp;
        if (wt[p] < 1e-12) wt[p] = 0;
        if (wt[p] > 0) n_frac++;
    }
    size_t used = 0;
    if (n_frac == n_samp - tot) {
        for (int p = 0; p < n_ranks; p++)
            if (wt[p] > 0) budgets[p]++;
        tot = n_samp;
    }
    if (tot < n_samp) {
        FRIES_REQUIRE(h_draws, "fries_piv_budget: draws needed");
        if (!(glob > 0)) {
            // nothing left to sample on any rank: the reference's sweep (unit = 0) still steps over one rank per unit and
            // draws twice each time (compress_utils.cpp:420-505); no budget changes
            used = 2 * (size_t)((uint32_t)n_ranks < n_samp - tot ? (uint32_t)n_ranks : n_samp - tot);
        } else
            piv_samp_host(wt, glob * (n_samp - tot) / n_samp, n_samp - tot, h_draws, used);
        for (int p = 0; p < n_ranks; p++)
            if (wt[p] > 0) budgets[p]++;
    }
    if (n_draws_used) *n_draws_used = used;
    return FRIES_OK;
}

// piv_comp_parallel (single rank) on a RESIDENT vector: find_preserve, budget, adjust_probs, pivotal sampling in place;
// d_keep out: 1 = zeroed element.  h_draws[n_draws]: the caller's generator outputs, *used is advanced.  Host round
// trips: the budget left by find_preserve and the number of sampling units.  Used by fries_apply_hbpp_piv (hbpp.cu).
int fries_piv_comp_resident(fries_ctx *c, double *d_vals, size_t n, uint32_t compress_size, uint8_t *d_keep,
                            const uint32_t *h_draws, size_t n_draws, size_t *used) {
    FRIES_REQUIRE(n < 0xfffffff0ull, "pivotal compression: at most 2^32 - 16 elements");
    int grid = piv_grid_size(c);
    DevBuf<uint32_t> draws;
    PivWork w;
    FRIES_TRY(draws.alloc(2 * (size_t)compress_size + 2));
    FRIES_TRY(w.alloc(n, compress_size));
    PivScratch s;
    FRIES_TRY(piv_scratch(c, grid, s));
    CUDA_TRY(cudaMemsetAsync(s.st, 0, sizeof(CompState), c->stream));
    FRIES_TRY(fries_find_preserve_launch(c, d_vals, n, nullptr, compress_size, d_keep, s.st, s.pd, s.pc, 0, nullptr));
    CompState st;
    CUDA_TRY(cudaMemcpyAsync(&st, s.st, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const unsigned n_samp = st.n_samp_left;
    const double loc = st.loc_norm;
    uint32_t loc_samp = 0;
    const double *d_par = nullptr;
    if (n_samp != 0) {
        size_t bu = 0;
        FRIES_REQUIRE(*used + 2 <= n_draws, "pivotal compression: out of draws");
        FRIES_TRY(fries_piv_budget(&loc, 1, n_samp, h_draws + *used, &bu, &loc_samp));
        *used += bu;
        double exp_loc = n_samp * loc / loc;
        if (exp_loc > 0) {
            FRIES_TRY(piv_adjust_launch(c, d_vals, n, d_keep, loc_samp, exp_loc, n_samp, loc, s, grid, s.par));
            d_par = s.par;
        }
    }
    FRIES_REQUIRE(*used + 2 * (size_t)loc_samp <= n_draws, "pivotal compression: out of draws (2 per sample)");
    if (loc_samp)
        CUDA_TRY(cudaMemcpyAsync(draws.p, h_draws + *used, 8 * (size_t)loc_samp, cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(piv_samp_launch(c, d_vals, n, d_keep, 0.0, loc_samp, d_par, draws.p, w.view(), s, grid));
    PivResult r;
    CUDA_TRY(cudaMemcpyAsync(&r, s.res, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *used += 2 * (size_t)r.n_units;
    return FRIES_OK;
}

// piv_comp_parallel (compress_utils.cpp:354-387) for one rank of n_ranks.  h_loc_norms[n_ranks]: the residual norms of
// the OTHER ranks' find_preserve (the reference all-gathers them, :365); this rank's entry is filled in here.  With
// n_ranks > 1 the preserved set is decided by the caller's collective find_preserve beforehand: pass preserved = 1 and
// h_keep / n_samp_left / h_loc_norms[rank] from it.  Single rank: preserved = 0, h_loc_norms may be NULL.
extern "C" int fries_piv_comp(fries_ctx *c, double *h_values, size_t count, uint32_t compress_size, uint8_t *h_keep,
                              const uint32_t *h_draws, size_t *n_draws_used, double *h_loc_norms, int n_ranks, int rank,
                              int preserved, uint32_t n_samp_left) {
    FRIES_REQUIRE(c && (count == 0 || (h_values && h_keep)), "fries_piv_comp: NULL argument");
    FRIES_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks && (n_ranks == 1 || (h_loc_norms && preserved)),
                  "fries_piv_comp: with %d ranks the caller passes the all-gathered norms and the preserved set", n_ranks);
    FRIES_REQUIRE(count < 0xfffffff0ull, "fries_piv_comp: at most 2^32 - 16 elements");
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = piv_grid_size(c);
    DevBuf<double> vals;
    DevBuf<uint8_t> keep;
    DevBuf<uint32_t> draws;
    PivWork w;
    FRIES_TRY(vals.alloc(count));
    FRIES_TRY(keep.alloc(count));
    FRIES_TRY(draws.alloc(2 * (size_t)compress_size + 2));
    FRIES_TRY(w.alloc(count, compress_size));
    PivScratch s;
    FRIES_TRY(piv_scratch(c, grid, s));
    CUDA_TRY(cudaMemcpyAsync(vals.p, h_values, count * 8, cudaMemcpyHostToDevice, c->stream));
    unsigned n_samp = compress_size;
    double loc = 0;
    if (!preserved) {
        CUDA_TRY(cudaMemsetAsync(s.st, 0, sizeof(CompState), c->stream));
        FRIES_TRY(fries_find_preserve_launch(c, vals.p, count, nullptr, compress_size, keep.p, s.st, s.pd, s.pc, 0,
                                             nullptr));
        CompState st;
        CUDA_TRY(cudaMemcpyAsync(&st, s.st, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        n_samp = st.n_samp_left;
        loc = st.loc_norm;
    } else {
        CUDA_TRY(cudaMemcpyAsync(keep.p, h_keep, count, cudaMemcpyHostToDevice, c->stream));
        n_samp = n_samp_left;
        loc = h_loc_norms[rank];
    }
    std::vector<double> norms(n_ranks, 0.0);
    for (int p = 0; p < n_ranks; p++) norms[p] = p == rank ? loc : h_loc_norms[p];
    double glob = 0;
    for (int p = 0; p < n_ranks; p++) glob += norms[p];
    size_t used = 0;
    uint32_t loc_samp = 0;
    double new_norm = 0;
    const double *d_par = nullptr;
    if (n_samp != 0) {
        std::vector<uint32_t> budgets(n_ranks);
        FRIES_TRY(fries_piv_budget(norms.data(), n_ranks, n_samp, h_draws, &used, budgets.data()));
        loc_samp = budgets[rank];
        double exp_loc = n_samp * loc / glob;
        if (exp_loc > 0) {
            FRIES_TRY(piv_adjust_launch(c, vals.p, count, keep.p, loc_samp, exp_loc, n_samp, glob, s, grid, s.par));
            d_par = s.par;  // {new_norm, loc_samp} stay on the device
        }
    }
    size_t n_up = 2 * (size_t)(loc_samp < compress_size ? loc_samp : compress_size);
    if (n_up) CUDA_TRY(cudaMemcpyAsync(draws.p, h_draws + used, 4 * n_up, cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(piv_samp_launch(c, vals.p, count, keep.p, new_norm, loc_samp, d_par, draws.p, w.view(), s, grid));
    PivResult r;
    CUDA_TRY(cudaMemcpyAsync(&r, s.res, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
    if (count) {
        CUDA_TRY(cudaMemcpyAsync(h_values, vals.p, count * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(h_keep, keep.p, count, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (n_draws_used) *n_draws_used = used + 2 * (size_t)r.n_units;
    if (h_loc_norms) h_loc_norms[rank] = r.new_norm;
    return FRIES_OK;
}

// compress_vecs / compress_vecs_sys (FRIES/vec_utils.cpp:10-70) on the resident store: rows [start_row, end_row) are
// compressed one after the other to compress_size elements (method 0: piv_comp_parallel, draws consumed as there;
// method 1: find_preserve + sys_comp with one draw per row), then every element that is zero in ALL rows is deleted
// (del_at_pos only removes such elements, vec_utils.hpp:458-476, so the reference's del_arr bookkeeping reduces to
// this).  Single rank.
// ---- compress_vecs_multi (FRIES/vec_utils.cpp:73-127): multinomial compression of a row with the alias method --------
// p_i = |v_i / norm| (the sign kept aside), alias table of p (setup_alias, compress_utils.cpp:823-857: a sequential
// two-stack construction -- on the host, as in the reference: the table decides which draws land where, so it has to be
// that construction exactly), compress_size draws of two uniforms each (sample_alias :882-897), v_i <- norm * count_i *
// sign_i / compress_size.
__global__ void multi_normalise_kernel(double *vals, size_t n, double norm, uint8_t *positive) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double q = vals[i] / norm;
        positive[i] = q > 0 ? 1 : 0;
        vals[i] = fabs(q);
    }
}
__global__ void multi_sample_kernel(const uint32_t *__restrict__ aliases, const double *__restrict__ alias_probs, size_t n,
                                    const uint32_t *__restrict__ draws, uint32_t n_samp, uint32_t *counts) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_samp; k += gridDim.x * blockDim.x) {
        const uint16_t chosen = (uint16_t)(draws[2 * k] / 4294967296.0 * (double)n);
        const double u = draws[2 * k + 1] / 4294967296.0;
        atomicAdd(&counts[u < alias_probs[chosen] ? (uint32_t)chosen : aliases[chosen]], 1u);
    }
}
__global__ void multi_apply_kernel(double *vals, size_t n, double norm, const uint32_t *__restrict__ counts,
                                   const uint8_t *__restrict__ positive, uint32_t compress_size) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        vals[i] = norm * (double)(uint16_t)counts[i] * (positive[i] ? 1 : -1) / (double)compress_size;
}
static void host_setup_alias(const double *probs, uint32_t *aliases, double *alias_probs, size_t n) {
    std::vector<uint32_t> smaller, bigger;
    smaller.reserve(n);
    bigger.reserve(n);
    for (size_t i = 0; i < n; i++) {
        aliases[i] = (uint32_t)i;
        alias_probs[i] = (double)n * probs[i];
        (alias_probs[i] < 1 ? smaller : bigger).push_back((uint32_t)i);
    }
    while (!smaller.empty() && !bigger.empty()) {
        const uint32_t sidx = smaller.back(), b = bigger.back();
        aliases[sidx] = b;
        alias_probs[b] += alias_probs[sidx] - 1;
        if (alias_probs[b] < 1) {
            smaller.back() = b;
            bigger.pop_back();
        } else {
            smaller.pop_back();
        }
    }
}
static int compress_row_multi(fries_vec *vec, unsigned row, size_t n, uint32_t compress_size, const uint32_t *h_draws,
                              size_t n_draws, size_t *used) {
    fries_ctx *c = vec->ctx;
    FRIES_REQUIRE(n <= 65535 && compress_size <= 65535,
                  "fries_vec_compress (multi): the reference's sample_alias counts in uint16 (%zu states, %u samples)", n,
                  compress_size);
    FRIES_REQUIRE(*used + 4 * (size_t)compress_size <= n_draws, "fries_vec_compress (multi): out of draws (4 per sample and row)");
    if (n == 0) {
        *used += 4 * (size_t)compress_size;
        return FRIES_OK;
    }
    double norm = 0;
    FRIES_TRY(fries_vec_local_norm(vec, row, &norm));
    double *d_vals = vec->vals[vec->cur].p + (size_t)row * vec->cap;
    DevBuf<uint8_t> positive;
    DevBuf<uint32_t> aliases, counts, draws;
    DevBuf<double> aprobs;
    FRIES_TRY(positive.alloc(n));
    FRIES_TRY(aliases.alloc(n));
    FRIES_TRY(counts.alloc(n));
    FRIES_TRY(aprobs.alloc(n));
    FRIES_TRY(draws.alloc(2 * (size_t)compress_size + 2));
    const int grid = c->sm_count * 4;
    multi_normalise_kernel<<<grid, 256, 0, c->stream>>>(d_vals, n, norm, positive.p);
    c->launch_count++;
    std::vector<double> hp(n), hap(n);
    std::vector<uint32_t> hal(n);
    CUDA_TRY(cudaMemcpyAsync(hp.data(), d_vals, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // one rank: the reference first spreads the samples over the ranks with a one-state alias table (vec_utils.cpp:101-109),
    // which consumes two draws per sample and gives this rank all of them
    *used += 2 * (size_t)compress_size;
    host_setup_alias(hp.data(), hal.data(), hap.data(), n);
    CUDA_TRY(cudaMemcpyAsync(aliases.p, hal.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(aprobs.p, hap.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(draws.p, h_draws + *used, 8 * (size_t)compress_size, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(counts.p, 0, n * 4, c->stream));
    multi_sample_kernel<<<grid, 256, 0, c->stream>>>(aliases.p, aprobs.p, n, draws.p, compress_size, counts.p);
    c->launch_count++;
    multi_apply_kernel<<<grid, 256, 0, c->stream>>>(d_vals, n, norm, counts.p, positive.p, compress_size);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // the staging vectors go out of scope
    *used += 2 * (size_t)compress_size;
    return FRIES_OK;
}

extern "C" int fries_vec_compress(fries_vec *vec, unsigned start_row, unsigned end_row, uint32_t compress_size, int method,
                                  const uint32_t *h_draws, size_t n_draws, size_t *n_draws_used) {
    FRIES_REQUIRE(vec && (h_draws || n_draws == 0), "fries_vec_compress: NULL argument");
    FRIES_REQUIRE(start_row <= end_row && end_row <= vec->n_vecs, "fries_vec_compress: rows [%u, %u) of %u", start_row,
                  end_row, vec->n_vecs);
    FRIES_REQUIRE(method >= 0 && method <= 2, "fries_vec_compress: method 0 (pivotal), 1 (systematic) or 2 (multinomial)");
    FRIES_REQUIRE(vec->n_ranks == 1, "fries_vec_compress: single rank");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    const size_t n = (size_t)cnt.n;
    int grid = piv_grid_size(c);
    DevBuf<uint8_t> keep;
    DevBuf<uint32_t> draws;
    PivWork w;
    FRIES_TRY(keep.alloc(n));
    if (method == 0) {
        FRIES_TRY(draws.alloc(2 * (size_t)compress_size + 2));
        FRIES_TRY(w.alloc(n, compress_size));
    }
    PivScratch s;
    FRIES_TRY(piv_scratch(c, grid, s));
    size_t used = 0;
    for (unsigned row = start_row; row < end_row; row++) {
        if (method == 2) {
            FRIES_TRY(compress_row_multi(vec, row, n, compress_size, h_draws, n_draws, &used));
            continue;
        }
        double *d_vals = vec->vals[vec->cur].p + (size_t)row * vec->cap;
        CUDA_TRY(cudaMemsetAsync(s.st, 0, sizeof(CompState), c->stream));
        FRIES_TRY(fries_find_preserve_launch(c, d_vals, n, nullptr, compress_size, keep.p, s.st, s.pd, s.pc, 0, nullptr));
        CompState st;
        CUDA_TRY(cudaMemcpyAsync(&st, s.st, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const unsigned n_samp = st.n_samp_left;
        const double loc = st.loc_norm;
        if (method == 1) {
            FRIES_REQUIRE(used < n_draws, "fries_vec_compress: out of draws (one per row)");
            double rn = h_draws[used++] / 4294967296.0;
            FRIES_TRY(fries_sys_comp_launch(c, d_vals, n, nullptr, keep.p, nullptr, 0.0, loc, (long long)n_samp, rn, s.st,
                                            s.pd, s.pc, 0, nullptr));
            continue;
        }
        uint32_t loc_samp = 0;
        const double *d_par = nullptr;
        if (n_samp != 0) {
            size_t bu = 0;
            FRIES_REQUIRE(used + 2 <= n_draws, "fries_vec_compress: out of draws");
            FRIES_TRY(fries_piv_budget(&loc, 1, n_samp, h_draws + used, &bu, &loc_samp));
            used += bu;
            double exp_loc = n_samp * loc / loc;
            if (exp_loc > 0) {
                FRIES_TRY(piv_adjust_launch(c, d_vals, n, keep.p, loc_samp, exp_loc, n_samp, loc, s, grid, s.par));
                d_par = s.par;
            }
        }
        FRIES_REQUIRE(used + 2 * (size_t)loc_samp <= n_draws, "fries_vec_compress: out of draws (2 per sample)");
        if (loc_samp)
            CUDA_TRY(cudaMemcpyAsync(draws.p, h_draws + used, 8 * (size_t)loc_samp, cudaMemcpyHostToDevice, c->stream));
        FRIES_TRY(piv_samp_launch(c, d_vals, n, keep.p, 0.0, loc_samp, d_par, draws.p, w.view(), s, grid));
        PivResult r;
        CUDA_TRY(cudaMemcpyAsync(&r, s.res, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // also keeps h_draws' staging buffer safe for the next row
        used += 2 * (size_t)r.n_units;
    }
    FRIES_TRY(fries_vec_compact_dev(vec));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (n_draws_used) *n_draws_used = used;
    return FRIES_OK;
}
