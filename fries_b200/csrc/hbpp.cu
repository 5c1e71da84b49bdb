// a10: apply_HBPP_sys (heat_bathPP.cpp:686-992) on the device.  See hbpp.cuh.
#include "hbpp.cuh"

extern __shared__ __align__(16) double fr_dyn_smem[];

#ifndef FR_STAGE_MIN_CTAS
#define FR_STAGE_MIN_CTAS 2  // two 512-thread CTAs per SM (<= 64 registers): the stage passes are latency bound
#endif

// MINCTAS = 2 is the product's configuration.  MINCTAS = 1 (one CTA per SM, up to 128 registers: no spills, half the
// resident warps) is compiled in as a measurement variant, selected with FRIES_STAGE_CTAS=1 in the environment; same
// arithmetic, same results.
template <int S, int MINCTAS>
__global__ void __launch_bounds__(FR_COMP_BLOCK, MINCTAS)
hbpp_stage_kernel(MolView gm, HbStageIO io, CompSubBufs bufs, unsigned n_samp, double rn) {
    HbProvider<S> prov;
    prov.m = mol_stage_shared(gm, fr_dyn_smem);
    prov.io = io;
    comp_sub_engine(prov, bufs, n_samp, rn);
}
// Second-generation engine (compress2.cuh): one 512-thread CTA per SM, up to 128 registers, four inputs per thread; the
// table blob arrives by one bulk asynchronous copy.  This is the product's configuration; FRIES_ENGINE=1 in the
// environment selects the first-generation kernels above (measurement / regression variant, same results up to FP ties).
template <int S, int MINCTAS, bool MULTI>
__global__ void __launch_bounds__(FR2_NT, MINCTAS)
hbpp_stage2_kernel(MolView gm, HbStageIO io, CompSubBufs2 bufs, unsigned n_samp, double rn) {
    HbProvider<S> prov;
    prov.m = mol_stage_shared_bulk(gm, fr_dyn_smem);
    prov.io = io;
    comp_sub_engine2<MULTI>(prov, bufs, n_samp, rn);
}
// CTAs per SM of the second-generation kernels.  One (128 registers, no spills) when an SM holds a few thousand inputs:
// the stage is a chain of short phases and extra warps add skeleton work.  Two (64 registers, ~1 kB of spill traffic per
// thread, twice the warps) when the scratch was sized for millions of samples: the row loops are then bound by
// per-warp latency chains and the second CTA hides them (measured round 2 at 1.25e7 samples: 9.8 -> 7.4 ms for the five
// stages; at 2.6e5: 0.49 -> 0.51 ms).  FRIES_STAGE2_CTAS=1|2 in the environment overrides the choice.
// (end of round 2, same-box A/B at 2.6e5 samples, scratch for 1.04e6: two CTAs gain 4 % on stage 2 and 11 % on stage 3 -- the
// stages with the longest rows -- and lose 3-15 % on the others: those two switch at 1e6)
static int stage2_ctas(size_t cap, int stage) {
    static int forced = [] {
        const char *e = getenv("FRIES_STAGE2_CTAS");
        return (e && (e[0] == '1' || e[0] == '2') && e[1] == 0) ? e[0] - '0' : 0;
    }();
    if (forced) return forced;
    if (cap >= (size_t)4000000) return 2;
    return (cap >= (size_t)1000000 && (stage == 2 || stage == 3)) ? 2 : 1;
}
static int stage_engine() {
    static int v = [] {
        const char *e = getenv("FRIES_ENGINE");
        return (e && e[0] == '1' && e[1] == 0) ? 1 : 2;
    }();
    return v;
}
extern "C" int fries_debug_stage_engine(int *generation) {
    if (generation) *generation = stage_engine();
    return FRIES_OK;
}
static int stage_min_ctas() {
    static int v = [] {
        const char *e = getenv("FRIES_STAGE_CTAS");
        return (e && e[0] == '1' && e[1] == 0) ? 1 : FR_STAGE_MIN_CTAS;
    }();
    return v;
}
extern "C" int fries_debug_stage_ctas(int *ctas_per_sm) {
    if (ctas_per_sm) *ctas_per_sm = stage_min_ctas();
    return FRIES_OK;
}
template <int S>
static const void *stage_kernel_ptr() {
    return stage_min_ctas() == 1 ? (const void *)hbpp_stage_kernel<S, 1> : (const void *)hbpp_stage_kernel<S, FR_STAGE_MIN_CTAS>;
}

// ---------------------------------------------------------------------------------------------------
// finalize :917-991 (+ the spawn loop body of frisys_mol.cpp:436-461 when sp.out_keys != nullptr).
// One thread per sample; tables in shared memory, integrals through L2.
// ---------------------------------------------------------------------------------------------------
// four CTAs of 256 threads per SM (<= 64 registers): the kernel is a chain of dependent gathers per sample (output list ->
// path -> determinant -> integrals in L2), so it wants resident warps; at the compiler's free choice (114 registers) only two
// CTAs fit
__global__ void __launch_bounds__(256, 4)
hbpp_finalize_kernel(MolView gm, const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ n_ptr,
                     unsigned long long in_cap, const double *__restrict__ pv, const uint32_t *__restrict__ pw,
                     const uint32_t *__restrict__ ps, const uint32_t *__restrict__ pdet,
                     const uint32_t *__restrict__ ppath, double p_doub, int new_hb, double *__restrict__ fin_val,
                     uint32_t *__restrict__ fin_det, uint32_t *__restrict__ fin_orbs, HbSpawnArgs sp, CompState *st,
                     double cutoff) {
    MolView m = mol_stage_shared(gm, fr_dyn_smem);
    __shared__ uint32_t s_pscr[64];
    if (sp.n_ranks > 1) {
        if (threadIdx.x < 64) s_pscr[threadIdx.x] = sp.proc_scr[threadIdx.x];
        __syncthreads();
    }
    unsigned long long n = *n_ptr;
    if (n > in_cap) n = in_cap;
    unsigned long long ok = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    // warp-uniform trip count: the routed path uses warp-wide match/shuffle
    size_t n_round = (n + 31) & ~(size_t)31;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_round; i0 += stride) {
      const bool live = i0 < n;
      const size_t i = live ? i0 : 0;
      double el = 0;
      uint64_t nk = FRIES_EMPTY_KEY;
      double add = 0;
      if (live) {
        const uint32_t widx = pw[i], sub = ps[i];
        const uint32_t d = pdet[widx], pp = ppath[widx];
        const uint64_t key = keys[d];
        uint8_t orbs[4];
        bool is_doub;
        el = hbpp_finalize_sample(m, key, pp, sub, pv[i], p_doub, new_hb, cutoff, orbs, is_doub);
        if (el != 0) ok++;
        if (sp.out_keys || sp.n_ranks > 1) {
            // spawn loop body frisys_mol.cpp:436-461
            if (el != 0) nk = hbpp_spawn_element(key, orbs, is_doub, el, sp.v0[d], sp.eps, sp.init_thresh, add);
            if (sp.n_ranks <= 1) {
                sp.out_keys[i] = nk;
                sp.out_vals[i] = add;
            }
        } else {
            fin_val[i] = el;
            fin_det[i] = d;
            fin_orbs[i] = pk(orbs[0], orbs[1], orbs[2], orbs[3]);
        }
      }  // live
      if (sp.n_ranks > 1) {
        // route: one atomicAdd per (warp, destination) instead of one per element
        int owner = -1;
        if (nk != FRIES_EMPTY_KEY) owner = (int)(fr_det_hash(nk & ~FRIES_INI_FLAG, s_pscr) % (unsigned)sp.n_ranks);
        unsigned peers = __match_any_sync(0xffffffffu, owner);
        if (owner >= 0) {
            unsigned lane = threadIdx.x & 31;
            int leader = __ffs(peers) - 1;
            unsigned rank_in = __popc(peers & ((1u << lane) - 1));
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(&sp.send_counts[owner], (unsigned long long)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            unsigned long long slot = base + rank_in;
            if (slot < sp.seg_cap) {
                uint64_t *seg = sp.peer_win[owner] ? sp.peer_win[owner] + (size_t)sp.rank * 2 * sp.seg_cap
                                                   : sp.send_buf + (size_t)owner * 2 * sp.seg_cap;
                seg[slot] = nk;
                seg[sp.seg_cap + slot] = (uint64_t)__double_as_longlong(add);
            } else {
                atomicAdd(&sp.send_counts[sp.n_ranks], 1ull);
            }
        }
      }
    }
    ok = warp_sum_u64(ok);
    if ((threadIdx.x & 31) == 0 && ok) atomicAdd(&st->n_out, ok);
    if (blockIdx.x == 0 && threadIdx.x == 0) st->n_in = n;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int fries_hbpp_alloc(fries_ctx *c, size_t cap, fries_hbpp **out, bool stages) {
    FRIES_REQUIRE(cap >= 1 && cap < 0x7fffffffull, "fries_hbpp: spawn capacity %zu out of range", cap);
    fries_hbpp *hb = new fries_hbpp();
    hb->ctx = c;
    hb->cap = cap;
    int rc = FRIES_OK;
#define A(buf, n) if (rc == FRIES_OK) rc = hb->buf.alloc(n)
    if (stages) {
        A(veff, cap); A(wtr, cap); A(lb, cap); A(rinv, cap); A(ndiv, cap); A(keep, cap); A(kcnt, cap); A(nsub, cap);
        for (int b = 0; b < 2; b++) {
            A(oval[b], cap); A(owidx[b], cap); A(osub[b], cap); A(det[b], cap); A(path[b], cap);
        }
        A(fin_val, cap); A(fin_det, cap); A(fin_orbs, cap);
    }
    A(part_d, FR_RED_PART_LEN); A(part_c, FR_RED_PART_LEN); A(st, 8); A(n_scalar, 4); A(scal, 64);
    A(cand_x, FR_CAND_GCAP); A(cand_m, FR_CAND_GCAP); A(cand_idx, FR_CAND_GCAP); A(pred, 8); A(gcomb, GC_STATE_WORDS);
#undef A
    if (rc == FRIES_OK && cudaMemset(hb->scal.p, 0, 64 * sizeof(double)) != cudaSuccess) {
        fries_set_error("fries_hbpp_alloc: cudaMemset failed");
        rc = FRIES_ERR_CUDA;
    }
    if (rc == FRIES_OK && cudaMemset(hb->gcomb.p, 0, (size_t)GC_STATE_WORDS * 8) != cudaSuccess) {
        fries_set_error("fries_hbpp_alloc: cudaMemset failed");
        rc = FRIES_ERR_CUDA;
    }
    if (rc == FRIES_OK && cudaMemset(hb->pred.p, 0, 8 * sizeof(KeepPred)) != cudaSuccess) {
        fries_set_error("fries_hbpp_alloc: cudaMemset failed");
        rc = FRIES_ERR_CUDA;
    }
    if (rc == FRIES_OK) {  // bracket predictors: how long a prediction error is remembered (FRIES_PRED_DECAY: measurement knob)
        KeepPred init[8];
        memset(init, 0, sizeof(init));
        const char *e = getenv("FRIES_PRED_DECAY");
        const double decay = e ? atof(e) : 0.7;
        const char *ef = getenv("FRIES_PRED_FACTOR");
        const double factor = ef ? atof(ef) : 4.0;  // same-box A/B, round 2: 6 / 4 / 3 -> 839 / 850 / 858 it/s, first misses at 3
        for (int k = 0; k < 8; k++) {
            init[k].decay = decay;
            init[k].factor = factor;
        }
        if (cudaMemcpy(hb->pred.p, init, sizeof(init), cudaMemcpyHostToDevice) != cudaSuccess) {
            fries_set_error("fries_hbpp_alloc: cudaMemcpy failed");
            rc = FRIES_ERR_CUDA;
        }
    }
    if (rc != FRIES_OK) {
        delete hb;
        return rc;
    }
    *out = hb;
    return FRIES_OK;
}

// Diagnostics: run every standalone compression `n` times on the same inputs and return the last run, so that the
// parity tests can exercise the bracketed threshold solve (which needs the fixed point of a previous run).
int fr_debug_repeat = 1;
int fr_bracket_on = 1;
extern "C" int fries_debug_set_repeat(int n) {
    fr_debug_repeat = n < 1 ? 1 : n;
    return FRIES_OK;
}
double fr_debug_perturb = 0;  // warm-up runs of a repeated standalone call see the values scaled by (1 + perturb)
extern "C" int fries_debug_set_perturb(double rel) {
    fr_debug_perturb = rel;
    return FRIES_OK;
}
// scale a device array (warm-up runs of the repeated standalone calls)
__global__ void fr_scale_kernel(double *v, size_t n, double f) {
    // per-element factor 1 + (f - 1) u_i with u_i in [0, 2): the fixed point moves by about f - 1, not exactly
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double u = (double)(((unsigned)i * 2654435761u) >> 16) / 32768.0;
        v[i] *= 1.0 + (f - 1.0) * u;
    }
}
int fr_debug_upload_vals(fries_ctx *c, double *d_vals, const double *h_vals, size_t n, bool warmup) {
    CUDA_TRY(cudaMemcpyAsync(d_vals, h_vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (warmup && fr_debug_perturb != 0 && n) fr_scale_kernel<<<256, 256, 0, c->stream>>>(d_vals, n, 1.0 + fr_debug_perturb);
    return FRIES_OK;
}
int fr_last_fast = 0;  // number of compressions of the last standalone call decided by the bracketed solve
extern "C" int fries_debug_last_fast(int *n_fast) {
    if (n_fast) *n_fast = fr_last_fast;
    return FRIES_OK;
}
extern "C" int fries_debug_set_bracket(int on) {
    fr_bracket_on = on ? 1 : 0;
    return FRIES_OK;
}

extern "C" int fries_hbpp_destroy(fries_hbpp *hb) {
    if (hb) {
        cudaSetDevice(hb->ctx->device);
        delete hb;
    }
    return FRIES_OK;
}

template <int S>
static int launch_stage(fries_hbpp *hb, fries_mol *mol, HbStageIO &io, CompSubBufs &bufs, unsigned n_samp, double rn) {
    fries_ctx *c = hb->ctx;
    size_t smem = (size_t)mol->view.d.blob_doubles * 8;
    if (hb->grid == 0) {
        // one grid size for all stages: the smallest co-resident grid among them (set on first use)
        int g = c->coop_grid(stage_kernel_ptr<0>(), FR_COMP_BLOCK, smem);
        int t;
        t = c->coop_grid(stage_kernel_ptr<1>(), FR_COMP_BLOCK, smem); g = t < g ? t : g;
        t = c->coop_grid(stage_kernel_ptr<2>(), FR_COMP_BLOCK, smem); g = t < g ? t : g;
        t = c->coop_grid(stage_kernel_ptr<3>(), FR_COMP_BLOCK, smem); g = t < g ? t : g;
        t = c->coop_grid(stage_kernel_ptr<4>(), FR_COMP_BLOCK, smem); g = t < g ? t : g;
        hb->grid = g;
    }
    MolView gm = mol->view;
    static const char *names[] = {"hbpp_stage0", "hbpp_stage1", "hbpp_stage2", "hbpp_stage3", "hbpp_stage4"};
    ProfScope ps(c, names[S]);
    if (stage_engine() == 2) {
        const int ctas = stage2_ctas(hb->cap, S);
        const bool multi = bufs.cm.n_ranks > 1;
        const void *kern = multi ? (ctas == 2 ? (const void *)hbpp_stage2_kernel<S, 2, true> : (const void *)hbpp_stage2_kernel<S, 1, true>)
                                 : (ctas == 2 ? (const void *)hbpp_stage2_kernel<S, 2, false> : (const void *)hbpp_stage2_kernel<S, 1, false>);
        static bool attr_set[4] = {false, false, false, false};  // per instantiation: opt in to more than 48 kB of shared memory
        const int ai = (ctas - 1) + (multi ? 2 : 0);
        if (!attr_set[ai]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_set[ai] = true;
        }
        if (hb->grid2_s[S] != c->sm_count * ctas) {
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FR2_NT, smem));
            FRIES_REQUIRE(per_sm >= ctas, "hbpp_stage2_kernel<%d>: %d CTAs per SM do not fit (%zu B of dynamic shared memory)", S,
                          ctas, smem);
            hb->grid2_s[S] = c->sm_count * ctas;
            hb->grid2 = hb->grid2_s[S];
        }
        static const bool marks = getenv("FRIES_CTA_MARKS") != nullptr;
        if (marks && !hb->cta_marks.p) FRIES_TRY(hb->cta_marks.alloc((size_t)7 * 8 * 1024));
        CompSubBufs2 b2{bufs, hb->cand_idx.p, hb->gcomb.p, marks ? hb->cta_marks.p + (size_t)S * 8 * 1024 : nullptr};
        void *args2[] = {(void *)&gm, (void *)&io, (void *)&b2, (void *)&n_samp, (void *)&rn};
        CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(hb->grid2_s[S]), dim3(FR2_NT), args2, smem, c->stream));
        c->launch_count++;
        return FRIES_OK;
    }
    void *args[] = {(void *)&gm, (void *)&io, (void *)&bufs, (void *)&n_samp, (void *)&rn};
    CUDA_TRY(cudaLaunchCooperativeKernel(stage_kernel_ptr<S>(), dim3(hb->grid), dim3(FR_COMP_BLOCK), args,
                                         smem, c->stream));
    c->launch_count++;
    return FRIES_OK;
}

int fries_hbpp_stages_dev(fries_hbpp *hb, fries_mol *mol, const uint64_t *d_keys, const double *d_vals,
                          const unsigned long long *d_n, double p_doub, int new_hb, const double *u5, unsigned n_samp) {
    fries_ctx *c = hb->ctx;
    CUDA_TRY(cudaMemsetAsync(hb->st.p, 0, 8 * sizeof(CompState), c->stream));
    for (int s = 0; s < 5; s++) {
        int o = s & 1, p = o ^ 1;
        HbStageIO io;
        io.keys = d_keys;
        io.vals = d_vals;
        io.n_in = s == 0 ? d_n : &hb->st.p[s - 1].n_out;
        io.pv = hb->oval[p].p;
        io.pw = hb->owidx[p].p;
        io.ps = hb->osub[p].p;
        io.pdet = hb->det[p].p;
        io.ppath = hb->path[p].p;
        io.det = hb->det[o].p;
        io.path = hb->path[o].p;
        io.p_doub = p_doub;
        io.new_hb = new_hb;
        io.in_cap = hb->cap;
        CompSubBufs bufs{hb->veff.p, hb->wtr.p, hb->lb.p, hb->rinv.p, hb->ndiv.p, hb->keep.p, hb->kcnt.p, hb->nsub.p,
                         hb->oval[o].p, hb->owidx[o].p, hb->osub[o].p, (unsigned long long)hb->cap,
                         hb->part_d.p, hb->part_c.p, hb->st.p + s, fries_comm_view(hb->comm),
                         fr_bracket_on ? hb->pred.p + s : nullptr, CandList{hb->cand_x.p, hb->cand_m.p, &hb->st.p[s].n_cand}};
        switch (s) {
            case 0: FRIES_TRY(launch_stage<0>(hb, mol, io, bufs, n_samp, u5[0])); break;
            case 1: FRIES_TRY(launch_stage<1>(hb, mol, io, bufs, n_samp, u5[1])); break;
            case 2: FRIES_TRY(launch_stage<2>(hb, mol, io, bufs, n_samp, u5[2])); break;
            case 3: FRIES_TRY(launch_stage<3>(hb, mol, io, bufs, n_samp, u5[3])); break;
            case 4: FRIES_TRY(launch_stage<4>(hb, mol, io, bufs, n_samp, u5[4])); break;
        }
    }
    return FRIES_OK;
}

int fries_hbpp_finalize_dev(fries_hbpp *hb, fries_mol *mol, const uint64_t *d_keys, double p_doub, int new_hb,
                            const HbSpawnArgs *spawn, double cutoff) {
    fries_ctx *c = hb->ctx;
    HbSpawnArgs sp{nullptr, 0, 0, nullptr, nullptr, 1, nullptr, nullptr, nullptr, 0, {nullptr}, 0};
    if (spawn) sp = *spawn;
    size_t smem = (size_t)mol->view.d.blob_doubles * 8;
    int grid = c->sm_count * 4;
    ProfScope ps(c, "hbpp_finalize");
    // stage 4 wrote outputs to buffer set 0 and its path state to set 0
    hbpp_finalize_kernel<<<grid, 256, smem, c->stream>>>(mol->view, d_keys, &hb->st.p[4].n_out,
                                                         (unsigned long long)hb->cap, hb->oval[0].p, hb->owidx[0].p,
                                                         hb->osub[0].p, hb->det[0].p, hb->path[0].p, p_doub, new_hb,
                                                         hb->fin_val.p, hb->fin_det.p, hb->fin_orbs.p, sp, hb->st.p + 5, cutoff);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

extern "C" int fries_apply_hbpp_sys(fries_mol *mol, const uint64_t *h_keys, const double *h_vals, size_t n,
                                    double p_doub, int new_hb, const double *h_uniforms5, unsigned n_samp,
                                    size_t spawn_cap, double *h_out_val, uint64_t *h_out_det, uint8_t *h_out_orbs,
                                    size_t out_cap, size_t *n_out) {
    FRIES_REQUIRE(mol && h_uniforms5 && n_out && (n == 0 || (h_keys && h_vals)), "fries_apply_hbpp_sys: NULL argument");
    FRIES_REQUIRE(n <= spawn_cap, "fries_apply_hbpp_sys: %zu inputs exceed the scratch capacity %zu", n, spawn_cap);
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    *n_out = 0;
    if (n == 0) return FRIES_OK;
    const MolDims &d = mol->view.d;
    uint64_t half = (1ull << d.n_orb) - 1;
    for (size_t i = 0; i < n; i++) {
        uint64_t k = h_keys[i];
        FRIES_REQUIRE((k >> (2 * d.n_orb)) == 0 && (unsigned)__builtin_popcountll(k & half) == d.n_elec / 2 &&
                          (unsigned)__builtin_popcountll(k >> d.n_orb) == d.n_elec / 2,
                      "fries_apply_hbpp_sys: determinant %zu has the wrong electron count", i);
    }
    fries_hbpp *hb = nullptr;
    FRIES_TRY(fries_hbpp_alloc(c, spawn_cap, &hb));
    DevBuf<uint64_t> keys;
    DevBuf<double> vals;
    int rc = keys.alloc(n);
    if (rc == FRIES_OK) rc = vals.alloc(n);
    if (rc != FRIES_OK) {
        fries_hbpp_destroy(hb);
        return rc;
    }
    unsigned long long n64 = n;
    auto fail = [&](int code) {
        fries_hbpp_destroy(hb);
        return code;
    };
#define CU(expr)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            fries_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));                          \
            return fail(FRIES_ERR_CUDA);                                                               \
        }                                                                                              \
    } while (0)
    CU(cudaMemcpyAsync(keys.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(vals.p, h_vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(hb->n_scalar.p, &n64, 8, cudaMemcpyHostToDevice, c->stream));
    for (int rep = 0; rep < fr_debug_repeat && rc == FRIES_OK; rep++) {
        rc = fr_debug_upload_vals(c, vals.p, h_vals, n, rep + 1 < fr_debug_repeat);
        if (rc == FRIES_OK)
            rc = fries_hbpp_stages_dev(hb, mol, keys.p, vals.p, hb->n_scalar.p, p_doub, new_hb, h_uniforms5, n_samp);
    }
    if (rc == FRIES_OK) rc = fries_hbpp_finalize_dev(hb, mol, keys.p, p_doub, new_hb, nullptr);
    if (rc != FRIES_OK) return fail(rc);
    CompState st[6];
    CU(cudaMemcpyAsync(st, hb->st.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    fr_last_fast = 0;
    for (int s = 0; s < 5; s++) fr_last_fast += (int)st[s].fast;
    for (int s = 0; s < 5; s++)
        if (st[s].overflow) {
            // the reference prints "insufficient memory allocated for matrix compression" (:732,768,814,862,913)
            fries_set_error("fries_apply_hbpp_sys: stage %d produced %llu samples beyond spawn_cap %zu", s,
                            st[s].overflow, spawn_cap);
            return fail(FRIES_ERR_CAPACITY);
        }
    size_t m = (size_t)st[5].n_in;
    std::vector<double> v(m ? m : 1);
    std::vector<uint32_t> dd(m ? m : 1), oo(m ? m : 1);
    if (m) {
        CU(cudaMemcpyAsync(v.data(), hb->fin_val.p, m * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(dd.data(), hb->fin_det.p, m * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(oo.data(), hb->fin_orbs.p, m * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
#undef CU
    // marshal the successful samples, in order, into the caller's arrays (num_success compaction :917-991)
    size_t k = 0;
    for (size_t i = 0; i < m; i++) {
        if (v[i] == 0) continue;
        if (k < out_cap) {
            h_out_val[k] = v[i];
            h_out_det[k] = dd[i];
            memcpy(h_out_orbs + 4 * k, &oo[i], 4);
        }
        k++;
    }
    *n_out = k;
    fries_hbpp_destroy(hb);
    if (k > out_cap) {
        fries_set_error("fries_apply_hbpp_sys: %zu samples, output buffers hold %zu", k, out_cap);
        return FRIES_ERR_CAPACITY;
    }
    return FRIES_OK;
}

// ---------------------------------------------------------------------------------------------------
// apply_HBPP_piv (heat_bathPP.cpp:1014-1419, spin_parity = 0): the pivotal twin of apply_HBPP_sys.  Every factor is
// multiplied out into a "long" vector -- one group of entries per input: value / n_div for the uniform rows, value x
// weight_j for the explicit ones, streamed by the same providers as the systematic stages --, compressed in place by
// piv_comp_parallel (piv.cu) and collapsed into the next stage's input list (collapse_long_ :994-1012).  The outputs of
// a stage have the systematic pipeline's format (value, input index, sub index), so the stage providers and the
// finalize kernel are shared.  Plain launches and a one-CTA scan: this is the API-completeness path, not the hot one.
// ---------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256)
hbpp_piv_prep_kernel(MolView gm, HbStageIO io, double *veff, uint32_t *ndiv, uint8_t *nsub, double *rinv, uint32_t *gsize) {
    HbProvider<S> prov;
    prov.m = mol_stage_shared(gm, fr_dyn_smem);
    prov.io = io;
    const size_t n = prov.count(), stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        gsize[i] = hbpp_piv_group_prep(prov, i, veff[i], ndiv[i], nsub[i], rinv[i]);
}

template <int S>
__global__ void __launch_bounds__(256)
hbpp_piv_fill_kernel(MolView gm, HbStageIO io, const double *veff, const uint32_t *ndiv, const uint8_t *nsub,
                     const double *rinv, const unsigned long long *goff, double *lng) {
    HbProvider<S> prov;
    prov.m = mol_stage_shared(gm, fr_dyn_smem);
    prov.io = io;
    const size_t n = prov.count(), stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        hbpp_piv_group_fill(prov, i, veff[i], ndiv[i], nsub[i], rinv[i], lng + goff[i]);
}

// exclusive prefix of `in[0 .. n)` (n on the device) into out, total to *tot; one CTA of 1024 threads
__global__ void __launch_bounds__(1024)
hbpp_piv_scan_kernel(const uint32_t *in, const unsigned long long *n_ptr, unsigned long long n_cap,
                     unsigned long long *out, unsigned long long *tot) {
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    unsigned long long n = *n_ptr;
    if (n > n_cap) n = n_cap;
    unsigned long long carry = 0;
    for (unsigned long long base = 0; base < n; base += blockDim.x) {
        unsigned long long i = base + threadIdx.x;
        unsigned long long c = i < n ? in[i] : 0ull, ec, tc;
        double ex, t;
        block_excl_scan(0.0, c, ex, ec, t, tc, sh_d, sh_c);
        if (i < n) out[i] = carry + ec;
        carry += tc;
    }
    if (threadIdx.x == 0) *tot = carry;
}

// survivors per group (zeroed flag 0), then -- after the scan -- the next stage's input list
__global__ void __launch_bounds__(256)
hbpp_piv_count_kernel(const unsigned long long *n_ptr, unsigned long long n_cap, const uint32_t *gsize,
                      const unsigned long long *goff, const uint8_t *zeroed, uint32_t *cnt) {
    unsigned long long n = *n_ptr;
    if (n > n_cap) n = n_cap;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint8_t *z = zeroed + goff[i];
        uint32_t c = 0;
        for (uint32_t j = 0; j < gsize[i]; j++) c += z[j] ? 0u : 1u;
        cnt[i] = c;
    }
}
__global__ void __launch_bounds__(256)
hbpp_piv_collapse_kernel(const unsigned long long *n_ptr, unsigned long long n_cap, const uint32_t *gsize,
                         const unsigned long long *goff, const uint8_t *zeroed, const double *lng,
                         const unsigned long long *ooff, unsigned long long out_cap, double *out_val, uint32_t *out_widx,
                         uint32_t *out_sub) {
    unsigned long long n = *n_ptr;
    if (n > n_cap) n = n_cap;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long b = goff[i];
        unsigned long long o = ooff[i];
        for (uint32_t j = 0; j < gsize[i]; j++) {
            if (zeroed[b + j]) continue;
            if (o < out_cap) {
                out_val[o] = lng[b + j];
                out_widx[o] = (uint32_t)i;
                out_sub[o] = j;
            }
            o++;
        }
    }
}
__global__ void hbpp_piv_set_nout_kernel(CompState *st, const unsigned long long *tot, unsigned long long cap) {
    unsigned long long t = *tot;
    st->n_out = t < cap ? t : cap;
    st->overflow = t > cap ? t - cap : 0;
}

template <int S>
static int piv_stage(fries_hbpp *hb, fries_mol *mol, HbStageIO &io, int o, double *lng, uint8_t *zeroed, size_t long_cap,
                     uint32_t *gsize, unsigned long long *goff, uint32_t *cnt, unsigned long long *ooff,
                     unsigned long long *tot2, unsigned n_samp, const uint32_t *h_draws, size_t n_draws, size_t *used) {
    fries_ctx *c = hb->ctx;
    const size_t smem = (size_t)mol->view.d.blob_doubles * 8;
    const int grid = c->sm_count * 4;
    const unsigned long long cap = hb->cap;
    hbpp_piv_prep_kernel<S><<<grid, 256, smem, c->stream>>>(mol->view, io, hb->veff.p, hb->ndiv.p, hb->nsub.p, hb->rinv.p, gsize);
    hbpp_piv_scan_kernel<<<1, 1024, 0, c->stream>>>(gsize, io.n_in, cap, goff, tot2);
    c->launch_count += 2;
    unsigned long long n_long = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_long, tot2, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaGetLastError());
    FRIES_REQUIRE(n_long <= long_cap, "fries_apply_hbpp_piv: stage %d expands to %llu entries, the scratch holds %zu", S,
                  n_long, long_cap);
    if (n_long == 0) {  // no input left (every earlier sample was dropped): an empty stage
        CUDA_TRY(cudaMemsetAsync(tot2 + 1, 0, 8, c->stream));
        hbpp_piv_set_nout_kernel<<<1, 1, 0, c->stream>>>(hb->st.p + S, tot2 + 1, cap);
        c->launch_count++;
        CUDA_TRY(cudaGetLastError());
        return FRIES_OK;
    }
    hbpp_piv_fill_kernel<S><<<grid, 256, smem, c->stream>>>(mol->view, io, hb->veff.p, hb->ndiv.p, hb->nsub.p, hb->rinv.p, goff, lng);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    FRIES_TRY(fries_piv_comp_resident(c, lng, (size_t)n_long, n_samp, zeroed, h_draws, n_draws, used));
    hbpp_piv_count_kernel<<<grid, 256, 0, c->stream>>>(io.n_in, cap, gsize, goff, zeroed, cnt);
    hbpp_piv_scan_kernel<<<1, 1024, 0, c->stream>>>(cnt, io.n_in, cap, ooff, tot2 + 1);
    hbpp_piv_collapse_kernel<<<grid, 256, 0, c->stream>>>(io.n_in, cap, gsize, goff, zeroed, lng, ooff, cap, hb->oval[o].p,
                                                          hb->owidx[o].p, hb->osub[o].p);
    hbpp_piv_set_nout_kernel<<<1, 1, 0, c->stream>>>(hb->st.p + S, tot2 + 1, cap);
    c->launch_count += 4;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

extern "C" int fries_apply_hbpp_piv(fries_mol *mol, const uint64_t *h_keys, const double *h_vals, size_t n, double p_doub,
                                    int new_hb, const uint32_t *h_draws, size_t n_draws, size_t *n_draws_used,
                                    unsigned n_samp, size_t spawn_cap, double *h_out_val, uint64_t *h_out_det,
                                    uint8_t *h_out_orbs, size_t out_cap, size_t *n_out) {
    FRIES_REQUIRE(mol && h_draws && n_out && (n == 0 || (h_keys && h_vals)), "fries_apply_hbpp_piv: NULL argument");
    FRIES_REQUIRE(n <= spawn_cap, "fries_apply_hbpp_piv: %zu inputs exceed the scratch capacity %zu", n, spawn_cap);
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    *n_out = 0;
    if (n_draws_used) *n_draws_used = 0;
    if (n == 0) return FRIES_OK;
    const MolDims &d = mol->view.d;
    uint64_t half = (1ull << d.n_orb) - 1;
    for (size_t i = 0; i < n; i++) {
        uint64_t k = h_keys[i];
        FRIES_REQUIRE((k >> (2 * d.n_orb)) == 0 && (unsigned)__builtin_popcountll(k & half) == d.n_elec / 2 &&
                          (unsigned)__builtin_popcountll(k >> d.n_orb) == d.n_elec / 2,
                      "fries_apply_hbpp_piv: determinant %zu has the wrong electron count", i);
    }
    size_t n_states = d.n_elec > d.n_orb - d.n_elec / 2 ? d.n_elec : d.n_orb - d.n_elec / 2;
    if (n_states < d.max_n_symm) n_states = d.max_n_symm;
    if (n_states < 2) n_states = 2;
    const size_t long_cap = spawn_cap * n_states;  // HBCompressPiv::long_vec (heat_bathPP.hpp:303-310)
    struct Guard {
        fries_hbpp *hb = nullptr;
        ~Guard() {
            if (hb) fries_hbpp_destroy(hb);
        }
    } g;
    FRIES_TRY(fries_hbpp_alloc(c, spawn_cap, &g.hb));
    fries_hbpp *hb = g.hb;
    DevBuf<uint64_t> keys;
    DevBuf<double> vals, lng;
    DevBuf<uint8_t> zeroed;
    DevBuf<uint32_t> gsize, cnt;
    DevBuf<unsigned long long> goff, ooff, tot2;
    FRIES_TRY(keys.alloc(n));
    FRIES_TRY(vals.alloc(n));
    FRIES_TRY(lng.alloc(long_cap));
    FRIES_TRY(zeroed.alloc(long_cap));
    FRIES_TRY(gsize.alloc(spawn_cap));
    FRIES_TRY(cnt.alloc(spawn_cap));
    FRIES_TRY(goff.alloc(spawn_cap));
    FRIES_TRY(ooff.alloc(spawn_cap));
    FRIES_TRY(tot2.alloc(2));
    unsigned long long n64 = n;
    CUDA_TRY(cudaMemcpyAsync(keys.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(vals.p, h_vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(hb->n_scalar.p, &n64, 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(hb->st.p, 0, 8 * sizeof(CompState), c->stream));
    size_t used = 0;
    for (int s = 0; s < 5; s++) {
        int o = s & 1, p = o ^ 1;
        HbStageIO io;
        io.keys = keys.p;
        io.vals = vals.p;
        io.n_in = s == 0 ? hb->n_scalar.p : &hb->st.p[s - 1].n_out;
        io.pv = hb->oval[p].p;
        io.pw = hb->owidx[p].p;
        io.ps = hb->osub[p].p;
        io.pdet = hb->det[p].p;
        io.ppath = hb->path[p].p;
        io.det = hb->det[o].p;
        io.path = hb->path[o].p;
        io.p_doub = p_doub;
        io.new_hb = new_hb;
        io.in_cap = hb->cap;
#define PIV_STAGE(S)                                                                                                       \
    FRIES_TRY(piv_stage<S>(hb, mol, io, o, lng.p, zeroed.p, long_cap, gsize.p, goff.p, cnt.p, ooff.p, tot2.p, n_samp, h_draws, \
                           n_draws, &used))
        switch (s) {
            case 0: PIV_STAGE(0); break;
            case 1: PIV_STAGE(1); break;
            case 2: PIV_STAGE(2); break;
            case 3: PIV_STAGE(3); break;
            default: PIV_STAGE(4); break;
        }
#undef PIV_STAGE
    }
    FRIES_TRY(fries_hbpp_finalize_dev(hb, mol, keys.p, p_doub, new_hb, nullptr, 1e-12));  // cutoff of :1409
    CompState st[6];
    CUDA_TRY(cudaMemcpyAsync(st, hb->st.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < 5; s++)
        if (st[s].overflow) {
            fries_set_error("fries_apply_hbpp_piv: stage %d produced %llu samples beyond spawn_cap %zu", s, st[s].overflow,
                            spawn_cap);
            return FRIES_ERR_CAPACITY;
        }
    size_t m = (size_t)st[5].n_in;
    std::vector<double> v(m ? m : 1);
    std::vector<uint32_t> dd(m ? m : 1), oo(m ? m : 1);
    if (m) {
        CUDA_TRY(cudaMemcpyAsync(v.data(), hb->fin_val.p, m * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(dd.data(), hb->fin_det.p, m * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(oo.data(), hb->fin_orbs.p, m * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    size_t k = 0;
    for (size_t i = 0; i < m; i++) {
        if (v[i] == 0) continue;
        if (k < out_cap) {
            h_out_val[k] = v[i];
            h_out_det[k] = dd[i];
            memcpy(h_out_orbs + 4 * k, &oo[i], 4);
        }
        k++;
    }
    *n_out = k;
    if (n_draws_used) *n_draws_used = used;
    if (k > out_cap) {
        fries_set_error("fries_apply_hbpp_piv: %zu samples, output buffers hold %zu", k, out_cap);
        return FRIES_ERR_CAPACITY;
    }
    return FRIES_OK;
}

// Diagnostics: run stages 0..stage and return that stage's raw output list: value, parent index, the
// 4-byte path state of the parent item and the chosen sub-index (what the reference holds in vec1/vec2,
// det_indices*, orb_indices* and comp_idx after the corresponding comp_sub call).
extern "C" int fries_debug_hbpp_stage(fries_mol *mol, const uint64_t *h_keys, const double *h_vals, size_t n,
                                      double p_doub, int new_hb, const double *h_uniforms5, unsigned n_samp,
                                      size_t spawn_cap, int stage, double *h_val, uint32_t *h_det, uint32_t *h_path,
                                      uint32_t *h_sub, size_t *n_out) {
    FRIES_REQUIRE(mol && h_keys && h_vals && h_uniforms5 && n_out && stage >= 0 && stage <= 4, "bad argument");
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    fries_hbpp *hb = nullptr;
    FRIES_TRY(fries_hbpp_alloc(c, spawn_cap, &hb));
    DevBuf<uint64_t> keys;
    DevBuf<double> vals;
    FRIES_TRY(keys.alloc(n));
    FRIES_TRY(vals.alloc(n));
    unsigned long long n64 = n;
    CUDA_TRY(cudaMemcpyAsync(keys.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(vals.p, h_vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(hb->n_scalar.p, &n64, 8, cudaMemcpyHostToDevice, c->stream));
    for (int rep = 0; rep < fr_debug_repeat; rep++) {
    FRIES_TRY(fr_debug_upload_vals(c, vals.p, h_vals, n, rep + 1 < fr_debug_repeat));
    CUDA_TRY(cudaMemsetAsync(hb->st.p, 0, 8 * sizeof(CompState), c->stream));
    for (int s = 0; s <= stage; s++) {
        int o = s & 1, p = o ^ 1;
        HbStageIO io;
        io.keys = keys.p; io.vals = vals.p;
        io.n_in = s == 0 ? hb->n_scalar.p : &hb->st.p[s - 1].n_out;
        io.pv = hb->oval[p].p; io.pw = hb->owidx[p].p; io.ps = hb->osub[p].p;
        io.pdet = hb->det[p].p; io.ppath = hb->path[p].p;
        io.det = hb->det[o].p; io.path = hb->path[o].p;
        io.p_doub = p_doub; io.new_hb = new_hb; io.in_cap = hb->cap;
        CompSubBufs bufs{hb->veff.p, hb->wtr.p, hb->lb.p, hb->rinv.p, hb->ndiv.p, hb->keep.p, hb->kcnt.p, hb->nsub.p,
                         hb->oval[o].p, hb->owidx[o].p, hb->osub[o].p, (unsigned long long)hb->cap,
                         hb->part_d.p, hb->part_c.p, hb->st.p + s, fries_comm_view(hb->comm),
                         fr_bracket_on ? hb->pred.p + s : nullptr, CandList{hb->cand_x.p, hb->cand_m.p, &hb->st.p[s].n_cand}};
        switch (s) {
            case 0: FRIES_TRY(launch_stage<0>(hb, mol, io, bufs, n_samp, h_uniforms5[0])); break;
            case 1: FRIES_TRY(launch_stage<1>(hb, mol, io, bufs, n_samp, h_uniforms5[1])); break;
            case 2: FRIES_TRY(launch_stage<2>(hb, mol, io, bufs, n_samp, h_uniforms5[2])); break;
            case 3: FRIES_TRY(launch_stage<3>(hb, mol, io, bufs, n_samp, h_uniforms5[3])); break;
            case 4: FRIES_TRY(launch_stage<4>(hb, mol, io, bufs, n_samp, h_uniforms5[4])); break;
        }
    }
    }
    CompState s1;
    CUDA_TRY(cudaMemcpyAsync(&s1, hb->st.p + stage, sizeof(s1), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    size_t m = (size_t)s1.n_out;
    *n_out = m;
    int o = stage & 1;
    std::vector<uint32_t> widx(m ? m : 1), det(hb->cap), path(hb->cap);
    if (m) {
        CUDA_TRY(cudaMemcpy(h_val, hb->oval[o].p, m * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(widx.data(), hb->owidx[o].p, m * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_sub, hb->osub[o].p, m * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(det.data(), hb->det[o].p, hb->cap * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(path.data(), hb->path[o].p, hb->cap * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < m; i++) {
            h_det[i] = det[widx[i]];
            h_path[i] = path[widx[i]];
        }
    }
    fries_hbpp_destroy(hb);
    return FRIES_OK;
}

// Diagnostics: the CompState records of the last iteration -- per state 20 doubles:
// loc_norm, glob_norm, new_norm, n_samp_left, rounds, n_kept, n_out, n_in, anomalies, n_cand, fast, overflow,
// then 8 phase time stamps in ns relative to ts[0] (comp_sub_engine: 1 prep, 2 preserved set, 3 line scan,
// 4 count + offsets, 5 emit).  States 0-4: HB-PP stages,
// 5: finalize, 6: find_preserve, 7: sys_comp.
// Diagnostics: time stamps inside the distributed candidate rounds of state s (CompState::rts), ns relative to the first
extern "C" int fries_hbpp_round_stamps(fries_hbpp *hb, int s, double *h_out16) {
    FRIES_REQUIRE(hb && h_out16 && s >= 0 && s < 8, "fries_hbpp_round_stamps: bad argument");
    fries_ctx *c = hb->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    CompState st;
    CUDA_TRY(cudaMemcpyAsync(&st, hb->st.p + s, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 16; k++) h_out16[k] = st.rts[k] >= st.rts[0] ? (double)(st.rts[k] - st.rts[0]) : -1.0;
    return FRIES_OK;
}

// Diagnostics: the timeline of thread 0 of CTA 0 through state s's last run of comp_sub_engine2 (CompState::tl, SM cycles
// relative to mark 0; -1 = mark not passed)
extern "C" int fries_hbpp_timeline(fries_hbpp *hb, int s, double *h_out48) {
    FRIES_REQUIRE(hb && h_out48 && s >= 0 && s < 8, "fries_hbpp_timeline: bad argument");
    fries_ctx *c = hb->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    CompState st;
    CUDA_TRY(cudaMemcpyAsync(&st, hb->st.p + s, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 48; k++) h_out48[k] = st.tl[k] >= st.tl[0] && st.tl[k] != 0 ? (double)(st.tl[k] - st.tl[0]) : -1.0;
    return FRIES_OK;
}

// Diagnostics (FRIES_CTA_MARKS=1): %globaltimer of every CTA at the phase ends of stage s's last run, [8][grid] in ns
// relative to the earliest CTA's first mark
extern "C" int fries_hbpp_cta_marks(fries_hbpp *hb, int s, double *h_out, int *grid) {
    FRIES_REQUIRE(hb && h_out && grid && s >= 0 && s < 7, "fries_hbpp_cta_marks: bad argument");  // 5, 6: the fused vector kernel (marks 0-7, 8-15)
    FRIES_REQUIRE(hb->cta_marks.p && hb->grid2 > 0, "fries_hbpp_cta_marks: set FRIES_CTA_MARKS=1 before the first iteration");
    fries_ctx *c = hb->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    const int g = s >= 5 ? hb->grid_vp : hb->grid2_s[s];
    FRIES_REQUIRE(g > 0 && g <= 1024, "fries_hbpp_cta_marks: that kernel has not run");
    std::vector<unsigned long long> m((size_t)8 * g);
    CUDA_TRY(cudaMemcpyAsync(m.data(), hb->cta_marks.p + (size_t)s * 8 * 1024, m.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    unsigned long long t0 = ~0ull;
    if (s == 6) {  // the second block of the vector kernel's marks counts from the same origin as the first
        std::vector<unsigned long long> m0((size_t)g);
        CUDA_TRY(cudaMemcpyAsync(m0.data(), hb->cta_marks.p + (size_t)5 * 8 * 1024, m0.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int q = 0; q < g; q++) t0 = m0[q] < t0 ? m0[q] : t0;
    } else
        for (int q = 0; q < g; q++) t0 = m[q] < t0 ? m[q] : t0;
    for (size_t q = 0; q < m.size(); q++) h_out[q] = m[q] >= t0 ? (double)(m[q] - t0) : -1.0;
    *grid = g;
    return FRIES_OK;
}

extern "C" int fries_hbpp_states(fries_hbpp *hb, double *h_out64) {
    FRIES_REQUIRE(hb && h_out64, "fries_hbpp_states: NULL argument");
    fries_ctx *c = hb->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    CompState st[8];
    CUDA_TRY(cudaMemcpyAsync(st, hb->st.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < 8; s++) {
        double *o = h_out64 + 20 * s;
        o[0] = st[s].loc_norm; o[1] = st[s].glob_norm; o[2] = st[s].new_norm; o[3] = st[s].n_samp_left;
        o[4] = st[s].rounds; o[5] = (double)st[s].n_kept; o[6] = (double)st[s].n_out; o[7] = (double)st[s].n_in;
        o[8] = (double)st[s].anomalies; o[9] = (double)st[s].n_cand; o[10] = (double)st[s].fast;
        o[11] = (double)st[s].overflow;
        for (int k = 0; k < 8; k++) o[12 + k] = st[s].ts[k] >= st[s].ts[0] ? (double)(st[s].ts[k] - st[s].ts[0]) : 0.0;
    }
    return FRIES_OK;
}
