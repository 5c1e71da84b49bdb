// a19: Hubbard-Holstein pieces and the frisys_hh loop body (FRIES_bin/frisys_hh.cpp:186-368).
// Keys: bits [0, n) spin-up sites, [n, 2n) spin-down sites, then n phonon fields of ph_bits bits (hh_vec.hpp).
#include "hbpp.cuh"
extern int fr_bracket_on;  // hbpp.cu
#include "vec.cuh"
#include "hh_prov.cuh"

int fries_vec_compact_flags_dev(fries_vec *vec, const uint8_t *d_flags);

// ---- providers for the two comp_sub stages (frisys_hh.cpp:187-226) -----------------------------------------------
struct HhStage1 {
    const double *vals;
    const unsigned long long *n_in;
    unsigned long long cap;
    double hub_t, elec_ph;
    __device__ size_t count() const {
        unsigned long long n = *n_in;
        return n < cap ? (size_t)n : (size_t)cap;
    }
    __device__ void prep(size_t i, double &v, uint32_t &nd, uint32_t &ns, double &rinv, double &wmax) const {
        v = fabs(vals[i]);
        nd = v > 0 ? 0u : 1u;
        ns = 2;
        rinv = 1.0;
        wmax = fmax(hub_t, elec_ph);  // the row is not normalised
    }
    template <class F>
    __device__ void visit(size_t, double, F &&f) const {
        if (fr_emit(f, 0u, hub_t)) fr_emit(f, 1u, elec_ph);  // NOT normalised in the reference (:193-194)
    }
};
struct HhStage2 {
    const uint64_t *keys;
    const unsigned long long *n_in;
    unsigned long long cap;
    const double *pv;
    const uint32_t *pw, *ps;
    uint32_t *det, *ph_ex;
    HhDims d;
    double hub_t, elec_ph;
    __device__ size_t count() const {
        unsigned long long n = *n_in;
        return n < cap ? (size_t)n : (size_t)cap;
    }
    __device__ void prep(size_t i, double &v, uint32_t &nd, uint32_t &ns, double &rinv, double &wmax) const {
        rinv = 1.0;
        wmax = fmax(hub_t, elec_ph);
        uint32_t di = pw[i], ex = ps[i];
        det[i] = di;
        ph_ex[i] = ex;
        if (ex) {
            nd = 2 * d.n_elec;
        } else {
            uint64_t plus, minus;
            hh_neighbors(keys[di], d.n_sites, plus, minus);
            nd = fr_popc(plus) + fr_popc(minus);
        }
        v = pv[i] * nd;  // comp_vec2[samp_idx] *= ndiv (:217)
        ns = 2;
    }
    template <class F>
    __device__ void visit(size_t, double, F &&f) const {
        if (fr_emit(f, 0u, hub_t)) fr_emit(f, 1u, elec_ph);  // only reached with ndiv == 0 (value 0)
    }
};

__global__ void __launch_bounds__(FR_COMP_BLOCK) hh_stage1_kernel(HhStage1 prov, CompSubBufs bufs, unsigned n_samp, double rn) {
    comp_sub_engine(prov, bufs, n_samp, rn);
}
__global__ void __launch_bounds__(FR_COMP_BLOCK) hh_stage2_kernel(HhStage2 prov, CompSubBufs bufs, unsigned n_samp, double rn) {
    comp_sub_engine(prov, bufs, n_samp, rn);
}

// spawn loop body frisys_hh.cpp:246-296
__global__ void hh_spawn_kernel(VecView v, HhDims d, const unsigned long long *n_ptr, unsigned long long cap,
                                const double *pv, const uint32_t *pw, const uint32_t *ps, const uint32_t *det,
                                const uint32_t *ph_ex, double eps, double init_thresh, uint64_t *out_keys, double *out_vals,
                                CompState *st) {
    unsigned long long n = *n_ptr;
    if (n > cap) n = cap;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long ok = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t prev = pw[i], exc = ps[i];
        uint32_t di = det[prev];
        double cv = v.vals[di];
        uint64_t key = v.keys[di], nk = key;
        double el = pv[i] * -eps;
        if (cv < 0) el *= -1;
        if (ph_ex[prev]) {
            uint8_t occ[FRIES_MAX_ELEC + 1];
            fr_occ_list(key & ((1ull << (2 * d.n_sites)) - 1), occ);
            unsigned site = occ[exc % d.n_elec] % d.n_sites;
            unsigned ph = hh_phonon(key, d, site);
            unsigned shift = 2 * d.n_sites + site * d.ph_bits;
            if (exc < d.n_elec && ph > 0) {
                nk = key - (1ull << shift);  // det_from_ph(..., -1)
                el *= sqrt((double)ph);
            } else if (exc >= d.n_elec && ph + 1 < (1u << d.ph_bits)) {
                nk = key + (1ull << shift);  // det_from_ph(..., +1)
                el *= sqrt((double)(ph + 1));
            } else {
                el = 0;
            }
        } else {
            uint64_t plus, minus;
            hh_neighbors(key, d.n_sites, plus, minus);
            unsigned n_plus = fr_popc(plus), orig, dest;
            if (exc < n_plus) {
                orig = hh_nth_bit(plus, exc);
                dest = orig + 1;
            } else {
                orig = hh_nth_bit(minus, exc - n_plus);
                dest = orig - 1;
            }
            nk = (key & ~(1ull << orig)) | (1ull << dest);
            el *= -1;  // hub_t
        }
        uint64_t outk = FRIES_EMPTY_KEY;
        if (fabs(el) > 1e-9) {
            outk = nk | (fabs(cv) >= init_thresh ? FRIES_INI_FLAG : 0ull);
            ok++;
        }
        out_keys[i] = outk;
        out_vals[i] = el;
    }
    ok = warp_sum_u64(ok);
    if ((threadIdx.x & 31) == 0 && ok) atomicAdd(&st->n_out, ok);
}

// diagonal step + add_vecs(0, 1) + zero row 1 (frisys_hh.cpp:308-318; row 1 is zeroed at :229-230)
__global__ void hh_diag_kernel(VecView v, HhDims d, double eps, double hub_u, double ph_freq, double hf_en, double shift) {
    unsigned long long n64 = v.cnt->n;
    size_t n = n64 < v.cap ? (size_t)n64 : v.cap;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double *v0 = v.vals, *v1 = v.vals + v.cap;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = v0[i], b = v1[i];
        if (a != 0) {
            uint64_t key = v.keys[i];
            double diag_el = hh_hub_diag(key, d.n_sites);
            double phonon_diag = hh_total_ph(key, d) * ph_freq;
            a *= 1 - eps * (diag_el * hub_u + phonon_diag - hf_en - shift);
        }
        v0[i] = a + b;
        if (b != 0) v1[i] = 0;
    }
}

// numer/denom (frisys_hh.cpp:333-345): one CTA, fixed-order sum
__global__ void __launch_bounds__(1024)
hh_energy_kernel(VecView v, HhDims d, uint64_t ref, double g_over_t, double hub_t, double hub_u, double hf_en, double *out2) {
    __shared__ double sh[34];
    unsigned long long n64 = v.cnt->n;
    size_t n = n64 < v.cap ? (size_t)n64 : v.cap;
    double acc = 0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += hh_ref_ovlp_term(v.keys[i], v.vals[i], ref, d, g_over_t);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        // the reference reads position 0 (the Neel state never leaves it); here it is found by key
        double ref_el = 0;
        for (size_t i = 0; i < n && i < 1; i++)
            if (v.keys[i] == ref) ref_el = v.vals[i];
        double numer = (hh_hub_diag(ref, d.n_sites) * hub_u - hf_en) * ref_el + acc * -hub_t;
        out2[0] = numer;
        out2[1] = ref_el;
    }
}

__global__ void hh_batch_kernel(int what, const uint64_t *keys, const double *vals, size_t n, HhDims d, uint64_t ref,
                                double g_over_t, double *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (what == 0) out[i] = hh_hub_diag(keys[i], d.n_sites);
    else if (what == 1) {
        uint64_t p, m;
        hh_neighbors(keys[i], d.n_sites, p, m);
        out[2 * i] = (double)p;
        out[2 * i + 1] = (double)m;
    } else out[i] = hh_ref_ovlp_term(keys[i], vals[i], ref, d, g_over_t);
}

// ---- C-ABI -----------------------------------------------------------------------------------------------------------
// what 0: hub_diag (hub_holstein.cpp:101-136); 1: neighbour masks (hh_vec.hpp:139-175) -> out[2n] (hop+1, hop-1);
// 2: per-state terms of calc_ref_ovlp (hub_holstein.hpp:93-182)
extern "C" int fries_hh_batch(fries_ctx *c, int what, const uint64_t *h_keys, const double *h_vals, size_t n, unsigned n_sites,
                              unsigned n_elec, unsigned ph_bits, uint64_t ref_key, double g_over_t, double *h_out) {
    FRIES_REQUIRE(c && h_out && (n == 0 || h_keys) && what >= 0 && what <= 2, "fries_hh_batch: bad argument");
    FRIES_REQUIRE(what != 2 || h_vals, "fries_hh_batch: values needed");
    if (n == 0) return FRIES_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> k;
    DevBuf<double> v, o;
    size_t no = what == 1 ? 2 * n : n;
    FRIES_TRY(k.alloc(n));
    FRIES_TRY(v.alloc(n));
    FRIES_TRY(o.alloc(no));
    CUDA_TRY(cudaMemcpyAsync(k.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (h_vals) CUDA_TRY(cudaMemcpyAsync(v.p, h_vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    HhDims d{n_sites, n_elec, ph_bits};
    hh_batch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(what, k.p, v.p, n, d, ref_key, g_over_t, o.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_out, o.p, no * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

extern "C" int fries_frisys_hh_setup(fries_vec *vec, size_t spawn_cap, fries_hbpp **out) {
    FRIES_REQUIRE(vec && out && vec->hh_sites, "fries_frisys_hh_setup: needs a vector made by fries_vec_create_hh");
    FRIES_REQUIRE(vec->n_vecs >= 2, "fries_frisys_hh_setup: the vector needs two value rows");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    fries_hbpp *hb = nullptr;
    FRIES_TRY(fries_hbpp_alloc(c, spawn_cap, &hb));
    int rc = hb->spawn_keys.alloc(spawn_cap);
    if (rc == FRIES_OK) rc = hb->spawn_vals.alloc(spawn_cap);
    if (rc == FRIES_OK) rc = hb->keep_flags.alloc(vec->cap);
    if (rc != FRIES_OK) {
        fries_hbpp_destroy(hb);
        return rc;
    }
    CUDA_TRY(cudaMemsetAsync(hb->keep_flags.p, 0, vec->cap, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    vec->min_del_idx = 1;  // the Neel state at position 0 is never deleted (frisys_hh.cpp:353)
    *out = hb;
    return FRIES_OK;
}

__global__ void hh_stats_kernel(const CompState *st, const VecCounters *cnt, const double *scal, double *out) {
    out[0] = st[6].glob_norm;
    out[1] = scal[4];
    out[2] = scal[5];
    out[3] = (double)st[6].n_kept;
    out[4] = (double)st[5].n_out;
    out[5] = (double)cnt->n;
    out[6] = (double)(st[0].overflow + st[1].overflow);
    out[7] = (double)cnt->overflow;
}
__global__ void hh_state_to_r4(const CompState *st, double *r4) {
    r4[0] = st->loc_norm;
    r4[1] = st->glob_norm;
    r4[2] = (double)st->n_samp_left;
    r4[3] = (double)st->n_kept;
}

extern "C" int fries_frisys_hh_iterate(fries_vec *vec, fries_hbpp *hb, const fries_frisys_hh_params *p,
                                       const double *u3, fries_iter_stats *stats) {
    FRIES_REQUIRE(vec && hb && p && u3 && vec->hh_sites, "fries_frisys_hh_iterate: bad argument");
    FRIES_REQUIRE(vec->n_ranks == 1, "fries_frisys_hh_iterate: single-rank entry point");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    VecView v = vec->view();
    HhDims d{vec->hh_sites, vec->n_elec, vec->hh_ph_bits};
    CUDA_TRY(cudaMemsetAsync(hb->st.p, 0, 8 * sizeof(CompState), c->stream));
    const double hub_t = 1;
    int grid = c->coop_grid((const void *)hh_stage2_kernel, FR_COMP_BLOCK, 0);
    int g1 = c->coop_grid((const void *)hh_stage1_kernel, FR_COMP_BLOCK, 0);
    if (g1 < grid) grid = g1;
    auto bufs_for = [&](int o, int s) {
        return CompSubBufs{hb->veff.p, hb->wtr.p, hb->lb.p, hb->rinv.p, hb->ndiv.p, hb->keep.p, hb->kcnt.p, hb->nsub.p,
                           hb->oval[o].p, hb->owidx[o].p, hb->osub[o].p, (unsigned long long)hb->cap,
                           hb->part_d.p, hb->part_c.p, hb->st.p + s, fries_comm_view(nullptr),
                           fr_bracket_on ? hb->pred.p + s : nullptr, CandList{hb->cand_x.p, hb->cand_m.p, &hb->st.p[s].n_cand}};
    };
    unsigned n_samp = p->target_nonz;
    {   // stage 1: hop vs phonon (:187-206)
        HhStage1 pr{v.vals, &vec->cnt.p->n, (unsigned long long)hb->cap, hub_t, p->elec_ph};
        CompSubBufs b = bufs_for(0, 0);
        double rn = u3[0];
        void *args[] = {(void *)&pr, (void *)&b, (void *)&n_samp, (void *)&rn};
        ProfScope ps(c, "hh_stage1");
        CUDA_TRY(cudaLaunchCooperativeKernel((const void *)hh_stage1_kernel, dim3(grid), dim3(FR_COMP_BLOCK), args, 0, c->stream));
        c->launch_count++;
    }
    {   // stage 2: which neighbour / which phonon move (:208-226)
        HhStage2 pr{v.keys, &hb->st.p[0].n_out, (unsigned long long)hb->cap, hb->oval[0].p, hb->owidx[0].p, hb->osub[0].p,
                    hb->det[0].p, hb->path[0].p, d, hub_t, p->elec_ph};
        CompSubBufs b = bufs_for(1, 1);
        double rn = u3[1];
        void *args[] = {(void *)&pr, (void *)&b, (void *)&n_samp, (void *)&rn};
        ProfScope ps(c, "hh_stage2");
        CUDA_TRY(cudaLaunchCooperativeKernel((const void *)hh_stage2_kernel, dim3(grid), dim3(FR_COMP_BLOCK), args, 0, c->stream));
        c->launch_count++;
    }
    hh_spawn_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(v, d, &hb->st.p[1].n_out, (unsigned long long)hb->cap,
                                                             hb->oval[1].p, hb->owidx[1].p, hb->osub[1].p, hb->det[0].p,
                                                             hb->path[0].p, p->eps, p->init_thresh, hb->spawn_keys.p,
                                                             hb->spawn_vals.p, hb->st.p + 5);
    c->launch_count++;
    FRIES_TRY(fries_vec_merge_dev(vec, hb->spawn_keys.p, hb->spawn_vals.p, hb->cap, &hb->st.p[1].n_out, 0, 1));
    hh_diag_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(v, d, p->eps, p->hub_u, p->ph_freq, p->hf_en, p->en_shift);
    c->launch_count++;
    FRIES_TRY(fries_find_preserve_launch(c, v.vals, vec->cap, &vec->cnt.p->n, p->target_nonz, hb->keep_flags.p, hb->st.p + 6,
                                         hb->part_d.p, hb->part_c.p, 0, nullptr, hb->pred.p + 5, hb->cand_x.p,
                                         hb->cand_m.p));
    hh_state_to_r4<<<1, 1, 0, c->stream>>>(hb->st.p + 6, hb->scal.p);
    hh_energy_kernel<<<1, 1024, 0, c->stream>>>(v, d, p->ref_key, p->elec_ph / hub_t, hub_t, p->hub_u, p->hf_en, hb->scal.p + 4);
    c->launch_count += 2;
    FRIES_TRY(fries_sys_comp_launch(c, v.vals, vec->cap, &vec->cnt.p->n, hb->keep_flags.p, hb->scal.p, 0.0, 0.0, -1LL, u3[2],
                                    hb->st.p + 7, hb->part_d.p, hb->part_c.p, 0, nullptr));
    FRIES_TRY(fries_vec_compact_flags_dev(vec, hb->keep_flags.p));
    hh_stats_kernel<<<1, 1, 0, c->stream>>>(hb->st.p, vec->cnt.p, hb->scal.p, hb->scal.p + 8);
    c->launch_count++;
    CUDA_TRY(cudaMemcpyAsync(c->h_pinned, hb->scal.p + 8, 8 * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const double *h = c->h_pinned;
    if (stats) {
        stats->glob_norm = h[0];
        stats->numer = h[1];
        stats->denom = h[2];
        stats->n_kept = (uint64_t)h[3];
        stats->n_matrix_samples = stats->n_spawned = (uint64_t)h[4];
        stats->curr_size = (uint64_t)h[5];
    }
    if (h[6] != 0) {
        fries_set_error("fries_frisys_hh_iterate: insufficient memory allocated for matrix compression");
        return FRIES_ERR_CAPACITY;
    }
    if (h[7] != 0) {
        fries_set_error("fries_frisys_hh_iterate: determinant store is full (capacity %zu)", vec->cap);
        return FRIES_ERR_CAPACITY;
    }
    return FRIES_OK;
}

// ---- frifull_hh (FRIES_bin/frifull_hh.cpp:186-330): no matrix compression -- every state spawns all of its hops
// (hub_all hub_holstein.cpp:83-98) and phonon moves (:224-257) ----------------------------------------------------------
// One thread per parent; a parent owns `slots` = 4 * n_elec consecutive places of the spawn window (2 hops per electron at
// most, one phonon creation and one annihilation per up electron and per down electron on a singly occupied site), unused
// places carry the empty key.  Elements of value 0 (elec_ph = 0) are not spawned: the reference adds them as zeros.
__global__ void hh_full_spawn_kernel(VecView v, HhDims d, size_t first, size_t count, unsigned slots, double eps, double hub_t,
                                     double elec_ph, double init_thresh, uint64_t *out_keys, double *out_vals,
                                     unsigned long long *n_spawned) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long ok = 0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += stride) {
        const size_t i = first + q;
        uint64_t *ok_ = out_keys + q * slots;
        double *ov = out_vals + q * slots;
        unsigned w = 0;
        const double cv = v.vals[i];
        if (cv != 0) {
            const uint64_t key = v.keys[i];
            const uint64_t ini = fabs(cv) > init_thresh ? FRIES_INI_FLAG : 0ull;
            auto put = [&](uint64_t nk, double el) {
                if (el != 0 && w < slots) {
                    ok_[w] = nk | ini;
                    ov[w] = el;
                    w++;
                }
            };
            uint64_t plus, minus;
            hh_neighbors(key, d.n_sites, plus, minus);
            for (uint64_t mk = plus; mk; mk &= mk - 1) {
                const unsigned orig = (unsigned)fr_ctz(mk);
                put((key & ~(1ull << orig)) | (1ull << (orig + 1)), eps * hub_t * cv);
            }
            for (uint64_t mk = minus; mk; mk &= mk - 1) {
                const unsigned orig = (unsigned)fr_ctz(mk);
                put((key & ~(1ull << orig)) | (1ull << (orig - 1)), eps * hub_t * cv);
            }
            const uint64_t site_mask = (1ull << d.n_sites) - 1;
            const uint64_t up = key & site_mask, dn = (key >> d.n_sites) & site_mask;
            for (uint64_t mk = up; mk; mk &= mk - 1) {  // :224-238
                const unsigned site = (unsigned)fr_ctz(mk);
                const unsigned ph = hh_phonon(key, d, site);
                const int docc = (int)((dn >> site) & 1ull);
                const unsigned shift = 2 * d.n_sites + site * d.ph_bits;
                if (ph > 0) put(key - (1ull << shift), -eps * elec_ph * sqrt((double)ph) * (docc + 1) * cv);
                if (ph + 1 < (1u << d.ph_bits)) put(key + (1ull << shift), -eps * elec_ph * sqrt((double)(ph + 1)) * (docc + 1) * cv);
            }
            for (uint64_t mk = dn & ~up; mk; mk &= mk - 1) {  // :240-256: down electrons on singly occupied sites
                const unsigned site = (unsigned)fr_ctz(mk);
                const unsigned ph = hh_phonon(key, d, site);
                const unsigned shift = 2 * d.n_sites + site * d.ph_bits;
                if (ph > 0) put(key - (1ull << shift), -eps * elec_ph * sqrt((double)ph) * cv);
                if (ph + 1 < (1u << d.ph_bits)) put(key + (1ull << shift), -eps * elec_ph * sqrt((double)(ph + 1)) * cv);
            }
        }
        ok += w;
        for (; w < slots; w++) {
            ok_[w] = FRIES_EMPTY_KEY;
            ov[w] = 0.0;
        }
    }
    ok = warp_sum_u64(ok);
    if ((threadIdx.x & 31) == 0 && ok) atomicAdd(n_spawned, ok);
}

extern "C" int fries_frifull_hh_iterate(fries_vec *vec, fries_hbpp *hb, const fries_frisys_hh_params *p, double uniform,
                                        fries_iter_stats *stats) {
    FRIES_REQUIRE(vec && hb && p && vec->hh_sites, "fries_frifull_hh_iterate: bad argument");
    FRIES_REQUIRE(vec->n_ranks == 1, "fries_frifull_hh_iterate: single-rank entry point");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    HhDims d{vec->hh_sites, vec->n_elec, vec->hh_ph_bits};
    const unsigned slots = 4 * vec->n_elec;
    FRIES_REQUIRE(hb->cap >= slots, "fries_frifull_hh_iterate: spawn window smaller than one state's %u moves", slots);
    CUDA_TRY(cudaMemsetAsync(hb->st.p, 0, 8 * sizeof(CompState), c->stream));
    const double hub_t = 1;
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    const size_t n_parents = (size_t)cnt.n, win = hb->cap / slots;
    // the parents are the states stored before the multiplication (:193-196); what the merges append has value 0 in row 0
    for (size_t first = 0; first < n_parents; first += win) {
        const size_t count = n_parents - first < win ? n_parents - first : win;
        VecView v = vec->view();
        {
            ProfScope ps(c, "hh_full_spawn");
            hh_full_spawn_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(v, d, first, count, slots, p->eps, hub_t, p->elec_ph,
                                                                          p->init_thresh, hb->spawn_keys.p, hb->spawn_vals.p,
                                                                          &hb->st.p[5].n_out);
            c->launch_count++;
        }
        CUDA_TRY(cudaGetLastError());
        FRIES_TRY(fries_vec_merge_dev(vec, hb->spawn_keys.p, hb->spawn_vals.p, count * slots, nullptr, 0, 1));
    }
    VecView v = vec->view();
    hh_diag_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(v, d, p->eps, p->hub_u, p->ph_freq, p->hf_en, p->en_shift);
    c->launch_count++;
    FRIES_TRY(fries_find_preserve_launch(c, v.vals, vec->cap, &vec->cnt.p->n, p->target_nonz, hb->keep_flags.p, hb->st.p + 6,
                                         hb->part_d.p, hb->part_c.p, 0, nullptr, hb->pred.p + 5, hb->cand_x.p,
                                         hb->cand_m.p));
    hh_state_to_r4<<<1, 1, 0, c->stream>>>(hb->st.p + 6, hb->scal.p);
    hh_energy_kernel<<<1, 1024, 0, c->stream>>>(v, d, p->ref_key, p->elec_ph / hub_t, hub_t, p->hub_u, p->hf_en, hb->scal.p + 4);
    c->launch_count += 2;
    FRIES_TRY(fries_sys_comp_launch(c, v.vals, vec->cap, &vec->cnt.p->n, hb->keep_flags.p, hb->scal.p, 0.0, 0.0, -1LL, uniform,
                                    hb->st.p + 7, hb->part_d.p, hb->part_c.p, 0, nullptr));
    FRIES_TRY(fries_vec_compact_flags_dev(vec, hb->keep_flags.p));
    hh_stats_kernel<<<1, 1, 0, c->stream>>>(hb->st.p, vec->cnt.p, hb->scal.p, hb->scal.p + 8);
    c->launch_count++;
    CUDA_TRY(cudaMemcpyAsync(c->h_pinned, hb->scal.p + 8, 8 * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const double *h = c->h_pinned;
    if (stats) {
        stats->glob_norm = h[0];
        stats->numer = h[1];
        stats->denom = h[2];
        stats->n_kept = (uint64_t)h[3];
        stats->n_matrix_samples = stats->n_spawned = (uint64_t)h[4];
        stats->curr_size = (uint64_t)h[5];
    }
    if (h[7] != 0) {
        fries_set_error("fries_frifull_hh_iterate: determinant store is full (capacity %zu)", vec->cap);
        return FRIES_ERR_CAPACITY;
    }
    return FRIES_OK;
}
