// TEST INFRASTRUCTURE ONLY -- never loaded by the product.
//
// This container has no GPU.  The __host__ __device__ inline functions of fries_b200/csrc/mol.cuh and
// common.cuh (the pure arithmetic the kernels are made of) are compiled here FOR THE HOST so that the
// CPU-only test tier can check them against the compiled reference before GPU time is spent.  The
// kernels themselves (launch geometry, scans, atomics, staging) are only checked by the -m gpu tests.
#include "../../fries_b200/csrc/mol.cuh"
#include "../../fries_b200/csrc/piv.cuh"
#include "../../fries_b200/csrc/hbpp_prov.cuh"
#include "../../fries_b200/csrc/hh_prov.cuh"
#include "../../fries_b200/csrc/hv_prov.cuh"
#include <vector>

struct HcMol {
    MolView v;
    std::vector<double> eris, hcore, blob;
};

extern "C" {

// tables (hb_info) are supplied by the caller (the GPU path builds them with its own kernels)
void *hc_mol_create(unsigned n_orb, unsigned n_elec_total, unsigned n_frz, const double *hcore, const double *eris_packed,
                    size_t n_packed, const uint8_t *symm, const double *d_diff, const double *d_same,
                    const double *s_tens, double s_norm, const double *exch_sqrt, const double *diag_sqrt,
                    const double *exch_norms) {
    HcMol *h = new HcMol();
    unsigned M = n_orb, T = n_orb + n_frz / 2, TT = M * (M - 1) / 2;
    MolDims &d = h->v.d;
    d.n_orb = M; d.n_elec = n_elec_total - n_frz; d.n_frz = n_frz; d.tot_orb = T; d.s_norm = s_norm;
    const unsigned off = mol_blob_layout(d);
    h->blob.assign(off, 0.0);
    {
        uint32_t *irr = (uint32_t *)&h->blob[d.off_irr];  // as fries_mol_create (mol.cu)
        for (unsigned i = 0; i < M; i++) irr[symm[i] % FR_N_IRREPS] |= 1u << i;
    }
    memcpy(&h->blob[d.off_d_diff], d_diff, M * M * 8);
    memcpy(&h->blob[d.off_d_same], d_same, TT * 8);
    memcpy(&h->blob[d.off_s_tens], s_tens, M * 8);
    memcpy(&h->blob[d.off_exch_sqrt], exch_sqrt, TT * 8);
    memcpy(&h->blob[d.off_diag_sqrt], diag_sqrt, M * 8);
    memcpy(&h->blob[d.off_exch_norms], exch_norms, M * 8);
    for (unsigned i = 0; i < M; i++)
        for (unsigned j = 0; j < M; j++) mol_square_entry(d, h->blob.data(), i, j);
    uint8_t *sy = (uint8_t *)&h->blob[d.off_symm], *lk = (uint8_t *)&h->blob[d.off_lookup];
    memcpy(sy, symm, M);
    d.max_n_symm = 0;
    for (unsigned i = 0; i < M; i++) {
        uint8_t s = symm[i], c = lk[s * (M + 1)];
        lk[s * (M + 1) + 1 + c] = (uint8_t)i;
        lk[s * (M + 1)] = c + 1;
        if ((unsigned)c + 1 > d.max_n_symm) d.max_n_symm = c + 1;
    }
    h->eris.assign(eris_packed, eris_packed + n_packed);
    h->hcore.assign(hcore, hcore + (size_t)T * T);
    h->v.eris = h->eris.data();
    h->v.hcore = h->hcore.data();
    mol_bind_blob(h->v, h->blob.data());
    return h;
}
void hc_mol_destroy(void *p) { delete (HcMol *)p; }

uint64_t hc_hash(uint64_t key, const uint32_t *scr) { return fr_det_hash(key, scr); }
int hc_bit_op(int op, uint64_t *key, const uint8_t *o) {
    switch (op) {
        case 0: return fr_sing_det_parity(*key, o[0], o[1]);
        case 1: return fr_doub_det_parity(*key, o[0], o[1], o[2], o[3]);
        case 2: return fr_sing_parity(*key, o[0], o[1]);
        case 3: return fr_doub_parity(*key, o[0], o[1], o[2], o[3]);
        case 4: return fr_bits_between(*key, o[0], o[1]);
    }
    return 0;
}
double hc_diag(void *p, uint64_t key) {
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    return mol_diag(((HcMol *)p)->v, occ);
}
double hc_sing_el(void *p, uint64_t key, const uint8_t *o) {
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    return mol_sing_el(((HcMol *)p)->v, o[0], o[1], occ);
}
double hc_doub_el(void *p, const uint8_t *o) { return mol_doub_el(((HcMol *)p)->v, o); }
size_t hc_sing_ex(void *p, uint64_t key, uint8_t *out) {
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    size_t n = 0;
    mol_for_each_sing(((HcMol *)p)->v, key, occ, [&](unsigned a, unsigned b) { out[2 * n] = a; out[2 * n + 1] = b; n++; });
    return n;
}
size_t hc_doub_ex(void *p, uint64_t key, uint8_t *out) {
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    size_t n = 0;
    mol_for_each_doub(((HcMol *)p)->v, key, occ, [&](unsigned a, unsigned b, unsigned k, unsigned l) {
        out[4 * n] = a; out[4 * n + 1] = b; out[4 * n + 2] = k; out[4 * n + 3] = l; n++;
    });
    return n;
}
size_t hc_count_singex(void *p, uint64_t key) {
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    return mol_count_singex(((HcMol *)p)->v, key, occ);
}
int hc_find_nth_virt(const uint8_t *occ, int spin, int n_elec, int n_orb, int n) {
    return mol_find_nth_virt(occ, spin, n_elec, n_orb, n);
}
double hc_hb_row(void *p, int which, uint64_t key, int a0, int a1, int a2, double *row, int *len) {
    const MolView &m = ((HcMol *)p)->v;
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    unsigned L = 0;
    double r = 0;
    switch (which) {
        case 0: r = hb_o1_probs(m, row, occ, a0); L = m.d.n_elec - (a0 > 0); break;
        case 1: r = hb_o2_probs(m, row, occ, a0); L = m.d.n_elec; break;
        case 2: r = hb_o2_probs_half(m, row, occ, a0); L = a0; break;
        case 3: r = hb_u1_probs(m, row, a0, occ, a1); L = m.d.n_orb - m.d.n_elec / 2; break;
        case 4: r = hb_u2_probs(m, row, a0, a1, a2, &L); break;
        case 5: r = hb_u2_probs_half(m, row, a0, a1, a2, key, &L); break;
    }
    *len = (int)L;
    return r;
}
double hc_hb_wt(void *p, int normalized, uint64_t key, const uint8_t *orbs) {
    const MolView &m = ((HcMol *)p)->v;
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    return normalized ? hb_norm_wt(m, orbs, occ, key) : hb_unnorm_wt(m, orbs);
}
// symmetry counters for singles: returns n_occ allowed; *n_virt for the occ_choice-th allowed electron
unsigned hc_sing_counts(void *p, uint64_t key, unsigned occ_choice, unsigned *elec_idx, unsigned *n_virt) {
    const MolView &m = ((HcMol *)p)->v;
    uint8_t occ[FRIES_MAX_ELEC + 1], cnt[FR_N_IRREPS][2];
    mol_occ_list(key, occ);
    mol_count_symm_virt(m, occ, cnt);
    uint8_t ch = (uint8_t)occ_choice;
    *n_virt = mol_count_sing_virt(m, occ, cnt, &ch);
    *elec_idx = ch;
    return mol_count_sing_allowed(m, occ, cnt);
}

// ---- pivotal family (piv.cuh): the kernels' phases of piv.cu, run one after the other on the host with a sequential
// prefix sum (the device sums in chunks: same values up to the last bit) --------------------------------------------------
size_t hc_piv_samp(double *vals, size_t n, uint8_t *keep, double seg_norm, uint32_t n_samp, const uint32_t *draws,
                   uint64_t *anomalies) {
    const PivGrid g = piv_grid(seg_norm, n_samp);
    uint32_t n_units = 0;
    *anomalies = 0;
    if (n_samp > 0 && n > 0) {
        std::vector<double> E(n);
        double run = 0;
        for (size_t i = 0; i < n; i++) E[i] = run += keep[i] ? 0.0 : fabs(vals[i]);
        std::vector<uint32_t> cross(n_samp, PIV_NONE), sample(n_samp, PIV_NONE), carry(n_samp, PIV_NONE);
        for (size_t i = 0; i < n; i++)
            piv_mark_cross(g, i ? E[i - 1] : 0.0, E[i], (uint32_t)i, [&](uint32_t k, uint32_t el) { cross[k] = el; });
        uint32_t n_crossed = g.borders_le(E[n - 1]);
        auto loadE = [&](uint32_t j) { return E[j]; };
        auto loadC = [&](uint32_t k) {
            uint32_t c = cross[k];
            if (c == PIV_NONE) c = piv_search(loadE, 0u, (uint32_t)(n - 1), g.border((uint64_t)k + 1));
            return c;
        };
        n_units = piv_n_units(g, n_crossed, n_crossed ? loadC(n_crossed - 1) : 0, n);
        for (uint32_t k = 0; k < n_units; k++)
            piv_unit(g, k, n_crossed, n, loadE, loadC, draws[2 * k] / 4294967296.0, draws[2 * k + 1] / 4294967296.0,
                     sample[k], carry[k]);
        for (uint32_t k = 0; k < n_units; k++) {
            uint32_t s = sample[k];
            if (s == PIV_CARRIED) s = piv_resolve(k, [&](uint32_t j) { return carry[j]; });
            if (s < n && keep[s] != 1) keep[s] = 2;
            else (*anomalies)++;
        }
    }
    for (size_t i = 0; i < n; i++) piv_finish(g.unit, n_samp, vals[i], keep[i]);
    return 2 * (size_t)n_units;
}

double hc_adjust_probs(double *vals, size_t n, uint8_t *keep, uint32_t *n_loc, double exp_loc, uint32_t n_tot,
                       double tot_norm) {
    const PivAdjust a = piv_adjust_setup(*n_loc, exp_loc, n_tot, tot_norm);
    const double thresh = a.loc_norm / ceil(exp_loc);
    bool big = false;
    for (size_t i = 0; i < n; i++)
        if (!keep[i] && fabs(vals[i]) >= thresh) big = true;
    if (!big) return a.loc_norm;
    double g = a.g0();
    uint32_t exact_cnt = 0;
    for (size_t i = 0; i < n; i++) {
        if (keep[i]) continue;
        double dg;
        unsigned long long dk;
        a.delta(fabs(vals[i]), dg, dk);
        double g_before = g, g_after = g + dg;
        g = g_after;
        if (!a.reached(g_before)) continue;
        bool exact;
        double v = vals[i], nv = a.apply(v, exact);
        if (a.last(g_after)) nv = fma((v > 0 ? 1.0 : -1.0) * a.unit, -g_after, nv);
        vals[i] = nv;
        if (exact) {
            keep[i] = 1;
            exact_cnt++;
        }
    }
    *n_loc -= exact_cnt;
    return *n_loc * a.loc_norm / exp_loc;
}

// ---- apply_HBPP_piv: the device pipeline of fries_apply_hbpp_piv (hbpp.cu) step by step on the host.  The per-input
// arithmetic is the product's (hbpp_prov.cuh: providers, group prep / fill, finalize of a sample); scans, collapse and
// buffer ping-pong are restated here as the kernels do them; the pivotal compression between expand and collapse is
// done by the caller (the oracle's piv_comp_parallel in tests/test_hostcheck_hbpp_piv.py). ----
struct HcPiv {
    HcMol *mol;
    std::vector<uint64_t> keys;
    std::vector<double> vals;
    size_t cap;
    std::vector<double> oval[2], veff, rinv;
    std::vector<uint32_t> owidx[2], osub[2], det[2], path[2], ndiv, gsize;
    std::vector<uint8_t> nsub;
    std::vector<unsigned long long> goff;
    unsigned long long n_in;
    double p_doub;
    int new_hb;
};
void *hc_hbpiv_begin(void *mol, const uint64_t *keys, const double *vals, size_t n, double p_doub, int new_hb, size_t cap) {
    HcPiv *h = new HcPiv();
    h->mol = (HcMol *)mol;
    h->keys.assign(keys, keys + n);
    h->vals.assign(vals, vals + n);
    h->cap = cap;
    for (int k = 0; k < 2; k++) {
        h->oval[k].assign(cap, 0.0);
        h->owidx[k].assign(cap, 0);
        h->osub[k].assign(cap, 0);
        h->det[k].assign(cap, 0);
        h->path[k].assign(cap, 0);
    }
    h->veff.assign(cap, 0.0);
    h->rinv.assign(cap, 0.0);
    h->ndiv.assign(cap, 0);
    h->gsize.assign(cap, 0);
    h->nsub.assign(cap, 0);
    h->goff.assign(cap, 0);
    h->n_in = n;
    h->p_doub = p_doub;
    h->new_hb = new_hb;
    return h;
}
void hc_hbpiv_end(void *p) { delete (HcPiv *)p; }

static HbStageIO hbpiv_io(HcPiv *h, int s) {
    int o = s & 1, p = o ^ 1;  // fries_apply_hbpp_piv's ping-pong
    HbStageIO io;
    io.keys = h->keys.data();
    io.vals = h->vals.data();
    io.n_in = &h->n_in;
    io.pv = h->oval[p].data();
    io.pw = h->owidx[p].data();
    io.ps = h->osub[p].data();
    io.pdet = h->det[p].data();
    io.ppath = h->path[p].data();
    io.det = h->det[o].data();
    io.path = h->path[o].data();
    io.p_doub = h->p_doub;
    io.new_hb = h->new_hb;
    io.in_cap = h->cap;
    return io;
}
}  // extern "C"
template <int S>
static size_t hbpiv_expand(HcPiv *h, double *lng, size_t long_cap) {
    HbProvider<S> prov;
    prov.m = h->mol->v;
    prov.io = hbpiv_io(h, S);
    const size_t n = prov.count();
    unsigned long long tot = 0;
    for (size_t i = 0; i < n; i++) {  // prep kernel + scan kernel
        h->gsize[i] = hbpp_piv_group_prep(prov, i, h->veff[i], h->ndiv[i], h->nsub[i], h->rinv[i]);
        h->goff[i] = tot;
        tot += h->gsize[i];
    }
    if (tot > long_cap) return (size_t)-1;
    for (size_t i = 0; i < n; i++)  // fill kernel
        hbpp_piv_group_fill(prov, i, h->veff[i], h->ndiv[i], h->nsub[i], h->rinv[i], lng + h->goff[i]);
    return (size_t)tot;
}
extern "C" {
size_t hc_hbpiv_expand(void *p, int stage, double *lng, size_t long_cap) {
    HcPiv *h = (HcPiv *)p;
    switch (stage) {
        case 0: return hbpiv_expand<0>(h, lng, long_cap);
        case 1: return hbpiv_expand<1>(h, lng, long_cap);
        case 2: return hbpiv_expand<2>(h, lng, long_cap);
        case 3: return hbpiv_expand<3>(h, lng, long_cap);
        default: return hbpiv_expand<4>(h, lng, long_cap);
    }
}
// count / scan / collapse kernels + set_nout; returns the uncapped number of survivors
size_t hc_hbpiv_collapse(void *p, int stage, const double *lng, const uint8_t *zeroed) {
    HcPiv *h = (HcPiv *)p;
    const int o = stage & 1;
    unsigned long long n = h->n_in < h->cap ? h->n_in : h->cap, out = 0;
    for (unsigned long long i = 0; i < n; i++) {
        const unsigned long long b = h->goff[i];
        for (uint32_t j = 0; j < h->gsize[i]; j++) {
            if (zeroed[b + j]) continue;
            if (out < h->cap) {
                h->oval[o][out] = lng[b + j];
                h->owidx[o][out] = (uint32_t)i;
                h->osub[o][out] = j;
            }
            out++;
        }
    }
    h->n_in = out < h->cap ? out : h->cap;
    return (size_t)out;
}
// hbpp_finalize_kernel without spawn arguments + the host marshalling of fries_apply_hbpp_piv
size_t hc_hbpiv_finalize(void *p, double cutoff, double *out_val, uint64_t *out_det, uint8_t *out_orbs) {
    HcPiv *h = (HcPiv *)p;
    size_t k = 0;
    for (unsigned long long i = 0; i < h->n_in; i++) {
        const uint32_t widx = h->owidx[0][i], sub = h->osub[0][i];
        const uint32_t d = h->det[0][widx], pp = h->path[0][widx];
        uint8_t orbs[4];
        bool is_doub;
        double el = hbpp_finalize_sample(h->mol->v, h->keys[d], pp, sub, h->oval[0][i], h->p_doub, h->new_hb, cutoff, orbs,
                                         is_doub);
        if (el == 0) continue;
        out_val[k] = el;
        out_det[k] = d;
        memcpy(out_orbs + 4 * k, orbs, 4);
        k++;
    }
    return k;
}

// ---- apply_HBPP_sys: the stage providers in the systematic pipeline.  rows: values, n_div, the sub-weight matrix
// (n_in x cols, the row of a uniform input is unused) and the row lengths, i.e. what the reference hands to comp_sub
// (heat_bathPP.cpp:713-915); accept: comp_sub's output (value, input, sub index) becomes the next stage's input. ----
}  // extern "C"
template <int S>
static size_t hbsys_rows(HcPiv *h, double *values, uint32_t *ndiv, double *subwts, size_t cols, uint16_t *nsub) {
    HbProvider<S> prov;
    prov.m = h->mol->v;
    prov.io = hbpiv_io(h, S);
    const size_t n = prov.count();
    for (size_t i = 0; i < n; i++) {
        double v, ri, wmax;
        uint32_t nd, ns;
        prov.prep(i, v, nd, ns, ri, wmax);
        values[i] = v;
        ndiv[i] = nd;
        nsub[i] = (uint16_t)ns;
        for (size_t j = 0; j < cols; j++) subwts[i * cols + j] = 0;
        if (nd == 0) {
            double mx = 0;
            prov.visit(i, ri, [&](unsigned j, double w) {
                if (j < cols) subwts[i * cols + j] = w;
                if (w > mx) mx = w;
            });
            if (mx > wmax * (1 + 1e-12)) return (size_t)-1;  // wmax must bound the row (the engine skips rows by it)
        }
    }
    return n;
}
extern "C" {
size_t hc_hbsys_rows(void *p, int stage, double *values, uint32_t *ndiv, double *subwts, size_t cols, uint16_t *nsub) {
    HcPiv *h = (HcPiv *)p;
    switch (stage) {
        case 0: return hbsys_rows<0>(h, values, ndiv, subwts, cols, nsub);
        case 1: return hbsys_rows<1>(h, values, ndiv, subwts, cols, nsub);
        case 2: return hbsys_rows<2>(h, values, ndiv, subwts, cols, nsub);
        case 3: return hbsys_rows<3>(h, values, ndiv, subwts, cols, nsub);
        default: return hbsys_rows<4>(h, values, ndiv, subwts, cols, nsub);
    }
}
void hc_hbsys_accept(void *p, int stage, const double *new_vals, const uint64_t *new_idx, size_t n_out) {
    HcPiv *h = (HcPiv *)p;
    const int o = stage & 1;
    if (n_out > h->cap) n_out = h->cap;
    for (size_t k = 0; k < n_out; k++) {
        h->oval[o][k] = new_vals[k];
        h->owidx[o][k] = (uint32_t)new_idx[2 * k];
        h->osub[o][k] = (uint32_t)new_idx[2 * k + 1];
    }
    h->n_in = n_out;
}

// ---- a19: Hubbard-Holstein arithmetic (hh_prov.cuh) ----
unsigned hc_hh_hub_diag(uint64_t key, unsigned n_sites) { return hh_hub_diag(key, n_sites); }
void hc_hh_neighbors(uint64_t key, unsigned n_sites, uint64_t *plus, uint64_t *minus) { hh_neighbors(key, n_sites, *plus, *minus); }
double hc_hh_ref_ovlp(const uint64_t *keys, const double *vals, size_t n, uint64_t ref, unsigned n_elec, unsigned n_sites,
                      unsigned ph_bits, double g_over_t) {
    HhDims d{n_sites, n_elec, ph_bits};
    double s = 0;
    for (size_t i = 0; i < n; i++) s += hh_ref_ovlp_term(keys[i], vals[i], ref, d, g_over_t);
    return s;
}
unsigned hc_hh_total_ph(uint64_t key, unsigned n_sites, unsigned n_elec, unsigned ph_bits) {
    HhDims d{n_sites, n_elec, ph_bits};
    return hh_total_ph(key, d);
}

// ---- a17: the spawn loop body of one sample (hbpp_prov.cuh) ----
uint64_t hc_spawn_element(uint64_t key, const uint8_t *orbs4, int is_doub, double el, double parent_val, double eps,
                          double init_thresh, double *add) {
    uint8_t o[4] = {orbs4[0], orbs4[1], orbs4[2], orbs4[3]};
    return hbpp_spawn_element(key, o, is_doub != 0, el, parent_val, eps, init_thresh, *add);
}

// ---- a16: the off-diagonal connections of one parent as the H.v kernels produce them: the 32 lanes' shares one after the
// other (hv_prov.cuh); returns their number, fills out_keys / out_vals up to cap ----
size_t hc_hv_parent(void *p, uint64_t key, double val, double h_fac, uint64_t *out_keys, double *out_vals, size_t cap) {
    const MolView &m = ((HcMol *)p)->v;
    ParentCtx pc;
    parent_ctx(m, key, pc);
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(key, occ);
    size_t n = 0;
    unsigned counted = 0;
    for (unsigned lane = 0; lane < 32; lane++)
        counted += lane_excitations(m, pc, lane, [&](bool dbl, unsigned o0, unsigned o1, unsigned v0, unsigned v1) {
            uint64_t nk;
            double el = hv_connection(m, key, occ, dbl, o0, o1, v0, v1, val, h_fac, nk);
            if (n < cap) {
                out_keys[n] = nk;
                out_vals[n] = el;
            }
            n++;
        });
    return counted == n ? n : (size_t)-1;
}
}
