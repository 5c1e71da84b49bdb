/* C-ABI shim over the REFERENCE's own C++ API (sgreene8/FRIES), for ctypes.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with the reference's sources
 * (straight from /root/reference, nothing copied) into oracle/_ref/libfries_ref.so.  Every function
 * here only marshals flat arrays into the reference's containers, calls the reference function
 * named in its comment, and marshals the result back.  It is used (a) to pin the plain-C
 * restatement in oracle/fries_oracle.c, (b) to generate tests/golden/, (c) as the strongest checker
 * of the CUDA path and (d) as the timed CPU baseline.  The product never loads it.
 */
#include <cstdint>
#include <cstring>
#include <vector>
#include <random>
#include <functional>
#include <mpi.h>
#include <FRIES/math_utils.h>
#include <FRIES/det_store.h>
#include <FRIES/fci_utils.h>
#include <FRIES/ndarr.hpp>
#include <FRIES/det_hash.hpp>
#include <FRIES/vec_utils.hpp>
#include <FRIES/compress_utils.hpp>
#include <FRIES/Hamiltonians/molecule.hpp>
#include <FRIES/Hamiltonians/heat_bathPP.hpp>
#include <FRIES/Hamiltonians/near_uniform.hpp>
#include <FRIES/Hamiltonians/hub_holstein.hpp>
#include <FRIES/hh_vec.hpp>

namespace {
inline void key_to_bytes(uint64_t key, uint8_t *bytes, size_t n_bytes) {
    for (size_t b = 0; b < n_bytes; b++) bytes[b] = (uint8_t)(key >> (8 * b));
}
inline uint64_t bytes_to_key(const uint8_t *bytes, size_t n_bytes) {
    uint64_t k = 0;
    for (size_t b = 0; b < n_bytes; b++) k |= (uint64_t)bytes[b] << (8 * b);
    return k;
}
}  // namespace

extern "C" {

/* ---- L1 bit utilities ------------------------------------------------------------------- */

/* find_bits, FRIES/math_utils.c:62-98 */
int ref_find_bits(uint64_t key, int n_bytes, uint8_t *out) {
    uint8_t bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    return find_bits(bytes, out, (uint8_t)n_bytes);
}

/* bits_between, FRIES/math_utils.c:9-58 */
unsigned ref_bits_between(uint64_t key, int a, int b) {
    uint8_t bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    return bits_between(bytes, (uint8_t)a, (uint8_t)b);
}

/* gen_hf_bitstring, FRIES/fci_utils.c:10-43 */
uint64_t ref_gen_hf_bitstring(unsigned n_orb, unsigned n_elec) {
    uint8_t bytes[16] = {0};
    gen_hf_bitstring(n_orb, n_elec, bytes);
    return bytes_to_key(bytes, 8);
}

/* sing_det_parity / doub_det_parity, FRIES/fci_utils.c:46-75: returns sign, *key updated */
int ref_sing_det_parity(uint64_t *key, const uint8_t *orbs) {
    uint8_t bytes[16] = {0}, o[2] = {orbs[0], orbs[1]};
    key_to_bytes(*key, bytes, 8);
    int s = sing_det_parity(bytes, o);
    *key = bytes_to_key(bytes, 8);
    return s;
}
int ref_doub_det_parity(uint64_t *key, const uint8_t *orbs) {
    uint8_t bytes[16] = {0}, o[4] = {orbs[0], orbs[1], orbs[2], orbs[3]};
    key_to_bytes(*key, bytes, 8);
    int s = doub_det_parity(bytes, o);
    *key = bytes_to_key(bytes, 8);
    return s;
}
/* sing_parity / doub_parity, FRIES/fci_utils.c:54-57,86-94 (determinant unchanged) */
int ref_sing_parity(uint64_t key, const uint8_t *orbs) {
    uint8_t bytes[16] = {0}, o[2] = {orbs[0], orbs[1]};
    key_to_bytes(key, bytes, 8);
    return sing_parity(bytes, o);
}
int ref_doub_parity(uint64_t key, const uint8_t *orbs) {
    uint8_t bytes[16] = {0}, o[4] = {orbs[0], orbs[1], orbs[2], orbs[3]};
    key_to_bytes(key, bytes, 8);
    return doub_parity(bytes, o);
}
/* find_nth_virt, FRIES/fci_utils.c:138-148 */
int ref_find_nth_virt(const uint8_t *occ, int spin, int n_elec, int n_orb, int n) {
    std::vector<uint8_t> o(occ, occ + n_elec);
    o.push_back(255);
    return find_nth_virt(o.data(), spin, (uint8_t)n_elec, (uint8_t)n_orb, (uint8_t)n);
}

/* ---- a1: hash + owner --------------------------------------------------------------------- */

/* HashTable::hash_fxn det_hash.hpp:160-170 via DistVec::idx_to_hash / idx_to_proc semantics
 * (vec_utils.hpp:360-400): occupied list from find_bits, owner = hash % n_procs. */
void ref_hash_keys(const uint64_t *keys, size_t n, int n_bits, const uint32_t *scrambler,
                   int n_procs, uint64_t *hash_out, int32_t *owner_out) {
    std::vector<uint32_t> scr(scrambler, scrambler + n_bits);
    HashTable<ssize_t> ht(0, scr);
    int n_bytes = CEILING(n_bits, 8);
    for (size_t i = 0; i < n; i++) {
        uint8_t bytes[16] = {0}, occ[64];
        key_to_bytes(keys[i], bytes, 8);
        uint8_t n_elec = find_bits(bytes, occ, (uint8_t)n_bytes);
        uintmax_t h = ht.hash_fxn(occ, n_elec, NULL, 0);
        if (hash_out) hash_out[i] = (uint64_t)h;
        if (owner_out) owner_out[i] = (int32_t)(h % (uintmax_t)n_procs);
    }
}

/* Attach this process to the multi-process communicator of oracle/mpi_shim (FRIES_SHIM_* in the environment, set by
 * shimrun.py): afterwards the reference's collectives inside find_preserve / sys_comp / piv_comp_parallel are real. */
int ref_mpi_init(void) {
    MPI_Init(nullptr, nullptr);
    int n = 1;
    MPI_Comm_size(MPI_COMM_WORLD, &n);
    return n;
}
int ref_mpi_rank(void) {
    int r = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &r);
    return r;
}

/* MPI_Allgather(MPI_IN_PLACE) of `count` doubles per rank, as the drivers do for loc_norms (frisys_mol.cpp:532) */
void ref_allgather_doubles(double *buf, int count) {
    MPI_Allgather(MPI_IN_PLACE, 0, MPI_DOUBLE, buf, count, MPI_DOUBLE, MPI_COMM_WORLD);
}

/* ---- a4/a5: vector compression -------------------------------------------------------------- */

/* find_preserve, FRIES/compress_utils.cpp:29-105.  keep_out[i] in {0,1}. Returns the local
 * residual one-norm. */
double ref_find_preserve(const double *values, size_t count, unsigned *n_samp, double *glob_norm,
                         uint8_t *keep_out) {
    std::vector<size_t> srt(count);
    std::vector<bool> keep(count, false);
    std::vector<double> v(values, values + count);
    double r = find_preserve(v.data(), srt, keep, count, n_samp, glob_norm);
    for (size_t i = 0; i < count; i++) keep_out[i] = keep[i];
    return r;
}

/* sys_comp, FRIES/compress_utils.cpp:278-327 (single rank: loc_norms has 1 entry).
 * keep[i] in: preserve exactly; out: 1 = element was zeroed ("delete me"). */
void ref_sys_comp(double *values, size_t count, double *loc_norm, unsigned n_samp, uint8_t *keep,
                  double rn) {
    std::vector<bool> k(count);
    for (size_t i = 0; i < count; i++) k[i] = keep[i] != 0;
    sys_comp(values, count, loc_norm, n_samp, k, rn);
    for (size_t i = 0; i < count; i++) keep[i] = k[i];
}

/* ---- pivotal family ---------------------------------------------------------------------------
 * The reference draws from a std::mt19937&; the bindings seed one, discard `skip` outputs and report how many
 * outputs the call consumed (by stepping a copy of the generator until the states agree). */
namespace {
size_t draws_between(std::mt19937 before, const std::mt19937 &after, size_t limit) {
    size_t k = 0;
    while (!(before == after) && k < limit) {
        before();
        k++;
    }
    return k;
}
}  // namespace

/* first n outputs of std::mt19937(seed) */
void ref_mt19937_fill(uint32_t seed, size_t n, uint32_t *out) {
    std::mt19937 mt(seed);
    for (size_t i = 0; i < n; i++) out[i] = (uint32_t)mt();
}

/* piv_samp_serial, FRIES/compress_utils.cpp:390-530 */
size_t ref_piv_samp_serial(double *values, size_t count, double seg_norm, uint32_t n_samp, uint8_t *keep,
                           uint32_t seed, uint64_t skip) {
    std::vector<bool> k(count);
    for (size_t i = 0; i < count; i++) k[i] = keep[i] != 0;
    std::mt19937 mt(seed);
    mt.discard(skip);
    std::mt19937 before = mt;
    piv_samp_serial(values, count, seg_norm, n_samp, k, mt);
    for (size_t i = 0; i < count; i++) keep[i] = k[i];
    return draws_between(before, mt, 2 * (size_t)n_samp + 8);
}

/* piv_budget, FRIES/compress_utils.cpp:560-608, as rank 0 of n_procs ranks (see oracle/mpi_shim/mpi.h): all budgets */
size_t ref_piv_budget(const double *loc_norms, int n_procs, uint32_t n_samp, uint32_t seed, uint32_t *budgets) {
    std::vector<double> ln(loc_norms, loc_norms + n_procs);
    std::mt19937 mt(seed);
    std::mt19937 before = mt;
    fries_shim_world_size = n_procs;
    uint32_t mine = piv_budget(ln.data(), n_samp, mt);
    if (n_procs > 1) std::memcpy(budgets, fries_shim_scatter_log, sizeof(uint32_t) * n_procs);
    else budgets[0] = mine;
    fries_shim_world_size = 1;
    return draws_between(before, mt, 2 * (size_t)n_procs + 8);
}

/* adjust_probs, FRIES/compress_utils.cpp:617-681 */
double ref_adjust_probs(double *values, size_t count, uint32_t *n_samp_loc, double exp_nsamp_loc, uint32_t n_samp_tot,
                        double tot_norm, uint8_t *keep) {
    std::vector<bool> k(count);
    for (size_t i = 0; i < count; i++) k[i] = keep[i] != 0;
    double r = adjust_probs(values, count, n_samp_loc, exp_nsamp_loc, n_samp_tot, tot_norm, k);
    for (size_t i = 0; i < count; i++) keep[i] = k[i];
    return r;
}

/* piv_comp_parallel, FRIES/compress_utils.cpp:354-387, single rank.  keep out: 1 = zeroed element */
size_t ref_piv_comp_parallel(double *values, size_t count, uint32_t compress_size, uint8_t *keep, uint32_t seed) {
    std::vector<size_t> srt(count);
    std::vector<bool> k(count, false);
    std::mt19937 mt(seed);
    std::mt19937 before = mt;
    piv_comp_parallel(values, count, compress_size, srt, k, mt);
    for (size_t i = 0; i < count; i++) keep[i] = k[i];
    return draws_between(before, mt, 2 * (size_t)compress_size + 8);
}

/* seed_sys, FRIES/compress_utils.cpp:107-127, single rank */
double ref_seed_sys(double *norms, double *rn, unsigned n_samp) { return seed_sys(norms, rn, n_samp); }

/* comp_sub, FRIES/compress_utils.cpp:797-820 (= find_keep_sub :130-276 + sys_sub :702-794).
 * sub_weights is row-major count x n_sub; sub_sizes may be NULL.  Outputs: new_vals[<= cap],
 * new_idx[<= cap][2] (u64).  Returns number of outputs; *n_samp_left and *loc_norm report the
 * state after find_keep_sub. */
size_t ref_comp_sub(const double *values, size_t count, const uint32_t *n_div, const double *sub_weights,
                    size_t n_sub, const uint16_t *sub_sizes, unsigned n_samp, double rn, double *new_vals,
                    uint64_t *new_idx, size_t cap) {
    std::vector<double> v(values, values + count);
    std::vector<uint32_t> nd(n_div, n_div + count);
    Matrix<double> sw(count ? count : 1, n_sub);
    if (count) std::memcpy(sw.data(), sub_weights, sizeof(double) * count * n_sub);
    Matrix<bool> keep(count ? count : 1, n_sub);
    std::vector<uint16_t> ss;
    if (sub_sizes) ss.assign(sub_sizes, sub_sizes + count);
    std::vector<double> wt_remain(count + 1);
    std::vector<double> nv(cap);
    size_t (*ni)[2] = (size_t (*)[2])malloc(sizeof(size_t) * 2 * cap);
    size_t n = comp_sub(v.data(), count, nd.data(), sw, keep, sub_sizes ? ss.data() : NULL, n_samp,
                        wt_remain.data(), rn, nv.data(), ni);
    for (size_t i = 0; i < n && i < cap; i++) {
        new_vals[i] = nv[i];
        new_idx[2 * i] = ni[i][0];
        new_idx[2 * i + 1] = ni[i][1];
    }
    free(ni);
    return n;
}

/* find_keep_sub alone, FRIES/compress_utils.cpp:130-276: keep_out is count x n_sub bytes */
double ref_find_keep_sub(const double *values, size_t count, const uint32_t *n_div, const double *sub_weights,
                         size_t n_sub, const uint16_t *sub_sizes, unsigned *n_samp, double *wt_remain,
                         uint8_t *keep_out) {
    std::vector<double> v(values, values + count);
    std::vector<uint32_t> nd(n_div, n_div + count);
    Matrix<double> sw(count ? count : 1, n_sub);
    if (count) std::memcpy(sw.data(), sub_weights, sizeof(double) * count * n_sub);
    Matrix<bool> keep(count ? count : 1, n_sub);
    std::vector<uint16_t> ss;
    if (sub_sizes) ss.assign(sub_sizes, sub_sizes + count);
    double r = find_keep_sub(v.data(), nd.data(), sw, keep, sub_sizes ? ss.data() : NULL, count, n_samp, wt_remain);
    for (size_t i = 0; i < count; i++)
        for (size_t j = 0; j < n_sub; j++) keep_out[i * n_sub + j] = keep(i, j);
    return r;
}

/* adjust_shift, FRIES/compress_utils.cpp:684-693 */
void ref_adjust_shift(double *shift, double one_norm, double *last_norm, double target_norm, double damp) {
    adjust_shift(shift, one_norm, last_norm, target_norm, damp);
}

/* ---- molecular Hamiltonian handle ------------------------------------------------------------- */

struct RefMol {
    unsigned n_orb, n_elec, n_frz, tot_orb; /* n_orb = unfrozen spatial, n_elec = TOTAL electrons */
    SymmERIs eris;
    Matrix<double> hcore;
    std::vector<uint8_t> symm;
    SymmInfo *symm_info;
    hb_info *hb;
    RefMol(unsigned tot) : eris(tot), hcore(tot, tot) {}
};

/* eris_chem: dense tot_orb^4 chemist-notation (ij|kl), must be 8-fold symmetric; packed with
 * SymmERIs::chemist_ordered ndarr.hpp:232-236.  symm: n_orb irreps of the UNFROZEN orbitals. */
void *ref_mol_create(unsigned n_orb, unsigned n_elec_total, unsigned n_frz, const double *hcore,
                     const double *eris_chem, const uint8_t *symm) {
    unsigned tot = n_orb + n_frz / 2;
    RefMol *m = new RefMol(tot);
    m->n_orb = n_orb; m->n_elec = n_elec_total; m->n_frz = n_frz; m->tot_orb = tot;
    for (unsigned i = 0; i < tot; i++)
        for (unsigned j = 0; j < tot; j++) m->hcore(i, j) = hcore[i * tot + j];
    for (unsigned i = 0; i < tot; i++)
        for (unsigned j = 0; j <= i; j++)
            for (unsigned k = 0; k < tot; k++)
                for (unsigned l = 0; l <= k; l++) {
                    size_t p1 = I_J_TO_TRI_WDIAG(j, i), p2 = I_J_TO_TRI_WDIAG(l, k);
                    if (p1 <= p2) m->eris.chemist_ordered(j, i, l, k) = eris_chem[((i * (size_t)tot + j) * tot + k) * tot + l];
                }
    m->symm.assign(symm, symm + n_orb);
    m->symm_info = new SymmInfo(m->symm.data(), n_orb);
    m->hb = set_up(tot, n_orb, m->eris); /* heat_bathPP.cpp:99-179 */
    return m;
}
void ref_mol_destroy(void *h) { delete (RefMol *)h; }

/* hb_info tables, heat_bathPP.hpp:25-34; each out pointer may be NULL */
void ref_mol_hb_tables(void *h, double *d_diff, double *d_same, double *s_tens, double *s_norm,
                       double *exch_sqrt, double *diag_sqrt, double *exch_norms) {
    RefMol *m = (RefMol *)h;
    unsigned M = m->n_orb, T = M * (M - 1) / 2;
    if (d_diff) std::memcpy(d_diff, m->hb->d_diff, sizeof(double) * M * M);
    if (d_same) std::memcpy(d_same, m->hb->d_same, sizeof(double) * T);
    if (s_tens) std::memcpy(s_tens, m->hb->s_tens, sizeof(double) * M);
    if (s_norm) *s_norm = m->hb->s_norm;
    if (exch_sqrt) std::memcpy(exch_sqrt, m->hb->exch_sqrt, sizeof(double) * T);
    if (diag_sqrt) std::memcpy(diag_sqrt, m->hb->diag_sqrt, sizeof(double) * M);
    if (exch_norms) std::memcpy(exch_norms, m->hb->exch_norms, sizeof(double) * M);
}

static void occ_of(RefMol *m, uint64_t key, uint8_t *occ) {
    uint8_t bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    find_bits(bytes, occ, (uint8_t)CEILING(2 * m->n_orb, 8));
}

/* diag_matrel (SymmERIs), molecule.cpp:983-1029 */
void ref_mol_diag(void *h, const uint64_t *keys, size_t n, double *out) {
    RefMol *m = (RefMol *)h;
    for (size_t i = 0; i < n; i++) {
        uint8_t occ[64];
        occ_of(m, keys[i], occ);
        out[i] = diag_matrel(occ, m->tot_orb, m->eris, m->hcore, m->n_frz, m->n_elec);
    }
}

/* sing_matr_el_nosgn molecule.cpp:76-105; orbs n x 2 */
void ref_mol_sing_el(void *h, const uint64_t *keys, const uint8_t *orbs, size_t n, double *out) {
    RefMol *m = (RefMol *)h;
    for (size_t i = 0; i < n; i++) {
        uint8_t occ[64], o[2] = {orbs[2 * i], orbs[2 * i + 1]};
        occ_of(m, keys[i], occ);
        out[i] = sing_matr_el_nosgn(o, occ, m->tot_orb, m->eris, m->hcore, m->n_frz, m->n_elec - m->n_frz);
    }
}
/* doub_matr_el_nosgn molecule.cpp:26-42; orbs n x 4 */
void ref_mol_doub_el(void *h, const uint8_t *orbs, size_t n, double *out) {
    RefMol *m = (RefMol *)h;
    for (size_t i = 0; i < n; i++) {
        uint8_t o[4] = {orbs[4 * i], orbs[4 * i + 1], orbs[4 * i + 2], orbs[4 * i + 3]};
        out[i] = doub_matr_el_nosgn(o, m->tot_orb, m->eris, m->n_frz);
    }
}

/* sing_ex_symm molecule.cpp:178-203 / doub_ex_symm :108-175: list for ONE determinant */
size_t ref_mol_sing_ex(void *h, uint64_t key, uint8_t *out, size_t cap) {
    RefMol *m = (RefMol *)h;
    unsigned ne = m->n_elec - m->n_frz;
    std::vector<uint8_t> buf(2 * (size_t)ne * m->n_orb + 16);
    uint8_t occ[64], bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    occ_of(m, key, occ);
    size_t n = sing_ex_symm(bytes, occ, ne, m->n_orb, (uint8_t(*)[2])buf.data(), m->symm.data());
    std::memcpy(out, buf.data(), 2 * (n < cap ? n : cap));
    return n;
}
size_t ref_mol_doub_ex(void *h, uint64_t key, uint8_t *out, size_t cap) {
    RefMol *m = (RefMol *)h;
    unsigned ne = m->n_elec - m->n_frz;
    std::vector<uint8_t> buf(4 * (size_t)ne * ne * m->n_orb * m->n_orb + 16);
    uint8_t occ[64], bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    occ_of(m, key, occ);
    size_t n = doub_ex_symm(bytes, occ, ne, m->n_orb, (uint8_t(*)[4])buf.data(), m->symm.data());
    std::memcpy(out, buf.data(), 4 * (n < cap ? n : cap));
    return n;
}
/* count_singex molecule.cpp:914-933 */
size_t ref_mol_count_singex(void *h, uint64_t key) {
    RefMol *m = (RefMol *)h;
    uint8_t occ[64], bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    occ_of(m, key, occ);
    return count_singex(bytes, occ, m->n_elec - m->n_frz, m->symm_info);
}

/* HB-PP weight rows, heat_bathPP.cpp:182-412.  which: 0 o1, 1 o2, 2 o2_half, 3 u1, 4 u2, 5 u2_half.
 * args: a0..a3 meaning per function (see oracle tests).  Returns the norm; *len = row length. */
double ref_mol_hb_row(void *h, int which, uint64_t key, int a0, int a1, int a2, int a3, double *row, int *len) {
    RefMol *m = (RefMol *)h;
    unsigned ne = m->n_elec - m->n_frz;
    uint8_t occ[65], bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    occ_of(m, key, occ);
    occ[ne] = 255;
    uint16_t plen = 0;
    double r = 0;
    switch (which) {
        case 0: r = calc_o1_probs(m->hb, row, ne, occ, a0); *len = ne - (a0 > 0); break;
        case 1: r = calc_o2_probs(m->hb, row, ne, occ, (uint8_t)a0); *len = ne; break;
        case 2: r = calc_o2_probs_half(m->hb, row, ne, occ, (uint8_t)a0); *len = a0; break;
        case 3: r = calc_u1_probs(m->hb, row, (uint8_t)a0, occ, (uint8_t)ne, a1); *len = m->n_orb - ne / 2; break;
        case 4: r = calc_u2_probs(m->hb, row, (uint8_t)a0, (uint8_t)a1, (uint8_t)a2, m->symm_info, &plen); *len = plen; break;
        case 5: r = calc_u2_probs_half(m->hb, row, (uint8_t)a0, (uint8_t)a1, (uint8_t)a2, bytes, m->symm_info, &plen); *len = plen; break;
    }
    (void)a3;
    return r;
}
/* calc_unnorm_wt :414-439 / calc_norm_wt :442-598 */
double ref_mol_hb_wt(void *h, int normalized, uint64_t key, const uint8_t *orbs) {
    RefMol *m = (RefMol *)h;
    unsigned ne = m->n_elec - m->n_frz;
    uint8_t occ[65], bytes[16] = {0}, o[4] = {orbs[0], orbs[1], orbs[2], orbs[3]};
    key_to_bytes(key, bytes, 8);
    occ_of(m, key, occ);
    occ[ne] = 255;
    if (normalized) return calc_norm_wt(m->hb, o, occ, ne, bytes, m->symm_info);
    return calc_unnorm_wt(m->hb, o);
}

/* apply_HBPP_sys, heat_bathPP.cpp:686-992, on a list of determinants with weights.
 * uniforms: the 5 numbers the function would draw from mt19937 (injected by constructing a
 * generator that replays them is impossible, so the shim seeds std::mt19937 with `seed` and ALSO
 * returns the 5 uniforms it produced so that the device path can be driven by the same numbers).
 * Outputs (capacity cap): out_val[], out_det[] (index into keys), out_orbs[][4].
 * Returns number of successes. */
size_t ref_mol_apply_hbpp_sys(void *h, const uint64_t *keys, const double *vals, size_t n, double p_doub,
                              int new_hb, unsigned seed, unsigned n_samp, size_t spawn_length, double *uniforms_out,
                              double *out_val, uint64_t *out_det, uint8_t *out_orbs, size_t cap) {
    RefMol *m = (RefMol *)h;
    unsigned ne = m->n_elec - m->n_frz;
    unsigned n_bytes = CEILING(2 * m->n_orb, 8);
    /* one spare (zeroed) row: calc_u1_probs and find_nth_virt read occ_orbs[n_elec], i.e. the first entry of the
     * next row; in the drivers the matrix is DistVec::occ_orbs_ with max_size rows, so that entry always exists */
    Matrix<uint8_t> all_orbs(n + 1, ne);
    Matrix<uint8_t> all_dets(n + 1, n_bytes);
    for (size_t i = 0; i < n; i++) {
        key_to_bytes(keys[i], all_dets[i], n_bytes);
        find_bits(all_dets[i], all_orbs[i], (uint8_t)n_bytes);
    }
    size_t n_states = ne > (m->n_orb - ne / 2) ? ne : m->n_orb - ne / 2;
    if (n_states < m->symm_info->max_n_symm) n_states = m->symm_info->max_n_symm;
    HBCompressSys comp(spawn_length, n_states);
    for (size_t i = 0; i < n; i++) {
        comp.vec1[i] = vals[i];
        comp.det_indices1[i] = i;
    }
    comp.vec_len = n;
    std::mt19937 mt(seed);
    if (uniforms_out) {
        std::mt19937 mt2(seed);
        for (int i = 0; i < 5; i++) uniforms_out[i] = mt2() / (1. + UINT32_MAX);
    }
    unsigned tot_orb = m->tot_orb, n_frz = m->n_frz;
    SymmERIs *eris = &m->eris;
    Matrix<double> *hc = &m->hcore;
    std::function<double(uint8_t *, uint8_t *)> sing_fn = [=](uint8_t *ex, uint8_t *occ) {
        return sing_matr_el_nosgn(ex, occ, tot_orb, *eris, *hc, n_frz, ne);
    };
    std::function<double(uint8_t *)> doub_fn = [=](uint8_t *ex) { return doub_matr_el_nosgn(ex, tot_orb, *eris, n_frz); };
    apply_HBPP_sys(all_orbs, all_dets, &comp, m->hb, m->symm_info, p_doub, new_hb != 0, mt, n_samp, sing_fn, doub_fn);
    size_t ns = comp.vec_len;
    for (size_t i = 0; i < ns && i < cap; i++) {
        out_val[i] = comp.vec1[i];
        out_det[i] = comp.det_indices2[i];
        std::memcpy(out_orbs + 4 * i, comp.orb_indices1[i], 4);
    }
    return ns;
}

/* apply_HBPP_piv, FRIES/Hamiltonians/heat_bathPP.cpp:1014-1419, spin_parity = 0, with the real matrix-element
 * functions.  mt19937(seed); *n_draws receives the number of generator outputs consumed. */
size_t ref_mol_apply_hbpp_piv(void *h, const uint64_t *keys, const double *vals, size_t n, double p_doub, int new_hb,
                              unsigned seed, unsigned n_samp, size_t spawn_length, double *out_val, uint64_t *out_det,
                              uint8_t *out_orbs, size_t cap, size_t *n_draws) {
    RefMol *m = (RefMol *)h;
    unsigned ne = m->n_elec - m->n_frz;
    unsigned n_bytes = CEILING(2 * m->n_orb, 8);
    Matrix<uint8_t> all_orbs(n + 1, ne);
    Matrix<uint8_t> all_dets(n + 1, n_bytes);
    for (size_t i = 0; i < n; i++) {
        key_to_bytes(keys[i], all_dets[i], n_bytes);
        find_bits(all_dets[i], all_orbs[i], (uint8_t)n_bytes);
    }
    size_t n_states = ne > (m->n_orb - ne / 2) ? ne : m->n_orb - ne / 2;
    if (n_states < m->symm_info->max_n_symm) n_states = m->symm_info->max_n_symm;
    if (n_states < 2) n_states = 2;
    HBCompressPiv comp(spawn_length, n_states);
    for (size_t i = 0; i < n; i++) {
        comp.vec1[i] = vals[i];
        comp.det_indices1[i] = i;
    }
    comp.vec_len = n;
    std::mt19937 mt(seed);
    std::mt19937 before = mt;
    unsigned tot_orb = m->tot_orb, n_frz = m->n_frz;
    SymmERIs *eris = &m->eris;
    Matrix<double> *hc = &m->hcore;
    std::function<double(uint8_t *, uint8_t *)> sing_fn = [=](uint8_t *ex, uint8_t *occ) {
        return sing_matr_el_nosgn(ex, occ, tot_orb, *eris, *hc, n_frz, ne);
    };
    std::function<double(uint8_t *)> doub_fn = [=](uint8_t *ex) { return doub_matr_el_nosgn(ex, tot_orb, *eris, n_frz); };
    apply_HBPP_piv(all_orbs, all_dets, &comp, m->hb, m->symm_info, p_doub, new_hb != 0, mt, n_samp, sing_fn, doub_fn, 0);
    if (n_draws) *n_draws = draws_between(before, mt, 12 * (size_t)n_samp + 64);
    size_t ns = comp.vec_len;
    for (size_t i = 0; i < ns && i < cap; i++) {
        out_val[i] = comp.vec1[i];
        out_det[i] = comp.det_indices2[i];
        std::memcpy(out_orbs + 4 * i, comp.orb_indices1[i], 4);
    }
    return ns;
}

/* ---- a2/a3: DistVec add / perform_add / add_elements ------------------------------------------ */

struct RefVec {
    DistVec<double> *vec;
    unsigned n_bits, n_bytes;
};

/* DistVec ctor vec_utils.hpp:181-187 (n_vecs rows, Adder of add_size) */
void *ref_vec_create(size_t size, size_t add_size, unsigned n_bits, unsigned n_elec, unsigned n_vecs,
                     const uint32_t *proc_scr, const uint32_t *vec_scr) {
    std::vector<uint32_t> ps(proc_scr, proc_scr + n_bits), vs(vec_scr, vec_scr + n_bits);
    RefVec *r = new RefVec;
    r->n_bits = n_bits;
    r->n_bytes = CEILING(n_bits, 8);
    r->vec = new DistVec<double>(size, add_size, (uint8_t)n_bits, n_elec, 1, nullptr, (uint8_t)n_vecs, ps, vs);
    return r;
}
void ref_vec_destroy(void *h) {
    RefVec *r = (RefVec *)h;
    delete r->vec;
    delete r;
}
/* DistVec::add x n then perform_add(origin) with curr_vec_idx = dest: vec_utils.hpp:418-440,606-641 */
void ref_vec_add(void *h, const uint64_t *keys, const double *vals, const uint8_t *ini, size_t n, unsigned origin,
                 unsigned dest) {
    RefVec *r = (RefVec *)h;
    r->vec->set_curr_vec_idx((uint8_t)dest);
    size_t i = 0;
    while (i < n) {
        bool room = true;
        while (i < n && room) {
            uint8_t bytes[16] = {0};
            key_to_bytes(keys[i], bytes, 8);
            room = r->vec->add(bytes, vals[i], ini[i]);
            i++;
        }
        r->vec->perform_add(origin);
    }
    r->vec->perform_add(origin);
}
size_t ref_vec_curr_size(void *h) { return ((RefVec *)h)->vec->curr_size(); }
int ref_vec_n_nonz(void *h) { return ((RefVec *)h)->vec->n_nonz(); }
uint64_t ref_vec_nonini_occ_add(void *h) { return ((RefVec *)h)->vec->tot_sgn_coh(); }
/* dump storage: keys[curr_size], vals[n_vecs][curr_size] */
void ref_vec_dump(void *h, uint64_t *keys, double *vals, unsigned n_vecs) {
    RefVec *r = (RefVec *)h;
    size_t n = r->vec->curr_size();
    for (size_t i = 0; i < n; i++) keys[i] = bytes_to_key(r->vec->indices()[i], r->n_bytes);
    for (unsigned v = 0; v < n_vecs; v++)
        for (size_t i = 0; i < n; i++) vals[v * n + i] = *(*r->vec)(v, i);
}
/* del_at_pos vec_utils.hpp:458-476 for every position whose flag is set */
void ref_vec_del(void *h, const uint8_t *flags) {
    RefVec *r = (RefVec *)h;
    size_t n = r->vec->curr_size();
    for (size_t i = 0; i < n; i++)
        if (flags[i]) r->vec->del_at_pos(i);
}

/* compress_vecs_multi vec_utils.cpp:73-127 on rows [start, end) with std::mt19937(seed); setup_alias / sample_alias
 * compress_utils.cpp:823-897 on plain arrays */
void ref_vec_compress_multi(void *h, unsigned start, unsigned end, unsigned compress_size, uint32_t seed) {
    RefVec *r = (RefVec *)h;
    size_t n = r->vec->curr_size();
    std::vector<size_t> srt(n + 1);
    std::vector<bool> keep(n + 1, false), del(n + 1, true);
    std::mt19937 mt(seed);
    compress_vecs_multi(*r->vec, start, end, compress_size, srt, keep, del, mt);
}
void ref_setup_alias(const double *probs, uint32_t *aliases, double *alias_probs, size_t n) {
    std::vector<double> p(probs, probs + n);
    setup_alias(p.data(), aliases, alias_probs, n);
}
void ref_sample_alias(uint32_t *aliases, double *alias_probs, size_t n, uint16_t *counts, uint32_t n_samp, uint32_t seed) {
    std::mt19937 mt(seed);
    sample_alias(aliases, alias_probs, n, counts, n_samp, mt);
}

/* ---- a16: h_op_offdiag + h_op_diag (SymmERIs variant molecule.cpp:448-665, :205-219) ------------
 * Computes dest = id_fac*v + h_fac*H*v for the list (keys, vals) and returns the result as a list.
 * The diagonal uses diag_matrel - 0 (no HF shift). */
size_t ref_mol_h_apply(void *h, const uint64_t *keys, const double *vals, size_t n, double id_fac, double h_fac,
                       size_t max_dets, const uint32_t *proc_scr, const uint32_t *vec_scr, uint64_t *out_keys,
                       double *out_vals, size_t cap) {
    RefMol *m = (RefMol *)h;
    unsigned ne = m->n_elec - m->n_frz;
    unsigned n_bits = 2 * m->n_orb;
    std::vector<uint32_t> ps(proc_scr, proc_scr + n_bits), vs(vec_scr, vec_scr + n_bits);
    unsigned tot_orb = m->tot_orb, n_frz = m->n_frz, n_elec = m->n_elec;
    SymmERIs *eris = &m->eris;
    Matrix<double> *hc = &m->hcore;
    std::function<double(const uint8_t *)> diag_fn = [=](const uint8_t *occ) {
        return diag_matrel(occ, tot_orb, *eris, *hc, n_frz, n_elec);
    };
    size_t n_ex = (size_t)m->n_orb * m->n_orb * ne * ne;
    size_t adder = n_ex * 8 < 100000 ? 100000 : n_ex * 8;
    DistVec<double> vec(max_dets, adder, (uint8_t)n_bits, ne, 1, diag_fn, 2, ps, vs);
    for (size_t i = 0; i < n; i++) {
        uint8_t bytes[16] = {0};
        key_to_bytes(keys[i], bytes, 8);
        if (!vec.add(bytes, vals[i], 1)) vec.perform_add(0);
    }
    vec.perform_add(0);
    std::vector<uint8_t> scratch(4 * n_ex + 64);
    h_op_diag(vec, 1, id_fac, h_fac);
    vec.set_curr_vec_idx(0);
    h_op_offdiag(vec, m->symm.data(), tot_orb, *eris, *hc, scratch.data(), scratch.size(), n_frz, ne, 1, h_fac, 0);
    size_t cs = vec.curr_size();
    unsigned nb = CEILING(n_bits, 8);
    for (size_t i = 0; i < cs && i < cap; i++) {
        out_keys[i] = bytes_to_key(vec.indices()[i], nb);
        out_vals[i] = *vec(1, i);
    }
    return cs;
}

/* ---- a19: Hubbard helpers ---------------------------------------------------------------------- */
/* hub_diag hub_holstein.cpp:101-136 */
unsigned ref_hub_diag(uint64_t key, unsigned n_sites) {
    uint8_t bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    return hub_diag(bytes, n_sites);
}
/* gen_neel_det_1D hub_holstein.cpp:139-171 */
uint64_t ref_gen_neel_det_1D(unsigned n_sites, unsigned n_elec, unsigned ph_bits) {
    uint8_t bytes[32] = {0};
    gen_neel_det_1D(n_sites, n_elec, ph_bits, bytes);
    return bytes_to_key(bytes, 8);
}


/* HubHolVec helpers, FRIES/hh_vec.hpp: find_neighbors_1D :139-175 (out: 2 x (n_elec + 1) bytes), idx_to_hash :72-88 */
static HubHolVec<double> *make_hh(unsigned n_sites, unsigned ph_bits, unsigned n_elec, const uint32_t *scr) {
    std::vector<uint32_t> s(scr, scr + 2 * n_sites);
    std::function<double(const uint8_t *)> diag = [n_sites](const uint8_t *det) { return (double)hub_diag((uint8_t *)det, n_sites); };
    return new HubHolVec<double>(16, 16, (uint8_t)n_sites, (uint8_t)ph_bits, n_elec, 1, diag, 1, s, s);
}
void ref_hh_neighbors(uint64_t key, unsigned n_sites, unsigned ph_bits, unsigned n_elec, uint8_t *out) {
    uint32_t scr[64] = {0};
    HubHolVec<double> *v = make_hh(n_sites, ph_bits, n_elec, scr);
    uint8_t bytes[16] = {0};
    key_to_bytes(key, bytes, 8);
    v->find_neighbors_1D(bytes, out);
    delete v;
}
uint64_t ref_hh_hash(uint64_t key, unsigned n_sites, unsigned ph_bits, unsigned n_elec, const uint32_t *scr) {
    HubHolVec<double> *v = make_hh(n_sites, ph_bits, n_elec, scr);
    uint8_t bytes[16] = {0}, orbs[64];
    key_to_bytes(key, bytes, 8);
    uint64_t h = (uint64_t)v->idx_to_hash(bytes, orbs);
    delete v;
    return h;
}
/* calc_ref_ovlp hub_holstein.hpp:93-182 over a list */
double ref_hh_ref_ovlp(const uint64_t *keys, const double *vals, size_t n, uint64_t ref, unsigned n_elec, unsigned n_sites,
                       unsigned ph_bits, double g_over_t) {
    uint32_t scr[64] = {0};
    HubHolVec<double> *v = make_hh(n_sites, ph_bits, n_elec, scr);
    size_t nb = CEILING(n_sites * (2 + ph_bits), 8);
    Matrix<uint8_t> dets(n + 1, nb), ph(n + 1, n_sites);
    std::vector<double> vv(vals, vals + n);
    for (size_t i = 0; i < n; i++) {
        key_to_bytes(keys[i], dets[i], nb);
        v->decode_phonons(dets[i], ph[i]);
    }
    uint8_t ref_b[16] = {0}, occ[64];
    key_to_bytes(ref, ref_b, 8);
    v->gen_orb_list(ref_b, occ);
    double r = calc_ref_ovlp(dets, vv.data(), ph, n, ref_b, occ, (uint8_t)n_elec, n_sites, g_over_t);
    delete v;
    return r;
}

}  // extern "C"
