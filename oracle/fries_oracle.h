/* fries_oracle -- plain-C restatement of the reference's algorithm for the FRI hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and only
 * as the checker.  Each function cites the reference file:line (relative to sgreene8/FRIES) that it
 * restates.  The restatement is pinned against the reference's own known-answer tests
 * (tests/test_bitstrings.cpp, test_vector.cpp, test_hamiltonian.cpp, test_compression.cpp) and
 * against the compiled reference itself (oracle/_ref/libfries_ref.so) in tests/test_oracle_*.py, and
 * against the committed fixtures in tests/golden/ (generated from the compiled reference by
 * tests/golden/make_golden.py).
 *
 * Determinants are uint64_t keys: bit i = spin-orbital i (little-endian load of the reference's
 * byte string).
 */
#ifndef FRIES_ORACLE_H
#define FRIES_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- L1 bit utilities ---- */
int fo_find_bits(uint64_t key, uint8_t *occ);                      /* math_utils.c:62-98 */
unsigned fo_bits_between(uint64_t key, int a, int b);              /* math_utils.c:9-58 */
uint64_t fo_gen_hf_bitstring(unsigned n_orb, unsigned n_elec);     /* fci_utils.c:10-43 */
int fo_excite_sign(int cre, int des, uint64_t key);                /* fci_utils.c:128-135 */
int fo_sing_det_parity(uint64_t *key, const uint8_t *orbs);        /* fci_utils.c:46-51 */
int fo_doub_det_parity(uint64_t *key, const uint8_t *orbs);        /* fci_utils.c:67-75 */
int fo_sing_parity(uint64_t key, const uint8_t *orbs);             /* fci_utils.c:54-57 */
int fo_doub_parity(uint64_t key, const uint8_t *orbs);             /* fci_utils.c:86-94 */
int fo_find_nth_virt(const uint8_t *occ, int spin, int n_elec, int n_orb, int n); /* fci_utils.c:138-148 */

/* ---- a1 ---- */
uint64_t fo_hash(uint64_t key, const uint32_t *scrambler);         /* det_hash.hpp:160-170 */
void fo_hash_keys(const uint64_t *keys, size_t n, const uint32_t *scrambler, int n_procs, uint64_t *hash_out,
                  int32_t *owner_out);                             /* vec_utils.hpp:360-400 */

/* ---- a4/a5 ---- */
double fo_find_preserve(const double *values, size_t count, unsigned *n_samp, double *glob_norm,
                        uint8_t *keep);                            /* compress_utils.cpp:29-105 */
double fo_seed_sys(const double *norms, int n_procs, int rank, double *rn, unsigned n_samp); /* :107-127 */
void fo_sys_comp(double *values, size_t count, double *loc_norms, int n_procs, int rank, unsigned n_samp,
                 uint8_t *keep, double rn);                        /* compress_utils.cpp:278-327 */

/* ---- a6 ---- */
void fo_set_keep_chunk(size_t chunk); /* 8 = reference behaviour (default); 1 = self-consistent fixed point */
double fo_find_keep_sub(const double *values, const uint32_t *n_div, const double *sub_weights, size_t n_sub,
                        uint8_t *keep, const uint16_t *sub_sizes, size_t count, unsigned *n_samp,
                        double *wt_remain);                        /* compress_utils.cpp:130-276 */
size_t fo_sys_sub(const double *values, const uint32_t *n_div, const double *sub_weights, size_t n_sub,
                  uint8_t *keep, const uint16_t *sub_sizes, size_t count, unsigned n_samp,
                  const double *wt_remain, double loc_norm, double rn, double *new_vals,
                  uint64_t *new_idx);                              /* compress_utils.cpp:702-794 */
size_t fo_comp_sub(const double *values, size_t count, const uint32_t *n_div, const double *sub_weights,
                   size_t n_sub, const uint16_t *sub_sizes, unsigned n_samp, double rn, double *new_vals,
                   uint64_t *new_idx, unsigned *n_samp_left, double *loc_norm); /* compress_utils.cpp:797-820 */
/* ---- pivotal family: piv_samp_serial / piv_budget / adjust_probs / piv_comp_parallel ---- */
void fo_setup_alias(const double *probs, uint32_t *aliases, double *alias_probs, size_t n);   /* compress_utils.cpp:823-857 */
void fo_sample_alias(const uint32_t *aliases, const double *alias_probs, size_t n, uint16_t *counts, uint32_t n_samp,
                     const uint32_t *draws);                                                  /* :882-897 */
size_t fo_compress_multi_row(double *values, size_t n, uint32_t compress_size, const uint32_t *draws); /* vec_utils.cpp:73-127 */
void fo_mt19937_fill(uint32_t seed, size_t n, uint32_t *out);     /* std::mt19937(seed): first n outputs */
void fo_piv_samp_serial(double *v, size_t n, double seg_norm, uint32_t n_samp, uint8_t *keep, const uint32_t *draws,
                        size_t *used);                             /* compress_utils.cpp:389-520 */
void fo_piv_budget(const double *loc_norms, int n_procs, uint32_t n_samp, const uint32_t *draws, size_t *used,
                   uint32_t *budgets);                             /* compress_utils.cpp:552-608 */
double fo_adjust_probs(double *v, size_t n, uint32_t *n_loc, double exp_loc, uint32_t n_tot, double tot_norm,
                       uint8_t *keep);                             /* compress_utils.cpp:610-681 */
void fo_piv_comp(double *v, size_t n, uint32_t compress_size, uint8_t *keep, const uint32_t *draws,
                 size_t *used);                                    /* compress_utils.cpp:354-386, single rank */
void fo_adjust_shift(double *shift, double one_norm, double *last_norm, double target_norm,
                     double damp);                                 /* compress_utils.cpp:684-693 */

/* ---- molecular Hamiltonian ---- */
typedef struct fo_mol fo_mol;
/* eris_chem: dense tot_orb^4 chemist (ij|kl), 8-fold symmetric; n_elec = TOTAL electrons */
fo_mol *fo_mol_create(unsigned n_orb, unsigned n_elec, unsigned n_frz, const double *hcore, const double *eris_chem,
                      const uint8_t *symm);
void fo_mol_destroy(fo_mol *m);
size_t fo_mol_packed_len(const fo_mol *m);
const double *fo_mol_packed_eris(const fo_mol *m);                 /* SymmERIs layout ndarr.hpp:206-244 */
void fo_mol_hb_tables(const fo_mol *m, double *d_diff, double *d_same, double *s_tens, double *s_norm,
                      double *exch_sqrt, double *diag_sqrt, double *exch_norms); /* heat_bathPP.cpp:99-179 */
double fo_mol_diag(const fo_mol *m, uint64_t key);                 /* molecule.cpp:983-1029 */
double fo_mol_sing_el(const fo_mol *m, uint64_t key, const uint8_t *orbs); /* molecule.cpp:76-105 */
double fo_mol_doub_el(const fo_mol *m, const uint8_t *orbs);       /* molecule.cpp:26-42 */
size_t fo_mol_sing_ex(const fo_mol *m, uint64_t key, uint8_t *out); /* molecule.cpp:178-203 */
size_t fo_mol_doub_ex(const fo_mol *m, uint64_t key, uint8_t *out); /* molecule.cpp:108-175 */
size_t fo_mol_count_singex(const fo_mol *m, uint64_t key);         /* molecule.cpp:914-933 */
double fo_mol_hb_row(const fo_mol *m, int which, uint64_t key, int a0, int a1, int a2, double *row,
                     int *len);                                    /* heat_bathPP.cpp:182-412 */
double fo_mol_hb_wt(const fo_mol *m, int normalized, uint64_t key, const uint8_t *orbs); /* :414-598 */
size_t fo_mol_apply_hbpp_sys(const fo_mol *m, const uint64_t *keys, const double *vals, size_t n, double p_doub,
                             int new_hb, const double *uniforms5, unsigned n_samp, size_t spawn_length,
                             double *out_val, uint64_t *out_det, uint8_t *out_orbs); /* heat_bathPP.cpp:686-992 */
size_t fo_mol_apply_hbpp_piv(const fo_mol *m, const uint64_t *keys, const double *vals, size_t n, double p_doub,
                             int new_hb, const uint32_t *draws, size_t *used, unsigned n_samp, size_t spawn_length,
                             double *out_val, uint64_t *out_det, uint8_t *out_orbs); /* heat_bathPP.cpp:1014-1419 */
size_t fo_debug_hbpp_stage(const fo_mol *m, const uint64_t *keys, const double *vals, size_t n, double p_doub,
                           int new_hb, const double *uniforms5, unsigned n_samp, size_t spawn_length, int stage,
                           double *out_val, uint64_t *out_det, uint8_t *out_orbs, uint32_t *out_sub);
/* full H.v of a list (h_op_diag molecule.cpp:205-219 + h_op_offdiag :448-665): out must hold
 * n * (1 + n_sing + n_doub) entries; duplicates are NOT merged (caller sorts + sums). */
size_t fo_mol_h_apply_list(const fo_mol *m, const uint64_t *keys, const double *vals, size_t n, double id_fac,
                           double h_fac, uint64_t *out_keys, double *out_vals, size_t cap);

/* ---- a19: Hubbard-Holstein ---- */
unsigned fo_hub_diag(uint64_t key, unsigned n_sites);                         /* hub_holstein.cpp:101-136 */
uint64_t fo_gen_neel_det_1D(unsigned n_sites, unsigned n_elec);               /* hub_holstein.cpp:139-171 */
void fo_hh_neighbors(uint64_t key, unsigned n_sites, unsigned n_elec, uint8_t *out); /* hh_vec.hpp:139-175 */
uint64_t fo_hash_hh(uint64_t key, const uint32_t *scr, unsigned n_sites, unsigned ph_bits); /* hh_vec.hpp:72-88 */
double fo_hh_ref_ovlp(const uint64_t *keys, const double *vals, size_t n, uint64_t ref, unsigned n_elec, unsigned n_sites,
                      unsigned ph_bits, double g_over_t);                     /* hub_holstein.hpp:93-182 */

#ifdef __cplusplus
}
#endif
#endif
