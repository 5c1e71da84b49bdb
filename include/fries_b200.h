/* fries_b200 -- C-ABI of the B200-native FRI hot path.
 *
 * This header is the drop-in boundary.  The reference (sgreene8/FRIES) has no
 * FFI: its boundary is the C++ header API of libfries plus three `extern "C"` C units
 * (FRIES/fci_utils.h, FRIES/det_store.h, FRIES/math_utils.h).  Every entry point below replaces
 * one reference function (cited file:line, relative to the reference root); INTEGRATION.md shows
 * the C++ shim a maintainer adds under the reference's own signatures.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns FRIES_OK (0) or a negative error code and
 *    never throws; fries_last_error() returns the message of the last failure on this thread.
 *  - a Slater determinant is one uint64_t "key": bit i = spin-orbital i occupied, i.e. the
 *    little-endian load of the reference's uint8_t bit string (FRIES/det_store.h:23-26).  All
 *    configurations in scope need <= 52 bits; bit 63 is reserved for the initiator flag that the
 *    reference stores at bit n_bits of its Adder buffers (FRIES/vec_utils.hpp:965-967).
 *  - pointers named h_* are HOST pointers, d_* are DEVICE pointers (cuda:device of the context).
 *    Functions without a _dev suffix take host buffers and do their own host<->device copies (they
 *    are the reference-facing calls); *_dev functions work on resident HBM buffers, asynchronously
 *    on the context's stream.
 *  - there is NO CPU fallback: every entry point fails with FRIES_ERR_CUDA when no sm_100 device is
 *    usable.
 */
#ifndef FRIES_B200_H
#define FRIES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRIES_OK 0
#define FRIES_ERR_ARG (-1)
#define FRIES_ERR_CUDA (-2)
#define FRIES_ERR_CAPACITY (-3)
#define FRIES_ERR_STATE (-4)

#define FRIES_INI_FLAG (1ull << 63)
#define FRIES_MAX_ELEC 32
#define FRIES_MAX_SUB 32

typedef struct fries_ctx fries_ctx;   /* device, stream, workspace */
typedef struct fries_vec fries_vec;   /* device-resident DistVec<double> (FRIES/vec_utils.hpp:121-953) */
typedef struct fries_mol fries_mol;   /* integrals, symmetry and HB-PP tables in HBM */
typedef struct fries_hbpp fries_hbpp; /* HBCompressSys scratch (heat_bathPP.hpp:279-297) in HBM */

const char *fries_last_error(void);
int fries_version(void);

/* ---- context ----------------------------------------------------------------------------------- */
int fries_ctx_create(int device, fries_ctx **out);
int fries_ctx_destroy(fries_ctx *ctx);
/* run on an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream */
int fries_ctx_set_stream(fries_ctx *ctx, void *cuda_stream);
int fries_ctx_sync(fries_ctx *ctx);
int fries_ctx_sm_count(fries_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t fries_ctx_launch_count(fries_ctx *ctx);
/* elapsed ms between two internal CUDA events recorded around the LAST call of the named kernel
 * family when profiling is on (fries_ctx_set_profile).  name: "comp_sub", "merge", ... */
int fries_ctx_set_profile(fries_ctx *ctx, int on);
int fries_ctx_kernel_ms(fries_ctx *ctx, const char *name, double *total_ms, uint64_t *launches);

/* ---- a1: hash and owner ---------------------------------------------------------------------------
 * HashTable::hash_fxn FRIES/det_hash.hpp:160-170; DistVec::idx_to_proc FRIES/vec_utils.hpp:360-379;
 * idx_to_hash :389-400.  scrambler has n_bits uint32 entries. d_hash / d_owner may be NULL. */
int fries_hash_owner(fries_ctx *ctx, const uint64_t *h_keys, size_t n, const uint32_t *h_scrambler, int n_bits,
                     int n_ranks, uint64_t *h_hash, int32_t *h_owner);
int fries_hash_owner_dev(fries_ctx *ctx, const uint64_t *d_keys, size_t n, const uint32_t *h_scrambler, int n_bits,
                         int n_ranks, uint64_t *d_hash, int32_t *d_owner);

/* ---- a15: bit-string utilities (batch, one thread per item) ---------------------------------------
 * op 0: sing_det_parity FRIES/fci_utils.c:46-51   (orbs n x 2; keys updated, sign out)
 * op 1: doub_det_parity FRIES/fci_utils.c:67-75   (orbs n x 4; keys updated, sign out)
 * op 2: sing_parity     FRIES/fci_utils.c:54-57   (keys unchanged)
 * op 3: doub_parity     FRIES/fci_utils.c:86-94   (keys unchanged)
 * op 4: bits_between    FRIES/math_utils.c:9-58   (orbs n x 2 = (a,b); result in sign) */
int fries_bit_op(fries_ctx *ctx, int op, uint64_t *h_keys, const uint8_t *h_orbs, size_t n, int32_t *h_sign);

/* ---- a4/a5: vector compression -----------------------------------------------------------------------
 * find_preserve FRIES/compress_utils.cpp:29-105.  n_ranks/rank describe the position of this shard;
 * with n_ranks == 1 no collective is used.  *n_samp in: budget, out: budget left.  keep: 0/1 bytes.
 * Returns local residual one-norm in *loc_norm and the global one-norm in *glob_norm. */
int fries_find_preserve(fries_ctx *ctx, const double *h_values, size_t count, unsigned *n_samp, double *glob_norm,
                        uint8_t *h_keep, double *loc_norm);
/* sys_comp FRIES/compress_utils.cpp:278-327 (+ seed_sys :107-127).  loc_norms[n_ranks] in/out,
 * keep in: preserved flags, out: 1 = zeroed element ("delete me"). */
int fries_sys_comp(fries_ctx *ctx, double *h_values, size_t count, double *loc_norms, int n_ranks, int rank,
                   unsigned n_samp, uint8_t *h_keep, double rand_num);
/* device-resident variants; d_result receives {loc_norm, glob_norm, (double)n_samp_left, n_kept} */
int fries_find_preserve_dev(fries_ctx *ctx, const double *d_values, size_t count, unsigned n_samp, uint8_t *d_keep,
                            double *d_result4);
int fries_sys_comp_dev(fries_ctx *ctx, double *d_values, size_t count, const double *d_result4, uint8_t *d_keep,
                       double rand_num, double *d_new_norm);

/* ---- pivotal compression family ------------------------------------------------------
 * Randomness: the reference hands these functions a std::mt19937& and draws uniforms as mt() / 2^32
 * (compress_utils.cpp:436,468).  Here the caller passes the generator's next raw 32-bit outputs in h_draws and
 * advances its generator by *n_draws_used afterwards (std::mt19937::discard); unit k of the sampler uses draws
 * 2k and 2k + 1, exactly as the sequential sweep does.
 *
 * piv_samp_serial FRIES/compress_utils.cpp:390-530: ordered pivotal sampling of n_samp of the elements with
 * keep == 0, each becoming +-seg_norm / n_samp; every one of them must be smaller than that unit.  keep in:
 * preserved flags, out: 1 = zeroed element.  h_draws: 2 * n_samp entries. */
int fries_piv_samp_serial(fries_ctx *ctx, double *h_values, size_t count, double seg_norm, uint32_t n_samp,
                          uint8_t *h_keep, const uint32_t *h_draws, size_t *n_draws_used);
/* resident form; d_work: 8 bytes per element + 12 per sample (+ 1 kB of alignment slack);
 * d_result4 (may be NULL) receives {one-norm after, elements drawn, units processed, anomalies} */
int fries_piv_samp_dev(fries_ctx *ctx, double *d_values, size_t count, double seg_norm, uint32_t n_samp, uint8_t *d_keep,
                       const uint32_t *d_draws, void *d_work, size_t work_bytes, double *d_result4);
/* adjust_probs FRIES/compress_utils.cpp:617-681: returns the norm for the sampler in *new_norm, updates
 * *n_samp_loc, values and keep flags */
int fries_adjust_probs(fries_ctx *ctx, double *h_values, size_t count, uint32_t *n_samp_loc, double exp_nsamp_loc,
                       uint32_t n_samp_tot, double tot_norm, uint8_t *h_keep, double *new_norm);
/* piv_budget FRIES/compress_utils.cpp:560-608: rank 0's host arithmetic on n_ranks norms; budgets of ALL ranks
 * (the reference scatters them).  h_draws: up to 2 * n_ranks entries.  No device work. */
int fries_piv_budget(const double *loc_norms, int n_ranks, uint32_t n_samp, const uint32_t *h_draws,
                     size_t *n_draws_used, uint32_t *budgets);
/* piv_comp_parallel FRIES/compress_utils.cpp:354-387 = find_preserve + piv_budget + adjust_probs +
 * piv_samp_serial.  Single rank: n_ranks = 1, preserved = 0, h_loc_norms may be NULL.  As one rank of several
 * (preserved = 1): h_keep, n_samp_left and h_loc_norms[n_ranks] come from the collective find_preserve and the
 * all-gather of its residual norms (:364-365).  h_draws: 2 * (compress_size + n_ranks) entries.  On return
 * h_loc_norms[rank] (if given) is this rank's one-norm after compression. */
int fries_piv_comp(fries_ctx *ctx, double *h_values, size_t count, uint32_t compress_size, uint8_t *h_keep,
                   const uint32_t *h_draws, size_t *n_draws_used, double *h_loc_norms, int n_ranks, int rank,
                   int preserved, uint32_t n_samp_left);

/* compress_vecs (method 0, pivotal) / compress_vecs_sys (method 1, systematic) FRIES/vec_utils.cpp:10-70 /
 * compress_vecs_multi (method 2, multinomial with the alias method, :73-127; setup_alias / sample_alias
 * FRIES/compress_utils.cpp:823-897) on the resident store: rows [start_row, end_row) are each compressed to compress_size
 * elements (method 2: compress_size samples), then the elements that are zero in every row are deleted.  Draws as above
 * (method 1: one per row; method 2: four per sample and row, <= 65535 stored elements and samples as in the reference's
 * uint16 counters).  Single rank. */
int fries_vec_compress(fries_vec *vec, unsigned start_row, unsigned end_row, uint32_t compress_size, int method,
                       const uint32_t *h_draws, size_t n_draws, size_t *n_draws_used);

/* ---- a6: hierarchical compression with explicit sub-weights -----------------------------------------
 * comp_sub FRIES/compress_utils.cpp:797-820 = find_keep_sub :130-276 + sys_sub :702-794.
 * sub_weights row-major count x n_sub (n_sub <= FRIES_MAX_SUB); sub_sizes may be NULL.
 * new_idx is [n_out][2] uint64 (weight index, sub index) as in the reference. */
int fries_comp_sub(fries_ctx *ctx, const double *h_values, size_t count, const uint32_t *h_ndiv,
                   const double *h_sub_weights, size_t n_sub, const uint16_t *h_sub_sizes, unsigned n_samp,
                   double rand_num, double *h_new_vals, uint64_t *h_new_idx, size_t out_cap, size_t *n_out,
                   unsigned *n_samp_left, double *loc_norm);

/* ---- molecular Hamiltonian tables -------------------------------------------------------------------
 * n_orb: unfrozen spatial orbitals; n_elec: TOTAL electrons; n_frz: frozen electrons;
 * h_hcore: (n_orb+n_frz/2)^2; h_eris_packed: SymmERIs layout FRIES/ndarr.hpp:206-244 over
 * tot_orb = n_orb + n_frz/2 orbitals; h_symm: irreps (0..7) of the n_orb unfrozen orbitals.
 * Builds SymmInfo (FRIES/Hamiltonians/molecule.hpp:265-280) and the HB-PP tables of set_up
 * (FRIES/Hamiltonians/heat_bathPP.cpp:99-179) on the device. */
int fries_mol_create(fries_ctx *ctx, unsigned n_orb, unsigned n_elec, unsigned n_frz, const double *h_hcore,
                     const double *h_eris_packed, const uint8_t *h_symm, fries_mol **out);
int fries_mol_destroy(fries_mol *mol);
/* hb_info tables (heat_bathPP.hpp:25-34); any pointer may be NULL */
int fries_mol_hb_tables(fries_mol *mol, double *h_d_diff, double *h_d_same, double *h_s_tens, double *h_s_norm,
                        double *h_exch_sqrt, double *h_diag_sqrt, double *h_exch_norms);
/* a12: diag_matrel FRIES/Hamiltonians/molecule.cpp:983-1029 */
int fries_mol_diag(fries_mol *mol, const uint64_t *h_keys, size_t n, double *h_out);
/* a13: sing_matr_el_nosgn molecule.cpp:76-105 (orbs n x 2), doub_matr_el_nosgn :26-42 (orbs n x 4) */
int fries_mol_sing_el(fries_mol *mol, const uint64_t *h_keys, const uint8_t *h_orbs, size_t n, double *h_out);
int fries_mol_doub_el(fries_mol *mol, const uint8_t *h_orbs, size_t n, double *h_out);
/* a14: sing_ex_symm molecule.cpp:178-203 / doub_ex_symm :108-175 for n determinants, in the
 * reference's enumeration order.  h_offsets[n+1] receives the start of each determinant's list in
 * h_orbs (n_ex x 2 or x 4).  Pass h_orbs == NULL to only count. */
int fries_mol_sing_ex(fries_mol *mol, const uint64_t *h_keys, size_t n, uint64_t *h_offsets, uint8_t *h_orbs,
                      size_t cap);
int fries_mol_doub_ex(fries_mol *mol, const uint64_t *h_keys, size_t n, uint64_t *h_offsets, uint8_t *h_orbs,
                      size_t cap);
/* a8/a9: HB-PP weight rows (heat_bathPP.cpp:182-412) and total weights (:414-598), one item per
 * call row; `which` 0 o1 (a0 = exclude_first), 1 o2 (a0 = o1_idx), 2 o2_half (a0 = o1_idx),
 * 3 u1 (a0 = o1_orb, a1 = exclude_first), 4 u2 (a0,a1,a2 = o1,o2,u1 orbitals), 5 u2_half (same).
 * h_rows is n x FRIES_MAX_SUB, h_len / h_norm have n entries. */
int fries_mol_hb_rows(fries_mol *mol, int which, const uint64_t *h_keys, const int32_t *h_args4, size_t n,
                      double *h_rows, int32_t *h_len, double *h_norm);
int fries_mol_hb_wt(fries_mol *mol, int normalized, const uint64_t *h_keys, const uint8_t *h_orbs, size_t n,
                    double *h_out);

/* ---- a10: apply_HBPP_sys FRIES/Hamiltonians/heat_bathPP.cpp:686-992 -----------------------------------
 * Host-buffer form: determinants + weights in, compressed Hamiltonian samples out, driven by the 5
 * uniforms the reference draws from mt19937 (:728,764,810,858,909).  spawn_cap = capacity of the
 * scratch (reference: spawn_length).  Outputs in the reference's order. */
int fries_apply_hbpp_sys(fries_mol *mol, const uint64_t *h_keys, const double *h_vals, size_t n, double p_doub,
                         int new_hb, const double *h_uniforms5, unsigned n_samp, size_t spawn_cap, double *h_out_val,
                         uint64_t *h_out_det, uint8_t *h_out_orbs, size_t out_cap, size_t *n_out);

/* apply_HBPP_piv FRIES/Hamiltonians/heat_bathPP.cpp:1014-1419 (spin_parity = 0): the pivotal twin of
 * fries_apply_hbpp_sys.  h_draws: the caller's next mt19937 outputs, consumed by the five piv_comp_parallel calls
 * (at most about 2 * (n_samp + 1) each); outputs as fries_apply_hbpp_sys, values carry the excitation's sign. */
int fries_apply_hbpp_piv(fries_mol *mol, const uint64_t *h_keys, const double *h_vals, size_t n, double p_doub,
                         int new_hb, const uint32_t *h_draws, size_t n_draws, size_t *n_draws_used, unsigned n_samp,
                         size_t spawn_cap, double *h_out_val, uint64_t *h_out_det, uint8_t *h_out_orbs, size_t out_cap,
                         size_t *n_out);

/* ---- a2/a3: the determinant store ---------------------------------------------------------------------
 * DistVec<double> ctor FRIES/vec_utils.hpp:154-198: capacity, n_bits (= 2 n_orb), n_elec, n_vecs rows,
 * proc_scrambler / vec_scrambler (n_bits uint32 each). */
int fries_vec_create(fries_ctx *ctx, size_t capacity, unsigned n_bits, unsigned n_elec, unsigned n_vecs,
                     const uint32_t *h_proc_scrambler, const uint32_t *h_vec_scrambler, int n_ranks, int rank,
                     fries_vec **out);
int fries_vec_destroy(fries_vec *vec);
/* DistVec::add + perform_add(origin) with curr_vec_idx = dest (vec_utils.hpp:418-440,606-641,991-1019)
 * for elements this rank owns.  ini flags: 0/1 bytes.  Elements are merged on the device. */
int fries_vec_add(fries_vec *vec, const uint64_t *h_keys, const double *h_vals, const uint8_t *h_ini, size_t n,
                  unsigned origin, unsigned dest);
/* same with resident spawn buffers: d_keys carry FRIES_INI_FLAG in bit 63; key == ~0 is skipped */
int fries_vec_add_dev(fries_vec *vec, const uint64_t *d_keys, const double *d_vals, size_t n, const uint32_t *d_n,
                      unsigned origin, unsigned dest);
int fries_vec_curr_size(fries_vec *vec, size_t *curr_size);
int fries_vec_n_nonz(fries_vec *vec, size_t *n_nonz);               /* DistVec::n_nonz */
int fries_vec_nonini_occ_add(fries_vec *vec, uint64_t *count);      /* DistVec::tot_sgn_coh :546-551 */
/* copy out storage: keys[curr_size], vals[n_vecs][curr_size] (row-major with row stride curr_size) */
int fries_vec_download(fries_vec *vec, uint64_t *h_keys, double *h_vals, size_t cap, size_t *n);
/* DistVec::load :761-844 from host arrays instead of files: replace the contents by n elements (vals = n_vecs rows
 * with row stride n), drop elements that are zero in every row, reset the diagonal cache, rebuild the index */
int fries_vec_upload(fries_vec *vec, const uint64_t *h_keys, const double *h_vals, size_t n);
/* DistVec::del_at_pos :458-476 for all flagged positions, followed by compaction of the storage
 * (stable) and a rebuild of the hash index */
int fries_vec_del(fries_vec *vec, const uint8_t *h_flags, size_t n);
/* DistVec::dot :228-238 against a (replicated) trial vector */
int fries_vec_dot(fries_vec *vec, const uint64_t *h_keys, const double *h_vals, size_t n, unsigned row, double *out);
/* DistVec::local_norm :683-689 of a row */
int fries_vec_local_norm(fries_vec *vec, unsigned row, double *out);
int fries_vec_two_norm(fries_vec *vec, unsigned row, double *out);   /* DistVec::two_norm :695-701: the sum of squares (the reference takes no root) */
/* row arithmetic on the stored elements (vec_utils.hpp:547-579).  op 0: add_vecs(dst, src, c)  dst += c * src;
 * 1: copy_vec(src, dst); 2: weight_vec(dst, src, expo = c)  dst *= (1 + |src|)^c; 3: zero_vec of row dst */
int fries_vec_row_op(fries_vec *vec, int op, unsigned dst, unsigned src, double c);
int fries_vec_set_diag_mol(fries_vec *vec, fries_mol *mol, double hf_en); /* diag_calc_ = diag_matrel - hf_en */

/* ---- a16: deterministic H.v (h_op_diag molecule.cpp:205-219 + h_op_offdiag :448-665) ---------------------
 * row dest <- id_fac * row src + h_fac * H * row src, for all elements of the store. */
int fries_h_apply(fries_vec *vec, fries_mol *mol, unsigned src, unsigned dest, double id_fac, double h_fac);
/* number of off-diagonal elements generated by the last fries_h_apply (the spawned H.v elements) */
int fries_h_apply_last_spawned(fries_vec *vec, uint64_t *n_spawned);
/* The same on a vector partitioned over ranks (collective): windows of <= seg_cap connections are stored straight into
 * their owners' receive windows (hb carries the direct route, fries_hbpp_set_route_p2p) and merged there. */
int fries_h_apply_routed(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, unsigned src, unsigned dest, double id_fac,
                         double h_fac, uint64_t *n_spawned);

/* ---- a17/a18 + drivers: one FRI iteration ---------------------------------------------------------------
 * frisys_mol loop body FRIES_bin/frisys_mol.cpp:405-552 (steps 1-11 of SURVEY.md 3.1) on resident
 * state.  uniforms6 = the 5 draws of apply_HBPP_sys followed by the vector-compression draw. */
typedef struct {
    double eps;          /* --epsilon */
    double init_thresh;  /* --initiator */
    double p_doub;       /* frisys_mol.cpp:217-220 */
    int new_hb;          /* --distribution HB_unnorm */
    unsigned matr_samp;  /* --mat_nonz */
    unsigned target_nonz;/* --vec_nonz */
    double en_shift;     /* current shift S */
} fries_frisys_params;
typedef struct {
    double glob_norm;    /* one-norm before compression (frisys_mol.cpp:503-504) */
    double numer, denom; /* projected energy pieces (:517-520) */
    uint64_t n_kept;     /* target_nonz - n_samp after find_preserve (:506) */
    uint64_t n_matrix_samples; /* comp_len after apply_HBPP_sys (:422) */
    uint64_t n_spawned;  /* elements handed to DistVec::add (:461) */
    uint64_t curr_size;  /* stored elements after the iteration */
} fries_iter_stats;
int fries_frisys_mol_setup(fries_vec *vec, fries_mol *mol, size_t spawn_cap, const uint64_t *h_trial_keys,
                           const double *h_trial_vals, size_t n_trial, const uint64_t *h_htrial_keys,
                           const double *h_htrial_vals, size_t n_htrial, fries_hbpp **out);
int fries_hbpp_destroy(fries_hbpp *hb);
int fries_frisys_mol_iterate(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, const fries_frisys_params *p,
                             const double *h_uniforms6, fries_iter_stats *stats);
/* frifull_mol loop body FRIES_bin/frifull_mol.cpp:256-320 */
typedef struct {
    double eps;
    unsigned target_nonz;
    double en_shift;        /* in: current shift; out: shift after this iteration's adjust_shift */
    /* adjust_shift (compress_utils.cpp:684-693) happens between find_preserve and sys_comp of the SAME iteration
     * (frifull_mol.cpp:270-276) and feeds h_op_diag of that iteration, so it is done inside the call */
    int adjust_shift;       /* nonzero on iterations with (iterat + 1) % shift_interval == 0 */
    double damp_factor;     /* shift_damping / shift_interval / eps */
    double target_norm;     /* --target */
    double last_one_norm;   /* in/out */
} fries_frifull_params;
/* single rank, or -- with the direct route set on hb -- collective over the ranks of a partitioned vector */
int fries_frifull_mol_iterate(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, fries_frifull_params *p,
                              double uniform, fries_iter_stats *stats);

/* ---- a19: Hubbard-Holstein (frisys_hh) ----------------------------------------------------------------------
 * Keys: bits [0, n) spin-up sites, [n, 2n) spin-down sites, then n phonon fields of ph_bits bits
 * (HubHolVec FRIES/hh_vec.hpp:27-29); hashes include the phonon numbers (hh_vec.hpp:56-88). */
int fries_vec_create_hh(fries_ctx *ctx, size_t capacity, unsigned n_sites, unsigned ph_bits, unsigned n_elec,
                        unsigned n_vecs, const uint32_t *h_proc_scrambler, const uint32_t *h_vec_scrambler, int n_ranks,
                        int rank, fries_vec **out);
int fries_vec_set_min_del_idx(fries_vec *vec, size_t idx); /* DistVec::set_min_del_idx */
/* Semi-stochastic calculations (DistVec::init_dense vec_utils.hpp:858-897, frisys_mol.cpp:347-401,414-419,479-485,
 * 503-504,533-536): the first n_dense stored determinants are the deterministic subspace.  fries_frisys_mol_iterate
 * then compresses and resamples only the rest, applies the dense determinants' columns of H exactly, and reports the
 * one-norm including the dense part.  Single rank. */
/* Reproducible merges (off by default; FRIES_DETERMINISTIC=1 in the environment turns it on for every new vector): new
 * determinants are appended and values added in the order of the batch -- the order of the reference's sequential
 * DistVec::add_elements (vec_utils.hpp:606-641) -- instead of the order in which the SMs reach them, so that two runs with
 * one seed agree bit for bit.  Single rank; slower (a scan and a sort of the batch per merge). */
int fries_vec_set_deterministic(fries_vec *vec, int on);
int fries_vec_set_dense(fries_vec *vec, size_t n_dense);
/* Several ranks: each rank's first n_dense stored determinants are its share of the subspace (DistVec::init_dense is
 * collective, vec_utils.hpp:858-897); n_dense_total = the sum over the ranks (one entry per rank in dense.txt). */
int fries_vec_set_dense_total(fries_vec *vec, size_t n_dense_total);
/* what 0: hub_diag hub_holstein.cpp:101-136 -> out[n]; 1: find_neighbors_1D hh_vec.hpp:139-175 as two bit masks per
 * state (hop to orb+1, hop to orb-1) -> out[2n]; 2: per-state terms of calc_ref_ovlp hub_holstein.hpp:93-182 -> out[n] */
int fries_hh_batch(fries_ctx *ctx, int what, const uint64_t *h_keys, const double *h_vals, size_t n, unsigned n_sites,
                   unsigned n_elec, unsigned ph_bits, uint64_t ref_key, double g_over_t, double *h_out);
typedef struct {
    double eps;            /* imaginary time step */
    double init_thresh;    /* --initiator */
    double hub_u, ph_freq, elec_ph, hf_en; /* U, omega, g, gs_energy of the parameter file (io_utils.cpp:320-408) */
    unsigned target_nonz;  /* --vec_nonz (also the matrix-compression budget, frisys_hh.cpp:202,222) */
    double en_shift;
    uint64_t ref_key;      /* Neel state gen_neel_det_1D hub_holstein.cpp:139-171 */
} fries_frisys_hh_params;
int fries_frisys_hh_setup(fries_vec *vec, size_t spawn_cap, fries_hbpp **out);
/* frisys_hh loop body FRIES_bin/frisys_hh.cpp:186-368; uniforms3 = stage 1, stage 2, vector compression;
 * stats: numer/denom as written to projnum.txt/projden.txt (denom = weight of the Neel state) */
int fries_frisys_hh_iterate(fries_vec *vec, fries_hbpp *hb, const fries_frisys_hh_params *p, const double *h_uniforms3,
                            fries_iter_stats *stats);
/* frifull_hh loop body FRIES_bin/frifull_hh.cpp:186-330: every stored state spawns all its hops (hub_all
 * hub_holstein.cpp:83-98) and phonon moves (no matrix compression), then death / cloning, find_preserve, the projected
 * energy against the Neel state and sys_comp with `uniform`.  Same parameter block (target_nonz = --vec_nonz); the
 * scratch of fries_frisys_hh_setup with spawn_cap >= 4 * n_elec. */
int fries_frifull_hh_iterate(fries_vec *vec, fries_hbpp *hb, const fries_frisys_hh_params *p, double uniform,
                             fries_iter_stats *stats);

/* ---- multi-GPU (one process per GPU; owner = hash_fxn(occ; proc_scrambler) % n_ranks, vec_utils.hpp:373-379) ----
 * Global reductions (sum_mpi compress_utils.hpp:179-231, the loc_norms Allgather) happen INSIDE the kernels through
 * peer-mapped inboxes: fries_comm_create returns a 64-byte CUDA IPC handle, the host all-gathers the handles
 * (torch.distributed) and passes all n_ranks x 64 bytes to fries_comm_connect.  The all-to-all of spawned elements
 * (Adder::perform_add vec_utils.hpp:991-1019) is issued by the host between _spawn and _finish on caller-owned
 * device buffers: send/recv_buf int64[n_ranks][2 * seg_cap] (keys | value bits per destination), send_counts
 * int64[n_ranks + 1] (last entry = elements that did not fit). */
typedef struct fries_comm fries_comm;
int fries_comm_create(fries_ctx *ctx, int n_ranks, int rank, fries_comm **out, void *h_ipc_handle64);
int fries_comm_connect(fries_comm *comm, const void *h_all_handles);
int fries_comm_destroy(fries_comm *comm);
int fries_comm_error(fries_comm *comm, uint64_t *epoch_of_failure);
/* diagnostics: average cost (us) of one in-kernel all-gather of n doubles + grid barrier, `iters` in a row, in a
 * cooperative grid of `ctas` CTAs (0 = the compression kernels' shape); collective over the ranks */
int fries_comm_pingpong(fries_comm *comm, int ctas, int iters, int n, double *us_per_exchange);
/* make fries_find_preserve_dev / fries_sys_comp_dev of this context collective over the ranks of comm */
int fries_ctx_set_comm(fries_ctx *ctx, fries_comm *comm);
int fries_hbpp_set_route(fries_hbpp *hb, fries_comm *comm, void *d_send_buf, void *d_recv_buf, void *d_send_counts,
                         size_t seg_cap);
/* Direct route (preferred): every rank owns a receive window [n_ranks sources][keys[seg_cap] | value bits[seg_cap]]
 * mapped into its peers (same IPC handshake as the inboxes).  The spawn kernel stores each element straight into its
 * owner's window over NVLink / NVSwitch -- routing is fused into the kernel that produces the elements -- and
 * _finish waits on per-source epoch flags; there is no collective call and no host round trip between _spawn and
 * _finish, and fries_frisys_mol_finish takes d_recv_counts = NULL. */
int fries_comm_route_create(fries_comm *comm, size_t seg_cap, void *h_ipc_handle64);
int fries_comm_route_connect(fries_comm *comm, const void *h_all_handles);
int fries_hbpp_set_route_p2p(fries_hbpp *hb, fries_comm *comm);
/* frisys_mol.cpp:405-471 up to Adder::add; then, after the all-to-all, :465-539 from add_elements on.
 * stats of _finish are global (summed over ranks in rank order). */
int fries_frisys_mol_spawn(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, const fries_frisys_params *p,
                           const double *h_uniforms6);
int fries_frisys_mol_finish(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, const fries_frisys_params *p,
                            const double *h_uniforms6, const void *d_recv_counts, fries_iter_stats *stats);

/* ---- diagnostics ------------------------------------------------------------------------------------------
 * Output list of stage `stage` (0..4) of apply_HBPP_sys: value, parent index, packed path bytes of the
 * parent item (orb_indices state) and chosen sub-index (comp_idx[.][1]).  Used by the parity tests to
 * localise a mismatch to one comp_sub call (heat_bathPP.cpp:731,767,813,861,912). */
int fries_debug_hbpp_stage(fries_mol *mol, const uint64_t *h_keys, const double *h_vals, size_t n, double p_doub,
                           int new_hb, const double *h_uniforms5, unsigned n_samp, size_t spawn_cap, int stage,
                           double *h_val, uint32_t *h_det, uint32_t *h_path, uint32_t *h_sub, size_t *n_out);

/* CompState records of the last iteration, 8 states x 20 doubles (loc_norm, glob_norm, new_norm, n_samp_left,
 * rounds, n_kept, n_out, n_in, anomalies, n_cand, fast, overflow, then 8 in-kernel phase time stamps in ns relative
 * to the first: 1 prep, 2 preserved set decided, 3 line scan, 4 count + offsets, 5 emit); states 0-4 = HB-PP stages, 5 = finalize,
 * 6 = find_preserve, 7 = sys_comp.  fast = 1: the preserved set came from the bracketed threshold solve (one data
 * pass + rounds on n_cand candidates), 0: from the plain rounds over the data. */
int fries_hbpp_states(fries_hbpp *hb, double *h_out160);
/* %globaltimer stamps (ns, relative to the first) inside the distributed candidate rounds 0-3 of state `state`:
 * per round start, before the grid barrier, after it, after the cross-rank exchange */
int fries_hbpp_round_stamps(fries_hbpp *hb, int state, double *h_out16);

/* Run every standalone (host-buffer) compression n times on the same inputs and return the last run: the bracketed
 * threshold solve starts from the fixed point of a previous run, which a one-shot call does not have. */
int fries_debug_set_repeat(int n);
/* the warm-up runs (all but the last) of a repeated call see their input values scaled by (1 + rel): the last run
 * then starts from a bracket that is off by about that much, as in consecutive FRI iterations */
int fries_debug_set_perturb(double rel);
/* on = 0: every compression uses the plain rounds (the reference's own iteration); on = 1 (default): bracketed solve
 * whenever a previous fixed point is available.  Both give the same preserved sets up to FP ties. */
int fries_debug_set_bracket(int on);
/* how many compressions of the last fries_comp_sub / fries_apply_hbpp_sys / fries_find_preserve call were decided by
 * the bracketed solve */
int fries_debug_last_fast(int *n_fast);
/* which build of the HB-PP stage kernels this process launches: 2 = two CTAs per SM (the product's configuration), 1 = the
 * one-CTA-per-SM measurement variant (FRIES_STAGE_CTAS=1 in the environment when the library first asks) */
int fries_debug_stage_ctas(int *ctas_per_sm);
/* generation of the compression engine the HB-PP stage kernels run (2 = csrc/compress2.cuh, the default; 1 with
 * FRIES_ENGINE=1 in the environment: the first-generation kernels, kept as a regression / measurement variant) */
int fries_debug_stage_engine(int *generation);
/* diagnostics / parity: the compression half of the fused vector kernel (csrc/vecphase.cu: find_preserve
 * compress_utils.cpp:29-105 -> sys_comp :278-327 -> del_at_pos vec_utils.hpp:458-476 + compaction) on the stored vector,
 * row 0; hb = the scratch of fries_frisys_mol_setup; h_state4: loc_norm, glob_norm, n_samp_left, n_kept */
int fries_debug_vec_phase(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, unsigned target_nonz, double uniform,
                          double *h_state4);
/* diagnostics: clock64 timeline (SM cycles) of thread 0 of CTA 0 through the last compression of state s (0-4: HB-PP stages) */
int fries_hbpp_timeline(fries_hbpp *hb, int s, double *h_out48);
/* diagnostics (FRIES_CTA_MARKS=1 in the environment): per-CTA phase-end times of stage s, h_out[8][*grid] in ns */
int fries_hbpp_cta_marks(fries_hbpp *hb, int s, double *h_out, int *grid);

#ifdef __cplusplus
}
#endif
#endif /* FRIES_B200_H */
