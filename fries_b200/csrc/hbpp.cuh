// a10: heat-bath Power-Pitzer hierarchical compression of H's columns (apply_HBPP_sys
// heat_bathPP.cpp:686-992) as five persistent cooperative kernels + a finalize/spawn kernel.
//
// The reference materialises a spawn_length x n_states sub-weight matrix per stage (608 MB at
// mat_nonz = 1e6) and walks it sequentially.  Here every stage is one instance of comp_sub_engine
// (compress.cuh) whose provider recomputes a row of sub-weights on the fly from the parent
// determinant (8 B key) and the 4-byte choice path, with the small HB-PP tables staged in shared
// memory.  Per sample and stage the HBM traffic is the 40 B of SURVEY.md 8d (value, parent index,
// path; read + written), not the n_states x 8 B row.
#pragma once
#include "compress.cuh"
#include "compress2.cuh"
#include "molhost.cuh"
#include "hbpp_prov.cuh"


struct fries_hbpp {
    fries_ctx *ctx = nullptr;
    size_t cap = 0;
    DevBuf<double> veff, wtr, lb, rinv, oval[2], fin_val;
    DevBuf<uint32_t> ndiv, keep, kcnt, owidx[2], osub[2], det[2], path[2], fin_det, fin_orbs;
    DevBuf<uint8_t> nsub;
    DevBuf<double> part_d;
    DevBuf<unsigned long long> part_c;
    DevBuf<CompState> st;  // [0..4] stages, [5] finalize counters, [6] vector compression
    DevBuf<unsigned long long> n_scalar;
    // bracketed threshold solve (compress.cuh): candidate list (shared by all compressions of an iteration, which
    // run one after the other) and one prediction per compression site: [0..4] HB-PP stages, [5] find_preserve
    DevBuf<double> cand_x;
    DevBuf<uint32_t> cand_m, cand_idx;  // cand_idx: input index of every candidate (compress2.cuh)
    DevBuf<KeepPred> pred;
    DevBuf<unsigned long long> gcomb;      // GridComb state of the second-generation stage kernels (gridcomb.cuh)
    DevBuf<unsigned long long> cta_marks;  // diagnostics: [5 stages + vec_phase][8 marks][grid] (FRIES_CTA_MARKS=1)
    int grid_vp = 0;                       // grid of the last vec_phase launch
    // frisys/frifull driver state (iter.cu)
    DevBuf<uint64_t> trial_keys, htrial_keys, spawn_keys;
    DevBuf<double> trial_vals, htrial_vals, spawn_vals, scal;
    DevBuf<uint8_t> keep_flags;
    size_t n_trial = 0, n_htrial = 0;
    int grid = 0, grid2 = 0;  // cooperative grid of the stage kernels: first / second generation engine
    int grid2_s[5] = {0, 0, 0, 0, 0};  // ... of every stage (one or two CTAs per SM, hbpp.cu: stage2_ctas)
    fries_comm *comm = nullptr;  // multi-rank: peer-mapped inboxes (comm.cuh); not owned
    // multi-rank routing: per-destination send segments (keys | vals) and counters
    int64_t *send_buf = nullptr, *recv_buf = nullptr;  // caller-owned device buffers [n_ranks][2 * seg_cap]
    unsigned long long *send_counts_ext = nullptr;     // caller-owned [n_ranks + 1]: per-destination counts + overflow
    size_t seg_cap = 0;
    // direct route: the windows live in `comm` (fries_comm_route_create); counters are owned here
    bool dense_norm_set = false;  // scal[DENSE_NORM] holds a nonzero contribution of a dense subspace
    bool p2p = false;
    DevBuf<unsigned long long> p2p_send_counts, p2p_recv_counts;
};

// stages = false: only the reduction scratch and counters (spawn buffers are added by the caller)
int fries_hbpp_alloc(fries_ctx *c, size_t spawn_cap, fries_hbpp **out, bool stages = true);
extern "C" int fries_hbpp_destroy(fries_hbpp *hb);
// run the five stages on resident inputs; results: st[4].n_out samples in oval[1]/owidx[1]/osub[1]
// with paths det[0]/path[0] (stage 4 state) -- see hbpp.cu for the ping-pong convention
int fries_hbpp_stages_dev(fries_hbpp *hb, fries_mol *mol, const uint64_t *d_keys, const double *d_vals,
                          const unsigned long long *d_n, double p_doub, int new_hb, const double *uniforms5,
                          unsigned n_samp);
// finalize (:917-991).  spawn == nullptr: write (value, parent, orbitals) per sample, value 0 = failed.
struct HbSpawnArgs {
    const double *v0;        // parent values (sign + initiator test), frisys_mol.cpp:441-449
    double eps, init_thresh;
    uint64_t *out_keys;      // new determinant | FRIES_INI_FLAG, FRIES_EMPTY_KEY for failed samples
    double *out_vals;
    // multi-rank routing (Adder::add vec_utils.hpp:957-971): pack by owner = hash(proc_scrambler) % n_ranks into
    // per-destination segments [keys[seg_cap] | vals[seg_cap]] of the all-to-all send buffer
    int n_ranks;
    const uint32_t *proc_scr;              // device, 64 entries
    uint64_t *send_buf;                    // [n_ranks][2 * seg_cap]
    unsigned long long *send_counts;       // [n_ranks] + [n_ranks] = overflow counter
    unsigned long long seg_cap;
    // direct route (comm.cuh RouteView): peer_win[p] != nullptr -> elements owned by rank p are stored straight into
    // segment `rank` of p's receive window over NVLink instead of the local send buffer
    uint64_t *peer_win[FR_MAX_RANKS];
    int rank;
};
// cutoff: samples with |value| at or below it are dropped (1e-9 in apply_HBPP_sys :960,986; 1e-12 in apply_HBPP_piv)
int fries_hbpp_finalize_dev(fries_hbpp *hb, fries_mol *mol, const uint64_t *d_keys, double p_doub, int new_hb,
                            const HbSpawnArgs *spawn, double cutoff = 1e-9);

// piv.cu: piv_comp_parallel (single rank) in place on a resident vector; see there
int fries_piv_comp_resident(fries_ctx *c, double *d_vals, size_t n, uint32_t compress_size, uint8_t *d_keep,
                            const uint32_t *h_draws, size_t n_draws, size_t *used);
