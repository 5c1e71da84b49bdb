// The vector half of a frisys_mol iteration as ONE persistent cooperative kernel (single rank):
//   step 7  death / cloning, add_vecs(0, 1), zero_vec(1)                  frisys_mol.cpp:488-499
//   step 8  find_preserve                                                  compress_utils.cpp:29-105, frisys_mol.cpp:503
//   step 10 <H trial | v>, <trial | v>                                     frisys_mol.cpp:517-520, vec_utils.hpp:228-253
//   step 11 sys_comp + del_at_pos of the zeroed elements                   compress_utils.cpp:278-327, frisys_mol.cpp:528-539
// Round 1 ran these as seven launches (death_axpy, find_preserve, state_to_r4, two dots, sys_comp, compact), each with its
// own grid barriers and re-reductions: 0.14 ms of a 0.68 ms iteration at 2.4e5 stored determinants, 3.4 ms of 15 at 1.25e7
// (2.2 ms of it the compaction, which cleared and rebuilt the 2^26-slot index one slot / one element per thread).
//
// Here: four passes over the vector, three barriers with payload (gridcomb.cuh), everything four elements per thread.
//   P1  death / cloning (lazily computed diagonal elements), |v| statistics against the threshold bracket of the previous
//       iteration (compress.cuh: bracket_solve), candidates inside the bracket staged in shared memory; grid-stride over
//       the elements (the determinants whose diagonal is still to be computed are contiguous at the end of the store)
//   --  barrier A: sums; the threshold is solved on the candidates -- by CTA 0 alone on a list of <= 4096
//       (compress2.cuh: bracket_solve2_local; result in every CTA's line), else by every CTA on the candidates it staged,
//       one more barrier per Newton round.  The preserved set is { |v| >= cut }: no keep flags are stored
//   P2  residual one-norm of the chunk; the two dot products through the still intact index, trial entry t on CTA t % grid
//   --  barrier B: residual norm + every chunk's place on the resampling line (+ <H trial|v>)
//   P3  the CTA's share of the index cleared; systematic resampling in storage order (one block scan per 2048 elements);
//       survivors counted
//   --  barrier C (with fence: the cleared index): survivor prefix (+ <trial|v>)
//   P4  stable compaction into the spare buffers + insertion into the index (four independent probe chains per thread)
// Without a valid bracket (first iteration, jump of the vector) the plain rounds of find_preserve run instead of the solve.
// Same arithmetic as the separate kernels (which stay: multi-rank, frifull_mol, the C-ABI's stand-alone entry points);
// prefix sums are associated per thread of four instead of per element, so resampling decisions agree up to FP-boundary
// ties (tests/test_gpu_vecphase.py counts them against the oracle).
#include "hbpp.cuh"
#include "vec.cuh"

extern int fr_bracket_on;  // hbpp.cu

extern __shared__ __align__(16) double fr_dyn_smem[];

#define VP_NT FR2_NT
#define VP_ITEMS FR2_ITEMS
#define VP_TILE FR2_TILE
#define VP_CAND_MASK ((1ull << 40) - 1)

struct VecPhaseArgs {
    VecView v;
    uint64_t *keys_b;
    double *vals_b, *diag_b;  // spare buffers (vals_b: n_vecs rows of cap)
    uint8_t *flags;           // [cap]: keep flags, then delete flags
    double hf_en, eps, shift;
    unsigned target_nonz;
    double uniform;
    unsigned long long n_dense;
    const uint64_t *trial_keys, *htrial_keys;
    const double *trial_vals, *htrial_vals;
    unsigned long long n_trial, n_htrial;
    KeepPred *pred;
    CandList cand;
    uint32_t *cand_idx;
    unsigned long long *gcomb;
    CompState *st6, *st7;
    double *scal;  // IterScalars layout (iter.cu): [0..3] R4, [4] numer, [5] denom, [43] dense norm
    int do_death;  // 0: compress only (diagnostics / parity test)
    unsigned long long *cta_marks;  // diagnostics (may be nullptr): [8][gridDim.x] %globaltimer at the phase ends
};
#define VP_MARK(k)                                                                                          \
    do {                                                                                                    \
        if (a.cta_marks && threadIdx.x == 0)                                                                \
            a.cta_marks[((k) < 8 ? (size_t)(k) * gridDim.x : (size_t)8 * 1024 + (size_t)((k)-8) * gridDim.x) + blockIdx.x] = fr_globaltimer(); \
    } while (0)

template <int MINCTAS>
__global__ void __launch_bounds__(VP_NT, MINCTAS) vec_phase_kernel(MolView gm, VecPhaseArgs a) {
    cg::grid_group grid = cg::this_grid();
    (void)grid;
    MolView m = mol_stage_shared_bulk(gm, fr_dyn_smem);
    __shared__ StageShared stg;
    __shared__ unsigned long long sm_bc[4];
    __shared__ double sm_wsum[FR2_NW + 1];
    __shared__ unsigned long long sm_wcnt[FR2_NW + 1];
    __shared__ GridCombShared gsh;
    __shared__ double sh_sd[6 * 33];
    __shared__ unsigned long long sh_sc[6 * 33];
    __shared__ uint32_t s_scr[64];
    const int tid = threadIdx.x;
    const VecView &v = a.v;
    const GridComb gcb{a.gcomb};
    CommView cm1;  // single rank
    cm1.n_ranks = 1;
    cm1.rank = 0;
    if (tid < 64) s_scr[tid] = v.scr_vec[tid];
    if (tid == 0) {
        unsigned long long n64 = *((volatile unsigned long long *)&v.cnt->n);
        sm_bc[0] = n64 < v.cap ? n64 : v.cap;
        sm_bc[1] = a.pred ? (unsigned long long)__double_as_longlong(__ldcg(&a.pred->t)) : 0ull;
        sm_bc[2] = a.pred ? (unsigned long long)__double_as_longlong(__ldcg(&a.pred->h)) : 0ull;
        sm_bc[3] = (unsigned long long)grid_comb_begin(gcb).epoch;
        stg.n_stage = 0;
    }
    __syncthreads();
    const size_t n = (size_t)sm_bc[0];
    const size_t nd = a.n_dense < n ? (size_t)a.n_dense : n;
    GridCombCursor gcur;
    gcur.epoch = (unsigned)sm_bc[3];
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 127) & ~(size_t)127;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;
    double *v0 = v.vals, *v1 = v.vals + v.cap;
    const double t_pred = __longlong_as_double((long long)sm_bc[1]), h_pred = __longlong_as_double((long long)sm_bc[2]);
    const bool try_fast = t_pred > 0 && h_pred > 0 && h_pred < 0.25;
    const double t_lo = t_pred * (1.0 - h_pred), t_hi = t_pred * (1.0 + h_pred);
    CandList2 cand{a.cand, a.cand_idx};
    const unsigned n_samp_in = a.target_nonz;

    VP_MARK(0);
    // ---- P1: death / cloning + statistics ----
    double s = 0, s_hi = 0;
    unsigned long long c_hi = 0, n_app = 0;
    // Grid-stride over the elements (this pass has no scan): the determinants added by the previous iteration's spawn sit at
    // the end of the store and are the ones whose diagonal element is still to be computed -- with a contiguous chunk per
    // CTA the last three CTAs did all of that (P1 ended at 10 us on the median CTA and at 65 us on them; with whole tiles
    // dealt round-robin still at 43 us; H2O-sized run, round 2).  Consecutive elements now go to consecutive threads of
    // consecutive CTAs.
    const size_t gthreads1 = (size_t)gridDim.x * VP_NT, gt1 = (size_t)blockIdx.x * VP_NT + tid;
    for (size_t base = gt1; base < n; base += (size_t)VP_ITEMS * gthreads1) {
        double av[VP_ITEMS], bv[VP_ITEMS], dv[VP_ITEMS];
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) {
            const size_t i = base + (size_t)k * gthreads1;
            av[k] = i < n ? v0[i] : 0.0;
            bv[k] = (i < n && a.do_death) ? v1[i] : 0.0;
            dv[k] = (i < n && a.do_death) ? v.diag[i] : 0.0;
        }
#pragma unroll 1
        for (int k = 0; k < VP_ITEMS; k++) {
            const size_t i = base + (size_t)k * gthreads1;
            if (i >= n) break;
            double x = av[k];
            if (a.do_death) {
                if (x != 0) {
                    double d = dv[k];
                    if (__builtin_expect(isnan(d), 0)) {  // DistVec::matr_el_at_pos vec_utils.hpp:672-677: computed on first use
                        uint8_t occ[FRIES_MAX_ELEC + 1];
                        mol_occ_list(v.keys[i], occ);
                        d = mol_diag(m, occ) - a.hf_en;
                        v.diag[i] = d;
                    }
                    x *= 1 - a.eps * (d - a.shift);
                }
                x += bv[k];
                v0[i] = x;
                if (bv[k] != 0) v1[i] = 0;
            }
            if (i >= nd) {
                const double mm = fabs(x);
                s += mm;
                if (try_fast && mm >= t_lo) {
                    if (mm >= t_hi) {
                        c_hi++;
                        s_hi += mm;
                    } else {
                        cand_stage(stg, cand, cm1, mm, 1u, (uint32_t)i);
                        n_app++;
                    }
                }
            }
        }
    }
    if (try_fast) cand_stage_flush(stg, cand, cm1);
    VP_MARK(1);
    double pre_d;
    unsigned long long pre_c, my_cand = 0, stage_overflows = 0;
    BracketResult br;
    br.valid = false;
    br.n_cand = 0;
    bool have_br = false;
    {
        double dd[2] = {s, s_hi};
        // the candidate count, and above bit 40 the number of CTAs whose staging area overflowed
        unsigned long long cc[2] = {c_hi, n_app};
        fr2_sum<2>(dd, cc, sh_sd, sh_sc);
        if (stg.n_stage > FR2_CAND_STAGE) cc[1] |= 1ull << 40;
        const unsigned tag = grid_comb_next_tag(gcur);
        gc_post<2>(gcb, tag, dd, cc, true);
        unsigned long long ex[6] = {0, 0, 0, 0, 0, 0};
        if (blockIdx.x == 0) {
            double td[2];
            unsigned long long tc[2];
            gc_reduce<2>(gcb, gsh, tag, td, tc, true);
            // lists beyond one CTA's registers are solved by all CTAs on their own staged candidates (below)
            if (try_fast && (tc[1] & VP_CAND_MASK) <= FR_CAND_CAP && (tc[1] >> 40) == 0) {
                BracketResult r = bracket_solve2_local(a.cand, td[0] - td[1], (long long)n_samp_in - (long long)tc[0], t_lo, t_hi,
                                                       sh_sc, cm1, nullptr, true, tc[1]);
                ex[0] = (unsigned long long)__double_as_longlong(r.x_cut);
                ex[1] = (unsigned long long)__double_as_longlong(r.R);
                ex[2] = (unsigned long long)r.nrem | ((unsigned long long)r.rounds << 32);
                ex[3] = r.kept_cand;
                ex[4] = (r.valid ? 1ull : 0ull) | (1ull << 8);
                ex[5] = r.n_cand;
            }
            gc_publish<2, 6>(gcb, tag, td, tc, ex);
        }
        gc_wait<2, 6>(gcb, gsh, tag, dd, cc, pre_d, pre_c, ex, true);
        gcur.epoch = tag;
        s = dd[0];
        s_hi = dd[1];
        c_hi = cc[0];
        my_cand = cc[1] & VP_CAND_MASK;
        stage_overflows = cc[1] >> 40;
        if ((ex[4] >> 8) == 1) {
            have_br = true;
            br.x_cut = __longlong_as_double((long long)ex[0]);
            br.R = __longlong_as_double((long long)ex[1]);
            br.nrem = (unsigned)ex[2];
            br.rounds = (unsigned)(ex[2] >> 32);
            br.kept_cand = ex[3];
            br.valid = (ex[4] & 1ull) != 0;
            br.n_cand = ex[5];
        }
    }
    VP_MARK(2);
    const double glob_total = s;
    unsigned nrem = n_samp_in;
    double R = 0, thr = INFINITY;
    unsigned rounds = 0;
    unsigned long long kept_total = 0, n_cand = 0;
    bool fast_done = false;
    if (try_fast) {
        if (!have_br && stage_overflows == 0) {
            // the list is longer than one CTA holds in registers (find_preserve at 1e6 elements brackets 1e4 - 6e4
            // candidates): every CTA runs the rounds on the candidates it staged itself (compress2.cuh)
            br = bracket_solve_staged(stg, gcb, gsh, gcur, glob_total - s_hi, (long long)n_samp_in - (long long)c_hi, t_lo, t_hi,
                                      sh_sd, sh_sc, my_cand);
        } else if (!have_br) {  // a staging area overflowed: the grid-distributed rounds of bracket_solve on the global list
            br = bracket_solve(grid, a.cand, a.st6->gacc, glob_total - s_hi, (long long)n_samp_in - (long long)c_hi, t_lo, t_hi,
                               sh_sd, sh_sc, cm1, nullptr, my_cand <= FR_CAND_GCAP);
        }
        n_cand = br.n_cand;
        if (br.valid) {
            // the preserved set is { |v| >= x_cut }: everything above the bracket, and the candidates the solve kept
            // (x_cut is the smallest of them, or the bracket's upper end) -- no per-element keep flag is stored
            kept_total = c_hi + br.kept_cand;
            thr = br.x_cut;
            nrem = br.nrem;
            R = br.R;
            rounds = br.rounds;
            fast_done = true;
            if (blockIdx.x == 0 && tid == 0) a.st6->fast = 1;
            __syncthreads();
        }
    }
    if (!fast_done) {
        // plain rounds (compress_utils.cpp:52-92): Newton from above on the threshold, one probe per round
        double loc = glob_total, R_next = glob_total;
        unsigned long long glob_sampled = 1;
        bool recalc = false;
        while (glob_sampled > 0 && rounds < 100000) {
            R = R_next;
            const double t0 = R / nrem;
            double sk[1] = {0.0};
            unsigned long long ck[1] = {0ull};
            if (R >= 0) {
                for (size_t i = (lo > nd ? lo : nd) + tid; i < hi; i += VP_NT) {
                    const double mm = fabs(v0[i]);
                    if (mm < thr && mm >= t0) {
                        sk[0] += mm;
                        ck[0]++;
                    }
                }
            }
            fr2_sum<1>(sk, ck, sh_sd, sh_sc);
            grid_comb<1>(gcb, gsh, gcur, sk, ck, false, false, pre_d, pre_c);
            if (ck[0] > 0) thr = t0;
            loc -= sk[0];
            R_next = loc;
            glob_sampled = ck[0];
            nrem -= (unsigned)ck[0];
            kept_total += ck[0];
            rounds++;
            if (glob_sampled == 0 && !recalc) {
                // exact recomputation of the residual norm (:78-90); the preserved set is { |v| >= thr }
                double t[1] = {0.0};
                unsigned long long dum[1] = {0ull};
                for (size_t i = (lo > nd ? lo : nd) + tid; i < hi; i += VP_NT) {
                    const double mm = fabs(v0[i]);
                    if (!(mm >= thr)) t[0] += mm;
                }
                fr2_sum<1>(t, dum, sh_sd, sh_sc);
                grid_comb<1>(gcb, gsh, gcur, t, dum, false, false, pre_d, pre_c);
                loc = t[0];
                R_next = t[0];
                glob_sampled = 1;
                recalc = true;
            } else {
                recalc = false;
            }
        }
        __syncthreads();
    }
    if (a.pred && blockIdx.x == 0 && tid == 0)
        keep_pred_update(a.pred, try_fast ? t_pred : 0.0, h_pred, nrem > 0 ? R / nrem : 0.0, n_cand);
    if (R < 1e-9) nrem = 0;  // compress_utils.cpp:94-96

    // ---- P2: residual one-norm of the chunk (fixed order) ----
    double cs = 0;
    for (size_t base = lo; base < hi; base += VP_TILE) {
        const size_t i0 = base + (size_t)tid * VP_ITEMS;
        double w4[VP_ITEMS];
        fr2_ld4(v0, i0, hi, w4);
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) {
            const size_t i = i0 + k;
            const bool live = i < hi && i >= nd && !(fabs(w4[k]) >= thr);
            cs += live ? fabs(w4[k]) : 0.0;
        }
    }
    // step 10 (frisys_mol.cpp:517-520), through the index as it is before this iteration's deletions: trial entry t is looked
    // up by CTA t % grid (one round of dependent loads per CTA instead of n_trial / 512 rounds on one CTA: 25 us there for the
    // ~3400 entries of H * HF, H2O-sized run, round 2).  <H trial|v> travels with barrier B, <trial|v> with barrier C.
    double dot_h = 0, dot_t = 0;
    if (a.do_death) {
        const unsigned long long tstep = (unsigned long long)VP_NT * gridDim.x;
        for (unsigned long long t = blockIdx.x + (unsigned long long)tid * gridDim.x; t < a.n_htrial; t += tstep) {
            const uint32_t pos = vec_lookup(v, a.htrial_keys[t], s_scr);
            if (pos != FRIES_NO_POS) dot_h += a.htrial_vals[t] * __ldcg(&v0[pos]);
        }
        for (unsigned long long t = blockIdx.x + (unsigned long long)tid * gridDim.x; t < a.n_trial; t += tstep) {
            const uint32_t pos = vec_lookup(v, a.trial_keys[t], s_scr);
            if (pos != FRIES_NO_POS) dot_t += a.trial_vals[t] * __ldcg(&v0[pos]);
        }
        if (blockIdx.x == 0) {  // DistVec::dense_norm vec_utils.hpp:903-917 (semi-stochastic runs only)
            double dn[1] = {0.0};
            unsigned long long dum[1] = {0ull};
            for (size_t i = tid; i < nd; i += VP_NT) dn[0] += fabs(__ldcg(&v0[i]));
            if (nd > 0) fr2_sum<1>(dn, dum, sh_sd, sh_sc);
            if (tid == 0) a.scal[43] = dn[0];
        }
    }
    VP_MARK(3);
    double blk_lb, loc_final;
    {
        double dd[2] = {cs, dot_h};
        unsigned long long cc[2] = {0ull, 0ull};
        fr2_sum<2>(dd, cc, sh_sd, sh_sc);
        grid_comb<2>(gcb, gsh, gcur, dd, cc, true, false, blk_lb, pre_c);
        loc_final = dd[0];
        if (a.do_death && blockIdx.x == 0 && tid == 0) a.scal[4] = dd[1];
    }
    VP_MARK(4);
    if (nrem == 0) loc_final = 0;
    const double G = loc_final;
    SysGrid sg;
    if (nrem > 0) {
        sg.unit = G / nrem;
        sg.rn0 = a.uniform * sg.unit;  // seed_sys :107-127 with nothing before this rank
        sg.inv = 1.0 / sg.unit;
        sg.n = (long long)nrem;
    } else {
        sg.rn0 = INFINITY;
        sg.unit = INFINITY;
        sg.inv = 0;
        sg.n = 0;
    }
    if (blockIdx.x == 0 && tid == 0) {
        a.st6->loc_norm = loc_final;
        a.st6->glob_norm = glob_total;
        a.st6->n_samp_left = nrem;
        a.st6->rounds = rounds;
        a.st6->n_kept = kept_total;
        a.st6->n_in = n - nd;
        a.st6->n_cand = n_cand;
        a.scal[0] = loc_final;
        a.scal[1] = glob_total;
        a.scal[2] = (double)nrem;
        a.scal[3] = (double)kept_total;
    }

    // ---- P3: systematic resampling in storage order ----
    // (+ this CTA's share of the index cleared: not before barrier B, because P2 reads the index for the dot products;
    // barrier C orders the clear before P4's insertions)
    {
        // 16 bytes per store, every thread of the grid; tkeys (8 B) and tpos (4 B) separately
        const size_t T = (size_t)v.tmask + 1, gthreads = (size_t)gridDim.x * VP_NT, gt = (size_t)blockIdx.x * VP_NT + tid;
        ulonglong2 *tk2 = reinterpret_cast<ulonglong2 *>(v.tkeys);
        const ulonglong2 e2 = make_ulonglong2(FRIES_EMPTY_KEY, FRIES_EMPTY_KEY);
        for (size_t q = gt; q < T / 2; q += gthreads) tk2[q] = e2;
        uint4 *tp4 = reinterpret_cast<uint4 *>(v.tpos);
        const uint4 p4 = make_uint4(FRIES_NO_POS, FRIES_NO_POS, FRIES_NO_POS, FRIES_NO_POS);
        for (size_t q = gt; q < T / 4; q += gthreads) tp4[q] = p4;
    }
    VP_MARK(11);  // index cleared
    double carry = blk_lb, new_norm = 0;
    unsigned long long n_samples = 0, n_surv = 0;
    const double nv = G / nrem;
    double nx4[VP_ITEMS];  // software pipeline: the next tile's loads are issued before this tile's scan and barrier
    {
        const size_t i0 = lo + (size_t)tid * VP_ITEMS;
        fr2_ld4(v0, i0, hi, nx4);
    }
    for (size_t base = lo; base < hi; base += VP_TILE) {
        const size_t i0 = base + (size_t)tid * VP_ITEMS;
        double x4[VP_ITEMS], m4[VP_ITEMS];
        uint8_t f4[VP_ITEMS];
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) {
            x4[k] = nx4[k];
            f4[k] = (i0 + k < hi && !(fabs(nx4[k]) >= thr)) ? 0 : 1;  // preserved exactly (or beyond the chunk)
        }
        if (base + VP_TILE < hi) fr2_ld4(v0, i0 + VP_TILE, hi, nx4);
        double tsum = 0;
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) {
            const size_t i = i0 + k;
            m4[k] = (i < hi && i >= nd && !f4[k]) ? fabs(x4[k]) : 0.0;
            tsum += m4[k];
        }
        double ex, tot;
        fr2_scan_d(tsum, ex, tot, sm_wsum);
        double start = carry + ex;
        bool changed = false;
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) {
            const size_t i = i0 + k;
            uint8_t del = 0;
            if (i < hi) {
                if (i < nd) {
                    // the dense subspace is neither compressed nor deleted
                } else if (f4[k]) {
                    new_norm += fabs(x4[k]);
                } else if (x4[k] != 0) {
                    const double lbound = start + m4[k];
                    const double g = sg.point_d(sg.count_below_d(start));
                    if (g < lbound) {
                        x4[k] = x4[k] > 0 ? nv : -nv;
                        new_norm += nv;
                        n_samples++;
                    } else {
                        x4[k] = 0;
                        del = 1;
                    }
                    changed = true;
                }
                a.flags[i] = del;
                n_surv += del ? 0 : 1;
            }
            start += m4[k];
        }
        if (changed) fr2_st4(v0, i0, hi, x4);
        carry += tot;
        __syncthreads();  // sm_wsum is rewritten by the next tile
    }
    VP_MARK(5);
    unsigned long long blk_off, total;
    {
        double dd[2] = {dot_t, new_norm};
        unsigned long long cc[2] = {n_surv, n_samples};
        fr2_sum<2>(dd, cc, sh_sd, sh_sc);
        double d0;
        grid_comb<2>(gcb, gsh, gcur, dd, cc, true, true, d0, blk_off);  // fence: the cleared index
        total = cc[0];
        if (blockIdx.x == 0 && tid == 0) {
            a.st7->new_norm = dd[1];
            a.st7->n_out = cc[1];
            if (a.do_death) a.scal[5] = dd[0];
        }
    }

    VP_MARK(6);
    // ---- P4: stable compaction into the spare buffers + index insertion ----
    unsigned long long ocarry = blk_off;
    for (size_t base = lo; base < hi; base += VP_TILE) {
        const size_t i0 = base + (size_t)tid * VP_ITEMS;
        uint8_t f4[VP_ITEMS];
        unsigned ts = 0;
        uint64_t key4[VP_ITEMS];
        double val4[VP_ITEMS], dg4[VP_ITEMS];
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) {  // everything the tile needs is in flight before the scan's barrier
            const bool in = i0 + k < hi;
            f4[k] = in ? a.flags[i0 + k] : 1;
            key4[k] = in ? v.keys[i0 + k] : FRIES_EMPTY_KEY;
            val4[k] = in ? v0[i0 + k] : 0.0;
            dg4[k] = in ? v.diag[i0 + k] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) ts += f4[k] ? 0u : 1u;
        unsigned ex, tot;
        fr2_scan_u(ts, ex, tot, sm_wcnt);
        if (base == lo) VP_MARK(8);  // first tile: loads + scan
        unsigned long long o = ocarry + ex;
        uint64_t slot4[VP_ITEMS];
        uint32_t pos4[VP_ITEMS];
#pragma unroll
        for (int k = 0; k < VP_ITEMS; k++) {
            if (!f4[k]) {
                const uint64_t key = key4[k];
                pos4[k] = (uint32_t)o;
                a.keys_b[o] = key;
                a.vals_b[o] = val4[k];
                a.vals_b[v.cap + o] = 0.0;  // row 1 is zero after step 7
                a.diag_b[o] = dg4[k];
                slot4[k] = vec_hash(v, key, s_scr) & v.tmask;
                o++;
            } else {
                key4[k] = FRIES_EMPTY_KEY;
            }
        }
        if (base == lo) VP_MARK(9);  // first tile: compacted copies stored, slots hashed
        // four independent probe chains per thread
        bool pending = true;
        while (pending) {
            pending = false;
#pragma unroll
            for (int k = 0; k < VP_ITEMS; k++) {
                if (key4[k] != FRIES_EMPTY_KEY) {
                    unsigned long long old = atomicCAS((unsigned long long *)&v.tkeys[slot4[k]], FRIES_EMPTY_KEY, key4[k]);
                    if (old == FRIES_EMPTY_KEY) {
                        v.tpos[slot4[k]] = pos4[k];
                        key4[k] = FRIES_EMPTY_KEY;
                    } else {
                        slot4[k] = (slot4[k] + 1) & v.tmask;
                        pending = true;
                    }
                }
            }
        }
        if (base == lo) VP_MARK(10);  // first tile: inserted into the index
        ocarry += tot;
        __syncthreads();
    }
    VP_MARK(7);
    if (blockIdx.x == 0 && tid == 0) v.cnt->n = total;
    grid_comb_end(gcb, gcur);
}

int fries_vec_phase_dev(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, double eps, double shift, unsigned target_nonz,
                        double uniform, bool do_death) {
    fries_ctx *c = vec->ctx;
    FRIES_REQUIRE(vec->n_ranks == 1 && vec->n_vecs == 2 && vec->hh_sites == 0, "vec_phase: single-rank molecular vectors with two rows");
    FRIES_REQUIRE(((vec->tsize) & 3) == 0, "vec_phase: index size");
    VecPhaseArgs a;
    a.v = vec->view();
    const int nb = vec->cur ^ 1;
    a.keys_b = vec->keys[nb].p;
    a.vals_b = vec->vals[nb].p;
    a.diag_b = vec->diag[nb].p;
    a.flags = hb->keep_flags.p;
    a.hf_en = vec->hf_en;
    a.eps = eps;
    a.shift = shift;
    a.target_nonz = target_nonz;
    a.uniform = uniform;
    a.n_dense = vec->n_dense;
    a.trial_keys = hb->trial_keys.p;
    a.trial_vals = hb->trial_vals.p;
    a.n_trial = hb->n_trial;
    a.htrial_keys = hb->htrial_keys.p;
    a.htrial_vals = hb->htrial_vals.p;
    a.n_htrial = hb->n_htrial;
    a.pred = fr_bracket_on ? hb->pred.p + 5 : nullptr;
    a.cand = CandList{hb->cand_x.p, hb->cand_m.p, &hb->st.p[6].n_cand};
    a.cand_idx = hb->cand_idx.p;
    a.gcomb = hb->gcomb.p;
    a.st6 = hb->st.p + 6;
    a.st7 = hb->st.p + 7;
    a.scal = hb->scal.p;
    a.do_death = do_death ? 1 : 0;
    const size_t smem = (size_t)mol->view.d.blob_doubles * 8;
    // two CTAs per SM (<= 64 registers) once an SM holds several tiles: the blocked passes keep one tile per CTA in flight
    static const int forced = [] {
        const char *e = getenv("FRIES_VECPHASE_CTAS");
        return (e && (e[0] == '1' || e[0] == '2') && e[1] == 0) ? e[0] - '0' : 0;
    }();
    const int ctas = forced ? forced : (vec->cap >= (size_t)1000000 ? 2 : 1);
    const void *kern = ctas == 2 ? (const void *)vec_phase_kernel<2> : (const void *)vec_phase_kernel<1>;
    static bool attr_set[2] = {false, false};
    if (!attr_set[ctas - 1]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set[ctas - 1] = true;
    }
    static const bool marks = getenv("FRIES_CTA_MARKS") != nullptr;
    if (marks && !hb->cta_marks.p) FRIES_TRY(hb->cta_marks.alloc((size_t)7 * 8 * 1024));
    a.cta_marks = marks ? hb->cta_marks.p + (size_t)5 * 8 * 1024 : nullptr;
    hb->grid_vp = c->sm_count * ctas;
    MolView gm = mol->view;
    void *args[] = {(void *)&gm, (void *)&a};
    {
        ProfScope ps(c, "vec_phase");
        CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(c->sm_count * ctas), dim3(VP_NT), args, smem, c->stream));
        c->launch_count++;
    }
    vec->cur = nb;
    return FRIES_OK;
}

// Diagnostics / parity: steps 8 + 11 (find_preserve -> sys_comp -> deletion + compaction) of the fused kernel alone, on the
// vector as stored (no death / cloning, no dot products).  hb: the scratch of fries_frisys_mol_setup.
extern "C" int fries_debug_vec_phase(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, unsigned target_nonz, double uniform,
                                     double *h_state4) {
    FRIES_REQUIRE(vec && mol && hb, "fries_debug_vec_phase: NULL argument");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemsetAsync(hb->st.p, 0, 8 * sizeof(CompState), c->stream));
    FRIES_TRY(fries_vec_phase_dev(vec, mol, hb, 0.0, 0.0, target_nonz, uniform, false));
    if (h_state4) {  // loc_norm, glob_norm, n_samp_left, n_kept
        CUDA_TRY(cudaMemcpyAsync(h_state4, hb->scal.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}
