// Molecular Hamiltonian on u64 determinants: Slater-Condon elements (a12/a13), symmetry bookkeeping
// (a11), excitation enumeration (a14) and the heat-bath Power-Pitzer weight rows (a8/a9).
//
// Everything here is a pure __host__ __device__ inline function of a MolView, so the kernels in
// mol.cu / hbpp.cu / iter.cu share one definition.  Reference: FRIES/Hamiltonians/molecule.cpp,
// heat_bathPP.cpp, near_uniform.cpp, FRIES/fci_utils.c (cited per function).  Loop and summation
// orders follow the reference so that FP64 results agree to the last bits, not just to 1e-12.
#pragma once
#include "common.cuh"

#define FR_N_IRREPS 8  // molecule.hpp: n_irreps

// tri-index macros of FRIES/math_utils.h:15-17
#define FR_TRI_N(n) ((n) * ((n) + 1) / 2)
#define FR_TRI_NODIAG(i, j) (FR_TRI_N((j)-1) + (i))
#define FR_TRI_WDIAG(i, j) (FR_TRI_N(j) + (i))

// Offsets (in doubles) of the small tables inside one contiguous blob, so that a kernel can stage
// the whole blob in shared memory with one cooperative copy.
struct MolDims {
    uint32_t n_orb;     // M: unfrozen spatial orbitals
    uint32_t n_elec;    // unfrozen electrons
    uint32_t n_frz;     // frozen electrons
    uint32_t tot_orb;   // M + n_frz / 2
    uint32_t max_n_symm;
    uint32_t blob_doubles;  // size of the small-table blob in doubles (incl. the byte tables)
    uint32_t off_d_diff, off_d_same, off_s_tens, off_exch_sqrt, off_diag_sqrt, off_exch_norms, off_symm, off_lookup;
    uint32_t off_irr;   // [8] u32: spatial orbitals of each irrep as a bit mask
    uint32_t off_xfull, off_dsfull;  // [M][M] square forms of exch_sqrt (diag_sqrt on the diagonal) and d_same (0 on it)
    double s_norm;
};

struct MolView {
    MolDims d;
    const double *eris;   // SymmERIs packed (FRIES/ndarr.hpp:206-244), tot_orb orbitals
    const double *hcore;  // tot_orb x tot_orb
    const double *d_diff, *d_same, *s_tens, *exch_sqrt, *diag_sqrt, *exch_norms;  // hb_info heat_bathPP.hpp:25-34
    const uint8_t *symm;    // [M] irreps
    const uint8_t *lookup;  // [8][M + 1] gen_symm_lookup molecule.cpp:1050-1065
    const uint32_t *irr_mask;  // [8] spatial orbitals of each irrep as a bit mask
    // square, symmetric copies for the row generators: one load at [a * M + b] instead of an ordered triangular index (the
    // index arithmetic was a third of a row entry's instructions); same values, so the rows are bit-identical
    const double *xfull, *dsfull;
};

__host__ __device__ __forceinline__ void mol_bind_blob(MolView &m, const double *blob) {
    m.d_diff = blob + m.d.off_d_diff;
    m.d_same = blob + m.d.off_d_same;
    m.s_tens = blob + m.d.off_s_tens;
    m.exch_sqrt = blob + m.d.off_exch_sqrt;
    m.diag_sqrt = blob + m.d.off_diag_sqrt;
    m.exch_norms = blob + m.d.off_exch_norms;
    m.symm = (const uint8_t *)(blob + m.d.off_symm);
    m.lookup = (const uint8_t *)(blob + m.d.off_lookup);
    m.irr_mask = (const uint32_t *)(blob + m.d.off_irr);
    m.xfull = blob + m.d.off_xfull;
    m.dsfull = blob + m.d.off_dsfull;
}

// offsets of the small tables inside the blob (one definition for fries_mol_create and the CPU test tier); returns the
// blob's size in doubles, a multiple of 2 (16 bytes: the stage kernels fetch it with one cp.async.bulk, molhost.cuh)
inline unsigned mol_blob_layout(MolDims &d) {
    const unsigned M = d.n_orb, TT = M * (M - 1) / 2;
    unsigned off = 0;
    d.off_d_diff = off; off += M * M;
    d.off_d_same = off; off += TT;
    d.off_s_tens = off; off += M;
    d.off_exch_sqrt = off; off += TT;
    d.off_diag_sqrt = off; off += M;
    d.off_exch_norms = off; off += M;
    d.off_symm = off; off += (M + 7) / 8;
    d.off_lookup = off; off += (FR_N_IRREPS * (M + 1) + 7) / 8;
    d.off_irr = off; off += (FR_N_IRREPS * 4 + 7) / 8;
    d.off_xfull = off; off += M * M;
    d.off_dsfull = off; off += M * M;
    off = (off + 1) & ~1u;
    d.blob_doubles = off;
    return off;
}
// entry (i, j) of the square forms from the triangular tables
__host__ __device__ __forceinline__ void mol_square_entry(const MolDims &d, double *blob, unsigned i, unsigned j) {
    const unsigned M = d.n_orb;
    const unsigned lo = i < j ? i : j, hi = i < j ? j : i;
    blob[d.off_xfull + i * M + j] = i == j ? blob[d.off_diag_sqrt + i] : blob[d.off_exch_sqrt + FR_TRI_NODIAG(lo, hi)];
    blob[d.off_dsfull + i * M + j] = i == j ? 0.0 : blob[d.off_d_same + FR_TRI_NODIAG(lo, hi)];
}

__host__ __device__ __forceinline__ uint8_t mol_lookup(const MolView &m, unsigned irrep, unsigned col) {
    return m.lookup[irrep * (m.d.n_orb + 1) + col];
}

// ---- SymmERIs::chemist / physicist ndarr.hpp:219-239 ------------------------------------------------
__host__ __device__ __forceinline__ double eri_chem(const MolView &m, unsigned i1, unsigned i2, unsigned i3,
                                                    unsigned i4) {
    unsigned min1 = i1 < i2 ? i1 : i2, max1 = i1 < i2 ? i2 : i1;
    unsigned p1 = FR_TRI_WDIAG(min1, max1);
    unsigned min2 = i3 < i4 ? i3 : i4, max2 = i3 < i4 ? i4 : i3;
    unsigned p2 = FR_TRI_WDIAG(min2, max2);
    size_t minp = p1 < p2 ? p1 : p2, maxp = p1 < p2 ? p2 : p1;
#ifdef __CUDA_ARCH__
    return __ldg(&m.eris[FR_TRI_WDIAG(minp, maxp)]);
#else
    return m.eris[FR_TRI_WDIAG(minp, maxp)];
#endif
}
__host__ __device__ __forceinline__ double eri_phys(const MolView &m, unsigned i1, unsigned i2, unsigned i3,
                                                    unsigned i4) {
    return eri_chem(m, i1, i3, i2, i4);
}
__host__ __device__ __forceinline__ double mol_hcore(const MolView &m, unsigned i, unsigned j) {
#ifdef __CUDA_ARCH__
    return __ldg(&m.hcore[i * m.d.tot_orb + j]);
#else
    return m.hcore[i * m.d.tot_orb + j];
#endif
}

// ---- a12: diag_matrel molecule.cpp:983-1029 -----------------------------------------------------------
__host__ __device__ inline double mol_diag(const MolView &m, const uint8_t *occ) {
    const unsigned hf = m.d.n_frz / 2, ne = m.d.n_elec, M = m.d.n_orb;
    double sum = 0;
    for (unsigned j = 0; j < hf; j++) {
        sum += mol_hcore(m, j, j) * 2;
        sum += eri_phys(m, j, j, j, j);
        for (unsigned k = j + 1; k < hf; k++) {
            sum += eri_phys(m, j, k, j, k) * 4;
            sum -= eri_phys(m, j, k, k, j) * 2;
        }
    }
    for (unsigned j = 0; j < ne / 2; j++) {
        unsigned e1 = occ[j] + hf;
        sum += mol_hcore(m, e1, e1);
        for (unsigned k = 0; k < hf; k++) {
            sum += eri_phys(m, e1, k, e1, k) * 2;
            sum -= eri_phys(m, e1, k, k, e1);
        }
        for (unsigned k = j + 1; k < ne / 2; k++) {
            unsigned e2 = occ[k] + hf;
            sum += eri_phys(m, e1, e2, e1, e2);
            sum -= eri_phys(m, e1, e2, e2, e1);
        }
        for (unsigned k = ne / 2; k < ne; k++) {
            unsigned e2 = occ[k] - M + hf;  // occ + n_frozen - n_orbs with n_orbs = tot_orb
            sum += eri_phys(m, e1, e2, e1, e2);
        }
    }
    for (unsigned j = ne / 2; j < ne; j++) {
        unsigned e1 = occ[j] - M + hf;
        sum += mol_hcore(m, e1, e1);
        for (unsigned k = 0; k < hf; k++) {
            sum += eri_phys(m, e1, k, e1, k) * 2;
            sum -= eri_phys(m, e1, k, k, e1);
        }
        for (unsigned k = j + 1; k < ne; k++) {
            unsigned e2 = occ[k] - M + hf;
            sum += eri_phys(m, e1, e2, e1, e2);
            sum -= eri_phys(m, e1, e2, e2, e1);
        }
    }
    return sum;
}

// ---- a13: sing_matr_el_nosgn molecule.cpp:76-105, doub_matr_el_nosgn :26-42 -----------------------------
__host__ __device__ inline double mol_sing_el(const MolView &m, unsigned o, unsigned v, const uint8_t *occ) {
    const unsigned hf = m.d.n_frz / 2, ne = m.d.n_elec, M = m.d.n_orb;
    unsigned occ_spa = (o % M) + hf, unocc_spa = (v % M) + hf, occ_spin = o / M;
    double el = mol_hcore(m, occ_spa, unocc_spa);
    for (unsigned j = 0; j < hf; j++) {
        el += eri_phys(m, occ_spa, j, unocc_spa, j) * 2;
        el -= eri_phys(m, occ_spa, j, j, unocc_spa);
    }
    for (unsigned j = 0; j < ne / 2; j++) {
        unsigned q = occ[j] + hf;
        el += eri_phys(m, occ_spa, q, unocc_spa, q);
        if (occ_spin == 0) el -= eri_phys(m, occ_spa, q, q, unocc_spa);
    }
    for (unsigned j = ne / 2; j < ne; j++) {
        unsigned q = occ[j] - M + hf;
        el += eri_phys(m, occ_spa, q, unocc_spa, q);
        if (occ_spin == 1) el -= eri_phys(m, occ_spa, q, q, unocc_spa);
    }
    return el;
}
// the same sum, same order, with the occupied orbitals taken from the two spin masks (ascending) instead of a list
__host__ __device__ inline double mol_sing_el_bits(const MolView &m, unsigned o, unsigned v, uint32_t occ_a, uint32_t occ_b) {
    const unsigned hf = m.d.n_frz / 2, M = m.d.n_orb;
    unsigned occ_spa = (o % M) + hf, unocc_spa = (v % M) + hf, occ_spin = o / M;
    double el = mol_hcore(m, occ_spa, unocc_spa);
    for (unsigned j = 0; j < hf; j++) {
        el += eri_phys(m, occ_spa, j, unocc_spa, j) * 2;
        el -= eri_phys(m, occ_spa, j, j, unocc_spa);
    }
    for (uint32_t am = occ_a; am; am &= am - 1) {
        unsigned q = (unsigned)fr_ctz(am) + hf;
        el += eri_phys(m, occ_spa, q, unocc_spa, q);
        if (occ_spin == 0) el -= eri_phys(m, occ_spa, q, q, unocc_spa);
    }
    for (uint32_t bm = occ_b; bm; bm &= bm - 1) {
        unsigned q = (unsigned)fr_ctz(bm) + hf;
        el += eri_phys(m, occ_spa, q, unocc_spa, q);
        if (occ_spin == 1) el -= eri_phys(m, occ_spa, q, q, unocc_spa);
    }
    return el;
}
__host__ __device__ inline double mol_doub_el(const MolView &m, const uint8_t *orbs) {
    const unsigned hf = m.d.n_frz / 2, M = m.d.n_orb;
    bool same_sp = (orbs[0] / M) == (orbs[1] / M);
    unsigned sp0 = (orbs[0] % M) + hf, sp1 = (orbs[1] % M) + hf, sp2 = (orbs[2] % M) + hf, sp3 = (orbs[3] % M) + hf;
    double el = eri_phys(m, sp0, sp1, sp2, sp3);
    if (same_sp) el -= eri_phys(m, sp0, sp1, sp3, sp2);
    return el;
}

// ---- a11: symmetry bookkeeping ----------------------------------------------------------------------------
// count_symm_virt near_uniform.cpp:14-28: unoccupied orbitals per (irrep, spin)
__host__ __device__ inline void mol_count_symm_virt(const MolView &m, const uint8_t *occ, uint8_t cnt[FR_N_IRREPS][2]) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb;
    for (unsigned i = 0; i < FR_N_IRREPS; i++) cnt[i][0] = cnt[i][1] = mol_lookup(m, i, 0);
    unsigned i = 0;
    for (; i < ne / 2; i++) cnt[m.symm[occ[i]]][0] -= 1;
    for (; i < ne; i++) cnt[m.symm[occ[i] - M]][1] -= 1;
}
// count_sing_allowed near_uniform.cpp:316-327
__host__ __device__ inline unsigned mol_count_sing_allowed(const MolView &m, const uint8_t *occ,
                                                           const uint8_t cnt[FR_N_IRREPS][2]) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb;
    unsigned n = 0;
    for (unsigned e = 0; e < ne; e++)
        if (cnt[m.symm[occ[e] % M]][e / (ne / 2)] != 0) n++;
    return n;
}
// count_sing_virt near_uniform.cpp:330-347: occ_choice in: index among allowed electrons; out: electron index
__host__ __device__ inline unsigned mol_count_sing_virt(const MolView &m, const uint8_t *occ,
                                                        const uint8_t cnt[FR_N_IRREPS][2], uint8_t *occ_choice) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb;
    unsigned n = 0;
    for (unsigned e = 0; e < ne; e++) {
        unsigned va = cnt[m.symm[occ[e] % M]][e / (ne / 2)];
        if (va != 0) {
            if (n == *occ_choice) {
                *occ_choice = (uint8_t)e;
                return va;
            }
            n++;
        }
    }
    return 0;
}
// The same three on bit masks (no occupied list, no per-irrep counter array in local memory): the virtual orbitals of
// irrep r and spin s are popc(virt_s & irr_mask[r]); an electron is "allowed" when its irrep has a virtual orbital of
// its spin.  mol_sing_allowed_masks returns the allowed occupied orbitals of both spins.
__host__ __device__ __forceinline__ void mol_sing_allowed_masks(const MolView &m, uint32_t occ_a, uint32_t occ_b,
                                                                uint32_t &al_a, uint32_t &al_b) {
    const uint32_t all = (uint32_t)((1ull << m.d.n_orb) - 1);
    const uint32_t va = ~occ_a & all, vb = ~occ_b & all;
    uint32_t oa = 0, ob = 0;
    for (unsigned r = 0; r < FR_N_IRREPS; r++) {
        const uint32_t im = m.irr_mask[r];
        if (va & im) oa |= im;
        if (vb & im) ob |= im;
    }
    al_a = occ_a & oa;
    al_b = occ_b & ob;
}
// count_sing_allowed
__host__ __device__ __forceinline__ unsigned mol_count_sing_allowed_bits(const MolView &m, uint32_t occ_a, uint32_t occ_b) {
    uint32_t al_a, al_b;
    mol_sing_allowed_masks(m, occ_a, occ_b, al_a, al_b);
    return (unsigned)(fr_popc((uint64_t)al_a) + fr_popc((uint64_t)al_b));
}

// virt_from_idx near_uniform.cpp:419-433
__host__ __device__ inline unsigned mol_virt_from_idx(const MolView &m, uint64_t det, unsigned irrep, unsigned spin_shift,
                                                      unsigned index) {
    unsigned n = mol_lookup(m, irrep, 0);
    for (unsigned s = 0; s < n; s++) {
        unsigned orb = spin_shift + mol_lookup(m, irrep, 1 + s);
        if (!fr_read_bit(det, orb)) {
            if (index == 0) return orb;
            index--;
        }
    }
    return 255;
}
// find_nth_virt fci_utils.c:138-148 (occ must be readable at index n_elec: callers pad with 255)
__host__ __device__ inline unsigned mol_find_nth_virt(const uint8_t *occ, unsigned spin, unsigned n_elec, unsigned n_orb,
                                                      unsigned n) {
    unsigned virt = (n_orb * spin + n) & 0xff;
    for (unsigned i = n_elec / 2 * spin; i < n_elec && occ[i] <= virt; i++) virt = (virt + 1) & 0xff;
    return virt;
}
// count_singex molecule.cpp:914-933
__host__ __device__ inline unsigned mol_count_singex(const MolView &m, uint64_t det, const uint8_t *occ) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb;
    unsigned n = 0;
    for (unsigned e = 0; e < ne; e++) {
        unsigned orb = occ[e], irrep = m.symm[orb % M], spin = orb / M;
        unsigned ns = mol_lookup(m, irrep, 0);
        for (unsigned s = 0; s < ns; s++)
            if (!fr_read_bit(det, mol_lookup(m, irrep, s + 1) + M * spin)) n++;
    }
    return n;
}

// ---- a14: sing_ex_symm molecule.cpp:178-203 / doub_ex_symm :108-175 ------------------------------------------
// Visitors are called in the reference's enumeration order; returning the count.  F: void(o, v) / void(o0,o1,v0,v1)
template <class F>
__host__ __device__ inline unsigned mol_for_each_sing(const MolView &m, uint64_t det, const uint8_t *occ, F &&f) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb;
    unsigned n = 0;
    for (unsigned i = 0; i < ne / 2; i++) {
        unsigned io = occ[i];
        for (unsigned a = 0; a < M; a++)
            if (!fr_read_bit(det, a) && m.symm[io] == m.symm[a]) {
                f(io, a);
                n++;
            }
    }
    for (unsigned i = ne / 2; i < ne; i++) {
        unsigned io = occ[i];
        for (unsigned a = M; a < 2 * M; a++)
            if (!fr_read_bit(det, a) && m.symm[io - M] == m.symm[a - M]) {
                f(io, a);
                n++;
            }
    }
    return n;
}
template <class F>
__host__ __device__ inline unsigned mol_for_each_doub(const MolView &m, uint64_t det, const uint8_t *occ, F &&f) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb;
    unsigned n = 0;
    for (unsigned i = 0; i < ne / 2; i++) {  // different spin
        unsigned io = occ[i];
        for (unsigned j = ne / 2; j < ne; j++) {
            unsigned jo = occ[j];
            unsigned sij = m.symm[io] ^ m.symm[jo - M];
            for (unsigned k = 0; k < M; k++) {
                if (fr_read_bit(det, k)) continue;
                unsigned sijk = sij ^ m.symm[k];
                for (unsigned l = M; l < 2 * M; l++)
                    if (!fr_read_bit(det, l) && (sijk ^ m.symm[l - M]) == 0) {
                        f(io, jo, k, l);
                        n++;
                    }
            }
        }
    }
    for (unsigned i = 0; i < ne / 2; i++) {  // up-up
        unsigned io = occ[i];
        for (unsigned j = i + 1; j < ne / 2; j++) {
            unsigned jo = occ[j];
            unsigned sij = m.symm[io] ^ m.symm[jo];
            for (unsigned k = 0; k < M; k++) {
                if (fr_read_bit(det, k)) continue;
                unsigned sijk = sij ^ m.symm[k];
                for (unsigned l = k + 1; l < M; l++)
                    if (!fr_read_bit(det, l) && (sijk ^ m.symm[l]) == 0) {
                        f(io, jo, k, l);
                        n++;
                    }
            }
        }
    }
    for (unsigned i = ne / 2; i < ne; i++) {  // down-down
        unsigned io = occ[i];
        for (unsigned j = i + 1; j < ne; j++) {
            unsigned jo = occ[j];
            unsigned sij = m.symm[io - M] ^ m.symm[jo - M];
            for (unsigned k = M; k < 2 * M; k++) {
                if (fr_read_bit(det, k)) continue;
                unsigned sijk = sij ^ m.symm[k - M];
                for (unsigned l = k + 1; l < 2 * M; l++)
                    if (!fr_read_bit(det, l) && (sijk ^ m.symm[l - M]) == 0) {
                        f(io, jo, k, l);
                        n++;
                    }
            }
        }
    }
    return n;
}

// ---- a8: HB-PP weight rows heat_bathPP.cpp:182-412 -------------------------------------------------------------
// Streaming generators: hbs_*(..., f) call f(j, raw_j) for the entries j = 0, 1, ... of one row in index order with
// the UN-normalised weight; the reference stores raw_j * (1 / norm).  Everything is derived from the u64 key with
// popcount / find-nth-set-bit, so a kernel never materialises the row or the occupied list (no local memory).
__host__ __device__ __forceinline__ unsigned fr_nth_bit32(uint32_t mask, unsigned n) {
#ifdef __CUDA_ARCH__
    // position of the n-th (0-based) set bit by a popcount binary search: 5 branch-free steps (__fns is a
    // software loop of ~100 instructions)
    unsigned pos = 0, t;
    t = __popc(mask & 0xffffu);
    if (n >= t) { n -= t; pos = 16; mask >>= 16; }
    t = __popc(mask & 0xffu);
    if (n >= t) { n -= t; pos += 8; mask >>= 8; }
    t = __popc(mask & 0xfu);
    if (n >= t) { n -= t; pos += 4; mask >>= 4; }
    t = __popc(mask & 0x3u);
    if (n >= t) { n -= t; pos += 2; mask >>= 2; }
    if (n >= (mask & 1u)) pos += 1;
    return pos;
#else
    for (unsigned i = 0; i < n; i++) mask &= mask - 1;
    return (unsigned)__builtin_ctz(mask);
#endif
}
struct OccMask {
    uint32_t a, b;  // spatial-orbital occupation of the two spin blocks
};
__host__ __device__ __forceinline__ OccMask mol_occ_mask(const MolView &m, uint64_t key) {
    uint32_t all = (uint32_t)((1ull << m.d.n_orb) - 1);
    OccMask o;
    o.a = (uint32_t)(key & all);
    o.b = (uint32_t)((key >> m.d.n_orb) & all);
    return o;
}
// spin orbital of the idx-th electron (alpha block first, ascending)
__host__ __device__ __forceinline__ unsigned mol_elec_orb(const MolView &m, const OccMask &o, unsigned idx) {
    const unsigned h = m.d.n_elec / 2;
    return idx < h ? fr_nth_bit32(o.a, idx) : m.d.n_orb + fr_nth_bit32(o.b, idx - h);
}


// Rows are streamed in chunks of FR_ROW_CHUNK entries: the table look-ups of a chunk are issued together (independent
// shared-memory loads in flight), then the chunk's entries go to the visitor in order.  A visitor's arithmetic is a serial
// FP64 chain (running norm, running prefix); with one look-up per iteration every entry paid the load latency inside
// that chain (~100 cycles per entry and warp, measured round 2), which bounded every stage kernel.
// gen(j) returns the raw weight of entry j and is called for j = 0 .. n - 1 in order (it may carry state, e.g. a bit
// mask whose lowest set bit it consumes); a negative return value ends the row before entry j.
#define FR_ROW_CHUNK 4
template <class G, class F>
__host__ __device__ __forceinline__ void fr_row_chunked(unsigned n, G &&gen, F &&f) {
    for (unsigned j0 = 0; j0 < n; j0 += FR_ROW_CHUNK) {
        double raw[FR_ROW_CHUNK];
#pragma unroll
        for (int q = 0; q < FR_ROW_CHUNK; q++) raw[q] = j0 + q < n ? gen(j0 + q) : -1.0;
#pragma unroll
        for (int q = 0; q < FR_ROW_CHUNK; q++) {
            if (j0 + q >= n || raw[q] < 0) return;
            if (!fr_emit(f, j0 + q, raw[q])) return;
        }
    }
}

// calc_o1_probs :182-200: one entry per electron (the first is skipped when exclude_first); norm = sum in index order
template <class F>
__host__ __device__ __forceinline__ void hbs_o1(const MolView &m, uint64_t key, int exclude_first, F &&f) {
    OccMask o = mol_occ_mask(m, key);
    uint32_t am = o.a, bm = o.b;
    if (exclude_first > 0) am &= am - 1;
    const unsigned na = (unsigned)fr_popc((uint64_t)am), n = na + (unsigned)fr_popc((uint64_t)bm);
    fr_row_chunked(n, [&](unsigned j) {
        uint32_t &mk = j < na ? am : bm;
        const unsigned q = (unsigned)fr_ctz(mk);
        mk &= mk - 1;
        return m.s_tens[q];
    }, f);
}
// calc_o2_probs_half :236-270: entries for the electrons before o1_idx
template <class F>
__host__ __device__ __forceinline__ void hbs_o2_half(const MolView &m, uint64_t key, unsigned o1_idx, F &&f) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb, h = ne / 2;
    OccMask o = mol_occ_mask(m, key);
    const unsigned o1 = mol_elec_orb(m, o, o1_idx);
    // entries j < o1_idx: the alpha electrons before o1 (j < h), then the beta electrons before it (j >= h)
    uint32_t am = o.a, bm = o.b;
    fr_row_chunked(o1_idx, [&](unsigned j) {
        if (j < h) {
            const unsigned q = (unsigned)fr_ctz(am);
            am &= am - 1;
            return o1 < M ? m.dsfull[o1 * M + q] : m.d_diff[(o1 - M) * M + q];
        }
        const unsigned q = (unsigned)fr_ctz(bm);
        bm &= bm - 1;
        return o1 < M ? m.d_diff[o1 * M + q] : m.dsfull[(o1 - M) * M + q];
    }, f);
}
// calc_o2_probs :203-233: ne entries in index order (entry o1_idx is 0)
template <class F>
__host__ __device__ __forceinline__ void hbs_o2(const MolView &m, uint64_t key, unsigned o1_idx, F &&f) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb, h = ne / 2;
    OccMask o = mol_occ_mask(m, key);
    unsigned o1 = mol_elec_orb(m, o, o1_idx), o1_spin = o1 / M, o1s = o1 % M;
    for (unsigned j = 0; j < ne; j++) {
        unsigned q = mol_elec_orb(m, o, j) % M;
        double raw;
        if (j / h != o1_spin) raw = m.d_diff[o1s * M + q];
        else if (j < o1_idx) raw = m.d_same[FR_TRI_NODIAG(q, o1s)];
        else if (j > o1_idx) raw = m.d_same[FR_TRI_NODIAG(o1s, q)];
        else raw = 0;
        if (!fr_emit(f, j, raw)) return;
    }
}
// its norm is accumulated opposite-spin block first, then the same-spin entries (:211-224)
__host__ __device__ __forceinline__ double hbs_o2_norm(const MolView &m, uint64_t key, unsigned o1_idx) {
    const unsigned ne = m.d.n_elec, M = m.d.n_orb, h = ne / 2;
    OccMask o = mol_occ_mask(m, key);
    unsigned o1 = mol_elec_orb(m, o, o1_idx), o1_spin = o1 / M, o1s = o1 % M;
    double norm = 0;
    unsigned off = (1 - o1_spin) * h;
    for (unsigned j = off; j < h + off; j++) norm += m.d_diff[o1s * M + mol_elec_orb(m, o, j) % M];
    off = o1_spin * h;
    for (unsigned j = off; j < o1_idx; j++) norm += m.d_same[FR_TRI_NODIAG(mol_elec_orb(m, o, j) % M, o1s)];
    for (unsigned j = o1_idx + 1; j < h + off; j++) norm += m.d_same[FR_TRI_NODIAG(o1s, mol_elec_orb(m, o, j) % M)];
    return norm;
}
// calc_u1_probs :273-319: one entry per virtual orbital of o1's spin, ascending (raw values; the caller applies
// exclude_first: norm -= raw_0, entry 0 = 0)
template <class F>
__host__ __device__ __forceinline__ void hbs_u1(const MolView &m, uint64_t key, unsigned o1_orb, F &&f) {
    const unsigned M = m.d.n_orb;
    OccMask o = mol_occ_mask(m, key);
    const unsigned o1s = o1_orb % M;
    uint32_t vm = ~(o1_orb / M ? o.b : o.a) & (uint32_t)((1ull << M) - 1);
    fr_row_chunked((unsigned)fr_popc((uint64_t)vm), [&](unsigned) {
        const unsigned k = (unsigned)fr_ctz(vm);
        vm &= vm - 1;
        return m.xfull[o1s * M + k];
    }, f);
}
__host__ __device__ __forceinline__ double hbs_u2_weight(const MolView &m, unsigned o2s, unsigned u2) {
    return m.xfull[o2s * m.d.n_orb + u2];  // diag_sqrt on the diagonal, exch_sqrt off it
}
// calc_u2_probs :322-365: one entry per orbital of the irrep that completes the double (0 for u2 == u1, same spin)
template <class F>
__host__ __device__ __forceinline__ void hbs_u2(const MolView &m, unsigned o1_orb, unsigned o2_orb, unsigned u1_orb, F &&f) {
    const unsigned M = m.d.n_orb;
    const unsigned o2s = o2_orb % M, u1s = u1_orb % M;
    const bool same = (o1_orb / M) == (o2_orb / M);
    const unsigned irrep = m.symm[o1_orb % M] ^ m.symm[o2s] ^ m.symm[u1s];
    fr_row_chunked(mol_lookup(m, irrep, 0), [&](unsigned i) {
        const unsigned u2 = mol_lookup(m, irrep, i + 1);
        return ((same && u2 != u1s) || !same) ? hbs_u2_weight(m, o2s, u2) : 0.0;
    }, f);
}
// calc_u2_probs_half :368-412: stops at u2 >= u1 for same-spin pairs; occupied u2 get 0
template <class F>
__host__ __device__ __forceinline__ void hbs_u2_half(const MolView &m, unsigned o1_orb, unsigned o2_orb, unsigned u1_orb,
                                                     uint64_t det, F &&f) {
    const unsigned M = m.d.n_orb;
    const unsigned o2s = o2_orb % M, u1s = u1_orb % M, u2_spin = o2_orb / M;
    const bool same = (o1_orb / M) == u2_spin;
    const unsigned irrep = m.symm[o1_orb % M] ^ m.symm[o2s] ^ m.symm[u1s];
    fr_row_chunked(mol_lookup(m, irrep, 0), [&](unsigned i) {
        const unsigned u2 = mol_lookup(m, irrep, i + 1);
        if (same && u2 >= u1s) return -1.0;  // the row ends here (the weights themselves are never negative)
        const bool ok = ((same && u2 != u1s) || !same) && !fr_read_bit(det, u2 + M * u2_spin);
        return ok ? hbs_u2_weight(m, o2s, u2) : 0.0;
    }, f);
}

// Array forms with the reference's signatures and return values (parity entry points, finalize-free code).
__host__ __device__ __forceinline__ uint64_t mol_key_of_occ(const MolView &m, const uint8_t *occ) {
    uint64_t k = 0;
    for (unsigned i = 0; i < m.d.n_elec; i++) k |= 1ull << occ[i];
    return k;
}
__host__ __device__ inline double hb_o1_probs(const MolView &m, double *p, const uint8_t *occ, int exclude_first) {
    uint64_t key = mol_key_of_occ(m, occ);
    double norm = 0;
    unsigned n = 0;
    hbs_o1(m, key, exclude_first, [&](unsigned j, double raw) { p[j] = raw; norm += raw; n = j + 1; });
    double inv = 1. / norm;
    for (unsigned j = 0; j < n; j++) p[j] *= inv;
    return norm / m.d.s_norm;
}
__host__ __device__ inline double hb_o2_probs(const MolView &m, double *p, const uint8_t *occ, unsigned o1_idx) {
    uint64_t key = mol_key_of_occ(m, occ);
    double norm = hbs_o2_norm(m, key, o1_idx), inv = 1. / norm;
    hbs_o2(m, key, o1_idx, [&](unsigned j, double raw) { p[j] = raw * inv; });
    return norm / m.s_tens[occ[o1_idx] % m.d.n_orb];
}
__host__ __device__ inline double hb_o2_probs_half(const MolView &m, double *p, const uint8_t *occ, unsigned o1_idx) {
    uint64_t key = mol_key_of_occ(m, occ);
    double norm = 0;
    hbs_o2_half(m, key, o1_idx, [&](unsigned j, double raw) { p[j] = raw; norm += raw; });
    double inv = 1. / norm;
    for (unsigned j = 0; j < o1_idx; j++) p[j] *= inv;
    return norm / m.s_tens[occ[o1_idx] % m.d.n_orb];
}
__host__ __device__ inline double hb_u1_probs(const MolView &m, double *p, unsigned o1_orb, const uint8_t *occ,
                                              int exclude_first) {
    uint64_t key = mol_key_of_occ(m, occ);
    double norm = 0;
    unsigned n = 0;
    hbs_u1(m, key, o1_orb, [&](unsigned j, double raw) { p[j] = raw; norm += raw; n = j + 1; });
    if (exclude_first) {
        norm -= p[0];
        p[0] = 0;
    }
    double inv = 1. / norm;
    for (unsigned j = 0; j < n; j++) p[j] *= inv;
    return norm / m.exch_norms[o1_orb % m.d.n_orb];
}
__host__ __device__ inline double hb_u2_probs(const MolView &m, double *p, unsigned o1_orb, unsigned o2_orb,
                                              unsigned u1_orb, unsigned *len) {
    double norm = 0;
    unsigned n = 0;
    hbs_u2(m, o1_orb, o2_orb, u1_orb, [&](unsigned j, double raw) { p[j] = raw; norm += raw; n = j + 1; });
    *len = mol_lookup(m, m.symm[o1_orb % m.d.n_orb] ^ m.symm[o2_orb % m.d.n_orb] ^ m.symm[u1_orb % m.d.n_orb], 0);
    if (norm != 0) {
        double inv = 1 / norm;
        for (unsigned j = 0; j < n; j++) p[j] *= inv;
    }
    return norm / m.exch_norms[o2_orb % m.d.n_orb];
}
__host__ __device__ inline double hb_u2_probs_half(const MolView &m, double *p, unsigned o1_orb, unsigned o2_orb,
                                                   unsigned u1_orb, uint64_t det, unsigned *len) {
    double norm = 0;
    unsigned n = 0;
    hbs_u2_half(m, o1_orb, o2_orb, u1_orb, det, [&](unsigned j, double raw) { p[j] = raw; norm += raw; n = j + 1; });
    *len = n;
    if (norm != 0) {
        double inv = 1 / norm;
        for (unsigned j = 0; j < n; j++) p[j] *= inv;
    }
    return norm / m.exch_norms[o2_orb % m.d.n_orb];
}

// ---- a9: total weights -----------------------------------------------------------------------------------------
// calc_unnorm_wt :414-439
__host__ __device__ inline double hb_unnorm_wt(const MolView &m, const uint8_t *orbs) {
    const unsigned M = m.d.n_orb;
    unsigned o1 = orbs[0] % M, o2 = orbs[1] % M, u1 = orbs[2] % M, u2 = orbs[3] % M;
    unsigned mn1 = o1 < u1 ? o1 : u1, mx1 = o1 > u1 ? o1 : u1, mn2 = o2 < u2 ? o2 : u2, mx2 = o2 > u2 ? o2 : u2;
    bool same = (orbs[0] / M) == (orbs[1] / M);
    // NOTE the reference evaluates I_J_TO_TRI_NODIAG with uint8_t operands promoted to int; when
    // mn == mx (o == u as spatial orbitals, opposite-spin partner) TRI_N(j-1)+i indexes as written.
    int o1u1 = FR_TRI_NODIAG((int)mn1, (int)mx1), o2u2 = FR_TRI_NODIAG((int)mn2, (int)mx2);
    double w;
    if (same) {
        int o1o2 = FR_TRI_NODIAG((int)o1, (int)o2);
        w = m.d_same[o1o2] * (m.exch_sqrt[o1u1] * m.exch_sqrt[o2u2]) / m.d.s_norm / m.exch_norms[o1] / m.exch_norms[o2];
    } else {
        w = (m.d_diff[o2 * M + o1]) * m.exch_sqrt[o1u1] * m.exch_sqrt[o2u2] / m.d.s_norm / m.exch_norms[o1] /
            m.exch_norms[o2];
    }
    return w;
}
// calc_norm_wt :442-598
__host__ __device__ inline double hb_norm_wt(const MolView &m, const uint8_t *orbs, const uint8_t *occ, uint64_t det) {
    const unsigned M = m.d.n_orb, ne = m.d.n_elec;
    int o1 = orbs[0] % M, o1_spin = orbs[0] / M, o2 = orbs[1] % M, o2_spin = orbs[1] / M;
    int u1 = orbs[2] % M, u2 = orbs[3] % M;
    int mn_o1u1 = o1 < u1 ? o1 : u1, mx_o1u1 = o1 > u1 ? o1 : u1, mn_o2u2 = o2 < u2 ? o2 : u2,
        mx_o2u2 = o2 > u2 ? o2 : u2;
    bool same = o1_spin == o2_spin;
    uint8_t os[FRIES_MAX_ELEC + 1];
    for (unsigned i = 0; i < ne; i++) os[i] = occ[i] % M;
    os[ne] = 255;
    double s_denom = 0;
    for (unsigned i = 0; i < ne; i++) s_denom += m.s_tens[os[i]];
    unsigned i;
    double d1 = 0;
    unsigned off = (1 - o1_spin) * ne / 2;
    for (i = off; i < ne / 2 + off; i++) d1 += m.d_diff[o1 * M + os[i]];
    off = o1_spin * ne / 2;
    for (i = off; (int)os[i] < o1; i++) d1 += m.d_same[FR_TRI_NODIAG((int)os[i], o1)];
    for (i++; i < ne / 2 + off; i++) d1 += m.d_same[FR_TRI_NODIAG(o1, (int)os[i])];
    double d2 = 0;
    off = (1 - o2_spin) * ne / 2;
    for (i = off; i < ne / 2 + off; i++) d2 += m.d_diff[o2 * M + os[i]];
    off = o2_spin * ne / 2;
    for (i = off; (int)os[i] < o2; i++) d2 += m.d_same[FR_TRI_NODIAG((int)os[i], o2)];
    for (i++; i < ne / 2 + off; i++) d2 += m.d_same[FR_TRI_NODIAG(o2, (int)os[i])];

    double e1_virt = 0;
    unsigned offset = o1_spin * M;
    for (int k = 0; k < o1; k++)
        if (!fr_read_bit(det, k + offset)) e1_virt += m.exch_sqrt[FR_TRI_NODIAG(k, o1)];
    for (int k = o1 + 1; k < (int)M; k++)
        if (!fr_read_bit(det, k + offset)) e1_virt += m.exch_sqrt[FR_TRI_NODIAG(o1, k)];
    double e2_virt = 0;
    offset = o2_spin * M;
    for (int k = 0; k < o2; k++)
        if (!fr_read_bit(det, k + offset)) e2_virt += m.exch_sqrt[FR_TRI_NODIAG(k, o2)];
    for (int k = o2 + 1; k < (int)M; k++)
        if (!fr_read_bit(det, k + offset)) e2_virt += m.exch_sqrt[FR_TRI_NODIAG(o2, k)];

    unsigned u1_irrep = m.symm[u1], u2_irrep = m.symm[u2];
    double e2_no1 = 0, e2_no2 = 0, e1_no1 = 0, e1_no2 = 0;
    unsigned n = mol_lookup(m, u2_irrep, 0);
    for (unsigned k = 0; k < n; k++) {
        int so = mol_lookup(m, u2_irrep, k + 1);
        if ((same && so != u1) || !same) {
            if (o2 == so) {
                e2_no1 += m.diag_sqrt[o2];
            } else {
                int mn = o2 < so ? o2 : so, mx = o2 > so ? o2 : so;
                e2_no1 += m.exch_sqrt[FR_TRI_NODIAG(mn, mx)];
            }
            if (o1 == so) {
                e1_no1 += m.diag_sqrt[o1];
            } else {
                int mn = o1 < so ? o1 : so, mx = o1 > so ? o1 : so;
                e1_no1 += m.exch_sqrt[FR_TRI_NODIAG(mn, mx)];
            }
        }
    }
    n = mol_lookup(m, u1_irrep, 0);
    for (unsigned k = 0; k < n; k++) {
        int so = mol_lookup(m, u1_irrep, k + 1);
        if ((same && so != u2) || !same) {
            if (o2 == so) {
                e2_no2 += m.diag_sqrt[o2];
            } else {
                int mn = o2 < so ? o2 : so, mx = o2 > so ? o2 : so;
                e2_no2 += m.exch_sqrt[FR_TRI_NODIAG(mn, mx)];
            }
            if (o1 == so) {
                e1_no2 += m.diag_sqrt[o1];
            } else {
                int mn = o1 < so ? o1 : so, mx = o1 > so ? o1 : so;
                e1_no2 += m.exch_sqrt[FR_TRI_NODIAG(mn, mx)];
            }
        }
    }
    int o1u1 = FR_TRI_NODIAG(mn_o1u1, mx_o1u1), o2u2 = FR_TRI_NODIAG(mn_o2u2, mx_o2u2);
    double w;
    if (same) {
        int mn_o1u2 = o1 < u2 ? o1 : u2, mx_o1u2 = o1 > u2 ? o1 : u2, mn_o2u1 = o2 < u1 ? o2 : u1,
            mx_o2u1 = o2 > u1 ? o2 : u1;
        int o1o2 = FR_TRI_NODIAG(o1, o2), o1u2 = FR_TRI_NODIAG(mn_o1u2, mx_o1u2), o2u1 = FR_TRI_NODIAG(mn_o2u1, mx_o2u1);
        w = m.d_same[o1o2] / s_denom *
            (m.s_tens[o1] / d1 / e1_virt *
                 (m.exch_sqrt[o1u1] * m.exch_sqrt[o2u2] / e2_no1 + m.exch_sqrt[o1u2] * m.exch_sqrt[o2u1] / e2_no2) +
             m.s_tens[o2] / d2 / e2_virt *
                 (m.exch_sqrt[o2u1] * m.exch_sqrt[o1u2] / e1_no1 + m.exch_sqrt[o2u2] * m.exch_sqrt[o1u1] / e1_no2));
    } else {
        w = (m.s_tens[o1] * m.d_diff[o1 * M + o2] / d1 / e1_virt / e2_no1 +
             m.s_tens[o2] * m.d_diff[o2 * M + o1] / d2 / e2_virt / e1_no2) *
            m.exch_sqrt[o1u1] * m.exch_sqrt[o2u2] / s_denom;
    }
    return w;
}

// occupied list + sentinel (find_bits math_utils.c:62-98); returns the electron count
__host__ __device__ __forceinline__ int mol_occ_list(uint64_t key, uint8_t *occ) {
    int n = fr_occ_list(key, occ);
    occ[n] = 255;
    return n;
}
