// Providers of the five HB-PP stages (apply_HBPP_sys heat_bathPP.cpp:686-992; shared by the pivotal pipeline
// apply_HBPP_piv :1014-1419): the per-sample set-up that precedes each compression and the on-the-fly sub-weight rows.
// __host__ __device__ so that the CPU-only test tier can compile the same arithmetic for the host; the product only ever
// instantiates them in kernels (hbpp.cu).
#pragma once
#include "mol.cuh"

__host__ __device__ __forceinline__ unsigned hb_popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (unsigned)__popc(x);
#else
    return (unsigned)__builtin_popcount(x);
#endif
}

struct HbStageIO {
    const uint64_t *keys;            // parent determinants (storage order)
    const double *vals;              // stage 0 input: vector values
    const unsigned long long *n_in;  // number of inputs of this stage (device)
    // outputs of the previous stage (inputs of this one)
    const double *pv;
    const uint32_t *pw, *ps;
    const uint32_t *pdet, *ppath;
    // per-item path state written by this stage
    uint32_t *det, *path;
    double p_doub;
    int new_hb;
    unsigned long long in_cap;       // inputs beyond this index were dropped by the previous stage
};

__host__ __device__ __forceinline__ uint32_t pk(unsigned p0, unsigned p1, unsigned p2, unsigned p3) {
    return (p0 & 0xff) | ((p1 & 0xff) << 8) | ((p2 & 0xff) << 16) | ((p3 & 0xff) << 24);
}

// Provider of stage S of the hierarchy.  prep() restates the per-sample set-up loop that precedes each
// comp_sub call in apply_HBPP_sys; visit() streams the sub-weight row the reference stores in subwts
// (mol.cuh hbs_* generators: masks + popcount, no row array, no occupied list).
template <int S>
struct HbProvider {
    MolView m;  // tables in shared memory
    HbStageIO io;

    __host__ __device__ size_t count() const {
        unsigned long long n = *io.n_in;
        return n < io.in_cap ? (size_t)n : (size_t)io.in_cap;
    }

    // singles bookkeeping (count_symm_virt + count_sing_allowed / count_sing_virt, near_uniform.cpp:14-28,316-347) on bit
    // masks: no occupied list and no per-irrep counter array (local memory) in the hot loop
    __host__ __device__ unsigned sing_allowed(uint64_t key) const {
        OccMask o = mol_occ_mask(m, key);
        return mol_count_sing_allowed_bits(m, o.a, o.b);
    }
    // choice in: index among the allowed electrons (alpha block first); out: electron index; returns its virtual count
    __host__ __device__ unsigned sing_virt(uint64_t key, unsigned &choice) const {
        const unsigned M = m.d.n_orb, h = m.d.n_elec / 2;
        OccMask o = mol_occ_mask(m, key);
        uint32_t al_a, al_b;
        mol_sing_allowed_masks(m, o.a, o.b, al_a, al_b);
        const unsigned na = (unsigned)hb_popc32(al_a), nb = (unsigned)hb_popc32(al_b);
        const uint32_t all = (uint32_t)((1ull << M) - 1);
        if (choice < na) {
            unsigned orb = fr_nth_bit32(al_a, choice);
            choice = (unsigned)hb_popc32(o.a & ((1u << orb) - 1u));
            return (unsigned)hb_popc32(~o.a & all & m.irr_mask[m.symm[orb]]);
        }
        if (choice < na + nb) {
            unsigned orb = fr_nth_bit32(al_b, choice - na);
            choice = h + (unsigned)hb_popc32(o.b & ((1u << orb) - 1u));
            return (unsigned)hb_popc32(~o.b & all & m.irr_mask[m.symm[orb]]);
        }
        return 0;  // count_sing_virt leaves occ_choice untouched and returns 0 when the index is out of range
    }

    // wmax: an upper bound of the sub-weights visit() will stream for this input (exactly the largest one where the
    // row is traversed here anyway); the engine skips the row of an input whose v * wmax is below the threshold bracket
    // The inputs of prep(): three levels of dependent global loads (output list -> path state of the parent item -> parent
    // determinant).  The second-generation engine fetches them level by level for all the inputs of a thread before it
    // calls prep_core (compress2.cuh), so that a thread pays three load latencies per tile instead of three per input.
    struct Pre {
        double v;
        uint32_t widx, sub, d, pp;
        uint64_t key;
    };
    __host__ __device__ __forceinline__ void fetch1(size_t i, Pre &p) const {
        if (S == 0) {
            p.v = io.vals[i];
            p.widx = p.sub = p.d = p.pp = 0;
            p.key = 0;
        } else {
            p.widx = io.pw[i];
            p.sub = io.ps[i];
            p.v = io.pv[i];
        }
    }
    __host__ __device__ __forceinline__ void fetch2(Pre &p) const {
        if (S != 0) {
            p.d = io.pdet[p.widx];
            p.pp = io.ppath[p.widx];
        }
    }
    __host__ __device__ __forceinline__ void fetch3(Pre &p) const {
        if (S != 0) p.key = io.keys[p.d];
    }
    __host__ __device__ void prep(size_t i, double &v, uint32_t &nd, uint32_t &ns, double &rinv, double &wmax) const {
        Pre p;
        fetch1(i, p);
        fetch2(p);
        fetch3(p);
        prep_core(i, p, v, nd, ns, rinv, wmax);
    }
    __host__ __device__ void prep_core(size_t i, const Pre &pre, double &v, uint32_t &nd, uint32_t &ns, double &rinv,
                                       double &wmax) const {
        const unsigned ne = m.d.n_elec, M = m.d.n_orb;
        rinv = 1.0;
        wmax = 1.0;
        if (S == 0) {  // singles vs doubles :713-727
            double w = fabs(pre.v);
            wmax = fmax(io.p_doub, 1 - io.p_doub);
            v = w;
            nd = w > 0 ? 0u : 1u;
            ns = 2;
            io.det[i] = (uint32_t)i;
            io.path[i] = 0;
            return;
        }
        const uint32_t sub = pre.sub;
        const uint32_t d = pre.d, pp = pre.pp;
        v = pre.v;
        io.det[i] = d;
        const uint64_t key = pre.key;
        unsigned p0 = pp & 0xff, p1 = (pp >> 8) & 0xff, p2 = (pp >> 16) & 0xff, p3 = pp >> 24;
        if (S == 1) {  // first occupied orbital :738-763
            p0 = sub;
            ns = ne - (io.new_hb ? 1 : 0);
            if (p0 == 0) {
                nd = 0;
                double norm = 0, mx = 0;
                hbs_o1(m, key, io.new_hb, [&](unsigned, double raw) {
                    norm += raw;
                    mx = fmax(mx, raw);
                });
                rinv = 1. / norm;
                wmax = mx * rinv;
                if (io.new_hb) v *= norm / m.d.s_norm;
            } else {
                unsigned n_occ = sing_allowed(key);
                if (n_occ == 0) {
                    nd = 1;
                    v = 0;
                } else {
                    nd = n_occ;
                }
            }
            io.path[i] = pk(p0, 0, 0, 0);
        } else if (S == 2) {  // 2nd occupied (double) / virtual count (single) :772-809
            p1 = sub;
            ns = ne - (io.new_hb ? 1 : 0);
            if (p1 >= ne) {
                v = 0;
                nd = 1;
            } else if (p0 == 0) {
                nd = 0;
                OccMask o = mol_occ_mask(m, key);
                if (io.new_hb) {
                    p1++;
                    ns = p1;
                    double norm = 0, mx = 0;
                    hbs_o2_half(m, key, p1, [&](unsigned, double raw) {
                        norm += raw;
                        mx = fmax(mx, raw);
                    });
                    rinv = 1. / norm;
                    wmax = mx * rinv;
                    v *= norm / m.s_tens[mol_elec_orb(m, o, p1) % M];
                } else {
                    rinv = 1. / hbs_o2_norm(m, key, p1);
                }
            } else {
                unsigned n_virt = sing_virt(key, p1);
                if (n_virt == 0) {
                    nd = 1;
                    v = 0;
                } else {
                    nd = n_virt;
                    p3 = n_virt;
                }
            }
            io.path[i] = pk(p0, p1, 0, p3);
        } else if (S == 3) {  // 1st virtual (double) :818-857
            p2 = sub;
            ns = M - ne / 2;
            if (p0 == 0) {
                if (p2 >= ne) {
                    v = 0;
                    nd = 1;
                } else {
                    nd = 0;
                    OccMask o = mol_occ_mask(m, key);
                    unsigned o1_orb = mol_elec_orb(m, o, p1);
                    bool excl = io.new_hb && (p1 / (ne / 2) == mol_elec_orb(m, o, p2) / M);
                    double norm = 0, first = 0, mx = 0;
                    hbs_u1(m, key, o1_orb, [&](unsigned j, double raw) {
                        if (j == 0) first = raw;
                        norm += raw;
                        mx = fmax(mx, raw);
                    });
                    if (excl) norm -= first;
                    rinv = 1. / norm;
                    wmax = mx * rinv;
                    if (io.new_hb) v *= norm / m.exch_norms[o1_orb % M];
                }
                p3 = 0;
            } else {
                nd = 1;
            }
            io.path[i] = pk(p0, p1, p2, p3);
        } else {  // S == 4: 2nd virtual (double) :866-908
            ns = m.d.max_n_symm;
            if (p0 == 0) {
                OccMask o = mol_occ_mask(m, key);
                unsigned spin = p1 / (ne / 2);
                uint32_t vm = ~(spin ? o.b : o.a) & (uint32_t)((1ull << M) - 1);
                if (sub >= (unsigned)hb_popc32(vm)) {  // find_nth_virt (fci_utils.c:138-148) would leave the orbital range
                    v = 0;
                    nd = 1;
                } else {
                    unsigned u1 = fr_nth_bit32(vm, sub) + M * spin;
                    nd = 0;
                    p3 = u1;
                    unsigned o1_orb = mol_elec_orb(m, o, p1), o2_orb = mol_elec_orb(m, o, p2);
                    double norm = 0, mx = 0;
                    unsigned len = 0;
                    if (io.new_hb) {
                        hbs_u2_half(m, o1_orb, o2_orb, u1, key, [&](unsigned j, double raw) {
                            norm += raw;
                            mx = fmax(mx, raw);
                            len = j + 1;
                        });
                    } else {
                        hbs_u2(m, o1_orb, o2_orb, u1, [&](unsigned j, double raw) {
                            norm += raw;
                            mx = fmax(mx, raw);
                            len = j + 1;
                        });
                    }
                    ns = len;
                    rinv = norm != 0 ? 1 / norm : 1.0;
                    wmax = mx * rinv;
                    double tot = norm / m.exch_norms[o2_orb % M];
                    if (io.new_hb || tot == 0) v *= tot;
                }
            } else {
                nd = 1;
            }
            io.path[i] = pk(p0, p1, p2, p3);
        }
    }

    template <class F>
    __host__ __device__ void visit(size_t i, double rinv, F &&f) const {
        if (S == 0) {
            if (fr_emit(f, 0u, io.p_doub)) fr_emit(f, 1u, 1 - io.p_doub);
            return;
        }
        const unsigned ne = m.d.n_elec, M = m.d.n_orb;
        const uint32_t pp = io.path[i];
        const uint64_t key = io.keys[io.det[i]];
        unsigned p1 = (pp >> 8) & 0xff, p2 = (pp >> 16) & 0xff, p3 = pp >> 24;
        if (S == 1) {
            hbs_o1(m, key, io.new_hb, [&](unsigned j, double raw) { return fr_emit(f, j, raw * rinv); });
        } else if (S == 2) {
            if (io.new_hb)
                hbs_o2_half(m, key, p1, [&](unsigned j, double raw) { return fr_emit(f, j, raw * rinv); });
            else
                hbs_o2(m, key, p1, [&](unsigned j, double raw) { return fr_emit(f, j, raw * rinv); });
        } else if (S == 3) {
            OccMask o = mol_occ_mask(m, key);
            bool excl = io.new_hb && (p1 / (ne / 2) == mol_elec_orb(m, o, p2) / M);
            hbs_u1(m, key, mol_elec_orb(m, o, p1),
                   [&](unsigned j, double raw) { return fr_emit(f, j, (excl && j == 0) ? 0.0 : raw * rinv); });
        } else {
            OccMask o = mol_occ_mask(m, key);
            unsigned o1_orb = mol_elec_orb(m, o, p1), o2_orb = mol_elec_orb(m, o, p2);
            if (io.new_hb)
                hbs_u2_half(m, o1_orb, o2_orb, p3, key, [&](unsigned j, double raw) { return fr_emit(f, j, raw * rinv); });
            else
                hbs_u2(m, o1_orb, o2_orb, p3, [&](unsigned j, double raw) { return fr_emit(f, j, raw * rinv); });
        }
    }
};

// finalize of one sample (apply_HBPP_sys :917-991; the last collapse of apply_HBPP_piv :1250-1415): orbitals of the
// excitation from the 4-byte path `pp` and the last stage's choice `sub`, total sampling weight, matrix element with the
// excitation's sign.  Returns value x element / weight, or 0 when the sample fails or is at or below `cutoff`.
__host__ __device__ __forceinline__ double hbpp_finalize_sample(const MolView &m, uint64_t key, uint32_t pp, uint32_t sub,
                                                                double val, double p_doub, int new_hb, double cutoff,
                                                                uint8_t (&orbs)[4], bool &is_doub) {
    const unsigned M = m.d.n_orb;
    double el = 0;
    const OccMask om = mol_occ_mask(m, key);  // occupied orbitals by spin: no occupied list in the common paths
    unsigned p0 = pp & 0xff, p1 = (pp >> 8) & 0xff, p2 = (pp >> 16) & 0xff, p3 = pp >> 24;
    orbs[0] = orbs[1] = orbs[2] = orbs[3] = 0;
    is_doub = p0 == 0;
    if (is_doub) {
        unsigned o1 = mol_elec_orb(m, om, p1), o2 = mol_elec_orb(m, om, p2), u1 = p3;
        unsigned u2_symm = m.symm[o1 % M] ^ m.symm[o2 % M] ^ m.symm[u1 % M];
        unsigned u2 = mol_lookup(m, u2_symm, sub + 1) + M * (o2 / M);
        if (!fr_read_bit(key, u2) && u1 != u2) {
            if (u1 > u2) {
                unsigned t = u1;
                u1 = u2;
                u2 = t;
            }
            if (o1 > o2) {
                unsigned t = o1;
                o1 = o2;
                o2 = t;
            }
            orbs[0] = (uint8_t)o1;
            orbs[1] = (uint8_t)o2;
            orbs[2] = (uint8_t)u1;
            orbs[3] = (uint8_t)u2;
            double tot;
            if (new_hb) {
                tot = hb_unnorm_wt(m, orbs);
            } else {
                uint8_t occ[FRIES_MAX_ELEC + 1];
                mol_occ_list(key, occ);
                tot = hb_norm_wt(m, orbs, occ, key);
            }
            el = mol_doub_el(m, orbs) * val / tot / p_doub;
            if (fabs(el) > cutoff)
                el *= fr_doub_parity(key, o1, o2, u1, u2);
            else
                el = 0;
        }
    } else {
        unsigned o1 = mol_elec_orb(m, om, p1);
        unsigned u1 = mol_virt_from_idx(m, key, m.symm[o1 % M], M * (o1 / M), p2);
        if (u1 != 255) {
            orbs[0] = (uint8_t)o1;
            orbs[1] = (uint8_t)u1;
            unsigned n_occ = mol_count_sing_allowed_bits(m, om.a, om.b);
            el = mol_sing_el_bits(m, o1, u1, om.a, om.b);
            el *= val / (1 - p_doub) * n_occ * p3;
            if (fabs(el) > cutoff)
                el *= fr_sing_parity(key, o1, u1);
            else
                el = 0;
        }
    }
    return el;
}

// spawn loop body frisys_mol.cpp:436-461 for one successful sample: the new determinant (with the initiator flag of its
// parent in bit 63: |parent value| >= init_thresh) and the element -eps x value x sign(parent value) to add to it
__host__ __device__ __forceinline__ uint64_t hbpp_spawn_element(uint64_t key, const uint8_t (&orbs)[4], bool is_doub, double el,
                                                               double parent_val, double eps, double init_thresh, double &add) {
    add = -eps * el;
    if (parent_val < 0) add *= -1;
    uint64_t nk = key;
    if (is_doub)
        nk = (nk & ~((1ull << orbs[0]) | (1ull << orbs[1]))) | (1ull << orbs[2]) | (1ull << orbs[3]);
    else
        nk = (nk & ~(1ull << orbs[0])) | (1ull << orbs[1]);
    if (fabs(parent_val) >= init_thresh) nk |= FRIES_INI_FLAG;
    return nk;
}

// ---- apply_HBPP_piv (heat_bathPP.cpp:1014-1419): one group of the "long" vector per input ----
// set-up of input i: effective value, row shape and normalisation; returns the group's length (a uniform row of n_div
// equal pieces, or the explicit row of n_sub weights; an input without continuation is one entry of value 0, which
// the pivotal compression zeroes like any other element of weight 0)
template <class P>
__host__ __device__ __forceinline__ uint32_t hbpp_piv_group_prep(const P &prov, size_t i, double &veff, uint32_t &ndiv,
                                                                 uint8_t &nsub, double &rinv) {
    double v, ri, wmax;
    uint32_t nd, ns;
    prov.prep(i, v, nd, ns, ri, wmax);
    if (ns > FRIES_MAX_SUB) ns = FRIES_MAX_SUB;
    veff = v;
    ndiv = nd;
    nsub = (uint8_t)ns;
    rinv = ri;
    return nd > 0 ? nd : ns;
}
// the group's entries: value / n_div each, or value x weight_j (:1046-1246)
template <class P>
__host__ __device__ __forceinline__ void hbpp_piv_group_fill(const P &prov, size_t i, double v, uint32_t nd, uint32_t ns,
                                                             double rinv, double *dst) {
    if (nd > 0) {
        const double piece = v / nd;
        for (uint32_t j = 0; j < nd; j++) dst[j] = piece;
    } else {
        for (uint32_t j = 0; j < ns; j++) dst[j] = 0.0;
        prov.visit(i, rinv, [&](unsigned j, double w) {
            if (j < ns) dst[j] = v * w;
        });
    }
}
