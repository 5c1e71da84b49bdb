// a16: the deterministic H.v (h_op_offdiag molecule.cpp:448-665) per lane and per connection, as __host__ __device__ code so
// that the CPU-only test tier can compile the same arithmetic for the host; the product uses it in the kernels of iter.cu.
#pragma once
#include "mol.cuh"

struct ParentCtx {
    uint64_t key;
    uint32_t occ_a, occ_b, virt_a, virt_b;  // spatial-orbital masks per spin
    uint32_t irr[FR_N_IRREPS];              // spatial orbitals of each irrep
};

__host__ __device__ __forceinline__ void parent_ctx(const MolView &m, uint64_t key, ParentCtx &p) {
    const unsigned M = m.d.n_orb;
    uint32_t all = (uint32_t)((1ull << M) - 1);
    p.key = key;
    p.occ_a = (uint32_t)(key & all);
    p.occ_b = (uint32_t)((key >> M) & all);
    p.virt_a = ~p.occ_a & all;
    p.virt_b = ~p.occ_b & all;
    for (unsigned r = 0; r < FR_N_IRREPS; r++) {
        uint32_t mk = 0;
        unsigned n = mol_lookup(m, r, 0);
        for (unsigned s = 0; s < n; s++) mk |= 1u << mol_lookup(m, r, s + 1);
        p.irr[r] = mk;
    }
}

// n-th set bit (0-based) of a 32-bit mask
__host__ __device__ __forceinline__ unsigned nth_bit(uint32_t mask, unsigned n) { return fr_nth_bit32(mask, n); }

// Visit this lane's share of the off-diagonal connections of the parent.  F(is_double, o0, o1, v0, v1).
template <class F>
__host__ __device__ __forceinline__ unsigned lane_excitations(const MolView &m, const ParentCtx &p, unsigned lane, F &&f) {
    const unsigned M = m.d.n_orb, h = m.d.n_elec / 2, nv = M - h;
    unsigned cnt = 0;
    // singles: (electron, spin) pairs over lanes
    for (unsigned e = lane; e < 2 * h; e += 32) {
        unsigned spin = e / h;
        unsigned o = nth_bit(spin ? p.occ_b : p.occ_a, e % h);
        uint32_t vm = (spin ? p.virt_b : p.virt_a) & p.irr[m.symm[o]];
        while (vm) {
            unsigned a = (unsigned)fr_ctz(vm);
            vm &= vm - 1;
            f(false, o + spin * M, a + spin * M, 0u, 0u);
            cnt++;
        }
    }
    // opposite-spin doubles: (i alpha, j beta, k alpha virtual) triples over lanes
    unsigned n_ab = h * h * nv;
    for (unsigned t = lane; t < n_ab; t += 32) {
        unsigned k_i = t % nv, ij = t / nv, i = ij / h, j = ij % h;
        unsigned io = nth_bit(p.occ_a, i), jo = nth_bit(p.occ_b, j), k = nth_bit(p.virt_a, k_i);
        uint32_t lm = p.virt_b & p.irr[m.symm[io] ^ m.symm[jo] ^ m.symm[k]];
        while (lm) {
            unsigned l = (unsigned)fr_ctz(lm);
            lm &= lm - 1;
            f(true, io, jo + M, k, l + M);
            cnt++;
        }
    }
    // same-spin doubles
    unsigned n_pair = h * (h - 1) / 2;
    for (unsigned spin = 0; spin < 2; spin++) {
        uint32_t om = spin ? p.occ_b : p.occ_a, vmask = spin ? p.virt_b : p.virt_a;
        for (unsigned t = lane; t < n_pair * nv; t += 32) {
            unsigned k_i = t % nv, pr = t / nv;
            // pair index -> (i < j)
            unsigned j = 1;
            while (j * (j + 1) / 2 <= pr) j++;
            unsigned i = pr - j * (j - 1) / 2;
            unsigned io = nth_bit(om, i), jo = nth_bit(om, j), k = nth_bit(vmask, k_i);
            uint32_t lm = vmask & p.irr[m.symm[io] ^ m.symm[jo] ^ m.symm[k]] & ~((2u << k) - 1u);
            while (lm) {
                unsigned l = (unsigned)fr_ctz(lm);
                lm &= lm - 1;
                f(true, io + spin * M, jo + spin * M, k + spin * M, l + spin * M);
                cnt++;
            }
        }
    }
    return cnt;
}

// one connection: new determinant, sign x matrix element x parent value x h_fac (molecule.cpp:590-601,644-655)
__host__ __device__ __forceinline__ double hv_connection(const MolView &m, uint64_t key, const uint8_t *occ, bool dbl, unsigned o0,
                                                         unsigned o1, unsigned v0, unsigned v1, double val, double h_fac,
                                                         uint64_t &nk) {
    nk = key;
    double el;
    if (dbl) {
        uint8_t ob[4] = {(uint8_t)o0, (uint8_t)o1, (uint8_t)v0, (uint8_t)v1};
        el = mol_doub_el(m, ob);
        el *= fr_doub_det_parity(nk, o0, o1, v0, v1);
    } else {
        el = mol_sing_el(m, o0, o1, occ);
        el *= fr_sing_det_parity(nk, o0, o1);
    }
    el *= val * h_fac;  // matr_el *= curr_el * h_fac (molecule.cpp:601,655)
    return el;
}
