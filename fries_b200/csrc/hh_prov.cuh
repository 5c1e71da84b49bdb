// a19: Hubbard-Holstein arithmetic on u64 keys (hub_diag, neighbour lists, phonon numbers, the overlap with H x reference):
// __host__ __device__ so that the CPU-only test tier can compile the same code for the host; the product only uses it in
// the kernels of hh.cu.  Keys: bits [0, n) spin-up sites, [n, 2n) spin-down sites, then n phonon fields of ph_bits bits.
#pragma once
#include "common.cuh"

struct HhDims {
    unsigned n_sites, n_elec, ph_bits;
};

// hub_diag hub_holstein.cpp:101-136: number of doubly occupied sites
__host__ __device__ __forceinline__ unsigned hh_hub_diag(uint64_t key, unsigned n) {
    uint64_t m = (1ull << n) - 1;
    return fr_popc((key & m) & ((key >> n) & m));
}
// HubHolVec::find_neighbors_1D hh_vec.hpp:139-175 as bit masks over the 2n electron bits: `right` = occupied
// orbitals whose neighbour at +1 is empty (first list, hop to orb + 1), `left` = those whose neighbour at -1 is
// empty (second list, hop to orb - 1); open boundary conditions, no hop between the spin blocks
__host__ __device__ __forceinline__ void hh_neighbors(uint64_t key, unsigned n, uint64_t &plus, uint64_t &minus) {
    uint64_t E = (1ull << (2 * n)) - 1, occ = key & E;
    plus = occ & ~(occ >> 1) & ~(1ull << (n - 1)) & ~(1ull << (2 * n - 1));
    minus = occ & ((~occ << 1) & E) & ~(1ull << n);
}
__host__ __device__ __forceinline__ unsigned hh_nth_bit(uint64_t mask, unsigned k) {
    for (unsigned i = 0; i < k; i++) mask &= mask - 1;
    return fr_ctz(mask);
}
// HubHolVec::decode_phonons hh_vec.hpp:185-197
__host__ __device__ __forceinline__ unsigned hh_phonon(uint64_t key, const HhDims &d, unsigned site) {
    return (unsigned)((key >> (2 * d.n_sites + site * d.ph_bits)) & ((1u << d.ph_bits) - 1));
}
__host__ __device__ __forceinline__ unsigned hh_total_ph(uint64_t key, const HhDims &d) {
    unsigned s = 0;
    for (unsigned i = 0; i < d.n_sites; i++) s += hh_phonon(key, d, i);
    return s;
}
// calc_ref_ovlp hub_holstein.hpp:93-182 for ONE basis state: its contribution to the sum (the reference walks the
// electron bytes; the byte-wise neighbour tests, including the open-boundary byte rule, are kept as written)
__host__ __device__ inline double hh_ref_ovlp_term(uint64_t curr, double val, uint64_t ref, const HhDims &d, double g_over_t) {
    const unsigned n = d.n_sites;
    uint64_t E = (1ull << (2 * n)) - 1;
    if ((curr & E) == (ref & E)) {
        unsigned sites_found = 0, site_elecs = 0;
        for (unsigned site = 0; site < n && sites_found < 2; site++) {
            unsigned ph = hh_phonon(curr, d, site);
            unsigned n_occ = fr_read_bit(ref, site) + fr_read_bit(ref, site + n);
            if (ph > 1 || (ph == 1 && n_occ == 0)) {
                site_elecs = 0;
                break;
            } else if (ph == 1) {
                site_elecs = n_occ;
                sites_found++;
            }
        }
        if (sites_found == 2) site_elecs = 0;
        return -(val * g_over_t * site_elecs);
    }
    if (hh_total_ph(curr, d) != 0) return 0.0;
    unsigned n_hop = 0, n_common = 0;
    const unsigned n_bytes = (2 * n + 7) / 8;
    for (unsigned b = 0; b < n_bytes && n_hop <= 1; b++) {
        uint8_t c = (uint8_t)(curr >> (8 * b)), r = (uint8_t)(ref >> (8 * b));
        uint8_t c_prev = b ? (uint8_t)(curr >> (8 * (b - 1))) : 0, r_prev = b ? (uint8_t)(ref >> (8 * (b - 1))) : 0;
        uint8_t c_next = (uint8_t)(curr >> (8 * (b + 1))), r_next = (uint8_t)(ref >> (8 * (b + 1)));
        uint8_t not_occ = c & ~r;
        uint8_t ref_left = c & (r >> 1);
        uint8_t not_occ_left = (uint8_t)(~c) >> 1;
        uint8_t ref_right = c & (uint8_t)(r << 1);
        uint8_t not_occ_right = (uint8_t)((uint8_t)(~c) << 1);
        if (b > 0) {
            ref_right |= c & ((r_prev >> 7) & 1);
            not_occ_right |= ((uint8_t)(~c_prev) >> 7) & 1;
        }
        if (b < n_bytes - 1) {
            ref_left |= c & (uint8_t)(r_next << 7);
            not_occ_left |= (uint8_t)((uint8_t)(~c_next) << 7);
        }
        if (b == (n + 7) / 8) ref_left &= ~(1 << ((n - 1) % 8));
        uint8_t mask = not_occ & ((ref_left & not_occ_left) | (ref_right & not_occ_right));
        if (b == n_bytes - 1 && (2 * n) % 8 != 0) mask &= (1 << ((2 * n) % 8)) - 1;
        n_hop += fr_popc(mask);
        if (n_hop > 1) break;
        n_common += fr_popc((uint64_t)(r & c));
    }
    return (n_hop == 1 && n_common == d.n_elec - 1) ? val : 0.0;
}
