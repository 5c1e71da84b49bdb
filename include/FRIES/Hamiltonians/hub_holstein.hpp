// FRIES/Hamiltonians/hub_holstein.hpp: the functions a caller of the vector API needs (the in-scope driver frisys_hh uses
// the device versions through the C-ABI: fries_hh_*, include/fries_b200.h).
#pragma once
#include "../hh_vec.hpp"

// hub_holstein.hpp:73 / hub_holstein.cpp:139-171 -- Neel state of a 1-D lattice: up spins on the even sites, down spins on
// the odd sites, no phonons; bit string of (2 + ph_bits) n_sites bits
inline void gen_neel_det_1D(unsigned int n_sites, unsigned int n_elec, uint8_t ph_bits, uint8_t *det) {
    const unsigned n_bytes = CEILING((2 + ph_bits) * n_sites, 8);
    std::memset(det, 0, n_bytes);
    const uint64_t k = fries::gen_neel_det_1D(n_sites, n_elec);
    for (unsigned b = 0; b < 2 * n_sites; b++)
        if ((k >> b) & 1) det[b / 8] |= (uint8_t)(1u << (b % 8));
}
// hub_holstein.hpp:61 / hub_holstein.cpp:101-136 -- number of doubly occupied sites
inline unsigned int hub_diag(uint8_t *det, unsigned int n_sites) {
    unsigned n = 0;
    for (unsigned s = 0; s < n_sites; s++) {
        const unsigned up = s, dn = n_sites + s;
        n += ((det[up / 8] >> (up % 8)) & 1) & ((det[dn / 8] >> (dn % 8)) & 1);
    }
    return n;
}
