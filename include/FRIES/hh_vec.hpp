// HubHolVec<el_type> (FRIES/hh_vec.hpp:14-264): a DistVec whose bit strings carry, after the 2 n_sites electron bits,
// one ph_bits-wide phonon number per site.  The phonon field arithmetic is host bit manipulation (as in the reference);
// the store itself is the device-resident HubHolVec of the C-ABI (fries_vec_create_hh), created on first use.
#pragma once
#include "vec_utils.hpp"

template <class el_type>
class HubHolVec : public DistVec<el_type> {
    uint8_t n_sites_, ph_bits_;
    void create_() override {
        this->dev_.reset(new fries::DistVec(fries_default_context(), this->size_, n_sites_, ph_bits_, this->n_elec_, this->n_vecs_,
                                            this->rns_common_, this->rns_distinct_));
    }
    unsigned field_bit(uint8_t site) const { return 2u * n_sites_ + (unsigned)site * ph_bits_; }

  public:
    // hh_vec.hpp:27-29
    HubHolVec(size_t size, size_t add_size, uint8_t n_sites, uint8_t max_ph, unsigned int n_elec, int n_procs,
              std::function<double(const uint8_t *)> diag_fxn, uint8_t n_vecs, std::vector<uint32_t> rns_common,
              std::vector<uint32_t> rns_distinct)
        : DistVec<el_type>(size, add_size, (uint8_t)(n_sites * 2 + n_sites * max_ph), n_elec, n_procs, diag_fxn, n_vecs, rns_common,
                           rns_distinct),
          n_sites_(n_sites), ph_bits_(max_ph) {}
    // :185-197: the phonon number of every site
    void decode_phonons(uint8_t *det, uint8_t *numbers) {
        for (uint8_t site = 0; site < n_sites_; site++) {
            unsigned v = 0;
            for (unsigned b = 0; b < ph_bits_; b++) {
                const unsigned bit = field_bit(site) + b;
                v |= (unsigned)((det[bit / 8] >> (bit % 8)) & 1) << b;
            }
            numbers[site] = (uint8_t)v;
        }
    }
    // :207-233: the bit string with the phonon number of one site changed by +-1; 0 if that leaves 0 .. 2^ph_bits - 1
    int det_from_ph(uint8_t *orig, uint8_t *new_det, uint8_t site_idx, int change) {
        unsigned v = 0;
        for (unsigned b = 0; b < ph_bits_; b++) {
            const unsigned bit = field_bit(site_idx) + b;
            v |= (unsigned)((orig[bit / 8] >> (bit % 8)) & 1) << b;
        }
        if (change == 1 && v == (1u << ph_bits_) - 1) {
            std::cerr << "Warning: maximum phonon number reached\n";
            return 0;
        }
        if (change == -1 && v == 0) return 0;
        v += change;
        std::memcpy(new_det, orig, CEILING(this->n_bits_, 8));
        for (unsigned b = 0; b < ph_bits_; b++) {
            const unsigned bit = field_bit(site_idx) + b;
            if ((v >> b) & 1)
                new_det[bit / 8] |= (uint8_t)(1u << (bit % 8));
            else
                new_det[bit / 8] &= (uint8_t)~(1u << (bit % 8));
        }
        return 1;
    }
    // :43-45: electrons only
    uint8_t gen_orb_list(uint8_t *det, uint8_t *occ) override {
        uint8_t n = 0;
        for (unsigned bit = 0; bit < 2u * n_sites_; bit++)
            if ((det[bit / 8] >> (bit % 8)) & 1) occ[n++] = (uint8_t)bit;
        return n;
    }
};
