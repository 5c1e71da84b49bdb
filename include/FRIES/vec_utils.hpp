// DistVec<el_type> with the reference's constructor and member signatures (FRIES/vec_utils.hpp:121-953) over the
// device-resident store (fries_vec, include/fries_b200.h).  The store holds doubles; el_type = int values are exact in
// them.  Pointers the reference hands out into its own arrays (operator[], values(), indices(), occ_orbs()) point into a
// host snapshot that is refreshed after every mutating call.
#pragma once
#include "fries_global.hpp"

template <class el_type>
class DistVec {
  protected:
    std::unique_ptr<fries::DistVec> dev_;
    std::vector<el_type> snap_;  // [n_vecs][curr_size], refreshed lazily
    bool snap_ok_ = false;
    uint8_t n_bits_;
    unsigned n_elec_;
    uint8_t n_vecs_;
    size_t size_, add_size_;
    std::vector<uint32_t> rns_common_, rns_distinct_;
    std::function<double(const uint8_t *)> diag_fxn_;
    virtual void create_() {
        dev_.reset(new fries::DistVec(fries_default_context(), size_, n_bits_, n_elec_, n_vecs_, rns_common_, rns_distinct_));
        if (diag_fxn_) dev_->set_diag_calc(diag_fxn_);
    }
    fries::DistVec &dev() {
        if (!dev_) create_();
        return *dev_;
    }
    void refresh_() {
        if (snap_ok_) return;
        const double *v = dev().values();
        const size_t n = dev().curr_size();
        snap_.resize((size_t)n_vecs_ * (n ? n : 1));
        for (size_t i = 0; i < (size_t)n_vecs_ * n; i++) snap_[i] = (el_type)v[i];
        snap_ok_ = true;
    }

  public:
    uint64_t nonini_occ_add = 0;
    // vec_utils.hpp:184-198
    DistVec(size_t size, size_t add_size, uint8_t n_bits, unsigned int n_elec, int /*n_procs*/,
            std::function<double(const uint8_t *)> diag_fxn, uint8_t n_vecs, std::vector<uint32_t> rns_common,
            std::vector<uint32_t> rns_distinct)
        : n_bits_(n_bits), n_elec_(n_elec), n_vecs_(n_vecs), size_(size), add_size_(add_size), rns_common_(rns_common),
          rns_distinct_(rns_distinct), diag_fxn_(diag_fxn) {}
    DistVec(size_t size, size_t add_size, uint8_t n_bits, unsigned int n_elec, int n_procs, std::vector<uint32_t> rns_common,
            std::vector<uint32_t> rns_distinct)
        : DistVec(size, add_size, n_bits, n_elec, n_procs, nullptr, 1, rns_common, rns_distinct) {}
    virtual ~DistVec() = default;
    uint8_t n_bits() { return n_bits_; }                                                       // :200
    size_t max_size() { return size_; }
    size_t curr_size() { return dev().curr_size(); }
    int n_nonz() { return (int)dev().n_nonz(); }
    uint8_t num_vecs() const { return n_vecs_; }
    size_t adder_size() { return add_size_; }
    void set_curr_vec_idx(uint8_t new_idx) { dev().set_curr_vec_idx(new_idx); }               // :585-599
    // add :418-431, perform_add :433-440: buffered on the host, merged on the device (merge_insert / merge_accum kernels)
    bool add(uint8_t *idx, el_type val, int ini_flag) {
        snap_ok_ = false;
        return dev().add(idx, (double)val, ini_flag);
    }
    void perform_add(uint8_t origin_idx) {
        dev().perform_add(origin_idx);
        snap_ok_ = false;
    }
    void del_at_pos(size_t pos) {                                                              // :458-476
        std::vector<bool> flags(dev().curr_size(), false);
        flags[pos] = true;
        dev().del_at_pos(flags);
        snap_ok_ = false;
    }
    void cleanup() {
        dev().cleanup();
        snap_ok_ = false;
    }
    el_type *operator[](size_t pos) {                                                          // :485-487
        refresh_();
        return &snap_[(size_t)dev().curr_vec_idx() * dev().curr_size() + pos];
    }
    el_type *operator()(uint8_t row, size_t pos) {
        refresh_();
        return &snap_[(size_t)row * dev().curr_size() + pos];
    }
    el_type *values() {
        refresh_();
        return snap_.data();
    }
    Matrix<uint8_t> &indices() { return dev().indices(); }                                     // :497
    Matrix<uint8_t> &occ_orbs() { return dev().occ_orbs(); }
    uint8_t *orbs_at_pos(size_t pos) { return dev().orbs_at_pos(pos); }
    double matr_el_at_pos(size_t pos) { return dev().matr_el_at_pos(pos); }                   // :672-677
    double local_norm() { return dev().local_norm(); }                                         // :683-689
    double two_norm() { return dev().two_norm(); }                                             // :695-701
    double dense_norm() { return dev().dense_norm(); }
    void zero_vec() {
        dev().zero_vec();
        snap_ok_ = false;
    }
    void save(const std::string &path) { dev().save(path); }                                   // :713-750
    void load(const std::string &path) {
        dev().load(path);
        snap_ok_ = false;
    }
    virtual uint8_t gen_orb_list(uint8_t *det, uint8_t *occ) { return dev().gen_orb_list(det, occ); }
    uintmax_t idx_to_hash(uint8_t *idx, uint8_t *orbs) { return dev().idx_to_hash(idx, orbs); }
    virtual int idx_to_proc(uint8_t *idx) { return dev().idx_to_proc(idx); }
    uint64_t tot_sgn_coh() { return dev().tot_sgn_coh(); }
    // rows [start, end) compressed on the device (the free functions below)
    void compress_rows_(size_t start, size_t end, unsigned compress_size, int method, std::mt19937 &rn_gen) {
        dev().compress_rows((unsigned)start, (unsigned)end, compress_size, method, rn_gen);
        snap_ok_ = false;
    }
};

// FRIES/vec_utils.cpp:10-127 (declared vec_utils.hpp:1036-1071): compress rows [start_idx, end_idx) of a DistVec to
// compress_size elements each -- pivotal, systematic, multinomial (alias method) -- and delete what is zero everywhere.
// The scratch arguments of the reference are unused (the device has its own).
inline void compress_vecs(DistVec<double> &vectors, size_t start_idx, size_t end_idx, unsigned int compress_size,
                          std::vector<size_t> &, std::vector<bool> &, std::vector<bool> &, std::mt19937 &rn_gen) {
    vectors.compress_rows_(start_idx, end_idx, compress_size, 0, rn_gen);
}
inline void compress_vecs_sys(DistVec<double> &vectors, size_t start_idx, size_t end_idx, unsigned int compress_size,
                              std::vector<size_t> &, std::vector<bool> &, std::vector<bool> &, std::mt19937 &rn_gen) {
    vectors.compress_rows_(start_idx, end_idx, compress_size, 1, rn_gen);
}
inline void compress_vecs_multi(DistVec<double> &vectors, size_t start_idx, size_t end_idx, unsigned int compress_size,
                                std::vector<size_t> &, std::vector<bool> &, std::vector<bool> &, std::mt19937 &rn_gen) {
    vectors.compress_rows_(start_idx, end_idx, compress_size, 2, rn_gen);
}
