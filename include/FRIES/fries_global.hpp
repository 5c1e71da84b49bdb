// Reference-signature layer: the global functions and class templates of the reference's public headers
// (FRIES/CMakeLists.txt:16-18: compress_utils.hpp det_store.h fci_utils.h math_utils.h ndarr.hpp vec_utils.hpp hh_vec.hpp,
// Hamiltonians/hub_holstein.hpp) over libfries_b200.so, so that a program written against <FRIES/...> -- the reference's own
// tests/test_compression.cpp, tests/test_bitstrings.cpp, tests/test_vector.cpp, examples/fries_test.cpp -- compiles
// UNMODIFIED with -I<repo>/include and links with -lfries_b200.  Everything is a thin inline wrapper:
//   * device work goes through the C-ABI (include/fries_b200.h) on a process-wide default context (GPU FRIES_DEVICE, or 0);
//   * the small byte-string / sorted-list helpers are host arithmetic, as they are in the reference (FRIES/*.c);
//   * <mpi.h> is replaced by the single-process stand-in below: one process drives the GPU(s), MPI_COMM_WORLD has one rank
//     (several GPUs: host/fries_launch).
// Each function cites the reference declaration it stands for.
#pragma once
#include "../../fries_b200/host/fries_host.hpp"

#include <chrono>
#include <limits>

// ---- single-process MPI stand-in (the symbols the reference's headers and tests use) ----------------------
#ifndef MPI_VERSION
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_DATATYPE_NULL 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_UNSIGNED 5
#define MPI_UINT8_T 1
#define MPI_UINT16_T 2
#define MPI_UINT32_T 6
#define MPI_UINT64_T 9
#define MPI_VERSION 0
inline int fries_mpi_size_(MPI_Datatype t) { return t == MPI_UINT8_T ? 1 : t == MPI_UINT16_T ? 2 : (t == MPI_INT || t == MPI_UNSIGNED || t == MPI_UINT32_T) ? 4 : 8; }
inline int MPI_Init(int *, char ***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
inline int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm) {
    if (s != MPI_IN_PLACE) std::memcpy(r, s, (size_t)n * fries_mpi_size_(t));
    return 0;
}
inline int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm) {
    if (s != MPI_IN_PLACE) std::memcpy(r, s, (size_t)n * fries_mpi_size_(t));
    return 0;
}
inline int MPI_Scatter(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm) {
    if (r != MPI_IN_PLACE) std::memcpy(r, s, (size_t)n * fries_mpi_size_(t));
    return 0;
}
#endif

inline fries::Context &fries_default_context() {
    static fries::Context ctx(std::getenv("FRIES_DEVICE") ? std::atoi(std::getenv("FRIES_DEVICE")) : 0);
    return ctx;
}

// ---- math_utils.h ---------------------------------------------------------------------------------------------------------
#define CEILING(x, y) ((x + y - 1) / y)                                   /* math_utils.h:13 */
#define TRI_N(n) (((n) * (n + 1)) / 2)                                    /* :15 */
#define I_J_TO_TRI_NODIAG(i, j) (TRI_N(j - 1) + i)                        /* :16 */
#define I_J_TO_TRI_WDIAG(i, j) (TRI_N(j) + i)                             /* :17 */
inline uint8_t find_bits(const uint8_t *bit_str, uint8_t *bits, uint8_t n_bytes) { return fries::find_bits(bit_str, bits, n_bytes); }  // :38
inline uint8_t find_diff_bits(const uint8_t *str1, const uint8_t *str2, uint8_t *bits, uint8_t n_bytes) {  // :48
    return fries::find_diff_bits(str1, str2, bits, n_bytes);
}
inline unsigned int bits_between(uint8_t *bit_str, uint8_t a, uint8_t b) { return fries::bits_between(bit_str, a, b); }  // :60
// :72 -- the sorted list `orig_list` with the element at del_idx replaced by new_el, sorted again
inline void new_sorted(uint8_t *orig_list, uint8_t *new_list, uint8_t length, uint8_t del_idx, uint8_t new_el) {
    uint8_t w = 0;
    bool placed = false;
    for (uint8_t r = 0; r < length; r++) {
        if (r == del_idx) continue;
        if (!placed && new_el < orig_list[r]) {
            new_list[w++] = new_el;
            placed = true;
        }
        new_list[w++] = orig_list[r];
    }
    if (!placed) new_list[w] = new_el;
}
inline void repl_sorted(uint8_t *srt_list, uint8_t length, uint8_t del_idx, uint8_t new_el) {  // :83, in place
    uint8_t tmp[256];
    new_sorted(srt_list, tmp, length, del_idx, new_el);
    std::memcpy(srt_list, tmp, length);
}

// ---- det_store.h ----------------------------------------------------------------------------------------------------------
inline int read_bit(const uint8_t *bit_str, uint8_t bit_idx) { return fries::read_bit(bit_str, bit_idx); }  // det_store.h:23-26
inline void zero_bit(uint8_t *bit_str, uint8_t bit_idx) { fries::zero_bit(bit_str, bit_idx); }          // :33
inline void set_bit(uint8_t *bit_str, uint8_t bit_idx) { fries::set_bit(bit_str, bit_idx); }            // :40
inline void print_str(uint8_t *bit_str, uint8_t n_bytes, char *out_str) { fries::print_str(bit_str, n_bytes, out_str); }  // :48

// ---- fci_utils.h ----------------------------------------------------------------------------------------------------------
inline void gen_hf_bitstring(unsigned int n_orb, unsigned int n_elec, uint8_t *det) { fries::gen_hf_bitstring(n_orb, n_elec, det); }  // :26
inline int doub_det_parity(uint8_t *det, uint8_t *orbs) { return fries::doub_det_parity(det, orbs); }   // :43
inline int doub_parity(uint8_t *det, uint8_t *orbs) { return fries::doub_parity(det, orbs); }           // :58
inline void doub_det(uint8_t *det, uint8_t *orbs) { fries::doub_det(det, orbs); }                        // :69
inline int sing_det_parity(uint8_t *det, uint8_t *orbs) { return fries::sing_det_parity(det, orbs); }   // :97
inline int sing_parity(uint8_t *det, uint8_t *orbs) { return fries::sing_parity(det, orbs); }           // :112
inline void sing_det(uint8_t *det, uint8_t *orbs) { fries::sing_det(det, orbs); }                        // :122
inline int excite_sign(uint8_t cre_op, uint8_t des_op, uint8_t *det) { return fries::excite_sign(cre_op, des_op, det); }  // :151
// :162 -- sign of moving the electron at position occ_idx of the sorted list to orbital virt_orb: (-1)^(electrons passed)
inline int excite_sign_occ(uint8_t occ_idx, uint8_t virt_orb, const uint8_t *occ_orbs, uint32_t n_elec) {
    unsigned passed = 0;
    const uint8_t from = occ_orbs[occ_idx];
    for (uint32_t e = 0; e < n_elec; e++) {
        if (e == occ_idx) continue;
        const uint8_t o = occ_orbs[e];
        if ((from < virt_orb && o > from && o < virt_orb && e > occ_idx) || (from > virt_orb && o < from && o > virt_orb && e < occ_idx)) passed++;
    }
    return (passed & 1) ? -1 : 1;
}
inline uint8_t find_nth_virt(uint8_t *occ_orbs, int spin, uint8_t n_elec, uint8_t n_orb, uint8_t n) {    // :174
    return fries::find_nth_virt(occ_orbs, spin, n_elec, n_orb, n);
}
inline void flip_spins(uint8_t *det_in, uint8_t *det_out, uint8_t n_orb) { fries::flip_spins(det_in, det_out, n_orb); }  // :184
// :134 -- occupied list (two sorted spin halves) after the single excitation electron index ex_orbs[0] -> orbital ex_orbs[1]
inline void sing_ex_orbs(uint8_t *curr_orbs, uint8_t *new_orbs, uint8_t *ex_orbs, uint8_t n_elec) {
    const uint8_t half = n_elec / 2, shift = (uint8_t)((ex_orbs[0] / half) * half), other = (uint8_t)(half - shift);
    new_sorted(curr_orbs + shift, new_orbs + shift, half, (uint8_t)(ex_orbs[0] - shift), ex_orbs[1]);
    std::memcpy(new_orbs + other, curr_orbs + other, half);
}
// :81 -- the same for a double excitation: electron indices ex_orbs[0], ex_orbs[1] -> orbitals ex_orbs[2], ex_orbs[3]; for
// two electrons of one spin the second index addresses the list as it is AFTER the first replacement (fci_utils.c:96-108)
inline void doub_ex_orbs(uint8_t *curr_orbs, uint8_t *new_orbs, uint8_t *ex_orbs, uint8_t n_elec) {
    const uint8_t half = n_elec / 2, s1 = (uint8_t)((ex_orbs[0] / half) * half), s2 = (uint8_t)((ex_orbs[1] / half) * half);
    new_sorted(curr_orbs + s1, new_orbs + s1, half, (uint8_t)(ex_orbs[0] - s1), ex_orbs[2]);
    if (s1 == s2) {
        std::memcpy(new_orbs + (half - s1), curr_orbs + (half - s1), half);
        repl_sorted(new_orbs + s1, half, (uint8_t)(ex_orbs[1] - s1), ex_orbs[3]);
    } else {
        new_sorted(curr_orbs + s2, new_orbs + s2, half, (uint8_t)(ex_orbs[1] - s2), ex_orbs[3]);
    }
}
// :196 -- orbitals by which two determinants differ: those only in str1 first, then those only in str2 (ascending);
// returns the excitation rank 0, 1, 2, or UINT8_MAX when more than two orbitals differ
inline uint8_t find_excitation(const uint8_t *str1, const uint8_t *str2, uint8_t *orbs, uint8_t n_bytes) {
    uint8_t only1[3], only2[3], n1 = 0, n2 = 0;
    for (unsigned bit = 0; bit < 8u * n_bytes; bit++) {
        const int a = (str1[bit / 8] >> (bit % 8)) & 1, b = (str2[bit / 8] >> (bit % 8)) & 1;
        if (a && !b) {
            if (n1 == 2) return UINT8_MAX;
            only1[n1++] = (uint8_t)bit;
        } else if (b && !a) {
            if (n2 < 2) only2[n2] = (uint8_t)bit;
            n2++;
        }
    }
    if (n1 == 0) return 0;
    for (uint8_t k = 0; k < n1; k++) orbs[k] = only1[k];
    for (uint8_t k = 0; k < n1 && k < n2; k++) orbs[n1 + k] = only2[k];
    return n1;
}
// :207 -- is a determinant connected to its time-reversed partner by a double excitation?  0: it IS its partner, 1: yes
// (diff_idx: the positions of the two differing electrons in occ_orbs), 2: no
inline int tr_doub_connect(const uint8_t *occ_orbs, uint32_t n_orb, uint32_t n_elec, uint8_t *diff_idx) {
    const uint32_t half = n_elec / 2;
    uint64_t a = 0, b = 0;
    for (uint32_t e = 0; e < half; e++) {
        a |= 1ull << occ_orbs[e];
        b |= 1ull << (occ_orbs[half + e] - n_orb);
    }
    if (a == b) return 0;
    const uint64_t only_a = a & ~b, only_b = b & ~a;
    if (__builtin_popcountll(only_a) != 1 || __builtin_popcountll(only_b) != 1) return 2;
    for (uint32_t e = 0; e < half; e++) {
        if ((only_a >> occ_orbs[e]) & 1) diff_idx[0] = (uint8_t)e;
        if ((only_b >> (occ_orbs[half + e] - n_orb)) & 1) diff_idx[1] = (uint8_t)(half + e);
    }
    return 1;
}

// ---- ndarr.hpp ------------------------------------------------------------------------------------------------------------
template <class T>
using Matrix = fries::Matrix<T>;  // ndarr.hpp:16-150
// FourDArr ndarr.hpp:152-204: dense rank-four array of doubles
class FourDArr {
    size_t len_[4];
    std::vector<double> data_;

  public:
    FourDArr(size_t len1, size_t len2, size_t len3, size_t len4) : len_{len1, len2, len3, len4}, data_(len1 * len2 * len3 * len4) {}
    double &operator()(size_t i1, size_t i2, size_t i3, size_t i4) { return data_[((i1 * len_[1] + i2) * len_[2] + i3) * len_[3] + i4]; }
    double operator()(size_t i1, size_t i2, size_t i3, size_t i4) const { return data_[((i1 * len_[1] + i2) * len_[2] + i3) * len_[3] + i4]; }
    double *data() { return data_.data(); }
    FourDArr(const FourDArr &) = delete;
    FourDArr &operator=(const FourDArr &) = delete;
};
// SymmERIs ndarr.hpp:206-244: two-electron integrals with eight-fold symmetry, packed (the layout fries_mol_create takes)
class SymmERIs {
    size_t n_orb_;
    std::vector<double> data_;
    static size_t tri(size_t a, size_t b) { return a < b ? b * (b + 1) / 2 + a : a * (a + 1) / 2 + b; }

  public:
    SymmERIs(size_t n_orb) : n_orb_(n_orb), data_(tri(tri(n_orb - 1, n_orb - 1), tri(n_orb - 1, n_orb - 1)) + 1) {}
    double &chemist(size_t i1, size_t i2, size_t i3, size_t i4) { return data_[tri(tri(i1, i2), tri(i3, i4))]; }   // :219-230
    double chemist(size_t i1, size_t i2, size_t i3, size_t i4) const { return data_[tri(tri(i1, i2), tri(i3, i4))]; }
    double &physicist(size_t i1, size_t i2, size_t i3, size_t i4) { return chemist(i1, i3, i2, i4); }             // :232-239
    double physicist(size_t i1, size_t i2, size_t i3, size_t i4) const { return chemist(i1, i3, i2, i4); }
    double *data() { return data_.data(); }
    size_t n_orb() const { return n_orb_; }
};

// ---- compress_utils.hpp ---------------------------------------------------------------------------------------------------
inline int round_binomially(double p, unsigned int n, std::mt19937 &mt_obj) { return fries::round_binomially(p, n, mt_obj); }  // :28
inline double find_preserve(double *values, std::vector<size_t> &srt_idx, std::vector<bool> &keep_idx, size_t count,
                            unsigned int *n_samp, double *global_norm) {  // :54 -> find_preserve_kernel
    return fries::find_preserve(fries_default_context(), values, srt_idx, keep_idx, count, n_samp, global_norm);
}
inline void sys_comp(double *vec_vals, size_t vec_len, double *loc_norms, unsigned int n_samp, std::vector<bool> &keep_exact,
                     double rand_num) {  // :74 -> sys_comp_kernel
    fries::sys_comp(fries_default_context(), vec_vals, vec_len, loc_norms, n_samp, keep_exact, rand_num);
}
inline void sys_comp(double *vec_vals, size_t vec_len, double *loc_norms, unsigned int n_samp, std::vector<bool> &keep_exact,
                     double rand_num, MPI_Comm) {  // :76
    sys_comp(vec_vals, vec_len, loc_norms, n_samp, keep_exact, rand_num);
}
// :92 -- one segment resampled on its own: the grid of seg_norm / n_samp, selected elements get the magnitude sampl_val.
// The device kernel resamples (it assigns seg_norm / n_samp); the magnitude is replaced here when the caller wants another.
inline void sys_comp_serial(double *vec_vals, size_t vec_len, double seg_norm, double sampl_val, uint32_t n_samp,
                            std::vector<bool> &keep_exact, double rand_num) {
    std::vector<double> before(vec_vals, vec_vals + vec_len);
    std::vector<bool> was_kept(keep_exact.begin(), keep_exact.begin() + vec_len);
    double norms[1] = {n_samp ? seg_norm : 0.0};
    sys_comp(vec_vals, vec_len, norms, n_samp, keep_exact, rand_num);
    for (size_t i = 0; i < vec_len; i++)
        if (!was_kept[i] && vec_vals[i] != 0) vec_vals[i] = before[i] > 0 ? sampl_val : -sampl_val;
}
inline void piv_samp_serial(double *vec_vals, size_t vec_len, double seg_norm, uint32_t n_samp, std::vector<bool> &keep_exact,
                            std::mt19937 &mt_obj) {  // :119 -> piv_samp_kernel
    fries::piv_samp_serial(fries_default_context(), vec_vals, vec_len, seg_norm, n_samp, keep_exact, mt_obj);
}
inline double seed_sys(double *norms, double *rn, unsigned int n_samp) { return fries::seed_sys(norms, rn, n_samp); }  // :42
inline void adjust_shift(double *shift, double one_norm, double *last_norm, double target_norm, double damp_factor) {  // :170
    fries::adjust_shift(shift, one_norm, last_norm, target_norm, damp_factor);
}
inline double sum_mpi(double local, int my_rank, int n_procs) { return fries::sum_mpi(local, my_rank, n_procs); }  // :179-231
inline int sum_mpi(int local, int my_rank, int n_procs) { return fries::sum_mpi(local, my_rank, n_procs); }
inline uint64_t sum_mpi(uint64_t local, int my_rank, int n_procs) { return fries::sum_mpi(local, my_rank, n_procs); }
// Walker's alias method, :394 / :411 (setup) and :414-429 (sampling): host arithmetic on a handful of states, as in the
// reference; two generator outputs per sample, in the reference's order (column, then the accept test)
inline void setup_alias(double *probs, unsigned int *aliases, double *alias_probs, size_t n_states) {
    std::vector<unsigned int> light, heavy;
    for (unsigned int i = 0; i < n_states; i++) {
        aliases[i] = i;
        alias_probs[i] = n_states * probs[i];
        (alias_probs[i] < 1 ? light : heavy).push_back(i);
    }
    while (!light.empty() && !heavy.empty()) {
        const unsigned int s = light.back(), b = heavy.back();
        aliases[s] = b;                          // column s: itself with alias_probs[s], else b
        alias_probs[b] += alias_probs[s] - 1;    // b gave away the rest of column s
        if (alias_probs[b] < 1) {
            light.back() = b;
            heavy.pop_back();
        } else {
            light.pop_back();
        }
    }
}
inline void sample_alias(unsigned int *aliases, double *alias_probs, size_t n_states, uint8_t *samples, unsigned int n_samp,
                         size_t samp_int, std::mt19937 &mt_obj) {
    if (n_states > std::numeric_limits<uint8_t>::max()) throw std::runtime_error("sample_alias: more than 255 states");
    for (unsigned int s = 0; s < n_samp; s++) {
        const uint8_t col = (uint8_t)(mt_obj() / (1. + UINT32_MAX) * n_states);
        samples[s * samp_int] = (mt_obj() / (1. + UINT32_MAX) < alias_probs[col]) ? col : (uint8_t)aliases[col];
    }
}
inline void sample_alias(unsigned int *aliases, double *alias_probs, size_t n_states, uint16_t *counts, unsigned int n_samp,
                         std::mt19937 &mt_obj) {
    if (n_states > std::numeric_limits<uint16_t>::max()) throw std::runtime_error("sample_alias: more than 65535 states");
    for (unsigned int s = 0; s < n_samp; s++) {
        const uint16_t col = (uint16_t)(mt_obj() / (1. + UINT32_MAX) * n_states);
        counts[(mt_obj() / (1. + UINT32_MAX) < alias_probs[col]) ? col : aliases[col]]++;
    }
}
