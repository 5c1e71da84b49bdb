// fries_host.hpp -- C++ host layer above the C-ABI (include/fries_b200.h).
//
// Mirrors the parts of the reference's host interface that the in-scope drivers use, with the reference's
// names, argument meaning and error behaviour (std::runtime_error, caught in main):
//   DistVec            FRIES/vec_utils.hpp:121-953      -> fries::DistVec (device-resident store)
//   find_preserve ...  FRIES/compress_utils.hpp:28-429  -> fries::find_preserve / sys_comp / comp_sub / adjust_shift /
//                      piv_samp_serial / piv_budget / adjust_probs / piv_comp_parallel
//   parse_fcidump ...  FRIES/io_utils.hpp:23-207        -> fries::parse_fcidump / parse_hf_input / load_vec_txt / ...
//   argparse kwargs    FRIES/Ext_Libs/argparse.hpp      -> fries::Args (same --key value syntax and failure modes)
// All arithmetic on vectors happens in libfries_b200.so (CUDA); this file is marshalling, file formats, control flow.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fries_b200.h"

namespace fries {

inline void check(int rc) {
    if (rc != FRIES_OK) throw std::runtime_error(fries_last_error());
}

// ---- FRIES/math_utils.h macros ---------------------------------------------------------------------------------
inline size_t ceiling(size_t x, size_t y) { return (x + y - 1) / y; }
inline size_t tri_wdiag(size_t i, size_t j) { return j * (j + 1) / 2 + i; }

// ---- determinant bit strings: the reference's uint8_t[] <-> the library's u64 key -------------------------------
inline uint64_t key_from_bytes(const uint8_t *det, size_t n_bytes) {
    uint64_t k = 0;
    for (size_t b = 0; b < n_bytes && b < 8; b++) k |= (uint64_t)det[b] << (8 * b);
    return k;
}
inline void key_to_bytes(uint64_t k, uint8_t *det, size_t n_bytes) {
    for (size_t b = 0; b < n_bytes; b++) det[b] = b < 8 ? (uint8_t)(k >> (8 * b)) : 0;
}
// gen_hf_bitstring FRIES/fci_utils.c:10-43
inline uint64_t gen_hf_bitstring(unsigned n_orb, unsigned n_elec) {
    uint64_t k = 0;
    for (unsigned i = 0; i < n_elec / 2; i++) k |= (1ull << i) | (1ull << (i + n_orb));
    return k;
}

// ---- Matrix<T> FRIES/ndarr.hpp:20-140: row-major 2-D array with the reference's accessors ---------------------------
template <class T>
class Matrix {
    size_t rows_ = 0, cols_ = 0;
    std::vector<T> data_;

  public:
    Matrix() = default;
    Matrix(size_t rows, size_t cols) : rows_(rows), cols_(cols), data_(rows * cols) {}
    T &operator()(size_t row, size_t col) { return data_[row * cols_ + col]; }
    const T &operator()(size_t row, size_t col) const { return data_[row * cols_ + col]; }
    T *operator[](size_t row) { return data_.data() + row * cols_; }
    const T *operator[](size_t row) const { return data_.data() + row * cols_; }
    void reshape(size_t new_rows, size_t new_cols) {
        rows_ = new_rows;
        cols_ = new_cols;
        data_.resize(new_rows * new_cols);
    }
    size_t rows() const { return rows_; }
    size_t cols() const { return cols_; }
    T *data() { return data_.data(); }
    const T *data() const { return data_.data(); }
};

// find_bits FRIES/math_utils.c:62-98: positions of the 1 bits of a byte string, ascending; returns their number
inline uint8_t find_bits(const uint8_t *bit_str, uint8_t *bits, uint8_t n_bytes) {
    uint8_t n = 0;
    for (unsigned b = 0; b < n_bytes; b++)
        for (unsigned i = 0; i < 8; i++)
            if (bit_str[b] >> i & 1) bits[n++] = (uint8_t)(8 * b + i);
    return n;
}
// HashTable::hash_fxn FRIES/det_hash.hpp:160-170 (the device twin is fr_det_hash): the product (i + 1) * scrambler is
// taken in 32 bits, the recurrence in 64
inline uintmax_t hash_fxn(const uint32_t *scrambler, const uint8_t *occ_orbs, unsigned n_elec, const uint8_t *phonon_nums = nullptr,
                          unsigned n_phonon = 0) {
    uintmax_t hash = 0;
    for (unsigned i = 0; i < n_elec; i++) hash = 1099511628211ULL * hash + (uint32_t)((i + 1) * scrambler[occ_orbs[i]]);
    for (unsigned i = 0; i < n_phonon; i++) hash = 1099511628211ULL * hash + (uint32_t)((i + 1) * scrambler[phonon_nums[i]]);
    return hash;
}

// ---- FRIES/det_store.h, math_utils.h, fci_utils.h on the reference's byte strings (host twins of the device functions
// fr_* in csrc/common.cuh; determinants of up to 64 bits) --------------------------------------------------------------
inline int read_bit(const uint8_t *bit_str, uint8_t bit_idx) { return bit_str[bit_idx / 8] >> (bit_idx % 8) & 1; }   // det_store.h:23-26
inline void zero_bit(uint8_t *bit_str, uint8_t bit_idx) { bit_str[bit_idx / 8] &= (uint8_t)~(1u << (bit_idx % 8)); }  // det_store.c:11-15
inline void set_bit(uint8_t *bit_str, uint8_t bit_idx) { bit_str[bit_idx / 8] |= (uint8_t)(1u << (bit_idx % 8)); }    // det_store.c:17-21
// print_str det_store.c:23-29: hexadecimal, most significant byte first
inline void print_str(const uint8_t *bit_str, uint8_t n_bytes, char *out_str) {
    for (unsigned b = 0; b < n_bytes; b++) snprintf(out_str + 2 * b, 3, "%02x", bit_str[n_bytes - 1 - b]);
    out_str[2 * n_bytes] = 0;
}
// bits_between math_utils.c:9-58: 1 bits strictly between positions a and b
inline unsigned bits_between(const uint8_t *bit_str, uint8_t a, uint8_t b) {
    const unsigned lo = a < b ? a : b, hi = a < b ? b : a;
    const uint64_t key = key_from_bytes(bit_str, hi / 8 + 1);
    const uint64_t below_hi = hi ? (~0ull >> (64 - hi)) : 0ull, upto_lo = ~0ull >> (63 - lo);
    return (unsigned)__builtin_popcountll(key & below_hi & ~upto_lo);
}
// find_diff_bits math_utils.c:100-131: positions of the first 4 differing bits; UINT8_MAX when more than 4 differ
inline uint8_t find_diff_bits(const uint8_t *str1, const uint8_t *str2, uint8_t *bits, uint8_t n_bytes) {
    uint8_t n = 0;
    for (unsigned b = 0; b < n_bytes; b++) {
        uint8_t x = str1[b] ^ str2[b];
        for (unsigned i = 0; i < 8; i++)
            if (x >> i & 1) {
                if (n == 4) return UINT8_MAX;
                bits[n++] = (uint8_t)(8 * b + i);
            }
    }
    return n;
}
// excite_sign fci_utils.c:128-135: (-1)^(electrons strictly between the two operators)
inline int excite_sign(uint8_t cre_op, uint8_t des_op, const uint8_t *det) {
    return (bits_between(det, cre_op, des_op) & 1) ? -1 : 1;
}
// sing_det_parity fci_utils.c:46-51; orbs = {occupied, virtual}
inline int sing_det_parity(uint8_t *det, const uint8_t *orbs) {
    zero_bit(det, orbs[0]);
    int sign = excite_sign(orbs[1], orbs[0], det);
    set_bit(det, orbs[1]);
    return sign;
}
inline int sing_parity(const uint8_t *det, const uint8_t *orbs) { return excite_sign(orbs[1], orbs[0], det); }  // :54-57
inline void sing_det(uint8_t *det, const uint8_t *orbs) {                                                          // :59-62
    zero_bit(det, orbs[0]);
    set_bit(det, orbs[1]);
}
// doub_det_parity fci_utils.c:67-75; orbs = {occ, occ, virt, virt}: both electrons are removed first, then each
// replacement orbs[i] -> orbs[i + 2] contributes its sign on that doubly-annihilated string
inline int doub_det_parity(uint8_t *det, const uint8_t *orbs) {
    zero_bit(det, orbs[0]);
    zero_bit(det, orbs[1]);
    int sign = excite_sign(orbs[2], orbs[0], det) * excite_sign(orbs[3], orbs[1], det);
    set_bit(det, orbs[2]);
    set_bit(det, orbs[3]);
    return sign;
}
inline int doub_parity(const uint8_t *det, const uint8_t *orbs) {  // :86-94: the string is left as it is
    uint8_t tmp[8] = {0};
    const unsigned hi = std::max(std::max(orbs[0], orbs[1]), std::max(orbs[2], orbs[3]));
    memcpy(tmp, det, hi / 8 + 1);
    zero_bit(tmp, orbs[0]);
    zero_bit(tmp, orbs[1]);
    return excite_sign(orbs[2], orbs[0], tmp) * excite_sign(orbs[3], orbs[1], tmp);
}
inline void doub_det(uint8_t *det, const uint8_t *orbs) {  // :77-84
    zero_bit(det, orbs[0]);
    zero_bit(det, orbs[1]);
    set_bit(det, orbs[2]);
    set_bit(det, orbs[3]);
}
// find_nth_virt fci_utils.c:138-148: the n-th orbital of the given spin that is not in the occupied list
inline uint8_t find_nth_virt(const uint8_t *occ_orbs, int spin, uint8_t n_elec, uint8_t n_orb, uint8_t n) {
    uint8_t virt = (uint8_t)(n_orb * spin + n);
    for (unsigned i = n_elec / 2 * spin; i < n_elec && occ_orbs[i] <= virt; i++) virt++;
    return virt;
}
// flip_spins fci_utils.c:150-165: exchange the alpha (bits [0, n_orb)) and beta (bits [n_orb, 2 n_orb)) strings
inline void flip_spins(const uint8_t *det_in, uint8_t *det_out, uint8_t n_orb) {
    const size_t nb = ceiling(2 * (size_t)n_orb, 8);
    const uint64_t k = key_from_bytes(det_in, nb), half = (1ull << n_orb) - 1;
    key_to_bytes(((k & half) << n_orb) | (k >> n_orb & half), det_out, nb);
}
// gen_hf_bitstring fci_utils.c:10-43, byte-string form
inline void gen_hf_bitstring(unsigned n_orb, unsigned n_elec, uint8_t *det) {
    key_to_bytes(gen_hf_bitstring(n_orb, n_elec), det, ceiling(2 * (size_t)n_orb, 8));
}

// ---- command line: argparse::Args semantics (Ext_Libs/argparse.hpp:347-393) ---------------------------------------
class Args {
    std::map<std::string, std::string> kv_;
    std::vector<std::string> known_;
    std::string errors_;

  public:
    Args(int argc, char **argv) {
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i];
            if (a.rfind("--", 0) != 0) continue;
            a = a.substr(2);
            size_t eq = a.find('=');
            if (eq != std::string::npos) {
                kv_[a.substr(0, eq)] = a.substr(eq + 1);
            } else if (i + 1 < argc && (argv[i + 1][0] != '-' || std::isdigit((unsigned char)argv[i + 1][1]))) {
                kv_[a] = argv[++i];
            } else {
                kv_[a] = "";
                errors_ += "No value provided for: " + a + "\n";
            }
        }
    }
    bool has(const std::string &k) {
        known_.push_back(k);
        return kv_.count(k) != 0;
    }
    std::string str(const std::string &k) {  // required
        if (!has(k)) errors_ += "Argument missing: --" + k + "\n";
        return kv_[k];
    }
    std::string str(const std::string &k, const std::string &dflt) { return has(k) ? kv_[k] : dflt; }
    double num(const std::string &k) {
        std::string s = str(k);
        return s.empty() ? 0 : std::stod(s);
    }
    double num(const std::string &k, double dflt) { return has(k) ? std::stod(kv_[k]) : dflt; }
    void validate() {  // unknown flag -> warning and continue; missing required -> exit(-1)
        for (auto &p : kv_) {
            bool ok = false;
            for (auto &k : known_) ok = ok || k == p.first;
            if (!ok) std::cerr << "unrecognised commandline argument: " << p.first << std::endl;
        }
        if (!errors_.empty()) {
            std::cerr << errors_;
            std::exit(-1);
        }
    }
};

// ---- input files -----------------------------------------------------------------------------------------------------
struct MolInput {
    unsigned n_elec = 0;  // TOTAL electrons
    unsigned n_frz = 0;
    unsigned n_orb = 0;   // unfrozen spatial orbitals
    double eps = 0, hf_en = 0, core_en = 0;
    std::vector<uint8_t> symm;          // irreps of the unfrozen orbitals
    std::vector<double> hcore;          // tot_orb^2
    std::vector<double> eris_packed;    // SymmERIs layout (FRIES/ndarr.hpp:206-244) over tot_orb orbitals
    unsigned tot_orb() const { return n_orb + n_frz / 2; }
};

// convert_symm FRIES/io_utils.cpp:189-239
inline void convert_symm(std::vector<uint8_t> &irreps, const std::string &pg) {
    auto apply = [&](const std::vector<uint8_t> &map, unsigned maxi) {
        for (auto &x : irreps) {
            if (x > maxi || x < 1) {
                std::stringstream msg;
                msg << "irrep index " << (unsigned)x << " read from the FCIDUMP file exceeds the maximum allowed irrep index ("
                    << maxi << ") for point group " << pg;
                throw std::runtime_error(msg.str());
            }
            x = map[x - 1];
        }
    };
    if (pg == "D2h" || pg == "d2h") apply({0, 7, 6, 1, 5, 2, 3, 4}, 8);
    else if (pg == "C2v" || pg == "c2v" || pg == "c2h" || pg == "C2h") apply({0, 2, 3, 1}, 4);
    else if (pg == "D2" || pg == "d2") apply({0, 3, 2, 1}, 4);
    else if (pg == "Cs" || pg == "cs" || pg == "C2" || pg == "c2" || pg == "ci" || pg == "Ci") apply({0, 1}, 2);
    else if (pg == "C1" || pg == "c1") apply({0}, 1);
    else throw std::runtime_error("Point group " + pg + " not recognized");
}

// parse_fcidump FRIES/io_utils.cpp:241-318
inline MolInput parse_fcidump(const std::string &path, const std::string &point_group) {
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error("Could not open FCIDUMP file " + path);
    std::string line;
    std::getline(in, line);
    auto field = [&](const std::string &key) {
        size_t p = line.find(key);
        if (p == std::string::npos) throw std::runtime_error("Could not find " + key + " in the first line of the FCIDUMP file");
        size_t e = line.find(",", p);
        return std::stoi(line.substr(p + key.size(), e - (p + key.size())));
    };
    MolInput m;
    m.n_orb = field("NORB=");
    m.n_elec = field("NELEC=");
    if (field("MS2=") != 0) throw std::runtime_error("MS2 is not zero in FCIDUMP file.");
    std::getline(in, line);
    size_t p = line.find("ORBSYM=");
    std::stringstream ss(line.substr(p == std::string::npos ? 0 : p + 7));
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        try {
            if (!tok.empty()) m.symm.push_back((uint8_t)std::stoi(tok));
        } catch (std::invalid_argument &) {
        }
    }
    if (m.symm.size() != m.n_orb)
        throw std::runtime_error("Number of irrep labels read in after ORBSYM in FCIDUMP file does not equal number of orbitals");
    convert_symm(m.symm, point_group);
    std::getline(in, line);  // ISYM
    std::getline(in, line);  // &END
    size_t T = m.n_orb, n_pair = T * (T + 1) / 2;
    m.hcore.assign(T * T, 0.0);
    m.eris_packed.assign(n_pair * (n_pair + 1) / 2, 0.0);
    double v;
    unsigned o[4];
    while (in >> v >> o[0] >> o[1] >> o[2] >> o[3]) {
        if (!o[0] && !o[1] && !o[2] && !o[3]) m.core_en = v;
        else if (!o[1] && !o[2] && !o[3]) continue;
        else if (!o[2] && !o[3]) m.hcore[(o[1] - 1) * T + o[0] - 1] = m.hcore[(o[0] - 1) * T + o[1] - 1] = v;
        else  // chemist_ordered(o3-1, o2-1, o1-1, o0-1): stored exactly where the reference stores it
            m.eris_packed[tri_wdiag(tri_wdiag(o[3] - 1, o[2] - 1), tri_wdiag(o[1] - 1, o[0] - 1))] = v;
    }
    return m;
}

inline size_t read_csv(std::vector<double> &out, const std::string &path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("Could not open file " + path);
    std::string tok;
    out.clear();
    char c;
    while (f.get(c)) {
        if (c == ',' || c == '\n' || c == ' ' || c == '\r' || c == '\t') {
            if (!tok.empty()) out.push_back(std::stod(tok));
            tok.clear();
        } else {
            tok += c;
        }
    }
    if (!tok.empty()) out.push_back(std::stod(tok));
    return out.size();
}

// parse_hf_input FRIES/io_utils.cpp:98-187 (legacy directory: sys_params.txt, symm.txt, hcore.txt, eris.txt)
inline MolInput parse_hf_input(const std::string &dir) {
    std::ifstream in(dir + "sys_params.txt");
    if (!in.is_open()) throw std::runtime_error("Could not open file sys_params.txt");
    MolInput m;
    std::string key;
    auto expect = [&](const char *name, auto &dst) {
        std::getline(in, key);
        if (key.empty()) std::getline(in, key);
        if (key != name) throw std::runtime_error(std::string("Could not find ") + name + " parameter in sys_params.txt");
        in >> dst;
        std::getline(in, key);
    };
    expect("n_elec", m.n_elec);
    expect("n_frozen", m.n_frz);
    expect("n_orb", m.n_orb);
    expect("eps", m.eps);
    expect("hf_energy", m.hf_en);
    size_t T = m.tot_orb();
    std::vector<double> tmp;
    read_csv(tmp, dir + "symm.txt");
    if (tmp.size() < T) throw std::runtime_error("Could not read the orbital irreps from symm.txt");
    for (size_t i = m.n_frz / 2; i < T; i++) m.symm.push_back((uint8_t)tmp[i]);
    if (read_csv(m.hcore, dir + "hcore.txt") < T * T) {
        std::stringstream msg;
        msg << "Could not read " << T * T << " elements from " << dir << "hcore.txt";
        throw std::runtime_error(msg.str());
    }
    if (read_csv(tmp, dir + "eris.txt") < T * T * T * T) {
        std::stringstream msg;
        msg << "Could not read " << T * T * T * T << " elements from " << dir << "eris.txt";
        throw std::runtime_error(msg.str());
    }
    // eris.txt is the dense FourDArr eris(i,j,a,b) = <ij|ab> = (ia|jb): pack the canonical representatives
    size_t n_pair = T * (T + 1) / 2;
    m.eris_packed.assign(n_pair * (n_pair + 1) / 2, 0.0);
    for (size_t i = 0; i < T; i++)
        for (size_t a = i; a < T; a++)
            for (size_t j = 0; j < T; j++)
                for (size_t b = j; b < T; b++) {
                    size_t p1 = tri_wdiag(i, a), p2 = tri_wdiag(j, b);
                    if (p1 <= p2) m.eris_packed[tri_wdiag(p1, p2)] = tmp[((i * T + j) * T + a) * T + b];
                }
    return m;
}

// parse_hh_input FRIES/io_utils.cpp:320-408: key / value lines in a fixed order
struct HhInput {
    unsigned n_elec = 0, lat_len = 0, n_dim = 0;
    double eps = 0, elec_int = 0, ph_freq = 0, elec_ph = 0, hf_en = 0;
};
inline HhInput parse_hh_input(const std::string &path) {
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error("Could not open file containing Hubbard-Holstein parameters");
    HhInput h;
    std::string key;
    auto expect = [&](const char *name, const char *what, auto &dst) {
        std::getline(in, key);
        if (key.empty()) std::getline(in, key);
        if (key != name)
            throw std::runtime_error(std::string("Could not find ") + what + " in file containing Hubbard-Holstein parameters");
        in >> dst;
        std::getline(in, key);
    };
    expect("n_elec", "n_elec parameter", h.n_elec);
    expect("lat_len", "lat_len parameter", h.lat_len);
    expect("n_dim", "n_dim parameter", h.n_dim);
    expect("eps", "eps parameter", h.eps);
    expect("U", "electron interaction parameter (U)", h.elec_int);
    expect("omega", "phonon frequency parameter (omega)", h.ph_freq);
    expect("g", "electron-phonon interaction parameter (g)", h.elec_ph);
    expect("gs_energy", "gs_energy parameter", h.hf_en);
    return h;
}
// gen_neel_det_1D FRIES/Hamiltonians/hub_holstein.cpp:139-171: up spins on even sites, down spins on odd sites
inline uint64_t gen_neel_det_1D(unsigned n_sites, unsigned n_elec) {
    uint64_t k = 0;
    for (unsigned e = 0; e < n_elec / 2; e++) k |= (1ull << (2 * e)) | (1ull << (n_sites + 2 * e + 1));
    return k;
}

// read_dets + load_vec_txt FRIES/io_utils.cpp:410-482,565-586
inline size_t load_vec_txt(const std::string &prefix, std::vector<uint64_t> &dets, std::vector<double> &vals) {
    std::ifstream fd(prefix + "dets");
    if (!fd.is_open()) throw std::runtime_error("Could not open file: " + prefix + "dets");
    long long d;
    dets.clear();
    while (fd >> d) dets.push_back((uint64_t)d);
    std::ifstream fv(prefix + "vals");
    if (!fv.is_open()) throw std::runtime_error("Could not open file: " + prefix + "vals");
    double v;
    vals.clear();
    while (fv >> v) vals.push_back(v);
    if (vals.size() > dets.size()) {
        std::cerr << "Warning: fewer determinants (" << dets.size() << ") than values (" << vals.size() << ") read in\n";
        vals.resize(dets.size());
    } else if (vals.size() < dets.size()) {
        std::cerr << "Warning: fewer values (" << vals.size() << ") than determinants (" << dets.size() << ") read in\n";
        dets.resize(vals.size());
    }
    return vals.size();
}

// save_proc_hash / load_proc_hash FRIES/io_utils.cpp:589-619
inline void save_proc_hash(const std::string &dir, const std::vector<uint32_t> &scr) {
    std::ofstream f(dir + "hash.dat", std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Could not save file at path " + dir + "hash.dat");
    f.write((const char *)scr.data(), sizeof(uint32_t) * scr.size());
}
inline void load_proc_hash(const std::string &dir, std::vector<uint32_t> &scr) {
    std::ifstream f(dir + "hash.dat", std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Could not open saved hash scrambler at " + dir + "hash.dat");
    f.read((char *)scr.data(), sizeof(uint32_t) * scr.size());
}
// load_last_line FRIES/io_utils.cpp:636-663 (one number per line files)
inline bool load_last_line(const std::string &path, double *val) {
    std::ifstream f(path);
    if (!f.is_open()) return false;
    std::string line, last;
    while (std::getline(f, line))
        if (!line.empty()) last = line;
    if (last.empty()) return false;
    *val = std::stod(last);
    return true;
}

// ---- RAII handles -------------------------------------------------------------------------------------------------------
struct Context {
    fries_ctx *h = nullptr;
    explicit Context(int device = 0) { check(fries_ctx_create(device, &h)); }
    ~Context() { fries_ctx_destroy(h); }
    Context(const Context &) = delete;
};

// ---- several GPUs: one process per GPU, started by host/fries_launch (the reference: mpirun + MPI_COMM_WORLD) -------------
// The launcher exports FRIES_NRANKS, FRIES_RANK and FRIES_RDV (a private directory); the ranks exchange small byte strings
// (CUDA IPC handles, the seed) through files in that directory -- the only host-side "collective" the GPU path needs: all
// per-iteration communication happens inside kernels over peer-mapped memory (csrc/comm.cuh).
struct Ranks {
    int n = 1, rank = 0;
    std::string rdv;
    Ranks() {
        const char *a = std::getenv("FRIES_NRANKS"), *b = std::getenv("FRIES_RANK"), *c = std::getenv("FRIES_RDV");
        if (a && b && c && std::atoi(a) > 1) {
            n = std::atoi(a);
            rank = std::atoi(b);
            rdv = c;
            if (rank < 0 || rank >= n) throw std::runtime_error("FRIES_RANK out of range");
        }
    }
    // all-gather of `bytes` bytes per rank (MPI_Allgather): out = n x bytes in rank order
    std::vector<uint8_t> allgather(const std::string &tag, const void *mine, size_t bytes) const {
        std::vector<uint8_t> out(n * bytes);
        if (n == 1) {
            std::memcpy(out.data(), mine, bytes);
            return out;
        }
        const std::string base = rdv + "/" + tag + ".";
        {
            const std::string tmp = base + std::to_string(rank) + ".tmp", fin = base + std::to_string(rank);
            std::ofstream f(tmp, std::ios::binary);
            f.write((const char *)mine, (std::streamsize)bytes);
            f.close();
            if (std::rename(tmp.c_str(), fin.c_str()) != 0) throw std::runtime_error("rendezvous: cannot write " + fin);
        }
        for (int r = 0; r < n; r++) {
            const std::string fin = base + std::to_string(r);
            for (long spin = 0;; spin++) {
                std::ifstream f(fin, std::ios::binary);
                if (f.is_open()) {
                    f.read((char *)out.data() + (size_t)r * bytes, (std::streamsize)bytes);
                    if ((size_t)f.gcount() == bytes) break;
                }
                if (spin > 600000) throw std::runtime_error("rendezvous: rank " + std::to_string(r) + " did not arrive (" + tag + ")");
                struct timespec ts = {0, 200000};
                nanosleep(&ts, nullptr);
            }
        }
        return out;
    }
    void barrier(const std::string &tag) const {
        uint8_t b = 1;
        allgather(tag, &b, 1);
    }
    // MPI_Bcast from rank 0
    template <class T>
    T bcast0(const std::string &tag, T v) const {
        std::vector<uint8_t> all = allgather(tag, &v, sizeof(T));
        T r;
        std::memcpy(&r, all.data(), sizeof(T));
        return r;
    }
};

// peer-mapped inboxes + spawn-route windows of the ranks (fries_comm, include/fries_b200.h)
struct Comm {
    fries_comm *h = nullptr;
    Comm(Context &c, const Ranks &rk, size_t seg_cap) {
        uint8_t handle[64];
        check(fries_comm_create(c.h, rk.n, rk.rank, &h, handle));
        std::vector<uint8_t> all = rk.allgather("comm", handle, 64);
        check(fries_comm_connect(h, all.data()));
        rk.barrier("comm_ok");
        check(fries_comm_route_create(h, seg_cap, handle));
        all = rk.allgather("route", handle, 64);
        check(fries_comm_route_connect(h, all.data()));
        rk.barrier("route_ok");
    }
    ~Comm() {
        if (h) fries_comm_destroy(h);
    }
    Comm(const Comm &) = delete;
};

struct Molecule {
    fries_mol *h = nullptr;
    unsigned n_orb, n_elec_total, n_frz;
    Molecule(Context &c, const MolInput &m) : n_orb(m.n_orb), n_elec_total(m.n_elec), n_frz(m.n_frz) {
        check(fries_mol_create(c.h, m.n_orb, m.n_elec, m.n_frz, m.hcore.data(), m.eris_packed.data(), m.symm.data(), &h));
    }
    ~Molecule() { fries_mol_destroy(h); }
    double diag_matrel(uint64_t det) {  // molecule.cpp:983-1029
        double out;
        check(fries_mol_diag(h, &det, 1, &out));
        return out;
    }
    size_t count_doub_ex(uint64_t det) {  // doub_ex_symm molecule.cpp:108-175 (count only)
        uint64_t off[2];
        check(fries_mol_doub_ex(h, &det, 1, off, nullptr, 0));
        return off[1];
    }
    size_t count_singex(uint64_t det) {  // molecule.cpp:914-933
        uint64_t off[2];
        check(fries_mol_sing_ex(h, &det, 1, off, nullptr, 0));
        return off[1];
    }
    // sing_ex_symm molecule.cpp:178-203 / doub_ex_symm :108-175 on the reference's byte strings: every symmetry-allowed
    // excitation of `det` in the reference's order, rows {occ, virt} / {occ, occ, virt, virt}; returns their number
    size_t sing_ex_symm(const uint8_t *det, uint8_t (*res_arr)[2], size_t cap) {
        uint64_t key = key_from_bytes(det, ceiling(2 * (size_t)n_orb, 8)), off[2];
        check(fries_mol_sing_ex(h, &key, 1, off, (uint8_t *)res_arr, cap));
        return off[1];
    }
    size_t doub_ex_symm(const uint8_t *det, uint8_t (*res_arr)[4], size_t cap) {
        uint64_t key = key_from_bytes(det, ceiling(2 * (size_t)n_orb, 8)), off[2];
        check(fries_mol_doub_ex(h, &key, 1, off, (uint8_t *)res_arr, cap));
        return off[1];
    }
    // sing_matr_el_nosgn molecule.cpp:76-105 (needs the determinant for the sum over occupied orbitals),
    // doub_matr_el_nosgn :26-42
    double sing_matr_el_nosgn(const uint8_t *ex_orbs, const uint8_t *det) {
        uint64_t key = key_from_bytes(det, ceiling(2 * (size_t)n_orb, 8));
        double out;
        check(fries_mol_sing_el(h, &key, ex_orbs, 1, &out));
        return out;
    }
    double doub_matr_el_nosgn(const uint8_t *ex_orbs) {
        double out;
        check(fries_mol_doub_el(h, ex_orbs, 1, &out));
        return out;
    }
    double diag_matrel(const uint8_t *det) { return diag_matrel(key_from_bytes(det, ceiling(2 * (size_t)n_orb, 8))); }
    // calc_o1_probs / calc_o2_probs / calc_o2_probs_half / calc_u1_probs / calc_u2_probs / calc_u2_probs_half
    // (heat_bathPP.cpp:182-412) with the reference's arguments: the normalised row goes to prob_arr, the un-normalised
    // total is returned (fries_mol_hb_rows; the determinant is rebuilt from the occupied list)
  private:
    double hb_row(int which, uint64_t key, int a0, int a1, int a2, double *prob_arr, uint16_t *prob_len) {
        double rows[FRIES_MAX_SUB], norm;
        int32_t args[4] = {a0, a1, a2, 0}, len;
        check(fries_mol_hb_rows(h, which, &key, args, 1, rows, &len, &norm));
        for (int32_t j = 0; j < len; j++) prob_arr[j] = rows[j];
        if (prob_len) *prob_len = (uint16_t)len;
        return norm;
    }
    static uint64_t key_of_occ(const uint8_t *occ_orbs, unsigned n) {
        uint64_t k = 0;
        for (unsigned i = 0; i < n; i++) k |= 1ull << occ_orbs[i];
        return k;
    }

  public:
    double calc_o1_probs(double *prob_arr, unsigned n_elec, const uint8_t *occ_orbs, int exclude_first) {
        return hb_row(0, key_of_occ(occ_orbs, n_elec), exclude_first, 0, 0, prob_arr, nullptr);
    }
    double calc_o2_probs(double *prob_arr, unsigned n_elec, const uint8_t *occ_orbs, uint8_t o1_idx) {
        return hb_row(1, key_of_occ(occ_orbs, n_elec), o1_idx, 0, 0, prob_arr, nullptr);
    }
    double calc_o2_probs_half(double *prob_arr, unsigned n_elec, const uint8_t *occ_orbs, uint8_t o1_idx) {
        return hb_row(2, key_of_occ(occ_orbs, n_elec), o1_idx, 0, 0, prob_arr, nullptr);
    }
    double calc_u1_probs(double *prob_arr, uint8_t o1_orb, const uint8_t *occ_orbs, uint8_t n_elec, int exclude_first) {
        return hb_row(3, key_of_occ(occ_orbs, n_elec), o1_orb, exclude_first, 0, prob_arr, nullptr);
    }
    double calc_u2_probs(double *prob_arr, uint8_t o1_orb, uint8_t o2_orb, uint8_t u1_orb, uint16_t *prob_len) {
        return hb_row(4, gen_hf_bitstring(n_orb, n_elec_total - n_frz), o1_orb, o2_orb, u1_orb, prob_arr, prob_len);
    }
    double calc_u2_probs_half(double *prob_arr, uint8_t o1_orb, uint8_t o2_orb, uint8_t u1_orb, const uint8_t *det,
                              uint16_t *prob_len) {
        return hb_row(5, key_from_bytes(det, ceiling(2 * (size_t)n_orb, 8)), o1_orb, o2_orb, u1_orb, prob_arr, prob_len);
    }
    // calc_unnorm_wt heat_bathPP.cpp:414-439 / calc_norm_wt :442-598: total HB-PP sampling weight of a double excitation
    double calc_unnorm_wt(const uint8_t *orbs) {
        uint64_t key = gen_hf_bitstring(n_orb, n_elec_total - n_frz);  // not used by the un-normalised weight; must be valid
        double out;
        check(fries_mol_hb_wt(h, 0, &key, orbs, 1, &out));
        return out;
    }
    double calc_norm_wt(const uint8_t *orbs, const uint8_t *det) {
        uint64_t key = key_from_bytes(det, ceiling(2 * (size_t)n_orb, 8));
        double out;
        check(fries_mol_hb_wt(h, 1, &key, orbs, 1, &out));
        return out;
    }
};

// DistVec<double> FRIES/vec_utils.hpp:121-953, single rank, resident on the GPU.
class DistVec {
  public:
    fries_vec *h = nullptr;
    fries_hbpp *hb = nullptr;
    unsigned n_bits, n_elec, n_vecs;
    size_t max_size_;
    uint8_t curr_vec_idx_ = 0;
    std::vector<uint64_t> buf_dets_;
    std::vector<double> buf_vals_;
    std::vector<uint8_t> buf_ini_;
    std::vector<uint32_t> proc_scr_, vec_scr_;  // hash.dat contents (HashTable scramblers, det_hash.hpp:41-58)
    unsigned hh_sites_ = 0, hh_ph_bits_ = 0;    // HubHolVec only
    int n_ranks_ = 1, rank_ = 0;  // owner partitioning: hash_fxn(occ; proc_scrambler) % n_ranks (vec_utils.hpp:373-379)
    Context *ctx_ = nullptr;
    DistVec(Context &c, size_t size, unsigned n_bits_, unsigned n_elec_, unsigned n_vecs_, const std::vector<uint32_t> &proc_scr,
            const std::vector<uint32_t> &vec_scr, int n_ranks = 1, int rank = 0)
        : n_bits(n_bits_), n_elec(n_elec_), n_vecs(n_vecs_), max_size_(size), proc_scr_(proc_scr), vec_scr_(vec_scr),
          n_ranks_(n_ranks), rank_(rank), ctx_(&c) {
        check(fries_vec_create(c.h, size, n_bits, n_elec, n_vecs, proc_scr.data(), vec_scr.data(), n_ranks, rank, &h));
    }
    // the elements of (dets, vals) this rank owns (DistVec::add routes by idx_to_proc; at set-up every rank holds the list)
    void owned(const std::vector<uint64_t> &dets, const std::vector<double> &vals, std::vector<uint64_t> &od,
               std::vector<double> &ov) {
        od.clear();
        ov.clear();
        if (n_ranks_ == 1) {
            od = dets;
            ov = vals;
            return;
        }
        std::vector<int32_t> own(dets.size());
        check(fries_hash_owner(ctx_->h, dets.data(), dets.size(), proc_scr_.data(), (int)n_bits, n_ranks_, nullptr, own.data()));
        for (size_t i = 0; i < dets.size(); i++)
            if (own[i] == rank_) {
                od.push_back(dets[i]);
                ov.push_back(vals[i]);
            }
    }
    // HubHolVec FRIES/hh_vec.hpp:27-29
    DistVec(Context &c, size_t size, unsigned n_sites, unsigned ph_bits, unsigned n_elec_, unsigned n_vecs_,
            const std::vector<uint32_t> &proc_scr, const std::vector<uint32_t> &vec_scr)
        : n_bits(n_sites * (2 + ph_bits)), n_elec(n_elec_), n_vecs(n_vecs_), max_size_(size), proc_scr_(proc_scr),
          vec_scr_(vec_scr), hh_sites_(n_sites), hh_ph_bits_(ph_bits) {
        check(fries_vec_create_hh(c.h, size, n_sites, ph_bits, n_elec, n_vecs, proc_scr.data(), vec_scr.data(), 1, 0, &h));
    }
    ~DistVec() {
        if (hb) fries_hbpp_destroy(hb);
        fries_vec_destroy(h);
    }
    DistVec(const DistVec &) = delete;
    size_t max_size() const { return max_size_; }
    size_t curr_size() {
        size_t n;
        check(fries_vec_curr_size(h, &n));
        return n;
    }
    // add x n + perform_add(origin) with curr_vec_idx = dest (vec_utils.hpp:418-440)
    void add(const std::vector<uint64_t> &dets_all, const std::vector<double> &vals_all, uint8_t ini_flag, unsigned origin = 0,
             unsigned dest = 0) {
        std::vector<uint64_t> dets;
        std::vector<double> vals;
        owned(dets_all, vals_all, dets, vals);
        std::vector<uint8_t> ini(dets.size(), ini_flag);
        if (!dets.empty()) check(fries_vec_add(h, dets.data(), vals.data(), ini.data(), dets.size(), origin, dest));
        host_valid_ = false;
    }
    double local_norm(unsigned row) {
        double n;
        check(fries_vec_local_norm(h, row, &n));
        return n;
    }
    double local_norm() { return local_norm(curr_vec_idx_); }  // vec_utils.hpp:683-689
    double two_norm() {                                         // :695-701
        double n;
        check(fries_vec_two_norm(h, curr_vec_idx_, &n));
        return n;
    }
    // the row that add / perform_add / local_norm / dot / zero_vec work on (vec_utils.hpp:585-599)
    void set_curr_vec_idx(uint8_t new_idx) {
        if (new_idx >= n_vecs) {
            std::stringstream error;
            error << "Argument to set_curr_vec_idx (" << (unsigned)new_idx << ") is out of bounds";
            throw std::runtime_error(error.str());
        }
        curr_vec_idx_ = new_idx;
    }
    uint8_t curr_vec_idx() const { return curr_vec_idx_; }
    size_t n_nonz() { return curr_size(); }
    // add(idx, val, ini_flag) buffers one element, perform_add(origin) sends the buffer to the store (:418-440, 957-1019)
    void add(uint64_t det, double val, uint8_t ini_flag) {
        if (val == 0) return;
        buf_dets_.push_back(det);
        buf_vals_.push_back(val);
        buf_ini_.push_back(ini_flag);
    }
    void perform_add(size_t origin = 0) {
        if (!buf_dets_.empty())
            check(fries_vec_add(h, buf_dets_.data(), buf_vals_.data(), buf_ini_.data(), buf_dets_.size(), (unsigned)origin,
                                curr_vec_idx_));
        buf_dets_.clear();
        buf_vals_.clear();
        buf_ini_.clear();
        host_valid_ = false;
    }
    // compress_vecs / compress_vecs_sys / compress_vecs_multi FRIES/vec_utils.cpp:10-127 (method 0 / 1 / 2) on rows
    // [start, end): the generator is advanced by exactly the draws the reference would have consumed
    void compress_rows(unsigned start, unsigned end, uint32_t compress_size, int method, std::mt19937 &mt_obj);
    void add_vecs(uint8_t idx1, uint8_t idx2, double c = 1.0) { host_valid_ = false; check(fries_vec_row_op(h, 0, idx1, idx2, c)); }   // :547-557
    void copy_vec(uint8_t src, uint8_t dst) { host_valid_ = false; check(fries_vec_row_op(h, 1, dst, src, 0.0)); }                   // :561-565
    void weight_vec(uint8_t idx1, uint8_t idx2, double expo) { host_valid_ = false; check(fries_vec_row_op(h, 2, idx1, idx2, expo)); }  // :569-573
    void zero_vec() { host_valid_ = false; check(fries_vec_row_op(h, 3, curr_vec_idx_, curr_vec_idx_, 0.0)); }                       // :577-579
    // dot with a (replicated) list of determinants (:228-253)
    double dot(const std::vector<uint64_t> &dets, const std::vector<double> &vals) {
        double out;
        check(fries_vec_dot(h, dets.data(), vals.data(), dets.size(), curr_vec_idx_, &out));
        return out;
    }
    // del_at_pos for every flagged position + cleanup (:458-497): elements that are zero in every row disappear
    void del_at_pos(const std::vector<bool> &flags) {
        std::vector<uint8_t> f(flags.begin(), flags.end());
        check(fries_vec_del(h, f.data(), f.size()));
        host_valid_ = false;
    }
    void cleanup() {
        std::vector<uint8_t> f(curr_size(), 1);
        check(fries_vec_del(h, f.data(), f.size()));
        host_valid_ = false;
    }
    // ---- the reference's raw-pointer view of the storage ------------------------------------------------------------------
    // values() / indices() / occ_orbs() / operator[] / operator() / orbs_at_pos hand out pointers into the reference's own
    // arrays (vec_utils.hpp:505-516,643-663).  Here the storage is resident in HBM, so the pointers go into a host
    // snapshot: sync_host() takes it (one device -> host copy of keys and values), push_host() writes the snapshot's
    // values back for callers that modified them through the pointers (elements that became zero in every row are dropped,
    // as by del_at_pos).  Any call that changes the store on the device (perform_add, the iterate calls, del_at_pos,
    // row operations) invalidates the snapshot; the accessors then take a new one.  Callers that drive the store through
    // the C-ABI directly (the iterate calls on `h`) call invalidate_host() themselves.
  private:
    std::vector<uint64_t> host_keys_;
    std::vector<double> host_vals_;  // n_vecs rows, row stride host_keys_.size()
    Matrix<uint8_t> host_idx_, host_occ_;
    std::vector<double> host_matr_el_;
    bool host_valid_ = false;
    std::function<double(const uint8_t *)> diag_calc_;

  public:
    void invalidate_host() { host_valid_ = false; }
    void sync_host() {
        download(host_keys_, host_vals_);
        const size_t n = host_keys_.size(), nb = ceiling(n_bits, 8);
        host_idx_.reshape(n ? n : 1, nb);
        host_occ_.reshape(n ? n : 1, n_elec);
        for (size_t i = 0; i < n; i++) {
            key_to_bytes(host_keys_[i], host_idx_[i], nb);
            gen_orb_list(host_idx_[i], host_occ_[i]);
        }
        host_matr_el_.assign(n, std::nan(""));
        host_valid_ = true;
    }
    void push_host() {
        if (!host_valid_) throw std::runtime_error("DistVec::push_host without a host snapshot (call sync_host first)");
        upload(host_keys_, host_vals_);
        host_valid_ = false;
    }
    double *values() {  // :505-507: the current row
        if (!host_valid_) sync_host();
        return host_vals_.data() + (size_t)curr_vec_idx_ * host_keys_.size();
    }
    uint8_t num_vecs() const { return (uint8_t)n_vecs; }
    Matrix<uint8_t> &indices() {  // :513-515: curr_size x n_bytes bit strings
        if (!host_valid_) sync_host();
        return host_idx_;
    }
    Matrix<uint8_t> &occ_orbs() {  // :661-663
        if (!host_valid_) sync_host();
        return host_occ_;
    }
    uint8_t *orbs_at_pos(size_t pos) { return occ_orbs()[pos]; }  // :657-659
    double *operator[](size_t pos) { return values() + pos; }     // :643-645
    double *operator()(size_t vec_idx, size_t pos) {               // :647-649
        if (!host_valid_) sync_host();
        return host_vals_.data() + vec_idx * host_keys_.size() + pos;
    }
    // the diagonal matrix element function the reference's constructor takes (vec_utils.hpp:154-198)
    void set_diag_calc(std::function<double(const uint8_t *)> f) { diag_calc_ = std::move(f); }
    double matr_el_at_pos(size_t pos) {  // :672-677, cached per snapshot
        if (!host_valid_) sync_host();
        if (!diag_calc_) throw std::runtime_error("DistVec::matr_el_at_pos: no diagonal matrix element function (set_diag_calc)");
        if (std::isnan(host_matr_el_[pos])) host_matr_el_[pos] = diag_calc_(host_occ_[pos]);
        return host_matr_el_[pos];
    }
    double internal_dot(uint8_t idx1, uint8_t idx2) {  // :324-340
        if (idx1 >= n_vecs || idx2 >= n_vecs) {
            std::stringstream error;
            error << "Error: idx" << (idx1 >= n_vecs ? 1 : 2) << " argument to internal_dot ("
                  << (unsigned)(idx1 >= n_vecs ? idx1 : idx2) << ") exceeds bounds of value matrix (" << n_vecs << ")";
            throw std::runtime_error(error.str());
        }
        if (!host_valid_) sync_host();
        const size_t n = host_keys_.size();
        double dprod = 0;
        for (size_t i = 0; i < n; i++) dprod += host_vals_[idx1 * n + i] * host_vals_[idx2 * n + i];
        return dprod;
    }
    double dense_norm() {  // :903-917 (single rank)
        if (n_dense == 0) return 0;
        const double *v = values();
        double result = 0;
        for (size_t i = 0; i < n_dense && i < host_keys_.size(); i++) result += std::fabs(v[i]);
        return result;
    }
    uint8_t gen_orb_list(const uint8_t *det, uint8_t *occ) {  // :212-214 (HubHolVec: electrons only, hh_vec.hpp:43-45)
        if (hh_sites_) {
            uint8_t tmp[64];
            uint8_t n = find_bits(det, tmp, (uint8_t)ceiling(n_bits, 8)), k = 0;
            for (uint8_t i = 0; i < n; i++)
                if (tmp[i] < 2 * hh_sites_) occ[k++] = tmp[i];
            return k;
        }
        return find_bits(det, occ, (uint8_t)ceiling(n_bits, 8));
    }
    uintmax_t idx_to_hash(const uint8_t *idx, uint8_t *orbs) {  // :389-400
        if (hh_sites_) throw std::runtime_error("DistVec::idx_to_hash: HubHolVec hashes live on the device (fries_hash_owner)");
        if (gen_orb_list(idx, orbs) != n_elec) {
            std::stringstream error;
            char det_txt[2 * 8 + 1];
            print_str(idx, (uint8_t)ceiling(n_bits, 8), det_txt);
            error << "Determinant " << det_txt;
            error << " created with an incorrect number of electrons";
            throw std::runtime_error(error.str());
        }
        return hash_fxn(vec_scr_.data(), orbs, n_elec);
    }
    int idx_to_proc(const uint8_t *idx, const uint8_t *orbs, int n_procs = 1) {  // :373-379
        (void)idx;
        return (int)(hash_fxn(proc_scr_.data(), orbs, n_elec) % (uintmax_t)n_procs);
    }
    int idx_to_proc(const uint8_t *idx, int n_procs = 1) {  // :360-365
        uint8_t orbs[64];
        gen_orb_list(idx, orbs);
        return idx_to_proc(idx, orbs, n_procs);
    }
    // pointer forms of add (:418-436) and dot (:242-253)
    bool add(const uint8_t *idx, double val, uint8_t ini_flag) {
        add(key_from_bytes(idx, ceiling(n_bits, 8)), val, ini_flag);
        return true;  // the buffer grows on demand: never full
    }
    bool add(const uint8_t *idx, const uint8_t *, double val, uint8_t ini_flag) { return add(idx, val, ini_flag); }
    double dot(Matrix<uint8_t> &idx2, const double *vals2, size_t num2) {
        std::vector<uint64_t> keys(num2);
        for (size_t i = 0; i < num2; i++) keys[i] = key_from_bytes(idx2[i], idx2.cols());
        double out;
        check(fries_vec_dot(h, keys.data(), vals2, num2, curr_vec_idx_, &out));
        return out;
    }
    double dot(Matrix<uint8_t> &idx2, const double *vals2, size_t num2, const uintmax_t *) { return dot(idx2, vals2, num2); }
    // add_elements :606-641: `count` received elements, bit n_bits of an index = initiator flag (Adder::add :965)
    void add_elements(const uint8_t *indices, const double *vals, size_t count, size_t origin) {
        const size_t nb = ceiling(n_bits + 1, 8);
        std::vector<uint64_t> keys(count);
        std::vector<uint8_t> ini(count);
        for (size_t i = 0; i < count; i++) {
            uint64_t k = key_from_bytes(indices + i * nb, nb);
            ini[i] = (uint8_t)(k >> n_bits & 1);
            keys[i] = k & ~(1ull << n_bits);
        }
        check(fries_vec_add(h, keys.data(), vals, ini.data(), count, (unsigned)origin, curr_vec_idx_));
        host_valid_ = false;
    }
    size_t adder_size() const { return max_size_; }  // :523-525: add() never reports a full buffer here
    void set_min_del_idx(size_t idx) { check(fries_vec_set_min_del_idx(h, idx)); }  // :501-503
    void fix_min_del_idx() { set_min_del_idx(curr_size()); }                         // :497-499
    // :343-353 doubles the arrays; the device store is allocated once (size it with --max_dets)
    void expand() { throw std::runtime_error("DistVec::expand: the device store has a fixed capacity (max_size)"); }
    void collect_procs() {}  // :920-952: single rank, nothing to gather
    void print_ht() {         // :405-407: the device index is one open-addressing table, not per-bucket chains
        std::cout << "hash index: " << curr_size() << " of " << max_size_ << " elements stored\n";
    }
    uint64_t tot_sgn_coh() {
        uint64_t n;
        check(fries_vec_nonini_occ_add(h, &n));
        return n;
    }
    void download(std::vector<uint64_t> &dets, std::vector<double> &vals) {
        size_t n = curr_size(), m;
        dets.resize(n ? n : 1);
        vals.resize((n ? n : 1) * n_vecs);
        check(fries_vec_download(h, dets.data(), vals.data(), n ? n : 1, &m));
        dets.resize(n);
        vals.resize(n * n_vecs);
    }
    void upload(const std::vector<uint64_t> &dets, const std::vector<double> &vals) {
        check(fries_vec_upload(h, dets.data(), vals.data(), dets.size()));
        host_valid_ = false;
    }
    // DistVec::save vec_utils.hpp:721-745: dets<rank>.dat = curr_size x n_bytes, vals<rank>.dat = n_vecs rows, dense.txt
    void save(const std::string &path) {
        std::vector<uint64_t> dets;
        std::vector<double> vals;
        download(dets, vals);
        size_t nb = ceiling(n_bits, 8);
        std::vector<uint8_t> bytes(dets.size() * nb);
        for (size_t i = 0; i < dets.size(); i++) key_to_bytes(dets[i], &bytes[i * nb], nb);
        const std::string r = std::to_string(rank_);  // one pair of files per rank, as the reference (:714-733)
        std::ofstream fd(path + "dets" + r + ".dat", std::ios::binary);
        fd.write((const char *)bytes.data(), bytes.size());
        std::ofstream fv(path + "vals" + r + ".dat", std::ios::binary);
        fv.write((const char *)vals.data(), vals.size() * sizeof(double));
        if (rank_ == 0) {  // one entry per rank (:735-745): the sizes fixed by init_dense / load
            std::ofstream fz(path + "dense.txt");
            if (n_ranks_ == 1) {
                fz << n_dense << "," << '\n';
            } else {
                for (int r = 0; r < n_ranks_; r++)
                    fz << (r < (int)dense_sizes.size() ? dense_sizes[r] : 0) << (r + 1 < n_ranks_ ? "," : "\n");
            }
        }
    }
    // DistVec::init_dense vec_utils.hpp:858-897 (single rank): the determinants of the file (read_dets io_utils.cpp:565-586,
    // one integer per determinant) become the first stored elements, with value 0, and are never deleted
    size_t n_dense = 0;        // this rank's share of the dense subspace
    size_t n_dense_total = 0;  // over all ranks
    std::vector<long long> dense_sizes;  // one per rank
    // Several ranks (rk given): every rank reads the file and keeps the determinants it owns -- the reference reads on rank 0
    // and routes them through add / perform_add, with the same result: each rank's share comes first in its storage
    // (:866-873) and dense.txt holds one size per rank (:876-893).
    size_t init_dense(const std::string &read_path, const std::string &save_dir, const Ranks *rk = nullptr) {
        std::ifstream fdets(read_path);
        if (!fdets.is_open()) throw std::runtime_error("Could not open file: " + read_path);
        std::vector<uint64_t> all_dets, dets;
        long long in_det;
        while (fdets >> in_det) all_dets.push_back((uint64_t)in_det);
        std::vector<double> none(all_dets.size(), 0.0), dummy;
        owned(all_dets, none, dets, dummy);
        std::vector<double> zeros(dets.size() * n_vecs, 0.0);
        check(fries_vec_set_min_del_idx(h, dets.size()));  // an upload drops all-zero elements beyond this index
        upload(dets, zeros);
        n_dense = dets.size();
        check(fries_vec_set_dense(h, n_dense));
        std::vector<long long> sizes{(long long)n_dense};
        if (rk && rk->n > 1) {
            std::vector<uint8_t> raw = rk->allgather("dense_sizes", &sizes[0], sizeof(long long));
            sizes.assign((const long long *)raw.data(), (const long long *)raw.data() + rk->n);
        }
        n_dense_total = 0;
        for (long long z : sizes) n_dense_total += (size_t)z;
        dense_sizes = sizes;
        check(fries_vec_set_dense_total(h, n_dense_total));
        if (!rk || rk->rank == 0) {
            std::ofstream dense_f(save_dir + "dense.txt");
            if (!dense_f.is_open())
                throw std::runtime_error("Could not load deterministic subspace from file at path " + save_dir + "dense.txt");
            for (long long z : sizes) dense_f << z << ",";
            dense_f << '\n';
        }
        return n_dense;
    }
    // DistVec::load vec_utils.hpp:761-844 (single rank): re-hash, drop |v| <= 1e-9
    void load(const std::string &path) {
        const std::string rs = std::to_string(rank_);  // a checkpoint is restarted on the rank count it was saved with
        std::ifstream fd(path + "dets" + rs + ".dat", std::ios::binary);
        if (!fd.is_open()) throw std::runtime_error("Error: could not open saved binary vector file at " + path + "dets" + rs + ".dat");
        std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(fd)), std::istreambuf_iterator<char>());
        size_t nb = ceiling(n_bits, 8), n = bytes.size() / nb;
        std::ifstream fv(path + "vals" + rs + ".dat", std::ios::binary);
        if (!fv.is_open()) throw std::runtime_error("Error: could not open saved binary vector file at " + path + "vals" + rs + ".dat");
        std::vector<double> all(n * n_vecs, 0.0);
        fv.read((char *)all.data(), all.size() * sizeof(double));
        {   // sizes of the dense subspaces, one per rank ("a, b, c, "): the first n_dense entries are kept as they are
            std::ifstream fz(path + "dense.txt");
            n_dense = 0;
            n_dense_total = 0;
            dense_sizes.clear();
            if (fz.is_open()) {
                std::string tok;
                for (int r = 0; std::getline(fz, tok, ','); r++) {
                    size_t start = tok.find_first_not_of(" \t\r\n");
                    if (start == std::string::npos) break;
                    const size_t z = (size_t)std::stoull(tok.substr(start));
                    if (r == rank_) n_dense = z;
                    n_dense_total += z;
                    dense_sizes.push_back((long long)z);
                }
            }
            if (n_dense > n) n_dense = n;
        }
        std::vector<uint64_t> dets;
        std::vector<double> rows[8];
        for (size_t i = 0; i < n; i++) {
            bool keep = i < n_dense;
            for (unsigned r = 0; r < n_vecs; r++) keep = keep || std::fabs(all[r * n + i]) > 1e-9;
            if (!keep) continue;
            dets.push_back(key_from_bytes(&bytes[i * nb], nb));
            for (unsigned r = 0; r < n_vecs; r++) rows[r].push_back(all[r * n + i]);
        }
        std::vector<double> vals;
        for (unsigned r = 0; r < n_vecs; r++) vals.insert(vals.end(), rows[r].begin(), rows[r].end());
        check(fries_vec_set_min_del_idx(h, n_dense));
        upload(dets, vals);
        if (n_dense) check(fries_vec_set_dense(h, n_dense));
        check(fries_vec_set_dense_total(h, n_dense_total));
    }
};

// ---- compress_utils.hpp mirrors on host arrays -----------------------------------------------------------------------------
// find_preserve compress_utils.cpp:29-105 (srt_idx is scratch in the reference; unused here)
inline double find_preserve(Context &c, double *values, std::vector<size_t> &, std::vector<bool> &keep_idx, size_t count,
                            unsigned *n_samp, double *global_norm) {
    std::vector<uint8_t> keep(count ? count : 1);
    double loc;
    check(fries_find_preserve(c.h, values, count, n_samp, global_norm, keep.data(), &loc));
    for (size_t i = 0; i < count; i++) keep_idx[i] = keep[i];
    return loc;
}
// sys_comp compress_utils.cpp:278-327
inline void sys_comp(Context &c, double *vec_vals, size_t vec_len, double *loc_norms, unsigned n_samp,
                     std::vector<bool> &keep_exact, double rand_num) {
    std::vector<uint8_t> keep(vec_len ? vec_len : 1);
    for (size_t i = 0; i < vec_len; i++) keep[i] = keep_exact[i];
    check(fries_sys_comp(c.h, vec_vals, vec_len, loc_norms, 1, 0, n_samp, keep.data(), rand_num));
    for (size_t i = 0; i < vec_len; i++) keep_exact[i] = keep[i];
}
// ---- pivotal family compress_utils.cpp:354-681.  The reference's functions draw from the caller's std::mt19937; the C-ABI
// takes the generator's next raw outputs and reports how many it used, and the generator is advanced by exactly that.
namespace detail {
inline std::vector<uint32_t> peek(const std::mt19937 &mt, size_t n) {
    std::mt19937 copy = mt;
    std::vector<uint32_t> d(n ? n : 1);
    for (size_t i = 0; i < n; i++) d[i] = (uint32_t)copy();
    return d;
}
}  // namespace detail
inline void DistVec::compress_rows(unsigned start, unsigned end, uint32_t compress_size, int method, std::mt19937 &mt_obj) {
    const size_t rows = end > start ? end - start : 0;
    const size_t per_row = method == 0 ? 2 * (size_t)compress_size + 2 : method == 1 ? 1 : 4 * (size_t)compress_size;
    std::vector<uint32_t> draws = detail::peek(mt_obj, rows * per_row);
    size_t used = 0;
    host_valid_ = false;
    check(fries_vec_compress(h, start, end, compress_size, method, draws.data(), rows * per_row, &used));
    mt_obj.discard(used);
}
// piv_samp_serial compress_utils.cpp:390-530
inline void piv_samp_serial(Context &c, double *vec_vals, size_t vec_len, double seg_norm, uint32_t n_samp,
                            std::vector<bool> &keep_exact, std::mt19937 &mt_obj) {
    std::vector<uint8_t> keep(vec_len ? vec_len : 1);
    for (size_t i = 0; i < vec_len; i++) keep[i] = keep_exact[i];
    std::vector<uint32_t> draws = detail::peek(mt_obj, 2 * (size_t)n_samp);
    size_t used = 0;
    check(fries_piv_samp_serial(c.h, vec_vals, vec_len, seg_norm, n_samp, keep.data(), draws.data(), &used));
    mt_obj.discard(used);
    for (size_t i = 0; i < vec_len; i++) keep_exact[i] = keep[i];
}
// piv_budget compress_utils.cpp:560-608: budgets of all ranks (the reference returns the caller's after an MPI_Scatter)
inline std::vector<uint32_t> piv_budget(const double *loc_norms, int n_ranks, uint32_t n_samp, std::mt19937 &mt_obj) {
    std::vector<uint32_t> budgets(n_ranks), draws = detail::peek(mt_obj, 2 * (size_t)n_ranks);
    size_t used = 0;
    check(fries_piv_budget(loc_norms, n_ranks, n_samp, draws.data(), &used, budgets.data()));
    mt_obj.discard(used);
    return budgets;
}
// adjust_probs compress_utils.cpp:617-681
inline double adjust_probs(Context &c, double *vec_vals, size_t vec_len, uint32_t *n_samp_loc, double exp_nsamp_loc,
                           uint32_t n_samp_tot, double tot_norm, std::vector<bool> &keep_exact) {
    std::vector<uint8_t> keep(vec_len ? vec_len : 1);
    for (size_t i = 0; i < vec_len; i++) keep[i] = keep_exact[i];
    double norm = 0;
    check(fries_adjust_probs(c.h, vec_vals, vec_len, n_samp_loc, exp_nsamp_loc, n_samp_tot, tot_norm, keep.data(), &norm));
    for (size_t i = 0; i < vec_len; i++) keep_exact[i] = keep[i];
    return norm;
}
// piv_comp_parallel compress_utils.cpp:354-387 (single rank; srt_scratch is unused here)
inline void piv_comp_parallel(Context &c, double *vec_vals, size_t vec_len, uint32_t compress_size, std::vector<size_t> &,
                              std::vector<bool> &keep_scratch, std::mt19937 &rn_gen) {
    std::vector<uint8_t> keep(vec_len ? vec_len : 1);
    std::vector<uint32_t> draws = detail::peek(rn_gen, 2 * ((size_t)compress_size + 1));
    size_t used = 0;
    check(fries_piv_comp(c.h, vec_vals, vec_len, compress_size, keep.data(), draws.data(), &used, nullptr, 1, 0, 0, 0));
    rn_gen.discard(used);
    for (size_t i = 0; i < vec_len; i++) keep_scratch[i] = keep[i];
}
// comp_sub compress_utils.cpp:797-820 (find_keep_sub + sys_sub; keep_idx and wt_remain are scratch in the reference and
// unused here).  new_vals / new_idx must hold n_samp entries, as there.
inline size_t comp_sub(Context &c, double *values, size_t count, unsigned int *n_div, Matrix<double> &sub_weights,
                       Matrix<bool> &, uint16_t *sub_sizes, unsigned int n_samp, double *, double rand_num, double *new_vals,
                       size_t new_idx[][2]) {
    static_assert(sizeof(size_t) == sizeof(uint64_t), "new_idx is passed through as uint64_t pairs");
    size_t n_out = 0;
    check(fries_comp_sub(c.h, values, count, n_div, sub_weights.data(), sub_weights.cols(), sub_sizes, n_samp, rand_num, new_vals,
                         (uint64_t *)new_idx, n_samp, &n_out, nullptr, nullptr));
    return n_out;
}
// compress_vecs / compress_vecs_sys FRIES/vec_utils.cpp:10-70 (the three scratch arrays are unused here): rows
// [start_idx, end_idx) of the resident store, pivotal / systematic; the generator advances as in the reference
inline void compress_vecs_impl(DistVec &vectors, size_t start_idx, size_t end_idx, unsigned int compress_size, int method,
                               std::mt19937 &rn_gen) {
    const size_t rows = end_idx > start_idx ? end_idx - start_idx : 0;
    std::vector<uint32_t> draws = detail::peek(rn_gen, rows * (method == 0 ? 2 * ((size_t)compress_size + 1) : 1) + 2);
    size_t used = 0;
    check(fries_vec_compress(vectors.h, (unsigned)start_idx, (unsigned)end_idx, compress_size, method, draws.data(), draws.size(),
                             &used));
    rn_gen.discard(used);
    vectors.invalidate_host();
}
inline void compress_vecs(DistVec &vectors, size_t start_idx, size_t end_idx, unsigned int compress_size, std::vector<size_t> &,
                          std::vector<bool> &, std::vector<bool> &, std::mt19937 &rn_gen) {
    compress_vecs_impl(vectors, start_idx, end_idx, compress_size, 0, rn_gen);
}
inline void compress_vecs_sys(DistVec &vectors, size_t start_idx, size_t end_idx, unsigned int compress_size,
                              std::vector<size_t> &, std::vector<bool> &, std::vector<bool> &, std::mt19937 &rn_gen) {
    compress_vecs_impl(vectors, start_idx, end_idx, compress_size, 1, rn_gen);
}

// ---- heat_bathPP.hpp: HBCompress* (:247-310) and apply_HBPP_sys (:686-992) / apply_HBPP_piv (:1014-1419) ---------------
// The caller-visible fields of the reference's structs; the sub-weight matrices and scratch arrays the reference keeps in
// them live on the device here.
struct HBCompress {
    std::vector<double> vec1;                       // values before and after compression
    size_t vec_len = 0;                             // number of elements before and after
    std::vector<size_t> det_indices1, det_indices2;  // row of all_dets: of each input / of each sample
    uint8_t (*orb_indices1)[4];                     // per sample: {occ, occ, virt, virt}, or {occ, virt, 0, 0} for a single
    explicit HBCompress(size_t length) : vec1(length), det_indices1(length), det_indices2(length), orb_store_(4 * length) {
        orb_indices1 = (uint8_t(*)[4])orb_store_.data();
    }

  private:
    std::vector<uint8_t> orb_store_;
};
struct HBCompressSys : HBCompress {
    HBCompressSys(size_t length, size_t) : HBCompress(length) {}
};
struct HBCompressPiv : HBCompress {
    HBCompressPiv(size_t length, size_t) : HBCompress(length) {}
};
namespace detail {
inline std::vector<uint64_t> hb_keys(Matrix<uint8_t> &all_dets, const HBCompress &cs) {
    std::vector<uint64_t> keys(cs.vec_len ? cs.vec_len : 1);
    for (size_t i = 0; i < cs.vec_len; i++) keys[i] = key_from_bytes(all_dets[cs.det_indices1[i]], all_dets.cols());
    return keys;
}
inline void hb_store(HBCompress &cs, const std::vector<double> &val, const std::vector<uint64_t> &det,
                     const std::vector<uint8_t> &orbs, size_t n_out) {
    std::vector<size_t> parent(n_out);
    for (size_t k = 0; k < n_out; k++) parent[k] = cs.det_indices1[det[k]];  // input position -> row of all_dets
    for (size_t k = 0; k < n_out; k++) {
        cs.vec1[k] = val[k];
        cs.det_indices2[k] = parent[k];
        memcpy(cs.orb_indices1[k], &orbs[4 * k], 4);
    }
    cs.vec_len = n_out;
}
}  // namespace detail
// inputs: comp_scratch->vec1[0 .. vec_len), det_indices1; outputs: vec1, det_indices2, orb_indices1, vec_len.  One uniform
// per compression stage is drawn from mt_obj, as in the reference (:729,765,811,859,910).
inline void apply_HBPP_sys(Molecule &mol, Matrix<uint8_t> &all_dets, HBCompressSys *comp_scratch, double p_doub, bool new_hb,
                           std::mt19937 &mt_obj, uint32_t n_samp) {
    HBCompress &cs = *comp_scratch;
    double u[5];
    for (double &x : u) x = mt_obj() / (1. + UINT32_MAX);
    std::vector<uint64_t> keys = detail::hb_keys(all_dets, cs), det(cs.vec1.size());
    std::vector<double> val(cs.vec1.size());
    std::vector<uint8_t> orbs(4 * cs.vec1.size());
    size_t n_out = 0;
    check(fries_apply_hbpp_sys(mol.h, keys.data(), cs.vec1.data(), cs.vec_len, p_doub, new_hb, u, n_samp, cs.vec1.size(),
                               val.data(), det.data(), orbs.data(), cs.vec1.size(), &n_out));
    detail::hb_store(cs, val, det, orbs, n_out);
}
// spin_parity = 0 only (the time-reversal-symmetric variant serves drivers outside the scope)
inline void apply_HBPP_piv(Molecule &mol, Matrix<uint8_t> &all_dets, HBCompressPiv *comp_scratch, double p_doub, bool new_hb,
                           std::mt19937 &mt_obj, uint32_t n_samp, int spin_parity = 0) {
    if (spin_parity != 0) throw std::runtime_error("apply_HBPP_piv: spin_parity != 0 is not supported");
    HBCompress &cs = *comp_scratch;
    std::vector<uint32_t> draws = detail::peek(mt_obj, 10 * ((size_t)n_samp + 1) + 64);  // five compressions
    std::vector<uint64_t> keys = detail::hb_keys(all_dets, cs), det(cs.vec1.size());
    std::vector<double> val(cs.vec1.size());
    std::vector<uint8_t> orbs(4 * cs.vec1.size());
    size_t n_out = 0, used = 0;
    check(fries_apply_hbpp_piv(mol.h, keys.data(), cs.vec1.data(), cs.vec_len, p_doub, new_hb, draws.data(), draws.size(), &used,
                               n_samp, cs.vec1.size(), val.data(), det.data(), orbs.data(), cs.vec1.size(), &n_out));
    mt_obj.discard(used);
    detail::hb_store(cs, val, det, orbs, n_out);
}

// seed_sys compress_utils.cpp:107-127: position of the first systematic-sampling grid point of rank my_rank, whose lower
// ranks hold the one-norms norms[0 .. my_rank); returns that lower bound (scalar control logic, as adjust_shift)
inline double seed_sys(const double *norms, double *rn, unsigned int n_samp, int n_procs = 1, int my_rank = 0) {
    double lbound = 0;
    for (int p = 0; p < my_rank; p++) lbound += norms[p];
    double global_norm = lbound;
    for (int p = my_rank; p < n_procs; p++) global_norm += norms[p];
    const double unit = global_norm / n_samp;
    *rn *= unit;
    *rn += unit * (int)(lbound * n_samp / global_norm);
    if (*rn < lbound) *rn += unit;
    return lbound;
}
// sum_mpi compress_utils.hpp:179-232: sum of one number per rank.  The host layer drives one GPU from one process (the
// multi-GPU exchange lives inside the library, csrc/comm.cuh), so the sum over ranks is the local value.
inline double sum_mpi(double local, int, int) { return local; }
inline int sum_mpi(int local, int, int) { return local; }
inline uint64_t sum_mpi(uint64_t local, int, int) { return local; }

// adjust_shift compress_utils.cpp:684-693 (scalar control logic of the drivers)
inline void adjust_shift(double *shift, double one_norm, double *last_norm, double target_norm, double damp_factor) {
    if (*last_norm) {
        *shift -= damp_factor * std::log(one_norm / *last_norm);
        *last_norm = one_norm;
    }
    if (*last_norm == 0 && one_norm > target_norm) *last_norm = one_norm;
}
// round_binomially compress_utils.cpp:19-27 (used by examples/fries_test.cpp, the reference's link smoke test)
inline int round_binomially(double p, unsigned n, std::mt19937 &mt_obj) {
    int flr = (int)std::floor(p);
    double prob = p - flr;
    int ret = flr * (int)n;
    for (unsigned i = 0; i < n; i++) ret += (mt_obj() / (1. + UINT32_MAX)) < prob;
    return ret;
}

inline unsigned seed_from_clock_or_env() {
    // the reference seeds mt19937 from the wall clock (frisys_mol.cpp:104-106); FRIES_SEED makes runs reproducible
    const char *s = std::getenv("FRIES_SEED");
    if (s) return (unsigned)std::atoll(s);
    return (unsigned)std::random_device{}();
}

// H * trial on the device, diagonal shifted by hf_en (frisys_mol.cpp:155-214)
inline void h_times(Context &c, Molecule &mol, double hf_en, const std::vector<uint64_t> &dets, const std::vector<double> &vals,
                    const std::vector<uint32_t> &proc_scr, const std::vector<uint32_t> &vec_scr, unsigned n_bits,
                    unsigned n_elec, std::vector<uint64_t> &out_dets, std::vector<double> &out_vals) {
    size_t n_ex = (size_t)mol.n_orb * mol.n_orb * n_elec * n_elec;
    DistVec tmp(c, std::max<size_t>(1 << 16, 2 * dets.size() * n_ex / 8), n_bits, n_elec, 2, proc_scr, vec_scr);
    check(fries_vec_set_diag_mol(tmp.h, mol.h, hf_en));
    tmp.add(dets, vals, 1);
    check(fries_h_apply(tmp.h, mol.h, 0, 1, 0.0, 1.0));
    std::vector<double> all;
    tmp.download(out_dets, all);
    out_vals.assign(all.begin() + out_dets.size(), all.end());
}

}  // namespace fries
