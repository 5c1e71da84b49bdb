// TEST INFRASTRUCTURE ONLY.  The pure-host pieces of the C++ mirror (fries_b200/host/fries_host.hpp) that need no GPU:
// Matrix<T>, find_bits, hash_fxn, key <-> byte string.  stdin: n_bits n_elec n_scr scr[...] n_keys keys[...];
// stdout: one line per key "hash n_bits_found" -- compared with the oracle by tests/test_hostapi_cpu.py.
#include "../../fries_b200/host/fries_host.hpp"
int main() {
    unsigned n_bits, n_elec, n_scr;
    size_t n_keys;
    std::cin >> n_bits >> n_elec >> n_scr;
    std::vector<uint32_t> scr(n_scr);
    for (auto &x : scr) std::cin >> x;
    std::cin >> n_keys;
    // Matrix: row-major storage, both accessors, reshape keeps the leading entries of the flat array
    fries::Matrix<uint8_t> m(3, 4);
    for (size_t r = 0; r < 3; r++)
        for (size_t c = 0; c < 4; c++) m(r, c) = (uint8_t)(10 * r + c);
    if (m[2][3] != 23 || m.data()[5] != 11 || m.rows() != 3 || m.cols() != 4) return 2;
    m.reshape(6, 2);
    if (m(2, 1) != 11 || m.rows() != 6 || m.cols() != 2) return 3;
    const size_t nb = fries::ceiling(n_bits, 8);
    for (size_t i = 0; i < n_keys; i++) {
        uint64_t k;
        std::cin >> k;
        uint8_t det[8], occ[64];
        fries::key_to_bytes(k, det, nb);
        if (fries::key_from_bytes(det, nb) != k) return 4;
        unsigned n = fries::find_bits(det, occ, (uint8_t)nb);
        for (unsigned j = 1; j < n; j++)
            if (occ[j] <= occ[j - 1]) return 5;
        std::cout << fries::hash_fxn(scr.data(), occ, n) << " " << n << "\n";
    }
    (void)n_elec;
    return 0;
}
