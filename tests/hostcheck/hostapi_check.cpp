// TEST INFRASTRUCTURE ONLY.  The pure-host pieces of the C++ mirror (fries_b200/host/fries_host.hpp) that need no GPU:
// Matrix<T>, find_bits, hash_fxn, key <-> byte string.  stdin: n_bits n_elec n_scr scr[...] n_keys keys[...];
// stdout: one line per key "hash n_bits_found" -- compared with the oracle by tests/test_hostapi_cpu.py.
#include "../../fries_b200/host/fries_host.hpp"
// mode "bits": stdin n_orb n_rec, then records "key o0 o1 v2 v3 a b"; stdout per record:
// bits_between(a,b) sing_sign sing_key doub_sign doub_key sing_parity doub_parity flipped_key excite_sign(a,b)
static int bits_mode() {
    unsigned n_orb;
    size_t n;
    std::cin >> n_orb >> n;
    const size_t nb = fries::ceiling(2 * n_orb, 8);
    for (size_t r = 0; r < n; r++) {
        uint64_t k;
        unsigned o[4], a, b;
        std::cin >> k >> o[0] >> o[1] >> o[2] >> o[3] >> a >> b;
        uint8_t det[8], orbs[4] = {(uint8_t)o[0], (uint8_t)o[1], (uint8_t)o[2], (uint8_t)o[3]}, so[2] = {orbs[0], orbs[2]};
        fries::key_to_bytes(k, det, 8);
        std::cout << fries::bits_between(det, (uint8_t)a, (uint8_t)b) << " ";
        uint8_t d1[8], d2[8], d3[8];
        memcpy(d1, det, 8);
        int s1 = fries::sing_det_parity(d1, so);
        memcpy(d3, det, 8);
        fries::sing_det(d3, so);
        if (memcmp(d1, d3, 8)) return 6;
        memcpy(d2, det, 8);
        int s2 = fries::doub_det_parity(d2, orbs);
        memcpy(d3, det, 8);
        fries::doub_det(d3, orbs);
        if (memcmp(d2, d3, 8)) return 7;
        uint8_t fl[8] = {0};
        fries::flip_spins(det, fl, (uint8_t)n_orb);
        for (unsigned i = 0; i < 64; i++)
            if (fries::read_bit(det, (uint8_t)i) != (int)(k >> i & 1)) return 8;
        std::cout << s1 << " " << fries::key_from_bytes(d1, 8) << " " << s2 << " " << fries::key_from_bytes(d2, 8) << " "
                  << fries::sing_parity(det, so) << " " << fries::doub_parity(det, orbs) << " " << fries::key_from_bytes(fl, nb) << " "
                  << fries::excite_sign((uint8_t)a, (uint8_t)b, det) << "\n";
    }
    char txt[17];
    uint8_t three[3] = {0x01, 0x23, 0xab};
    fries::print_str(three, 3, txt);
    if (std::string(txt) != "ab2301") return 9;
    uint8_t x[3] = {0x81, 0, 0x10}, y[3] = {0x80, 0, 0x00}, bits[4];
    if (fries::find_diff_bits(x, y, bits, 3) != 2 || bits[0] != 0 || bits[1] != 20) return 10;
    uint8_t z[3] = {0x7e, 0, 0};
    if (fries::find_diff_bits(x, z, bits, 3) != UINT8_MAX) return 11;
    return 0;
}

// mode "parse": argv = parse fcidump <path> <point group> | parse hf <dir>; prints the MolInput: one header line
// "n_orb n_elec n_frz eps hf_en", the irreps, then hcore and the packed integrals as hexadecimal floats
static int parse_mode(int argc, char **argv) {
    if (argc < 4) return 20;
    if (std::string(argv[2]) == "hh") {  // Hubbard-Holstein parameter file + the Neel determinant of its lattice
        try {
            fries::HhInput h = fries::parse_hh_input(argv[3]);
            printf("%u %u %u %a %a %a %a %a %llu\n", h.n_elec, h.lat_len, h.n_dim, h.eps, h.elec_int, h.ph_freq, h.elec_ph, h.hf_en,
                   (unsigned long long)fries::gen_neel_det_1D(h.lat_len, h.n_elec));
        } catch (std::exception &e) {
            std::cout << "Exception : " << e.what() << std::endl;
            return 21;
        }
        return 0;
    }
    fries::MolInput m;
    try {
        if (std::string(argv[2]) == "fcidump") {
            if (argc < 5) return 20;
            m = fries::parse_fcidump(argv[3], argv[4]);
        } else {
            m = fries::parse_hf_input(argv[3]);
        }
    } catch (std::exception &e) {
        std::cout << "Exception : " << e.what() << std::endl;
        return 21;
    }
    printf("%u %u %u %a %a\n", m.n_orb, m.n_elec, m.n_frz, m.eps, m.hf_en);
    for (uint8_t x : m.symm) printf("%u ", (unsigned)x);
    printf("\n");
    for (double x : m.hcore) printf("%a ", x);
    printf("\n");
    for (double x : m.eris_packed) printf("%a ", x);
    printf("\n");
    return 0;
}

// mode "files <dir>": load_vec_txt of <dir>ini_ (dets / vals text files), hash.dat round trip, load_last_line of <dir>S.txt
static int files_mode(int argc, char **argv) {
    if (argc < 3) return 30;
    const std::string dir = argv[2];
    std::vector<uint64_t> dets;
    std::vector<double> vals;
    size_t n = fries::load_vec_txt(dir + "ini_", dets, vals);
    printf("%zu\n", n);
    for (size_t i = 0; i < n; i++) printf("%llu %a\n", (unsigned long long)dets[i], vals[i]);
    std::vector<uint32_t> scr(44), back(44);
    fries::load_proc_hash(dir, scr);           // written by the test
    fries::save_proc_hash(dir + "copy_", scr);  // read back by the test
    for (uint32_t x : scr) printf("%u ", x);
    printf("\n");
    double last = 0;
    bool ok = fries::load_last_line(dir + "S.txt", &last);
    printf("%d %a\n", (int)ok, last);
    ok = fries::load_last_line(dir + "missing.txt", &last);
    printf("%d\n", (int)ok);
    return 0;
}
// mode "args ...": the drivers' command-line handling (argparse.hpp semantics)
static int args_mode(int argc, char **argv) {
    fries::Args a(argc - 1, argv + 1);
    std::string path = a.str("fcidump_path");
    double vec_nonz = a.num("vec_nonz");
    double target = a.num("target", 0);
    std::string rd = a.str("result_dir", "./");
    a.validate();
    printf("%s %g %g %s\n", path.c_str(), vec_nonz, target, rd.c_str());
    return 0;
}

// mode "seed": stdin records "n_procs rank n_samp rn norms..."; stdout "lbound rn" as hexadecimal floats
static int seed_mode() {
    int n_procs, rank;
    unsigned n_samp;
    double rn;
    while (std::cin >> n_procs >> rank >> n_samp >> rn) {
        std::vector<double> norms(n_procs);
        for (double &x : norms) std::cin >> x;
        double lb = fries::seed_sys(norms.data(), &rn, n_samp, n_procs, rank);
        printf("%a %a\n", lb, rn);
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "seed") return seed_mode();
    if (argc > 1 && std::string(argv[1]) == "files") return files_mode(argc, argv);
    if (argc > 1 && std::string(argv[1]) == "args") return args_mode(argc, argv);
    if (argc > 1 && std::string(argv[1]) == "bits") return bits_mode();
    if (argc > 1 && std::string(argv[1]) == "parse") return parse_mode(argc, argv);
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
