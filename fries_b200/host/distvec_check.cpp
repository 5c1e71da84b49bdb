// Self-check of the raw-pointer view of fries::DistVec (values / indices / occ_orbs / operator[] / operator() /
// orbs_at_pos / matr_el_at_pos / internal_dot / dense_norm / idx_to_hash / idx_to_proc / add_elements / push_host;
// FRIES/vec_utils.hpp:200-953) on the GPU: a small store is filled through add + perform_add, read back through the
// pointers and compared with the values that went in.  Exit code 0 = all checks passed; every failure prints a line.
#include "fries_host.hpp"
using namespace fries;

static int n_fail = 0;
#define CHECK(cond)                                                                    \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            std::cout << "FAILED line " << __LINE__ << ": " #cond << std::endl;        \
            n_fail++;                                                                  \
        }                                                                              \
    } while (0)

int main() {
    try {
        Context ctx(0);
        const unsigned n_orb = 10, n_elec = 6, n_bits = 2 * n_orb;
        std::mt19937 mt(7);
        std::vector<uint32_t> pscr(n_bits), vscr(n_bits);
        for (auto &x : pscr) x = mt();
        for (auto &x : vscr) x = mt();
        DistVec vec(ctx, 4096, n_bits, n_elec, 2, pscr, vscr);
        // 200 distinct determinants with 3 alpha + 3 beta electrons
        std::map<uint64_t, double> truth;
        while (truth.size() < 200) {
            uint64_t k = 0;
            for (int spin = 0; spin < 2; spin++) {
                unsigned got = 0;
                while (got < n_elec / 2) {
                    unsigned o = mt() % n_orb + spin * n_orb;
                    if (!(k >> o & 1)) {
                        k |= 1ull << o;
                        got++;
                    }
                }
            }
            if (truth.count(k)) continue;
            double v = (double)(mt() % 2000) / 100.0 - 10.0;
            if (v == 0) v = 0.5;
            truth[k] = v;
        }
        uint8_t det[8];
        for (auto &kv : truth) {
            key_to_bytes(kv.first, det, 3);
            CHECK(vec.add(det, kv.second, 1));  // pointer form
        }
        vec.perform_add(0);
        CHECK(vec.curr_size() == truth.size());
        CHECK(vec.num_vecs() == 2 && vec.max_size() == 4096 && vec.adder_size() == 4096);
        // raw view
        Matrix<uint8_t> &idx = vec.indices();
        Matrix<uint8_t> &occ = vec.occ_orbs();
        CHECK(idx.cols() == 3 && occ.cols() == n_elec);
        double *vals = vec.values();
        double norm = 0;
        for (size_t i = 0; i < vec.curr_size(); i++) {
            uint64_t k = key_from_bytes(idx[i], 3);
            CHECK(truth.count(k) == 1);
            CHECK(vals[i] == truth[k] && *vec[i] == truth[k] && *vec(0, i) == truth[k] && *vec(1, i) == 0);
            uint8_t o2[64];
            CHECK(vec.gen_orb_list(idx[i], o2) == n_elec && memcmp(o2, occ[i], n_elec) == 0 && vec.orbs_at_pos(i) == occ[i]);
            uint8_t o3[64];
            CHECK(vec.idx_to_hash(idx[i], o3) == hash_fxn(vscr.data(), occ[i], n_elec));
            CHECK(vec.idx_to_proc(idx[i], 8) == (int)(hash_fxn(pscr.data(), occ[i], n_elec) % 8));
            norm += std::fabs(vals[i]);
        }
        CHECK(std::fabs(norm - vec.local_norm()) <= 1e-12 * norm);
        // diagonal elements through the user's function, cached
        int calls = 0;
        vec.set_diag_calc([&](const uint8_t *o) {
            calls++;
            double s = 0;
            for (unsigned i = 0; i < n_elec; i++) s += o[i];
            return s;
        });
        double d5 = vec.matr_el_at_pos(5);
        CHECK(d5 == vec.matr_el_at_pos(5) && calls == 1);
        // internal_dot / two_norm
        double tn = vec.two_norm();
        CHECK(std::fabs(vec.internal_dot(0, 0) - tn) <= 1e-12 * tn && vec.internal_dot(0, 1) == 0);
        bool threw = false;
        try {
            vec.internal_dot(0, 2);
        } catch (std::runtime_error &) {
            threw = true;
        }
        CHECK(threw);
        // write through the pointers, push back: element 3 doubled, element 4 deleted (zero in every row)
        uint64_t k3 = key_from_bytes(idx[3], 3), k4 = key_from_bytes(idx[4], 3);
        *vec[3] *= 2;
        *vec[4] = 0;
        vec.push_host();
        CHECK(vec.curr_size() == truth.size() - 1);
        std::vector<uint64_t> one{k3}, gone{k4};
        std::vector<double> w{1.0};
        CHECK(vec.dot(one, w) == 2 * truth[k3] && vec.dot(gone, w) == 0);
        // Matrix form of dot
        Matrix<uint8_t> tm(2, 3);
        key_to_bytes(k3, tm[0], 3);
        key_to_bytes(k4, tm[1], 3);
        double tv[2] = {1.0, 1.0};
        CHECK(vec.dot(tm, tv, 2) == 2 * truth[k3]);
        // add_elements: a received buffer, bit n_bits = initiator flag; a flagged new determinant is created, an unflagged
        // one is dropped, an unflagged existing one accumulates (vec_utils.hpp:606-641)
        uint64_t k_new1 = 0x7ull | (0x7ull << n_orb), k_new2 = 0x38ull | (0x38ull << n_orb);
        std::vector<uint8_t> buf(3 * 3);
        std::vector<double> bv{1.25, 2.5, 0.75};
        bool new1 = !truth.count(k_new1), new2 = !truth.count(k_new2);
        key_to_bytes(k_new1 | (1ull << n_bits), &buf[0], 3);
        key_to_bytes(k_new2, &buf[3], 3);
        key_to_bytes(k3, &buf[6], 3);
        size_t before = vec.curr_size();
        vec.add_elements(buf.data(), bv.data(), 3, 0);
        CHECK(vec.curr_size() == before + (new1 ? 1 : 0));
        std::vector<uint64_t> q{k_new1, k_new2, k3};
        std::vector<double> q1{1, 0, 0}, q2{0, 1, 0}, q3{0, 0, 1};
        if (new1) CHECK(vec.dot(q, q1) == 1.25);
        if (new2) CHECK(vec.dot(q, q2) == 0);
        CHECK(vec.dot(q, q3) == 2 * truth[k3] + 0.75);
        // fix_min_del_idx + dense_norm without a dense subspace, print_ht, collect_procs, expand
        vec.fix_min_del_idx();
        vec.set_min_del_idx(0);
        CHECK(vec.dense_norm() == 0);
        vec.collect_procs();
        vec.print_ht();
        threw = false;
        try {
            vec.expand();
        } catch (std::runtime_error &) {
            threw = true;
        }
        CHECK(threw);
        // wrong electron count
        threw = false;
        uint8_t bad[3] = {1, 0, 0}, o4[64];
        try {
            vec.idx_to_hash(bad, o4);
        } catch (std::runtime_error &e) {
            threw = std::string(e.what()).find("incorrect number of electrons") != std::string::npos;
        }
        CHECK(threw);
        // ---- Molecule: the reference's molecule.hpp / heat_bathPP.hpp entry points on byte strings ----
        {
            MolInput mi;
            mi.n_orb = n_orb;
            mi.n_elec = n_elec;
            mi.n_frz = 0;
            mi.symm.resize(n_orb);
            for (unsigned i = 0; i < n_orb; i++) mi.symm[i] = (uint8_t)(i % 4);
            mi.hcore.assign((size_t)n_orb * n_orb, 0.0);
            for (unsigned i = 0; i < n_orb; i++)
                for (unsigned j = 0; j <= i; j++)
                    mi.hcore[i * n_orb + j] = mi.hcore[j * n_orb + i] = i == j ? -2.0 + 0.3 * i : 0.01 * (i + j);
            const size_t n_pair = (size_t)n_orb * (n_orb + 1) / 2;
            mi.eris_packed.assign(n_pair * (n_pair + 1) / 2, 0.0);
            for (size_t i = 0; i < mi.eris_packed.size(); i++) mi.eris_packed[i] = 0.05 + 0.001 * (double)(i % 97);
            Molecule mol(ctx, mi);
            uint8_t hf[3];
            gen_hf_bitstring(n_orb, n_elec, hf);
            std::vector<uint8_t> sing(2 * 256), doub(4 * 4096);
            size_t ns = mol.sing_ex_symm(hf, (uint8_t(*)[2])sing.data(), 256);
            size_t nd = mol.doub_ex_symm(hf, (uint8_t(*)[4])doub.data(), 4096);
            CHECK(ns == mol.count_singex(key_from_bytes(hf, 3)) && nd == mol.count_doub_ex(key_from_bytes(hf, 3)));
            CHECK(ns > 0 && nd > 0);
            for (size_t e = 0; e < ns; e++) {
                const uint8_t *o = &sing[2 * e];
                CHECK(read_bit(hf, o[0]) && !read_bit(hf, o[1]) && o[0] / n_orb == o[1] / n_orb &&
                      mi.symm[o[0] % n_orb] == mi.symm[o[1] % n_orb]);
            }
            for (size_t e = 0; e < nd; e++) {
                const uint8_t *o = &doub[4 * e];
                CHECK(read_bit(hf, o[0]) && read_bit(hf, o[1]) && !read_bit(hf, o[2]) && !read_bit(hf, o[3]));
                CHECK((mi.symm[o[0] % n_orb] ^ mi.symm[o[1] % n_orb] ^ mi.symm[o[2] % n_orb] ^ mi.symm[o[3] % n_orb]) == 0);
            }
            CHECK(std::isfinite(mol.diag_matrel(hf)) && mol.diag_matrel(hf) == mol.diag_matrel(key_from_bytes(hf, 3)));
            CHECK(std::isfinite(mol.sing_matr_el_nosgn(&sing[0], hf)) && std::isfinite(mol.doub_matr_el_nosgn(&doub[0])));
            CHECK(mol.calc_unnorm_wt(&doub[0]) > 0 && mol.calc_norm_wt(&doub[0], hf) > 0);
            // calc_*_probs: normalised rows (sum 1) of the documented lengths, positive totals
            {
                uint8_t occ[6] = {0, 1, 2, 10, 11, 12};
                double row[FRIES_MAX_SUB];
                auto sums_to_one = [&](unsigned len) {
                    double t = 0;
                    for (unsigned j = 0; j < len; j++) t += row[j];
                    return std::fabs(t - 1) <= 1e-12;
                };
                CHECK(mol.calc_o1_probs(row, n_elec, occ, 0) > 0 && sums_to_one(n_elec));
                CHECK(mol.calc_o1_probs(row, n_elec, occ, 1) > 0 && sums_to_one(n_elec - 1));
                CHECK(mol.calc_o2_probs(row, n_elec, occ, 4) > 0 && sums_to_one(n_elec) && row[4] == 0);
                CHECK(mol.calc_o2_probs_half(row, n_elec, occ, 4) > 0 && sums_to_one(4));
                CHECK(mol.calc_u1_probs(row, 1, occ, (uint8_t)n_elec, 0) > 0 && sums_to_one(n_orb - n_elec / 2));
                uint16_t len = 0;
                CHECK(mol.calc_u2_probs(row, 1, 11, 4, &len) > 0 && len > 0 && sums_to_one(len));
                len = 0;
                CHECK(mol.calc_u2_probs_half(row, 1, 11, 4, hf, &len) > 0 && len > 0 && sums_to_one(len));
            }
            // apply_HBPP_sys / apply_HBPP_piv from one determinant (row 1 of all_dets): every sample is an excitation of it
            Matrix<uint8_t> all_dets(2, 3);
            memcpy(all_dets[1], hf, 3);
            auto valid = [&](HBCompress &cs, const char *what) {
                CHECK(cs.vec_len > 0 && cs.vec_len <= cs.vec1.size());
                for (size_t k = 0; k < cs.vec_len; k++) {
                    const uint8_t *o = cs.orb_indices1[k];
                    bool single = o[2] == 0 && o[3] == 0;  // frisys_mol.cpp:451
                    bool ok = cs.det_indices2[k] == 1 && cs.vec1[k] != 0 && std::isfinite(cs.vec1[k]) && read_bit(hf, o[0]) &&
                              (single ? !read_bit(hf, o[1]) : (read_bit(hf, o[1]) && !read_bit(hf, o[2]) && !read_bit(hf, o[3])));
                    if (!ok) {
                        std::cout << what << ": sample " << k << " is not an excitation of the input determinant" << std::endl;
                        n_fail++;
                        break;
                    }
                }
            };
            {
                HBCompressSys cs(4096, 32);
                cs.vec1[0] = 1.0;
                cs.det_indices1[0] = 1;
                cs.vec_len = 1;
                std::mt19937 g(3), g_ref(3);
                apply_HBPP_sys(mol, all_dets, &cs, 0.9, true, g, 300);
                g_ref.discard(5);  // one uniform per compression stage
                CHECK(g() == g_ref());
                valid(cs, "apply_HBPP_sys");
                std::cout << "apply_HBPP_sys: " << cs.vec_len << " samples" << std::endl;
            }
            try {
                HBCompressPiv cs(4096, 32);
                cs.vec1[0] = 1.0;
                cs.det_indices1[0] = 1;
                cs.vec_len = 1;
                std::mt19937 g(3);
                apply_HBPP_piv(mol, all_dets, &cs, 0.9, true, g, 300);
                valid(cs, "apply_HBPP_piv");
                std::cout << "apply_HBPP_piv: " << cs.vec_len << " samples" << std::endl;
            } catch (std::exception &e) {
                std::cout << "apply_HBPP_piv: Exception : " << e.what() << std::endl;
                n_fail++;
            }
        }
        // comp_sub with more samples than sub-weights is the identity (tests/test_compression.cpp:64-118 in the reference)
        {
            const size_t count = 5, n_sub = 4;
            std::vector<double> values{1.0, 2.0, 0.5, 3.0, 1.5}, new_vals(64), wt_remain(count);
            std::vector<unsigned int> n_div(count, 0);
            Matrix<double> sub_weights(count, n_sub);
            Matrix<bool> keep_idx(count, n_sub);
            for (size_t i = 0; i < count; i++)
                for (size_t j = 0; j < n_sub; j++) sub_weights(i, j) = 0.25;
            std::vector<size_t> new_idx(2 * 64);
            size_t n = comp_sub(ctx, values.data(), count, n_div.data(), sub_weights, keep_idx, nullptr, 64, wt_remain.data(), 0.37,
                                new_vals.data(), (size_t(*)[2])new_idx.data());
            CHECK(n == count * n_sub);
            double tot = 0;
            for (size_t k = 0; k < n && k < 64; k++) {
                CHECK(new_idx[2 * k] < count && new_idx[2 * k + 1] < n_sub);
                CHECK(std::fabs(new_vals[k] - values[new_idx[2 * k]] * 0.25) <= 1e-14);
                tot += new_vals[k];
            }
            CHECK(std::fabs(tot - 8.0) <= 1e-12);
        }
        // compress_vecs_sys / compress_vecs on the store: at most compress_size (+ preserved) elements stay, the norm of
        // the row is conserved by the systematic scheme
        {
            std::vector<size_t> srt;
            std::vector<bool> keep, del;
            std::mt19937 g(5);
            double before = vec.local_norm();
            compress_vecs_sys(vec, 0, 1, 50, srt, keep, del, g);
            CHECK(vec.curr_size() <= 50 && vec.curr_size() > 0);
            CHECK(std::fabs(vec.local_norm() - before) <= 1e-9 * before);
            compress_vecs(vec, 0, 1, 20, srt, keep, del, g);
            CHECK(vec.curr_size() <= 20 && vec.curr_size() > 0);
            CHECK(std::fabs(vec.local_norm() - before) <= 1e-9 * before);
        }
    } catch (std::exception &e) {
        std::cout << "Exception : " << e.what() << std::endl;
        return 2;
    }
    std::cout << (n_fail ? "distvec_check: FAILED" : "distvec_check: all checks passed") << std::endl;
    return n_fail ? 1 : 0;
}
