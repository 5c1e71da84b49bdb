// frisys_mol -- FRI with systematic compression for a molecular Hamiltonian (FRIES_bin/frisys_mol.cpp), same
// command line, stdout lines and output files; the loop body runs on the GPU (fries_frisys_mol_iterate).
#include "fries_host.hpp"

using namespace fries;

int main(int argc, char *argv[]) {
    Args args(argc, argv);
    // --hf_path <dir>/ (legacy input directory, io_utils.cpp:98-187) is accepted in place of --fcidump_path, as in the
    // reference's examples/run_neon.sh and Benchmarks/Results.tex command lines: epsilon then defaults to the eps of
    // sys_params.txt, the distribution to HB_unnorm, frozen core electrons are honoured and energies are relative to
    // hf_energy.  Everything else is the HEAD command line (FRIES_bin/frisys_mol.cpp:17-32).
    bool legacy = args.has("hf_path");
    std::string hf_path = args.str("hf_path", "");
    std::string fcidump_path = legacy ? args.str("fcidump_path", "") : args.str("fcidump_path");
    double target_norm = args.num("target", 0);
    std::string dist_str = legacy ? args.str("distribution", "HB_unnorm") : args.str("distribution");
    uint32_t max_iter = (uint32_t)args.num("max_iter", 1000000);
    uint32_t target_nonz = (uint32_t)args.num("vec_nonz");
    uint32_t matr_samp = (uint32_t)args.num("mat_nonz");
    std::string result_dir = args.str("result_dir", "./");
    size_t max_n_dets = (size_t)args.num("max_dets");
    double init_thresh = args.num("initiator", 0);
    bool has_load = args.has("load_dir"), has_ini = args.has("ini_vec"), has_trial = args.has("trial_vec");
    std::string load_dir = args.str("load_dir", ""), ini_path = args.str("ini_vec", ""), trial_path = args.str("trial_vec", "");
    bool has_det_space = args.has("det_space");
    std::string det_space_path = args.str("det_space", "");
    double eps = legacy ? args.num("epsilon", -1) : args.num("epsilon");
    std::string point_group = args.str("point_group", "C1");
    bool has_shift = args.has("ham_shift");
    double ham_shift = args.num("ham_shift", 0);
    int device = (int)args.num("device", 0);
    args.validate();

    int new_hb;
    if (dist_str == "HB") new_hb = 0;
    else if (dist_str == "HB_unnorm") new_hb = 1;
    else {
        std::cerr << "\nError parsing command line: \"dist_str\" argument must be either \"NU\" or \"HB_unnorm\"\n\n";
        return 1;
    }
    try {
        // several GPUs: started by fries_launch -n N, one process per GPU (rank r drives GPU r unless --device says otherwise)
        Ranks rk;
        const bool multi = rk.n > 1;
        if (multi && !args.has("device")) device = rk.rank;
        Context ctx(device);
        double shift_damping = 0.05;
        unsigned shift_interval = 10, save_interval = 100;
        double en_shift = 0;

        MolInput in_data = legacy ? parse_hf_input(hf_path) : parse_fcidump(fcidump_path, point_group);
        unsigned n_elec = in_data.n_elec, n_frz = legacy ? in_data.n_frz : 0, n_orb = in_data.n_orb;
        unsigned n_elec_unf = n_elec - n_frz;
        if (legacy && eps < 0) eps = in_data.eps;
        Molecule mol(ctx, in_data);
        uint64_t hf_det = gen_hf_bitstring(n_orb, n_elec_unf);
        double hf_en = has_shift ? ham_shift - in_data.core_en : (legacy ? in_data.hf_en : mol.diag_matrel(hf_det));

        // one generator state on every rank: the reference draws on rank 0 and broadcasts (frisys_mol.cpp:132-139,528-530)
        unsigned seed = rk.bcast0("seed", seed_from_clock_or_env());
        if (rk.rank == 0) std::cout << "seed on process 0 is " << seed << std::endl;
        std::mt19937 mt_obj(seed);

        unsigned spawn_length = matr_samp * 4 / rk.n;  // frisys_mol.cpp:109: spawn_length = matr_samp * 4 / n_procs
        std::vector<uint32_t> proc_scrambler(2 * n_orb), vec_scrambler(2 * n_orb);
        if (has_load) {
            load_proc_hash(load_dir, proc_scrambler);
        } else {
            for (auto &x : proc_scrambler) x = mt_obj();
            if (rk.rank == 0) save_proc_hash(result_dir, proc_scrambler);
        }
        for (auto &x : vec_scrambler) x = mt_obj();

        DistVec sol_vec(ctx, max_n_dets, 2 * n_orb, n_elec_unf, 2, proc_scrambler, vec_scrambler, rk.n, rk.rank);
        // the rank that owns the Hartree-Fock determinant writes the text files and stdout (frisys_mol.cpp:288,521)
        int hf_proc = 0;
        if (multi) {
            int32_t own = 0;
            check(fries_hash_owner(ctx.h, &hf_det, 1, proc_scrambler.data(), (int)(2 * n_orb), rk.n, nullptr, &own));
            hf_proc = own;
        }
        const bool writer = rk.rank == hf_proc;
        check(fries_vec_set_diag_mol(sol_vec.h, mol.h, hf_en));

        // trial vector and H * trial
        std::vector<uint64_t> trial_dets{hf_det}, htrial_dets;
        std::vector<double> trial_vals{1.0}, htrial_vals;
        if (has_trial) load_vec_txt(trial_path, trial_dets, trial_vals);
        h_times(ctx, mol, hf_en, trial_dets, trial_vals, proc_scrambler, vec_scrambler, 2 * n_orb, n_elec_unf, htrial_dets,
                htrial_vals);

        size_t n_hf_doub = mol.count_doub_ex(hf_det), n_hf_sing = mol.count_singex(hf_det);
        double p_doub = (double)n_hf_doub / (n_hf_sing + n_hf_doub);

        // deterministic subspace of a semi-stochastic calculation (frisys_mol.cpp:233-252): its determinants come first
        size_t n_determ = 0;
        if (!has_load) {
            if (has_det_space) {
                n_determ = sol_vec.init_dense(det_space_path, result_dir, &rk);  // collective: every rank keeps its share
            } else if (rk.rank == 0) {
                std::ofstream dense_f(result_dir + "dense.txt");
                if (!dense_f.is_open()) throw std::runtime_error("Error opening file containing sizes of deterministic subspaces");
                for (int r = 0; r < rk.n; r++) dense_f << 0 << ", ";  // one size per rank (frisys_mol.cpp:243-251)
                dense_f << '\n';
            }
        }
        // initial vector
        if (has_load) {
            sol_vec.load(load_dir);
            n_determ = sol_vec.n_dense;
            load_last_line(load_dir + "S.txt", &en_shift);
        } else if (has_ini) {
            std::vector<uint64_t> d;
            std::vector<double> v;
            load_vec_txt(ini_path, d, v);
            sol_vec.add(d, v, 1);
        } else if (!multi) {
            sol_vec.add(hf_det, 100.0, 1);  // DistVec::add + perform_add, as the reference does
            sol_vec.perform_add(0);
        } else {
            sol_vec.add(std::vector<uint64_t>{hf_det}, std::vector<double>{100.0}, 1);  // lands on its owner only
        }
        double glob_norm = sol_vec.local_norm();
        double last_one_norm = 0;
        (void)glob_norm;

        auto open_app = [&](const char *name) {
            std::ofstream f(writer ? result_dir + name : std::string("/dev/null"), std::ofstream::app);
            if (!f.is_open()) throw std::runtime_error("Could not open file for writing in directory " + result_dir);
            return f;
        };
        std::ofstream num_file = open_app("projnum.txt"), den_file = open_app("projden.txt"), shift_file = open_app("S.txt"),
                      norm_file = open_app("norm.txt"), nkept_file = open_app("nkept.txt"), ini_file = open_app("nini.txt");
        if (writer) {
            std::ofstream param_f(result_dir + "params.txt");
            param_f << "FRI calculation\n" << (legacy ? "HF path: " : "FCIDUMP path: ") << (legacy ? hf_path : fcidump_path) << "\nepsilon (imaginary time step): " << eps
                    << "\nTarget norm " << target_norm << "\nInitiator threshold: " << init_thresh
                    << "\nMatrix nonzero: " << matr_samp << "\nVector nonzero: " << target_nonz << "\n";
            if (has_load) param_f << "Restarting calculation from " << load_dir << "\n";
            else if (has_ini) param_f << "Initializing calculation from vector files with prefix " << ini_path << '\n';
            else param_f << "Initializing calculation from HF unit vector\n";
        }
        check(fries_frisys_mol_setup(sol_vec.h, mol.h, spawn_length, trial_dets.data(), trial_vals.data(), trial_dets.size(),
                                     htrial_dets.data(), htrial_vals.data(), htrial_dets.size(), &sol_vec.hb));
        // several GPUs: the peer-mapped inboxes and spawn-route windows (Adder::perform_add vec_utils.hpp:991-1019 becomes
        // stores into the owner's window from inside the spawn kernel; csrc/comm.cuh)
        std::unique_ptr<Comm> comm;
        if (multi) {
            const size_t seg_cap = 2 * (size_t)matr_samp / ((size_t)rk.n * rk.n) + 8192;
            comm.reset(new Comm(ctx, rk, seg_cap));
            check(fries_hbpp_set_route_p2p(sol_vec.hb, comm->h));
        }
        // frisys_mol.cpp:398-401: size of the pre-computed dense part of H = all connections of the dense determinants;
        // the stochastic compression of H gets the rest of the budget (:421: matr_samp - tot_dense_h)
        size_t tot_dense_h = 0;
        {
            if (n_determ) {
                std::vector<uint64_t> d;
                std::vector<double> vv;
                sol_vec.download(d, vv);
                std::vector<uint64_t> off(n_determ + 1);
                check(fries_mol_sing_ex(mol.h, d.data(), n_determ, off.data(), nullptr, 0));
                tot_dense_h += off[n_determ];
                check(fries_mol_doub_ex(mol.h, d.data(), n_determ, off.data(), nullptr, 0));
                tot_dense_h += off[n_determ];
            }
            if (multi) {  // sum_mpi over the ranks (frisys_mol.cpp:398)
                unsigned long long mine = tot_dense_h;
                std::vector<uint8_t> raw = rk.allgather("dense_h", &mine, sizeof(mine));
                tot_dense_h = 0;
                for (int r = 0; r < rk.n; r++) tot_dense_h += ((const unsigned long long *)raw.data())[r];
            }
            if (writer) std::cout << "Elements in dense H: " << tot_dense_h << "\n";
        }

        for (unsigned iterat = 0; iterat < max_iter; iterat++) {
            size_t n_ini = 0;
            // RNG consumption order of the reference: 5 uniforms inside apply_HBPP_sys, then one for sys_comp
            double u[6];
            for (int k = 0; k < 6; k++) u[k] = mt_obj() / (1. + UINT32_MAX);
            // the reference leaves `matr_samp - tot_dense_h` to unsigned wrap-around when the dense part alone exceeds the
            // budget; here that is an error
            if (tot_dense_h >= matr_samp)
                throw std::runtime_error("the dense subspace's part of H (" + std::to_string(tot_dense_h) +
                                         " elements) leaves nothing of --mat_nonz for the stochastic part");
            fries_frisys_params p{eps, init_thresh, p_doub, new_hb, (uint32_t)(matr_samp - tot_dense_h), target_nonz, en_shift};
            fries_iter_stats st;
            if (!multi) {
                check(fries_frisys_mol_iterate(sol_vec.h, mol.h, sol_vec.hb, &p, u, &st));
            } else {  // spawn (elements go straight to their owners), then merge + vector half; the statistics are global
                check(fries_frisys_mol_spawn(sol_vec.h, mol.h, sol_vec.hb, &p, u));
                check(fries_frisys_mol_finish(sol_vec.h, mol.h, sol_vec.hb, &p, u, nullptr, &st));
            }
            nkept_file << st.n_kept << '\n';
            if ((iterat + 1) % shift_interval == 0) {
                // NOTE the shift used inside iteration k+1 is the one adjusted after iteration k, as in the reference
                adjust_shift(&en_shift, st.glob_norm, &last_one_norm, target_norm, shift_damping / shift_interval / eps);
                shift_file << en_shift << "\n";
                norm_file << st.glob_norm << "\n";
            }
            num_file << st.numer << '\n';
            den_file << st.denom << '\n';
            if (writer)
                std::cout << iterat << ", en est: " << st.numer / st.denom << ", shift: " << en_shift << ", norm: " << st.glob_norm
                          << '\n';
            ini_file << n_ini << '\n';
            if ((iterat + 1) % save_interval == 0) {
                sol_vec.save(result_dir);
                uint64_t tot_add = sol_vec.tot_sgn_coh();
                num_file.flush();
                den_file.flush();
                shift_file.flush();
                nkept_file.flush();
                if (writer) std::cout << "Total additions to nonzero: " << tot_add << "\n";
            }
        }
        sol_vec.save(result_dir);
        if (multi) rk.barrier("done");  // nobody unmaps its windows while a peer may still be inside an exchange
    } catch (std::exception &ex) {
        std::cerr << "\nException : " << ex.what() << "\n\n";
        // the reference returns 0 here as well (frisys_mol.cpp:562-565); under fries_launch a failed rank must be seen, or
        // its peers wait for it in the next exchange
        if (std::getenv("FRIES_NRANKS") && std::atoi(std::getenv("FRIES_NRANKS")) > 1) return 1;
    }
    return 0;
}
