// frifull_mol -- FRI with a deterministic, full H.v and systematic vector compression
// (FRIES_bin/frifull_mol.cpp), same command line and output files; loop body = fries_frifull_mol_iterate.
#include "fries_host.hpp"

using namespace fries;

int main(int argc, char *argv[]) {
    Args args(argc, argv);
    std::string hf_path = args.str("hf_path");
    double target_norm = args.num("target", 0);
    uint32_t max_iter = (uint32_t)args.num("max_iter", 1000000);
    uint32_t target_nonz = (uint32_t)args.num("vec_nonz");
    std::string result_dir = args.str("result_dir", "./");
    size_t max_n_dets = (size_t)args.num("max_dets");
    bool has_load = args.has("load_dir"), has_ini = args.has("ini_vec"), has_trial = args.has("trial_vec");
    std::string load_dir = args.str("load_dir", ""), ini_path = args.str("ini_vec", ""), trial_path = args.str("trial_vec", "");
    int device = (int)args.num("device", 0);
    args.validate();
    try {
        // several GPUs: started by fries_launch -n N, one process per GPU (the reference: mpirun -n N, frifull_mol.cpp:34-38)
        Ranks rk;
        const bool multi = rk.n > 1;
        if (multi && !args.has("device")) device = rk.rank;
        Context ctx(device);
        double shift_damping = 0.05;
        unsigned shift_interval = 10, save_interval = 100;
        double en_shift = 0;
        MolInput in_data = parse_hf_input(hf_path);
        double eps = in_data.eps;
        unsigned n_elec = in_data.n_elec, n_frz = in_data.n_frz, n_orb = in_data.n_orb;
        unsigned n_elec_unf = n_elec - n_frz;
        double hf_en = in_data.hf_en;
        Molecule mol(ctx, in_data);

        unsigned seed = rk.bcast0("seed", seed_from_clock_or_env());
        if (rk.rank == 0) std::cout << "seed on process 0 is " << seed << std::endl;
        std::mt19937 mt_obj(seed);
        std::vector<uint32_t> proc_scrambler(2 * n_orb), vec_scrambler(2 * n_orb);
        if (has_load) {
            load_proc_hash(load_dir, proc_scrambler);
        } else {
            for (auto &x : proc_scrambler) x = mt_obj();
            if (rk.rank == 0) save_proc_hash(result_dir, proc_scrambler);
        }
        for (auto &x : vec_scrambler) x = mt_obj();
        DistVec sol_vec(ctx, max_n_dets, 2 * n_orb, n_elec_unf, 2, proc_scrambler, vec_scrambler, rk.n, rk.rank);
        check(fries_vec_set_diag_mol(sol_vec.h, mol.h, hf_en));
        uint64_t hf_det = gen_hf_bitstring(n_orb, n_elec_unf);
        // the rank that owns the Hartree-Fock determinant writes the text files and stdout (frifull_mol.cpp:302)
        int hf_proc = 0;
        if (multi) {
            int32_t own = 0;
            check(fries_hash_owner(ctx.h, &hf_det, 1, proc_scrambler.data(), (int)(2 * n_orb), rk.n, nullptr, &own));
            hf_proc = own;
        }
        const bool writer = rk.rank == hf_proc;

        std::vector<uint64_t> trial_dets{hf_det};
        std::vector<double> trial_vals{1.0};
        if (has_trial) load_vec_txt(trial_path, trial_dets, trial_vals);

        if (has_load) {
            sol_vec.load(load_dir);
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
        double last_one_norm = 0;
        auto open_app = [&](const char *name) {
            std::ofstream f(writer ? result_dir + name : std::string("/dev/null"), std::ofstream::app);
            if (!f.is_open()) throw std::runtime_error("Could not open file for writing in directory " + result_dir);
            return f;
        };
        std::ofstream num_file = open_app("projnum.txt"), den_file = open_app("projden.txt"), shift_file = open_app("S.txt"),
                      norm_file = open_app("norm.txt"), nkept_file = open_app("nkept.txt");
        if (writer) {
            std::ofstream param_f(result_dir + "params.txt");
            param_f << "FRI calculation\nHF path: " << hf_path << "\nepsilon (imaginary time step): " << eps
                    << "\nTarget norm " << target_norm << "\nVector nonzero: " << target_nonz << "\n";
            if (has_load) param_f << "Restarting calculation from " << load_dir << "\n";
            else if (has_ini) param_f << "Initializing calculation from vector files with prefix " << ini_path << '\n';
            else param_f << "Initializing calculation from HF unit vector\n";
        }
        // spawn window: the reference sizes its adder at min(1e6, target_nonz * num_ex / 4) (frifull_mol.cpp:66-68)
        size_t num_ex = (size_t)n_elec_unf * n_elec_unf * (n_orb - n_elec_unf / 2) * (n_orb - n_elec_unf / 2);
        size_t window = std::max<size_t>(1 << 22, std::min<size_t>((size_t)target_nonz * num_ex / 4, (size_t)1 << 26));
        std::vector<uint64_t> none_k;
        std::vector<double> none_v;
        check(fries_frisys_mol_setup(sol_vec.h, mol.h, window, trial_dets.data(), trial_vals.data(), trial_dets.size(),
                                     none_k.data(), none_v.data(), 0, &sol_vec.hb));
        // several GPUs: spawned elements are stored into their owner's window from inside the H.v kernel (the reference:
        // Adder::perform_add's MPI_Alltoallv, vec_utils.hpp:991-1019); a rank's segment holds one chunk of parents' spawns
        std::unique_ptr<Comm> comm;
        if (multi) {
            const size_t seg_cap = std::min<size_t>(window / (size_t)rk.n, (size_t)1 << 22);  // connections per round and rank
            comm.reset(new Comm(ctx, rk, seg_cap));
            check(fries_hbpp_set_route_p2p(sol_vec.hb, comm->h));
        }
        for (unsigned iterat = 0; iterat < max_iter; iterat++) {
            double rn_sys = mt_obj() / (1. + UINT32_MAX);
            int adjust = (iterat + 1) % shift_interval == 0;
            fries_frifull_params p{eps, target_nonz, en_shift, adjust, shift_damping / shift_interval / eps, target_norm,
                                   last_one_norm};
            fries_iter_stats st;
            // adjust_shift sits between find_preserve and sys_comp of the same iteration (frifull_mol.cpp:270-276)
            // and feeds this iteration's h_op_diag, so the call performs it and returns the new shift
            check(fries_frifull_mol_iterate(sol_vec.h, mol.h, sol_vec.hb, &p, rn_sys, &st));
            en_shift = p.en_shift;
            last_one_norm = p.last_one_norm;
            nkept_file << st.n_kept << '\n';
            if (adjust) {
                shift_file << en_shift << "\n";
                norm_file << st.glob_norm << "\n";
            }
            num_file << st.numer << '\n';
            den_file << st.denom << "\n";
            if (writer)
                std::cout << iterat << ", en est: " << st.numer / st.denom << ", shift: " << en_shift << ", norm: " << st.glob_norm
                          << '\n';
            if ((iterat + 1) % save_interval == 0) {
                sol_vec.save(result_dir);
                num_file.flush();
                den_file.flush();
                shift_file.flush();
                nkept_file.flush();
            }
        }
        sol_vec.save(result_dir);
        if (multi) rk.barrier("done");  // nobody unmaps its windows while a peer may still be inside an exchange
    } catch (std::exception &ex) {
        std::cerr << "\nException : " << ex.what() << "\n\n";
        // the reference returns 0 here as well (frifull_mol.cpp:330-333); under fries_launch a failed rank must be seen
        if (std::getenv("FRIES_NRANKS") && std::atoi(std::getenv("FRIES_NRANKS")) > 1) return 1;
    }
    return 0;
}
