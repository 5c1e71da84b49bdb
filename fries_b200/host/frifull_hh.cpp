// frifull_hh -- FRI without matrix compression for the 1-D Hubbard-Holstein model, systematic vector compression
// (FRIES_bin/frifull_hh.cpp), same command line, stdout lines and output files; loop body = fries_frifull_hh_iterate.
#include "fries_host.hpp"

using namespace fries;

int main(int argc, char *argv[]) {
    Args args(argc, argv);
    std::string params_path = args.str("params_path");
    double target_norm = args.num("target", 0);
    uint32_t max_iter = (uint32_t)args.num("max_iter", 1000000);
    uint32_t target_nonz = (uint32_t)args.num("vec_nonz");
    std::string result_dir = args.str("result_dir", "./");
    size_t max_n_dets = (size_t)args.num("max_dets");
    double init_thresh = args.num("initiator", 0);
    bool has_load = args.has("load_dir");
    std::string load_dir = args.str("load_dir", "");
    int device = (int)args.num("device", 0);
    args.validate();
    try {
        Context ctx(device);
        double shift_damping = 0.05;
        unsigned shift_interval = 10, save_interval = 1000;
        double en_shift = 0;
        HhInput in_data = parse_hh_input(params_path);
        double eps = in_data.eps;
        unsigned hub_len = in_data.lat_len, n_elec = in_data.n_elec;
        if (in_data.n_dim != 1) {
            fprintf(stderr, "Error: only 1-D Hubbard calculations supported right now.\n");
            return 0;
        }
        unsigned n_orb = hub_len;
        unsigned seed = seed_from_clock_or_env();
        std::cout << "seed on process 0 is " << seed << std::endl;
        std::mt19937 mt_obj(seed);
        std::vector<uint32_t> proc_scrambler(2 * n_orb), vec_scrambler(2 * n_orb);
        if (has_load) {
            load_proc_hash(load_dir, proc_scrambler);
        } else {
            for (auto &x : proc_scrambler) x = mt_obj();
            save_proc_hash(result_dir, proc_scrambler);
        }
        for (auto &x : vec_scrambler) x = mt_obj();
        // frifull_hh.cpp:97-100: spawn_length = n_elec * 4 * max_n_dets / n_procs, at most 200000; here the window of one
        // spawn launch, a whole number of states' 4 * n_elec places
        size_t spawn_length = std::min<size_t>((size_t)n_elec * 4 * max_n_dets, (size_t)1 << 24);
        if (spawn_length < (size_t)4 * n_elec) spawn_length = (size_t)4 * n_elec;
        unsigned ph_bits = 3;
        DistVec sol_vec(ctx, max_n_dets, hub_len, ph_bits, n_elec, 2, proc_scrambler, vec_scrambler);
        uint64_t neel_det = gen_neel_det_1D(n_orb, n_elec);
        double last_one_norm = 0;
        if (has_load) {
            sol_vec.load(load_dir);
            last_one_norm = sol_vec.local_norm();
        } else {
            sol_vec.add(neel_det, 100.0, 1);  // DistVec::add + perform_add, as the reference does
            sol_vec.perform_add(0);
        }
        auto open_app = [&](const char *name) {
            std::ofstream f(result_dir + name, std::ofstream::app);
            if (!f.is_open()) throw std::runtime_error("Could not open file for writing in directory " + result_dir);
            return f;
        };
        std::ofstream num_file = open_app("projnum.txt"), den_file = open_app("projden.txt"), shift_file = open_app("S.txt"),
                      norm_file = open_app("norm.txt");
        {
            std::ofstream param_f(result_dir + "params.txt");
            param_f << "FRI calculation\nHubbard-Holstein parameters path: " << params_path
                    << "\nepsilon (imaginary time step): " << eps << "\nTarget norm " << target_norm
                    << "\nInitiator threshold: " << init_thresh << "\nVector nonzero: " << target_nonz << "\n";
            if (has_load) param_f << "Restarting calculation from " << load_dir << "\n";
            else param_f << "Initializing calculation from Neel unit vector\n";
        }
        check(fries_frisys_hh_setup(sol_vec.h, spawn_length, &sol_vec.hb));
        for (unsigned iterat = 0; iterat < max_iter; iterat++) {
            double rn_sys = mt_obj() / (1. + UINT32_MAX);  // one draw per iteration, for sys_comp (frifull_hh.cpp:305-307)
            fries_frisys_hh_params p{eps, init_thresh, in_data.elec_int, in_data.ph_freq, in_data.elec_ph, in_data.hf_en,
                                     target_nonz, en_shift, neel_det};
            fries_iter_stats st;
            check(fries_frifull_hh_iterate(sol_vec.h, sol_vec.hb, &p, rn_sys, &st));
            if ((iterat + 1) % shift_interval == 0) {
                adjust_shift(&en_shift, st.glob_norm, &last_one_norm, target_norm, shift_damping / shift_interval / eps);
                shift_file << en_shift << '\n';
                norm_file << st.glob_norm << '\n';
            }
            num_file << st.numer << '\n';
            den_file << st.denom << '\n';
            std::cout << iterat << ", norm: " << st.glob_norm << ", en est: " << st.numer / st.denom << ", shift: " << en_shift
                      << ", n_neel: " << st.denom << '\n';
            if ((iterat + 1) % save_interval == 0) {
                sol_vec.save(result_dir);
                num_file.flush();
                den_file.flush();
                shift_file.flush();
            }
        }
        sol_vec.save(result_dir);
    } catch (std::exception &ex) {
        std::cerr << "\nException : " << ex.what() << "\n\n";
    }
    return 0;
}
