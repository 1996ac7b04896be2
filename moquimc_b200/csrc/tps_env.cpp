// tps_env -- treatment-planning front end of the B200 proton transport path: same command line
// (one optional argument, the input-parameter file, default ./moqui_tps.in), same input keys and the
// same output files as the reference's tests/mc/tps/tps_env.cpp:25-43, over the C ABI of
// libmqi_b200.so.  The RTPLAN / CT DICOM inputs are replaced by a text plan and an .mha volume
// (mqi_tps_host.hpp explains why).  `--dry-run` parses everything and prints the beam source
// (beamlets, histories per spot, CT edges) as JSON without touching a GPU: used by the CPU tests.
#include "mqi_tps_host.hpp"

#include <iterator>

static void
dump_json(mqib::tps_env& env) {
    // the builders narrate on stdout like the reference: run them before the one-line JSON starts
    std::vector<std::vector<mqib::tps_env::beamline_node>> all_nodes;
    for (size_t q = 0; q < env.beam_numbers.size(); ++q) all_nodes.push_back(env.build_beamline(env.plan.beams[env.beam_numbers[q] - 1]));
    unsigned long long   roi_scoring = 0, roi_stat = 0;
    const uint64_t       nvox = (uint64_t) env.ct.nx * env.ct.ny * env.ct.nz;
    std::vector<uint8_t> m_scoring, m_stat;
    env.roi_masks(m_scoring, m_stat);
    if (!m_scoring.empty()) roi_scoring = mqib::mask_to_roi(m_scoring.data(), nvox).size();
    if (!m_stat.empty()) roi_stat = mqib::mask_to_roi(m_stat.data(), nvox).size();
    if (const char* dump = getenv("MQI_DRYRUN_DUMP_MASKS")) {   // test hook: the summed mask volumes as raw uint8
        if (!m_scoring.empty()) std::ofstream(std::string(dump) + "/scoring_mask.raw", std::ios::binary).write((const char*) m_scoring.data(), m_scoring.size());
        if (!m_stat.empty()) std::ofstream(std::string(dump) + "/stat_mask.raw", std::ios::binary).write((const char*) m_stat.data(), m_stat.size());
    }
    printf("DRYRUN {\"seed\": %d, \"n_fractions\": %d, \"sim_type\": %d, \"beams\": [", env.master_seed, env.n_fractions, (int) env.sim_type);
    for (size_t q = 0; q < env.beam_numbers.size(); ++q) {
        const mqib::plan_beam& b = env.plan.beams[env.beam_numbers[q] - 1];
        env.sid = b.snout + 50;
        std::vector<mqi_beamlet> bl;
        std::vector<uint64_t>    nh;
        env.build_source(b, bl, nh);
        printf("%s{\"name\": \"%s\", \"sid\": %.9g, \"spots\": [", q ? ", " : "", b.name.c_str(), env.sid);
        for (size_t i = 0; i < bl.size(); ++i) {
            printf("%s{\"histories\": %llu, \"energy\": %.9g, \"sigma_energy\": %.9g, \"mean\": [", i ? ", " : "",
                   (unsigned long long) nh[i], bl[i].energy, bl[i].sigma_energy);
            for (int k = 0; k < 6; ++k) printf("%s%.9g", k ? ", " : "", bl[i].mean[k]);
            printf("], \"sigma\": [");
            for (int k = 0; k < 6; ++k) printf("%s%.9g", k ? ", " : "", bl[i].sigma[k]);
            printf("], \"rot\": [");
            for (int k = 0; k < 9; ++k) printf("%s%.9g", k ? ", " : "", bl[i].rot[k]);
            printf("], \"trans\": [%.9g, %.9g, %.9g]}", bl[i].trans[0], bl[i].trans[1], bl[i].trans[2]);
        }
        printf("], \"beamline\": [");
        const std::vector<mqib::tps_env::beamline_node>& nodes = all_nodes[q];
        for (size_t i = 0; i < nodes.size(); ++i) {
            const auto& n = nodes[i];
            size_t n_open = 0;
            for (float r : n.rho) n_open += r < 1e-7f;
            double sx = 0, sy = 0;   // centroid of the open voxels of the first z layer (aperture orientation check)
            const size_t nx = n.xe.size() - 1, ny = n.ye.size() - 1;
            size_t       c0 = 0;
            for (size_t j = 0; j < ny; ++j)
                for (size_t k = 0; k < nx; ++k)
                    if (n.rho[j * nx + k] < 1e-7f) { sx += n.xe[k] + 0.5; sy += n.ye[j] + 0.5; ++c0; }
            printf("%s{\"pos_z\": %.9g, \"n\": [%zu, %zu, %zu], \"xe\": [%.9g, %.9g], \"ye\": [%.9g, %.9g], \"ze\": [%.9g, %.9g], "
                   "\"rho0\": %.9g, \"open_voxels\": %zu, \"open_centroid\": [%.9g, %.9g]}",
                   i ? ", " : "", n.pos_z, nx, ny, n.ze.size() - 1, n.xe.front(), n.xe.back(), n.ye.front(), n.ye.back(), n.ze.front(),
                   n.ze.back(), n.rho[0], n_open, c0 ? sx / c0 : 0.0, c0 ? sy / c0 : 0.0);
        }
        printf("]}");
    }
    printf("], \"scoring_roi_size\": %llu, \"stat_roi_size\": %llu", roi_scoring, roi_stat);
    {
        long long hu_sum = 0;
        int       hu_min = 32767, hu_max = -32768;
        for (int16_t v : env.ct.hu) { hu_sum += v; hu_min = std::min<int>(hu_min, v); hu_max = std::max<int>(hu_max, v); }
        printf(", \"hu\": {\"sum\": %lld, \"min\": %d, \"max\": %d}", hu_sum, hu_min, hu_max);
    }
    printf(", \"grid\": {\"n\": [%d, %d, %d], \"xe\": [%.9g, %.9g], \"ye\": [%.9g, %.9g], \"ze\": [%.9g, %.9g]}}\n", env.ct.nx, env.ct.ny,
           env.ct.nz, env.grid.xe.front(), env.grid.xe.back(), env.grid.ye.front(), env.grid.ye.back(), env.grid.ze.front(),
           env.grid.ze.back());
}

int
main(int argc, char* argv[]) {
    auto        start      = std::chrono::high_resolution_clock::now();
    std::string input_file = "./moqui_tps.in";
    bool        dry_run    = false;
    for (int i = 1; i < argc; ++i) {
        if (std::string(argv[i]) == "--dry-run") dry_run = true;
        else if (std::string(argv[i]) == "--npz-selftest" && i + 1 < argc) {
            // writer self-test for the CPU suite: 3 spots x 10 voxels, entries in "slot order"
            const std::vector<uint32_t> vox { 7, 2, 9, 2, 0, 5 }, spot { 2, 0, 2, 1, 0, 2 };
            const std::vector<double>   val { 0.5, 1.5, 2.5, 3.5, 4.5, 5.5 };
            mqib::save_csr_npz(argv[i + 1], 3, 10, vox, spot, val);
            return 0;
        } else if (std::string(argv[i]) == "--parse-selftest" && i + 1 < argc) {
            // input-parameter format self-test for the CPU suite: the queries of oracle/ref_kat.cpp section 11, answered by
            // this file_parser; the reference's own file_parser answered them in tests/golden/fmt_writers.npz
            mqib::file_parser p(argv[i + 1], " ");
            std::ostream&     o = std::cout;
            const char*       skeys[] = { "GPUID", "randomseed", "ParentDir", "SCORER", "Mask", "OutputDir", "Machine", "UnitWeights",
                                    "Duplicate", "NoValue", "Missing", "EmptyList", "Trailing", "ZShift" };
            for (const char* k : skeys) o << k << "|s|" << p.get_string(k, "<default>") << "|\n";
            const char* ikeys[] = { "GPUID", "RandomSeed", "BeamNumbers", "Missing", "Trailing", "XShift", "ParticlesPerHistory" };
            for (const char* k : ikeys) o << k << "|i|" << p.get_int(k, -7) << "|\n";
            const char* fkeys[] = { "XShift", "YShift", "ZShift", "ParticlesPerHistory", "Missing", "RandomSeed" };
            for (const char* k : fkeys) o << k << "|f|" << std::setprecision(9) << p.get_float(k, 0.125f) << "|\n";
            const char* bkeys[] = { "OverwriteResults", "SaveMap", "ReadStructure", "ScoringMask", "StoppingStatistics", "Missing",
                                    "UnitWeights", "GPUID", "XShift" };
            for (const char* k : bkeys) o << k << "|b|" << p.get_bool(k, false) << p.get_bool(k, true) << "|\n";
            const char* vkeys[] = { "scorer", "Mask", "BeamNumbers", "Missing", "EmptyList", "GPUID" };
            for (const char* k : vkeys) {
                o << k << "|v|";
                for (const auto& t : p.get_string_vector(k, ",")) o << "[" << t << "]";
                o << "|";
                for (int t : p.get_int_vector(k, ",")) o << t << ";";
                o << "|\n";
            }
            return 0;
        } else if (std::string(argv[i]) == "--mask-selftest" && i + 4 < argc) {
            // mask files self-test for the CPU suite: --mask-selftest a.mha,b.mha nx ny nz reads the uint8 .mha masks,
            // adds them and prints the run-length roi (the reference's own mask_reader did the same on the same files:
            // oracle/ref_kat.cpp section 12)
            std::vector<std::string> files;
            std::stringstream        ss(argv[i + 1]);
            for (std::string f; std::getline(ss, f, ',');) files.push_back(f);
            const std::vector<uint8_t> m = mqib::read_mask_files(files, std::atoi(argv[i + 2]), std::atoi(argv[i + 3]), std::atoi(argv[i + 4]));
            const mqib::RoiRuns        r = mqib::mask_to_roi(m.data(), m.size());
            printf("runs %zu size %u\n", r.start.size(), r.size());
            for (size_t k = 0; k < r.start.size(); ++k) printf("run %u %u %u\n", r.start[k], r.stride[k], r.acc_stride[k]);
            printf("total");
            for (uint8_t v : m) printf(" %u", (unsigned) v);
            printf("\n");
            return 0;
        } else if (std::string(argv[i]) == "--roi-selftest" && i + 1 < argc) {
            // run-length roi of a raw uint8 summed-mask file (CPU suite): prints "start stride acc_stride" per run,
            // then the compressed index of every 7th voxel
            std::ifstream        f(argv[i + 1], std::ios::binary);
            std::vector<uint8_t> m((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            const mqib::RoiRuns  r = mqib::mask_to_roi(m.data(), m.size());
            printf("runs %zu size %u\n", r.start.size(), r.size());
            for (size_t k = 0; k < r.start.size(); ++k) printf("run %u %u %u\n", r.start[k], r.stride[k], r.acc_stride[k]);
            for (size_t v = 0; v < m.size(); v += 7) printf("idx %zu %lld\n", v, (long long) r.compressed_index((uint32_t) v));
            const std::vector<uint32_t> bits = r.bitmask();
            size_t                      pop  = 0;
            for (uint32_t w : bits) pop += (size_t) __builtin_popcount(w);
            printf("bits %zu\n", pop);
            return 0;
        } else input_file = argv[i];
    }
    try {
        mqib::tps_env myenv(input_file);
        if (dry_run) {
            dump_json(myenv);
            return 0;
        }
        myenv.initialize_and_run();
    } catch (const std::exception& e) {
        std::cerr << "tps_env: " << e.what() << std::endl;
        return 1;
    }
    auto                                      stop     = std::chrono::high_resolution_clock::now();
    std::chrono::duration<double, std::milli> duration = stop - start;
    std::cout << "Time taken by MC engine: " << duration.count() << " milli-seconds\n";
    return 0;
}
