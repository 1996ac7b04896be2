// phantom_env -- drop-in for the reference's tests/mc/phantom/phantom_env.cpp:5-21 (same flags, same
// stdout keys, same output files), running the transport on B200 GPUs through libmqi_b200.so.
//
//   phantom_env --lxyz 100 100 350 --pxyz 0 0 -175 --nxyz 200 200 350 --spot_energy 200 0
//               --spot_position 0 0 0.5 --spot_size 30 30 --histories 100000
//               --phantom_path water_phantom.raw --output_prefix out --random_seed 12345 --gpu_id 0
//
// Extensions: --gpu_id may list several devices (histories sharded, one NCCL reduce of the dose grid);
// --physics debug|release selects which compile-time physics of the reference to reproduce
// (default debug = the reference's phantom CMake); --output_format raw|mhd|mha is honoured.
#include "mqi_host.hpp"

#include <chrono>
#include <iostream>

// phantom_env --write-selftest <dir>: the output writers on a small fixed volume, without a device.  The files are
// compared byte for byte with the ones the reference's io::save_to_mhd / save_to_mha wrote for the same volume
// (oracle/ref_kat.cpp section 9 -> tests/golden/fmt_writers.npz, tests/test_cli.py).
static int
write_selftest(const std::string& dir) {
    mqib::grid_desc g;
    g.xe = { -1.25f, -0.75f, -0.25f, 0.25f, 0.75f };
    g.ye = { 10.f, 10.75f, 11.5f, 12.25f };
    g.ze = { -3.7f, -2.45f, -1.2f };
    double src[24];
    for (int i = 0; i < 24; ++i)
        src[i] = 0.001 * i * i - 0.0137 * i + (i % 5 == 0 ? 0.0 : 1e-9 * i);
    mqib::save_to_mhd(g, src, (double) 2.5f, dir, "fmt_mhd", 24);
    mqib::save_to_mha(g, src, (double) 2.5f, dir, "fmt_mha", 24);
    return 0;
}

int
main(int argc, char* argv[]) {
    if (argc == 3 && std::string(argv[1]) == "--write-selftest") return write_selftest(argv[2]);
    auto      start = std::chrono::high_resolution_clock::now();
    mqib::cli cl;
    cl.read(argc, argv);
    try {
        mqib::phantom_env myenv(cl);
        myenv.initialize();
        myenv.run();
        myenv.finalize();
        myenv.save_reshaped_files();
    } catch (const std::exception& e) {
        // the reference lets std::runtime_error escape main() (abort); report and fail instead
        std::cerr << "phantom_env: " << e.what() << std::endl;
        return 1;
    }
    auto                                      stop     = std::chrono::high_resolution_clock::now();
    std::chrono::duration<double, std::milli> duration = stop - start;
    std::cout << "Time taken by MC engine: " << duration.count() << " milli-seconds\n";
    return 0;
}
