// mqi_host.hpp -- C++ host side of the B200 transport path: the reference's command-line front end
// and x_environment life cycle (initialize -> run -> finalize -> save) re-expressed over the C ABI
// of include/mqi_b200.h.  Header-only, no CUDA types: everything device-side happens behind
// libmqi_b200.so.
//
// Mirrors (file:line under /root/reference/moqui):
//   mqi::cli                         base/mqi_cli.hpp:32-99          flag table, "values until the next --flag"
//   mqi::phantom_env<R>              base/environments/mqi_phantom_env.hpp:37-428
//   x_environment::save_reshaped_files   base/environments/mqi_xenvironment.hpp:169-209
//   io::save_to_bin / _mhd / _mha    base/mqi_io.hpp:165-183, 493-591
//   coordinate_transform             base/mqi_coordinate_transform.hpp:52-58 (+ mat3x3 rotate_x/y/z, mqi_matrix.hpp:214-273)
//   grid3d(min, max, n) edge rule    base/mqi_grid3d.hpp:152-161
#pragma once

#include "mqi_b200.h"

#include <array>
#include <chrono>
#include <sstream>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mqib
{

// ---------------------------------------------------------------------------------------------
// mqi::cli: fixed table of options; every token after an option up to the next "--" token is one
// of its values; unknown options are ignored (base/mqi_cli.hpp:41-99)
// ---------------------------------------------------------------------------------------------
class cli
{
protected:
    std::map<const std::string, std::vector<std::string>> parameters;

public:
    cli() {
        for (const char* k :
             { "--dicom_path", "--bname", "--bnumber", "--spots", "--beamlets", "--pph", "--sid", "--output_prefix",
               "--nhistory", "--pxyz", "--source_energy", "--energy_variance", "--rxyz", "--lxyz", "--nxyz",
               "--spot_position", "--spot_size", "--spot_angles", "--spot_energy", "--histories", "--threads",
               "--score_variance", "--gpu_id", "--output_format", "--random_seed", "--phantom_path",
               // extension of this build (not in the reference table): which compile-time physics of the
               // reference to reproduce; phantom_env is built with __PHYSICS_DEBUG__ (tests/mc/phantom/CMakeLists.txt:10)
               "--physics" })
            parameters[k] = {};
    }
    virtual ~cli() {}

    void
    read(int argc, char** argv) {
        std::cout << "# of arguments: " << argc << std::endl;
        for (int i = 1; i < argc; ++i) {
            auto it = parameters.find(argv[i]);
            if (it == parameters.end()) continue;
            // the reference dereferences argv[argc] when the last option has no value (:86-90); guarded here
            for (int j = i + 1; j < argc; ++j) {
                if (std::string(argv[j]).compare(0, 2, "--") == 0) break;
                it->second.push_back(argv[j]);
            }
            std::cout << it->first << " : ";
            for (const auto& parm : it->second) std::cout << parm << " ";
            std::cout << std::endl;
        }
    }

    const std::vector<std::string>
    operator[](const std::string& t) const {
        auto it = parameters.find(t);
        return it == parameters.end() ? std::vector<std::string>() : it->second;
    }
};

// mat3x3(a, b, c) = identity rotated about x, then y, then z (mqi_matrix.hpp:69-76, 214-273), row-major
struct mat3 {
    float m[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    void
    rotate_x(float a) {
        const float c1 = std::cos(a), s1 = std::sin(a);
        const float x1 = m[3], y1 = m[4], z1 = m[5];
        m[3] = c1 * x1 - s1 * m[6]; m[4] = c1 * y1 - s1 * m[7]; m[5] = c1 * z1 - s1 * m[8];
        m[6] = s1 * x1 + c1 * m[6]; m[7] = s1 * y1 + c1 * m[7]; m[8] = s1 * z1 + c1 * m[8];
    }
    void
    rotate_y(float a) {
        const float c1 = std::cos(a), s1 = std::sin(a);
        const float x1 = m[6], y1 = m[7], z1 = m[8];
        m[6] = c1 * x1 - s1 * m[0]; m[7] = c1 * y1 - s1 * m[1]; m[8] = c1 * z1 - s1 * m[2];
        m[0] = s1 * x1 + c1 * m[0]; m[1] = s1 * y1 + c1 * m[1]; m[2] = s1 * z1 + c1 * m[2];
    }
    void
    rotate_z(float a) {
        const float c1 = std::cos(a), s1 = std::sin(a);
        const float x1 = m[0], y1 = m[1], z1 = m[2];
        m[0] = c1 * x1 - s1 * m[3]; m[1] = c1 * y1 - s1 * m[4]; m[2] = c1 * z1 - s1 * m[5];
        m[3] = s1 * x1 + c1 * m[3]; m[4] = s1 * y1 + c1 * m[4]; m[5] = s1 * z1 + c1 * m[5];
    }
    static mat3
    euler(float a, float b, float c) {
        mat3 r;
        if (a != 0) r.rotate_x(a);
        if (b != 0) r.rotate_y(b);
        if (c != 0) r.rotate_z(c);
        return r;
    }
    mat3
    operator*(const mat3& o) const {
        mat3 r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[3 * i + j] = m[3 * i] * o.m[j] + m[3 * i + 1] * o.m[3 + j] + m[3 * i + 2] * o.m[6 + j];
        return r;
    }
};

// coordinate_transform(angles = {collimator, gantry, couch, iec2dicom} in degrees, position)
inline mat3
coordinate_rotation(const std::array<float, 4>& angles) {
    const float deg2rad         = M_PI / 180.0;
    const mat3  collimator      = mat3::euler(0, 0, angles[0] * deg2rad);
    const mat3  gantry          = mat3::euler(0, angles[1] * deg2rad, 0);
    const mat3  patient_support = mat3::euler(0, 0, angles[2] * deg2rad);
    const mat3  iec2dicom       = mat3::euler(angles[3] * deg2rad, 0, 0);
    return iec2dicom * patient_support * gantry * collimator;
}

// grid3d(xe_min, xe_max, n_xe): xe[i] = xe_min + i * dx in fp32
inline std::vector<float>
uniform_edges(float lo, float hi, int n_cells) {
    std::vector<float> e(n_cells + 1);
    const float        dx = (hi - lo) / n_cells;
    for (int i = 0; i <= n_cells; ++i) e[i] = lo + i * dx;
    return e;
}

// ---------------------------------------------------------------------------------------------
// output writers, byte-for-byte the reference's layouts (base/mqi_io.hpp)
// ---------------------------------------------------------------------------------------------
struct grid_desc {
    std::vector<float> xe, ye, ze;
    int nx() const { return (int) xe.size() - 1; }
    int ny() const { return (int) ye.size() - 1; }
    int nz() const { return (int) ze.size() - 1; }
};

// voxel spacing and centre of voxel (0, 0, 0) as the reference derives them from the first two edges: the spacing is
// a float difference, the centre is formed in double (the 0.5 literal) and rounded to float (mqi_io.hpp:497-503)
struct voxel_frame {
    float spacing[3], centre0[3];
    int   n[3];
};

inline voxel_frame
frame_of(const grid_desc& g) {
    voxel_frame              f;
    const std::vector<float>* e[3] = { &g.xe, &g.ye, &g.ze };
    for (int a = 0; a < 3; ++a) {
        f.spacing[a] = (*e[a])[1] - (*e[a])[0];
        f.centre0[a] = (float) ((double) (*e[a])[0] + (double) f.spacing[a] * 0.5);
        f.n[a]       = (int) e[a]->size() - 1;
    }
    return f;
}

// The MetaImage header of a dose file.  The text is an output contract: the two flavours the reference writes differ in
// more than the data line (mqi_io.hpp:493-591) -- the detached .mhd header omits '=' after TransformMatrix and
// CenterOfRotation, calls the origin "Offset" and prints floats with the stream's default precision; the .mha header
// says "Origin", adds HeaderSize and prints nine significant digits -- and readers of the reference's files get
// exactly those bytes (tests/golden/fmt_writers.npz).
inline std::string
meta_image_header(const voxel_frame& f, bool embedded, const std::string& data_file) {
    std::ostringstream h;
    if (embedded) h << std::setprecision(9);
    const char* sep = embedded ? " = " : " ";
    h << "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\nCompressedData = False\n"
      << "TransformMatrix" << sep << "1 0 0 0 1 0 0 0 1\n"
      << (embedded ? "Origin = " : "Offset ") << f.centre0[0] << " " << f.centre0[1] << " " << f.centre0[2] << "\n"
      << "CenterOfRotation" << sep << "0 0 0\n"
      << "AnatomicOrientation = RAI\n"
      << "DimSize = " << f.n[0] << " " << f.n[1] << " " << f.n[2] << "\n"
      << "ElementType = MET_DOUBLE\n";
    if (embedded) h << "HeaderSize = -1\n";
    h << "ElementSpacing = " << f.spacing[0] << " " << f.spacing[1] << " " << f.spacing[2] << "\n"
      << "ElementDataFile = " << data_file << "\n";
    return h.str();
}

// header text (may be empty) followed by length doubles times scale
inline void
write_dose_file(const std::string& path, const std::string& header, const double* src, double scale, size_t length) {
    std::vector<double> scaled;
    if (src && scale != 1.0) {
        scaled.assign(src, src + length);
        for (auto& v : scaled) v *= scale;
        src = scaled.data();
    }
    std::ofstream out(path, std::ios::out | std::ios::binary);
    if (!out) {
        std::cout << "Cannot write :" << path << std::endl;
        return;
    }
    out.write(header.data(), (std::streamsize) header.size());
    if (src) out.write(reinterpret_cast<const char*>(src), (std::streamsize) (length * sizeof(double)));
    out.close();
    if (!out.good()) std::cout << "Error occurred at writing time!" << std::endl;
}

inline void
save_to_bin(const double* src, double scale, const std::string& filepath, const std::string& filename, size_t length) {
    write_dose_file(filepath + "/" + filename + ".raw", "", src, scale, length);
}

inline void
save_to_mhd(const grid_desc& g, const double* src, double scale, const std::string& filepath, const std::string& filename,
            size_t length) {
    write_dose_file(filepath + "/" + filename + ".mhd", meta_image_header(frame_of(g), false, filename + ".raw"), nullptr, 1.0, 0);
    save_to_bin(src, scale, filepath, filename, length);
}

inline void
save_to_mha(const grid_desc& g, const double* src, double scale, const std::string& filepath, const std::string& filename,
            size_t length) {
    const voxel_frame f = frame_of(g);
    printf("x0 %.9g y0 %.9g z0 %.9g\n", f.centre0[0], f.centre0[1], f.centre0[2]);   // the reference prints the origin, :553
    write_dose_file(filepath + "/" + filename + ".mha", meta_image_header(f, true, "LOCAL"), src, scale, length);
}

// ---------------------------------------------------------------------------------------------
// phantom_env: box phantom from a raw int16 HU file, one uniform-square beamlet, one dose-to-water
// scorer "water_dE_total" (mqi_phantom_env.hpp).  --gpu_id may list several devices: histories are
// then sharded over them and the per-GPU dose grids summed with one NCCL reduce (mqi_reduce_dense).
// ---------------------------------------------------------------------------------------------
class phantom_env
{
public:
    std::vector<int>     gpu_ids;
    float                lxyz[3], pos[3];
    int                  nxyz[3];
    float                spot_position[3], spot_size[2], spot_energy[2];
    std::array<float, 4> spot_angles;
    long long            n_histories;
    std::string          output_path, phantom_path, output_format = "raw";
    int                  random_seed;
    int                  physics = MQI_PHYSICS_DEBUG;
    grid_desc            grid;
    std::vector<mqi_handle*> handles;
    std::vector<double>      dose;
    uint64_t                 tracked = 0;
    float                    kernel_ms = 0.f;

    static void
    check(int rc, const char* what) {
        if (rc < 0) throw std::runtime_error(std::string(what) + ": " + mqi_last_error());
    }

    explicit phantom_env(const cli& c) {
        auto f = [](const std::string& s) { return std::stof(s); };
        auto gid = c["--gpu_id"];
        if (gid.size() >= 1) {
            for (const auto& s : gid) gpu_ids.push_back(std::stoi(s));
        } else {
            gpu_ids.push_back(0);
            printf("gpu_id 0\n");
        }
        auto v = c["--lxyz"];
        if (v.size() >= 1) { lxyz[0] = f(v.at(0)); lxyz[1] = f(v.at(1)); lxyz[2] = f(v.at(2)); }
        else { lxyz[0] = 512.0; lxyz[1] = 512.0; lxyz[2] = 400.0; }
        v = c["--pxyz"];
        if (v.size() >= 1) { pos[0] = f(v.at(0)); pos[1] = f(v.at(1)); pos[2] = f(v.at(2)); }
        else { pos[0] = -256.0; pos[1] = 0.0; pos[2] = 0.0; }
        v = c["--nxyz"];
        if (v.size() >= 1) { nxyz[0] = std::stoi(v.at(0)); nxyz[1] = std::stoi(v.at(1)); nxyz[2] = std::stoi(v.at(2)); }
        else { nxyz[0] = 512; nxyz[1] = 512; nxyz[2] = 200; }
        v = c["--spot_position"];
        if (v.size() >= 1) { spot_position[0] = f(v.at(0)); spot_position[1] = f(v.at(1)); spot_position[2] = f(v.at(2)); }
        else { spot_position[0] = 1.0; spot_position[1] = 0.0; spot_position[2] = 0.0; }
        v = c["--spot_size"];
        if (v.size() >= 1) { spot_size[0] = f(v.at(0)); spot_size[1] = f(v.at(1)); }
        else { spot_size[0] = 0.0; spot_size[1] = 0.0; }
        v = c["--spot_energy"];
        if (v.size() >= 1) { spot_energy[0] = f(v.at(0)); spot_energy[1] = f(v.at(1)); }
        else { spot_energy[0] = 230; spot_energy[1] = 0.0; }
        v = c["--histories"];
        // the reference parses with stoi (int); scientific notation such as 1e8 is accepted here too
        n_histories = v.size() >= 1 ? (long long) std::llround(std::stod(v[0])) : 10000;
        v = c["--spot_angles"];
        if (v.size() >= 1) spot_angles = { f(v.at(0)), f(v.at(1)), f(v.at(2)), f(v.at(3)) };
        else spot_angles = { 0.f, 0.f, 0.f, 0.f };
        // --threads t [b]: launch shape of the reference kernel; the persistent grid of this build is
        // sized from the SM count, so the flag is accepted and ignored
        v = c["--output_prefix"];
        if (v.size() >= 1) { output_path = v[0]; printf("%s\n", output_path.c_str()); }
        else throw std::runtime_error("output_path is required.");
        v = c["--phantom_path"];
        if (v.size() >= 1) { phantom_path = v[0]; printf("phantom path: %s\n", output_path.c_str()); }
        else throw std::runtime_error("phantom_path is required.");
        v = c["--random_seed"];
        if (v.size() >= 1) { random_seed = std::stoi(v[0]); printf("random seed input %d\n", random_seed); }
        else random_seed = static_cast<int>(std::chrono::system_clock::now().time_since_epoch().count());
        printf("random seed %d\n", random_seed);
        // the reference never copies --output_format into output_format (always .raw); honoured here
        v = c["--output_format"];
        if (v.size() >= 1) output_format = v[0];
        v = c["--physics"];
        if (v.size() >= 1) {
            if (v[0] == "release") physics = MQI_PHYSICS_RELEASE;
            else if (v[0] == "debug") physics = MQI_PHYSICS_DEBUG;
            else throw std::runtime_error("--physics must be debug or release");
        }
    }

    ~phantom_env() {
        for (auto* h : handles) mqi_destroy(h);
    }

    // setup_world + setup_materials + setup_beamsource + upload (mqi_xenvironment.hpp:89-131)
    void
    initialize() {
        auto start = std::chrono::high_resolution_clock::now();
        grid.xe = uniform_edges(pos[0] - 0.5 * lxyz[0], pos[0] + 0.5 * lxyz[0], nxyz[0]);
        grid.ye = uniform_edges(pos[1] - 0.5 * lxyz[1], pos[1] + 0.5 * lxyz[1], nxyz[1]);
        grid.ze = uniform_edges(pos[2] - 0.5 * lxyz[2], pos[2] + 0.5 * lxyz[2], nxyz[2]);
        const size_t nvox = (size_t) nxyz[0] * nxyz[1] * nxyz[2];
        std::vector<int16_t> ph(nvox, 0);
        {
            std::ifstream ph_fid(phantom_path, std::ios::in | std::ios::binary);
            if (!ph_fid) throw std::runtime_error("cannot open phantom_path " + phantom_path);
            ph_fid.read(reinterpret_cast<char*>(ph.data()), nvox * sizeof(int16_t));
        }
        mqi_beamlet b {};
        b.phsp_uniform  = 1;   // phsp_6d_uniform
        b.energy_normal = 0;   // const_1d: the sigma of --spot_energy is ignored by the reference
        b.energy        = spot_energy[0];
        b.sigma_energy  = spot_energy[1];
        const float mean[6]  = { spot_position[0], spot_position[1], spot_position[2], 0.f, 0.f, -1.f };
        const float sigma[6] = { spot_size[0], spot_size[1], 0.f, 0.f, 0.f, 0.f };
        for (int i = 0; i < 6; ++i) { b.mean[i] = mean[i]; b.sigma[i] = sigma[i]; }
        const mat3 R = coordinate_rotation(spot_angles);
        for (int i = 0; i < 9; ++i) b.rot[i] = R.m[i];
        const uint64_t hist = (uint64_t) n_histories;
        printf("total histories %lu\n", (unsigned long) hist);
        for (int id : gpu_ids) {
            mqi_handle* h = nullptr;
            check(mqi_create(id, &h), "mqi_create");
            handles.push_back(h);
            check(mqi_set_physics(h, physics, 0), "mqi_set_physics");
            check(mqi_set_grid_hu(h, grid.xe.data(), (int) grid.xe.size(), grid.ye.data(), (int) grid.ye.size(),
                                  grid.ze.data(), (int) grid.ze.size(), ph.data(), 1.0f, nullptr, nullptr),
                  "mqi_set_grid_hu");
            check(mqi_add_scorer(h, MQI_SCORER_DOSE, "water_dE_total", nvox), "mqi_add_scorer");
            check(mqi_set_beamlets(h, &b, 1, &hist), "mqi_set_beamlets");
        }
        auto stop = std::chrono::high_resolution_clock::now();
        printf("Initialization for geometry done %f s\n", std::chrono::duration<double>(stop - start).count());
    }

    // run(): history range sharded over the devices, one launch each, then one reduce to device 0
    void
    run() {
        auto start = std::chrono::high_resolution_clock::now();
        printf("num spots %d\n", 1);
        const uint64_t n = (uint64_t) n_histories, g = handles.size();
        for (uint64_t r = 0; r < g; ++r) {
            const uint64_t first = n * r / g, last = n * (r + 1) / g;
            check(mqi_run_async(handles[r], (uint64_t) (int64_t) random_seed, first, last - first, 0), "mqi_run_async");
        }
        tracked   = 0;
        kernel_ms = 0.f;
        for (auto* h : handles) {
            mqi_run_stats st;
            check(mqi_get_run_stats(h, &st), "mqi_get_run_stats");
            tracked += st.histories;
            kernel_ms = std::max(kernel_ms, st.kernel_ms);
        }
        if (g > 1) check(mqi_reduce_dense(handles.data(), (int) g, 0, 0), "mqi_reduce_dense");
        printf("Number of particles tracked %lu\n", (unsigned long) tracked);
        auto stop = std::chrono::high_resolution_clock::now();
        printf("Run done %f s\n", std::chrono::duration<double>(stop - start).count());
        printf("Transport kernel %f ms on %d GPU(s): %e histories/s\n", kernel_ms, (int) g,
               kernel_ms > 0 ? 1e3 * (double) tracked / kernel_ms : 0.0);
    }

    // download_node + reshape_data (mqi_xenvironment.hpp:134-167)
    void
    finalize() {
        printf("finalizing\n");
        dose.resize((size_t) nxyz[0] * nxyz[1] * nxyz[2]);
        check(mqi_get_dense(handles[0], 0, dose.data(), 1.0), "mqi_get_dense");
    }

    // save_reshaped_files (:169-209): "<child>_<scorer>.<ext>", child index 0
    void
    save_reshaped_files() {
        auto start = std::chrono::high_resolution_clock::now();
        const std::string filename = "0_water_dE_total";
        if (!output_format.compare("mhd")) save_to_mhd(grid, dose.data(), 1.0, output_path, filename, dose.size());
        else if (!output_format.compare("mha")) save_to_mha(grid, dose.data(), 1.0, output_path, filename, dose.size());
        else save_to_bin(dose.data(), 1.0, output_path, filename, dose.size());
        auto stop = std::chrono::high_resolution_clock::now();
        printf("Reshape and save done %f s\n", std::chrono::duration<double>(stop - start).count());
    }
};

}   // namespace mqib
