// mqi_tps_host.hpp -- C++ host side of the treatment-planning front end (the reference's tps_env) over
// the C ABI of include/mqi_b200.h: input-parameter file, CT volume from .mha, generic PBS beam-model
// file, spot list, per-beam / per-spot / statistical-stopping loops, dense and sparse outputs.
// Header-only, no CUDA types.
//
// Mirrors (file:line under /root/reference/moqui):
//   mqi::file_parser                    base/mqi_file_handler.hpp:220-453    "Key value # comment" lines
//   tps_env<R> constructor / keys       base/environments/mqi_tps_env.hpp:155-370
//   tps_env::read_ct_image (.mha)       :615-701 ; CT edge construction :468-531
//   tps_env::setup_world (patient node) :712-962 ; setup_beamsource :965-993
//   initialize_and_run / run            :996-1036
//   run_by_beam / run_by_beam_stat / calculate_stat / calculate_average_results / run_by_spot
//                                       :1163-1240, :1242-1336, :1339-1426, :1428-1476, :1478-1603
//   save_reshaped_files / save_sparse_file  :1851-1928
//   mqi::pbs<T> (generic PBS machine)   base/mqi_treatment_machine_pbs.hpp:94-143, 238-276, 400-516
//   treatment_machine_ion::create_coordinate_transform / beam_starting_position
//                                       base/mqi_treatment_machine_ion.hpp:42-70, 324-333
//   io::save_to_npz / save_npz          base/mqi_io.hpp:249-320, base/mqi_sparse_io.hpp:281-399
//
// What is NOT here (stated in DESIGN.md): DICOM.  The reference reads the spot list, the beam angles
// and the isocentre from an RTPLAN through GDCM, which is neither vendored by the reference nor
// installed; this front end reads the same quantities from a small text plan (key `PlanFile`).
// Mask files (ScoringMask + Mask, StatROIMaskFilename) are supported: mask_reader
// (base/mqi_file_handler.hpp:13-217).  The RTSTRUCT options (ReadStructure + BodyContourName,
// StatROIStructFromRT + StatROI) read their contours from a text structure file (key `StructureFile`)
// instead of the DICOM RTSTRUCT and rasterise them like fill_contour (mqi_tps_env.hpp:1769-1826).
// Beamline children (range shifter, aperture block) are described in the text plan by the quantities
// characterize_rangeshifter / characterize_aperture (base/mqi_treatment_machine_pbs.hpp:279-331, 375-397)
// take from the RTPLAN, and built like create_rangeshifter / create_voxelized_aperture
// (base/environments/mqi_tps_env.hpp:1605-1736).
#pragma once

#include "mqi_host.hpp"
#include "mqi_roi.hpp"

#include <algorithm>
#include <cstring>
#include <ctime>
#include <limits>
#include <memory>
#include <map>
#include <sstream>
#include <strings.h>
#include <sys/stat.h>

namespace mqib
{

inline std::string
trim_copy(std::string s) {
    const char* ws = " \t\r\n";
    const size_t b = s.find_first_not_of(ws);
    if (b == std::string::npos) return "";
    const size_t e = s.find_last_not_of(ws);
    return s.substr(b, e - b + 1);
}

// ---------------------------------------------------------------------------------------------
// file_parser: every non-comment line is "Option<delimiter>value"; options are matched without
// regard to case, the first delimiter splits, text after '#' is dropped
// ---------------------------------------------------------------------------------------------
class file_parser
{
public:
    std::string              filename, delimiter;
    std::vector<std::string> lines;

    file_parser(const std::string& filename_, const std::string& delimiter_) : filename(filename_), delimiter(delimiter_) {
        std::ifstream fid(filename);
        if (!fid.is_open()) throw std::runtime_error("Cannot open input parameter file.");
        std::string line;
        while (std::getline(fid, line)) {
            line = trim_copy(line);
            if (line.empty() || line[0] == '#') continue;
            const size_t c = line.find('#');
            if (c != std::string::npos) line = line.substr(0, c);
            lines.push_back(line);
        }
    }

    std::string
    get_string(const std::string& option, const std::string& default_value) const {
        for (const auto& l : lines) {
            const size_t pos = l.find(delimiter);
            if (pos == std::string::npos) continue;
            if (strcasecmp(option.c_str(), trim_copy(l.substr(0, pos)).c_str()) == 0) return trim_copy(l.substr(pos + 1));
        }
        return default_value;
    }
    std::vector<std::string>
    get_string_vector(const std::string& option, const std::string& sep) const {
        std::vector<std::string> out;
        const std::string        value = get_string(option, "");
        size_t                   prev = 0, cur;
        while ((cur = value.find(sep, prev)) != std::string::npos) {
            const std::string t = trim_copy(value.substr(prev, cur - prev));
            if (!t.empty()) out.push_back(t);
            prev = cur + 1;
        }
        const std::string t = trim_copy(value.substr(prev));
        if (!t.empty()) out.push_back(t);
        return out;
    }
    float get_float(const std::string& o, float d) const { return (float) std::atof(get_string(o, std::to_string(d)).c_str()); }
    int   get_int(const std::string& o, int d) const { return std::atoi(get_string(o, std::to_string(d)).c_str()); }
    bool
    get_bool(const std::string& o, bool d) const {
        const std::string t = get_string(o, std::to_string(d));
        return strcasecmp(t.c_str(), "true") == 0 || std::atoi(t.c_str()) != 0;
    }
    std::vector<int>
    get_int_vector(const std::string& o, const std::string& sep) const {
        std::vector<int> out;
        for (const auto& s : get_string_vector(o, sep)) out.push_back(std::atoi(s.c_str()));
        return out;
    }
};

// ---------------------------------------------------------------------------------------------
// CT volume from a MetaImage (.mha, ElementDataFile = LOCAL).  The reference reads float voxels and casts them to
// int16 (mqi_tps_env.hpp:615-701); MET_SHORT -- the common CT export -- is accepted as well, anything else, a
// compressed payload or big-endian data is refused instead of being misread.  Values are clamped to int16 before
// the cast (an out-of-range float -> int16 conversion is undefined behaviour).
// ---------------------------------------------------------------------------------------------
struct ct_volume {
    int                  nx = 0, ny = 0, nz = 0;
    float                dx = 0, dy = 0, dz = 0;
    float                origin[3] = { 0, 0, 0 };   // centre of voxel (0,0,0)
    std::vector<int16_t> hu;
};

inline ct_volume
read_mha_ct(const std::string& path) {
    ct_volume     ct;
    std::ifstream fid(path, std::ios::binary);
    if (!fid) throw std::runtime_error("cannot open CT volume " + path);
    std::string line, element_type = "MET_FLOAT";
    bool        local = false;
    auto is_true = [](const std::string& v) { return strcasecmp(v.c_str(), "True") == 0 || v == "1"; };
    auto three = [](const std::string& v, double out[3]) {
        std::stringstream ss(v);
        ss >> out[0] >> out[1] >> out[2];
    };
    while (std::getline(fid, line)) {
        line = trim_copy(line);
        const size_t pos = line.find('=');
        if (pos == std::string::npos) continue;
        const std::string key = trim_copy(line.substr(0, pos)), value = trim_copy(line.substr(pos + 1));
        double            t[3] = { 0, 0, 0 };
        if (strcasecmp(key.c_str(), "Offset") == 0) {
            three(value, t);
            ct.origin[0] = (float) t[0]; ct.origin[1] = (float) t[1]; ct.origin[2] = (float) t[2];
        } else if (strcasecmp(key.c_str(), "ElementSpacing") == 0) {
            three(value, t);
            ct.dx = (float) t[0]; ct.dy = (float) t[1]; ct.dz = (float) t[2];
        } else if (strcasecmp(key.c_str(), "DimSize") == 0) {
            three(value, t);
            ct.nx = (int) t[0]; ct.ny = (int) t[1]; ct.nz = (int) t[2];
        } else if (strcasecmp(key.c_str(), "ElementType") == 0) {
            element_type = value;
        } else if (strcasecmp(key.c_str(), "CompressedData") == 0) {
            if (is_true(value)) throw std::runtime_error("compressed .mha CT volumes are not supported: " + path);
        } else if (strcasecmp(key.c_str(), "BinaryDataByteOrderMSB") == 0 || strcasecmp(key.c_str(), "ElementByteOrderMSB") == 0) {
            if (is_true(value)) throw std::runtime_error("big-endian .mha CT volumes are not supported: " + path);
        } else if (strcasecmp(key.c_str(), "ElementDataFile") == 0) {
            if (strcasecmp(value.c_str(), "LOCAL") != 0) throw std::runtime_error("Mask files does not contain data.");
            local = true;
            break;   // the voxel data follow this line
        }
    }
    if (!local || ct.nx <= 0 || ct.ny <= 0 || ct.nz <= 0) throw std::runtime_error("bad .mha header in " + path);
    const size_t n = (size_t) ct.nx * ct.ny * ct.nz;
    ct.hu.resize(n);
    if (strcasecmp(element_type.c_str(), "MET_SHORT") == 0) {
        fid.read(reinterpret_cast<char*>(ct.hu.data()), n * sizeof(int16_t));
        if ((size_t) fid.gcount() != n * sizeof(int16_t)) throw std::runtime_error("CT volume is truncated: " + path);
    } else if (strcasecmp(element_type.c_str(), "MET_FLOAT") == 0) {
        std::vector<float> tmp(n);
        fid.read(reinterpret_cast<char*>(tmp.data()), n * sizeof(float));
        if ((size_t) fid.gcount() != n * sizeof(float)) throw std::runtime_error("CT volume is truncated: " + path);
        for (size_t i = 0; i < n; ++i) {
            const float v = tmp[i];
            ct.hu[i]      = v != v ? (int16_t) 0 : (int16_t) std::min(32767.0f, std::max(-32768.0f, v));
        }
    } else {
        throw std::runtime_error("unsupported ElementType " + element_type + " in " + path + " (MET_FLOAT or MET_SHORT)");
    }
    return ct;
}

// mask_reader::read_mha_file (base/mqi_file_handler.hpp:38-99): uint8 voxels after the ElementDataFile =
// LOCAL line; only DimSize is looked at (it must match the CT here; the reference does not check)
inline std::vector<uint8_t>
read_mha_mask(const std::string& path, int nx, int ny, int nz) {
    std::ifstream fid(path, std::ios::binary);
    if (!fid) throw std::runtime_error("cannot open mask file " + path);
    std::string line;
    int         mx = 0, my = 0, mz = 0;
    bool        local = false;
    while (std::getline(fid, line)) {
        line = trim_copy(line);
        const size_t pos = line.find('=');
        if (pos == std::string::npos) continue;
        const std::string key = trim_copy(line.substr(0, pos)), value = trim_copy(line.substr(pos + 1));
        if (strcasecmp(key.c_str(), "DimSize") == 0) {
            std::stringstream ss(value);
            ss >> mx >> my >> mz;
        } else if (strcasecmp(key.c_str(), "ElementDataFile") == 0) {
            if (strcasecmp(value.c_str(), "LOCAL") != 0) throw std::runtime_error("Mask files does not contain data.");
            local = true;
            break;
        }
    }
    if (!local) throw std::runtime_error("bad .mha header in " + path);
    if (mx != nx || my != ny || mz != nz) throw std::runtime_error("mask " + path + " does not have the CT's dimensions");
    std::vector<uint8_t> m((size_t) nx * ny * nz);
    fid.read(reinterpret_cast<char*>(m.data()), m.size());
    if ((size_t) fid.gcount() != m.size()) throw std::runtime_error("mask file is truncated: " + path);
    return m;
}

// mask_reader::read_mask_files (:101-113): the 0/1 masks of all files are ADDED (overlaps give 2, which
// mask_to_roi treats as neither opening nor closing a run, see mqi_roi.hpp)
inline std::vector<uint8_t>
read_mask_files(const std::vector<std::string>& files, int nx, int ny, int nz) {
    if (files.empty()) throw std::runtime_error("Mask filelist are required for masking scorers.");
    std::vector<uint8_t> total((size_t) nx * ny * nz, 0);
    for (const auto& f : files) {
        printf("Reading maskfile %s\n", f.c_str());
        const std::vector<uint8_t> m = read_mha_mask(f, nx, ny, nz);
        for (size_t i = 0; i < total.size(); ++i) {
            if (m[i] > 1) throw std::runtime_error("mask values must be 0 or 1: " + f);
            total[i] = (uint8_t) (total[i] + m[i]);
        }
    }
    return total;
}

// ---------------------------------------------------------------------------------------------
// Text structure set (stands in for the RTSTRUCT):  [roi] name <text>  then one [contour] section per
// closed planar contour with rows "x y z" (mm, patient coordinates; ContourData of the RTSTRUCT).
// ---------------------------------------------------------------------------------------------
struct text_structures {
    typedef std::vector<std::array<float, 3>> contour_t;
    std::vector<std::pair<std::string, std::vector<contour_t>>> rois;

    static text_structures
    load(const std::string& path) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("RT STRCUTURE does not exist");   // the reference's message (:609)
        text_structures t;
        std::string     line, section;
        while (std::getline(f, line)) {
            line = trim_copy(line.substr(0, line.find_first_of('#')));
            if (line.empty()) continue;
            if (line.front() == '[' && line.back() == ']') {
                section = line.substr(1, line.size() - 2);
                std::transform(section.begin(), section.end(), section.begin(), ::tolower);
                if (section == "roi") t.rois.emplace_back();
                else if (section == "contour") {
                    if (t.rois.empty()) throw std::runtime_error("[contour] before [roi] in " + path);
                    t.rois.back().second.emplace_back();
                } else throw std::runtime_error("unknown structure section [" + section + "]");
                continue;
            }
            std::stringstream ss(line);
            if (section == "roi") {
                std::string k;
                ss >> k;
                if (k == "name") {
                    std::string rest;
                    std::getline(ss, rest);
                    t.rois.back().first = trim_copy(rest);
                }
            } else if (section == "contour") {
                std::array<float, 3> p { 0, 0, 0 };
                ss >> p[0] >> p[1] >> p[2];
                if (ss.fail()) throw std::runtime_error("bad contour row in " + path + ": " + line);
                t.rois.back().second.back().push_back(p);
            }
        }
        return t;
    }

    const std::vector<contour_t>*
    find(const std::string& name) const {   // ROIName matched without regard to case (:563)
        for (const auto& r : rois)
            if (strcasecmp(r.first.c_str(), name.c_str()) == 0) return &r.second;
        return nullptr;
    }
};

// fill_contour + sol1_1 (mqi_tps_env.hpp:1769-1826) for every contour of a roi: the contour's slice is
// the first i < nz - 1 with ze[i] < z < ze[i+1] (strict, and the last slab is never searched); inside it
// the pixel CENTRES (edge + half a pixel) of columns / rows [0, n - 1) are tested with the even-odd
// rule; hits are OR-ed (inner contours do not cut holes).  The volume starts from zeros here (the
// reference leaves it uninitialised).
inline std::vector<uint8_t>
rasterize_contours(const std::vector<text_structures::contour_t>& contours, int nx, int ny, int nz, const float* xe,
                   const float* ye, const float* ze, float dx, float dy) {
    std::vector<uint8_t> vol((size_t) nx * ny * nz, 0);
    for (const auto& c : contours) {
        if (c.empty()) continue;
        int z_ind = -1;
        for (int i = 0; i < nz - 1; i++) {
            if (c[0][2] > ze[i] && c[0][2] < ze[i + 1]) {
                z_ind = i;
                break;
            }
        }
        if (z_ind < 0) continue;
        const int n = (int) c.size();
        for (int x_ind = 0; x_ind < nx - 1; x_ind++) {
            for (int y_ind = 0; y_ind < ny - 1; y_ind++) {
                const float px = xe[x_ind] + dx * 0.5;
                const float py = ye[y_ind] + dy * 0.5;
                int         in = 0;
                for (int i = 0, j = n - 1; i < n; j = i++) {
                    const float x0 = c[i][0], y0 = c[i][1], x1 = c[j][0], y1 = c[j][1];
                    if ((((y0 <= py) && (py < y1)) || ((y1 <= py) && (py < y0))) && (px < (x1 - x0) * (py - y0) / (y1 - y0) + x0)) in = !in;
                }
                if (in) vol[(size_t) z_ind * nx * ny + (size_t) y_ind * nx + x_ind] = 1;
            }
        }
    }
    return vol;
}

// CT edges as read_dcm_dir builds them: first edge = voxel centre - half a voxel (+ robust shift),
// x / y edges = e0 + i * d, z edges by cumulative addition (all in fp32)
inline grid_desc
ct_edges(const ct_volume& ct, float shift_x, float shift_y, float shift_z) {
    grid_desc g;
    float xe0 = ct.origin[0] - ct.dx / 2.0;
    float ye0 = ct.origin[1] - ct.dy / 2.0;
    float ze0 = ct.origin[2] - ct.dz / 2.0;
    xe0 += shift_x; ye0 += shift_y; ze0 += shift_z;
    g.xe.resize(ct.nx + 1); g.ye.resize(ct.ny + 1); g.ze.resize(ct.nz + 1);
    for (int i = 0; i <= ct.nx; ++i) g.xe[i] = xe0 + i * ct.dx;
    for (int i = 0; i <= ct.ny; ++i) g.ye[i] = ye0 + i * ct.dy;
    for (int i = 0; i <= ct.nz; ++i) g.ze[i] = i == 0 ? ze0 : g.ze[i - 1] + ct.dz;
    return g;
}

// ---------------------------------------------------------------------------------------------
// Generic PBS machine: beam-model text file -> per-spot beamlet and history count
// ---------------------------------------------------------------------------------------------
struct plan_spot {
    float e = 0, x = 0, y = 0, meterset = 0;   // nominal energy [MeV], position at isocentre plane [mm], weight
};

class pbs_machine
{
public:
    struct spot_spec { float E = 0, dE = 0, x = 0, y = 0, xp = 0, yp = 0, ratio = 0; };
    float SAD[2]          = { 0, 0 };
    float rangeshifter[2] = { 0, 0 }, aperture[2] = { 0, 0 }, rangeshifter_snout_gap = 0;
    std::map<std::string, float> rangeshifter_thickness;
    std::map<float, spot_spec>   beamdata;
    std::map<std::string, float> time_spec;

    static float
    intpl(float x, float x0, float x1, float y0, float y1) {
        return (x1 == x0) ? y0 : y0 + (x - x0) * (y1 - y0) / (x1 - x0);
    }

    explicit pbs_machine(const std::string& f) {
        load_beamdata(f);
        // SAD 0 means a parallel beam (mqi_treatment_machine_pbs.hpp:105-118)
        if (SAD[0] == 0) SAD[0] = std::numeric_limits<float>::infinity();
        if (SAD[1] == 0) SAD[1] = std::numeric_limits<float>::infinity();
        if (beamdata.size() < 2) throw std::runtime_error("beam model " + f + " needs at least two [spot] rows");
    }

    // sections are recognised like in the reference: a section body ends at the next "[...]" line, and
    // that line is then tested by the following section checks of the same pass (B14: the file must
    // list its sections in the order geometry, rangeshifter_thickness, spot, time)
    void
    load_beamdata(const std::string& f) {
        std::ifstream file(f);
        if (!file) throw std::runtime_error("cannot open beam model file " + f);
        auto is_header = [](const std::string& l) {
            const size_t a = l.find('['), b = l.find(']');
            return a != std::string::npos && b != std::string::npos && a < b;
        };
        auto next = [&](std::string& line) -> bool {   // next non-empty line with comments stripped
            while (std::getline(file, line)) {
                line = line.substr(0, line.find_first_of('#'));
                if (line.size() == 0) continue;
                return true;
            }
            return false;
        };
        std::string line;
        while (next(line)) {
            if (line.compare("[geometry]") == 0) {
                while (next(line)) {
                    if (is_header(line)) break;
                    std::stringstream data(line);
                    std::string       key;
                    float             v1 = 0, v2 = 0;
                    data >> key >> v1 >> v2;
                    if (key == "SAD(mm)") { SAD[0] = v1; SAD[1] = v2; }
                    if (key == "rangeshifter(mm)") { rangeshifter[0] = v1; rangeshifter[1] = v2; }
                    if (key == "rangeshifter_snout_gap(mm)") rangeshifter_snout_gap = v1;
                    if (key == "aperture(mm)") { aperture[0] = v1; aperture[1] = v2; }
                }
            }
            if (line.compare("[rangeshifter_thickness]") == 0) {
                while (next(line)) {
                    if (is_header(line)) break;
                    const size_t from = line.find_first_of('"'), to = line.find_last_of('"');
                    if (from == std::string::npos || to == from) continue;
                    rangeshifter_thickness[line.substr(from + 1, to - from - 1)] = std::stof(line.substr(to + 1));
                }
            }
            if (line.compare("[spot]") == 0) {
                while (next(line)) {
                    if (is_header(line)) break;
                    std::stringstream data(line);
                    spot_spec         s;
                    float             e = 0;
                    data >> e >> s.E >> s.dE >> s.x >> s.y >> s.xp >> s.yp >> s.ratio;
                    if (data.fail()) continue;
                    beamdata.insert(std::make_pair(e, s));
                }
            }
            if (line.compare("[time]") == 0) {
                while (next(line)) {
                    if (is_header(line)) break;
                    std::stringstream data(line);
                    std::string       k;
                    float             v = 0;
                    data >> k >> v;
                    time_spec[k] = v;
                }
            }
        }
    }

    // the two table rows around a nominal energy: lower_bound(e) and its predecessor
    void
    bracket(float e, float& k_down, float& k_up, spot_spec& down, spot_spec& up) const {
        auto it = beamdata.lower_bound(e);
        if (it == beamdata.begin()) {   // the reference prints "out-of-bound" and dereferences prev(begin)
            std::cerr << "out-of-bound\n";
            ++it;
        }
        if (it == beamdata.end()) --it;
        auto dn = std::prev(it, 1);
        k_down = dn->first; k_up = it->first;
        down = dn->second; up = it->second;
    }

    // histories of a spot: meterset * ratio(E) / ParticlesPerHistory, truncated (characterize_history :134-143)
    size_t
    characterize_history(const plan_spot& s, float scale) const {
        float     kd, ku;
        spot_spec down, up;
        bracket(s.e, kd, ku, down, up);
        const float mid_ratio = intpl(s.e, kd, ku, down.ratio, up.ratio);
        return (size_t) (s.meterset * mid_ratio / scale);
    }

    // beamlet of a MODULATED spot in the beam frame (characterize_beamlet :238-276): energy ~ N(E, dE),
    // phase space phsp_6d around the point where the ray to (x, y, 0) crosses z = source_to_isocenter_mm
    mqi_beamlet
    characterize_beamlet(const plan_spot& s, float source_to_isocenter_mm) const {
        float     kd, ku;
        spot_spec down, up;
        bracket(s.e, kd, ku, down, up);
        const float mid_e  = intpl(s.e, kd, ku, down.E, up.E);
        const float mid_de = intpl(s.e, down.E, up.E, down.dE, up.dE);
        const float mid_x  = intpl(s.e, down.E, up.E, down.x, up.x);
        const float mid_y  = intpl(s.e, down.E, up.E, down.y, up.y);
        const float mid_xp = intpl(s.e, down.E, up.E, down.xp, up.xp);
        const float mid_yp = intpl(s.e, down.E, up.E, down.yp, up.yp);
        // beam_starting_position (mqi_treatment_machine_ion.hpp:324-333)
        const float z  = source_to_isocenter_mm;
        const float bx = s.x * (SAD[0] - z) / SAD[0];
        const float by = s.y * (SAD[1] - z) / SAD[1];
        float       dir[3] = { s.x - bx, s.y - by, 0.f - z };
        const float n      = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        dir[0] /= n; dir[1] /= n; dir[2] /= n;
        mqi_beamlet b {};
        b.phsp_uniform  = 0;   // phsp_6d
        b.energy_normal = 1;   // norm_1d
        b.energy        = mid_e;
        b.sigma_energy  = mid_de;
        const float mean[6]  = { bx, by, z, dir[0], dir[1], dir[2] };
        const float sigma[6] = { mid_x, mid_y, 0.f, mid_xp, mid_yp, 0.f };
        for (int i = 0; i < 6; ++i) { b.mean[i] = mean[i]; b.sigma[i] = sigma[i]; }
        b.corr[0] = b.corr[1] = 0.f;
        const float I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        for (int i = 0; i < 9; ++i) b.rot[i] = I[i];
        return b;
    }
};

// ---------------------------------------------------------------------------------------------
// Text plan (stands in for the RTPLAN the reference reads through GDCM).  Sections:
//   [plan]   name <text> | fractions <n>
//   [beam]   name <text> | gantry_angle <deg> | couch_angle <deg> | collimator_angle <deg> |
//            isocenter <x> <y> <z> | snout_position <mm>
//            optional beamline: rangeshifter_id <ID> [<ID> ...]  |  rangeshifter_wet <WET mm> <IsocenterToRangeShifterDistance mm>
//            | block_thickness <mm> | block_tray_distance <mm>
//   [block]  rows "x(mm) y(mm)": polygon of one aperture opening of the beam above (one section per block)
//   [spots]  rows "E(MeV nominal) x(mm) y(mm) meterset" of the beam above
// '#' starts a comment.  The quantities are the ones create_coordinate_transform, beam_module_ion and
// setup_beamsource take from the RTPLAN (BeamLimitingDeviceAngle, GantryAngle, PatientSupportAngle,
// IsocenterPosition, SnoutPosition, NominalBeamEnergy, ScanSpotPositionMap, ScanSpotMetersetWeights).
// ---------------------------------------------------------------------------------------------
struct plan_beam {
    std::string            name = "beam";
    float                  gantry = 0, couch = 0, collimator = 0, snout = 0;
    float                  iso[3] = { 0, 0, 0 };
    std::vector<plan_spot> spots;
    // beamline (RTPLAN: RangeShifterSequence / RangeShifterSettingsSequence / IonBlockSequence)
    std::vector<std::string> rangeshifter_ids;                          // RangeShifterID of every range shifter
    float                    rangeshifter_wet = 0, rangeshifter_distance = 0;   // RangeShifterWaterEquivalentThickness, IsocenterToRangeShifterDistance
    bool                     has_rangeshifter = false;
    float                    block_thickness = 0, block_tray_distance = 0;      // BlockThickness, IsocenterToBlockTrayDistance
    std::vector<std::vector<std::array<float, 2>>> block_data;          // BlockData of every block: (x, y) polygon points
};

struct text_plan {
    std::string            name = "plan";
    int                    fractions = 1;
    std::vector<plan_beam> beams;

    static text_plan
    load(const std::string& path) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("cannot open plan file " + path);
        text_plan   p;
        std::string line, section;
        while (std::getline(f, line)) {
            line = trim_copy(line.substr(0, line.find_first_of('#')));
            if (line.empty()) continue;
            if (line.front() == '[' && line.back() == ']') {
                section = line.substr(1, line.size() - 2);
                std::transform(section.begin(), section.end(), section.begin(), ::tolower);
                if (section == "beam") p.beams.emplace_back();
                else if ((section == "spots" || section == "block") && p.beams.empty()) throw std::runtime_error("[" + section + "] before [beam] in " + path);
                else if (section == "block") p.beams.back().block_data.emplace_back();
                else if (section != "plan" && section != "spots") throw std::runtime_error("unknown plan section [" + section + "]");
                continue;
            }
            std::stringstream ss(line);
            if (section == "plan") {
                std::string k;
                ss >> k;
                if (k == "name") ss >> p.name;
                else if (k == "fractions") ss >> p.fractions;
            } else if (section == "beam") {
                plan_beam&  b = p.beams.back();
                std::string k;
                ss >> k;
                if (k == "name") ss >> b.name;
                else if (k == "gantry_angle") ss >> b.gantry;
                else if (k == "couch_angle") ss >> b.couch;
                else if (k == "collimator_angle") ss >> b.collimator;
                else if (k == "isocenter") ss >> b.iso[0] >> b.iso[1] >> b.iso[2];
                else if (k == "snout_position") ss >> b.snout;
                else if (k == "rangeshifter_id") {
                    std::string id;
                    while (ss >> id) b.rangeshifter_ids.push_back(id);
                    b.has_rangeshifter = true;
                } else if (k == "rangeshifter_wet") {
                    ss >> b.rangeshifter_wet >> b.rangeshifter_distance;
                    b.has_rangeshifter = true;
                } else if (k == "block_thickness") ss >> b.block_thickness;
                else if (k == "block_tray_distance") ss >> b.block_tray_distance;
                else throw std::runtime_error("unknown beam key " + k + " in " + path);
            } else if (section == "block") {
                std::array<float, 2> xy { 0.f, 0.f };
                ss >> xy[0] >> xy[1];
                if (ss.fail()) throw std::runtime_error("bad block row in " + path + ": " + line);
                p.beams.back().block_data.back().push_back(xy);
            } else if (section == "spots") {
                plan_spot s;
                ss >> s.e >> s.x >> s.y >> s.meterset;
                if (ss.fail()) throw std::runtime_error("bad spot row in " + path + ": " + line);
                p.beams.back().spots.push_back(s);
            } else {
                throw std::runtime_error("text outside a section in " + path);
            }
        }
        if (p.beams.empty()) throw std::runtime_error("plan has no beam: " + path);
        return p;
    }
};

// ---------------------------------------------------------------------------------------------
// .npz (uncompressed zip of .npy members) for the sparse Dij output: scipy.sparse CSR members
// indices / indptr / shape / data / format exactly as save_to_npz names and types them
// (uint32 indices and indptr, uint32 shape = [n_spots, n_voxels], float64 data).  format.npy holds
// the 3-byte string "csr" (the reference writes 96 bytes of which numpy reads 3: B12).
// ---------------------------------------------------------------------------------------------
inline uint32_t
crc32_update(uint32_t crc, const void* data, size_t n) {
    static uint32_t table[256];
    static bool     init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    const unsigned char* p = static_cast<const unsigned char*>(data);
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

class npz_writer
{
    std::ofstream     out_;
    std::vector<char> central_;
    uint16_t          n_ = 0;
    template<typename T> static void
    put(std::vector<char>& v, T x) {
        const char* p = reinterpret_cast<const char*>(&x);
        v.insert(v.end(), p, p + sizeof(T));
    }
    static void put(std::vector<char>& v, const std::string& s) { v.insert(v.end(), s.begin(), s.end()); }

public:
    explicit npz_writer(const std::string& path) : out_(path, std::ios::binary) {
        if (!out_) throw std::runtime_error("cannot write " + path);
    }
    // one .npy member: descr e.g. "<u4", "<f8", "|S3"; shape text e.g. "12,", "" (0-d)
    void
    add(const std::string& name, const std::string& descr, const std::string& shape, const void* data, size_t nbytes) {
        std::string dict = "{'descr': '" + descr + "', 'fortran_order': False, 'shape': (" + shape + "), }";
        dict.append(16 - (10 + dict.size() + 1) % 16, ' ');
        dict.push_back('\n');
        std::vector<char> npy;
        npy.push_back((char) 0x93);
        put(npy, std::string("NUMPY"));
        npy.push_back(1); npy.push_back(0);
        put(npy, (uint16_t) dict.size());
        put(npy, dict);
        const uint64_t total = npy.size() + nbytes;
        if (total > 0xffffffffull) throw std::runtime_error("npz member " + name + " exceeds 4 GiB (zip64 is not written)");
        uint32_t crc = crc32_update(0, npy.data(), npy.size());
        crc          = crc32_update(crc, data, nbytes);
        const uint32_t    offset = (uint32_t) out_.tellp();
        std::vector<char> local;
        put(local, std::string("PK")); put(local, (uint16_t) 0x0403);
        put(local, (uint16_t) 20); put(local, (uint16_t) 0); put(local, (uint16_t) 0); put(local, (uint16_t) 0); put(local, (uint16_t) 0);
        put(local, crc); put(local, (uint32_t) total); put(local, (uint32_t) total);
        put(local, (uint16_t) name.size()); put(local, (uint16_t) 0);
        put(local, name);
        out_.write(local.data(), local.size());
        out_.write(npy.data(), npy.size());
        out_.write(static_cast<const char*>(data), nbytes);
        put(central_, std::string("PK")); put(central_, (uint16_t) 0x0201); put(central_, (uint16_t) 20);
        central_.insert(central_.end(), local.begin() + 4, local.begin() + 30);
        put(central_, (uint16_t) 0); put(central_, (uint16_t) 0); put(central_, (uint16_t) 0); put(central_, (uint32_t) 0);
        put(central_, offset);
        put(central_, name);
        ++n_;
    }
    void
    close() {
        const uint32_t    cd_offset = (uint32_t) out_.tellp();
        std::vector<char> footer;
        put(footer, std::string("PK")); put(footer, (uint16_t) 0x0605);
        put(footer, (uint16_t) 0); put(footer, (uint16_t) 0); put(footer, n_); put(footer, n_);
        put(footer, (uint32_t) central_.size()); put(footer, cd_offset); put(footer, (uint16_t) 0);
        out_.write(central_.data(), central_.size());
        out_.write(footer.data(), footer.size());
        out_.close();
    }
};

// triplets (voxel, spot, value) in table-slot order -> CSR [n_spots x n_voxels]; within a row the
// columns keep the slot order, as in the reference (save_to_npz scans the table once and appends)
inline void
save_csr_npz(const std::string& path, uint32_t n_spots, uint32_t n_vox, const std::vector<uint32_t>& vox,
             const std::vector<uint32_t>& spot, const std::vector<double>& val) {
    std::vector<uint32_t> indptr(n_spots + 1, 0);
    for (size_t i = 0; i < spot.size(); ++i) {
        if (spot[i] >= n_spots) throw std::runtime_error("Dij entry with a spot index beyond the beam's spots");
        ++indptr[spot[i] + 1];
    }
    for (uint32_t s = 0; s < n_spots; ++s) indptr[s + 1] += indptr[s];
    std::vector<uint32_t> indices(vox.size());
    std::vector<double>   data(vox.size());
    std::vector<uint32_t> cursor(indptr.begin(), indptr.end() - 1);
    for (size_t i = 0; i < vox.size(); ++i) {
        const uint32_t o = cursor[spot[i]]++;
        indices[o] = vox[i];
        data[o]    = val[i];
    }
    const uint32_t shape[2] = { n_spots, n_vox };
    npz_writer     z(path);
    z.add("indices.npy", "<u4", std::to_string(indices.size()) + ",", indices.data(), indices.size() * 4);
    z.add("indptr.npy", "<u4", std::to_string(indptr.size()) + ",", indptr.data(), indptr.size() * 4);
    z.add("shape.npy", "<u4", "2,", shape, 8);
    z.add("data.npy", "<f8", std::to_string(data.size()) + ",", data.data(), data.size() * 8);
    z.add("format.npy", "|S3", "", "csr", 3);
    z.close();
}

// ---------------------------------------------------------------------------------------------
// tps_env
// ---------------------------------------------------------------------------------------------
enum sim_type_t { PER_BEAM = 0, PER_SPOT = 1 };

class tps_env
{
public:
    // ---- input parameters (same keys, defaults and checks as the reference constructor)
    std::vector<int> gpu_ids;
    int              master_seed = 0;
    bool             use_absolute_path = false;
    std::string      beam_prefix, parent_dir, dicom_dir, ct_name, ct_path, plan_path;
    long long        max_histories_per_batch = 0;
    std::string      source_type, machine_name, calibration_name;
    sim_type_t       sim_type = PER_BEAM;
    float            particles_per_history = -1.f, rbe = 1.1f, rangeshifter_density = 1.19f;
    int              n_fractions = -1, unit_weights = -1;
    std::vector<std::string> scorer_string;
    float            density_scale = 1.f, shift[3] = { 0, 0, 0 };
    std::string      output_path, output_format;
    bool             sparse_output = false, overwrite_results = false;
    bool             record_statistics = false, save_statistics = false;
    float            stat_criteria = -1.f, stat_threshold = 0.f;
    std::vector<int> beam_numbers;   // 1-based, like the reference's bnb
    bool             scoring_mask = false, save_scorer_map = false;
    std::vector<std::string> mask_filenames, stat_roi_mask_filenames;
    std::string      scorer_map_prefix;
    bool             read_structure = false, set_stat_roi_from_rtstruct = false;
    std::string      body_contour_name, stat_roi, structure_path;
    text_structures  structures;
    bool             reference_quirks = false;   // extension: reproduce B2 (double scoring with >= 3 scorers)
    int              max_stat_passes = 1000;     // extension: bound on the stopping loop
    uint64_t         dij_capacity = 0;           // extension: slots of the Dij table (0 = sized from the work, at most the reference's 393 216 000)

    // ---- data
    ct_volume   ct;
    grid_desc   grid;
    text_plan   plan;
    std::unique_ptr<pbs_machine> machine;

    // ---- per-beam state
    struct scorer_slot { int id; int kind; std::string name; bool save; };
    std::vector<mqi_handle*> handles;
    std::vector<scorer_slot> scorers;
    int                      stat_sum = -1, stat_sumsq = -1;
    std::vector<uint64_t>    spot_histories;
    uint64_t                 total_histories = 0, tracked = 0;
    uint32_t                 num_spots = 0;
    int                      n_beamline = 0;          // children of the world in front of the patient grid
    uint64_t                 scoring_roi_size = 0, stat_roi_size = 0;
    int                      bnb = 0;
    float                    sid = 0.f;
    float                    last_stat_percent = 100.f;
    int                      stat_passes = 0;
    float                    kernel_ms_total = 0.f;

    static void
    check(int rc, const char* what) {
        if (rc < 0) throw std::runtime_error(std::string(what) + ": " + mqi_last_error());
    }

    explicit tps_env(const std::string& input_name) {
        file_parser parser(input_name, " ");
        // GPUID: one id like the reference, or a comma list (histories / spots sharded over the devices)
        for (int id : parser.get_int_vector("GPUID", ",")) gpu_ids.push_back(id);
        if (gpu_ids.empty()) gpu_ids.push_back(0);
        master_seed = parser.get_int("RandomSeed", -1);
        if (master_seed == -1) master_seed = (int) std::time(nullptr);
        printf("master seed %d\n", master_seed);
        use_absolute_path       = parser.get_bool("UseAbsolutePath", false);
        beam_prefix             = parser.get_string("BeamPrefix", "beam");
        max_histories_per_batch = parser.get_int("MaxHistoriesPerBatch", 0);
        parent_dir              = parser.get_string("ParentDir", "");
        if (parent_dir.empty()) throw std::runtime_error("ParentDir is not provided");
        ct_name = parser.get_string("CTVolumeName", "");
        const std::string plan_name = parser.get_string("PlanFile", "");
        if (use_absolute_path) {
            dicom_dir = parser.get_string("DicomDir", "");
            ct_path   = ct_name;
            plan_path = plan_name;
        } else {
            dicom_dir = parent_dir + "/" + parser.get_string("DicomDir", "");
            ct_path   = parent_dir + "/" + ct_name;
            plan_path = parent_dir + "/" + plan_name;
        }
        if (ct_name.empty())
            throw std::runtime_error("CTVolumeName (.mha) is required: reading a DICOM CT series needs GDCM, which this build does not have");
        if (plan_name.empty())
            throw std::runtime_error("PlanFile (text plan) is required: reading an RTPLAN needs GDCM, which this build does not have");

        source_type = parser.get_string("SourceType", "FluenceMap");
        const std::string st = parser.get_string("SimulationType", "perBeam");
        if (strcasecmp(st.c_str(), "perBeam") == 0) sim_type = PER_BEAM;
        else if (strcasecmp(st.c_str(), "perSpot") == 0) sim_type = PER_SPOT;
        else throw std::runtime_error("SimulationType must be perBeam or perSpot");
        particles_per_history = parser.get_float("ParticlesPerHistory", -1.0);
        rbe                   = parser.get_float("RBE", 1.1);
        n_fractions           = parser.get_int("NumberOfFraction", -1);
        rangeshifter_density  = parser.get_float("RangeshifterDensity", 1.19);
        unit_weights          = parser.get_int("UnitWeights", -1);
        machine_name          = parser.get_string("Machine", "");
        calibration_name      = parser.get_string("Calibration", "default");
        scorer_string         = parser.get_string_vector("Scorer", ",");
        for (const auto& s : scorer_string) {
            static const char* known[] = { "EnergyDeposition", "Dose", "LETd", "LETt", "Dij", "TrackLength" };
            bool ok = false;
            for (const char* k : known) ok = ok || strcasecmp(s.c_str(), k) == 0;
            if (!ok) throw std::runtime_error("Unrecognized scorer name");
            if (strcasecmp(s.c_str(), "Dij") == 0) {   // Dij forces per-spot simulation and must be alone (:213-220)
                sim_type = PER_SPOT;
                if (scorer_string.size() > 1) throw std::runtime_error("Dij cannot be scored with the other quantities");
            }
        }
        body_contour_name          = parser.get_string("BodyContourName", "External");   // :225-226
        read_structure             = parser.get_bool("ReadStructure", false);
        set_stat_roi_from_rtstruct = parser.get_bool("StatROIStructFromRT", false);      // :290-296
        if (set_stat_roi_from_rtstruct) {
            stat_roi = parser.get_string("StatROI", "");
            if (stat_roi.empty()) throw std::runtime_error("Statistical ROI is not provided.");
        }
        if (read_structure || set_stat_roi_from_rtstruct) {
            std::string sf = parser.get_string("StructureFile", "");
            if (sf.empty())
                throw std::runtime_error("StructureFile (text structure set) is required: reading an RTSTRUCT needs GDCM, which this build does not have");
            if (!use_absolute_path && sf[0] != '/') sf = parent_dir + "/" + sf;
            structure_path = sf;
        }
        scoring_mask = parser.get_bool("ScoringMask", false);   // :224-241
        if (scoring_mask) {
            save_scorer_map = parser.get_bool("SaveMap", true);
            mask_filenames  = parser.get_string_vector("Mask", ",");
            if (mask_filenames.empty()) throw std::runtime_error("Mask filename is missing");
        }
        if (save_scorer_map) scorer_map_prefix = parser.get_string("ScorerMapName", "scorer_map");
        stat_roi_mask_filenames = parser.get_string_vector("StatROIMaskFilename", ",");   // :289
        density_scale = parser.get_float("DensityScaling", 1);
        shift[0]      = parser.get_float("XShift", 0);
        shift[1]      = parser.get_float("YShift", 0);
        shift[2]      = parser.get_float("ZShift", 0);
        output_path   = parser.get_string("OutputDir", "");
        output_format = parser.get_string("OutputFormat", "raw");
        sparse_output = strcasecmp(output_format.c_str(), "npz") == 0;
        if (output_path.empty()) throw std::runtime_error("Output directory is not provided.");
        overwrite_results = parser.get_bool("OverwriteResults", false);
        struct stat info;
        if (stat(output_path.c_str(), &info) != 0) mkdir(output_path.c_str(), 0755);
        else if (!overwrite_results) throw std::runtime_error("Output directory exists.");
        record_statistics = parser.get_bool("StoppingStatistics", false);
        save_statistics   = parser.get_bool("SaveStoppingStatistics", false);
        if (record_statistics) {
            stat_criteria = parser.get_float("StoppingCriteria", -1.0);
            if (stat_criteria < 0) throw std::runtime_error("Statistical criteria must be positive float");
        }
        stat_threshold = parser.get_float("StatThreshold", 0.0);
        if (record_statistics && stat_threshold <= 0) {   // :299-315
            if (!set_stat_roi_from_rtstruct && stat_roi_mask_filenames.empty())
                throw std::runtime_error("If no contour or mask is selected, the statical threshold cannot be zero");
            printf("If statistical threshold is zero, the uncertainty might be biased\n");
        }
        // accepted and without effect here: TotalThreads sizes the reference's launch (<<<threads/512, 512>>>, :1107);
        // this path launches one persistent CTA per SM whatever the history count.  ScoreToCTGrid is forced to
        // true by the reference itself (:243).
        (void) parser.get_int("TotalThreads", -1);
        (void) parser.get_bool("ScoreToCTGrid", true);
        reference_quirks = parser.get_bool("ReferenceQuirks", false);
        max_stat_passes  = parser.get_int("MaxStatPasses", 1000);
        dij_capacity     = (uint64_t) std::strtoull(parser.get_string("DijCapacity", "0").c_str(), nullptr, 10);
        std::cout << parent_dir << std::endl;

        // ---- data: CT, plan, machine (treatment_session::create_machine accepts only "pbs:<file>", ts:129-166)
        ct   = read_mha_ct(ct_path);
        grid = ct_edges(ct, shift[0], shift[1], shift[2]);
        printf("dcm.dim nx %d ny %d nz %d\n", ct.nx, ct.ny, ct.nz);
        plan = text_plan::load(plan_path);
        printf("%s\n", plan_path.c_str());
        if (!structure_path.empty()) {
            printf("Loading RTSTRUCT from %s\n", structure_path.c_str());
            structures = text_structures::load(structure_path);
            printf("roi seq size %lu\n", (unsigned long) structures.rois.size());
        }
        {
            const size_t deli = machine_name.find(":");
            std::string  site = machine_name.substr(0, deli);
            std::transform(site.begin(), site.end(), site.begin(), ::tolower);
            std::string cal = calibration_name;
            std::transform(cal.begin(), cal.end(), cal.begin(), ::tolower);
            if (deli == std::string::npos || site != "pbs") throw std::runtime_error("Valid machine is not available.");
            if (cal != "default") throw std::runtime_error("Unknown calibration method. annony");
            std::string model = machine_name.substr(deli + 1);
            if (!use_absolute_path && !model.empty() && model[0] != '/') model = parent_dir + "/" + model;
            std::cout << "Creating a generic PBS machine from : " << model << "\n";
            machine.reset(new pbs_machine(model));
        }
        if (n_fractions <= 0) n_fractions = plan.fractions;
        printf("n fraction %d rbe %f\n", n_fractions, rbe);
        // BeamNumbers: empty or "0" selects every beam that is not called "Setup" (:320-368)
        beam_numbers = parser.get_int_vector("BeamNumbers", ",");
        if (beam_numbers.empty() || (beam_numbers.size() == 1 && beam_numbers[0] == 0)) {
            beam_numbers.clear();
            for (int k = 1; k <= (int) plan.beams.size(); ++k)
                if (plan.beams[k - 1].name != "Setup") beam_numbers.push_back(k);
        }
        for (int k : beam_numbers)
            if (k < 1 || k > (int) plan.beams.size()) throw std::runtime_error("BeamNumbers entry out of range");
    }

    ~tps_env() { release(); }

    void
    release() {
        for (auto* h : handles) mqi_destroy(h);
        handles.clear();
    }

    // coordinate transform of a beam: angles {collimator, gantry, -couch, iec2dicom = 90}, translation =
    // isocentre (create_coordinate_transform tmi:42-70, setup_beamsource :975-983)
    static mat3
    beam_rotation(const plan_beam& b) {
        return coordinate_rotation({ b.collimator, b.gantry, -1.0f * b.couch, 90.0f });
    }

    // beamlets + histories per spot of beam bnb (create_beamsource tmi:83-134)
    void
    build_source(const plan_beam& b, std::vector<mqi_beamlet>& beamlets, std::vector<uint64_t>& histories) const {
        const mat3 R = beam_rotation(b);
        beamlets.clear();
        histories.clear();
        for (const auto& s : b.spots) {
            mqi_beamlet bl = machine->characterize_beamlet(s, sid);
            for (int i = 0; i < 9; ++i) bl.rot[i] = R.m[i];
            for (int i = 0; i < 3; ++i) bl.trans[i] = b.iso[i];
            size_t n = (particles_per_history == -1.f) ? 1 : machine->characterize_history(s, particles_per_history);
            if (sim_type == PER_SPOT && unit_weights > 0) n = (size_t) unit_weights;   // run_by_spot :1488-1491,1532-1536
            beamlets.push_back(bl);
            histories.push_back((uint64_t) n);
        }
    }

    // One beamline child of the world in the beam frame (create_beamline tmi:238-276 + the builders
    // create_rangeshifter / create_voxelized_aperture of the environment)
    struct beamline_node {
        std::vector<float> xe, ye, ze, rho;
        float              pos_z = 0;
    };

    // characterize_rangeshifter (pbs:279-331) + create_rangeshifter (:1605-1649): a slab of one voxel,
    // grid3d(pos - vol/2, pos + vol/2, 2 edges per axis), density RangeshifterDensity * 1e-3 g/mm^3
    beamline_node
    make_rangeshifter(const plan_beam& b) const {
        float lx = machine->rangeshifter[0], ly = machine->rangeshifter[1], lz = 0.f, pz = 0.f;
        if (machine->rangeshifter_thickness.empty()) {
            std::cout << "Thickness is defined from RangeShifter Setting sequence\n";
            lz = b.rangeshifter_wet / 1.15;
            pz = b.rangeshifter_distance;
            pz -= lz;
        } else {
            std::cout << "Thickness is defined from Rangeshifter ID\n";
            pz = b.snout;
            for (const auto& id : b.rangeshifter_ids) {
                auto it = machine->rangeshifter_thickness.find(trim_copy(id));
                if (it != machine->rangeshifter_thickness.end()) lz += it->second;
            }
            if (!(lz > 0)) throw std::runtime_error("range shifter thickness is zero (unknown RangeShifterID?)");
            pz -= (lz * 0.5 + machine->rangeshifter_snout_gap);
        }
        if (!(lx > 0) || !(ly > 0) || !(lz > 0)) throw std::runtime_error("range shifter needs rangeshifter(mm) in the beam model and a positive thickness");
        std::cout << "Range shifter thickness: " << lz << " (mm) and position: " << pz << " (mm)" << std::endl;
        beamline_node n;
        auto two = [](float lo, float hi) {   // grid3d(e_min, e_max, n_e = 2): e_min + i * (e_max - e_min) / 1
            const float d = (hi - lo) / 1;
            return std::vector<float> { lo + 0 * d, lo + 1 * d };
        };
        n.xe = two(0.f - lx / 2, 0.f + lx / 2);
        n.ye = two(0.f - ly / 2, 0.f + ly / 2);
        n.ze = two(pz - lz / 2, pz + lz / 2);
        printf("Rangeshifter density %.4f\n", rangeshifter_density * 1e-3);
        n.rho.assign(1, (float) (rangeshifter_density * 1e-3));
        n.pos_z = pz;
        return n;
    }

    // is_inside / sol1_1 (:1738-1768): even-odd test of the voxel centre against the block polygon.  The
    // reference overwrites `inside` per opening, so with several blocks only the LAST one counts (kept).
    static bool
    inside_block(float x, float y, const std::vector<std::vector<std::array<float, 2>>>& blocks) {
        bool inside = false;
        for (const auto& seg : blocks) {
            int          c = 0;
            const size_t n = seg.size();
            for (size_t i = 0, j = n - 1; i < n; j = i++) {
                const float x0 = seg[i][0], y0 = seg[i][1], x1 = seg[j][0], y1 = seg[j][1];
                if ((((y0 <= y) && (y < y1)) || ((y1 <= y) && (y < y0))) && (x < (x1 - x0) * (y - y0) / (y1 - y0) + x0)) c = !c;
            }
            inside = c != 0;
        }
        return inside;
    }

    // characterize_aperture (pbs:375-397) + create_voxelized_aperture (:1651-1736): 1 mm voxels, 1e-8 g/mm^3
    // where the voxel centre lies inside the opening, 100 elsewhere
    beamline_node
    make_aperture(const plan_beam& b) const {
        const float lx = machine->aperture[0], ly = machine->aperture[1], lz = b.block_thickness;
        const float pz = b.block_tray_distance + lz * 0.5f;
        if (!(lx > 0) || !(ly > 0) || !(lz > 0)) throw std::runtime_error("aperture needs aperture(mm) in the beam model and block_thickness in the plan");
        std::cout << "BlockThickness and Position (center) : " << lz << ", " << pz << " mm" << std::endl;
        const size_t nx = (size_t) std::ceil(lx / 1.0f), ny = (size_t) std::ceil(ly / 1.0f), nz = (size_t) std::ceil(lz / 1.0f);
        beamline_node n;
        n.xe.resize(nx + 1); n.ye.resize(ny + 1); n.ze.resize(nz + 1);
        for (size_t i = 0; i <= nx; ++i) n.xe[i] = (0.f - lx / 2) + i * 1.0f;
        for (size_t i = 0; i <= ny; ++i) n.ye[i] = (0.f - ly / 2) + i * 1.0f;
        for (size_t i = 0; i <= nz; ++i) n.ze[i] = (pz - lz / 2) + i * 1.0f;
        n.rho.resize(nx * ny * nz);
        for (size_t i = 0; i < nx; ++i) {
            const float x = n.xe[i] + 1.0f * 0.5;
            for (size_t j = 0; j < ny; ++j) {
                const float y    = n.ye[j] + 1.0f * 0.5;
                const float rho  = inside_block(x, y, b.block_data) ? 1e-8f : 100.0f;
                for (size_t k = 0; k < nz; ++k) n.rho[k * nx * ny + j * nx + i] = rho;
            }
        }
        n.pos_z = pz;
        return n;
    }

    // create_beamline (tmi:238-276): range shifter first, then the block, sorted by z descending (upstream first)
    std::vector<beamline_node>
    build_beamline(const plan_beam& b) const {
        std::vector<beamline_node> v;
        std::cout << "number of range shifter: " << (b.has_rangeshifter ? 1 : 0) << std::endl;
        if (b.has_rangeshifter) {
            v.push_back(make_rangeshifter(b));
            printf("RANGE SHIFTER added\n");
        }
        std::cout << "number of blocks: " << b.block_data.size() << std::endl;
        if (!b.block_data.empty()) {
            v.push_back(make_aperture(b));
            printf("APERTURE added\n");
        }
        std::stable_sort(v.begin(), v.end(), [](const beamline_node& a, const beamline_node& c) { return a.pos_z > c.pos_z; });
        return v;
    }

    // the summed mask volumes behind the scoring roi and the stat roi (setup_world :771-815): mask files,
    // else the rasterised body / StatROI contour, else empty (= roi_t(DIRECT))
    std::vector<uint8_t>
    contour_mask(const std::string& name) const {
        const auto* c = structures.find(name);
        if (!c) throw std::runtime_error("structure " + name + " is not in the structure file");
        printf("%s\n", name.c_str());
        return rasterize_contours(*c, ct.nx, ct.ny, ct.nz, grid.xe.data(), grid.ye.data(), grid.ze.data(), ct.dx, ct.dy);
    }
    void
    roi_masks(std::vector<uint8_t>& scoring, std::vector<uint8_t>& stat) const {
        scoring.clear();
        stat.clear();
        if (scoring_mask) scoring = read_mask_files(mask_filenames, ct.nx, ct.ny, ct.nz);
        else if (read_structure) scoring = contour_mask(body_contour_name);
        if (record_statistics) {
            if (set_stat_roi_from_rtstruct) stat = contour_mask(stat_roi);
            else if (!stat_roi_mask_filenames.empty()) stat = read_mask_files(stat_roi_mask_filenames, ct.nx, ct.ny, ct.nz);
            else printf("Statistical ROI is set to entire patient.\n");
        }
    }

    // initialize(): setup_world + setup_materials + setup_beamsource + upload, per device
    void
    initialize() {
        const plan_beam& b = plan.beams[bnb - 1];
        printf("There are %d beams\n", (int) plan.beams.size());
        printf("Selecting %d: %s\n", bnb, b.name.c_str());
        sid = b.snout + 50;
        printf("sid %f\n", sid);
        std::vector<mqi_beamlet> beamlets;
        build_source(b, beamlets, spot_histories);
        num_spots       = (uint32_t) beamlets.size();
        total_histories = 0;
        for (auto n : spot_histories) total_histories += n;
        if (sim_type == PER_SPOT && unit_weights > 0) particles_per_history = 1;
        printf("beamlets %lu\n", (unsigned long) num_spots);
        printf("histories %lu\n", (unsigned long) total_histories);
        printf("dcm dim %d %d %d\n", ct.nx, ct.ny, ct.nz);

        release();
        scorers.clear();
        stat_sum = stat_sumsq = -1;
        const uint64_t nvox = (uint64_t) ct.nx * ct.ny * ct.nz;
        // beamline children in the beam's frame: rotation / translation of the coordinate transform with
        // iec2dicom = 90 (setup_world :736-737, create_rangeshifter :1611-1618)
        const std::vector<beamline_node> beamline = build_beamline(b);
        const mat3                       R        = beam_rotation(b);
        n_beamline = (int) beamline.size();
        // regions of interest (setup_world :771-815)
        std::vector<uint8_t> scoring_mask_total, stat_mask_total;
        roi_masks(scoring_mask_total, stat_mask_total);
        for (size_t d = 0; d < gpu_ids.size(); ++d) {
            mqi_handle* h = nullptr;
            check(mqi_create(gpu_ids[d], &h), "mqi_create");
            handles.push_back(h);
            // tps_env is built without __PHYSICS_DEBUG__ (tests/mc/tps/CMakeLists.txt)
            // B2 (scorers [0, n-2) scored twice) belongs to transport_particles_patient only: the _stat kernel a run
            // with StoppingStatistics launches scores them once (mqi_transport.hpp:342-352)
            check(mqi_set_physics(h, MQI_PHYSICS_RELEASE, reference_quirks && !record_statistics ? MQI_QUIRK_B2_DOUBLE_SCORE : 0u),
                  "mqi_set_physics");
            check(mqi_set_grid_hu(h, grid.xe.data(), (int) grid.xe.size(), grid.ye.data(), (int) grid.ye.size(),
                                  grid.ze.data(), (int) grid.ze.size(), ct.hu.data(), density_scale, nullptr, nullptr),
                  "mqi_set_grid_hu");
            for (const auto& n : beamline)
                check(mqi_add_beamline_node(h, n.xe.data(), (int) n.xe.size(), n.ye.data(), (int) n.ye.size(), n.ze.data(),
                                            (int) n.ze.size(), n.rho.data(), R.m, b.iso),
                      "mqi_add_beamline_node");
            auto add = [&](int kind, const std::string& name, uint64_t cap, bool save) {
                const int id = mqi_add_scorer(h, kind, name.c_str(), cap);
                check(id, "mqi_add_scorer");
                if (d == 0) scorers.push_back({ id, kind, name, save });
                const bool is_stat = name == "Dose_stat" || name == "DoseSquare_stat";
                const std::vector<uint8_t>& m = is_stat ? stat_mask_total : scoring_mask_total;
                if (!m.empty()) {
                    uint64_t roi_n = 0;
                    check(mqi_set_scorer_roi(h, id, m.data(), m.size(), &roi_n), "mqi_set_scorer_roi");
                    (is_stat ? stat_roi_size : scoring_roi_size) = roi_n;
                }
                return id;
            };
            for (const auto& s : scorer_string) {
                if (strcasecmp(s.c_str(), "Dose") == 0) add(MQI_SCORER_DOSE, s, nvox, true);
                else if (strcasecmp(s.c_str(), "EnergyDeposition") == 0) add(MQI_SCORER_EDEP, s, nvox, true);
                else if (strcasecmp(s.c_str(), "LETd") == 0) {
                    add(MQI_SCORER_LETD_NUMER, "LETd_numer", nvox, true);
                    add(MQI_SCORER_LETD_DENOM, "LETd_denom", nvox, true);
                } else if (strcasecmp(s.c_str(), "LETt") == 0) {
                    add(MQI_SCORER_LETT_NUMER, "LETt_numer", nvox, true);
                    add(MQI_SCORER_LETT_DENOM, "LETt_denom", nvox, true);
                } else if (strcasecmp(s.c_str(), "Dij") == 0) {
                    // the reference hard-codes 512*512*300*5 slots (:922); sized here from the work instead:
                    // at most one entry per scored step of this device's share of the histories, bounded below
                    // and by a quarter of the device's free HBM.  The probe sequence is bound by uncoalesced
                    // 32-byte sectors, so the load factor is what counts: C4 (2.9e8 entries) runs at 4.7e7
                    // histories/s in the reference's 393 216 000 slots (load 0.73) and at 6.4e7 in 1.6e9
                    // slots (26 GB, load 0.18) -- profiles/r1_experiments.md.
                    uint64_t dev_share = (total_histories + gpu_ids.size() - 1) / gpu_ids.size();
                    uint64_t free_b = 0, total_b = 0;
                    check(mqi_device_memory(h, &free_b, &total_b), "mqi_device_memory");
                    uint64_t cap = std::max<uint64_t>(1u << 20, std::min<uint64_t>(dev_share * 64ull, free_b / 4 / 16));
                    if (reference_quirks) cap = std::min<uint64_t>(cap, 393216000ull);
                    if (dij_capacity > 0) cap = dij_capacity;   // extension: DijCapacity (slots of 16 B)
                    add(MQI_SCORER_DIJ, s, cap | 1ull, true);
                }   // TrackLength: accepted by the parser, creates no scorer (as in the reference :850-933)
            }
            if (record_statistics) {
                const int a = add(MQI_SCORER_DOSE, "Dose_stat", nvox, save_statistics);
                const int c = add(MQI_SCORER_DOSE_SQ, "DoseSquare_stat", nvox, save_statistics);
                if (d == 0) { stat_sum = a; stat_sumsq = c; }
            }
            if (scorers.empty()) throw std::runtime_error("no scorer was created from the Scorer list");
            check(mqi_set_beamlets(h, beamlets.data(), num_spots, spot_histories.data()), "mqi_set_beamlets");
        }
        printf("total scorer %d current scorer %d\n", (int) scorers.size(), (int) scorers.size());
    }

    // The history ranges of device r.  Per beam: the whole range, of which the device transports its interleaved
    // share.  Per spot (Dij): whole
    // SPOTS -- the rows of the matrix stay on one device, so the shards need no reduction -- dealt in kSpotBlocks
    // contiguous blocks per device, round-robin: a plan lists its spots energy layer by energy layer and the cost of a
    // spot grows with its range, so one contiguous block per device leaves the device with the highest layers working
    // twice as long as the one with the lowest (measured on C4 over two devices: 76 against 212 million entries).
    static constexpr uint64_t kSpotBlocks = 4;
    std::vector<std::pair<uint64_t, uint64_t>>
    device_ranges(size_t r) const {
        const uint64_t g = handles.size();
        std::vector<std::pair<uint64_t, uint64_t>> out;
        if (sim_type == PER_SPOT) {
            std::vector<uint64_t> cum(num_spots + 1, 0);
            for (uint64_t s = 0; s < num_spots; ++s) cum[s + 1] = cum[s] + spot_histories[s];
            const uint64_t nb = g > 1 ? g * kSpotBlocks : 1;
            for (uint64_t blk = r; blk < nb; blk += g) {
                const uint64_t s0 = (uint64_t) num_spots * blk / nb, s1 = (uint64_t) num_spots * (blk + 1) / nb;
                if (cum[s1] > cum[s0]) out.emplace_back(cum[s0], cum[s1]);
            }
        } else {
            // per beam: every device takes its interleaved share of the WHOLE range (mqi_run_async_sharded: chunks of
            // 32 histories dealt round-robin), so each gets the same mix of energy layers
            out.emplace_back(0, total_histories);
        }
        return out;
    }

    // all histories of the beam with `seed`: every device works through its own ranges in batches of
    // MaxHistoriesPerBatch (run_by_beam / run_by_spot batching), the devices run concurrently
    void
    run_all(uint64_t seed) {
        const uint64_t g     = handles.size();
        const uint64_t batch = max_histories_per_batch > 0 ? (uint64_t) max_histories_per_batch : total_histories;
        if (max_histories_per_batch <= 0) printf("Uploading %lu histories\n", (unsigned long) total_histories);
        else printf("Upload %lu histories per batch, %d batches expected\n", (unsigned long) batch,
                    (int) ((total_histories + batch - 1) / batch));
        std::vector<std::vector<std::pair<uint64_t, uint64_t>>> todo(g);
        std::vector<size_t>   at(g, 0);
        std::vector<float>    dev_ms(g, 0.f);
        for (size_t r = 0; r < g; ++r) todo[r] = device_ranges(r);
        bool more = true;
        while (more) {
            more = false;
            printf("Transporting particles...\n");
            std::vector<char> launched(g, 0);
            for (size_t r = 0; r < g; ++r) {
                if (at[r] >= todo[r].size()) continue;
                auto&          rg = todo[r][at[r]];
                const uint64_t n  = std::min<uint64_t>(batch, rg.second - rg.first);
                if (sim_type == PER_SPOT) check(mqi_run_async(handles[r], seed, rg.first, n, 1), "mqi_run_async");
                else check(mqi_run_async_sharded(handles[r], seed, rg.first, n, 0, (uint32_t) g, (uint32_t) r), "mqi_run_async_sharded");
                rg.first += n;
                if (rg.first >= rg.second) ++at[r];
                launched[r] = 1;
                more        = more || at[r] < todo[r].size();
            }
            for (size_t r = 0; r < g; ++r) {
                if (!launched[r]) continue;
                mqi_run_stats st;
                check(mqi_get_run_stats(handles[r], &st), "mqi_get_run_stats");
                tracked += st.histories;
                dev_ms[r] += st.kernel_ms;
                if (st.dij_table_full) throw std::runtime_error("Dij table is full");
            }
        }
        kernel_ms_total += *std::max_element(dev_ms.begin(), dev_ms.end());   // the devices run concurrently: the slowest counts
    }

    // dense scorers of devices 1.. are added into device 0 (one NCCL reduce each): device 0 then holds the total
    void
    gather_dense() {
        if (handles.size() < 2) return;
        for (const auto& s : scorers)
            if (s.kind != MQI_SCORER_DIJ) check(mqi_reduce_dense(handles.data(), (int) handles.size(), s.id, 0), "mqi_reduce_dense");
        // the other devices keep what they hold: their Dij tables are their rows of the matrix (save() concatenates
        // them), and gather_dense() runs once per beam, after the last batch or pass
    }

    // calculate_stat: mean over the voxels with mean dose > StatThreshold * max of sigma / mu.  The devices keep
    // their own running sums; mqi_stat_multi reduce-scatters the stat pair and evaluates the criterion in slices, so
    // no grid travels to one device between passes (the dense scorers are gathered once, after the last pass)
    float
    calculate_stat(uint64_t n_histories) {
        double     out[3] = { 0, 0, 0 };
        const auto t0     = std::chrono::high_resolution_clock::now();
        check(mqi_stat_multi(handles.data(), (int) handles.size(), stat_sum, stat_sumsq, n_histories, stat_threshold, out), "mqi_stat_multi");
        stat_ms_total += std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        return out[1] > 0 ? (float) (out[0] / out[1]) : 0.f;
    }
    double stat_ms_total = 0.0, gather_ms_total = 0.0;

    void
    run() {
        tracked         = 0;
        kernel_ms_total = 0.f;
        const uint64_t seed = (uint64_t) (int64_t) master_seed;
        printf("scorer %d sim type %d\n", (int) scorers.size(), (int) sim_type);
        if (sim_type == PER_BEAM && record_statistics) {
            // run_by_beam_stat: passes over all histories until the mean relative uncertainty falls to
            // StoppingCriteria %.  The reference replays identical seeds on every pass (B7); each pass
            // here draws fresh counter-based streams (seed + pass).
            printf("Run by beam\n");
            float current_stat = 100.0f;
            stat_passes        = 0;
            while (current_stat > stat_criteria && stat_passes < max_stat_passes) {
                if (current_stat == 100.0f) printf("Running %lu histories for the first run\n", (unsigned long) total_histories);
                else printf("Running additional %lu histories\n", (unsigned long) total_histories);
                run_all(seed + (uint64_t) stat_passes);
                ++stat_passes;
                current_stat = calculate_stat(tracked) * 100.0f;
                printf("Number of particles tracked %lu\n", (unsigned long) tracked);
                printf("Run %d: current uncertainty %f %%\n", stat_passes, current_stat);
            }
            last_stat_percent = current_stat;
            {
                const auto t0 = std::chrono::high_resolution_clock::now();
                gather_dense();
                gather_ms_total += std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
            }
            printf("Stopping criterion: %d evaluations %f ms (reduce-scatter of the stat pair + partial sums), final gather of the dense scorers %f ms\n",
                   stat_passes, stat_ms_total, gather_ms_total);
            // calculate_average_results: scorers [0, n-2) scaled by target / tracked histories
            const double w = (double) total_histories / (double) tracked;
            for (const auto& s : scorers)
                if (s.id != stat_sum && s.id != stat_sumsq) check(mqi_scale_scorer(handles[0], s.id, w), "mqi_scale_scorer");
        } else {
            printf(sim_type == PER_SPOT ? "Run by spot\n" : "Run by beam\n");
            if (sim_type == PER_SPOT) printf("num spots %d\n", (int) num_spots);
            run_all(seed);
            gather_dense();
            printf("Number of particles tracked %lu\n", (unsigned long) tracked);
        }
        printf("Transport kernels %f ms on %d GPU(s): %e histories/s\n", kernel_ms_total, (int) handles.size(),
               kernel_ms_total > 0 ? 1e3 * (double) tracked / kernel_ms_total : 0.0);
    }

    // save_reshaped_files / save_sparse_file: "<BeamName>_<child>_<scorer>.<ext>", values times
    // ParticlesPerHistory * RBE * NumberOfFraction.  The patient is the last child of the world: child
    // 0 without beamline nodes, else the number of beamline nodes (:1860-1866).
    void
    save() {
        const std::string beam_name = plan.beams[bnb - 1].name;
        const double      scale     = (double) (particles_per_history * rbe * n_fractions);
        const size_t      nvox      = (size_t) ct.nx * ct.ny * ct.nz;
        std::vector<double> dense;
        for (const auto& s : scorers) {
            if (!s.save) continue;
            const std::string filename = beam_name + "_" + std::to_string(n_beamline) + "_" + s.name;
            if (s.kind == MQI_SCORER_DIJ || sparse_output) {
                if (s.kind != MQI_SCORER_DIJ) throw std::runtime_error("OutputFormat npz is for the Dij scorer");
                std::vector<uint32_t> vox, spot;
                std::vector<double>   val;
                for (auto* h : handles) {   // spots are sharded by device: rows are disjoint, concatenate
                    uint64_t nnz = 0;
                    check(mqi_get_sparse_count(h, s.id, &nnz), "mqi_get_sparse_count");
                    const size_t o = vox.size();
                    vox.resize(o + nnz); spot.resize(o + nnz); val.resize(o + nnz);
                    if (nnz) check(mqi_get_sparse(h, s.id, vox.data() + o, spot.data() + o, val.data() + o, nnz, scale), "mqi_get_sparse");
                }
                printf("%d\n", (int) num_spots);
                save_csr_npz(output_path + "/" + filename + ".npz", num_spots, (uint32_t) nvox, vox, spot, val);
                continue;
            }
            dense.resize(nvox);
            check(mqi_get_dense(handles[0], s.id, dense.data(), 1.0), "mqi_get_dense");
            if (!output_format.compare("mhd")) save_to_mhd(grid, dense.data(), scale, output_path, filename, nvox);
            else if (!output_format.compare("mha")) save_to_mha(grid, dense.data(), scale, output_path, filename, nvox);
            else save_to_bin(dense.data(), scale, output_path, filename, nvox);
        }
    }

    void
    initialize_and_run() {
        for (size_t q = 0; q < beam_numbers.size(); ++q) {
            bnb = beam_numbers[q];
            master_seed += (int) q * 10000;
            printf("bnb %d seed %d\n", bnb, master_seed);
            initialize();
            run();
            save();
        }
    }
};

}   // namespace mqib
