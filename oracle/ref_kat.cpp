// Known-answer dumper: drives the REFERENCE's own headers (from /root/reference, never copied)
// and writes deterministic input/output vectors as raw little-endian arrays into a directory.
// oracle/gen_golden.py packs them into tests/golden/kat_<variant>.npz.
//
// Test infrastructure only. Built by oracle/build_ref.sh into oracle/_ref/ref_kat_{debug,release}.
//
// Covered reference entry points (file:line under /root/reference/moqui):
//   patient_material_t::hu_to_density          base/materials/mqi_patient_materials.hpp:514-542
//   spr_default / radiation_length_default     base/materials/mqi_patient_materials.hpp:414-473
//   grid3d::index / intersect / ijk2cnb        base/mqi_grid3d.hpp:403-413,490-626,631-743,745-877
//   mc::hash_fun(k1,k2,cap)                    kernel_functions/mqi_transport.hpp:32-51
//   start_and_length                           base/mqi_utils.hpp:138-146
//   relativistic_quantities                    base/mqi_relativistic_quantities.hpp:27-44
//   tabulated cross sections / dEdx            base/mqi_p_ionization.hpp:254-286, mqi_pp_elastic.hpp:221-235,
//                                              mqi_po_elastic.hpp:243-256, mqi_po_inelastic.hpp:141-155
//   track_t::update_post_vertex_direction      base/mqi_track.hpp:163-172 (+ mqi_matrix.hpp:69-150)
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <cstdint>
#include <string>
#include <vector>

#include <moqui/base/environments/mqi_phantom_env.hpp>
#include <moqui/base/mqi_file_handler.hpp>       // scratch copy cut before class file_parser (patch 3): mask_reader
#include <moqui/base/mqi_file_parser_only.hpp>   // scratch header cut from mqi_file_handler.hpp by build_ref.sh (patch 4)

static std::string g_dir;

template<typename T>
static void
dump(const char* name, const std::vector<T>& v) {
    std::string path = g_dir + "/" + name;
    FILE*       f    = fopen(path.c_str(), "wb");
    if (!f) {
        perror(path.c_str());
        exit(1);
    }
    fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

// splitmix64: input generator private to this dumper (inputs are stored next to outputs)
static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static uint64_t
next_u64() {
    uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
    z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double
next_unit() {
    return (next_u64() >> 11) * (1.0 / 9007199254740992.0);
}

int
main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: ref_kat <outdir>\n");
        return 1;
    }
    g_dir = argv[1];
    typedef float R;

    // ---- 1. HU -> density for every int16 the clamp can see
    mqi::patient_material_t<R> mat;
    {
        std::vector<int16_t> hu;
        std::vector<float>   rho;
        for (int h = -1100; h <= 3100; ++h) {
            hu.push_back((int16_t) h);
            rho.push_back(mat.hu_to_density((int16_t) h));
        }
        dump("hu.i16", hu);
        dump("hu_rho.f32", rho);
    }
    // ---- 2. rsp(rho, Ek), radiation length(rho)
    {
        const int   hus[] = { -1000, -990, -950, -800, -741, -500, -300, -200, -120, -98, -50, 0,   7,
                              15,    20,   23,   60,   100,  101,  300,  500,  800,  1000, 1500, 2000, 2001,
                              2500,  2995 };
        const float eks[] = { 0.1f, 0.3f, 0.5f, 1.f,  2.f,   3.7f,  5.5f,  10.f,  20.f,  35.f,  50.f,
                              70.f, 85.f, 100.f, 120.f, 150.f, 175.f, 200.f, 215.f, 230.f, 250.f, 299.f };
        std::vector<float> rho_in, ek_in, rsp, rl;
        for (int h : hus) {
            float r = mat.hu_to_density((int16_t) h);
            rl.push_back(mqi::radiation_length_default<R>(r, 1.0f / 1000.0f, 360.863f));
            for (float e : eks) {
                rho_in.push_back(r);
                ek_in.push_back(e);
                rsp.push_back(mqi::spr_default<R>(r, e));
            }
        }
        // a few literal densities that exercise the piecewise branches, incl. the debug shortcut
        const float rhos[] = { 1.0e-3f, 1.0005e-3f, 0.9995e-3f, 1.21e-6f, 1.0e-6f, 0.26e-3f, 0.2e-3f,
                               0.9e-3f, 0.899e-3f,  1.19e-3f,   1.0e-8f,  100.f };
        for (float r : rhos) {
            rl.push_back(mqi::radiation_length_default<R>(r, 1.0f / 1000.0f, 360.863f));
            for (float e : eks) {
                rho_in.push_back(r);
                ek_in.push_back(e);
                rsp.push_back(mqi::spr_default<R>(r, e));
            }
        }
        std::vector<float> rl_rho;
        for (int h : hus)
            rl_rho.push_back(mat.hu_to_density((int16_t) h));
        for (float r : rhos)
            rl_rho.push_back(r);
        dump("rsp_rho.f32", rho_in);
        dump("rsp_ek.f32", ek_in);
        dump("rsp_out.f32", rsp);
        dump("rl_rho.f32", rl_rho);
        dump("rl_out.f32", rl);
    }
    // ---- 3. hash keys
    {
        std::vector<uint32_t> k1, k2, out;
        std::vector<uint64_t> cap;
        const uint64_t        caps[] = { 1000003ull, 393216000ull, 14000000ull * 4ull, 65536ull, 7ull };
        uint32_t fixed[][2] = { { 0, 0 }, { 123456, 7 }, { 13999999, 4999 }, { 0xfffffffeu, 0xfffffffeu } };
        for (auto& f : fixed)
            for (uint64_t c : caps) {
                k1.push_back(f[0]);
                k2.push_back(f[1]);
                cap.push_back(c);
                out.push_back(mc::hash_fun(f[0], f[1], c));
            }
        for (int i = 0; i < 4000; ++i) {
            uint32_t a = (uint32_t)(next_u64() % 52428800ull);
            uint32_t b = (uint32_t)(next_u64() % 5000ull);
            uint64_t c = caps[i % 5];
            k1.push_back(a);
            k2.push_back(b);
            cap.push_back(c);
            out.push_back(mc::hash_fun(a, b, c));
        }
        dump("hash_k1.u32", k1);
        dump("hash_k2.u32", k2);
        dump("hash_cap.u64", cap);
        dump("hash_out.u32", out);
    }
    // ---- 4. start_and_length
    {
        std::vector<uint32_t> in, out;
        uint32_t              cases[][2] = { { 2, 11 }, { 7, 100 }, { 512, 100000 }, { 3, 2 }, { 1, 5 } };
        for (auto& c : cases)
            for (uint32_t t = 0; t < c[0] && t < 16; ++t) {
                auto r = mqi::start_and_length(c[0], c[1], t);
                in.push_back(c[0]);
                in.push_back(c[1]);
                in.push_back(t);
                out.push_back(r.x);
                out.push_back(r.y);
            }
        dump("sal_in.u32", in);
        dump("sal_out.u32", out);
    }
    // ---- 5. relativistic quantities + tabulated cross-sections / stopping power at rho = 1e-3
    {
        mqi::fippel_physics<R>  phys;
        mqi::material_t<R>*     m = &mat;
        std::vector<float>      ek, out;
        for (int i = 0; i < 1500; ++i) {
            float e = (i < 1400) ? 0.05f + 0.2143f * i : (float) (0.05 + 300.0 * next_unit());
            mqi::relativistic_quantities<R> rel(e, 938.272046f);
            ek.push_back(e);
            out.push_back(rel.beta_sq);
            out.push_back(rel.gamma);
            out.push_back(rel.Te_max);
            out.push_back(rel.momentum());
            out.push_back(phys.p_ion.cross_section(rel, m, 1.0e-3f));
            out.push_back(phys.p_ion.dEdx(rel, m));
            out.push_back(phys.pp_e.cross_section(rel, m, 1.0e-3f));
            out.push_back(phys.po_e.cross_section(rel, m, 1.0e-3f));
            out.push_back(phys.po_i.cross_section(rel, m, 1.0e-3f));
        }
        dump("phys_ek.f32", ek);
        dump("phys_out.f32", out);   // [n][9]
        // raw tables (600 each) so the oracle's private copies can be verified entry by entry
        std::vector<float> tabs;
        for (int i = 0; i < 600; ++i) tabs.push_back(mqi::cs_p_ion_table[i]);
        for (int i = 0; i < 600; ++i) tabs.push_back(mqi::restricted_stopping_power_table[i]);
        for (int i = 0; i < 600; ++i) tabs.push_back(mqi::range_steps[i]);
        for (int i = 0; i < 600; ++i) tabs.push_back(mqi::cs_pp_e_g4_table[i]);
        for (int i = 0; i < 600; ++i) tabs.push_back(mqi::cs_po_e_g4_table[i]);
        for (int i = 0; i < 600; ++i) tabs.push_back(mqi::cs_po_i_g4_table[i]);
        dump("tables.f32", tabs);   // [6][600]
        std::vector<float> corr(mqi::density_correction, mqi::density_correction + 3996);
        dump("density_correction.f32", corr);
    }
    // ---- 6. geometry: the C1 grid (200x200x350 over 100x100x350 mm centred at (0,0,-175))
    {
        const int nx = 200, ny = 200, nz = 350;
        mqi::grid3d<mqi::density_t, R> g(-50.f, 50.f, nx + 1, -50.f, 50.f, ny + 1, -350.f, 0.f, nz + 1);
        std::vector<float> edges;
        for (int i = 0; i <= nx; ++i) edges.push_back(g.get_x_edges()[i]);
        for (int i = 0; i <= ny; ++i) edges.push_back(g.get_y_edges()[i]);
        for (int i = 0; i <= nz; ++i) edges.push_back(g.get_z_edges()[i]);
        dump("geo_edges.f32", edges);
        const int          N = 6000;
        std::vector<float> pin, din, dist_out, dir_out, p1_out;
        std::vector<int>   idx_out, idx1_out;
        std::vector<uint64_t> cnb_out;
        for (int n = 0; n < N; ++n) {
            mqi::vec3<R> p, d;
            int          kind = n % 6;
            // interior points; every 6th family pins coordinates on / next to edges
            p.x = (float) (-50.0 + 100.0 * next_unit());
            p.y = (float) (-50.0 + 100.0 * next_unit());
            p.z = (float) (-350.0 + 350.0 * next_unit());
            if (kind == 1) p.x = g.get_x_edges()[next_u64() % (nx + 1)];
            if (kind == 2) p.y = g.get_y_edges()[next_u64() % (ny + 1)] + (float) (2e-3 * (next_unit() - 0.5));
            if (kind == 3) p.z = g.get_z_edges()[next_u64() % (nz + 1)] + (float) (4e-4 * (next_unit() - 0.5));
            if (kind == 4) {
                p.x = g.get_x_edges()[next_u64() % (nx + 1)];
                p.z = g.get_z_edges()[next_u64() % (nz + 1)];
            }
            d.x = (float) (2.0 * next_unit() - 1.0);
            d.y = (float) (2.0 * next_unit() - 1.0);
            d.z = (float) (2.0 * next_unit() - 1.0);
            if (kind == 5) {   // beam-like, nearly axis-parallel incl. exact zeros
                d.x = (n % 12 == 5) ? 0.f : (float) (2e-4 * (next_unit() - 0.5));
                d.y = (float) (1e-2 * (next_unit() - 0.5));
                d.z = -1.f;
            }
            d.normalize();
            pin.push_back(p.x); pin.push_back(p.y); pin.push_back(p.z);
            din.push_back(d.x); din.push_back(d.y); din.push_back(d.z);
            mqi::vec3<mqi::ijk_t> c = g.index(p, d);
            idx_out.push_back(c.x); idx_out.push_back(c.y); idx_out.push_back(c.z);
            if (g.is_valid(c)) {
                cnb_out.push_back(g.ijk2cnb(c));
                mqi::vec3<R>       pp = p, dd = d;
                mqi::intersect_t<R> its = g.intersect(pp, dd, c);
                dist_out.push_back(its.dist);
                dir_out.push_back(dd.x); dir_out.push_back(dd.y); dir_out.push_back(dd.z);
                // move to the exit point and update the cell incrementally
                float        len = its.dist > 0 ? its.dist : 0.f;
                mqi::vec3<R> p1  = pp + dd * len;
                p1_out.push_back(p1.x); p1_out.push_back(p1.y); p1_out.push_back(p1.z);
                mqi::vec3<mqi::ijk_t> c1 = c;
                g.index(p1, dd, c1);
                idx1_out.push_back(c1.x); idx1_out.push_back(c1.y); idx1_out.push_back(c1.z);
            } else {
                cnb_out.push_back(~0ull);
                dist_out.push_back(-2.f);
                dir_out.push_back(d.x); dir_out.push_back(d.y); dir_out.push_back(d.z);
                p1_out.push_back(p.x); p1_out.push_back(p.y); p1_out.push_back(p.z);
                idx1_out.push_back(c.x); idx1_out.push_back(c.y); idx1_out.push_back(c.z);
            }
        }
        dump("geo_p.f32", pin);
        dump("geo_d.f32", din);
        dump("geo_idx.i32", idx_out);
        dump("geo_cnb.u64", cnb_out);
        dump("geo_dist.f32", dist_out);
        dump("geo_dir_after.f32", dir_out);
        dump("geo_p1.f32", p1_out);
        dump("geo_idx1.i32", idx1_out);

        // entry from outside the grid
        std::vector<float> epin, edin, edist;
        std::vector<int>   ecell;
        for (int n = 0; n < 2000; ++n) {
            mqi::vec3<R> p, d;
            p.x = (float) (-120.0 + 240.0 * next_unit());
            p.y = (float) (-120.0 + 240.0 * next_unit());
            p.z = (float) (-450.0 + 550.0 * next_unit());
            if (n % 4 == 0) {   // the phantom_env source plane: z = +0.5, heading -z
                p.x = (float) (-30.0 + 60.0 * next_unit());
                p.y = (float) (-30.0 + 60.0 * next_unit());
                p.z = 0.5f;
                d.x = 0; d.y = 0; d.z = -1;
            } else {
                mqi::vec3<R> target((float) (-50.0 + 100.0 * next_unit()),
                                    (float) (-50.0 + 100.0 * next_unit()),
                                    (float) (-350.0 + 350.0 * next_unit()));
                d = target - p;
                if (n % 4 == 3) { d.x = (float) (2 * next_unit() - 1); d.y = (float) (2 * next_unit() - 1); d.z = (float) (2 * next_unit() - 1); }
                d.normalize();
            }
            epin.push_back(p.x); epin.push_back(p.y); epin.push_back(p.z);
            edin.push_back(d.x); edin.push_back(d.y); edin.push_back(d.z);
            mqi::vec3<R>        pp = p, dd = d;
            mqi::intersect_t<R> its = g.intersect(pp, dd);
            edist.push_back(its.dist);
            ecell.push_back(its.cell.x); ecell.push_back(its.cell.y); ecell.push_back(its.cell.z);
        }
        dump("entry_p.f32", epin);
        dump("entry_d.f32", edin);
        dump("entry_dist.f32", edist);
        dump("entry_cell.i32", ecell);
    }
    // ---- 7. direction update (scattering rotation)
    {
        std::vector<float> din, ang, dout;
        for (int n = 0; n < 3000; ++n) {
            mqi::track_t<R> t;
            mqi::vec3<R>    d((float) (2 * next_unit() - 1), (float) (2 * next_unit() - 1), (float) (2 * next_unit() - 1));
            if (n % 3 == 0) { d.x *= 0.02f; d.y *= 0.02f; d.z = -1.f; }   // beam-like: Householder branch
            if (n % 50 == 1) { d.x = 0; d.y = 0; d.z = (n % 100 == 1) ? 1.f : -1.f; }
            d.normalize();
            t.vtx0.dir = d;
            t.vtx1.dir = d;
            float th  = (n % 5 == 0) ? (float) (3.14159 * next_unit()) : (float) (0.05 * next_unit());
            float phi = (float) (6.2831853 * next_unit());
            t.update_post_vertex_direction(th, phi);
            din.push_back(d.x); din.push_back(d.y); din.push_back(d.z);
            ang.push_back(th); ang.push_back(phi);
            dout.push_back(t.vtx1.dir.x); dout.push_back(t.vtx1.dir.y); dout.push_back(t.vtx1.dir.z);
        }
        dump("rot_d.f32", din);
        dump("rot_ang.f32", ang);
        dump("rot_out.f32", dout);
    }
    // ---- 8. beam frame: coordinate_transform(collimator, gantry, couch, iec2dicom) (mqi_coordinate_transform.hpp:52-58),
    // with the angles as treatment_machine_ion::create_coordinate_transform passes them (couch negated, iec2dicom 90)
    // and as phantom_env --spot_angles passes them (iec2dicom 0)
    {
        std::vector<float> ang, rot, moved;
        const float        colls[] = { 0.f, 15.f, -30.f, 90.f }, gans[] = { 0.f, 45.f, 90.f, 180.f, 270.f, 333.f },
                    couches[] = { 0.f, 10.f, -20.f, 90.f }, iecs[] = { 0.f, 90.f };
        for (float c : colls)
            for (float ga : gans)
                for (float co : couches)
                    for (float ie : iecs) {
                        std::array<R, 4>             a = { c, ga, co, ie };
                        mqi::vec3<R>                 pos(1.5f, -2.5f, 40.f);
                        mqi::coordinate_transform<R> t(a, pos);
                        ang.push_back(c); ang.push_back(ga); ang.push_back(co); ang.push_back(ie);
                        const R* m = &t.rotation.xx;
                        for (int i = 0; i < 9; ++i) rot.push_back(m[i]);
                        // a source-frame point and direction mapped to the patient frame: R * p + T, R * d
                        mqi::vec3<R> p(3.f, -4.f, 465.f), q = t.rotation * p + t.translation;
                        moved.push_back(q.x); moved.push_back(q.y); moved.push_back(q.z);
                    }
        dump("ct_ang.f32", ang);
        dump("ct_rot.f32", rot);
        dump("ct_moved.f32", moved);
    }
    // ---- 9. output writers: io::save_to_mhd / save_to_mha (mqi_io.hpp:493-591) on a small ragged-offset grid; the
    // files themselves are the vectors (gen_golden.py packs fmt_* into tests/golden/fmt_writers.npz as bytes)
    {
        const float xe[5] = { -1.25f, -0.75f, -0.25f, 0.25f, 0.75f };
        const float ye[4] = { 10.f, 10.75f, 11.5f, 12.25f };
        const float ze[3] = { -3.7f, -2.45f, -1.2f };
        mqi::node_t<R> node;
        node.geo = new mqi::grid3d<mqi::density_t, R>(xe, 5, ye, 4, ze, 3);
        double src[24];
        for (int i = 0; i < 24; ++i)
            src[i] = 0.001 * i * i - 0.0137 * i + (i % 5 == 0 ? 0.0 : 1e-9 * i);
        mqi::io::save_to_mhd<R>(&node, src, (R) 2.5, g_dir, "fmt_mhd", 24);
        mqi::io::save_to_mha<R>(&node, src, (R) 2.5, g_dir, "fmt_mha", 24);
    }
    // ---- 10. sparse output: io::save_to_npz (mqi_io.hpp:249-320) of a small (voxel, spot) table: the six entries of
    // tps_env --npz-selftest, in the same slot order, 3 spots x 10 voxels
    {
        const uint32_t  cap = 16;
        mqi::scorer<R>* sc  = new mqi::scorer<R>("Dij", cap, mqi::dose_to_water<R>);
        mqi::key_value* t   = new mqi::key_value[cap];
        mqi::init_table(t, cap);
        const uint32_t vox[6] = { 7, 2, 9, 2, 0, 5 }, spot[6] = { 2, 0, 2, 1, 0, 2 }, slot[6] = { 1, 3, 4, 8, 9, 14 };
        for (int i = 0; i < 6; ++i) {
            t[slot[i]].key1  = vox[i];
            t[slot[i]].key2  = spot[i];
            t[slot[i]].value = 0.5 + i;
        }
        sc->data_ = t;
        mqi::vec3<mqi::ijk_t> dim(10, 1, 1);
        mqi::io::save_to_npz<R>(sc, (R) 1.0, g_dir, "fmt_npz", dim, 3);
    }
    // ---- 11. the moqui input-parameter format: file_parser (mqi_file_handler.hpp:220-380) on a file with comments,
    // tabs, odd spacing, mixed-case and repeated keys, empty values and lists; the queries of tps_env --parse-selftest
    {
        const std::string in = g_dir + "/fmt_parser_in.txt";
        {
            std::ofstream f(in);
            f << "# a full-line comment\n\nGPUID 0\n   RandomSeed   -1932   # trailing comment\n"
              << "ParentDir\t/data/case one/\nscorer Dose, LETd ,Dij\nMask a.mha,b.mha , c.mha\nBeamNumbers 1, 3,5\n"
              << "XShift 1.5e0\nYShift -2\nZShift .25abc\nOverwriteResults True\nSaveMap 1\nReadStructure false\n"
              << "ScoringMask 0\nStoppingStatistics tRuE\nUnitWeights\nDUPLICATE first\nDuplicate second\nNoValue\n"
              << "OutputDir ./out # x\nParticlesPerHistory 2.5e4\nMachine pbs:/a b/machine.txt\nEmptyList \nTrailing 7   \n";
        }
        mqi::file_parser   p(in, " ");
        std::ofstream      o(g_dir + "/fmt_parser_out.txt");
        const char*        skeys[] = { "GPUID", "randomseed", "ParentDir", "SCORER", "Mask", "OutputDir", "Machine", "UnitWeights",
                                "Duplicate", "NoValue", "Missing", "EmptyList", "Trailing", "ZShift" };
        for (const char* k : skeys) o << k << "|s|" << p.get_string(k, "<default>") << "|\n";
        const char* ikeys[] = { "GPUID", "RandomSeed", "BeamNumbers", "Missing", "Trailing", "XShift", "ParticlesPerHistory" };
        for (const char* k : ikeys) o << k << "|i|" << p.get_int(k, -7) << "|\n";
        const char* fkeys[] = { "XShift", "YShift", "ZShift", "ParticlesPerHistory", "Missing", "RandomSeed" };
        for (const char* k : fkeys) o << k << "|f|" << std::setprecision(9) << p.get_float(k, 0.125f) << "|\n";
        const char* bkeys[] = { "OverwriteResults", "SaveMap", "ReadStructure", "ScoringMask", "StoppingStatistics", "Missing", "UnitWeights",
                                "GPUID", "XShift" };
        for (const char* k : bkeys) o << k << "|b|" << p.get_bool(k, false) << p.get_bool(k, true) << "|\n";
        const char* vkeys[] = { "scorer", "Mask", "BeamNumbers", "Missing", "EmptyList", "GPUID" };
        for (const char* k : vkeys) {
            o << k << "|v|";
            for (const auto& t : p.get_string_vector(k, ",")) o << "[" << t << "]";
            o << "|";
            for (int t : p.get_int_vector(k, ",")) o << t << ";";
            o << "|\n";
        }
    }
    // ---- 12. mask files: mask_reader::read_mha_file + read_mask_files + mask_to_roi (mqi_file_handler.hpp:38-217) on two
    // overlapping uint8 .mha masks of a 7 x 5 x 4 volume written here (one header in the order ITK writes it, one with
    // odd spacing and upper-case keys); the summed mask and the run-length roi are the vectors
    {
        const int nx = 7, ny = 5, nz = 4, n = nx * ny * nz;
        std::vector<uint8_t> a(n, 0), b(n, 0);
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const int v = (k * ny + j) * nx + i;
                    a[v] = (i >= 1 && i <= 4 && j >= 1 && j <= 3 && k <= 2) ? 1 : 0;
                    b[v] = ((i + j + k) % 3 == 0 || (i >= 4 && k >= 2)) ? 1 : 0;
                }
        a[0] = 1;        // a run that starts at voxel 0
        b[n - 1] = 1;    // and one still open at the end of the volume
        a[n - 1] = 0;
        const std::string fa = g_dir + "/fmt_mask_a.mha", fb = g_dir + "/fmt_mask_b.mha";
        {
            std::ofstream f(fa, std::ios::binary);
            f << "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\nCompressedData = False\n"
              << "TransformMatrix = 1 0 0 0 1 0 0 0 1\nOffset = 0 0 0\nCenterOfRotation = 0 0 0\nAnatomicalOrientation = RAI\n"
              << "ElementSpacing = 1 1 1\nDimSize = 7 5 4\nElementType = MET_UCHAR\nElementDataFile = LOCAL\n";
            f.write((const char*) a.data(), n);
        }
        {
            std::ofstream f(fb, std::ios::binary);
            f << "NDims=3\nDIMSIZE =   7 5 4\n  ElementType = MET_UCHAR  \nelementdatafile =  local\n";
            f.write((const char*) b.data(), n);
        }
        mqi::vec3<mqi::ijk_t>    dim(nx, ny, nz);
        std::vector<std::string> files = { fa, fb };
        mqi::mask_reader         mr(files, dim);
        mr.read_mask_files();
        mqi::roi_t*           roi = mr.mask_to_roi();
        std::vector<uint32_t> runs;
        runs.push_back(roi->length_);
        for (uint32_t k = 0; k < roi->length_; ++k) {
            runs.push_back(roi->start_[k]);
            runs.push_back(roi->stride_[k]);
        }
        dump("mask_runs.u32", runs);
        std::vector<uint32_t> tot(mr.mask_total, mr.mask_total + n);
        dump("mask_total.u32", tot);
    }
    printf("ref_kat: wrote KATs to %s\n", g_dir.c_str());
    return 0;
}
