// Harness that runs the REFERENCE's own CPU transport (headers included from /root/reference,
// never copied) for the scorer sets and the per-spot (Dij) mode that the phantom_env CLI does not
// expose.  It subclasses mqi::phantom_env<float>, replaces setup_world()'s scorer list and, for Dij,
// calls mc::transport_particles_patient directly with a scorer_offset_vector, exactly as
// mqi_tps_env.hpp:1119-1135 does.  Test infrastructure only (oracle/_ref/ref_harness_*).
//
//   ref_harness <phantom_env flags...> --scorers dose|edep|letd|lett|dose+letd|dij [--nspots N] [--spot_pitch mm]
//               [--gauss sx sy sxp syp sigmaE]
//               [--rangeshifter zlo zhi half density_g_cm3]      one-voxel slab, create_rangeshifter style
//               [--aperture zlo zhi half open_hx open_hy]        1 mm voxels, 1e-8 open / 100 closed
//               [--frame r00 .. r22 tx ty tz]                    rotation_matrix_fwd / translation of both
//               [--roi_mask file]                                uint8 summed mask -> mask_to_roi CONTOUR roi
// Beamline children are inserted in front of the phantom (which becomes the last child, as in
// tps_env::setup_world mqi_tps_env.hpp:732-758), so dense outputs are named <n_beamline>_<name>.raw.
// Output (into --output_prefix):
//   0_<name>.raw            dense float64 [nz][ny][nx] per scorer (reference save_reshaped_files)
//   dij_key1.raw/_key2.raw/_value.raw    occupied (voxel, spot, value) triplets in slot order
//   harness_stats.txt       wall seconds of the transport call, histories
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <moqui/base/environments/mqi_phantom_env.hpp>
#include <moqui/base/mqi_file_handler.hpp>

namespace
{
struct extra_opts {
    std::string scorers    = "dose";
    int         nspots     = 1;
    float       spot_pitch = 10.f;
    bool        gauss      = false;
    float       g[5]       = { 0, 0, 0, 0, 0 };   // sx sy sxp syp sigmaE
    bool        has_rs = false, has_ap = false, has_frame = false;
    float       rs[4]  = { 0, 0, 0, 0 };          // zlo zhi half density
    float       ap[5]  = { 0, 0, 0, 0, 0 };       // zlo zhi half open_hx open_hy
    float       frame[12] = { 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0 };
    std::string roi_mask;
};

class harness_env : public mqi::phantom_env<float>
{
public:
    typedef float R;
    extra_opts    opt;
    double        transport_seconds = 0;

    harness_env(mqi::cli& c, const extra_opts& o) : mqi::phantom_env<float>(c), opt(o) {}

    mqi::scorer<R>*
    make_scorer(const char* name, uint32_t capacity, mqi::fp_compute_hit<R> fp, uint32_t nvox) {
        mqi::scorer<R>* s = new mqi::scorer<R>(name, capacity, fp);
        mqi::key_value* t = new mqi::key_value[capacity];
        mqi::init_table(t, capacity);
        s->data_ = t;
        s->roi_  = roi_override ? roi_override : new mqi::roi_t(mqi::DIRECT, nvox);
        return s;
    }

    mqi::roi_t* roi_override = nullptr;

    mqi::node_t<R>*
    bare_node(mqi::grid3d<mqi::density_t, R>* geo) {
        mqi::node_t<R>* n = new mqi::node_t<R>;
        n->n_scorers  = 0;
        n->scorers    = nullptr;
        n->n_children = 0;
        n->children   = nullptr;
        n->geo        = geo;
        geo->translation_vector = mqi::vec3<R>(opt.frame[9], opt.frame[10], opt.frame[11]);
        return n;
    }

    // beamline children in front of the phantom, built like create_rangeshifter / create_voxelized_aperture
    // (mqi_tps_env.hpp:1605-1736): a one-voxel slab (grid3d with 2 edges per axis, fill_data) and a grid
    // voxelised at 1 mm holding 1e-8 (open) or 100 (closed)
    void
    insert_beamline() {
        std::vector<mqi::node_t<R>*> nodes;
        mqi::mat3x3<R> rot(opt.frame[0], opt.frame[1], opt.frame[2], opt.frame[3], opt.frame[4], opt.frame[5],
                           opt.frame[6], opt.frame[7], opt.frame[8]);
        if (opt.has_rs) {
            auto* geo = new mqi::grid3d<mqi::density_t, R>(-opt.rs[2], opt.rs[2], 2, -opt.rs[2], opt.rs[2], 2,
                                                           opt.rs[0], opt.rs[1], 2, rot);
            geo->fill_data(opt.rs[3] * 1e-3);
            nodes.push_back(bare_node(geo));
        }
        if (opt.has_ap) {
            const int nxy = (int) std::ceil(2 * opt.ap[2]), nz = (int) std::ceil(opt.ap[1] - opt.ap[0]);
            R* xe = new R[nxy + 1];
            R* ze = new R[nz + 1];
            for (int i = 0; i <= nxy; ++i) xe[i] = -opt.ap[2] + i * 1.0f;
            for (int i = 0; i <= nz; ++i) ze[i] = opt.ap[0] + i * 1.0f;
            auto* geo = new mqi::grid3d<mqi::density_t, R>(xe, nxy + 1, xe, nxy + 1, ze, nz + 1, rot);
            mqi::density_t* data = new mqi::density_t[(size_t) nxy * nxy * nz];
            for (int k = 0; k < nz; ++k)
                for (int j = 0; j < nxy; ++j)
                    for (int i = 0; i < nxy; ++i) {
                        const float x = xe[i] + 0.5f, y = xe[j] + 0.5f;
                        const bool  open = std::fabs(x) < opt.ap[3] && std::fabs(y) < opt.ap[4];
                        data[((size_t) k * nxy + j) * nxy + i] = open ? 1e-8 : 100.0;
                    }
            geo->set_data(data);
            nodes.push_back(bare_node(geo));
        }
        if (nodes.empty()) return;
        mqi::node_t<R>* ph = this->world->children[0];
        this->world->n_children = nodes.size() + 1;
        this->world->children   = new mqi::node_t<R>*[this->world->n_children];
        for (size_t i = 0; i < nodes.size(); ++i) this->world->children[i] = nodes[i];
        this->world->children[nodes.size()] = ph;
    }

    virtual void
    setup_world() {
        mqi::phantom_env<float>::setup_world();   // geometry + density + the default scorer
        mqi::node_t<R>* ph   = this->world->children[0];
        const uint32_t  nvox = nxyz.x * nxyz.y * nxyz.z;
        insert_beamline();
        if (!opt.roi_mask.empty()) {
            // mask_reader::set_mask + mask_to_roi (mqi_file_handler.hpp:160-217) on a summed mask volume
            std::vector<uint8_t>* m = new std::vector<uint8_t>(nvox);
            std::ifstream f(opt.roi_mask, std::ios::binary);
            f.read((char*) m->data(), nvox);
            mqi::vec3<mqi::ijk_t> dim(nxyz.x, nxyz.y, nxyz.z);
            mqi::mask_reader      mr(dim);
            mr.set_mask(m->data());
            roi_override = mr.mask_to_roi();
            printf("roi runs %u size %d\n", roi_override->length_, roi_override->get_mask_size());
            ph->scorers[0]->roi_ = roi_override;
        }
        if (opt.scorers == "dose") return;
        delete[] ph->scorers[0]->data_;
        ph->scorers[0]->data_ = nullptr;
        std::vector<mqi::scorer<R>*> v;
        if (opt.scorers == "edep") {
            v.push_back(make_scorer("Edep", nvox, mqi::energy_deposit<R>, nvox));
        } else if (opt.scorers == "letd") {
            v.push_back(make_scorer("LETd_numer", nvox, mqi::LETd_weight1<R>, nvox));
            v.push_back(make_scorer("LETd_denom", nvox, mqi::LETd_weight2<R>, nvox));
        } else if (opt.scorers == "lett") {   // track-averaged LET, scorers/mqi_scorer_energy_deposit.hpp:141-177
            v.push_back(make_scorer("LETt_numer", nvox, mqi::LETt_weight1<R>, nvox));
            v.push_back(make_scorer("LETt_denom", nvox, mqi::LETt_weight2<R>, nvox));
        } else if (opt.scorers == "dose+letd") {   // 3 scorers: exercises the double-scoring quirk
            v.push_back(make_scorer("Dose", nvox, mqi::dose_to_water<R>, nvox));
            v.push_back(make_scorer("LETd_numer", nvox, mqi::LETd_weight1<R>, nvox));
            v.push_back(make_scorer("LETd_denom", nvox, mqi::LETd_weight2<R>, nvox));
        } else if (opt.scorers == "dij") {
            // the reference hard-codes 512*512*300*5 slots (mqi_tps_env.hpp:922); a smaller table keeps
            // the CPU harness in memory while exercising the same hash + probe code
            uint32_t cap = nvox / 4 * (uint32_t) opt.nspots + 1024;
            v.push_back(make_scorer("Dij", cap, mqi::dose_to_water<R>, nvox));
        } else {
            throw std::runtime_error("unknown --scorers");
        }
        ph->n_scorers = v.size();
        ph->scorers   = new mqi::scorer<R>*[v.size()];
        for (size_t i = 0; i < v.size(); ++i)
            ph->scorers[i] = v[i];
    }

    virtual void
    setup_beamsource() {
        mqi::coordinate_transform<R> p_coord(spot_angles, { 0, 0, 0 });
        size_t per_spot = n_histories / opt.nspots;
        for (int s = 0; s < opt.nspots; ++s) {
            float              off  = (s - 0.5f * (opt.nspots - 1)) * opt.spot_pitch;
            std::array<R, 6>   mean = { spot_position[0] + off, spot_position[1], spot_position[2], 0, 0, -1 };
            std::array<R, 2>   corr = { 0.0, 0.0 };
            mqi::pdf_Md<R, 6>* phsp;
            mqi::pdf_Md<R, 1>* energy;
            if (opt.gauss) {
                std::array<R, 6> sig = { opt.g[0], opt.g[1], 0.0, opt.g[2], opt.g[3], 0.0 };
                phsp                 = new mqi::phsp_6d<R>(mean, sig, corr);
                energy               = new mqi::norm_1d<R>({ spot_energy[0] }, { opt.g[4] });
            } else {
                std::array<R, 6> sig = { spot_size[0], spot_size[1], 0.0, 0.0, 0.0, 0.0 };
                phsp                 = new mqi::phsp_6d_uniform<R>(mean, sig, corr);
                energy               = new mqi::const_1d<R>({ spot_energy[0] }, { spot_energy[1] });
            }
            this->beamsource.append_beamlet(mqi::beamlet<R>(energy, phsp), per_spot, p_coord);
        }
        uint32_t h1    = this->beamsource.total_histories();
        this->vertices = new mqi::vertex_t<R>[h1];
        for (size_t i = 0; i < h1; ++i) {
            auto bl           = this->beamsource(i);
            this->vertices[i] = bl(&this->beam_rng);
        }
    }

    virtual void
    run() {
        uint32_t h1     = this->beamsource.total_histories();
        this->num_spots = this->beamsource.total_beamlets();
        uint32_t  tracked = 0;
        std::vector<mqi::key_t> spot_of(h1);
        uint32_t idx = 0;
        for (uint32_t s = 0; s < this->num_spots; ++s) {
            size_t n = std::get<1>(this->beamsource[s]);
            for (size_t k = 0; k < n; ++k)
                spot_of[idx++] = s;
        }
        std::vector<int32_t> seeds(h1, 0);   // unused on the CPU path
        mc::mc_world     = this->world;
        mc::mc_vertices  = this->vertices;
        mc::mc_materials = this->materials;
        mqi::thrd_t* wt  = new mqi::thrd_t[1];
        wt[0].rnd_generator.seed((unsigned) this->random_seed * 2654435761u + 12345u);
        auto t0 = std::chrono::high_resolution_clock::now();
        mc::transport_particles_patient<R>(wt,
                                           mc::mc_world,
                                           mc::mc_vertices,
                                           mc::mc_materials,
                                           h1,
                                           &tracked,
                                           seeds.data(),
                                           opt.scorers == "dij" ? spot_of.data() : nullptr);
        auto t1           = std::chrono::high_resolution_clock::now();
        transport_seconds = std::chrono::duration<double>(t1 - t0).count();
        std::cout << "Number of particles tracked " << tracked << std::endl;
        std::ofstream st(this->output_path + "/harness_stats.txt");
        st << "histories " << h1 << "\ntransport_seconds " << transport_seconds << "\n";
    }

    void
    save_dij_triplets() {
        mqi::scorer<R>*        s = this->world->children[this->world->n_children - 1]->scorers[0];
        std::vector<uint32_t> k1, k2;
        std::vector<double>   val;
        for (uint32_t i = 0; i < s->max_capacity_; ++i) {
            if (s->data_[i].key1 != mqi::empty_pair && s->data_[i].key2 != mqi::empty_pair) {
                k1.push_back(s->data_[i].key1);
                k2.push_back(s->data_[i].key2);
                val.push_back(s->data_[i].value);
            }
        }
        std::ofstream a(this->output_path + "/dij_key1.raw", std::ios::binary);
        a.write((const char*) k1.data(), k1.size() * 4);
        std::ofstream b(this->output_path + "/dij_key2.raw", std::ios::binary);
        b.write((const char*) k2.data(), k2.size() * 4);
        std::ofstream c(this->output_path + "/dij_value.raw", std::ios::binary);
        c.write((const char*) val.data(), val.size() * 8);
        printf("dij nnz %lu capacity %u\n", k1.size(), s->max_capacity_);
    }
};
}   // namespace

int
main(int argc, char* argv[]) {
    extra_opts         o;
    std::vector<char*> pass;
    pass.push_back(argv[0]);
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--scorers" && i + 1 < argc) {
            o.scorers = argv[++i];
        } else if (a == "--nspots" && i + 1 < argc) {
            o.nspots = std::stoi(argv[++i]);
        } else if (a == "--spot_pitch" && i + 1 < argc) {
            o.spot_pitch = std::stof(argv[++i]);
        } else if (a == "--rangeshifter" && i + 4 < argc) {
            o.has_rs = true;
            for (int k = 0; k < 4; ++k)
                o.rs[k] = std::stof(argv[++i]);
        } else if (a == "--aperture" && i + 5 < argc) {
            o.has_ap = true;
            for (int k = 0; k < 5; ++k)
                o.ap[k] = std::stof(argv[++i]);
        } else if (a == "--frame" && i + 12 < argc) {
            o.has_frame = true;
            for (int k = 0; k < 12; ++k)
                o.frame[k] = std::stof(argv[++i]);
        } else if (a == "--roi_mask" && i + 1 < argc) {
            o.roi_mask = argv[++i];
        } else if (a == "--gauss" && i + 5 < argc) {
            o.gauss = true;
            for (int k = 0; k < 5; ++k)
                o.g[k] = std::stof(argv[++i]);
        } else {
            pass.push_back(argv[i]);
        }
    }
    mqi::cli cl;
    cl.read((int) pass.size(), pass.data());
    harness_env env(cl, o);
    env.initialize();
    env.run();
    env.finalize();
    if (o.scorers == "dij") {
        env.save_dij_triplets();
    } else {
        env.save_reshaped_files();
    }
    return 0;
}
