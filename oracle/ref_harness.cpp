// Harness that runs the REFERENCE's own CPU transport (headers included from /root/reference,
// never copied) for the scorer sets and the per-spot (Dij) mode that the phantom_env CLI does not
// expose.  It subclasses mqi::phantom_env<float>, replaces setup_world()'s scorer list and, for Dij,
// calls mc::transport_particles_patient directly with a scorer_offset_vector, exactly as
// mqi_tps_env.hpp:1119-1135 does.  Test infrastructure only (oracle/_ref/ref_harness_*).
//
//   ref_harness <phantom_env flags...> --scorers dose|edep|letd|lett|dose+letd|dose+dose2|stat|dij [--nspots N] [--spot_pitch mm]
//               [--gauss sx sy sxp syp sigmaE]
//               [--spot_grid nx ny pitch]      nx*ny spots on a grid (overrides --nspots / --spot_pitch)
//               [--energy_step dE]             spot s gets spot_energy + s * dE
//               [--batches B]                  B passes of --histories each, vertices resampled per pass, scorer tables
//                                              accumulate (the reference's own batching, mqi_tps_env.hpp:1086-1140)
//               [--sample_threads T]           host threads that run the reference sampler (each with its own
//                                              beamsource copy and std::default_random_engine)
//               [--stat_threshold t]           "stat": Dose + the two stat scorers through transport_particles_patient_stat,
//                                              then (CUDA build) calculate_standard_deviation + the host part of calculate_stat
// The same file compiles with nvcc -x cu (oracle/build_ref.sh: ref_harness_gpu_<variant>): run() then uploads and
// launches the reference's own CUDA kernel (mqi_phantom_env.hpp:337-412) with CUDA events around the launch.
//               [--rangeshifter zlo zhi half density_g_cm3]      one-voxel slab, create_rangeshifter style
//               [--aperture zlo zhi half open_hx open_hy]        1 mm voxels, 1e-8 open / 100 closed
//               [--frame r00 .. r22 tx ty tz]                    rotation_matrix_fwd / translation of both
//               [--roi_mask file]                                uint8 summed mask -> mask_to_roi CONTOUR roi
// Beamline children are inserted in front of the phantom (which becomes the last child, as in
// tps_env::setup_world mqi_tps_env.hpp:732-758), so dense outputs are named <n_beamline>_<name>.raw.
// Output (into --output_prefix):
//   0_<name>.raw            dense float64 [nz][ny][nx] per scorer (reference save_reshaped_files)
//   dij_key1.raw/_key2.raw/_value.raw    occupied (voxel, spot, value) triplets in slot order
//   harness_stats.txt       histories, transport_seconds (CPU: wall of the call; CUDA: sum of the kernel's event times),
//                           init_threads_seconds, run_seconds (sampling + uploads + kernels), blocks, threads
//   stat_sd.raw / stat_mean.raw / stat_sum.raw / stat_sumsq.raw   "stat" mode (CUDA build)
#include <chrono>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <thread>
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
    int         grid[2]     = { 0, 0 };
    float       grid_pitch  = 5.f;
    float       energy_step = 0.f;
    int         batches     = 1;
    int         sample_threads = 1;
    float       stat_threshold = 0.5f;
};

// the reference's hit functions: host addresses, or (CUDA build) the device addresses held by the
// reference's own __device__ pointer symbols, scorers/mqi_scorer_energy_deposit.hpp:180-189
#if defined(__CUDACC__)
#define MQI_HIT(host_fn, dev_sym) hit_from_symbol(dev_sym)
template<typename S>
mqi::fp_compute_hit<float>
hit_from_symbol(const S& sym) {
    mqi::fp_compute_hit<float> fp;
    gpu_err_chk(cudaMemcpyFromSymbol(&fp, sym, sizeof(fp)));
    return fp;
}
#else
#define MQI_HIT(host_fn, dev_sym) (host_fn)
#endif

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
        const mqi::fp_compute_hit<R> f_dw    = MQI_HIT(mqi::dose_to_water<R>, mqi::Dw_pointer);
        const mqi::fp_compute_hit<R> f_dw2   = MQI_HIT(mqi::dose_to_water_square<R>, mqi::Dw_square_pointer);
        const mqi::fp_compute_hit<R> f_edep  = MQI_HIT(mqi::energy_deposit<R>, mqi::energy_deposit_pointer);
        const mqi::fp_compute_hit<R> f_letd1 = MQI_HIT(mqi::LETd_weight1<R>, mqi::LETd_weight1_pointer);
        const mqi::fp_compute_hit<R> f_letd2 = MQI_HIT(mqi::LETd_weight2<R>, mqi::LETd_weight2_pointer);
        const mqi::fp_compute_hit<R> f_lett1 = MQI_HIT(mqi::LETt_weight1<R>, mqi::LETt_weight1_pointer);
        const mqi::fp_compute_hit<R> f_lett2 = MQI_HIT(mqi::LETt_weight2<R>, mqi::LETt_weight2_pointer);
        if (opt.scorers == "edep") {
            v.push_back(make_scorer("Edep", nvox, f_edep, nvox));
        } else if (opt.scorers == "letd") {
            v.push_back(make_scorer("LETd_numer", nvox, f_letd1, nvox));
            v.push_back(make_scorer("LETd_denom", nvox, f_letd2, nvox));
        } else if (opt.scorers == "lett") {   // track-averaged LET, scorers/mqi_scorer_energy_deposit.hpp:141-177
            v.push_back(make_scorer("LETt_numer", nvox, f_lett1, nvox));
            v.push_back(make_scorer("LETt_denom", nvox, f_lett2, nvox));
        } else if (opt.scorers == "dose+letd") {   // 3 scorers: exercises the double-scoring quirk
            v.push_back(make_scorer("Dose", nvox, f_dw, nvox));
            v.push_back(make_scorer("LETd_numer", nvox, f_letd1, nvox));
            v.push_back(make_scorer("LETd_denom", nvox, f_letd2, nvox));
        } else if (opt.scorers == "dose+dose2") {   // dose_to_water_square, :64-77 (two scorers: no double scoring)
            v.push_back(make_scorer("Dose", nvox, f_dw, nvox));
            v.push_back(make_scorer("Dose2", nvox, f_dw2, nvox));
        } else if (opt.scorers == "stat") {
            // the scorer list of a tps run with StoppingStatistics (mqi_tps_env.hpp:846-905): the user's Dose, then the
            // stat pair (dose, dose^2) that transport_particles_patient_stat keys by roi_->get_mask_idx(cnb)
            v.push_back(make_scorer("Dose", nvox, f_dw, nvox));
            v.push_back(make_scorer("StatDose", nvox, f_dw, nvox));
            v.push_back(make_scorer("StatDose2", nvox, f_dw2, nvox));
        } else if (opt.scorers == "dij") {
            // the reference hard-codes 512*512*300*5 slots (mqi_tps_env.hpp:922); a smaller table keeps
            // the CPU harness in memory while exercising the same hash + probe code
            uint32_t cap = nvox / 4 * (uint32_t) n_spots() + 1024;
            v.push_back(make_scorer("Dij", cap, f_dw, nvox));
        } else {
            throw std::runtime_error("unknown --scorers");
        }
        ph->n_scorers = v.size();
        ph->scorers   = new mqi::scorer<R>*[v.size()];
        for (size_t i = 0; i < v.size(); ++i)
            ph->scorers[i] = v[i];
    }

    int
    n_spots() const { return opt.grid[0] > 0 ? opt.grid[0] * opt.grid[1] : opt.nspots; }

    // the beamlets of the run, appended to `bs` (called once per sampling thread: the reference's pdf objects hold
    // stateful std distributions, so every thread samples from its own copy)
    void
    build_beamsource(mqi::beamsource<R>& bs) {
        mqi::coordinate_transform<R> p_coord(spot_angles, { 0, 0, 0 });
        const int ns = n_spots();
        size_t per_spot = n_histories / ns;
        for (int s = 0; s < ns; ++s) {
            float offx, offy = 0.f;
            if (opt.grid[0] > 0) {
                offx = (s % opt.grid[0] - 0.5f * (opt.grid[0] - 1)) * opt.grid_pitch;
                offy = (s / opt.grid[0] - 0.5f * (opt.grid[1] - 1)) * opt.grid_pitch;
            } else {
                offx = (s - 0.5f * (opt.nspots - 1)) * opt.spot_pitch;
            }
            const R            e0   = spot_energy[0] + s * opt.energy_step;
            std::array<R, 6>   mean = { spot_position[0] + offx, spot_position[1] + offy, spot_position[2], 0, 0, -1 };
            std::array<R, 2>   corr = { 0.0, 0.0 };
            mqi::pdf_Md<R, 6>* phsp;
            mqi::pdf_Md<R, 1>* energy;
            if (opt.gauss) {
                std::array<R, 6> sig = { opt.g[0], opt.g[1], 0.0, opt.g[2], opt.g[3], 0.0 };
                phsp                 = new mqi::phsp_6d<R>(mean, sig, corr);
                energy               = new mqi::norm_1d<R>({ e0 }, { opt.g[4] });
            } else {
                std::array<R, 6> sig = { spot_size[0], spot_size[1], 0.0, 0.0, 0.0, 0.0 };
                phsp                 = new mqi::phsp_6d_uniform<R>(mean, sig, corr);
                energy               = new mqi::const_1d<R>({ e0 }, { spot_energy[1] });
            }
            bs.append_beamlet(mqi::beamlet<R>(energy, phsp), per_spot, p_coord);
        }
    }

    std::vector<mqi::beamsource<R>*>           thread_sources;
    std::vector<std::default_random_engine*>   thread_rngs;

    // beamsource(h)(rng) for every history of a pass, mqi_phantom_env.hpp:225-232; with --sample_threads T the
    // history range is cut into T contiguous parts, each sampled by its own thread
    void
    sample_vertices(int pass) {
        const uint32_t h1 = this->beamsource.total_histories();
        const int      T  = std::max(1, opt.sample_threads);
        if (T == 1) {
            for (size_t i = 0; i < h1; ++i) {
                auto bl           = this->beamsource(i);
                this->vertices[i] = bl(&this->beam_rng);
            }
            return;
        }
        if (thread_sources.empty()) {
            for (int t = 0; t < T; ++t) {
                thread_sources.push_back(new mqi::beamsource<R>);
                build_beamsource(*thread_sources.back());
                thread_rngs.push_back(new std::default_random_engine);
            }
        }
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) {
            // (run seed, pass, thread) -> one stream each: a seed sequence, so that runs whose seeds differ by a
            // multiple of some stride never share a stream (an additive combination did: run k / thread t met run
            // k + 1 / thread t - 1, and the "independent" runs of the first GPU fixtures shared most of their vertices)
            std::seed_seq sq { (uint32_t) this->random_seed, (uint32_t) pass, (uint32_t) t, 0x6d716931u };
            thread_rngs[t]->seed(sq);
            const size_t a = (size_t) h1 * t / T, b = (size_t) h1 * (t + 1) / T;
            th.emplace_back([this, t, a, b]() {
                for (size_t i = a; i < b; ++i) {
                    auto bl           = (*thread_sources[t])(i);
                    this->vertices[i] = bl(thread_rngs[t]);
                }
            });
        }
        for (auto& x : th) x.join();
    }

    virtual void
    setup_beamsource() {
        build_beamsource(this->beamsource);
        uint32_t h1    = this->beamsource.total_histories();
        this->vertices = new mqi::vertex_t<R>[h1];
        sample_vertices(0);
    }

    virtual void
    run() {
        auto     r0     = std::chrono::high_resolution_clock::now();
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
        const bool per_spot = opt.scorers == "dij";
        const bool stat     = opt.scorers == "stat";
        double     init_threads_seconds = 0;
#if defined(__CUDACC__)
        // the reference's own launch sequence, mqi_phantom_env.hpp:337-412, with CUDA events around the kernel
        std::vector<int32_t>       seeds(h1);
        std::default_random_engine tmp_rng;
        tmp_rng.seed(this->random_seed);
        mqi::key_t* d_spot = nullptr;
        if (per_spot) mc::upload_scorer_offset_vector(spot_of.data(), d_spot, h1);
        mqi::thrd_t* wt;
        gpu_err_chk(cudaMalloc(&wt, (size_t) threads[1] * threads[0] * sizeof(mqi::thrd_t)));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        initialize_threads<<<threads[1], threads[0]>>>(wt, threads[1] * threads[0], 0);
        cudaEventRecord(e1);
        gpu_err_chk(cudaDeviceSynchronize());
        {
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            init_threads_seconds = ms * 1e-3;
        }
        int32_t*  d_seed;
        uint32_t* d_tracked;
        gpu_err_chk(cudaMalloc(&d_seed, sizeof(int32_t) * (size_t) h1));
        gpu_err_chk(cudaMalloc(&d_tracked, sizeof(uint32_t)));
        gpu_err_chk(cudaMemset(d_tracked, 0, sizeof(uint32_t)));
        printf("blocks %d threads %d\n", threads[1], threads[0]);
        for (int pass = 0; pass < opt.batches; ++pass) {
            if (pass > 0) sample_vertices(pass);
            for (uint32_t i = 0; i < h1; ++i) seeds[i] = tmp_rng();
            gpu_err_chk(cudaMemcpy(d_seed, seeds.data(), sizeof(int32_t) * (size_t) h1, cudaMemcpyHostToDevice));
            mc::upload_vertices(this->vertices, mc::mc_vertices, 0, h1);
            cudaEventRecord(e0);
            if (stat)
                mc::transport_particles_patient_stat<R><<<threads[1], threads[0]>>>(
                  wt, mc::mc_world, mc::mc_vertices, mc::mc_materials, h1, d_tracked, d_seed, d_spot);
            else
                mc::transport_particles_patient<R><<<threads[1], threads[0]>>>(
                  wt, mc::mc_world, mc::mc_vertices, mc::mc_materials, h1, d_tracked, d_seed, d_spot);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) {
                std::cout << "CUDA error (transport): " << cudaGetErrorString(err) << std::endl;
                exit(-1);
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            transport_seconds += ms * 1e-3;
            gpu_err_chk(cudaFree(mc::mc_vertices));
        }
        gpu_err_chk(cudaMemcpy(&tracked, d_tracked, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (stat) stat_on_device((int) tracked);
        if (d_spot) gpu_err_chk(cudaFree(d_spot));
        gpu_err_chk(cudaFree(wt));
        gpu_err_chk(cudaFree(d_seed));
        gpu_err_chk(cudaFree(d_tracked));
#else
        std::vector<int32_t> seeds(h1, 0);   // unused on the CPU path
        mc::mc_world     = this->world;
        mc::mc_vertices  = this->vertices;
        mc::mc_materials = this->materials;
        mqi::thrd_t* wt  = new mqi::thrd_t[1];
        wt[0].rnd_generator.seed((unsigned) this->random_seed * 2654435761u + 12345u);
        for (int pass = 0; pass < opt.batches; ++pass) {
            if (pass > 0) sample_vertices(pass);
            auto t0 = std::chrono::high_resolution_clock::now();
            if (stat)
                mc::transport_particles_patient_stat<R>(wt, mc::mc_world, mc::mc_vertices, mc::mc_materials, h1, &tracked,
                                                        seeds.data(), nullptr);
            else
                mc::transport_particles_patient<R>(wt, mc::mc_world, mc::mc_vertices, mc::mc_materials, h1, &tracked,
                                                   seeds.data(), per_spot ? spot_of.data() : nullptr);
            auto t1 = std::chrono::high_resolution_clock::now();
            transport_seconds += std::chrono::duration<double>(t1 - t0).count();
        }
#endif
        auto r1 = std::chrono::high_resolution_clock::now();
        std::cout << "Number of particles tracked " << tracked << std::endl;
        std::ofstream st(this->output_path + "/harness_stats.txt");
        st << "histories " << (size_t) h1 * opt.batches << "\ntransport_seconds " << transport_seconds
           << "\ninit_threads_seconds " << init_threads_seconds
           << "\nrun_seconds " << std::chrono::duration<double>(r1 - r0).count()
           << "\nblocks " << threads[1] << "\nthreads " << threads[0] << "\ntracked " << tracked << "\n";
        {
            mqi::coordinate_transform<R> pc(spot_angles, { 0, 0, 0 });   // the beam frame of build_beamsource
            const mqi::mat3x3<R>& m = pc.rotation;
            const R e[9] = { m.xx, m.xy, m.xz, m.yx, m.yy, m.yz, m.zx, m.zy, m.zz };
            st << "rot";
            for (int i = 0; i < 9; ++i) st << " " << std::setprecision(9) << e[i];
            st << "\n";
        }
        if (stat_value >= 0) st << "stat_value " << std::setprecision(17) << stat_value << "\nstat_count " << stat_count
                                << "\nstat_dose_max " << stat_dose_max << "\n";
    }

    double stat_value = -1, stat_dose_max = 0;
    long   stat_count = 0;
#if defined(__CUDACC__)
    // calculate_stat of tps_env (mqi_tps_env.hpp:1340-1426; that class needs GDCM and cannot be built): the reference's
    // own calculate_standard_deviation kernel (kernel_functions/mqi_variables.hpp:20-48) launched as calculate_stat
    // launches it for num_total_threads < 0 (:1371-1379), then the host reduction of :1409-1425 restated here.
    void
    stat_on_device(int n_histories) {
        const int c_ind = this->world->n_children - 1;
        const int roi   = this->world->children[c_ind]->scorers[1]->roi_->get_mask_size();
        std::vector<float> sd(roi, 0.f), mean(roi);   // dose_mean is uploaded uninitialised by the reference; zeros here
        float *            d_sd, *d_mean;
        gpu_err_chk(cudaMalloc(&d_sd, sizeof(float) * roi));
        gpu_err_chk(cudaMalloc(&d_mean, sizeof(float) * roi));
        gpu_err_chk(cudaMemset(d_sd, 0, sizeof(float) * roi));
        gpu_err_chk(cudaMemset(d_mean, 0, sizeof(float) * roi));
        uint32_t n_threads = mqi::thread_limit;
        uint32_t n_blocks  = (int) std::ceil(roi * 1.0 / n_threads);
        if (n_blocks > mqi::block_limit) n_blocks = mqi::block_limit;
        uint32_t vpt = (int) std::ceil(roi * 1.0 / (n_threads * n_blocks));
        if (roi % (n_threads * n_blocks * vpt) > 0) n_blocks += 1;
        calculate_standard_deviation<R><<<n_blocks, n_threads>>>(mc::mc_world, d_sd, d_mean, n_histories, roi, c_ind);
        gpu_err_chk(cudaDeviceSynchronize());
        gpu_err_chk(cudaMemcpy(sd.data(), d_sd, sizeof(float) * roi, cudaMemcpyDeviceToHost));
        gpu_err_chk(cudaMemcpy(mean.data(), d_mean, sizeof(float) * roi, cudaMemcpyDeviceToHost));
        cudaFree(d_sd);
        cudaFree(d_mean);
        double current = 0, dmax = 0;
        long   count = 0;
        for (int i = 0; i < roi; ++i)
            if (dmax < mean[i]) dmax = mean[i];
        for (int i = 0; i < roi; ++i)
            if (mean[i] > dmax * opt.stat_threshold) {
                current += (sd[i] / mean[i]);
                count += 1;
            }
        stat_value    = (float) (current / count);
        stat_count    = count;
        stat_dose_max = dmax;
        std::ofstream a(this->output_path + "/stat_sd.raw", std::ios::binary);
        a.write((const char*) sd.data(), sizeof(float) * roi);
        std::ofstream b(this->output_path + "/stat_mean.raw", std::ios::binary);
        b.write((const char*) mean.data(), sizeof(float) * roi);
        printf("stat %.9g count %ld dose_max %.9g n %d\n", stat_value, count, dmax, n_histories);
    }
#endif

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
        } else if (a == "--spot_grid" && i + 3 < argc) {
            o.grid[0]    = std::stoi(argv[++i]);
            o.grid[1]    = std::stoi(argv[++i]);
            o.grid_pitch = std::stof(argv[++i]);
        } else if (a == "--energy_step" && i + 1 < argc) {
            o.energy_step = std::stof(argv[++i]);
        } else if (a == "--batches" && i + 1 < argc) {
            o.batches = std::stoi(argv[++i]);
        } else if (a == "--sample_threads" && i + 1 < argc) {
            o.sample_threads = std::stoi(argv[++i]);
        } else if (a == "--stat_threshold" && i + 1 < argc) {
            o.stat_threshold = std::stof(argv[++i]);
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
