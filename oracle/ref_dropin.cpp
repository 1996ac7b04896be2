// Drop-in proof from the reference's side: the reference's OWN phantom_env (headers included from /root/reference,
// never copied; host compiler only) with run() routed through the C ABI of libmqi_b200.so exactly as INTEGRATION.md
// section 1 shows a maintainer how to do it.  Everything around the hot path is the reference's code: the CLI, the
// world and density set-up (hu_to_density), the host beam sampler, finalize() and save_reshaped_files(), which
// writes <prefix>/0_water_dE_total.raw from the reference's scorer table.  Test infrastructure
// (oracle/_ref/ref_dropin_<variant>, built by oracle/build_ref.sh; tests/test_gpu_dropin.py).
//
//   ref_dropin <phantom_env flags...> [--dump_vertices file]     file: the sampled vertex_t<float>[n] as raw bytes
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <moqui/base/environments/mqi_phantom_env.hpp>

#include "mqi_b200.h"

class dropin_env : public mqi::phantom_env<float>
{
public:
    typedef float R;
    std::string   dump_vertices;
    dropin_env(mqi::cli& c) : mqi::phantom_env<float>(c) {}

    // replaces mqi_phantom_env.hpp:301-428
    virtual void
    run() {
        const uint32_t h1 = this->beamsource.total_histories();
        static_assert(sizeof(mqi::vertex_t<float>) == sizeof(mqi_vertex), "vertex_t<float> {ke, pos, dir} is mqi_vertex");
        if (!dump_vertices.empty())
            std::ofstream(dump_vertices, std::ios::binary).write((const char*) this->vertices, (size_t) h1 * sizeof(mqi::vertex_t<float>));
        mqi_handle* h = nullptr;
        if (mqi_create(this->gpu_id, &h)) throw std::runtime_error(mqi_last_error());
#ifdef __PHYSICS_DEBUG__
        mqi_set_physics(h, MQI_PHYSICS_DEBUG, 0);
#else
        mqi_set_physics(h, MQI_PHYSICS_RELEASE, 0);
#endif
        mqi::grid3d<mqi::density_t, R>* g = this->world->children[0]->geo;   // mqi_phantom_env.hpp:263-264
        const mqi::vec3<mqi::ijk_t>     n = g->get_nxyz();
        const size_t                    nvox = (size_t) n.x * n.y * n.z;
        // edges and densities exactly as setup_world() produced them
        if (mqi_set_grid_density(h, g->get_x_edges(), n.x + 1, g->get_y_edges(), n.y + 1, g->get_z_edges(), n.z + 1, g->get_data(),
                                 nullptr, nullptr))
            throw std::runtime_error(mqi_last_error());
        const int s = mqi_add_scorer(h, MQI_SCORER_DOSE, "water_dE_total", nvox);
        if (s < 0) throw std::runtime_error(mqi_last_error());
        if (mqi_set_vertices(h, reinterpret_cast<const mqi_vertex*>(this->vertices), h1, nullptr)) throw std::runtime_error(mqi_last_error());
        if (mqi_run(h, (uint64_t) this->random_seed, 0, h1, 0)) throw std::runtime_error(mqi_last_error());
        mqi_run_stats st;
        mqi_get_run_stats(h, &st);
        std::cout << "Number of particles tracked " << st.histories << std::endl;   // :424
        printf("Run done %f s\n", st.kernel_ms * 0.0001);                          // the reference's scale, :427
        // the dense result goes into the reference's scorer table the way insert_hashtable's dense mode leaves it
        // (slot = voxel, key1 = voxel, key2 = 0, mqi_transport.hpp:78-111): finalize() and save_reshaped_files() then
        // run unchanged
        std::vector<double> dose(nvox);
        if (mqi_get_dense(h, s, dose.data(), 1.0)) throw std::runtime_error(mqi_last_error());
        mqi::key_value* t = this->world->children[0]->scorers[0]->data_;
        for (size_t i = 0; i < nvox; ++i)
            if (dose[i] > 0.0) {
                t[i].key1  = (mqi::key_t) i;
                t[i].key2  = 0;
                t[i].value = dose[i];
            }
        mqi_destroy(h);
    }
};

int
main(int argc, char* argv[]) {
    std::string        dump;
    std::vector<char*> pass;
    pass.push_back(argv[0]);
    for (int i = 1; i < argc; ++i) {
        if (std::string(argv[i]) == "--dump_vertices" && i + 1 < argc) dump = argv[++i];
        else pass.push_back(argv[i]);
    }
    // from here on: tests/mc/phantom/phantom_env.cpp:6-22
    mqi::cli cl_opts;
    cl_opts.read((int) pass.size(), pass.data());
    dropin_env myenv(cl_opts);
    myenv.dump_vertices = dump;
    myenv.initialize();
    myenv.run();
    myenv.finalize();
    myenv.save_reshaped_files();
    return 0;
}
