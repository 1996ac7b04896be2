// mqi_capi.cu -- the C ABI of include/mqi_b200.h on top of the sm_100a kernels.
//
// Host-side responsibilities only: device buffers, the calibration LUT (HU -> density -> RSP /
// radiation-length coefficients), launch configuration, downloads.  No transport arithmetic runs on
// the host and there is no CPU fallback: without a usable CUDA device every compute entry point
// returns MQI_ENODEVICE.
#include "../../include/mqi_b200.h"
#include "mqi_device.cuh"
#include "mqi_kernels.h"
#include "mqi_roi.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace mqib;

namespace
{
#include "mqi_tables_data.inc"   // generated from moquimc_b200/data/mqi_tables_v1.bin by build.py

thread_local std::string g_err;

int
fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(MQI_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
    } while (0)

struct HostScorer {
    int         kind = 0;
    std::string name;
    uint64_t    capacity = 0;
    double*     d_dense  = nullptr;
    bool        external = false;
    DijSlot*    d_table  = nullptr;
    uint32_t*   d_roi    = nullptr;   // CONTOUR roi as one bit per voxel, nullptr = DIRECT roi
    uint64_t    roi_size = 0;         // voxels inside the roi (get_mask_size)
};

// one child of the world in front of the scored grid (range shifter, aperture)
struct HostNode {
    int                nx = 0, ny = 0, nz = 0;
    std::vector<float> edges;   // xe | ye | ze
    uint16_t*          d_mat = nullptr;
    MatEntry*          d_lut = nullptr;
    int                lut_size = 0;
    float              rot[9]   = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    float              trans[3] = { 0, 0, 0 };
    int                identity = 1;
    float              inv_w[3] = { 0, 0, 0 };
};

const float* table_ptr(int t) { return reinterpret_cast<const float*>(k_tables_blob + 16) + 600 * t; }
const float* correction_ptr() { return reinterpret_cast<const float*>(k_tables_blob + 16) + 3600; }

// ---- calibration on the host side, only to fill the LUT -----------------------------------------
// patient_material_t::hu_to_density  materials/mqi_patient_materials.hpp:514-542
float
host_hu_to_density(int hu) {
    hu = hu < -1000 ? -1000 : (hu > 2995 ? 2995 : hu);
    float rho_mass;
    if (hu < -98) rho_mass = 0.00121 + 0.001029700665188 * (1000.0 + hu);
    else if (hu < 15) rho_mass = 1.018 + 0.000893 * hu;
    else if (hu < 23) rho_mass = 1.03;
    else if (hu < 101) rho_mass = 1.003 + 0.001169 * hu;
    else if (hu < 2001) rho_mass = 1.017 + 0.000592 * hu;
    else if (hu < 2995) rho_mass = 2.201 + 0.0005 * (-2000.0 + hu);
    else rho_mass = 4.54;
    rho_mass *= correction_ptr()[hu + 1000];
    rho_mass /= 1000.0;
    return rho_mass;
}

inline float
lerp_ref(float x, float x0, float x1, float y0, float y1) {
    return (x1 == x0) ? y0 : y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}

// spr_default / radiation_length_default (:414-473): the density-only parts, evaluated once per
// distinct density in the reference's precision (R = float variables, double literals)
MatEntry
make_mat_entry(float rho, int variant) {
    MatEntry m;
    std::memset(&m, 0, sizeof(m));
    m.rho     = rho;
    m.inv_rho = rho > 0.f ? 1.0f / rho : 0.f;
    const float d = rho * 1000.0;
    const bool  water_shortcut = (variant == MQI_PHYSICS_DEBUG) && std::fabs(d - 1.0) < 1e-3;
    if (water_shortcut) {
        m.mode = 0;
        m.a    = 1.0f;
    } else if (d <= 0.26) {
        m.mode = 0;
        m.a    = d < 0.0012 ? 0.0f : lerp_ref(d, 0.0012f, 0.26f, 0.8815f, 0.9925f);
    } else {
        const float pd = powf(d, -0.7f);
        m.P = (float) ((double) pd - 1.0);
        if (d >= 0.9) {
            m.mode = 1;
            m.a    = pd;   // rsp_eval_exact: above 2.69 g/cm3 (pd < 0.5) pd - 1 is not representable in fp32
        } else {
            m.mode = 2;
            m.a    = d - 0.26f;
        }
    }
    float x0;
    if (water_shortcut) {
        x0 = 360.863f;
    } else {
        float f = 0.f;
        if (d <= 0.26) f = 0.9857 + 0.0085 * d;
        else if (d <= 0.9) f = 1.0446 - 0.2180 * d;
        else f = 1.19 + 0.44 * std::log((double) d - 0.44);
        x0 = (0.001f * 360.863f) / (d * 0.001 * f);
    }
    m.x0       = x0;
    m.inv_x0   = 1.0f / x0;
    m.inv_rsp0 = (m.mode == 0 && m.a > 0.f) ? 1.0f / m.a : 0.f;
    return m;
}

float
dedx_term0() {   // physics_constants::two_pi_re2_mc2_h2o  base/mqi_physics_constants.hpp:30-33
    const float cm3            = 10.0f * 10.0f * 10.0f;
    const float re             = 2.8179403262e-12;
    const float re_sq          = re * re;
    const float two_pi_re2_mc2 = 2.0 * M_PI * re_sq * 0.510998928f;
    return two_pi_re2_mc2 * 3.3428e+23 / cm3;
}
}   // namespace

struct mqi_handle {
    int          device = 0;
    cudaStream_t stream = nullptr;       // the stream kernels are launched on
    cudaStream_t own_stream = nullptr;   // created by mqi_create
    bool         pending = false;        // an mqi_run_async launch has not been collected yet
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    int          sm_count = 148;
    int          variant  = MQI_PHYSICS_RELEASE;
    uint32_t     quirks   = 0;
    int          accum    = MQI_ACCUM_ATOMIC;
    int          count_steps = 0;
    int          fetch_order = 0;              // option "fetch_order": 1 = the chunks of a launch from the last to the first, 0 = first to last
    int          dij_write_combine = 1;        // option "dij_write_combine": consecutive hits of a lane on one (voxel, spot) key are summed in registers and inserted once
    int          rsp_exact = 0;                // option "rsp_exact": mqi_dev_rsp evaluates spr_default in the reference's precision (bit-exact KAT)
    int          blocks_per_sm_override = 0;
    int          l2_persist = 0;               // option "l2_persist": pin the material volume in L2 with a persisting access window (measured: no gain at C1)
    size_t       l2_persist_max = 0, l2_window_max = 0;
    bool         l2_window_set = false;   // this handle has put an access-policy window on its stream
    // physics tables
    float4* d_tab_a0 = nullptr;
    float4* d_tab_a1 = nullptr;
    float2* d_tab_bs = nullptr;
    float4* d_tab_n0 = nullptr;
    float2* d_tab_n1 = nullptr;
    float*  d_correction = nullptr;
    // grid
    bool      has_grid = false;
    int       nx = 0, ny = 0, nz = 0;
    float*    d_edges = nullptr;
    uint16_t* d_mat   = nullptr;
    MatEntry* d_lut   = nullptr;
    int       lut_size = 0;
    float     rot[9]   = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    float     trans[3] = { 0, 0, 0 };
    int       identity = 1;
    float     inv_w[3] = { 0, 0, 0 };
    std::vector<float> edges_host;
    // beamline children in transport order (multi-node launches), and their device mirror
    std::vector<HostNode> beamline;
    GridDev*  d_nodes = nullptr;
    uint32_t* d_adv_raw = nullptr;      // hand-over buffers of the multi-node kernels (transport_handover_bytes), grown on demand
    size_t    adv_raw_bytes = 0;
    float*    d_edges_all = nullptr;
    int       n_edge_floats_all = 0;
    bool      nodes_dirty = true;
    // source
    BeamletDev*         d_beamlets = nullptr;
    unsigned long long* d_cum      = nullptr;
    uint32_t            n_spots    = 0;
    uint64_t            total_histories = 0;
    VertexDev*          d_vertices = nullptr;
    uint32_t*           d_spot_ids = nullptr;
    uint64_t            n_vertices = 0;
    // scorers
    std::vector<HostScorer> scorers;
    unsigned long long*     d_counters = nullptr;
    // device buffers released by a geometry change, kept for the next one of the same size: a caller that
    // re-uploads the CT per run (x_environment::initialize per beam, bench.py's end-to-end leg) then pays no
    // cudaMalloc / cudaFree (both synchronise the device) per run
    std::multimap<size_t, void*> pool;
    size_t                       pool_bytes = 0;
    // this device's slice of the summed stat grids (mqi_stat_multi): sum | sum of squares, stat_slice_n doubles each
    double*                      d_stat_slice = nullptr;
    size_t                       stat_slice_n = 0;
    unsigned char*               d_stat_flags = nullptr;   // chunk flags and list of kept chunks (mqi_stat_multi)
    unsigned int*                d_stat_list  = nullptr;
    size_t                       stat_flags_n = 0;
    double*                      d_stat3 = nullptr;   // scratch of the stopping criterion: sum of sigma/mu, count, max mean
                                                       // (kept in the handle: cudaMalloc / cudaFree per evaluation cost milliseconds)
    mqi_run_stats           stats {};
};

namespace
{
size_t nvox(const mqi_handle* h) { return (size_t) h->nx * h->ny * h->nz; }

int
activate(mqi_handle* h) {
    if (!h) return fail(MQI_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    return MQI_OK;
}

template<typename T>
cudaError_t
pool_alloc(mqi_handle* h, T** p, size_t bytes) {
    bytes = std::max<size_t>(bytes, 1);
    auto it = h->pool.find(bytes);
    if (it != h->pool.end()) {
        *p = static_cast<T*>(it->second);
        h->pool_bytes -= bytes;
        h->pool.erase(it);
        return cudaSuccess;
    }
    return cudaMalloc(p, bytes);
}

void
pool_free(mqi_handle* h, void* p, size_t bytes) {
    if (!p) return;
    bytes = std::max<size_t>(bytes, 1);
    if (h->pool.size() >= 16 || h->pool_bytes + bytes > (4ull << 30)) {
        cudaFree(p);
        return;
    }
    h->pool.emplace(bytes, p);
    h->pool_bytes += bytes;
}

void
pool_clear(mqi_handle* h) {
    for (auto& kv : h->pool) cudaFree(kv.second);
    h->pool.clear();
    h->pool_bytes = 0;
}

void
free_grid(mqi_handle* h) {   // sizes are those of the grid being dropped (h->nx ... are still its dimensions)
    pool_free(h, h->d_edges, (size_t) (h->nx + h->ny + h->nz + 3) * sizeof(float));
    pool_free(h, h->d_mat, nvox(h) * sizeof(uint16_t));
    pool_free(h, h->d_lut, (size_t) h->lut_size * sizeof(MatEntry));
    h->d_edges = nullptr;
    h->d_mat   = nullptr;
    h->d_lut   = nullptr;
    h->has_grid = false;
}

int
set_grid_common(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze, int n_ze,
                const float* rot, const float* trans) {
    if (!xe || !ye || !ze || n_xe < 2 || n_ye < 2 || n_ze < 2) return fail(MQI_EINVAL, "bad grid edges");
    const size_t nv = (size_t) (n_xe - 1) * (n_ye - 1) * (n_ze - 1);
    if (nv >= 0x7fffffffull) return fail(MQI_EINVAL, "grid has more than 2^31-1 voxels (scorer keys are 32-bit, B6)");
    for (int i = 1; i < n_xe; ++i) if (!(xe[i] > xe[i - 1])) return fail(MQI_EINVAL, "x edges must increase");
    for (int i = 1; i < n_ye; ++i) if (!(ye[i] > ye[i - 1])) return fail(MQI_EINVAL, "y edges must increase");
    for (int i = 1; i < n_ze; ++i) if (!(ze[i] > ze[i - 1])) return fail(MQI_EINVAL, "z edges must increase");
    if (transport_smem_bytes(n_xe + n_ye + n_ze, 1) > 200 * 1024) return fail(MQI_EINVAL, "too many grid edges for shared memory");
    // scorers are sized by the grid: drop dense accumulators of a previous grid
    for (auto& s : h->scorers) {
        if (s.d_dense && !s.external) pool_free(h, s.d_dense, nvox(h) * sizeof(double));
        if (!s.external) s.d_dense = nullptr;
        cudaFree(s.d_roi);   // a region of interest belongs to the grid it was defined on
        s.d_roi    = nullptr;
        s.roi_size = 0;
    }
    free_grid(h);
    h->nodes_dirty = true;
    h->nx = n_xe - 1; h->ny = n_ye - 1; h->nz = n_ze - 1;
    std::vector<float> e;
    e.insert(e.end(), xe, xe + n_xe);
    e.insert(e.end(), ye, ye + n_ye);
    e.insert(e.end(), ze, ze + n_ze);
    h->edges_host = e;
    CU(pool_alloc(h, &h->d_edges, e.size() * sizeof(float)));
    CU(cudaMemcpyAsync(h->d_edges, e.data(), e.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->inv_w[0] = (float) (n_xe - 1) / (xe[n_xe - 1] - xe[0]);
    h->inv_w[1] = (float) (n_ye - 1) / (ye[n_ye - 1] - ye[0]);
    h->inv_w[2] = (float) (n_ze - 1) / (ze[n_ze - 1] - ze[0]);
    const float I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    std::memcpy(h->rot, rot ? rot : I, sizeof(I));
    h->trans[0] = trans ? trans[0] : 0.f; h->trans[1] = trans ? trans[1] : 0.f; h->trans[2] = trans ? trans[2] : 0.f;
    h->identity = (std::memcmp(h->rot, I, sizeof(I)) == 0 && h->trans[0] == 0.f && h->trans[1] == 0.f && h->trans[2] == 0.f) ? 1 : 0;
    return MQI_OK;
}

int
upload_hu_lut(mqi_handle* h, float density_scale) {
    std::vector<MatEntry> lut(3996);
    for (int i = 0; i < 3996; ++i) {
        float rho = host_hu_to_density(i - 1000);
        if (density_scale != 1.0f) rho *= density_scale;
        lut[i] = make_mat_entry(rho, h->variant);
    }
    CU(pool_alloc(h, &h->d_lut, lut.size() * sizeof(MatEntry)));
    CU(cudaMemcpyAsync(h->d_lut, lut.data(), lut.size() * sizeof(MatEntry), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->lut_size = (int) lut.size();
    return MQI_OK;
}

int
ensure_scorer_buffers(mqi_handle* h) {
    for (auto& s : h->scorers) {
        if (s.kind == MQI_SCORER_DIJ) {
            if (!s.d_table) {
                CU(cudaMalloc(&s.d_table, s.capacity * sizeof(DijSlot)));
                CU(launch_dij_clear(s.d_table, s.capacity, h->stream));
            }
        } else if (!s.d_dense) {
            CU(pool_alloc(h, &s.d_dense, nvox(h) * sizeof(double)));
            CU(cudaMemsetAsync(s.d_dense, 0, nvox(h) * sizeof(double), h->stream));
        }
    }
    return MQI_OK;
}

void
fill_params(const mqi_handle* h, Params& p) {
    std::memset(&p, 0, sizeof(p));
    p.g.nx = h->nx; p.g.ny = h->ny; p.g.nz = h->nz;
    p.g.edges = h->d_edges;
    p.g.mat   = h->d_mat;
    p.g.lut   = h->d_lut;
    p.g.lut_size = h->lut_size;
    p.g.identity = h->identity;
    std::memcpy(p.g.rot_fwd, h->rot, sizeof(h->rot));
    std::memcpy(p.g.trans, h->trans, sizeof(h->trans));
    std::memcpy(p.g.inv_w, h->inv_w, sizeof(h->inv_w));
    p.g.edge_off = 0;
    if (h->beamline.empty()) {
        p.nodes         = nullptr;
        p.n_nodes       = 1;
        p.edges_all     = h->d_edges;
        p.n_edge_floats = h->nx + h->ny + h->nz + 3;
    } else {   // mirror built by sync_nodes()
        p.nodes         = h->d_nodes;
        p.n_nodes       = (int) h->beamline.size() + 1;
        p.edges_all     = h->d_edges_all;
        p.n_edge_floats = h->n_edge_floats_all;
    }
    p.src.beamlets = h->d_beamlets;
    p.src.cum      = h->d_cum;
    p.src.n_spots  = h->n_spots;
    p.src.vertices = h->d_vertices;
    p.src.spot_ids = h->d_spot_ids;
    p.n_scorers = (int) h->scorers.size();
    for (int i = 0; i < p.n_scorers; ++i) {
        p.sc[i].kind     = h->scorers[i].kind;
        p.sc[i].dense    = h->scorers[i].d_dense;
        p.sc[i].table    = h->scorers[i].d_table;
        p.sc[i].capacity = h->scorers[i].capacity;
        p.sc[i].cap_magic = remainder_magic(h->scorers[i].capacity);
        p.sc[i].roi      = h->scorers[i].d_roi;
    }
    p.n_shards    = 1;
    p.shard       = 0;
    // option "fetch_order" = 1: the launch hands out its chunks from the last to the first, so that a plan listed by ascending
    // energy starts its longest histories first and the tail of the persistent kernel is made of the short ones -- worth it
    // only for launches of a few histories per lane: on whole plans it measured SLOWER (C3 pass -2.8 %, C4 -2 ... -20 %)
    p.reverse     = h->fetch_order == 1 ? 1 : 0;
    p.adv_raw     = nullptr;
    p.quirks      = h->quirks;
    p.accum_mode  = h->accum;
    p.count_steps = h->count_steps;
    p.dij_wc_scorer = -1;
    if (h->dij_write_combine)
        for (int i = 0; i < p.n_scorers; ++i)
            if (p.sc[i].kind == MQI_SCORER_DIJ) { p.dij_wc_scorer = i; break; }
    p.dedx_term0  = dedx_term0();
    p.tab_a0      = h->d_tab_a0;
    p.tab_a1      = h->d_tab_a1;
    p.tab_bs      = h->d_tab_bs;
    p.tab_n0      = h->d_tab_n0;
    p.tab_n1      = h->d_tab_n1;
    p.counters    = h->d_counters;
}

// raw densities -> 16-bit material indices + LUT of the distinct values
int
build_density_lut(const float* rho, size_t nv, int variant, std::vector<uint16_t>& mat, std::vector<MatEntry>& lut) {
    std::map<uint32_t, uint16_t> dict;
    mat.resize(nv);
    lut.clear();
    for (size_t i = 0; i < nv; ++i) {
        uint32_t bits;
        std::memcpy(&bits, &rho[i], 4);
        auto it = dict.find(bits);
        if (it == dict.end()) {
            if (lut.size() >= 65536) return fail(MQI_EINVAL, "more than 65536 distinct densities; pass HU instead");
            it = dict.emplace(bits, (uint16_t) lut.size()).first;
            lut.push_back(make_mat_entry(rho[i], variant));
        }
        mat[i] = it->second;
    }
    return MQI_OK;
}

void
free_beamline(mqi_handle* h) {
    for (auto& n : h->beamline) {
        cudaFree(n.d_mat);
        cudaFree(n.d_lut);
    }
    h->beamline.clear();
    cudaFree(h->d_nodes);
    cudaFree(h->d_edges_all);
    h->d_nodes     = nullptr;
    h->d_edges_all = nullptr;
    h->nodes_dirty = true;
}

// device mirror of the world's children for multi-node launches: descriptors + all edges back to back
int
sync_nodes(mqi_handle* h) {
    if (h->beamline.empty() || !h->nodes_dirty) return MQI_OK;
    std::vector<GridDev> nodes;
    std::vector<float>   edges;
    auto add = [&](int nx, int ny, int nz, const std::vector<float>& e, const uint16_t* mat, const MatEntry* lut, int lut_size,
                   const float* rot, const float* trans, int identity, const float* inv_w) {
        GridDev g;
        std::memset(&g, 0, sizeof(g));
        g.nx = nx; g.ny = ny; g.nz = nz;
        g.edges = nullptr;
        g.mat = mat; g.lut = lut; g.lut_size = lut_size; g.identity = identity;
        std::memcpy(g.rot_fwd, rot, 9 * sizeof(float));
        std::memcpy(g.trans, trans, 3 * sizeof(float));
        std::memcpy(g.inv_w, inv_w, 3 * sizeof(float));
        g.edge_off = (int) edges.size();
        edges.insert(edges.end(), e.begin(), e.end());
        nodes.push_back(g);
    };
    for (auto& n : h->beamline) add(n.nx, n.ny, n.nz, n.edges, n.d_mat, n.d_lut, n.lut_size, n.rot, n.trans, n.identity, n.inv_w);
    add(h->nx, h->ny, h->nz, h->edges_host, h->d_mat, h->d_lut, h->lut_size, h->rot, h->trans, h->identity, h->inv_w);
    cudaFree(h->d_nodes);
    cudaFree(h->d_edges_all);
    h->d_nodes = nullptr; h->d_edges_all = nullptr;
    CU(cudaMalloc(&h->d_nodes, nodes.size() * sizeof(GridDev)));
    CU(cudaMalloc(&h->d_edges_all, edges.size() * sizeof(float)));
    CU(cudaMemcpyAsync(h->d_nodes, nodes.data(), nodes.size() * sizeof(GridDev), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_edges_all, edges.data(), edges.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->n_edge_floats_all = (int) edges.size();
    h->nodes_dirty       = false;
    return MQI_OK;
}

template<typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};
}   // namespace

extern "C" {
static int collect_run(mqi_handle* h);
static int collect_run_fwd(mqi_handle* h) { return collect_run(h); }

const char* mqi_last_error(void) { return g_err.c_str(); }
const char* mqi_version(void) { return "moquimc_b200 0.1 (sm_100a)"; }

int
mqi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int
mqi_device_memory(mqi_handle* h, uint64_t* free_bytes, uint64_t* total_bytes) {
    if (!h) return fail(MQI_EINVAL, "handle is null");
    CU(cudaSetDevice(h->device));
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (uint64_t) f;
    if (total_bytes) *total_bytes = (uint64_t) t;
    return MQI_OK;
}

static int create_impl(mqi_handle* h, int device_id);

int
mqi_create(int device_id, mqi_handle** out) {
    if (!out) return fail(MQI_EINVAL, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MQI_ENODEVICE, "no usable CUDA device (this library has no CPU fallback)");
    }
    if (device_id < 0 || device_id >= n) return fail(MQI_EINVAL, "device id out of range");
    if (std::memcmp(k_tables_blob, "MQITBL1", 7) != 0) return fail(MQI_ESTATE, "embedded physics tables are corrupt");
    mqi_handle* h = new mqi_handle;
    h->device     = device_id;
    const int rc  = create_impl(h, device_id);
    if (rc != MQI_OK) {   // a failed step must not leak the handle, its stream and events or the tables already uploaded
        const std::string why = g_err;
        mqi_destroy(h);
        return fail(rc, why);
    }
    *out = h;
    return MQI_OK;
}

static int
create_impl(mqi_handle* h, int device_id) {
    CU(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device_id));
    h->sm_count = prop.multiProcessorCount;
    h->l2_persist_max = (size_t) prop.persistingL2CacheMaxSize;
    if (const char* e = std::getenv("MQI_L2_PERSIST")) h->l2_persist = std::atoi(e);   // tuning aid: default of the option "l2_persist"
    h->l2_window_max  = (size_t) prop.accessPolicyMaxWindowSize;
    CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    // physics tables -> value + slope rows (one shared-memory row per interpolation); tables 0-2 live
    // on the p-ionisation grid (Ei = 0.1), 3-5 on the nuclear grid (Ei = 0.5), step 0.5 MeV
    std::vector<float4> a0(kTableN), a1(kTableN), n0(kTableN);
    std::vector<float2> bs(kTableN), n1(kTableN);
    auto slope = [](const float* t, int i) { return i + 1 < kTableN ? (t[i + 1] - t[i]) * 2.0f : 0.0f; };
    for (int i = 0; i < kTableN; ++i) {
        const float* cs = table_ptr(0); const float* sp = table_ptr(1); const float* rg = table_ptr(2);
        const float* pp = table_ptr(3); const float* pe = table_ptr(4); const float* pi = table_ptr(5);
        const float dedr = (i + 1 < kTableN && rg[i + 1] > rg[i]) ? 0.5f / (rg[i + 1] - rg[i]) : 0.0f;
        a0[i] = make_float4(cs[i], slope(cs, i), sp[i], slope(sp, i));
        a1[i] = make_float4(rg[i], slope(rg, i), dedr, 0.f);
        n0[i] = make_float4(pp[i], slope(pp, i), pe[i], slope(pe, i));
        n1[i] = make_float2(pi[i], slope(pi, i));
        bs[i] = make_float2(pp[i] + pe[i] + pi[i], slope(pp, i) + slope(pe, i) + slope(pi, i));
    }
    CU(cudaMalloc(&h->d_tab_a0, kTableN * sizeof(float4)));
    CU(cudaMalloc(&h->d_tab_a1, kTableN * sizeof(float4)));
    CU(cudaMalloc(&h->d_tab_n0, kTableN * sizeof(float4)));
    CU(cudaMalloc(&h->d_tab_bs, kTableN * sizeof(float2)));
    CU(cudaMalloc(&h->d_tab_n1, kTableN * sizeof(float2)));
    CU(cudaMemcpy(h->d_tab_a0, a0.data(), kTableN * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_tab_a1, a1.data(), kTableN * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_tab_n0, n0.data(), kTableN * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_tab_bs, bs.data(), kTableN * sizeof(float2), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_tab_n1, n1.data(), kTableN * sizeof(float2), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&h->d_correction, 3996 * sizeof(float)));
    CU(cudaMalloc(&h->d_counters, C_COUNT * sizeof(unsigned long long)));
    CU(cudaMemcpy(h->d_correction, correction_ptr(), 3996 * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemset(h->d_counters, 0, C_COUNT * sizeof(unsigned long long)));
    return MQI_OK;
}

int
mqi_destroy(mqi_handle* h) {
    if (!h) return MQI_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_grid(h);
    pool_clear(h);
    free_beamline(h);
    for (auto& s : h->scorers) {
        if (s.d_dense && !s.external) cudaFree(s.d_dense);
        cudaFree(s.d_table);
        cudaFree(s.d_roi);
    }
    cudaFree(h->d_tab_a0); cudaFree(h->d_tab_a1); cudaFree(h->d_tab_bs); cudaFree(h->d_tab_n0); cudaFree(h->d_tab_n1); cudaFree(h->d_correction); cudaFree(h->d_counters);
    cudaFree(h->d_beamlets); cudaFree(h->d_cum); cudaFree(h->d_vertices); cudaFree(h->d_spot_ids);
    cudaFree(h->d_adv_raw);
    cudaFree(h->d_stat_slice);
    cudaFree(h->d_stat_flags);
    cudaFree(h->d_stat_list);
    cudaFree(h->d_stat3);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    cudaGetLastError();
    delete h;
    return MQI_OK;
}

int
mqi_set_physics(mqi_handle* h, int variant, uint32_t quirks) {
    if (!h) return fail(MQI_EINVAL, "null handle");
    if (variant != MQI_PHYSICS_RELEASE && variant != MQI_PHYSICS_DEBUG) return fail(MQI_EINVAL, "unknown physics variant");
    if (h->has_grid && variant != h->variant) return fail(MQI_ESTATE, "set the physics variant before the grid (the calibration LUT depends on it)");
    h->variant = variant;
    h->quirks  = quirks;
    return MQI_OK;
}

static int
set_grid_hu_impl(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze, int n_ze,
                 const void* hu, bool on_device, float density_scale, const float* rot, const float* trans) {
    int rc = activate(h);
    if (rc) return rc;
    if (!hu) return fail(MQI_EINVAL, "hu is null");
    rc = set_grid_common(h, xe, n_xe, ye, n_ye, ze, n_ze, rot, trans);
    if (rc) return rc;
    const size_t nv = nvox(h);
    CU(pool_alloc(h, &h->d_mat, nv * sizeof(uint16_t)));
    if (on_device) {
        // the conversion kernel reads the volume as 16-byte vectors
        if (reinterpret_cast<uintptr_t>(hu) & 15u) return fail(MQI_EINVAL, "the device HU volume must be 16-byte aligned");
        CU(launch_hu_to_material(static_cast<const int16_t*>(hu), h->d_mat, nv, h->stream));
    } else {
        int16_t* tmp = nullptr;   // staging buffer of the HU upload (the size differs from d_mat's by the +1 below, so the pool tells them apart)
        CU(pool_alloc(h, &tmp, nv * sizeof(int16_t) + 1));
        // a failed copy or launch hands the staging buffer back before the error is returned
        cudaError_t e = cudaMemcpyAsync(tmp, hu, nv * sizeof(int16_t), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = launch_hu_to_material(tmp, h->d_mat, nv, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        pool_free(h, tmp, nv * sizeof(int16_t) + 1);
        if (e != cudaSuccess) return fail(MQI_ECUDA, std::string("HU volume upload: ") + cudaGetErrorString(e));
    }
    rc = upload_hu_lut(h, density_scale);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    h->has_grid = true;
    return MQI_OK;
}

int
mqi_set_grid_hu(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze, int n_ze,
                const int16_t* hu, float density_scale, const float* rot, const float* trans) {
    return set_grid_hu_impl(h, xe, n_xe, ye, n_ye, ze, n_ze, hu, false, density_scale, rot, trans);
}

int
mqi_set_grid_hu_device(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze,
                       int n_ze, const void* d_hu, float density_scale, const float* rot, const float* trans) {
    return set_grid_hu_impl(h, xe, n_xe, ye, n_ye, ze, n_ze, d_hu, true, density_scale, rot, trans);
}

int
mqi_set_grid_density(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze, int n_ze,
                     const float* rho, const float* rot, const float* trans) {
    int rc = activate(h);
    if (rc) return rc;
    if (!rho) return fail(MQI_EINVAL, "rho is null");
    rc = set_grid_common(h, xe, n_xe, ye, n_ye, ze, n_ze, rot, trans);
    if (rc) return rc;
    const size_t nv = nvox(h);
    std::vector<uint16_t> mat;
    std::vector<MatEntry> lut;
    rc = build_density_lut(rho, nv, h->variant, mat, lut);
    if (rc) return rc;
    CU(pool_alloc(h, &h->d_mat, nv * sizeof(uint16_t)));
    CU(pool_alloc(h, &h->d_lut, lut.size() * sizeof(MatEntry)));
    CU(cudaMemcpyAsync(h->d_mat, mat.data(), nv * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_lut, lut.data(), lut.size() * sizeof(MatEntry), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->lut_size = (int) lut.size();
    h->has_grid = true;
    return MQI_OK;
}

int
mqi_add_scorer(mqi_handle* h, int kind, const char* name, uint64_t capacity) {
    if (!h) return fail(MQI_EINVAL, "null handle");
    if (kind < MQI_SCORER_DOSE || kind > MQI_SCORER_LETT_DENOM) return fail(MQI_EINVAL, "unknown scorer kind");
    if ((int) h->scorers.size() >= kMaxScorers) return fail(MQI_EINVAL, "too many scorers");
    if (kind == MQI_SCORER_DIJ && capacity == 0) return fail(MQI_EINVAL, "Dij scorer needs a capacity");
    HostScorer s;
    s.kind     = kind;
    s.name     = name ? name : "";
    s.capacity = capacity;
    h->scorers.push_back(s);
    return (int) h->scorers.size() - 1;
}

int
mqi_bind_scorer_buffer(mqi_handle* h, int scorer, void* d_buffer) {
    if (!h || scorer < 0 || scorer >= (int) h->scorers.size()) return fail(MQI_EINVAL, "bad scorer index");
    HostScorer& s = h->scorers[scorer];
    if (s.kind == MQI_SCORER_DIJ) return fail(MQI_EINVAL, "Dij scorers own their table");
    if (s.d_dense && !s.external) cudaFree(s.d_dense);
    s.d_dense  = static_cast<double*>(d_buffer);
    s.external = d_buffer != nullptr;
    return MQI_OK;
}

int
mqi_set_scorer_roi(mqi_handle* h, int scorer, const uint8_t* mask_total, uint64_t n_voxels, uint64_t* roi_size) {
    int rc = activate(h);
    if (rc) return rc;
    if (scorer < 0 || scorer >= (int) h->scorers.size()) return fail(MQI_EINVAL, "bad scorer index");
    if (!h->has_grid) return fail(MQI_ESTATE, "set the grid before a region of interest");
    HostScorer& s = h->scorers[scorer];
    rc = collect_run_fwd(h);
    if (rc) return rc;
    cudaFree(s.d_roi);
    s.d_roi    = nullptr;
    s.roi_size = 0;
    if (!mask_total) {   // back to roi_t(DIRECT)
        if (roi_size) *roi_size = nvox(h);
        return MQI_OK;
    }
    if (n_voxels != nvox(h)) return fail(MQI_EINVAL, "mask size differs from the grid");
    const RoiRuns               runs = mask_to_roi(mask_total, n_voxels);
    const std::vector<uint32_t> bits = runs.bitmask();
    CU(cudaMalloc(&s.d_roi, std::max<size_t>(bits.size(), 1) * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(s.d_roi, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    s.roi_size = runs.size();
    if (roi_size) *roi_size = s.roi_size;
    return MQI_OK;
}

int
mqi_add_beamline_node(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze, int n_ze,
                      const float* rho, const float* rot, const float* trans) {
    int rc = activate(h);
    if (rc) return rc;
    if (!xe || !ye || !ze || n_xe < 2 || n_ye < 2 || n_ze < 2 || !rho) return fail(MQI_EINVAL, "bad beamline node");
    for (int i = 1; i < n_xe; ++i) if (!(xe[i] > xe[i - 1])) return fail(MQI_EINVAL, "x edges must increase");
    for (int i = 1; i < n_ye; ++i) if (!(ye[i] > ye[i - 1])) return fail(MQI_EINVAL, "y edges must increase");
    for (int i = 1; i < n_ze; ++i) if (!(ze[i] > ze[i - 1])) return fail(MQI_EINVAL, "z edges must increase");
    if ((int) h->beamline.size() >= 7) return fail(MQI_EINVAL, "too many beamline nodes");
    rc = collect_run_fwd(h);
    if (rc) return rc;
    HostNode n;
    n.nx = n_xe - 1; n.ny = n_ye - 1; n.nz = n_ze - 1;
    n.edges.insert(n.edges.end(), xe, xe + n_xe);
    n.edges.insert(n.edges.end(), ye, ye + n_ye);
    n.edges.insert(n.edges.end(), ze, ze + n_ze);
    n.inv_w[0] = (float) n.nx / (xe[n_xe - 1] - xe[0]);
    n.inv_w[1] = (float) n.ny / (ye[n_ye - 1] - ye[0]);
    n.inv_w[2] = (float) n.nz / (ze[n_ze - 1] - ze[0]);
    const float I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    std::memcpy(n.rot, rot ? rot : I, sizeof(I));
    for (int k = 0; k < 3; ++k) n.trans[k] = trans ? trans[k] : 0.f;
    n.identity = (std::memcmp(n.rot, I, sizeof(I)) == 0 && n.trans[0] == 0.f && n.trans[1] == 0.f && n.trans[2] == 0.f) ? 1 : 0;
    const size_t          nv = (size_t) n.nx * n.ny * n.nz;
    std::vector<uint16_t> mat;
    std::vector<MatEntry> lut;
    rc = build_density_lut(rho, nv, h->variant, mat, lut);
    if (rc) return rc;
    CU(cudaMalloc(&n.d_mat, nv * sizeof(uint16_t)));
    CU(cudaMalloc(&n.d_lut, lut.size() * sizeof(MatEntry)));
    CU(cudaMemcpyAsync(n.d_mat, mat.data(), nv * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(n.d_lut, lut.data(), lut.size() * sizeof(MatEntry), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    n.lut_size = (int) lut.size();
    h->beamline.push_back(n);
    h->nodes_dirty = true;
    return (int) h->beamline.size() - 1;
}

int
mqi_clear_beamline(mqi_handle* h) {
    int rc = activate(h);
    if (rc) return rc;
    rc = collect_run_fwd(h);
    if (rc) return rc;
    free_beamline(h);
    return MQI_OK;
}

int
mqi_clear_scorers(mqi_handle* h) {
    int rc = activate(h);
    if (rc) return rc;
    for (auto& s : h->scorers) {
        if (s.d_table) CU(launch_dij_clear(s.d_table, s.capacity, h->stream));
        if (s.d_dense) CU(cudaMemsetAsync(s.d_dense, 0, nvox(h) * sizeof(double), h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_set_accumulation(mqi_handle* h, int mode) {
    if (!h) return fail(MQI_EINVAL, "null handle");
    if (mode != MQI_ACCUM_ATOMIC && mode != MQI_ACCUM_WARP_MATCH) return fail(MQI_EINVAL, "unknown accumulation mode");
    h->accum = mode;
    return MQI_OK;
}

int
mqi_set_option(mqi_handle* h, const char* key, int64_t value) {
    if (!h || !key) return fail(MQI_EINVAL, "null argument");
    const std::string k(key);
    if (k == "count_steps") h->count_steps = value != 0;
    else if (k == "blocks_per_sm") h->blocks_per_sm_override = (int) value;
    else if (k == "fetch_order") h->fetch_order = (int) value;
    else if (k == "l2_persist") h->l2_persist = (int) value;   // 0 off, 1 material volume, 2 the first dense scorer's grid
    else if (k == "dij_write_combine") h->dij_write_combine = value != 0;
    else if (k == "rsp_exact") h->rsp_exact = value != 0;
    else return fail(MQI_EINVAL, "unknown option " + k);
    return MQI_OK;
}

int
mqi_set_beamlets(mqi_handle* h, const mqi_beamlet* beamlets, uint32_t n_spots, const uint64_t* histories_per_spot) {
    int rc = activate(h);
    if (rc) return rc;
    if (!beamlets || !histories_per_spot || n_spots == 0) return fail(MQI_EINVAL, "empty beam source");
    static_assert(sizeof(mqi_beamlet) == sizeof(BeamletDev), "beamlet layout");
    std::vector<unsigned long long> cum(n_spots);
    unsigned long long              acc = 0;
    for (uint32_t i = 0; i < n_spots; ++i) {
        acc += histories_per_spot[i];
        cum[i] = acc;
    }
    // the buffers of the previous source go back to the handle's pool and come out of it again when the size is the same
    // (a new batch of the same plan): cudaFree would wait for every kernel on the device, another handle's included
    pool_free(h, h->d_beamlets, h->n_spots * sizeof(BeamletDev));
    pool_free(h, h->d_cum, h->n_spots * sizeof(unsigned long long));
    h->d_beamlets = nullptr; h->d_cum = nullptr; h->n_spots = 0;
    if (h->d_vertices) {
        cudaFree(h->d_vertices); cudaFree(h->d_spot_ids);
        h->d_vertices = nullptr; h->d_spot_ids = nullptr; h->n_vertices = 0;
    }
    CU(pool_alloc(h, &h->d_beamlets, n_spots * sizeof(BeamletDev)));
    CU(pool_alloc(h, &h->d_cum, n_spots * sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(h->d_beamlets, beamlets, n_spots * sizeof(BeamletDev), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_cum, cum.data(), n_spots * sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->n_spots         = n_spots;
    h->total_histories = acc;
    return MQI_OK;
}

int
mqi_set_vertices(mqi_handle* h, const mqi_vertex* vertices, uint64_t n, const uint32_t* spot_ids) {
    int rc = activate(h);
    if (rc) return rc;
    if (!vertices || n == 0) return fail(MQI_EINVAL, "empty vertex list");
    static_assert(sizeof(mqi_vertex) == sizeof(VertexDev), "vertex layout");
    cudaFree(h->d_vertices); cudaFree(h->d_spot_ids);
    h->d_vertices = nullptr; h->d_spot_ids = nullptr;
    CU(cudaMalloc(&h->d_vertices, n * sizeof(VertexDev)));
    CU(cudaMemcpyAsync(h->d_vertices, vertices, n * sizeof(VertexDev), cudaMemcpyHostToDevice, h->stream));
    if (spot_ids) {
        CU(cudaMalloc(&h->d_spot_ids, n * sizeof(uint32_t)));
        CU(cudaMemcpyAsync(h->d_spot_ids, spot_ids, n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    h->n_vertices = n;
    return MQI_OK;
}

static int
collect_run(mqi_handle* h) {
    if (!h->pending) return MQI_OK;
    CU(cudaStreamSynchronize(h->stream));
    unsigned long long c[C_COUNT];
    CU(cudaMemcpy(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.histories       = c[C_DONE];
    h->stats.steps           = c[C_STEPS];
    h->stats.secondaries     = c[C_SECONDARIES];
    h->stats.stack_overflows = c[C_OVERFLOW];
    h->stats.dij_table_full  = c[C_DIJ_FULL];
    h->stats.kernel_ms       = ms;
    h->stats.launches        = 1;
    h->pending               = false;
    return MQI_OK;
}

int
mqi_run_async(mqi_handle* h, uint64_t seed, uint64_t first_history, uint64_t count, int per_spot) {
    return mqi_run_async_sharded(h, seed, first_history, count, per_spot, 1, 0);
}

int
mqi_run_async_sharded(mqi_handle* h, uint64_t seed, uint64_t first_history, uint64_t count, int per_spot, uint32_t n_shards,
                      uint32_t shard) {
    int rc = activate(h);
    if (rc) return rc;
    if (n_shards == 0 || shard >= n_shards) return fail(MQI_EINVAL, "shard index out of range");
    if (!h->has_grid) return fail(MQI_ESTATE, "no grid set");
    if (h->scorers.empty()) return fail(MQI_ESTATE, "no scorer added");
    if (h->d_vertices) {
        if (first_history + count > h->n_vertices) return fail(MQI_EINVAL, "history range exceeds the vertex list");
    } else {
        if (!h->d_beamlets) return fail(MQI_ESTATE, "no beam source set");
        if (first_history + count > h->total_histories) return fail(MQI_EINVAL, "history range exceeds the beam source");
    }
    // a Dij scorer without spot ids runs in the reference's dense mode (slot = voxel, mqi_transport.hpp:78-81): the
    // table must then cover the grid
    if (!per_spot || (h->d_vertices && !h->d_spot_ids))
        for (const auto& s : h->scorers)
            if (s.kind == MQI_SCORER_DIJ && s.capacity < nvox(h))
                return fail(MQI_EINVAL, "a Dij scorer run without per-spot keys uses slot = voxel: capacity must be >= the number of voxels");
    rc = collect_run(h);   // counters of a previous asynchronous launch are overwritten below
    if (rc) return rc;
    rc = ensure_scorer_buffers(h);
    if (rc) return rc;
    rc = sync_nodes(h);
    if (rc) return rc;
    std::memset(&h->stats, 0, sizeof(h->stats));
    if (count == 0) return MQI_OK;
    Params p;
    fill_params(h, p);
    p.seed     = seed;
    for (int i = 0; i < 10; ++i) {   // Philox4x32 key schedule
        p.rk[2 * i]     = (uint32_t) seed + (uint32_t) i * 0x9E3779B9u;
        p.rk[2 * i + 1] = (uint32_t) (seed >> 32) + (uint32_t) i * 0xBB67AE85u;
    }
    p.first    = first_history;
    p.count    = count;
    p.n_shards = n_shards;
    p.shard    = shard;
    p.per_spot = per_spot;
    if (h->d_vertices) {   // explicit vertices are addressed relative to the launch
        p.src.vertices = h->d_vertices + first_history;
        p.src.spot_ids = h->d_spot_ids ? h->d_spot_ids + first_history : nullptr;
    }
    const size_t smem = transport_smem_bytes(p.n_edge_floats, p.n_nodes);
    if (smem > 200 * 1024) return fail(MQI_EINVAL, "the world's grid edges do not fit into shared memory");
    int          bps  = 0;
    CU(transport_occupancy(p, h->variant, smem, &bps));
    if (bps < 1) return fail(MQI_ECUDA, "transport kernel does not fit on an SM");
    if (h->blocks_per_sm_override > 0) bps = std::min(bps, h->blocks_per_sm_override);
    // persistent grid: a whole number of CTAs per SM, never more lanes than histories
    const unsigned long long blk = (unsigned long long) transport_block(p);
    unsigned long long want = (count / n_shards + 32 + blk - 1) / blk;   // this shard's share of the range
    int                grid = (int) std::min<unsigned long long>((unsigned long long) h->sm_count * bps, want);
    {   // subsystem 3: the 16-bit material volume of the scored grid is read once per voxel step by every
        // lane; keep it resident in L2 (persisting access window on the launching stream) while the fp64
        // dose grid streams through the rest.  Larger volumes than the persisting carve-out get a
        // proportional hit ratio.
        // (the stream attribute is only touched when the option is on: the caller's own access-policy window on a
        // stream bound with mqi_set_stream stays as it is otherwise)
        cudaStreamAttrValue attr;
        std::memset(&attr, 0, sizeof(attr));
        bool set_window = false;
        if (h->l2_persist && h->l2_persist_max > 0 && h->l2_window_max > 0) {
            const void* base  = h->d_mat;
            size_t      bytes = nvox(h) * sizeof(uint16_t);
            if (h->l2_persist == 2) {   // the dose grid instead: its RED sectors are what DRAM sees written back 204 times over
                for (const auto& sc : h->scorers)
                    if (sc.kind != MQI_SCORER_DIJ && sc.d_dense) { base = sc.d_dense; bytes = nvox(h) * sizeof(double); break; }
            }
            bytes = std::min(bytes, h->l2_window_max);
            const size_t carve  = std::min(h->l2_persist_max, bytes);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                attr.accessPolicyWindow.base_ptr  = const_cast<void*>(base);
                attr.accessPolicyWindow.num_bytes = bytes;
                attr.accessPolicyWindow.hitRatio  = (float) std::min(1.0, (double) carve / (double) bytes);
                attr.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
                set_window = true;
            } else {
                cudaGetLastError();
            }
        }
        if (set_window || h->l2_window_set) {   // switching the option off clears the window this handle set before
            if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
            h->l2_window_set = set_window;
        }
    }
    {
        const size_t hb = transport_handover_bytes(grid, p.n_nodes);
        if (hb > h->adv_raw_bytes) {
            CU(cudaStreamSynchronize(h->stream));   // an earlier launch may still use the smaller buffer
            cudaFree(h->d_adv_raw);
            h->d_adv_raw = nullptr; h->adv_raw_bytes = 0;
            CU(cudaMalloc(&h->d_adv_raw, hb));
            h->adv_raw_bytes = hb;
        }
        p.adv_raw = h->d_adv_raw;
    }
    CU(cudaMemsetAsync(h->d_counters, 0, C_COUNT * sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(launch_transport(p, h->variant, grid, smem, h->stream));
    CU(cudaEventRecord(h->ev1, h->stream));
    h->pending = true;
    return MQI_OK;
}

int
mqi_run(mqi_handle* h, uint64_t seed, uint64_t first_history, uint64_t count, int per_spot) {
    int rc = mqi_run_async(h, seed, first_history, count, per_spot);
    if (rc) return rc;
    return collect_run(h);
}

int
mqi_get_run_stats(mqi_handle* h, mqi_run_stats* out) {
    if (!h || !out) return fail(MQI_EINVAL, "null argument");
    int rc = activate(h);
    if (rc) return rc;
    rc = collect_run(h);
    if (rc) return rc;
    *out = h->stats;
    return MQI_OK;
}

int
mqi_set_stream(mqi_handle* h, void* cuda_stream) {
    int rc = activate(h);
    if (rc) return rc;
    rc = collect_run(h);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return MQI_OK;
}

int
mqi_get_dense(mqi_handle* h, int scorer, double* out, double scale) {
    int rc = activate(h);
    if (rc) return rc;
    if (scorer < 0 || scorer >= (int) h->scorers.size() || !out) return fail(MQI_EINVAL, "bad scorer index");
    HostScorer& s = h->scorers[scorer];
    if (s.kind == MQI_SCORER_DIJ) return fail(MQI_EINVAL, "use mqi_get_sparse for Dij scorers");
    const size_t nv = nvox(h);
    if (!s.d_dense) {
        std::memset(out, 0, nv * sizeof(double));
        return MQI_OK;
    }
    CU(cudaMemcpyAsync(out, s.d_dense, nv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (scale != 1.0)
        for (size_t i = 0; i < nv; ++i) out[i] *= scale;
    return MQI_OK;
}

int
mqi_get_scorer_device_ptr(mqi_handle* h, int scorer, void** d_ptr, uint64_t* n_elements) {
    int rc = activate(h);
    if (rc) return rc;
    if (scorer < 0 || scorer >= (int) h->scorers.size() || !d_ptr) return fail(MQI_EINVAL, "bad scorer index");
    if (!h->has_grid) return fail(MQI_ESTATE, "no grid set");
    rc = ensure_scorer_buffers(h);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    HostScorer& s = h->scorers[scorer];
    if (s.kind == MQI_SCORER_DIJ) {
        *d_ptr = s.d_table;
        if (n_elements) *n_elements = s.capacity;
    } else {
        *d_ptr = s.d_dense;
        if (n_elements) *n_elements = nvox(h);
    }
    return MQI_OK;
}

static const size_t kDijChunk = 1u << 16;

int
mqi_get_sparse_count(mqi_handle* h, int scorer, uint64_t* nnz) {
    int rc = activate(h);
    if (rc) return rc;
    if (scorer < 0 || scorer >= (int) h->scorers.size() || !nnz) return fail(MQI_EINVAL, "bad scorer index");
    HostScorer& s = h->scorers[scorer];
    if (s.kind != MQI_SCORER_DIJ) return fail(MQI_EINVAL, "not a Dij scorer");
    *nnz = 0;
    if (!s.d_table) return MQI_OK;
    DevBuf<unsigned long long> cnt;
    CU(cnt.alloc(1));
    CU(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), h->stream));
    CU(launch_dij_count(s.d_table, s.capacity, cnt.p, h->stream));
    unsigned long long c = 0;
    CU(cudaMemcpyAsync(&c, cnt.p, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *nnz = c;
    return MQI_OK;
}

int
mqi_get_sparse(mqi_handle* h, int scorer, uint32_t* key1, uint32_t* key2, double* value, uint64_t nnz, double scale) {
    int rc = activate(h);
    if (rc) return rc;
    if (scorer < 0 || scorer >= (int) h->scorers.size()) return fail(MQI_EINVAL, "bad scorer index");
    HostScorer& s = h->scorers[scorer];
    if (s.kind != MQI_SCORER_DIJ) return fail(MQI_EINVAL, "not a Dij scorer");
    if (nnz == 0 || !s.d_table) return MQI_OK;
    if (!key1 || !key2 || !value) return fail(MQI_EINVAL, "null output");
    const size_t nchunks = (s.capacity + kDijChunk - 1) / kDijChunk;
    DevBuf<unsigned long long> d_cnt;
    CU(d_cnt.alloc(nchunks));
    CU(cudaMemsetAsync(d_cnt.p, 0, nchunks * sizeof(unsigned long long), h->stream));
    CU(launch_dij_chunk_count(s.d_table, s.capacity, kDijChunk, d_cnt.p, h->stream));
    std::vector<unsigned long long> cnt(nchunks);
    CU(cudaMemcpyAsync(cnt.data(), d_cnt.p, nchunks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    unsigned long long acc = 0;
    for (size_t i = 0; i < nchunks; ++i) {
        const unsigned long long c = cnt[i];
        cnt[i] = acc;
        acc += c;
    }
    if (acc != nnz) return fail(MQI_EINVAL, "nnz does not match the table (call mqi_get_sparse_count first)");
    CU(cudaMemcpyAsync(d_cnt.p, cnt.data(), nchunks * sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
    DevBuf<uint32_t> d_k1, d_k2;
    DevBuf<double>   d_v;
    CU(d_k1.alloc(nnz)); CU(d_k2.alloc(nnz)); CU(d_v.alloc(nnz));
    CU(launch_dij_chunk_write(s.d_table, s.capacity, kDijChunk, d_cnt.p, d_k1.p, d_k2.p, d_v.p, scale, h->stream));
    CU(cudaMemcpyAsync(key1, d_k1.p, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(key2, d_k2.p, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(value, d_v.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_stat_partial(mqi_handle* h, int scorer_sum, int scorer_sumsq, uint64_t n_histories, double threshold_fraction,
                 double max_mean, double out[3]) {
    int rc = activate(h);
    if (rc) return rc;
    const int ns = (int) h->scorers.size();
    if (scorer_sum < 0 || scorer_sum >= ns || scorer_sumsq < 0 || scorer_sumsq >= ns || !out) return fail(MQI_EINVAL, "bad scorer index");
    const double* sum = h->scorers[scorer_sum].d_dense;
    const double* sq  = h->scorers[scorer_sumsq].d_dense;
    if (!sum || !sq || n_histories < 2) return fail(MQI_ESTATE, "stat scorers are empty");
    if (!h->d_stat3) CU(cudaMalloc(&h->d_stat3, 3 * sizeof(double)));
    struct { double* p; } d { h->d_stat3 };
    CU(cudaMemsetAsync(d.p, 0, 3 * sizeof(double), h->stream));
    if (max_mean < 0.0) {
        CU(launch_stat_max(sum, nvox(h), 1.0 / (double) n_histories, d.p + 2, h->stream));
        CU(cudaMemcpyAsync(&max_mean, d.p + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(launch_stat_partial(sum, sq, nvox(h), (double) n_histories, threshold_fraction * max_mean, d.p, h->stream));
    CU(cudaMemcpyAsync(out, d.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    out[2] = max_mean;
    return MQI_OK;
}

int
mqi_stat_partial_buffers(mqi_handle* h, const void* d_sum, const void* d_sumsq, uint64_t n_voxels, uint64_t n_histories,
                         double threshold_fraction, double max_mean, double out[3]) {
    int rc = activate(h);
    if (rc) return rc;
    if (!d_sum || !d_sumsq || !out || n_voxels == 0) return fail(MQI_EINVAL, "null buffer");
    if (n_histories < 2) return fail(MQI_ESTATE, "stat buffers need at least two histories");
    const double* sum = static_cast<const double*>(d_sum);
    const double* sq  = static_cast<const double*>(d_sumsq);
    if (!h->d_stat3) CU(cudaMalloc(&h->d_stat3, 3 * sizeof(double)));
    struct { double* p; } d { h->d_stat3 };
    CU(cudaMemsetAsync(d.p, 0, 3 * sizeof(double), h->stream));
    if (max_mean < 0.0) {
        CU(launch_stat_max(sum, n_voxels, 1.0 / (double) n_histories, d.p + 2, h->stream));
        CU(cudaMemcpyAsync(&max_mean, d.p + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(launch_stat_partial(sum, sq, n_voxels, (double) n_histories, threshold_fraction * max_mean, d.p, h->stream));
    CU(cudaMemcpyAsync(out, d.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    out[2] = max_mean;
    return MQI_OK;
}

int
mqi_scale_scorer(mqi_handle* h, int scorer, double factor) {
    int rc = activate(h);
    if (rc) return rc;
    if (scorer < 0 || scorer >= (int) h->scorers.size()) return fail(MQI_EINVAL, "bad scorer index");
    HostScorer& s = h->scorers[scorer];
    if (s.kind == MQI_SCORER_DIJ || !s.d_dense) return fail(MQI_ESTATE, "scorer has no dense buffer");
    CU(launch_scale(s.d_dense, nvox(h), factor, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

// ---- multi-GPU reduction over NCCL (loaded at run time) ----------------------------------------
namespace
{
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::vector<int>        g_comm_devices;
std::vector<ncclComm_t> g_comms;
const int kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h (2.x)

int
load_nccl() {
    if (g_nccl.ok) return MQI_OK;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail(MQI_ESTATE, "NCCL is not available (dlopen libnccl.so.2 failed)");
#define MQI_SYM(field, name)                                                                           \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(g_nccl.lib, name);                                \
    if (!g_nccl.field) return fail(MQI_ESTATE, std::string("NCCL symbol missing: ") + name)
    MQI_SYM(CommInitAll, "ncclCommInitAll");
    MQI_SYM(CommDestroy, "ncclCommDestroy");
    MQI_SYM(Reduce, "ncclReduce");
    MQI_SYM(AllReduce, "ncclAllReduce");
    MQI_SYM(ReduceScatter, "ncclReduceScatter");
    MQI_SYM(GroupStart, "ncclGroupStart");
    MQI_SYM(GroupEnd, "ncclGroupEnd");
    MQI_SYM(GetErrorString, "ncclGetErrorString");
#undef MQI_SYM
    g_nccl.ok = true;
    return MQI_OK;
}

#define NC(call)                                                                                       \
    do {                                                                                               \
        int r__ = (call);                                                                              \
        if (r__ != 0) return fail(MQI_ECUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r__));   \
    } while (0)

// the handles of one multi-GPU group: distinct devices, the same grid, dense scorers; communicators are created
// on first use and cached for the device list
int
prepare_group(mqi_handle* const* handles, int n, const int* scorer_ids, int n_scorers, size_t* count_out) {
    if (!handles || n < 1) return fail(MQI_EINVAL, "no handles");
    std::vector<int> devs(n);
    size_t           count = 0;
    for (int i = 0; i < n; ++i) {
        mqi_handle* h = handles[i];
        if (!h) return fail(MQI_EINVAL, "null handle");
        for (int k = 0; k < n_scorers; ++k) {
            const int scorer = scorer_ids[k];
            if (scorer < 0 || scorer >= (int) h->scorers.size()) return fail(MQI_EINVAL, "bad scorer index");
            if (h->scorers[scorer].kind == MQI_SCORER_DIJ) return fail(MQI_EINVAL, "Dij tables are sharded by spot, not reduced");
        }
        if (!h->has_grid) return fail(MQI_ESTATE, "no grid set");
        if (i == 0) count = nvox(h);
        else if (nvox(h) != count) return fail(MQI_EINVAL, "handles have different grids");
        devs[i] = h->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(MQI_EINVAL, "two handles on the same device");
        int rc = activate(h);
        if (rc) return rc;
        rc = collect_run(h);
        if (rc) return rc;
        rc = ensure_scorer_buffers(h);
        if (rc) return rc;
    }
    *count_out = count;
    if (n == 1) return MQI_OK;
    int rc = load_nccl();
    if (rc) return rc;
    if (devs != g_comm_devices) {
        for (auto c : g_comms) g_nccl.CommDestroy(c);
        g_comms.assign(n, nullptr);
        NC(g_nccl.CommInitAll(g_comms.data(), n, devs.data()));
        g_comm_devices = devs;
    }
    return MQI_OK;
}

int
reduce_impl(mqi_handle* const* handles, int n, int scorer, int root, bool all) {
    if (root < 0 || root >= std::max(n, 1)) return fail(MQI_EINVAL, "root out of range");
    size_t count = 0;
    int    rc    = prepare_group(handles, n, &scorer, 1, &count);
    if (rc || n == 1) return rc;
    NC(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        mqi_handle* h = handles[i];
        CU(cudaSetDevice(h->device));
        double* buf = h->scorers[scorer].d_dense;
        if (all) NC(g_nccl.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, g_comms[i], h->stream));
        else NC(g_nccl.Reduce(buf, buf, count, kNcclFloat64, kNcclSum, root, g_comms[i], h->stream));
    }
    NC(g_nccl.GroupEnd());
    for (int i = 0; i < n; ++i) {
        CU(cudaSetDevice(handles[i]->device));
        CU(cudaStreamSynchronize(handles[i]->stream));
    }
    return MQI_OK;
}

// The stopping criterion over several devices without moving grids to one of them.  Every device keeps its own running
// sums.  calculate_stat averages sigma/mu over the voxels whose mean dose exceeds threshold x the largest mean dose, so
// only voxels that CAN exceed it are exchanged: with M_lb = the largest local sum of any device (a lower bound of the
// largest summed value), a voxel whose local sum is at most threshold * M_lb / n on every device sums to at most
// threshold * M_lb and cannot qualify.  The grids are cut into chunks of 4 096 voxels; a chunk is kept if any device holds
// a larger value in it; the kept chunks of both grids are packed, one ncclReduceScatter per grid leaves each device with
// the summed values of 1/n of them (every link carries 1/n of the packed values, all links at once), each device
// evaluates its slice, the host combines n x 3 doubles.  The selected voxels and the result are exactly those of an
// evaluation on whole summed grids (tests/test_gpu_multi.py); at config C3 a few per cent of the grid travel.
int
stat_multi_impl(mqi_handle* const* handles, int n, int s_sum, int s_sq, uint64_t n_histories, double threshold_fraction,
                double out[3]) {
    if (!out) return fail(MQI_EINVAL, "null argument");
    if (n_histories < 2) return fail(MQI_ESTATE, "the stopping criterion needs at least two histories");
    const int ids[2] = { s_sum, s_sq };
    size_t    count  = 0;
    int       rc     = prepare_group(handles, n, ids, 2, &count);
    if (rc) return rc;
    if (n == 1) return mqi_stat_partial(handles[0], s_sum, s_sq, n_histories, threshold_fraction, -1.0, out);
    const size_t chunk = 4096, nchunks = (count + chunk - 1) / chunk;
    // 1. the largest local sum of any device
    std::vector<double> local_max(n, 0.0);
    for (int i = 0; i < n; ++i) {
        mqi_handle* h = handles[i];
        CU(cudaSetDevice(h->device));
        if (!h->d_stat3) CU(cudaMalloc(&h->d_stat3, 3 * sizeof(double)));
        CU(cudaMemsetAsync(h->d_stat3, 0, 3 * sizeof(double), h->stream));
        CU(launch_stat_max(h->scorers[s_sum].d_dense, count, 1.0, h->d_stat3 + 2, h->stream));
        CU(cudaMemcpyAsync(&local_max[i], h->d_stat3 + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    double m_lb = 0.0;
    for (int i = 0; i < n; ++i) {
        CU(cudaSetDevice(handles[i]->device));
        CU(cudaStreamSynchronize(handles[i]->stream));
        m_lb = std::max(m_lb, local_max[i]);
    }
    if (!(m_lb > 0.0)) return fail(MQI_ESTATE, "stat scorers are empty");
    // 2. chunks in which some device holds a value above threshold * M_lb / n
    const double bound = threshold_fraction * m_lb / (double) n;
    std::vector<std::vector<unsigned char>> flags(n, std::vector<unsigned char>(nchunks));
    for (int i = 0; i < n; ++i) {
        mqi_handle* h = handles[i];
        CU(cudaSetDevice(h->device));
        if (h->stat_flags_n < nchunks) {
            cudaFree(h->d_stat_flags);
            cudaFree(h->d_stat_list);
            h->d_stat_flags = nullptr;
            h->d_stat_list  = nullptr;
            h->stat_flags_n = 0;
            CU(cudaMalloc(&h->d_stat_flags, nchunks));
            CU(cudaMalloc(&h->d_stat_list, nchunks * sizeof(unsigned int)));
            h->stat_flags_n = nchunks;
        }
        CU(launch_chunk_above(h->scorers[s_sum].d_dense, count, chunk, bound, h->d_stat_flags, h->stream));
        CU(cudaMemcpyAsync(flags[i].data(), h->d_stat_flags, nchunks, cudaMemcpyDeviceToHost, h->stream));
    }
    std::vector<unsigned int> list;
    for (int i = 0; i < n; ++i) {
        CU(cudaSetDevice(handles[i]->device));
        CU(cudaStreamSynchronize(handles[i]->stream));
    }
    for (size_t c = 0; c < nchunks; ++c) {
        unsigned char any = 0;
        for (int i = 0; i < n; ++i) any |= flags[i][c];
        if (any) list.push_back((unsigned int) c);
    }
    // 3. pack the kept chunks of both grids (padded with zeros to a multiple of n) and reduce-scatter them
    const size_t packed = ((list.size() * chunk + (size_t) n - 1) / (size_t) n) * (size_t) n, per = packed / (size_t) n;
    if (per == 0) return fail(MQI_ESTATE, "stat scorers are empty");
    for (int i = 0; i < n; ++i) {
        mqi_handle* h = handles[i];
        CU(cudaSetDevice(h->device));
        if (h->stat_slice_n < packed) {   // [packed sum | packed sq | slice sum | slice sq]
            cudaFree(h->d_stat_slice);
            h->d_stat_slice = nullptr;
            h->stat_slice_n = 0;
            const size_t cap = packed + packed / 4 + chunk * (size_t) n;   // head room: the kept set grows slowly from pass to pass
            CU(cudaMalloc(&h->d_stat_slice, (2 * cap + 2 * (cap / (size_t) n + 1)) * sizeof(double)));
            h->stat_slice_n = cap;
        }
        double* pa = h->d_stat_slice;
        double* pb = pa + h->stat_slice_n;
        CU(cudaMemcpyAsync(h->d_stat_list, list.data(), list.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, h->stream));
        CU(launch_pack_chunks(h->scorers[s_sum].d_dense, h->scorers[s_sq].d_dense, count, chunk, h->d_stat_list, list.size(), pa, pb, h->stream));
        if (packed > list.size() * chunk) {
            CU(cudaMemsetAsync(pa + list.size() * chunk, 0, (packed - list.size() * chunk) * sizeof(double), h->stream));
            CU(cudaMemsetAsync(pb + list.size() * chunk, 0, (packed - list.size() * chunk) * sizeof(double), h->stream));
        }
    }
    NC(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        mqi_handle* h = handles[i];
        CU(cudaSetDevice(h->device));
        double* pa = h->d_stat_slice;
        double* pb = pa + h->stat_slice_n;
        double* sa = pb + h->stat_slice_n;
        double* sb = sa + (h->stat_slice_n / (size_t) n + 1);
        NC(g_nccl.ReduceScatter(pa, sa, per, kNcclFloat64, kNcclSum, g_comms[i], h->stream));
        NC(g_nccl.ReduceScatter(pb, sb, per, kNcclFloat64, kNcclSum, g_comms[i], h->stream));
    }
    NC(g_nccl.GroupEnd());
    // 4. pass 1: the largest mean dose; pass 2: sum of sigma / mu and the number of voxels above the threshold
    std::vector<double> host(3 * (size_t) n, 0.0);
    double              max_mean = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < n; ++i) {
            mqi_handle* h = handles[i];
            CU(cudaSetDevice(h->device));
            double* sa = h->d_stat_slice + 2 * h->stat_slice_n;
            double* sb = sa + (h->stat_slice_n / (size_t) n + 1);
            double* d3 = h->d_stat3;
            if (pass == 0) {
                CU(cudaMemsetAsync(d3, 0, 3 * sizeof(double), h->stream));
                CU(launch_stat_max(sa, per, 1.0 / (double) n_histories, d3 + 2, h->stream));
                CU(cudaMemcpyAsync(&host[3 * i + 2], d3 + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            } else {
                CU(launch_stat_partial(sa, sb, per, (double) n_histories, threshold_fraction * max_mean, d3, h->stream));
                CU(cudaMemcpyAsync(&host[3 * i], d3, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            }
        }
        for (int i = 0; i < n; ++i) {
            CU(cudaSetDevice(handles[i]->device));
            CU(cudaStreamSynchronize(handles[i]->stream));
        }
        if (pass == 0)
            for (int i = 0; i < n; ++i) max_mean = std::max(max_mean, host[3 * i + 2]);
    }
    out[0] = out[1] = 0.0;
    for (int i = 0; i < n; ++i) {
        out[0] += host[3 * i];
        out[1] += host[3 * i + 1];
    }
    out[2] = max_mean;
    return MQI_OK;
}
}   // namespace

extern "C" {
int
mqi_reduce_dense(mqi_handle* const* handles, int n, int scorer, int root) {
    return reduce_impl(handles, n, scorer, root, false);
}
int
mqi_allreduce_dense(mqi_handle* const* handles, int n, int scorer) {
    return reduce_impl(handles, n, scorer, 0, true);
}
int
mqi_stat_multi(mqi_handle* const* handles, int n, int scorer_sum, int scorer_sumsq, uint64_t n_histories,
               double threshold_fraction, double out[3]) {
    return stat_multi_impl(handles, n, scorer_sum, scorer_sumsq, n_histories, threshold_fraction, out);
}
}

// ---- deterministic device pieces ---------------------------------------------------------------
int
mqi_dev_hu_to_density(mqi_handle* h, const int16_t* hu, uint64_t n, float density_scale, float* rho_out) {
    int rc = activate(h);
    if (rc) return rc;
    if (!hu || !rho_out) return fail(MQI_EINVAL, "null argument");
    if (n == 0) return MQI_OK;
    DevBuf<int16_t> d_hu;
    DevBuf<float>   d_rho;
    CU(d_hu.alloc(n)); CU(d_rho.alloc(n));
    CU(cudaMemcpyAsync(d_hu.p, hu, n * sizeof(int16_t), cudaMemcpyHostToDevice, h->stream));
    CU(launch_hu_to_density(d_hu.p, d_rho.p, n, h->d_correction, density_scale, h->stream));
    CU(cudaMemcpyAsync(rho_out, d_rho.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_dev_rsp(mqi_handle* h, const float* rho, const float* ek, uint64_t n, float* rsp_out, float* rl_out) {
    int rc = activate(h);
    if (rc) return rc;
    if (!rho || !ek || !rsp_out || !rl_out) return fail(MQI_EINVAL, "null argument");
    if (n == 0) return MQI_OK;
    std::vector<MatEntry> m(n);
    for (uint64_t i = 0; i < n; ++i) m[i] = make_mat_entry(rho[i], h->variant);
    DevBuf<MatEntry> d_m;
    DevBuf<float>    d_ek, d_rsp, d_rl;
    CU(d_m.alloc(n)); CU(d_ek.alloc(n)); CU(d_rsp.alloc(n)); CU(d_rl.alloc(n));
    CU(cudaMemcpyAsync(d_m.p, m.data(), n * sizeof(MatEntry), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_ek.p, ek, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(launch_dev_rsp(d_m.p, d_ek.p, n, d_rsp.p, d_rl.p, h->stream, h->rsp_exact != 0));
    CU(cudaMemcpyAsync(rsp_out, d_rsp.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(rl_out, d_rl.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_dev_grid_step(mqi_handle* h, const float* p, const float* d, uint64_t n, int32_t* cell, uint64_t* cnb, float* dist,
                  float* dir_after, float* p_exit, int32_t* cell_after) {
    int rc = activate(h);
    if (rc) return rc;
    if (!h->has_grid) return fail(MQI_ESTATE, "no grid set");
    if (!p || !d || !cell || !cnb || !dist || !dir_after || !p_exit || !cell_after) return fail(MQI_EINVAL, "null argument");
    if (n == 0) return MQI_OK;
    DevBuf<float>              dp, dd, ddist, ddir, dpe;
    DevBuf<int32_t>            dc, dca;
    DevBuf<unsigned long long> dcnb;
    CU(dp.alloc(3 * n)); CU(dd.alloc(3 * n)); CU(ddist.alloc(n)); CU(ddir.alloc(3 * n)); CU(dpe.alloc(3 * n));
    CU(dc.alloc(3 * n)); CU(dca.alloc(3 * n)); CU(dcnb.alloc(n));
    CU(cudaMemcpyAsync(dp.p, p, 3 * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dd.p, d, 3 * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    Params prm;
    fill_params(h, prm);
    CU(launch_dev_grid_step(prm, dp.p, dd.p, n, dc.p, dcnb.p, ddist.p, ddir.p, dpe.p, dca.p, h->stream));
    CU(cudaMemcpyAsync(cell, dc.p, 3 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(cnb, dcnb.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(dist, ddist.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(dir_after, ddir.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(p_exit, dpe.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(cell_after, dca.p, 3 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_dev_grid_entry(mqi_handle* h, const float* p, const float* d, uint64_t n, float* dist, int32_t* cell) {
    int rc = activate(h);
    if (rc) return rc;
    if (!h->has_grid) return fail(MQI_ESTATE, "no grid set");
    if (!p || !d || !dist || !cell) return fail(MQI_EINVAL, "null argument");
    if (n == 0) return MQI_OK;
    DevBuf<float>   dp, dd, ddist;
    DevBuf<int32_t> dc;
    CU(dp.alloc(3 * n)); CU(dd.alloc(3 * n)); CU(ddist.alloc(n)); CU(dc.alloc(3 * n));
    CU(cudaMemcpyAsync(dp.p, p, 3 * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dd.p, d, 3 * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    Params prm;
    fill_params(h, prm);
    CU(launch_dev_grid_entry(prm, dp.p, dd.p, n, ddist.p, dc.p, h->stream));
    CU(cudaMemcpyAsync(dist, ddist.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(cell, dc.p, 3 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_dev_hash(mqi_handle* h, const uint32_t* k1, const uint32_t* k2, const uint64_t* capacity, uint64_t n, uint32_t* out) {
    int rc = activate(h);
    if (rc) return rc;
    if (!k1 || !k2 || !capacity || !out) return fail(MQI_EINVAL, "null argument");
    if (n == 0) return MQI_OK;
    DevBuf<uint32_t>           d1, d2, dout;
    DevBuf<unsigned long long> dc;
    CU(d1.alloc(n)); CU(d2.alloc(n)); CU(dout.alloc(n)); CU(dc.alloc(n));
    CU(cudaMemcpyAsync(d1.p, k1, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d2.p, k2, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dc.p, capacity, n * 8, cudaMemcpyHostToDevice, h->stream));
    CU(launch_dev_hash(d1.p, d2.p, dc.p, n, dout.p, h->stream));
    CU(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_dev_insert(mqi_handle* h, int scorer, const uint32_t* key1, const uint32_t* key2, const double* value, uint64_t n) {
    int rc = activate(h);
    if (rc) return rc;
    if (scorer < 0 || scorer >= (int) h->scorers.size()) return fail(MQI_EINVAL, "bad scorer index");
    if (!key1 || !key2 || !value) return fail(MQI_EINVAL, "null argument");
    if (!h->has_grid) return fail(MQI_ESTATE, "no grid set");
    if (n == 0) return MQI_OK;
    if (h->scorers[scorer].kind != MQI_SCORER_DIJ) {
        for (uint64_t i = 0; i < n; ++i)
            if (key1[i] >= nvox(h)) return fail(MQI_EINVAL, "voxel key out of range");
    } else {   // the reference's dense mode inside a Dij table: slot = key1, which must lie inside the table
        for (uint64_t i = 0; i < n; ++i)
            if (key2[i] == 0xffffffffu && key1[i] >= h->scorers[scorer].capacity)
                return fail(MQI_EINVAL, "dense-mode key (key2 = 0xffffffff) beyond the Dij table capacity");
    }
    rc = ensure_scorer_buffers(h);
    if (rc) return rc;
    DevBuf<uint32_t> d1, d2;
    DevBuf<double>   dv;
    CU(d1.alloc(n)); CU(d2.alloc(n)); CU(dv.alloc(n));
    CU(cudaMemcpyAsync(d1.p, key1, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d2.p, key2, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dv.p, value, n * 8, cudaMemcpyHostToDevice, h->stream));
    Params prm;
    fill_params(h, prm);
    CU(launch_dev_insert(prm, scorer, d1.p, d2.p, dv.p, n, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

int
mqi_dev_sample_vertices(mqi_handle* h, uint64_t seed, uint64_t first, uint64_t n, mqi_vertex* out, uint32_t* spot_out) {
    int rc = activate(h);
    if (rc) return rc;
    if (!h->d_beamlets) return fail(MQI_ESTATE, "no beam source set");
    if (!out || !spot_out) return fail(MQI_EINVAL, "null argument");
    if (first + n > h->total_histories) return fail(MQI_EINVAL, "history range exceeds the beam source");
    if (n == 0) return MQI_OK;
    DevBuf<VertexDev> dv;
    DevBuf<uint32_t>  ds;
    CU(dv.alloc(n)); CU(ds.alloc(n));
    Params prm;
    fill_params(h, prm);
    prm.seed = seed;
    CU(launch_dev_sample(prm, first, n, dv.p, ds.p, h->stream));
    CU(cudaMemcpyAsync(out, dv.p, n * sizeof(VertexDev), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(spot_out, ds.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MQI_OK;
}

}   // extern "C"
