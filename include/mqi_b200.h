/* mqi_b200.h -- C ABI of the B200-native moqui proton-transport hot path (libmqi_b200.so).
 *
 * The reference (MGHPhysicsResearch/moquimc) has no FFI layer: its "API" for this path is the C++
 * template kernel mc::transport_particles_patient<R> plus three global device pointers
 * (moqui/kernel_functions/mqi_transport.hpp:113-126, mqi_variables.hpp:24-26), driven by the
 * x_environment life cycle initialize() -> run() -> finalize() (moqui/base/environments/
 * mqi_xenvironment.hpp:89-146).  This header is the flat, POD-only replacement of that contract;
 * every entry point cites the reference interface it replaces.  All functions return 0 on success
 * or a negative MQI_E* code; mqi_last_error() returns the message of the calling thread's last
 * failure.  A handle is not thread-safe; calls are blocking unless stated.  No exceptions cross the
 * boundary and there is no CPU fallback: every compute entry point fails with MQI_ENODEVICE when no
 * CUDA device is usable.
 */
#ifndef MQI_B200_H
#define MQI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MQI_API __attribute__((visibility("default")))
#else
#define MQI_API
#endif

#define MQI_OK 0
#define MQI_EINVAL -1
#define MQI_ENODEVICE -2
#define MQI_ECUDA -3
#define MQI_ENOMEM -4
#define MQI_ESTATE -5

/* physics variant = the reference's compile-time macro __PHYSICS_DEBUG__:
 * phantom_env is built WITH it (tests/mc/phantom/CMakeLists.txt:10), tps_env without. */
#define MQI_PHYSICS_RELEASE 0
#define MQI_PHYSICS_DEBUG 1

/* scorer kinds: moqui/base/mqi_scorer.hpp:13-21 and scorers/mqi_scorer_energy_deposit.hpp */
#define MQI_SCORER_DOSE 0        /* dose_to_water            :43-60  (tps "Dose") */
#define MQI_SCORER_EDEP 1        /* energy_deposit           :14-19  ("EnergyDeposition") */
#define MQI_SCORER_LETD_NUMER 2  /* LETd_weight1             :93-114 */
#define MQI_SCORER_LETD_DENOM 3  /* LETd_weight2             :117-137 */
#define MQI_SCORER_DOSE_SQ 4     /* dose_to_water_square     :64-77  (stopping criterion) */
#define MQI_SCORER_DIJ 5         /* dose_to_water keyed by (voxel, spot): sparse hash table */
#define MQI_SCORER_LETT_NUMER 6  /* LETt_weight1             :141-158 */
#define MQI_SCORER_LETT_DENOM 7  /* LETt_weight2             :161-177 */

/* reference quirks that can be switched on for bug-for-bug comparisons (default: off) */
#define MQI_QUIRK_B2_DOUBLE_SCORE 1u /* mqi_transport.hpp:204-225: scorers [0,n-2) scored twice when n >= 3 */

/* Dense-scorer accumulation strategies (subsystem 4); results are identical up to fp64 summation order */
#define MQI_ACCUM_ATOMIC 0     /* one red.global.add.f64 per scored step */
#define MQI_ACCUM_WARP_MATCH 1 /* warp-aggregated: lanes hitting the same voxel are combined first */

typedef struct mqi_handle mqi_handle;

/* vertex_t<float>: moqui/base/mqi_vertex.hpp:13-29 */
typedef struct {
    float ke;
    float pos[3];
    float dir[3];
} mqi_vertex;

/* One spot = beamlet (energy pdf + 6-D phase-space pdf + coordinate transform):
 * moqui/base/mqi_beamlet.hpp:33-90, distributions/mqi_phsp6d.hpp, mqi_phsp6d_uniform.hpp,
 * mqi_const_1d.hpp, mqi_norm_1d.hpp, mqi_coordinate_transform.hpp:52-58. */
typedef struct {
    int32_t phsp_uniform;  /* 1: phsp_6d_uniform (U(-1,1)*sigma), 0: phsp_6d (N(0,1)*sigma) */
    int32_t energy_normal; /* 1: norm_1d(energy, sigma_energy), 0: const_1d(energy) */
    float   energy;
    float   sigma_energy;
    float   mean[6];  /* x y z x' y' z' */
    float   sigma[6]; /* sigma[2] scales Uz, sigma[5] ignored (z' from normalisation) */
    float   corr[2];  /* x-x', y-y' correlation */
    float   rot[9];   /* coordinate_transform::rotation, row-major */
    float   trans[3]; /* coordinate_transform::translation */
} mqi_beamlet;

typedef struct {
    uint64_t histories;        /* primaries transported by the last mqi_run (atomicAdd(tracked_particles), :245) */
    uint64_t steps;            /* scored voxel steps (optional counters, see mqi_set_option "count_steps") */
    uint64_t secondaries;      /* secondary protons pushed */
    uint64_t stack_overflows;  /* secondaries dropped because the stack was full (B10) */
    uint64_t dij_table_full;   /* Dij inserts dropped because the table was full (reference: infinite loop, B5) */
    float    kernel_ms;        /* device time of the transport kernel(s) of the last mqi_run (CUDA events) */
    uint32_t launches;         /* kernels launched by the last mqi_run */
} mqi_run_stats;

MQI_API const char* mqi_last_error(void);
MQI_API const char* mqi_version(void);
/* number of usable CUDA devices (0 if none; never fails) */
MQI_API int mqi_device_count(void);

/* cudaSetDevice(gpu_id): mqi_phantom_env.hpp:45, mqi_tps_env.hpp:163 */
MQI_API int mqi_create(int device_id, mqi_handle** out);
MQI_API int mqi_destroy(mqi_handle* h);
/* free and total bytes of HBM on the handle's device (cudaMemGetInfo); the tps front end sizes the Dij
 * table from it where the reference hard-codes 512*512*300*5 slots (mqi_tps_env.hpp:922) */
MQI_API int mqi_device_memory(mqi_handle* h, uint64_t* free_bytes, uint64_t* total_bytes);

/* __PHYSICS_DEBUG__ switch + quirk mask; must be called before mqi_set_grid_* (the calibration LUT
 * depends on it: materials/mqi_patient_materials.hpp:420-424,457-461). */
MQI_API int mqi_set_physics(mqi_handle* h, int variant, uint32_t quirks);

/* Patient / phantom grid (the world's child node): grid3d edges + HU volume.  Replaces
 * setup_world()'s host loop rho[i] = hu_to_density(hu[i]) * DensityScaling
 * (mqi_phantom_env.hpp:263-276, mqi_tps_env.hpp:759-769) and upload_node()
 * (kernel_functions/mqi_upload_data.hpp:148-398): the HU->density->(RSP, radiation length)
 * calibration is evaluated on the device into a 16-bit material volume + a material LUT.
 * hu: host pointer, int16 [nz][ny][nx] (x fastest).  rot: rotation_matrix_fwd row-major or NULL,
 * trans: translation_vector or NULL. */
MQI_API int mqi_set_grid_hu(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze,
                    int n_ze, const int16_t* hu, float density_scale, const float* rot, const float* trans);
/* Same but the HU volume is already resident in device memory (HBM-resident leg of bench.py). */
MQI_API int mqi_set_grid_hu_device(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye,
                           const float* ze, int n_ze, const void* d_hu, float density_scale,
                           const float* rot, const float* trans);
/* Raw mass densities (g/mm^3) instead of HU, e.g. range-shifter / aperture style nodes; at most
 * 65536 distinct values. */
MQI_API int mqi_set_grid_density(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye,
                         const float* ze, int n_ze, const float* rho, const float* rot, const float* trans);

/* scorer<R>(name, max_capacity, fp_compute_hit) + roi_t(DIRECT): mqi_scorer.hpp:64-73,
 * mqi_phantom_env.hpp:280-294, mqi_tps_env.hpp:850-933.  capacity is used by MQI_SCORER_DIJ only
 * (number of hash slots; the reference hard-codes 512*512*300*5, mqi_tps_env.hpp:922).
 * Returns the scorer index (>= 0) or a negative error. */
MQI_API int mqi_add_scorer(mqi_handle* h, int kind, const char* name, uint64_t capacity);
/* Let a dense scorer accumulate into caller-owned device memory (nvox doubles), e.g. a torch tensor
 * that is then reduced with NCCL.  The buffer is NOT cleared. */
MQI_API int mqi_bind_scorer_buffer(mqi_handle* h, int scorer, void* d_buffer);
/* Region of interest of a scorer: roi_t(CONTOUR) built by mask_reader::mask_to_roi from the summed 0/1
 * mask volumes (moqui/base/mqi_file_handler.hpp:107-113,176-217; mqi_roi.hpp:48-58,127-137; used by
 * ScoringMask / StatROIMaskFilename, mqi_tps_env.hpp:776-812).  mask_total: host pointer, one byte per
 * voxel of the scored grid ([nz][ny][nx]); a step is scored only if its voxel lies inside a run of the
 * mask.  NULL restores roi_t(DIRECT) (every voxel but voxel 0).  roi_size (optional)
 * receives get_mask_size().  Must be called after mqi_set_grid_*; a new grid drops the roi. */
MQI_API int mqi_set_scorer_roi(mqi_handle* h, int scorer, const uint8_t* mask_total, uint64_t n_voxels,
                       uint64_t* roi_size);
/* Beamline children of the world in front of the scored grid, in transport order: the range shifter
 * slab and the voxelised aperture of create_rangeshifter / create_voxelized_aperture
 * (mqi_tps_env.hpp:1605-1736), traversed by the c_ind loop of transport_particles_patient
 * (mqi_transport.hpp:162-240) before the patient grid.  rho: raw mass densities in g/mm^3
 * ([nz][ny][nx]; 1e-8 = open, 100 = closed aperture voxel: stepping stops the track at rho > 99.9,
 * mqi_fippel_physics.hpp:77-85); rot = rotation_matrix_fwd (row-major) or NULL, trans =
 * translation_vector or NULL.  Beamline nodes carry no scorers (mqi_tps_env.hpp:751).  Returns the
 * node index (>= 0) or a negative error; mqi_clear_beamline removes them all. */
MQI_API int mqi_add_beamline_node(mqi_handle* h, const float* xe, int n_xe, const float* ye, int n_ye, const float* ze,
                          int n_ze, const float* rho, const float* rot, const float* trans);
MQI_API int mqi_clear_beamline(mqi_handle* h);
MQI_API int mqi_clear_scorers(mqi_handle* h);
MQI_API int mqi_set_accumulation(mqi_handle* h, int mode);

/* Device-side beam source (subsystem 1).  Replaces beamsource::append_beamlet / operator()(h)
 * (mqi_beamsource.hpp:49-112) and the host sampling loops mqi_phantom_env.hpp:228-231,
 * mqi_tps_env.hpp:1202-1219,1547-1552: history h belongs to the spot whose cumulative history
 * count first exceeds h. */
MQI_API int mqi_set_beamlets(mqi_handle* h, const mqi_beamlet* beamlets, uint32_t n_spots,
                     const uint64_t* histories_per_spot);
/* Reference-compatible source: explicit vertices sampled by the caller (upload_vertices,
 * mqi_upload_data.hpp:415-445) and optional per-history spot ids (scorer_offset_vector). */
MQI_API int mqi_set_vertices(mqi_handle* h, const mqi_vertex* vertices, uint64_t n, const uint32_t* spot_ids);

/* transport_particles_patient<<<...>>> + cudaDeviceSynchronize (mqi_phantom_env.hpp:391-397,
 * mqi_tps_env.hpp:1129-1139).  Transports histories [first, first + count) of the source.
 * per_spot != 0 keys Dij entries by spot id (scorer_offset_vector != nullptr, :150-154).
 * Scorers accumulate across calls until mqi_clear_scorers (device tables persist across batches,
 * mqi_upload_data.hpp:250-253). */
MQI_API int mqi_run(mqi_handle* h, uint64_t seed, uint64_t first_history, uint64_t count, int per_spot);
/* Same launch without the trailing synchronisation: returns as soon as the kernel is queued on the
 * handle's stream; mqi_get_run_stats (or any download) waits for it.  Lets the caller overlap
 * transport with NCCL reductions / copies on other streams and time it with its own events. */
MQI_API int mqi_run_async(mqi_handle* h, uint64_t seed, uint64_t first_history, uint64_t count, int per_spot);
/* One shard of an interleaved partition of [first, first + count) over n_shards devices (new: the reference is
 * single-GPU): the range is cut into chunks of 32 histories and chunk c is transported by shard c % n_shards, so every
 * device gets the same mix of spots and energies -- contiguous sub-ranges of a plan sorted by energy layer leave the
 * device with the highest layers working longest.  The shards of all devices together transport every history of the
 * range exactly once, with the streams it would have drawn in a single launch; mqi_run_stats::histories counts this
 * shard's.  Not for per-spot (Dij) runs sharded by spot: rows must stay on one device. */
MQI_API int mqi_run_async_sharded(mqi_handle* h, uint64_t seed, uint64_t first_history, uint64_t count, int per_spot,
                          uint32_t n_shards, uint32_t shard);
/* waits for the last mqi_run_async and returns its counters and device time */
MQI_API int mqi_get_run_stats(mqi_handle* h, mqi_run_stats* out);
/* options: "count_steps" (0/1: fill mqi_run_stats::steps; runs the general kernel, ~2 % slower),
 * "blocks_per_sm" (cap on resident CTAs per SM, 0 = occupancy limit),
 * "l2_persist" (0/1: persisting L2 access window over the material volume; off by default, it measured
 * -0.25 % on the C1 workload whose hot part of the volume is cache resident anyway),
 * "dij_write_combine" (0/1, default 1: consecutive hits of a track on one (voxel, spot) key are summed in
 * registers and inserted into the Dij table once, when the voxel changes or the track ends; +46 % histories/s on the
 * 5 000-spot configuration),
 * "rsp_exact" (0/1: mqi_dev_rsp evaluates spr_default in the reference's own precision -- fp64 energy term, correctly
 * rounded pow -- and is then bit-exact against the reference; the transport kernel keeps the fp32 evaluation, within
 * 4 ulp, because the exact one costs 32 % of the C1 throughput (+46 % kernel time): DESIGN.md section 6),
 * "fetch_order" (0 default / 1: a launch hands out its chunks of 32 histories first to last (0) or last to first (1), so that a
 * plan listed by ascending energy starts its longest histories first and the tail of the persistent kernel is made of the
 * short ones; the same histories with the same streams either way.  Measured slower on whole plans) */
MQI_API int mqi_set_option(mqi_handle* h, const char* key, int64_t value);
/* Launch on a caller-owned cudaStream_t (e.g. the framework's current stream, so that its events
 * bracket the kernels) instead of the handle's own stream; NULL restores the handle's stream.  The
 * reference uses the default stream throughout (mqi_phantom_env.hpp:391-397). */
MQI_API int mqi_set_stream(mqi_handle* h, void* cuda_stream);

/* download_node + reshape_data (mqi_download_data.hpp:37-121, mqi_xenvironment.hpp:150-167):
 * dense float64 [nz][ny][nx] of a scorer, times scale. */
MQI_API int mqi_get_dense(mqi_handle* h, int scorer, double* out, double scale);
/* Dij: number of occupied slots, then the (voxel, spot, value) triplets in slot order
 * (mqi_io.hpp:105-160). */
MQI_API int mqi_get_sparse_count(mqi_handle* h, int scorer, uint64_t* nnz);
MQI_API int mqi_get_sparse(mqi_handle* h, int scorer, uint32_t* key1, uint32_t* key2, double* value, uint64_t nnz,
                   double scale);
/* device pointer / element count of a dense scorer's accumulator (for NCCL reduction by the caller) */
MQI_API int mqi_get_scorer_device_ptr(mqi_handle* h, int scorer, void** d_ptr, uint64_t* n_elements);

/* Stopping criterion (subsystem 5): calculate_standard_deviation (mqi_variables.hpp:20-48) +
 * host calculate_stat (mqi_tps_env.hpp:1339-1426) fused on the device.  sum / sumsq are dense
 * scorers of kinds DOSE and DOSE_SQ; out[0] = sum over selected voxels of sigma/mu, out[1] = number
 * of selected voxels (mean dose > threshold * max mean dose), out[2] = max mean dose.  The caller
 * all-reduces these partial sums across ranks before forming the percentage. */
MQI_API int mqi_stat_partial(mqi_handle* h, int scorer_sum, int scorer_sumsq, uint64_t n_histories,
                     double threshold_fraction, double max_mean_dose_or_negative, double out[3]);
/* Same criterion on caller-owned device buffers (n_voxels doubles each), e.g. the all-reduced sum /
 * sum-of-squares grids of a one-process-per-GPU deployment (moquimc_b200/parallel.py). */
MQI_API int mqi_stat_partial_buffers(mqi_handle* h, const void* d_sum, const void* d_sumsq, uint64_t n_voxels,
                             uint64_t n_histories, double threshold_fraction, double max_mean_dose_or_negative,
                             double out[3]);
/* calculate_average (mqi_variables.hpp:50-66): scale a dense scorer in place */
MQI_API int mqi_scale_scorer(mqi_handle* h, int scorer, double factor);

/* ---- multi-GPU inside one process (subsystem 5) ----
 * Sum dense scorer `scorer` of n handles (one per GPU, same grid) into the handle at index `root`
 * with ONE ncclReduce over NVLink (ncclCommInitAll communicator, created on first use and cached
 * for the handle set; NCCL is loaded with dlopen so the library also loads where NCCL is absent).
 * The reference has no multi-GPU mode (one cudaSetDevice per process, mqi_phantom_env.hpp:45).
 * With all_ranks != 0 an all-reduce leaves the sum on every handle (stopping-criterion passes).
 * One process per GPU deployments reduce the device pointer of mqi_get_scorer_device_ptr with their
 * own communicator instead (bench.py does, through torch.distributed). */
MQI_API int mqi_reduce_dense(mqi_handle* const* handles, int n, int scorer, int root);
MQI_API int mqi_allreduce_dense(mqi_handle* const* handles, int n, int scorer);
/* The stopping criterion of a run sharded over n devices (calculate_stat, mqi_tps_env.hpp:1339-1426, on the sums
 * of all devices) without gathering grids on one of them: the Dose / Dose^2 stat grids stay where they are and keep
 * accumulating.  Only the 4 096-voxel chunks in which some device holds a value that could exceed the criterion's
 * dose threshold after summation are packed and exchanged, one ncclReduceScatter per grid gives every device the summed
 * values of 1/n of them, each device evaluates its part, the host adds n x 3 doubles.  Result and selected voxels are
 * those of an evaluation on whole summed grids.  out as in mqi_stat_partial.  n == 1 is mqi_stat_partial. */
MQI_API int mqi_stat_multi(mqi_handle* const* handles, int n, int scorer_sum, int scorer_sumsq, uint64_t n_histories,
                   double threshold_fraction, double out[3]);

/* ---- deterministic device pieces, exposed for bit-exact parity tests ---- */
/* patient_material_t::hu_to_density on the device (materials/mqi_patient_materials.hpp:514-542) */
MQI_API int mqi_dev_hu_to_density(mqi_handle* h, const int16_t* hu, uint64_t n, float density_scale, float* rho_out);
/* spr_default / radiation_length_default (:414-473) evaluated through the device material LUT */
MQI_API int mqi_dev_rsp(mqi_handle* h, const float* rho, const float* ek, uint64_t n, float* rsp_out, float* rl_out);
/* grid3d::index(p,dir) + ijk2cnb + intersect(p,d,idx) + index(vtx1,dir1,idx) (mqi_grid3d.hpp:403-413,
 * 490-626, 745-877) on the grid set by mqi_set_grid_*: cell[3n], cnb[n] (~0 if outside), dist[n],
 * dir_after[3n], p_exit[3n], cell_after[3n] */
MQI_API int mqi_dev_grid_step(mqi_handle* h, const float* p, const float* d, uint64_t n, int32_t* cell,
                      uint64_t* cnb, float* dist, float* dir_after, float* p_exit, int32_t* cell_after);
/* grid3d::intersect(p,d) entry from outside (:631-743): dist[n] (-1 on a miss), cell[3n] */
MQI_API int mqi_dev_grid_entry(mqi_handle* h, const float* p, const float* d, uint64_t n, float* dist, int32_t* cell);
/* mc::hash_fun(k1,k2,capacity) (mqi_transport.hpp:32-51) */
MQI_API int mqi_dev_hash(mqi_handle* h, const uint32_t* k1, const uint32_t* k2, const uint64_t* capacity, uint64_t n,
                 uint32_t* out);
/* mc::insert_hashtable (mqi_transport.hpp:68-111) on its own: score n caller-supplied hits
 * (key1 = voxel, key2 = spot or 0xffffffff for the dense mode, value) into a scorer; hits with
 * value <= 0 are skipped like in the reference.  Host pointers. */
MQI_API int mqi_dev_insert(mqi_handle* h, int scorer, const uint32_t* key1, const uint32_t* key2, const double* value,
                   uint64_t n);
/* beamlet::operator() on the device for history ids [first, first+n) (subsystem 1) */
MQI_API int mqi_dev_sample_vertices(mqi_handle* h, uint64_t seed, uint64_t first, uint64_t n, mqi_vertex* out,
                            uint32_t* spot_out);

#ifdef __cplusplus
}
#endif
#endif
