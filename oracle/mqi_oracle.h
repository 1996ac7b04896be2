/* mqi_oracle.h -- CPU restatement (plain C99) of moqui's per-history proton transport path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under moquimc_b200/ may include, link or execute this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker.
 *
 * Parity status: PINNED.  The deterministic functions are checked bit-for-bit against vectors
 * produced by the reference's own headers (oracle/ref_kat.cpp -> tests/golden/kat_*.npz) and the
 * stochastic transport is checked against dose files written by the reference's own CPU build
 * (oracle/_ref/phantom_env_cpu_*, oracle/ref_run.py -> tests/golden/c1_*.npz).  The reference ships
 * no tests or golden vectors of its own.
 *
 * Every function cites the reference file:line (relative to /root/reference/moqui) it follows.
 * Random numbers: the reference uses std::default_random_engine (CPU) / curand XORWOW (GPU); this
 * restatement uses the counter-based Philox4x32-7 protocol of DESIGN.md ("RNG protocol") so that
 * it can be compared history by history with the CUDA path.
 */
#ifndef MQI_ORACLE_H
#define MQI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQO_VARIANT_RELEASE 0 /* tps_env build: no __PHYSICS_DEBUG__ */
#define MQO_VARIANT_DEBUG 1   /* phantom_env build: -D__PHYSICS_DEBUG__ (tests/mc/phantom/CMakeLists.txt:10) */

#define MQO_SCORER_DOSE 0   /* dose_to_water      scorers/mqi_scorer_energy_deposit.hpp:43-60 */
#define MQO_SCORER_EDEP 1   /* energy_deposit     :14-19 */
#define MQO_SCORER_LETD_NUMER 2 /* LETd_weight1   :93-114 */
#define MQO_SCORER_LETD_DENOM 3 /* LETd_weight2   :117-137 */
#define MQO_SCORER_DOSE_SQ 4    /* dose_to_water_square :64-77 */
#define MQO_SCORER_DIJ 5        /* dose_to_water keyed by (voxel, spot), hashed */
#define MQO_SCORER_LETT_NUMER 6 /* LETt_weight1   :141-158 */
#define MQO_SCORER_LETT_DENOM 7 /* LETt_weight2   :161-177 */

#define MQO_QUIRK_B2_DOUBLE_SCORE 1u /* mqi_transport.hpp:204-225 scores scorers [0,n-2) twice when n>=3 */

typedef struct {
    int    nx, ny, nz;
    const float* xe; /* nx+1 */
    const float* ye;
    const float* ze;
    const float* rho;      /* g/mm^3, [nz][ny][nx] */
    float        rot_fwd[9]; /* row-major; identity for the patient grid */
    float        trans[3];
} mqo_grid;

typedef struct {
    int   phsp_uniform;  /* 1: phsp_6d_uniform, 0: phsp_6d (gaussian) */
    int   energy_normal; /* 1: norm_1d, 0: const_1d */
    float energy, sigma_energy;
    float mean[6];
    float sigma[6];
    float corr[2];
    float rot[9]; /* coordinate_transform rotation, row-major */
    float trans[3];
} mqo_beamlet;

typedef struct {
    float ke, pos[3], dir[3];
} mqo_vertex;

typedef struct {
    uint32_t key1, key2;
    double   value;
} mqo_key_value;

typedef struct {
    int            kind;
    double*        dense;    /* nvox doubles (NULL for DIJ) */
    mqo_key_value* table;    /* DIJ only */
    uint64_t       capacity; /* DIJ only */
} mqo_scorer;

typedef struct {
    uint64_t histories;
    uint64_t steps;        /* scored-step loop iterations (== reference intersect(p,d,idx) calls) */
    uint64_t along_steps;
    uint64_t delta_events;
    uint64_t pp_events, poe_events, poi_events;
    uint64_t secondaries_pushed;
    uint64_t max_stack;
} mqo_stats;

/* ---- deterministic pieces (bit-exact against the reference) ---- */
int      mqo_load_tables(const char* path); /* moquimc_b200/data/mqi_tables_v1.bin */
float    mqo_hu_to_density(int16_t hu);
float    mqo_spr(float rho_mass, float ek, int variant);
float    mqo_radiation_length(float rho_mass, int variant);
uint32_t mqo_hash(uint32_t k1, uint32_t k2, uint64_t capacity);
void     mqo_start_and_length(uint32_t n_threads, uint32_t n_jobs, uint32_t tid, uint32_t out[2]);
void     mqo_grid_index(const mqo_grid* g, const float p[3], const float d[3], int cell[3]);
void     mqo_grid_index_update(const mqo_grid* g, const float p[3], const float d[3], int cell[3]);
float    mqo_grid_intersect_cell(const mqo_grid* g, const float p[3], float d[3], const int cell[3]);
float    mqo_grid_intersect_entry(const mqo_grid* g, const float p[3], float d[3], int cell[3]);
void     mqo_rotate_direction(const float dir_in[3], float theta, float phi, float dir_out[3]);
/* out[9] = beta_sq, gamma, Te_max, momentum, cs_pion, dEdx, cs_pp, cs_poe, cs_poi (rho = 1e-3) */
void     mqo_physics_probe(float ek, float out[9]);

/* ---- RNG protocol (shared with the CUDA path) ---- */
#define MQO_PHILOX_ROUNDS 7   /* rounds of every Philox4x32 block of the protocol (== MQI_K_PHILOX_ROUNDS of the kernel) */
void  mqo_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
void  mqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float mqo_u32_to_uniform(uint32_t x);
void  mqo_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]);

/* ---- source sampling: beamlet::operator() mqi_beamlet.hpp:81-90 ---- */
void mqo_sample_vertex(const mqo_beamlet* b, uint64_t seed, uint64_t history, mqo_vertex* out);

/* ---- the transport path: transport_particles_patient mqi_transport.hpp:113-250 ----
 * histories [h0, h0+n) of the beam source (spot of history h found through cum_histories) or, if
 * vertices != NULL, explicit vertices[i] / spot_ids[i] for i in [0,n) with history id h0+i. */
void mqo_insert(mqo_scorer* s, const uint32_t* key1, const uint32_t* key2, const double* value, uint64_t n);
int mqo_transport(const mqo_grid* g, int variant, uint32_t quirks, const mqo_beamlet* beamlets,
                  const uint64_t* cum_histories, uint32_t n_beamlets, const mqo_vertex* vertices,
                  const uint32_t* spot_ids, int per_spot, uint64_t seed, uint64_t h0, uint64_t n,
                  mqo_scorer* scorers, int n_scorers, mqo_stats* stats);

/* Multi-node world (beamline children first, patient grid last, mqi_tps_env.hpp:732-758) and per-scorer
 * CONTOUR regions of interest (expanded masks, see mqo_mask_to_roi); scorers belong to the last node. */
int mqo_transport_nodes(const mqo_grid* nodes, int n_nodes, int variant, uint32_t quirks, const mqo_beamlet* beamlets,
                        const uint64_t* cum_histories, uint32_t n_beamlets, const mqo_vertex* vertices,
                        const uint32_t* spot_ids, int per_spot, uint64_t seed, uint64_t h0, uint64_t n,
                        mqo_scorer* scorers, int n_scorers, const uint8_t* const* roi_masks, mqo_stats* stats);
/* mask_reader::mask_to_roi (mqi_file_handler.hpp:176-217) on the summed mask volume */
uint32_t mqo_mask_to_roi(const uint8_t* mask_total, uint64_t n, uint32_t* start, uint32_t* stride, uint32_t max_runs,
                         uint8_t* member);

#ifdef __cplusplus
}
#endif
#endif
