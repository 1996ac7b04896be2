// mqi_kernels.h -- host-visible launch interface of the CUDA kernels (internal to libmqi_b200.so)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#define MQI_K_RELEASE 0
#define MQI_K_DEBUG 1

#define MQI_K_DOSE 0
#define MQI_K_EDEP 1
#define MQI_K_LETD_NUMER 2
#define MQI_K_LETD_DENOM 3
#define MQI_K_DOSE_SQ 4
#define MQI_K_DIJ 5
#define MQI_K_LETT_NUMER 6
#define MQI_K_LETT_DENOM 7

#define MQI_K_QUIRK_B2 1u
#define MQI_K_ACCUM_ATOMIC 0
#define MQI_K_ACCUM_WARP_MATCH 1

#ifndef MQI_K_BLOCK
#define MQI_K_BLOCK 768
#endif
#ifndef MQI_K_BLOCK_DIJ
#define MQI_K_BLOCK_DIJ MQI_K_BLOCK   /* threads per CTA of the single-Dij-scorer kernel (SET_DIJ with write-combining); C4 at the reference's
                                         table size: 640: 8.11e7, 768: 8.68e7 histories/s (profiles/r2_experiments.md) */
#endif
#ifndef MQI_K_BLOCK_MULTI
#define MQI_K_BLOCK_MULTI 640   /* threads per CTA of the multi-node kernels (worlds with beamline children): 96 registers per thread
                                   instead of 80 take the per-lane node descriptor without spilling (8 B against 84 B); measured on
                                   the range shifter + aperture workload 768: 1.12e8, 640: 1.15e8, 512: 1.07e8 histories/s */
#endif
#ifndef MQI_K_MIN_BLOCKS
#define MQI_K_MIN_BLOCKS 1   /* 24 warps/SM at <= 80 registers/thread in ONE CTA: one copy of the shared-memory tables leaves the most L1; measured best of 128x5 ... 768x1 on B200 (profiles/r1_experiments.md) */
#endif


#ifndef MQI_K_ADV_BATCH
#define MQI_K_ADV_BATCH 6    /* multi-node worlds: lanes of a warp that hand their track over to the next child together ... */
#endif
#ifndef MQI_K_ADV_TURNS
#define MQI_K_ADV_TURNS 10   /* ... or after the first of them has waited this many turns (1 / 0: 1.04e8, 12 / 6: 1.12e8, 6 / 10: 1.18e8
                                with 640 threads per CTA, profiles/r2_experiments.md) */
#endif

#ifndef MQI_K_ADV_QUEUE
#define MQI_K_ADV_QUEUE 1    /* multi-node worlds: 1 = tracks that leave a beamline child go through a per-warp hand-over buffer and are
                                located in the next child 32 at a time (process_handovers); 0 = batched hand-overs in the owning lane */
#endif
#ifndef MQI_K_ADV_PUSH_INLINE
#define MQI_K_ADV_PUSH_INLINE __noinline__
#endif
#ifndef MQI_K_ADV_MIN
#define MQI_K_ADV_MIN 32     /* buffered hand-overs that make the warp process them instead of fetching primaries when its queue is empty */
#endif
#ifndef MQI_K_ADV_RAW_CAP
#define MQI_K_ADV_RAW_CAP 96 /* entries of a warp's hand-over buffer (48 B each, global memory); a lane that finds it full falls back to
                                restart_lane */
#endif

#ifndef MQI_K_RSP_EXACT
#define MQI_K_RSP_EXACT 0    /* 1: the transport kernel evaluates spr_default in the reference's precision (rsp_eval_exact) */
#endif

#ifndef MQI_K_PROBE_WIDTH
#define MQI_K_PROBE_WIDTH 1  /* slots of the Dij table whose keys the FIRST step of a probe sequence loads together (dij_probe_from) */
#endif
#ifndef MQI_K_PROBE_WIDTH2
#define MQI_K_PROBE_WIDTH2 4 /* ... and every later step.  C4 at the reference's table size / at 1.6e9 slots, histories/s: every step
                                1 slot 6.3e7, 2 slots 8.56e7 / 1.02e8, 4 slots 7.3e7; first step 1 slot and later steps 2: 8.61e7,
                                3: 8.32e7, 4: 8.80e7 / 1.05e8 (kept), 5: 8.55e7, 6: 8.53e7 / 1.00e8, 8: 7.98e7; first 2 and later 4: 8.40e7,
                                6: 8.26e7, 8: 7.86e7 (profiles/r2_experiments.md) */
#endif

#ifndef MQI_K_EXACT_DIV
#define MQI_K_EXACT_DIV 1    /* 1: the three voxel-exit divisions of a step are correctly rounded (bit-exact geometry against the reference's
                                CPU build); 0: rcp.approx * n as in the reference's own --use_fast_math CUDA build (measured: see
                                profiles/r2_experiments.md) */
#endif

#ifndef MQI_K_LATE_LUT
#define MQI_K_LATE_LUT 1   /* delay the material LUT load behind the step's random numbers (see mqi_transport.cu) */
#endif

#ifndef MQI_K_PHILOX_ROUNDS
#define MQI_K_PHILOX_ROUNDS 7    /* rounds of every Philox4x32 block of the RNG protocol; the oracle uses the same number (MQO_PHILOX_ROUNDS) */
#endif

namespace mqib
{
struct Params;
struct MatEntry;
struct BeamletDev;
struct VertexDev;

int         transport_block(bool multi, bool dij_set);
int         transport_block(const Params& p);
size_t      transport_smem_bytes(int n_edge_floats, int n_nodes);
size_t      transport_handover_bytes(int grid, int n_nodes);   // per-launch scratch of a multi-node world (0: none)
bool        transport_is_simple(const Params& p);
cudaError_t transport_occupancy(const Params& p, int variant, size_t smem, int* blocks_per_sm);
cudaError_t launch_transport(const Params& p, int variant, int grid, size_t smem, cudaStream_t st);

// HU volume -> 16-bit material index volume (index = clamp(hu) + 1000)
cudaError_t launch_hu_to_material(const int16_t* d_hu, uint16_t* d_mat, size_t n, cudaStream_t st);
// HU -> density on the device (bit-exact restatement of hu_to_density)
cudaError_t launch_hu_to_density(const int16_t* d_hu, float* d_rho, size_t n, const float* d_correction,
                                 float density_scale, cudaStream_t st);
cudaError_t launch_dev_rsp(const MatEntry* lut_entries, const float* ek, size_t n, float* rsp, float* rl,
                           cudaStream_t st, bool exact = false);
cudaError_t launch_dev_grid_step(const Params& p, const float* pin, const float* din, size_t n, int32_t* cell,
                                 unsigned long long* cnb, float* dist, float* dir_after, float* p_exit,
                                 int32_t* cell_after, cudaStream_t st);
cudaError_t launch_dev_grid_entry(const Params& p, const float* pin, const float* din, size_t n, float* dist,
                                  int32_t* cell, cudaStream_t st);
cudaError_t launch_dev_hash(const uint32_t* k1, const uint32_t* k2, const unsigned long long* cap, size_t n,
                            uint32_t* out, cudaStream_t st);
cudaError_t launch_dev_sample(const Params& p, unsigned long long first, size_t n, VertexDev* out,
                              uint32_t* spot, cudaStream_t st);
cudaError_t launch_dev_insert(const Params& p, int scorer, const uint32_t* k1, const uint32_t* k2, const double* v,
                              size_t n, cudaStream_t st);
cudaError_t launch_fill_u64(unsigned long long* p, unsigned long long v, size_t n, cudaStream_t st);
cudaError_t launch_dij_clear(void* table, size_t capacity, cudaStream_t st);
cudaError_t launch_dij_count(const void* table, size_t capacity, unsigned long long* d_count, cudaStream_t st);
cudaError_t launch_dij_chunk_count(const void* table, size_t capacity, size_t chunk,
                                   unsigned long long* d_chunk_count, cudaStream_t st);
cudaError_t launch_dij_chunk_write(const void* table, size_t capacity, size_t chunk,
                                   const unsigned long long* d_chunk_off, uint32_t* k1, uint32_t* k2, double* val,
                                   double scale, cudaStream_t st);
cudaError_t launch_scale(double* p, size_t n, double f, cudaStream_t st);
// stopping criterion: per-voxel mean / sigma from sum and sum of squares, partial reductions
cudaError_t launch_stat_max(const double* sum, size_t n, double inv_n, double* d_out_max, cudaStream_t st);
// stopping criterion over several devices: flag[c] = any value of chunk c above bound; gather of listed chunks
cudaError_t launch_chunk_above(const double* a, size_t n, size_t chunk, double bound, unsigned char* d_flag, cudaStream_t st);
cudaError_t launch_pack_chunks(const double* a, const double* b, size_t n, size_t chunk, const unsigned int* d_list, size_t n_list,
                               double* pa, double* pb, cudaStream_t st);
cudaError_t launch_stat_partial(const double* sum, const double* sumsq, size_t n, double n_hist, double cut,
                                double* d_out2, cudaStream_t st);
}   // namespace mqib
