// mqi_transport.cu -- the per-history proton transport kernel for sm_100a, fused with the device
// beam source and the scorers.  Replaces mc::transport_particles_patient(_stat)
// (kernel_functions/mqi_transport.hpp:113-389), fippel_physics::stepping
// (base/mqi_fippel_physics.hpp:67-216) and the four interaction classes, the host-side vertex
// sampling loops, initialize_threads / per-history curand_init, and insert_hashtable.
//
// Design: persistent grid, one lane = one history at a time; every loop iteration is
// exactly one voxel step of whichever track the lane currently owns, so lanes stay converged on the
// step body and only re-arm (pop a secondary / fetch the next history) under a short divergent
// prologue.  Physics tables and grid edges live in shared memory, the 16-bit material volume is read
// through the read-only path, the material LUT through L1.  Delta electrons of the debug variant are
// folded into the parent's step instead of being pushed as zero-energy tracks.
#include "mqi_device.cuh"
#include "mqi_kernels.h"

#include <algorithm>

namespace mqib
{

// ---------------------------------------------------------------------------------------------
// shared-memory layout: physics tables | grid edges of every node | node descriptors (multi-node
// launches) | per-warp queues of pre-sampled primaries
// ---------------------------------------------------------------------------------------------
struct Smem {
    const float4* a0;   // {cs_p_ion, slope, restricted stopping power, slope}   per 0.5 MeV row, Ei = 0.1
    const float4* a1;   // {csda range, slope, dE/drange (inverse slope), 0}
    const float2* bs;   // {cs_pp + cs_pO_el + cs_pO_inel, slope}                  Ei = 0.5
    const float*  edges;   // node k: xe | ye | ze at edges + nodes[k].edge_off (single node: offset 0)
    const GridDev* nodes;  // multi-node launches only
    uint32_t*     queue;   // queue_words(multi) words per warp
};

// queue of pre-sampled primaries, one per warp, structure of arrays: field f of entry e at
// [f * kQueueCap + e].  Filled by the whole warp when it has run empty (refill_queue: up to 32 entries),
// drained by the lanes whose track ended.  Shared memory is taken from the L1 carve-out, so the queue
// is kept as small as one refill.
constexpr int kQueueCap    = 32;
// Q_BLK (multi-node launches only): the Philox block the track continues with -- a queue entry is then either a fresh
// primary or a track handed over from one beamline child to the next (process_handovers)
enum QueueField { Q_PX = 0, Q_PY, Q_PZ, Q_DX, Q_DY, Q_DZ, Q_KE, Q_IX, Q_IY, Q_IZ, Q_H0, Q_H1, Q_SPOT, Q_NODE, Q_FIELDS, Q_BLK = Q_FIELDS };
__host__ __device__ constexpr int queue_words(bool multi) { return (multi ? Q_FIELDS + 1 : Q_FIELDS) * kQueueCap; }

// Hand-over buffer (multi-node launches, MQI_K_ADV_QUEUE): tracks that left a beamline child alive, in that child's frame,
// waiting to be mapped to the world and located in the next child by the whole warp at once.  One buffer per warp in global
// memory (written and read through L2: 2 pushes per history), structure of arrays like the queue.
constexpr int kRawCap = MQI_K_ADV_RAW_CAP;
enum RawField { R_PX = 0, R_PY, R_PZ, R_DX, R_DY, R_DZ, R_KE, R_H0, R_H1, R_SPOT, R_NODE, R_BLK, R_FIELDS };
constexpr int kRawWords = R_FIELDS * kRawCap;
constexpr size_t kTableBytes = kTableN * (2 * sizeof(float4) + sizeof(float2));

__host__ __device__ __forceinline__ size_t
smem_nodes_offset(int n_edge_floats) { return (kTableBytes + (size_t) n_edge_floats * sizeof(float) + 15) & ~(size_t) 15; }
__host__ __device__ __forceinline__ size_t
smem_queue_offset(int n_edge_floats, int n_nodes) {
    return smem_nodes_offset(n_edge_floats) + ((n_nodes > 1 ? (size_t) n_nodes * sizeof(GridDev) : 0) + 15 & ~(size_t) 15);
}

__device__ __forceinline__ Smem
smem_view(unsigned char* raw, int n_edge_floats, int n_nodes) {
    Smem sm;
    float4* a0 = reinterpret_cast<float4*>(raw);
    float4* a1 = a0 + kTableN;
    float2* bs = reinterpret_cast<float2*>(a1 + kTableN);
    sm.a0 = a0; sm.a1 = a1; sm.bs = bs;
    sm.edges = reinterpret_cast<float*>(bs + kTableN);
    sm.nodes = reinterpret_cast<const GridDev*>(raw + smem_nodes_offset(n_edge_floats));
    sm.queue = reinterpret_cast<uint32_t*>(raw + smem_queue_offset(n_edge_floats, n_nodes));
    return sm;
}

// the node a lane is in: a shared-memory descriptor in multi-node launches, else the kernel parameter
template<bool MULTI>
__device__ __forceinline__ const GridDev&
node_ref(const Params& P, const Smem& sm, int node) {
    if (MULTI) return sm.nodes[node];
    return P.g;
}

// Cross sections [mm^2/g]: delta production on the p-ion grid plus the three nuclear channels
// (pre-summed for the mean free path: linear interpolation commutes with the sum), mqi_p_ionization.hpp:
// 254-268, mqi_pp_elastic.hpp:221-235, mqi_po_elastic.hpp:243-256, mqi_po_inelastic.hpp:141-155.
// row of the p-ionisation grid (Ei = 0.1, step 0.5): uint16_t((Ek - Ei) / 0.5)
// float -> unsigned conversion saturates negative arguments (and NaN) to 0: one clamp instead of two
__device__ __forceinline__ int row_a(float ek) { return (int) min(__float2uint_rz((ek - 0.1f) * 2.0f), (unsigned) (kTableN - 1)); }
__device__ __forceinline__ int row_b(float ek) { return (int) min(__float2uint_rz((ek - 0.5f) * 2.0f), (unsigned) (kTableN - 1)); }

// |dEdx| in water (restricted stopping power), mqi_p_ionization.hpp:271-286
__device__ __forceinline__ float
stopping_power(const Smem& sm, float ek) {
    if (ek >= 0.1f && ek <= 299.6f) {
        const int    i = row_a(ek);
        const float4 a = sm.a0[i];
        return fmaf(ek - (0.1f + i * 0.5f), a.w, a.z);
    }
    if (ek < 0.1f && ek > 0.f) return sm.a0[0].z;
    return 0.f;
}

// ---------------------------------------------------------------------------------------------
// RNG plumbing.  Protocol (shared with the oracle): Philox4x32-7, key = seed, counter =
// (block, 0, history_lo, history_hi); one aligned block {u_mfp, u_a, u_b, u_phi} per physics step,
// further blocks on demand inside a discrete interaction.  The per-step block is generated inline;
// every other use goes through one out-of-line copy to keep the hot loop small.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ uint4
philox_block(uint32_t blk, uint32_t h0, uint32_t h1, uint32_t k0, uint32_t k1) {
    uint32_t o[4];
    philox4x32<kPhiloxRounds>(blk, 0u, h0, h1, k0, k1, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}

struct RngBuf {
    uint4    w;
    int      pos;
    uint32_t blk, h0, h1, k0, k1;
};
__device__ __forceinline__ float
rb_uniform(RngBuf& r) {
    if (r.pos == 4) {
        r.w   = philox_block(r.blk, r.h0, r.h1, r.k0, r.k1);
        r.blk += 1;
        r.pos = 0;
    }
    const int      p = r.pos++;
    const uint32_t x = p == 0 ? r.w.x : (p == 1 ? r.w.y : (p == 2 ? r.w.z : r.w.w));
    return u32_to_uniform(x);
}

// ---------------------------------------------------------------------------------------------
// secondary stack (per lane, local memory): base/mqi_track_stack.hpp:11-67
// ---------------------------------------------------------------------------------------------
template<int VARIANT>
struct StackCfg {
    static constexpr int depth = (VARIANT == MQI_K_DEBUG) ? 16 : 10;
};

struct Secondary {
    float px, py, pz, dx, dy, dz;
    float ke0;      // vtx0.ke
    float ke1_off;  // vtx1.ke - vtx0.ke at creation (debug recoil daughters: -ke0), else 0
    float dE_pre;   // energy already carried as trk.dE (debug recoil daughters), else 0
};

// in/out record of the (rare) nuclear interactions, kept in local memory
struct NucIO {
    float px, py, pz, dx, dy, dz;   // pre-step vertex (vtx0)
    float p1x, p1y, p1z;            // post-step position (vtx1.pos)
    float d1x, d1y, d1z;            // vtx1.dir, in/out
    float ke1;                      // vtx1.ke, in/out
    float dE, local_dE;             // deposits, in/out
    float u, e_cs, rho;             // selector (already reduced by the delta channel), energy of the cross sections
    float c1, c2, c3;               // nuclear channel cross sections (filled by nuclear_event)
    int   stopped;
    int   sp;                       // stack pointer, in/out
    int   node;                     // node the interaction happens in (multi-node launches)
    unsigned n_sec, n_ovf;          // counters, out
    RngBuf rb;
};

template<int VARIANT, bool MULTI>
__device__ __forceinline__ void
push_secondary(const Params& P, Secondary* stack, NucIO& io, float x, float y, float z, float ux, float uy,
               float uz, float ke0, float ke1_off, float dE_pre) {
    constexpr int DEPTH = StackCfg<VARIANT>::depth;
    if (io.sp >= DEPTH) {   // overflow silently drops the secondary: mqi_track_stack.hpp:36-41 (B10)
        ++io.n_ovf;
        return;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem     sm = smem_view(smem_raw, P.n_edge_floats, P.n_nodes);
    const GridDev& G  = node_ref<MULTI>(P, sm, io.node);
    if (!G.identity) {
        // daughters are mapped with Rfwd * (x - T) + T, mqi_pp_elastic.hpp:188-195
        const float* R  = G.rot_fwd;
        const float  qx = x - G.trans[0], qy = y - G.trans[1], qz = z - G.trans[2];
        x = R[0] * qx + R[1] * qy + R[2] * qz + G.trans[0];
        y = R[3] * qx + R[4] * qy + R[5] * qz + G.trans[1];
        z = R[6] * qx + R[7] * qy + R[8] * qz + G.trans[2];
        const float ex = ux, ey = uy, ez = uz;
        ux = R[0] * ex + R[1] * ey + R[2] * ez;
        uy = R[3] * ex + R[4] * ey + R[5] * ez;
        uz = R[6] * ex + R[7] * ey + R[8] * ez;
    }
    Secondary& s = stack[io.sp++];
    s.px = x; s.py = y; s.pz = z; s.dx = ux; s.dy = uy; s.dz = uz;
    s.ke0 = ke0; s.ke1_off = ke1_off; s.dE_pre = dE_pre;
    ++io.n_sec;
}

// p-p elastic, p-O elastic and p-O inelastic post-step interactions (0.5 events per 200 MeV history):
// out of line so that the voxel-step loop stays small.
template<int VARIANT, bool MULTI>
__device__ __noinline__ void
nuclear_event(const Params& P, Secondary* stack, NucIO& io) {
    RngBuf& rb = io.rb;
    {   // the three nuclear channels at the energy whose total was the larger one (:111-118)
        io.c1 = io.c2 = io.c3 = 0.f;
        const float e = io.e_cs;
        if (e >= 0.5f && e <= 300.0f) {
            const int    i  = row_b(e);
            const float  t  = e - (0.5f + i * 0.5f);
            const float4 n0 = __ldg(P.tab_n0 + i);
            const float2 n1 = __ldg(P.tab_n1 + i);
            io.c1 = fmaf(t, n0.y, n0.x) * io.rho;
            io.c2 = fmaf(t, n0.w, n0.z) * io.rho;
            io.c3 = fmaf(t, n1.y, n1.x) * io.rho;
        }
    }
    if (io.u < io.c1) {
        // p-p elastic, pp_elastic_tabulated::post_step mqi_pp_elastic.hpp:119-219
        const Rel   r1   = rel_make(io.ke1);
        const float minv = kTpCut / r1.Ek;
        const float uu   = rb_uniform(rb) * (1.0f - 2.0f * minv) + minv;
        const float E1 = r1.Et;
        const float dE = r1.Ek * uu;
        const float E3 = (r1.Ek - dE) + kMp;
        const float E4 = dE + kMp;
        const float P1 = sqrtf(r1.Et * r1.Et - kMpSq);
        const float P3 = sqrtf(E3 * E3 - kMpSq);
        const float P4 = sqrtf(E4 * E4 - kMpSq);
        float cos_th3  = (E1 * E3 - kMpSq - kMp * (E1 - E3)) / (P1 * P3);
        float cos_th34 = (E3 * E4 - E1 * kMp) / (P3 * P4);
        cos_th3  = fminf(1.f, fmaxf(-1.f, cos_th3));
        cos_th34 = fminf(1.f, fmaxf(-1.f, cos_th34));
        const float th3 = acosf(cos_th3);
        const float th4 = th3 - acosf(cos_th34);
        const float phi = kTwoPi * rb_uniform(rb);
        io.ke1 -= dE;
        rotate_direction(io.d1x, io.d1y, io.d1z, th3, phi);
        // recoil proton: direction rotated from the already scattered primary direction
        float sx = io.d1x, sy = io.d1y, sz = io.d1z;
        rotate_direction(sx, sy, sz, th4, phi);
        push_secondary<VARIANT, MULTI>(P, stack, io, io.p1x, io.p1y, io.p1z, sx, sy, sz, dE, 0.f, 0.f);
    } else if (io.u < io.c1 + io.c2) {
        // p-O elastic, po_elastic::post_step mqi_po_elastic.hpp:97-217
        const Rel r1 = rel_make(io.ke1);
        if (r1.Ek <= 5.5f) {
            const float dE = r1.Ek;
            if (VARIANT == MQI_K_DEBUG)
                push_secondary<VARIANT, MULTI>(P, stack, io, io.p1x, io.p1y, io.p1z, io.d1x, io.d1y, io.d1z, dE, -dE, dE);
            else io.local_dE += dE;
            io.ke1 -= dE;
            io.stopped = 1;
        } else {
            const float Tp_avg = 0.65f * expf(-0.0013f * r1.Ek) - 0.71f * expf(-0.0177f * r1.Ek);
            const float Tp_max = (2.0f * kMo * r1.beta_sq * r1.gamma_sq) /
                                 (1.0f + 2.0f * r1.gamma * kMoMp + kMoMp * kMoMp);
            float dE;
            do {   // mqi_exponential (GPU definition, truncated) base/mqi_math.hpp:298-307
                dE = -Tp_avg * logf(1.0f - rb_uniform(rb));
            } while (dE > Tp_max || dE != dE);
            const float E1 = r1.Ek * (r1.Ek + 2.0f * kMp);
            const float E3 = (r1.Ek - dE) * (r1.Ek - dE + 2.0f * kMp);
            float cos_th3  = (E1 + E3 - dE * (dE + 2.0f * kMo)) / 2.0f / sqrtf(E1 * E3);
            cos_th3        = fminf(1.f, fmaxf(-1.f, cos_th3));
            const float th3 = acosf(cos_th3);
            const float phi = kTwoPi * rb_uniform(rb);
            if (VARIANT == MQI_K_DEBUG)   // daughter starts at the parent's PRE-step vertex
                push_secondary<VARIANT, MULTI>(P, stack, io, io.px, io.py, io.pz, io.dx, io.dy, io.dz, dE, -dE, dE);
            else io.local_dE += dE;
            io.ke1 -= dE;
            rotate_direction(io.d1x, io.d1y, io.d1z, th3, phi);
        }
    } else if (io.u < io.c1 + io.c2 + io.c3) {
        // p-O inelastic cascade, po_inelastic_tabulated::post_step mqi_po_inelastic.hpp:159-288
        const float Ek = io.ke1;
        float       Eb = 5.0f, Er = Ek;
        float       prob_2nd, prob_long, power;
        if (Ek <= 215.f && Ek > 200.f) { prob_2nd = 0.78f; prob_long = prob_2nd + (1.f - prob_2nd) * 0.9f; power = 0.4f; }
        else if (Ek > 215.f) { prob_2nd = 0.78f; prob_long = prob_2nd + (1.f - prob_2nd) * 1.0f; power = 0.4f; }
        else if (Ek <= 200.f && Ek > 150.f) { prob_2nd = 0.72f; prob_long = prob_2nd + (1.f - prob_2nd) * 0.83f; power = 0.45f; }
        else { prob_2nd = 0.7f; prob_long = prob_2nd + (1.f - prob_2nd) * 0.83f; power = 0.52f; }
        while ((Er - Eb) > 2.0f) {
            Er -= Eb;
            const float uu = rb_uniform(rb);
            float       dE = powf(uu, power) * (Er - 2.0f) + 2.0f;
            if (dE >= Er) dE = Er;
            Er -= dE;
            io.ke1 -= (dE + Eb);
            const float zeta = rb_uniform(rb);
            if (zeta < prob_2nd) {
                float cos_th = (2.0f * dE / Ek - 1.0f) + 2.0f * (1.f - dE / Ek) * rb_uniform(rb);
                cos_th       = fminf(1.f, fmaxf(-1.f, cos_th));
                const float th  = acosf(cos_th);
                const float phi = kTwoPi * rb_uniform(rb);
                float sx = io.d1x, sy = io.d1y, sz = io.d1z;
                rotate_direction(sx, sy, sz, th, phi);
                push_secondary<VARIANT, MULTI>(P, stack, io, io.p1x, io.p1y, io.p1z, sx, sy, sz, dE, 0.f, 0.f);
            } else if (zeta < prob_long) {
                // neutral / long-range: energy leaves
            } else if (VARIANT == MQI_K_DEBUG) {
                push_secondary<VARIANT, MULTI>(P, stack, io, io.px, io.py, io.pz, io.dx, io.dy, io.dz, dE, -dE, dE);
            }   // release: short-range energy dropped (:273-275, B4)
            Eb *= 0.65f;
        }
        io.dE += Er;
        io.ke1 -= Er;
        io.stopped = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// device beam source: beamlet::operator()  base/mqi_beamlet.hpp:81-90 with
// phsp_6d_uniform / phsp_6d (distributions/mqi_phsp6d_uniform.hpp:68-85, mqi_phsp6d.hpp:62-79),
// const_1d / norm_1d.  RNG protocol: block 0 = {Ux, Vx, Uy, Vy}, block 1 = {za, zb, uc, -}.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void
sample_vertex(const BeamletDev& b, unsigned long long seed, unsigned long long h, VertexDev& out) {
    const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32), h0 = (uint32_t) h, h1 = (uint32_t) (h >> 32);
    float Ux, Vx, Uy, Vy;
    {
        const uint4 w  = philox_block(0u, h0, h1, k0, k1);
        const float u0 = u32_to_uniform(w.x), u1 = u32_to_uniform(w.y), u2 = u32_to_uniform(w.z), u3 = u32_to_uniform(w.w);
        if (b.phsp_uniform) {
            Ux = 2.0f * u0 - 1.0f; Vx = 2.0f * u1 - 1.0f; Uy = 2.0f * u2 - 1.0f; Vy = 2.0f * u3 - 1.0f;
        } else {
            box_muller(u0, u1, Ux, Vx);
            box_muller(u2, u3, Uy, Vy);
        }
    }
    float za = 0.f, zb = 0.f, uc = 0.5f;
    if (b.energy_normal || b.sigma[2] != 0.f) {   // block 1 only feeds the energy and the z spread
        const uint4 w = philox_block(1u, h0, h1, k0, k1);
        box_muller(u32_to_uniform(w.x), u32_to_uniform(w.y), za, zb);
        uc = u32_to_uniform(w.z);
    }
    const float Uz = b.phsp_uniform ? 2.0f * uc - 1.0f : za;
    float       ph[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) ph[i] = b.mean[i];
    ph[0] += b.sigma[0] * Ux;
    ph[1] += b.sigma[1] * Uy;
    ph[2] += b.sigma[2] * Uz;
    ph[3] += b.sigma[3] * (b.corr[0] * Ux + Vx * sqrtf(1.0f - b.corr[0] * b.corr[0]));
    ph[4] += b.sigma[4] * (b.corr[1] * Uy + Vy * sqrtf(1.0f - b.corr[1] * b.corr[1]));
    ph[5] = -1.0f * sqrtf(1.0f - ph[3] * ph[3] - ph[4] * ph[4]);
    out.ke = b.energy_normal ? zb * b.sigma_energy + b.energy : b.energy;
    const float* R = b.rot;
    out.pos[0] = R[0] * ph[0] + R[1] * ph[1] + R[2] * ph[2] + b.trans[0];
    out.pos[1] = R[3] * ph[0] + R[4] * ph[1] + R[5] * ph[2] + b.trans[1];
    out.pos[2] = R[6] * ph[0] + R[7] * ph[1] + R[8] * ph[2] + b.trans[2];
    out.dir[0] = R[0] * ph[3] + R[1] * ph[4] + R[2] * ph[5];
    out.dir[1] = R[3] * ph[3] + R[4] * ph[4] + R[5] * ph[5];
    out.dir[2] = R[6] * ph[3] + R[7] * ph[4] + R[8] * ph[5];
}

// ---------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void
dense_add(double* __restrict__ acc, unsigned cnb, double v, int accum_mode) {
    if (accum_mode == MQI_K_ACCUM_WARP_MATCH) {
        // warp-aggregated: lanes that hit the same voxel in this step are summed by one leader
        const unsigned active = __activemask();
        const unsigned peers  = __match_any_sync(active, cnb);
        const int      leader = __ffs(peers) - 1;
        const int      lane   = threadIdx.x & 31;
        if (peers != (1u << lane)) {
            double sum = 0.0;
            for (unsigned m = peers; m; m &= m - 1) {
                const int src = __ffs(m) - 1;
                sum += __shfl_sync(peers, v, src);
            }
            v = sum;
        }
        if (lane == leader) atomicAdd(acc + cnb, v);
    } else {
        atomicAdd(acc + cnb, v);   // RED.E.ADD.F64 (result unused)
    }
}

// open-addressing (voxel, spot) -> dose table with a single 64-bit CAS per claim.  Same hash
// function, home slot and linear probing as insert_hashtable (mqi_transport.hpp:68-111), so the set
// of occupied keys is identical; the reference's two independent 32-bit CAS (race B5) are replaced.
// A probe step reads the keys of several consecutive slots at once (independent loads) and takes the first one in probe
// order that is free or holds the key: the same slot linear probing ends in -- a slot seen occupied by another key stays so
// (keys are never removed), a slot seen free is claimed through the CAS, which reports whoever won it.  The first step of a
// sequence looks at kProbeWidth slots (one: more than half of the inserts end at their home slot even at load factor 0.73) ...
constexpr int kProbeWidth = MQI_K_PROBE_WIDTH;
// ... and the later steps of a probe sequence kProbeWidth2 slots: the lanes of a warp that insert in a turn wait for the longest
// probe sequence among them, and at load factor 0.73 that tail is long (mean 2.4 slots, but the longest of ten is ~ 12): the
// unlucky pairs that are still looking after the first step take wider steps, the common case pays one load.
constexpr int kProbeWidth2 = MQI_K_PROBE_WIDTH2;

// one step of the probe sequence: the keys of W consecutive slots from `slot` (wrapping), loaded together; true if the pair
// went into one of them
template<int W>
__device__ __forceinline__ bool
dij_probe_step(DijSlot* table, unsigned long long capacity, unsigned long long& slot, unsigned long long key, double v) {
    DijSlot*           e[W];
    unsigned long long kk[W];
    if (slot + W <= capacity) {   // no wrap-around inside this group (all but the last home slots)
#pragma unroll
        for (int i = 0; i < W; ++i) e[i] = table + slot + i;
        slot = slot + W == capacity ? 0 : slot + W;
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            e[i] = table + slot;
            slot = slot + 1 == capacity ? 0 : slot + 1;
        }
    }
    // L2 is the point of coherence of the table (the CAS and the adds are performed there): a cache-global
    // load sees every claimed key; a system-scope volatile load costs more and buys nothing
#pragma unroll
    for (int i = 0; i < W; ++i) kk[i] = __ldcg(&e[i]->key);
#pragma unroll
    for (int i = 0; i < W; ++i) {
        unsigned long long prev = kk[i];
        if (prev == kEmptyKey64) prev = atomicCAS(&e[i]->key, kEmptyKey64, key);
        if (prev == kEmptyKey64 || prev == key) {
            red_add_f64(&e[i]->value, v);
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ void
dij_probe_from(DijSlot* table, unsigned long long capacity, unsigned long long slot, unsigned long long probes,
               unsigned long long key, double v, unsigned long long* counters) {
    if (probes < capacity) {
        if (dij_probe_step<kProbeWidth>(table, capacity, slot, key, v)) return;
        probes += kProbeWidth;
    }
    for (; probes < capacity; probes += kProbeWidth2)
        if (dij_probe_step<kProbeWidth2>(table, capacity, slot, key, v)) return;
    atomicAdd(counters + C_DIJ_FULL, 1ull);   // the reference would spin forever here
}

// home slot of a (voxel, spot) pair; key2 = 0xffffffff is the reference's dense mode: slot = voxel, key2 := 0
// (mqi_transport.hpp:78-81).  Returns false if that slot lies outside the table.
__device__ __forceinline__ bool
dij_home(unsigned long long capacity, unsigned long long magic, uint32_t key1, uint32_t& key2, unsigned long long& slot) {
    if (key2 == kEmptyKey32) {
        slot = key1;
        key2 = 0;
        return slot < capacity;
    }
    slot = hash_fun_magic(key1, key2, capacity, magic);
    return true;
}

__device__ __forceinline__ void
dij_add_inline(DijSlot* table, unsigned long long capacity, unsigned long long magic, uint32_t key1, uint32_t key2, double v,
               unsigned long long* counters) {
    unsigned long long slot;
    if (!dij_home(capacity, magic, key1, key2, slot)) {   // a table smaller than the grid: count the hit instead of writing outside it
        atomicAdd(counters + C_DIJ_FULL, 1ull);
        return;
    }
    dij_probe_from(table, capacity, slot, 0ull, ((unsigned long long) key2 << 32) | key1, v, counters);
}

__device__ __noinline__ void
dij_add(const ScorerDev& S, uint32_t key1, uint32_t key2, double v, unsigned long long* counters) {
    dij_add_inline(S.table, S.capacity, S.cap_magic, key1, key2, v, counters);
}

struct StepResult {
    float dE;        // trk.dE of the step (incl. dE_pre)
    float local_dE;  // trk.local_dE
    float te_debug;  // delta-electron energy carried by the (folded) zero-energy daughter, debug only
    float len;       // |vtx1.pos - vtx0.pos|
};

// LETd_weight1/2: scorers/mqi_scorer_energy_deposit.hpp:93-137 (out of line: fp64 divisions)
__device__ __noinline__ double
letd_hit(int kind, float dE, float len, float rho) {
    const double let = (double) dE / (double) len / (double) (rho * 1000.0f);
    if (!(let < 25.0)) return 0.0;
    return kind == MQI_K_LETD_NUMER ? (double) dE * let : (double) dE;
}

// LETt_weight1/2: scorers/mqi_scorer_energy_deposit.hpp:141-177
__device__ __noinline__ double
lett_hit(int kind, float dE, float len, float rho) {
    if (kind == MQI_K_LETT_DENOM) return (double) len;
    const double let = (double) dE / (double) len / (double) (rho * 1000.0f);
    return (double) len * let;
}

// one insert per scorer per step, keyed to the voxel occupied at step start: mqi_transport.hpp:204-225,
// hit functions scorers/mqi_scorer_energy_deposit.hpp:14-137
// Write-combining in front of the hash table: a track makes two to three steps inside a CT voxel (the step is
// limited to 1 mm of water, the voxels are 1.5-2.5 mm), so consecutive hits of a lane mostly carry the same
// (voxel, spot) key.  The lane sums them in registers and inserts once when the voxel changes or the track ends
// (flush_dij in the re-arm prologue): the insert -- a dependent chain of uncoalesced L2/DRAM sector accesses --
// is what bounds the sparse scorer (profiles/r1_experiments.md).  Same keys, same home slots, same sums up to
// the order of the additions.
struct DijCombine {
    uint32_t key;   // voxel of the pending hit, kEmptyKey32 = nothing pending
    double   val;
};

// The finished pair of a lane -- its voxel changed, or its track ended -- is inserted by that lane at once.  Parking the
// pairs in per-lane shared-memory slots and letting the whole warp insert them together (round 2 experiment,
// profiles/r2_experiments.md) ran the probe code at 13 - 17 lanes instead of 5 but was slower: one insert per lane and
// turn keeps more table accesses in flight than a burst every sixth turn, and 49 kB of parking slots came out of L1.
__device__ __forceinline__ void
flush_dij(const Params& P, DijCombine& wc, uint32_t spot_ind) {
    if (wc.key != kEmptyKey32) {
        dij_add(P.sc[P.dij_wc_scorer], wc.key, spot_ind, wc.val, P.counters);
        wc.key = kEmptyKey32;
    }
}

// Scorer sets known at compile time: the loop over P.sc with its kind dispatch, quirk and accumulation-mode tests
// (~ 25 issue slots per scorer and step, and 30 spilled registers) is what the general kernel pays; the three
// scorer lists the configs actually use get the same hit values from straight-line code.
//   SET_DOSE       one dense Dose scorer, DIRECT roi                       (phantom_env; tps "Dose")
//   SET_DOSE_LETD  Dose, LETd numerator, LETd denominator, DIRECT rois     (config 2; tps "Dose LETd")
//   SET_DOSE_STAT  Dose, Dose (stat), Dose^2 (stat), any rois              (config 3: tps "Dose" with StoppingStatistics)
//   SET_DIJ        one Dij scorer with write-combining (DIJWC)             (config 4: tps "Dij")
enum ScorerSet { SET_GENERIC = 0, SET_DOSE = 1, SET_DOSE_LETD = 2, SET_DOSE_STAT = 3, SET_DIJ = 4 };

// LETd_weight1 and LETd_weight2 of one step from one evaluation of the LET (same arithmetic as letd_hit)
__device__ __noinline__ double
letd_numer_hit(float dE, float len, float rho) {
    const double let = (double) dE / (double) len / (double) (rho * 1000.0f);
    if (!(let < 25.0)) return -1.0;   // neither scorer takes the step
    return (double) dE * let;
}

// roi_->idx(cnb) > 0 (mqi_transport.hpp:205,216): a DIRECT roi returns cnb, so voxel 0 is never scored (B1); a
// CONTOUR roi (mask_to_roi) accepts the voxels inside its runs
__device__ __forceinline__ bool
roi_accepts(const uint32_t* roi, unsigned cnb) {
    return roi ? ((__ldg(roi + (cnb >> 5)) >> (cnb & 31u)) & 1u) != 0u : (int) cnb > 0;
}

template<int VARIANT, int SET, bool DIJWC>
__device__ __forceinline__ void
score_step(const Params& P, const Smem& sm, const MatEntry& M, unsigned cnb, uint32_t spot_ind, float inv_vol, float rsp0,
           const StepResult& r, DijCombine& wc) {
    // dose_to_water: (dE + local_dE) * 1.60218e-10 / (V * rho * rsp(rho, vtx0.ke))
    const float  kdose   = 1.60218e-10f * inv_vol * M.inv_rho;
    const double dose    = (double) ((r.dE + r.local_dE) * kdose / rsp0);
    const double dose_te = (VARIANT == MQI_K_DEBUG) ? (double) (r.te_debug * kdose * inv_rsp_at_zero_energy(M)) : 0.0;
    if (SET == SET_DIJ) {
        const double v = dose + dose_te;
        if (!(v > 0.0) || !roi_accepts(P.sc[0].roi, cnb)) return;
        if (wc.key == cnb) {
            wc.val += v;
        } else {
            flush_dij(P, wc, spot_ind);
            wc.key = cnb;
            wc.val = v;
        }
        return;
    }
    if (SET == SET_DOSE_LETD) {
        if ((int) cnb <= 0) return;
        const double v = dose + dose_te;
        if (v > 0.0) atomicAdd(P.sc[0].dense + cnb, v);
        if (r.len > 0.f) {
            const double numer = letd_numer_hit(r.dE, r.len, M.rho);
            if (numer > 0.0) atomicAdd(P.sc[1].dense + cnb, numer);            // insert_hashtable: value <= 0 -> skip
            if (numer >= 0.0 && r.dE > 0.f) atomicAdd(P.sc[2].dense + cnb, (double) r.dE);
        }
        return;
    }
    if (SET == SET_DOSE_STAT) {
        const double v = dose + dose_te;
        if (!(v > 0.0)) return;   // then the square is not positive either
        const double v2 = dose * dose + dose_te * dose_te;
        if (!P.sc[0].roi && !P.sc[1].roi && !P.sc[2].roi) {   // three DIRECT rois (the common case): one test
            if ((int) cnb > 0) {
                atomicAdd(P.sc[0].dense + cnb, v);
                atomicAdd(P.sc[1].dense + cnb, v);
                atomicAdd(P.sc[2].dense + cnb, v2);
            }
            return;
        }
        if (roi_accepts(P.sc[0].roi, cnb)) atomicAdd(P.sc[0].dense + cnb, v);
        if (roi_accepts(P.sc[1].roi, cnb)) atomicAdd(P.sc[1].dense + cnb, v);
        if (roi_accepts(P.sc[2].roi, cnb)) atomicAdd(P.sc[2].dense + cnb, v2);
        return;
    }
    const int    n       = P.n_scorers;
#pragma unroll 1
    for (int s = 0; s < n; ++s) {
        const int kind = P.sc[s].kind;
        if (!roi_accepts(P.sc[s].roi, cnb)) continue;
        double    v    = 0.0;
        if (kind == MQI_K_DOSE || kind == MQI_K_DIJ) v = dose + dose_te;
        else if (kind == MQI_K_DOSE_SQ) v = dose * dose + dose_te * dose_te;
        else if (kind == MQI_K_EDEP) v = (double) (r.dE + r.local_dE) + (double) r.te_debug;
        else if (kind == MQI_K_LETT_NUMER || kind == MQI_K_LETT_DENOM) { if (r.len > 0.f) v = lett_hit(kind, r.dE, r.len, M.rho); }
        else if (r.len > 0.f) v = letd_hit(kind, r.dE, r.len, M.rho);
        if (!(v > 0.0)) continue;   // insert_hashtable: value <= 0 -> skip
        // quirk B2: the reference's non-stat kernel scores scorers [0, n-2) twice when n >= 3
        if ((P.quirks & MQI_K_QUIRK_B2) && s < n - 2) v += v;
        if (kind == MQI_K_DIJ) {
            if (DIJWC && s == P.dij_wc_scorer) {
                if (wc.key == cnb) {
                    wc.val += v;
                } else {
                    flush_dij(P, wc, spot_ind);
                    wc.key = cnb;
                    wc.val = v;
                }
            } else {
                dij_add(P.sc[s], cnb, spot_ind, v, P.counters);
            }
        } else {
            dense_add(P.sc[s].dense, cnb, v, P.accum_mode);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// starting a track: world -> node frame, locate the start cell (index(p, dir) or entry intersect),
// mqi_transport.hpp:162-190.  Out of the voxel-step loop (keeps it compact for the instruction cache).
// Primaries are started 32 at a time by the whole warp (refill_queue) and parked in a per-warp
// shared-memory queue; secondaries and node-to-node transitions are started by the lane that owns
// them (restart_lane).
// ---------------------------------------------------------------------------------------------
struct TrackIO {
    float px, py, pz, dx, dy, dz, ke;
    int   ix, iy, iz;
    int   recoil;   // debug recoil daughter before its first step: vtx1.ke = 0 and trk.dE = vtx0.ke
    int   node;     // in: first child to try, out: child entered
    int   sp;       // stack pointer (restart_lane)
};

// Try the world's children in order from T.node (the reference's c_ind loop): a child the track misses
// hands it back in the world frame to the next one, :176-185.  Returns false if no child is entered.
template<bool MULTI>
__device__ __forceinline__ bool
enter_nodes(const Params& P, const Smem& sm, TrackIO& T) {
    const int n_nodes = MULTI ? P.n_nodes : 1;
    float     px = T.px, py = T.py, pz = T.pz, dx = T.dx, dy = T.dy, dz = T.dz, ke = T.ke;
    bool      recoil = T.recoil != 0;
#pragma unroll 1
    for (int node = MULTI ? T.node : 0; node < n_nodes; ++node) {
        const GridDev& G  = node_ref<MULTI>(P, sm, node);
        const int      nx = G.nx, ny = G.ny, nz = G.nz;
        const float*   xe = sm.edges + (MULTI ? G.edge_off : 0);
        const float*   ye = xe + nx + 1;
        const float*   ze = ye + ny + 1;
        // world -> node frame, mqi_transport.hpp:165-170
        if (!G.identity) {
            const float* R  = G.rot_fwd;   // inverse = transpose
            const float  qx = px - G.trans[0], qy = py - G.trans[1], qz = pz - G.trans[2];
            px = R[0] * qx + R[3] * qy + R[6] * qz;
            py = R[1] * qx + R[4] * qy + R[7] * qz;
            pz = R[2] * qx + R[5] * qy + R[8] * qz;
            const float ex = dx, ey = dy, ez = dz;
            dx = R[0] * ex + R[3] * ey + R[6] * ez;
            dy = R[1] * ex + R[4] * ey + R[7] * ez;
            dz = R[2] * ex + R[5] * ey + R[8] * ez;
        }
        {
            const float n = rsqrtf(dx * dx + dy * dy + dz * dz);
            dx *= n; dy *= n; dz *= n;
        }
        // locate: index(p, dir) or entry intersect, :171-190.  A point farther than the geometry
        // tolerance outside the bounding box has no valid index on that axis: skip the search.
        int ix, iy, iz;
        if (px < xe[0] - 2.f * kGeomTol || px > xe[nx] + 2.f * kGeomTol || py < ye[0] - 2.f * kGeomTol ||
            py > ye[ny] + 2.f * kGeomTol || pz < ze[0] - 2.f * kGeomTol || pz > ze[nz] + 2.f * kGeomTol) {
            ix = iy = iz = -1;
        } else {
            ix = index_axis_fast(xe, nx, px, dx, G.inv_w[0]);
            iy = index_axis_fast(ye, ny, py, dy, G.inv_w[1]);
            iz = index_axis_fast(ze, nz, pz, dz, G.inv_w[2]);
        }
        bool  alive = true;
        float d[3]  = { dx, dy, dz };
        if (ix < 0 || iy < 0 || iz < 0 || ix >= nx || iy >= ny || iz >= nz) {
            const float p[3] = { px, py, pz };
            int         c[3];
            const float inv_w[3] = { G.inv_w[0], G.inv_w[1], G.inv_w[2] };
            const float dist = grid_entry(xe, ye, ze, nx, ny, nz, inv_w, p, d, c);
            if (dist < 0.f) {
                alive = false;
            } else {
                // update_post_vertex_position uses the (possibly zeroed) direction, move() then
                // restores the un-zeroed copy held in vtx1.dir; vtx1.ke becomes vtx0.ke.  The cell
                // of the moved point is the one intersect() already looked up (same point, same d
                // up to the zeroed components, which only matter exactly on an edge).
                px = __fadd_rn(px, __fmul_rn(d[0], dist));
                py = __fadd_rn(py, __fmul_rn(d[1], dist));
                pz = __fadd_rn(pz, __fmul_rn(d[2], dist));
                if (recoil) ke = 0.f;   // vtx0.ke := vtx1.ke, and the carried energy is dropped by move()
                recoil = false;
                if (d[0] == dx && d[1] == dy && d[2] == dz) {
                    ix = c[0]; iy = c[1]; iz = c[2];
                } else {
                    ix = index_axis_guess(xe, nx, px, dx, G.inv_w[0]);
                    iy = index_axis_guess(ye, ny, py, dy, G.inv_w[1]);
                    iz = index_axis_guess(ze, nz, pz, dz, G.inv_w[2]);
                }
                if (ix < 0 || iy < 0 || iz < 0 || ix >= nx || iy >= ny || iz >= nz) alive = false;
            }
        }
        if (alive) {
            T.px = px; T.py = py; T.pz = pz; T.dx = dx; T.dy = dy; T.dz = dz;
            T.ke = ke; T.recoil = recoil ? 1 : 0;
            T.ix = ix; T.iy = iy; T.iz = iz;
            T.node = node;
            return true;
        }
        if (!MULTI) break;
        // missed this child: node -> world frame with the direction intersect() left behind (tiny
        // components zeroed in place), :176-185, and on to the next child
        dx = d[0]; dy = d[1]; dz = d[2];
        if (!G.identity) {
            const float* R  = G.rot_fwd;
            const float  qx = px, qy = py, qz = pz;
            px = R[0] * qx + R[1] * qy + R[2] * qz + G.trans[0];
            py = R[3] * qx + R[4] * qy + R[5] * qz + G.trans[1];
            pz = R[6] * qx + R[7] * qy + R[8] * qz + G.trans[2];
            const float ex = dx, ey = dy, ez = dz;
            dx = R[0] * ex + R[1] * ey + R[2] * ez;
            dy = R[3] * ex + R[4] * ey + R[5] * ez;
            dz = R[6] * ex + R[7] * ey + R[8] * ez;
        }
    }
    return false;
}

// A lane whose track ended continues with (a) the same track in the next child of the world, if it
// left its node alive (multi-node launches: local -> world, mqi_transport.hpp:234-239, then the c_ind
// loop goes on), or (b) the secondary on top of its stack, which starts at child 0 again (:160-162).
// a track that left child G alive: child frame -> world frame, mqi_transport.hpp:234-239
__device__ __forceinline__ void
node_to_world(const GridDev& G, TrackIO& T) {
    if (!G.identity) {
        const float* R  = G.rot_fwd;
        const float  qx = T.px, qy = T.py, qz = T.pz;
        T.px = R[0] * qx + R[1] * qy + R[2] * qz + G.trans[0];
        T.py = R[3] * qx + R[4] * qy + R[5] * qz + G.trans[1];
        T.pz = R[6] * qx + R[7] * qy + R[8] * qz + G.trans[2];
        const float ex = T.dx, ey = T.dy, ez = T.dz;
        T.dx = R[0] * ex + R[1] * ey + R[2] * ez;
        T.dy = R[3] * ex + R[4] * ey + R[5] * ez;
        T.dz = R[6] * ex + R[7] * ey + R[8] * ez;
    }
}

template<bool MULTI>
__device__ __noinline__ bool
restart_lane(const Params& P, const Secondary* __restrict__ stack, TrackIO& __restrict__ T, int advance) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(smem_raw, P.n_edge_floats, P.n_nodes);
    if (MULTI && advance) {
        node_to_world(sm.nodes[T.node], T);
        T.node += 1;   // the caller only advances when a next child exists
    } else {
        const Secondary s = stack[--T.sp];   // by value: the nine loads are issued back to back, then the stores
        T.px = s.px; T.py = s.py; T.pz = s.pz; T.dx = s.dx; T.dy = s.dy; T.dz = s.dz;
        T.ke     = s.ke0;
        T.recoil = s.ke1_off != 0.f;   // pushed as (ke0, -ke0, ke0) by the debug variant only
        T.node   = 0;
    }
    return enter_nodes<MULTI>(P, sm, T);
}

// one track into a warp's hand-over buffer (out of line: twelve scattered stores that the step loop should not hold
// registers for)
__device__ MQI_K_ADV_PUSH_INLINE void
push_handover(uint32_t* w, float px, float py, float pz, float dx, float dy, float dz, float ke, uint32_t h0, uint32_t h1, uint32_t spot,
              uint32_t node, uint32_t blk) {
    __stcg(w + R_PX * kRawCap, __float_as_uint(px)); __stcg(w + R_PY * kRawCap, __float_as_uint(py));
    __stcg(w + R_PZ * kRawCap, __float_as_uint(pz)); __stcg(w + R_DX * kRawCap, __float_as_uint(dx));
    __stcg(w + R_DY * kRawCap, __float_as_uint(dy)); __stcg(w + R_DZ * kRawCap, __float_as_uint(dz));
    __stcg(w + R_KE * kRawCap, __float_as_uint(ke));
    __stcg(w + R_H0 * kRawCap, h0); __stcg(w + R_H1 * kRawCap, h1);
    __stcg(w + R_SPOT * kRawCap, spot); __stcg(w + R_NODE * kRawCap, node);
    __stcg(w + R_BLK * kRawCap, blk);
}

// The whole warp takes the top (up to 32) tracks of its hand-over buffer: every lane maps one from the child it left to
// the world frame (mqi_transport.hpp:234-239), offers it to the following children (enter_nodes, the c_ind loop) and
// appends it to the warp's EMPTY queue if it enters one.  What restart_lane's advance branch does for one lane at a
// time, at full SIMT width.  Returns the number of queue entries written.
__device__ __noinline__ int
process_handovers(const Params& P, uint32_t* q, const uint32_t* raw, int n_raw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm   = smem_view(smem_raw, P.n_edge_floats, P.n_nodes);
    const int  lane = threadIdx.x & 31;
    const int  e    = n_raw - 1 - lane;
    bool       alive = false;
    TrackIO    T;
    uint32_t   h0 = 0, h1 = 0, spot = 0, blk = 0;
    if (e >= 0) {
        const uint32_t* r = raw + e;
        T.px = __uint_as_float(__ldcg(r + R_PX * kRawCap)); T.py = __uint_as_float(__ldcg(r + R_PY * kRawCap));
        T.pz = __uint_as_float(__ldcg(r + R_PZ * kRawCap)); T.dx = __uint_as_float(__ldcg(r + R_DX * kRawCap));
        T.dy = __uint_as_float(__ldcg(r + R_DY * kRawCap)); T.dz = __uint_as_float(__ldcg(r + R_DZ * kRawCap));
        T.ke = __uint_as_float(__ldcg(r + R_KE * kRawCap));
        h0 = __ldcg(r + R_H0 * kRawCap); h1 = __ldcg(r + R_H1 * kRawCap);
        spot = __ldcg(r + R_SPOT * kRawCap); blk = __ldcg(r + R_BLK * kRawCap);
        T.node   = (int) __ldcg(r + R_NODE * kRawCap);
        T.recoil = 0;
        node_to_world(sm.nodes[T.node], T);
        T.node += 1;   // a track is only handed over when a next child exists
        alive = enter_nodes<true>(P, sm, T);
    }
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    if (alive) {
        uint32_t* w = q + __popc(m & ((1u << lane) - 1u));
        w[Q_PX * kQueueCap] = __float_as_uint(T.px); w[Q_PY * kQueueCap] = __float_as_uint(T.py);
        w[Q_PZ * kQueueCap] = __float_as_uint(T.pz); w[Q_DX * kQueueCap] = __float_as_uint(T.dx);
        w[Q_DY * kQueueCap] = __float_as_uint(T.dy); w[Q_DZ * kQueueCap] = __float_as_uint(T.dz);
        w[Q_KE * kQueueCap] = __float_as_uint(T.ke);
        w[Q_IX * kQueueCap] = (uint32_t) T.ix; w[Q_IY * kQueueCap] = (uint32_t) T.iy; w[Q_IZ * kQueueCap] = (uint32_t) T.iz;
        w[Q_H0 * kQueueCap] = h0; w[Q_H1 * kQueueCap] = h1;
        w[Q_SPOT * kQueueCap] = spot;
        w[Q_NODE * kQueueCap] = (uint32_t) T.node;
        w[Q_BLK * kQueueCap]  = blk;
    }
    __syncwarp();
    return __popc(m);
}

// The whole warp fetches the next 32 history ids with one atomic, samples (or loads) their primary
// vertices, locates them and appends the ones that enter the geometry to the warp's queue.  Every lane
// works on its own history: the source sampling and the cell search, which cost about as much as one
// voxel step, run at full SIMT width instead of once per lane and history.  Returns the new queue
// length | source exhausted << 16.
template<bool MULTI>
__device__ __noinline__ int
refill_queue(const Params& P, uint32_t* q, int q_n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm   = smem_view(smem_raw, P.n_edge_floats, P.n_nodes);
    const int  lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.counters + C_NEXT, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    // this launch's k-th fetch is chunk k * n_shards + shard of the range (n_shards = 1: chunk k) -- or, with P.reverse, the
    // shard's chunks from the last to the first: a plan that lists its energy layers in ascending order then starts its
    // longest histories first and its shortest last, which is what the tail of a persistent launch wants
    bool exhausted;
    if (P.reverse) {
        const unsigned long long n_chunks = (P.count + 31ull) >> 5;
        const unsigned long long mine     = n_chunks > P.shard ? (n_chunks - P.shard + P.n_shards - 1ull) / P.n_shards : 0ull;   // chunks of this shard
        const unsigned long long k        = base >> 5;
        exhausted = k + 1ull >= mine;
        base      = k < mine ? ((mine - 1ull - k) * P.n_shards + P.shard) << 5 : P.count;
    } else {
        base      = (base * P.n_shards) + 32ull * P.shard;
        exhausted = base + 32ull * P.n_shards >= P.count;   // the shard's next chunk lies beyond the range
    }
    const unsigned long long i    = base + lane;
    const bool               have = i < P.count;
    bool     alive = false;
    TrackIO  T;
    uint32_t spot = 0, h0 = 0, h1 = 0;
    if (have) {
        const unsigned long long h = P.first + i;
        h0 = (uint32_t) h;
        h1 = (uint32_t) (h >> 32);
        VertexDev v;
        if (P.src.vertices) {
            v    = P.src.vertices[i];
            spot = P.src.spot_ids ? P.src.spot_ids[i] : 0u;
        } else {
            // beamsource::operator()(h): first spot whose cumulative count exceeds h
            uint32_t lo = 0, hi = P.src.n_spots;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (P.src.cum[mid] > h) hi = mid; else lo = mid + 1;
            }
            spot = min(lo, P.src.n_spots - 1);
            sample_vertex(P.src.beamlets[spot], P.seed, h, v);
        }
        T.px = v.pos[0]; T.py = v.pos[1]; T.pz = v.pos[2];
        T.dx = v.dir[0]; T.dy = v.dir[1]; T.dz = v.dir[2];
        T.ke = v.ke;
        T.recoil = 0;
        T.node   = 0;
        alive    = enter_nodes<MULTI>(P, sm, T);
    }
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    const unsigned fetched = __ballot_sync(0xffffffffu, have);
    if (lane == 0 && fetched) atomicAdd(P.counters + C_DONE, (unsigned long long) __popc(fetched));   // :245 tracked_particles
    if (alive) {
        uint32_t* e = q + q_n + __popc(m & ((1u << lane) - 1u));
        e[Q_PX * kQueueCap] = __float_as_uint(T.px); e[Q_PY * kQueueCap] = __float_as_uint(T.py);
        e[Q_PZ * kQueueCap] = __float_as_uint(T.pz); e[Q_DX * kQueueCap] = __float_as_uint(T.dx);
        e[Q_DY * kQueueCap] = __float_as_uint(T.dy); e[Q_DZ * kQueueCap] = __float_as_uint(T.dz);
        e[Q_KE * kQueueCap] = __float_as_uint(T.ke);
        e[Q_IX * kQueueCap] = (uint32_t) T.ix; e[Q_IY * kQueueCap] = (uint32_t) T.iy; e[Q_IZ * kQueueCap] = (uint32_t) T.iz;
        e[Q_H0 * kQueueCap] = h0; e[Q_H1 * kQueueCap] = h1;
        e[Q_SPOT * kQueueCap] = P.per_spot ? spot : kEmptyKey32;
        e[Q_NODE * kQueueCap] = (uint32_t) T.node;
        if (MULTI) e[Q_BLK * kQueueCap] = P.src.vertices ? 0u : 2u;   // blocks 0-1 belong to the source sampling
    }
    __syncwarp();
    return (q_n + __popc(m)) | ((exhausted ? 1 : 0) << 16);
}

// further tries of the delta-electron energy rejection loop (about one event in ten needs them):
// (n, accept) pairs from Philox2x32-10, counter = (block, history_lo), one block number per pair
// returns {Te, next block number} in registers (a reference parameter would pin the lane's block
// counter to local memory for the whole loop)
__device__ __noinline__ float2
delta_retry(uint32_t blk, uint32_t h0, uint32_t key2, float T_cut, float Tmax1, float kb, float inv_2Et_sq, float g_max) {
    const float dT = T_cut - Tmax1, num = T_cut * Tmax1;   // kb = b1^2 / Tmax1, g_max = g(T_cut): see the first try
    while (true) {
        uint32_t wn, wa;
        philox2x32_10(blk, h0, key2, wn, wa);
        blk += 1;
        const float Te = num * rcp_fast(fmaf(u32_to_uniform(wn), dT, Tmax1));
        if (u32_to_uniform(wa) * g_max < fmaf(Te * Te, inv_2Et_sq, fmaf(-kb, Te, 1.0f))) return make_float2(Te, __uint_as_float(blk));
    }
}

// do { if (r >= r_steps[n]) break; } while (--n > 0) of the reference's range-table inversion (mqi_p_ionization.hpp:
// 298-420), started from a guessed row: the largest row n <= n0 whose range does not exceed r.  Out of line: the
// guess is right for nearly every step.
__device__ __noinline__ int
csda_row_fix(const float4* a1, int n, int n0, float r) {
    float bx = a1[n].x;
    while (n < n0) {
        const float nx = a1[n + 1].x;
        if (r < nx) break;
        bx = nx;
        ++n;
    }
    while (n > 0 && r < bx) bx = a1[--n].x;
    return n;
}

// ---------------------------------------------------------------------------------------------
// the transport kernel
// ---------------------------------------------------------------------------------------------
// SET (ScorerSet): SET_DOSE -- exactly one dense Dose scorer with a DIRECT roi (phantom_env, and the tps "Dose"
// case) -- compiles the scorer loop and the other hit functions out of the voxel-step loop; SET_DOSE_LETD and
// SET_DOSE_STAT replace the loop by straight-line code; SET_GENERIC is the loop over P.sc.
// MULTI: the world has beamline children (range shifter, aperture) in front of the scored grid; every
// lane carries the index of the child it is in and reads that child's descriptor from shared memory.
template<int VARIANT, int SET, bool MULTI, bool DIJWC = false>
__global__ void __launch_bounds__(MULTI ? MQI_K_BLOCK_MULTI : (SET == SET_DIJ && DIJWC ? MQI_K_BLOCK_DIJ : MQI_K_BLOCK), MQI_K_MIN_BLOCKS)
transport_kernel(const __grid_constant__ Params P) {
    constexpr bool SIMPLE  = SET == SET_DOSE;
    constexpr bool COUNTED = SET == SET_GENERIC;   // count_steps runs the general kernel
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = smem_view(smem_raw, P.n_edge_floats, P.n_nodes);
    {
        float4* s_a0    = reinterpret_cast<float4*>(smem_raw);
        float4* s_a1    = s_a0 + kTableN;
        float2* s_bs    = reinterpret_cast<float2*>(s_a1 + kTableN);
        float*  s_edges = reinterpret_cast<float*>(s_bs + kTableN);
        for (int i = threadIdx.x; i < kTableN; i += blockDim.x) {
            s_a0[i] = P.tab_a0[i];
            s_a1[i] = P.tab_a1[i];
            s_bs[i] = P.tab_bs[i];
        }
        for (int i = threadIdx.x; i < P.n_edge_floats; i += blockDim.x) s_edges[i] = P.edges_all[i];
        if (MULTI) {
            uint32_t*       dst = reinterpret_cast<uint32_t*>(smem_raw + smem_nodes_offset(P.n_edge_floats));
            const uint32_t* src = reinterpret_cast<const uint32_t*>(P.nodes);
            for (int i = threadIdx.x; i < P.n_nodes * (int) (sizeof(GridDev) / 4); i += blockDim.x) dst[i] = src[i];
        }
    }
    __syncthreads();

    constexpr float T_cut = (VARIANT == MQI_K_DEBUG) ? 0.08511f : 0.0815f;   // mqi_interaction.hpp:24-28
    constexpr int   DEPTH = StackCfg<VARIANT>::depth;
    Secondary stack[DEPTH];
    int       sp = 0;

    // lane state
    float    px = 0, py = 0, pz = 0, dx = 0, dy = 0, dz = 0, ke = 0;
    // lane flags in ONE register (separate bools end up as predicates, which every out-of-line call
    // saves and restores through local memory): FL_RECOIL see TrackIO::recoil (debug variant only),
    // FL_ADVANCE the track left `node` alive and a next child exists (MULTI), FL_DONE this lane found the
    // queue empty and the history counter exhausted
    constexpr unsigned FL_ALIVE = 1u, FL_DONE = 2u, FL_RECOIL = 4u, FL_ADVANCE = 8u;
    unsigned fl = 0u;
    int      ix = 0, iy = 0, iz = 0;
    int      node = 0;          // child of the world the lane's track is in (MULTI)
    uint32_t spot_ind = kEmptyKey32;
    uint32_t h0 = 0, h1 = 0, blk = 0;   // Philox counter of the current history
    const uint32_t k0 = (uint32_t) P.seed, k1 = (uint32_t) (P.seed >> 32);
    unsigned n_steps = 0;
    DijCombine wc;   // DIJWC instantiations only (a Dij scorer with write-combining): costs three registers
    wc.key = kEmptyKey32; wc.val = 0.0;

    // the warp's queue of pre-sampled primaries: entries in the queue | history counter exhausted << 16
    // (warp-uniform; one register)
    int q_state = 0;
#if MQI_K_ADV_QUEUE
    int n_raw = 0;      // MULTI: tracks in the warp's hand-over buffer (warp-uniform)
#else
    int adv_wait = 0;   // MULTI: turns the oldest lane waiting for a node hand-over has waited (warp-uniform)
    constexpr int kAdvBatch = MQI_K_ADV_BATCH, kAdvTurns = MQI_K_ADV_TURNS;
#endif

    // Warp-level reconvergence.  Restarting a lane is executed by the few lanes whose track just ended;
    // without an explicit join the compiler only reconverges them at the END of the iteration, i.e. the
    // whole step body ran twice per restart (once for the restarted lane alone: 13 % of all issue slots in
    // ncu).  The full-warp votes below are the join: every lane executes them once per turn, and the
    // warp leaves the loop together once all of its lanes found the source exhausted.
    while (true) {
        // One vote is the join of the turn and the fast path in one: while every lane of the warp owns a track
        // (86 % of the turns at C1) the re-arm prologue -- two more votes and the tests around them, ~20 issue
        // slots -- is skipped altogether.
        if (!__all_sync(0xffffffffu, fl & FL_ALIVE)) {
        // ------------------------------------------------------------------ restart the lane
        // Node-to-node hand-overs (MULTI) are batched: the tracks of a warp were started together, so they leave a
        // beamline child within a few turns of each other.  A lane whose track has left its child waits (idle) until
        // kAdvBatch lanes of the warp wait or the first of them has waited kAdvTurns turns; then all of them are
        // mapped to the world frame and located in the next child in ONE pass of restart_lane.  One lane at a time
        // (the whole warp waiting for ~ 300 instructions of a single lane, twice per history) this was 27 % of the
        // issued warp instructions at 1.5 active lanes (profiles/r2_experiments.md).
        bool hand_over = true;
#if MQI_K_ADV_QUEUE
        // Hand-over queue: a track that left a beamline child alive does not have to go on in the lane that brought it
        // there.  The lane writes it -- position and direction in the child's frame, energy, history id, Philox block --
        // to the warp's hand-over buffer and takes the next located track from the warp's queue (a primary or an earlier
        // hand-over); when the queue runs empty the whole warp maps 32 buffered tracks to the world frame and locates
        // them in their next child at once (process_handovers).  The stream protocol is untouched: a history's blocks
        // are consumed in the same order by whichever lane owns the track.  A lane that still holds secondaries of the
        // history (sp > 0; they follow the primary in the history's block sequence) keeps the track: restart_lane below.
        uint32_t* const raw = MULTI ? P.adv_raw + ((size_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kRawWords : nullptr;
        if (MULTI) {
            const bool     push = (fl & FL_ADVANCE) != 0u && sp == 0;
            const unsigned pm   = __ballot_sync(0xffffffffu, push);
            if (pm) {
                const int r = n_raw + __popc(pm & ((1u << (threadIdx.x & 31)) - 1u));
                if (push && r < kRawCap) {
                    if (DIJWC) flush_dij(P, wc, spot_ind);
                    push_handover(raw + r, px, py, pz, dx, dy, dz, ke, h0, h1, spot_ind, (uint32_t) node, blk);
                    fl = 0u;   // the lane is free and takes the next queue entry below
                }
                n_raw = min(n_raw + __popc(pm), kRawCap);   // lanes the buffer has no room for hand their track over themselves
                __syncwarp();
            }
        }
#else
        if (MULTI) {
            const unsigned adv = __ballot_sync(0xffffffffu, (fl & FL_ADVANCE) != 0u);
            if (adv) {
                adv_wait += 1;
                hand_over = __popc(adv) >= kAdvBatch || adv_wait > kAdvTurns;
                if (hand_over) adv_wait = 0;
            }
        }
#endif
        bool need = false;   // the lane needs a new primary
        if (!(fl & (FL_ALIVE | FL_DONE)) && !(MULTI && (fl & FL_ADVANCE) && !hand_over)) {
            if (DIJWC) flush_dij(P, wc, spot_ind);   // the track ended: insert its pending write-combined Dij hit
            if ((MULTI && (fl & FL_ADVANCE)) || sp > 0) {
                TrackIO T;
                T.px = px; T.py = py; T.pz = pz; T.dx = dx; T.dy = dy; T.dz = dz; T.ke = ke;
                T.recoil = 0; T.node = node; T.sp = sp;
                const bool ok = restart_lane<MULTI>(P, stack, T, MULTI && (fl & FL_ADVANCE) ? 1 : 0);
                sp = T.sp;
                fl = 0u;
                if (ok) {
                    fl = FL_ALIVE | ((VARIANT == MQI_K_DEBUG && T.recoil != 0) ? FL_RECOIL : 0u);
                    px = T.px; py = T.py; pz = T.pz; dx = T.dx; dy = T.dy; dz = T.dz;
                    ke = T.ke;
                    ix = T.ix; iy = T.iy; iz = T.iz;
                    if (MULTI) node = T.node;
                }   // else: the track never enters the geometry, try again next turn
            } else {
                need = true;
            }
        }
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
        if (need_mask) {
            const int n_need = __popc(need_mask);
            const int lane   = threadIdx.x & 31;
            uint32_t* q      = sm.queue + (threadIdx.x >> 5) * queue_words(MULTI);
            // warp-uniform: the queue is empty and the source is not exhausted -> every lane helps to refill.
            // Lanes a nearly empty queue cannot serve this turn idle for one pass and are served next turn.
#if MQI_K_ADV_QUEUE
            if (MULTI && (q_state & 0xffff) == 0 && n_raw > 0 && (n_raw >= MQI_K_ADV_MIN || (q_state >> 16) != 0)) {
                // buffered hand-overs first (a full warp's worth, or whatever is left once the source is exhausted)
                q_state = process_handovers(P, q, raw, n_raw) | (q_state & 0x10000);
                n_raw   = max(n_raw - 32, 0);
            } else
#endif
            if (q_state == 0) q_state = refill_queue<MULTI>(P, q, 0);
            const int  q_n       = q_state & 0xffff;
#if MQI_K_ADV_QUEUE
            const bool src_empty = (q_state >> 16) != 0 && !(MULTI && n_raw > 0);   // buffered hand-overs are work to come
#else
            const bool src_empty = (q_state >> 16) != 0;
#endif
            if (need) {
                const int e = q_n - 1 - __popc(need_mask & ((1u << lane) - 1u));
                if (e >= 0) {
                    const uint32_t* qe = q + e;
                    px = __uint_as_float(qe[Q_PX * kQueueCap]); py = __uint_as_float(qe[Q_PY * kQueueCap]);
                    pz = __uint_as_float(qe[Q_PZ * kQueueCap]); dx = __uint_as_float(qe[Q_DX * kQueueCap]);
                    dy = __uint_as_float(qe[Q_DY * kQueueCap]); dz = __uint_as_float(qe[Q_DZ * kQueueCap]);
                    ke = __uint_as_float(qe[Q_KE * kQueueCap]);
                    ix = (int) qe[Q_IX * kQueueCap]; iy = (int) qe[Q_IY * kQueueCap]; iz = (int) qe[Q_IZ * kQueueCap];
                    h0 = qe[Q_H0 * kQueueCap]; h1 = qe[Q_H1 * kQueueCap];
                    spot_ind = qe[Q_SPOT * kQueueCap];
                    if (MULTI) node = (int) qe[Q_NODE * kQueueCap];
                    if (MULTI) blk = qe[Q_BLK * kQueueCap];      // a primary's first block, or where a handed-over track goes on
                    else blk = P.src.vertices ? 0u : 2u;         // blocks 0-1 belong to the source sampling
                    fl = FL_ALIVE;
                } else if (src_empty) {
                    fl = FL_DONE;
                }
            }
            __syncwarp();   // the queue slots just read may be overwritten by the next refill
            q_state = max(q_n - n_need, 0) | (q_state & 0x10000);
        }
        if (__all_sync(0xffffffffu, fl & FL_DONE)) break;
        if (!(fl & FL_ALIVE)) continue;   // taken after the join: the lane idles this turn, the others are converged
        }

        // ------------------------------------------------------------------ one voxel step
        // the option "count_steps" runs the general kernel: the counter is a spilled register (LDL + IADD + STL
        // per step) that the single-Dose-scorer kernel does not pay for
        if (COUNTED) ++n_steps;
        const GridDev& G  = node_ref<MULTI>(P, sm, node);
        const int      nx = G.nx, ny = G.ny, nz = G.nz;
        const float*   xe = sm.edges + (MULTI ? G.edge_off : 0);
        const float*   ye = xe + nx + 1;
        const float*   ze = ye + ny + 1;
        const float ex0 = xe[ix], ex1 = xe[ix + 1];
        const float ey0 = ye[iy], ey1 = ye[iy + 1];
        const float ez0 = ze[iz], ez1 = ze[iz + 1];
        const unsigned cnb = ((unsigned) iz * (unsigned) ny + (unsigned) iy) * (unsigned) nx + (unsigned) ix;
        // material of the voxel: two dependent loads (index volume -> LUT entry); everything up to the
        // first use of M (random numbers, kinematics, table rows, voxel exit distance) depends on the
        // lane state alone
#if MQI_K_LATE_LUT
        const unsigned mat_idx = __ldg(G.mat + cnb);
#else
        const MatEntry M = G.lut[__ldg(G.mat + cnb)];
#endif

        // the per-step Philox block {u_mfp, u_a, u_b, u_phi}; consumed (blk advances) only if the step
        // turns out to be a condensed-history step
        uint32_t w[4];
        philox4x32_rk<kPhiloxRounds>(blk, 0u, h0, h1, P.rk, w);
        const float u_mfp = u32_to_uniform(w[0]);
        const float u_phi = u32_to_uniform(w[3]);
        float z_loss, z_theta;
        box_muller(u32_to_uniform(w[1]), u32_to_uniform(w[2]), z_loss, z_theta);

        // relativistic quantities of vtx0.ke, base/mqi_relativistic_quantities.hpp:27-44
        constexpr float MeMp = kMe / kMp;
        const float Et       = ke + kMp;
        const float gamma    = Et * (1.0f / kMp);
        const float gamma_sq = gamma * gamma;
        // beta^2 gamma^2 = gamma^2 - 1 =: x.  One reciprocal of x serves 1 / beta^2 = gamma^2 / x (straggling
        // variance) and 1 / (P^2 beta^2) = gamma^2 / (Mp x)^2 (Highland angle): two MUFU.RCP per step instead of four
        const float bg_sq    = gamma_sq - 1.0f;
        const float inv_bg   = rcp_fast(bg_sq);
        const float inv_beta_sq = gamma_sq * inv_bg;
        const float Te_max   = (2.0f * kMe) * bg_sq * rcp_fast(fmaf(2.0f * MeMp, gamma, 1.0f + MeMp * MeMp));
        // rows of vtx0.ke: one row of the p-ion grid serves the delta cross section, |dEdx| and the csda
        // range; clamped rows make every lookup safe, out-of-table energies are masked by selects
        const int    ia   = row_a(ke);
        const float  ta   = ke - (0.1f + ia * 0.5f);
        const float4 A0   = sm.a0[ia];
        const float4 A1   = sm.a1[ia];
        const bool   in_a = ke <= 299.6f;
        const float  sp_w    = in_a ? fmaf(ta, A0.w, A0.z) : 0.f;
        const float  cs1_ion = in_a ? fmaf(ta, A0.y, A0.x) : 0.f;
        const int    ib   = row_b(ke);
        const float2 Bs1  = sm.bs[ib];
        const float  cs1_sum = cs1_ion + (ke <= 300.0f ? fmaf(ke - (0.5f + ib * 0.5f), Bs1.y, Bs1.x) : 0.f);

#if MQI_K_LATE_LUT
        // The LUT entry is addressed only now: the address takes the (always clear) sign bit of u_mfp as
        // a data dependency, so that the in-order issue does not park the warp on the index load
        // before the ~100 independent instructions above have been issued.
        const MatEntry M = G.lut[mat_idx + (__float_as_uint(u_mfp) >> 31)];
#endif
        // vtx0.dir with its tiny components zeroed in place by intersect() (z*), next to the un-zeroed copy the
        // reference keeps in vtx1.dir (d1*); the lane's own dx, dy, dz are only read here
        float       zx = dx, zy = dy, zz = dz;
        float       d1x = dx, d1y = dy, d1z = dz;
        const float tx = cell_tmax_axis(ex0, ex1, nx, px, zx, ix);
        const float ty = cell_tmax_axis(ey0, ey1, ny, py, zy, iy);
        const float tz = cell_tmax_axis(ez0, ez1, nz, pz, zz, iz);
        const float d2b = min3_ref(tx, ty, tz);
        const float rho = M.rho;
        // intersect() failed (the reference poisons the track and breaks), or a closed aperture voxel
        // (rho > 99.9: mqi_fippel_physics.hpp:81-85): the track ends without scoring
        if (!(d2b > 0.f) || rho > 99.9f) {
            fl = 0u;
            continue;   // rare; the lane rejoins at the ballot of the next turn
        }

        bool  stopped = false;
        float p1x, p1y, p1z;              // vtx1.pos
        const bool recoil = VARIANT == MQI_K_DEBUG && (fl & FL_RECOIL);
        float ke1 = recoil ? 0.f : ke;    // vtx1.ke

        StepResult res;
        res.dE = recoil ? ke : 0.f; res.local_dE = 0.f; res.te_debug = 0.f; res.len = 0.f;
        float rsp0 = 1.0f;
        {
            // the two rare kinds of step (vacuum voxel, track below the cut) share one test in the step body
            const bool vacuum = rho < 1.0e-7f;
            if (vacuum | (ke <= kTpCut)) {
                if (vacuum) {
                    // vacuum: move to the boundary (mqi_fippel_physics.hpp:77-80); a zero deposit of zero
                    // length is skipped by every hit function below, like the reference's early return
                    res.dE = 0.f;
                    p1x = px + zx * d2b; p1y = py + zy * d2b; p1z = pz + zz * d2b;
                } else {
                    // below the tracking cut: dump the energy, :86-94 + last_step mqi_p_ionization.hpp:482-490
                    if (ke < 0.f) ke = 0.f;
                    rsp0 = rsp_eval(M, ke);
                    res.dE += ke;
                    ke1 -= ke;
                    float step_len = 0.f;
                    if (res.dE > 0.f && ke > 0.f) {
                        const float liw = res.dE / stopping_power(sm, ke);
                        step_len        = liw * kWaterRho / (rsp0 * rho);
                    }
                    p1x = px + zx * step_len; p1y = py + zy * step_len; p1z = pz + zz * step_len;
                    res.len = step_len;
                    stopped = true;
                }
            } else {
                // ---------------- class-II condensed-history step, fippel_physics::stepping :95-216
                blk += 1;
                rsp0            = rsp_eval(M, ke);
                const float cms = rsp0 * rho * (1.0f / kWaterRho);   // WEPL of the 1 mm max step
                const float e2  = ke - cms * sp_w;   // energy after the largest possible CSDA loss
                // cross sections at e2 (< ke, so inside the tables from above); zero below the grids
                const int    i2  = row_a(e2);
                const float4 a2  = sm.a0[i2];
                const int    j2  = row_b(e2);
                const float2 b2  = sm.bs[j2];
                float cs2_ion_v = fmaf(e2 - (0.1f + i2 * 0.5f), a2.y, a2.x);
                float cs2_nuc_v = fmaf(e2 - (0.5f + j2 * 0.5f), b2.y, b2.x);
                asm("" : "+f"(cs2_ion_v), "+f"(cs2_nuc_v));   // evaluated by every lane, then selected: no branch around two FFMA
                const float cs2_ion = e2 >= 0.1f ? cs2_ion_v : 0.f;
                const float cs2_sum = cs2_ion + (e2 >= 0.5f ? cs2_nuc_v : 0.f);
                const bool  use1   = cs1_sum >= cs2_sum;
                const float cs_sum = (use1 ? cs1_sum : cs2_sum) * rho;
                const float c0     = (use1 ? cs1_ion : cs2_ion) * rho;   // delta-electron channel

                const float mfp = -logf(u_mfp) / cs_sum;
                constexpr float step_limit = 1.0f;   // cms * rho_w / (rsp * rho): max_step, mqi_fippel_physics.hpp:20
                const bool  to_boundary = d2b < mfp && d2b < step_limit;
                // same decision without short-circuit evaluation (no divergent branch in the step body)
                const bool  discrete    = !to_boundary & ((mfp < d2b) | (fabsf(mfp - d2b) < kGeomTol)) &
                                          ((mfp < step_limit) | (fabsf(mfp - step_limit) < kGeomTol));
                const float len         = to_boundary ? d2b : (discrete ? mfp : step_limit);
                // ---------------- along step (CSDA + straggling + MCS), mqi_p_ionization.hpp:298-420
                {
                    const float liw = len * cms;
                    const float R0  = fmaf(ta, A1.y, A1.x);   // residual csda range in water
                    // do { if (r >= r_steps[n]) break; } while (--n > 0) from n = min(ia, 598): the
                    // largest row n <= ia whose range does not exceed r (ranges increase with the row).
                    // Same row, found from a first guess (the row of ke - liw |dEdx|, exact above
                    // ~70 MeV) corrected against the table in both directions.  When the residual range
                    // is shorter than the step (R0 < liw) all of ke is lost; r is clamped so that the
                    // (discarded) inversion stays in the table.
                    const float r  = fmaxf(R0 - liw, 0.f);
                    const int   n0 = min(ia, kTableN - 2);
                    int         n  = min(max((int) ((fmaf(-liw, sp_w, ke) - 0.1f) * 2.0f), 0), n0);
                    float4      B  = sm.a1[n];
                    // the guess is right for nearly every step: one test (row n holds r, or nothing above /
                    // below to move to); the correction loops are out of line
                    {
                        const float up = sm.a1[min(n + 1, n0)].x;
                        if (!(((r >= B.x) | (n == 0)) & ((r < up) | (n == n0)))) {
                            n = csda_row_fix(sm.a1, n, n0, r);
                            B = sm.a1[n];
                        }
                    }
                    const float dE_mean = ke - fmaf(r - B.x, B.z, 0.1f + n * 0.5f);
                    const float Te      = fminf(Te_max, 0.08511f);
                    const float var     = P.dedx_term0 * rho * (1.0f / kWaterRho) * liw * (Te * (inv_beta_sq - 0.5f));
                    const float dE      = R0 < liw ? ke : fabsf(fmaf(z_loss, sqrtf(var), dE_mean));
                    float rr = 1.0f;
                    if (dE >= ke) {
                        rr      = ke / dE;
                        stopped = true;
                    }
                    const float th_sq = (13.9f * 13.9f / kMpSq) * inv_beta_sq * inv_bg * len * M.inv_x0;
                    const float th    = fabsf(z_theta) * sqrtf(2.0f * th_sq);
                    rotate_direction(d1x, d1y, d1z, th, kTwoPi * u_phi);
                    res.dE += dE * rr;
                    const float sl = rr * len;
                    p1x = fmaf(zx, sl, px); p1y = fmaf(zy, sl, py); p1z = fmaf(zz, sl, pz);
                    res.len = sl;
                    ke1 -= dE * rr;
                }
                // ---------------- discrete interaction at the end of the step, :156-197
                if (discrete && ke1 > kTpCut) {
                    d1x = zx; d1y = zy; d1z = zz;   // vtx1.dir = vtx0.dir (B11)
                    const float u = cs_sum * u_phi;  // u_phi is unused on this step: selects the process
                    if (u < c0) {
                        // delta electron, p_ionization_tabulated::post_step mqi_p_ionization.hpp:425-477.
                        // First try of the rejection loop without another generator call: given that the
                        // delta channel was selected, u / c0 is uniform in [0,1); the acceptance deviate
                        // comes from the top bytes of the step's block (unused by the 23-bit uniforms).
                        // Kinematics of vtx1.ke (relativistic_quantities, :27-44) with the divisions shared: with
                        // g = Et/Mp, D = 1 + 2 g Me/Mp + (Me/Mp)^2:  b^2 = 1 - 1/g^2,  Tmax = 2 Me (g^2 - 1) / D,
                        // b^2 / Tmax = D / (2 Me g^2),  1 / (2 Et^2) = 1 / (2 Mp^2 g^2)  -- four reciprocals in all.
                        const float g1    = (ke1 + kMp) * (1.0f / kMp);
                        const float g1_sq = g1 * g1;
                        const float r_g   = rcp_fast(g1_sq);
                        const float D1    = fmaf(2.0f * MeMp, g1, 1.0f + MeMp * MeMp);
                        const float Tmax1 = (2.0f * kMe) * (g1_sq - 1.0f) * rcp_fast(D1);
                        const float kb    = D1 * r_g * (0.5f / kMe);          // b1^2 / Tmax1
                        const float inv_2Et_sq = (0.5f / kMpSq) * r_g;
                        const float n  = fminf(u * rcp_fast(c0), 1.0f);
                        float       Te = T_cut * Tmax1 * rcp_fast(fmaf(n, T_cut - Tmax1, Tmax1));   // T_cut Tmax / ((1 - n) Tmax + n T_cut)
                        // The reference accepts with probability g(Te) = 1 - b^2 Te/Tmax + Te^2/(2 Et^2) (:447-451).  g
                        // falls with Te on [T_cut, Tmax] (g' < 0 for Te < b^2 Et^2 / Tmax, which is ~1e6 MeV), so
                        // g(T_cut) bounds it: accepting with g(Te) / g(T_cut) samples the same density with half
                        // the rejections (4 % instead of 9 %).
                        const float g_max = fmaf(T_cut * T_cut, inv_2Et_sq, fmaf(-kb, T_cut, 1.0f));
                        if (!(spare_bytes_to_uniform(w[0], w[1], w[2]) * g_max < fmaf(Te * Te, inv_2Et_sq, fmaf(-kb, Te, 1.0f)))) {
                            const float2 rt = delta_retry(blk, h0, k0 ^ (k1 * 0x85EBCA6Bu) ^ (h1 * 0xC2B2AE35u), T_cut, Tmax1, kb, inv_2Et_sq, g_max);
                            Te  = rt.x;
                            blk = __float_as_uint(rt.y);
                        }
                        if (VARIANT == MQI_K_DEBUG) res.te_debug = Te;   // carried by a zero-energy daughter
                        else res.dE += Te;
                        ke1 -= Te;
                    } else {
                        NucIO io;
                        io.px = px; io.py = py; io.pz = pz; io.dx = zx; io.dy = zy; io.dz = zz;
                        io.p1x = p1x; io.p1y = p1y; io.p1z = p1z;
                        io.d1x = d1x; io.d1y = d1y; io.d1z = d1z;
                        io.ke1 = ke1; io.dE = res.dE; io.local_dE = res.local_dE;
                        io.u = u - c0; io.e_cs = use1 ? ke : e2; io.rho = rho;
                        io.stopped = stopped ? 1 : 0;
                        io.sp = sp; io.node = node; io.n_sec = 0; io.n_ovf = 0;
                        io.rb.blk = blk; io.rb.pos = 4; io.rb.h0 = h0; io.rb.h1 = h1; io.rb.k0 = k0; io.rb.k1 = k1;
                        nuclear_event<VARIANT, MULTI>(P, stack, io);
                        d1x = io.d1x; d1y = io.d1y; d1z = io.d1z;
                        ke1 = io.ke1; res.dE = io.dE; res.local_dE = io.local_dE;
                        stopped = io.stopped != 0;
                        sp = io.sp;
                        if (io.n_sec) atomicAdd(P.counters + C_SECONDARIES, (unsigned long long) io.n_sec);
                        if (io.n_ovf) atomicAdd(P.counters + C_OVERFLOW, (unsigned long long) io.n_ovf);
                        blk = io.rb.blk;
                    }
                }
            }

            // -------------------------------------------------------------- scoring, :204-225
            const float inv_vol = 1.0f / ((ex1 - ex0) * (ey1 - ey0) * (ez1 - ez0));
            if (MULTI && node != P.n_nodes - 1) {
                // beamline children carry no scorers (n_scorers = 0, mqi_tps_env.hpp:751)
            } else if (SIMPLE) {
                // dose_to_water: (dE + local_dE) * 1.60218e-10 / (V * rho * rsp(rho, vtx0.ke)); voxel 0 is
                // never scored (roi_->idx(cnb) > 0, B1); insert_hashtable skips value <= 0.  The debug
                // variant's zero-energy delta daughter scores its own hit with rsp(rho, 0).
                const float kdose = 1.60218e-10f * inv_vol * M.inv_rho;
                float       vf    = (res.dE + res.local_dE) * kdose / rsp0;
                if (VARIANT == MQI_K_DEBUG) vf = fmaf(res.te_debug * kdose, inv_rsp_at_zero_energy(M), vf);
                // cvt.f64.f32 directly: under -ftz the compiler first flushes a denormal deposit with a multiply by one
                double v;
                asm("cvt.f64.f32 %0, %1;" : "=d"(v) : "f"(vf));
                if (cnb != 0u && vf > 0.f) atomicAdd(P.sc[0].dense + cnb, v);   // warp-match accumulation runs the general kernel
            } else {
                score_step<VARIANT, SET, DIJWC>(P, sm, M, cnb, spot_ind, inv_vol, rsp0, res, wc);
            }
        }

        // ------------------------------------------------------------------ advance, :227-232
        if (stopped) {
            fl = 0u;
        } else {
            ix = index_update_axis(ex0, ex1, p1x, d1x, ix);
            iy = index_update_axis(ey0, ey1, p1y, d1y, iy);
            iz = index_update_axis(ez0, ez1, p1z, d1z, iz);
            px = p1x; py = p1y; pz = p1z;
            dx = d1x; dy = d1y; dz = d1z;
            ke = ke1;
            fl = FL_ALIVE;
            if ((unsigned) ix >= (unsigned) nx || (unsigned) iy >= (unsigned) ny || (unsigned) iz >= (unsigned) nz)
                fl = (MULTI && node + 1 < P.n_nodes) ? FL_ADVANCE : 0u;   // the c_ind loop hands the track to the next child
        }
    }

    // per-lane step counter -> global (one atomic per lane per launch)
    if (COUNTED && n_steps && P.count_steps) atomicAdd(P.counters + C_STEPS, (unsigned long long) n_steps);
}


// =============================================================================================
// auxiliary kernels
// =============================================================================================
// 128 threads of at most 32 registers: 4 096 registers per CTA, what a 768-thread transport CTA (80 registers) leaves free on an
// SM -- so the conversion of the NEXT batch's HU volume (another handle, another stream) runs beside a resident transport
// kernel instead of waiting for it, and the host call that uploads the volume returns while the kernel before is still busy.
__global__ void __launch_bounds__(128, 16)
hu_to_material_kernel(const int16_t* __restrict__ hu, uint16_t* __restrict__ mat, size_t n) {
    // 8 voxels (16 B) per thread per iteration, grid-stride
    const size_t n8 = n / 8;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n8; i += (size_t) gridDim.x * blockDim.x) {
        const int4 v = reinterpret_cast<const int4*>(hu)[i];
        int        w[4] = { v.x, v.y, v.z, v.w };
        int4       o;
        int*       ow = &o.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int a = (int) (short) (w[k] & 0xffff), b = (int) (short) ((unsigned) w[k] >> 16);
            a = min(max(a, -1000), 2995) + 1000;
            b = min(max(b, -1000), 2995) + 1000;
            ow[k] = a | (b << 16);
        }
        reinterpret_cast<int4*>(mat)[i] = o;
    }
    for (size_t i = n8 * 8 + blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        mat[i] = (uint16_t) (min(max((int) hu[i], -1000), 2995) + 1000);
}

__global__ void
hu_to_density_kernel(const int16_t* __restrict__ hu, float* __restrict__ rho, size_t n,
                     const float* __restrict__ correction, float density_scale) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float r = hu_to_density((int) hu[i], correction);
        if (density_scale != 1.0f) r = __fmul_rn(r, density_scale);   // rho *= DensityScaling, mqi_tps_env.hpp:768
        rho[i] = r;
    }
}

__global__ void
dev_rsp_kernel(const MatEntry* __restrict__ m, const float* __restrict__ ek, size_t n, float* rsp, float* rl, int exact) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        rsp[i] = exact ? rsp_eval_exact(m[i], ek[i]) : rsp_eval(m[i], ek[i]);
        rl[i]  = m[i].x0;
    }
}

__global__ void
dev_grid_step_kernel(GridDev g, const float* __restrict__ pin, const float* __restrict__ din, size_t n,
                     int32_t* cell, unsigned long long* cnb, float* dist, float* dir_after, float* p_exit,
                     int32_t* cell_after) {
    const float* xe = g.edges;
    const float* ye = xe + g.nx + 1;
    const float* ze = ye + g.ny + 1;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const float p[3] = { pin[3 * i], pin[3 * i + 1], pin[3 * i + 2] };
        float       d[3] = { din[3 * i], din[3 * i + 1], din[3 * i + 2] };
        int ix = index_axis(xe, g.nx, p[0], d[0]);
        int iy = index_axis(ye, g.ny, p[1], d[1]);
        int iz = index_axis(ze, g.nz, p[2], d[2]);
        cell[3 * i] = ix; cell[3 * i + 1] = iy; cell[3 * i + 2] = iz;
        const bool valid = ix >= 0 && iy >= 0 && iz >= 0 && ix < g.nx && iy < g.ny && iz < g.nz;
        float t = -2.f;
        float q[3] = { p[0], p[1], p[2] };
        if (valid) {
            cnb[i] = ((unsigned long long) iz * g.ny + iy) * g.nx + ix;
            const float tx = cell_tmax_axis(xe[ix], xe[ix + 1], g.nx, p[0], d[0], ix);
            const float ty = cell_tmax_axis(ye[iy], ye[iy + 1], g.ny, p[1], d[1], iy);
            const float tz = cell_tmax_axis(ze[iz], ze[iz + 1], g.nz, p[2], d[2], iz);
            const float u  = min3_ref(tx, ty, tz);
            t              = u > 0.f ? u : -1.f;
            const float len = t > 0.f ? t : 0.f;
            q[0] = __fadd_rn(p[0], __fmul_rn(d[0], len));
            q[1] = __fadd_rn(p[1], __fmul_rn(d[1], len));
            q[2] = __fadd_rn(p[2], __fmul_rn(d[2], len));
            const int jx = index_update_axis(xe[ix], xe[ix + 1], q[0], d[0], ix);
            const int jy = index_update_axis(ye[iy], ye[iy + 1], q[1], d[1], iy);
            const int jz = index_update_axis(ze[iz], ze[iz + 1], q[2], d[2], iz);
            ix = jx; iy = jy; iz = jz;
        } else {
            cnb[i] = ~0ull;
        }
        dist[i] = t;
        dir_after[3 * i] = d[0]; dir_after[3 * i + 1] = d[1]; dir_after[3 * i + 2] = d[2];
        p_exit[3 * i] = q[0]; p_exit[3 * i + 1] = q[1]; p_exit[3 * i + 2] = q[2];
        cell_after[3 * i] = ix; cell_after[3 * i + 1] = iy; cell_after[3 * i + 2] = iz;
    }
}

__global__ void
dev_grid_entry_kernel(GridDev g, const float* __restrict__ pin, const float* __restrict__ din, size_t n,
                      float* dist, int32_t* cell) {
    const float* xe = g.edges;
    const float* ye = xe + g.nx + 1;
    const float* ze = ye + g.ny + 1;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const float p[3] = { pin[3 * i], pin[3 * i + 1], pin[3 * i + 2] };
        float       d[3] = { din[3 * i], din[3 * i + 1], din[3 * i + 2] };
        int         c[3];
        dist[i] = grid_entry(xe, ye, ze, g.nx, g.ny, g.nz, g.inv_w, p, d, c);
        cell[3 * i] = c[0]; cell[3 * i + 1] = c[1]; cell[3 * i + 2] = c[2];
    }
}

__global__ void
dev_hash_kernel(const uint32_t* k1, const uint32_t* k2, const unsigned long long* cap, size_t n, uint32_t* out) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        out[i] = hash_fun_magic(k1[i], k2[i], cap[i], remainder_magic(cap[i]));   // the form the transport kernel uses
}

__global__ void
dev_sample_kernel(SourceDev src, unsigned long long seed, unsigned long long first, size_t n, VertexDev* out,
                  uint32_t* spot_out) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const unsigned long long h = first + i;
        uint32_t lo = 0, hi = src.n_spots;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (src.cum[mid] > h) hi = mid; else lo = mid + 1;
        }
        const uint32_t spot = min(lo, src.n_spots - 1);
        VertexDev v;
        sample_vertex(src.beamlets[spot], seed, h, v);
        out[i]      = v;
        spot_out[i] = spot;
    }
}

// scores externally computed hits (voxel, spot, value) into a scorer: the insert_hashtable half of
// the path on its own (deterministic parity tests; also usable to merge deposits computed elsewhere)
__global__ void
dev_insert_kernel(ScorerDev sc, const uint32_t* __restrict__ k1, const uint32_t* __restrict__ k2,
                  const double* __restrict__ v, size_t n, unsigned long long* counters) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        if (!(v[i] > 0.0)) continue;
        if (sc.kind == MQI_K_DIJ) dij_add(sc, k1[i], k2[i], v[i], counters);
        else atomicAdd(sc.dense + k1[i], v[i]);
    }
}

__global__ void
fill_u64_kernel(unsigned long long* p, unsigned long long v, size_t n) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

__global__ void
dij_clear_kernel(DijSlot* t, size_t cap) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < cap; i += (size_t) gridDim.x * blockDim.x) {
        t[i].key   = kEmptyKey64;
        t[i].value = 0.0;
    }
}

__global__ void
dij_count_kernel(const DijSlot* t, size_t cap, unsigned long long* count) {
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < cap; i += (size_t) gridDim.x * blockDim.x)
        c += (t[i].key != kEmptyKey64);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// ordered compaction (slot order, like the reference's host scan mqi_io.hpp:105-160): one block
// scans a contiguous chunk; chunk offsets come from a first counting pass (d_chunk_off)
__global__ void
dij_chunk_count_kernel(const DijSlot* t, size_t cap, size_t chunk, unsigned long long* chunk_count) {
    const size_t b0 = blockIdx.x * chunk, b1 = min(cap, b0 + chunk);
    unsigned long long c = 0;
    for (size_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) c += (t[i].key != kEmptyKey64);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(chunk_count + blockIdx.x, c);
}

__global__ void
dij_chunk_write_kernel(const DijSlot* t, size_t cap, size_t chunk, const unsigned long long* chunk_off,
                       uint32_t* k1, uint32_t* k2, double* val, double scale) {
    // one warp per chunk keeps slot order with a ballot-based running offset
    const size_t b0 = blockIdx.x * chunk, b1 = min(cap, b0 + chunk);
    unsigned long long off = chunk_off[blockIdx.x];
    const int lane = threadIdx.x;
    for (size_t base = b0; base < b1; base += 32) {
        const size_t i   = base + lane;
        const bool   occ = i < b1 && t[i].key != kEmptyKey64;
        const unsigned m = __ballot_sync(0xffffffffu, occ);
        if (occ) {
            const unsigned long long o = off + __popc(m & ((1u << lane) - 1));
            const unsigned long long k = t[i].key;
            k1[o]  = (uint32_t) k;
            k2[o]  = (uint32_t) (k >> 32);
            val[o] = t[i].value * scale;
        }
        off += __popc(m);
    }
}

__global__ void
scale_kernel(double* p, size_t n, double f) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] *= f;
}

__device__ __forceinline__ double
atomic_max_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long  old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long) assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long) __double_as_longlong(v));
    } while (assumed != old);
    return __longlong_as_double((long long) old);
}

// max over voxels of the mean dose (sum / n), mqi_tps_env.hpp:1396-1407
__global__ void
stat_max_kernel(const double* __restrict__ sum, size_t n, double inv_n, double* out_max) {
    double m = 0.0;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        m = fmax(m, sum[i] * inv_n);
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_double(out_max, m);
}

// calculate_standard_deviation (kernel_functions/mqi_variables.hpp:20-48) fused with the host
// reduction of calculate_stat (mqi_tps_env.hpp:1409-1425):
//   mean = sum/n ; var = (sumsq/n - mean^2) / (n-1) ; sigma = sqrt(var)
//   out[0] += sigma/mean, out[1] += 1 for voxels with mean > cut
__global__ void
stat_partial_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, size_t n, double n_hist,
                    double cut, double* out2) {
    double acc = 0.0, cnt = 0.0;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const double mean = sum[i] / n_hist;
        if (mean > cut && mean > 0.0) {
            const double var = (sumsq[i] / n_hist - mean * mean) / (n_hist - 1.0);
            acc += sqrt(fmax(var, 0.0)) / mean;
            cnt += 1.0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(0xffffffffu, acc, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
        atomicAdd(out2, acc);
        atomicAdd(out2 + 1, cnt);
    }
}

// chunk flags of the stopping criterion over several devices: flag[c] = 1 if any value of chunk c exceeds `bound`
// (one block per chunk, grid-stride over the chunks)
__global__ void
chunk_above_kernel(const double* __restrict__ a, size_t n, size_t chunk, double bound, unsigned char* __restrict__ flag) {
    const size_t nchunks = (n + chunk - 1) / chunk;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const size_t b0 = c * chunk, b1 = min(n, b0 + chunk);
        bool         any = false;
        for (size_t i = b0 + threadIdx.x; i < b1 && !any; i += blockDim.x) any = a[i] > bound;
        const int r = __syncthreads_or(any);
        if (threadIdx.x == 0) flag[c] = r ? 1 : 0;
    }
}

// gather the listed chunks of two grids into packed buffers (zeros behind the end of the grid)
__global__ void
pack_chunks_kernel(const double* __restrict__ a, const double* __restrict__ b, size_t n, size_t chunk,
                   const unsigned int* __restrict__ list, size_t n_list, double* __restrict__ pa, double* __restrict__ pb) {
    for (size_t k = blockIdx.x; k < n_list; k += gridDim.x) {
        const size_t src = (size_t) list[k] * chunk, dst = k * chunk;
        for (size_t i = threadIdx.x; i < chunk; i += blockDim.x) {
            const bool in = src + i < n;
            pa[dst + i] = in ? a[src + i] : 0.0;
            pb[dst + i] = in ? b[src + i] : 0.0;
        }
    }
}

// =============================================================================================
// host launchers
// =============================================================================================
static inline int grid_for(size_t n, int block = 256, int cap = 148 * 16) {
    size_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (size_t) cap) g = cap;
    return (int) g;
}

size_t
transport_smem_bytes(int n_edge_floats, int n_nodes) {
    // sized for the largest CTA of the world's kernels (the single-Dij-scorer kernel may run another CTA size)
    const size_t block = (size_t) std::max(transport_block(n_nodes > 1, false), transport_block(n_nodes > 1, true));
    return smem_queue_offset(n_edge_floats, n_nodes) + (block / 32) * queue_words(n_nodes > 1) * sizeof(uint32_t);
}

// hand-over buffers of a multi-node launch: kRawWords words per warp of the grid
size_t
transport_handover_bytes(int grid, int n_nodes) {
#if MQI_K_ADV_QUEUE
    if (n_nodes > 1) return (size_t) grid * (MQI_K_BLOCK_MULTI / 32) * kRawWords * sizeof(uint32_t);
#endif
    (void) grid; (void) n_nodes;
    return 0;
}

int transport_scorer_set(const Params& p);

// threads per CTA of the kernels of a world with / without beamline children (the multi-node kernels carry a
// per-lane node descriptor: fewer threads per CTA leave them more registers)
int
transport_block(bool multi, bool dij_set) { return multi ? MQI_K_BLOCK_MULTI : (dij_set ? MQI_K_BLOCK_DIJ : MQI_K_BLOCK); }
int
transport_block(const Params& p) { return transport_block(p.n_nodes > 1, p.dij_wc_scorer >= 0 && transport_scorer_set(p) == SET_DIJ); }

typedef void (*transport_fn)(const Params);
template<int SET, bool MULTI, bool DIJWC>
static transport_fn
pick_variant(int variant) {
    return variant == MQI_K_DEBUG ? transport_kernel<MQI_K_DEBUG, SET, MULTI, DIJWC> : transport_kernel<MQI_K_RELEASE, SET, MULTI, DIJWC>;
}
static transport_fn
pick_transport(int variant, int set, bool multi, bool dijwc) {
    if (dijwc && set == SET_DIJ) return multi ? pick_variant<SET_DIJ, true, true>(variant) : pick_variant<SET_DIJ, false, true>(variant);
    if (dijwc) return multi ? pick_variant<SET_GENERIC, true, true>(variant) : pick_variant<SET_GENERIC, false, true>(variant);
    switch (set) {
    case SET_DOSE: return multi ? pick_variant<SET_DOSE, true, false>(variant) : pick_variant<SET_DOSE, false, false>(variant);
    case SET_DOSE_LETD: return multi ? pick_variant<SET_DOSE_LETD, true, false>(variant) : pick_variant<SET_DOSE_LETD, false, false>(variant);
    case SET_DOSE_STAT: return multi ? pick_variant<SET_DOSE_STAT, true, false>(variant) : pick_variant<SET_DOSE_STAT, false, false>(variant);
    default: return multi ? pick_variant<SET_GENERIC, true, false>(variant) : pick_variant<SET_GENERIC, false, false>(variant);
    }
}

// the compile-time scorer set a launch qualifies for (see ScorerSet)
int
transport_scorer_set(const Params& p) {
    if ((p.quirks & MQI_K_QUIRK_B2) || p.accum_mode != MQI_K_ACCUM_ATOMIC || p.count_steps) return SET_GENERIC;
    const ScorerDev* sc = p.sc;
    if (p.n_scorers == 1 && sc[0].kind == MQI_K_DIJ && p.dij_wc_scorer == 0) return SET_DIJ;
    if (p.n_scorers == 1 && sc[0].kind == MQI_K_DOSE && !sc[0].roi) return SET_DOSE;
    if (p.n_scorers == 3 && sc[0].kind == MQI_K_DOSE && sc[1].kind == MQI_K_LETD_NUMER && sc[2].kind == MQI_K_LETD_DENOM &&
        !sc[0].roi && !sc[1].roi && !sc[2].roi)
        return SET_DOSE_LETD;
    if (p.n_scorers == 3 && sc[0].kind == MQI_K_DOSE && sc[1].kind == MQI_K_DOSE && sc[2].kind == MQI_K_DOSE_SQ)
        return SET_DOSE_STAT;
    return SET_GENERIC;
}

// one dense Dose scorer with a DIRECT roi -> the kernels with the scorer loop compiled out
bool
transport_is_simple(const Params& p) {
    return transport_scorer_set(p) == SET_DOSE;
}

cudaError_t
transport_occupancy(const Params& p, int variant, size_t smem, int* blocks_per_sm) {
    transport_fn f = pick_transport(variant, transport_scorer_set(p), p.n_nodes > 1, p.dij_wc_scorer >= 0);
    cudaError_t  e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, f, transport_block(p), smem);
}

cudaError_t
launch_transport(const Params& p, int variant, int grid, size_t smem, cudaStream_t st) {
    pick_transport(variant, transport_scorer_set(p), p.n_nodes > 1, p.dij_wc_scorer >= 0)<<<grid, transport_block(p), smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t
launch_hu_to_material(const int16_t* d_hu, uint16_t* d_mat, size_t n, cudaStream_t st) {
    hu_to_material_kernel<<<grid_for(n / 8 + 1, 128), 128, 0, st>>>(d_hu, d_mat, n);
    return cudaGetLastError();
}
cudaError_t
launch_hu_to_density(const int16_t* d_hu, float* d_rho, size_t n, const float* d_correction, float scale, cudaStream_t st) {
    hu_to_density_kernel<<<grid_for(n), 256, 0, st>>>(d_hu, d_rho, n, d_correction, scale);
    return cudaGetLastError();
}
cudaError_t
launch_dev_rsp(const MatEntry* m, const float* ek, size_t n, float* rsp, float* rl, cudaStream_t st, bool exact) {
    dev_rsp_kernel<<<grid_for(n), 256, 0, st>>>(m, ek, n, rsp, rl, exact ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t
launch_dev_grid_step(const Params& p, const float* pin, const float* din, size_t n, int32_t* cell,
                     unsigned long long* cnb, float* dist, float* dir_after, float* p_exit, int32_t* cell_after,
                     cudaStream_t st) {
    dev_grid_step_kernel<<<grid_for(n), 256, 0, st>>>(p.g, pin, din, n, cell, cnb, dist, dir_after, p_exit, cell_after);
    return cudaGetLastError();
}
cudaError_t
launch_dev_grid_entry(const Params& p, const float* pin, const float* din, size_t n, float* dist, int32_t* cell,
                      cudaStream_t st) {
    dev_grid_entry_kernel<<<grid_for(n), 256, 0, st>>>(p.g, pin, din, n, dist, cell);
    return cudaGetLastError();
}
cudaError_t
launch_dev_hash(const uint32_t* k1, const uint32_t* k2, const unsigned long long* cap, size_t n, uint32_t* out,
                cudaStream_t st) {
    dev_hash_kernel<<<grid_for(n), 256, 0, st>>>(k1, k2, cap, n, out);
    return cudaGetLastError();
}
cudaError_t
launch_dev_sample(const Params& p, unsigned long long first, size_t n, VertexDev* out, uint32_t* spot, cudaStream_t st) {
    dev_sample_kernel<<<grid_for(n), 256, 0, st>>>(p.src, p.seed, first, n, out, spot);
    return cudaGetLastError();
}
cudaError_t
launch_dev_insert(const Params& p, int scorer, const uint32_t* k1, const uint32_t* k2, const double* v, size_t n,
                  cudaStream_t st) {
    dev_insert_kernel<<<grid_for(n), 256, 0, st>>>(p.sc[scorer], k1, k2, v, n, p.counters);
    return cudaGetLastError();
}
cudaError_t
launch_fill_u64(unsigned long long* p, unsigned long long v, size_t n, cudaStream_t st) {
    fill_u64_kernel<<<grid_for(n), 256, 0, st>>>(p, v, n);
    return cudaGetLastError();
}
cudaError_t
launch_dij_clear(void* table, size_t capacity, cudaStream_t st) {
    dij_clear_kernel<<<grid_for(capacity), 256, 0, st>>>(static_cast<DijSlot*>(table), capacity);
    return cudaGetLastError();
}
cudaError_t
launch_dij_count(const void* table, size_t capacity, unsigned long long* d_count, cudaStream_t st) {
    dij_count_kernel<<<grid_for(capacity), 256, 0, st>>>(static_cast<const DijSlot*>(table), capacity, d_count);
    return cudaGetLastError();
}
cudaError_t
launch_scale(double* p, size_t n, double f, cudaStream_t st) {
    scale_kernel<<<grid_for(n), 256, 0, st>>>(p, n, f);
    return cudaGetLastError();
}
cudaError_t
launch_stat_max(const double* sum, size_t n, double inv_n, double* d_out_max, cudaStream_t st) {
    stat_max_kernel<<<grid_for(n), 256, 0, st>>>(sum, n, inv_n, d_out_max);
    return cudaGetLastError();
}
cudaError_t
launch_stat_partial(const double* sum, const double* sumsq, size_t n, double n_hist, double cut, double* d_out2,
                    cudaStream_t st) {
    stat_partial_kernel<<<grid_for(n), 256, 0, st>>>(sum, sumsq, n, n_hist, cut, d_out2);
    return cudaGetLastError();
}
cudaError_t
launch_chunk_above(const double* a, size_t n, size_t chunk, double bound, unsigned char* d_flag, cudaStream_t st) {
    const size_t nchunks = (n + chunk - 1) / chunk;
    chunk_above_kernel<<<(int) std::min<size_t>(nchunks, 148 * 8), 256, 0, st>>>(a, n, chunk, bound, d_flag);
    return cudaGetLastError();
}
cudaError_t
launch_pack_chunks(const double* a, const double* b, size_t n, size_t chunk, const unsigned int* d_list, size_t n_list, double* pa,
                   double* pb, cudaStream_t st) {
    if (n_list == 0) return cudaSuccess;
    pack_chunks_kernel<<<(int) std::min<size_t>(n_list, 148 * 8), 256, 0, st>>>(a, b, n, chunk, d_list, n_list, pa, pb);
    return cudaGetLastError();
}
cudaError_t
launch_dij_chunk_count(const void* table, size_t capacity, size_t chunk, unsigned long long* d_chunk_count, cudaStream_t st) {
    const int nchunks = (int) ((capacity + chunk - 1) / chunk);
    dij_chunk_count_kernel<<<nchunks, 128, 0, st>>>(static_cast<const DijSlot*>(table), capacity, chunk, d_chunk_count);
    return cudaGetLastError();
}
cudaError_t
launch_dij_chunk_write(const void* table, size_t capacity, size_t chunk, const unsigned long long* d_chunk_off,
                       uint32_t* k1, uint32_t* k2, double* val, double scale, cudaStream_t st) {
    const int nchunks = (int) ((capacity + chunk - 1) / chunk);
    dij_chunk_write_kernel<<<nchunks, 32, 0, st>>>(static_cast<const DijSlot*>(table), capacity, chunk, d_chunk_off, k1, k2, val, scale);
    return cudaGetLastError();
}

}   // namespace mqib
