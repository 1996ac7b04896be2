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

namespace mqib
{

// ---------------------------------------------------------------------------------------------
// shared-memory view
// ---------------------------------------------------------------------------------------------
struct Smem {
    const float4* tab_a;
    const float4* tab_b;
    const float*  xe;
    const float*  ye;
    const float*  ze;
};

__device__ __forceinline__ int tab_index(float ek, float e0) {
    // uint16_t((Ek - Ei) / 0.5)
    return (int) (unsigned short) (int) ((ek - e0) * 2.0f);
}

// tabulated cross sections of the four discrete processes at kinetic energy ek, times rho
// (mqi_p_ionization.hpp:254-268, mqi_pp_elastic.hpp:221-235, mqi_po_elastic.hpp:243-256,
// mqi_po_inelastic.hpp:141-155)
__device__ __forceinline__ void
cross_sections(const Smem& sm, float ek, float rho, float cs[4]) {
    cs[0] = cs[1] = cs[2] = cs[3] = 0.f;
    if (ek >= 0.1f && ek <= 299.6f) {
        const int   i0 = tab_index(ek, 0.1f);
        const int   i1 = min(i0 + 1, kTableN - 1);
        const float x0 = 0.1f + i0 * 0.5f;
        cs[0]          = intpl1d(ek, x0, x0 + 0.5f, sm.tab_a[i0].x, sm.tab_a[i1].x);
    }
    if (ek >= 0.5f && ek <= 300.0f) {
        const int    i0 = min(tab_index(ek, 0.5f), kTableN - 1);
        const int    i1 = min(i0 + 1, kTableN - 1);
        const float  x0 = 0.5f + i0 * 0.5f;
        const float  x1 = x0 + 0.5f;
        const float4 a = sm.tab_b[i0], b = sm.tab_b[i1];
        cs[1] = intpl1d(ek, x0, x1, a.x, b.x);
        cs[2] = intpl1d(ek, x0, x1, a.y, b.y);
        cs[3] = intpl1d(ek, x0, x1, a.z, b.z);
    }
    cs[0] *= rho; cs[1] *= rho; cs[2] *= rho; cs[3] *= rho;
}

// |dEdx| in water (restricted stopping power), mqi_p_ionization.hpp:271-286
__device__ __forceinline__ float
stopping_power(const Smem& sm, float ek) {
    if (ek >= 0.1f && ek <= 299.6f) {
        const int   i0 = tab_index(ek, 0.1f);
        const int   i1 = min(i0 + 1, kTableN - 1);
        const float x0 = 0.1f + i0 * 0.5f;
        return intpl1d(ek, x0, x0 + 0.5f, sm.tab_a[i0].y, sm.tab_a[i1].y);
    }
    if (ek < 0.1f && ek > 0.f) return sm.tab_a[0].y;
    return 0.f;
}

// CSDA energy loss over a water-equivalent length + Gaussian straggling,
// p_ionization_tabulated::energy_loss / energy_straggling  mqi_p_ionization.hpp:298-345
__device__ __forceinline__ float
energy_loss(const Smem& sm, const Rel& rel, float rho, float liw, float z, float dedx_term0) {
    int         n  = tab_index(rel.Ek, 0.1f);
    const float x0 = 0.1f + n * 0.5f;
    const float x1 = x0 + 0.5f;
    if (x0 > rel.Ek) n -= 1;
    if (x1 < rel.Ek) n += 1;
    n       = max(0, min(n, kTableN - 2));
    float r = intpl1d(rel.Ek, x0, x1, sm.tab_a[n].z, sm.tab_a[n + 1].z);
    if (r < liw) return rel.Ek;
    r -= liw;
    while (n > 0 && r < sm.tab_a[n].z) --n;   // do { if (r >= r_steps[n]) break; } while (--n > 0)
    const float y0      = 0.1f + n * 0.5f;
    const float dE_mean = rel.Ek - intpl1d(r, sm.tab_a[n].z, sm.tab_a[n + 1].z, y0, y0 + 0.5f);
    const float Te      = fminf(rel.Te_max, 0.08511f);
    const float var     = dedx_term0 * rho / kWaterRho * liw * (Te / rel.beta_sq * (1.0f - 0.5f * rel.beta_sq));
    return fabsf(z * sqrtf(var) + dE_mean);
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

template<int VARIANT>
__device__ __forceinline__ void
push_secondary(const Params& P, Secondary* stack, int& sp, float x, float y, float z, float ux, float uy,
               float uz, float ke0, float ke1_off, float dE_pre, unsigned long long& n_sec,
               unsigned long long& n_ovf) {
    constexpr int DEPTH = StackCfg<VARIANT>::depth;
    if (sp >= DEPTH) {   // overflow silently drops the secondary: mqi_track_stack.hpp:36-41 (B10)
        ++n_ovf;
        return;
    }
    if (!P.g.identity) {
        // daughters are mapped with Rfwd * (x - T) + T, mqi_pp_elastic.hpp:188-195
        const float* R  = P.g.rot_fwd;
        const float  qx = x - P.g.trans[0], qy = y - P.g.trans[1], qz = z - P.g.trans[2];
        x = R[0] * qx + R[1] * qy + R[2] * qz + P.g.trans[0];
        y = R[3] * qx + R[4] * qy + R[5] * qz + P.g.trans[1];
        z = R[6] * qx + R[7] * qy + R[8] * qz + P.g.trans[2];
        const float ex = ux, ey = uy, ez = uz;
        ux = R[0] * ex + R[1] * ey + R[2] * ez;
        uy = R[3] * ex + R[4] * ey + R[5] * ez;
        uz = R[6] * ex + R[7] * ey + R[8] * ez;
    }
    Secondary& s = stack[sp++];
    s.px = x; s.py = y; s.pz = z; s.dx = ux; s.dy = uy; s.dz = uz;
    s.ke0 = ke0; s.ke1_off = ke1_off; s.dE_pre = dE_pre;
    ++n_sec;
}

// ---------------------------------------------------------------------------------------------
// device beam source: beamlet::operator()  base/mqi_beamlet.hpp:81-90 with
// phsp_6d_uniform / phsp_6d (distributions/mqi_phsp6d_uniform.hpp:68-85, mqi_phsp6d.hpp:62-79),
// const_1d / norm_1d.  RNG protocol: block 0 = {Ux, Vx, Uy, Vy}, block 1 = {za, zb, uc, -}.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void
sample_vertex(const BeamletDev& b, Rng& rng, VertexDev& out) {
    float Ux, Vx, Uy, Vy;
    rng_begin_step(rng);
    {
        const float u0 = rng_uniform(rng), u1 = rng_uniform(rng), u2 = rng_uniform(rng), u3 = rng_uniform(rng);
        if (b.phsp_uniform) {
            Ux = 2.0f * u0 - 1.0f; Vx = 2.0f * u1 - 1.0f; Uy = 2.0f * u2 - 1.0f; Vy = 2.0f * u3 - 1.0f;
        } else {
            box_muller(u0, u1, Ux, Vx);
            box_muller(u2, u3, Uy, Vy);
        }
    }
    rng_begin_step(rng);
    float za, zb;
    {
        const float u0 = rng_uniform(rng), u1 = rng_uniform(rng);
        box_muller(u0, u1, za, zb);
    }
    const float uc = rng_uniform(rng);
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
__device__ __forceinline__ void
dij_add(const ScorerDev& s, uint32_t key1, uint32_t key2, double v, unsigned long long* counters) {
    unsigned long long slot;
    if (key2 == kEmptyKey32) {   // dense mode of the reference: slot = voxel, key2 := 0
        slot = key1;
        key2 = 0;
    } else {
        slot = hash_fun(key1, key2, s.capacity);
    }
    const unsigned long long key = ((unsigned long long) key2 << 32) | key1;
    for (unsigned long long probes = 0; probes < s.capacity; ++probes) {
        DijSlot*           e    = s.table + slot;
        unsigned long long prev = *reinterpret_cast<volatile unsigned long long*>(&e->key);
        if (prev == kEmptyKey64) prev = atomicCAS(&e->key, kEmptyKey64, key);
        if (prev == kEmptyKey64 || prev == key) {
            atomicAdd(&e->value, v);
            return;
        }
        slot = (slot + 1) % s.capacity;
    }
    atomicAdd(counters + C_DIJ_FULL, 1ull);   // the reference would spin forever here
}

struct StepResult {
    float dE;        // trk.dE of the step (incl. dE_pre)
    float local_dE;  // trk.local_dE
    float te_debug;  // delta-electron energy carried by the (folded) zero-energy daughter, debug only
    float len;       // |vtx1.pos - vtx0.pos|
    float ke0;       // vtx0.ke
};

template<int VARIANT>
__device__ __forceinline__ void
score_step(const Params& P, const MatEntry& M, unsigned cnb, uint32_t spot_ind, float inv_vol, float rsp0,
           const StepResult& r) {
    if ((int) cnb <= 0) return;   // roi_->idx(cnb) > 0 with a DIRECT roi: voxel 0 is never scored (B1)
    // dose_to_water: (dE + local_dE) * 1.60218e-10 / (V * rho * rsp(rho, vtx0.ke))
    const float  kdose   = 1.60218e-10f * inv_vol * M.inv_rho;
    const double dose    = (M.rho < 1.0e-7f) ? 0.0 : (double) ((r.dE + r.local_dE) * kdose / rsp0);
    const double dose_te = (VARIANT == MQI_K_DEBUG && M.rho >= 1.0e-7f) ? (double) (r.te_debug * kdose * M.inv_rsp0) : 0.0;
    const int n = P.n_scorers;
#pragma unroll 1
    for (int pass = (P.quirks & MQI_K_QUIRK_B2) ? 0 : 1; pass < 2; ++pass) {
        const int s_end = pass == 0 ? n - 2 : n;
#pragma unroll 1
        for (int s = 0; s < s_end; ++s) {
            const ScorerDev& sc = P.sc[s];
            double           v  = 0.0;
            switch (sc.kind) {
            case MQI_K_DOSE: v = dose + dose_te; break;
            case MQI_K_DIJ: v = dose + dose_te; break;
            case MQI_K_DOSE_SQ: v = dose * dose + dose_te * dose_te; break;
            case MQI_K_EDEP: v = (double) (r.dE + r.local_dE) + (double) r.te_debug; break;
            case MQI_K_LETD_NUMER:
            case MQI_K_LETD_DENOM: {
                // LETd_weight1/2: scorers/mqi_scorer_energy_deposit.hpp:93-137
                if (r.len > 0.f) {
                    const double let = (double) r.dE / (double) r.len / (double) (M.rho * 1000.0f);
                    if (let < 25.0) v = sc.kind == MQI_K_LETD_NUMER ? (double) r.dE * let : (double) r.dE;
                }
                break;
            }
            default: break;
            }
            if (!(v > 0.0)) continue;   // insert_hashtable: value <= 0 -> skip
            if (sc.kind == MQI_K_DIJ) dij_add(sc, cnb, spot_ind, v, P.counters);
            else dense_add(sc.dense, cnb, v, P.accum_mode);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the transport kernel
// ---------------------------------------------------------------------------------------------
template<int VARIANT>
__global__ void __launch_bounds__(MQI_K_BLOCK, MQI_K_MIN_BLOCKS)
transport_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_tab_a = reinterpret_cast<float4*>(smem_raw);
    float4* s_tab_b = s_tab_a + kTableN;
    float*  s_edges = reinterpret_cast<float*>(s_tab_b + kTableN);
    const int nx = P.g.nx, ny = P.g.ny, nz = P.g.nz;
    for (int i = threadIdx.x; i < kTableN; i += blockDim.x) {
        s_tab_a[i] = P.tab_a[i];
        s_tab_b[i] = P.tab_b[i];
    }
    for (int i = threadIdx.x; i < nx + ny + nz + 3; i += blockDim.x) s_edges[i] = P.g.edges[i];
    __syncthreads();
    Smem sm;
    sm.tab_a = s_tab_a;
    sm.tab_b = s_tab_b;
    sm.xe    = s_edges;
    sm.ye    = s_edges + nx + 1;
    sm.ze    = s_edges + nx + 1 + ny + 1;

    constexpr float T_cut = (VARIANT == MQI_K_DEBUG) ? 0.08511f : 0.0815f;   // mqi_interaction.hpp:24-28
    constexpr int   DEPTH = StackCfg<VARIANT>::depth;
    Secondary stack[DEPTH];
    int       sp = 0;

    // lane state
    float    px = 0, py = 0, pz = 0, dx = 0, dy = 0, dz = 0, ke = 0, ke1_off = 0, dE_pre = 0;
    int      ix = 0, iy = 0, iz = 0;
    bool     alive = false;
    uint32_t spot_ind = kEmptyKey32;
    Rng      rng;
    rng_init(rng, P.seed, 0);
    unsigned long long n_steps = 0, n_done = 0, n_sec = 0, n_ovf = 0;

    while (true) {
        // ------------------------------------------------------------------ re-arm the lane
        if (!alive) {
            if (sp > 0) {
                const Secondary& s = stack[--sp];
                px = s.px; py = s.py; pz = s.pz; dx = s.dx; dy = s.dy; dz = s.dz;
                ke = s.ke0; ke1_off = s.ke1_off; dE_pre = s.dE_pre;
            } else {
                const unsigned long long i = atomicAdd(P.counters + C_NEXT, 1ull);
                if (i >= P.count) break;
                const unsigned long long h = P.first + i;
                rng_init(rng, P.seed, h);
                uint32_t spot = 0;
                if (P.src.vertices) {
                    const VertexDev v = P.src.vertices[i];
                    px = v.pos[0]; py = v.pos[1]; pz = v.pos[2];
                    dx = v.dir[0]; dy = v.dir[1]; dz = v.dir[2];
                    ke = v.ke;
                    spot = P.src.spot_ids ? P.src.spot_ids[i] : 0u;
                } else {
                    // beamsource::operator()(h): first spot whose cumulative count exceeds h
                    uint32_t lo = 0, hi = P.src.n_spots;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (P.src.cum[mid] > h) hi = mid; else lo = mid + 1;
                    }
                    spot = min(lo, P.src.n_spots - 1);
                    VertexDev v;
                    sample_vertex(P.src.beamlets[spot], rng, v);
                    px = v.pos[0]; py = v.pos[1]; pz = v.pos[2];
                    dx = v.dir[0]; dy = v.dir[1]; dz = v.dir[2];
                    ke = v.ke;
                }
                spot_ind = P.per_spot ? spot : kEmptyKey32;
                ke1_off  = 0.f;
                dE_pre   = 0.f;
                ++n_done;
            }
            // world -> node frame, mqi_transport.hpp:165-170
            if (!P.g.identity) {
                const float* R  = P.g.rot_fwd;   // inverse = transpose
                const float  qx = px - P.g.trans[0], qy = py - P.g.trans[1], qz = pz - P.g.trans[2];
                px = R[0] * qx + R[3] * qy + R[6] * qz;
                py = R[1] * qx + R[4] * qy + R[7] * qz;
                pz = R[2] * qx + R[5] * qy + R[8] * qz;
                const float ex = dx, ey = dy, ez = dz;
                dx = R[0] * ex + R[3] * ey + R[6] * ez;
                dy = R[1] * ex + R[4] * ey + R[7] * ez;
                dz = R[2] * ex + R[5] * ey + R[8] * ez;
            }
            {
                const float n = sqrtf(dx * dx + dy * dy + dz * dz);
                dx /= n; dy /= n; dz /= n;
            }
            // locate: index(p, dir) or entry intersect, :171-190
            ix = index_axis(sm.xe, nx, px, dx);
            iy = index_axis(sm.ye, ny, py, dy);
            iz = index_axis(sm.ze, nz, pz, dz);
            alive = true;
            if (ix < 0 || iy < 0 || iz < 0 || ix >= nx || iy >= ny || iz >= nz) {
                const float p[3] = { px, py, pz };
                float       d[3] = { dx, dy, dz };
                int         c[3];
                const float dist = grid_entry(sm.xe, sm.ye, sm.ze, nx, ny, nz, p, d, c);
                if (dist < 0.f) {
                    alive = false;
                } else {
                    // update_post_vertex_position uses the (possibly zeroed) direction, move() then
                    // restores the un-zeroed copy held in vtx1.dir; vtx1.ke becomes vtx0.ke
                    px = __fadd_rn(px, __fmul_rn(d[0], dist));
                    py = __fadd_rn(py, __fmul_rn(d[1], dist));
                    pz = __fadd_rn(pz, __fmul_rn(d[2], dist));
                    ke += ke1_off;
                    ke1_off = 0.f;
                    dE_pre  = 0.f;
                    ix = index_axis(sm.xe, nx, px, dx);
                    iy = index_axis(sm.ye, ny, py, dy);
                    iz = index_axis(sm.ze, nz, pz, dz);
                    if (ix < 0 || iy < 0 || iz < 0 || ix >= nx || iy >= ny || iz >= nz) alive = false;
                }
            }
            if (!alive) continue;
        }

        // ------------------------------------------------------------------ one voxel step
        if (P.count_steps) ++n_steps;
        const float ex0 = sm.xe[ix], ex1 = sm.xe[ix + 1];
        const float ey0 = sm.ye[iy], ey1 = sm.ye[iy + 1];
        const float ez0 = sm.ze[iz], ez1 = sm.ze[iz + 1];
        const unsigned cnb = ((unsigned) iz * (unsigned) ny + (unsigned) iy) * (unsigned) nx + (unsigned) ix;
        float       d1x = dx, d1y = dy, d1z = dz;   // vtx1.dir: copy taken before intersect() zeroes tiny components
        const float tx = cell_tmax_axis(ex0, ex1, nx, px, dx, ix);
        const float ty = cell_tmax_axis(ey0, ey1, ny, py, dy, iy);
        const float tz = cell_tmax_axis(ez0, ez1, nz, pz, dz, iz);
        const float d2b = min3_ref(tx, ty, tz);
        if (!(d2b > 0.f)) {   // intersect() failed: the reference poisons the track and breaks
            alive = false;
            continue;
        }
        const MatEntry M   = P.g.lut[__ldg(P.g.mat + cnb)];
        const float    rho = M.rho;

        bool  stopped = false;
        float p1x, p1y, p1z;              // vtx1.pos
        float ke1 = ke + ke1_off;         // vtx1.ke
        StepResult res;
        res.dE = dE_pre; res.local_dE = 0.f; res.te_debug = 0.f; res.len = 0.f; res.ke0 = ke;
        float rsp0 = 1.f;

        if (rho < 1.0e-7f) {
            // vacuum: move to the boundary, nothing to score (mqi_fippel_physics.hpp:77-80)
            p1x = px + dx * d2b; p1y = py + dy * d2b; p1z = pz + dz * d2b;
        } else if (rho > 99.9f) {
            stopped = true;   // closed aperture (:81-85)
            p1x = px; p1y = py; p1z = pz;
        } else if (ke <= kTpCut) {
            // below the tracking cut: dump the energy, :86-94 + last_step mqi_p_ionization.hpp:482-490
            if (ke < 0.f) ke = 0.f;
            rsp0 = rsp_eval(M, ke);
            res.ke0 = ke;
            res.dE += ke;
            ke1 -= ke;
            float step_len = 0.f;
            if (res.dE > 0.f && ke > 0.f) {
                const float liw = res.dE / stopping_power(sm, ke);
                step_len        = liw * kWaterRho / (rsp0 * rho);
            }
            p1x = px + dx * step_len; p1y = py + dy * step_len; p1z = pz + dz * step_len;
            res.len = step_len;
            stopped = true;
        } else {
            // ---------------- class-II condensed-history step, fippel_physics::stepping :95-216
            const Rel   rel = rel_make(ke);
            rsp0            = rsp_eval(M, ke);
            const float cms = 1.0f * rsp0 * rho / kWaterRho;            // WEPL of the 1 mm max step
            const float max_loss = cms * stopping_power(sm, ke);
            float cs1[4], cs2[4];
            cross_sections(sm, ke, rho, cs1);
            cross_sections(sm, ke - max_loss, rho, cs2);
            const float cs1_sum = cs1[0] + cs1[1] + cs1[2] + cs1[3];
            const float cs2_sum = cs2[0] + cs2[1] + cs2[2] + cs2[3];
            const bool  use1    = cs1_sum >= cs2_sum;
            const float cs_sum  = use1 ? cs1_sum : cs2_sum;
            const float c0 = use1 ? cs1[0] : cs2[0], c1 = use1 ? cs1[1] : cs2[1];
            const float c2 = use1 ? cs1[2] : cs2[2], c3 = use1 ? cs1[3] : cs2[3];

            rng_begin_step(rng);
            const float u_mfp = rng_uniform(rng);
            const float u_a   = rng_uniform(rng);
            const float u_b   = rng_uniform(rng);
            const float u_phi = rng_uniform(rng);
            float z_loss, z_theta;
            box_muller(u_a, u_b, z_loss, z_theta);

            const float mfp        = -1.0f * logf(u_mfp) / cs_sum;
            const float step_limit = cms * kWaterRho / (rsp0 * rho);
            float len;
            bool  discrete = false;
            if (d2b < mfp && d2b < step_limit) {
                len = d2b;
            } else if ((mfp < d2b || fabsf(mfp - d2b) < kGeomTol) &&
                       (mfp < step_limit || fabsf(mfp - step_limit) < kGeomTol)) {
                len      = mfp;
                discrete = true;
            } else {
                len = step_limit;
            }
            // ---------------- along step (CSDA + straggling + MCS), mqi_p_ionization.hpp:349-420
            {
                const float liw = len * rsp0 * rho / kWaterRho;
                float       dE  = energy_loss(sm, rel, rho, liw, z_loss, P.dedx_term0);
                float       r   = 1.0f;
                if (dE >= ke) {
                    r       = ke / dE;
                    stopped = true;
                }
                const float P_sq  = rel.Et * rel.Et - kMpSq;
                const float th_sq = ((13.9f * 13.9f / P_sq) / rel.beta_sq) * len * M.inv_x0;
                const float th    = fabsf(z_theta * (1.41421356237f * sqrtf(th_sq)));
                const float phi   = kTwoPi * u_phi;
                rotate_direction(d1x, d1y, d1z, th, phi);
                res.dE += dE * r;
                const float sl = r * len;
                p1x = px + dx * sl; p1y = py + dy * sl; p1z = pz + dz * sl;
                res.len = sl;
                ke1 -= dE * r;
            }
            // ---------------- discrete interaction at the end of the step, :156-197
            if (discrete && ke1 > kTpCut) {
                const float u = cs_sum * rng_uniform(rng);
                d1x = dx; d1y = dy; d1z = dz;   // vtx1.dir = vtx0.dir (B11)
                if (u < c0) {
                    // delta electron, p_ionization_tabulated::post_step mqi_p_ionization.hpp:425-477
                    const Rel r1 = rel_make(ke1);
                    float     Te;
                    while (true) {
                        const float n = rng_uniform(rng);
                        Te = T_cut * r1.Te_max / ((1.0f - n) * r1.Te_max + n * T_cut);
                        if (rng_uniform(rng) < 1.0f - r1.beta_sq * Te / r1.Te_max + Te * Te / (2.0f * r1.Et * r1.Et)) break;
                    }
                    if (VARIANT == MQI_K_DEBUG) res.te_debug = Te;   // carried by a zero-energy daughter
                    else res.dE += Te;
                    ke1 -= Te;
                } else if (u < c0 + c1) {
                    // p-p elastic, pp_elastic_tabulated::post_step mqi_pp_elastic.hpp:119-219
                    const Rel   r1   = rel_make(ke1);
                    const float minv = kTpCut / r1.Ek;
                    const float uu   = rng_uniform(rng) * (1.0f - 2.0f * minv) + minv;
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
                    const float phi = kTwoPi * rng_uniform(rng);
                    ke1 -= dE;
                    rotate_direction(d1x, d1y, d1z, th3, phi);
                    // recoil proton: direction rotated from the already scattered primary direction
                    float sx = d1x, sy = d1y, sz = d1z;
                    rotate_direction(sx, sy, sz, th4, phi);
                    push_secondary<VARIANT>(P, stack, sp, p1x, p1y, p1z, sx, sy, sz, dE, 0.f, 0.f, n_sec, n_ovf);
                } else if (u < c0 + c1 + c2) {
                    // p-O elastic, po_elastic::post_step mqi_po_elastic.hpp:97-217
                    const Rel r1 = rel_make(ke1);
                    if (r1.Ek <= 5.5f) {
                        const float dE = r1.Ek;
                        if (VARIANT == MQI_K_DEBUG)
                            push_secondary<VARIANT>(P, stack, sp, p1x, p1y, p1z, d1x, d1y, d1z, dE, -dE, dE, n_sec, n_ovf);
                        else res.local_dE += dE;
                        ke1 -= dE;
                        stopped = true;
                    } else {
                        const float Tp_avg = 0.65f * expf(-0.0013f * r1.Ek) - 0.71f * expf(-0.0177f * r1.Ek);
                        const float Tp_max = (2.0f * kMo * r1.beta_sq * r1.gamma_sq) /
                                             (1.0f + 2.0f * r1.gamma * kMoMp + kMoMp * kMoMp);
                        float dE;
                        do {   // mqi_exponential (GPU definition, truncated) base/mqi_math.hpp:298-307
                            dE = -Tp_avg * logf(1.0f - rng_uniform(rng));
                        } while (dE > Tp_max || dE != dE);
                        const float E1 = r1.Ek * (r1.Ek + 2.0f * kMp);
                        const float E3 = (r1.Ek - dE) * (r1.Ek - dE + 2.0f * kMp);
                        float cos_th3  = (E1 + E3 - dE * (dE + 2.0f * kMo)) / 2.0f / sqrtf(E1 * E3);
                        cos_th3        = fminf(1.f, fmaxf(-1.f, cos_th3));
                        const float th3 = acosf(cos_th3);
                        const float phi = kTwoPi * rng_uniform(rng);
                        if (VARIANT == MQI_K_DEBUG)   // daughter starts at the parent's PRE-step vertex
                            push_secondary<VARIANT>(P, stack, sp, px, py, pz, dx, dy, dz, dE, -dE, dE, n_sec, n_ovf);
                        else res.local_dE += dE;
                        ke1 -= dE;
                        rotate_direction(d1x, d1y, d1z, th3, phi);
                    }
                } else if (u < c0 + c1 + c2 + c3) {
                    // p-O inelastic cascade, po_inelastic_tabulated::post_step mqi_po_inelastic.hpp:159-288
                    const float Ek = ke1;
                    float       Eb = 5.0f, Er = Ek;
                    float       prob_2nd, prob_long, power;
                    if (Ek <= 215.f && Ek > 200.f) { prob_2nd = 0.78f; prob_long = prob_2nd + (1.f - prob_2nd) * 0.9f; power = 0.4f; }
                    else if (Ek > 215.f) { prob_2nd = 0.78f; prob_long = prob_2nd + (1.f - prob_2nd) * 1.0f; power = 0.4f; }
                    else if (Ek <= 200.f && Ek > 150.f) { prob_2nd = 0.72f; prob_long = prob_2nd + (1.f - prob_2nd) * 0.83f; power = 0.45f; }
                    else { prob_2nd = 0.7f; prob_long = prob_2nd + (1.f - prob_2nd) * 0.83f; power = 0.52f; }
                    while ((Er - Eb) > 2.0f) {
                        Er -= Eb;
                        const float uu = rng_uniform(rng);
                        float       dE = powf(uu, power) * (Er - 2.0f) + 2.0f;
                        if (dE >= Er) dE = Er;
                        Er -= dE;
                        ke1 -= (dE + Eb);
                        const float zeta = rng_uniform(rng);
                        if (zeta < prob_2nd) {
                            float cos_th = (2.0f * dE / Ek - 1.0f) + 2.0f * (1.f - dE / Ek) * rng_uniform(rng);
                            cos_th       = fminf(1.f, fmaxf(-1.f, cos_th));
                            const float th  = acosf(cos_th);
                            const float phi = kTwoPi * rng_uniform(rng);
                            float sx = d1x, sy = d1y, sz = d1z;
                            rotate_direction(sx, sy, sz, th, phi);
                            push_secondary<VARIANT>(P, stack, sp, p1x, p1y, p1z, sx, sy, sz, dE, 0.f, 0.f, n_sec, n_ovf);
                        } else if (zeta < prob_long) {
                            // neutral / long-range: energy leaves
                        } else if (VARIANT == MQI_K_DEBUG) {
                            push_secondary<VARIANT>(P, stack, sp, px, py, pz, dx, dy, dz, dE, -dE, dE, n_sec, n_ovf);
                        }   // release: short-range energy dropped (:273-275, B4)
                        Eb *= 0.65f;
                    }
                    res.dE += Er;
                    ke1 -= Er;
                    stopped = true;
                }
            }
        }

        // ------------------------------------------------------------------ scoring, :204-225
        if (rho >= 1.0e-7f && rho <= 99.9f) {
            const float inv_vol = 1.0f / ((ex1 - ex0) * (ey1 - ey0) * (ez1 - ez0));
            score_step<VARIANT>(P, M, cnb, spot_ind, inv_vol, rsp0, res);
        }

        // ------------------------------------------------------------------ advance, :227-232
        if (stopped) {
            alive = false;
        } else {
            ix = index_update_axis(ex0, ex1, p1x, d1x, ix);
            iy = index_update_axis(ey0, ey1, p1y, d1y, iy);
            iz = index_update_axis(ez0, ez1, p1z, d1z, iz);
            px = p1x; py = p1y; pz = p1z;
            dx = d1x; dy = d1y; dz = d1z;
            ke = ke1;
            ke1_off = 0.f;
            dE_pre  = 0.f;
            if (ix < 0 || iy < 0 || iz < 0 || ix >= nx || iy >= ny || iz >= nz) alive = false;
        }
    }

    // per-lane counters -> global (one atomic per lane per launch)
    if (n_done) atomicAdd(P.counters + C_DONE, n_done);
    if (n_steps) atomicAdd(P.counters + C_STEPS, n_steps);
    if (n_sec) atomicAdd(P.counters + C_SECONDARIES, n_sec);
    if (n_ovf) atomicAdd(P.counters + C_OVERFLOW, n_ovf);
}


// =============================================================================================
// auxiliary kernels
// =============================================================================================
__global__ void
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
dev_rsp_kernel(const MatEntry* __restrict__ m, const float* __restrict__ ek, size_t n, float* rsp, float* rl) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        rsp[i] = rsp_eval(m[i], ek[i]);
        rl[i]  = 1.0f / m[i].inv_x0;
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
        dist[i] = grid_entry(xe, ye, ze, g.nx, g.ny, g.nz, p, d, c);
        cell[3 * i] = c[0]; cell[3 * i + 1] = c[1]; cell[3 * i + 2] = c[2];
    }
}

__global__ void
dev_hash_kernel(const uint32_t* k1, const uint32_t* k2, const unsigned long long* cap, size_t n, uint32_t* out) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        out[i] = hash_fun(k1[i], k2[i], cap[i]);
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
        Rng            rng;
        rng_init(rng, seed, h);
        VertexDev v;
        sample_vertex(src.beamlets[spot], rng, v);
        out[i]      = v;
        spot_out[i] = spot;
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
transport_smem_bytes(int nx, int ny, int nz) {
    return 2 * kTableN * sizeof(float4) + (size_t) (nx + ny + nz + 3) * sizeof(float);
}

template<int V>
static cudaError_t
prep_transport(size_t smem) {
    return cudaFuncSetAttribute(transport_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
}

cudaError_t
transport_occupancy(int variant, size_t smem, int* blocks_per_sm) {
    cudaError_t e = variant == MQI_K_DEBUG ? prep_transport<MQI_K_DEBUG>(smem) : prep_transport<MQI_K_RELEASE>(smem);
    if (e != cudaSuccess) return e;
    if (variant == MQI_K_DEBUG)
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, transport_kernel<MQI_K_DEBUG>, MQI_K_BLOCK, smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, transport_kernel<MQI_K_RELEASE>, MQI_K_BLOCK, smem);
}

cudaError_t
launch_transport(const Params& p, int variant, int grid, size_t smem, cudaStream_t st) {
    if (variant == MQI_K_DEBUG) transport_kernel<MQI_K_DEBUG><<<grid, MQI_K_BLOCK, smem, st>>>(p);
    else transport_kernel<MQI_K_RELEASE><<<grid, MQI_K_BLOCK, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t
launch_hu_to_material(const int16_t* d_hu, uint16_t* d_mat, size_t n, cudaStream_t st) {
    hu_to_material_kernel<<<grid_for(n / 8 + 1), 256, 0, st>>>(d_hu, d_mat, n);
    return cudaGetLastError();
}
cudaError_t
launch_hu_to_density(const int16_t* d_hu, float* d_rho, size_t n, const float* d_correction, float scale, cudaStream_t st) {
    hu_to_density_kernel<<<grid_for(n), 256, 0, st>>>(d_hu, d_rho, n, d_correction, scale);
    return cudaGetLastError();
}
cudaError_t
launch_dev_rsp(const MatEntry* m, const float* ek, size_t n, float* rsp, float* rl, cudaStream_t st) {
    dev_rsp_kernel<<<grid_for(n), 256, 0, st>>>(m, ek, n, rsp, rl);
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
