/* mqi_oracle.c -- CPU restatement of moqui's per-history proton transport (see mqi_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Plain C99, fp32 state exactly as the reference (R = float) with the
 * reference's implicit double promotions kept (double literals stay double literals).
 * Build: gcc -O2 -ffp-contract=off -std=c99 -shared -fPIC mqi_oracle.c -lm
 * Citations are relative to /root/reference/moqui.
 */
#include "mqi_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------- */
/* constants: base/mqi_physics_constants.hpp:16-38, base/mqi_math.hpp:17-24                      */
/* ------------------------------------------------------------------------------------------- */
static const float k_near_zero          = 1e-7;
static const float k_geometry_tolerance = 1e-3;
#define K_EMPTY_PAIR 0xffffffffu

static const float k_Mp    = 938.272046;
static const float k_Me    = 0.510998928;
static const float k_Mo    = 14903.3460795634;
static const float k_X0w   = 36.0863 * 10.0; /* radiation_length_water = 36.0863 * cm */
static const float k_Tp_cut = 0.5;

static float k_water_density; /* 1.0 / cm3 */
static float k_Mp_sq, k_MoMp, k_MoMp_sq, k_dedx_term0;

/* tables: base/mqi_p_ionization.hpp:11,74,145; mqi_pp_elastic.hpp:11; mqi_po_elastic.hpp:12;
 * mqi_po_inelastic.hpp:11; materials/mqi_patient_materials.hpp:11.  Loaded from the data file that
 * oracle/ref_kat.cpp dumped from the reference headers (never copied as source). */
static float t_cs_pion[600], t_pw[600], t_range[600], t_pp[600], t_poe[600], t_poi[600];
static float t_density_correction[3996];
static int   g_tables_loaded = 0;

static void
init_constants(void) {
    const float cm  = 10.0;
    const float cm3 = cm * cm * cm;
    k_water_density = 1.0 / cm3;
    k_Mp_sq         = k_Mp * k_Mp;
    k_MoMp          = k_Mo / k_Mp;
    k_MoMp_sq       = k_MoMp * k_MoMp;
    {
        const float re             = 2.8179403262e-12 * 1.0f;
        const float re_sq          = re * re;
        const float two_pi_re2_mc2 = 2.0 * M_PI * re_sq * k_Me;
        k_dedx_term0               = two_pi_re2_mc2 * 3.3428e+23 / cm3;
    }
}

int
mqo_load_tables(const char* path) {
    FILE* f = fopen(path, "rb");
    char  magic[8];
    uint32_t n[2];
    if (!f) return -1;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "MQITBL1", 7) != 0) { fclose(f); return -2; }
    if (fread(n, 4, 2, f) != 2 || n[0] != 600 || n[1] != 3996) { fclose(f); return -3; }
    if (fread(t_cs_pion, 4, 600, f) != 600 || fread(t_pw, 4, 600, f) != 600 ||
        fread(t_range, 4, 600, f) != 600 || fread(t_pp, 4, 600, f) != 600 ||
        fread(t_poe, 4, 600, f) != 600 || fread(t_poi, 4, 600, f) != 600 ||
        fread(t_density_correction, 4, 3996, f) != 3996) {
        fclose(f);
        return -4;
    }
    fclose(f);
    init_constants();
    g_tables_loaded = 1;
    return 0;
}

/* base/mqi_math.hpp:26-30 */
static inline float
intpl1d(float x, float x0, float x1, float y0, float y1) {
    return (x1 == x0) ? y0 : y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}

/* ------------------------------------------------------------------------------------------- */
/* calibration: materials/mqi_patient_materials.hpp                                              */
/* ------------------------------------------------------------------------------------------- */
/* hu_to_density :514-542 */
float
mqo_hu_to_density(int16_t hu) {
    float rho_mass = 0.0;
    if (hu < -1000) {
        hu = -1000;
    } else if (hu > 2995) {
        hu = 2995;
    }
    if (hu >= -1000 && hu < -98) {
        rho_mass = 0.00121 + 0.001029700665188 * (1000.0 + hu);
    } else if (hu >= -98 && hu < 15) {
        rho_mass = 1.018 + 0.000893 * hu;
    } else if (hu >= 15 && hu < 23) {
        rho_mass = 1.03;
    } else if (hu >= 23 && hu < 101) {
        rho_mass = 1.003 + 0.001169 * hu;
    } else if (hu >= 101 && hu < 2001) {
        rho_mass = 1.017 + 0.000592 * hu;
    } else if (hu >= 2001 && hu < 2995) {
        rho_mass = 2.201 + 0.0005 * (-2000.0 + hu);
    } else {
        rho_mass = 4.54;
    }
    {
        float correction_factor = t_density_correction[hu + 1000];
        rho_mass *= correction_factor;
    }
    rho_mass /= 1000.0;
    return rho_mass;
}

/* spr_default :414-450 */
float
mqo_spr(float rho_mass, float Ek, int variant) {
    float density_tmp = rho_mass * 1000.0;
    if (variant == MQO_VARIANT_DEBUG) {
        if (fabs(density_tmp - 1.0) < 1e-3) { /* mqi_abs<double> on a double expression */
            return 1.0;
        }
    }
    if (density_tmp <= 0.26) {
        if (density_tmp < 0.0012) {
            return 0.0;
        } else {
            return intpl1d(density_tmp, 0.0012, 0.26, 0.8815, 0.9925);
        }
    } else {
        float rsp = 1.0123 - 3.386e-5 * Ek;
        rsp += 0.291 * (1.0 + powf(Ek, (float) (-0.3421))) * (powf(density_tmp, (float) (-0.7)) - 1.0);
        if (density_tmp >= 0.9) {
            return rsp;
        } else {
            return intpl1d(density_tmp, 0.26, 0.9, 0.9925, rsp);
        }
    }
}

/* radiation_length_default :451-473 (water_density, radiation_length_water are the callers'
 * constants, mqi_p_ionization.hpp:381-382) */
float
mqo_radiation_length(float rho_mass, int variant) {
    float radiation_length_mat = 0.0;
    float f                    = 0.0;
    float density              = rho_mass * 1000.0;
    if (variant == MQO_VARIANT_DEBUG) {
        if (fabs(density - 1.0) < 1e-3) { return k_X0w; }
    }
    if (density <= 0.26) {
        f = 0.9857 + 0.0085 * density;
    } else if (density > 0.26 && density <= 0.9) {
        f = 1.0446 - 0.2180 * density;
    } else if (density > 0.9) {
        f = 1.19 + 0.44 * log(density - 0.44); /* mqi_ln<double>: the argument is a double expression */
    }
    radiation_length_mat = (k_water_density * k_X0w) / (density * 0.001 * f);
    return radiation_length_mat;
}

/* ------------------------------------------------------------------------------------------- */
/* hash + job split                                                                              */
/* ------------------------------------------------------------------------------------------- */
/* kernel_functions/mqi_transport.hpp:32-51 */
uint32_t
mqo_hash(uint32_t k1, uint32_t k2, uint64_t max_capacity) {
    k1 *= 0xcc9e2d5;
    k1 = (k1 << 15) | (k1 >> 17);
    k1 *= 0x1b873593;
    k2 ^= k1;
    k2 = (k2 << 13) | (k2 >> 19);
    k2 *= 5;
    k2 += 0xe6546b64;
    k2 ^= 4;
    k2 ^= k2 >> 16;
    k2 *= 0x85ebca6b;
    k2 ^= k2 >> 13;
    k2 *= 0xc2b2ae35;
    k2 ^= k2 >> 16;
    return (uint32_t) (k2 % (max_capacity));
}

/* base/mqi_utils.hpp:138-146 */
void
mqo_start_and_length(uint32_t n_threads, uint32_t n_jobs, uint32_t thread_id, uint32_t out[2]) {
    uint32_t quotient  = n_jobs / n_threads;
    uint32_t remainder = n_jobs % n_threads;
    out[0] = quotient * thread_id + ((thread_id >= remainder) ? remainder : thread_id);
    out[1] = quotient + 1 * (thread_id < remainder);
}

/* ------------------------------------------------------------------------------------------- */
/* vectors / rotation: base/mqi_vec.hpp, base/mqi_matrix.hpp                                     */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    float x, y, z;
} v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float v3_norm(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* mqi_vec.hpp:207-212 */
static inline v3 v3_normalize(v3 a) {
    float n = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3_make(a.x / n, a.y / n, a.z / n);
}
/* mqi_vec.hpp cross */
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

typedef struct {
    float xx, xy, xz, yx, yy, yz, zx, zy, zz;
} m33;

static inline v3
m33_mul(const m33* m, v3 r) {
    return v3_make(m->xx * r.x + m->xy * r.y + m->xz * r.z, m->yx * r.x + m->yy * r.y + m->yz * r.z,
                   m->zx * r.x + m->zy * r.y + m->zz * r.z);
}

/* mat3x3(0, theta, phi): mqi_matrix.hpp:69-76 with rotate_y :230-247, rotate_z :249-272 */
static m33
m33_from_y_z(float b, float c) {
    m33 m = { 1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0 };
    if (b != 0) {
        float c1 = cosf(b), s1 = sinf(b);
        float x1 = m.zx, y1 = m.zy, z1 = m.zz;
        m.zx = c1 * x1 - s1 * m.xx;
        m.zy = c1 * y1 - s1 * m.xy;
        m.zz = c1 * z1 - s1 * m.xz;
        m.xx = s1 * x1 + c1 * m.xx;
        m.xy = s1 * y1 + c1 * m.xy;
        m.xz = s1 * z1 + c1 * m.xz;
    }
    if (c != 0) {
        float c1 = cosf(c), s1 = sinf(c);
        float x1 = m.xx, y1 = m.xy, z1 = m.xz;
        m.xx = c1 * x1 - s1 * m.yx;
        m.xy = c1 * y1 - s1 * m.yy;
        m.xz = c1 * z1 - s1 * m.yz;
        m.yx = s1 * x1 + c1 * m.yx;
        m.yy = s1 * y1 + c1 * m.yy;
        m.yz = s1 * z1 + c1 * m.yz;
    }
    return m;
}

/* mat3x3(f, t): rotation aligning f to t, mqi_matrix.hpp:88-150 */
static m33
m33_align(v3 f, v3 t) {
    m33   m;
    v3    v = v3_cross(f, t);
    float c = v3_dot(f, t) / (v3_norm(f) * v3_norm(t));
    float h = 0;
    if (fabsf(c - 1) < k_geometry_tolerance || fabsf(c + 1) < k_geometry_tolerance) {
        v3    x  = v3_make(1, 0, 0);
        v3    uu = v3_normalize(v3_sub(x, f));
        v3    vv = v3_normalize(v3_sub(x, t));
        float dot_u  = v3_dot(uu, uu);
        float dot_v  = v3_dot(vv, vv);
        float dot_uv = v3_dot(vv, uu);
        m.xx = 1 - 2 / dot_u * uu.x * uu.x - 2 / dot_v * vv.x * vv.x + 4 * dot_uv / (dot_u * dot_v) * vv.x * uu.x;
        m.xy = 0 - 2 / dot_u * uu.x * uu.y - 2 / dot_v * vv.x * vv.y + 4 * dot_uv / (dot_u * dot_v) * vv.x * uu.y;
        m.xz = 0 - 2 / dot_u * uu.x * uu.z - 2 / dot_v * vv.x * vv.z + 4 * dot_uv / (dot_u * dot_v) * vv.x * uu.z;
        m.yx = 0 - 2 / dot_u * uu.y * uu.x - 2 / dot_v * vv.y * vv.x + 4 * dot_uv / (dot_u * dot_v) * vv.y * uu.x;
        m.yy = 1 - 2 / dot_u * uu.y * uu.y - 2 / dot_v * vv.y * vv.y + 4 * dot_uv / (dot_u * dot_v) * vv.y * uu.y;
        m.yz = 0 - 2 / dot_u * uu.y * uu.z - 2 / dot_v * vv.y * vv.z + 4 * dot_uv / (dot_u * dot_v) * vv.y * uu.z;
        m.zx = 0 - 2 / dot_u * uu.z * uu.x - 2 / dot_v * vv.z * vv.x + 4 * dot_uv / (dot_u * dot_v) * vv.z * uu.x;
        m.zy = 0 - 2 / dot_u * uu.z * uu.y - 2 / dot_v * vv.z * vv.y + 4 * dot_uv / (dot_u * dot_v) * vv.z * uu.y;
        m.zz = 1 - 2 / dot_u * uu.z * uu.z - 2 / dot_v * vv.z * vv.z + 4 * dot_uv / (dot_u * dot_v) * vv.z * uu.z;
    } else {
        h    = (1.0 - c) / (1.0 - c * c);
        m.xx = c + h * v.x * v.x;
        m.xy = h * v.x * v.y - v.z;
        m.xz = h * v.x * v.z + v.y;
        m.yx = h * v.x * v.y + v.z;
        m.yy = c + h * v.y * v.y;
        m.yz = h * v.y * v.z - v.x;
        m.zx = h * v.x * v.z - v.y;
        m.zy = h * v.y * v.z + v.x;
        m.zz = c + h * v.z * v.z;
    }
    return m;
}

/* track_t::update_post_vertex_direction base/mqi_track.hpp:163-172 */
static v3
rotate_direction(v3 dir, float theta, float phi) {
    const v3 ref_vector = { 0, 0, 1 };
    m33      m_local    = m33_from_y_z(theta, phi);
    v3       d_local    = v3_normalize(m33_mul(&m_local, ref_vector));
    m33      m_global   = m33_align(ref_vector, dir);
    return v3_normalize(m33_mul(&m_global, d_local));
}

void
mqo_rotate_direction(const float dir_in[3], float theta, float phi, float dir_out[3]) {
    v3 r       = rotate_direction(v3_make(dir_in[0], dir_in[1], dir_in[2]), theta, phi);
    dir_out[0] = r.x;
    dir_out[1] = r.y;
    dir_out[2] = r.z;
}

/* ------------------------------------------------------------------------------------------- */
/* geometry: base/mqi_grid3d.hpp                                                                 */
/* ------------------------------------------------------------------------------------------- */
/* one axis of index(p, dir) :745-844 (linear scan with boundary tie rules) */
static int
index_axis(const float* e, int dim, float p, float dir) {
    int idx = 0; /* the reference leaves it unset if dim == 0 */
    int ind;
    for (ind = 0; ind < dim; ind++) {
        if (fabsf(e[ind] - p) < k_geometry_tolerance) {
            if (dir > 0) { idx = ind; break; }
            else if (dir < 0) { idx = ind - 1; break; }
            else { idx = ind; break; }
        } else if (fabsf(e[ind + 1] - p) < k_geometry_tolerance) {
            if (dir > 0) { idx = ind + 1; break; }
            else if (dir < 0) { idx = ind; break; }
            else { idx = ind; break; }
        } else if (e[ind] - p < 0 && e[ind + 1] - p > 0) {
            idx = ind;
            break;
        } else {
            idx = -1;
        }
    }
    return idx;
}

void
mqo_grid_index(const mqo_grid* g, const float p[3], const float d[3], int cell[3]) {
    cell[0] = index_axis(g->xe, g->nx, p[0], d[0]);
    cell[1] = index_axis(g->ye, g->ny, p[1], d[1]);
    cell[2] = index_axis(g->ze, g->nz, p[2], d[2]);
}

/* index(vtx1, dir1, idx) :846-877 (incremental cell update after a step) */
static void
index_update_axis(const float* e, float v, float dir, int* idx) {
    if (dir < 0 && (fabsf(v - e[*idx]) < k_geometry_tolerance || v < e[*idx])) {
        *idx -= 1;
    } else if (dir > 0 && (fabsf(v - e[*idx + 1]) < k_geometry_tolerance || v > e[*idx + 1])) {
        *idx += 1;
    }
}

void
mqo_grid_index_update(const mqo_grid* g, const float p[3], const float d[3], int cell[3]) {
    index_update_axis(g->xe, p[0], d[0], &cell[0]);
    index_update_axis(g->ye, p[1], d[1], &cell[1]);
    index_update_axis(g->ze, p[2], d[2], &cell[2]);
}

static inline int
grid_is_valid(const mqo_grid* g, const int c[3]) { /* :881-889 */
    if (c[0] < 0 || c[1] < 0 || c[2] < 0) return 0;
    if (c[0] >= g->nx || c[1] >= g->ny || c[2] >= g->nz) return 0;
    return 1;
}

/* one axis of intersect(p, d, idx) :528-605; may zero *d (in place, as the reference does) */
static float
cell_tmax_axis(const float* e, int dim, float p, float* d, int idx) {
    float vox1 = e[idx], vox2 = e[idx + 1];
    float me = *d; /* d.dot(n100_) with the unit axis */
    float t_max;
    if (me * me > k_near_zero) {
        if (me < 0) {
            if (fabsf(-(p - vox1) / *d) < k_geometry_tolerance && idx > 0) {
                t_max = 1 / k_geometry_tolerance;
            } else {
                t_max = -(p - vox1) / *d;
            }
        } else {
            if (fabsf((vox2 - p) / *d) < k_geometry_tolerance && idx < dim) {
                t_max = 1 / k_geometry_tolerance;
            } else {
                t_max = (vox2 - p) / *d;
            }
        }
    } else {
        *d    = 0;
        t_max = HUGE_VALF;
    }
    return t_max;
}

/* intersect(p, d, idx) :490-626 -> distance to the exit of the current voxel, or -1 */
float
mqo_grid_intersect_cell(const mqo_grid* g, const float p[3], float d[3], const int cell[3]) {
    float tx = cell_tmax_axis(g->xe, g->nx, p[0], &d[0], cell[0]);
    float ty = cell_tmax_axis(g->ye, g->ny, p[1], &d[1], cell[1]);
    float tz = cell_tmax_axis(g->ze, g->nz, p[2], &d[2], cell[2]);
    float u_max;
    if (tx < ty) {
        u_max = (tx < tz) ? tx : tz;
    } else {
        u_max = (ty < tz) ? ty : tz;
    }
    if (u_max > 0) return u_max;
    return -1.0;
}

/* intersect(p, d) :631-743 -> entry distance from outside (0 if already inside), cell at entry;
 * -1 and cell = (-1,-1,-1) on a miss */
float
mqo_grid_intersect_entry(const mqo_grid* g, const float p[3], float d[3], int cell[3]) {
    float t_min[3], t_max[3];
    const float lo[3] = { g->xe[0], g->ye[0], g->ze[0] };
    const float hi[3] = { g->xe[g->nx], g->ye[g->ny], g->ze[g->nz] };
    float       u_min, u_max;
    int         a;
    if (p[0] >= lo[0] && p[0] <= hi[0] && p[1] >= lo[1] && p[1] <= hi[1] && p[2] >= lo[2] && p[2] <= hi[2]) {
        mqo_grid_index(g, p, d, cell);
        return 0;
    }
    cell[0] = cell[1] = cell[2] = -1;
    for (a = 0; a < 3; ++a) {
        float me = d[a];
        if (me * me > k_near_zero) {
            if (me > 0) {
                t_min[a] = (lo[a] - p[a]) / d[a];
                t_max[a] = (hi[a] - p[a]) / d[a];
            } else {
                t_max[a] = (lo[a] - p[a]) / d[a];
                t_min[a] = (hi[a] - p[a]) / d[a];
            }
        } else {
            d[a]     = 0;
            t_min[a] = -HUGE_VALF;
            t_max[a] = HUGE_VALF;
        }
    }
    if (t_min[0] > t_min[1]) {
        u_min = (t_min[0] > t_min[2]) ? t_min[0] : t_min[2];
    } else {
        u_min = (t_min[1] > t_min[2]) ? t_min[1] : t_min[2];
    }
    if (t_max[0] < t_max[1]) {
        u_max = (t_max[0] < t_max[2]) ? t_max[0] : t_max[2];
    } else {
        u_max = (t_max[1] < t_max[2]) ? t_max[1] : t_max[2];
    }
    if ((u_min < u_max || fabsf(u_min - u_max) < k_geometry_tolerance) && u_min >= 0 && u_max >= 0) {
        float p_on[3];
        p_on[0] = p[0] + d[0] * u_min;
        p_on[1] = p[1] + d[1] * u_min;
        p_on[2] = p[2] + d[2] * u_min;
        mqo_grid_index(g, p_on, d, cell);
        return u_min;
    }
    return -1.0;
}

/* ------------------------------------------------------------------------------------------- */
/* relativistic quantities + tabulated physics                                                   */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    float beta_sq, gamma_sq, gamma, Et, Et_sq, Ek, mc2, tau, Te_max;
} relq;

/* base/mqi_relativistic_quantities.hpp:27-44 */
static relq
rel_make(float kinetic_energy) {
    relq        r;
    const float Mp      = 938.272046;
    const float Me      = 0.510998928;
    const float MeMp    = Me / Mp;
    const float MeMp_sq = MeMp * MeMp;
    r.Ek       = kinetic_energy;
    r.mc2      = k_Mp;
    r.Et       = r.Ek + Mp;
    r.Et_sq    = r.Et * r.Et;
    r.gamma    = r.Et / Mp;
    r.gamma_sq = r.gamma * r.gamma;
    r.beta_sq  = 1.0 - 1.0 / r.gamma_sq;
    r.Te_max   = (2.0 * Me * r.beta_sq * r.gamma_sq);
    r.Te_max /= (1.0 + 2.0 * r.gamma * MeMp + MeMp_sq);
    r.tau = r.Ek / Mp;
    return r;
}
static inline float
rel_momentum(const relq* r) { /* :46-49 */
    return sqrtf(r->Et * r->Et - r->mc2 * r->mc2);
}

static const float pi_Ei = 0.1, pi_Ef = 299.6, pi_step = 0.5; /* mqi_fippel_physics.hpp:30-35 */

static inline int clamp_idx(int i) { return i < 0 ? 0 : (i > 599 ? 599 : i); }

/* p_ionization_tabulated::cross_section mqi_p_ionization.hpp:254-268 */
static float
cs_pion(const relq* rel, float rho_mass) {
    float cs = 0;
    if (rel->Ek >= pi_Ei && rel->Ek <= pi_Ef) {
        uint16_t idx0 = (uint16_t) ((rel->Ek - pi_Ei) / pi_step);
        uint16_t idx1 = idx0 + 1;
        float    x0   = pi_Ei + idx0 * pi_step;
        float    x1   = x0 + pi_step;
        cs            = intpl1d(rel->Ek, x0, x1, t_cs_pion[idx0], t_cs_pion[clamp_idx(idx1)]);
    }
    cs *= rho_mass;
    return cs;
}

/* p_ionization_tabulated::dEdx :271-286 (negative) */
static float
dEdx(const relq* rel) {
    float pw = 0;
    if (rel->Ek >= pi_Ei && rel->Ek <= pi_Ef) {
        uint16_t idx0 = (uint16_t) ((rel->Ek - pi_Ei) / pi_step);
        uint16_t idx1 = idx0 + 1;
        float    x0   = pi_Ei + idx0 * pi_step;
        float    x1   = x0 + pi_step;
        pw            = intpl1d(rel->Ek, x0, x1, t_pw[idx0], t_pw[clamp_idx(idx1)]);
    } else if (rel->Ek < pi_Ei && rel->Ek > 0) {
        pw = t_pw[0];
    }
    return -1.0 * pw;
}

/* pp_elastic_tabulated::cross_section mqi_pp_elastic.hpp:221-235 (same for po_e :243-256, po_i :141-155) */
static float
cs_nuclear(const float* table, const relq* rel, float rho_mass) {
    const float Ek_min = 0.5, Ek_max = 300.0, dEk = 0.5;
    float       cs = 0;
    if (rel->Ek >= Ek_min && rel->Ek <= Ek_max) {
        uint16_t idx0 = (uint16_t) ((rel->Ek - Ek_min) / dEk);
        uint16_t idx1 = idx0 + 1;
        float    x0   = Ek_min + idx0 * dEk;
        float    x1   = x0 + 0.5;
        cs            = intpl1d(rel->Ek, x0, x1, table[clamp_idx(idx0)], table[clamp_idx(idx1)]);
    }
    cs *= rho_mass;
    return cs;
}

void
mqo_physics_probe(float ek, float out[9]) {
    relq r = rel_make(ek);
    out[0] = r.beta_sq;
    out[1] = r.gamma;
    out[2] = r.Te_max;
    out[3] = rel_momentum(&r);
    out[4] = cs_pion(&r, 1.0e-3f);
    out[5] = dEdx(&r);
    out[6] = cs_nuclear(t_pp, &r, 1.0e-3f);
    out[7] = cs_nuclear(t_poe, &r, 1.0e-3f);
    out[8] = cs_nuclear(t_poi, &r, 1.0e-3f);
}

/* ------------------------------------------------------------------------------------------- */
/* RNG protocol: Philox4x32-7, key = seed, counter = (block, 0, history_lo, history_hi).    */
/* Seven rounds are the smallest Philox4x32 variant that passes BigCrush (Salmon et al., SC'11, Random123 */
/* philox4x32_R(7, ...)); the kernel spends 4 of its ~600 instructions per voxel step on each round.      */
/* The round function is pinned by the Random123 known-answer vectors of the 10-round generator.          */
/* ------------------------------------------------------------------------------------------- */
void
mqo_philox4x32_r(const uint32_t ctr_in[4], const uint32_t key_in[2], int rounds, uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    int      i;
    for (i = 0; i < rounds; ++i) {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
        uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void
mqo_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    mqo_philox4x32_r(ctr_in, key_in, 10, out);
}

/* Philox2x32-10 (Random123): the short generator of the delta-electron rejection loop */
void
mqo_philox2x32_10(const uint32_t ctr_in[2], uint32_t key, uint32_t out[2]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1];
    int      i;
    for (i = 0; i < 10; ++i) {
        uint64_t p  = (uint64_t) 0xD256D193u * c0;
        uint32_t n0 = (uint32_t) (p >> 32) ^ key ^ c1;
        c1          = (uint32_t) p;
        c0          = n0;
        key += 0x9E3779B9u;
    }
    out[0] = c0; out[1] = c1;
}

float
mqo_u32_to_uniform(uint32_t x) { /* open interval (0,1), 23 bits: (m + 0.5) * 2^-23 with m = x & 0x7fffff */
    uint32_t bits = 0x3f800000u | (x & 0x007fffffu);
    float    f;
    memcpy(&f, &bits, sizeof(f));
    return f - 0.99999994f; /* exact: f = 1 + m 2^-23, constant = 1 - 2^-24 */
}
/* 24-bit deviate from the top bytes of three Philox words (bits the mapping above never looks at) */
static inline float
low_bytes_to_uniform(uint32_t w0, uint32_t w1, uint32_t w2) {
    uint32_t v = (w0 >> 24) | ((w1 >> 24) << 8) | ((w2 >> 24) << 16);
    return ((float) v + 0.5f) * (1.0f / 16777216.0f);
}

typedef struct {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int      pos;
} rng_t;

static void
rng_init(rng_t* r, uint64_t seed, uint64_t history) {
    r->key[0] = (uint32_t) seed;
    r->key[1] = (uint32_t) (seed >> 32);
    r->ctr[0] = 0;
    r->ctr[1] = 0;
    r->ctr[2] = (uint32_t) history;
    r->ctr[3] = (uint32_t) (history >> 32);
    r->pos    = 4;
}
static inline void rng_begin_step(rng_t* r) { r->pos = 4; } /* discard the rest of the block */
static inline uint32_t
rng_u32(rng_t* r) {
    if (r->pos == 4) {
        mqo_philox4x32_r(r->ctr, r->key, MQO_PHILOX_ROUNDS, r->buf);
        r->ctr[0] += 1;
        r->pos = 0;
    }
    return r->buf[r->pos++];
}
static inline float rng_uniform(rng_t* r) { return mqo_u32_to_uniform(rng_u32(r)); }
/* one (n, accept) pair of the delta-electron sampler: Philox2x32-10, counter = (block, history_lo),
 * key = seed_lo ^ seed_hi * 0x85EBCA6B ^ history_hi * 0xC2B2AE35; consumes one block number */
static inline void
rng_pair2(rng_t* r, float* a, float* b) {
    uint32_t ctr[2], out[2];
    ctr[0] = r->ctr[0];
    ctr[1] = r->ctr[2];
    mqo_philox2x32_10(ctr, r->key[0] ^ (r->key[1] * 0x85EBCA6Bu) ^ (r->ctr[3] * 0xC2B2AE35u), out);
    r->ctr[0] += 1;
    r->pos = 4;
    *a = mqo_u32_to_uniform(out[0]);
    *b = mqo_u32_to_uniform(out[1]);
}
static inline void
rng_normal_pair(rng_t* r, float* z1, float* z2) { /* Box-Muller on two uniforms */
    float u1  = rng_uniform(r);
    float u2  = rng_uniform(r);
    float rad = sqrtf(-2.0f * logf(u1));
    float ang = 6.28318530717958647692f * u2;
    *z1       = rad * cosf(ang);
    *z2       = rad * sinf(ang);
}

/* ------------------------------------------------------------------------------------------- */
/* source: beamlet::operator() mqi_beamlet.hpp:81-90, phsp_6d(_uniform)::operator()              */
/* distributions/mqi_phsp6d_uniform.hpp:68-85, mqi_phsp6d.hpp:62-79, const_1d / norm_1d          */
/* ------------------------------------------------------------------------------------------- */
static void
sample_vertex(const mqo_beamlet* b, rng_t* rng, mqo_vertex* out) {
    float phsp[6];
    float Ux, Vx, Uy, Vy, Uz, za, zb, uc, ke;
    int   i;
    for (i = 0; i < 6; ++i) phsp[i] = b->mean[i];
    rng_begin_step(rng);
    if (b->phsp_uniform) {
        Ux = 2.0f * rng_uniform(rng) - 1.0f;
        Vx = 2.0f * rng_uniform(rng) - 1.0f;
        Uy = 2.0f * rng_uniform(rng) - 1.0f;
        Vy = 2.0f * rng_uniform(rng) - 1.0f;
    } else {
        rng_normal_pair(rng, &Ux, &Vx);
        rng_normal_pair(rng, &Uy, &Vy);
    }
    rng_begin_step(rng);
    rng_normal_pair(rng, &za, &zb);
    uc = rng_uniform(rng);
    Uz = b->phsp_uniform ? 2.0f * uc - 1.0f : za;
    phsp[0] += b->sigma[0] * Ux;
    phsp[1] += b->sigma[1] * Uy;
    phsp[2] += b->sigma[2] * Uz;
    phsp[3] += b->sigma[3] * (b->corr[0] * Ux + Vx * sqrt(1.0 - b->corr[0] * b->corr[0]));
    phsp[4] += b->sigma[4] * (b->corr[1] * Uy + Vy * sqrt(1.0 - b->corr[1] * b->corr[1]));
    phsp[5] = -1.0 * sqrt(1.0 - phsp[3] * phsp[3] - phsp[4] * phsp[4]);
    ke      = b->energy_normal ? zb * b->sigma_energy + b->energy : b->energy;
    out->ke = ke;
    {
        const float* R = b->rot;
        out->pos[0] = R[0] * phsp[0] + R[1] * phsp[1] + R[2] * phsp[2] + b->trans[0];
        out->pos[1] = R[3] * phsp[0] + R[4] * phsp[1] + R[5] * phsp[2] + b->trans[1];
        out->pos[2] = R[6] * phsp[0] + R[7] * phsp[1] + R[8] * phsp[2] + b->trans[2];
        out->dir[0] = R[0] * phsp[3] + R[1] * phsp[4] + R[2] * phsp[5];
        out->dir[1] = R[3] * phsp[3] + R[4] * phsp[4] + R[5] * phsp[5];
        out->dir[2] = R[6] * phsp[3] + R[7] * phsp[4] + R[8] * phsp[5];
    }
}

void
mqo_sample_vertex(const mqo_beamlet* b, uint64_t seed, uint64_t history, mqo_vertex* out) {
    rng_t rng;
    rng_init(&rng, seed, history);
    sample_vertex(b, &rng, out);
}

/* ------------------------------------------------------------------------------------------- */
/* tracks: base/mqi_track.hpp:48-205, base/mqi_track_stack.hpp:11-67                             */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    float ke;
    v3    pos, dir;
} vtx_t;

typedef struct {
    int   stopped;
    int   primary;
    vtx_t vtx0, vtx1;
    float dE, local_dE;
    int   cell[3];
    float its_dist;
} track_t;

#define STACK_MAX 200
typedef struct {
    track_t tracks[STACK_MAX];
    int     idx;
    int     limit; /* 200 with __PHYSICS_DEBUG__, else 10: mqi_track_stack.hpp:16-22 */
} tstack_t;

static inline void
stack_push(tstack_t* s, const track_t* t, mqo_stats* st) {
    if (s->idx < s->limit) { /* overflow silently drops: :36-41 */
        s->tracks[s->idx] = *t;
        ++s->idx;
        if (st) {
            st->secondaries_pushed++;
            if ((uint64_t) s->idx > st->max_stack) st->max_stack = s->idx;
        }
    }
}

typedef struct {
    const mqo_grid* g;
    int             variant;
    m33             rot_fwd, rot_inv;
    v3              trans;
    float           T_cut; /* mqi_interaction.hpp:24-28 */
    rng_t*          rng;
    tstack_t*       stk;
    mqo_stats*      st;
} ctx_t;

/* daughters are mapped with Rfwd*(x - T) + T: mqi_pp_elastic.hpp:188-195 */
static inline v3
daughter_pos(const ctx_t* c, v3 p) {
    return v3_add(m33_mul(&c->rot_fwd, v3_sub(p, c->trans)), c->trans);
}
static inline v3
daughter_dir(const ctx_t* c, v3 d) {
    return m33_mul(&c->rot_fwd, d);
}

/* p_ionization_tabulated::energy_straggling mqi_p_ionization.hpp:334-345 */
static float
energy_straggling(const relq* rel, float step_length, float rho_mass) {
    float Te   = (rel->Te_max >= 0.08511) ? 0.08511 : rel->Te_max;
    float O_sq = k_dedx_term0 * rho_mass / k_water_density * step_length;
    O_sq *= Te / rel->beta_sq * (1.0 - 0.5 * rel->beta_sq);
    return O_sq;
}

/* p_ionization_tabulated::energy_loss :298-331; z = standard normal deviate for the straggling */
static float
energy_loss(const ctx_t* c, const relq* rel, float rho_mass, float step_length, float z) {
    float    length_in_water = step_length * mqo_spr(rho_mass, rel->Ek, c->variant) * rho_mass / k_water_density;
    uint16_t n  = (uint16_t) ((rel->Ek - pi_Ei) / pi_step);
    float    x0 = pi_Ei + n * pi_step;
    float    x1 = x0 + pi_step;
    float    r, dE_mean, dE_var, ret;
    if (x0 > rel->Ek) n -= 1;
    if (x1 < rel->Ek) n += 1;
    if (n > 598) n = 598; /* keeps r_steps[n + 1] in range (Ek beyond the table is out of scope) */
    r = intpl1d(rel->Ek, x0, x1, t_range[n], t_range[n + 1]);
    if (r < length_in_water) return rel->Ek;
    r -= length_in_water;
    do { /* :318-320 with the n == 0 guard of oracle/build_ref.sh patch 2 */
        if (r >= t_range[n] || n == 0) break;
    } while (--n > 0);
    x0      = pi_Ei + n * pi_step;
    x1      = x0 + pi_step;
    dE_mean = rel->Ek - intpl1d(r, t_range[n], t_range[n + 1], x0, x1);
    dE_var  = energy_straggling(rel, length_in_water, rho_mass);
    ret     = z * sqrtf(dE_var) + dE_mean; /* mqi_normal(rng, dE_mean, sqrt(dE_var)) */
    if (ret < 0) ret *= -1.0;
    return ret;
}

/* p_ionization_tabulated::along_step :349-420 */
static void
along_step(ctx_t* c, track_t* trk, float len, float rho_mass, float z_loss, float z_theta, float u_phi) {
    relq  rel = rel_make(trk->vtx0.ke);
    float dE  = energy_loss(c, &rel, rho_mass, len, z_loss);
    float r   = 1.0;
    float P, radiation_length_mat, th_sq, th, phi;
    if (c->st) c->st->along_steps++;
    if (dE >= trk->vtx0.ke) {
        r            = trk->vtx0.ke / dE;
        trk->stopped = 1;
    }
    P                    = rel_momentum(&rel);
    radiation_length_mat = mqo_radiation_length(rho_mass, c->variant);
    th_sq = ((13.9f / P) * (13.9f / P) / rel.beta_sq) * len / radiation_length_mat;
    th    = sqrtf(th_sq);
    th    = z_theta * (sqrtf(2.0f) * th); /* mqi_normal(rng, 0, sqrt(2) * th) */
    if (th < 0) th *= -1.0;
    phi = 2.0 * M_PI * u_phi;
    trk->vtx1.dir = rotate_direction(trk->vtx1.dir, th, phi);
    trk->dE += dE * r;
    trk->vtx1.pos = v3_add(trk->vtx0.pos, v3_scale(trk->vtx0.dir, r * len));
    trk->vtx1.ke -= dE * r;
}

/* p_ionization_tabulated::last_step :482-490 */
static void
last_step(ctx_t* c, track_t* trk, float rho_mass) {
    relq  rel             = rel_make(trk->vtx0.ke);
    float length_in_water = 0;
    float step_length;
    if (trk->dE > 0) length_in_water = -trk->dE / dEdx(&rel);
    step_length   = length_in_water * k_water_density / (mqo_spr(rho_mass, trk->vtx0.ke, c->variant) * rho_mass);
    trk->vtx1.pos = v3_add(trk->vtx0.pos, v3_scale(trk->vtx0.dir, step_length));
}

/* p_ionization_tabulated::post_step (delta electron) :425-477 */
static void
delta_post_step(ctx_t* c, track_t* trk, float n_first, float acc_first) {
    relq  rel = rel_make(trk->vtx1.ke);
    float Te, n, acc;
    int   first = 1;
    if (c->st) c->st->delta_events++;
    while (1) {
        /* RNG protocol: the first try reuses the step's block (n = u / cs_delta, uniform given that the
         * delta channel was selected; accept deviate = top bytes of the block), later tries draw pairs */
        if (first) { n = n_first; acc = acc_first; first = 0; }
        else rng_pair2(c->rng, &n, &acc);
        Te = c->T_cut * rel.Te_max;
        Te /= ((1.0 - n) * rel.Te_max + n * c->T_cut);
        /* The reference accepts with probability g(Te) = 1 - b^2 Te/Tmax + Te^2/(2 Et^2) (:447-451).  g falls
         * with Te on [T_cut, Tmax], so g(T_cut) bounds it; like the CUDA path this restatement accepts with
         * g(Te) / g(T_cut): the same density from the same envelope with half the rejections (part of the
         * shared sampling protocol). */
        {
            float g_max = 1.0 - rel.beta_sq * c->T_cut / rel.Te_max + c->T_cut * c->T_cut / (2.0 * rel.Et_sq);
            if (acc * g_max < 1.0 - rel.beta_sq * Te / rel.Te_max + Te * Te / (2.0 * rel.Et_sq)) break;
        }
    }
    if (c->variant == MQO_VARIANT_DEBUG) {
        track_t d  = *trk;
        d.dE       = Te;
        d.primary  = 0;
        d.vtx0.ke  = 0;
        d.vtx1.ke  = 0;
        d.stopped  = 0;
        d.vtx0.pos = daughter_pos(c, d.vtx0.pos);
        d.vtx0.dir = daughter_dir(c, d.vtx0.dir);
        d.vtx1.pos = daughter_pos(c, d.vtx1.pos);
        d.vtx1.dir = daughter_dir(c, d.vtx1.dir);
        stack_push(c->stk, &d, c->st);
    } else {
        trk->dE += Te;
    }
    trk->vtx1.ke -= Te;
}

/* pp_elastic_tabulated::post_step mqi_pp_elastic.hpp:119-219 */
static void
pp_post_step(ctx_t* c, track_t* trk) {
    relq  rel       = rel_make(trk->vtx1.ke);
    float min_value = k_Tp_cut / rel.Ek;
    float u         = rng_uniform(c->rng) * (1.0 - 2.0 * min_value) + min_value;
    float E1 = rel.Et;
    float dE = rel.Ek * u;
    float E3 = (rel.Ek - dE) + k_Mp;
    float E4 = dE + k_Mp;
    float P1 = rel_momentum(&rel);
    float P3 = sqrtf(E3 * E3 - k_Mp_sq);
    float P4 = sqrtf(E4 * E4 - k_Mp_sq);
    float cos_th3, cos_th34, th3, th4, phi;
    track_t d;
    if (c->st) c->st->pp_events++;
    cos_th3 = E1 * E3 - k_Mp_sq - k_Mp * (E1 - E3);
    cos_th3 /= (P1 * P3);
    cos_th34 = E3 * E4 - E1 * k_Mp;
    cos_th34 /= (P3 * P4);
    if (cos_th3 > 1.0) cos_th3 = 1.0;
    else if (cos_th3 < -1.0) cos_th3 = -1.0;
    if (cos_th34 > 1.0) cos_th34 = 1.0;
    else if (cos_th34 < -1.0) cos_th34 = -1.0;
    th3 = acosf(cos_th3);
    th4 = th3 - acosf(cos_th34);
    phi = 2.0 * M_PI * rng_uniform(c->rng);
    trk->vtx1.ke -= dE;
    trk->vtx1.dir = rotate_direction(trk->vtx1.dir, th3, phi);
    d          = *trk;
    d.dE       = 0;
    d.local_dE = 0;
    d.primary  = 0;
    d.vtx0.ke  = dE;
    d.vtx1.ke  = dE;
    d.stopped  = 0;
    d.vtx1.dir = rotate_direction(d.vtx1.dir, th4, phi);
    d.vtx0.pos = daughter_pos(c, d.vtx1.pos);
    d.vtx0.dir = daughter_dir(c, d.vtx1.dir);
    d.vtx1.pos = daughter_pos(c, d.vtx1.pos);
    d.vtx1.dir = daughter_dir(c, d.vtx1.dir);
    stack_push(c->stk, &d, c->st);
}

/* mqi_exponential: the GPU definition (truncated), base/mqi_math.hpp:298-307.  The CPU definition
 * (:487-495) does not truncate and can trip assert(dE <= Tp_max); the truncated one is the
 * behaviour of the shipped (GPU) product, so both the oracle and the CUDA path use it. */
static float
rng_exponential_truncated(rng_t* rng, float avg, float up) {
    float x;
    do {
        x = -1.0 / avg * logf(1.0 - rng_uniform(rng));
    } while (x > up || isnan(x));
    return x;
}

/* po_elastic::post_step mqi_po_elastic.hpp:97-217 */
static void
poe_post_step(ctx_t* c, track_t* trk) {
    relq rel = rel_make(trk->vtx1.ke);
    if (c->st) c->st->poe_events++;
    if (rel.Ek <= 5.5) {
        float dE = rel.Ek;
        if (c->variant == MQO_VARIANT_DEBUG) {
            track_t d  = *trk;
            d.dE       = dE;
            d.primary  = 0;
            d.vtx0.ke  = dE;
            d.vtx1.ke  = 0;
            d.vtx0.pos = daughter_pos(c, d.vtx1.pos);
            d.vtx0.dir = daughter_dir(c, d.vtx1.dir);
            d.vtx1.pos = daughter_pos(c, d.vtx1.pos);
            d.vtx1.dir = daughter_dir(c, d.vtx1.dir);
            d.stopped  = 0;
            stack_push(c->stk, &d, c->st);
        } else {
            trk->local_dE += dE;
        }
        trk->vtx1.ke -= dE;
        trk->stopped = 1;
    } else {
        float Tp_avg = 0.65 * exp(-0.0013 * rel.Ek); /* mqi_exp<double>: double argument */
        float Tp_max, dE, E1, E3, cos_th3, th3, phi;
        Tp_avg -= 0.71 * exp(-0.0177 * rel.Ek);
        Tp_max = (2.0 * k_Mo * rel.beta_sq * rel.gamma_sq);
        Tp_max /= (1.0 + 2.0 * rel.gamma * k_MoMp + k_MoMp_sq);
        dE      = rng_exponential_truncated(c->rng, 1.0 / Tp_avg, Tp_max);
        E1      = rel.Ek * (rel.Ek + 2.0 * k_Mp);
        E3      = (rel.Ek - dE) * (rel.Ek - dE + 2.0 * k_Mp);
        cos_th3 = (E1 + E3 - dE * (dE + 2.0 * k_Mo)) / 2.0 / sqrtf(E1 * E3);
        if (cos_th3 > 1.0) cos_th3 = 1.0;
        if (cos_th3 < -1.0) cos_th3 = -1.0;
        th3 = acosf(cos_th3);
        phi = 2.0 * M_PI * rng_uniform(c->rng);
        if (c->variant == MQO_VARIANT_DEBUG) {
            track_t d  = *trk;
            d.dE       = dE;
            d.primary  = 0;
            d.vtx0.ke  = dE;
            d.vtx1.ke  = 0;
            d.stopped  = 0;
            d.vtx0.pos = daughter_pos(c, d.vtx0.pos);
            d.vtx0.dir = daughter_dir(c, d.vtx0.dir);
            d.vtx1.pos = daughter_pos(c, d.vtx1.pos);
            d.vtx1.dir = daughter_dir(c, d.vtx1.dir);
            stack_push(c->stk, &d, c->st);
        } else {
            trk->local_dE += dE;
        }
        trk->vtx1.ke -= dE;
        trk->vtx1.dir = rotate_direction(trk->vtx1.dir, th3, phi);
    }
}

/* po_inelastic_tabulated::post_step mqi_po_inelastic.hpp:159-288 */
static void
poi_post_step(ctx_t* c, track_t* trk) {
    const float Ek = trk->vtx1.ke;
    float       Eb = 5.0; /* E_bind :128 */
    float       Er = Ek;
    const float E_mini = 2.0, E_ratio = 0.65;
    float       Prob_2nd, Prob_long, power;
    if (c->st) c->st->poi_events++;
    if (Ek <= 215 && Ek > 200) {
        Prob_2nd  = 0.78;
        Prob_long = Prob_2nd + (1 - Prob_2nd) * 0.9;
        power     = 0.4;
    } else if (Ek > 215) {
        Prob_2nd  = 0.78;
        Prob_long = Prob_2nd + (1 - Prob_2nd) * 1.0;
        power     = 0.4;
    } else if (Ek <= 200 && Ek > 150) {
        Prob_2nd  = 0.72;
        Prob_long = Prob_2nd + (1 - Prob_2nd) * 0.83;
        power     = 0.45;
    } else {
        Prob_2nd  = 0.7;
        Prob_long = Prob_2nd + (1 - Prob_2nd) * 0.83;
        power     = 0.52;
    }
    while ((Er - Eb) > E_mini) {
        float u, dE, zeta;
        Er -= Eb;
        u  = rng_uniform(c->rng);
        dE = powf(u, power) * (Er - E_mini) + E_mini;
        if (dE >= Er) dE = Er;
        Er -= dE;
        trk->vtx1.ke -= (dE + Eb);
        zeta = rng_uniform(c->rng);
        if (zeta < Prob_2nd) {
            float   cos_th = (2.0 * dE / Ek - 1.0) + 2.0 * (1 - dE / Ek) * rng_uniform(c->rng);
            float   th, phi;
            track_t d;
            if (cos_th < -1) cos_th = -1;
            if (cos_th > 1) cos_th = 1;
            th  = acosf(cos_th);
            phi = 2.0 * M_PI * rng_uniform(c->rng);
            d          = *trk;
            d.dE       = 0;
            d.local_dE = 0;
            d.primary  = 0;
            d.vtx0.ke  = dE;
            d.vtx1.ke  = dE;
            d.stopped  = 0;
            d.vtx1.dir = rotate_direction(d.vtx1.dir, th, phi);
            d.vtx0.pos = daughter_pos(c, d.vtx1.pos);
            d.vtx0.dir = daughter_dir(c, d.vtx1.dir);
            d.vtx1.pos = daughter_pos(c, d.vtx1.pos);
            d.vtx1.dir = daughter_dir(c, d.vtx1.dir);
            stack_push(c->stk, &d, c->st);
        } else if (zeta < Prob_long) {
            /* long-range (neutral) energy leaves the geometry */
        } else {
            if (c->variant == MQO_VARIANT_DEBUG) {
                track_t d  = *trk;
                d.dE       = dE;
                d.primary  = 0;
                d.vtx0.ke  = dE;
                d.vtx1.ke  = 0;
                d.stopped  = 0;
                d.vtx0.pos = daughter_pos(c, d.vtx0.pos);
                d.vtx0.dir = daughter_dir(c, d.vtx0.dir);
                d.vtx1.pos = daughter_pos(c, d.vtx1.pos);
                d.vtx1.dir = daughter_dir(c, d.vtx1.dir);
                stack_push(c->stk, &d, c->st);
            } /* release: short-range energy is dropped (:273-275) */
        }
        Eb *= E_ratio;
    }
    trk->dE += Er;
    trk->vtx1.ke -= Er;
    trk->stopped = 1;
}

/* fippel_physics::stepping base/mqi_fippel_physics.hpp:67-216 */
static void
stepping(ctx_t* c, track_t* trk, float rho_mass, float distance_to_boundary) {
    relq  rel, rel_de;
    float current_min_step, max_loss_energy, cs1[4], cs2[4], cs1_sum, cs2_sum, cs_sum, prob, mfp, step_limit;
    float z1, z2, u_phi;
    float* cs;
    if (rho_mass < 1.0e-7) {
        trk->vtx1.pos = v3_add(trk->vtx0.pos, v3_scale(trk->vtx0.dir, distance_to_boundary));
        return;
    } else if (rho_mass > 99.9) {
        trk->stopped = 1;
        return;
    }
    if (trk->vtx0.ke <= k_Tp_cut) {
        if (trk->vtx0.ke < 0) trk->vtx0.ke = 0;
        trk->dE += trk->vtx0.ke;
        trk->vtx1.ke -= trk->vtx0.ke;
        last_step(c, trk, rho_mass);
        trk->stopped = 1;
        return;
    }
    rel              = rel_make(trk->vtx0.ke);
    current_min_step = 1.0f; /* max_step :20 */
    current_min_step = current_min_step * mqo_spr(rho_mass, rel.Ek, c->variant) * rho_mass / k_water_density;
    max_loss_energy  = -1.0 * current_min_step * dEdx(&rel);
    cs1[0] = cs_pion(&rel, rho_mass);
    cs1[1] = cs_nuclear(t_pp, &rel, rho_mass);
    cs1[2] = cs_nuclear(t_poe, &rel, rho_mass);
    cs1[3] = cs_nuclear(t_poi, &rel, rho_mass);
    cs1_sum = cs1[0] + cs1[1] + cs1[2] + cs1[3];
    rel_de  = rel_make(trk->vtx0.ke - max_loss_energy);
    cs2[0] = cs_pion(&rel_de, rho_mass);
    cs2[1] = cs_nuclear(t_pp, &rel_de, rho_mass);
    cs2[2] = cs_nuclear(t_poe, &rel_de, rho_mass);
    cs2[3] = cs_nuclear(t_poi, &rel_de, rho_mass);
    cs2_sum = cs2[0] + cs2[1] + cs2[2] + cs2[3];
    cs_sum  = (cs1_sum >= cs2_sum) ? cs1_sum : cs2_sum;
    cs      = (cs1_sum >= cs2_sum) ? cs1 : cs2;

    /* RNG protocol: one aligned Philox4x32 block per physics step = {u_mfp, u_a, u_b, u_phi}; a
     * discrete interaction is selected with u_phi; the first delta-electron try reuses the step's block
     * (see delta_post_step), later tries draw Philox2x32 pairs;
     * nuclear interactions draw from further Philox4x32 blocks */
    rng_begin_step(c->rng);
    prob = rng_uniform(c->rng);
    rng_normal_pair(c->rng, &z1, &z2);
    u_phi = rng_uniform(c->rng);

    mfp        = -1.0f * logf(prob) / cs_sum;
    step_limit = current_min_step * k_water_density / (mqo_spr(rho_mass, rel.Ek, c->variant) * rho_mass);

    if (distance_to_boundary < mfp && distance_to_boundary < step_limit) {
        along_step(c, trk, distance_to_boundary, rho_mass, z1, z2, u_phi);
    } else if ((mfp < distance_to_boundary || fabsf(mfp - distance_to_boundary) < k_geometry_tolerance) &&
               (mfp < step_limit || fabsf(mfp - step_limit) < k_geometry_tolerance)) {
        float u;
        along_step(c, trk, mfp, rho_mass, z1, z2, u_phi);
        if (trk->vtx1.ke <= k_Tp_cut) { return; }
        /* the step's multiple-scattering deflection is discarded below (B11), so its azimuth deviate
         * u_phi is otherwise unused on this step and selects the process */
        u             = cs_sum * u_phi;
        trk->vtx1.dir = trk->vtx0.dir;
        if (u < cs[0]) {
            float n_first = u / cs[0];
            if (n_first > 1.0f) n_first = 1.0f;
            delta_post_step(c, trk, n_first, low_bytes_to_uniform(c->rng->buf[0], c->rng->buf[1], c->rng->buf[2]));
        } else if (u < (cs[0] + cs[1])) {
            pp_post_step(c, trk);
        } else if (u < (cs[0] + cs[1] + cs[2])) {
            poe_post_step(c, trk);
        } else if (u < (cs[0] + cs[1] + cs[2] + cs[3])) {
            poi_post_step(c, trk);
        }
    } else {
        along_step(c, trk, step_limit, rho_mass, z1, z2, u_phi);
    }
}

/* ------------------------------------------------------------------------------------------- */
/* scoring: scorers/mqi_scorer_energy_deposit.hpp, kernel_functions/mqi_transport.hpp:68-111      */
/* ------------------------------------------------------------------------------------------- */
static float
grid_volume(const mqo_grid* g, uint64_t cnb) { /* get_volume(cnb) mqi_grid3d.hpp:470-477 */
    const uint64_t nxy = (uint64_t) g->nx * g->ny;
    int   k = (int) (cnb / nxy);
    int   j = (int) ((cnb % nxy) / g->nx);
    int   i = (int) ((cnb % nxy) % g->nx);
    float volume = g->xe[i + 1] - g->xe[i];
    volume *= g->ye[j + 1] - g->ye[j];
    volume *= g->ze[k + 1] - g->ze[k];
    return volume;
}

static double
compute_hit(const ctx_t* c, int kind, const track_t* trk, uint64_t cnb) {
    const mqo_grid* g = c->g;
    switch (kind) {
    case MQO_SCORER_EDEP: return trk->dE + trk->local_dE;
    case MQO_SCORER_DOSE:
    case MQO_SCORER_DIJ:
    case MQO_SCORER_DOSE_SQ: {
        float density = g->rho[cnb];
        float volume  = grid_volume(g, cnb);
        if (density < 1.0e-7) {
            return 0.0;
        } else {
            double dose = (trk->dE + trk->local_dE) * 1.60218e-10 /
                          (volume * density * mqo_spr(density, trk->vtx0.ke, c->variant));
            return kind == MQO_SCORER_DOSE_SQ ? dose * dose : dose;
        }
    }
    case MQO_SCORER_LETD_NUMER:
    case MQO_SCORER_LETD_DENOM: {
        float  density = g->rho[cnb];
        double length, let;
        density *= 1000.0;
        length = (trk->vtx1.pos.x - trk->vtx0.pos.x) * (trk->vtx1.pos.x - trk->vtx0.pos.x);
        length += (trk->vtx1.pos.y - trk->vtx0.pos.y) * (trk->vtx1.pos.y - trk->vtx0.pos.y);
        length += (trk->vtx1.pos.z - trk->vtx0.pos.z) * (trk->vtx1.pos.z - trk->vtx0.pos.z);
        length = sqrt(length);
        if (length <= 0) return 0.0;
        let = trk->dE / length / density;
        if (let >= 25.0) return 0;
        return kind == MQO_SCORER_LETD_NUMER ? trk->dE * let : trk->dE * 1.0;
    }
    case MQO_SCORER_LETT_NUMER:
    case MQO_SCORER_LETT_DENOM: { /* LETt_weight1/2 scorers/mqi_scorer_energy_deposit.hpp:141-177 */
        float  density = g->rho[cnb];
        double length, let;
        density *= 1000.0;
        length = (trk->vtx1.pos.x - trk->vtx0.pos.x) * (trk->vtx1.pos.x - trk->vtx0.pos.x);
        length += (trk->vtx1.pos.y - trk->vtx0.pos.y) * (trk->vtx1.pos.y - trk->vtx0.pos.y);
        length += (trk->vtx1.pos.z - trk->vtx0.pos.z) * (trk->vtx1.pos.z - trk->vtx0.pos.z);
        length = sqrt(length);
        if (length <= 0) return 0.0;
        if (kind == MQO_SCORER_LETT_DENOM) return length;
        let = trk->dE / length / density;
        return length * let;
    }
    default: return 0.0;
    }
}

/* insert_hashtable mqi_transport.hpp:68-111 */
static void
insert_scorer(mqo_scorer* s, uint32_t key1, uint32_t key2, double value) {
    uint32_t slot;
    if (!(value > 0)) { /* the reference tests value <= 0; NaN would pass there (B18) -- rejected here */
        return;
    }
    if (s->kind != MQO_SCORER_DIJ) {
        s->dense[key1] += value;
        return;
    }
    if (key2 == K_EMPTY_PAIR) {
        slot = key1;
        key2 = 0;
    } else {
        slot = mqo_hash(key1, key2, s->capacity);
    }
    while (1) {
        mqo_key_value* e = &s->table[slot];
        uint32_t       prev1 = e->key1, prev2 = e->key2;
        if (prev1 == K_EMPTY_PAIR) e->key1 = key1;
        if (prev2 == K_EMPTY_PAIR) e->key2 = key2;
        if ((prev1 == K_EMPTY_PAIR || prev1 == key1) && (prev2 == K_EMPTY_PAIR || prev2 == key2)) {
            e->value += value;
            return;
        }
        slot = (uint32_t) ((slot + 1) % s->capacity);
    }
}

/* insert_hashtable applied to a list of (key1, key2, value) hits, in order (KAT entry point) */
void
mqo_insert(mqo_scorer* s, const uint32_t* key1, const uint32_t* key2, const double* value, uint64_t n) {
    uint64_t i;
    for (i = 0; i < n; ++i) insert_scorer(s, key1[i], key2[i], value[i]);
}

/* ------------------------------------------------------------------------------------------- */
/* transport_particles_patient mqi_transport.hpp:113-250                                         */
/* ------------------------------------------------------------------------------------------- */
static void
ctx_set_node(ctx_t* c, const mqo_grid* g) {
    c->g = g;
    c->rot_fwd.xx = g->rot_fwd[0]; c->rot_fwd.xy = g->rot_fwd[1]; c->rot_fwd.xz = g->rot_fwd[2];
    c->rot_fwd.yx = g->rot_fwd[3]; c->rot_fwd.yy = g->rot_fwd[4]; c->rot_fwd.yz = g->rot_fwd[5];
    c->rot_fwd.zx = g->rot_fwd[6]; c->rot_fwd.zy = g->rot_fwd[7]; c->rot_fwd.zz = g->rot_fwd[8];
    /* inverse() is the transpose: mqi_matrix.hpp:291-295 */
    c->rot_inv.xx = c->rot_fwd.xx; c->rot_inv.xy = c->rot_fwd.yx; c->rot_inv.xz = c->rot_fwd.zx;
    c->rot_inv.yx = c->rot_fwd.xy; c->rot_inv.yy = c->rot_fwd.yy; c->rot_inv.yz = c->rot_fwd.zy;
    c->rot_inv.zx = c->rot_fwd.xz; c->rot_inv.zy = c->rot_fwd.yz; c->rot_inv.zz = c->rot_fwd.zz;
    c->trans = v3_make(g->trans[0], g->trans[1], g->trans[2]);
}

/* roi_t::idx(cnb) > 0 (mqi_transport.hpp:205,216; mqi_roi.hpp:48-58,127-137): a DIRECT roi returns cnb
 * itself (voxel 0 is never scored, B1); a CONTOUR roi (mask_reader::mask_to_roi, mqi_file_handler.hpp:
 * 176-217) returns 1 inside a run of the mask and -1 outside.  The mask here is the run-length ROI
 * expanded back to one byte per voxel (mqo_mask_to_roi). */
static inline int
roi_accepts(const uint8_t* mask, uint64_t cnb) {
    if (!mask) return (int32_t) (uint32_t) cnb > 0;
    return mask[cnb] != 0;
}

/* mask_reader::mask_to_roi mqi_file_handler.hpp:176-217 restated on the summed mask volume
 * (read_mask_files adds the 0/1 masks of all files, :107-113): a run starts at a voxel whose sum is
 * exactly 1 while no run is open and ends at the next voxel whose sum is 0.  Voxels with a sum >= 2
 * neither open nor close a run.  A run still open at the end of the volume has no stride in the
 * reference (uninitialised read); it is closed at the volume end here.  Returns the number of runs;
 * start/stride (capacity max_runs each) and the expanded 0/1 membership are written when non-NULL. */
uint32_t
mqo_mask_to_roi(const uint8_t* mask_total, uint64_t n, uint32_t* start, uint32_t* stride, uint32_t max_runs,
                uint8_t* member) {
    uint32_t runs = 0;
    int      open = 0;
    uint64_t i, s0 = 0;
    if (member) memset(member, 0, n);
    for (i = 0; i < n; ++i) {
        if (mask_total[i] == 1 && !open) {
            open = 1;
            s0   = i;
        }
        if (mask_total[i] == 0 && open) {
            open = 0;
            if (runs < max_runs) {
                if (start) start[runs] = (uint32_t) s0;
                if (stride) stride[runs] = (uint32_t) (i - s0);
            }
            if (member) memset(member + s0, 1, i - s0);
            ++runs;
        }
    }
    if (open) {
        if (runs < max_runs) {
            if (start) start[runs] = (uint32_t) s0;
            if (stride) stride[runs] = (uint32_t) (n - s0);
        }
        if (member) memset(member + s0, 1, n - s0);
        ++runs;
    }
    return runs;
}

/* The world's children in order (beamline objects first, the patient / phantom grid last:
 * mqi_tps_env.hpp:732-758); scorers belong to the LAST node (beamline nodes have n_scorers = 0, :751).
 * roi_masks[s] (may be NULL as a whole or per scorer) is the expanded CONTOUR roi of scorer s. */
int
mqo_transport_nodes(const mqo_grid* nodes, int n_nodes, int variant, uint32_t quirks, const mqo_beamlet* beamlets,
                    const uint64_t* cum_histories, uint32_t n_beamlets, const mqo_vertex* vertices,
                    const uint32_t* spot_ids, int per_spot, uint64_t seed, uint64_t h0, uint64_t n,
                    mqo_scorer* scorers, int n_scorers, const uint8_t* const* roi_masks, mqo_stats* stats) {
    ctx_t     c;
    rng_t     rng;
    tstack_t* stk;
    uint64_t  i;
    if (!g_tables_loaded) return -1;
    if (n_nodes < 1) return -4;
    stk = (tstack_t*) malloc(sizeof(tstack_t));
    if (!stk) return -2;
    memset(&c, 0, sizeof(c));
    c.variant = variant;
    c.T_cut = (variant == MQO_VARIANT_DEBUG) ? 0.08511 * 1.0f : 0.0815 * 1.0f;
    c.rng   = &rng;
    c.stk   = stk;
    c.st    = stats;
    stk->limit = (variant == MQO_VARIANT_DEBUG) ? 200 : 10;
    stk->idx   = 0;

    for (i = 0; i < n; ++i) {
        const uint64_t h = h0 + i;
        uint32_t       spot_ind;
        mqo_vertex     vtx;
        track_t        primary;
        uint32_t       spot = 0;
        rng_init(&rng, seed, h);
        if (vertices) {
            vtx  = vertices[i];
            spot = spot_ids ? spot_ids[i] : 0;
        } else {
            /* beamsource::operator()(h): cdf2beamlet_.upper_bound(h) mqi_beamsource.hpp:109-112 */
            uint32_t lo = 0, hi = n_beamlets;
            while (lo < hi) {
                uint32_t mid = (lo + hi) / 2;
                if (cum_histories[mid] > h) hi = mid; else lo = mid + 1;
            }
            if (lo >= n_beamlets) { free(stk); return -3; }
            spot = lo;
            sample_vertex(&beamlets[spot], &rng, &vtx);
        }
        spot_ind = per_spot ? spot : K_EMPTY_PAIR; /* scorer_offset_vector :150-154 */

        memset(&primary, 0, sizeof(primary));
        primary.primary  = 1;
        primary.vtx0.ke  = vtx.ke;
        primary.vtx0.pos = v3_make(vtx.pos[0], vtx.pos[1], vtx.pos[2]);
        primary.vtx0.dir = v3_make(vtx.dir[0], vtx.dir[1], vtx.dir[2]);
        primary.vtx1     = primary.vtx0;
        stk->tracks[0]   = primary; /* push_primary */
        stk->idx         = 1;

        while (stk->idx != 0) {
            track_t track = stk->tracks[--stk->idx];
            int     c_ind;
            for (c_ind = 0; c_ind < n_nodes; ++c_ind) { /* :162 */
                const mqo_grid* g = &nodes[c_ind];
                const int nb_of_scorers = (c_ind == n_nodes - 1) ? n_scorers : 0;
                float   p[3], d[3];
                int     checker[3];
                ctx_set_node(&c, g);
                /* world -> local :165-170 */
                track.vtx0.pos = m33_mul(&c.rot_inv, v3_sub(track.vtx0.pos, c.trans));
                track.vtx0.dir = v3_normalize(m33_mul(&c.rot_inv, track.vtx0.dir));
                track.vtx1.pos = track.vtx0.pos;
                track.vtx1.dir = track.vtx0.dir;
                p[0] = track.vtx0.pos.x; p[1] = track.vtx0.pos.y; p[2] = track.vtx0.pos.z;
                d[0] = track.vtx0.dir.x; d[1] = track.vtx0.dir.y; d[2] = track.vtx0.dir.z;
                mqo_grid_index(g, p, d, checker);
                if (!grid_is_valid(g, checker)) {
                    int   ecell[3];
                    float dist = mqo_grid_intersect_entry(g, p, d, ecell);
                    track.vtx0.dir = v3_make(d[0], d[1], d[2]); /* intersect() zeroes tiny components in place */
                    track.its_dist = dist;
                    if (dist < 0) { /* :176-185: back to the world frame, next child */
                        track.vtx0.pos = v3_add(m33_mul(&c.rot_fwd, track.vtx0.pos), c.trans);
                        track.vtx0.dir = m33_mul(&c.rot_fwd, track.vtx0.dir);
                        track.vtx1.pos = track.vtx0.pos;
                        track.vtx1.dir = track.vtx0.dir;
                        continue;
                    }
                    track.vtx1.pos = v3_add(track.vtx0.pos, v3_scale(track.vtx0.dir, dist));
                    track.vtx0     = track.vtx1; /* move(): note vtx1.dir still holds the un-zeroed copy */
                    track.dE       = 0;
                    track.local_dE = 0;
                    p[0] = track.vtx0.pos.x; p[1] = track.vtx0.pos.y; p[2] = track.vtx0.pos.z;
                    d[0] = track.vtx0.dir.x; d[1] = track.vtx0.dir.y; d[2] = track.vtx0.dir.z;
                    mqo_grid_index(g, p, d, track.cell);
                } else {
                    track.its_dist = 0.0;
                    track.cell[0] = checker[0]; track.cell[1] = checker[1]; track.cell[2] = checker[2];
                }
                while (grid_is_valid(g, track.cell) && !track.stopped) {
                    uint64_t cnb = (uint64_t) track.cell[2] * g->nx * g->ny + (uint64_t) track.cell[1] * g->nx + track.cell[0];
                    int      s, pass;
                    p[0] = track.vtx0.pos.x; p[1] = track.vtx0.pos.y; p[2] = track.vtx0.pos.z;
                    d[0] = track.vtx0.dir.x; d[1] = track.vtx0.dir.y; d[2] = track.vtx0.dir.z;
                    track.its_dist = mqo_grid_intersect_cell(g, p, d, track.cell);
                    track.vtx0.dir = v3_make(d[0], d[1], d[2]);
                    if (stats) stats->steps++;
                    if (track.its_dist < 0) {
                        /* the reference still calls stepping() with a negative distance (which poisons
                         * the track with NaN) and then breaks without scoring: :195-202.  The poisoned
                         * track fails every later index / intersect test, i.e. it is lost. */
                        track.stopped = 1;
                        break;
                    }
                    stepping(&c, &track, g->rho[cnb], track.its_dist);
                    /* scoring :204-225 */
                    for (pass = 0; pass < 2; ++pass) {
                        int s_end = nb_of_scorers;
                        if (pass == 0) {
                            if (!(quirks & MQO_QUIRK_B2_DOUBLE_SCORE)) continue;
                            s_end = nb_of_scorers - 2;
                        }
                        for (s = 0; s < s_end; ++s) {
                            if (roi_accepts(roi_masks ? roi_masks[s] : NULL, cnb)) {
                                insert_scorer(&scorers[s], (uint32_t) cnb, spot_ind,
                                              compute_hit(&c, scorers[s].kind, &track, cnb));
                            }
                        }
                    }
                    if (!track.stopped) {
                        float q[3] = { track.vtx1.pos.x, track.vtx1.pos.y, track.vtx1.pos.z };
                        float e[3] = { track.vtx1.dir.x, track.vtx1.dir.y, track.vtx1.dir.z };
                        mqo_grid_index_update(g, q, e, track.cell);
                        track.vtx0     = track.vtx1;
                        track.dE       = 0;
                        track.local_dE = 0;
                    }
                }
                /* local -> world :234-239 */
                track.vtx0.pos = v3_add(m33_mul(&c.rot_fwd, track.vtx0.pos), c.trans);
                track.vtx0.dir = m33_mul(&c.rot_fwd, track.vtx0.dir);
                track.vtx1.pos = track.vtx0.pos;
                track.vtx1.dir = track.vtx0.dir;
            }
        }
        if (stats) stats->histories++;
    }
    free(stk);
    return 0;
}

int
mqo_transport(const mqo_grid* g, int variant, uint32_t quirks, const mqo_beamlet* beamlets,
              const uint64_t* cum_histories, uint32_t n_beamlets, const mqo_vertex* vertices,
              const uint32_t* spot_ids, int per_spot, uint64_t seed, uint64_t h0, uint64_t n,
              mqo_scorer* scorers, int n_scorers, mqo_stats* stats) {
    return mqo_transport_nodes(g, 1, variant, quirks, beamlets, cum_histories, n_beamlets, vertices, spot_ids,
                               per_spot, seed, h0, n, scorers, n_scorers, NULL, stats);
}
