/* zdr_oracle.c — CPU ORACLE for the zdr hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a scalar float32 restatement, in plain C, of the reference algorithm in
 * /root/reference (111116/zdr): integrator.py, camera.py, pmj02bn.py, onb.py, vertex.py,
 * interaction.py, microfacet.py, light.py, envmap.py, collocated.py, direct.py, prb.py and the
 * scene assembly of render.py:73-148.  Each function cites the file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product (zdr_b200/) never does.
 *
 * PARITY PINNING.  The reference cannot run here as shipped (no `luisa`, sampler tables are
 * missing blobs, prb.py does not JIT — SURVEY.md §8c).  The oracle is pinned instead against
 * tests/golden/ (npz files), produced by executing the reference's OWN Python sources under a small
 * LuisaCompute emulator (oracle/lc_emu.py + tests/golden/gen_golden.py, run in the build
 * container where /root/reference exists).  What the emulator cannot pin — because the
 * arithmetic lives inside the un-vendored, un-pinned dependency `luisa-python` — is listed
 * under "UNPINNED" below and in DESIGN.md: for those items parity is unpinned.
 *
 * UNPINNED (defined here, LuisaCompute source not on this box):
 *   - ray/triangle intersection arithmetic and tie-breaking (OptiX/Embree inside LC): we
 *     specify the watertight test of Woop/Benthin/Wald 2013, closest hit = smallest t, ties
 *     to the lowest (instance, primitive) in scene order; valid hits satisfy tmin < t < tmax.
 *   - offset_ray_origin: Waechter & Binder, Ray Tracing Gems ch.6 constants 1/32, 1/65536, 256.
 *   - int/uint promotion in _permutation_element's final modulo: signed floor-mod.
 *   - sin/cos implementation, fused-multiply-add placement, pow lowering: see the numeric
 *     contract below.
 *   - out-of-bounds table reads of _blue_noise (pmj02bn.py:23 goes out of range for
 *     tex>=25): index taken modulo the table length.
 *
 * NUMERIC CONTRACT (shared with the CUDA path so that forward results can be compared
 * bit-for-bit): every operation is an IEEE-754 binary32 operation in source order, compiled
 * with contraction off, except the primitives dot3/cross3/lincomb3/lerp/interp/xform which use
 * fmaf exactly as written here; x**2 = x*x, x**5 = (x*x)*(x*x)*x; normalize(v) = v*(1/sqrt(v.v));
 * sin/cos(phi) are evaluated in binary64 by zo_sincos (Cody-Waite pi/2 reduction + Taylor
 * polynomials, fma Horner) and rounded once; tan(fov/2) is libm tanf on the host.
 * Gradient accumulation (the reference's float atomics have no defined order) is done in
 * binary64 and rounded once at the end.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp, NO -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/zdr_b200.h"

#define ZO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* small vector helpers (numeric contract primitives)                                   */
/* ------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } f3;
typedef struct { float x, y; } f2;
typedef struct { float x, y, z, w; } f4;

static const float PI_F = 3.14159274101257324f;        /* float(math.pi) */
static const float TWO_PI_F = 6.28318548202514648f;    /* float(2*math.pi) */
static const float ONE_MINUS_EPS = 0x1.fffffep-1f;     /* pmj02bn.py:76 */

static inline f3 v3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 add3(f3 a, f3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub3(f3 a, f3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 mul3(f3 a, f3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 scale3(f3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline f3 div3s(f3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline f3 neg3(f3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline f3 cross3(f3 a, f3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline f3 normalize3(f3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return scale3(v, inv); }
static inline float length3(f3 v) { return sqrtf(dot3(v, v)); }
/* a*A + b*B + c*C */
static inline f3 lincomb3(float a, f3 A, float b, f3 B, float c, f3 C) {
    return v3(fmaf(c, C.x, fmaf(b, B.x, a * A.x)), fmaf(c, C.y, fmaf(b, B.y, a * A.y)),
              fmaf(c, C.z, fmaf(b, B.z, a * A.z)));
}
static inline float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int any_nan3(f3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }
/* float -> int with CUDA cvt.rzi.s32.f32 semantics (saturating, NaN -> 0) */
static inline int f2i(float x) {
    if (isnan(x)) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
}

/* sin/cos of a binary32 angle, evaluated in binary64 (numeric contract) */
static void zo_sincos(float phi, float* s_out, float* c_out) {
    double x = (double)phi;
    double kd = rint(x * 0.63661977236758134308);
    double r = fma(-kd, 1.57079632679489655800e+00, x);
    r = fma(-kd, 6.12323399573676603587e-17, r);
    double r2 = r * r;
    double ps = -1.0 / 1307674368000.0;               /* -1/15! */
    ps = fma(ps, r2, 1.0 / 6227020800.0);
    ps = fma(ps, r2, -1.0 / 39916800.0);
    ps = fma(ps, r2, 1.0 / 362880.0);
    ps = fma(ps, r2, -1.0 / 5040.0);
    ps = fma(ps, r2, 1.0 / 120.0);
    ps = fma(ps, r2, -1.0 / 6.0);
    double sn = fma(ps * r2, r, r);
    double pc = 1.0 / 20922789888000.0;               /* 1/16! */
    pc = fma(pc, r2, -1.0 / 87178291200.0);
    pc = fma(pc, r2, 1.0 / 479001600.0);
    pc = fma(pc, r2, -1.0 / 3628800.0);
    pc = fma(pc, r2, 1.0 / 40320.0);
    pc = fma(pc, r2, -1.0 / 720.0);
    pc = fma(pc, r2, 1.0 / 24.0);
    pc = fma(pc, r2, -0.5);
    double cs = fma(pc, r2, 1.0);
    long long k = (long long)kd;
    double s, c;
    switch ((int)(k & 3)) {
        case 0: s = sn; c = cs; break;
        case 1: s = cs; c = -sn; break;
        case 2: s = -sn; c = -cs; break;
        default: s = -cs; c = sn; break;
    }
    *s_out = (float)s;
    *c_out = (float)c;
}

/* ------------------------------------------------------------------------------------ */
/* sampler tables (pmj02bn.py:9-16)                                                     */
/* ------------------------------------------------------------------------------------ */
static float* g_pmj = NULL;    /* [sets*n*2] */
static float* g_bn = NULL;     /* [ntex*res*res] */
static int g_pmj_sets = 0, g_pmj_n = 0, g_bn_ntex = 0, g_bn_res = 0;

ZO_API int zo_set_sampler_tables(const uint32_t* pmj, int sets, int n, const uint16_t* bn, int ntex, int res) {
    if (!pmj || !bn || sets <= 0 || n <= 0 || ntex <= 0 || res <= 0) return ZDR_ERR_ARG;
    free(g_pmj); free(g_bn);
    size_t np = (size_t)sets * n * 2, nb = (size_t)ntex * res * res;
    g_pmj = (float*)malloc(np * sizeof(float));
    g_bn = (float*)malloc(nb * sizeof(float));
    /* pmj02bn.py:9: torch.tensor(ints)/2**32 -> RN(u32->f32) * 2^-32 */
    for (size_t i = 0; i < np; ++i) g_pmj[i] = (float)pmj[i] * 0x1p-32f;
    /* pmj02bn.py:14 */
    for (size_t i = 0; i < nb; ++i) g_bn[i] = (float)bn[i] * 0x1p-16f;
    g_pmj_sets = sets; g_pmj_n = n; g_bn_ntex = ntex; g_bn_res = res;
    return ZDR_OK;
}

/* pmj02bn.py:19-24 — LITERAL index formula tex*(res+cx)*res+cy (differs from pbrt's);
 * out-of-range indices wrap modulo the table length (UNPINNED: UB upstream). */
static inline float blue_noise(int tex_id, int px, int py) {
    int tex_index = tex_id % g_bn_ntex;
    int cx = px % g_bn_res, cy = py % g_bn_res;      /* C truncating %, as LC lowers it */
    long long i = (long long)tex_index * (g_bn_res + cx) * g_bn_res + cy;
    long long len = (long long)g_bn_ntex * g_bn_res * g_bn_res;
    i %= len; if (i < 0) i += len;
    return g_bn[i];
}
/* pmj02bn.py:26-30 */
static inline f2 pmj02bn_sample(int set_id, int sample_id) {
    int set_index = set_id % g_pmj_sets;
    long long i = (long long)set_index * g_pmj_n + sample_id;
    long long len = (long long)g_pmj_sets * g_pmj_n;
    i %= len; if (i < 0) i += len;
    f2 r = {g_pmj[2 * i], g_pmj[2 * i + 1]};
    return r;
}
/* pmj02bn.py:32-57 (Kensler permutation; signed floor-mod on the last line) */
ZO_API int zo_permutation_element(int i_in, int l, int w_in, uint32_t p) {
    uint32_t i = (uint32_t)i_in, w = (uint32_t)w_in;
    for (;;) {
        i ^= p;
        i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;
        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;
        i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11;
        i *= 0x74dcb303u;
        i ^= (i & w) >> 2;
        i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;
        i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
        if (i < (uint32_t)l) break;
    }
    int32_t s = (int32_t)(i + p);
    return ((s % l) + l) % l;
}
/* pmj02bn.py:59-74 */
ZO_API uint32_t zo_xxhash32_int4(int x, int y, int z, int w) {
    const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t px = (uint32_t)x, py = (uint32_t)y, pz = (uint32_t)z, pw = (uint32_t)w;
    uint32_t h = pw + P5 + px * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h += py * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h += pz * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h = P2 * (h ^ (h >> 15));
    h = P3 * (h ^ (h >> 13));
    return h ^ (h >> 16);
}

typedef struct { int px, py, sample_index, dimension, seed, spp, w; } Sampler; /* pmj02bn.py:78-85 */

/* pmj02bn.py:87-102 */
static Sampler make_sampler(int px, int py, int seed, int spp, int sample_index) {
    int w = spp - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    Sampler t = {px, py, sample_index, 0, seed, spp, w};
    return t;
}
/* pmj02bn.py:104-112 */
static float sampler_next(Sampler* s) {
    uint32_t h = zo_xxhash32_int4(s->px, s->py, s->dimension, s->seed);
    int index = zo_permutation_element(s->sample_index, s->spp, s->w, h);
    float delta = blue_noise(s->dimension, s->px ^ s->seed, s->py ^ s->seed);
    float u = ((float)index + delta) / (float)s->spp;
    s->dimension += 1;
    return clampf(u, 0.0f, ONE_MINUS_EPS);
}
/* pmj02bn.py:114-126 */
static f2 sampler_next2f(Sampler* s) {
    int index = s->sample_index;
    int pmj_instance = s->dimension / 2;
    if (pmj_instance >= g_pmj_sets) {
        uint32_t h = zo_xxhash32_int4(s->px, s->py, s->dimension, s->seed);
        index = zo_permutation_element(s->sample_index, s->spp, s->w, h);
    }
    f2 u = pmj02bn_sample(pmj_instance, index);
    u.x = u.x + blue_noise(s->dimension, s->px ^ s->seed, s->py ^ s->seed);
    u.y = u.y + blue_noise(s->dimension + 1, s->px ^ s->seed, s->py ^ s->seed);
    s->dimension += 2;
    u.x = u.x - floorf(u.x);   /* fract */
    u.y = u.y - floorf(u.y);
    return u;
}

/* ------------------------------------------------------------------------------------ */
/* scene (render.py:73-148)                                                             */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int nv, nt;
    float* verts;       /* nv*8, Vertex rows (vertex.py:4) */
    int32_t* tris;      /* nt*3 */
    float M[16];        /* column-major float4x4 */
    float N[9];         /* inverse(transpose(M3x3)), column-major (interaction.py:28) */
    f3 emission;
    int tri_base;       /* first global triangle id */
    int material_set;   /* texture set that shades the instance, -1 = light / terminal surface; default: instance 0 -> 0, others -1
                           ("treat everything as light other than the first object", direct.py:30, prb.py:45) */
} Inst;

struct zo_scene {
    int ninst;
    Inst* inst;
    int committed;
    int ntri;
    f3* P;              /* world-space triangle corners, 3 per global triangle */
    int* tri_inst;      /* global tri -> instance */
    int light_count;
    int* lights;        /* render.py:121 */
    int env_count;
    int pdf_env_count;  /* what sample_light_pdf adds to light_count: 0 (light.py:89) or env_count (ZDR_FLAG_LIGHT_PDF_COUNTS_ENV) */
    /* envmap (envmap.py) */
    int env_w, env_h;
    float* env_tex;     /* rgba, env_h rows (after the row duplication of envmap.py:123-125) */
    float* env_alias_prob; int* env_alias_idx; float* env_pdf;
    float* env_scale_map;   /* weight map before MIS compensation (envmap.py:136-160), kept for the tests */
    /* optional acceleration of the brute-force scans for large scenes (same triangle test, same winner) */
    struct ONode* bvh; int* bvh_order; int bvh_nodes;
    /* counters */
    uint64_t n_closest, n_any;
};
typedef struct zo_scene zo_scene;
struct ONode { float lo[3], hi[3]; int left, right, first, count; };   /* count > 0: leaf over bvh_order[first..] */

ZO_API zo_scene* zo_scene_create(void) { return (zo_scene*)calloc(1, sizeof(zo_scene)); }
ZO_API void zo_scene_destroy(zo_scene* s) {
    if (!s) return;
    for (int i = 0; i < s->ninst; ++i) { free(s->inst[i].verts); free(s->inst[i].tris); }
    free(s->inst); free(s->P); free(s->tri_inst); free(s->lights);
    free(s->env_tex); free(s->env_alias_prob); free(s->env_alias_idx); free(s->env_pdf); free(s->env_scale_map);
    free(s->bvh); free(s->bvh_order);
    free(s);
}

/* (transform * float4(v,1)).xyz, interaction.py:19-21 — numeric contract `xform` */
static inline f3 xform_point(const float* M, f3 v) {
    return v3(fmaf(M[8], v.z, fmaf(M[4], v.y, fmaf(M[0], v.x, M[12]))),
              fmaf(M[9], v.z, fmaf(M[5], v.y, fmaf(M[1], v.x, M[13]))),
              fmaf(M[10], v.z, fmaf(M[6], v.y, fmaf(M[2], v.x, M[14]))));
}
/* inverse(transpose(make_float3x3(M[0].xyz, M[1].xyz, M[2].xyz))), interaction.py:28.
 * Cofactor expansion; column-major in and out. */
static void normal_matrix(const float* M, float* N) {
    /* A = transpose(M3): A[col][row] = M3[row][col]; write a_rc = element (row r, col c) of A */
    float a00 = M[0], a01 = M[1], a02 = M[2];      /* row 0 of A = column 0 of M */
    float a10 = M[4], a11 = M[5], a12 = M[6];
    float a20 = M[8], a21 = M[9], a22 = M[10];
    float c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    float c10 = a02 * a21 - a01 * a22, c11 = a00 * a22 - a02 * a20, c12 = a01 * a20 - a00 * a21;
    float c20 = a01 * a12 - a02 * a11, c21 = a02 * a10 - a00 * a12, c22 = a00 * a11 - a01 * a10;
    float det = a00 * c00 + a01 * c01 + a02 * c02;
    float id = 1.0f / det;
    /* inverse = adj/det, adj[r][c] = cofactor[c][r]; store column-major N[c*3+r] */
    N[0] = c00 * id; N[1] = c01 * id; N[2] = c02 * id;     /* column 0: inv[r][0] = cof[0][r] */
    N[3] = c10 * id; N[4] = c11 * id; N[5] = c12 * id;
    N[6] = c20 * id; N[7] = c21 * id; N[8] = c22 * id;
}
static inline f3 mat3_mul(const float* N, f3 v) {
    return v3(fmaf(N[6], v.z, fmaf(N[3], v.y, N[0] * v.x)), fmaf(N[7], v.z, fmaf(N[4], v.y, N[1] * v.x)),
              fmaf(N[8], v.z, fmaf(N[5], v.y, N[2] * v.x)));
}

ZO_API int zo_scene_add_mesh(zo_scene* s, const float* verts, int nv, const int32_t* tris, int nt,
                             const float* xform, const float* emission) {
    if (!s || !verts || !tris || nv <= 0 || nt <= 0) return ZDR_ERR_ARG;
    if (s->ninst >= 10000) return ZDR_ERR_LIMIT;   /* render.py:114 */
    s->inst = (Inst*)realloc(s->inst, sizeof(Inst) * (s->ninst + 1));
    Inst* in = &s->inst[s->ninst++];
    in->nv = nv; in->nt = nt;
    in->verts = (float*)malloc(sizeof(float) * 8 * nv);
    memcpy(in->verts, verts, sizeof(float) * 8 * nv);
    in->tris = (int32_t*)malloc(sizeof(int32_t) * 3 * nt);
    memcpy(in->tris, tris, sizeof(int32_t) * 3 * nt);
    static const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(in->M, xform ? xform : I16, sizeof(float) * 16);
    normal_matrix(in->M, in->N);
    in->emission = emission ? v3(emission[0], emission[1], emission[2]) : v3(0, 0, 0);
    in->material_set = (s->ninst == 1) ? 0 : -1;      /* direct.py:30: only the first object is textured */
    s->committed = 0;
    return ZDR_OK;
}

/* material sets ("polymorphic materials", README.md:5 — not implemented upstream): texture set per instance, -1 = light / terminal */
ZO_API int zo_scene_set_material_sets(zo_scene* s, const int32_t* set_of_instance, int n) {
    if (!s || !set_of_instance || n != s->ninst) return ZDR_ERR_ARG;
    for (int i = 0; i < n; ++i) if (set_of_instance[i] < -1) return ZDR_ERR_ARG;
    for (int i = 0; i < n; ++i) s->inst[i].material_set = set_of_instance[i];
    return ZDR_OK;
}

static void rebuild_lights(zo_scene* s) {
    free(s->lights);
    s->lights = (int*)calloc(s->ninst > 0 ? s->ninst : 1, sizeof(int));
    s->light_count = 0;
    for (int i = 0; i < s->ninst; ++i) {            /* render.py:89-90, :146 */
        f3 e = s->inst[i].emission;
        if (e.x > 0 || e.y > 0 || e.z > 0) s->lights[s->light_count++] = i;
    }
}

static void bvh_build(zo_scene* s);
ZO_API int zo_scene_commit(zo_scene* s) {
    if (!s || s->ninst == 0) return ZDR_ERR_STATE;
    int nt = 0;
    for (int i = 0; i < s->ninst; ++i) { s->inst[i].tri_base = nt; nt += s->inst[i].nt; }
    free(s->P); free(s->tri_inst);
    s->ntri = nt;
    s->P = (f3*)malloc(sizeof(f3) * 3 * nt);
    s->tri_inst = (int*)malloc(sizeof(int) * nt);
    for (int i = 0; i < s->ninst; ++i) {
        Inst* in = &s->inst[i];
        for (int t = 0; t < in->nt; ++t) {
            int g = in->tri_base + t;
            s->tri_inst[g] = i;
            for (int k = 0; k < 3; ++k) {
                const float* v = in->verts + 8 * in->tris[3 * t + k];
                s->P[3 * g + k] = xform_point(in->M, v3(v[0], v[1], v[2]));
            }
        }
    }
    rebuild_lights(s);
    bvh_build(s);
    s->committed = 1;
    return ZDR_OK;
}

ZO_API int zo_scene_set_emissions(zo_scene* s, const float* e3, int n) {   /* render.py:130-148 */
    if (!s || !e3 || n != s->ninst) return ZDR_ERR_ARG;
    for (int i = 0; i < n; ++i) s->inst[i].emission = v3(e3[3 * i], e3[3 * i + 1], e3[3 * i + 2]);
    rebuild_lights(s);
    return ZDR_OK;
}
ZO_API int zo_scene_light_count(const zo_scene* s) { return s->light_count; }
ZO_API int zo_scene_instance_count(const zo_scene* s) { return s->ninst; }

/* ------------------------------------------------------------------------------------ */
/* intersection (UNPINNED inside LuisaCompute; specified here): watertight Woop test    */
/* ------------------------------------------------------------------------------------ */
typedef struct { f3 o, d; float tmin, tmax; } Ray;
typedef struct { int inst, prim; float u, v, t; } Hit;   /* inst<0: miss; (u,v) weights of v1,v2 */

typedef struct { int kx, ky, kz; float Sx, Sy, Sz; } RayPre;

static inline float comp(f3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

static RayPre ray_pre(f3 d) {
    RayPre r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    if (comp(d, r.kz) < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    r.Sx = comp(d, r.kx) / comp(d, r.kz);
    r.Sy = comp(d, r.ky) / comp(d, r.kz);
    r.Sz = 1.0f / comp(d, r.kz);
    return r;
}

/* returns 1 and fills (t,u,v) when the triangle is hit with tmin < t < tmax */
static int tri_test(const RayPre* rp, f3 o, f3 p0, f3 p1, f3 p2, float tmin, float tmax,
                    float* t_out, float* u_out, float* v_out) {
    f3 A = sub3(p0, o), B = sub3(p1, o), C = sub3(p2, o);
    float Akz = comp(A, rp->kz), Bkz = comp(B, rp->kz), Ckz = comp(C, rp->kz);
    float Ax = fmaf(-rp->Sx, Akz, comp(A, rp->kx)), Ay = fmaf(-rp->Sy, Akz, comp(A, rp->ky));
    float Bx = fmaf(-rp->Sx, Bkz, comp(B, rp->kx)), By = fmaf(-rp->Sy, Bkz, comp(B, rp->ky));
    float Cx = fmaf(-rp->Sx, Ckz, comp(C, rp->kx)), Cy = fmaf(-rp->Sy, Ckz, comp(C, rp->ky));
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
    float det = U + V + W;
    if (det == 0.0f) return 0;
    float Az = rp->Sz * Akz, Bz = rp->Sz * Bkz, Cz = rp->Sz * Ckz;
    float T = fmaf(W, Cz, fmaf(V, Bz, U * Az));
    float rcp = 1.0f / det;
    float t = T * rcp;
    if (!(t > tmin && t < tmax)) return 0;
    *t_out = t; *u_out = V * rcp; *v_out = W * rcp;
    return 1;
}

/* ---- median-split BVH over the world-space triangles (TEST INFRASTRUCTURE for scenes too large to scan) ----
 * Only a pruning device: every candidate still goes through tri_test, and the winner is the smallest t with ties to
 * the lowest scene-order triangle — exactly what the in-order scan with a strict '<' keeps — so closest hits,
 * occlusion answers and ray counts are those of the brute-force scan.  Boxes are inflated and the slab test is
 * widened, as on the GPU.  Scenes below ZO_BVH_MIN_TRIS keep the scan; ZDR_ORACLE_BVH=0 / 1 forces either. */
#define ZO_BVH_MIN_TRIS 4096
#define ZO_BVH_LEAF 4
static float tri_centroid(const zo_scene* s, int g, int ax) {
    return comp(s->P[3 * g], ax) + comp(s->P[3 * g + 1], ax) + comp(s->P[3 * g + 2], ax);
}
static void bvh_select(const zo_scene* s, int* idx, int lo, int hi, int k, int ax) {   /* quickselect: idx[k] in place */
    while (hi - lo > 1) {
        float pivot = tri_centroid(s, idx[lo + (hi - lo) / 2], ax);
        int i = lo, j = hi - 1;
        while (i <= j) {
            while (tri_centroid(s, idx[i], ax) < pivot) ++i;
            while (tri_centroid(s, idx[j], ax) > pivot) --j;
            if (i <= j) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; ++i; --j; }
        }
        if (k <= j) hi = j + 1; else if (k >= i) lo = i; else return;
    }
}
static int bvh_build_rec(zo_scene* s, int first, int count, float eps) {
    int self = s->bvh_nodes++;
    struct ONode* n = &s->bvh[self];
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (int q = 0; q < 3; ++q) { n->lo[q] = 1e30f; n->hi[q] = -1e30f; }
    for (int i = first; i < first + count; ++i) {
        int g = s->bvh_order[i];
        for (int q = 0; q < 3; ++q) {
            float c = tri_centroid(s, g, q);
            clo[q] = fminf(clo[q], c); chi[q] = fmaxf(chi[q], c);
            for (int k = 0; k < 3; ++k) {
                float v = comp(s->P[3 * g + k], q);
                n->lo[q] = fminf(n->lo[q], v); n->hi[q] = fmaxf(n->hi[q], v);
            }
        }
    }
    for (int q = 0; q < 3; ++q) { n->lo[q] -= eps; n->hi[q] += eps; }
    n->first = first; n->count = 0; n->left = n->right = -1;
    if (count <= ZO_BVH_LEAF) { n->count = count; return self; }
    int ax = 0;
    if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
    if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
    int mid = first + count / 2;
    if (chi[ax] > clo[ax]) bvh_select(s, s->bvh_order, first, first + count, mid, ax);   /* else: coincident centroids, split by position */
    int l = bvh_build_rec(s, first, mid - first, eps);
    int r = bvh_build_rec(s, mid, first + count - mid, eps);
    s->bvh[self].left = l; s->bvh[self].right = r;      /* (s->bvh is not reallocated during the build) */
    return self;
}
static void bvh_build(zo_scene* s) {
    free(s->bvh); free(s->bvh_order);
    s->bvh = NULL; s->bvh_order = NULL; s->bvh_nodes = 0;
    const char* force = getenv("ZDR_ORACLE_BVH");
    int want = force ? atoi(force) != 0 : s->ntri >= ZO_BVH_MIN_TRIS;
    if (!want || s->ntri == 0) return;
    float lo = 1e30f, hi = -1e30f;
    for (int i = 0; i < 3 * s->ntri; ++i) {
        lo = fminf(lo, fminf(s->P[i].x, fminf(s->P[i].y, s->P[i].z)));
        hi = fmaxf(hi, fmaxf(s->P[i].x, fmaxf(s->P[i].y, s->P[i].z)));
    }
    float eps = 1e-5f * fmaxf(fmaxf(fabsf(lo), fabsf(hi)), 1e-20f);
    s->bvh = (struct ONode*)malloc(sizeof(struct ONode) * (size_t)(2 * s->ntri));
    s->bvh_order = (int*)malloc(sizeof(int) * (size_t)s->ntri);
    for (int g = 0; g < s->ntri; ++g) s->bvh_order[g] = g;
    bvh_build_rec(s, 0, s->ntri, eps);
}
static inline int bvh_box(const struct ONode* n, f3 o, f3 id, float tmin, float tmax, float* tnear) {
    float x0 = (n->lo[0] - o.x) * id.x, x1 = (n->hi[0] - o.x) * id.x;
    float y0 = (n->lo[1] - o.y) * id.y, y1 = (n->hi[1] - o.y) * id.y;
    float z0 = (n->lo[2] - o.z) * id.z, z1 = (n->hi[2] - o.z) * id.z;
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    *tnear = tn;
    /* NaN (0 * inf on a slab boundary) must not prune */
    return !(tn * 0.999999f > tf * 1.000001f);
}
static inline f3 bvh_inv_dir(f3 d) {
    f3 id = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    if (!(fabsf(id.x) <= 1e30f)) id.x = copysignf(1e30f, d.x);
    if (!(fabsf(id.y) <= 1e30f)) id.y = copysignf(1e30f, d.y);
    if (!(fabsf(id.z) <= 1e30f)) id.z = copysignf(1e30f, d.z);
    return id;
}
static Hit bvh_trace_closest(zo_scene* s, Ray r) {
    Hit h = {-1, -1, 0, 0, 0};
    RayPre rp = ray_pre(r.d);
    f3 id = bvh_inv_dir(r.d);
    float best = r.tmax;
    int best_g = 0x7fffffff, have = 0;
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const struct ONode* n = &s->bvh[stack[--sp]];
        float tn;
        if (!bvh_box(n, r.o, id, r.tmin, best, &tn)) continue;
        if (n->count == 0) {                         /* nearer child on top of the stack: best shrinks early */
            float tl, tr;
            bvh_box(&s->bvh[n->left], r.o, id, r.tmin, best, &tl);
            bvh_box(&s->bvh[n->right], r.o, id, r.tmin, best, &tr);
            if (tl < tr) { stack[sp++] = n->right; stack[sp++] = n->left; }
            else { stack[sp++] = n->left; stack[sp++] = n->right; }
            continue;
        }
        for (int i = n->first; i < n->first + n->count; ++i) {
            int g = s->bvh_order[i];
            float t, u, v;
            if (!tri_test(&rp, r.o, s->P[3 * g], s->P[3 * g + 1], s->P[3 * g + 2], r.tmin, INFINITY, &t, &u, &v)) continue;
            if (t < best || (have && t == best && g < best_g)) {
                best = t; best_g = g; have = 1;
                h.t = t; h.u = u; h.v = v; h.inst = s->tri_inst[g]; h.prim = g - s->inst[h.inst].tri_base;
            }
        }
    }
    return h;
}
static int bvh_trace_any(zo_scene* s, Ray r) {
    RayPre rp = ray_pre(r.d);
    f3 id = bvh_inv_dir(r.d);
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const struct ONode* n = &s->bvh[stack[--sp]];
        float tn;
        if (!bvh_box(n, r.o, id, r.tmin, r.tmax, &tn)) continue;
        if (n->count == 0) { stack[sp++] = n->left; stack[sp++] = n->right; continue; }
        for (int i = n->first; i < n->first + n->count; ++i) {
            int g = s->bvh_order[i];
            float t, u, v;
            if (tri_test(&rp, r.o, s->P[3 * g], s->P[3 * g + 1], s->P[3 * g + 2], r.tmin, r.tmax, &t, &u, &v)) return 1;
        }
    }
    return 0;
}

/* accel.trace_closest (collocated.py:16 ...): brute force in scene order; ties keep the first */
static Hit trace_closest(zo_scene* s, Ray r) {
    if (s->bvh) return bvh_trace_closest(s, r);
    Hit h = {-1, -1, 0, 0, 0};
    RayPre rp = ray_pre(r.d);
    float best = r.tmax;
    for (int g = 0; g < s->ntri; ++g) {
        float t, u, v;
        if (tri_test(&rp, r.o, s->P[3 * g], s->P[3 * g + 1], s->P[3 * g + 2], r.tmin, best, &t, &u, &v)) {
            best = t; h.t = t; h.u = u; h.v = v;
            h.inst = s->tri_inst[g]; h.prim = g - s->inst[h.inst].tri_base;
        }
    }
    return h;
}
/* accel.trace_any (direct.py:44) */
static int trace_any(zo_scene* s, Ray r) {
    if (s->bvh) return bvh_trace_any(s, r);
    RayPre rp = ray_pre(r.d);
    for (int g = 0; g < s->ntri; ++g) {
        float t, u, v;
        if (tri_test(&rp, r.o, s->P[3 * g], s->P[3 * g + 1], s->P[3 * g + 2], r.tmin, r.tmax, &t, &u, &v)) return 1;
    }
    return 0;
}

/* luisa.offset_ray_origin (direct.py:64): Waechter-Binder, RT Gems ch.6 (UNPINNED) */
static inline float int_as_float(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline int32_t float_as_int(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static f3 offset_ray_origin(f3 p, f3 n) {
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    int ox = f2i(int_scale * n.x), oy = f2i(int_scale * n.y), oz = f2i(int_scale * n.z);
    float px = int_as_float(float_as_int(p.x) + (p.x < 0 ? -ox : ox));
    float py = int_as_float(float_as_int(p.y) + (p.y < 0 ? -oy : oy));
    float pz = int_as_float(float_as_int(p.z) + (p.z < 0 ? -oz : oz));
    return v3(fabsf(p.x) < origin ? fmaf(float_scale, n.x, p.x) : px,
              fabsf(p.y) < origin ? fmaf(float_scale, n.y, p.y) : py,
              fabsf(p.z) < origin ? fmaf(float_scale, n.z, p.z) : pz);
}

/* ------------------------------------------------------------------------------------ */
/* onb.py, interaction.py                                                               */
/* ------------------------------------------------------------------------------------ */
typedef struct { f3 t, b, n; } Onb;
/* onb.py:21-28 */
static Onb make_onb(f3 n) {
    Onb r;
    r.b = normalize3(fabsf(n.x) > fabsf(n.z) ? v3(-n.y, n.x, 0.0f) : v3(0.0f, -n.z, n.y));
    r.t = normalize3(cross3(r.b, n));
    r.n = n;
    return r;
}
static inline f3 to_world(const Onb* o, f3 v) { return lincomb3(v.x, o->t, v.y, o->b, v.z, o->n); } /* onb.py:10 */
static inline f3 to_local(const Onb* o, f3 v) { return v3(dot3(v, o->t), dot3(v, o->b), dot3(v, o->n)); } /* onb.py:14 */

typedef struct { f3 p; f2 uv; f3 ns, ng; } Interaction;   /* interaction.py:6 */

/* hit.interpolate(a,b,c) = (1-u-v)a + u b + v c  [LC]; numeric contract `interp` */
static inline f3 interp3(const Hit* h, f3 a, f3 b, f3 c) {
    float w = 1.0f - h->u - h->v;
    return lincomb3(w, a, h->u, b, h->v, c);
}
/* interaction.py:8-30 */
static Interaction surface_interact(zo_scene* s, const Hit* h) {
    Inst* in = &s->inst[h->inst];
    const float* v0 = in->verts + 8 * in->tris[3 * h->prim + 0];
    const float* v1 = in->verts + 8 * in->tris[3 * h->prim + 1];
    const float* v2 = in->verts + 8 * in->tris[3 * h->prim + 2];
    f3 p0 = xform_point(in->M, v3(v0[0], v0[1], v0[2]));
    f3 p1 = xform_point(in->M, v3(v1[0], v1[1], v1[2]));
    f3 p2 = xform_point(in->M, v3(v2[0], v2[1], v2[2]));
    Interaction it;
    it.p = interp3(h, p0, p1, p2);
    float w = 1.0f - h->u - h->v;
    it.uv.x = fmaf(h->v, v2[3], fmaf(h->u, v1[3], w * v0[3]));
    it.uv.y = fmaf(h->v, v2[4], fmaf(h->u, v1[4], w * v0[4]));
    f3 ns0 = interp3(h, v3(v0[5], v0[6], v0[7]), v3(v1[5], v1[6], v1[7]), v3(v2[5], v2[6], v2[7]));
    it.ns = normalize3(mat3_mul(in->N, ns0));
    it.ng = normalize3(cross3(sub3(p1, p0), sub3(p2, p0)));
    return it;
}

typedef struct { const float* mat; double* dmat; int rx, ry; int bilinear; } Tex;
/* material sets ("polymorphic materials", README.md:5): set k of the material buffer starts k * rx * ry texels in */
static inline Tex tex_set(const Tex* t, int set) {
    Tex r = *t;
    const size_t off = (size_t)(set < 0 ? 0 : set) * (size_t)t->rx * (size_t)t->ry * 4;
    r.mat = t->mat + off;
    if (t->dmat) r.dmat = t->dmat + off;
    return r;
}

/* interaction.py:35-44 */
static inline f4 read_single_bsdf(const Tex* t, int cx, int cy) {
    cx = clampi(cx, 0, t->rx - 1); cy = clampi(cy, 0, t->ry - 1);
    int idx = cx + t->rx * cy;
    f4 r = {t->mat[idx * 4 + 0], t->mat[idx * 4 + 1], t->mat[idx * 4 + 2], t->mat[idx * 4 + 3]};
    return r;
}
static inline f4 lerp4(f4 a, f4 b, float t) {
    f4 r = {lerpf(a.x, b.x, t), lerpf(a.y, b.y, t), lerpf(a.z, b.z, t), lerpf(a.w, b.w, t)};
    return r;
}
/* interaction.py:46-60 */
static f4 read_bsdf(const Tex* t, f2 uv) {
    float px = uv.x * (float)(t->rx - 1), py = (1.0f - uv.y) * (float)(t->ry - 1);
    if (!t->bilinear) return read_single_bsdf(t, f2i(px + 0.4999f), f2i(py + 0.4999f));
    int ix = f2i(px), iy = f2i(py);
    float ox = px - (float)ix, oy = py - (float)iy;
    f4 c00 = read_single_bsdf(t, ix, iy), c01 = read_single_bsdf(t, ix, iy + 1);
    f4 c10 = read_single_bsdf(t, ix + 1, iy), c11 = read_single_bsdf(t, ix + 1, iy + 1);
    return lerp4(lerp4(c00, c01, oy), lerp4(c10, c11, oy), ox);
}
/* interaction.py:62-70: 4 float atomics per texel; order-free, so accumulated in binary64 */
static inline void write_single_bsdf_grad(const Tex* t, int cx, int cy, float k, f4 d) {
    cx = clampi(cx, 0, t->rx - 1); cy = clampi(cy, 0, t->ry - 1);
    int idx = cx + t->rx * cy;
    float gx = k * d.x, gy = k * d.y, gz = k * d.z, gw = k * d.w;
    double* q = t->dmat + (size_t)idx * 4;
#pragma omp atomic
    q[0] += (double)gx;
#pragma omp atomic
    q[1] += (double)gy;
#pragma omp atomic
    q[2] += (double)gz;
#pragma omp atomic
    q[3] += (double)gw;
}
/* interaction.py:72-89 */
static void write_bsdf_grad(const Tex* t, f2 uv, f4 d) {
    float px = uv.x * (float)(t->rx - 1), py = (1.0f - uv.y) * (float)(t->ry - 1);
    if (!t->bilinear) { write_single_bsdf_grad(t, f2i(px + 0.4999f), f2i(py + 0.4999f), 1.0f, d); return; }
    int ix = f2i(px), iy = f2i(py);
    float ox = px - (float)ix, oy = py - (float)iy;
    float k00 = (1.0f - ox) * (1.0f - oy), k01 = (1.0f - ox) * oy, k10 = ox * (1.0f - oy), k11 = ox * oy;
    write_single_bsdf_grad(t, ix, iy, k00, d);
    write_single_bsdf_grad(t, ix, iy + 1, k01, d);
    write_single_bsdf_grad(t, ix + 1, iy, k10, d);
    write_single_bsdf_grad(t, ix + 1, iy + 1, k11, d);
}

/* ------------------------------------------------------------------------------------ */
/* microfacet.py                                                                        */
/* ------------------------------------------------------------------------------------ */
static inline float ggx_distribution(f3 h, float alpha) {      /* microfacet.py:6-11 */
    float alpha2 = alpha * alpha;
    float nh = fmaxf(0.00001f, h.z);
    float t = nh * nh * (alpha2 - 1.0f) + 1.0f;
    return alpha2 / (PI_F * (t * t));
}
static inline float fresnel_schlick(float c, float specular) { /* microfacet.py:13-15 */
    float x = 1.0f - c, x2 = x * x;
    return specular + (1.0f - specular) * (x2 * x2 * x);
}
static inline float smith_geometry(f3 v, float alpha) {        /* microfacet.py:17-21 */
    float alpha2 = alpha * alpha;
    float nv = fmaxf(0.00001f, v.z);
    return 2.0f / (1.0f + sqrtf(1.0f + alpha2 * (1.0f - nv * nv) / (nv * nv)));
}
/* microfacet.py:23-30 */
static f3 ggx_brdf(f3 wo, f3 wi, f3 diffuse, float specular, float roughness) {
    float alpha = roughness * roughness;
    f3 h = normalize3(add3(wi, wo));
    float d = ggx_distribution(h, alpha);
    float f = fresnel_schlick(clampf(dot3(wo, h), 0.00001f, 1.0f), specular);
    float g = smith_geometry(wi, alpha) * smith_geometry(wo, alpha);
    float spec = (d * f * g) / (4.0f * fmaxf(0.00001f, wi.z) * fmaxf(0.00001f, wo.z));
    return v3((spec + diffuse.x / PI_F) * wi.z, (spec + diffuse.y / PI_F) * wi.z, (spec + diffuse.z / PI_F) * wi.z);
}
/* Analytic adjoint of ggx_brdf wrt mat=(diffuse.xyz, roughness) for cotangent g; replaces the
 * LuisaCompute autodiff blocks at collocated.py:44-56, direct.py:126-131,160-165,
 * prb.py:138-146,157-163 (SURVEY.md §8a formulae). */
static f4 ggx_brdf_grad(f3 wo, f3 wi, float specular, float roughness, f3 g) {
    float alpha = roughness * roughness;
    float a2 = alpha * alpha;
    f3 h = normalize3(add3(wi, wo));
    float nh = fmaxf(0.00001f, h.z);
    float nh2 = nh * nh;
    float t = nh2 * (a2 - 1.0f) + 1.0f;
    float D = a2 / (PI_F * (t * t));
    float dD = (t - 2.0f * a2 * nh2) / (PI_F * (t * t * t));
    float F = fresnel_schlick(clampf(dot3(wo, h), 0.00001f, 1.0f), specular);
    float nvi = fmaxf(0.00001f, wi.z), nvo = fmaxf(0.00001f, wo.z);
    float ki = (1.0f - nvi * nvi) / (nvi * nvi), ko = (1.0f - nvo * nvo) / (nvo * nvo);
    float si = sqrtf(1.0f + a2 * ki), so = sqrtf(1.0f + a2 * ko);
    float Gi = 2.0f / (1.0f + si), Go = 2.0f / (1.0f + so);
    float dGi = -ki / (si * ((1.0f + si) * (1.0f + si)));
    float dGo = -ko / (so * ((1.0f + so) * (1.0f + so)));
    float dDGG = dD * Gi * Go + D * dGi * Go + D * Gi * dGo;
    float da2_dr = 4.0f * roughness * roughness * roughness;
    float common = wi.z * F / (4.0f * nvi * nvo) * dDGG * da2_dr;
    float dk = wi.z / PI_F;
    f4 r = {g.x * dk, g.y * dk, g.z * dk, (g.x + g.y + g.z) * common};
    return r;
}
/* microfacet.py:33-37 */
static f3 cosine_sample_hemisphere(f2 u) {
    float r = sqrtf(u.x);
    float phi = TWO_PI_F * u.y;
    float s, c; zo_sincos(phi, &s, &c);
    return v3(r * c, r * s, sqrtf(1.0f - u.x));
}
/* microfacet.py:70-73 */
static float pdf_wm(f3 w, f3 wm, float alpha) {
    return smith_geometry(w, alpha) / fabsf(w.z) * ggx_distribution(wm, alpha) * fabsf(dot3(w, wm));
}
/* microfacet.py:75-92 */
static f3 sample_wm(f3 w, float alpha, f2 u) {
    f3 wh = normalize3(v3(alpha * w.x, alpha * w.y, w.z));
    if (wh.z < 0.0f) wh = neg3(wh);
    f3 T1 = wh.z < 0.99999f ? normalize3(cross3(v3(0.0f, 0.0f, 1.0f), wh)) : v3(1.0f, 0.0f, 0.0f);
    f3 T2 = cross3(wh, T1);
    float r = sqrtf(u.x);
    float theta = TWO_PI_F * u.y;
    float sn, cs; zo_sincos(theta, &sn, &cs);
    float px = r * cs, py = r * sn;
    float h = sqrtf(1.0f - px * px);
    py = lerpf(h, py, (1.0f + wh.z) / 2.0f);
    float pz = sqrtf(fmaxf(0.0f, 1.0f - fmaf(py, py, px * px)));
    f3 nh = lincomb3(px, T1, py, T2, pz, wh);
    return normalize3(v3(alpha * nh.x, alpha * nh.y, fmaxf(1e-6f, nh.z)));
}
/* microfacet.py:40-49 */
static f3 ggx_sample(f3 wo, float roughness, Sampler* sampler) {
    if (sampler_next(sampler) < 0.5f) return cosine_sample_hemisphere(sampler_next2f(sampler));
    float alpha = roughness * roughness;
    f3 wm = sample_wm(wo, alpha, sampler_next2f(sampler));
    /* reflect(-wo, wm) = i - 2 dot(n,i) n */
    f3 i = neg3(wo);
    float k = 2.0f * dot3(wm, i);
    return v3(fmaf(-k, wm.x, i.x), fmaf(-k, wm.y, i.y), fmaf(-k, wm.z, i.z));
}
/* microfacet.py:51-58 */
static float ggx_sample_pdf(f3 wo, f3 wi, float roughness) {
    float alpha = roughness * roughness;
    f3 wm = normalize3(add3(wi, wo));
    float diffuse_pdf = wi.z / PI_F;
    float glossy_pdf = pdf_wm(wo, wm, alpha) / (4.0f * fabsf(dot3(wo, wm)));
    return 0.5f * diffuse_pdf + 0.5f * glossy_pdf;
}

/* ------------------------------------------------------------------------------------ */
/* envmap.py (device side)                                                              */
/* ------------------------------------------------------------------------------------ */
#define ENV_MAP_W 512
#define ENV_MAP_H 256
/* heap.texture2d_sample(23332, uv): UNPINNED [LC]; specified as bilinear, REPEAT in u,
 * CLAMP in v, texel centres at (i+0.5)/size */
static f3 env_texture(const zo_scene* s, f2 uv) {
    float fx = uv.x * (float)s->env_w - 0.5f, fy = uv.y * (float)s->env_h - 0.5f;
    float flx = floorf(fx), fly = floorf(fy);
    float tx = fx - flx, ty = fy - fly;
    int x0 = f2i(flx), y0 = f2i(fly);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 %= s->env_w; if (x0 < 0) x0 += s->env_w;
    x1 %= s->env_w; if (x1 < 0) x1 += s->env_w;
    y0 = clampi(y0, 0, s->env_h - 1); y1 = clampi(y1, 0, s->env_h - 1);
    const float* a = s->env_tex + 4 * ((size_t)y0 * s->env_w + x0);
    const float* b = s->env_tex + 4 * ((size_t)y0 * s->env_w + x1);
    const float* c = s->env_tex + 4 * ((size_t)y1 * s->env_w + x0);
    const float* d = s->env_tex + 4 * ((size_t)y1 * s->env_w + x1);
    return v3(lerpf(lerpf(a[0], b[0], tx), lerpf(c[0], d[0], tx), ty),
              lerpf(lerpf(a[1], b[1], tx), lerpf(c[1], d[1], tx), ty),
              lerpf(lerpf(a[2], b[2], tx), lerpf(c[2], d[2], tx), ty));
}
/* envmap.py:206-213 */
static f3 uv_to_direction(f2 uv) {
    float phi = TWO_PI_F * (1.0f - uv.x);
    float theta = PI_F * uv.y;
    float sp, cp, st, ct;
    zo_sincos(phi, &sp, &cp); zo_sincos(theta, &st, &ct);
    return normalize3(v3(sp * st, ct, cp * st));
}
/* envmap.py:215-219 (acos/atan2: libm; tolerance-compared on the GPU side) */
static f2 direction_to_uv(f3 d) {
    float theta = acosf(d.y);
    float phi = atan2f(d.x, d.z);
    f2 r = {1.0f - phi / TWO_PI_F, theta / PI_F};
    return r;
}
/* envmap.py:240-248 */
static float env_sampled_light_pdf(const zo_scene* s, f3 dir) {
    f2 uv = direction_to_uv(dir);
    int index = clampi(f2i(uv.y * (float)ENV_MAP_H), 0, ENV_MAP_H - 1) * ENV_MAP_W +
                clampi(f2i(uv.x * (float)ENV_MAP_W), 0, ENV_MAP_W - 1);
    float pdf = s->env_pdf[index];
    float sn, cs; zo_sincos(PI_F * uv.y, &sn, &cs);
    float inv_s = sn > 0.0f ? 1.0f / sn : 0.0f;
    return pdf * inv_s / (2.0f * PI_F * PI_F);
}

typedef struct { f3 wi; float dist, pdf; f3 eval; } LightSample;   /* light.py:11 */

/* envmap.py:70-98 */
static void sample_alias(const zo_scene* s, int n, float u_in, int offset, int* index, float* uu) {
    float u = u_in * (float)n;
    int i = clampi(f2i(u), 0, n - 1);
    float ur = u - floorf(u);
    float prob = s->env_alias_prob[i + offset];
    int alias = s->env_alias_idx[i + offset];
    *index = ur < prob ? i : alias;
    *uu = ur < prob ? ur / prob : (ur - prob) / (1.0f - prob);
}
/* envmap.py:223-238 */
static LightSample sample_envmap(const zo_scene* s, f2 u) {
    int sy, sx; float uy, ux;
    sample_alias(s, ENV_MAP_H, u.y, 0, &sy, &uy);
    int offset = ENV_MAP_H + sy * ENV_MAP_W;
    sample_alias(s, ENV_MAP_W, u.x, offset, &sx, &ux);
    f2 uv = {((float)sx + ux) / (float)ENV_MAP_W, ((float)sy + uy) / (float)ENV_MAP_H};
    float pdf = s->env_pdf[sy * ENV_MAP_W + sx];
    LightSample t;
    t.wi = uv_to_direction(uv);
    t.dist = 1e30f;
    float sn, cs; zo_sincos(PI_F * uv.y, &sn, &cs);
    float inv_s = sn > 0.0f ? 1.0f / sn : 0.0f;
    t.pdf = pdf * inv_s / (2.0f * PI_F * PI_F);
    t.eval = env_texture(s, uv);
    return t;
}

/* ------------------------------------------------------------------------------------ */
/* light.py                                                                             */
/* ------------------------------------------------------------------------------------ */
/* light.py:15-20 */
static f3 sample_uniform_triangle(f2 u) {
    float a, b;
    if (u.x < u.y) { a = 0.5f * u.x; b = -0.5f * u.x + u.y; }
    else { a = -0.5f * u.y + u.x; b = 0.5f * u.y; }
    return v3(a, b, 1.0f - a - b);
}
static void inst_triangle(zo_scene* s, int inst, int prim, f3* p0, f3* p1, f3* p2) {
    int g = s->inst[inst].tri_base + prim;      /* = M*(v,1) recomputed, light.py:52-61 */
    *p0 = s->P[3 * g]; *p1 = s->P[3 * g + 1]; *p2 = s->P[3 * g + 2];
}
/* light.py:22-81 (point_light_count = 0, light.py:7) */
static LightSample sample_light(zo_scene* s, f3 origin, Sampler* sampler) {
    float u = sampler_next(sampler);
    int n = s->env_count + 0 + s->light_count;
    LightSample t;
    if (n == 0) {
        /* UNPINNED: upstream indexes lights[-1] (UB). Defined: consume the mesh-light
         * dimensions, return a null sample. */
        (void)sampler_next(sampler); (void)sampler_next2f(sampler);
        t.wi = v3(0, 0, 1); t.dist = 0.0f; t.pdf = 1.0f; t.eval = v3(0, 0, 0);
        return t;
    }
    int idx = clampi(f2i(u * (float)n), 0, n - 1);
    if (idx < s->env_count) return sample_envmap(s, sampler_next2f(sampler));
    idx -= s->env_count;
    int inst = s->lights[idx];
    int trig_count = s->inst[inst].nt;
    int prim = clampi(f2i(sampler_next(sampler) * (float)trig_count), 0, trig_count - 1);
    f3 p0, p1, p2; inst_triangle(s, inst, prim, &p0, &p1, &p2);
    f3 abc = sample_uniform_triangle(sampler_next2f(sampler));
    f3 p = lincomb3(abc.x, p0, abc.y, p1, abc.z, p2);
    f3 emission = s->inst[inst].emission;
    f3 wi_light = normalize3(sub3(p, origin));
    f3 c = cross3(sub3(p1, p0), sub3(p2, p0));
    f3 light_normal = normalize3(c);
    float cos_light = -dot3(light_normal, wi_light);
    f3 dd = sub3(p, origin);
    float sqr_dist = dot3(dd, dd);
    float area = length3(c) / 2.0f;
    t.pdf = sqr_dist / ((float)(n * trig_count) * area * cos_light);
    t.eval = cos_light > 1e-4f ? emission : v3(0, 0, 0);
    t.wi = wi_light;
    t.dist = 0.9999f * sqrtf(sqr_dist);
    return t;
}
/* light.py:83-111 — note n EXCLUDES env_count (light.py:89), reproduced; with ZDR_FLAG_LIGHT_PDF_COUNTS_ENV it
 * is the n of sample_light (light.py:26) */
static float sample_light_pdf(zo_scene* s, f3 origin, int inst, int prim, f3 p) {
    int n = s->pdf_env_count + s->light_count;
    int trig_count = s->inst[inst].nt;
    f3 p0, p1, p2; inst_triangle(s, inst, prim, &p0, &p1, &p2);
    f3 wi_light = normalize3(sub3(p, origin));
    f3 c = cross3(sub3(p1, p0), sub3(p2, p0));
    f3 light_normal = normalize3(c);
    float cos_light = -dot3(light_normal, wi_light);
    f3 dd = sub3(p, origin);
    float sqr_dist = dot3(dd, dd);
    float area = length3(c) / 2.0f;
    return sqr_dist / ((float)(n * trig_count) * area * cos_light);
}

/* ------------------------------------------------------------------------------------ */
/* estimators                                                                           */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    zo_scene* s;
    Tex tex;
    const zdr_params* p;
} Ctx;

static inline float balanced_heuristic(float a, float b) { return a / fmaxf(a + b, 1e-4f); } /* direct.py:16-18 */
static inline int backfacing(f3 d, const Interaction* it) {   /* collocated.py:20 */
    f3 nd = neg3(d);
    return dot3(nd, it->ng) < 1e-4f || dot3(nd, it->ns) < 1e-4f;
}
static inline f4 add4(f4 a, f4 b) { f4 r = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; return r; }

/* collocated.py:10-31 */
static f3 collocated_estimator(Ctx* c, Ray ray, Sampler* sampler) {
    (void)sampler;
    c->s->n_closest++;
    Hit hit = trace_closest(c->s, ray);
    if (hit.inst < 0) return v3(0, 0, 0);
    Interaction it = surface_interact(c->s, &hit);
    if (backfacing(ray.d, &it)) return v3(0, 0, 0);
    const Tex tex = tex_set(&c->tex, c->s->inst[hit.inst].material_set);      /* collocated.py:22 shades any instance */
    f4 mat = read_bsdf(&tex, it.uv);
    Onb onb = make_onb(it.ns);
    f3 wo_local = to_local(&onb, neg3(ray.d));
    f3 beta = ggx_brdf(wo_local, wo_local, v3(mat.x, mat.y, mat.z), 0.04f, mat.w);
    float inv = 1.0f / hit.t;
    float li = 1.0f * (inv * inv);
    return scale3(beta, li);
}
/* collocated.py:34-57 */
static void collocated_estimator_backward(Ctx* c, Ray ray, Sampler* sampler, f3 le_grad) {
    (void)sampler;
    c->s->n_closest++;
    Hit hit = trace_closest(c->s, ray);
    if (hit.inst < 0) return;
    Interaction it = surface_interact(c->s, &hit);
    if (backfacing(ray.d, &it)) return;
    const Tex tex = tex_set(&c->tex, c->s->inst[hit.inst].material_set);
    f4 mat = read_bsdf(&tex, it.uv);
    Onb onb = make_onb(it.ns);
    f3 wo_local = to_local(&onb, neg3(ray.d));
    float inv = 1.0f / hit.t;
    float li = 1.0f * (inv * inv);
    f4 mat_grad = ggx_brdf_grad(wo_local, wo_local, 0.04f, mat.w, scale3(le_grad, li));
    write_bsdf_grad(&tex, it.uv, mat_grad);
}

static f3 env_emission(Ctx* c, f3 dir) {   /* direct.py:24 */
    return c->s->env_count > 0 ? env_texture(c->s, direction_to_uv(dir)) : v3(0, 0, 0);
}

/* direct.py:20-85 */
static f3 direct_estimator(Ctx* c, Ray ray, Sampler* sampler) {
    zo_scene* s = c->s;
    s->n_closest++;
    Hit hit = trace_closest(s, ray);
    if (hit.inst < 0) return env_emission(c, ray.d);
    Interaction it = surface_interact(s, &hit);
    if (backfacing(ray.d, &it)) return v3(0, 0, 0);
    f3 emission = s->inst[hit.inst].emission;
    if (s->inst[hit.inst].material_set < 0) return emission;        /* direct.py:30: `hit.inst > 0` with one material set */
    const Tex tex = tex_set(&c->tex, s->inst[hit.inst].material_set);
    f4 mat = read_bsdf(&tex, it.uv);
    f3 diffuse = v3(mat.x, mat.y, mat.z);
    float roughness = mat.w;
    const float specular = 0.04f;
    f3 radiance = v3(0, 0, 0);
    LightSample light = sample_light(s, it.p, sampler);
    Ray shadow = {it.p, light.wi, 1e-4f, light.dist};
    s->n_any++;
    int occluded = trace_any(s, shadow);
    Onb onb = make_onb(it.ns);
    f3 wo_local = to_local(&onb, neg3(ray.d));
    f3 wi_light_local = to_local(&onb, light.wi);
    if (!occluded && wi_light_local.z > 0.0f) {
        f3 bsdf = ggx_brdf(wo_local, wi_light_local, diffuse, specular, roughness);
        float mis_weight = 1.0f;
        if (c->p->use_mis) {
            float pdf_bsdf = ggx_sample_pdf(wo_local, wi_light_local, roughness);
            mis_weight = balanced_heuristic(light.pdf, pdf_bsdf);
        }
        float dn = fmaxf(light.pdf, 1e-4f);
        radiance = add3(radiance, div3s(mul3(scale3(bsdf, mis_weight), light.eval), dn));
    }
    if (c->p->use_mis) {
        f3 wi_local = ggx_sample(wo_local, roughness, sampler);
        f3 wi = to_world(&onb, wi_local);
        if (dot3(wi, it.ng) < 1e-4f || wi_local.z < 1e-4f) return radiance;
        Ray r2 = {offset_ray_origin(it.p, it.ng), wi, 0.0f, 1e30f};
        f3 origin = it.p;
        s->n_closest++;
        Hit hit2 = trace_closest(s, r2);
        float pdf_light;
        if (hit2.inst < 0) {
            emission = env_emission(c, r2.d);
            pdf_light = s->env_count > 0 ? env_sampled_light_pdf(s, r2.d) : 0.0f; /* direct.py:71: unused when env_count==0 */
        } else {
            Interaction it2 = surface_interact(s, &hit2);
            if (backfacing(r2.d, &it2)) return radiance;
            emission = s->inst[hit2.inst].emission;
            pdf_light = sample_light_pdf(s, origin, hit2.inst, hit2.prim, it2.p);
        }
        if (emission.x > 0.0f || emission.y > 0.0f || emission.z > 0.0f) {
            float pdf_bsdf = ggx_sample_pdf(wo_local, wi_local, roughness);
            float mis_weight = balanced_heuristic(pdf_bsdf, pdf_light);
            f3 beta = div3s(ggx_brdf(wo_local, wi_local, diffuse, specular, roughness), pdf_bsdf);
            return add3(radiance, mul3(scale3(beta, mis_weight), emission));
        }
    }
    return radiance;
}

/* direct.py:88-167 */
static void direct_estimator_backward(Ctx* c, Ray ray, Sampler* sampler, f3 le_grad) {
    zo_scene* s = c->s;
    s->n_closest++;
    Hit hit = trace_closest(s, ray);
    if (hit.inst < 0) return;
    Interaction it = surface_interact(s, &hit);
    if (backfacing(ray.d, &it)) return;
    f3 emission;
    if (s->inst[hit.inst].material_set < 0) return;                   /* direct.py:98 */
    const Tex tex = tex_set(&c->tex, s->inst[hit.inst].material_set);  /* the gradient belongs to the PRIMARY hit's texture set */
    f4 mat_grad = {0, 0, 0, 0};
    f2 mat_grad_uv = it.uv;
    f4 mat = read_bsdf(&tex, it.uv);
    float roughness = mat.w;
    const float specular = 0.04f;
    LightSample light = sample_light(s, it.p, sampler);
    Ray shadow = {it.p, light.wi, 1e-4f, light.dist};
    s->n_any++;
    int occluded = trace_any(s, shadow);
    Onb onb = make_onb(it.ns);
    f3 wo_local = to_local(&onb, neg3(ray.d));
    f3 wi_light_local = to_local(&onb, light.wi);
    if (!occluded && wi_light_local.z > 0.0f) {
        float mis_weight = 1.0f;
        if (c->p->use_mis) {
            float pdf_bsdf = ggx_sample_pdf(wo_local, wi_light_local, roughness);
            mis_weight = balanced_heuristic(light.pdf, pdf_bsdf);
        }
        /* le = bsdf * mis_weight * light.eval / max(light.pdf,1e-4); backward(le, le_grad) */
        float dn = fmaxf(light.pdf, 1e-4f);
        f3 cot = mul3(le_grad, div3s(scale3(light.eval, mis_weight), dn));
        mat_grad = add4(mat_grad, ggx_brdf_grad(wo_local, wi_light_local, specular, roughness, cot));
    }
    if (c->p->use_mis) {
        f3 wi_local = ggx_sample(wo_local, roughness, sampler);
        f3 wi = to_world(&onb, wi_local);
        if (dot3(wi, it.ng) < 1e-4f || wi_local.z < 1e-4f) {
            write_bsdf_grad(&tex, it.uv, mat_grad);                       /* direct.py:138 */
            return;
        }
        Ray r2 = {offset_ray_origin(it.p, it.ng), wi, 0.0f, 1e30f};
        f3 origin = it.p;
        s->n_closest++;
        Hit hit2 = trace_closest(s, r2);
        float pdf_light;
        if (hit2.inst < 0) {
            emission = env_emission(c, r2.d);
            pdf_light = s->env_count > 0 ? env_sampled_light_pdf(s, r2.d) : 0.0f;
        } else {
            Interaction it2 = surface_interact(s, &hit2);
            if (backfacing(r2.d, &it2)) {
                /* direct.py:152 scatters at `it.uv` AFTER `it` was reassigned to the second
                 * hit (direct.py:149): reproduced literally. */
                write_bsdf_grad(&tex, it2.uv, mat_grad);
                return;
            }
            emission = s->inst[hit2.inst].emission;
            pdf_light = sample_light_pdf(s, origin, hit2.inst, hit2.prim, it2.p);
        }
        if (emission.x > 0.0f || emission.y > 0.0f || emission.z > 0.0f) {
            float pdf_bsdf = ggx_sample_pdf(wo_local, wi_local, roughness);
            float mis_weight = balanced_heuristic(pdf_bsdf, pdf_light);
            /* le = (bsdf/pdf_bsdf) * mis_weight * emission */
            f3 cot = div3s(mul3(le_grad, scale3(emission, mis_weight)), pdf_bsdf);
            mat_grad = add4(mat_grad, ggx_brdf_grad(wo_local, wi_local, specular, roughness, cot));
        }
    }
    write_bsdf_grad(&tex, mat_grad_uv, mat_grad);                         /* direct.py:167 */
}

static inline float luminance(f3 b) { return dot3(v3(0.212671f, 0.715160f, 0.072169f), b); } /* prb.py:80 */

/* prb.py:18-88, with the call-arity resolutions of SURVEY.md §8a (direct.py conventions) */
static f3 path_estimator(Ctx* c, Ray ray, Sampler* sampler) {
    zo_scene* s = c->s;
    const int corrected = c->p->prb_mode == ZDR_PRB_CORRECTED;
    f3 radiance = v3(0, 0, 0), beta = v3(1, 1, 1);
    float pdf_bsdf = 1e30f;
    const float specular = 0.04f;
    for (int depth = 0; depth < c->p->max_depth; ++depth) {
        s->n_closest++;
        Hit hit = trace_closest(s, ray);
        if (hit.inst < 0) {
            if (s->env_count > 0) {
                f3 tex = env_texture(s, direction_to_uv(ray.d));
                f3 emission = corrected ? tex : mul3(beta, tex);           /* prb.py:28: beta applied twice */
                float pdf_light = env_sampled_light_pdf(s, ray.d);
                float mis_weight = balanced_heuristic(pdf_bsdf, pdf_light);
                radiance = add3(radiance, mul3(scale3(beta, mis_weight), emission));
            }
            break;
        }
        Interaction it = surface_interact(s, &hit);
        if (backfacing(ray.d, &it)) break;
        f3 emission = s->inst[hit.inst].emission;
        if (emission.x > 0.0f || emission.y > 0.0f || emission.z > 0.0f) {
            float pdf_light = sample_light_pdf(s, ray.o, hit.inst, hit.prim, it.p);
            float mis_weight = balanced_heuristic(pdf_bsdf, pdf_light);
            radiance = add3(radiance, mul3(scale3(beta, mis_weight), emission));
            break;
        }
        if (s->inst[hit.inst].material_set < 0) break;              /* prb.py:45 */
        const Tex tex = tex_set(&c->tex, s->inst[hit.inst].material_set);
        f4 mat = read_bsdf(&tex, it.uv);
        f3 diffuse = v3(mat.x, mat.y, mat.z);
        float roughness = mat.w;
        Onb onb = make_onb(it.ns);
        f3 wo_local = to_local(&onb, neg3(ray.d));
        LightSample light = sample_light(s, it.p, sampler);
        Ray shadow = {it.p, light.wi, 1e-4f, light.dist};
        s->n_any++;
        int occluded = trace_any(s, shadow);
        f3 wi_light_local = to_local(&onb, light.wi);
        if (!occluded && wi_light_local.z >= 1e-4f) {
            f3 bsdf = ggx_brdf(wo_local, wi_light_local, diffuse, specular, roughness);
            pdf_bsdf = ggx_sample_pdf(wo_local, wi_light_local, roughness);
            float mis_weight = balanced_heuristic(light.pdf, pdf_bsdf);
            float dn = fmaxf(light.pdf, 1e-4f);
            radiance = add3(radiance, div3s(mul3(scale3(mul3(beta, bsdf), mis_weight), light.eval), dn));
        }
        f3 wi_local = ggx_sample(wo_local, roughness, sampler);
        pdf_bsdf = ggx_sample_pdf(wo_local, wi_local, roughness);
        f3 wi = to_world(&onb, wi_local);
        if (dot3(wi, it.ng) < 1e-4f || wi_local.z < 1e-4f) break;
        Ray nr = {offset_ray_origin(it.p, it.ng), wi, 0.0f, 1e30f};
        ray = nr;
        beta = mul3(beta, div3s(ggx_brdf(wo_local, wi_local, diffuse, specular, roughness), pdf_bsdf));
        if (depth >= c->p->rr_depth) {
            float l = luminance(beta);
            if (l == 0.0f) break;
            float q = fmaxf(l, 0.05f);
            float r = sampler_next(sampler);
            if (r >= q) break;
            beta = div3s(beta, q);
        }
    }
    return radiance;
}

/* prb.py:91-187 */
static void path_estimator_backward(Ctx* c, Ray ray, Sampler* sampler, f3 le_grad) {
    zo_scene* s = c->s;
    const int corrected = c->p->prb_mode == ZDR_PRB_CORRECTED;
    Sampler sampler_twin = *sampler;                     /* prb.py:97-98 */
    f3 Le = path_estimator(c, ray, &sampler_twin);
    if (any_nan3(Le)) return;
    f3 beta = v3(1, 1, 1);
    float pdf_bsdf = 1e30f;
    const float specular = 0.04f;
    for (int depth = 0; depth < c->p->max_depth; ++depth) {
        s->n_closest++;
        Hit hit = trace_closest(s, ray);
        if (hit.inst < 0) break;
        Interaction it = surface_interact(s, &hit);
        if (backfacing(ray.d, &it)) break;
        f3 emission = s->inst[hit.inst].emission;
        if (emission.x > 0.0f || emission.y > 0.0f || emission.z > 0.0f) {
            float pdf_light = sample_light_pdf(s, ray.o, hit.inst, hit.prim, it.p);
            float mis_weight = balanced_heuristic(pdf_bsdf, pdf_light);
            Le = sub3(Le, mul3(scale3(beta, mis_weight), emission));
            break;
        }
        if (corrected && s->inst[hit.inst].material_set < 0) break;             /* missing upstream (defect 2) */
        const Tex tex = tex_set(&c->tex, s->inst[hit.inst].material_set);      /* literal mode shades any instance: set 0 for -1 */
        f4 mat_grad = {0, 0, 0, 0};
        f4 mat = read_bsdf(&tex, it.uv);
        f3 diffuse = v3(mat.x, mat.y, mat.z);
        float roughness = mat.w;
        Onb onb = make_onb(it.ns);
        f3 wo_local = to_local(&onb, neg3(ray.d));
        LightSample light = sample_light(s, it.p, sampler);
        Ray shadow = {it.p, light.wi, 1e-4f, light.dist};
        s->n_any++;
        int occluded = trace_any(s, shadow);
        f3 wi_light_local = to_local(&onb, light.wi);
        int done = 0;
        if (!occluded && wi_light_local.z >= 1e-4f) {
            f3 bsdf = ggx_brdf(wo_local, wi_light_local, diffuse, specular, roughness);
            pdf_bsdf = ggx_sample_pdf(wo_local, wi_light_local, roughness);
            float mis_weight = balanced_heuristic(light.pdf, pdf_bsdf);
            float dn = fmaxf(light.pdf, 1e-4f);
            /* backward(bsdf, beta*mis_weight*light.eval/max(light.pdf,1e-4)*le_grad), prb.py:145 */
            f3 cot = mul3(div3s(mul3(scale3(beta, mis_weight), light.eval), dn), le_grad);
            mat_grad = add4(mat_grad, ggx_brdf_grad(wo_local, wi_light_local, specular, roughness, cot));
            Le = sub3(Le, div3s(mul3(scale3(mul3(beta, bsdf), mis_weight), light.eval), dn));
        }
        f3 wi_local = ggx_sample(wo_local, roughness, sampler);
        pdf_bsdf = ggx_sample_pdf(wo_local, wi_local, roughness);
        f3 wi = to_world(&onb, wi_local);
        if (dot3(wi, it.ng) < 1e-4f || wi_local.z < 1e-4f) done = 1;
        if (!done) {
            Ray nr = {offset_ray_origin(it.p, it.ng), wi, 0.0f, 1e30f};
            ray = nr;
            f3 bsdf = ggx_brdf(wo_local, wi_local, diffuse, specular, roughness);
            f3 cot;
            if (corrected) {
                /* d(Le_remaining)/d(bsdf) = Le_remaining / bsdf, per channel */
                cot = v3(bsdf.x != 0.0f ? Le.x / bsdf.x * le_grad.x : 0.0f,
                         bsdf.y != 0.0f ? Le.y / bsdf.y * le_grad.y : 0.0f,
                         bsdf.z != 0.0f ? Le.z / bsdf.z * le_grad.z : 0.0f);
            } else {
                cot = mul3(mul3(div3s(beta, pdf_bsdf), Le), le_grad);      /* prb.py:162 */
            }
            mat_grad = add4(mat_grad, ggx_brdf_grad(wo_local, wi_local, specular, roughness, cot));
            beta = mul3(beta, div3s(bsdf, pdf_bsdf));
            if (depth >= c->p->rr_depth) {
                float l = luminance(beta);
                if (l == 0.0f) done = 1;
                else {
                    float q = fmaxf(l, 0.05f);
                    float r = sampler_next(sampler);
                    if (r >= q) done = 1;
                    else beta = div3s(beta, q);
                }
            }
        }
        /* prb.py:178-187: per-vertex flush (in-loop or after-loop, same `it`), NaNs dropped */
        if (mat_grad.x != 0.0f || mat_grad.y != 0.0f || mat_grad.z != 0.0f || mat_grad.w != 0.0f) {
            if (!(isnan(mat_grad.x) || isnan(mat_grad.y) || isnan(mat_grad.z) || isnan(mat_grad.w)))
                write_bsdf_grad(&tex, it.uv, mat_grad);
        }
        if (done) break;
    }
}

/* ------------------------------------------------------------------------------------ */
/* camera.py + integrator.py                                                            */
/* ------------------------------------------------------------------------------------ */
static inline float tent_warp1(float u) {      /* camera.py:20-31, radius 1 */
    return u < 0.5f ? 1.0f * (sqrtf(2.0f * u) - 1.0f) : 1.0f * (1.0f - sqrtf(2.0f - 2.0f * u));
}
/* camera.py:5-17 */
static Ray generate_ray(const zdr_params* p, f2 px) {
    f3 origin = v3(p->cam_origin[0], p->cam_origin[1], p->cam_origin[2]);
    f3 target = v3(p->cam_target[0], p->cam_target[1], p->cam_target[2]);
    f3 up = v3(p->cam_up[0], p->cam_up[1], p->cam_up[2]);
    f3 forward = normalize3(sub3(target, origin));
    f3 right = normalize3(cross3(forward, up));
    f3 up_perp = cross3(right, forward);
    float tn = tanf(0.5f * p->cam_fov);
    float x = px.x * tn, y = px.y * tn;
    f3 d = v3(x * right.x - y * up_perp.x + forward.x, x * right.y - y * up_perp.y + forward.y,
              x * right.z - y * up_perp.z + forward.z);
    Ray r = {origin, normalize3(d), 0.0f, 1e30f};
    return r;
}
/* integrator.py:17-25 */
static Ray camera_sample(const zdr_params* p, int x, int y, Sampler* sampler) {
    f2 off = sampler_next2f(sampler);
    if (p->use_tent_filter) { off.x = tent_warp1(off.x) + 0.5f; off.y = tent_warp1(off.y) + 0.5f; }
    f2 px;
    px.x = 2.0f / (float)p->width * ((float)x + off.x) - 1.0f;
    px.y = 2.0f / (float)p->height * ((float)y + off.y) - 1.0f;
    px.y *= (float)p->height / (float)p->width;
    return generate_ray(p, px);
}

static f3 estimate(Ctx* c, int integrator, Ray ray, Sampler* sm) {
    switch (integrator) {
        case ZDR_COLLOCATED: return collocated_estimator(c, ray, sm);
        case ZDR_DIRECT: return direct_estimator(c, ray, sm);
        default: return path_estimator(c, ray, sm);
    }
}

/* zdr_params.shard_index/shard_count: 8-row tiles of [row0,row1) dealt round-robin */
static inline int row_in_shard(const zdr_params* p, int y) {
    if (p->shard_count <= 1) return 1;
    return ((y - p->row0) / 8) % p->shard_count == p->shard_index;
}

static int check_render_args(zo_scene* s, int integrator, const zdr_params* p) {
    if (!s || !p) return ZDR_ERR_ARG;
    if (!s->committed || !g_pmj || !g_bn) return ZDR_ERR_STATE;
    if (integrator < 0 || integrator > 2) return ZDR_ERR_ARG;
    if (p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->tex_res_x <= 0 || p->tex_res_y <= 0) return ZDR_ERR_ARG;
    if (p->row0 < 0 || p->row1 > p->height || p->row0 > p->row1 || p->it0 < 0 || p->it1 > p->spp || p->it0 > p->it1) return ZDR_ERR_ARG;
    return ZDR_OK;
}

/* integrator.py:10-29.  samples_out (optional): (H,W,spp,3) un-clamped per-sample radiance */
static int render_forward_impl(zo_scene* s, int integrator, const zdr_params* p, const float* material,
                               float* image, float* samples_out) {
    int rc = check_render_args(s, integrator, p);
    if (rc) return rc;
    uint64_t nc = 0, na = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nc, na)
    for (int y = p->row0; y < p->row1; ++y) {
        if (!row_in_shard(p, y)) continue;
        zo_scene local = *s; local.n_closest = 0; local.n_any = 0;
        local.pdf_env_count = (p->flags & ZDR_FLAG_LIGHT_PDF_COUNTS_ENV) ? s->env_count : 0;
        Ctx c = {&local, {material, NULL, p->tex_res_x, p->tex_res_y, p->bilinear}, p};
        for (int x = 0; x < p->width; ++x) {
            f3 acc = v3(0, 0, 0);
            for (int it = p->it0; it < p->it1; ++it) {
                Sampler sm = make_sampler(x, y, p->seed, p->spp, it);
                Ray ray = camera_sample(p, x, y, &sm);
                f3 radiance = estimate(&c, integrator, ray, &sm);
                if (samples_out) {
                    float* o = samples_out + (((size_t)y * p->width + x) * p->spp + it) * 3;
                    o[0] = radiance.x; o[1] = radiance.y; o[2] = radiance.z;
                }
                if (!any_nan3(radiance))
                    acc = add3(acc, v3(clampf(radiance.x, 0.0f, 100000.0f), clampf(radiance.y, 0.0f, 100000.0f),
                                       clampf(radiance.z, 0.0f, 100000.0f)));
            }
            if (image) {
                float* o = image + ((size_t)x + (size_t)y * p->width) * 4;
                o[0] = acc.x / (float)p->spp; o[1] = acc.y / (float)p->spp; o[2] = acc.z / (float)p->spp; o[3] = 1.0f;
            }
        }
        nc += local.n_closest; na += local.n_any;
    }
    s->n_closest += nc; s->n_any += na;
    return ZDR_OK;
}
ZO_API int zo_render_forward(zo_scene* s, int integrator, const zdr_params* p, const float* material, float* image) {
    if (!material || !image) return ZDR_ERR_ARG;
    return render_forward_impl(s, integrator, p, material, image, NULL);
}
ZO_API int zo_render_samples(zo_scene* s, int integrator, const zdr_params* p, const float* material, float* out) {
    if (!material || !out) return ZDR_ERR_ARG;
    return render_forward_impl(s, integrator, p, material, NULL, out);
}

/* integrator.py:34-52.  d_material is ACCUMULATED into (callers zero it, render.py:220). */
ZO_API int zo_render_backward(zo_scene* s, int integrator, const zdr_params* p, const float* material,
                              const float* d_image, float* d_material) {
    int rc = check_render_args(s, integrator, p);
    if (rc) return rc;
    if (!material || !d_image || !d_material) return ZDR_ERR_ARG;
    size_t nmat = (size_t)p->tex_res_x * p->tex_res_y * 4 * (size_t)(p->num_material_sets > 1 ? p->num_material_sets : 1);
    double* acc = (double*)calloc(nmat, sizeof(double));
    uint64_t nc = 0, na = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nc, na)
    for (int y = p->row0; y < p->row1; ++y) {
        if (!row_in_shard(p, y)) continue;
        zo_scene local = *s; local.n_closest = 0; local.n_any = 0;
        local.pdf_env_count = (p->flags & ZDR_FLAG_LIGHT_PDF_COUNTS_ENV) ? s->env_count : 0;
        Ctx c = {&local, {material, acc, p->tex_res_x, p->tex_res_y, p->bilinear}, p};
        for (int x = 0; x < p->width; ++x) {
            const float* g = d_image + ((size_t)x + (size_t)y * p->width) * 4;
            f3 le_grad = v3(g[0] / (float)p->spp, g[1] / (float)p->spp, g[2] / (float)p->spp);
            if (any_nan3(le_grad)) le_grad = v3(0, 0, 0);
            for (int it = p->it0; it < p->it1; ++it) {
                Sampler sm = make_sampler(x, y, p->seed, p->spp, it);
                Ray ray = camera_sample(p, x, y, &sm);
                switch (integrator) {
                    case ZDR_COLLOCATED: collocated_estimator_backward(&c, ray, &sm, le_grad); break;
                    case ZDR_DIRECT: direct_estimator_backward(&c, ray, &sm, le_grad); break;
                    default: path_estimator_backward(&c, ray, &sm, le_grad); break;
                }
            }
        }
        nc += local.n_closest; na += local.n_any;
    }
    for (size_t i = 0; i < nmat; ++i) d_material[i] = (float)((double)d_material[i] + acc[i]);
    free(acc);
    s->n_closest += nc; s->n_any += na;
    return ZDR_OK;
}

/* uvgrad.py:19-49 — screen-to-texture Jacobian (d u / d x, d v / d x, d u / d y, d v / d y) of the first
 * hit, by ray differentials.  Deviations: triangle corners are taken in WORLD space (the
 * reference mixes object-space corners with world-space rays, identical for identity transforms), and
 * the 3x3 products of uvgrad.py:43-46 are evaluated as iaTa * (aT * dp). */
static f4 trace_duvdxy(zo_scene* s, Ray ray, Ray rdx, Ray rdy) {
    f4 zero = {0, 0, 0, 0};
    s->n_closest++;
    Hit hit = trace_closest(s, ray);
    if (hit.inst < 0) return zero;
    Inst* in = &s->inst[hit.inst];
    int g = in->tri_base + hit.prim;
    f3 p0 = s->P[3 * g], p1 = s->P[3 * g + 1], p2 = s->P[3 * g + 2];
    const float* v0 = in->verts + 8 * in->tris[3 * hit.prim + 0];
    const float* v1 = in->verts + 8 * in->tris[3 * hit.prim + 1];
    const float* v2 = in->verts + 8 * in->tris[3 * hit.prim + 2];
    f3 p = interp3(&hit, p0, p1, p2);
    /* compute_dpduv, uvgrad.py:6-16 */
    f3 e1 = sub3(p1, p0), e2 = sub3(p2, p0);
    float a = v1[3] - v0[3], b = v1[4] - v0[4], c = v2[3] - v0[3], d = v2[4] - v0[4];   /* duvde columns (a,b), (c,d) */
    float det = a * d - c * b;
    float i00 = d / det, i10 = -b / det, i01 = -c / det, i11 = a / det;                  /* inverse, column-major */
    f3 dpdu = v3(fmaf(e2.x, i10, e1.x * i00), fmaf(e2.y, i10, e1.y * i00), fmaf(e2.z, i10, e1.z * i00));
    f3 dpdv = v3(fmaf(e2.x, i11, e1.x * i01), fmaf(e2.y, i11, e1.y * i01), fmaf(e2.z, i11, e1.z * i01));
    dpdv = neg3(dpdv);                                                                    /* inverted v */
    f3 ng = normalize3(cross3(e1, e2));
    float t_dx = dot3(sub3(p, rdx.o), ng) / dot3(rdx.d, ng);
    float t_dy = dot3(sub3(p, rdy.o), ng) / dot3(rdy.d, ng);
    f3 dpdx = sub3(v3(fmaf(t_dx, rdx.d.x, rdx.o.x), fmaf(t_dx, rdx.d.y, rdx.o.y), fmaf(t_dx, rdx.d.z, rdx.o.z)), p);
    f3 dpdy = sub3(v3(fmaf(t_dy, rdy.d.x, rdy.o.x), fmaf(t_dy, rdy.d.y, rdy.o.y), fmaf(t_dy, rdy.d.z, rdy.o.z)), p);
    float m00 = dot3(dpdu, dpdu), m01 = dot3(dpdu, dpdv), m11 = dot3(dpdv, dpdv);
    float mdet = m00 * m11 - m01 * m01;
    float j00 = m11 / mdet, j01 = -m01 / mdet, j11 = m00 / mdet;
    float bx0 = dot3(dpdu, dpdx), bx1 = dot3(dpdv, dpdx), by0 = dot3(dpdu, dpdy), by1 = dot3(dpdv, dpdy);
    f4 r = {fmaf(j01, bx1, j00 * bx0), fmaf(j11, bx1, j01 * bx0), fmaf(j01, by1, j00 * by0), fmaf(j11, by1, j01 * by0)};
    return r;
}
static Ray camera_ray_at(const zdr_params* p, int x, int y, f2 off, float ox, float oy) {
    f2 px;
    px.x = 2.0f / (float)p->width * ((float)x + off.x + ox) - 1.0f;
    px.y = 2.0f / (float)p->height * ((float)y + off.y + oy) - 1.0f;
    px.y *= (float)p->height / (float)p->width;
    return generate_ray(p, px);
}
/* uvgrad.py:75-98; sampler: the render kernels' pmj02bn sampler (integrator.py:17) instead of LuisaCompute's
 * built-in LCG (unpinned). out: (H,W,4) */
ZO_API int zo_render_duvdxy(zo_scene* s, const zdr_params* p, float* image) {
    int rc = check_render_args(s, ZDR_COLLOCATED, p);
    if (rc) return rc;
    if (!image) return ZDR_ERR_ARG;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = p->row0; y < p->row1; ++y) {
        if (!row_in_shard(p, y)) continue;
        zo_scene local = *s;
        for (int x = 0; x < p->width; ++x) {
            f4 acc = {0, 0, 0, 0};
            for (int it = p->it0; it < p->it1; ++it) {
                Sampler sm = make_sampler(x, y, p->seed, p->spp, it);
                f2 off = sampler_next2f(&sm);
                if (p->use_tent_filter) { off.x = tent_warp1(off.x) + 0.5f; off.y = tent_warp1(off.y) + 0.5f; }
                Ray ray = camera_ray_at(p, x, y, off, 0.0f, 0.0f);
                Ray rdx = camera_ray_at(p, x, y, off, 1.0f, 0.0f);
                Ray rdy = camera_ray_at(p, x, y, off, 0.0f, 1.0f);
                f4 g = trace_duvdxy(&local, ray, rdx, rdy);
                if (!(isnan(g.x) || isnan(g.y) || isnan(g.z) || isnan(g.w))) acc = add4(acc, g);
            }
            float* o = image + ((size_t)x + (size_t)y * p->width) * 4;
            o[0] = acc.x / (float)p->spp; o[1] = acc.y / (float)p->spp; o[2] = acc.z / (float)p->spp; o[3] = acc.w / (float)p->spp;
        }
    }
    return ZDR_OK;
}

/* rays: n x 18 (o,d of the centre ray, of the x-offset ray, of the y-offset ray) -> n x 4 */
ZO_API int zo_probe_duvdxy(zo_scene* s, const float* rays, int n, float* out) {
    if (!s || !s->committed) return ZDR_ERR_STATE;
    for (int i = 0; i < n; ++i) {
        const float* r = rays + 18 * i;
        Ray a = {v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), 0.0f, 1e30f};
        Ray bx = {v3(r[6], r[7], r[8]), v3(r[9], r[10], r[11]), 0.0f, 1e30f};
        Ray by = {v3(r[12], r[13], r[14]), v3(r[15], r[16], r[17]), 0.0f, 1e30f};
        f4 g = trace_duvdxy(s, a, bx, by);
        out[4 * i] = g.x; out[4 * i + 1] = g.y; out[4 * i + 2] = g.z; out[4 * i + 3] = g.w;
    }
    return ZDR_OK;
}

ZO_API void zo_get_ray_counters(zo_scene* s, uint64_t* out) { out[0] = s->n_closest; out[1] = s->n_any; }
ZO_API void zo_reset_ray_counters(zo_scene* s) { s->n_closest = 0; s->n_any = 0; }
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm of bench.py asks for all host threads explicitly */
ZO_API void zo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
ZO_API int zo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* envmap.py host side (create_alias_table :17-57, load_envmap :116-202)                */
/* ------------------------------------------------------------------------------------ */
/* envmap.py:17-57, float64 arithmetic like the Python original */
static void create_alias_table(const double* values, int n, float* prob_out, int* alias_out, double* pdf_out) {
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += fabs(values[i]);
    for (int i = 0; i < n; ++i) pdf_out[i] = sum == 0.0 ? 1.0 / n : fabs(values[i]) / sum;
    double ratio = sum > 0.0 ? n / sum : 1.0;
    double* prob = (double*)malloc(sizeof(double) * n);
    int* over = (int*)malloc(sizeof(int) * n); int* under = (int*)malloc(sizeof(int) * n);
    int no = 0, nu = 0;
    for (int i = 0; i < n; ++i) {
        double p = values[i] * ratio;
        prob[i] = p; alias_out[i] = i;
        if (p > 1.0) over[no++] = i; else if (p < 1.0) under[nu++] = i;
    }
    while (no > 0 && nu > 0) {
        int o = over[--no], u = under[--nu];
        prob[o] -= 1.0 - prob[u];
        alias_out[u] = o;
        if (prob[o] > 1.0) over[no++] = o; else if (prob[o] < 1.0) under[nu++] = o;
    }
    for (int i = 0; i < no; ++i) { prob[over[i]] = 1.0; alias_out[over[i]] = over[i]; }
    for (int i = 0; i < nu; ++i) { prob[under[i]] = 1.0; alias_out[under[i]] = under[i]; }
    for (int i = 0; i < n; ++i) prob_out[i] = (float)prob[i];
    free(prob); free(over); free(under);
}
ZO_API void zo_create_alias_table(const double* values, int n, float* prob_out, int* alias_out, double* pdf_out) {
    create_alias_table(values, n, prob_out, alias_out, pdf_out);
}

/* envmap.py:116-202 */
ZO_API int zo_scene_set_envmap(zo_scene* s, const float* rgba, int w, int h, int compensate_mis) {
    if (!s || !rgba || w <= 0 || h <= 0) return ZDR_ERR_ARG;
    int hh = h;
    if (h != w) { if (w == 2 * h) hh = 2 * h; else return ZDR_ERR_ARG; }   /* envmap.py:122-128 */
    free(s->env_tex);
    s->env_tex = (float*)malloc(sizeof(float) * 4 * (size_t)w * hh);
    for (int y = 0; y < hh; ++y) {
        int sy = (hh == h) ? y : y / 2;                                   /* img.repeat(2, axis=0) */
        memcpy(s->env_tex + 4 * (size_t)y * w, rgba + 4 * (size_t)sy * w, sizeof(float) * 4 * w);
    }
    s->env_w = w; s->env_h = hh;
    const int W = ENV_MAP_W, H = ENV_MAP_H, pixel_count = W * H;
    float* scale_map = (float*)malloc(sizeof(float) * pixel_count);
    /* generate_weight_map_kernel, envmap.py:136-158 */
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            float cx = (float)px + 0.5f, cy = (float)py + 0.5f;
            float sum_weight = 0.0f, sum_scale = 0.0f;
            const float filter_step = 0.125f;
            const int n = 8;
            for (int dy = -n; dy <= n; ++dy)
                for (int dx = -n; dx <= n; ++dx) {
                    float ox = (float)dx * filter_step, oy = (float)dy * filter_step;
                    f2 uv = {(cx + ox) / (float)W, (cy + oy) / (float)H};
                    f3 c = env_texture(s, uv);
                    float scale = 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z;
                    float sn, cs; zo_sincos(uv.y * PI_F, &sn, &cs);
                    float weight = expf(-4.0f * fmaf(oy, oy, ox * ox));
                    float value = weight * fminf(scale * sn, 1e8f);
                    sum_weight += weight;
                    sum_scale += value;
                }
            scale_map[py * W + px] = sum_scale / sum_weight;
        }
    free(s->env_scale_map);
    s->env_scale_map = (float*)malloc(sizeof(float) * pixel_count);
    memcpy(s->env_scale_map, scale_map, sizeof(float) * pixel_count);
    /* envmap.py:167-175 */
    if (compensate_mis) {
        double mean = 0.0;
        for (int i = 0; i < pixel_count; ++i) mean += scale_map[i];
        float average_scale = (float)(mean / pixel_count);
        double wavg = 0.0;
        for (int y = 0; y < H; ++y) wavg += sin((y + 0.5) / H * M_PI);
        wavg /= H;
        for (int y = 0; y < H; ++y) {
            double rw = sin((y + 0.5) / H * M_PI);
            float sub = (float)((double)average_scale * rw / wavg);
            for (int x = 0; x < W; ++x) {
                float v = scale_map[y * W + x] - sub;
                scale_map[y * W + x] = v > 0.0f ? v : 0.0f;
            }
        }
    }
    /* envmap.py:177-194 */
    free(s->env_alias_prob); free(s->env_alias_idx); free(s->env_pdf);
    s->env_alias_prob = (float*)malloc(sizeof(float) * (H + pixel_count));
    s->env_alias_idx = (int*)malloc(sizeof(int) * (H + pixel_count));
    s->env_pdf = (float*)malloc(sizeof(float) * pixel_count);
    double* pdfs = (double*)malloc(sizeof(double) * pixel_count);
    double* row = (double*)malloc(sizeof(double) * W);
    double* row_avg = (double*)malloc(sizeof(double) * H);
    double* pdf_y = (double*)malloc(sizeof(double) * H);
    for (int y = 0; y < H; ++y) {
        float m = 0.0f;   /* numpy float32 mean (pairwise in numpy; plain here — see DESIGN.md) */
        double md = 0.0;
        for (int x = 0; x < W; ++x) { row[x] = scale_map[y * W + x]; md += row[x]; }
        m = (float)(md / W);
        row_avg[y] = m;
        create_alias_table(row, W, s->env_alias_prob + H + y * W, s->env_alias_idx + H + y * W, pdfs + y * W);
    }
    create_alias_table(row_avg, H, s->env_alias_prob, s->env_alias_idx, pdf_y);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) s->env_pdf[y * W + x] = (float)(pdfs[y * W + x] * pdf_y[y] * pixel_count);
    free(pdfs); free(row); free(row_avg); free(pdf_y); free(scale_map);
    s->env_count = 1;
    return ZDR_OK;
}
ZO_API int zo_scene_get_env_scale_map(zo_scene* s, float* out) {
    if (!s || !s->env_scale_map || !out) return ZDR_ERR_STATE;
    memcpy(out, s->env_scale_map, sizeof(float) * ENV_MAP_W * ENV_MAP_H);
    return ZDR_OK;
}
/* envmap.py:223-238 on n inputs: out n x 8 = wi.xyz pdf eval.xyz dist */
ZO_API int zo_probe_sample_envmap(zo_scene* s, const float* u, int n, float* out) {
    if (!s || !s->env_count) return ZDR_ERR_STATE;
    for (int i = 0; i < n; ++i) {
        f2 uu = {u[2 * i], u[2 * i + 1]};
        LightSample t = sample_envmap(s, uu);
        float* o = out + 8 * (size_t)i;
        o[0] = t.wi.x; o[1] = t.wi.y; o[2] = t.wi.z; o[3] = t.pdf; o[4] = t.eval.x; o[5] = t.eval.y; o[6] = t.eval.z; o[7] = t.dist;
    }
    return ZDR_OK;
}
/* envmap.py:240-248 */
ZO_API int zo_probe_env_pdf(zo_scene* s, const float* dir, int n, float* out) {
    if (!s || !s->env_count) return ZDR_ERR_STATE;
    for (int i = 0; i < n; ++i) out[i] = env_sampled_light_pdf(s, v3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
    return ZDR_OK;
}
/* envmap.py:206-219: in n x 5 = uv.xy dir.xyz -> out n x 5 = uv_to_direction(uv).xyz direction_to_uv(dir).xy */
ZO_API void zo_probe_env_mapping(const float* in, int n, float* out) {
    for (int i = 0; i < n; ++i) {
        f2 uv = {in[5 * i], in[5 * i + 1]};
        f3 d = uv_to_direction(uv);
        f2 r = direction_to_uv(v3(in[5 * i + 2], in[5 * i + 3], in[5 * i + 4]));
        out[5 * i] = d.x; out[5 * i + 1] = d.y; out[5 * i + 2] = d.z; out[5 * i + 3] = r.x; out[5 * i + 4] = r.y;
    }
}
/* read back the envmap sampling tables (tests compare them with the GPU build) */
ZO_API int zo_scene_get_env_tables(zo_scene* s, float* prob, int* alias, float* pdf) {
    if (!s || !s->env_count) return ZDR_ERR_STATE;
    int n = ENV_MAP_H + ENV_MAP_W * ENV_MAP_H;
    memcpy(prob, s->env_alias_prob, sizeof(float) * n);
    memcpy(alias, s->env_alias_idx, sizeof(int) * n);
    memcpy(pdf, s->env_pdf, sizeof(float) * ENV_MAP_W * ENV_MAP_H);
    return ZDR_OK;
}

/* ------------------------------------------------------------------------------------ */
/* function-level probes for the unit tests                                             */
/* ------------------------------------------------------------------------------------ */
/* state: n x (px,py,seed,spp,sample_index); prog: '1' next(), '2' next2f() */
ZO_API int zo_probe_sampler(const int32_t* state, int n, const char* prog, float* out) {
    if (!g_pmj || !g_bn) return ZDR_ERR_STATE;
    size_t per = 0;
    for (const char* q = prog; *q; ++q) per += (*q == '2') ? 2 : 1;
    for (int i = 0; i < n; ++i) {
        const int32_t* st = state + 5 * i;
        Sampler sm = make_sampler(st[0], st[1], st[2], st[3], st[4]);
        float* o = out + per * i;
        for (const char* q = prog; *q; ++q) {
            if (*q == '2') { f2 u = sampler_next2f(&sm); *o++ = u.x; *o++ = u.y; }
            else *o++ = sampler_next(&sm);
        }
    }
    return ZDR_OK;
}
ZO_API int zo_probe_trace_closest(zo_scene* s, const float* rays, int n, int32_t* inst_prim, float* uvt) {
    if (!s || !s->committed) return ZDR_ERR_STATE;
    for (int i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray = {v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), r[6], r[7]};
        Hit h = trace_closest(s, ray);
        inst_prim[2 * i] = h.inst; inst_prim[2 * i + 1] = h.prim;
        uvt[3 * i] = h.u; uvt[3 * i + 1] = h.v; uvt[3 * i + 2] = h.t;
    }
    return ZDR_OK;
}
ZO_API int zo_probe_trace_any(zo_scene* s, const float* rays, int n, int32_t* occ) {
    if (!s || !s->committed) return ZDR_ERR_STATE;
    for (int i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray = {v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), r[6], r[7]};
        occ[i] = trace_any(s, ray);
    }
    return ZDR_OK;
}
/* in n x 13: wo wi diffuse roughness g ; out n x 7: f.xyz dmat.xyzw */
ZO_API void zo_probe_ggx(const float* in, int n, float* out) {
    for (int i = 0; i < n; ++i) {
        const float* a = in + 13 * i;
        f3 wo = v3(a[0], a[1], a[2]), wi = v3(a[3], a[4], a[5]), d = v3(a[6], a[7], a[8]), g = v3(a[10], a[11], a[12]);
        f3 f = ggx_brdf(wo, wi, d, 0.04f, a[9]);
        f4 dm = ggx_brdf_grad(wo, wi, 0.04f, a[9], g);
        float* o = out + 7 * i;
        o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = dm.x; o[4] = dm.y; o[5] = dm.z; o[6] = dm.w;
    }
}
/* in n x 6: wo.xyz roughness u1... ; generic probes used by the golden tests */
ZO_API void zo_probe_ggx_sample_pdf(const float* in, int n, float* out) {   /* n x 7: wo wi r -> pdf */
    for (int i = 0; i < n; ++i) {
        const float* a = in + 7 * i;
        out[i] = ggx_sample_pdf(v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), a[6]);
    }
}
ZO_API void zo_probe_sample_wm(const float* in, int n, float* out) {        /* n x 6: w alpha u.xy -> wm */
    for (int i = 0; i < n; ++i) {
        const float* a = in + 6 * i;
        f2 u = {a[4], a[5]};
        f3 r = sample_wm(v3(a[0], a[1], a[2]), a[3], u);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
ZO_API void zo_probe_cosine_hemisphere(const float* in, int n, float* out) { /* n x 2 -> n x 3 */
    for (int i = 0; i < n; ++i) {
        f2 u = {in[2 * i], in[2 * i + 1]};
        f3 r = cosine_sample_hemisphere(u);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
ZO_API void zo_probe_make_onb(const float* in, int n, float* out) {          /* n x 3 -> n x 9 (t,b,n) */
    for (int i = 0; i < n; ++i) {
        Onb o = make_onb(v3(in[3 * i], in[3 * i + 1], in[3 * i + 2]));
        float* q = out + 9 * i;
        q[0] = o.t.x; q[1] = o.t.y; q[2] = o.t.z; q[3] = o.b.x; q[4] = o.b.y; q[5] = o.b.z; q[6] = o.n.x; q[7] = o.n.y; q[8] = o.n.z;
    }
}
ZO_API void zo_probe_tent_warp(const float* in, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = tent_warp1(in[i]);
}
ZO_API void zo_probe_generate_ray(const zdr_params* p, const float* ndc, int n, float* out) { /* n x 2 -> n x 6 */
    for (int i = 0; i < n; ++i) {
        f2 q = {ndc[2 * i], ndc[2 * i + 1]};
        Ray r = generate_ray(p, q);
        float* o = out + 6 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
    }
}
ZO_API void zo_probe_uniform_triangle(const float* in, int n, float* out) {   /* n x 2 -> n x 3 */
    for (int i = 0; i < n; ++i) {
        f2 u = {in[2 * i], in[2 * i + 1]};
        f3 r = sample_uniform_triangle(u);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
ZO_API void zo_probe_sincos(const float* in, int n, float* out) {            /* n -> n x 2 (sin, cos) */
    for (int i = 0; i < n; ++i) zo_sincos(in[i], &out[2 * i], &out[2 * i + 1]);
}
ZO_API void zo_probe_offset_ray_origin(const float* in, int n, float* out) { /* n x 6 (p, n) -> n x 3 */
    for (int i = 0; i < n; ++i) {
        f3 r = offset_ray_origin(v3(in[6 * i], in[6 * i + 1], in[6 * i + 2]), v3(in[6 * i + 3], in[6 * i + 4], in[6 * i + 5]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
/* read_bsdf on a host texture: uv n x 2 -> n x 4 */
ZO_API void zo_probe_read_bsdf(const float* mat, int rx, int ry, int bilinear, const float* uv, int n, float* out) {
    Tex t = {mat, NULL, rx, ry, bilinear};
    for (int i = 0; i < n; ++i) {
        f2 q = {uv[2 * i], uv[2 * i + 1]};
        f4 r = read_bsdf(&t, q);
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
}
/* surface_interact for given hits: in n x (inst, prim) + n x (u,v) -> n x 11 (p, uv, ns, ng) */
ZO_API int zo_probe_surface_interact(zo_scene* s, const int32_t* ip, const float* uv, int n, float* out) {
    if (!s || !s->committed) return ZDR_ERR_STATE;
    for (int i = 0; i < n; ++i) {
        Hit h = {ip[2 * i], ip[2 * i + 1], uv[2 * i], uv[2 * i + 1], 0.0f};
        Interaction it = surface_interact(s, &h);
        float* o = out + 11 * i;
        o[0] = it.p.x; o[1] = it.p.y; o[2] = it.p.z; o[3] = it.uv.x; o[4] = it.uv.y;
        o[5] = it.ns.x; o[6] = it.ns.y; o[7] = it.ns.z; o[8] = it.ng.x; o[9] = it.ng.y; o[10] = it.ng.z;
    }
    return ZDR_OK;
}
