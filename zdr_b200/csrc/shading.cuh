// shading.cuh — device function library of the render kernels: surface interaction, texture
// fetch / gradient scatter, GGX microfacet BRDF (value, analytic adjoint, sampling, pdf),
// mesh-light and envmap sampling.  Reference: interaction.py, onb.py, microfacet.py, light.py,
// envmap.py.  Operation order follows the numeric contract (zdr_math.cuh).
#pragma once
#include "sampler.cuh"
#include "traverse.cuh"

namespace zdr {

// ---------------------------------------------------------------------------------------------
// onb.py
// ---------------------------------------------------------------------------------------------
struct Onb { float3 t, b, n; };
ZDR_DEV Onb make_onb(float3 n) {                                   // onb.py:21-28
    Onb r;
    r.b = normalize(fabsf(n.x) > fabsf(n.z) ? f3(-n.y, n.x, 0.0f) : f3(0.0f, -n.z, n.y));
    r.t = normalize(cross(r.b, n));
    r.n = n;
    return r;
}
ZDR_DEV float3 to_world(const Onb& o, float3 v) { return lincomb(v.x, o.t, v.y, o.b, v.z, o.n); }      // onb.py:10
ZDR_DEV float3 to_local(const Onb& o, float3 v) { return f3(dot(v, o.t), dot(v, o.b), dot(v, o.n)); }  // onb.py:14

// ---------------------------------------------------------------------------------------------
// interaction.py
// ---------------------------------------------------------------------------------------------
struct Interaction { float3 p; float2 uv; float3 ns, ng; };       // interaction.py:6

// interaction.py:8-30.  p0..p2 are the pre-transformed world corners (tris[] of the hit slot),
// normals are interpolated in object space and then mapped by inverse(transpose(M3x3)).
ZDR_DEV Interaction surface_interact(const SceneView& sc, const Hit& h) {
    const float4* tp = sc.tris + 3 * (size_t)h.slot;
    const float4* sp = sc.shade + 4 * (size_t)h.slot;
    const float4 a = ld_scene(tp), b = ld_scene(tp + 1), c = ld_scene(tp + 2);
    const float4 s0 = ld_scene(sp), s1 = ld_scene(sp + 1), s2 = ld_scene(sp + 2), s3 = ld_scene(sp + 3);
    const float3 p0 = f3(a.x, a.y, a.z), p1 = f3(b.x, b.y, b.z), p2 = f3(c.x, c.y, c.z);
    const float w = 1.0f - h.u - h.v;
    Interaction it;
    it.p = lincomb(w, p0, h.u, p1, h.v, p2);
    it.uv.x = fmaf(h.v, s2.w, fmaf(h.u, s1.w, w * s0.w));
    it.uv.y = fmaf(h.v, s3.z, fmaf(h.u, s3.y, w * s3.x));
    const float3 ns0 = lincomb(w, f3(s0.x, s0.y, s0.z), h.u, f3(s1.x, s1.y, s1.z), h.v, f3(s2.x, s2.y, s2.z));
    it.ns = normalize(mat3_mul(sc.inst[h.inst].N, ns0));
    it.ng = normalize(cross(p1 - p0, p2 - p0));
    return it;
}

struct TexView {
    const float4* __restrict__ mat;   // (ry rows? no: index = x + rx*y) float4 texels
    float4* __restrict__ dmat;        // gradient accumulator (backward only)
    int rx, ry;
    int bilinear;
    int set_texels;                   // texels of one texture set (rx * ry): set k starts at k * set_texels ("polymorphic materials")
};
// Material sets (README.md:5 "currently only one set of material"): the texture buffer holds K sets of (ry, rx) texels; an
// instance is shaded with the set its InstData names.  -1 = the reference's "treat everything as light other than the first
// object" (direct.py:30, prb.py:45); stages that upstream shade any instance (collocated.py:22, the literal PRB adjoint)
// fall back to set 0.
ZDR_DEV int inst_material_set(const SceneView& sc, int inst) { return sc.inst[inst].material_set; }
ZDR_DEV int inst_texture_set(const SceneView& sc, int inst) { const int m = sc.inst[inst].material_set; return m < 0 ? 0 : m; }

ZDR_DEV float4 read_single_bsdf(const TexView& t, int cx, int cy, int base) {      // interaction.py:35-44
    cx = clampi(cx, 0, t.rx - 1); cy = clampi(cy, 0, t.ry - 1);
    return __ldg(t.mat + (base + cx + t.rx * cy));
}
ZDR_DEV float4 read_bsdf(const TexView& t, float2 uv, int set = 0) {       // interaction.py:46-60
    const int base = set * t.set_texels;
    const float px = uv.x * (float)(t.rx - 1), py = (1.0f - uv.y) * (float)(t.ry - 1);
    if (!t.bilinear) return read_single_bsdf(t, f2i(px + 0.4999f), f2i(py + 0.4999f), base);
    const int ix = f2i(px), iy = f2i(py);
    const float ox = px - (float)ix, oy = py - (float)iy;
    const float4 c00 = read_single_bsdf(t, ix, iy, base), c01 = read_single_bsdf(t, ix, iy + 1, base);
    const float4 c10 = read_single_bsdf(t, ix + 1, iy, base), c11 = read_single_bsdf(t, ix + 1, iy + 1, base);
    return lerp(lerp(c00, c01, oy), lerp(c10, c11, oy), ox);
}

// read_bsdf in two halves (bilinear only): the four texel loads, and their interpolation — a stage can put independent work
// between them (experiment ZDR_WF_EARLY_TEXELS)
struct BsdfTexels { float4 c00, c01, c10, c11; float ox, oy; };
ZDR_DEV BsdfTexels fetch_bsdf(const TexView& t, float2 uv, int set = 0) {
    const int base = set * t.set_texels;
    const float px = uv.x * (float)(t.rx - 1), py = (1.0f - uv.y) * (float)(t.ry - 1);
    const int ix = f2i(px), iy = f2i(py);
    BsdfTexels r;
    r.ox = px - (float)ix; r.oy = py - (float)iy;
    r.c00 = read_single_bsdf(t, ix, iy, base); r.c01 = read_single_bsdf(t, ix, iy + 1, base);
    r.c10 = read_single_bsdf(t, ix + 1, iy, base); r.c11 = read_single_bsdf(t, ix + 1, iy + 1, base);
    return r;
}
ZDR_DEV float4 combine_bsdf(const BsdfTexels& r) { return lerp(lerp(r.c00, r.c01, r.oy), lerp(r.c10, r.c11, r.oy), r.ox); }

// one 16-byte vector reduction per texel instead of the reference's four scalar
// atomic_fetch_add (interaction.py:67-70): red.global.add.v4.f32 (sm_90+)
ZDR_DEV void red_add_f4(float4* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// Warp-aggregated scatter (replaces the contended atomic_fetch_add the reference's README flags).
// Lanes of the warp whose bilinear footprint starts at the SAME texel are found with ONE match.any
// on the footprint key; a lane that is alone (the common case on a uv-mapped, high-resolution
// texture) issues its four red.v4 directly, so the uncontended path costs one MATCH.  Lanes that
// share a footprint — neighbouring pixels' primary hits, and whole warps when the texture is
// uniform or the mesh has no uv (C2/C4: uv == 0) — sum their contributions in lane order through
// shuffles and the group's lowest lane issues the reductions: 32 serialised L2 atomics on one
// address become one.  Exact-zero contributions are not issued (adding 0.0 is the identity).
// Returns the number of red.v4 this lane issued.
ZDR_DEV bool is_zero4(float4 v) { return v.x == 0.0f && v.y == 0.0f && v.z == 0.0f && v.w == 0.0f; }
ZDR_DEV float4 group_sum4(unsigned peers, bool full, float4 v) {
    if (full) {        // all 32 lanes share the key: butterfly
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
        }
        return v;
    }
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned m = peers; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        sum.x += __shfl_sync(peers, v.x, src); sum.y += __shfl_sync(peers, v.y, src);
        sum.z += __shfl_sync(peers, v.z, src); sum.w += __shfl_sync(peers, v.w, src);
    }
    return sum;
}
ZDR_DEV int texel_index(const TexView& t, int cx, int cy) {                // interaction.py:37-39, address mode CLAMP
    return clampi(cx, 0, t.rx - 1) + t.rx * clampi(cy, 0, t.ry - 1);
}
ZDR_DEV float4 scale4(float k, float4 d) { return make_float4(k * d.x, k * d.y, k * d.z, k * d.w); }
// Block-level stage of the contended scatter (north star item 6: "per-SM duplicated gradient buffers"; the README's
// "Need to add duplicate buffer for that", README.md:21).  Warp aggregation stops at 32 lanes: on a uniform / uv-less
// texture every warp of every CTA still sends its sum to the SAME texel, and same-address reductions serialise in the L2
// atomic unit (0.65 G red.v4/s, profiles/r01q_microbench.json).  Each CTA therefore owns a small duplicate of the gradient
// buffer in SHARED memory — an open-addressing table of kGradCacheSlots (texel, float4) entries — that the group leaders add
// into with shared-memory atomics; it is folded into the global buffer ONCE, when the CTA ends (one red.v4 per distinct
// texel the CTA touched).  The persistent kernel keeps one CTA per SM alive for the whole pass, so there the table is
// literally a per-SM duplicate buffer.  A full table (more than ~kGradCacheSlots distinct texels in one CTA) falls back to
// the direct reduction.
constexpr int kGradCacheSlots = 256;
constexpr int kGradCacheBytes = kGradCacheSlots * 20;      // keys (int) + 4 floats per slot
struct GradCache {
    int* key;        // texel index, -1 = empty
    float* acc;      // 4 floats per slot
};
ZDR_DEV GradCache grad_cache_null() { GradCache c; c.key = nullptr; c.acc = nullptr; return c; }
// all threads of the CTA; ends with a barrier
ZDR_DEV void grad_cache_init(GradCache c) {
    for (int i = threadIdx.x; i < kGradCacheSlots; i += blockDim.x) {
        c.key[i] = -1;
        c.acc[4 * i] = 0.f; c.acc[4 * i + 1] = 0.f; c.acc[4 * i + 2] = 0.f; c.acc[4 * i + 3] = 0.f;
    }
    __syncthreads();
}
ZDR_DEV bool grad_cache_add(GradCache c, int texel, float4 v) {
    const unsigned h = ((unsigned)texel * 2654435761u) >> 24;
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        const int s = (int)((h + (unsigned)p) & (unsigned)(kGradCacheSlots - 1));
        const int old = atomicCAS(c.key + s, -1, texel);
        if (old == -1 || old == texel) {
            float* a = c.acc + 4 * s;
            if (v.x != 0.0f) atomicAdd(a, v.x);
            if (v.y != 0.0f) atomicAdd(a + 1, v.y);
            if (v.z != 0.0f) atomicAdd(a + 2, v.z);
            if (v.w != 0.0f) atomicAdd(a + 3, v.w);
            return true;
        }
    }
    return false;
}
// all threads of the CTA, after its last scatter; returns the number of red.v4 this thread issued
ZDR_DEV unsigned grad_cache_flush(GradCache c, float4* dmat) {
    __syncthreads();
    unsigned issued = 0;
    for (int i = threadIdx.x; i < kGradCacheSlots; i += blockDim.x) {
        const int texel = c.key[i];
        if (texel >= 0) {
            const float4 v = make_float4(c.acc[4 * i], c.acc[4 * i + 1], c.acc[4 * i + 2], c.acc[4 * i + 3]);
            if (!is_zero4(v)) { red_add_f4(dmat + texel, v); ++issued; }
        }
    }
    return issued;
}
ZDR_DEV unsigned scatter_one(const GradCache& c, float4* dmat, int texel, float4 v) {
    if (c.key && grad_cache_add(c, texel, v)) return 0u;
    red_add_f4(dmat + texel, v);
    return 1u;
}
static __device__ __noinline__ unsigned write_bsdf_grad_agg_nl(float4* __restrict__ dmat, int rx, int ry, int bilinear, float uvx, float uvy,
                                                            float dx, float dy, float dz, float dw, int* cache_key, float* cache_acc, int base) {  // interaction.py:62-89
    TexView t;
    t.mat = nullptr; t.dmat = dmat; t.rx = rx; t.ry = ry; t.bilinear = bilinear;
    GradCache cache;
    cache.key = cache_key; cache.acc = cache_acc;
    const float2 uv = make_float2(uvx, uvy);
    const float4 d = make_float4(dx, dy, dz, dw);
    const float px = uv.x * (float)(t.rx - 1), py = (1.0f - uv.y) * (float)(t.ry - 1);
    int idx[4];
    float4 v[4];
    int n;
    if (!t.bilinear) {
        n = 1;
        idx[0] = base + texel_index(t, f2i(px + 0.4999f), f2i(py + 0.4999f));
        v[0] = d;
    } else {
        n = 4;
        const int ix = f2i(px), iy = f2i(py);
        const float ox = px - (float)ix, oy = py - (float)iy;
        idx[0] = base + texel_index(t, ix, iy);         v[0] = scale4((1.0f - ox) * (1.0f - oy), d);
        idx[1] = base + texel_index(t, ix, iy + 1);     v[1] = scale4((1.0f - ox) * oy, d);
        idx[2] = base + texel_index(t, ix + 1, iy);     v[2] = scale4(ox * (1.0f - oy), d);
        idx[3] = base + texel_index(t, ix + 1, iy + 1); v[3] = scale4(ox * oy, d);
    }
    const unsigned active = __activemask();
    // lanes with equal first texel and equal far-corner texel have identical footprints
    const unsigned peers = __match_any_sync(active, (unsigned long long)(unsigned)idx[0] << 32 | (unsigned)idx[n - 1]);
    const int lane = threadIdx.x & 31;
    unsigned issued = 0;
    if (peers == (1u << lane)) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < n && !is_zero4(v[k])) issued += scatter_one(cache, t.dmat, idx[k], v[k]);
        return issued;
    }
    const bool full = peers == 0xffffffffu;
    const bool leader = lane == __ffs(peers) - 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < n) {
            const float4 sum = group_sum4(peers, full, v[k]);
            if (leader && !is_zero4(sum)) issued += scatter_one(cache, t.dmat, idx[k], sum);
        }
    }
    return issued;
}

ZDR_DEV unsigned write_bsdf_grad_aggregated(const TexView& t, float2 uv, float4 d, GradCache cache = grad_cache_null(), int set = 0) {
    return write_bsdf_grad_agg_nl(t.dmat, t.rx, t.ry, t.bilinear, uv.x, uv.y, d.x, d.y, d.z, d.w, cache.key, cache.acc, set * t.set_texels);
}
// Plain scatter: one red.v4 per texel of the footprint (the reference issues four scalar atomics per
// texel, interaction.py:67-70).  Used when the footprints of a warp are spread over the texture
// (uv-mapped mesh, high-resolution texture): measured on C3, 3.81 of 4 reductions per scatter
// survive aggregation, so the match + shuffle work is pure overhead there.
ZDR_DEV unsigned write_bsdf_grad(const TexView& t, float2 uv, float4 d, int set = 0) {
    float4* dm = t.dmat + set * t.set_texels;
    const float px = uv.x * (float)(t.rx - 1), py = (1.0f - uv.y) * (float)(t.ry - 1);
    if (!t.bilinear) { red_add_f4(dm + texel_index(t, f2i(px + 0.4999f), f2i(py + 0.4999f)), d); return 1u; }
    const int ix = f2i(px), iy = f2i(py);
    const float ox = px - (float)ix, oy = py - (float)iy;
    red_add_f4(dm + texel_index(t, ix, iy), scale4((1.0f - ox) * (1.0f - oy), d));
    red_add_f4(dm + texel_index(t, ix, iy + 1), scale4((1.0f - ox) * oy, d));
    red_add_f4(dm + texel_index(t, ix + 1, iy), scale4(ox * (1.0f - oy), d));
    red_add_f4(dm + texel_index(t, ix + 1, iy + 1), scale4(ox * oy, d));
    return 4u;
}

// ---------------------------------------------------------------------------------------------
// microfacet.py
// ---------------------------------------------------------------------------------------------
ZDR_DEV float ggx_distribution(float3 h, float alpha) {                  // microfacet.py:6-11
    const float alpha2 = alpha * alpha;
    const float nh = fmaxf(0.00001f, h.z);
    const float t = nh * nh * (alpha2 - 1.0f) + 1.0f;
    return alpha2 / (kPi * (t * t));
}
ZDR_DEV float fresnel_schlick(float c, float specular) {                 // microfacet.py:13-15
    const float x = 1.0f - c, x2 = x * x;
    return specular + (1.0f - specular) * (x2 * x2 * x);
}
ZDR_DEV float smith_geometry(float3 v, float alpha) {                    // microfacet.py:17-21
    const float alpha2 = alpha * alpha;
    const float nv = fmaxf(0.00001f, v.z);
    return 2.0f / (1.0f + sqrtf(1.0f + alpha2 * (1.0f - nv * nv) / (nv * nv)));
}
ZDR_DEV float3 ggx_brdf(float3 wo, float3 wi, float3 diffuse, float specular, float roughness) {   // microfacet.py:23-30
    const float alpha = roughness * roughness;
    const float3 h = normalize(wi + wo);
    const float d = ggx_distribution(h, alpha);
    const float f = fresnel_schlick(clampf(dot(wo, h), 0.00001f, 1.0f), specular);
    const float g = smith_geometry(wi, alpha) * smith_geometry(wo, alpha);
    const float spec = (d * f * g) / (4.0f * fmaxf(0.00001f, wi.z) * fmaxf(0.00001f, wo.z));
    return f3((spec + diffuse.x / kPi) * wi.z, (spec + diffuse.y / kPi) * wi.z, (spec + diffuse.z / kPi) * wi.z);
}
// Analytic adjoint of ggx_brdf wrt mat = (diffuse.xyz, roughness) for cotangent g: replaces
// LuisaCompute's in-kernel reverse-mode AD (collocated.py:44-56, direct.py:126-131,160-165,
// prb.py:138-146,157-163).  With a2 = r^4: dD/da2 = (t - 2 a2 nh^2)/(pi t^3),
// dG1/da2 = -k/(s (1+s)^2), k = (1-nv^2)/nv^2, s = sqrt(1 + a2 k), da2/dr = 4 r^3.
ZDR_DEV float4 ggx_brdf_grad(float3 wo, float3 wi, float specular, float roughness, float3 g) {
    const float alpha = roughness * roughness;
    const float a2 = alpha * alpha;
    const float3 h = normalize(wi + wo);
    const float nh = fmaxf(0.00001f, h.z);
    const float nh2 = nh * nh;
    const float t = nh2 * (a2 - 1.0f) + 1.0f;
    const float D = a2 / (kPi * (t * t));
    const float dD = (t - 2.0f * a2 * nh2) / (kPi * (t * t * t));
    const float F = fresnel_schlick(clampf(dot(wo, h), 0.00001f, 1.0f), specular);
    const float nvi = fmaxf(0.00001f, wi.z), nvo = fmaxf(0.00001f, wo.z);
    const float ki = (1.0f - nvi * nvi) / (nvi * nvi), ko = (1.0f - nvo * nvo) / (nvo * nvo);
    const float si = sqrtf(1.0f + a2 * ki), so = sqrtf(1.0f + a2 * ko);
    const float Gi = 2.0f / (1.0f + si), Go = 2.0f / (1.0f + so);
    const float dGi = -ki / (si * ((1.0f + si) * (1.0f + si)));
    const float dGo = -ko / (so * ((1.0f + so) * (1.0f + so)));
    const float dDGG = dD * Gi * Go + D * dGi * Go + D * Gi * dGo;
    const float da2_dr = 4.0f * roughness * roughness * roughness;
    const float common = wi.z * F / (4.0f * nvi * nvo) * dDGG * da2_dr;
    const float dk = wi.z / kPi;
    return make_float4(g.x * dk, g.y * dk, g.z * dk, (g.x + g.y + g.z) * common);
}
// (ggx_brdf, ggx_sample_pdf) of one direction pair; one out-of-line copy shared by the
// light-sample and BSDF-sample evaluations (instruction-cache budget, see traverse.cuh)
#ifndef ZDR_INLINE_GGX_EVAL
#define ZDR_INLINE_GGX_EVAL 0
#endif
#if ZDR_INLINE_GGX_EVAL
#define ZDR_GGX_LINKAGE static __device__ __forceinline__
#else
#define ZDR_GGX_LINKAGE static __device__ __noinline__
#endif
ZDR_GGX_LINKAGE float4 ggx_eval(float wox, float woy, float woz, float wix, float wiy, float wiz, float dr, float dg, float db,
                                        float specular, float roughness);
// out-of-line ggx_brdf_grad (adjoint kernels)
static __device__ __noinline__ float4 ggx_brdf_grad_nl(float wox, float woy, float woz, float wix, float wiy, float wiz, float specular,
                                                float roughness, float gx, float gy, float gz);
ZDR_DEV float3 cosine_sample_hemisphere(float2 u) {                      // microfacet.py:33-37
    const float r = sqrtf(u.x);
    const float phi = kTwoPi * u.y;
    float s, c;
    sincos_contract(phi, &s, &c);
    return f3(r * c, r * s, sqrtf(1.0f - u.x));
}
ZDR_DEV float pdf_wm(float3 w, float3 wm, float alpha) {                 // microfacet.py:70-73
    return smith_geometry(w, alpha) / fabsf(w.z) * ggx_distribution(wm, alpha) * fabsf(dot(w, wm));
}
ZDR_DEV float3 sample_wm(float3 w, float alpha, float2 u) {              // microfacet.py:75-92
    float3 wh = normalize(f3(alpha * w.x, alpha * w.y, w.z));
    if (wh.z < 0.0f) wh = -wh;
    const float3 T1 = wh.z < 0.99999f ? normalize(cross(f3(0.0f, 0.0f, 1.0f), wh)) : f3(1.0f, 0.0f, 0.0f);
    const float3 T2 = cross(wh, T1);
    const float r = sqrtf(u.x);
    const float theta = kTwoPi * u.y;
    float sn, cs;
    sincos_contract(theta, &sn, &cs);
    const float px = r * cs;
    float py = r * sn;
    const float h = sqrtf(1.0f - px * px);
    py = lerp(h, py, (1.0f + wh.z) / 2.0f);
    const float pz = sqrtf(fmaxf(0.0f, 1.0f - fmaf(py, py, px * px)));
    const float3 nh = lincomb(px, T1, py, T2, pz, wh);
    return normalize(f3(alpha * nh.x, alpha * nh.y, fmaxf(1e-6f, nh.z)));
}
ZDR_DEV float3 ggx_sample(const SceneView& sc, float3 wo, float roughness, Sampler& sampler) {   // microfacet.py:40-49
    if (sampler_next(sc, sampler) < 0.5f) return cosine_sample_hemisphere(sampler_next2f(sc, sampler));
    const float alpha = roughness * roughness;
    const float3 wm = sample_wm(wo, alpha, sampler_next2f(sc, sampler));
    const float3 i = -wo;                                                 // reflect(-wo, wm)
    const float k = 2.0f * dot(wm, i);
    return f3(fmaf(-k, wm.x, i.x), fmaf(-k, wm.y, i.y), fmaf(-k, wm.z, i.z));
}
ZDR_DEV float ggx_sample_pdf(float3 wo, float3 wi, float roughness) {    // microfacet.py:51-58
    const float alpha = roughness * roughness;
    const float3 wm = normalize(wi + wo);
    const float diffuse_pdf = wi.z / kPi;
    const float glossy_pdf = pdf_wm(wo, wm, alpha) / (4.0f * fabsf(dot(wo, wm)));
    return 0.5f * diffuse_pdf + 0.5f * glossy_pdf;
}

// The terms ggx_brdf and ggx_sample_pdf (and the adjoint) share for one direction pair, each computed ONCE: the half vector
// (one normalize), D, the Fresnel term, both Smith terms and dot(wo, h).  Same operations on the same inputs as the separate
// functions above, so every value has the same bits — the out-of-line normalize is simply no longer called twice (the compiler
// cannot merge calls of a __noinline__ function) and D no longer evaluated twice.
struct GgxTerms { float nh, voh, d, f, gi, go; };
ZDR_DEV GgxTerms ggx_terms(float3 wo, float3 wi, float specular, float alpha) {
    GgxTerms t;
    const float3 h = normalize(wi + wo);
    t.nh = h.z;
    t.d = ggx_distribution(h, alpha);
    t.voh = dot(wo, h);
    t.f = fresnel_schlick(clampf(t.voh, 0.00001f, 1.0f), specular);
    t.gi = smith_geometry(wi, alpha);
    t.go = smith_geometry(wo, alpha);
    return t;
}
ZDR_DEV float4 ggx_eval_from_terms(const GgxTerms& t, float3 wo, float3 wi, float3 diffuse) {
    const float g = t.gi * t.go;                                                              // ggx_brdf, microfacet.py:23-30
    const float spec = (t.d * t.f * g) / (4.0f * fmaxf(0.00001f, wi.z) * fmaxf(0.00001f, wo.z));
    const float diffuse_pdf = wi.z / kPi;                                                     // ggx_sample_pdf, microfacet.py:51-58
    const float glossy_pdf = (t.go / fabsf(wo.z) * t.d * fabsf(t.voh)) / (4.0f * fabsf(t.voh));   // pdf_wm(wo, h) / (4 |wo . h|)
    return make_float4((spec + diffuse.x / kPi) * wi.z, (spec + diffuse.y / kPi) * wi.z, (spec + diffuse.z / kPi) * wi.z,
                       0.5f * diffuse_pdf + 0.5f * glossy_pdf);
}
ZDR_GGX_LINKAGE float4 ggx_eval(float wox, float woy, float woz, float wix, float wiy, float wiz, float dr, float dg, float db,
                                        float specular, float roughness) {
    const float3 wo = f3(wox, woy, woz), wi = f3(wix, wiy, wiz);
    const GgxTerms t = ggx_terms(wo, wi, specular, roughness * roughness);
    return ggx_eval_from_terms(t, wo, wi, f3(dr, dg, db));
}
// (ggx_brdf, ggx_sample_pdf) AND the basis of the adjoint of one direction pair in one call (adjoint kernels): the adjoint is
// linear in the cotangent g, ggx_brdf_grad(.., g) = (g.x dk, g.y dk, g.z dk, (g.x + g.y + g.z) common), so (dk, common) can be
// computed next to the BRDF value, before the caller knows g, from the terms the value already needed (half vector, D, F, both
// Smith terms — the separate ggx_brdf_grad computes all of them again).  Gradients are compared at 1e-3 (float atomics are
// unordered anyway); the Smith terms reused here differ from ggx_brdf_grad's own by the rounding of a2 * k.
ZDR_GGX_LINKAGE float4 ggx_eval_adj(float wox, float woy, float woz, float wix, float wiy, float wiz, float dr, float dg, float db,
                                            float specular, float roughness, float2* basis) {
    const float3 wo = f3(wox, woy, woz), wi = f3(wix, wiy, wiz);
    const float alpha = roughness * roughness, a2 = alpha * alpha;
    const GgxTerms t = ggx_terms(wo, wi, specular, alpha);
    const float nh = fmaxf(0.00001f, t.nh), nh2 = nh * nh;
    const float tt = nh2 * (a2 - 1.0f) + 1.0f;
    const float dD = (tt - 2.0f * a2 * nh2) / (kPi * (tt * tt * tt));
    const float nvi = fmaxf(0.00001f, wi.z), nvo = fmaxf(0.00001f, wo.z);
    const float ki = (1.0f - nvi * nvi) / (nvi * nvi), ko = (1.0f - nvo * nvo) / (nvo * nvo);
    const float si = sqrtf(1.0f + a2 * ki), so = sqrtf(1.0f + a2 * ko);
    const float dGi = -ki / (si * ((1.0f + si) * (1.0f + si)));
    const float dGo = -ko / (so * ((1.0f + so) * (1.0f + so)));
    const float dDGG = dD * t.gi * t.go + t.d * dGi * t.go + t.d * t.gi * dGo;
    const float da2_dr = 4.0f * roughness * roughness * roughness;
    *basis = make_float2(wi.z / kPi, wi.z * t.f / (4.0f * nvi * nvo) * dDGG * da2_dr);
    return ggx_eval_from_terms(t, wo, wi, f3(dr, dg, db));
}
ZDR_DEV float4 ggx_eval_adj(float3 wo, float3 wi, float3 diffuse, float specular, float roughness, float2& basis) {
    return ggx_eval_adj(wo.x, wo.y, wo.z, wi.x, wi.y, wi.z, diffuse.x, diffuse.y, diffuse.z, specular, roughness, &basis);
}
ZDR_DEV float4 ggx_grad_from_basis(float2 basis, float3 g) {
    return make_float4(g.x * basis.x, g.y * basis.x, g.z * basis.x, (g.x + g.y + g.z) * basis.y);
}
static __device__ __noinline__ float4 ggx_brdf_grad_nl(float wox, float woy, float woz, float wix, float wiy, float wiz, float specular,
                                                float roughness, float gx, float gy, float gz) {
    return ggx_brdf_grad(f3(wox, woy, woz), f3(wix, wiy, wiz), specular, roughness, f3(gx, gy, gz));
}
ZDR_DEV float4 ggx_eval(float3 wo, float3 wi, float3 diffuse, float specular, float roughness) {
    return ggx_eval(wo.x, wo.y, wo.z, wi.x, wi.y, wi.z, diffuse.x, diffuse.y, diffuse.z, specular, roughness);
}
ZDR_DEV float4 ggx_brdf_grad_nl(float3 wo, float3 wi, float specular, float roughness, float3 g) {
    return ggx_brdf_grad_nl(wo.x, wo.y, wo.z, wi.x, wi.y, wi.z, specular, roughness, g.x, g.y, g.z);
}

// luisa.offset_ray_origin (direct.py:64): Waechter & Binder, Ray Tracing Gems ch. 6
ZDR_DEV float3 offset_ray_origin(float3 p, float3 n) {
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    const int ox = f2i(int_scale * n.x), oy = f2i(int_scale * n.y), oz = f2i(int_scale * n.z);
    const float px = __int_as_float(__float_as_int(p.x) + (p.x < 0.0f ? -ox : ox));
    const float py = __int_as_float(__float_as_int(p.y) + (p.y < 0.0f ? -oy : oy));
    const float pz = __int_as_float(__float_as_int(p.z) + (p.z < 0.0f ? -oz : oz));
    return f3(fabsf(p.x) < origin ? fmaf(float_scale, n.x, p.x) : px,
              fabsf(p.y) < origin ? fmaf(float_scale, n.y, p.y) : py,
              fabsf(p.z) < origin ? fmaf(float_scale, n.z, p.z) : pz);
}

// ---------------------------------------------------------------------------------------------
// envmap.py (device side)
// ---------------------------------------------------------------------------------------------
constexpr int kEnvMapW = 512, kEnvMapH = 256;      // envmap.py:114

// heap.texture2d_sample(23332, uv): software bilinear, REPEAT in u, CLAMP in v (specified)
ZDR_DEV float3 env_texture(const SceneView& sc, float2 uv) {
    const float fx = uv.x * (float)sc.env_w - 0.5f, fy = uv.y * (float)sc.env_h - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float tx = fx - flx, ty = fy - fly;
    int x0 = f2i(flx), y0 = f2i(fly);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 %= sc.env_w; if (x0 < 0) x0 += sc.env_w;
    x1 %= sc.env_w; if (x1 < 0) x1 += sc.env_w;
    y0 = clampi(y0, 0, sc.env_h - 1); y1 = clampi(y1, 0, sc.env_h - 1);
    const float4 a = __ldg(sc.env_tex + ((size_t)y0 * sc.env_w + x0)), b = __ldg(sc.env_tex + ((size_t)y0 * sc.env_w + x1));
    const float4 c = __ldg(sc.env_tex + ((size_t)y1 * sc.env_w + x0)), d = __ldg(sc.env_tex + ((size_t)y1 * sc.env_w + x1));
    return f3(lerp(lerp(a.x, b.x, tx), lerp(c.x, d.x, tx), ty), lerp(lerp(a.y, b.y, tx), lerp(c.y, d.y, tx), ty),
              lerp(lerp(a.z, b.z, tx), lerp(c.z, d.z, tx), ty));
}
ZDR_DEV float3 uv_to_direction(float2 uv) {                               // envmap.py:206-213
    const float phi = kTwoPi * (1.0f - uv.x);
    const float theta = kPi * uv.y;
    float sp, cp, st, ct;
    sincos_contract(phi, &sp, &cp);
    sincos_contract(theta, &st, &ct);
    return normalize(f3(sp * st, ct, cp * st));
}
ZDR_DEV float2 direction_to_uv(float3 d) {                                // envmap.py:215-219
    const float theta = acosf(d.y);
    const float phi = atan2f(d.x, d.z);
    return make_float2(1.0f - phi / kTwoPi, theta / kPi);
}
ZDR_DEV float env_sampled_light_pdf(const SceneView& sc, float3 dir) {    // envmap.py:240-248
    const float2 uv = direction_to_uv(dir);
    const int index = clampi(f2i(uv.y * (float)kEnvMapH), 0, kEnvMapH - 1) * kEnvMapW +
                      clampi(f2i(uv.x * (float)kEnvMapW), 0, kEnvMapW - 1);
    const float pdf = __ldg(sc.env_pdf + index);
    float sn, cs;
    sincos_contract(kPi * uv.y, &sn, &cs);
    const float inv_s = sn > 0.0f ? 1.0f / sn : 0.0f;
    return pdf * inv_s / (2.0f * kPi * kPi);
}

struct LightSample { float3 wi; float dist, pdf; float3 eval; };         // light.py:11

ZDR_DEV void sample_alias(const SceneView& sc, int n, float u_in, int offset, int& index, float& uu) {  // envmap.py:70-98
    const float u = u_in * (float)n;
    const int i = clampi(f2i(u), 0, n - 1);
    const float ur = u - floorf(u);
    const float2 e = __ldg(sc.env_alias + i + offset);
    const float prob = e.x;
    const int alias = __float_as_int(e.y);
    index = ur < prob ? i : alias;
    uu = ur < prob ? ur / prob : (ur - prob) / (1.0f - prob);
}
ZDR_DEV LightSample sample_envmap(const SceneView& sc, float2 u) {        // envmap.py:223-238
    int sy, sx;
    float uy, ux;
    sample_alias(sc, kEnvMapH, u.y, 0, sy, uy);
    const int offset = kEnvMapH + sy * kEnvMapW;
    sample_alias(sc, kEnvMapW, u.x, offset, sx, ux);
    const float2 uv = make_float2(((float)sx + ux) / (float)kEnvMapW, ((float)sy + uy) / (float)kEnvMapH);
    const float pdf = __ldg(sc.env_pdf + sy * kEnvMapW + sx);
    LightSample t;
    t.wi = uv_to_direction(uv);
    t.dist = 1e30f;
    float sn, cs;
    sincos_contract(kPi * uv.y, &sn, &cs);
    const float inv_s = sn > 0.0f ? 1.0f / sn : 0.0f;
    t.pdf = pdf * inv_s / (2.0f * kPi * kPi);
    t.eval = env_texture(sc, uv);
    return t;
}

// ---------------------------------------------------------------------------------------------
// light.py
// ---------------------------------------------------------------------------------------------
ZDR_DEV float3 sample_uniform_triangle(float2 u) {                        // light.py:15-20
    float a, b;
    if (u.x < u.y) { a = 0.5f * u.x; b = -0.5f * u.x + u.y; }
    else { a = -0.5f * u.y + u.x; b = 0.5f * u.y; }
    return f3(a, b, 1.0f - a - b);
}
ZDR_DEV void inst_triangle(const SceneView& sc, int inst, int prim, float3& p0, float3& p1, float3& p2) {
    const int slot = __ldg(sc.slot_of_gid + sc.inst[inst].tri_base + prim);
    const float4* tp = sc.tris + 3 * (size_t)slot;
    const float4 a = ld_scene(tp), b = ld_scene(tp + 1), c = ld_scene(tp + 2);
    p0 = f3(a.x, a.y, a.z); p1 = f3(b.x, b.y, b.z); p2 = f3(c.x, c.y, c.z);
}
ZDR_DEV float3 inst_emission(const SceneView& sc, int inst) {
    const InstData& d = sc.inst[inst];
    return f3(d.emission[0], d.emission[1], d.emission[2]);
}
// light.py:22-81 (point_light_count = 0, light.py:7)
template <bool ENV = true>
ZDR_DEV LightSample sample_light(const SceneView& sc, float3 origin, Sampler& sampler) {
    const float u = sampler_next(sc, sampler);
    const int n = (ENV ? sc.env_count : 0) + 0 + sc.light_count;
    LightSample t;
    if (n == 0) {   // upstream indexes lights[-1]; defined: consume the mesh-light dims, null sample
        (void)sampler_next(sc, sampler);
        (void)sampler_next2f(sc, sampler);
        t.wi = f3(0.0f, 0.0f, 1.0f); t.dist = 0.0f; t.pdf = 1.0f; t.eval = f3(0.0f, 0.0f, 0.0f);
        return t;
    }
    int idx = clampi(f2i(u * (float)n), 0, n - 1);
    if (ENV) {
        if (idx < sc.env_count) return sample_envmap(sc, sampler_next2f(sc, sampler));
        idx -= sc.env_count;
    }
    const int inst = __ldg(sc.lights + idx);
    const int trig_count = sc.inst[inst].tri_count;
    const int prim = clampi(f2i(sampler_next(sc, sampler) * (float)trig_count), 0, trig_count - 1);
    float3 p0, p1, p2;
    inst_triangle(sc, inst, prim, p0, p1, p2);
    const float3 abc = sample_uniform_triangle(sampler_next2f(sc, sampler));
    const float3 p = lincomb(abc.x, p0, abc.y, p1, abc.z, p2);
    const float3 emission = inst_emission(sc, inst);
    const float3 dd = p - origin;
    const float3 wi_light = normalize(dd);
    const float3 c = cross(p1 - p0, p2 - p0);
    const float3 light_normal = normalize(c);
    const float cos_light = -dot(light_normal, wi_light);
    const float sqr_dist = dot(dd, dd);
    const float area = length(c) / 2.0f;
    t.pdf = sqr_dist / ((float)(n * trig_count) * area * cos_light);
    t.eval = cos_light > 1e-4f ? emission : f3(0.0f, 0.0f, 0.0f);
    t.wi = wi_light;
    t.dist = 0.9999f * sqrtf(sqr_dist);
    return t;
}
// light.py:83-111 — n excludes env_count (light.py:89), reproduced unless ZDR_FLAG_LIGHT_PDF_COUNTS_ENV
ZDR_DEV float sample_light_pdf(const SceneView& sc, float3 origin, const Hit& h, float3 p) {
    const int n = sc.pdf_env_count + sc.light_count;
    const int trig_count = sc.inst[h.inst].tri_count;
    const float4* tp = sc.tris + 3 * (size_t)h.slot;
    const float4 a = ld_scene(tp), b = ld_scene(tp + 1), c4 = ld_scene(tp + 2);
    const float3 p0 = f3(a.x, a.y, a.z), p1 = f3(b.x, b.y, b.z), p2 = f3(c4.x, c4.y, c4.z);
    const float3 dd = p - origin;
    const float3 wi_light = normalize(dd);
    const float3 c = cross(p1 - p0, p2 - p0);
    const float3 light_normal = normalize(c);
    const float cos_light = -dot(light_normal, wi_light);
    const float sqr_dist = dot(dd, dd);
    const float area = length(c) / 2.0f;
    return sqr_dist / ((float)(n * trig_count) * area * cos_light);
}

ZDR_DEV float balanced_heuristic(float a, float b) { return a / fmaxf(a + b, 1e-4f); }   // direct.py:16-18
ZDR_DEV bool backfacing(float3 d, const Interaction& it) {                                // collocated.py:20
    const float3 nd = -d;
    return dot(nd, it.ng) < 1e-4f || dot(nd, it.ns) < 1e-4f;
}
ZDR_DEV float luminance(float3 b) { return dot(f3(0.212671f, 0.715160f, 0.072169f), b); }  // prb.py:80

}  // namespace zdr
