// sampler.cuh — pmj02bn sampler (pmj02bn.py:19-126), bit-exact integer hashing + table gathers.
#pragma once
#include "scene.cuh"

namespace zdr {

constexpr float kOneMinusEps = 0x1.fffffep-1f;   // pmj02bn.py:76

// pmj02bn.py:59-74 — xxHash32-style mix of (pixel.x, pixel.y, dimension, seed) as uint4
ZDR_HD uint32_t xxhash32_int4(int x, int y, int z, int w) {
    const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t h = (uint32_t)w + P5 + (uint32_t)x * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h += (uint32_t)y * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h += (uint32_t)z * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h = P2 * (h ^ (h >> 15));
    h = P3 * (h ^ (h >> 13));
    return h ^ (h >> 16);
}

// pmj02bn.py:32-57 — Kensler's cycle-walking permutation; the reference's last line
// ((i + p) % l + l) % l is evaluated with signed (floor-mod) semantics.
ZDR_HD int permutation_element(int i_in, int l, int w_in, uint32_t p) {
    uint32_t i = (uint32_t)i_in, w = (uint32_t)w_in;
    do {
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
    } while (i >= (uint32_t)l);
    int32_t s = (int32_t)(i + p);
    if ((l & (l - 1)) == 0) return s & (l - 1);      // floor-mod by a power of two (every BASELINE config)
    return ((s % l) + l) % l;
}

struct Sampler {     // pmj02bn.py:78-85, plus per-(pixel,seed) values hoisted out of the draws
    int px, py, sample_index, dimension, seed, spp, w;
    uint32_t h_xy;      // xxhash32_int4 state after mixing seed, pixel.x, pixel.y (pmj02bn.py:66-69)
    int bn_k, bn_cy;    // blue-noise index = tex * bn_k + bn_cy with bn_k = (res + cx) * res (pmj02bn.py:20-23)
};

// pmj02bn.py:87-102
ZDR_DEV Sampler make_sampler(const SceneView& sc, int px, int py, int seed, int spp, int sample_index) {
    int w = spp - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    Sampler s;
    s.px = px; s.py = py; s.sample_index = sample_index; s.dimension = 0; s.seed = seed; s.spp = spp; s.w = w;
    const uint32_t P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t h = (uint32_t)seed + P5 + (uint32_t)px * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h += (uint32_t)py * P3;
    s.h_xy = P4 * ((h << 17) | (h >> 15));
    const int qx = px ^ seed, qy = py ^ seed;              // pmj02bn.py:108: pixel ^ seed
    int cx, cy;                                            // truncating %, as the reference's int2 % lowers
    if ((sc.bn_res & (sc.bn_res - 1)) == 0) {              // power-of-two table (pbrt's is 128): sign * (|q| & (res - 1))
        const unsigned m = (unsigned)sc.bn_res - 1u;
        const unsigned ax = qx < 0 ? 0u - (unsigned)qx : (unsigned)qx, ay = qy < 0 ? 0u - (unsigned)qy : (unsigned)qy;
        cx = (int)(ax & m); cy = (int)(ay & m);
        if (qx < 0) cx = -cx;
        if (qy < 0) cy = -cy;
    } else {
        cx = qx % sc.bn_res; cy = qy % sc.bn_res;
    }
    s.bn_k = (sc.bn_res + cx) * sc.bn_res;
    s.bn_cy = cy;
    return s;
}
// pmj02bn.py:19-24 — literal index formula tex*(res+cx)*res+cy; indices past the table wrap
// modulo its length (upstream: undefined behaviour from dimension 25 on).  |index| is below
// 2*len + res, so the modulo is at most two conditional subtractions; zdr_set_sampler_tables keeps len below
// 2^29, so 32-bit arithmetic is exact.  small_mod: x mod n for an x that is almost always below 2n.
ZDR_DEV int small_mod(int x, int n) {
    unsigned t = (unsigned)x;
    if (t >= (unsigned)n) { t -= (unsigned)n; if (t >= (unsigned)n) t %= (unsigned)n; }
    return (int)t;
}
ZDR_DEV float blue_noise(const float* __restrict__ bn, int ntex, int res, int bn_k, int bn_cy, int tex_id) {
    const int tex_index = small_mod(tex_id, ntex);   // tex_id >= 0
    int i = tex_index * bn_k + bn_cy;
    const int len = ntex * res * res;
    if (i < 0) i += len;
    if (i >= len) i -= len;
    if (i >= len) i -= len;
    return __ldg(bn + i);
}
// pmj02bn.py:26-30
ZDR_DEV float2 pmj02bn_sample(const float2* __restrict__ pmj, int sets, int n, int set_id, int sample_id) {
    const int set_index = small_mod(set_id, sets);
    if (sample_id < n) return __ldg(pmj + (set_index * n + sample_id));
    long long i = (long long)set_index * n + sample_id;     // only for spp > table length
    const long long len = (long long)sets * n;
    if (i >= len) i %= len;
    return __ldg(pmj + i);
}
ZDR_DEV uint32_t hash_dim(uint32_t h_xy, int dimension) {
    const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u;
    uint32_t h = h_xy + (uint32_t)dimension * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h = P2 * (h ^ (h >> 15));
    h = P3 * (h ^ (h >> 13));
    return h ^ (h >> 16);
}
// The draws are out-of-line, by-value functions: seven call sites per path vertex would
// otherwise each inline ~100-150 instructions and push the render kernels out of the 32 KB
// instruction-cache level.
// pmj02bn.py:104-112
static __device__ __noinline__ float sampler_next_v(const float* __restrict__ bn, int ntex, int res, uint32_t h_xy, int bn_k, int bn_cy,
                                             int sample_index, int spp, int w, int dimension) {
    const uint32_t h = hash_dim(h_xy, dimension);
    const int index = permutation_element(sample_index, spp, w, h);
    const float delta = blue_noise(bn, ntex, res, bn_k, bn_cy, dimension);
    const float u = __fdiv_rn(__fadd_rn((float)index, delta), (float)spp);      // explicit IEEE operations: the sampler is bit-exact in every build
    return clampf(u, 0.0f, kOneMinusEps);
}
// pmj02bn.py:114-126
static __device__ __noinline__ float2 sampler_next2f_v(const float2* __restrict__ pmj, int sets, int n, const float* __restrict__ bn, int ntex,
                                                int res, uint32_t h_xy, int bn_k, int bn_cy, int sample_index, int spp, int w,
                                                int dimension) {
    int index = sample_index;
    const int pmj_instance = dimension / 2;
    if (pmj_instance >= sets) index = permutation_element(sample_index, spp, w, hash_dim(h_xy, dimension));
    float2 u = pmj02bn_sample(pmj, sets, n, pmj_instance, index);
    u.x = __fadd_rn(u.x, blue_noise(bn, ntex, res, bn_k, bn_cy, dimension));
    u.y = __fadd_rn(u.y, blue_noise(bn, ntex, res, bn_k, bn_cy, dimension + 1));
    u.x = __fsub_rn(u.x, floorf(u.x));
    u.y = __fsub_rn(u.y, floorf(u.y));
    return u;
}
ZDR_DEV float sampler_next(const SceneView& sc, Sampler& s) {
    const float u = sampler_next_v(sc.bn, sc.bn_ntex, sc.bn_res, s.h_xy, s.bn_k, s.bn_cy, s.sample_index, s.spp, s.w, s.dimension);
    s.dimension += 1;
    return u;
}
ZDR_DEV float2 sampler_next2f(const SceneView& sc, Sampler& s) {
    const float2 u = sampler_next2f_v(sc.pmj, sc.pmj_sets, sc.pmj_n, sc.bn, sc.bn_ntex, sc.bn_res, s.h_xy, s.bn_k, s.bn_cy,
                                      s.sample_index, s.spp, s.w, s.dimension);
    s.dimension += 2;
    return u;
}

}  // namespace zdr
