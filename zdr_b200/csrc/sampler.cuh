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
    return ((s % l) + l) % l;
}

struct Sampler {     // pmj02bn.py:78-85
    int px, py, sample_index, dimension, seed, spp, w;
};

// pmj02bn.py:87-102
ZDR_HD Sampler make_sampler(int px, int py, int seed, int spp, int sample_index) {
    int w = spp - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    Sampler s;
    s.px = px; s.py = py; s.sample_index = sample_index; s.dimension = 0; s.seed = seed; s.spp = spp; s.w = w;
    return s;
}

// pmj02bn.py:19-24 — literal index formula tex*(res+cx)*res+cy; indices past the table wrap
// modulo its length (upstream: undefined behaviour from dimension 25 on).
ZDR_DEV float blue_noise(const SceneView& sc, int tex_id, int px, int py) {
    int tex_index = tex_id % sc.bn_ntex;
    int cx = px % sc.bn_res, cy = py % sc.bn_res;
    long long i = (long long)tex_index * (sc.bn_res + cx) * sc.bn_res + cy;
    long long len = (long long)sc.bn_ntex * sc.bn_res * sc.bn_res;
    i %= len;
    if (i < 0) i += len;
    return __ldg(sc.bn + i);
}
// pmj02bn.py:26-30
ZDR_DEV float2 pmj02bn_sample(const SceneView& sc, int set_id, int sample_id) {
    int set_index = set_id % sc.pmj_sets;
    long long i = (long long)set_index * sc.pmj_n + sample_id;
    long long len = (long long)sc.pmj_sets * sc.pmj_n;
    i %= len;
    if (i < 0) i += len;
    return __ldg(sc.pmj + i);
}
// pmj02bn.py:104-112
ZDR_DEV float sampler_next(const SceneView& sc, Sampler& s) {
    uint32_t h = xxhash32_int4(s.px, s.py, s.dimension, s.seed);
    int index = permutation_element(s.sample_index, s.spp, s.w, h);
    float delta = blue_noise(sc, s.dimension, s.px ^ s.seed, s.py ^ s.seed);
    float u = ((float)index + delta) / (float)s.spp;
    s.dimension += 1;
    return clampf(u, 0.0f, kOneMinusEps);
}
// pmj02bn.py:114-126
ZDR_DEV float2 sampler_next2f(const SceneView& sc, Sampler& s) {
    int index = s.sample_index;
    int pmj_instance = s.dimension / 2;
    if (pmj_instance >= sc.pmj_sets) {
        uint32_t h = xxhash32_int4(s.px, s.py, s.dimension, s.seed);
        index = permutation_element(s.sample_index, s.spp, s.w, h);
    }
    float2 u = pmj02bn_sample(sc, pmj_instance, index);
    u.x = u.x + blue_noise(sc, s.dimension, s.px ^ s.seed, s.py ^ s.seed);
    u.y = u.y + blue_noise(sc, s.dimension + 1, s.px ^ s.seed, s.py ^ s.seed);
    s.dimension += 2;
    u.x = u.x - floorf(u.x);
    u.y = u.y - floorf(u.y);
    return u;
}

}  // namespace zdr
