// render_common.cuh — device functions shared by the render kernels of render.cu, wavefront.cu and persistent.cu: camera sample
// (integrator.py:17-25, camera.py:5-31), envmap emission on a miss, the packing of (depth, phase, sampler dimension), the
// pixel order of a shard.
#pragma once
#include "render.cuh"

namespace zdr {

ZDR_DEV Ray make_ray(float3 o, float3 d, float tmin, float tmax) { Ray r; r.o = o; r.d = d; r.tmin = tmin; r.tmax = tmax; return r; }

ZDR_DEV float3 env_emission(const SceneView& sc, float3 dir) {                     // direct.py:24
    return sc.env_count > 0 ? env_texture(sc, direction_to_uv(dir)) : f3(0.f, 0.f, 0.f);
}

// --------------------------------------------------------------------- camera.py, integrator.py
ZDR_DEV float tent_warp1(float u) {                                                // camera.py:20-31
    return u < 0.5f ? 1.0f * (sqrtf(2.0f * u) - 1.0f) : 1.0f * (1.0f - sqrtf(2.0f - 2.0f * u));
}
ZDR_DEV Ray camera_sample(const RenderArgs& a, int x, int y, Sampler& sampler) {    // integrator.py:17-25, camera.py:5-17
    float2 off = sampler_next2f(a.sc, sampler);
    if (a.use_tent) { off.x = tent_warp1(off.x) + 0.5f; off.y = tent_warp1(off.y) + 0.5f; }
    float px = 2.0f / (float)a.W * ((float)x + off.x) - 1.0f;
    float py = 2.0f / (float)a.H * ((float)y + off.y) - 1.0f;
    py *= (float)a.H / (float)a.W;
    const float sx = px * a.tan_half_fov, sy = py * a.tan_half_fov;
    const float3 d = f3(sx * a.cam_right.x - sy * a.cam_up.x + a.cam_fwd.x, sx * a.cam_right.y - sy * a.cam_up.y + a.cam_fwd.y,
                        sx * a.cam_right.z - sy * a.cam_up.z + a.cam_fwd.z);
    return make_ray(a.cam_o, normalize(d), 0.0f, 1e30f);
}

ZDR_DEV unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }      // one REDUX (was a 5-step shuffle loop: 3 % of a shade stage's instructions)
constexpr unsigned kWfPhaseBit = 1u << 8;

ZDR_DEV unsigned wf_pack(int depth, int phase, int dimension) { return (unsigned)depth | (phase ? kWfPhaseBit : 0u) | ((unsigned)dimension << 12); }
ZDR_DEV void wf_pixel(const RenderArgs& a, int pix, int& x, int& y) {
    const int lr = fastdiv(pix, a.wf.div_W);
    x = pix - lr * a.W;
    y = a.row0 + ((lr >> 3) * a.shard_count + a.shard_index) * 8 + (lr & 7);
}

}  // namespace zdr
