// render.cu — per-pixel render kernels (forward + adjoint) for the collocated, direct and path
// integrators.  Replaces the LuisaCompute kernels derived at integrator.py:8-53 from
// collocated.py / direct.py / prb.py and dispatched at render.py:168-171,193-196.  (The HBM wavefront schedule of the path
// integrator is wavefront.cu, the persistent one persistent.cu.)
#include <algorithm>
#include <type_traits>
#include "render_ctx.cuh"

namespace zdr {

#ifndef ZDR_FAST_MATH
std::atomic<unsigned long long> g_kernel_launches{0};
#endif

namespace {

// ------------------------------------------------------------------------------ collocated.py
template <class CTX>
ZDR_DEV float3 collocated_estimator(CTX& c, const Ray& ray) {               // collocated.py:10-31
    const Hit hit = c.closest(ray);
    if (hit.inst < 0) return f3(0.f, 0.f, 0.f);
    const Interaction it = surface_interact(c.a.sc, hit);
    if (backfacing(ray.d, it)) return f3(0.f, 0.f, 0.f);
    c.tl.shaded++;
    const float4 mat = read_bsdf(c.a.tex, it.uv, inst_texture_set(c.a.sc, hit.inst));
    const Onb onb = make_onb(it.ns);
    const float3 wo_local = to_local(onb, -ray.d);
    const float3 beta = ggx_brdf(wo_local, wo_local, f3(mat.x, mat.y, mat.z), 0.04f, mat.w);
    const float inv = 1.0f / hit.t;
    const float li = 1.0f * (inv * inv);
    return beta * li;
}
template <class CTX>
ZDR_DEV void collocated_estimator_backward(CTX& c, const Ray& ray, float3 le_grad) {   // collocated.py:34-57
    const Hit hit = c.closest(ray);
    if (hit.inst < 0) return;
    const Interaction it = surface_interact(c.a.sc, hit);
    if (backfacing(ray.d, it)) return;
    c.tl.shaded++;
    const int mset = inst_texture_set(c.a.sc, hit.inst);
    const float4 mat = read_bsdf(c.a.tex, it.uv, mset);
    const Onb onb = make_onb(it.ns);
    const float3 wo_local = to_local(onb, -ray.d);
    const float inv = 1.0f / hit.t;
    const float li = 1.0f * (inv * inv);
    c.scatter(it.uv, ggx_brdf_grad(wo_local, wo_local, 0.04f, mat.w, le_grad * li), mset);
}

// ---------------------------------------------------------------------------------- direct.py
// Shared walk of direct_estimator (direct.py:20-85) and direct_estimator_backward (:88-167):
// BWD=false returns radiance; BWD=true accumulates the BRDF adjoints and scatters once.
template <bool BWD, class CTX>
ZDR_DEV float3 direct_walk(CTX& c, const Ray& ray, Sampler& sampler, float3 le_grad) {
    const SceneView& sc = c.a.sc;
    const float3 zero = f3(0.f, 0.f, 0.f);
    const Hit hit = c.closest(ray);
    if (hit.inst < 0) return BWD ? zero : env_emission(sc, ray.d);
    const Interaction it = surface_interact(sc, hit);
    if (backfacing(ray.d, it)) return zero;
    float3 emission = inst_emission(sc, hit.inst);
    const int mset = inst_material_set(sc, hit.inst);
    if (mset < 0) return BWD ? zero : emission;                                    // direct.py:30-32 (`hit.inst > 0` upstream)
    c.tl.shaded++;
    const float4 mat = read_bsdf(c.a.tex, it.uv, mset);
    const float3 diffuse = f3(mat.x, mat.y, mat.z);
    const float roughness = mat.w;
    const float specular = 0.04f;
    float3 radiance = zero;
    float4 mat_grad = make_float4(0.f, 0.f, 0.f, 0.f);
    const LightSample light = sample_light(sc, it.p, sampler);
    const bool occluded = c.occluded(make_ray(it.p, light.wi, 1e-4f, light.dist));
    const Onb onb = make_onb(it.ns);
    const float3 wo_local = to_local(onb, -ray.d);
    const float3 wi_light_local = to_local(onb, light.wi);
    if (!occluded && wi_light_local.z > 0.0f) {
        float mis_weight = 1.0f;
        if (c.a.use_mis) mis_weight = balanced_heuristic(light.pdf, ggx_sample_pdf(wo_local, wi_light_local, roughness));
        const float dn = fmaxf(light.pdf, 1e-4f);
        if (BWD) {
            const float3 cot = le_grad * ((light.eval * mis_weight) / dn);
            mat_grad = mat_grad + ggx_brdf_grad(wo_local, wi_light_local, specular, roughness, cot);
        } else {
            const float3 bsdf = ggx_brdf(wo_local, wi_light_local, diffuse, specular, roughness);
            radiance = radiance + ((bsdf * mis_weight) * light.eval) / dn;
        }
    }
    if (c.a.use_mis) {
        const float3 wi_local = ggx_sample(sc, wo_local, roughness, sampler);
        const float3 wi = to_world(onb, wi_local);
        if (dot(wi, it.ng) < 1e-4f || wi_local.z < 1e-4f) {
            if (BWD) c.scatter(it.uv, mat_grad, mset);                             // direct.py:138
            return radiance;
        }
        const Ray r2 = make_ray(offset_ray_origin(it.p, it.ng), wi, 0.0f, 1e30f);
        const Hit hit2 = c.closest(r2);
        float pdf_light;
        if (hit2.inst < 0) {
            emission = env_emission(sc, r2.d);
            pdf_light = sc.env_count > 0 ? env_sampled_light_pdf(sc, r2.d) : 0.0f;
        } else {
            const Interaction it2 = surface_interact(sc, hit2);
            if (backfacing(r2.d, it2)) {
                // direct.py:152 scatters at `it.uv` after `it` was rebound to the second hit
                if (BWD) c.scatter(it2.uv, mat_grad, mset);
                return radiance;
            }
            emission = inst_emission(sc, hit2.inst);
            pdf_light = sample_light_pdf(sc, it.p, hit2, it2.p);
        }
        if (any_pos(emission)) {
            const float pdf_bsdf = ggx_sample_pdf(wo_local, wi_local, roughness);
            const float mis_weight = balanced_heuristic(pdf_bsdf, pdf_light);
            if (BWD) {
                const float3 cot = (le_grad * (emission * mis_weight)) / pdf_bsdf;
                mat_grad = mat_grad + ggx_brdf_grad(wo_local, wi_local, specular, roughness, cot);
            } else {
                const float3 beta = ggx_brdf(wo_local, wi_local, diffuse, specular, roughness) / pdf_bsdf;
                return radiance + (beta * mis_weight) * emission;
            }
        }
    }
    if (BWD) c.scatter(it.uv, mat_grad, mset);                                     // direct.py:167
    return radiance;
}

// Resident blocks per SM the per-pixel kernels (collocated, direct) are compiled for.  Round 1 left them uncapped (128
// registers, 16 warps per SM); at 8 blocks (64 registers, 32 warps) C2 4 861 -> 5 182, C4 7 268 -> 8 734, C4d 7 201 -> 7 638 Mrays/s
// (profiles/r02n_pixel_kernel_occupancy.txt; 6 and 10 blocks are in between).
#ifndef ZDR_PIXEL_MINB
#define ZDR_PIXEL_MINB 8
#endif
// block = 128 threads = 4 warps; each warp owns an 8x4 pixel tile, the block a 16x8 tile
constexpr int kBlockThreads = 128;
ZDR_DEV void pixel_of_thread(const RenderArgs& a, int& x, int& y) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    x = blockIdx.x * 16 + (w & 1) * 8 + (l & 7);
    y = a.row0 + (blockIdx.y * a.shard_count + a.shard_index) * 8 + (w >> 1) * 4 + (l >> 3);
}

template <int INTEGRATOR, bool STATS>
__global__ void __launch_bounds__(kBlockThreads, ZDR_PIXEL_MINB) k_forward(const __grid_constant__ RenderArgs a) {
    int x, y;
    pixel_of_thread(a, x, y);
    Tally tl = {};
    Ctx<STATS> c{a, tl};
    if (x < a.W && y < a.row1) {
        float3 s = f3(0.f, 0.f, 0.f);
        for (int it = a.it0; it < a.it1; ++it) {                                   // integrator.py:15
            Sampler sampler = make_sampler(a.sc, x, y, a.seed, a.spp, it);
            const Ray ray = camera_sample(a, x, y, sampler);
            tl.samples++;
            float3 radiance;
            if (INTEGRATOR == ZDR_COLLOCATED) radiance = collocated_estimator(c, ray);
            else radiance = direct_walk<false>(c, ray, sampler, f3(0.f, 0.f, 0.f));
            if (a.samples_out) {
                float* o = a.samples_out + (((size_t)y * a.W + x) * a.spp + it) * 3;
                o[0] = radiance.x; o[1] = radiance.y; o[2] = radiance.z;
            }
            if (!any_nan(radiance))                                                // integrator.py:27-28
                s = s + f3(clampf(radiance.x, 0.0f, 100000.0f), clampf(radiance.y, 0.0f, 100000.0f),
                           clampf(radiance.z, 0.0f, 100000.0f));
        }
        if (a.image) a.image[(size_t)x + (size_t)y * a.W] = make_float4(s.x / (float)a.spp, s.y / (float)a.spp, s.z / (float)a.spp, 1.0f);
    }
    flush_tally(a, tl);
}

template <int INTEGRATOR, bool STATS, bool AGG>
__global__ void __launch_bounds__(kBlockThreads, ZDR_PIXEL_MINB) k_backward(const __grid_constant__ RenderArgs a) {
    int x, y;
    pixel_of_thread(a, x, y);
    Tally tl = {};
    Ctx<STATS, AGG> c{a, tl};
    ZDR_GRAD_CACHE_BEGIN(c)
    if (x < a.W && y < a.row1) {
        const float4 g = __ldg(a.d_image + ((size_t)x + (size_t)y * a.W));          // integrator.py:38-40
        float3 le_grad = f3(g.x / (float)a.spp, g.y / (float)a.spp, g.z / (float)a.spp);
        if (any_nan(le_grad)) le_grad = f3(0.f, 0.f, 0.f);
        for (int it = a.it0; it < a.it1; ++it) {
            Sampler sampler = make_sampler(a.sc, x, y, a.seed, a.spp, it);
            const Ray ray = camera_sample(a, x, y, sampler);
            tl.samples++;
            if (INTEGRATOR == ZDR_COLLOCATED) collocated_estimator_backward(c, ray, le_grad);
            else direct_walk<true>(c, ray, sampler, le_grad);
        }
    }
    ZDR_GRAD_CACHE_END(c, a, tl)
    flush_tally(a, tl);
}

// ------------------------------------------------------------------------------------- prb.py
// Path integrator (prb.py:18-88) and its Path-Replay-Backprop adjoint (prb.py:91-187) as ONE
// bounce state machine.  The reference nests `for it in range(spp)` around `while depth <
// max_depth` in one thread per pixel, so a warp idles until its longest path ends and — with the
// replay and the adjoint walk inlined as separate copies — the kernel's code does not fit the
// instruction cache.  Here one loop iteration is one bounce of whatever sample the lane is on:
// a lane whose path ended regenerates its next sample (or, in the adjoint kernel, switches from
// the replay phase to the adjoint phase of the same sample) at the top of the loop, so the whole
// warp executes the single copy of closest-hit / shading / shadow-ray / BSDF-sampling code
// together.  Every pixel still runs its samples in order in one thread: the image keeps the
// reference's accumulation order (integrator.py:15-28) and is bit-identical to the oracle's.
//   phase 0: radiance walk (L accumulates);  phase 1 (BWD only): adjoint walk (L = Le remaining).
template <bool BWD, bool STATS, bool AGG, bool ENV>
__global__ void __launch_bounds__(kBlockThreads, 4) k_path(const __grid_constant__ RenderArgs a) {
    int x, y;
    pixel_of_thread(a, x, y);
    const SceneView& sc = a.sc;
    Tally tl = {};
    Ctx<STATS, AGG> c{a, tl};
    if constexpr (BWD) { ZDR_GRAD_CACHE_BEGIN(c) }
    const bool corrected = a.prb_mode == ZDR_PRB_CORRECTED;
    const float specular = 0.04f;
    const float3 zero = f3(0.f, 0.f, 0.f);
    if (x < a.W && y < a.row1) {
        float3 acc = zero, le_grad = zero;
        if (BWD) {                                                                 // integrator.py:38-40
            const float4 g = __ldg(a.d_image + ((size_t)x + (size_t)y * a.W));
            le_grad = f3(g.x / (float)a.spp, g.y / (float)a.spp, g.z / (float)a.spp);
            if (any_nan(le_grad)) le_grad = zero;
        }
        int it = a.it0, depth = 0, phase = 0;
        bool alive = false;
        Ray ray = make_ray(zero, zero, 0.f, 0.f);
        float3 cam_d = zero, beta = zero, L = zero;
        float pdf_bsdf = 0.f;
        Sampler sampler = make_sampler(sc, x, y, a.seed, a.spp, 0);
        for (;;) {
            if (!alive) {
                if (it >= a.it1) break;
                sampler.sample_index = it;                                         // make_pmj02bn_sampler(coord, seed, spp, it)
                sampler.dimension = 0;
                ray = camera_sample(a, x, y, sampler);
                cam_d = ray.d;
                beta = f3(1.f, 1.f, 1.f); L = zero; pdf_bsdf = 1e30f; depth = 0; phase = 0;   // prb.py:20-22
                alive = true;
                tl.samples++;
            }
            bool end = false;
            if (depth >= a.max_depth) {                                            // prb.py:23
                end = true;
            } else {
                const Hit hit = c.closest(ray);                                    // prb.py:25
                if (hit.inst < 0) {
                    if (ENV && phase == 0 && sc.env_count > 0) {                          // prb.py:26-31
                        const float3 tex = env_texture(sc, direction_to_uv(ray.d));
                        const float3 emission = corrected ? tex : beta * tex;     // prb.py:28 applies beta twice
                        const float mis_weight = balanced_heuristic(pdf_bsdf, env_sampled_light_pdf(sc, ray.d));
                        L = L + (beta * mis_weight) * emission;
                    }
                    end = true;
                } else {
                    const Interaction itx = surface_interact(sc, hit);
                    const float3 emission = inst_emission(sc, hit.inst);
                    if (backfacing(ray.d, itx)) {                                  // prb.py:35-36
                        end = true;
                    } else if (any_pos(emission)) {                                // prb.py:39-44 / :117-121
                        const float mis_weight = balanced_heuristic(pdf_bsdf, sample_light_pdf(sc, ray.o, hit, itx.p));
                        const float3 contrib = (beta * mis_weight) * emission;
                        L = (phase == 0) ? L + contrib : L - contrib;
                        end = true;
                    } else if ((phase == 0 || corrected) && inst_material_set(sc, hit.inst) < 0) {   // prb.py:45-46 (forward only upstream)
                        end = true;
                    } else {
                        tl.shaded++;
                        float4 mat_grad = make_float4(0.f, 0.f, 0.f, 0.f);
                        const int mset = inst_texture_set(sc, hit.inst);
                        const float4 mat = read_bsdf(a.tex, itx.uv, mset);
                        const float3 diffuse = f3(mat.x, mat.y, mat.z);
                        const float roughness = mat.w;
                        const Onb onb = make_onb(itx.ns);
                        const float3 wo_local = to_local(onb, -ray.d);
                        const LightSample light = sample_light<ENV>(sc, itx.p, sampler);                     // prb.py:57
                        const bool occluded = c.occluded(make_ray(itx.p, light.wi, 1e-4f, light.dist));  // prb.py:58-59
                        const float3 wi_light_local = to_local(onb, light.wi);
                        if (!occluded && wi_light_local.z >= 1e-4f) {                                   // prb.py:61-66
                            const float4 ev = ggx_eval(wo_local, wi_light_local, diffuse, specular, roughness);
                            const float3 bsdf = f3(ev.x, ev.y, ev.z);
                            pdf_bsdf = ev.w;
                            const float mis_weight = balanced_heuristic(light.pdf, pdf_bsdf);
                            const float dn = fmaxf(light.pdf, 1e-4f);
                            const float3 contrib = (((beta * bsdf) * mis_weight) * light.eval) / dn;
                            if (!BWD || phase == 0) {
                                L = L + contrib;
                            } else {
                                const float3 cot = (((beta * mis_weight) * light.eval) / dn) * le_grad;   // prb.py:145
                                mat_grad = mat_grad + ggx_brdf_grad_nl(wo_local, wi_light_local, specular, roughness, cot);
                                L = L - contrib;                                                        // prb.py:147
                            }
                        }
                        const float3 wi_local = ggx_sample(sc, wo_local, roughness, sampler);           // prb.py:69
                        const float4 ev = ggx_eval(wo_local, wi_local, diffuse, specular, roughness);
                        pdf_bsdf = ev.w;
                        const float3 wi = to_world(onb, wi_local);
                        if (dot(wi, itx.ng) < 1e-4f || wi_local.z < 1e-4f) {                            // prb.py:73-74
                            end = true;
                        } else {
                            ray = make_ray(offset_ray_origin(itx.p, itx.ng), wi, 0.0f, 1e30f);          // prb.py:75
                            const float3 bsdf = f3(ev.x, ev.y, ev.z);
                            if (BWD && phase == 1) {
                                float3 cot;
                                if (corrected)   // d(Le_remaining)/d(bsdf) = Le_remaining / bsdf per channel
                                    cot = f3(bsdf.x != 0.0f ? L.x / bsdf.x * le_grad.x : 0.0f, bsdf.y != 0.0f ? L.y / bsdf.y * le_grad.y : 0.0f,
                                             bsdf.z != 0.0f ? L.z / bsdf.z * le_grad.z : 0.0f);
                                else
                                    cot = ((beta / pdf_bsdf) * L) * le_grad;                            // prb.py:162
                                mat_grad = mat_grad + ggx_brdf_grad_nl(wo_local, wi_local, specular, roughness, cot);
                            }
                            beta = beta * (bsdf / pdf_bsdf);                                            // prb.py:76
                            if (depth >= a.rr_depth) {                                                  // prb.py:78-87
                                const float l = luminance(beta);
                                if (l == 0.0f) {
                                    end = true;
                                } else {
                                    const float q = fmaxf(l, 0.05f);
                                    const float r = sampler_next(sc, sampler);
                                    if (r >= q) end = true;
                                    else beta = beta / q;
                                }
                            }
                        }
                        if (BWD && phase == 1) {                                                        // prb.py:178-187
                            if ((mat_grad.x != 0.0f || mat_grad.y != 0.0f || mat_grad.z != 0.0f || mat_grad.w != 0.0f) && grad_ok(mat_grad))
                                c.scatter(itx.uv, mat_grad, mset);
                        }
                        ++depth;
                    }
                }
            }
            if (end) {
                if (BWD && phase == 0 && !any_nan(L)) {
                    // prb.py:97-101: Le is known; walk the same path again with the original sampler state
                    phase = 1;
                    ray = make_ray(a.cam_o, cam_d, 0.0f, 1e30f);
                    sampler.dimension = 2;
                    beta = f3(1.f, 1.f, 1.f); pdf_bsdf = 1e30f; depth = 0;
                } else {
                    if (!BWD) {
                        if (a.samples_out) {
                            float* o = a.samples_out + (((size_t)y * a.W + x) * a.spp + it) * 3;
                            o[0] = L.x; o[1] = L.y; o[2] = L.z;
                        }
                        if (!any_nan(L))                                                                // integrator.py:27-28
                            acc = acc + f3(clampf(L.x, 0.0f, 100000.0f), clampf(L.y, 0.0f, 100000.0f), clampf(L.z, 0.0f, 100000.0f));
                    }
                    alive = false;
                    ++it;
                }
            }
        }
        if (!BWD && a.image)
            a.image[(size_t)x + (size_t)y * a.W] = make_float4(acc.x / (float)a.spp, acc.y / (float)a.spp, acc.z / (float)a.spp, 1.0f);
    }
    if constexpr (BWD) { ZDR_GRAD_CACHE_END(c, a, tl) }
    flush_tally(a, tl);
}

// ---------------------------------------------------------------------------------- uvgrad.py
// Scene.render_duvdxy (render.py:243-257, uvgrad.py:19-98): screen-to-texture Jacobian of the first
// hit by ray differentials.  Deviations, stated in DESIGN.md: the pixel offsets come from the render
// kernels' pmj02bn sampler (integrator.py:17; the reference uses LuisaCompute's built-in LCG here),
// triangle corners are world-space, and the 3x3 products of uvgrad.py:43-46 are evaluated as
// iaTa * (aT * dp).
ZDR_DEV Ray camera_ray_at(const RenderArgs& a, int x, int y, float2 off, float ox, float oy) {
    float px = 2.0f / (float)a.W * ((float)x + off.x + ox) - 1.0f;
    float py = 2.0f / (float)a.H * ((float)y + off.y + oy) - 1.0f;
    py *= (float)a.H / (float)a.W;
    const float sx = px * a.tan_half_fov, sy = py * a.tan_half_fov;
    const float3 d = f3(sx * a.cam_right.x - sy * a.cam_up.x + a.cam_fwd.x, sx * a.cam_right.y - sy * a.cam_up.y + a.cam_fwd.y,
                        sx * a.cam_right.z - sy * a.cam_up.z + a.cam_fwd.z);
    return make_ray(a.cam_o, normalize(d), 0.0f, 1e30f);
}
__global__ void __launch_bounds__(kBlockThreads) k_uvgrad(const __grid_constant__ RenderArgs a) {
    int x, y;
    pixel_of_thread(a, x, y);
    Tally tl = {};
    Ctx<false> c{a, tl};
    const SceneView& sc = a.sc;
    if (x < a.W && y < a.row1) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = a.it0; it < a.it1; ++it) {
            Sampler sampler = make_sampler(sc, x, y, a.seed, a.spp, it);
            float2 off = sampler_next2f(sc, sampler);
            if (a.use_tent) { off.x = tent_warp1(off.x) + 0.5f; off.y = tent_warp1(off.y) + 0.5f; }
            const Ray ray = camera_ray_at(a, x, y, off, 0.0f, 0.0f);
            const Ray rdx = camera_ray_at(a, x, y, off, 1.0f, 0.0f), rdy = camera_ray_at(a, x, y, off, 0.0f, 1.0f);
            tl.samples++;
            const Hit hit = c.closest(ray);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hit.inst >= 0) {
                const float4* tp = sc.tris + 3 * (size_t)hit.slot;
                const float4* sp = sc.shade + 4 * (size_t)hit.slot;
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2), s3 = __ldg(sp + 3);
                const float3 p0 = f3(t0.x, t0.y, t0.z), p1 = f3(t1.x, t1.y, t1.z), p2 = f3(t2.x, t2.y, t2.z);
                const float w = 1.0f - hit.u - hit.v;
                const float3 p = lincomb(w, p0, hit.u, p1, hit.v, p2);
                const float3 e1 = p1 - p0, e2 = p2 - p0;                                   // compute_dpduv, uvgrad.py:6-16
                const float ua = s1.w - s0.w, ub = s3.y - s3.x, uc = s2.w - s0.w, ud = s3.z - s3.x;
                const float det = ua * ud - uc * ub;
                const float i00 = ud / det, i10 = -ub / det, i01 = -uc / det, i11 = ua / det;
                const float3 dpdu = f3(fmaf(e2.x, i10, e1.x * i00), fmaf(e2.y, i10, e1.y * i00), fmaf(e2.z, i10, e1.z * i00));
                const float3 dpdv = -f3(fmaf(e2.x, i11, e1.x * i01), fmaf(e2.y, i11, e1.y * i01), fmaf(e2.z, i11, e1.z * i01));
                const float3 ng = normalize(cross(e1, e2));
                const float t_dx = dot(p - rdx.o, ng) / dot(rdx.d, ng), t_dy = dot(p - rdy.o, ng) / dot(rdy.d, ng);
                const float3 dpdx = f3(fmaf(t_dx, rdx.d.x, rdx.o.x), fmaf(t_dx, rdx.d.y, rdx.o.y), fmaf(t_dx, rdx.d.z, rdx.o.z)) - p;
                const float3 dpdy = f3(fmaf(t_dy, rdy.d.x, rdy.o.x), fmaf(t_dy, rdy.d.y, rdy.o.y), fmaf(t_dy, rdy.d.z, rdy.o.z)) - p;
                const float m00 = dot(dpdu, dpdu), m01 = dot(dpdu, dpdv), m11 = dot(dpdv, dpdv);
                const float mdet = m00 * m11 - m01 * m01;
                const float j00 = m11 / mdet, j01 = -m01 / mdet, j11 = m00 / mdet;
                const float bx0 = dot(dpdu, dpdx), bx1 = dot(dpdv, dpdx), by0 = dot(dpdu, dpdy), by1 = dot(dpdv, dpdy);
                g = make_float4(fmaf(j01, bx1, j00 * bx0), fmaf(j11, bx1, j01 * bx0), fmaf(j01, by1, j00 * by0), fmaf(j11, by1, j01 * by0));
            }
            if (!any_nan(g)) acc = acc + g;                                                // uvgrad.py:96-97
        }
        a.image[(size_t)x + (size_t)y * a.W] = make_float4(acc.x / (float)a.spp, acc.y / (float)a.spp, acc.z / (float)a.spp, acc.w / (float)a.spp);
    }
    flush_tally(a, tl);
}

__global__ void k_probe_trace(const SceneView sc, const float* __restrict__ rays, int n, int any, int* __restrict__ ip,
                              float* __restrict__ uvt, int* __restrict__ occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rays + 8 * (size_t)i;
    const Ray ray = make_ray(f3(r[0], r[1], r[2]), f3(r[3], r[4], r[5]), r[6], r[7]);
    TraceStats ts = {};
    if (any) { occ[i] = trace<true, false>(sc, ray, ts).inst >= 0 ? 1 : 0; return; }
    const Hit h = trace<false, false>(sc, ray, ts);
    ip[2 * i] = h.inst; ip[2 * i + 1] = h.prim;
    uvt[3 * i] = h.u; uvt[3 * i + 1] = h.v; uvt[3 * i + 2] = h.t;
}

__global__ void k_probe_sampler(const SceneView sc, const int* __restrict__ state, int n, const char* __restrict__ prog,
                                int prog_len, int per, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int* st = state + 5 * (size_t)i;
    Sampler sm = make_sampler(sc, st[0], st[1], st[2], st[3], st[4]);
    float* o = out + (size_t)per * i;
    for (int k = 0; k < prog_len; ++k) {
        if (prog[k] == '2') { const float2 u = sampler_next2f(sc, sm); *o++ = u.x; *o++ = u.y; }
        else *o++ = sampler_next(sc, sm);
    }
}

__global__ void k_probe_ggx(const float* __restrict__ in, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = in + 13 * (size_t)i;
    const float3 wo = f3(a[0], a[1], a[2]), wi = f3(a[3], a[4], a[5]), d = f3(a[6], a[7], a[8]), g = f3(a[10], a[11], a[12]);
    const float3 f = ggx_brdf(wo, wi, d, 0.04f, a[9]);
    const float4 dm = ggx_brdf_grad(wo, wi, 0.04f, a[9], g);
    float* o = out + 7 * (size_t)i;
    o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = dm.x; o[4] = dm.y; o[5] = dm.z; o[6] = dm.w;
}

dim3 render_grid(const RenderArgs& a) {
    const int tiles = (a.row1 - a.row0 + 7) / 8;
    return dim3((a.W + 15) / 16, (tiles - a.shard_index + a.shard_count - 1) / a.shard_count);
}

template <int I>
cudaError_t launch_fwd_i(const RenderArgs& a, bool stats, cudaStream_t s) {
    ++g_kernel_launches;
    if (stats) k_forward<I, true><<<render_grid(a), kBlockThreads, 0, s>>>(a);
    else k_forward<I, false><<<render_grid(a), kBlockThreads, 0, s>>>(a);
    return cudaGetLastError();
}
template <int I>
cudaError_t launch_bwd_i(const RenderArgs& a, bool stats, cudaStream_t s) {
    ++g_kernel_launches;
    if (a.aggregate_scatter) {
        const size_t gc = a.block_scatter ? kGradCacheBytes : 0;
        if (stats) k_backward<I, true, true><<<render_grid(a), kBlockThreads, gc, s>>>(a);
        else k_backward<I, false, true><<<render_grid(a), kBlockThreads, gc, s>>>(a);
    } else {
        if (stats) k_backward<I, true, false><<<render_grid(a), kBlockThreads, 0, s>>>(a);
        else k_backward<I, false, false><<<render_grid(a), kBlockThreads, 0, s>>>(a);
    }
    return cudaGetLastError();
}

template <bool BWD, bool STATS, bool AGG>
void launch_path3(bool env, const RenderArgs& a, cudaStream_t s) {
    ++g_kernel_launches;
    const size_t gc = (BWD && AGG && a.block_scatter) ? kGradCacheBytes : 0;
    if (env) k_path<BWD, STATS, AGG, true><<<render_grid(a), kBlockThreads, gc, s>>>(a);
    else k_path<BWD, STATS, AGG, false><<<render_grid(a), kBlockThreads, gc, s>>>(a);
}
}  // namespace

cudaError_t launch_forward(int integrator, const RenderArgs& a, bool stats, cudaStream_t s) {
    if (a.row1 <= a.row0 || a.it1 <= a.it0 || render_grid(a).y == 0) return cudaSuccess;
    switch (integrator) {
        case ZDR_COLLOCATED: return launch_fwd_i<ZDR_COLLOCATED>(a, stats, s);
        case ZDR_DIRECT: return launch_fwd_i<ZDR_DIRECT>(a, stats, s);
        default:
            if (stats) launch_path3<false, true, false>(a.sc.env_count > 0, a, s);
            else launch_path3<false, false, false>(a.sc.env_count > 0, a, s);
            return cudaGetLastError();
    }
}
cudaError_t launch_backward(int integrator, const RenderArgs& a, bool stats, cudaStream_t s) {
    if (a.row1 <= a.row0 || a.it1 <= a.it0 || render_grid(a).y == 0) return cudaSuccess;
    switch (integrator) {
        case ZDR_COLLOCATED: return launch_bwd_i<ZDR_COLLOCATED>(a, stats, s);
        case ZDR_DIRECT: return launch_bwd_i<ZDR_DIRECT>(a, stats, s);
        default:
            if (a.aggregate_scatter) {
                if (stats) launch_path3<true, true, true>(a.sc.env_count > 0, a, s);
                else launch_path3<true, false, true>(a.sc.env_count > 0, a, s);
            } else {
                if (stats) launch_path3<true, true, false>(a.sc.env_count > 0, a, s);
                else launch_path3<true, false, false>(a.sc.env_count > 0, a, s);
            }
            return cudaGetLastError();
    }
}
#ifndef ZDR_FAST_MATH      // exist once (exact build): uv Jacobian, counters, workspace sizes
cudaError_t launch_duvdxy(const RenderArgs& a, cudaStream_t s) {
    if (a.row1 <= a.row0 || a.it1 <= a.it0 || render_grid(a).y == 0) return cudaSuccess;
    ++g_kernel_launches;
    k_uvgrad<<<render_grid(a), kBlockThreads, 0, s>>>(a);
    return cudaGetLastError();
}

unsigned long long kernel_launch_count() { return g_kernel_launches.load(); }
#endif

#ifndef ZDR_FAST_MATH      // the probes exist once (exact build)
cudaError_t launch_probe_trace(const SceneView& sc, const float* rays, int n, bool any, int* ip, float* uvt, int* occ,
                               cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    ++g_kernel_launches;
    k_probe_trace<<<(n + 127) / 128, 128, 0, s>>>(sc, rays, n, any ? 1 : 0, ip, uvt, occ);
    return cudaGetLastError();
}
cudaError_t launch_probe_sampler(const SceneView& sc, const int* state, int n, const char* prog, int prog_len, int per,
                                 float* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    ++g_kernel_launches;
    k_probe_sampler<<<(n + 127) / 128, 128, 0, s>>>(sc, state, n, prog, prog_len, per, out);
    return cudaGetLastError();
}
cudaError_t launch_probe_ggx(const float* in, int n, float* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    ++g_kernel_launches;
    k_probe_ggx<<<(n + 127) / 128, 128, 0, s>>>(in, n, out);
    return cudaGetLastError();
}
#endif

}  // namespace zdr
