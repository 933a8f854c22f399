// render_ctx.cuh — what every render kernel's translation unit shares: the per-thread tallies and their flush, the kernel
// context (scene queries + gradient scatter with the launch's scatter mode), the shared-memory gradient duplicate's set-up.
#pragma once
#include <atomic>
#include "render.cuh"
#include "render_common.cuh"

namespace zdr {

extern std::atomic<unsigned long long> g_kernel_launches;      // one counter, defined by render.cu (exact build)

namespace {

// per-thread tallies, reduced per warp and flushed with one atomic per counter per warp
struct Tally {
    unsigned closest, any, scatter, red, shaded, samples;
    TraceStats ts;
};

template <bool STATS, bool AGG = false>
struct Ctx {
    const RenderArgs& a;
    Tally& tl;
    GradCache gc = grad_cache_null();      // AGG kernels: the CTA's shared-memory duplicate of the gradient buffer (shading.cuh)
    ZDR_DEV Hit closest(const Ray& r) { tl.closest++; return trace<false, STATS>(a.sc, r, tl.ts); }
    ZDR_DEV bool occluded(const Ray& r) { tl.any++; return trace<true, STATS>(a.sc, r, tl.ts).inst >= 0; }
    ZDR_DEV void scatter(float2 uv, float4 g, int set) {
        tl.scatter++;
        tl.red += AGG ? write_bsdf_grad_aggregated(a.tex, uv, g, gc, set) : write_bsdf_grad(a.tex, uv, g, set);
    }
};

// AGG kernels: declare, clear and (at the end, by every thread of the CTA) fold the shared-memory gradient duplicate
// (dynamic shared memory, kGradCacheBytes, passed by the launcher only in block-scatter mode: the default launches carry none)
#define ZDR_GRAD_CACHE_BEGIN(c)                                               \
    extern __shared__ __align__(16) unsigned char gc_dyn_[];                 \
    if (AGG && a.block_scatter) {                                            \
        (c).gc.key = reinterpret_cast<int*>(gc_dyn_);                        \
        (c).gc.acc = reinterpret_cast<float*>(gc_dyn_ + 4 * kGradCacheSlots); \
        grad_cache_init((c).gc);                                             \
    }
#define ZDR_GRAD_CACHE_END(c, a, tl) \
    if (AGG && (a).block_scatter) (tl).red += grad_cache_flush((c).gc, (a).tex.dmat);

ZDR_DEV void flush_tally(const RenderArgs& a, const Tally& t) {
    const unsigned v[C_NUM] = {t.closest, t.any, t.ts.nodes, t.ts.tris, t.scatter, t.red, t.shaded, t.samples};
#pragma unroll 1
    for (int i = 0; i < C_NUM; ++i) {
        const unsigned s = warp_sum(v[i]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(a.counters + i, (unsigned long long)s);
    }
}

}  // namespace
}  // namespace zdr
