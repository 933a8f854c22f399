// persistent.cu — the path integrator (prb.py:18-88) and its Path-Replay-Backprop adjoint (prb.py:91-187) as ONE persistent
// kernel per pass whose path state and ray queues live in SHARED memory (BASELINE north star item 5).
//
// One resident CTA per SM owns a pool of `pool` paths (112 bytes of state each, structure-of-arrays in shared memory) and
// runs the stages of the wavefront as barrier-separated PHASES over shared-memory queues, round after round:
//
//   claim     one thread takes as many new camera samples from the launch's global counter as the pool has free slots
//   R  regen  free slots <- new camera samples (integrator.py:17-25); adjoint pass: replayed samples restart at the camera
//             with the ORIGINAL sampler state, carrying Le (prb.py:97-101)                          -> queue L0
//   E  extend closest-hit traversal of L0, rays fetched dynamically from the queue (Aila & Laine)   -> hit records
//   A  shade  miss / backface / emitter / inst>0 tests (prb.py:26-46), light sample (prb.py:57)     -> queue L1 | free | restart
//   S  shadow occlusion query of L1 (prb.py:58-59)
//   B  shade  NEE term, BSDF sample, throughput, Russian roulette (prb.py:61-87); adjoint walk: BRDF adjoints + gradient
//             scatter (prb.py:138-187)                                                               -> queue L0' | free | restart
//
// A path that ends frees its slot, which the next round refills: every phase of every round works on a full pool, the
// stages' code runs with the whole CTA in the same phase (instruction-cache friendly, unlike a one-kernel state machine),
// and nothing but finished samples (forward: 16 bytes per sample, summed per pixel IN SAMPLE ORDER by k_pw_accumulate,
// integrator.py:15-28, so images stay bit-identical to the oracle's) ever leaves the SM.  Against the HBM wavefront of
// wavefront.cu (k_wf_*: ~100 launches and ~1.7 kB of DRAM traffic per path and pass) a pass is 1 launch (+1 accumulate).
// Replay and adjoint-walk paths of the adjoint pass share the pool but never a warp of stage B: L1 is filled from both
// ends (replay paths bottom-up, adjoint paths top-down).
//
// Small scenes (Cornell box: 4 KB) are staged into shared memory with cp.async.bulk (TMA bulk copy, mbarrier
// completion) at kernel start, and the traversal stack is a column of shared memory per thread: the hot loops then touch
// no global memory at all except texels, sampler tables and the scatter.
//
// Arithmetic per sample is unchanged (the device functions of shading.cuh / sampler.cuh / traverse.cuh).
#define ZDR_SCENE_GENERIC_LOADS 1
#define ZDR_INLINE_RCP_LENGTH 1      // normalize() inline: 5 277 -> 5 440 Mrays/s on C3 at 64 spp (profiles/r03b)
#define ZDR_INLINE_GGX_EVAL 1
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include "render.cuh"
#include "render_common.cuh"

namespace zdr {

extern std::atomic<unsigned long long> g_kernel_launches;      // render.cu

namespace {

#ifndef ZDR_PW_THREADS
#define ZDR_PW_THREADS 512
#endif
#ifndef ZDR_PW_BLOCKS          // resident CTAs per SM (each with its own pool; their phases drift apart and fill each other's barrier tails)
#define ZDR_PW_BLOCKS 1
#endif
constexpr int kPwThreads = ZDR_PW_THREADS;
constexpr int kPwBlocks = ZDR_PW_BLOCKS;
#ifndef ZDR_PW_GSTATE          // 1: the path state records live in global memory (a per-CTA slice of the workspace, L2-resident:
#define ZDR_PW_GSTATE 0        //    148 pools x 4096 slots x 112 B = 68 MB of the 126 MB L2), only the queues in shared memory
#endif
constexpr bool kPwGlobalState = ZDR_PW_GSTATE != 0;
constexpr int kPwStateBytes = 7 * 16;                  // 7 float4 records per slot
constexpr int kPwBytesPerSlot = (kPwGlobalState ? 0 : kPwStateBytes) + 4 * 2;       // shared memory per slot: (state +) 4 queue entries
constexpr int kPwHeader = 256 + kGradCacheSlots * 20;      // counters + mbarrier, then the CTA's gradient duplicate (shading.cuh GradCache)
constexpr int kPwNodeBudget = 12;

// counters in shared memory (ints)
enum {
    PW_N0A = 0, PW_N0B, PW_N1A, PW_N1B, PW_NF, PW_NRS, PW_CUR_E, PW_CUR_S, PW_CLAIM_BASE, PW_CLAIM_N, PW_DRAINED, PW_ABORT,
    PW_TALLY0 = 16,            // C_NUM tallies
    PW_MBAR = 32               // 8-byte mbarrier at byte offset 128
};

ZDR_DEV int round32(int n) { return (n + 31) & ~31; }

// survivors of a warp are appended to a shared-memory queue with one atomic; `down`: the queue grows from index top downwards
ZDR_DEV void pw_append(unsigned short* list, int* counter, bool keep, int value, bool down = false, int top = 0) {
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
        const int k = base + __popc(m & ((1u << lane) - 1u));
        list[down ? top - k : k] = (unsigned short)value;
    }
}
ZDR_DEV void pw_tally(int* ctr, int which, unsigned v) {
    const unsigned s = warp_sum(v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(reinterpret_cast<unsigned*>(ctr) + PW_TALLY0 + which, s);
}
ZDR_DEV void pw_pixel(const RenderArgs& a, int pix, int& x, int& y) {
    const int lr = fastdiv(pix, a.pw.div_W);
    x = pix - lr * a.W;
    y = a.row0 + ((lr >> 3) * a.shard_count + a.shard_index) * 8 + (lr & 7);
}
ZDR_DEV Sampler pw_sampler(const RenderArgs& a, unsigned id) {
    const int j = fastdiv((int)id, a.pw.div_P), pix = (int)id - j * a.pw.P;
    int x, y;
    pw_pixel(a, pix, x, y);
    return make_sampler(a.sc, x, y, a.seed, a.spp, a.pw.it_base + j);
}

struct Pool {          // shared-memory path state, one float4 per slot and array
    float4* ro;        // ray origin.xyz | pdf_bsdf            (after stage A: hit point.xyz | light.dist = the shadow ray)
    float4* rd;        // ray direction.xyz | packed (depth, phase, sampler dimension)
    float4* bi;        // beta.xyz | camera-sample id (uint bits)
    float4* Ls;        // L.xyz (radiance accumulated / Le remaining) | -
    float4* hit;       // closest hit: slot | u | v | t
    float4* lw;        // light.wi.xyz | light.pdf
    float4* le;        // light.eval.xyz | occluded (written by stage S)
};

// per-thread traversal stack: a column of shared memory (small scenes) or local memory
template <bool SMEM>
struct TStack;
template <>
struct TStack<true> {
    int* col;
    ZDR_DEV void push(int& sp, int v) { col[sp * kPwThreads] = v; ++sp; }
    ZDR_DEV int pop(int& sp) { --sp; return col[sp * kPwThreads]; }
};
template <>
struct TStack<false> {
    int s[kStackSize];
    ZDR_DEV void push(int& sp, int v) { s[sp++] = v; }
    ZDR_DEV int pop(int& sp) { return s[--sp]; }
};

// Stage E (ANY = false: closest hit of queue `lo` -> pool.hit) and stage S (ANY = true: occlusion of the two-ended queue
// `lo` / `hi` -> pool.le.w): the persistent dynamic-fetch traversal of wavefront.cu's k_wf_trace with the queue, its cursor,
// the rays and (small scenes) the tree, the triangles and the stack in shared memory.  A lane that finishes its ray takes the
// next queue entry; lanes still traversing leave the inner loop to let idle ones refill once fewer than `refill` are left.
template <bool ANY, bool STATS, bool SMALL>
ZDR_DEV void pw_trace(const SceneView& sc, const Pool& pool, const unsigned short* lo, int n_lo, const unsigned short* hi_top, int n_hi,
                      int* cursor, int* stack_col, int refill, int* ctr) {
    const int n = n_lo + n_hi;
    const float4* nodes = sc.nodes;
    const float4* tris = sc.tris;
    const int lane = threadIdx.x & 31;
    constexpr bool BIG = !SMALL;
    unsigned n_rays = 0;
    TraceStats ts = {};
    bool has_ray = false, exhausted = false;
    int slot = 0, link = kLinkSentinel, sp = 0, hit_slot = -1, best_gid = 0x7fffffff;
    float tmin = 0.f, tmax = 0.f, best = 0.f, hit_u = 0.f, hit_v = 0.f, hit_t = 0.f;
    float3 id = f3(0.f, 0.f, 0.f), oid = id;
    RayPre rp = {};
    TStack<SMALL> stack;
    if constexpr (SMALL) stack.col = stack_col;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !has_ray && !exhausted);
        if (idle) {
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(cursor, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!has_ray && !exhausted) {
                const int i = base + __popc(idle & ((1u << lane) - 1u));
                if (i >= n) {
                    exhausted = true;
                } else {
                    slot = i < n_lo ? lo[i] : hi_top[-(i - n_lo)];
                    float3 o, d;
                    const float4 A = pool.ro[slot];
                    o = f3(A.x, A.y, A.z);
                    if (ANY) {
                        const float4 B = pool.lw[slot];
                        d = f3(B.x, B.y, B.z); tmin = 1e-4f; tmax = A.w;
                    } else {
                        const float4 B = pool.rd[slot];
                        d = f3(B.x, B.y, B.z); tmin = 0.0f; tmax = 1e30f;
                    }
                    n_rays++;
                    rp = ray_pre(o, d);
                    id = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    if (!(fabsf(id.x) <= 1e30f)) id.x = copysignf(1e30f, d.x);
                    if (!(fabsf(id.y) <= 1e30f)) id.y = copysignf(1e30f, d.y);
                    if (!(fabsf(id.z) <= 1e30f)) id.z = copysignf(1e30f, d.z);
                    oid = f3(o.x * id.x, o.y * id.y, o.z * id.z);
                    best = tmax; best_gid = 0x7fffffff; hit_slot = -1; hit_u = hit_v = hit_t = 0.0f;
                    sp = 0; stack.push(sp, kLinkSentinel); link = sc.root_link;
                    has_ray = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, !has_ray)) {
            if (__all_sync(0xffffffffu, exhausted)) break;
            continue;
        }
        while (has_ray) {
            int budget = kPwNodeBudget;
            while (link >= 0 && (!BIG || budget-- > 0)) {        // inner nodes
                const float4* nd = nodes + 4 * (size_t)link;
                float4 n0, n1, n2, n3;
                ld_node(nd, n0, n1, n2, n3);
                if (STATS) ts.nodes++;
                float t0, t1;
                const bool h0 = box_test(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, id, oid, tmin, best, t0);
                const bool h1 = box_test(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, id, oid, tmin, best, t1);
                const int l0 = __float_as_int(n3.x), l1 = __float_as_int(n3.y);
                if (h0 && h1) {
                    const bool swap = t1 < t0;
                    stack.push(sp, swap ? l0 : l1);
                    link = swap ? l1 : l0;
                } else if (h0 || h1) {
                    link = h0 ? l0 : l1;
                } else {
                    link = stack.pop(sp);
                }
            }
            bool done = link == kLinkSentinel;
            if (!done && link < 0) {                     // leaf
                const int code = ~link;
                const int first = code >> 2, count = (code & 3) + 1;
                for (int k = 0; k < count; ++k) {
                    const float* tp = reinterpret_cast<const float*>(tris + 3 * (size_t)(first + k));
                    if (STATS) ts.tris++;
                    float t, u, v;
                    if (tri_test(rp, tp, tmin, t, u, v)) {
                        if (ANY) {
                            if (t < tmax) { hit_slot = first + k; done = true; break; }
                            continue;
                        }
                        const int gid = __float_as_int(ld_scene(tp + 3));
                        if (t < best || (hit_slot >= 0 && t == best && gid < best_gid)) {
                            best = t; best_gid = gid;
                            hit_t = t; hit_u = u; hit_v = v; hit_slot = first + k;
                        }
                    }
                }
                if (!done) { link = stack.pop(sp); done = link == kLinkSentinel; }
            }
            if (done && (link == kLinkSentinel || (ANY && hit_slot >= 0))) {
                if (ANY) pool.le[slot].w = hit_slot >= 0 ? 1.0f : 0.0f;
                else pool.hit[slot] = make_float4(__int_as_float(hit_slot), hit_u, hit_v, hit_t);
                has_ray = false;
                break;
            }
            if (__popc(__activemask()) < refill) break;
        }
    }
    pw_tally(ctr, ANY ? C_ANY : C_CLOSEST, n_rays);
    if (STATS) { pw_tally(ctr, C_NODES, ts.nodes); pw_tally(ctr, C_TRIS, ts.tris); }
}

// End of a path.  Forward: the finished sample goes to R (integrator.py:27-28) and the slot is free (returns 2).  Adjoint
// pass: a replayed sample (phase 0) whose Le is not NaN restarts as an adjoint walk (prb.py:97-101; returns 3), the end of
// an adjoint walk frees the slot.
template <int MODE>
ZDR_DEV int pw_finish(const RenderArgs& a, const Pool& pool, int slot, unsigned id, int phase, float3 L) {
    if (MODE == 0) {
        if (a.samples_out) {
            const int j = fastdiv((int)id, a.pw.div_P), pix = (int)id - j * a.pw.P;
            int x, y;
            pw_pixel(a, pix, x, y);
            float* o = a.samples_out + (((size_t)y * a.W + x) * a.spp + (a.pw.it_base + j)) * 3;
            o[0] = L.x; o[1] = L.y; o[2] = L.z;
        }
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!any_nan(L)) r = make_float4(clampf(L.x, 0.0f, 100000.0f), clampf(L.y, 0.0f, 100000.0f), clampf(L.z, 0.0f, 100000.0f), 0.f);
        __stcs(a.pw.R + id, r);                       // streamed: read once, by k_pw_accumulate
        return 2;
    }
    if (phase == 0 && !any_nan(L)) {
        pool.Ls[slot] = make_float4(L.x, L.y, L.z, 0.f);
        return 3;
    }
    return 2;
}

// ---- staging of a small scene into shared memory: cp.async.bulk (TMA bulk copy) completing on an mbarrier
ZDR_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
ZDR_DEV void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

template <int MODE, bool STATS, bool ENV, bool SMALL>
__global__ void __launch_bounds__(kPwThreads, kPwBlocks) k_pw(const __grid_constant__ RenderArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const PwState& w = a.pw;
    const int N = w.pool, tid = threadIdx.x;
    int* ctr = reinterpret_cast<int*>(smem);
    Pool pool;
    GradCache gcache = grad_cache_null();
    if (MODE && a.aggregate_scatter && a.block_scatter) {      // per-SM duplicate of the gradient buffer, alive for the whole pass, folded once at the end
        gcache.key = reinterpret_cast<int*>(smem + 256);
        gcache.acc = reinterpret_cast<float*>(smem + 256 + kGradCacheSlots * 4);
        grad_cache_init(gcache);
    }
    if constexpr (kPwGlobalState) pool.ro = w.gstate + (size_t)blockIdx.x * 7 * (size_t)N;
    else pool.ro = reinterpret_cast<float4*>(smem + kPwHeader);
    pool.rd = pool.ro + N; pool.bi = pool.rd + N; pool.Ls = pool.bi + N; pool.hit = pool.Ls + N; pool.lw = pool.hit + N; pool.le = pool.lw + N;
    unsigned short* l0a = kPwGlobalState ? reinterpret_cast<unsigned short*>(smem + kPwHeader) : reinterpret_cast<unsigned short*>(pool.le + N);
    unsigned short* l0b = l0a + N;
    unsigned short* l1 = l0b + N;          // stage-B queue: replay paths from index 0 up, adjoint-walk paths from N-1 down
    unsigned short* fr = l1 + N;           // free slots from index 0 up, restart entries from N-1 down
    unsigned char* tail = reinterpret_cast<unsigned char*>(fr + N);
    int* stack_col = reinterpret_cast<int*>(tail) + tid;
    SceneView sc = a.sc;
    if (tid < 32) ctr[tid] = 0;
    for (int i = tid; i < N; i += kPwThreads) fr[i] = (unsigned short)i;
    if constexpr (SMALL) {
        float4* s_nodes = reinterpret_cast<float4*>(tail + (size_t)w.stack_depth * kPwThreads * sizeof(int));
        float4* s_tris = s_nodes + 4 * (size_t)sc.num_nodes;
        float4* s_shade = s_tris + 3 * (size_t)sc.num_tris;
        const unsigned b_nodes = 64u * (unsigned)sc.num_nodes, b_tris = 48u * (unsigned)sc.num_tris, b_shade = 64u * (unsigned)sc.num_tris;
        if (w.bulk_copy) {
            unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 4 * PW_MBAR);
            if (tid == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(b_nodes + b_tris + b_shade) : "memory");
                if (b_nodes) bulk_copy_g2s(s_nodes, a.sc.nodes, b_nodes, bar);
                bulk_copy_g2s(s_tris, a.sc.tris, b_tris, bar);
                bulk_copy_g2s(s_shade, a.sc.shade, b_shade, bar);
            }
            unsigned ok = 0;
            for (int spin = 0; !ok && spin < (1 << 22); ++spin) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok)
                             : "r"(smem_addr(bar))
                             : "memory");
            }
            if (!ok) {            // never seen; a stuck copy must not hang the device: fall back to the loop below (same bytes)
                for (int i = tid; i < 4 * sc.num_nodes; i += kPwThreads) s_nodes[i] = __ldg(a.sc.nodes + i);
                for (int i = tid; i < 3 * sc.num_tris; i += kPwThreads) s_tris[i] = __ldg(a.sc.tris + i);
                for (int i = tid; i < 4 * sc.num_tris; i += kPwThreads) s_shade[i] = __ldg(a.sc.shade + i);
            }
        } else {
            for (int i = tid; i < 4 * sc.num_nodes; i += kPwThreads) s_nodes[i] = __ldg(a.sc.nodes + i);
            for (int i = tid; i < 3 * sc.num_tris; i += kPwThreads) s_tris[i] = __ldg(a.sc.tris + i);
            for (int i = tid; i < 4 * sc.num_tris; i += kPwThreads) s_shade[i] = __ldg(a.sc.shade + i);
        }
        sc.nodes = s_nodes; sc.tris = s_tris; sc.shade = s_shade;
    }
    __syncthreads();
    if (tid == 0) ctr[PW_NF] = N;
    const bool corrected = a.prb_mode == ZDR_PRB_CORRECTED;
    const float specular = 0.04f;
    const float3 zero = f3(0.f, 0.f, 0.f);
    int cur = 0;                                  // which of l0a / l0b is the live queue of this round
    const long long t_start = clock64();
    // phase profile (w.phase_cycles != null): thread 0 sums the cycles between the barriers of every phase
    long long t_mark = t_start, t_ph[6] = {0, 0, 0, 0, 0, 0};
    long long n_rounds = 0, n_paths = 0;
#define PW_MARK(k) do { if (w.phase_cycles && tid == 0) { const long long t_ = clock64(); t_ph[k] += t_ - t_mark; t_mark = t_; } } while (0)
    for (;;) {
        // ---------------------------------------------------------------------------------- claim
        __syncthreads();
        if (tid == 0) {
            // watchdog: a pass takes well under a second; a kernel still here after ~2^36 cycles (half a minute) gives up
            // instead of hanging the device (its samples are then missing: parity tests fail loudly)
            if (clock64() - t_start > (1ll << 36)) ctr[PW_ABORT] = 1;
            const int want = ctr[PW_DRAINED] ? 0 : ctr[PW_NF];
            int got = 0;
            unsigned base = 0;
            if (want > 0) {
                base = atomicAdd(w.next, (unsigned)want);
                got = base >= w.total ? 0 : (int)min((unsigned)want, w.total - base);
                if (got < want) ctr[PW_DRAINED] = 1;
            }
            ctr[PW_CLAIM_BASE] = (int)base;
            ctr[PW_CLAIM_N] = got;
        }
        __syncthreads();
        unsigned short* l0 = cur ? l0b : l0a;
        int* n0 = ctr + (cur ? PW_N0B : PW_N0A);
        unsigned short* l0n = cur ? l0a : l0b;
        int* n0n = ctr + (cur ? PW_N0A : PW_N0B);
        // ---------------------------------------------------------------------------------- R: regenerate / restart
        {
            const int nrs = MODE ? ctr[PW_NRS] : 0, got = ctr[PW_CLAIM_N];
            const unsigned base = (unsigned)ctr[PW_CLAIM_BASE];
            const int nwork = nrs + got, nround = round32(nwork);
            unsigned samples = 0;
            for (int i = tid; i < nround; i += kPwThreads) {
                const bool act = i < nwork;
                int slot = 0;
                if (act) {
                    const bool restart = i < nrs;
                    unsigned id;
                    if (restart) { slot = fr[N - 1 - i]; id = __float_as_uint(pool.bi[slot].w); }
                    else { slot = fr[i - nrs]; id = base + (unsigned)(i - nrs); samples++; }
                    const int j = fastdiv((int)id, w.div_P), pix = (int)id - j * w.P;
                    int x, y;
                    pw_pixel(a, pix, x, y);
                    Sampler sampler = make_sampler(a.sc, x, y, a.seed, a.spp, w.it_base + j);
                    const Ray ray = camera_sample(a, x, y, sampler);
                    pool.ro[slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, 1e30f);                                  // prb.py:22
                    pool.rd[slot] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(wf_pack(0, restart ? 1 : 0, sampler.dimension)));
                    pool.bi[slot] = make_float4(1.f, 1.f, 1.f, __uint_as_float(id));
                    if (!restart) pool.Ls[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                pw_append(l0, n0, act, slot);
            }
            pw_tally(ctr, C_SAMPLES, samples);
        }
        __syncthreads();
        PW_MARK(0);
        const int n_live = *n0;
        if (n_live == 0 || ctr[PW_ABORT]) break;        // nothing alive and nothing left to start
        n_rounds++; n_paths += n_live;
        // ---------------------------------------------------------------------------------- E: extend
        if (tid == 0) { ctr[PW_NF] = 0; ctr[PW_NRS] = 0; ctr[PW_N1A] = 0; ctr[PW_N1B] = 0; ctr[PW_CUR_S] = 0; *n0n = 0; }
        pw_trace<false, STATS, SMALL>(sc, pool, l0, n_live, nullptr, 0, ctr + PW_CUR_E, stack_col, w.refill_closest, ctr);
        __syncthreads();
        PW_MARK(1);
        const bool drained = ctr[PW_DRAINED] != 0;
        // ---------------------------------------------------------------------------------- A: classify, light sample
        {
            if (tid == 0) ctr[PW_CUR_E] = 0;
            const int nround = round32(n_live);
            unsigned shaded = 0;
            for (int i = tid; i < nround; i += kPwThreads) {
                int slot = 0, dest = 0, phase = 0;       // dest: 1 -> stage B queue, 2 -> free, 3 -> restart
                if (i < n_live) {
                    slot = l0[i];
                    const float4 RO = pool.ro[slot], RD = pool.rd[slot], H = pool.hit[slot];
                    const unsigned packed = __float_as_uint(RD.w);
                    const int depth = (int)(packed & 0xffu);
                    phase = (packed & kWfPhaseBit) ? 1 : 0;
                    const Ray ray = make_ray(f3(RO.x, RO.y, RO.z), f3(RD.x, RD.y, RD.z), 0.0f, 1e30f);
                    const float pdf_bsdf = RO.w;
                    const unsigned id = __float_as_uint(pool.bi[slot].w);
                    Hit hit;
                    hit.slot = __float_as_int(H.x); hit.u = H.y; hit.v = H.z; hit.t = H.w; hit.inst = -1; hit.prim = -1;
                    bool end = false;
                    float3 beta = zero, L = zero;
                    if (hit.slot < 0) {
                        const float4 B4 = pool.bi[slot], L4 = pool.Ls[slot];
                        beta = f3(B4.x, B4.y, B4.z); L = f3(L4.x, L4.y, L4.z);
                        if (ENV && phase == 0 && sc.env_count > 0) {                 // prb.py:26-31
                            const float3 tex = env_texture(sc, direction_to_uv(ray.d));
                            const float3 emission = corrected ? tex : beta * tex;
                            const float mis_weight = balanced_heuristic(pdf_bsdf, env_sampled_light_pdf(sc, ray.d));
                            L = L + (beta * mis_weight) * emission;
                        }
                        end = true;
                    } else {
                        const float4* tp = sc.tris + 3 * (size_t)hit.slot;
                        hit.inst = __float_as_int(ld_scene(tp + 1).w);
                        hit.prim = __float_as_int(ld_scene(tp + 2).w);
                        const Interaction itx = surface_interact(sc, hit);
                        const float3 emission = inst_emission(sc, hit.inst);
                        const bool back = backfacing(ray.d, itx);                  // prb.py:35-36
                        const bool emit = !back && any_pos(emission);              // prb.py:39-44 / :117-121
                        if (back || emit || ((phase == 0 || corrected) && inst_material_set(sc, hit.inst) < 0)) {      // prb.py:45-46
                            const float4 B4 = pool.bi[slot], L4 = pool.Ls[slot];
                            beta = f3(B4.x, B4.y, B4.z); L = f3(L4.x, L4.y, L4.z);
                            if (emit) {
                                const float mis_weight = balanced_heuristic(pdf_bsdf, sample_light_pdf(sc, ray.o, hit, itx.p));
                                const float3 contrib = (beta * mis_weight) * emission;
                                L = (phase == 0) ? L + contrib : L - contrib;
                            }
                            end = true;
                        } else {
                            shaded++;
                            Sampler sampler = pw_sampler(a, id);
                            sampler.dimension = (int)(packed >> 12);
                            const LightSample light = sample_light<ENV>(sc, itx.p, sampler);                    // prb.py:57
                            pool.ro[slot] = make_float4(itx.p.x, itx.p.y, itx.p.z, light.dist);
                            pool.lw[slot] = make_float4(light.wi.x, light.wi.y, light.wi.z, light.pdf);
                            pool.le[slot] = make_float4(light.eval.x, light.eval.y, light.eval.z, 0.f);
                            pool.rd[slot] = make_float4(RD.x, RD.y, RD.z, __uint_as_float(wf_pack(depth, phase, sampler.dimension)));
                            dest = 1;
                        }
                    }
                    if (end) dest = pw_finish<MODE>(a, pool, slot, id, phase, L);
                }
                pw_append(l1, ctr + PW_N1A, dest == 1 && phase == 0, slot);
                if (MODE) pw_append(l1, ctr + PW_N1B, dest == 1 && phase == 1, slot, true, N - 1);
                pw_append(fr, ctr + PW_NF, dest == 2 && !drained, slot);
                if (MODE) pw_append(fr, ctr + PW_NRS, dest == 3, slot, true, N - 1);
            }
            pw_tally(ctr, C_SHADED, shaded);
        }
        __syncthreads();
        PW_MARK(2);
        const int n1a = ctr[PW_N1A], n1b = MODE ? ctr[PW_N1B] : 0;
        // ---------------------------------------------------------------------------------- S: shadow rays
        pw_trace<true, STATS, SMALL>(sc, pool, l1, n1a, l1 + (N - 1), n1b, ctr + PW_CUR_S, stack_col, w.refill_any, ctr);
        __syncthreads();
        PW_MARK(3);
        // ---------------------------------------------------------------------------------- B: NEE, BSDF sample, adjoints
        {
            const int ra = round32(n1a), rb = round32(n1b);
            unsigned scatters = 0, reds = 0;
            for (int i = tid; i < ra + rb; i += kPwThreads) {
                const bool ADJ = MODE && i >= ra;              // warp-uniform: replay and adjoint-walk paths never share a warp
                bool act;
                int slot = 0, dest = 0;
                if (!ADJ) { act = i < n1a; if (act) slot = l1[i]; }
                else { const int k = i - ra; act = k < n1b; if (act) slot = l1[N - 1 - k]; }
                if (act) {
                    const float4 RD = pool.rd[slot], H = pool.hit[slot], B4 = pool.bi[slot], L4 = pool.Ls[slot];
                    const float4 LW = pool.lw[slot], LE = pool.le[slot];
                    const unsigned packed = __float_as_uint(RD.w);
                    int depth = (int)(packed & 0xffu);
                    const unsigned id = __float_as_uint(B4.w);
                    const float3 ray_d = f3(RD.x, RD.y, RD.z);
                    float3 beta = f3(B4.x, B4.y, B4.z), L = f3(L4.x, L4.y, L4.z);
                    const int j = fastdiv((int)id, w.div_P), pix = (int)id - j * w.P;
                    int x, y;
                    pw_pixel(a, pix, x, y);
                    Sampler sampler = make_sampler(a.sc, x, y, a.seed, a.spp, w.it_base + j);
                    sampler.dimension = (int)(packed >> 12);
                    Hit hit;
                    hit.slot = __float_as_int(H.x); hit.u = H.y; hit.v = H.z; hit.t = H.w;
                    {
                        const float4* tp = sc.tris + 3 * (size_t)hit.slot;
                        hit.inst = __float_as_int(ld_scene(tp + 1).w);
                        hit.prim = __float_as_int(ld_scene(tp + 2).w);
                    }
                    float3 le_grad = zero;
                    if (ADJ) {                                         // integrator.py:38-40
                        const float4 g = __ldg(a.d_image + ((size_t)x + (size_t)y * a.W));
                        le_grad = f3(g.x / (float)a.spp, g.y / (float)a.spp, g.z / (float)a.spp);
                        if (any_nan(le_grad)) le_grad = zero;
                    }
                    const Interaction itx = surface_interact(sc, hit);
                    const int mset = inst_texture_set(sc, hit.inst);
                    const float4 mat = read_bsdf(a.tex, itx.uv, mset);
                    const float3 diffuse = f3(mat.x, mat.y, mat.z);
                    const float roughness = mat.w;
                    const Onb onb = make_onb(itx.ns);
                    const float3 wo_local = to_local(onb, -ray_d);
                    const float3 light_wi = f3(LW.x, LW.y, LW.z), light_eval = f3(LE.x, LE.y, LE.z);
                    const float light_pdf = LW.w;
                    const bool occluded = LE.w != 0.0f;
                    float4 mat_grad = make_float4(0.f, 0.f, 0.f, 0.f);
                    bool end = false;
                    float pdf_bsdf;
                    float3 ray_o = zero, ray_nd = zero;
                    const float3 wi_light_local = to_local(onb, light_wi);
                    if (!occluded && wi_light_local.z >= 1e-4f) {                                       // prb.py:61-66
                        float2 gb = make_float2(0.f, 0.f);      // adjoint: (dk, common) of ggx_brdf_grad, computed with the value (shading.cuh ggx_eval_adj)
                        const float4 ev = ADJ ? ggx_eval_adj(wo_local, wi_light_local, diffuse, specular, roughness, gb)
                                              : ggx_eval(wo_local, wi_light_local, diffuse, specular, roughness);
                        const float3 bsdf = f3(ev.x, ev.y, ev.z);
                        const float mis_weight = balanced_heuristic(light_pdf, ev.w);
                        const float dn = fmaxf(light_pdf, 1e-4f);
                        const float3 contrib = (((beta * bsdf) * mis_weight) * light_eval) / dn;
                        if (!ADJ) {
                            L = L + contrib;
                        } else {
                            const float3 cot = (((beta * mis_weight) * light_eval) / dn) * le_grad;     // prb.py:145
                            mat_grad = mat_grad + ggx_grad_from_basis(gb, cot);
                            L = L - contrib;                                                        // prb.py:147
                        }
                    }
                    const float3 wi_local = ggx_sample(a.sc, wo_local, roughness, sampler);             // prb.py:69
                    float2 gb = make_float2(0.f, 0.f);
                    const float4 ev = ADJ ? ggx_eval_adj(wo_local, wi_local, diffuse, specular, roughness, gb) : ggx_eval(wo_local, wi_local, diffuse, specular, roughness);
                    pdf_bsdf = ev.w;
                    const float3 wi = to_world(onb, wi_local);
                    if (dot(wi, itx.ng) < 1e-4f || wi_local.z < 1e-4f) {                                // prb.py:73-74
                        end = true;
                    } else {
                        ray_o = offset_ray_origin(itx.p, itx.ng);                                       // prb.py:75
                        ray_nd = wi;
                        const float3 bsdf = f3(ev.x, ev.y, ev.z);
                        if (ADJ) {
                            float3 cot;
                            if (corrected)
                                cot = f3(bsdf.x != 0.0f ? L.x / bsdf.x * le_grad.x : 0.0f, bsdf.y != 0.0f ? L.y / bsdf.y * le_grad.y : 0.0f,
                                         bsdf.z != 0.0f ? L.z / bsdf.z * le_grad.z : 0.0f);
                            else
                                cot = ((beta / pdf_bsdf) * L) * le_grad;                                // prb.py:162
                            mat_grad = mat_grad + ggx_grad_from_basis(gb, cot);
                        }
                        beta = beta * (bsdf / pdf_bsdf);                                                // prb.py:76
                        if (depth >= a.rr_depth) {                                                      // prb.py:78-87
                            const float l = luminance(beta);
                            if (l == 0.0f) {
                                end = true;
                            } else {
                                const float q = fmaxf(l, 0.05f);
                                const float r = sampler_next(a.sc, sampler);
                                if (r >= q) end = true;
                                else beta = beta / q;
                            }
                        }
                    }
                    if (ADJ) {                                                          // prb.py:178-187
                        if ((mat_grad.x != 0.0f || mat_grad.y != 0.0f || mat_grad.z != 0.0f || mat_grad.w != 0.0f) && grad_ok(mat_grad)) {
                            scatters++;
                            reds += a.aggregate_scatter ? write_bsdf_grad_aggregated(a.tex, itx.uv, mat_grad, gcache, mset) : write_bsdf_grad(a.tex, itx.uv, mat_grad, mset);
                        }
                    }
                    ++depth;
                    if (depth >= a.max_depth) end = true;                                               // prb.py:23
                    if (end) {
                        dest = pw_finish<MODE>(a, pool, slot, id, ADJ ? 1 : 0, L);
                    } else {
                        dest = 1;
                        pool.ro[slot] = make_float4(ray_o.x, ray_o.y, ray_o.z, pdf_bsdf);
                        pool.rd[slot] = make_float4(ray_nd.x, ray_nd.y, ray_nd.z, __uint_as_float(wf_pack(depth, ADJ ? 1 : 0, sampler.dimension)));
                        pool.bi[slot] = make_float4(beta.x, beta.y, beta.z, B4.w);
                        pool.Ls[slot] = make_float4(L.x, L.y, L.z, 0.f);
                    }
                }
                pw_append(l0n, n0n, dest == 1, slot);
                pw_append(fr, ctr + PW_NF, dest == 2 && !drained, slot);
                if (MODE) pw_append(fr, ctr + PW_NRS, dest == 3, slot, true, N - 1);
            }
            if (MODE) { pw_tally(ctr, C_SCATTER, scatters); pw_tally(ctr, C_RED, reds); }
        }
        cur ^= 1;
        __syncthreads();
        PW_MARK(4);
    }
    __syncthreads();
    if (w.phase_cycles && tid == 0) {          // [0..4] claim+regen, extend, shade A, shadow, shade B; [5] rounds; [6] live paths summed over rounds; [7] total
        for (int k = 0; k < 5; ++k) atomicAdd(w.phase_cycles + k, (unsigned long long)t_ph[k]);
        atomicAdd(w.phase_cycles + 5, (unsigned long long)n_rounds);
        atomicAdd(w.phase_cycles + 6, (unsigned long long)n_paths);
        atomicAdd(w.phase_cycles + 7, (unsigned long long)(clock64() - t_start));
    }
    if (MODE && a.aggregate_scatter && a.block_scatter) {
        pw_tally(ctr, C_RED, grad_cache_flush(gcache, a.tex.dmat));
        __syncthreads();
    }
    if (tid < C_NUM) {
        const unsigned v = reinterpret_cast<unsigned*>(ctr)[PW_TALLY0 + tid];
        if (v) atomicAdd(a.counters + tid, (unsigned long long)v);
    }
}

// integrator.py:15-29: s += clamp(radiance) for it ascending; image = s / spp.  The running sum lives in the image buffer
// between batches.
__global__ void __launch_bounds__(256) k_pw_accumulate(const __grid_constant__ RenderArgs a) {
    const PwState& w = a.pw;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < w.P; pix += gridDim.x * blockDim.x) {
        int x, y;
        pw_pixel(a, pix, x, y);
        float4* px = a.image + ((size_t)x + (size_t)y * a.W);
        float3 s = f3(0.f, 0.f, 0.f);
        if (!w.first) { const float4 v = *px; s = f3(v.x, v.y, v.z); }
        for (int j = 0; j < w.S; ++j) {
            const float4 r = __ldcs(w.R + ((size_t)j * w.P + pix));
            s = s + f3(r.x, r.y, r.z);
        }
        *px = w.last ? make_float4(s.x / (float)a.spp, s.y / (float)a.spp, s.z / (float)a.spp, 1.0f) : make_float4(s.x, s.y, s.z, 0.f);
    }
}

template <int MODE, bool STATS>
cudaError_t launch_k_pw(const RenderArgs& a, const PwPlan& plan, cudaStream_t s) {
    const bool env = a.sc.env_count > 0;
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
        if (e != cudaSuccess) return e;
        kernel<<<plan.grid, kPwThreads, plan.smem_bytes, s>>>(a);
        return cudaGetLastError();
    };
    if (plan.small) return go(k_pw<MODE, STATS, false, true>);
    if (env) return go(k_pw<MODE, STATS, true, false>);
    return go(k_pw<MODE, STATS, false, false>);
}

}  // namespace

#ifndef ZDR_FAST_MATH      // exists once (exact build): depends on constants only
// Pool size, traversal-stack placement and shared-memory footprint of the persistent kernel for a scene.
cudaError_t pw_plan(const SceneView& sc, int bvh_depth, int num_sms, PwPlan* out) {
    int dev = 0, max_smem = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    PwPlan p = {};
    p.threads = kPwThreads;
    p.grid = num_sms * kPwBlocks;
    max_smem = std::min(max_smem, (228 * 1024) / kPwBlocks - 1024);      // 228 KB per SM, 1 KB per resident CTA is the system's
    const size_t scene_bytes = 64 * (size_t)sc.num_nodes + (48 + 64) * (size_t)sc.num_tris;
    static const bool no_small = getenv("ZDR_PW_NO_SMEM_SCENE") != nullptr;
    p.small = !no_small && sc.env_count == 0 && scene_bytes <= 32 * 1024 && bvh_depth + 2 <= 20;
    p.stack_depth = p.small ? bvh_depth + 2 : 0;
    const size_t fixed = kPwHeader + (p.small ? (size_t)p.stack_depth * kPwThreads * sizeof(int) + scene_bytes : 0);
    static const int pool_env = getenv("ZDR_PW_POOL") ? atoi(getenv("ZDR_PW_POOL")) : 0;
    int pool = pool_env > 0 ? pool_env : (kPwGlobalState ? 4096 : (p.small ? 1536 : 1024));
    const int fit = (int)(((size_t)max_smem - fixed) / kPwBytesPerSlot);
    pool = std::min(pool, fit) & ~31;
    if (pool < 64 || pool > 65535) return cudaErrorInvalidConfiguration;
    p.pool = pool;
    p.smem_bytes = fixed + (size_t)pool * kPwBytesPerSlot;
    p.gstate_bytes = kPwGlobalState ? (size_t)p.grid * pool * kPwStateBytes : 0;
    *out = p;
    return cudaSuccess;
}
#endif

// Host side: batches of S consecutive samples of every pixel of the shard; per batch ONE persistent launch (+ the ordered
// accumulation of the forward pass).  The adjoint pass needs no buffer and runs as one batch.
cudaError_t launch_path_persistent(bool backward, RenderArgs a, bool stats, const PwPlan& plan, size_t r_capacity_slots,
                                   unsigned* counters_dev, int max_batches, cudaStream_t s) {
    const int tiles = (a.row1 - a.row0 + 7) / 8;
    int rows = 0;
    for (int t = a.shard_index; t < tiles; t += a.shard_count) rows += std::min(8, a.row1 - (a.row0 + 8 * t));
    if (rows <= 0 || a.it1 <= a.it0) return cudaSuccess;
    PwState& w = a.pw;
    w.P = a.W * rows;
    w.div_P = make_fastdiv(w.P);
    w.div_W = make_fastdiv(a.W);
    w.pool = plan.pool;
    w.stack_depth = plan.stack_depth;
    static const bool no_bulk = getenv("ZDR_PW_NO_BULK_COPY") != nullptr;
    w.bulk_copy = no_bulk ? 0 : 1;
    static const int refill_c = getenv("ZDR_PW_REFILL") ? atoi(getenv("ZDR_PW_REFILL")) : 20;
    static const int refill_a = getenv("ZDR_PW_REFILL_ANY") ? atoi(getenv("ZDR_PW_REFILL_ANY")) : 20;
    w.refill_closest = refill_c;
    w.refill_any = refill_a;
    static const bool profile = getenv("ZDR_PW_PROFILE") != nullptr;     // phase cycles -> 8 x u64 behind the 64 batch counters (zdr_debug_phase_cycles)
    w.phase_cycles = profile ? reinterpret_cast<unsigned long long*>(counters_dev + 64) : nullptr;
    const long long its = a.it1 - a.it0;
    long long smax = std::min<long long>(its, ((1ll << 31) - 1) / w.P);
    if (!backward) smax = std::min<long long>(smax, (long long)(r_capacity_slots / (size_t)w.P));
    if (smax < 1) return cudaErrorInvalidValue;
    const int batches = (int)((its + smax - 1) / smax);
    if (batches > max_batches) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(counters_dev, 0, sizeof(unsigned) * 64 + sizeof(unsigned long long) * 8, s);
    if (e != cudaSuccess) return e;
    int b = 0;
    for (long long base = a.it0; base < a.it1; base += smax, ++b) {
        w.S = (int)std::min<long long>(smax, a.it1 - base);
        w.it_base = (int)base;
        w.total = (unsigned)((long long)w.P * w.S);
        w.next = counters_dev + b;
        w.first = base == a.it0;
        w.last = base + smax >= a.it1;
        ++g_kernel_launches;
        if (backward) e = stats ? launch_k_pw<1, true>(a, plan, s) : launch_k_pw<1, false>(a, plan, s);
        else e = stats ? launch_k_pw<0, true>(a, plan, s) : launch_k_pw<0, false>(a, plan, s);
        if (e != cudaSuccess) return e;
        if (!backward && a.image) {
            ++g_kernel_launches;
            const int grid = std::min((w.P + 255) / 256, plan.grid * 8);
            k_pw_accumulate<<<grid, 256, 0, s>>>(a);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

}  // namespace zdr
