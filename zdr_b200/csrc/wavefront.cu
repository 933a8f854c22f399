// wavefront.cu — the HBM wavefront schedule of the path integrator (prb.py:18-88) and its Path-Replay-Backprop adjoint
// (prb.py:91-187): stage kernels over compacted lists of live paths whose state lives in global memory (render.cuh WfState).
// Its own translation unit because its kernels want other inlining than the per-pixel kernels of render.cu: here the IEEE
// sqrt + divide of normalize() and the BRDF evaluation are inline (stage B: no calls, no register saves around them:
// C3 +1.6 %, profiles/r03a-b), there the out-of-line copies are faster (direct / collocated / one-kernel path: 2-3 %).
#define ZDR_INLINE_RCP_LENGTH 1
#define ZDR_INLINE_GGX_EVAL 1
#include <algorithm>
#include <type_traits>
#include "render_ctx.cuh"

namespace zdr {

namespace {

// ------------------------------------------------------------------------------ wavefront path tracer
// The same path integrator / PRB adjoint as k_path, cut into small kernels that run over compacted
// lists of live paths whose state lives in HBM (render.cuh: WfState).  Per batch of S samples x P pixels:
//   k_wf_generate          camera samples -> path state, list[0] = all slots
//   per bounce (max_depth rounds):
//     k_wf_trace<closest>  closest-hit traversal of list[0]                      (ray -> hit record)
//     k_wf_shade_a         miss / backface / emitter tests, light sample + shadow-ray record; list[0] -> list[1]
//     k_wf_trace<any>      occlusion query of list[1]
//     k_wf_shade_b         NEE term, BSDF sample, throughput, Russian roulette, (adjoint walk) BRDF adjoints +
//                          gradient scatter; list[1] -> list[0] with one warp-aggregated atomic per warp
//   forward:  k_wf_accumulate adds the batch's finished samples per pixel IN SAMPLE ORDER (integrator.py:15-28)
//   adjoint:  the rounds run twice — MODE 1 replays every sample to get Le (prb.py:97-99), k_wf_restart puts
//             the samples back at the camera with the original sampler state, MODE 2 is the adjoint walk.
// Why: in the one-kernel form only ~15 of 32 lanes are active (finished lanes idle through shading,
// traversal trip counts diverge), 128 registers cap occupancy at 16 warps/SM and the code overflows the
// 32 KB instruction-cache level (profiles/r01d).  Here every warp of a shade stage starts with 32 live
// paths, each kernel is small (64 registers), and the state traffic (~1.5 kB of DRAM per path) is what the
// 7.7 TB/s of HBM3e are for.  Arithmetic per sample is unchanged (same device functions), so images stay
// bit-identical to the oracle.
#ifndef ZDR_WF_EARLY_SHADOW_LOADS      // stage B: load the light-sample records right after the list entry, with the rest of the
#define ZDR_WF_EARLY_SHADOW_LOADS 1      // slot's state (measured r02y: C3 at 64 spp 7067 -> 7103 Mrays/s, reproducible; 256 spp: +0.1 %)
#endif
#ifndef ZDR_WF_EARLY_TEXELS      // stage B: texel loads issued before the shading frame is built.  Measured r02y: C3 7115 -> 6730 Mrays/s
#define ZDR_WF_EARLY_TEXELS 0      // (16 more live registers at the 64-register cap: spills); off
#endif
#ifndef ZDR_WF_SHADE_A_MINB
#define ZDR_WF_SHADE_A_MINB 8
#endif
#ifndef ZDR_WF_SHADE_B_MINB
#define ZDR_WF_SHADE_B_MINB 8
#endif
#ifndef ZDR_WF_SHADE_BB_MINB      // adjoint variant of stage B (BRDF adjoints + scatter)
#define ZDR_WF_SHADE_BB_MINB 8
#endif
#ifndef ZDR_WF_EXTEND_MINB
#define ZDR_WF_EXTEND_MINB 8
#endif
#ifndef ZDR_WF_GRID_MULT
#define ZDR_WF_GRID_MULT 64
#endif
// Blocks per SM of the persistent traversal kernels.  They fetch rays dynamically, so one resident wave (16 blocks of
// 128 threads at 64 registers) is enough; the 64-per-SM grid of the grid-stride stages only added launch / drain time
// (sweep r01u: C3 adjoint pass -2 %, one rank's share of an 8-GPU run -4 % forward / -7 % adjoint, C5 -2 %).
#ifndef ZDR_WF_TRACE_GRID_MULT
#define ZDR_WF_TRACE_GRID_MULT 16
#endif
constexpr int kWfThreads = 128;
ZDR_DEV int wf_stride() { return gridDim.x * blockDim.x; }

template <bool BWD>
__global__ void __launch_bounds__(kWfThreads) k_wf_generate(const __grid_constant__ RenderArgs a) {
    const WfState& w = a.wf;
    Tally tl = {};
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < w.nslots; slot += wf_stride()) {
        const int j = fastdiv(slot, w.div_P), pix = slot - j * w.P;
        int x, y;
        wf_pixel(a, pix, x, y);
        Sampler sampler = make_sampler(a.sc, x, y, a.seed, a.spp, w.it_base + j);
        const Ray ray = camera_sample(a, x, y, sampler);
        tl.samples++;
        w.AB[2 * (size_t)(slot)] = make_float4(ray.o.x, ray.o.y, ray.o.z, 1e30f);                              // prb.py:22
        w.AB[2 * (size_t)(slot) + 1] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(wf_pack(0, 0, sampler.dimension)));
        w.CD[2 * (size_t)(slot)] = make_float4(1.f, 1.f, 1.f, 0.f);
        w.CD[2 * (size_t)(slot) + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!BWD) w.R[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
        w.list[0][slot] = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { w.counts[0] = a.max_depth > 0 ? w.nslots : 0; w.counts[1] = 0; w.counts[2] = 0; w.counts[3] = 0; }
    if (w.bin_mask) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * kWfBins; i += wf_stride()) w.hist[0][i] = 0;   // hist[0], hist[1], cur
    flush_tally(a, tl);
}

// Extend (ANY = false: closest hit of list[0] -> H) and shadow (ANY = true: occlusion of list[1] ->
// SC.w) stages: persistent traversal with dynamic ray fetch (Aila & Laine 2009).  Rays of one warp
// need very different numbers of node / triangle steps (secondary rays: 14 of 32 lanes active in
// the one-ray-per-lane form, profiles/r01h); here a lane that finishes its ray takes the next one
// from the stage's list, and the lanes still traversing leave the inner loop to let the idle ones
// refill once fewer than kWfRefill of them are left.
#ifndef ZDR_WF_REFILL
#define ZDR_WF_REFILL 20
#endif
constexpr int kWfRefill = ZDR_WF_REFILL;
#ifndef ZDR_WF_FETCH_CHUNK
#define ZDR_WF_FETCH_CHUNK 128
#endif
// Staged ray fetch of the traversal kernels: a warp claims list entries a chunk AHEAD (the cursor atomic's round trip is off
// the critical path), reads the chunk's list entries once, and copies the chunk's ray records from HBM into its own
// shared-memory staging buffer with cp.async while it is still traversing the previous chunk; refills then take rays from
// shared memory.  profiles/r01w (ncu source page of k_wf_trace): 24 % of the warps' time went to the refill path — cursor
// atomic -> list entry -> ray record, three dependent trips to L2 / HBM per refill.
// Measured (profiles/r02l): correct, but C3 7081 -> 6718 Mrays/s and the C5-shaped scene 1587 -> 1492: the 18 KB of staging per
// block cost more L1 and issue slots than the hidden latency was worth (the other resident warps already covered it).  Off.
#ifndef ZDR_WF_STAGE_RAYS
#define ZDR_WF_STAGE_RAYS 0
#endif
#ifndef ZDR_WF_STAGE_CHUNK
#define ZDR_WF_STAGE_CHUNK 64
#endif
#ifndef ZDR_WF_TRACE_PREFETCH      // measured r02h: 6917 (on) vs 6936 (off) Mrays/s on C3 at 64 spp — no gain, off
#define ZDR_WF_TRACE_PREFETCH 0
#endif
#ifndef ZDR_WF_REFILL_ANY_SMALL      // shadow rays of small scenes also refill idle lanes (cheap once the records are prefetched)
#define ZDR_WF_REFILL_ANY_SMALL 0
#endif
// the stage that consumed a binned list leaves the list's bin counters and the cursors zero for their next use
ZDR_DEV void wf_bins_reset(const WfState& w, int which) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kWfBins; i += wf_stride()) { w.hist[which][i] = 0; w.cur[i] = 0; }
}
constexpr int kWfNodeBudget = 12;
template <bool ANY, bool STATS, bool BIG>      // BIG: scene with more than kWfBigSceneNodes nodes (launch_wf_round)
__global__ void __launch_bounds__(kWfThreads, ZDR_WF_EXTEND_MINB) k_wf_trace(const __grid_constant__ RenderArgs a) {
    const WfState& w = a.wf;
    const int n = *(volatile int*)(w.counts + (ANY ? 1 : 0));
    int* cursor = w.counts + (ANY ? 3 : 2);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (ANY) { w.counts[0] = 0; w.counts[2] = 0; }       // shade B appends to list[0]; cursor of the next extend
        else { w.counts[1] = 0; w.counts[3] = 0; }           // shade A appends to list[1]; cursor of the next shadow stage
    }
    const int* __restrict__ list = w.bin_in ? w.blist : w.list[ANY ? 1 : 0];
    if (w.bin_in) wf_bins_reset(w, ANY ? 1 : 0);
    const float4* __restrict__ nodes = a.sc.nodes;
    const float4* __restrict__ tris = a.sc.tris;
    Tally tl = {};
    const int lane = threadIdx.x & 31;
    constexpr bool refill_any = BIG || ZDR_WF_REFILL_ANY_SMALL;      // traversals long enough for refills to pay on every kind of ray
    // per-lane traversal state
    bool has_ray = false, exhausted = false;
    int slot = 0, link = kLinkSentinel, sp = 0, hit_slot = -1, best_gid = 0x7fffffff;
    float tmin = 0.f, tmax = 0.f, best = 0.f, hit_u = 0.f, hit_v = 0.f, hit_t = 0.f;
    float3 id = f3(0.f, 0.f, 0.f), oid = id;
    RayPre rp = {};
    int stack[kStackSize];
#if ZDR_WF_STAGE_RAYS
    constexpr int CH = BIG ? ZDR_WF_STAGE_CHUNK / 2 : ZDR_WF_STAGE_CHUNK;      // large scenes keep more of the SM's memory as L1 for the tree
    __shared__ float4 s_ray[kWfThreads / 32][2][CH][2];      // per warp: two chunks of ray records (current / being staged)
    __shared__ int s_slot[kWfThreads / 32][2][CH];
    const int wid = threadIdx.x >> 5;
    const int total_warps = gridDim.x * (kWfThreads / 32);
    const int chunk = max(32, min(CH, (n / (4 * total_warps)) & ~31));     // shrinks with the list so that short lists spread over all warps
    int cb = 1, c_n = 0, c_pos = 0, n_n = 0;              // warp-uniform: current buffer, its entries, next unread entry; entries of the staged chunk
    int pend = 0;                                        // lane 0: base of the chunk claimed ahead (atomicAdd issued, result not yet consumed)
    bool claims_done = false;
    // stage the chunk whose base lane 0 holds in `pend` into buffer `b`, then claim the chunk after it
    auto stage_next = [&](int b) -> int {
        if (claims_done) return 0;
        const int base = __shfl_sync(0xffffffffu, pend, 0);
        const int cnt = min(max(n - base, 0), chunk);
        if (cnt <= 0) { claims_done = true; return 0; }
        for (int k = lane; k < cnt; k += 32) {
            const int sl = list[base + k];
            const float4* rec = ANY ? w.SAB + 2 * (size_t)sl : w.AB + 2 * (size_t)sl;
            s_slot[wid][b][k] = sl;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_ray[wid][b][k][0]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rec) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(rec + 1) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (lane == 0) pend = atomicAdd(cursor, chunk);
        return cnt;
    };
    if (lane == 0) pend = atomicAdd(cursor, chunk);
    n_n = stage_next(0);                                 // chunk 0 -> buffer 0 (the "staged" one while cb == 1)
#elif ZDR_WF_FETCH_CHUNK
    // A warp takes list entries from the shared cursor a chunk at a time and hands them to its lanes over several
    // refills: every refill hitting ONE address with an atomic is what a same-address atomic unit serialises
    // (tools/microbench: ~1-2 G/s).  The chunk shrinks with the list so that short lists still spread over all warps.
    const int total_warps = gridDim.x * (kWfThreads / 32);
    const int chunk = max(32, min(ZDR_WF_FETCH_CHUNK, (n / (4 * total_warps)) & ~31));
    int chunk_next = 0, chunk_end = 0;                  // warp-uniform
#endif
    for (;;) {
        // ---- refill idle lanes
        const unsigned idle = __ballot_sync(0xffffffffu, !has_ray && !exhausted);
        if (idle) {
            const bool mine = !has_ray && !exhausted;
            const int rank = __popc(idle & ((1u << lane) - 1u));
            float4 RA = make_float4(0.f, 0.f, 0.f, 0.f), RB = RA;      // the lane's new ray record
            bool got = false;
#if ZDR_WF_STAGE_RAYS
            const int need = __popc(idle), have = c_n - c_pos;
            if (mine && rank < have) {                   // from the current chunk
                const int e = c_pos + rank;
                RA = s_ray[wid][cb][e][0]; RB = s_ray[wid][cb][e][1]; slot = s_slot[wid][cb][e];
                got = true;
            }
            if (have < need) {                           // switch to the staged chunk, stage the one after it into the freed buffer
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                cb ^= 1; c_n = n_n; c_pos = min(need - have, c_n);
                if (mine && rank >= have) {
                    const int e = rank - have;
                    if (e < c_n) { RA = s_ray[wid][cb][e][0]; RB = s_ray[wid][cb][e][1]; slot = s_slot[wid][cb][e]; got = true; }
                }
                __syncwarp();
                n_n = stage_next(cb ^ 1);
            } else {
                c_pos += need;
            }
#else
#if ZDR_WF_FETCH_CHUNK
            const int need = __popc(idle), have = chunk_end - chunk_next;
            int base = 0;
            if (have < need) {
                if (lane == 0) base = atomicAdd(cursor, chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
            }
            const int i_next = rank < have ? chunk_next + rank : base + (rank - have);
            if (have < need) {
#if ZDR_WF_TRACE_PREFETCH
                // ask L2 for the chunk's ray records at claim time (measured r02h: no gain; off)
                for (int k = lane; k < chunk; k += 32) {
                    const int j = base + k;
                    if (j < n) {
                        const int sl = list[j];
                        const float4* rec = ANY ? w.SAB + 2 * (size_t)sl : w.AB + 2 * (size_t)sl;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(rec));
                    }
                }
#endif
                chunk_next = base + (need - have); chunk_end = base + chunk;
            } else chunk_next += need;
#else
            int base = 0;
            if (lane == __ffs(idle) - 1) base = atomicAdd(cursor, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, __ffs(idle) - 1);
            const int i_next = base + rank;
#endif
            if (mine && i_next < n) {
                slot = list[i_next];
                const float4* rec = ANY ? w.SAB + 2 * (size_t)slot : w.AB + 2 * (size_t)slot;
                RA = rec[0]; RB = rec[1];
                got = true;
            }
#endif
            if (mine) {
                if (!got) {
                    exhausted = true;
                } else {
                    const float3 o = f3(RA.x, RA.y, RA.z), d = f3(RB.x, RB.y, RB.z);
                    if (ANY) { tmin = 1e-4f; tmax = RA.w; tl.any++; }           // SAB: origin | light.dist, light.wi | light.pdf
                    else { tmin = 0.0f; tmax = 1e30f; tl.closest++; }           // AB: origin | pdf_bsdf, direction | packed state
                    rp = ray_pre(o, d);
                    id = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    if (!(fabsf(id.x) <= 1e30f)) id.x = copysignf(1e30f, d.x);
                    if (!(fabsf(id.y) <= 1e30f)) id.y = copysignf(1e30f, d.y);
                    if (!(fabsf(id.z) <= 1e30f)) id.z = copysignf(1e30f, d.z);
                    oid = f3(o.x * id.x, o.y * id.y, o.z * id.z);
                    best = tmax; best_gid = 0x7fffffff; hit_slot = -1; hit_u = hit_v = hit_t = 0.0f;
                    sp = 0; stack[sp++] = kLinkSentinel; link = a.sc.root_link;
                    has_ray = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, !has_ray)) {
            if (__all_sync(0xffffffffu, exhausted)) break;
            continue;                                   // only pass-through entries were fetched
        }
        // ---- traverse until done, or until too few lanes are left working
        while (has_ray) {
            int budget = kWfNodeBudget;      // BIG: bounded, so that the refill test below is reached regularly even when
                                             // a ray walks hundreds of nodes between two leaves (9.6 M-triangle scene:
                                             // 9.6 of 32 lanes were active without the bound, profiles/r01p)
            while (link >= 0 && (!BIG || budget-- > 0)) {    // inner nodes
                const float4* nd = nodes + 4 * (size_t)link;
                float4 n0, n1, n2, n3;
                ld_node(nd, n0, n1, n2, n3);
                if (STATS) tl.ts.nodes++;
                float t0, t1;
                const bool h0 = box_test(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, id, oid, tmin, best, t0);
                const bool h1 = box_test(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, id, oid, tmin, best, t1);
                const int l0 = __float_as_int(n3.x), l1 = __float_as_int(n3.y);
                if (h0 && h1) {
                    const bool swap = t1 < t0;
                    stack[sp++] = swap ? l0 : l1;
                    link = swap ? l1 : l0;
                } else if (h0 || h1) {
                    link = h0 ? l0 : l1;
                } else {
                    link = stack[--sp];
                }
            }
            bool done = link == kLinkSentinel;
            if (!done && link < 0) {                     // leaf
                const int code = ~link;
                const int first = code >> 2, count = (code & 3) + 1;
                for (int k = 0; k < count; ++k) {
                    const float* tp = reinterpret_cast<const float*>(tris + 3 * (size_t)(first + k));
                    if (STATS) tl.ts.tris++;
                    float t, u, v;
                    if (tri_test(rp, tp, tmin, t, u, v)) {
                        if (ANY) {
                            if (t < tmax) { hit_slot = first + k; done = true; break; }
                            continue;
                        }
                        const int gid = __float_as_int(__ldg(tp + 3));
                        if (t < best || (hit_slot >= 0 && t == best && gid < best_gid)) {
                            best = t; best_gid = gid;
                            hit_t = t; hit_u = u; hit_v = v; hit_slot = first + k;
                        }
                    }
                }
                if (!done) { link = stack[--sp]; done = link == kLinkSentinel; }
            }
            if (done && (link == kLinkSentinel || (ANY && hit_slot >= 0))) {
                if (ANY) w.SC[slot].w = hit_slot >= 0 ? 1.0f : 0.0f;
                else w.H[slot] = make_float4(__int_as_float(hit_slot), hit_u, hit_v, hit_t);
                has_ray = false;
                break;
            }
            // shadow rays in a small scene are too short for refills to pay (r01i)
            if ((!ANY || refill_any) && __popc(__activemask()) < kWfRefill) break;
        }
    }
    flush_tally(a, tl);
}

// The traversal stages for scenes of a few kilobytes (the Cornell box: 35 nodes + 36 triangles = 4 KB): every block keeps the
// tree and the triangles in SHARED memory (LDS with 32-bit addresses instead of L1-cached global loads) and every lane its
// traversal stack in NREG registers (traverse.cuh RegStack: links fit 16 bits, the tree is at most 2 * NREG levels deep) — the
// pop -> node load -> box test chain no longer goes through local memory and L1.  Same fetch policy and the same arithmetic as
// k_wf_trace (tri_test_vals, box_test): bit-identical hits.
#ifndef ZDR_WF_TINY
#define ZDR_WF_TINY 1
#endif
#ifndef ZDR_WF_TINY_LOCAL_STACK      // measured (profiles/r02q): the PRMT register stack costs +3-6 % warp instructions (its predicated push / pop
#define ZDR_WF_TINY_LOCAL_STACK 1      // sequences issue every node step) and is 0.6 % slower than the local-memory stack on C3: local stack shipped
#endif
template <bool ANY, bool STATS, int NREG>
__global__ void __launch_bounds__(kWfThreads, ZDR_WF_EXTEND_MINB) k_wf_trace_tiny(const __grid_constant__ RenderArgs a) {
    extern __shared__ __align__(16) float4 s_scene[];
    const WfState& w = a.wf;
    const int n = *(volatile int*)(w.counts + (ANY ? 1 : 0));
    int* cursor = w.counts + (ANY ? 3 : 2);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (ANY) { w.counts[0] = 0; w.counts[2] = 0; }
        else { w.counts[1] = 0; w.counts[3] = 0; }
    }
    const int* __restrict__ list = w.bin_in ? w.blist : w.list[ANY ? 1 : 0];
    if (w.bin_in) wf_bins_reset(w, ANY ? 1 : 0);
    const int n_node4 = 4 * a.sc.num_nodes, n_tri4 = 3 * a.sc.num_tris;
    for (int i = threadIdx.x; i < n_node4; i += kWfThreads) s_scene[i] = __ldg(a.sc.nodes + i);
    for (int i = threadIdx.x; i < n_tri4; i += kWfThreads) s_scene[n_node4 + i] = __ldg(a.sc.tris + i);
    __syncthreads();
    unsigned s_nodes = (unsigned)__cvta_generic_to_shared(s_scene), s_tris = s_nodes + 16u * (unsigned)n_node4;
    asm volatile("" : "+r"(s_nodes), "+r"(s_tris));      // keep both bases in registers (ptxas otherwise recomputes the shared window base,
                                                         // an S2R, inside the node loop)
    Tally tl = {};
    const int lane = threadIdx.x & 31;
    bool has_ray = false, exhausted = false;
    int slot = 0, link = kShortSentinel, hit_slot = -1, best_gid = 0x7fffffff;
    float tmin = 0.f, tmax = 0.f, best = 0.f, hit_u = 0.f, hit_v = 0.f, hit_t = 0.f;
    float3 id = f3(0.f, 0.f, 0.f), oid = id;
    RayPre rp = {};
#if ZDR_WF_TINY_LOCAL_STACK
    int stack_entries[16];
    LocalStack16 stack{stack_entries, 0};
#else
    RegStack<NREG> stack;
#endif
    stack.clear();
    const int total_warps = gridDim.x * (kWfThreads / 32);
    const int chunk = max(32, min(ZDR_WF_FETCH_CHUNK, (n / (4 * total_warps)) & ~31));
    int chunk_next = 0, chunk_end = 0;                  // warp-uniform
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !has_ray && !exhausted);
        if (idle) {
            const bool mine = !has_ray && !exhausted;
            const int rank = __popc(idle & ((1u << lane) - 1u));
            const int need = __popc(idle), have = chunk_end - chunk_next;
            int base = 0;
            if (have < need) {
                if (lane == 0) base = atomicAdd(cursor, chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
            }
            const int i_next = rank < have ? chunk_next + rank : base + (rank - have);
            if (have < need) { chunk_next = base + (need - have); chunk_end = base + chunk; }
            else chunk_next += need;
            if (mine) {
                if (i_next >= n) {
                    exhausted = true;
                } else {
                    slot = list[i_next];
                    const float4* rec = ANY ? w.SAB + 2 * (size_t)slot : w.AB + 2 * (size_t)slot;
                    const float4 RA = rec[0], RB = rec[1];
                    const float3 o = f3(RA.x, RA.y, RA.z), d = f3(RB.x, RB.y, RB.z);
                    if (ANY) { tmin = 1e-4f; tmax = RA.w; tl.any++; }
                    else { tmin = 0.0f; tmax = 1e30f; tl.closest++; }
                    rp = ray_pre(o, d);
                    id = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    if (!(fabsf(id.x) <= 1e30f)) id.x = copysignf(1e30f, d.x);
                    if (!(fabsf(id.y) <= 1e30f)) id.y = copysignf(1e30f, d.y);
                    if (!(fabsf(id.z) <= 1e30f)) id.z = copysignf(1e30f, d.z);
                    oid = f3(o.x * id.x, o.y * id.y, o.z * id.z);
                    best = tmax; best_gid = 0x7fffffff; hit_slot = -1; hit_u = hit_v = hit_t = 0.0f;
                    stack.clear(); link = a.sc.root_link;
                    has_ray = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, !has_ray)) {
            if (__all_sync(0xffffffffu, exhausted)) break;
            continue;
        }
        while (has_ray) {
            while (link >= 0) {                                  // inner nodes
                const unsigned nd = s_nodes + 64u * (unsigned)link;
                const float4 n0 = lds_f4(nd), n1 = lds_f4(nd + 16u), n2 = lds_f4(nd + 32u);
                const float2 n3 = lds_f2(nd + 48u);
                if (STATS) tl.ts.nodes++;
                float t0, t1;
                const bool h0 = box_test(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, id, oid, tmin, best, t0);
                const bool h1 = box_test(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, id, oid, tmin, best, t1);
                const int l0 = __float_as_int(n3.x), l1 = __float_as_int(n3.y);
                if (h0 && h1) {
                    const bool swap = t1 < t0;
                    stack.push(swap ? l0 : l1);
                    link = swap ? l1 : l0;
                } else if (h0 || h1) {
                    link = h0 ? l0 : l1;
                } else {
                    link = stack.pop();
                }
            }
            bool done = link == kShortSentinel;
            if (!done) {                                         // leaf
                const int code = ~link;
                const int first = code >> 2, count = (code & 3) + 1;
                for (int k = 0; k < count; ++k) {
                    const unsigned tp = s_tris + 48u * (unsigned)(first + k);
                    const unsigned ax = tp + 4u * (unsigned)rp.kx, ay = tp + 4u * (unsigned)rp.ky, az = tp + 4u * (unsigned)rp.kz;
                    if (STATS) tl.ts.tris++;
                    float t, u, v;
                    if (tri_test_vals(rp, lds_f(ax), lds_f(ay), lds_f(az), lds_f(ax + 16u), lds_f(ay + 16u), lds_f(az + 16u), lds_f(ax + 32u),
                                      lds_f(ay + 32u), lds_f(az + 32u), tmin, t, u, v)) {
                        if (ANY) {
                            if (t < tmax) { hit_slot = first + k; done = true; break; }
                            continue;
                        }
                        const int gid = __float_as_int(lds_f(tp + 12u));
                        if (t < best || (hit_slot >= 0 && t == best && gid < best_gid)) {
                            best = t; best_gid = gid;
                            hit_t = t; hit_u = u; hit_v = v; hit_slot = first + k;
                        }
                    }
                }
                if (!done) { link = stack.pop(); done = link == kShortSentinel; }
            }
            if (done) {
                if (ANY) w.SC[slot].w = hit_slot >= 0 ? 1.0f : 0.0f;
                else w.H[slot] = make_float4(__int_as_float(hit_slot), hit_u, hit_v, hit_t);
                has_ray = false;
                break;
            }
            if (!ANY && __popc(__activemask()) < kWfRefill) break;
        }
    }
    flush_tally(a, tl);
}

// survivors of a warp are appended to `list` with one atomic
ZDR_DEV void wf_append(int* __restrict__ list, int* counter, bool keep, int value) {
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}
// Ray binning (WfState): bin of a ray = (cell of its origin in an 8^3 grid over the scene, octant of its direction), or for large
// scenes the cell of its origin in a 16^3 grid.
ZDR_DEV int wf_key(const WfState& w, float3 o, float3 d) {
    const int nc = w.key_cells;
    const int cx = min(max((int)((o.x - w.key_lo.x) * w.key_scale.x), 0), nc - 1);
    const int cy = min(max((int)((o.y - w.key_lo.y) * w.key_scale.y), 0), nc - 1);
    const int cz = min(max((int)((o.z - w.key_lo.z) * w.key_scale.z), 0), nc - 1);
    const int cell = (cx * nc + cy) * nc + cz;
    if (nc == 16) return cell;
    const int oct = (d.x < 0.0f ? 1 : 0) | (d.y < 0.0f ? 2 : 0) | (d.z < 0.0f ? 4 : 0);
    return cell * 8 + oct;
}
// wf_append that also records the bin of every appended entry and counts the bins (one reduction per distinct bin of the warp)
ZDR_DEV void wf_append_binned(const WfState& w, int which, bool keep, int value, int key) {
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(w.counts + which, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned same = __match_any_sync(0xffffffffu, keep ? key : -1 - lane);
    if (keep) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        w.list[which][pos] = value;
        w.keys[which][pos] = (unsigned short)key;
        if (lane == __ffs(same) - 1) atomicAdd(w.hist[which] + key, __popc(same));
    }
}

// The binned list of a traversal stage: blist = list[WHICH] grouped by bin (counting sort; the order inside a bin is whatever
// the atomics give — the image is summed per pixel in sample order by k_wf_accumulate, so no result depends on the order rays
// are traced in).  Every block owns one contiguous tile of the list: it counts the tile's keys in shared memory, reserves the
// tile's share of every bin it touches with ONE global atomic per bin (against the global prefix of the stage's histogram,
// which each thread holds for its own 16 bins), and places the entries with shared-memory atomics.  (The first version placed
// every entry with a global atomic on its bin's cursor: a Cornell box fills a few hundred bins and the atomics of a hot bin
// serialise — 200 us for 3.3 M entries, profiles/r02r; this one is a streaming pass.)
constexpr int kWfBinThreads = 256;
template <int WHICH>
__global__ void __launch_bounds__(kWfBinThreads) k_wf_bin(const __grid_constant__ RenderArgs a) {
    const WfState& w = a.wf;
    __shared__ int s_pos[kWfBins];      // count of the tile's entries per bin, then the next free position of the bin's reserved range
    __shared__ int s_part[kWfBinThreads];
    const int n = *(volatile int*)(w.counts + WHICH);
    constexpr int per = kWfBins / kWfBinThreads;
    const int tid = threadIdx.x;
    // global exclusive prefix of this thread's bins [tid * per, tid * per + per)
    int loc[per], sum = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) { loc[k] = sum; sum += w.hist[WHICH][tid * per + k]; s_pos[tid * per + k] = 0; }
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < kWfBinThreads; o <<= 1) {          // inclusive scan of the per-thread sums
        const int v = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    const int before = s_part[tid] - sum;
    // the tile of this block
    const long long t0 = (long long)n * blockIdx.x / gridDim.x, t1 = (long long)n * (blockIdx.x + 1) / gridDim.x;
    const unsigned short* __restrict__ keys = w.keys[WHICH];
    const int* __restrict__ list = w.list[WHICH];
    for (long long i = t0 + tid; i < t1; i += kWfBinThreads) atomicAdd(&s_pos[keys[i]], 1);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < per; ++k) {
        const int bin = tid * per + k, c = s_pos[bin];
        if (c) s_pos[bin] = before + loc[k] + atomicAdd(w.cur + bin, c);
    }
    __syncthreads();
    for (long long i = t0 + tid; i < t1; i += kWfBinThreads) w.blist[atomicAdd(&s_pos[keys[i]], 1)] = list[i];
}

struct WfPath {        // registers of one path inside a stage
    int slot, x, y, it, depth, phase;
    Ray ray;
    float pdf_bsdf;
    float3 beta, L;
    Sampler sampler;
};
ZDR_DEV WfPath wf_load(const RenderArgs& a, int slot, bool with_throughput) {
    const WfState& w = a.wf;
    WfPath p;
    p.slot = slot;
    const int j = fastdiv(slot, w.div_P), pix = slot - j * w.P;
    wf_pixel(a, pix, p.x, p.y);
    p.it = w.it_base + j;
    const float4 A = w.AB[2 * (size_t)(slot)], B = w.AB[2 * (size_t)(slot) + 1];
    p.ray = make_ray(f3(A.x, A.y, A.z), f3(B.x, B.y, B.z), 0.0f, 1e30f);
    p.pdf_bsdf = A.w;
    const unsigned packed = __float_as_uint(B.w);
    p.depth = (int)(packed & 0xffu);
    p.phase = (packed & kWfPhaseBit) ? 1 : 0;
    p.sampler = make_sampler(a.sc, p.x, p.y, a.seed, a.spp, p.it);
    p.sampler.dimension = (int)(packed >> 12);
    if (with_throughput) {
        const float4 C4 = w.CD[2 * (size_t)(slot)], D4 = w.CD[2 * (size_t)(slot) + 1];
        p.beta = f3(C4.x, C4.y, C4.z); p.L = f3(D4.x, D4.y, D4.z);
    } else {
        p.beta = f3(0.f, 0.f, 0.f); p.L = f3(0.f, 0.f, 0.f);
    }
    return p;
}
ZDR_DEV void wf_load_throughput(const RenderArgs& a, WfPath& p) {
    const float4 C4 = a.wf.CD[2 * (size_t)(p.slot)], D4 = a.wf.CD[2 * (size_t)(p.slot) + 1];
    p.beta = f3(C4.x, C4.y, C4.z); p.L = f3(D4.x, D4.y, D4.z);
}
ZDR_DEV Hit wf_hit(const RenderArgs& a, int slot) {
    const float4 H = a.wf.H[slot];
    Hit hit;
    hit.slot = __float_as_int(H.x); hit.u = H.y; hit.v = H.z; hit.t = H.w; hit.inst = -1; hit.prim = -1;
    if (hit.slot >= 0) {
        const float4* tp = a.sc.tris + 3 * (size_t)hit.slot;
        hit.inst = __float_as_int(__ldg(tp + 1).w);
        hit.prim = __float_as_int(__ldg(tp + 2).w);
    }
    return hit;
}
// End of a path.  MODE 0 (forward): finished sample -> R.  MODE 1 (adjoint pass, replay): Le -> D,
// read back by k_wf_restart (prb.py:97-101).  MODE 2 (adjoint walk): nothing left to do.
template <int MODE>
ZDR_DEV void wf_finish(const RenderArgs& a, WfPath& p) {
    const WfState& w = a.wf;
    if (MODE == 1) { w.CD[2 * (size_t)(p.slot) + 1] = make_float4(p.L.x, p.L.y, p.L.z, 1.0f); return; }
    if (MODE == 2) return;
    if (a.samples_out) {
        float* o = a.samples_out + (((size_t)p.y * a.W + p.x) * a.spp + p.it) * 3;
        o[0] = p.L.x; o[1] = p.L.y; o[2] = p.L.z;
    }
    if (!any_nan(p.L))                                                             // integrator.py:27-28
        w.R[p.slot] = make_float4(clampf(p.L.x, 0.0f, 100000.0f), clampf(p.L.y, 0.0f, 100000.0f), clampf(p.L.z, 0.0f, 100000.0f), 0.f);
}

// Stage A: classify the closest hit (prb.py:26-46); paths that go on get their light sample
// (prb.py:57) and a shadow-ray record.  list[0] -> list[1].
template <int MODE, bool ENV>
__global__ void __launch_bounds__(kWfThreads, ZDR_WF_SHADE_A_MINB) k_wf_shade_a(const __grid_constant__ RenderArgs a) {
    const WfState& w = a.wf;
    const SceneView& sc = a.sc;
    const int n = *(volatile int*)(w.counts + 0);
    const int n_round = (n + 31) & ~31;
    const bool corrected = a.prb_mode == ZDR_PRB_CORRECTED;
    constexpr int phase = MODE == 2 ? 1 : 0;
    Tally tl = {};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += wf_stride()) {
        bool keep = false;
        int entry = 0, key = 0;
        if (i < n) {
            WfPath p = wf_load(a, w.list[0][i], false);       // beta / L are only needed when the path ends here
            const Hit hit = wf_hit(a, p.slot);
            bool end = false;
            if (hit.inst < 0) {
                wf_load_throughput(a, p);
                if (ENV && phase == 0 && sc.env_count > 0) {                     // prb.py:26-31
                    const float3 tex = env_texture(sc, direction_to_uv(p.ray.d));
                    const float3 emission = corrected ? tex : p.beta * tex;
                    const float mis_weight = balanced_heuristic(p.pdf_bsdf, env_sampled_light_pdf(sc, p.ray.d));
                    p.L = p.L + (p.beta * mis_weight) * emission;
                }
                end = true;
            } else {
                const Interaction itx = surface_interact(sc, hit);
                const float3 emission = inst_emission(sc, hit.inst);
                if (backfacing(p.ray.d, itx)) {                                    // prb.py:35-36
                    wf_load_throughput(a, p);
                    end = true;
                } else if (any_pos(emission)) {                                    // prb.py:39-44 / :117-121
                    wf_load_throughput(a, p);
                    const float mis_weight = balanced_heuristic(p.pdf_bsdf, sample_light_pdf(sc, p.ray.o, hit, itx.p));
                    const float3 contrib = (p.beta * mis_weight) * emission;
                    p.L = (phase == 0) ? p.L + contrib : p.L - contrib;
                    end = true;
                } else if ((phase == 0 || corrected) && inst_material_set(sc, hit.inst) < 0) {          // prb.py:45-46
                    wf_load_throughput(a, p);
                    end = true;
                } else {
                    tl.shaded++;
                    const LightSample light = sample_light<ENV>(sc, itx.p, p.sampler);                 // prb.py:57
                    w.SAB[2 * (size_t)(p.slot)] = make_float4(itx.p.x, itx.p.y, itx.p.z, light.dist);
                    w.SAB[2 * (size_t)(p.slot) + 1] = make_float4(light.wi.x, light.wi.y, light.wi.z, light.pdf);
                    w.SC[p.slot] = make_float4(light.eval.x, light.eval.y, light.eval.z, 0.f);
                    w.AB[2 * (size_t)(p.slot) + 1].w = __uint_as_float(wf_pack(p.depth, phase, p.sampler.dimension));
                    keep = true;
                    entry = p.slot;
                    if (w.bin_out) key = wf_key(w, itx.p, light.wi);
                }
            }
            if (end) wf_finish<MODE>(a, p);
        }
        if (w.bin_out) wf_append_binned(w, 1, keep, entry, key);
        else wf_append(w.list[1], w.counts + 1, keep, entry);
    }
    flush_tally(a, tl);
}

// Stage B: light-sample contribution given the occlusion result, BSDF sample, throughput update,
// Russian roulette (prb.py:61-87), and in the adjoint walk the BRDF adjoints + scatter
// (prb.py:138-187).  The interaction and the texel fetch are recomputed from the hit record (a few
// hundred instructions) instead of travelling through HBM.  list[1] -> list[0].
template <int MODE, bool STATS, bool AGG>
__global__ void __launch_bounds__(kWfThreads, MODE == 2 ? ZDR_WF_SHADE_BB_MINB : ZDR_WF_SHADE_B_MINB) k_wf_shade_b(const __grid_constant__ RenderArgs a) {
    const WfState& w = a.wf;
    const SceneView& sc = a.sc;
    const int n = *(volatile int*)(w.counts + 1);
    const int n_round = (n + 31) & ~31;
    Tally tl = {};
    Ctx<STATS, AGG> c{a, tl};
    if constexpr (MODE == 2) { ZDR_GRAD_CACHE_BEGIN(c) }
    const bool corrected = a.prb_mode == ZDR_PRB_CORRECTED;
    const float specular = 0.04f;
    const float3 zero = f3(0.f, 0.f, 0.f);
    constexpr bool ADJ = MODE == 2;      // adjoint walk: BRDF adjoints + scatter
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += wf_stride()) {
        bool survive = false;
        int slot = 0, key = 0;
        if (i < n) {
            slot = w.list[1][i];
            {
#if ZDR_WF_EARLY_SHADOW_LOADS
                const float4 SB = w.SAB[2 * (size_t)(slot) + 1], SC = w.SC[slot];      // issued with the other state loads of the slot
#endif
                WfPath p = wf_load(a, slot, true);
                const Hit hit = wf_hit(a, slot);
                float3 le_grad = zero;
                if (ADJ) {                                         // integrator.py:38-40
                    const float4 g = __ldg(a.d_image + ((size_t)p.x + (size_t)p.y * a.W));
                    le_grad = f3(g.x / (float)a.spp, g.y / (float)a.spp, g.z / (float)a.spp);
                    if (any_nan(le_grad)) le_grad = zero;
                }
                const Interaction itx = surface_interact(sc, hit);
                const int mset = inst_texture_set(sc, hit.inst);
#if ZDR_WF_EARLY_TEXELS
                BsdfTexels tx;
                if (a.tex.bilinear) tx = fetch_bsdf(a.tex, itx.uv, mset);      // texel loads in flight while the frame is built
                const Onb onb = make_onb(itx.ns);
                const float3 wo_local = to_local(onb, -p.ray.d);
                const float4 mat = a.tex.bilinear ? combine_bsdf(tx) : read_bsdf(a.tex, itx.uv, mset);
                const float3 diffuse = f3(mat.x, mat.y, mat.z);
                const float roughness = mat.w;
#else
                const float4 mat = read_bsdf(a.tex, itx.uv, mset);
                const float3 diffuse = f3(mat.x, mat.y, mat.z);
                const float roughness = mat.w;
                const Onb onb = make_onb(itx.ns);
                const float3 wo_local = to_local(onb, -p.ray.d);
#endif
#if !ZDR_WF_EARLY_SHADOW_LOADS
                const float4 SB = w.SAB[2 * (size_t)(slot) + 1], SC = w.SC[slot];
#endif
                const float3 light_wi = f3(SB.x, SB.y, SB.z), light_eval = f3(SC.x, SC.y, SC.z);
                const float light_pdf = SB.w;
                const bool occluded = SC.w != 0.0f;
                float4 mat_grad = make_float4(0.f, 0.f, 0.f, 0.f);
                bool end = false;
                const float3 wi_light_local = to_local(onb, light_wi);
                if (!occluded && wi_light_local.z >= 1e-4f) {                                       // prb.py:61-66
                    float2 gb = make_float2(0.f, 0.f);      // adjoint: (dk, common) of ggx_brdf_grad, computed with the value (shading.cuh ggx_eval_adj)
                    const float4 ev = ADJ ? ggx_eval_adj(wo_local, wi_light_local, diffuse, specular, roughness, gb)
                                          : ggx_eval(wo_local, wi_light_local, diffuse, specular, roughness);
                    const float3 bsdf = f3(ev.x, ev.y, ev.z);
                    const float mis_weight = balanced_heuristic(light_pdf, ev.w);
                    const float dn = fmaxf(light_pdf, 1e-4f);
                    const float3 contrib = (((p.beta * bsdf) * mis_weight) * light_eval) / dn;
                    if (!ADJ) {
                        p.L = p.L + contrib;
                    } else {
                        const float3 cot = (((p.beta * mis_weight) * light_eval) / dn) * le_grad;     // prb.py:145
                        mat_grad = mat_grad + ggx_grad_from_basis(gb, cot);
                        p.L = p.L - contrib;                                                        // prb.py:147
                    }
                }
                const float3 wi_local = ggx_sample(sc, wo_local, roughness, p.sampler);             // prb.py:69
                float2 gb = make_float2(0.f, 0.f);
                const float4 ev = ADJ ? ggx_eval_adj(wo_local, wi_local, diffuse, specular, roughness, gb) : ggx_eval(wo_local, wi_local, diffuse, specular, roughness);
                p.pdf_bsdf = ev.w;
                const float3 wi = to_world(onb, wi_local);
                if (dot(wi, itx.ng) < 1e-4f || wi_local.z < 1e-4f) {                                // prb.py:73-74
                    end = true;
                } else {
                    p.ray = make_ray(offset_ray_origin(itx.p, itx.ng), wi, 0.0f, 1e30f);            // prb.py:75
                    const float3 bsdf = f3(ev.x, ev.y, ev.z);
                    if (ADJ) {
                        float3 cot;
                        if (corrected)
                            cot = f3(bsdf.x != 0.0f ? p.L.x / bsdf.x * le_grad.x : 0.0f, bsdf.y != 0.0f ? p.L.y / bsdf.y * le_grad.y : 0.0f,
                                     bsdf.z != 0.0f ? p.L.z / bsdf.z * le_grad.z : 0.0f);
                        else
                            cot = ((p.beta / p.pdf_bsdf) * p.L) * le_grad;                          // prb.py:162
                        mat_grad = mat_grad + ggx_grad_from_basis(gb, cot);
                    }
                    p.beta = p.beta * (bsdf / p.pdf_bsdf);                                          // prb.py:76
                    if (p.depth >= a.rr_depth) {                                                    // prb.py:78-87
                        const float l = luminance(p.beta);
                        if (l == 0.0f) {
                            end = true;
                        } else {
                            const float q = fmaxf(l, 0.05f);
                            const float r = sampler_next(sc, p.sampler);
                            if (r >= q) end = true;
                            else p.beta = p.beta / q;
                        }
                    }
                }
                if (ADJ) {                                                          // prb.py:178-187
                    if ((mat_grad.x != 0.0f || mat_grad.y != 0.0f || mat_grad.z != 0.0f || mat_grad.w != 0.0f) && grad_ok(mat_grad))
                        c.scatter(itx.uv, mat_grad, mset);
                }
                ++p.depth;
                if (p.depth >= a.max_depth) end = true;                                             // prb.py:23
                if (end) {
                    wf_finish<MODE>(a, p);
                } else {
                    survive = true;
                    w.AB[2 * (size_t)(slot)] = make_float4(p.ray.o.x, p.ray.o.y, p.ray.o.z, p.pdf_bsdf);
                    w.AB[2 * (size_t)(slot) + 1] = make_float4(p.ray.d.x, p.ray.d.y, p.ray.d.z, __uint_as_float(wf_pack(p.depth, ADJ ? 1 : 0, p.sampler.dimension)));
                    w.CD[2 * (size_t)(slot)] = make_float4(p.beta.x, p.beta.y, p.beta.z, 0.f);
                    w.CD[2 * (size_t)(slot) + 1] = make_float4(p.L.x, p.L.y, p.L.z, 0.f);
                    if (w.bin_out) key = wf_key(w, p.ray.o, p.ray.d);
                }
            }
        }
        if (w.bin_out) wf_append_binned(w, 0, survive, slot, key);
        else wf_append(w.list[0], w.counts + 0, survive, slot);
    }
    if constexpr (MODE == 2) { ZDR_GRAD_CACHE_END(c, a, tl) }
    flush_tally(a, tl);
}

// integrator.py:15-29: s += clamp(radiance) for it ascending; image = s / spp.  The running sum
// lives in the image buffer between batches.
__global__ void __launch_bounds__(kWfThreads) k_wf_accumulate(const __grid_constant__ RenderArgs a) {
    const WfState& w = a.wf;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < w.P; pix += wf_stride()) {
        int x, y;
        wf_pixel(a, pix, x, y);
        float4* px = a.image + ((size_t)x + (size_t)y * a.W);
        float3 s = f3(0.f, 0.f, 0.f);
        if (!w.first) { const float4 v = *px; s = f3(v.x, v.y, v.z); }
        for (int j = 0; j < w.S; ++j) {
            const float4 r = w.R[(size_t)j * w.P + pix];
            s = s + f3(r.x, r.y, r.z);
        }
        *px = w.last ? make_float4(s.x / (float)a.spp, s.y / (float)a.spp, s.z / (float)a.spp, 1.0f) : make_float4(s.x, s.y, s.z, 0.f);
    }
}

// Between the replay rounds and the adjoint rounds of the adjoint pass (prb.py:97-101): every sample
// whose replayed radiance Le is not NaN starts again at the camera with the ORIGINAL sampler state,
// carrying Le as the "radiance still to be explained".
__global__ void __launch_bounds__(kWfThreads) k_wf_restart(const __grid_constant__ RenderArgs a) {
    const WfState& w = a.wf;
    const int n_round = (w.nslots + 31) & ~31;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_round; slot += wf_stride()) {
        bool keep = false;
        if (slot < w.nslots && a.max_depth > 0) {
            const float4 D4 = w.CD[2 * (size_t)(slot) + 1];
            const float3 Le = f3(D4.x, D4.y, D4.z);
            if (!any_nan(Le)) {
                const int j = fastdiv(slot, w.div_P), pix = slot - j * w.P;
                int x, y;
                wf_pixel(a, pix, x, y);
                Sampler cam = make_sampler(a.sc, x, y, a.seed, a.spp, w.it_base + j);
                const Ray ray = camera_sample(a, x, y, cam);
                w.AB[2 * (size_t)(slot)] = make_float4(ray.o.x, ray.o.y, ray.o.z, 1e30f);
                w.AB[2 * (size_t)(slot) + 1] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(wf_pack(0, 1, cam.dimension)));
                w.CD[2 * (size_t)(slot)] = make_float4(1.f, 1.f, 1.f, 0.f);
                keep = true;
            }
        }
        wf_append(w.list[0], w.counts + 0, keep, slot);
    }
}

template <bool ANY>
void launch_wf_trace(bool stats, int tgrid, cudaStream_t s, const RenderArgs& a) {
    const bool big = a.sc.num_nodes > kWfBigSceneNodes;
    if (ZDR_WF_TINY && a.wf.tiny_regs > 0) {
        const size_t sm = 64 * (size_t)a.sc.num_nodes + 48 * (size_t)a.sc.num_tris;
#if ZDR_WF_TINY_LOCAL_STACK
        if (stats) k_wf_trace_tiny<ANY, true, 4><<<tgrid, kWfThreads, sm, s>>>(a); else k_wf_trace_tiny<ANY, false, 4><<<tgrid, kWfThreads, sm, s>>>(a);
#else      // the register-stack build has a second instantiation for deeper trees
        if (a.wf.tiny_regs <= 4) { if (stats) k_wf_trace_tiny<ANY, true, 4><<<tgrid, kWfThreads, sm, s>>>(a); else k_wf_trace_tiny<ANY, false, 4><<<tgrid, kWfThreads, sm, s>>>(a); }
        else { if (stats) k_wf_trace_tiny<ANY, true, 6><<<tgrid, kWfThreads, sm, s>>>(a); else k_wf_trace_tiny<ANY, false, 6><<<tgrid, kWfThreads, sm, s>>>(a); }
#endif
        return;
    }
    if (stats) { if (big) k_wf_trace<ANY, true, true><<<tgrid, kWfThreads, 0, s>>>(a); else k_wf_trace<ANY, true, false><<<tgrid, kWfThreads, 0, s>>>(a); }
    else { if (big) k_wf_trace<ANY, false, true><<<tgrid, kWfThreads, 0, s>>>(a); else k_wf_trace<ANY, false, false><<<tgrid, kWfThreads, 0, s>>>(a); }
}
// One bounce (round r of a pass).  From the second bounce on the rays of a traversal stage are binned first (WfState): the
// first bounce's rays are coherent as they come (camera rays / shadow rays of neighbouring pixels).
template <int MODE>
void launch_wf_round(int r, bool stats, bool agg, bool env, int grid, cudaStream_t s, const RenderArgs& a0) {
    RenderArgs a = a0;
    const int tgrid = min(grid, a.wf.num_sms * ZDR_WF_TRACE_GRID_MULT);
    const int bgrid = min(grid, a.wf.num_sms * 2);      // k_wf_bin: one tile of the list per block
    const bool bin_closest = (a.wf.bin_mask & 1) && r >= 1, bin_any = (a.wf.bin_mask & 2) && r >= 1;
    const bool bin_next = (a.wf.bin_mask & 1) && r + 1 < a.max_depth;      // stage B feeds the next bounce's extension rays
    g_kernel_launches += 4 + (bin_closest ? 1 : 0) + (bin_any ? 1 : 0);
    a.wf.bin_out = 0;
    if (bin_closest) k_wf_bin<0><<<bgrid, kWfBinThreads, 0, s>>>(a);
    a.wf.bin_in = bin_closest ? 1 : 0;
    launch_wf_trace<false>(stats, tgrid, s, a);
    a.wf.bin_in = 0; a.wf.bin_out = bin_any ? 1 : 0;
    if (env) k_wf_shade_a<MODE, true><<<grid, kWfThreads, 0, s>>>(a);
    else k_wf_shade_a<MODE, false><<<grid, kWfThreads, 0, s>>>(a);
    a.wf.bin_out = 0;
    if (bin_any) k_wf_bin<1><<<bgrid, kWfBinThreads, 0, s>>>(a);
    a.wf.bin_in = bin_any ? 1 : 0;
    launch_wf_trace<true>(stats, tgrid, s, a);
    a.wf.bin_in = 0; a.wf.bin_out = bin_next ? 1 : 0;
    if constexpr (MODE == 2) {
        if (agg) {
            const size_t gc = a.block_scatter ? kGradCacheBytes : 0;
            if (stats) k_wf_shade_b<MODE, true, true><<<grid, kWfThreads, gc, s>>>(a); else k_wf_shade_b<MODE, false, true><<<grid, kWfThreads, gc, s>>>(a);
            return;
        }
    }
    { if (stats) k_wf_shade_b<MODE, true, false><<<grid, kWfThreads, 0, s>>>(a); else k_wf_shade_b<MODE, false, false><<<grid, kWfThreads, 0, s>>>(a); }
}
}  // namespace

#ifndef ZDR_FAST_MATH      // exist once (exact build): workspace sizes
// 9 state records + list[2]; with ray binning also blist + keys[2] per slot and hist[2] + cur per set
size_t wf_bytes_per_slot(bool binned) { return 9 * sizeof(float4) + 2 * sizeof(int) + (binned ? sizeof(int) + 2 * sizeof(unsigned short) : 0); }
size_t wf_fixed_bytes(bool binned) { return 256 + (binned ? 3 * kWfBins * sizeof(int) : 0); }
#endif

// Host side of the wavefront path tracer: batches of S consecutive samples of every pixel of the
// shard; per batch generate, then max_depth (forward) or 2*max_depth (adjoint: replay + adjoint
// walk) rounds of extend + shade, then the ordered accumulation.  No host synchronisation: list
// lengths stay on the device, the grids are persistent (grid-stride over the live list).
namespace {
void wf_shift_set(WfState& w, size_t bytes) {
    auto sh = [bytes](auto*& p) { if (p) p = reinterpret_cast<std::remove_reference_t<decltype(p)>>(reinterpret_cast<char*>(p) + bytes); };
    sh(w.AB); sh(w.CD); sh(w.H); sh(w.R); sh(w.SAB); sh(w.SC); sh(w.list[0]); sh(w.list[1]); sh(w.counts);
    sh(w.blist); sh(w.keys[0]); sh(w.keys[1]); sh(w.hist[0]); sh(w.hist[1]); sh(w.cur);
}
}  // namespace

cudaError_t launch_path_wavefront(bool backward, RenderArgs a, bool stats, cudaStream_t s0, cudaStream_t* side, cudaEvent_t* ev) {
    const int tiles = (a.row1 - a.row0 + 7) / 8;
    int rows = 0;
    for (int t = a.shard_index; t < tiles; t += a.shard_count) rows += min(8, a.row1 - (a.row0 + 8 * t));
    if (rows <= 0 || a.it1 <= a.it0) return cudaSuccess;
    WfState& w = a.wf;
    w.P = a.W * rows;
    w.div_P = make_fastdiv(w.P);
    w.div_W = make_fastdiv(a.W);
    const int smax = w.S;      // capacity in samples per batch (buffers hold P * smax slots)
    // persistent grids: ZDR_WF_GRID_MULT blocks per SM, but never more blocks than the batch has work for (a near-empty
    // kernel of 9 472 blocks costs ~30 us, x ~100 launches per batch: 3 ms on a job of 2^18 paths)
    const long long max_slots = (long long)w.P * min(smax, a.it1 - a.it0);
    const int grid = (int)std::min<long long>((long long)w.num_sms * ZDR_WF_GRID_MULT,
                                               std::max<long long>(w.num_sms, (max_slots + kWfThreads - 1) / kWfThreads));
    // Several state sets: batch b runs on stream b % K with set b % K.  Kernels of one batch are ordered by their
    // stream; the only dependency between batches is the running sum in the image (k_wf_accumulate, sample order).
    const int batches = (a.it1 - a.it0 + smax - 1) / smax;
    const int K = (side && ev) ? min(min(w.num_sets, kWfMaxSets), batches) : 1;
    RenderArgs sets[kWfMaxSets];
    cudaStream_t streams[kWfMaxSets];
    for (int k = 0; k < K; ++k) {
        sets[k] = a;
        if (k) wf_shift_set(sets[k].wf, w.set_bytes * k);
        streams[k] = k ? side[k - 1] : s0;
    }
    if (K > 1) {                                   // fork
        cudaError_t e = cudaEventRecord(ev[0], s0);
        for (int k = 1; k < K && e == cudaSuccess; ++k) e = cudaStreamWaitEvent(streams[k], ev[0], 0);
        if (e != cudaSuccess) return e;
    }
    int b = 0;
    for (int base = a.it0; base < a.it1; base += smax, ++b) {
        RenderArgs& ab = sets[b % K];
        cudaStream_t s = streams[b % K];
        WfState& wb = ab.wf;
        wb.P = w.P; wb.div_P = w.div_P; wb.div_W = w.div_W;
        wb.S = min(smax, a.it1 - base);
        wb.it_base = base;
        wb.nslots = wb.P * wb.S;
        wb.first = base == a.it0;
        wb.last = base + smax >= a.it1;
        g_kernel_launches += (!backward && a.image) ? 2 : 1;      // generate (+ accumulate)
        if (backward) k_wf_generate<true><<<grid, kWfThreads, 0, s>>>(ab);
        else k_wf_generate<false><<<grid, kWfThreads, 0, s>>>(ab);
        const bool env = a.sc.env_count > 0;
        if (!backward) {
            for (int r = 0; r < a.max_depth; ++r) launch_wf_round<0>(r, stats, false, env, grid, s, ab);
        } else {
            for (int r = 0; r < a.max_depth; ++r) launch_wf_round<1>(r, stats, false, env, grid, s, ab);      // replay: Le of every sample
            ++g_kernel_launches;
            k_wf_restart<<<grid, kWfThreads, 0, s>>>(ab);
            for (int r = 0; r < a.max_depth; ++r) launch_wf_round<2>(r, stats, a.aggregate_scatter != 0, env, grid, s, ab);   // adjoint walk
        }
        if (!backward && a.image) {
            if (K > 1 && b > 0) { cudaError_t e = cudaStreamWaitEvent(s, ev[1 + (b - 1) % K], 0); if (e != cudaSuccess) return e; }
            k_wf_accumulate<<<grid, kWfThreads, 0, s>>>(ab);
            if (K > 1) { cudaError_t e = cudaEventRecord(ev[1 + b % K], s); if (e != cudaSuccess) return e; }
        }
    }
    for (int k = 1; k < K; ++k) {                  // join: the caller's stream continues after all of them
        cudaError_t e = cudaEventRecord(ev[1 + kWfMaxSets], streams[k]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s0, ev[1 + kWfMaxSets], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace zdr
