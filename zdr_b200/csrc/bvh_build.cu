// bvh_build.cu — GPU LBVH build (replaces luisa.Accel().add/.update, render.py:74,109,127,
// which is an OptiX library call inside LuisaCompute).
//
// Pipeline (all on the caller's stream, no host round-trips except one size read-back):
//   1. 63-bit Morton code of every triangle centroid            (k_morton)
//   2. radix sort of (code, gid)                                 (cub::DeviceRadixSort)
//   3. Karras 2012 binary radix tree, one thread per inner node  (k_hierarchy)
//   4. bottom-up box fit with per-node arrival counters          (k_fit)
//   5. collapse of subtrees with <= 4 triangles into leaves, prefix-sum compaction of the
//      surviving inner nodes, emission of the 64-byte traversal nodes with both child boxes
//      inlined and slightly inflated                             (k_mark, cub scan, k_emit)
//   6. gather of the triangle / shading records into leaf order  (k_gather)
// Karras leaves cover contiguous ranges of the sorted order, so a collapsed subtree is just
// (first_slot, count).
//
// Default for large scenes (build_bvh_clustered_sah): steps 1-4 as above, then the radix tree is cut
// into Morton-prefix-aligned clusters of <= kClusterMax triangles; one thread block per cluster builds a
// binned-SAH subtree in shared memory (k_sah_cluster), a full-sweep SAH tree over the cluster boxes is
// built on the host meanwhile, and the two are stitched into one flat node array.  The Morton order
// only decides which triangles share a cluster; every split above and below is a SAH split.
#include <cub/cub.cuh>
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "bvh_build.h"

namespace zdr {

namespace {

__device__ __forceinline__ uint64_t expand21(uint32_t v) {   // spread 21 bits to every 3rd bit
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ tri_in, int n, float3 lo, float3 inv_ext,
                         uint64_t* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = tri_in[3 * i], b = tri_in[3 * i + 1], c = tri_in[3 * i + 2];
    float cx = (fminf(fminf(a.x, b.x), c.x) + fmaxf(fmaxf(a.x, b.x), c.x)) * 0.5f;
    float cy = (fminf(fminf(a.y, b.y), c.y) + fmaxf(fmaxf(a.y, b.y), c.y)) * 0.5f;
    float cz = (fminf(fminf(a.z, b.z), c.z) + fmaxf(fmaxf(a.z, b.z), c.z)) * 0.5f;
    const float S = 2097151.0f;   // 2^21 - 1
    uint32_t qx = (uint32_t)fminf(fmaxf((cx - lo.x) * inv_ext.x * S, 0.0f), S);
    uint32_t qy = (uint32_t)fminf(fmaxf((cy - lo.y) * inv_ext.y * S, 0.0f), S);
    uint32_t qz = (uint32_t)fminf(fmaxf((cz - lo.z) * inv_ext.z * S, 0.0f), S);
    keys[i] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
    vals[i] = i;
}

// length of the common prefix of sorted keys i and j (index-augmented to break duplicate codes)
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012, Algorithm "construct binary radix tree". Children < 0 encode leaves: ~child = sorted index.
__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ child,
                            int2* __restrict__ range, int* __restrict__ parent_inner, int* __restrict__ parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? ~gamma : gamma;
    int right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    child[i] = make_int2(left, right);
    range[i] = make_int2(first, last);
    if (left < 0) parent_leaf[~left] = i; else parent_inner[left] = i;
    if (right < 0) parent_leaf[~right] = i; else parent_inner[right] = i;
    if (i == 0) parent_inner[0] = -1;
}

struct Box { float3 lo, hi; };
__device__ __forceinline__ Box box_union(Box a, Box b) {
    Box r;
    r.lo = make_float3(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z));
    r.hi = make_float3(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z));
    return r;
}

// boxes: [0, n-1) inner nodes, [n-1, 2n-1) leaves (sorted order)
__global__ void k_fit(const float4* __restrict__ tri_in, const int* __restrict__ sorted_gid, int n,
                      const int2* __restrict__ child, const int* __restrict__ parent_inner,
                      const int* __restrict__ parent_leaf, float4* __restrict__ box_lo, float4* __restrict__ box_hi,
                      int* __restrict__ arrive, int* __restrict__ max_depth) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int g = sorted_gid[s];
    float4 a = tri_in[3 * g], b = tri_in[3 * g + 1], c = tri_in[3 * g + 2];
    Box bx;
    bx.lo = make_float3(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z));
    bx.hi = make_float3(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z));
    box_lo[n - 1 + s] = make_float4(bx.lo.x, bx.lo.y, bx.lo.z, 0.f);
    box_hi[n - 1 + s] = make_float4(bx.hi.x, bx.hi.y, bx.hi.z, 0.f);
    if (n == 1) return;
    int node = parent_leaf[s];
    int depth = 1;
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&arrive[node], 1) == 0) return;   // first child to arrive stops; second continues
        int2 ch = child[node];
        int li = ch.x < 0 ? n - 1 + ~ch.x : ch.x, ri = ch.y < 0 ? n - 1 + ~ch.y : ch.y;
        // volatile-style reads through L2: written by another thread before its fence+atomic
        float4 llo = __ldcg(&box_lo[li]), lhi = __ldcg(&box_hi[li]);
        float4 rlo = __ldcg(&box_lo[ri]), rhi = __ldcg(&box_hi[ri]);
        Box u;
        u.lo = make_float3(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z));
        u.hi = make_float3(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z));
        box_lo[node] = make_float4(u.lo.x, u.lo.y, u.lo.z, 0.f);
        box_hi[node] = make_float4(u.hi.x, u.hi.y, u.hi.z, 0.f);
        ++depth;
        node = parent_inner[node];
    }
    atomicMax(max_depth, depth);
}

#ifndef ZDR_BVH_LEAF_MAX
#define ZDR_BVH_LEAF_MAX 4
#endif
constexpr int kLeafMax = ZDR_BVH_LEAF_MAX;      // triangles per leaf (2 bits of the leaf link hold count - 1)

// an inner node survives iff it covers more than kLeafMax triangles
__global__ void k_mark(const int2* __restrict__ range, int n, int* __restrict__ keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int2 r = range[i];
    keep[i] = (r.y - r.x + 1) > kLeafMax ? 1 : 0;
}

__device__ __forceinline__ int make_link(int c, const int2* __restrict__ range, const int* __restrict__ new_index) {
    if (c < 0) return ~(((~c) << 2) | 0);                 // single-triangle leaf
    int2 r = range[c];
    int cnt = r.y - r.x + 1;
    if (cnt <= kLeafMax) return ~((r.x << 2) | (cnt - 1));
    return new_index[c];
}

__global__ void k_emit(const int2* __restrict__ child, const int2* __restrict__ range, const int* __restrict__ keep,
                       const int* __restrict__ new_index, const float4* __restrict__ box_lo,
                       const float4* __restrict__ box_hi, int n, float eps, float4* __restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !keep[i]) return;
    int2 ch = child[i];
    int li = ch.x < 0 ? n - 1 + ~ch.x : ch.x, ri = ch.y < 0 ? n - 1 + ~ch.y : ch.y;
    float4 llo = box_lo[li], lhi = box_hi[li], rlo = box_lo[ri], rhi = box_hi[ri];
    float4* o = nodes + 4 * (size_t)new_index[i];
    o[0] = make_float4(llo.x - eps, lhi.x + eps, llo.y - eps, lhi.y + eps);
    o[1] = make_float4(rlo.x - eps, rhi.x + eps, rlo.y - eps, rhi.y + eps);
    o[2] = make_float4(llo.z - eps, lhi.z + eps, rlo.z - eps, rhi.z + eps);
    o[3] = make_float4(__int_as_float(make_link(ch.x, range, new_index)),
                       __int_as_float(make_link(ch.y, range, new_index)), 0.f, 0.f);
}

__global__ void k_gather(const float4* __restrict__ tri_in, const float4* __restrict__ shade_in,
                         const int* __restrict__ sorted_gid, int n, float4* __restrict__ tris,
                         float4* __restrict__ shade, int* __restrict__ slot_of_gid) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int g = sorted_gid[s];
    slot_of_gid[g] = s;
#pragma unroll
    for (int k = 0; k < 3; ++k) tris[3 * (size_t)s + k] = tri_in[3 * (size_t)g + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) shade[4 * (size_t)s + k] = shade_in[4 * (size_t)g + k];
}


// ---------------------------------------------------------------------------------------------
// Clustered SAH build
constexpr int kClusterMax = 2048;      // triangles per cluster (shared-memory resident: 32 B each)
constexpr int kSahThreads = 256;
constexpr int kBins = 16;

// Cluster roots of the radix tree: the highest nodes that cover <= kClusterMax triangles.  Written at the
// first sorted position of the cluster, so that an exclusive scan of the start flags numbers them in order.
__global__ void k_cluster_mark(const int2* __restrict__ range, const int* __restrict__ parent_inner,
                               const int* __restrict__ parent_leaf, int n, int* __restrict__ start_flag,
                               int* __restrict__ count_at, int* __restrict__ box_at) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) {                                   // inner node i
        const int2 r = range[i];
        const int cnt = r.y - r.x + 1;
        if (cnt <= kClusterMax) {
            const int p = parent_inner[i];
            bool top = p < 0;
            if (!top) { const int2 pr = range[p]; top = pr.y - pr.x + 1 > kClusterMax; }
            if (top) { start_flag[r.x] = 1; count_at[r.x] = cnt; box_at[r.x] = i; }
        }
    }
    if (i < n) {                                       // sorted triangle i hanging directly off a big node
        const int2 pr = range[parent_leaf[i]];
        if (pr.y - pr.x + 1 > kClusterMax) { start_flag[i] = 1; count_at[i] = 1; box_at[i] = n - 1 + i; }
    }
}

__global__ void k_cluster_compact(const int* __restrict__ start_flag, const int* __restrict__ cluster_id,
                                  const int* __restrict__ count_at, const int* __restrict__ box_at,
                                  const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, int n,
                                  int* __restrict__ cl_first, int* __restrict__ cl_count, float4* __restrict__ cl_box) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || !start_flag[s]) return;
    const int c = cluster_id[s];
    cl_first[c] = s;
    cl_count[c] = count_at[s];
    cl_box[2 * c] = box_lo[box_at[s]];
    cl_box[2 * c + 1] = box_hi[box_at[s]];
}

// float <-> int whose signed order is the float order (shared-memory atomicMin/atomicMax on boxes)
__device__ __forceinline__ int f2o(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct SahShared {
    float lo[3][kClusterMax], hi[3][kClusterMax];      // triangle boxes by position in the cluster
    int idx[kClusterMax], idx2[kClusterMax];           // current order / partition scratch
    int bin_cnt[3][kBins];
    int bin_lo[3][kBins][3], bin_hi[3][kBins][3];
    int cb_lo[3], cb_hi[3];                            // centroid bounds of the range
    float cost[3 * (kBins - 1) + 3];
    int stack[32][5];                                  // a, b, parent node, side, depth
    int sp, num_nodes, max_depth, cur_node;
    int best_axis, best_split, make_leaf;
    float child_box[2][6];
    typename cub::BlockScan<int, kSahThreads>::TempStorage scan;
};

// One block per cluster: top-down binned SAH (16 bins x 3 axes) over the cluster's triangles, entirely
// in shared memory.  Writes inner nodes to tmp_nodes[first .. first + used) with cluster-local inner
// links and absolute leaf links, the cluster's root link, the number of nodes used, and the cluster's
// part of the final triangle order.  The tree depends only on the set of triangles in each range
// (min / max / count are order-independent; the partition is stable), so builds are reproducible.
__global__ void __launch_bounds__(kSahThreads) k_sah_cluster(const int* __restrict__ cl_first, const int* __restrict__ cl_count,
                                                             const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                                                             const int* __restrict__ sorted_gid, int n, float eps,
                                                             float4* __restrict__ tmp_nodes, int* __restrict__ cl_root,
                                                             int* __restrict__ cl_used, int* __restrict__ final_gid,
                                                             int* __restrict__ max_depth_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SahShared& sh = *reinterpret_cast<SahShared*>(smem_raw);
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int first = cl_first[c], cnt = cl_count[c];
    for (int j = tid; j < cnt; j += kSahThreads) {
        const float4 l = box_lo[n - 1 + first + j], h = box_hi[n - 1 + first + j];
        sh.lo[0][j] = l.x; sh.lo[1][j] = l.y; sh.lo[2][j] = l.z;
        sh.hi[0][j] = h.x; sh.hi[1][j] = h.y; sh.hi[2][j] = h.z;
        sh.idx[j] = j;
    }
    if (tid == 0) {
        sh.sp = 1; sh.num_nodes = 0; sh.max_depth = 0;
        sh.stack[0][0] = 0; sh.stack[0][1] = cnt; sh.stack[0][2] = -1; sh.stack[0][3] = 0; sh.stack[0][4] = 1;
    }
    __syncthreads();
    for (;;) {
        const int sp = sh.sp;
        if (sp == 0) break;
        int a = sh.stack[sp - 1][0], b = sh.stack[sp - 1][1], parent = sh.stack[sp - 1][2], side = sh.stack[sp - 1][3],
            depth = sh.stack[sp - 1][4];
        __syncthreads();
        if (tid == 0) sh.sp = sp - 1;
        for (;;) {                                     // this range, then its smaller child, ...
            const int m = b - a;
            // ---- centroid bounds of the range
            if (tid < 3) { sh.cb_lo[tid] = 0x7fffffff; sh.cb_hi[tid] = (int)0x80000000; }
            for (int k = tid; k < 3 * kBins; k += kSahThreads) {
                const int ax = k / kBins, bn = k % kBins;
                sh.bin_cnt[ax][bn] = 0;
                for (int q = 0; q < 3; ++q) { sh.bin_lo[ax][bn][q] = 0x7fffffff; sh.bin_hi[ax][bn][q] = (int)0x80000000; }
            }
            __syncthreads();
            {
                float cmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                for (int j = a + tid; j < b; j += kSahThreads) {
                    const int e = sh.idx[j];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        const float ctr = sh.lo[ax][e] + sh.hi[ax][e];
                        cmin[ax] = fminf(cmin[ax], ctr); cmax[ax] = fmaxf(cmax[ax], ctr);
                    }
                }
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    for (int o = 16; o > 0; o >>= 1) {
                        cmin[ax] = fminf(cmin[ax], __shfl_xor_sync(0xffffffffu, cmin[ax], o));
                        cmax[ax] = fmaxf(cmax[ax], __shfl_xor_sync(0xffffffffu, cmax[ax], o));
                    }
                    if (lane == 0 && cmin[ax] <= cmax[ax]) { atomicMin(&sh.cb_lo[ax], f2o(cmin[ax])); atomicMax(&sh.cb_hi[ax], f2o(cmax[ax])); }
                }
            }
            __syncthreads();
            float cmin[3], scale[3];
            bool degenerate = true;                    // all centroids coincide: split by position instead
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                cmin[ax] = o2f(sh.cb_lo[ax]);
                const float ext = o2f(sh.cb_hi[ax]) - cmin[ax];
                const float sc = (float)kBins / ext;
                scale[ax] = (ext > 0.0f && sc <= FLT_MAX) ? sc : 0.0f;
                if (scale[ax] > 0.0f) degenerate = false;
            }
            auto bin_of = [&](int j, int e, int ax) -> int {
                if (degenerate) return ax == 0 ? (int)(((long long)(j - a) * kBins) / m) : 0;
                if (scale[ax] == 0.0f) return 0;
                const float ctr = sh.lo[ax][e] + sh.hi[ax][e];
                return min(kBins - 1, max(0, (int)((ctr - cmin[ax]) * scale[ax])));
            };
            // ---- binning
            for (int j = a + tid; j < b; j += kSahThreads) {
                const int e = sh.idx[j];
                int lo_o[3], hi_o[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) { lo_o[q] = f2o(sh.lo[q][e]); hi_o[q] = f2o(sh.hi[q][e]); }
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const int bn = bin_of(j, e, ax);
                    atomicAdd(&sh.bin_cnt[ax][bn], 1);
#pragma unroll
                    for (int q = 0; q < 3; ++q) { atomicMin(&sh.bin_lo[ax][bn][q], lo_o[q]); atomicMax(&sh.bin_hi[ax][bn][q], hi_o[q]); }
                }
            }
            __syncthreads();
            // ---- sweep: candidate t = (axis, split after bin s)
            if (tid < 3 * (kBins - 1)) {
                const int ax = tid / (kBins - 1), s = tid % (kBins - 1);
                float bl[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
                float br[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
                int nl = 0, nr = 0;
                for (int bn = 0; bn < kBins; ++bn) {
                    const int k = sh.bin_cnt[ax][bn];
                    if (k == 0) continue;
                    float* bx = bn <= s ? bl : br;
                    (bn <= s ? nl : nr) += k;
                    for (int q = 0; q < 3; ++q) {
                        bx[q] = fminf(bx[q], o2f(sh.bin_lo[ax][bn][q]));
                        bx[3 + q] = fmaxf(bx[3 + q], o2f(sh.bin_hi[ax][bn][q]));
                    }
                }
                float cost = FLT_MAX;
                if (nl > 0 && nr > 0) {
                    const float lx = bl[3] - bl[0], ly = bl[4] - bl[1], lz = bl[5] - bl[2];
                    const float rx = br[3] - br[0], ry = br[4] - br[1], rz = br[5] - br[2];
                    cost = (lx * ly + ly * lz + lz * lx) * (float)nl + (rx * ry + ry * rz + rz * rx) * (float)nr;
                }
                sh.cost[tid] = cost;
            }
            __syncthreads();
            if (tid == 0) {
                int best = -1;
                float best_cost = FLT_MAX;
                for (int t = 0; t < 3 * (kBins - 1); ++t)
                    if (sh.cost[t] < best_cost) { best_cost = sh.cost[t]; best = t; }
                float pb[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
                for (int bn = 0; bn < kBins; ++bn) {
                    if (sh.bin_cnt[0][bn] == 0) continue;
                    for (int q = 0; q < 3; ++q) {
                        pb[q] = fminf(pb[q], o2f(sh.bin_lo[0][bn][q]));
                        pb[3 + q] = fmaxf(pb[3 + q], o2f(sh.bin_hi[0][bn][q]));
                    }
                }
                const float px = pb[3] - pb[0], py = pb[4] - pb[1], pz = pb[5] - pb[2];
                const float parent_area = fmaxf(px * py + py * pz + pz * px, 1e-30f);
                // same leaf rule as the host builder: a node visit costs one triangle test
                const float split_cost = 1.0f + best_cost / parent_area;
                const bool leaf = m == 1 || best < 0 || (m <= kLeafMax && (float)m <= split_cost);
                sh.make_leaf = leaf ? 1 : 0;
                if (!leaf) {
                    const int ax = best / (kBins - 1), s = best % (kBins - 1);
                    sh.best_axis = ax; sh.best_split = s;
                    for (int ch = 0; ch < 2; ++ch) {
                        float bx[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
                        for (int bn = ch == 0 ? 0 : s + 1; bn <= (ch == 0 ? s : kBins - 1); ++bn) {
                            if (sh.bin_cnt[ax][bn] == 0) continue;
                            for (int q = 0; q < 3; ++q) {
                                bx[q] = fminf(bx[q], o2f(sh.bin_lo[ax][bn][q]));
                                bx[3 + q] = fmaxf(bx[3 + q], o2f(sh.bin_hi[ax][bn][q]));
                            }
                        }
                        for (int q = 0; q < 6; ++q) sh.child_box[ch][q] = bx[q];
                    }
                    sh.cur_node = sh.num_nodes++;
                }
                const int link = leaf ? ~(((first + a) << 2) | (m - 1)) : sh.cur_node;
                if (parent < 0) cl_root[c] = link;
                else reinterpret_cast<int*>(tmp_nodes + 4 * (size_t)(first + parent) + 3)[side] = link;
                sh.max_depth = max(sh.max_depth, depth);
            }
            __syncthreads();
            if (sh.make_leaf) break;
            // a range of more than kLeafMax coincident-centroid triangles always splits (by position), so m > kLeafMax
            // cannot reach here with best < 0
            const int ax = sh.best_axis, split = sh.best_split, node = sh.cur_node;
            // ---- stable partition of idx[a, b) by bin <= split
            const int per = (m + kSahThreads - 1) / kSahThreads;
            const int j0 = min(b, a + tid * per), j1 = min(b, j0 + per);
            int my_left = 0;
            for (int j = j0; j < j1; ++j) my_left += bin_of(j, sh.idx[j], ax) <= split ? 1 : 0;
            int left_before, num_left;
            cub::BlockScan<int, kSahThreads>(sh.scan).ExclusiveSum(my_left, left_before, num_left);
            {
                int lpos = a + left_before, rpos = a + num_left + (j0 - a) - left_before;
                for (int j = j0; j < j1; ++j) {
                    const int e = sh.idx[j];
                    if (bin_of(j, e, ax) <= split) sh.idx2[lpos++] = e; else sh.idx2[rpos++] = e;
                }
            }
            __syncthreads();
            for (int j = a + tid; j < b; j += kSahThreads) sh.idx[j] = sh.idx2[j];
            if (tid == 0) {
                const float* l = sh.child_box[0];
                const float* r = sh.child_box[1];
                float4* o = tmp_nodes + 4 * (size_t)(first + node);
                o[0] = make_float4(l[0] - eps, l[3] + eps, l[1] - eps, l[4] + eps);
                o[1] = make_float4(r[0] - eps, r[3] + eps, r[1] - eps, r[4] + eps);
                o[2] = make_float4(l[2] - eps, l[5] + eps, r[2] - eps, r[5] + eps);
                // o[3] (the links) is written by the children
            }
            const int mid = a + num_left;
            const bool left_small = (mid - a) <= (b - mid);
            if (tid == 0) {                            // push the larger child: the stack stays <= log2(kClusterMax) deep
                const int t = sh.sp++;
                sh.stack[t][0] = left_small ? mid : a; sh.stack[t][1] = left_small ? b : mid;
                sh.stack[t][2] = node; sh.stack[t][3] = left_small ? 1 : 0; sh.stack[t][4] = depth + 1;
            }
            if (left_small) b = mid; else a = mid;
            parent = node; side = left_small ? 0 : 1; depth += 1;
            __syncthreads();
        }
        __syncthreads();
    }
    for (int j = tid; j < cnt; j += kSahThreads) final_gid[first + j] = sorted_gid[first + sh.idx[j]];
    if (tid == 0) { cl_used[c] = sh.num_nodes; atomicMax(max_depth_out, sh.max_depth); }
}

// Packs the clusters' nodes behind the top-level nodes and makes their inner links absolute.
__global__ void k_cluster_pack(const float4* __restrict__ tmp_nodes, const int* __restrict__ start_flag,
                               const int* __restrict__ cluster_id, const int* __restrict__ cl_first,
                               const int* __restrict__ cl_used, const int* __restrict__ cl_base, int n, int num_top,
                               float4* __restrict__ nodes) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int c = cluster_id[p] + start_flag[p] - 1;
    const int j = p - cl_first[c];
    if (j >= cl_used[c]) return;
    const int base = num_top + cl_base[c];
    const float4* src = tmp_nodes + 4 * (size_t)p;
    float4* dst = nodes + 4 * (size_t)(base + j);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    float4 l = src[3];
    int l0 = __float_as_int(l.x), l1 = __float_as_int(l.y);
    if (l0 >= 0) l0 += base;
    if (l1 >= 0) l1 += base;
    dst[3] = make_float4(__int_as_float(l0), __int_as_float(l1), 0.f, 0.f);
}

#define BVH_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

}  // namespace

// ---------------------------------------------------------------------------------------------
// Scene flatten on the GPU (large scenes): world-space triangle + shading records of every instance,
// p = (M * float4(v,1)).xyz as interaction.py:19-21, with the same fmaf chain as the host path
// (xform_point), so both produce identical bits.  One thread per triangle; its mesh is found by
// binary search in the (L1-resident) table of first-triangle indices.
namespace {
__device__ __forceinline__ float3 dev_xform_point(const float* __restrict__ M, float x, float y, float z) {
    return make_float3(fmaf(M[8], z, fmaf(M[4], y, fmaf(M[0], x, M[12]))),
                       fmaf(M[9], z, fmaf(M[5], y, fmaf(M[1], x, M[13]))),
                       fmaf(M[10], z, fmaf(M[6], y, fmaf(M[2], x, M[14]))));
}

__global__ void k_flatten(const float* __restrict__ verts, const int* __restrict__ idx, const int* __restrict__ tri_base,
                          const int* __restrict__ vert_base, const float* __restrict__ xforms, int num_meshes, int nt,
                          float4* __restrict__ tri, float4* __restrict__ shade, int* __restrict__ bounds) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool finite = true;
    if (g < nt) {
        int a = 0, b = num_meshes;                    // last mesh with tri_base <= g
        while (b - a > 1) { const int m = (a + b) >> 1; if (tri_base[m] <= g) a = m; else b = m; }
        const int t = g - tri_base[a];
        const float* M = xforms + 16 * a;
        const float* v[3];
        float3 p[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = verts + 8 * (size_t)(vert_base[a] + idx[3 * (size_t)g + k]);
            p[k] = dev_xform_point(M, v[k][0], v[k][1], v[k][2]);
            lo[0] = fminf(lo[0], p[k].x); lo[1] = fminf(lo[1], p[k].y); lo[2] = fminf(lo[2], p[k].z);
            hi[0] = fmaxf(hi[0], p[k].x); hi[1] = fmaxf(hi[1], p[k].y); hi[2] = fmaxf(hi[2], p[k].z);
            finite = finite && isfinite(p[k].x) && isfinite(p[k].y) && isfinite(p[k].z);
        }
        tri[3 * (size_t)g + 0] = make_float4(p[0].x, p[0].y, p[0].z, __int_as_float(g));
        tri[3 * (size_t)g + 1] = make_float4(p[1].x, p[1].y, p[1].z, __int_as_float(a));
        tri[3 * (size_t)g + 2] = make_float4(p[2].x, p[2].y, p[2].z, __int_as_float(t));
#pragma unroll
        for (int k = 0; k < 3; ++k) shade[4 * (size_t)g + k] = make_float4(v[k][5], v[k][6], v[k][7], v[k][3]);
        shade[4 * (size_t)g + 3] = make_float4(v[0][4], v[1][4], v[2][4], 0.0f);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
        for (int o = 16; o > 0; o >>= 1) {
            lo[q] = fminf(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], o));
            hi[q] = fmaxf(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], o));
        }
    const bool all_finite = __all_sync(0xffffffffu, finite);
    if ((threadIdx.x & 31) == 0) {
        if (lo[0] <= hi[0])
            for (int q = 0; q < 3; ++q) { atomicMin(&bounds[q], f2o(lo[q])); atomicMax(&bounds[3 + q], f2o(hi[q])); }
        if (!all_finite) atomicOr(&bounds[6], 1);
    }
}
}  // namespace

cudaError_t flatten_scene_gpu(const FlattenInput* meshes, int num_meshes, cudaStream_t stream, float4** tri_out,
                              float4** shade_out, float3* lo, float3* hi, bool* finite) {
    *tri_out = nullptr; *shade_out = nullptr;
    std::vector<int> tri_base(num_meshes), vert_base(num_meshes);
    std::vector<float> xf(16 * (size_t)num_meshes);
    size_t nt = 0, nv = 0;
    for (int i = 0; i < num_meshes; ++i) {
        tri_base[i] = (int)nt; vert_base[i] = (int)nv;
        nt += meshes[i].num_tris; nv += meshes[i].num_verts;
        memcpy(&xf[16 * (size_t)i], meshes[i].xform, sizeof(float) * 16);
    }
    float *d_verts = nullptr, *d_xf = nullptr;
    int *d_idx = nullptr, *d_tri_base = nullptr, *d_vert_base = nullptr, *d_bounds = nullptr;
    cudaError_t e = cudaMalloc(&d_verts, sizeof(float) * 8 * nv);
    if (e == cudaSuccess) e = cudaMalloc(&d_idx, sizeof(int) * 3 * nt);
    if (e == cudaSuccess) e = cudaMalloc(&d_tri_base, sizeof(int) * num_meshes);
    if (e == cudaSuccess) e = cudaMalloc(&d_vert_base, sizeof(int) * num_meshes);
    if (e == cudaSuccess) e = cudaMalloc(&d_xf, sizeof(float) * xf.size());
    if (e == cudaSuccess) e = cudaMalloc(&d_bounds, sizeof(int) * 8);
    if (e == cudaSuccess) e = cudaMalloc(tri_out, sizeof(float4) * 3 * nt);
    if (e == cudaSuccess) e = cudaMalloc(shade_out, sizeof(float4) * 4 * nt);
    for (int i = 0; i < num_meshes && e == cudaSuccess; ++i) {
        e = cudaMemcpyAsync(d_verts + 8 * (size_t)vert_base[i], meshes[i].verts, sizeof(float) * 8 * meshes[i].num_verts, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_idx + 3 * (size_t)tri_base[i], meshes[i].tris, sizeof(int) * 3 * meshes[i].num_tris, cudaMemcpyHostToDevice, stream);
    }
    const int init[8] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000, 0, 0};
    int h_bounds[8] = {};
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tri_base, tri_base.data(), sizeof(int) * num_meshes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_vert_base, vert_base.data(), sizeof(int) * num_meshes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_xf, xf.data(), sizeof(float) * xf.size(), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
        k_flatten<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(d_verts, d_idx, d_tri_base, d_vert_base, d_xf, num_meshes, (int)nt,
                                                                    *tri_out, *shade_out, d_bounds);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_bounds, d_bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_verts); cudaFree(d_idx); cudaFree(d_tri_base); cudaFree(d_vert_base); cudaFree(d_xf); cudaFree(d_bounds);
    if (e != cudaSuccess) { cudaFree(*tri_out); cudaFree(*shade_out); *tri_out = nullptr; *shade_out = nullptr; return e; }
    auto dec = [](int i) { const int b = i >= 0 ? i : i ^ 0x7fffffff; float f; memcpy(&f, &b, 4); return f; };
    *lo = make_float3(dec(h_bounds[0]), dec(h_bounds[1]), dec(h_bounds[2]));
    *hi = make_float3(dec(h_bounds[3]), dec(h_bounds[4]), dec(h_bounds[5]));
    *finite = h_bounds[6] == 0;
    return cudaSuccess;
}

// full-sweep SAH tree over cluster boxes (defined below, next to the host builder)
void build_top_level(const std::vector<float4>& cl_box, const std::vector<int>& cl_count, float eps,
                     std::vector<float4>* nodes, std::vector<int>* leaf_cluster_of_link, int* root_link, int* depth);

cudaError_t build_bvh(const float4* tri_in, const float4* shade_in, int n, float3 lo, float3 hi,
                      cudaStream_t stream, BvhResult* out, bool clustered_sah) {
    cudaError_t err = cudaSuccess;
    int *start_flag = nullptr, *count_at = nullptr, *box_at = nullptr, *cluster_id = nullptr, *cl_first = nullptr,
        *cl_count = nullptr, *cl_root = nullptr, *cl_used = nullptr, *cl_base = nullptr, *final_gid = nullptr;
    float4 *cl_box = nullptr, *tmp_nodes = nullptr;
    uint64_t *keys = nullptr, *keys_alt = nullptr;
    int *vals = nullptr, *vals_alt = nullptr, *parent_inner = nullptr, *parent_leaf = nullptr, *arrive = nullptr,
        *keep = nullptr, *new_index = nullptr, *d_scalars = nullptr;
    int2 *child = nullptr, *range = nullptr;
    float4 *box_lo = nullptr, *box_hi = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    const int T = 256;
    const int gb = (n + T - 1) / T;
    memset(out, 0, sizeof(*out));
    {
        float3 ext = make_float3(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z);
        float3 inv_ext = make_float3(ext.x > 0 ? 1.0f / ext.x : 0.0f, ext.y > 0 ? 1.0f / ext.y : 0.0f,
                                     ext.z > 0 ? 1.0f / ext.z : 0.0f);
        float mag = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))),
                          fmaxf(fmaxf(fabsf(lo.z), fabsf(hi.z)), sqrtf(ext.x * ext.x + ext.y * ext.y + ext.z * ext.z)));
        const float eps = 1e-5f * fmaxf(mag, 1e-20f);

        BVH_CK(cudaMalloc(&out->tris, sizeof(float4) * 3 * (size_t)n));
        BVH_CK(cudaMalloc(&out->shade, sizeof(float4) * 4 * (size_t)n));
        BVH_CK(cudaMalloc(&out->slot_of_gid, sizeof(int) * (size_t)n));
        BVH_CK(cudaMalloc(&keys, sizeof(uint64_t) * n));
        BVH_CK(cudaMalloc(&keys_alt, sizeof(uint64_t) * n));
        BVH_CK(cudaMalloc(&vals, sizeof(int) * n));
        BVH_CK(cudaMalloc(&vals_alt, sizeof(int) * n));
        k_morton<<<gb, T, 0, stream>>>(tri_in, n, lo, inv_ext, keys, vals);
        {
            cub::DoubleBuffer<uint64_t> dk(keys, keys_alt);
            cub::DoubleBuffer<int> dv(vals, vals_alt);
            BVH_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, 63, stream));
            BVH_CK(cudaMalloc(&tmp, tmp_bytes));
            BVH_CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, n, 0, 63, stream));
            if (dk.Current() != keys) { uint64_t* t = keys; keys = keys_alt; keys_alt = t; }
            if (dv.Current() != vals) { int* t = vals; vals = vals_alt; vals_alt = t; }
            cudaFree(tmp); tmp = nullptr;
        }
        if (n <= kLeafMax || !clustered_sah)
            k_gather<<<gb, T, 0, stream>>>(tri_in, shade_in, vals, n, out->tris, out->shade, out->slot_of_gid);
        if (n <= kLeafMax) {
            out->root_link = ~((0 << 2) | (n - 1));
            out->num_nodes = 0;
            out->max_depth = 1;
            BVH_CK(cudaMalloc(&out->nodes, sizeof(float4) * 4));   // never read
            BVH_CK(cudaMemsetAsync(out->nodes, 0, sizeof(float4) * 4, stream));
            BVH_CK(cudaStreamSynchronize(stream));
            goto done;
        }
        const int ni = n - 1;
        const int gi = (ni + T - 1) / T;
        BVH_CK(cudaMalloc(&child, sizeof(int2) * ni));
        BVH_CK(cudaMalloc(&range, sizeof(int2) * ni));
        BVH_CK(cudaMalloc(&parent_inner, sizeof(int) * ni));
        BVH_CK(cudaMalloc(&parent_leaf, sizeof(int) * n));
        BVH_CK(cudaMalloc(&arrive, sizeof(int) * ni));
        BVH_CK(cudaMalloc(&keep, sizeof(int) * ni));
        BVH_CK(cudaMalloc(&new_index, sizeof(int) * ni));
        BVH_CK(cudaMalloc(&box_lo, sizeof(float4) * (2 * (size_t)n - 1)));
        BVH_CK(cudaMalloc(&box_hi, sizeof(float4) * (2 * (size_t)n - 1)));
        BVH_CK(cudaMalloc(&d_scalars, sizeof(int) * 4));
        BVH_CK(cudaMemsetAsync(arrive, 0, sizeof(int) * ni, stream));
        BVH_CK(cudaMemsetAsync(d_scalars, 0, sizeof(int) * 4, stream));
        k_hierarchy<<<gi, T, 0, stream>>>(keys, n, child, range, parent_inner, parent_leaf);
        k_fit<<<gb, T, 0, stream>>>(tri_in, vals, n, child, parent_inner, parent_leaf, box_lo, box_hi, arrive, d_scalars);
        if (clustered_sah) {
            // ---- cut into clusters
            BVH_CK(cudaMalloc(&start_flag, sizeof(int) * n));
            BVH_CK(cudaMalloc(&count_at, sizeof(int) * n));
            BVH_CK(cudaMalloc(&box_at, sizeof(int) * n));
            BVH_CK(cudaMalloc(&cluster_id, sizeof(int) * n));
            BVH_CK(cudaMalloc(&final_gid, sizeof(int) * n));
            BVH_CK(cudaMemsetAsync(start_flag, 0, sizeof(int) * n, stream));
            k_cluster_mark<<<gb, T, 0, stream>>>(range, parent_inner, parent_leaf, n, start_flag, count_at, box_at);
            BVH_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, start_flag, cluster_id, n, stream));
            BVH_CK(cudaMalloc(&tmp, tmp_bytes));
            BVH_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, start_flag, cluster_id, n, stream));
            int h_last_flag = 0, h_last_id = 0;
            BVH_CK(cudaMemcpyAsync(&h_last_flag, start_flag + n - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaMemcpyAsync(&h_last_id, cluster_id + n - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaStreamSynchronize(stream));
            const int nc = h_last_id + h_last_flag;
            BVH_CK(cudaMalloc(&cl_first, sizeof(int) * nc));
            BVH_CK(cudaMalloc(&cl_count, sizeof(int) * nc));
            BVH_CK(cudaMalloc(&cl_root, sizeof(int) * nc));
            BVH_CK(cudaMalloc(&cl_used, sizeof(int) * nc));
            BVH_CK(cudaMalloc(&cl_base, sizeof(int) * nc));
            BVH_CK(cudaMalloc(&cl_box, sizeof(float4) * 2 * (size_t)nc));
            BVH_CK(cudaMalloc(&tmp_nodes, sizeof(float4) * 4 * (size_t)n));
            k_cluster_compact<<<gb, T, 0, stream>>>(start_flag, cluster_id, count_at, box_at, box_lo, box_hi, n, cl_first, cl_count, cl_box);
            std::vector<float4> h_box(2 * (size_t)nc);
            std::vector<int> h_count(nc), h_root(nc), h_base(nc);
            BVH_CK(cudaMemcpyAsync(h_box.data(), cl_box, sizeof(float4) * h_box.size(), cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaMemcpyAsync(h_count.data(), cl_count, sizeof(int) * nc, cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaStreamSynchronize(stream));
            // ---- per-cluster SAH subtrees on the GPU, the tree over the clusters on the host meanwhile
            BVH_CK(cudaFuncSetAttribute(k_sah_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SahShared)));
            k_sah_cluster<<<nc, kSahThreads, sizeof(SahShared), stream>>>(cl_first, cl_count, box_lo, box_hi, vals, n, eps, tmp_nodes,
                                                                          cl_root, cl_used, final_gid, d_scalars + 1);
            BVH_CK(cudaGetLastError());
            std::vector<float4> top_nodes;
            std::vector<int> top_leaf;
            int top_root = 0, top_depth = 0;
            build_top_level(h_box, h_count, eps, &top_nodes, &top_leaf, &top_root, &top_depth);
            const int num_top = (int)(top_nodes.size() / 4);
            cudaFree(tmp); tmp = nullptr;
            BVH_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cl_used, cl_base, nc, stream));
            BVH_CK(cudaMalloc(&tmp, tmp_bytes));
            BVH_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cl_used, cl_base, nc, stream));
            int h_last_used = 0, h_last_base = 0, h_cl_depth = 0;
            BVH_CK(cudaMemcpyAsync(&h_last_used, cl_used + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaMemcpyAsync(&h_last_base, cl_base + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaMemcpyAsync(&h_cl_depth, d_scalars + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaMemcpyAsync(h_root.data(), cl_root, sizeof(int) * nc, cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaMemcpyAsync(h_base.data(), cl_base, sizeof(int) * nc, cudaMemcpyDeviceToHost, stream));
            BVH_CK(cudaStreamSynchronize(stream));
            const int num_cluster_nodes = h_last_base + h_last_used;
            // a cluster's root is its local node 0, or a leaf link when the whole cluster is one leaf
            auto cluster_link = [&](int c) { return h_root[c] >= 0 ? num_top + h_base[c] + h_root[c] : h_root[c]; };
            for (int k = 0; k < num_top; ++k) {
                int l[2];
                memcpy(l, &top_nodes[4 * k + 3], sizeof(l));
                for (int q = 0; q < 2; ++q)
                    if (l[q] < 0) l[q] = cluster_link(top_leaf[(~l[q]) >> 2]);
                memcpy(&top_nodes[4 * k + 3], l, sizeof(l));
            }
            out->num_nodes = num_top + num_cluster_nodes;
            out->max_depth = top_depth + h_cl_depth;
            out->root_link = num_top > 0 ? top_root : cluster_link(top_leaf[(~top_root) >> 2]);
            BVH_CK(cudaMalloc(&out->nodes, sizeof(float4) * 4 * (size_t)std::max(out->num_nodes, 1)));
            if (num_top > 0)
                BVH_CK(cudaMemcpyAsync(out->nodes, top_nodes.data(), sizeof(float4) * top_nodes.size(), cudaMemcpyHostToDevice, stream));
            k_cluster_pack<<<gb, T, 0, stream>>>(tmp_nodes, start_flag, cluster_id, cl_first, cl_used, cl_base, n, num_top, out->nodes);
            k_gather<<<gb, T, 0, stream>>>(tri_in, shade_in, final_gid, n, out->tris, out->shade, out->slot_of_gid);
            BVH_CK(cudaGetLastError());
            BVH_CK(cudaStreamSynchronize(stream));
            goto done;
        }
        k_mark<<<gi, T, 0, stream>>>(range, n, keep);
        BVH_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, keep, new_index, ni, stream));
        BVH_CK(cudaMalloc(&tmp, tmp_bytes));
        BVH_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, keep, new_index, ni, stream));
        int h_last_keep = 0, h_last_idx = 0, h_depth = 0;
        BVH_CK(cudaMemcpyAsync(&h_last_keep, keep + ni - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
        BVH_CK(cudaMemcpyAsync(&h_last_idx, new_index + ni - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
        BVH_CK(cudaMemcpyAsync(&h_depth, d_scalars, sizeof(int), cudaMemcpyDeviceToHost, stream));
        BVH_CK(cudaStreamSynchronize(stream));
        out->num_nodes = h_last_idx + h_last_keep;
        out->max_depth = h_depth;
        BVH_CK(cudaMalloc(&out->nodes, sizeof(float4) * 4 * (size_t)out->num_nodes));
        k_emit<<<gi, T, 0, stream>>>(child, range, keep, new_index, box_lo, box_hi, n, eps, out->nodes);
        out->root_link = 0;   // Karras root is inner node 0; it is kept (n > kLeafMax) and scans to index 0
        BVH_CK(cudaGetLastError());
        BVH_CK(cudaStreamSynchronize(stream));
    }
done:
    cudaFree(keys); cudaFree(keys_alt); cudaFree(vals); cudaFree(vals_alt); cudaFree(parent_inner);
    cudaFree(parent_leaf); cudaFree(arrive); cudaFree(keep); cudaFree(new_index); cudaFree(d_scalars);
    cudaFree(child); cudaFree(range); cudaFree(box_lo); cudaFree(box_hi); cudaFree(tmp);
    cudaFree(start_flag); cudaFree(count_at); cudaFree(box_at); cudaFree(cluster_id); cudaFree(cl_first); cudaFree(cl_count);
    cudaFree(cl_root); cudaFree(cl_used); cudaFree(cl_base); cudaFree(final_gid); cudaFree(cl_box); cudaFree(tmp_nodes);
    if (err != cudaSuccess) free_bvh(out);
    return err;
}

// ---------------------------------------------------------------------------------------------
// Full-sweep SAH builder for small scenes (host).  The Cornell-box class of scenes (tens of
// large, mutually overlapping triangles) is where Morton-order LBVH trees are worst: measured on
// cboxuv+light, the LBVH visited 9.5 nodes and tested 8.4 triangles per ray.  Cost model: one
// node visit (two box tests) costs about as much as one watertight triangle test.
namespace {

struct HBox {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; } }
    void grow(const HBox& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx < 0 ? 0.0f : 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct SahBuilder {
    const std::vector<HBox>& tb;      // per-triangle boxes (scene order)
    std::vector<int> order;           // triangle ids, leaf order
    std::vector<float4> nodes;
    std::vector<float> right_area;
    float eps;
    int max_depth = 0;
    int leaf_max = kLeafMax;          // 1: every leaf is a single element (tree over clusters)
    const float* weight = nullptr;    // per-element cost in the SAH sum (null: 1)
    SahBuilder(const std::vector<HBox>& b, float e) : tb(b), eps(e) {
        order.resize(b.size());
        std::iota(order.begin(), order.end(), 0);
        right_area.resize(b.size());
    }
    HBox range_box(int a, int b) const {
        HBox r; r.reset();
        for (int i = a; i < b; ++i) r.grow(tb[order[i]]);
        return r;
    }
    // builds [a,b); returns its link
    int build(int a, int b, int depth) {
        max_depth = std::max(max_depth, depth);
        const int n = b - a;
        const HBox box = range_box(a, b);
        float best_cost = FLT_MAX;
        int best_axis = -1, best_split = -1;
        if (n > 1) {
            for (int axis = 0; axis < 3; ++axis) {
                std::sort(order.begin() + a, order.begin() + b, [&](int x, int y) {
                    const float cx = tb[x].lo[axis] + tb[x].hi[axis], cy = tb[y].lo[axis] + tb[y].hi[axis];
                    return cx < cy || (cx == cy && x < y);
                });
                HBox acc; acc.reset();
                float wtotal = 0.0f, wl = 0.0f;
                for (int i = b - 1; i > a; --i) { acc.grow(tb[order[i]]); right_area[i] = acc.area(); }
                for (int i = a; i < b; ++i) wtotal += weight ? weight[order[i]] : 1.0f;
                acc.reset();
                for (int i = a; i < b - 1; ++i) {
                    acc.grow(tb[order[i]]);
                    wl += weight ? weight[order[i]] : 1.0f;
                    const float c = acc.area() * wl + right_area[i + 1] * (wtotal - wl);
                    if (c < best_cost) { best_cost = c; best_axis = axis; best_split = i + 1; }
                }
            }
        }
        const float parent_area = std::max(box.area(), 1e-30f);
        static const float node_cost = getenv("ZDR_SAH_NODE_COST") ? (float)atof(getenv("ZDR_SAH_NODE_COST")) : 1.0f;
        const float split_cost = node_cost + best_cost / parent_area;     // node visit + expected children
        if (n <= leaf_max && (n == 1 || (float)n <= split_cost)) return ~((a << 2) | (n - 1));
        std::sort(order.begin() + a, order.begin() + b, [&](int x, int y) {
            const float cx = tb[x].lo[best_axis] + tb[x].hi[best_axis], cy = tb[y].lo[best_axis] + tb[y].hi[best_axis];
            return cx < cy || (cx == cy && x < y);
        });
        const int self = (int)(nodes.size() / 4);
        nodes.resize(nodes.size() + 4);
        const int l = build(a, best_split, depth + 1);
        const int r = build(best_split, b, depth + 1);
        const HBox bl = range_box(a, best_split), br = range_box(best_split, b);
        nodes[4 * self + 0] = make_float4(bl.lo[0] - eps, bl.hi[0] + eps, bl.lo[1] - eps, bl.hi[1] + eps);
        nodes[4 * self + 1] = make_float4(br.lo[0] - eps, br.hi[0] + eps, br.lo[1] - eps, br.hi[1] + eps);
        nodes[4 * self + 2] = make_float4(bl.lo[2] - eps, bl.hi[2] + eps, br.lo[2] - eps, br.hi[2] + eps);
        nodes[4 * self + 3] = make_float4(__int_as_float_host(l), __int_as_float_host(r), 0.f, 0.f);
        return self;
    }
    static float __int_as_float_host(int i) { float f; memcpy(&f, &i, 4); return f; }
};

float inflate_eps(float3 lo, float3 hi) {
    const float3 ext = make_float3(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z);
    const float mag = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))),
                            fmaxf(fmaxf(fabsf(lo.z), fabsf(hi.z)), sqrtf(ext.x * ext.x + ext.y * ext.y + ext.z * ext.z)));
    return 1e-5f * fmaxf(mag, 1e-20f);
}

}  // namespace

// Tree over the clusters of the large-scene build: elements are cluster boxes, every leaf is one cluster
// (link ~(position << 2); *leaf_cluster maps position -> cluster), an element's SAH weight is the expected
// cost of entering its subtree (~ log2 of its triangle count).
void build_top_level(const std::vector<float4>& cl_box, const std::vector<int>& cl_count, float eps,
                     std::vector<float4>* nodes, std::vector<int>* leaf_cluster, int* root_link, int* depth) {
    const int nc = (int)cl_count.size();
    std::vector<HBox> tb(nc);
    std::vector<float> w(nc);
    for (int c = 0; c < nc; ++c) {
        const float4 l = cl_box[2 * c], h = cl_box[2 * c + 1];
        tb[c].lo[0] = l.x; tb[c].lo[1] = l.y; tb[c].lo[2] = l.z;
        tb[c].hi[0] = h.x; tb[c].hi[1] = h.y; tb[c].hi[2] = h.z;
        w[c] = 1.0f + log2f((float)std::max(cl_count[c], 1));
    }
    SahBuilder b(tb, eps);
    b.leaf_max = 1;
    b.weight = w.data();
    *root_link = b.build(0, nc, 1);
    *nodes = std::move(b.nodes);
    *leaf_cluster = std::move(b.order);
    *depth = b.max_depth;
}

cudaError_t build_bvh_sah_host(const float4* tri, const float4* shade, int n, float3 lo, float3 hi, cudaStream_t stream,
                               BvhResult* out) {
    memset(out, 0, sizeof(*out));
    std::vector<HBox> tb(n);
    for (int i = 0; i < n; ++i) {
        tb[i].reset();
        for (int k = 0; k < 3; ++k) {
            const float4 v = tri[3 * i + k];
            const float p[3] = {v.x, v.y, v.z};
            for (int c = 0; c < 3; ++c) { tb[i].lo[c] = std::min(tb[i].lo[c], p[c]); tb[i].hi[c] = std::max(tb[i].hi[c], p[c]); }
        }
    }
    SahBuilder b(tb, inflate_eps(lo, hi));
    const int root = b.build(0, n, 1);
    if (b.nodes.empty()) b.nodes.resize(4, make_float4(0.f, 0.f, 0.f, 0.f));    // single-leaf scene: never read
    std::vector<float4> tris(3 * (size_t)n), sh(4 * (size_t)n);
    std::vector<int> slot_of_gid(n);
    for (int s = 0; s < n; ++s) {
        const int g = b.order[s];
        slot_of_gid[g] = s;
        for (int k = 0; k < 3; ++k) tris[3 * (size_t)s + k] = tri[3 * (size_t)g + k];
        for (int k = 0; k < 4; ++k) sh[4 * (size_t)s + k] = shade[4 * (size_t)g + k];
    }
    cudaError_t err = cudaMalloc(&out->nodes, sizeof(float4) * b.nodes.size());
    if (err == cudaSuccess) err = cudaMalloc(&out->tris, sizeof(float4) * tris.size());
    if (err == cudaSuccess) err = cudaMalloc(&out->shade, sizeof(float4) * sh.size());
    if (err == cudaSuccess) err = cudaMalloc(&out->slot_of_gid, sizeof(int) * (size_t)n);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out->nodes, b.nodes.data(), sizeof(float4) * b.nodes.size(), cudaMemcpyHostToDevice, stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out->tris, tris.data(), sizeof(float4) * tris.size(), cudaMemcpyHostToDevice, stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out->shade, sh.data(), sizeof(float4) * sh.size(), cudaMemcpyHostToDevice, stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out->slot_of_gid, slot_of_gid.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
    if (err != cudaSuccess) { free_bvh(out); return err; }
    out->root_link = root;
    out->num_nodes = root >= 0 ? (int)(b.nodes.size() / 4) : 0;
    out->max_depth = b.max_depth;
    return cudaSuccess;
}

void free_bvh(BvhResult* r) {
    if (!r) return;
    cudaFree(r->nodes); cudaFree(r->tris); cudaFree(r->shade); cudaFree(r->slot_of_gid);
    memset(r, 0, sizeof(*r));
}

}  // namespace zdr
