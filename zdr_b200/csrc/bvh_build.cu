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

constexpr int kLeafMax = 4;

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

#define BVH_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

}  // namespace

cudaError_t build_bvh(const float4* tri_in, const float4* shade_in, int n, float3 lo, float3 hi,
                      cudaStream_t stream, BvhResult* out) {
    cudaError_t err = cudaSuccess;
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
                for (int i = b - 1; i > a; --i) { acc.grow(tb[order[i]]); right_area[i] = acc.area(); }
                acc.reset();
                for (int i = a; i < b - 1; ++i) {
                    acc.grow(tb[order[i]]);
                    const int nl = i - a + 1, nr = n - nl;
                    const float c = acc.area() * (float)nl + right_area[i + 1] * (float)nr;
                    if (c < best_cost) { best_cost = c; best_axis = axis; best_split = i + 1; }
                }
            }
        }
        const float parent_area = std::max(box.area(), 1e-30f);
        static const float node_cost = getenv("ZDR_SAH_NODE_COST") ? (float)atof(getenv("ZDR_SAH_NODE_COST")) : 1.0f;
        const float split_cost = node_cost + best_cost / parent_area;     // node visit + expected children
        if (n <= 4 && (n == 1 || (float)n <= split_cost)) return ~((a << 2) | (n - 1));
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
