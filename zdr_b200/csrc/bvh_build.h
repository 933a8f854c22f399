// bvh_build.h — interface of the BVH builders (bvh_build.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zdr {

struct BvhResult {
    float4* nodes;        // 4 float4 per inner node (traverse.cuh layout)
    float4* tris;         // 3 float4 per triangle, leaf order
    float4* shade;        // 4 float4 per triangle, leaf order
    int* slot_of_gid;     // scene-order triangle id -> leaf-order slot
    int root_link;
    int num_nodes;
    int max_depth;        // upper bound of the tree depth (sizes the traversal stack check)
};

// tri_in / shade_in: device arrays in scene order (3 resp. 4 float4 per triangle).
// lo/hi: bounds of all vertices. Synchronises `stream` before returning.
// clustered_sah: Morton order only groups the triangles into clusters; SAH trees inside the clusters (GPU,
// one block each) and over them (host).  false: plain LBVH (the radix tree itself).
cudaError_t build_bvh(const float4* tri_in, const float4* shade_in, int n, float3 lo, float3 hi,
                      cudaStream_t stream, BvhResult* out, bool clustered_sah);
void free_bvh(BvhResult* r);

// Large scenes: instances -> world-space triangle / shading records on the GPU (scene order), same bits as the
// host flatten of zdr_scene_commit.  Host pointers are borrowed for the call.  *tri_out / *shade_out are
// cudaMalloc'ed (caller frees); lo/hi = bounds of all vertices; *finite = no NaN/Inf position.
struct FlattenInput {
    const float* verts;      // num_verts x 8 (vertex.py:4)
    const int32_t* tris;     // num_tris x 3, mesh-local vertex ids
    int num_verts, num_tris;
    const float* xform;      // 16 floats, column-major
};
cudaError_t flatten_scene_gpu(const FlattenInput* meshes, int num_meshes, cudaStream_t stream, float4** tri_out,
                              float4** shade_out, float3* lo, float3* hi, bool* finite);

// Scenes up to this many triangles get a full-sweep SAH tree (built on the host in microseconds to
// milliseconds, uploaded once); larger ones the GPU build.  tri_host / shade_host: scene order.
constexpr int kSahMaxTris = 1 << 14;
cudaError_t build_bvh_sah_host(const float4* tri_host, const float4* shade_host, int n, float3 lo, float3 hi,
                               cudaStream_t stream, BvhResult* out);

}  // namespace zdr
