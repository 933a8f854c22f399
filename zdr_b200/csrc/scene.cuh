// scene.cuh — device-side view of a committed scene (data layout in HBM).
//
// The reference reaches geometry through a bindless heap (render.py:110-126): slot 2i = index
// buffer, 2i+1 = 32-byte Vertex buffer, 23333 emission, 23334 light list, 23335 triangle
// counts, plus a per-instance float4x4 fetched from the accel (interaction.py:18).  Here the
// instances are flattened once at commit into world-space triangles stored in BVH-leaf order
// as 16-byte aligned float4 records, so a hit needs no index/vertex/matrix indirection:
//
//   tris  [3*slot+0..2] : v0.xyz|gid   v1.xyz|inst   v2.xyz|prim      (48 B, traversal + p0..p2)
//   shade [4*slot+0..3] : n0.xyz|u0    n1.xyz|u1     n2.xyz|u2    (v0,v1,v2,-)   (64 B)
//   nodes [4*node+0..3] : BVH2, both child boxes + two links          (64 B, see bvh.cuh)
//
// Vertex normals stay in object space and the per-instance normal matrix
// inverse(transpose(M3x3)) is applied after interpolation exactly as interaction.py:28 does.
#pragma once
#include "zdr_math.cuh"

namespace zdr {

struct InstData {            // 64 B
    float N[9];              // inverse(transpose(M3x3)), column-major
    float emission[3];       // heap slot 23333 (render.py:120)
    int tri_base;            // first gid of the instance
    int tri_count;           // heap slot 23335 (render.py:126)
    int material_set;        // texture set that shades this instance, or -1: a light / terminal surface ("everything but the
                             // first object", direct.py:30, prb.py:45).  Default: instance 0 -> set 0, all others -1 (upstream)
    int pad;
};

struct SceneView {
    const float4* __restrict__ nodes;
    const float4* __restrict__ tris;
    const float4* __restrict__ shade;
    const int* __restrict__ slot_of_gid;
    const InstData* __restrict__ inst;
    const int* __restrict__ lights;      // heap slot 23334 (render.py:121)
    int root_link;                       // link of the root (leaf link if the scene is one leaf)
    int num_tris, num_nodes, num_inst;
    int light_count, env_count;
    int pdf_env_count;                   // added to light_count by sample_light_pdf: 0 (light.py:89) or env_count
    // sampler tables (pmj02bn.py:9-16)
    const float2* __restrict__ pmj;
    const float* __restrict__ bn;
    int pmj_sets, pmj_n, bn_ntex, bn_res;
    // envmap (envmap.py:116-202): heap slots 23332 (texture), 23330 (alias), 23331 (pdf)
    const float4* __restrict__ env_tex;
    const float2* __restrict__ env_alias;   // (prob, alias bits)
    const float* __restrict__ env_pdf;
    int env_w, env_h;
};

// Loads of scene records (nodes / tris / shade).  Default: read-only global loads.  A translation unit that may stage a small
// scene in shared memory (persistent.cu) defines ZDR_SCENE_GENERIC_LOADS before including the headers: the same
// device functions then work on either address space (ld.global.nc on a shared-memory address would fault).
#ifdef ZDR_SCENE_GENERIC_LOADS
ZDR_DEV float4 ld_scene(const float4* p) { return *p; }
ZDR_DEV float ld_scene(const float* p) { return *p; }
#else
ZDR_DEV float4 ld_scene(const float4* p) { return __ldg(p); }
ZDR_DEV float ld_scene(const float* p) { return __ldg(p); }
#endif

struct Ray { float3 o, d; float tmin, tmax; };
struct Hit { int inst, prim, slot; float u, v, t; };   // inst < 0: miss; (u,v) weights of v1, v2

// kernel-side counters (zdr_get_counters)
enum { C_CLOSEST = 0, C_ANY, C_NODES, C_TRIS, C_SCATTER, C_RED, C_SHADED, C_SAMPLES, C_NUM };

}  // namespace zdr
