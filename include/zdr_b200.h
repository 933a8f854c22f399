/* zdr_b200.h — C ABI of the B200-native differentiable ray-tracer core.
 *
 * This is the drop-in boundary for the ONE hot path of 111116/zdr: the per-pixel render
 * kernels (forward + adjoint) that the reference launches through LuisaCompute from
 *   render.py:168-171  (forward kernel launch, 11-argument operator signature)
 *   render.py:193-196  (backward kernel launch, 12-argument operator signature, seed+1)
 * and the scene assembly those kernels read (render.py:73-128, :130-148, :150-156).
 * Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *  - plain C types only; no torch/ATen types cross this boundary.
 *  - every function returns 0 on success, a negative zdr_status on failure, never throws;
 *    zdr_last_error() returns a thread-local human readable message for the last failure.
 *  - *_dev pointers are CUDA device pointers owned by the caller (torch owns them in the
 *    Python surface, borrowed for the duration of the call, as the reference borrows them
 *    through DLPack at render.py:162-164,179-181).
 *  - *_host pointers are ordinary host memory.
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it, no device-wide sync
 *    (the reference does two full barriers per call, render.py:165,172 — we do none).  A scene
 *    owns its counters and the wavefront workspace: use it from one stream at a time.
 *  - there is NO CPU fallback: with no usable CUDA device every compute entry fails with
 *    ZDR_ERR_CUDA.
 */
#ifndef ZDR_B200_H
#define ZDR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZDR_ABI_VERSION 1

typedef enum zdr_status {
    ZDR_OK = 0,
    ZDR_ERR_ARG = -1,     /* bad argument (NULL, negative size, unknown integrator ...) */
    ZDR_ERR_STATE = -2,   /* call order violated (render before commit, tables not set ...) */
    ZDR_ERR_CUDA = -3,    /* CUDA runtime failure / no device */
    ZDR_ERR_LIMIT = -4    /* reference limits exceeded (render.py:114: > 10000 instances) */
} zdr_status;

/* integrator ids == the keys of the dict at render.py:65-69 */
typedef enum zdr_integrator {
    ZDR_COLLOCATED = 0,   /* collocated.py:10-57 */
    ZDR_DIRECT = 1,       /* direct.py:20-167    */
    ZDR_PATH = 2          /* prb.py:18-187       */
} zdr_integrator;

/* prb_mode (SURVEY.md §8a "PRB defects"): how the path adjoint is weighted */
#define ZDR_PRB_LITERAL 0    /* prb.py:162 cotangent beta/pdf*Le, no inst>0 break in bwd    */
#define ZDR_PRB_CORRECTED 1  /* textbook PRB: cotangent Le_remaining/f; fwd/bwd control flow equal */

/* Camera == luisa.StructType(fov, origin, target, up) at render.py:28; plus every scalar the
 * reference passes to the kernel (render.py:168-171) and the module-level constants that
 * LuisaCompute bakes in at JIT time (direct.py:14, prb.py:15-16, interaction.py:33). */
typedef struct zdr_params {
    float cam_fov;            /* full horizontal field of view, radians (camera.py:15) */
    float cam_origin[3];
    float cam_target[3];
    float cam_up[3];
    int32_t width, height;    /* dispatch_size=res (render.py:171): res=(W,H) */
    int32_t tex_res_x;        /* texture_res.x == material.shape[0] (render.py:161,170) */
    int32_t tex_res_y;        /* texture_res.y == material.shape[1] */
    int32_t spp;
    int32_t seed;             /* forward: seed; backward: caller passes seed+1 (render.py:196) */
    int32_t use_tent_filter;  /* render.py:71 */
    int32_t max_depth;        /* prb.py:15 (16 upstream; BASELINE configs use 8 and 16) */
    int32_t rr_depth;         /* prb.py:16 */
    int32_t use_mis;          /* direct.py:14 */
    int32_t bilinear;         /* interaction.py:33 */
    int32_t prb_mode;         /* ZDR_PRB_* */
    /* data-parallel shard of the dispatch grid (multi-GPU): rows
     * [row0,row1) of the image and samples [it0,it1) of every pixel. Full job: 0,H,0,spp. */
    int32_t row0, row1;
    int32_t it0, it1;
    /* second sharding level used by the multi-GPU path: the 8-row tiles of [row0,row1) are
     * dealt round-robin, shard `shard_index` of `shard_count` renders tiles i with
     * i % shard_count == shard_index (load balance; every pixel keeps its whole sample loop on
     * one GPU, so the image is bit-identical to the unsharded one). 0 or 1 = no interleave. */
    int32_t shard_index, shard_count;
    int32_t flags;            /* ZDR_FLAG_* */
    /* texture sets in the material buffer (0 or 1: one, the reference's layout).  K > 1: material / d_material are
     * (K, tex_res_x, tex_res_y, 4), and zdr_scene_set_material_sets says which set shades which instance. */
    int32_t num_material_sets;
} zdr_params;
#define ZDR_FLAG_STATS 1      /* also count BVH nodes visited / triangles tested (slower) */
/* texture-gradient scatter (replaces the atomic_fetch_add of interaction.py:67-70). Default (neither
 * bit): chosen per launch from the uv extent of instance 0 in texels — warp-aggregated when the
 * textured object covers fewer than 64 texels in u or v (uniform / uv-less meshes), plain otherwise. */
#define ZDR_FLAG_SCATTER_AGGREGATE 2
#define ZDR_FLAG_SCATTER_PLAIN 4
/* on top of the warp aggregation: every CTA adds into a duplicate of the gradient buffer in its SHARED memory (an
 * open-addressing table of 256 texels) and folds it into d_material once, when it ends — the "per-SM duplicated gradient
 * buffers" of the README's to-do (README.md:21); with the persistent schedule the table lives for the whole pass.  Cuts the
 * global reductions per scatter to ~1e-4 (C2) but is NOT the default: after warp aggregation the L2 reductions are off the
 * critical path, and the shared-memory atomics cost more than they save (profiles/r02e). */
#define ZDR_FLAG_SCATTER_BLOCK 128
/* schedule of the path integrator / PRB adjoint. Default (no bit), by camera samples per launch: below 2^14 the
 * one-kernel bounce state machine; up to 3 * 2^20 the persistent shared-memory schedule (ZDR_FLAG_PERSISTENT); above
 * that the wavefront kernels (path state in HBM, compacted live lists). All produce bit-identical images. */
#define ZDR_FLAG_WAVEFRONT 8
#define ZDR_FLAG_MEGAKERNEL 16
/* persistent schedule: one resident CTA per SM, path state and ray queues in SHARED memory, the stages as phases of one
 * kernel per pass (persistent.cu). */
#define ZDR_FLAG_PERSISTENT 64
/* MIS weight of a BSDF-sampled emitter hit in a scene that has BOTH mesh lights and an envmap: the reference's
 * sample_light_pdf divides by light_count only (light.py:89) while sample_light picks among
 * env_count + light_count lights (light.py:26), so the two densities disagree by (n_mesh + 1) / n_mesh.
 * Default: as the reference.  With this bit sample_light_pdf uses the same n as sample_light. */
#define ZDR_FLAG_LIGHT_PDF_COUNTS_ENV 32
/* "fast" build of the render kernels: approximate division / square root / sin / cos (2 ulp), floating-point contraction,
 * denormals flushed — in the SHADING arithmetic only.  The sampler (pmj02bn.py) and ray / triangle intersection (primitive ids,
 * t, u, v) are the same explicit IEEE operations in both builds and stay bit-exact.  Default (bit clear): every operation is the
 * IEEE operation the numeric contract names and forward images are bit-identical to the CPU oracle's.  With the bit set the
 * images / gradients agree with the exact build to ~1e-6 / ~1e-4 relative L2 where the material is well conditioned
 * (roughness >= 0.3, the bar under which the oracle itself is pinned to the reference's sources) and deviate on very smooth
 * texels, where binary32 GGX is ill conditioned and any two float32 implementations differ.  1.1-1.4x faster. */
#define ZDR_FLAG_FAST_MATH 256

typedef struct zdr_scene zdr_scene;

/* --- library-wide ---------------------------------------------------------------------- */
int zdr_abi_version(void);
const char* zdr_last_error(void);

/* Sampler tables: replaces the module-level uploads at pmj02bn.py:9-16.
 * pmj:  [sets][n][2] uint32 (pbrt-v4 PMJ02bnSamples), converted as pmj02bn.py:9 does
 *       (round-to-nearest u32->f32, times 2^-32).
 * bn:   [ntex][res][res] uint16 (pbrt-v4 BlueNoiseTextures), /65536 (pmj02bn.py:14).
 * Uploaded to `device`; must be called once per device before any render. */
int zdr_set_sampler_tables(int device, const uint32_t* pmj_host, int sets, int n,
                           const uint16_t* bn_host, int ntex, int res);

/* --- scene assembly: replaces Scene.load_geometry (render.py:73-128) ------------------- */
int zdr_scene_create(zdr_scene** out, int device);           /* luisa.Accel()+BindlessArray(), render.py:74-75 */
void zdr_scene_destroy(zdr_scene* s);
/* accel.add(vertex_buffer, triangle_buffer, transform=) + heap.emplace(2i / 2i+1), render.py:96-111.
 * verts: nv rows of the 32-byte Vertex (vertex.py:4): px py pz  u v  nx ny nz.
 * tris: nt*3 int32 indices (load_obj.py:63-68). xform: float4x4, column-major (16 floats).
 * Instance id = order of calls (render.py:81). */
int zdr_scene_add_mesh(zdr_scene* s, const float* verts_nx8_host, int nv,
                       const int32_t* tris_host, int nt,
                       const float xform_colmajor[16], const float emission[3]);
/* accel.update(); heap.update() (render.py:127-128): flattens the instances to world space and builds the BVH —
 * up to 16 K triangles on the host (full-sweep SAH), above that on the GPU (flatten kernel, Morton clusters,
 * per-cluster binned SAH, SAH tree over the clusters).  May be called again after more zdr_scene_add_mesh calls. */
int zdr_scene_commit(zdr_scene* s);
/* Back to the state after zdr_scene_create: Scene.load_geometry builds a fresh luisa.Accel() and BindlessArray()
 * on every call (render.py:74-76), so a second load_geometry replaces the instances instead of appending. */
int zdr_scene_clear(zdr_scene* s);
/* Scene.update_lights (render.py:130-148): n must equal the instance count. */
int zdr_scene_set_emissions(zdr_scene* s, const float* emission_nx3_host, int n);
/* Scene.add_envmap -> load_envmap (render.py:150-156, envmap.py:116-202). rgba float, h rows. */
int zdr_scene_set_envmap(zdr_scene* s, const float* rgba_host, int w, int h, int compensate_mis);
/* "Polymorphic" materials — several texture sets, the item the reference lists as not implemented (README.md:5: "currently
 * only one set of material").  set_of_instance[i] >= 0: instance i is shaded with that texture set of the material buffer;
 * -1: the instance is a light / terminal surface.  The default (instance 0 -> 0, all others -1) is the reference's rule
 * "treat everything as light other than the first object" (direct.py:30, prb.py:45). n must equal the instance count. */
int zdr_scene_set_material_sets(zdr_scene* s, const int32_t* set_of_instance_host, int n);
/* Workspace of the path-integrator schedules (no counterpart upstream: LuisaCompute's per-pixel kernels keep their paths in
 * registers).  HBM wavefront: 152 B per resident path, by default as many samples of every pixel as fit a quarter of the free
 * memory (up to 2^28 paths = 41 GB); persistent schedule: 16 B per camera sample of a forward batch (up to 2 GiB).  Allocated
 * with cudaMalloc on first use and kept, unless the caller
 *   - bounds it:            zdr_scene_set_workspace_limit(bytes)   (0 = default policy; batches shrink to fit),
 *   - provides it:          zdr_scene_set_workspace(ptr, bytes)    (256-byte aligned device buffer the CALLER owns, e.g. a torch
 *                           tensor, so that the caller's allocator accounts for it; NULL, 0 = back to the library's own),
 *   - or releases it:       zdr_scene_trim()                       (re-allocated on demand).
 * A launch that does not fit fails with ZDR_ERR_LIMIT. */
int zdr_scene_set_workspace_limit(zdr_scene* s, uint64_t bytes);
int zdr_scene_set_workspace(zdr_scene* s, void* workspace_dev, uint64_t bytes);
int zdr_scene_trim(zdr_scene* s);
uint64_t zdr_scene_workspace_bytes(const zdr_scene* s);
int zdr_scene_instance_count(const zdr_scene* s);
int zdr_scene_light_count(const zdr_scene* s);
int zdr_scene_triangle_count(const zdr_scene* s);

/* --- the hot path: replaces the kernel launches at render.py:168-171 / :193-196 -------- */
/* image_dev: (H,W,4) f32, written for rows [row0,row1). material_dev: (tex_res_x*tex_res_y*4) f32.
 * stream: the caller's cudaStream_t; the call only enqueues (no synchronisation, no allocation once the scene's
 * workspace holds the job).  Large path-integrator jobs fork onto a stream owned by the scene and join back with
 * events before returning, so ordering and CUDA-graph capture on `stream` behave as for a single-stream call. */
int zdr_render_forward(zdr_scene* s, int integrator, const zdr_params* p,
                       const float* material_dev, float* image_dev, void* stream);
/* d_image_dev: (H,W,4) f32 cotangent; d_material_dev: same shape as material, ACCUMULATED into
 * (the reference zero-fills it first at render.py:220; callers here do the same). */
int zdr_render_backward(zdr_scene* s, int integrator, const zdr_params* p,
                        const float* material_dev, const float* d_image_dev,
                        float* d_material_dev, void* stream);

/* Scene.render_duvdxy (render.py:243-257, kernel uvgrad.py:75-98): screen-to-texture Jacobian
 * (du/dx, dv/dx, du/dy, dv/dy) per pixel, averaged over spp; image_dev: (H,W,4) f32. Not differentiable.
 * Pixel offsets are drawn with the render kernels' pmj02bn sampler (the reference uses LuisaCompute's
 * built-in LCG here, which is not reproducible without LuisaCompute). */
int zdr_render_duvdxy(zdr_scene* s, const zdr_params* p, float* image_dev, void* stream);

/* Same operators with HOST buffers: the call a non-torch embedder (or bench.py's e2e leg)
 * makes; performs H2D of the inputs and D2H of the result inside the call, then syncs. */
int zdr_render_forward_host(zdr_scene* s, int integrator, const zdr_params* p,
                            const float* material_host, float* image_host);
int zdr_render_backward_host(zdr_scene* s, int integrator, const zdr_params* p,
                             const float* material_host, const float* d_image_host,
                             float* d_material_host);

/* --- instrumentation ------------------------------------------------------------------ */
/* counters accumulated by the render kernels since the last reset (flushed once per block):
 * [0] closest-hit rays  [1] any-hit rays  [2] BVH nodes visited  [3] triangles tested
 * [4] gradient scatters (write_bsdf_grad calls)  [5] red.v4 issued  [6] shaded vertices
 * [7] camera samples.  Blocks until the scene's last launch finished. */
#define ZDR_NUM_COUNTERS 8
int zdr_get_counters(zdr_scene* s, uint64_t out[ZDR_NUM_COUNTERS]);
int zdr_reset_counters(zdr_scene* s);
/* number of CUDA kernels this library has launched so far in this process (render + probe kernels) */
uint64_t zdr_kernel_launches(void);
/* profile of the scene's last persistent-schedule launch (only filled when the environment has ZDR_PW_PROFILE=1): SM cycles
 * summed over the CTAs for [0] claim + regenerate [1] extend [2] shade A [3] shadow [4] shade B, [5] rounds, [6] live
 * paths summed over rounds, [7] total cycles.  Blocks until the launch finished. */
int zdr_debug_phase_cycles(zdr_scene* s, uint64_t out[8]);

/* --- probes (parity tests only; small n). All pointers HOST. ---------------------------- */
/* closest hit for n rays (o.xyz, d.xyz, tmin, tmax) -> inst, prim, bary u, v, t (inst=-1: miss) */
int zdr_probe_trace_closest(zdr_scene* s, const float* rays_nx8_host, int n,
                            int32_t* inst_prim_nx2_host, float* uvt_nx3_host);
int zdr_probe_trace_any(zdr_scene* s, const float* rays_nx8_host, int n, int32_t* occluded_host);
/* sampler: for each of n (px,py,seed,spp,sample_index) run the draw program `prog`
 * (string of '1' = next(), '2' = next2f()) and write one float per '1', two per '2'. */
int zdr_probe_sampler(int device, const int32_t* state_nx5_host, int n, const char* prog,
                      float* out_host);
/* per-sample (unclamped) radiance of the forward estimator: out (H,W,spp,3) f32 */
int zdr_probe_render_samples(zdr_scene* s, int integrator, const zdr_params* p,
                             const float* material_host, float* out_host);
/* ggx_brdf value and its analytic adjoint wrt (diffuse.xyz, roughness), cotangent g:
 * in  n x 13: wo.xyz wi.xyz diffuse.xyz roughness g.xyz ; out n x 7: f.xyz dmat.xyzw */
int zdr_probe_ggx(int device, const float* in_nx13_host, int n, float* out_nx7_host);

/* --- scene ingest (host only, needs no device) ---------------------------------------------
 * Wavefront OBJ -> the (verts N x 8, tris T x 3) arrays zdr_scene_add_mesh takes.  Replaces read_obj +
 * concat_triangles (load_obj.py:1-68) and the recompute_normal call of render.py:101-103, with the same vertex
 * numbering (first use of each distinct position/uv/normal triple), primitive order (fan triangulation in face
 * order), defaults (missing vt = (0,0); missing vn = recomputed area-weighted normals).  *verts_out / *tris_out
 * are malloc'ed; release them with zdr_obj_free.  Errors: ZDR_ERR_ARG, text in zdr_obj_last_error(). */
int zdr_obj_load(const char* path, float** verts_out, int* num_verts_out, int32_t** tris_out, int* num_tris_out);
void zdr_obj_free(void* p);
const char* zdr_obj_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ZDR_B200_H */
