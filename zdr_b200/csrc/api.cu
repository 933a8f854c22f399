// api.cu — the C ABI of include/zdr_b200.h: scene assembly (render.py:73-148), sampler-table
// upload (pmj02bn.py:9-16), the forward / adjoint operator launches (render.py:168-171,
// 193-196) and the parity probes.  Nothing here has a CPU fallback: every compute entry needs a
// CUDA device and returns ZDR_ERR_CUDA without one.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "bvh_build.h"
#include "envmap.cuh"
#include "render.cuh"

#define ZDR_EXPORT extern "C" __attribute__((visibility("default")))

using namespace zdr;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ZDR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ---- sampler tables, one immutable copy per device (pmj02bn.py:9-16) ---------------------------
constexpr int kMaxDevices = 64;
struct Tables {
    float2* pmj = nullptr;
    float* bn = nullptr;
    int sets = 0, n = 0, ntex = 0, res = 0;
};
Tables g_tables[kMaxDevices];
std::mutex g_tables_mu;

struct Mesh {
    std::vector<float> verts;     // nv x 8 (vertex.py:4)
    std::vector<int32_t> tris;    // nt x 3
    float M[16];
    float N[9];
    float emission[3];
    int tri_base = 0;
    int material_set = -2;        // -2: default (instance 0 -> set 0, others -1 = upstream); else zdr_scene_set_material_sets
};

// inverse(transpose(M3x3)), interaction.py:28 — cofactor expansion, plain binary32 operations
void normal_matrix(const float* M, float* N) {
    const float a00 = M[0], a01 = M[1], a02 = M[2];
    const float a10 = M[4], a11 = M[5], a12 = M[6];
    const float a20 = M[8], a21 = M[9], a22 = M[10];
    const float c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const float c10 = a02 * a21 - a01 * a22, c11 = a00 * a22 - a02 * a20, c12 = a01 * a20 - a00 * a21;
    const float c20 = a01 * a12 - a02 * a11, c21 = a02 * a10 - a00 * a12, c22 = a00 * a11 - a01 * a10;
    const float det = a00 * c00 + a01 * c01 + a02 * c02;
    const float id = 1.0f / det;
    N[0] = c00 * id; N[1] = c01 * id; N[2] = c02 * id;
    N[3] = c10 * id; N[4] = c11 * id; N[5] = c12 * id;
    N[6] = c20 * id; N[7] = c21 * id; N[8] = c22 * id;
}

template <typename T>
struct DevBuf {   // grow-only device scratch; `external`: a buffer the caller owns (zdr_scene_set_workspace), never reallocated
    T* p = nullptr;
    size_t cap = 0;
    bool external = false;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (external) return cudaErrorMemoryAllocation;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, sizeof(T) * n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p && !external) cudaFree(p); p = nullptr; cap = 0; external = false; }
};

}  // namespace

struct zdr_scene {
    int device = 0;
    std::vector<Mesh> meshes;
    bool committed = false;
    BvhResult bvh = {};
    InstData* d_inst = nullptr;
    int* d_lights = nullptr;
    int light_count = 0;
    int num_tris = 0;
    EnvMap env = {};
    unsigned long long* d_counters = nullptr;
    float uv_lo[2] = {0.f, 0.f}, uv_hi[2] = {0.f, 0.f};   // uv bounds of instance 0, the textured object (direct.py:30)
    float3 bounds_lo = {0.f, 0.f, 0.f}, bounds_hi = {1.f, 1.f, 1.f};      // world-space bounds of the committed geometry
    cudaStream_t own_stream = nullptr;       // for the *_host entry points and the probes
    cudaStream_t side_streams[kWfMaxSets - 1] = {};      // further lanes of the wavefront schedule (launch_path_wavefront)
    cudaEvent_t wf_events[kWfMaxSets + 2] = {};
    DevBuf<float> s_mat, s_img, s_dimg, s_dmat, s_misc, s_misc2;
    DevBuf<char> s_wf;                       // workspace of the path schedules: wavefront path state (render.cuh WfState) / persistent sample buffer
    size_t ws_limit = 0;                     // zdr_scene_set_workspace_limit: upper bound of that workspace in bytes (0: default policy)
    int num_sms = 148;
    DevBuf<int> s_int, s_int2;
};

namespace {

// Every entry point makes the scene's device current for its own duration only and puts the caller's
// device back on return: torch reads its current device with cudaGetDevice, so a Scene on cuda:1 must not
// silently move a caller that works on cuda:0.
struct DevGuard {
    int prev = -1;
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
int use_device(int device, DevGuard& g) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ZDR_ERR_CUDA, "no CUDA device (%s); zdr_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count || device >= kMaxDevices) return fail(ZDR_ERR_ARG, "bad device %d", device);
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
    if (cur != device) {
        CK(cudaSetDevice(device));
        if (g.prev < 0) g.prev = cur;
    }
    return ZDR_OK;
}

// uv bounds of the textured instances (instance 0 upstream, direct.py:30): decides the default scatter strategy
void update_uv_bounds(zdr_scene* s) {
    s->uv_lo[0] = s->uv_lo[1] = INFINITY; s->uv_hi[0] = s->uv_hi[1] = -INFINITY;
    for (size_t i = 0; i < s->meshes.size(); ++i) {
        const Mesh& m = s->meshes[i];
        const int ms = m.material_set == -2 ? (i == 0 ? 0 : -1) : m.material_set;
        if (ms < 0 && i != 0) continue;
        for (size_t v = 0; v < m.verts.size() / 8; ++v)
            for (int k = 0; k < 2; ++k) {
                s->uv_lo[k] = fminf(s->uv_lo[k], m.verts[8 * v + 3 + k]);
                s->uv_hi[k] = fmaxf(s->uv_hi[k], m.verts[8 * v + 3 + k]);
            }
    }
}

int upload_lights(zdr_scene* s) {   // render.py:116-126, :141-148
    const int n = (int)s->meshes.size();
    std::vector<InstData> inst(n);
    std::vector<int> lights;
    for (int i = 0; i < n; ++i) {
        const Mesh& m = s->meshes[i];
        memset(&inst[i], 0, sizeof(InstData));
        memcpy(inst[i].N, m.N, sizeof(m.N));
        memcpy(inst[i].emission, m.emission, sizeof(m.emission));
        inst[i].tri_base = m.tri_base;
        inst[i].tri_count = (int)(m.tris.size() / 3);
        inst[i].material_set = m.material_set == -2 ? (i == 0 ? 0 : -1) : m.material_set;
        if (m.emission[0] > 0.f || m.emission[1] > 0.f || m.emission[2] > 0.f) lights.push_back(i);   // render.py:89-90
    }
    s->light_count = (int)lights.size();
    if (lights.empty()) lights.push_back(0);
    lights.resize(n > 0 ? n : 1, 0);
    // the kernels of an in-flight launch may still read the old tables
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(s->d_inst, inst.data(), sizeof(InstData) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_lights, lights.data(), sizeof(int) * lights.size(), cudaMemcpyHostToDevice));
    return ZDR_OK;
}

int scene_view(zdr_scene* s, SceneView* v) {
    if (!s->committed) return fail(ZDR_ERR_STATE, "scene not committed (zdr_scene_commit)");
    const Tables& t = g_tables[s->device];
    if (!t.pmj || !t.bn) return fail(ZDR_ERR_STATE, "sampler tables not set for device %d (zdr_set_sampler_tables)", s->device);
    memset(v, 0, sizeof(*v));
    v->nodes = s->bvh.nodes; v->tris = s->bvh.tris; v->shade = s->bvh.shade; v->slot_of_gid = s->bvh.slot_of_gid;
    v->inst = s->d_inst; v->lights = s->d_lights;
    v->root_link = s->bvh.root_link;
    v->num_tris = s->num_tris; v->num_nodes = s->bvh.num_nodes; v->num_inst = (int)s->meshes.size();
    v->light_count = s->light_count;
    v->env_count = s->env.valid ? 1 : 0;
    v->pmj = t.pmj; v->bn = t.bn; v->pmj_sets = t.sets; v->pmj_n = t.n; v->bn_ntex = t.ntex; v->bn_res = t.res;
    v->env_tex = s->env.tex; v->env_alias = s->env.alias; v->env_pdf = s->env.pdf; v->env_w = s->env.w; v->env_h = s->env.h;
    return ZDR_OK;
}

int check_params(const zdr_scene* s, int integrator, const zdr_params* p) {
    if (!s || !p) return fail(ZDR_ERR_ARG, "null scene/params");
    if (integrator < ZDR_COLLOCATED || integrator > ZDR_PATH) return fail(ZDR_ERR_ARG, "unknown integrator %d", integrator);
    if (p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->tex_res_x <= 0 || p->tex_res_y <= 0)
        return fail(ZDR_ERR_ARG, "non-positive width/height/spp/texture size");
    if (p->row0 < 0 || p->row1 > p->height || p->row0 > p->row1 || p->it0 < 0 || p->it1 > p->spp || p->it0 > p->it1)
        return fail(ZDR_ERR_ARG, "bad shard rows [%d,%d) its [%d,%d)", p->row0, p->row1, p->it0, p->it1);
    if (p->max_depth < 0 || p->rr_depth < 0) return fail(ZDR_ERR_ARG, "negative depth");
    if (p->shard_count > 1 && (p->shard_index < 0 || p->shard_index >= p->shard_count))
        return fail(ZDR_ERR_ARG, "bad shard %d of %d", p->shard_index, p->shard_count);
    return ZDR_OK;
}

int make_args(zdr_scene* s, const zdr_params* p, RenderArgs* a) {
    memset(a, 0, sizeof(*a));
    int rc = scene_view(s, &a->sc);
    if (rc) return rc;
    // camera.py:7-10, evaluated once per launch (pure function of the camera)
    const float3 origin = f3(p->cam_origin[0], p->cam_origin[1], p->cam_origin[2]);
    const float3 target = f3(p->cam_target[0], p->cam_target[1], p->cam_target[2]);
    const float3 up = f3(p->cam_up[0], p->cam_up[1], p->cam_up[2]);
    a->cam_o = origin;
    a->cam_fwd = normalize(target - origin);
    a->cam_right = normalize(cross(a->cam_fwd, up));
    a->cam_up = cross(a->cam_right, a->cam_fwd);
    a->tan_half_fov = tanf(0.5f * p->cam_fov);
    a->W = p->width; a->H = p->height; a->spp = p->spp; a->seed = p->seed;
    a->use_tent = p->use_tent_filter; a->max_depth = p->max_depth; a->rr_depth = p->rr_depth;
    a->use_mis = p->use_mis; a->prb_mode = p->prb_mode;
    a->sc.pdf_env_count = (p->flags & ZDR_FLAG_LIGHT_PDF_COUNTS_ENV) ? a->sc.env_count : 0;
    a->row0 = p->row0; a->row1 = p->row1; a->it0 = p->it0; a->it1 = p->it1;
    a->shard_count = p->shard_count > 1 ? p->shard_count : 1;
    a->shard_index = p->shard_count > 1 ? p->shard_index : 0;
    a->tex.rx = p->tex_res_x; a->tex.ry = p->tex_res_y; a->tex.bilinear = p->bilinear;
    a->tex.set_texels = p->tex_res_x * p->tex_res_y;
    {
        const int num_sets = p->num_material_sets > 1 ? p->num_material_sets : 1;
        for (size_t i = 0; i < s->meshes.size(); ++i) {
            const int ms = s->meshes[i].material_set;
            if (ms >= num_sets) return fail(ZDR_ERR_ARG, "instance %zu uses material set %d but the material buffer holds %d set(s)", i, ms, num_sets);
        }
        if ((long long)num_sets * a->tex.set_texels >= (1ll << 31)) return fail(ZDR_ERR_LIMIT, "material sets too large");
    }
    a->counters = s->d_counters;
    // gradient scatter strategy: warp aggregation pays when many lanes of a warp share a footprint,
    // i.e. when the textured object's uv range covers only a few texels (uniform / uv-less meshes)
    a->block_scatter = (p->flags & ZDR_FLAG_SCATTER_BLOCK) ? 1 : 0;
    if (p->flags & (ZDR_FLAG_SCATTER_AGGREGATE | ZDR_FLAG_SCATTER_BLOCK)) a->aggregate_scatter = 1;
    else if (p->flags & ZDR_FLAG_SCATTER_PLAIN) a->aggregate_scatter = 0;
    else {
        const float span_x = (s->uv_hi[0] - s->uv_lo[0]) * (float)(p->tex_res_x - 1);
        const float span_y = (s->uv_hi[1] - s->uv_lo[1]) * (float)(p->tex_res_y - 1);
        a->aggregate_scatter = !(span_x >= 64.0f && span_y >= 64.0f);
    }
    return ZDR_OK;
}

// Path integrator schedule: the wavefront kernels pay a fixed ~20 launches per batch, the one-kernel
// state machine does not; small jobs stay on the latter.  Returns 1 and fills a->wf when the
// wavefront path is to be used.
int setup_wavefront(zdr_scene* s, int integrator, const zdr_params* p, RenderArgs* a, bool* use) {
    *use = false;
    if (integrator != ZDR_PATH) return ZDR_OK;
    long long rows = 0;
    for (int t = a->shard_index, tiles = (p->row1 - p->row0 + 7) / 8; t < tiles; t += a->shard_count)
        rows += (p->row1 - (p->row0 + 8 * t)) < 8 ? (p->row1 - (p->row0 + 8 * t)) : 8;
    if (rows <= 0) return ZDR_OK;
    const long long pixels = (long long)p->width * rows;
    const long long samples = pixels * (p->it1 - p->it0);
    if ((p->flags & ZDR_FLAG_MEGAKERNEL) || p->max_depth > 255) return ZDR_OK;      // path depth is packed into 8 bits of the state
    if (!(p->flags & ZDR_FLAG_WAVEFRONT) && samples < (3ll << 20)) return ZDR_OK;      // (smaller jobs: persistent schedule, setup_persistent)
    // path-state budget: as many samples per batch as fit a quarter of the free HBM, at most 2^28 slots
    // (41 GB at 152 B per slot: all 256 samples of a 1024^2 image in one batch); ZDR_WF_SLOTS_LOG2 overrides.
    // Larger batches amortise the ~4*depth launches per batch and the thin tail rounds (measured: 2^24 -> 2^28
    // slots = 4.9 -> 5.5 Grays/s on C3).
    // Several state sets of a fraction of the samples each, worked on from as many streams (wavefront.cu
    // launch_path_wavefront), once the job is large enough for the parts to fill the GPU; ZDR_WF_LANES=1 keeps one.
    // ray binning (render.cuh WfState): which traversal stages read binned lists (ZDR_WF_BIN: bit 0 extension rays, bit 1 shadow
    // rays; read per call: tests switch it; default: on for large scenes).  Its arrays are part of the workspace only when it is on.
    const int bin_mask = (getenv("ZDR_WF_BIN") ? atoi(getenv("ZDR_WF_BIN")) : (s->bvh.num_nodes > kWfBigSceneNodes ? ZDR_WF_BIN_DEFAULT_BIG : ZDR_WF_BIN_DEFAULT)) & 3;
    const bool binned = bin_mask != 0;
    static const int lanes = getenv("ZDR_WF_LANES") ? atoi(getenv("ZDR_WF_LANES")) : 2;
    long long S = 0;
    int sets = 1;
    size_t slots = 0, set_bytes = 0, bytes = 0;
    auto plan = [&](long long target_slots) {
        S = target_slots / (pixels > 0 ? pixels : 1);
        if (S < 1) S = 1;
        if (S > p->it1 - p->it0) S = p->it1 - p->it0;
        sets = 1;                                   // every set keeps at least 2^22 slots
        while (sets < lanes && sets < kWfMaxSets && S >= 2 * (sets + 1) && pixels * S >= (long long)(sets + 1) << 22) ++sets;
        S = (S + sets - 1) / sets;
        slots = (size_t)(pixels * S);
        set_bytes = (slots * wf_bytes_per_slot(binned) + wf_fixed_bytes(binned) + 255) & ~(size_t)255;
        bytes = set_bytes * sets;
    };
    const size_t ws_cap = s->s_wf.external ? s->s_wf.cap : s->ws_limit;      // caller-provided buffer / caller-set bound
    if (ws_cap) {
        const size_t fixed = 2 * (wf_fixed_bytes(binned) + 256);      // up to two state sets
        long long fit = (long long)((ws_cap > fixed ? ws_cap - fixed : 0) / wf_bytes_per_slot(binned));
        if (fit < pixels) return fail(ZDR_ERR_LIMIT, "workspace of %zu MiB cannot hold one sample per pixel of this launch (%lld MiB needed)",
                                      ws_cap >> 20, (long long)(pixels * wf_bytes_per_slot(binned) + fixed) >> 20);
        plan(fit < (1ll << 28) ? fit : (1ll << 28));
        while (bytes > ws_cap && S > 1) { plan((long long)(slots - pixels)); }
    } else if (const char* slots_env = getenv("ZDR_WF_SLOTS_LOG2")) {
        plan(1ll << atoi(slots_env));
    } else {
        plan(1ll << 28);
        // (no query when the workspace already holds the plan: the call then makes no CUDA API call besides kernel
        // launches and event records, and can be captured into a CUDA graph)
        if (bytes > s->s_wf.cap) {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                const long long fit = (long long)((free_b + s->s_wf.cap) / 4 / wf_bytes_per_slot(binned));
                if (fit < (1ll << 28)) plan(fit);
            }
        }
    }
    cudaError_t e = s->s_wf.reserve(bytes);
    if (e != cudaSuccess) return fail(ZDR_ERR_CUDA, "wavefront state (%zu MiB): %s", bytes >> 20, cudaGetErrorString(e));
    char* base = s->s_wf.p;
    WfState& w = a->wf;
    w.AB = reinterpret_cast<float4*>(base); base += 2 * slots * sizeof(float4);
    w.CD = reinterpret_cast<float4*>(base); base += 2 * slots * sizeof(float4);
    w.H = reinterpret_cast<float4*>(base); base += slots * sizeof(float4);
    w.R = reinterpret_cast<float4*>(base); base += slots * sizeof(float4);
    w.SAB = reinterpret_cast<float4*>(base); base += 2 * slots * sizeof(float4);
    w.SC = reinterpret_cast<float4*>(base); base += slots * sizeof(float4);
    w.list[0] = reinterpret_cast<int*>(base); base += slots * sizeof(int);
    w.list[1] = reinterpret_cast<int*>(base); base += slots * sizeof(int);
    w.blist = nullptr; w.keys[0] = w.keys[1] = nullptr; w.hist[0] = w.hist[1] = w.cur = nullptr;
    if (binned) {
        w.blist = reinterpret_cast<int*>(base); base += slots * sizeof(int);
        w.keys[0] = reinterpret_cast<unsigned short*>(base); base += slots * sizeof(unsigned short);
        w.keys[1] = reinterpret_cast<unsigned short*>(base); base += slots * sizeof(unsigned short);
    }
    base = s->s_wf.p + (((base - s->s_wf.p) + 15) & ~(size_t)15);
    w.counts = reinterpret_cast<int*>(base); base += 64;
    if (binned) {
        w.hist[0] = reinterpret_cast<int*>(base); base += kWfBins * sizeof(int);      // hist[0], hist[1], cur: contiguous (k_wf_generate zeroes them)
        w.hist[1] = reinterpret_cast<int*>(base); base += kWfBins * sizeof(int);
        w.cur = reinterpret_cast<int*>(base);
    }
    w.bin_mask = bin_mask;
    w.bin_in = w.bin_out = 0;
    w.key_lo = s->bounds_lo;
    w.key_cells = s->bvh.num_nodes > kWfBigSceneNodes ? 16 : 8;
    w.key_scale = f3((float)w.key_cells / fmaxf(s->bounds_hi.x - s->bounds_lo.x, 1e-30f), (float)w.key_cells / fmaxf(s->bounds_hi.y - s->bounds_lo.y, 1e-30f),
                     (float)w.key_cells / fmaxf(s->bounds_hi.z - s->bounds_lo.z, 1e-30f));
    // shared-memory traversal (k_wf_trace_tiny): scenes of at most 8 KB whose links fit 16 bits and whose tree leaves at most 12
    // pending links (its stack: 16 local-memory entries, or 8 / 12 entries in registers); ZDR_WF_TINY_OFF=1: the general kernel
    const bool tiny_off = getenv("ZDR_WF_TINY_OFF") != nullptr;
    const int pending = s->bvh.max_depth - 1;
    w.tiny_regs = 0;
    if (!tiny_off && 64 * (size_t)s->bvh.num_nodes + 48 * (size_t)s->num_tris <= 8192 && pending <= 12) w.tiny_regs = pending <= 8 ? 4 : 6;
    w.S = (int)S;
    w.num_sms = s->num_sms;
    w.num_sets = sets;
    w.set_bytes = set_bytes;
    *use = true;
    return ZDR_OK;
}

// Persistent shared-memory schedule (persistent.cu).  Workspace: 64 batch counters + (forward only) one float4 per camera
// sample of a batch; the batch shrinks to fit `ZDR_PW_R_BYTES` (default 2 GiB) — the adjoint pass needs no buffer at all.
constexpr int kPwMaxBatches = 64;
constexpr size_t kPwHeaderBytes = 512;      // 64 batch counters + 8 profile accumulators, then the sample buffer
int setup_persistent(zdr_scene* s, int integrator, const zdr_params* p, RenderArgs* a, bool backward, bool* use, PwPlan* plan,
                     size_t* r_slots, unsigned** counters) {
    *use = false;
    if (integrator != ZDR_PATH || p->max_depth < 1 || p->max_depth > 255) return ZDR_OK;
    if (p->flags & (ZDR_FLAG_WAVEFRONT | ZDR_FLAG_MEGAKERNEL)) return ZDR_OK;
    long long rows = 0;
    for (int t = a->shard_index, tiles = (p->row1 - p->row0 + 7) / 8; t < tiles; t += a->shard_count)
        rows += (p->row1 - (p->row0 + 8 * t)) < 8 ? (p->row1 - (p->row0 + 8 * t)) : 8;
    if (rows <= 0 || p->it1 <= p->it0) return ZDR_OK;
    const long long pixels = (long long)p->width * rows, its = p->it1 - p->it0;
    // Default schedule by job size (tools/schedule_sweep.py, profiles/r02g_schedule_sweep.json: C3 scene, forward + adjoint):
    //   < 2^14 camera samples: one-kernel state machine (everything is launch-bound there);
    //   2^14 .. 3 * 2^20:      this persistent shared-memory schedule (2 launches per pass: 1.3-3x the others);
    //   above:                 HBM wavefront (its ~100 launches per pass amortise; 7.0 vs 5.1 Grays/s at 2^26+).
    static const long long auto_lo = 1ll << (getenv("ZDR_PW_AUTO_LOG2") ? atoi(getenv("ZDR_PW_AUTO_LOG2")) : 14);
    static const long long auto_hi = getenv("ZDR_PW_AUTO_MAX") ? atoll(getenv("ZDR_PW_AUTO_MAX")) : 3ll << 20;
    if (!(p->flags & ZDR_FLAG_PERSISTENT) && (pixels * its < auto_lo || pixels * its >= auto_hi)) return ZDR_OK;
    cudaError_t e = pw_plan(a->sc, s->bvh.max_depth, s->num_sms, plan);
    if (e != cudaSuccess) return fail(ZDR_ERR_CUDA, "persistent schedule: %s", cudaGetErrorString(e));
    static const long long r_cap_default = getenv("ZDR_PW_R_BYTES") ? atoll(getenv("ZDR_PW_R_BYTES")) : (2ll << 30);
    long long r_cap = r_cap_default;
    const size_t ws_cap = s->s_wf.external ? s->s_wf.cap : s->ws_limit;
    if (ws_cap) {
        const long long room = (long long)ws_cap - (long long)kPwHeaderBytes - (long long)plan->gstate_bytes;
        if (room < (backward ? 0 : pixels * 16)) return fail(ZDR_ERR_LIMIT, "workspace of %zu MiB is too small for this launch", ws_cap >> 20);
        if (room < r_cap) r_cap = room;
    }
    size_t slots = 0;
    if (!backward) {
        long long smax = r_cap / 16 / pixels;
        if (smax < 1) smax = 1;
        if (smax > its) smax = its;
        while ((its + smax - 1) / smax > kPwMaxBatches) ++smax;
        slots = (size_t)(pixels * smax);
    }
    const size_t bytes = kPwHeaderBytes + plan->gstate_bytes + slots * sizeof(float4);
    if ((e = s->s_wf.reserve(bytes)) != cudaSuccess) return fail(ZDR_ERR_CUDA, "persistent workspace (%zu MiB): %s", bytes >> 20, cudaGetErrorString(e));
    *counters = reinterpret_cast<unsigned*>(s->s_wf.p);
    a->pw.gstate = reinterpret_cast<float4*>(s->s_wf.p + kPwHeaderBytes);
    a->pw.R = reinterpret_cast<float4*>(s->s_wf.p + kPwHeaderBytes + plan->gstate_bytes);
    *r_slots = slots;
    *use = true;
    return ZDR_OK;
}

inline bool want_stats(const zdr_params* p) { return (p->flags & ZDR_FLAG_STATS) != 0; }
inline bool fast_math(const zdr_params* p) { return (p->flags & ZDR_FLAG_FAST_MATH) != 0; }      // the render_fast.cu / persistent_fast.cu twins

}  // namespace

// ---- library-wide ---------------------------------------------------------------------------
ZDR_EXPORT int zdr_abi_version(void) { return ZDR_ABI_VERSION; }
ZDR_EXPORT const char* zdr_last_error(void) { return g_err; }

ZDR_EXPORT int zdr_set_sampler_tables(int device, const uint32_t* pmj_host, int sets, int n, const uint16_t* bn_host,
                                      int ntex, int res) {
    if (!pmj_host || !bn_host || sets <= 0 || n <= 0 || ntex <= 0 || res <= 0) return fail(ZDR_ERR_ARG, "bad sampler tables");
    if ((long long)ntex * res * res >= (1ll << 29) || (long long)sets * n >= (1ll << 29))
        return fail(ZDR_ERR_LIMIT, "sampler tables too large (the kernels index them with 32-bit arithmetic)");
    DevGuard dev_guard;
    int rc = use_device(device, dev_guard);
    if (rc) return rc;
    const size_t np = (size_t)sets * n * 2, nb = (size_t)ntex * res * res;
    std::vector<float> pf(np), bf(nb);
    for (size_t i = 0; i < np; ++i) pf[i] = (float)pmj_host[i] * 0x1p-32f;   // pmj02bn.py:9: RN(u32->f32) * 2^-32
    for (size_t i = 0; i < nb; ++i) bf[i] = (float)bn_host[i] * 0x1p-16f;    // pmj02bn.py:14
    std::lock_guard<std::mutex> lock(g_tables_mu);
    Tables& t = g_tables[device];
    CK(cudaDeviceSynchronize());
    if (t.pmj) cudaFree(t.pmj);
    if (t.bn) cudaFree(t.bn);
    t = Tables();
    CK(cudaMalloc(&t.pmj, sizeof(float) * np));
    CK(cudaMalloc(&t.bn, sizeof(float) * nb));
    CK(cudaMemcpy(t.pmj, pf.data(), sizeof(float) * np, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t.bn, bf.data(), sizeof(float) * nb, cudaMemcpyHostToDevice));
    t.sets = sets; t.n = n; t.ntex = ntex; t.res = res;
    return ZDR_OK;
}

// ---- scene assembly -------------------------------------------------------------------------
ZDR_EXPORT int zdr_scene_create(zdr_scene** out, int device) {
    if (!out) return fail(ZDR_ERR_ARG, "null out");
    *out = nullptr;
    DevGuard dev_guard;
    int rc = use_device(device, dev_guard);
    if (rc) return rc;
    zdr_scene* s = new (std::nothrow) zdr_scene();
    if (!s) return fail(ZDR_ERR_ARG, "out of host memory");
    s->device = device;
    cudaError_t e = cudaMalloc(&s->d_counters, sizeof(unsigned long long) * C_NUM);
    if (e == cudaSuccess) e = cudaMemset(s->d_counters, 0, sizeof(unsigned long long) * C_NUM);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    for (int k = 0; k < kWfMaxSets - 1 && e == cudaSuccess; ++k) e = cudaStreamCreateWithFlags(&s->side_streams[k], cudaStreamNonBlocking);
    for (int k = 0; k < kWfMaxSets + 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&s->wf_events[k], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete s;
        return fail(ZDR_ERR_CUDA, "scene_create: %s", cudaGetErrorString(e));
    }
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, device);
    *out = s;
    return ZDR_OK;
}

ZDR_EXPORT void zdr_scene_destroy(zdr_scene* s) {
    if (!s) return;
    DevGuard dev_guard;
    if (use_device(s->device, dev_guard) != ZDR_OK) return;
    cudaDeviceSynchronize();
    free_bvh(&s->bvh);
    free_envmap(&s->env);
    cudaFree(s->d_inst); cudaFree(s->d_lights); cudaFree(s->d_counters);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    for (cudaStream_t st : s->side_streams) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t ev : s->wf_events) if (ev) cudaEventDestroy(ev);
    s->s_mat.release(); s->s_img.release(); s->s_dimg.release(); s->s_dmat.release(); s->s_misc.release();
    s->s_misc2.release(); s->s_int.release(); s->s_int2.release(); s->s_wf.release();
    delete s;
}

ZDR_EXPORT int zdr_scene_add_mesh(zdr_scene* s, const float* verts, int nv, const int32_t* tris, int nt,
                                  const float xform[16], const float emission[3]) {
    if (!s || !verts || !tris || nv <= 0 || nt <= 0) return fail(ZDR_ERR_ARG, "add_mesh: bad arguments");
    if (s->meshes.size() >= 10000) return fail(ZDR_ERR_LIMIT, "too many instances (render.py:114: 10000)");
    for (int i = 0; i < 3 * nt; ++i)
        if (tris[i] < 0 || tris[i] >= nv) return fail(ZDR_ERR_ARG, "add_mesh: index %d out of range", tris[i]);
    static const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    s->meshes.emplace_back();
    Mesh& m = s->meshes.back();
    m.verts.assign(verts, verts + 8 * (size_t)nv);
    m.tris.assign(tris, tris + 3 * (size_t)nt);
    memcpy(m.M, xform ? xform : I16, sizeof(m.M));
    normal_matrix(m.M, m.N);
    for (int k = 0; k < 3; ++k) m.emission[k] = emission ? emission[k] : 0.0f;
    s->committed = false;
    return ZDR_OK;
}

// Scene.load_geometry starts from a fresh luisa.Accel() / BindlessArray() on every call (render.py:74-76): the
// instances, the tree and the envmap of an earlier call are dropped.
ZDR_EXPORT int zdr_scene_clear(zdr_scene* s) {
    if (!s) return fail(ZDR_ERR_ARG, "scene_clear: null");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    s->meshes.clear();
    free_bvh(&s->bvh);
    free_envmap(&s->env);
    cudaFree(s->d_inst); cudaFree(s->d_lights);
    s->d_inst = nullptr; s->d_lights = nullptr;
    s->light_count = 0; s->num_tris = 0;
    s->committed = false;
    return ZDR_OK;
}

ZDR_EXPORT int zdr_scene_commit(zdr_scene* s) {
    if (!s || s->meshes.empty()) return fail(ZDR_ERR_STATE, "commit: empty scene");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    size_t nt = 0;
    for (Mesh& m : s->meshes) { m.tri_base = (int)nt; nt += m.tris.size() / 3; }
    if (nt >= (1u << 29)) return fail(ZDR_ERR_LIMIT, "too many triangles");
    const char* force = getenv("ZDR_BVH");     // "lbvh" | "csah" | "sah": builder override for experiments and tests
    const bool sah = force ? strcmp(force, "sah") == 0 : nt <= (size_t)kSahMaxTris;
    const bool lbvh = force && strcmp(force, "lbvh") == 0;
    update_uv_bounds(s);
    CK(cudaDeviceSynchronize());
    free_bvh(&s->bvh);
    cudaFree(s->d_inst); cudaFree(s->d_lights);
    s->d_inst = nullptr; s->d_lights = nullptr;
    s->committed = false;
    cudaError_t e = cudaSuccess;
    const bool verbose = getenv("ZDR_VERBOSE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto seconds_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    if (sah) {
        // small scene: flatten the instances to world space on the host, p = (M * float4(v,1)).xyz as
        // interaction.py:19-21, and build the full-sweep SAH tree there
        std::vector<float4> tri(3 * nt), shade(4 * nt);
        float3 lo = f3(INFINITY, INFINITY, INFINITY), hi = f3(-INFINITY, -INFINITY, -INFINITY);
        bool finite = true;
        for (size_t i = 0; i < s->meshes.size(); ++i) {
            const Mesh& m = s->meshes[i];
            const size_t mt = m.tris.size() / 3;
            for (size_t t = 0; t < mt; ++t) {
                const size_t g = m.tri_base + t;
                const float* v[3];
                float3 p[3];
                for (int k = 0; k < 3; ++k) {
                    v[k] = m.verts.data() + 8 * (size_t)m.tris[3 * t + k];
                    p[k] = xform_point(m.M, f3(v[k][0], v[k][1], v[k][2]));
                    lo = f3(fminf(lo.x, p[k].x), fminf(lo.y, p[k].y), fminf(lo.z, p[k].z));
                    hi = f3(fmaxf(hi.x, p[k].x), fmaxf(hi.y, p[k].y), fmaxf(hi.z, p[k].z));
                    finite = finite && std::isfinite(p[k].x) && std::isfinite(p[k].y) && std::isfinite(p[k].z);
                }
                tri[3 * g + 0] = make_float4(p[0].x, p[0].y, p[0].z, __builtin_bit_cast(float, (int)g));
                tri[3 * g + 1] = make_float4(p[1].x, p[1].y, p[1].z, __builtin_bit_cast(float, (int)i));
                tri[3 * g + 2] = make_float4(p[2].x, p[2].y, p[2].z, __builtin_bit_cast(float, (int)t));
                for (int k = 0; k < 3; ++k) shade[4 * g + k] = make_float4(v[k][5], v[k][6], v[k][7], v[k][3]);
                shade[4 * g + 3] = make_float4(v[0][4], v[1][4], v[2][4], 0.0f);
            }
        }
        if (!(finite && lo.x <= hi.x)) return fail(ZDR_ERR_ARG, "commit: non-finite geometry");
        s->bounds_lo = lo; s->bounds_hi = hi;
        e = build_bvh_sah_host(tri.data(), shade.data(), (int)nt, lo, hi, s->own_stream, &s->bvh);
    } else {
        // large scene: the same records are produced on the GPU (k_flatten), then the GPU build
        std::vector<FlattenInput> in(s->meshes.size());
        for (size_t i = 0; i < s->meshes.size(); ++i) {
            const Mesh& m = s->meshes[i];
            in[i] = FlattenInput{m.verts.data(), m.tris.data(), (int)(m.verts.size() / 8), (int)(m.tris.size() / 3), m.M};
        }
        float4 *d_tri = nullptr, *d_shade = nullptr;
        float3 lo, hi;
        bool finite = true;
        e = flatten_scene_gpu(in.data(), (int)in.size(), s->own_stream, &d_tri, &d_shade, &lo, &hi, &finite);
        if (e == cudaSuccess && !(finite && lo.x <= hi.x)) {
            cudaFree(d_tri); cudaFree(d_shade);
            return fail(ZDR_ERR_ARG, "commit: non-finite geometry");
        }
        const double t_flat = seconds_since(t_start);
        if (e == cudaSuccess) { s->bounds_lo = lo; s->bounds_hi = hi; }
        if (e == cudaSuccess) e = build_bvh(d_tri, d_shade, (int)nt, lo, hi, s->own_stream, &s->bvh, !lbvh);
        cudaFree(d_tri); cudaFree(d_shade);
        if (verbose)
            fprintf(stderr, "[zdr] commit: %zu triangles, upload + GPU flatten %.3f s, %s build %.3f s (%d nodes, depth <= %d)\n", nt, t_flat,
                    lbvh ? "LBVH" : "clustered-SAH", seconds_since(t_start) - t_flat, s->bvh.num_nodes, s->bvh.max_depth);
    }
    if (e != cudaSuccess) return fail(ZDR_ERR_CUDA, "commit/build_bvh: %s", cudaGetErrorString(e));
    if (s->bvh.max_depth + 2 > kStackSize) return fail(ZDR_ERR_LIMIT, "BVH depth %d exceeds the traversal stack", s->bvh.max_depth);
    s->num_tris = (int)nt;
    CK(cudaMalloc(&s->d_inst, sizeof(InstData) * s->meshes.size()));
    CK(cudaMalloc(&s->d_lights, sizeof(int) * s->meshes.size()));
    rc = upload_lights(s);
    if (rc) return rc;
    s->committed = true;
    return ZDR_OK;
}

ZDR_EXPORT int zdr_scene_set_emissions(zdr_scene* s, const float* e3, int n) {   // render.py:130-148
    if (!s || !e3) return fail(ZDR_ERR_ARG, "set_emissions: null");
    if (n != (int)s->meshes.size()) return fail(ZDR_ERR_ARG, "set_emissions: %d emissions for %d instances (render.py:141)", n, (int)s->meshes.size());
    bool changed = false;
    for (int i = 0; i < n; ++i) {
        changed |= memcmp(s->meshes[i].emission, e3 + 3 * i, sizeof(float) * 3) != 0;
        memcpy(s->meshes[i].emission, e3 + 3 * i, sizeof(float) * 3);
    }
    // RenderOperator.backward re-applies the forward call's emissions on every step (render.py:218): nothing to
    // upload, and no reason to synchronise the device, when they are what the device already holds
    if (!s->committed || !changed) return ZDR_OK;
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    return upload_lights(s);
}

// Material sets ("polymorphic materials", README.md:5): set_of_instance[i] = the texture set that shades instance i, or -1 for
// a light / terminal surface.  The default — instance 0 -> set 0, every other instance -1 — is the reference's rule "treat
// everything as light other than the first object" (direct.py:30, prb.py:45).
ZDR_EXPORT int zdr_scene_set_material_sets(zdr_scene* s, const int32_t* set_of_instance, int n) {
    if (!s || !set_of_instance) return fail(ZDR_ERR_ARG, "set_material_sets: null");
    if (n != (int)s->meshes.size()) return fail(ZDR_ERR_ARG, "set_material_sets: %d entries for %d instances", n, (int)s->meshes.size());
    for (int i = 0; i < n; ++i)
        if (set_of_instance[i] < -1) return fail(ZDR_ERR_ARG, "set_material_sets: entry %d is %d (>= -1 expected)", i, set_of_instance[i]);
    for (int i = 0; i < n; ++i) s->meshes[i].material_set = set_of_instance[i];
    update_uv_bounds(s);
    if (!s->committed) return ZDR_OK;
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    return upload_lights(s);
}

ZDR_EXPORT int zdr_scene_set_envmap(zdr_scene* s, const float* rgba, int w, int h, int compensate_mis) {
    if (!s || !rgba || w <= 0 || h <= 0) return fail(ZDR_ERR_ARG, "set_envmap: bad arguments");
    if (h != w && w != 2 * h) return fail(ZDR_ERR_ARG, "set_envmap: image must be 1:1 or 2:1 (envmap.py:122-128)");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    cudaError_t e = build_envmap(rgba, w, h, compensate_mis != 0, s->own_stream, &s->env);
    if (e != cudaSuccess) return fail(ZDR_ERR_CUDA, "set_envmap: %s", cudaGetErrorString(e));
    return ZDR_OK;
}

// Workspace of the path-integrator schedules (wavefront path state: up to 152 B per resident path; persistent: 16 B per
// camera sample of a forward batch).  By default the library allocates it with cudaMalloc, outside the caller's allocator:
//   zdr_scene_set_workspace_limit  bounds it (batches shrink to fit);
//   zdr_scene_set_workspace        makes the scene use a buffer the CALLER owns (a torch tensor: the caching allocator then
//                                  accounts for it); NULL / 0 goes back to the library's own allocation;
//   zdr_scene_trim                 frees what the library allocated (it is re-allocated on demand).
ZDR_EXPORT int zdr_scene_set_workspace_limit(zdr_scene* s, uint64_t bytes) {
    if (!s) return fail(ZDR_ERR_ARG, "set_workspace_limit: null");
    s->ws_limit = (size_t)bytes;
    return ZDR_OK;
}
ZDR_EXPORT int zdr_scene_set_workspace(zdr_scene* s, void* workspace_dev, uint64_t bytes) {
    if (!s) return fail(ZDR_ERR_ARG, "set_workspace: null");
    if ((workspace_dev == nullptr) != (bytes == 0)) return fail(ZDR_ERR_ARG, "set_workspace: pointer and size must both be given or both be zero");
    if (((uintptr_t)workspace_dev & 255) != 0) return fail(ZDR_ERR_ARG, "set_workspace: the buffer must be 256-byte aligned");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    s->s_wf.release();
    if (workspace_dev) { s->s_wf.p = static_cast<char*>(workspace_dev); s->s_wf.cap = (size_t)bytes; s->s_wf.external = true; }
    return ZDR_OK;
}
ZDR_EXPORT int zdr_scene_trim(zdr_scene* s) {
    if (!s) return fail(ZDR_ERR_ARG, "scene_trim: null");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    if (!s->s_wf.external) s->s_wf.release();
    s->s_mat.release(); s->s_img.release(); s->s_dimg.release(); s->s_dmat.release(); s->s_misc.release(); s->s_misc2.release();
    s->s_int.release(); s->s_int2.release();
    return ZDR_OK;
}
ZDR_EXPORT uint64_t zdr_scene_workspace_bytes(const zdr_scene* s) { return s ? (uint64_t)s->s_wf.cap : 0; }

ZDR_EXPORT int zdr_scene_instance_count(const zdr_scene* s) { return s ? (int)s->meshes.size() : 0; }
ZDR_EXPORT int zdr_scene_light_count(const zdr_scene* s) {
    if (!s) return 0;
    int n = 0;
    for (const Mesh& m : s->meshes) n += (m.emission[0] > 0.f || m.emission[1] > 0.f || m.emission[2] > 0.f) ? 1 : 0;
    return n;
}
ZDR_EXPORT int zdr_scene_triangle_count(const zdr_scene* s) {
    if (!s) return 0;
    size_t n = 0;
    for (const Mesh& m : s->meshes) n += m.tris.size() / 3;
    return (int)n;
}

// ---- the hot path ---------------------------------------------------------------------------
ZDR_EXPORT int zdr_render_forward(zdr_scene* s, int integrator, const zdr_params* p, const float* material_dev,
                                  float* image_dev, void* stream) {
    int rc = check_params(s, integrator, p);
    if (rc) return rc;
    if (!material_dev || !image_dev) return fail(ZDR_ERR_ARG, "render_forward: null buffer");
    DevGuard dev_guard;
    if ((rc = use_device(s->device, dev_guard))) return rc;
    RenderArgs a;
    if ((rc = make_args(s, p, &a))) return rc;
    a.tex.mat = reinterpret_cast<const float4*>(material_dev);
    a.image = reinterpret_cast<float4*>(image_dev);
    bool persistent, wavefront;
    PwPlan plan;
    size_t r_slots = 0;
    unsigned* pw_counters = nullptr;
    if ((rc = setup_persistent(s, integrator, p, &a, false, &persistent, &plan, &r_slots, &pw_counters))) return rc;
    if (persistent) {
        CK((fast_math(p) ? launch_path_persistent_fast : launch_path_persistent)(false, a, want_stats(p), plan, r_slots, pw_counters, kPwMaxBatches, (cudaStream_t)stream));
        return ZDR_OK;
    }
    if ((rc = setup_wavefront(s, integrator, p, &a, &wavefront))) return rc;
    if (wavefront) CK((fast_math(p) ? launch_path_wavefront_fast : launch_path_wavefront)(false, a, want_stats(p), (cudaStream_t)stream, s->side_streams, s->wf_events));
    else CK((fast_math(p) ? launch_forward_fast : launch_forward)(integrator, a, want_stats(p), (cudaStream_t)stream));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_render_backward(zdr_scene* s, int integrator, const zdr_params* p, const float* material_dev,
                                   const float* d_image_dev, float* d_material_dev, void* stream) {
    int rc = check_params(s, integrator, p);
    if (rc) return rc;
    if (!material_dev || !d_image_dev || !d_material_dev) return fail(ZDR_ERR_ARG, "render_backward: null buffer");
    DevGuard dev_guard;
    if ((rc = use_device(s->device, dev_guard))) return rc;
    RenderArgs a;
    if ((rc = make_args(s, p, &a))) return rc;
    a.tex.mat = reinterpret_cast<const float4*>(material_dev);
    a.tex.dmat = reinterpret_cast<float4*>(d_material_dev);
    a.d_image = reinterpret_cast<const float4*>(d_image_dev);
    bool persistent, wavefront;
    PwPlan plan;
    size_t r_slots = 0;
    unsigned* pw_counters = nullptr;
    if ((rc = setup_persistent(s, integrator, p, &a, true, &persistent, &plan, &r_slots, &pw_counters))) return rc;
    if (persistent) {
        CK((fast_math(p) ? launch_path_persistent_fast : launch_path_persistent)(true, a, want_stats(p), plan, 0, pw_counters, kPwMaxBatches, (cudaStream_t)stream));
        return ZDR_OK;
    }
    if ((rc = setup_wavefront(s, integrator, p, &a, &wavefront))) return rc;
    if (wavefront) CK((fast_math(p) ? launch_path_wavefront_fast : launch_path_wavefront)(true, a, want_stats(p), (cudaStream_t)stream, s->side_streams, s->wf_events));
    else CK((fast_math(p) ? launch_backward_fast : launch_backward)(integrator, a, want_stats(p), (cudaStream_t)stream));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_render_duvdxy(zdr_scene* s, const zdr_params* p, float* image_dev, void* stream) {
    int rc = check_params(s, ZDR_COLLOCATED, p);
    if (rc) return rc;
    if (!image_dev) return fail(ZDR_ERR_ARG, "render_duvdxy: null buffer");
    DevGuard dev_guard;
    if ((rc = use_device(s->device, dev_guard))) return rc;
    RenderArgs a;
    if ((rc = make_args(s, p, &a))) return rc;
    a.image = reinterpret_cast<float4*>(image_dev);
    CK(launch_duvdxy(a, (cudaStream_t)stream));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_render_forward_host(zdr_scene* s, int integrator, const zdr_params* p, const float* material_host,
                                       float* image_host) {
    int rc = check_params(s, integrator, p);
    if (rc) return rc;
    if (!material_host || !image_host) return fail(ZDR_ERR_ARG, "render_forward_host: null buffer");
    DevGuard dev_guard;
    if ((rc = use_device(s->device, dev_guard))) return rc;
    const size_t nmat = (size_t)p->tex_res_x * p->tex_res_y * 4 * (p->num_material_sets > 1 ? p->num_material_sets : 1), nimg = (size_t)p->width * p->height * 4;
    CK(s->s_mat.reserve(nmat));
    CK(s->s_img.reserve(nimg));
    cudaStream_t st = s->own_stream;
    CK(cudaMemcpyAsync(s->s_mat.p, material_host, sizeof(float) * nmat, cudaMemcpyHostToDevice, st));
    // rows of the other shards, and every row when the launch has no samples to render (it0 == it1)
    if (p->shard_count > 1 || p->it1 <= p->it0) CK(cudaMemsetAsync(s->s_img.p, 0, sizeof(float) * nimg, st));
    if ((rc = zdr_render_forward(s, integrator, p, s->s_mat.p, s->s_img.p, st))) return rc;
    const size_t off = (size_t)p->row0 * p->width * 4, cnt = (size_t)(p->row1 - p->row0) * p->width * 4;
    CK(cudaMemcpyAsync(image_host + off, s->s_img.p + off, sizeof(float) * cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_render_backward_host(zdr_scene* s, int integrator, const zdr_params* p, const float* material_host,
                                        const float* d_image_host, float* d_material_host) {
    int rc = check_params(s, integrator, p);
    if (rc) return rc;
    if (!material_host || !d_image_host || !d_material_host) return fail(ZDR_ERR_ARG, "render_backward_host: null buffer");
    DevGuard dev_guard;
    if ((rc = use_device(s->device, dev_guard))) return rc;
    const size_t nmat = (size_t)p->tex_res_x * p->tex_res_y * 4 * (p->num_material_sets > 1 ? p->num_material_sets : 1), nimg = (size_t)p->width * p->height * 4;
    CK(s->s_mat.reserve(nmat));
    CK(s->s_dmat.reserve(nmat));
    CK(s->s_dimg.reserve(nimg));
    cudaStream_t st = s->own_stream;
    CK(cudaMemcpyAsync(s->s_mat.p, material_host, sizeof(float) * nmat, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->s_dimg.p, d_image_host, sizeof(float) * nimg, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(s->s_dmat.p, 0, sizeof(float) * nmat, st));                                  // render.py:220
    if ((rc = zdr_render_backward(s, integrator, p, s->s_mat.p, s->s_dimg.p, s->s_dmat.p, st))) return rc;
    CK(cudaMemcpyAsync(d_material_host, s->s_dmat.p, sizeof(float) * nmat, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ZDR_OK;
}

// ---- instrumentation ------------------------------------------------------------------------
ZDR_EXPORT int zdr_get_counters(zdr_scene* s, uint64_t out[ZDR_NUM_COUNTERS]) {
    if (!s || !out) return fail(ZDR_ERR_ARG, "get_counters: null");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    static_assert(C_NUM == ZDR_NUM_COUNTERS, "counter count");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, s->d_counters, sizeof(uint64_t) * C_NUM, cudaMemcpyDeviceToHost));
    return ZDR_OK;
}
ZDR_EXPORT uint64_t zdr_kernel_launches(void) { return kernel_launch_count(); }

// Phase profile of the scene's last persistent launch (environment ZDR_PW_PROFILE=1): SM cycles summed over the CTAs for
// claim + regenerate, extend, shade A, shadow, shade B; rounds; live paths summed over rounds; total cycles.
ZDR_EXPORT int zdr_debug_phase_cycles(zdr_scene* s, uint64_t out[8]) {
    if (!s || !out) return fail(ZDR_ERR_ARG, "debug_phase_cycles: null");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    if (s->s_wf.cap < kPwHeaderBytes) return fail(ZDR_ERR_STATE, "debug_phase_cycles: no persistent launch yet");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, s->s_wf.p + 256, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_reset_counters(zdr_scene* s) {
    if (!s) return fail(ZDR_ERR_ARG, "reset_counters: null");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(s->d_counters, 0, sizeof(uint64_t) * C_NUM));
    return ZDR_OK;
}

// ---- probes ---------------------------------------------------------------------------------
ZDR_EXPORT int zdr_probe_trace_closest(zdr_scene* s, const float* rays, int n, int32_t* inst_prim, float* uvt) {
    if (!s || !rays || !inst_prim || !uvt || n < 0) return fail(ZDR_ERR_ARG, "probe_trace_closest: bad arguments");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    SceneView sc;
    if ((rc = scene_view(s, &sc))) return rc;
    if (n == 0) return ZDR_OK;
    CK(s->s_misc.reserve(8 * (size_t)n));
    CK(s->s_misc2.reserve(3 * (size_t)n));
    CK(s->s_int.reserve(2 * (size_t)n));
    cudaStream_t st = s->own_stream;
    CK(cudaMemcpyAsync(s->s_misc.p, rays, sizeof(float) * 8 * n, cudaMemcpyHostToDevice, st));
    CK(launch_probe_trace(sc, s->s_misc.p, n, false, s->s_int.p, s->s_misc2.p, nullptr, st));
    CK(cudaMemcpyAsync(inst_prim, s->s_int.p, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(uvt, s->s_misc2.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_probe_trace_any(zdr_scene* s, const float* rays, int n, int32_t* occluded) {
    if (!s || !rays || !occluded || n < 0) return fail(ZDR_ERR_ARG, "probe_trace_any: bad arguments");
    DevGuard dev_guard;
    int rc = use_device(s->device, dev_guard);
    if (rc) return rc;
    SceneView sc;
    if ((rc = scene_view(s, &sc))) return rc;
    if (n == 0) return ZDR_OK;
    CK(s->s_misc.reserve(8 * (size_t)n));
    CK(s->s_int.reserve((size_t)n));
    cudaStream_t st = s->own_stream;
    CK(cudaMemcpyAsync(s->s_misc.p, rays, sizeof(float) * 8 * n, cudaMemcpyHostToDevice, st));
    CK(launch_probe_trace(sc, s->s_misc.p, n, true, nullptr, nullptr, s->s_int.p, st));
    CK(cudaMemcpyAsync(occluded, s->s_int.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_probe_sampler(int device, const int32_t* state, int n, const char* prog, float* out) {
    if (!state || !prog || !out || n < 0) return fail(ZDR_ERR_ARG, "probe_sampler: bad arguments");
    DevGuard dev_guard;
    int rc = use_device(device, dev_guard);
    if (rc) return rc;
    const Tables& t = g_tables[device];
    if (!t.pmj || !t.bn) return fail(ZDR_ERR_STATE, "sampler tables not set for device %d", device);
    if (n == 0) return ZDR_OK;
    SceneView sc;
    memset(&sc, 0, sizeof(sc));
    sc.pmj = t.pmj; sc.bn = t.bn; sc.pmj_sets = t.sets; sc.pmj_n = t.n; sc.bn_ntex = t.ntex; sc.bn_res = t.res;
    const int len = (int)strlen(prog);
    int per = 0;
    for (int k = 0; k < len; ++k) per += prog[k] == '2' ? 2 : 1;
    int* d_state = nullptr; char* d_prog = nullptr; float* d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_state, sizeof(int) * 5 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_prog, (size_t)len + 1);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(float) * (size_t)per * n + 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_state, state, sizeof(int) * 5 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_prog, prog, (size_t)len + 1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_probe_sampler(sc, d_state, n, d_prog, len, per, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(float) * (size_t)per * n, cudaMemcpyDeviceToHost);
    cudaFree(d_state); cudaFree(d_prog); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ZDR_ERR_CUDA, "probe_sampler: %s", cudaGetErrorString(e));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_probe_render_samples(zdr_scene* s, int integrator, const zdr_params* p, const float* material_host,
                                        float* out_host) {
    int rc = check_params(s, integrator, p);
    if (rc) return rc;
    if (!material_host || !out_host) return fail(ZDR_ERR_ARG, "probe_render_samples: null buffer");
    DevGuard dev_guard;
    if ((rc = use_device(s->device, dev_guard))) return rc;
    const size_t nmat = (size_t)p->tex_res_x * p->tex_res_y * 4 * (p->num_material_sets > 1 ? p->num_material_sets : 1), nout = (size_t)p->width * p->height * p->spp * 3;
    CK(s->s_mat.reserve(nmat));
    CK(s->s_misc.reserve(nout));
    RenderArgs a;
    if ((rc = make_args(s, p, &a))) return rc;
    cudaStream_t st = s->own_stream;
    CK(cudaMemcpyAsync(s->s_mat.p, material_host, sizeof(float) * nmat, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(s->s_misc.p, 0, sizeof(float) * nout, st));
    a.tex.mat = reinterpret_cast<const float4*>(s->s_mat.p);
    a.samples_out = s->s_misc.p;
    bool persistent, wavefront;
    PwPlan plan;
    size_t r_slots = 0;
    unsigned* pw_counters = nullptr;
    if ((rc = setup_persistent(s, integrator, p, &a, false, &persistent, &plan, &r_slots, &pw_counters))) return rc;
    if (persistent) {
        CK((fast_math(p) ? launch_path_persistent_fast : launch_path_persistent)(false, a, want_stats(p), plan, r_slots, pw_counters, kPwMaxBatches, st));
    } else {
        if ((rc = setup_wavefront(s, integrator, p, &a, &wavefront))) return rc;
        if (wavefront) CK((fast_math(p) ? launch_path_wavefront_fast : launch_path_wavefront)(false, a, want_stats(p), st, s->side_streams, s->wf_events));
        else CK((fast_math(p) ? launch_forward_fast : launch_forward)(integrator, a, want_stats(p), st));
    }
    CK(cudaMemcpyAsync(out_host, s->s_misc.p, sizeof(float) * nout, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ZDR_OK;
}

ZDR_EXPORT int zdr_probe_ggx(int device, const float* in, int n, float* out) {
    if (!in || !out || n < 0) return fail(ZDR_ERR_ARG, "probe_ggx: bad arguments");
    DevGuard dev_guard;
    int rc = use_device(device, dev_guard);
    if (rc) return rc;
    if (n == 0) return ZDR_OK;
    float *d_in = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_in, sizeof(float) * 13 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(float) * 7 * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, in, sizeof(float) * 13 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_probe_ggx(d_in, n, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(float) * 7 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ZDR_ERR_CUDA, "probe_ggx: %s", cudaGetErrorString(e));
    return ZDR_OK;
}
