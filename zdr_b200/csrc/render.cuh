// render.cuh — kernel argument block shared by the render kernels (render.cu, wavefront.cu, persistent.cu) and api.cu.
#pragma once
#include "scene.cuh"
#include "shading.cuh"
#include "../../include/zdr_b200.h"

namespace zdr {

// Global-memory path state of the wavefront path tracer (wavefront.cu, k_wf_*): one slot per
// (sample-in-batch j, pixel pix), slot = j * P + pix, so a warp works on 32 neighbouring pixels.
struct WfState {
    // pairs of records that are always read together share one 32-byte sector (slots of a compacted list are
    // not contiguous, so a lone 16-byte record would waste half of every sector fetched):
    float4* AB;       // [2*slot] ray origin.xyz | pdf_bsdf      [2*slot+1] ray direction.xyz | packed (depth, phase, sampler dimension)
    float4* CD;       // [2*slot] beta.xyz | -                   [2*slot+1] L.xyz (radiance accumulated / Le remaining) | -
    float4* H;        // closest hit: slot | u | v | t
    float4* R;        // finished sample: clamped radiance.xyz | -   (forward only)
    float4* SAB;      // [2*slot] shadow ray origin.xyz (= hit point) | light.dist   [2*slot+1] light.wi.xyz | light.pdf
    float4* SC;       // light.eval.xyz | occluded (written by the shadow stage)
    int* list[2];     // list[0]: live paths (input of extend); list[1]: paths to shade (output of shade A)
    int* counts;      // counts[0], counts[1]
    int P;            // pixels of this launch (width * rows of the shard)
    FastDiv div_P, div_W;   // slot -> (sample in batch, pixel), pixel -> (row, column)
    int S;            // samples of the current batch
    int it_base;      // first sample index of the batch
    int nslots;       // P * S
    int first, last;  // first / last batch of the launch
    int num_sms;
    int num_sets;           // 1..kWfMaxSets copies of the state (set k = set 0 + k * set_bytes), one stream each
    size_t set_bytes;
    // Ray binning (k_wf_bin): the rays of a traversal stage are grouped by (origin cell, direction octant) so that the 32 rays
    // of a warp walk the same part of the tree.  The producing stage writes a key per list entry and counts the bins; one
    // kernel per stage turns that into the binned list the traversal reads.  Shade stages keep the slot-ordered list.
    unsigned short* keys[2];   // bin of list[k][i]
    int* blist;             // binned copy of the list the next traversal stage reads
    int* hist[2];           // kWfBins counters per list, zero between uses
    int* cur;               // kWfBins cursors of k_wf_bin, zero between uses
    float3 key_lo, key_scale;   // origin cell = (o - key_lo) * key_scale, clamped to [0, key_cells)
    int key_cells;              // 8: bin = (8^3 origin cells, direction octant); 16: bin = 16^3 origin cells (large scenes: +1.6 % on C5, r03m)
    int bin_mask;           // scene-level switch: bit 0 = extension rays, bit 1 = shadow rays (from bounce 1 on)
    int bin_in;             // this launch (traversal): read blist instead of list[k]
    int bin_out;            // this launch (generate / shade): write keys + hist for the list it appends to
    int tiny_regs;          // > 0: scene fits the shared-memory traversal kernel with a register stack of 2 * tiny_regs entries
};
// Which traversal stages read binned lists when ZDR_WF_BIN is not set (bit 0 extension rays, bit 1 shadow rays), for scenes of at
// most / more than kWfBigSceneNodes nodes.  Measured (profiles/r03j, r03k): the Cornell box loses (C3 7 360 -> 6 859 Mrays/s: its
// traversal is short, the binned traversal gathers ray records at random from a 4 GB array), the 9.6 M-triangle scene gains
// (C5 1 974 -> 2 085: rays that start in the same cell and octant share the top of a deep tree and its cache lines).
#ifndef ZDR_WF_BIN_DEFAULT
#define ZDR_WF_BIN_DEFAULT 0
#endif
#ifndef ZDR_WF_BIN_DEFAULT_BIG
#define ZDR_WF_BIN_DEFAULT_BIG 3
#endif
constexpr int kWfBigSceneNodes = 1024;
constexpr int kWfBins = 4096;      // 8^3 origin cells x 8 direction octants, or 16^3 origin cells (WfState.key_cells)

// Persistent per-SM wavefront (persistent.cu, k_pw): one resident CTA per SM keeps a pool of paths in SHARED memory and
// runs the stages of the wavefront as barrier-separated phases over shared-memory queues.  Nothing but finished samples
// (forward) leaves the SM.
struct PwState {
    float4* R;            // forward: finished samples of the batch, [sample-in-batch * P + pixel] (clamped radiance.xyz | -)
    float4* gstate;       // build variant ZDR_PW_GSTATE: the pools' state records in global memory (7 * pool float4 per CTA)
    unsigned* next;       // work counter of the launch: camera samples handed out so far (zeroed by the launcher)
    unsigned total;       // P * S camera samples of the batch
    int P;                // pixels of this launch (width * rows of the shard)
    FastDiv div_P, div_W; // sample id -> (sample in batch, pixel), pixel -> (row, column)
    int S, it_base;       // samples per pixel of the batch, first sample index
    int first, last;      // first / last batch of the launch
    int pool;             // path slots per CTA
    int stack_depth;      // SMALL scenes: entries of the per-thread traversal stack in shared memory
    int bulk_copy;        // SMALL scenes: stage the scene with cp.async.bulk + mbarrier (else a cooperative copy loop)
    int refill_closest, refill_any;   // a warp of the traversal phases refills its idle lanes once fewer than this many are busy
    unsigned long long* phase_cycles; // profile (ZDR_PW_PROFILE): 8 accumulators, or null
};

struct RenderArgs {
    SceneView sc;
    TexView tex;
    // camera basis, computed once per launch on the host exactly as camera.py:7-10 does per ray
    float3 cam_o, cam_fwd, cam_right, cam_up;
    float tan_half_fov;
    int W, H, spp, seed, use_tent, max_depth, rr_depth, use_mis, prb_mode;
    int row0, row1, it0, it1;
    int shard_index, shard_count;
    int block_scatter;        // with aggregate_scatter: also the CTA-level shared-memory duplicate (shading.cuh GradCache)
    int aggregate_scatter;    // warp-aggregated gradient scatter (shading.cuh) instead of plain red.v4   // 8-row CTA tiles of [row0,row1) are dealt round-robin to shard_count shards
    float4* image;            // forward: (H,W) float4
    const float4* d_image;    // backward: (H,W) float4 cotangent
    float* samples_out;       // probe: (H,W,spp,3) un-clamped per-sample radiance, or null
    unsigned long long* counters;
    WfState wf;
    PwState pw;
};

// number of kernels this library has launched (render kernels and probes), for bench.py's gpu_launches
unsigned long long kernel_launch_count();
// launchers (render.cu; the wavefront schedule: wavefront.cu)
cudaError_t launch_forward(int integrator, const RenderArgs& a, bool stats, cudaStream_t stream);
cudaError_t launch_backward(int integrator, const RenderArgs& a, bool stats, cudaStream_t stream);
// wavefront variant of the path integrator / PRB adjoint; `wf` buffers sized for wf_slot_capacity() slots
size_t wf_bytes_per_slot(bool binned);      // binned: with the arrays of the ray binning (WfState)
size_t wf_fixed_bytes(bool binned);         // per state set, on top of the slots
// side / events: the scene's kWfMaxSets - 1 extra streams and kWfMaxSets + 2 timing-less events, used when
// wf.num_sets > 1 to run consecutive batches from several streams (the tail of every stage kernel overlaps the other
// streams' work); both may be null otherwise
constexpr int kWfMaxSets = 4;
cudaError_t launch_path_wavefront(bool backward, RenderArgs a, bool stats, cudaStream_t stream, cudaStream_t* side,
                                  cudaEvent_t* events);
// persistent shared-memory schedule of the path integrator / PRB adjoint (persistent.cu).  `a.pw.R` (forward) holds
// pw_forward_slots() float4 records; `counter_dev`: one unsigned per batch, zeroed by the call.
struct PwPlan {
    int pool, threads, stack_depth, small, grid;
    size_t smem_bytes;
    size_t gstate_bytes;      // global-memory path state (build variant ZDR_PW_GSTATE), else 0
};
cudaError_t pw_plan(const SceneView& sc, int bvh_depth, int num_sms, PwPlan* out);
cudaError_t launch_path_persistent(bool backward, RenderArgs a, bool stats, const PwPlan& plan, size_t r_capacity_slots,
                                   unsigned* counters_dev, int max_batches, cudaStream_t stream);
cudaError_t launch_duvdxy(const RenderArgs& a, cudaStream_t stream);
// The "fast" build of the render kernels (ZDR_FLAG_FAST_MATH): render_fast.cu / persistent_fast.cu are the same sources
// compiled with --use_fast_math and ZDR_FAST_MATH (approximate division / square root / sin / cos, contraction, denormals
// flushed) — except the sampler and the ray / triangle test, which are written with explicit IEEE operations and stay
// bit-exact.  Same arguments as their exact twins.
cudaError_t launch_forward_fast(int integrator, const RenderArgs& a, bool stats, cudaStream_t stream);
cudaError_t launch_backward_fast(int integrator, const RenderArgs& a, bool stats, cudaStream_t stream);
cudaError_t launch_path_wavefront_fast(bool backward, RenderArgs a, bool stats, cudaStream_t stream, cudaStream_t* side,
                                       cudaEvent_t* events);
cudaError_t launch_path_persistent_fast(bool backward, RenderArgs a, bool stats, const PwPlan& plan, size_t r_capacity_slots,
                                        unsigned* counters_dev, int max_batches, cudaStream_t stream);
cudaError_t launch_probe_trace(const SceneView& sc, const float* rays_dev, int n, bool any, int* ip_dev,
                               float* uvt_dev, int* occ_dev, cudaStream_t stream);
cudaError_t launch_probe_sampler(const SceneView& sc, const int* state_dev, int n, const char* prog_dev, int prog_len,
                                 int per, float* out_dev, cudaStream_t stream);
cudaError_t launch_probe_ggx(const float* in_dev, int n, float* out_dev, cudaStream_t stream);

}  // namespace zdr
