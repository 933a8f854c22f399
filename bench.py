#!/usr/bin/env python
"""bench.py — PRB fwd+bwd Mrays/s on the Cornell box (BASELINE.json metric).

A step = one forward render + one adjoint (PRB) render of the workload, `d_image = 1`
(benchmark.py:37-39: `I.sum().backward()`).  Default workload = BASELINE.json configs[2] (C3):
cboxuv.obj + cbox-light.obj (17,12,4), path integrator, max depth 8, rr depth 2, 1024x1024,
256 spp, 1024^2 random texture, seed 0 (fwd) / 1 (bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our CUDA path
    python bench.py --impl reference ...                          # CPU restatement (oracle), all host threads
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0.  `value` = rays (closest + any queries of forward pass, adjoint replay
pass and adjoint walk, device counters) of ALL ranks / max-over-ranks device time; inputs
resident in HBM.  `e2e` = same metric through the torch surface (`Scene.render` +
`.backward()`) with the material coming from pinned host memory and the gradient + loss going
back to the host every step.  N > 1: strong scaling — the image's 8-row tiles are dealt
round-robin to the ranks, texture gradients are summed with one NCCL all-reduce inside the step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "prb_fwd_bwd_mrays_per_s"
UNIT = "Mrays/s"

# algorithmic bytes per unit of work (DESIGN.md "Algorithmic bytes")
B_NODE = 64            # one BVH2 node record (both child boxes + links)
B_TRI = 48             # one triangle record (3 x float4)
B_SHADE = 48 + 64 + 64 + 36      # hit triangle + shading record + 4 texels + instance normal matrix
B_LIGHT = 4 + 48                 # slot_of_gid + light triangle (mesh-light NEE sample)
B_SAMPLER_VERTEX = 4 * 4 + 2 * 16  # 4 one-dimensional + 2 two-dimensional draws per path vertex
B_SCATTER = 128        # 4 texels x 16 B, read-modify-write
B_CAMERA = 16          # the camera sample's 2-D draw
B_PIXEL = 16           # image write (fwd) / d_image read (bwd)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--tex", type=int, default=1024)
    ap.add_argument("--integrator", default="path", choices=["path", "direct", "collocated"])
    ap.add_argument("--ref-spp", type=int, default=2, help="samples per pixel (of --spp) one reference step renders")
    ap.add_argument("--cpu-spp", type=int, default=32, help="samples per pixel of the cpu_baseline sample")
    ap.add_argument("--schedule", default="auto", choices=["auto", "wavefront", "megakernel"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(args):
    from tests import scenes
    models = {"path": scenes.cbox_path_models, "direct": scenes.cbox_direct_models,
              "collocated": scenes.cbox_collocated_models}[args.integrator]()
    material = scenes.random_material((args.tex, args.tex), 0)
    name = (f"cornell-box {args.integrator} (cboxuv.obj + cbox-light.obj) max_depth {args.depth} rr_depth 2, "
            f"{args.res}x{args.res}, {args.spp} spp, texture {args.tex}^2 random, fwd seed 0 / bwd seed 1, d_image=1")
    return models, scenes.CBOX_CAMERA, material, name


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(c: dict, pixels: int) -> int:
    """Σ bytes the hot path has to touch for the counted work."""
    return (c["nodes_visited"] * B_NODE + c["tris_tested"] * B_TRI + c["shaded_vertices"] * (B_SHADE + B_LIGHT + B_SAMPLER_VERTEX)
            + c["grad_scatters"] * B_SCATTER + c["camera_samples"] * B_CAMERA + pixels * B_PIXEL)


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path: the plain-C restatement in oracle/
    (the reference itself is LuisaCompute-JIT Python and cannot be installed here),
    all host threads, each step = samples [0, ref_spp) of every pixel of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as o
    from zdr_b200._abi import INTEGRATORS, make_params
    o.build()
    o.lib().zo_set_num_threads(os.cpu_count() or 1)      # torchrun sets OMP_NUM_THREADS=1 for its workers
    models, cam, material, name = workload(args)
    sc = o.OracleScene(models)
    integ = INTEGRATORS[args.integrator]
    k = min(args.ref_spp, args.spp)
    pf = make_params(cam, (args.res, args.res), material.shape, args.spp, 0, max_depth=args.depth, its=(0, k))
    pb = make_params(cam, (args.res, args.res), material.shape, args.spp, 1, max_depth=args.depth, its=(0, k))
    d_img = np.ones((args.res, args.res, 4), np.float32)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        sc.reset_ray_counters()
        t0 = time.perf_counter()
        sc.render_forward(integ, pf, material)
        sc.render_backward(integ, pb, material, d_img)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays = sum(sc.ray_counters())
    ms = 1e3 * sum(times) / len(times)
    value = rays / (ms * 1e-3) / 1e6
    cores = o.lib().zo_num_threads()
    sample = f"{args.res}x{args.res}, samples [0,{k}) of {args.spp} spp per pixel, fwd+bwd, {rays} rays/step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": name, "sample_per_step": sample},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(args):
    from oracle import oracle as o
    from zdr_b200._abi import INTEGRATORS, make_params
    o.build()
    o.lib().zo_set_num_threads(os.cpu_count() or 1)
    models, cam, material, _ = workload(args)
    sc = o.OracleScene(models)
    integ = INTEGRATORS[args.integrator]
    k = min(args.cpu_spp, args.spp)
    pf = make_params(cam, (args.res, args.res), material.shape, args.spp, 0, max_depth=args.depth, its=(0, k))
    pb = make_params(cam, (args.res, args.res), material.shape, args.spp, 1, max_depth=args.depth, its=(0, k))
    d_img = np.ones((args.res, args.res, 4), np.float32)
    sc.reset_ray_counters()
    t0 = time.perf_counter()
    sc.render_forward(integ, pf, material)
    sc.render_backward(integ, pb, material, d_img)
    dt = time.perf_counter() - t0
    rays = sum(sc.ray_counters())
    return {"value": round(rays / dt / 1e6, 3), "unit": UNIT, "cores": o.lib().zo_num_threads(), "kind": "port",
            "sample": f"{args.res}x{args.res}, samples [0,{k}) of {args.spp} spp per pixel, fwd+bwd, {rays} rays in {dt:.2f} s"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from zdr_b200 import Scene
    from zdr_b200._abi import INTEGRATORS, check, make_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    models, cam, mat_np, name = workload(args)
    scene = Scene(models, integrator=args.integrator, device=local_rank)
    scene.camera = cam.copy()
    scene.max_depth = args.depth
    scene.schedule = args.schedule
    if world > 1:
        scene.shard = (rank, world)
    lib, h = scene._lib, scene._handle
    integ = INTEGRATORS[args.integrator]
    res = (args.res, args.res)
    shard = (rank, world) if world > 1 else None
    pf = make_params(cam, res, mat_np.shape, args.spp, 0, max_depth=args.depth, shard=shard, schedule=args.schedule)
    pb = make_params(cam, res, mat_np.shape, args.spp, 1, max_depth=args.depth, shard=shard, schedule=args.schedule)
    material = torch.from_numpy(mat_np).to(dev)
    image = torch.zeros((args.res, args.res, 4), dtype=torch.float32, device=dev)
    d_image = torch.ones_like(image)
    d_material = torch.zeros_like(material)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def step(ev=None):
        d_material.zero_()
        if ev:
            ev[0].record(stream)
        check(lib.zdr_render_forward(h, integ, C.byref(pf), material.data_ptr(), image.data_ptr(), stream.cuda_stream))
        if ev:
            ev[1].record(stream)
        check(lib.zdr_render_backward(h, integ, C.byref(pb), material.data_ptr(), d_image.data_ptr(), d_material.data_ptr(),
                                      stream.cuda_stream))
        if ev:
            ev[2].record(stream)
        if world > 1:
            dist.all_reduce(d_material)
        if ev:
            ev[3].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # one instrumented pass (untimed): rays per step and the algorithmic bytes of each kernel
    scene.reset_counters()
    pfs = make_params(cam, res, mat_np.shape, args.spp, 0, max_depth=args.depth, shard=shard, stats=True, schedule=args.schedule)
    pbs = make_params(cam, res, mat_np.shape, args.spp, 1, max_depth=args.depth, shard=shard, stats=True, schedule=args.schedule)
    check(lib.zdr_render_forward(h, integ, C.byref(pfs), material.data_ptr(), image.data_ptr(), stream.cuda_stream))
    c_fwd = scene.counters()
    scene.reset_counters()
    check(lib.zdr_render_backward(h, integ, C.byref(pbs), material.data_ptr(), d_image.data_ptr(), d_material.data_ptr(),
                                  stream.cuda_stream))
    c_bwd = scene.counters()
    rays_local = c_fwd["closest_rays"] + c_fwd["any_rays"] + c_bwd["closest_rays"] + c_bwd["any_rays"]
    for _ in range(max(0, args.warmup - 1)):
        step()
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("ZDR_BENCH_NO_CLOCKS"):      # (experiment switch: is the sampler itself a disturbance?)
        clocks.start()
    launches0 = int(lib.zdr_kernel_launches())
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                       # L2 flush between timed iterations (outside the events)
        step(evs[k])
    barrier()
    launches = int(lib.zdr_kernel_launches()) - launches0
    clock_rec = clocks.stop() if rank == 0 else None
    t_fwd = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_bwd = sum(e[1].elapsed_time(e[2]) for e in evs)
    t_all = sum(e[0].elapsed_time(e[3]) for e in evs)

    # ---- e2e: torch surface, host buffers in the timed region
    e2e = None
    if not args.no_e2e:
        host_mat = torch.from_numpy(mat_np).pin_memory()
        host_grad = torch.empty_like(host_mat).pin_memory()

        def e2e_step():
            m = host_mat.to(dev, non_blocking=True).requires_grad_()
            img = scene.render(m, res=res, spp=args.spp, seed=0)
            loss = img.sum()
            loss.backward()
            g = m.grad
            if world > 1:
                dist.all_reduce(g)
            host_grad.copy_(g, non_blocking=True)
            return float(loss.item())
        for _ in range(1):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        t_e2e = (time.perf_counter() - t0) / n_e2e
    # ---- reductions over ranks
    stats = torch.tensor([t_all, t_fwd, t_bwd, float(rays_local), t_e2e if not args.no_e2e else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_all, t_fwd, t_bwd, t_e2e = mx[0].item(), mx[1].item(), mx[2].item(), mx[4].item()
        rays_total = sm[3].item()
    else:
        rays_total = float(rays_local)
        t_e2e = stats[4].item()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = t_all / args.steps
    value = rays_total / (ms_per_step * 1e-3) / 1e6
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    pixels_local = (args.res * args.res + world - 1) // world
    bytes_bwd = algorithmic_bytes(c_bwd, pixels_local)
    ms_bwd = t_bwd / args.steps
    achieved = bytes_bwd / (ms_bwd * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath))
        if rec.get("res") == args.res and rec.get("integrator") == args.integrator and world == 1:
            # DRAM bytes of one adjoint pass measured with ncu at rec["spp"]; batches of 16 samples are identical, so it scales with spp
            traffic = int(rec["backward_dram_bytes"] * args.spp / rec["spp"])
    # the other two denominators SURVEY.md 8(d) asks for, measured by tools/microbench (profiles/r01q_microbench.json)
    other = None
    mpath = os.path.join(ROOT, "profiles", "r01q_microbench.json")
    if os.path.exists(mpath):
        mb = json.load(open(mpath))
        l2_64 = next(g["gb_per_s"] for g in mb["gather"] if g["working_set_mib"] == 16 and g["record_bytes"] == 64)
        red_peak = next(r["gops_per_s"] for r in mb["red"] if r["op"].endswith("v4.f32") and r["addresses"].startswith("random"))
        red_rate = c_bwd["red_v4"] / (ms_bwd * 1e-3) / 1e9
        other = {"l2_gather_peak_gbs_64B_records": l2_64, "achieved_over_l2_gather": round(achieved / l2_64, 3),
                 "red_v4_g_per_s": round(red_rate, 2), "red_v4_peak_g_per_s": red_peak, "red_frac": round(red_rate / red_peak, 4),
                 "binding": ("none of the three: the BVH (4 KB) is L1-resident, texel gathers are ~2 % of the L2 gather rate and the scatter "
                             "~2 % of the reduction rate; the stage kernels issue 60-76 % of their cycles (ncu, profiles/r01m) at 19-30 of 32 "
                             "lanes, i.e. the pass is bound by instruction issue under divergence, with HBM latency on the path state next")}
    wavefront = args.integrator == "path" and args.schedule != "megakernel" and (args.schedule == "wavefront" or pixels_local * args.spp >= 1 << 21)
    if wavefront:
        kernel_name = ("adjoint pass of the wavefront path tracer: per batch k_wf_generate + 2*depth x [k_wf_trace<closest>, k_wf_shade_a, "
                       "k_wf_trace<any>, k_wf_shade_b] + k_wf_restart (PRB replay + adjoint walk + gradient scatter); two half-batches "
                       "run concurrently on two streams")
    else:
        kernel_name = f"adjoint kernel of the {args.integrator} integrator (k_path<BWD> / k_backward: replay + adjoint + gradient scatter)"
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "l2": "flushed between timed iterations (256 MiB fill)", "parallelism": f"tile-dp{world}",
                   "rays_per_step": int(rays_total), "paths_per_step": 2 * args.res * args.res * args.spp,
                   "ms_forward": round(t_fwd / args.steps, 3), "ms_backward": round(ms_bwd, 3),
                   "ms_each_step_rank0": [round(e[0].elapsed_time(e[3]), 2) for e in evs],
                   "ms_allreduce": round((t_all - t_fwd - t_bwd) / args.steps, 3),
                   "bvh_nodes_per_ray": round((c_fwd["nodes_visited"] + c_bwd["nodes_visited"]) / max(rays_local, 1), 2),
                   "tris_per_ray": round((c_fwd["tris_tested"] + c_bwd["tris_tested"]) / max(rays_local, 1), 2),
                   "red_v4_per_scatter": round(c_bwd["red_v4"] / max(c_bwd["grad_scatters"], 1), 3)},
        "clocks": clock_rec,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": kernel_name,
                     "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bytes_bwd),
                     "ms_per_launch": round(ms_bwd, 3), "other_denominators": other,
                     "note": ("algorithmic bytes = BVH nodes/triangles visited + shading records + texels + scatter footprints (DESIGN.md 6); on "
                              "the Cornell box these are L1/L2 resident. `traffic` = DRAM bytes ncu measured for the same pass: with the "
                              "wavefront schedule it is the path state streamed through HBM between stages")},
    }
    if not args.no_e2e:
        out["e2e"] = {"value": round(rays_total / t_e2e / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(mat_np.nbytes),
                      "d2h_bytes_per_step": int(mat_np.nbytes) + 4, "ms_per_step": round(t_e2e * 1e3, 3),
                      "api": "zdr_b200.Scene.render(...).sum().backward() with pinned-host material in, gradient + loss out"}
    if not args.no_cpu_baseline and world == 1:          # reported at N=1 only (it times the host, not the GPUs)
        out["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
