"""GPU next to CPU for BASELINE.json configs C1-C4 in one invocation (SURVEY.md §8d "CPU baseline beside it").

    python tools/run_configs.py [--json gpurun_out/configs.json]

Per config: scene + material as §8d names them, one forward + one adjoint render = a step.  GPU: the C ABI with
device buffers on torch's stream, CUDA events, best of 5 steps after 3 warm-ups; rays from the device counters of
one instrumented step.  CPU: the oracle (plain-C restatement, all host threads) on the bounded sample stated in
the row (full size where that takes seconds).  C5 has its own script (tools/run_c5.py: 9.6 M triangles).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import oracle as o
from bench import scenes
from zdr_b200 import Scene
from zdr_b200._abi import INTEGRATORS, check, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
ap.add_argument("--only", default=None, help="comma-separated config names")
args = ap.parse_args()
o.build()
o.lib().zo_set_num_threads(os.cpu_count() or 1)
dev = torch.device("cuda", 0)


def light_stage_models():
    quad = scenes.asset("quad.obj")
    lights = [(quad, scenes.translate_scale(2.5 * np.cos(a), 2.0, 2.5 * np.sin(a), 0.6), scenes.float3(0.0))
              for a in np.linspace(0, 2 * np.pi, 8, endpoint=False)]
    return [(scenes.asset("sphere.obj"), None, scenes.float3(0.0))] + lights


LS_CAMERA = scenes.Camera(fov=0.7, origin=scenes.float3(0.0, 0.6, 3.2), target=scenes.float3(0.0, 0.0, 0.0), up=scenes.float3(0, 1, 0))


def emissions_for(k, n=8):
    return [(0.0, 0.0, 0.0)] + [(12.0, 12.0, 12.0) if j == k else (0.0, 0.0, 0.0) for j in range(n)]


# name, models, camera, integrator, res, spp, depth, material, cpu its, light-stage loop
CONFIGS = [
    ("C1 collocated 256^2 1 spp", scenes.cbox_collocated_models(), scenes.CBOX_CAMERA, "collocated", 256, 1, 8,
     scenes.random_material((1024, 1024), 0), None, 0),
    ("C2 direct 512^2 64 spp (uv = 0)", scenes.cbox_direct_models(), scenes.CBOX_CAMERA, "direct", 512, 64, 8,
     scenes.random_material((1024, 1024), 0), None, 0),
    ("C3 path/PRB depth 8 1024^2 256 spp", scenes.cbox_path_models(), scenes.CBOX_CAMERA, "path", 1024, 256, 8,
     scenes.random_material((1024, 1024), 0), (0, 16), 0),
    ("C4 light stage: 8 lights x direct 1024^2 4 spp, constant 16^2 texture", light_stage_models(), LS_CAMERA, "direct", 1024, 4, 8,
     scenes.constant_material((16, 16)), None, 8),
    ("C4 light stage: 8 lights x collocated 1024^2 1 spp, 1x1 texture", light_stage_models(), LS_CAMERA, "collocated", 1024, 1, 8,
     scenes.constant_material((1, 1)), None, 8),
]

rows = []
for name, models, cam, integ_name, res, spp, depth, mat_np, cpu_its, n_lights in CONFIGS:
    if args.only and not any(name.startswith(k) for k in args.only.split(",")):
        continue
    integ = INTEGRATORS[integ_name]
    sc = Scene(models, integrator=integ_name, device=0)
    lib, h = sc._lib, sc._handle
    material = torch.from_numpy(mat_np).to(dev)
    image = torch.zeros((res, res, 4), dtype=torch.float32, device=dev)
    d_image = torch.ones_like(image)
    d_material = torch.zeros_like(material)
    stream = torch.cuda.current_stream(dev)
    loops = max(1, n_lights)

    def step(stats=False):
        for k in range(loops):
            if n_lights:
                sc.update_lights(emissions_for(k))          # test_lightstage.py:46-54: one light on per render
            pf = make_params(cam, (res, res), mat_np.shape, spp, 0, max_depth=depth, stats=stats)
            pb = make_params(cam, (res, res), mat_np.shape, spp, 1, max_depth=depth, stats=stats)
            check(lib.zdr_render_forward(h, integ, C.byref(pf), material.data_ptr(), image.data_ptr(), stream.cuda_stream))
            check(lib.zdr_render_backward(h, integ, C.byref(pb), material.data_ptr(), d_image.data_ptr(), d_material.data_ptr(), stream.cuda_stream))

    sc.reset_counters()
    step(stats=True)
    torch.cuda.synchronize()
    c = sc.counters()
    rays = c["closest_rays"] + c["any_rays"]
    for _ in range(3):
        step()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # ---- CPU oracle, bounded sample
    osc = o.OracleScene(models)
    its = cpu_its or (0, spp)
    cpu_loops = min(loops, 2)
    osc.reset_ray_counters()
    t0 = time.perf_counter()
    for k in range(cpu_loops):
        if n_lights:
            osc.update_lights(emissions_for(k))
        pf = make_params(cam, (res, res), mat_np.shape, spp, 0, max_depth=depth, its=its)
        pb = make_params(cam, (res, res), mat_np.shape, spp, 1, max_depth=depth, its=its)
        osc.render_forward(integ, pf, mat_np)
        osc.render_backward(integ, pb, mat_np, np.ones((res, res, 4), np.float32))
    dt = time.perf_counter() - t0
    cpu_rays = sum(osc.ray_counters())
    row = {"config": name, "gpu_ms_per_step": round(best, 3), "rays_per_step": rays, "gpu_mrays_per_s": round(rays / best / 1e3, 1),
           "red_v4_per_scatter": round(c["red_v4"] / max(c["grad_scatters"], 1), 4), "grad_scatters": c["grad_scatters"],
           "cpu_mrays_per_s": round(cpu_rays / dt / 1e6, 2), "cpu_threads": o.lib().zo_num_threads(),
           "cpu_sample": f"samples [{its[0]},{its[1]}) of {spp}" + (f", {cpu_loops} of {loops} lights" if n_lights else "") + f", {dt:.1f} s",
           "speedup": round((rays / best / 1e3) / (cpu_rays / dt / 1e6), 1)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    sc.close()

print("\n| config | GPU ms/step | rays/step | GPU Mrays/s | red.v4 per scatter | CPU Mrays/s (threads; sample) | GPU/CPU |")
print("|---|---:|---:|---:|---:|---|---:|")
for r in rows:
    print(f"| {r['config']} | {r['gpu_ms_per_step']} | {r['rays_per_step']:,} | {r['gpu_mrays_per_s']} | {r['red_v4_per_scatter']} | "
          f"{r['cpu_mrays_per_s']} ({r['cpu_threads']}; {r['cpu_sample']}) | {r['speedup']}x |")
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
