"""Which schedule of the path integrator wins at which job size: C3 scene / camera / depth at image sizes from 64^2 x 4 spp to
1024^2 x 64 spp, forward + adjoint, every schedule.  usage: python tools/schedule_sweep.py [--json out.json]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import scenes
from zdr_b200 import Scene
from zdr_b200._abi import INTEGRATORS, check, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
args = ap.parse_args()
w = scenes.workload("c3")
sc = Scene(w.models(), integrator="path")
mat_np = w.material()
mat = torch.from_numpy(mat_np).cuda()
st = torch.cuda.current_stream().cuda_stream
rows = []
for res, spp in ((64, 4), (64, 16), (128, 16), (256, 16), (256, 64), (512, 64), (362, 256), (724, 256), (1024, 64)):
    img = torch.zeros((res, res, 4), device="cuda")
    dimg, dmat = torch.ones_like(img), torch.zeros_like(mat)
    row = {"res": res, "spp": spp, "log2_samples": round(float(torch.log2(torch.tensor(float(res * res * spp)))), 1)}
    for sched in ("megakernel", "wavefront", "persistent"):
        pf = make_params(w.camera, (res, res), mat_np.shape, spp, 0, max_depth=w.depth, schedule=sched)
        pb = make_params(w.camera, (res, res), mat_np.shape, spp, 1, max_depth=w.depth, schedule=sched)

        def step():
            check(sc._lib.zdr_render_forward(sc._handle, INTEGRATORS["path"], C.byref(pf), mat.data_ptr(), img.data_ptr(), st))
            check(sc._lib.zdr_render_backward(sc._handle, INTEGRATORS["path"], C.byref(pb), mat.data_ptr(), dimg.data_ptr(), dmat.data_ptr(), st))
        if sched == "megakernel" and res * res * spp > (1 << 25):
            continue
        sc.reset_counters()
        step()
        c = sc.counters()
        rays = c["closest_rays"] + c["any_rays"]
        for _ in range(2):
            step()
        best = 1e30
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row[sched] = {"ms": round(best, 3), "mrays_per_s": round(rays / best / 1e3, 1)}
    rows.append(row)
    print(json.dumps(row), flush=True)
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
