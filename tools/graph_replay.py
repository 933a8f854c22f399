"""Launch-bound jobs: the operator captured into a CUDA graph vs eager launches (forward + adjoint per step).

    python tools/graph_replay.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import scenes
from zdr_b200 import Scene
from zdr_b200._abi import INTEGRATORS, check, make_params

JOBS = [("C1 collocated 256^2 1 spp", scenes.cbox_collocated_models(), "collocated", 256, 1, "auto"),
        ("direct 256^2 4 spp", scenes.cbox_path_models(), "direct", 256, 4, "auto"),
        ("path depth 8 256^2 4 spp, one kernel", scenes.cbox_path_models(), "path", 256, 4, "megakernel"),
        ("path depth 8 256^2 4 spp, wavefront (100 launches)", scenes.cbox_path_models(), "path", 256, 4, "wavefront"),
        ("path depth 8 512^2 16 spp, wavefront", scenes.cbox_path_models(), "path", 512, 16, "wavefront")]
print("| job | eager ms/step | graph replay ms/step |\n|---|---:|---:|")
for name, models, integ_name, res, spp, sched in JOBS:
    sc = Scene(models, integrator=integ_name)
    lib, h, integ = sc._lib, sc._handle, INTEGRATORS[integ_name]
    material = torch.from_numpy(scenes.random_material((1024, 1024), 0)).cuda()
    image = torch.zeros((res, res, 4), device="cuda")
    d_image, d_material = torch.ones_like(image), torch.zeros_like(material)
    pf = make_params(scenes.CBOX_CAMERA, (res, res), (1024, 1024), spp, 0, max_depth=8, schedule=sched)
    pb = make_params(scenes.CBOX_CAMERA, (res, res), (1024, 1024), spp, 1, max_depth=8, schedule=sched)

    def run():
        st = torch.cuda.current_stream().cuda_stream
        check(lib.zdr_render_forward(h, integ, C.byref(pf), material.data_ptr(), image.data_ptr(), st))
        check(lib.zdr_render_backward(h, integ, C.byref(pb), material.data_ptr(), d_image.data_ptr(), d_material.data_ptr(), st))

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            run()

        def timed(fn, n=50):
            for _ in range(5):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        t_eager, t_graph = timed(run), timed(graph.replay)
    print(f"| {name} | {t_eager:.3f} | {t_graph:.3f} |")
    sc.close()
