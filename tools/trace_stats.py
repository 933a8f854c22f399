"""Prints BVH work per ray (nodes visited, triangles tested) for the bench workload at low spp."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import scenes
from zdr_b200 import Scene

for integ, models in (("path", scenes.cbox_path_models()), ("direct", scenes.cbox_direct_models())):
    sc = Scene(models, integrator=integ)
    sc.camera = scenes.CBOX_CAMERA.copy()
    sc.max_depth = 8
    sc.collect_stats = True
    m = torch.from_numpy(scenes.random_material((1024, 1024), 0)).cuda().requires_grad_()
    sc.reset_counters()
    img = sc.render(m, res=(512, 512), spp=8)
    torch.cuda.synchronize()
    c = sc.counters()
    rays = c["closest_rays"] + c["any_rays"]
    print(integ, "fwd", c, "nodes/ray %.2f tris/ray %.2f" % (c["nodes_visited"] / rays, c["tris_tested"] / rays))
