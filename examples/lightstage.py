"""The reference's test_lightstage.py on zdr_b200: an object inside a light stage of `nlight` quad emitters placed on a
sphere by rotate_mat (test_lightstage.py:27-55), lights switched with Scene.update_lights (:66), one ground-truth render per
lighting condition.  bunnyuv.obj and the wood_olive textures are not in the reference's checkout: sphere.obj and the
Cornell textures stand in.  One Scene serves every lighting condition — update_lights rewrites 12 bytes per instance.

    python examples/lightstage.py [--nlight 30] [--res 1024] [--spp 1024] [--out results]
"""
import argparse
import os
import time
from math import acos, pi

from common import asset, load_material, rotate_mat, save_png

import torch

from zdr_b200 import Camera, Scene, float3          # reference: from zdr import Scene, float3, Camera

ap = argparse.ArgumentParser()
ap.add_argument("--nlight", type=int, default=30)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--out", default="results")
args = ap.parse_args()

models = [(asset("sphere.obj"), rotate_mat(0, -0.4, float3(0, 0, 0)), None)]
nlight = args.nlight
for i in range(nlight):                                       # golden-angle spiral, test_lightstage.py:52-55 (all of them, not just i == 29)
    models.append((asset("quad.obj"), rotate_mat(acos((i + 0.5) / nlight * 2 - 1), pi * 2 * 0.618 * (i + 1), float3(0)), 50))

sphere_camera1 = Camera(fov=50 / 180 * 3.1415926, origin=float3(0, 0.5, 3), target=float3(0, 0, 0), up=float3(0.0, 1.0, 0.0))
scene = Scene(models, integrator="direct")
scene.camera = sphere_camera1
material_GT = load_material(asset("cboxd.png"), asset("cboxr.png"))
ImgRes = args.res, args.res
print("Image resolution:", ImgRes)

torch.cuda.synchronize()
t0 = time.perf_counter()
means = []
for k in range(nlight):                                       # one light at a time (OLAT)
    scene.update_lights([None] + [50 if i == k else 0 for i in range(nlight)])
    I_GT = scene.render(material_GT, res=ImgRes, spp=args.spp)          # seed defaults to 0
    means.append(float(I_GT[..., 0:3].mean()))
    if k in (0, nlight - 1):
        save_png(os.path.join(args.out, f"lightstage_{k:02d}.png"), I_GT)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{nlight} lighting conditions at {args.spp} spp in {dt:.2f} s ({dt / nlight * 1e3:.1f} ms each); "
      f"lit conditions: {sum(m > 0 for m in means)}, mean radiance {sum(means) / nlight:.5f}")
scene.update_lights([None] + [50] * nlight)                   # all on
save_png(os.path.join(args.out, "lightstage_all.png"), scene.render(material_GT, res=ImgRes, spp=args.spp))
