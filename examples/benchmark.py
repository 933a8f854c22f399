"""The reference's benchmark.py on zdr_b200: forward + backward of the direct integrator on a textured sphere under one
area light, 512 x 512 at 256 spp, a fresh random seed per iteration (benchmark.py:20-41).

Assets the reference's checkout lacks are replaced: `assets/lightstage/l00.obj` by quad.obj (a 0.42 x 0.42 emitter 5 units above the sphere), the
wood_olive textures by the Cornell textures cboxd.png / cboxr.png that ARE in its checkout (example.py:47-48).

    python examples/benchmark.py [--iters 100] [--res 512] [--spp 256] [--math fast]
"""
import argparse
import random
import time

from common import asset, load_material

import torch

from zdr_b200 import Scene          # reference: from zdr import Scene

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=100)            # benchmark.py:39 runs 1000
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--math", default="exact", choices=["exact", "fast"])
args = ap.parse_args()

scene = Scene([(asset("sphere.obj"), None, 0), (asset("quad.obj"), None, 100)], integrator="direct")
scene.math = args.math
material_GT = load_material(asset("cboxd.png"), asset("cboxr.png"))

ImgRes = args.res, args.res
spp = args.spp
TexRes = material_GT.shape[0:2]
print("Image resolution:", ImgRes)
print("Texture resolution:", tuple(TexRes))
print("spp:", spp)

print("Forward + backward:")
material = material_GT.clone().requires_grad_()
for it in range(3):                                          # warm-up (allocations, clocks)
    scene.render(material, res=ImgRes, spp=spp, seed=it).sum().backward()
scene.reset_counters()
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(args.iters):
    I = scene.render(material, res=ImgRes, spp=spp, seed=random.randint(0, 2147483647))
    I.sum().backward()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
c = scene.counters()
rays = c["closest_rays"] + c["any_rays"]
print(f"{args.iters} iterations in {dt:.2f} s: {args.iters / dt:.1f} it/s, {rays / dt / 1e6:.0f} Mrays/s "
      f"(gradient finite: {bool(torch.isfinite(material.grad).all())}, non-zero texels: {int((material.grad.abs().sum(-1) > 0).sum())})")
