"""Inverse rendering with zdr_b200, the usage contract of the reference's example.py (example.py:20-32, 56, 74-91):
render a ground-truth image of the Cornell box from its textures, then recover a random texture by Adam on the image
loss through `Scene.render(...)` / `.backward()`.  Only the import line differs from the reference.

With `--textures png` the ground truth is the reference's own cboxd.png / cboxr.png loaded as example.py:13-18 does
(`ToTensor() ** 2.2`, 1024^2 texels); with `--out DIR` the script also writes what example.py writes: gt.png, the
screen-to-texture Jacobian images of `render_duvdxy` (duvdx_dudy.png, footprints.png; example.py:59-66), the reconstruction
a.png and the recovered textures (diffuse.npy / roughness.npy in place of the .exr files: imageio is optional here).

    python examples/inverse_cornell.py [--res 512] [--spp 64] [--iters 200] [--integrator path] [--textures png] [--out results]
"""
import argparse
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from zdr_b200 import Camera, Scene, float3          # reference: from zdr import Scene, float3, Camera

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--tex", type=int, default=256)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--gt-spp", type=int, default=1024)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--integrator", default="path", choices=["path", "direct", "collocated"])
ap.add_argument("--textures", default="procedural", choices=["procedural", "png"])
ap.add_argument("--out", default=None)
args = ap.parse_args()

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zdr_b200", "assets")
cbox_model = [(os.path.join(ASSETS, "cboxuv.obj"), None, float3(0.0)),
              (os.path.join(ASSETS, "cbox-light.obj"), None, float3(17, 12, 4))]
cbox_camera1 = Camera(fov=50 / 180 * 3.1415926, origin=float3(-0.2, 2.6, 6.0), target=float3(-0.2, 2.6, -2.5), up=float3(0.0, 1.0, 0.0))

scene = Scene(cbox_model, integrator=args.integrator)
scene.camera = cbox_camera1
scene.max_depth = 8

torch.manual_seed(0)
# ground truth: a smooth procedural texture (the reference loads cboxd.png / cboxr.png with torchvision)
u = torch.linspace(0, 1, args.tex, device="cuda")
vv, uu = torch.meshgrid(u, u, indexing="ij")
material_GT = torch.stack([0.5 + 0.4 * torch.sin(9 * uu), 0.5 + 0.4 * torch.cos(7 * vv), 0.5 + 0.4 * torch.sin(5 * (uu + vv)),
                           0.45 + 0.25 * torch.cos(6 * uu * vv)], dim=-1).clamp(1e-3, 1).contiguous()
if args.textures == "png":
    from common import load_material, save_png
    material_GT = load_material(os.path.join(ASSETS, "cboxd.png"), os.path.join(ASSETS, "cboxr.png"))
    args.tex = material_GT.shape[0]
ImgRes = args.res, args.res
I_GT = scene.render(material_GT, res=ImgRes, spp=args.gt_spp)
if args.out:
    from common import save_png
    save_png(os.path.join(args.out, "gt.png"), I_GT)
    # duv/dxy (screen space to texture space jacobian), example.py:59-66
    duvdxy = scene.render_duvdxy(material_GT, res=ImgRes, spp=128)
    save_png(os.path.join(args.out, "duvdx_dudy.png"), duvdxy[..., 0:3] * 1000 + 0.5)
    footprints = sum(torch.det(scene.render_duvdxy(material_GT, res=ImgRes, spp=1, seed=f).reshape(ImgRes[1], ImgRes[0], 2, 2)).abs()
                     * args.tex * args.tex for f in range(16)) / 16
    save_png(os.path.join(args.out, "footprints.png"), footprints)

material = torch.rand((args.tex, args.tex, 4), device="cuda").requires_grad_()
optimizer = torch.optim.Adam([material], lr=0.01)
with torch.no_grad():
    mse0 = ((scene.render(material, res=ImgRes, spp=args.gt_spp) - I_GT) ** 2).mean().item()
t0 = time.perf_counter()
for it in range(args.iters):
    I = scene.render(material, res=ImgRes, spp=args.spp, seed=random.randint(0, 2147483647))
    ((I - I_GT) ** 2).sum().backward()
    optimizer.step()
    optimizer.zero_grad()
    material.data.clamp_(min=1e-3, max=1)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
with torch.no_grad():
    I = scene.render(material, res=ImgRes, spp=args.gt_spp)
mse = ((I - I_GT) ** 2).mean().item()
if args.out:
    import numpy as np
    save_png(os.path.join(args.out, "a.png"), I)
    np.save(os.path.join(args.out, "diffuse.npy"), material[..., 0:3].detach().cpu().numpy())
    np.save(os.path.join(args.out, "roughness.npy"), material[..., 3].detach().cpu().numpy())
print(f"{args.integrator} {ImgRes[0]}x{ImgRes[1]} {args.spp} spp: {args.iters} Adam steps in {dt:.2f} s ({args.iters / dt:.1f} it/s); "
      f"image MSE {mse0:.5f} -> {mse:.5f}")
