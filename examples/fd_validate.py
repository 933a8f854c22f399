"""The reference's fd_validate.py on zdr_b200: finite differences against the autograd (PRB adjoint) gradient of ONE pixel with
respect to ONE texel, for spp = 2^0 .. 2^max_exp and the script's five seeds (fd_validate.py:57-114); the (pixel, texel) pair is
importance-picked as at :116-186.  "good if values in last row of AD and FD are similar" (:114).

Scene: the Cornell box the script keeps commented in (:21-33, 41-42) with cboxd.png / cboxr.png; `--scene sphere` is its default
(sphere.obj + envmap; a procedural envmap stands in for assets/empty_workshop_4k.exr, which its checkout lacks).
tests/test_gpu_fd.py asserts what this script prints.

    python examples/fd_validate.py [--scene cbox|sphere] [--integrator path|direct] [--res 256] [--max-exp 12] [--prb-mode corrected|literal]
"""
import argparse
import random

from common import asset, load_material, procedural_envmap

import torch

from zdr_b200 import Camera, Scene, float3          # reference: from zdr import Scene, Camera, float3

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cbox", choices=["cbox", "sphere"])
ap.add_argument("--integrator", default="path", choices=["path", "direct", "collocated"])
ap.add_argument("--res", type=int, default=256)              # fd_validate.py renders 1024 x 1024; the pair is one pixel
ap.add_argument("--tex", type=int, default=32, help="texture down-sampled to tex x tex so that one texel carries a visible gradient")
ap.add_argument("--max-exp", type=int, default=12)
ap.add_argument("--prb-mode", default="corrected", choices=["corrected", "literal"])
ap.add_argument("--seed", type=int, default=None)
args = ap.parse_args()
if args.seed is not None:
    random.seed(args.seed)
    torch.manual_seed(args.seed)

cbox_model = [(asset("cboxuv.obj"), None, float3(0.0)), (asset("cbox-light.obj"), None, float3(20.0))]
sphere_model = [(asset("sphere.obj"), None, float3(0.0))]
cbox_camera1 = Camera(fov=50 / 180 * 3.1415926, origin=float3(-0.2, 2.6, 6.0), target=float3(-0.2, 2.6, -2.5), up=float3(0.0, 1.0, 0.0))
sphere_camera1 = Camera(fov=50 / 180 * 3.1415926, origin=float3(3.0, 0.0, 0.0), target=float3(0.0, 0.0, 0.0), up=float3(0.0, 1.0, 0.0))

if args.scene == "cbox":
    scene = Scene(cbox_model, integrator=args.integrator)
    scene.camera = cbox_camera1
else:
    scene = Scene(sphere_model, integrator=args.integrator)
    scene.add_envmap(procedural_envmap())
    scene.camera = sphere_camera1
scene.prb_mode = args.prb_mode
RES = (args.res, args.res)

material_GT = load_material(asset("cboxd.png"), asset("cboxr.png"))
material_GT = torch.nn.functional.interpolate(material_GT.permute(2, 0, 1)[None], size=(args.tex, args.tex), mode="area")[0].permute(1, 2, 0)
material_GT = (0.02 + 0.96 * material_GT).contiguous()          # inside (fd_eps, 1 - fd_eps), fd_validate.py:93-94
material_black = torch.zeros_like(material_GT)
material_black[..., 3] = 1.0
print("scene loaded")


def compare_fdad(scene, material, texidx, imgidx, res):
    """fd_validate.py:57-114."""
    def fd_grad(scene, material, texidx, imgidx, res, spp, seed, fd_eps):
        m0 = material.detach().clone()
        m0[texidx] -= fd_eps
        I0 = scene.render(m0, res=res, spp=spp, seed=seed)
        m1 = material.detach().clone()
        m1[texidx] += fd_eps
        I1 = scene.render(m1, res=res, spp=spp, seed=seed)
        return (I1[imgidx] - I0[imgidx]).item() / (2 * fd_eps)      # the same seed on both sides

    def ad_grad(scene, material, texidx, imgidx, res, spp, seed):
        material.requires_grad_()
        material.grad = None
        I = scene.render(material, res=res, spp=spp, seed=seed)
        I[imgidx].backward()
        return material.grad[texidx].item()

    fd_eps = 0.01
    if material[texidx] < fd_eps or material[texidx] > 1 - fd_eps:
        raise RuntimeError("material too close to boundary, can not FD")
    seeds = [0, 12345, 853402567, 19260817, 948377263]
    rows = {}
    print("Increasing spps, each 5 different seeds")
    for name, fn in (("FD", lambda spp, seed: fd_grad(scene, material, texidx, imgidx, res, spp, seed, fd_eps)),
                     ("AD", lambda spp, seed: ad_grad(scene, material, texidx, imgidx, res, spp, seed))):
        print(name + ":")
        for exp in range(args.max_exp + 1):
            rows[name] = [fn(2 ** exp, seed) for seed in seeds]
            print(" ".join("%0.6f" % v for v in rows[name]))
    print("good if values in last row of AD and FD are similar")
    def mean_se(v):
        m = sum(v) / len(v)
        return m, (sum((x - m) ** 2 for x in v) / (len(v) - 1) / len(v)) ** 0.5
    (fd, fd_se), (ad, ad_se) = mean_se(rows["FD"]), mean_se(rows["AD"])
    print(f"last rows, mean over seeds +- standard error: FD {fd:.6f} +- {fd_se:.6f}   AD {ad:.6f} +- {ad_se:.6f}   "
          f"difference {abs(fd - ad) / max((fd_se ** 2 + ad_se ** 2) ** 0.5, 1e-30):.1f} standard errors")


def importance_sample_tensor(values):
    flat = values.flatten().double().clamp_min(0)
    return tuple(int(x) for x in torch.unravel_index(torch.multinomial(flat / flat.sum(), 1)[0], values.shape))


try_roughness = random.random() < 0.5              # mandate differentiating roughness with a probability (fd_validate.py:128)
print("Selecting pair: ", end="")
while True:
    material_GT.requires_grad_()
    material_GT.grad = None
    I_GT = scene.render(material_GT, res=RES, spp=128)           # seed defaults to 0
    I_black = scene.render(material_black, res=RES, spp=128)
    print("#", end="")
    img_weight = I_GT.detach().clone()
    light_mask = (I_GT.detach() == I_black).all(axis=-1)         # pixels on the light: material derivatives are zero
    img_weight[light_mask] = 0.0
    imgidx = importance_sample_tensor(img_weight[..., 0:3])
    I_GT[imgidx].backward()
    g = material_GT.grad
    if (g.min().item() == 0.0 and g.max().item() == 0.0) or g.isnan().any().item():
        print("BAD! imgidx", imgidx, "pixel value", I_GT[imgidx].item())
        raise SystemExit(1)
    texidx = importance_sample_tensor(g.abs())
    if texidx[-1] == 3 or not try_roughness:
        print()
        break

print("Image index:", imgidx)
print("Pixel brightness:", I_GT[imgidx].item())
print("Texture index:", texidx)
print("Texel gradient:", material_GT.grad[texidx].item())
print()
compare_fdad(scene, material_GT.detach().clone(), texidx, imgidx, res=RES)
