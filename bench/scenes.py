"""Synthetic workloads shared by the tests, bench.py and smoke().

Pure host-side descriptions: model lists, cameras and seeded materials.  Nothing here touches the
oracle or the CUDA library, so both arms of every comparison are built from the same inputs.
"""
from __future__ import annotations

import math
import os

import numpy as np

from zdr_b200.mathtypes import Camera, float3, float4x4

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zdr_b200", "assets")


def asset(name: str) -> str:
    return os.path.join(ASSETS, name)


# example.py:27-32
CBOX_CAMERA = Camera(fov=50 / 180 * 3.1415926, origin=float3(-0.2, 2.6, 6.0), target=float3(-0.2, 2.6, -2.5),
                     up=float3(0.0, 1.0, 0.0))
# example.py:33-38
SPHERE_CAMERA = Camera(fov=50 / 180 * 3.1415926, origin=float3(1.0, 0.0, 0.0), target=float3(0.0, 0.0, 0.0),
                       up=float3(0.0, 1.0, 0.0))


def cbox_collocated_models():            # C1: cboxuv.obj only
    return [(asset("cboxuv.obj"), None, float3(0.0))]


def cbox_direct_models():                # C2: cbox-combined (no vt => uv = 0) + light (17,12,4)
    return [(asset("cbox-combined.obj"), None, float3(0.0)), (asset("cbox-light.obj"), None, float3(17, 12, 4))]


def cbox_path_models():                  # C3: cboxuv + light (example.py:20-23)
    return [(asset("cboxuv.obj"), None, float3(0.0)), (asset("cbox-light.obj"), None, float3(17, 12, 4))]


def random_material(res=(1024, 1024), seed=0, lo=1e-3) -> np.ndarray:
    """example.py:77,86: uniform random texture clamped to [1e-3, 1]; numpy-seeded so that the
    oracle and the CUDA path get identical bytes without torch's device RNG."""
    rng = np.random.default_rng(seed)
    m = rng.random((res[0], res[1], 4), dtype=np.float32)
    return np.clip(m, lo, 1.0).astype(np.float32)


def constant_material(res=(16, 16), value=(0.5, 0.5, 0.5, 0.5)) -> np.ndarray:
    """C4: spatially uniform texture (atomic-contention worst case)."""
    m = np.empty((res[0], res[1], 4), dtype=np.float32)
    m[...] = np.asarray(value, dtype=np.float32)
    return m


def translate_scale(tx, ty, tz, s) -> float4x4:
    return float4x4(s, 0, 0, 0, 0, s, 0, 0, 0, 0, s, 0, tx, ty, tz, 1)


def rotation_y(angle: float) -> float4x4:
    c, s = math.cos(angle), math.sin(angle)
    return float4x4(c, 0, -s, 0, 0, 1, 0, 0, s, 0, c, 0, 0, 0, 0, 1)


def cbox_sheared_models():
    """Cornell box + light under ONE non-rigid transform (rotation about y, non-uniform scale, shift about the box
    centre): exercises p = M v (interaction.py:19-21) and the normal matrix inverse(transpose(M3x3)) (interaction.py:28)."""
    c, s_ = math.cos(0.25), math.sin(0.25)
    sx, sy, sz = 1.1, 0.9, 1.05
    cx, cy, cz = -0.2, 2.6, -2.5                       # rotate / scale about the box centre
    rows = [[c * sx, 0.0, s_ * sz, 0.0], [0.0, sy, 0.0, 0.0], [-s_ * sx, 0.0, c * sz, 0.0], [0.0, 0.0, 0.0, 1.0]]
    lin = np.asarray(rows, np.float64)[:3, :3]
    t = np.asarray([cx, cy, cz]) - lin @ np.asarray([cx, cy, cz]) + np.asarray([0.1, -0.05, 0.0])
    for i in range(3):
        rows[i][3] = float(np.float32(t[i]))
        for j in range(3):
            rows[i][j] = float(np.float32(rows[i][j]))
    m = float4x4.from_rows(rows)
    return [(asset("cboxuv.obj"), m, float3(0.0)), (asset("cbox-light.obj"), m, float3(17, 12, 4))]


def cbox_two_light_models():
    """Cornell box, its ceiling light (on) and a second copy of the light moved down and sideways (off): the scene of the
    update_lights / RenderOperator fixtures."""
    return [(asset("cboxuv.obj"), None, float3(0.0)), (asset("cbox-light.obj"), None, float3(17, 12, 4)),
            (asset("cbox-light.obj"), translate_scale(-1.2, -1.5, 0.4, 1.0), float3(0.0))]


def cbox_material_sets_models():
    """Scene of the material-set ("polymorphic materials") fixtures: Cornell box (texture set 0), its light, sphere.obj inside
    the box (texture set 1) and a second, switched-off copy of the light quad (no texture set: a terminal surface).
    Returns (models, set of every instance)."""
    models = [(asset("cboxuv.obj"), None, float3(0.0)), (asset("cbox-light.obj"), None, float3(17, 12, 4)),
              (asset("sphere.obj"), translate_scale(-0.2, 1.4, -2.4, 1.1), float3(0.0)),
              (asset("cbox-light.obj"), translate_scale(-1.2, -2.5, 0.4, 1.0), float3(0.0))]
    return models, [0, -1, 1, -1]


def sphere_grid_models(n_inst: int, seed: int = 0, emissive_every: int = 0):
    """C5 shape: instances of sphere.obj on a jittered lattice (uniform scale 0.4, spacing 1.0,
    jitter +-0.25), first n cells of a 22x22x21 grid; instance 0 is the textured object."""
    rng = np.random.default_rng(seed)
    models = []
    for i in range(n_inst):
        x, y, z = i % 22, (i // 22) % 22, i // (22 * 22)
        j = rng.uniform(-0.25, 0.25, 3)
        em = float3(8.0, 8.0, 8.0) if (emissive_every and i > 0 and i % emissive_every == 0) else float3(0.0)
        models.append((asset("sphere.obj"), translate_scale(x + j[0], y + j[1], z + j[2], 0.4), em))
    return models


def sphere_grid_camera(n_inst: int) -> Camera:
    ext = min(22.0, max(2.0, n_inst ** (1 / 3) * 1.2))
    return Camera(fov=50 / 180 * 3.1415926, origin=float3(-ext * 0.9, ext * 1.1, -ext * 1.3),
                  target=float3(ext * 0.4, ext * 0.4, ext * 0.3), up=float3(0.0, 1.0, 0.0))


def procedural_envmap(w=256, h=128, seed=0) -> np.ndarray:
    """Seeded sum of 8 gaussians + constant 0.1; RGBA float, 2:1."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
    img = np.full((h, w, 4), 0.1, dtype=np.float64)
    for _ in range(8):
        cx, cy, sg = rng.uniform(0, 1), rng.uniform(0.1, 0.9), rng.uniform(0.02, 0.15)
        col = rng.uniform(0.5, 6.0, 3)
        dx = np.minimum(np.abs(xx - cx), 1 - np.abs(xx - cx))
        g = np.exp(-(dx * dx + (yy - cy) ** 2) / (2 * sg * sg))
        img[..., :3] += g[..., None] * col
    img[..., 3] = 1.0
    return img.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs as workloads (SURVEY.md §8d "Synthetic inputs"); bench.py --config, bench/run.py
# ------------------------------------------------------------------------------------------------
def light_stage_models(n_lights: int = 8):
    """C4 (test_lightstage.py:24-54 shape): sphere.obj as the textured object + n quad lights on a ring, all off;
    the renders switch ONE light on each with update_lights (test_lightstage.py:46-54)."""
    quad = asset("quad.obj")
    lights = [(quad, translate_scale(2.5 * np.cos(a), 2.0, 2.5 * np.sin(a), 0.6), float3(0.0))
              for a in np.linspace(0, 2 * np.pi, n_lights, endpoint=False)]
    return [(asset("sphere.obj"), None, float3(0.0))] + lights


LIGHT_STAGE_CAMERA = Camera(fov=0.7, origin=float3(0.0, 0.6, 3.2), target=float3(0.0, 0.0, 0.0), up=float3(0, 1, 0))


def light_stage_emissions(k: int, n_lights: int = 8):
    return [(0.0, 0.0, 0.0)] + [(12.0, 12.0, 12.0) if j == k else (0.0, 0.0, 0.0) for j in range(n_lights)]


def merged_sphere_grid_models(n_inst: int = 10000, seed: int = 0, emissive_every: int = 97):
    """C5 scene.  The reference textures instance 0 only and ends paths on every other instance (direct.py:30,
    prb.py:45), so 10 000 separate instances would shade almost nothing: the non-emissive spheres are merged into
    ONE textured mesh (instance 0, ~9.5 M triangles), every `emissive_every`-th sphere stays an emissive instance."""
    from zdr_b200.load_obj import load_mesh_python
    grid = sphere_grid_models(n_inst, seed=seed, emissive_every=emissive_every)
    sv, st = load_mesh_python(grid[0][0])
    mv, mt, lights = [], [], []
    for obj, xf, em in grid:
        if any(c > 0 for c in em):
            lights.append((obj, xf, em))
            continue
        m = np.asarray(xf.flat(), np.float32).reshape(4, 4).T                     # column-major -> rows
        v = sv.copy()
        v[:, 0:3] = sv[:, 0:3] @ m[:3, :3].T + m[:3, 3]
        mt.append(st + len(mv) * sv.shape[0])
        mv.append(v)
    return [((np.concatenate(mv), np.concatenate(mt)), None, float3(0.0))] + lights


class Workload:
    """One BASELINE.json config: scene, camera, integrator, sizes, material; `lights` > 0 = light-stage loop (one
    render per light, update_lights between them); `cpu` = the bounded sample the CPU arm renders
    ((width, height), sample range) — rays/s does not depend on spp or on the image size at a fixed view."""

    def __init__(self, key, title, models, camera, integrator, res, spp, depth, material, *, envmap=None, lights=0,
                 cpu_res=None, cpu_its=None, ref_its=None):
        self.key, self.title, self._models, self.camera, self.integrator = key, title, models, camera, integrator
        self.res, self.spp, self.depth, self._material, self._envmap, self.lights = res, spp, depth, material, envmap, lights
        self.cpu_res, self.cpu_its, self.ref_its = cpu_res or res, cpu_its, ref_its

    def models(self):
        return self._models() if callable(self._models) else self._models

    def material(self):
        return self._material()

    def envmap(self):
        return self._envmap() if self._envmap else None

    def name(self):
        return (f"{self.key}: {self.title}, {self.integrator} integrator, max_depth {self.depth} rr_depth 2, {self.res[0]}x{self.res[1]}, "
                f"{self.spp} spp, fwd seed 0 / bwd seed 1, d_image=1")


def workload(key: str, res=None, spp=None, depth=None, tex=None, integrator=None) -> Workload:
    """BASELINE.json configs[0..4] = c1..c5 (c4 = the collocated 1x1-texture light stage, c4d = its direct variant).
    res / spp / depth / tex / integrator override the config's values (reduced-size runs under ncu, tests)."""
    key = key.lower()
    if key == "c1":
        w = Workload("c1", "cornell box cboxuv.obj, random 1024^2 texture", cbox_collocated_models, CBOX_CAMERA, "collocated",
                     (256, 256), 1, 8, lambda: random_material((tex or 1024,) * 2, 0))
    elif key == "c2":
        w = Workload("c2", "cornell box cbox-combined.obj (uv = 0) + cbox-light.obj (17,12,4), random 1024^2 texture",
                     cbox_direct_models, CBOX_CAMERA, "direct", (512, 512), 64, 8, lambda: random_material((tex or 1024,) * 2, 0),
                     ref_its=(0, 16))
    elif key == "c3":
        w = Workload("c3", "cornell box cboxuv.obj + cbox-light.obj (17,12,4), random 1024^2 texture", cbox_path_models, CBOX_CAMERA,
                     "path", (1024, 1024), 256, 8, lambda: random_material((tex or 1024,) * 2, 0), cpu_its=(0, 32), ref_its=(0, 2))
    elif key == "c4":
        w = Workload("c4", "light stage: sphere.obj + 8 quad lights switched on one per render, spatially uniform 1x1 texture",
                     light_stage_models, LIGHT_STAGE_CAMERA, "collocated", (1024, 1024), 1, 8, lambda: constant_material((tex or 1,) * 2),
                     lights=8)
    elif key == "c4d":
        w = Workload("c4d", "light stage: sphere.obj + 8 quad lights switched on one per render, constant 16^2 texture",
                     light_stage_models, LIGHT_STAGE_CAMERA, "direct", (1024, 1024), 4, 8, lambda: constant_material((tex or 16,) * 2),
                     lights=8, ref_its=(0, 1))
    elif key == "c5":
        w = Workload("c5", "9.6 M triangles: 9 896 spheres merged into the textured instance + 104 emissive sphere instances, "
                           "procedural 2048x1024 envmap, random 2048^2 texture", merged_sphere_grid_models, sphere_grid_camera(10000),
                     "path", (3840, 2160), 64, 16, lambda: random_material((tex or 2048,) * 2, 0),
                     envmap=lambda: procedural_envmap(2048, 1024, seed=0), cpu_res=(480, 270), cpu_its=(0, 2), ref_its=(0, 1))
    else:
        raise KeyError(f"unknown config {key!r} (c1, c2, c3, c4, c4d, c5)")
    if integrator:
        w.integrator = integrator
    if res:
        w.res = (int(res[0]), int(res[1]))
        if key != "c5":
            w.cpu_res = w.res
    if spp:
        w.spp = int(spp)
    if depth:
        w.depth = int(depth)
    if w.cpu_its:
        w.cpu_its = (0, min(w.cpu_its[1], w.spp))
    if w.ref_its:
        w.ref_its = (0, min(w.ref_its[1], w.spp))
    return w
