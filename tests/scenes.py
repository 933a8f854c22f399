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
