"""Helpers the reference's scripts define inline (example.py:13-18, test_lightstage.py:27-44), shared by the examples here."""
import os
import sys
from math import cos, sin

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ASSETS = os.path.join(ROOT, "zdr_b200", "assets")


def asset(name):
    return os.path.join(ASSETS, name)


def load_material(diffuse_file, roughness_file):
    """example.py:13-18 — `ToTensor()(Image.open(f))` (uint8 / 255, channels first), diffuse RGB over one roughness channel,
    `** 2.2`, (Ht, Wt, 4) contiguous on the GPU.  PIL + torch only (no torchvision needed)."""
    import torch
    from PIL import Image

    def to_tensor(f):
        a = np.asarray(Image.open(f))
        if a.ndim == 2:
            a = a[..., None]
        scale = 255.0 if a.dtype == np.uint8 else 65535.0
        return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).cuda().float() / scale
    diffuse_img = to_tensor(diffuse_file)[0:3]
    roughness_img = to_tensor(roughness_file)[0:1]
    return (torch.vstack((diffuse_img, roughness_img)).permute((1, 2, 0)) ** 2.2).contiguous()


def procedural_envmap(w=512, h=256):
    """Stand-in for assets/empty_workshop_4k.exr (not in the reference's checkout): sky gradient, a warm sun, a dim floor."""
    v, u = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
    sky = np.stack([0.35 + 0.25 * (1 - v), 0.45 + 0.3 * (1 - v), 0.6 + 0.4 * (1 - v)], -1) * (v < 0.5)[..., None]
    floor = np.stack([0.12 + 0 * v, 0.1 + 0 * v, 0.08 + 0 * v], -1) * (v >= 0.5)[..., None]
    sun = 60.0 * np.exp(-(((u - 0.3) * 2) ** 2 + (v - 0.25) ** 2) / 0.002)[..., None] * np.array([1.0, 0.9, 0.7])
    return (sky + floor + sun).astype(np.float32)


def rotate_mat(theta, phi, offset):
    """test_lightstage.py:27-44, returning this package's float4x4 (column-major, as luisa.float4x4(*m.transpose().flatten()))."""
    from zdr_b200 import float4x4
    pitch = np.array([[cos(theta), -sin(theta), 0, 0], [sin(theta), cos(theta), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    yaw = np.array([[cos(phi), 0, -sin(phi), 0], [0, 1, 0, 0], [sin(phi), 0, cos(phi), 0], [0, 0, 0, 1]])
    translate = np.array([[1, 0, 0, offset[0]], [0, 1, 0, offset[1]], [0, 0, 1, offset[2]], [0, 0, 0, 1]])
    m = yaw @ pitch @ translate
    return float4x4(*m.transpose().flatten())


def save_png(path, img):
    """`Image.fromarray((I[..., 0:3].clamp(0, 1) ** 0.454 * 255).to(torch.uint8)...)` of example.py:56."""
    import torch
    from PIL import Image
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    x = img if img.ndim == 2 else img[..., 0:3]
    Image.fromarray((x.detach().clamp(min=0, max=1) ** 0.454 * 255).to(torch.uint8).cpu().numpy()).save(path)
