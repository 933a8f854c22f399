"""zdr_b200 — B200-native differentiable ray-tracer core with zdr's Python surface.

``from zdr_b200 import Scene, Camera, float3`` mirrors ``from zdr import Scene, Camera, float3``
(reference ``__init__.py:1``).
"""
from .mathtypes import Camera, float2, float3, float4, float4x4, int2  # noqa: F401
from .load_obj import read_obj, concat_triangles, load_mesh  # noqa: F401


def __getattr__(name):
    # Scene pulls in torch and the CUDA library; keep `import zdr_b200` light for host-only users
    if name == "Scene":
        from .render import Scene
        return Scene
    raise AttributeError(name)


__all__ = ["Scene", "Camera", "float2", "float3", "float4", "float4x4", "int2", "read_obj", "concat_triangles", "load_mesh"]
