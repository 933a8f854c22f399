"""Wavefront OBJ ingest with the reference's indexing semantics.

Behaviour follows ``load_obj.py:1-68`` of the reference, because it fixes what the hot path
observes: instance-local vertex ids are assigned in first-use order of the distinct
``(position, uv, normal)`` triples (``load_obj.py:47-55``), primitive ids follow face order with
fan triangulation ``(f0, f[i-1], f[i])`` (``load_obj.py:63-68``), a missing ``vt`` becomes
``(0, 0)`` (``:48``) and a missing ``vn`` becomes NaN (``:49``) which triggers
``recompute_normal`` at load (``render.py:101-103``).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np


def _parse_corner(token: str, pos, tex, nrm):
    fields = token.split("/")
    p = pos[int(fields[0]) - 1]
    t = tex[int(fields[1]) - 1] if len(fields) > 1 and fields[1] else (0.0, 0.0)
    # a FRESH NaN object per corner (load_obj.py:49): distinct NaN objects never compare equal,
    # so corners without a normal are never merged and such meshes end up flat-shaded.
    n = nrm[int(fields[2]) - 1] if len(fields) > 2 and fields[2] else (float("nan"),) * 3
    return p, t, n


def read_obj(file_path: str) -> Tuple[List[tuple], List[List[int]]]:
    """Returns ``(vertices, faces)``: ``vertices[i] = ((px,py,pz), (u,v), (nx,ny,nz))`` and
    ``faces[j]`` = list of vertex ids of polygon j (same return contract as the reference)."""
    pos, tex, nrm = [], [], []
    vertices: List[tuple] = []
    index_of = {}
    faces: List[List[int]] = []
    with open(file_path, "r") as fh:
        for raw in fh:
            tok = raw.split()
            if not tok:
                continue
            kind = tok[0]
            if kind == "v":
                pos.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif kind == "vt":
                tex.append((float(tok[1]), float(tok[2])))
            elif kind == "vn":
                nrm.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif kind == "f":
                poly = []
                for corner in tok[1:]:
                    key = _parse_corner(corner, pos, tex, nrm)
                    vid = index_of.get(key)
                    if vid is None:
                        vid = len(vertices)
                        index_of[key] = vid
                        vertices.append(key)
                    poly.append(vid)
                faces.append(poly)
    return vertices, faces


def concat_triangles(faces: List[List[int]]) -> List[int]:
    """Fan triangulation, flattened (``load_obj.py:63-68``)."""
    out: List[int] = []
    for poly in faces:
        a = poly[0]
        for k in range(2, len(poly)):
            out.extend((a, poly[k - 1], poly[k]))
    return out


def recompute_normal(verts: np.ndarray, tris: np.ndarray) -> np.ndarray:
    """Area-weighted vertex normals (``recompute_normal.py:4-51``): zero, scatter-add the
    un-normalised face normal ``cross(p1-p0, p2-p0)`` to the three corners, normalise.
    Load-time only (out of scope as a perf target); float32, triangle order."""
    v = verts.copy()
    p = v[:, 0:3]
    t = tris.reshape(-1, 3)
    e1 = p[t[:, 1]] - p[t[:, 0]]
    e2 = p[t[:, 2]] - p[t[:, 0]]
    fn = np.cross(e1, e2).astype(np.float32)
    acc = np.zeros((v.shape[0], 3), dtype=np.float32)
    for k in range(3):
        np.add.at(acc, t[:, k], fn)
    with np.errstate(invalid="ignore", divide="ignore"):
        acc /= np.sqrt((acc * acc).sum(axis=1, keepdims=True, dtype=np.float32))
    v[:, 5:8] = acc
    return v


def load_mesh_native(file_path: str) -> Tuple[np.ndarray, np.ndarray]:
    """The same result as :func:`load_mesh_python` from the library's C++ parser (``zdr_obj_load``,
    ``csrc/obj_loader.cu``): seconds instead of minutes on meshes of millions of triangles."""
    import ctypes as C

    from . import _abi
    lib = _abi.load_library()
    vp, tp = C.POINTER(C.c_float)(), C.POINTER(C.c_int32)()
    nv, nt = C.c_int(), C.c_int()
    rc = lib.zdr_obj_load(file_path.encode(), C.byref(vp), C.byref(nv), C.byref(tp), C.byref(nt))
    if rc != 0:
        raise ValueError(f"{file_path}: {lib.zdr_obj_last_error().decode()}")
    try:
        verts = np.ctypeslib.as_array(vp, shape=(nv.value, 8)).copy()
        tris = np.ctypeslib.as_array(tp, shape=(nt.value * 3,)).copy()
    finally:
        lib.zdr_obj_free(vp)
        lib.zdr_obj_free(tp)
    return verts, tris


def load_mesh(file_path: str) -> Tuple[np.ndarray, np.ndarray]:
    """OBJ -> (``verts`` float32 (N,8) rows ``px py pz u v nx ny nz`` = the 32-byte ``Vertex`` of
    ``vertex.py:4`` as filled at ``render.py:97``; ``tris`` int32 (T*3,)).  Normals are
    recomputed when the file has none (``render.py:101-103``).  Parsed by the library's native loader;
    :func:`load_mesh_python` is the line-by-line statement of the same algorithm (tests hold them equal)."""
    return load_mesh_native(file_path)


def load_mesh_python(file_path: str) -> Tuple[np.ndarray, np.ndarray]:
    """The line-by-line statement of the ingest (``read_obj`` + ``concat_triangles`` + ``recompute_normal``): what the
    reference does at ``render.py:93-103``.  ``load_mesh_native`` must return the same bytes."""
    vertices, faces = read_obj(file_path)
    if not vertices or not faces:
        raise ValueError(f"{file_path}: no geometry")
    verts = np.array([(*p, *t, *n) for p, t, n in vertices], dtype=np.float32)
    tris = np.array(concat_triangles(faces), dtype=np.int32)
    if math.isnan(vertices[0][2][0]):
        verts = recompute_normal(verts, tris)
    return verts, tris
