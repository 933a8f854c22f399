"""ctypes front-end of the CPU oracle (oracle/zdr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs import this module; the product package ``zdr_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from zdr_b200._abi import ZdrParams
from zdr_b200.load_obj import load_mesh_python as load_mesh   # the line-by-line Python statement, not the product library
from zdr_b200.mathtypes import float3, float4x4
from zdr_b200.tables import get_tables

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libzdr_oracle.so")
_lib = None

_f32p, _f64p, _i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
_vp, _pp = C.c_void_p, C.POINTER(ZdrParams)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zdr_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.zo_scene_create.restype = _vp
        L.zo_scene_destroy.argtypes = [_vp]
        L.zo_scene_add_mesh.argtypes = [_vp, _f32p, C.c_int, _i32p, C.c_int, _f32p, _f32p]
        L.zo_scene_commit.argtypes = [_vp]
        L.zo_scene_set_emissions.argtypes = [_vp, _f32p, C.c_int]
        L.zo_scene_set_material_sets.argtypes = [_vp, _i32p, C.c_int]
        L.zo_scene_set_envmap.argtypes = [_vp, _f32p, C.c_int, C.c_int, C.c_int]
        L.zo_scene_get_env_tables.argtypes = [_vp, _f32p, _i32p, _f32p]
        L.zo_scene_get_env_scale_map.argtypes = [_vp, _f32p]
        L.zo_probe_sample_envmap.argtypes = [_vp, _f32p, C.c_int, _f32p]
        L.zo_probe_env_pdf.argtypes = [_vp, _f32p, C.c_int, _f32p]
        L.zo_probe_env_mapping.argtypes = [_f32p, C.c_int, _f32p]
        L.zo_probe_env_mapping.restype = None
        L.zo_scene_light_count.argtypes = [_vp]
        L.zo_scene_instance_count.argtypes = [_vp]
        L.zo_set_sampler_tables.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_int, C.c_int]
        L.zo_render_forward.argtypes = [_vp, C.c_int, _pp, _f32p, _f32p]
        L.zo_render_samples.argtypes = [_vp, C.c_int, _pp, _f32p, _f32p]
        L.zo_render_backward.argtypes = [_vp, C.c_int, _pp, _f32p, _f32p, _f32p]
        L.zo_render_duvdxy.argtypes = [_vp, _pp, _f32p]
        L.zo_probe_duvdxy.argtypes = [_vp, _f32p, C.c_int, _f32p]
        L.zo_get_ray_counters.argtypes = [_vp, C.POINTER(C.c_uint64)]
        L.zo_reset_ray_counters.argtypes = [_vp]
        L.zo_set_num_threads.argtypes = [C.c_int]
        L.zo_set_num_threads.restype = None
        L.zo_xxhash32_int4.restype = C.c_uint32
        L.zo_xxhash32_int4.argtypes = [C.c_int] * 4
        L.zo_permutation_element.restype = C.c_int
        L.zo_permutation_element.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.zo_probe_sampler.argtypes = [_i32p, C.c_int, C.c_char_p, _f32p]
        L.zo_probe_trace_closest.argtypes = [_vp, _f32p, C.c_int, _i32p, _f32p]
        L.zo_probe_trace_any.argtypes = [_vp, _f32p, C.c_int, _i32p]
        L.zo_probe_surface_interact.argtypes = [_vp, _i32p, _f32p, C.c_int, _f32p]
        L.zo_probe_generate_ray.argtypes = [_pp, _f32p, C.c_int, _f32p]
        L.zo_probe_read_bsdf.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _f32p]
        L.zo_create_alias_table.argtypes = [_f64p, C.c_int, _f32p, _i32p, _f64p]
        for name in ("zo_probe_ggx", "zo_probe_ggx_sample_pdf", "zo_probe_sample_wm", "zo_probe_cosine_hemisphere",
                     "zo_probe_make_onb", "zo_probe_tent_warp", "zo_probe_uniform_triangle", "zo_probe_sincos",
                     "zo_probe_offset_ray_origin"):
            getattr(L, name).argtypes = [_f32p, C.c_int, _f32p]
            getattr(L, name).restype = None
        _lib = L
        set_tables(*get_tables()[:2])
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def set_tables(pmj: np.ndarray, bn: np.ndarray):
    L = _lib if _lib is not None else lib()
    pmj = np.ascontiguousarray(pmj, dtype=np.uint32)
    bn = np.ascontiguousarray(bn, dtype=np.uint16)
    rc = L.zo_set_sampler_tables(_p(pmj, C.c_uint32), pmj.shape[0], pmj.shape[1], _p(bn, C.c_uint16), bn.shape[0], bn.shape[1])
    assert rc == 0


def _rowfn(name, inp, out_cols):
    inp = _f(inp)
    n = inp.shape[0]
    out = np.empty((n, out_cols) if out_cols > 1 else (n,), dtype=np.float32)
    getattr(lib(), name)(_p(inp), n, _p(out))
    return out


# ---- function-level probes -------------------------------------------------------------
def xxhash32_int4(x, y, z, w) -> int:
    return int(lib().zo_xxhash32_int4(int(x), int(y), int(z), int(w)))


def permutation_element(i, l, w, p) -> int:
    return int(lib().zo_permutation_element(int(i), int(l), int(w), int(p) & 0xFFFFFFFF))


def sampler_draws(states, prog: str) -> np.ndarray:
    """states: (n,5) int32 rows (px, py, seed, spp, sample_index); prog e.g. '2121'."""
    st = np.ascontiguousarray(states, dtype=np.int32).reshape(-1, 5)
    per = sum(2 if c == "2" else 1 for c in prog)
    out = np.empty((st.shape[0], per), dtype=np.float32)
    rc = lib().zo_probe_sampler(_p(st, C.c_int32), st.shape[0], prog.encode(), _p(out))
    assert rc == 0
    return out


def ggx(inp):
    return _rowfn("zo_probe_ggx", inp, 7)


def ggx_sample_pdf(inp):
    return _rowfn("zo_probe_ggx_sample_pdf", inp, 1)


def sample_wm(inp):
    return _rowfn("zo_probe_sample_wm", inp, 3)


def cosine_hemisphere(inp):
    return _rowfn("zo_probe_cosine_hemisphere", inp, 3)


def make_onb(inp):
    return _rowfn("zo_probe_make_onb", inp, 9)


def tent_warp(inp):
    return _rowfn("zo_probe_tent_warp", np.asarray(inp, dtype=np.float32).reshape(-1, 1), 1)


def uniform_triangle(inp):
    return _rowfn("zo_probe_uniform_triangle", inp, 3)


def sincos(inp):
    return _rowfn("zo_probe_sincos", np.asarray(inp, dtype=np.float32).reshape(-1, 1), 2)


def offset_ray_origin(inp):
    return _rowfn("zo_probe_offset_ray_origin", inp, 3)


def generate_ray(params: ZdrParams, ndc):
    ndc = _f(ndc)
    out = np.empty((ndc.shape[0], 6), dtype=np.float32)
    lib().zo_probe_generate_ray(C.byref(params), _p(ndc), ndc.shape[0], _p(out))
    return out


def read_bsdf(material, uv, bilinear=True):
    m, uv = _f(material), _f(uv)
    out = np.empty((uv.shape[0], 4), dtype=np.float32)
    lib().zo_probe_read_bsdf(_p(m), m.shape[0], m.shape[1], int(bilinear), _p(uv), uv.shape[0], _p(out))
    return out


def env_mapping(inp):
    """(n, 5) uv.xy dir.xyz -> (n, 5) uv_to_direction(uv).xyz direction_to_uv(dir).xy (envmap.py:206-219)."""
    a = _f(inp).reshape(-1, 5)
    out = np.empty_like(a)
    lib().zo_probe_env_mapping(_p(a), a.shape[0], _p(out))
    return out


def create_alias_table(values):
    v = np.ascontiguousarray(values, dtype=np.float64)
    n = v.shape[0]
    prob, alias, pdf = np.empty(n, np.float32), np.empty(n, np.int32), np.empty(n, np.float64)
    lib().zo_create_alias_table(_p(v, C.c_double), n, _p(prob), _p(alias, C.c_int32), _p(pdf, C.c_double))
    return prob, alias, pdf


# ---- scene -----------------------------------------------------------------------------
def _emission3(e):
    if e is None:
        return [0.0, 0.0, 0.0]
    if isinstance(e, (int, float)):
        return [float(e)] * 3
    return [float(c) for c in e]


class OracleScene:
    """Same ``models`` contract as ``zdr_b200.Scene`` (render.py:55,81-91)."""

    def __init__(self, models):
        L = lib()
        self._h = L.zo_scene_create()
        self._mesh_cache = {}
        for obj_file, transform, emission in models:
            if isinstance(obj_file, str):
                if obj_file not in self._mesh_cache:
                    self._mesh_cache[obj_file] = load_mesh(obj_file)
                verts, tris = self._mesh_cache[obj_file]
            else:
                verts, tris = obj_file                     # (verts Nx8, tris) given directly
            verts, tris = _f(verts), np.ascontiguousarray(tris, dtype=np.int32).reshape(-1)
            xf = _f((transform if transform is not None else float4x4(1.0)).flat())
            em = _f(_emission3(emission))
            rc = L.zo_scene_add_mesh(self._h, _p(verts), verts.shape[0], _p(tris, C.c_int32), tris.shape[0] // 3, _p(xf), _p(em))
            assert rc == 0, rc
        assert L.zo_scene_commit(self._h) == 0

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.zo_scene_destroy(self._h)
            self._h = None

    @property
    def light_count(self):
        return lib().zo_scene_light_count(self._h)

    def update_lights(self, emissions):
        e = _f([_emission3(x) for x in emissions])
        assert lib().zo_scene_set_emissions(self._h, _p(e), e.shape[0]) == 0

    def set_material_sets(self, sets):
        a = np.ascontiguousarray(sets, dtype=np.int32)
        assert lib().zo_scene_set_material_sets(self._h, a.ctypes.data_as(_i32p), a.shape[0]) == 0

    def set_envmap(self, rgba, compensate_mis=True):
        img = _f(rgba)
        assert img.ndim == 3 and img.shape[2] == 4
        assert lib().zo_scene_set_envmap(self._h, _p(img), img.shape[1], img.shape[0], int(compensate_mis)) == 0

    def env_tables(self):
        n = 256 + 512 * 256
        prob, alias, pdf = np.empty(n, np.float32), np.empty(n, np.int32), np.empty(512 * 256, np.float32)
        assert lib().zo_scene_get_env_tables(self._h, _p(prob), _p(alias, C.c_int32), _p(pdf)) == 0
        return prob, alias, pdf

    def env_scale_map(self):
        """(256, 512) weight map before MIS compensation (envmap.py:136-160)."""
        out = np.empty((256, 512), np.float32)
        assert lib().zo_scene_get_env_scale_map(self._h, _p(out)) == 0
        return out

    def sample_envmap(self, u):
        """envmap.py:223-238 on (n, 2) uniform numbers -> (n, 8): wi.xyz pdf eval.xyz dist."""
        u = _f(u).reshape(-1, 2)
        out = np.empty((u.shape[0], 8), np.float32)
        assert lib().zo_probe_sample_envmap(self._h, _p(u), u.shape[0], _p(out)) == 0
        return out

    def env_pdf(self, dirs):
        """envmap.py:240-248."""
        d = _f(dirs).reshape(-1, 3)
        out = np.empty(d.shape[0], np.float32)
        assert lib().zo_probe_env_pdf(self._h, _p(d), d.shape[0], _p(out)) == 0
        return out

    def render_forward(self, integrator: int, params: ZdrParams, material) -> np.ndarray:
        m = _f(material)
        img = np.zeros((params.height, params.width, 4), dtype=np.float32)
        rc = lib().zo_render_forward(self._h, integrator, C.byref(params), _p(m), _p(img))
        assert rc == 0, rc
        return img

    def render_samples(self, integrator: int, params: ZdrParams, material) -> np.ndarray:
        m = _f(material)
        out = np.zeros((params.height, params.width, params.spp, 3), dtype=np.float32)
        rc = lib().zo_render_samples(self._h, integrator, C.byref(params), _p(m), _p(out))
        assert rc == 0, rc
        return out

    def render_backward(self, integrator: int, params: ZdrParams, material, d_image) -> np.ndarray:
        m, g = _f(material), _f(d_image)
        dm = np.zeros_like(m)
        rc = lib().zo_render_backward(self._h, integrator, C.byref(params), _p(m), _p(g), _p(dm))
        assert rc == 0, rc
        return dm

    def render_duvdxy(self, params: ZdrParams) -> np.ndarray:
        img = np.zeros((params.height, params.width, 4), dtype=np.float32)
        rc = lib().zo_render_duvdxy(self._h, C.byref(params), _p(img))
        assert rc == 0, rc
        return img

    def duvdxy(self, rays18):
        r = _f(rays18).reshape(-1, 18)
        out = np.empty((r.shape[0], 4), np.float32)
        assert lib().zo_probe_duvdxy(self._h, _p(r), r.shape[0], _p(out)) == 0
        return out

    def ray_counters(self):
        out = (C.c_uint64 * 2)()
        lib().zo_get_ray_counters(self._h, out)
        return int(out[0]), int(out[1])

    def reset_ray_counters(self):
        lib().zo_reset_ray_counters(self._h)

    def trace_closest(self, rays):
        r = _f(rays).reshape(-1, 8)
        ip = np.empty((r.shape[0], 2), np.int32)
        uvt = np.empty((r.shape[0], 3), np.float32)
        assert lib().zo_probe_trace_closest(self._h, _p(r), r.shape[0], _p(ip, C.c_int32), _p(uvt)) == 0
        return ip, uvt

    def trace_any(self, rays):
        r = _f(rays).reshape(-1, 8)
        occ = np.empty(r.shape[0], np.int32)
        assert lib().zo_probe_trace_any(self._h, _p(r), r.shape[0], _p(occ, C.c_int32)) == 0
        return occ

    def surface_interact(self, inst_prim, uv):
        ip = np.ascontiguousarray(inst_prim, dtype=np.int32).reshape(-1, 2)
        uv = _f(uv).reshape(-1, 2)
        out = np.empty((ip.shape[0], 11), np.float32)
        assert lib().zo_probe_surface_interact(self._h, _p(ip, C.c_int32), _p(uv), ip.shape[0], _p(out)) == 0
        return out
