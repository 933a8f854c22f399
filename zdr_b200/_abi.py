"""ctypes view of ``include/zdr_b200.h`` and loader of the in-tree CUDA library.

There is no CPU fallback: if ``zdr_b200/csrc/libzdr_b200.so`` is missing the import of the
library fails loudly with the build command (``python -c 'import __graft_entry__ as g; g.build()'``).
"""
from __future__ import annotations

import ctypes as C
import os

ZDR_COLLOCATED, ZDR_DIRECT, ZDR_PATH = 0, 1, 2
INTEGRATORS = {"collocated": ZDR_COLLOCATED, "direct": ZDR_DIRECT, "path": ZDR_PATH}  # render.py:65-69
ZDR_PRB_LITERAL, ZDR_PRB_CORRECTED = 0, 1
PRB_MODES = {"literal": ZDR_PRB_LITERAL, "corrected": ZDR_PRB_CORRECTED}
ZDR_NUM_COUNTERS = 8
ZDR_FLAG_STATS = 1
ZDR_FLAG_SCATTER_AGGREGATE, ZDR_FLAG_SCATTER_PLAIN = 2, 4
ZDR_FLAG_WAVEFRONT, ZDR_FLAG_MEGAKERNEL = 8, 16
ZDR_FLAG_LIGHT_PDF_COUNTS_ENV = 32
ZDR_FLAG_PERSISTENT = 64
SCHEDULES = {"auto": 0, "wavefront": ZDR_FLAG_WAVEFRONT, "megakernel": ZDR_FLAG_MEGAKERNEL, "persistent": ZDR_FLAG_PERSISTENT}
ZDR_FLAG_SCATTER_BLOCK = 128
ZDR_FLAG_FAST_MATH = 256
MATH_MODES = {"exact": 0, "fast": ZDR_FLAG_FAST_MATH}
SCATTER_MODES = {"auto": 0, "aggregate": ZDR_FLAG_SCATTER_AGGREGATE, "plain": ZDR_FLAG_SCATTER_PLAIN,
                 "block": ZDR_FLAG_SCATTER_AGGREGATE | ZDR_FLAG_SCATTER_BLOCK}
COUNTER_NAMES = ("closest_rays", "any_rays", "nodes_visited", "tris_tested",
                 "grad_scatters", "red_v4", "shaded_vertices", "camera_samples")


class ZdrParams(C.Structure):
    """``zdr_params`` of include/zdr_b200.h (field order is ABI)."""
    _fields_ = [
        ("cam_fov", C.c_float),
        ("cam_origin", C.c_float * 3),
        ("cam_target", C.c_float * 3),
        ("cam_up", C.c_float * 3),
        ("width", C.c_int32), ("height", C.c_int32),
        ("tex_res_x", C.c_int32), ("tex_res_y", C.c_int32),
        ("spp", C.c_int32), ("seed", C.c_int32),
        ("use_tent_filter", C.c_int32),
        ("max_depth", C.c_int32), ("rr_depth", C.c_int32),
        ("use_mis", C.c_int32), ("bilinear", C.c_int32), ("prb_mode", C.c_int32),
        ("row0", C.c_int32), ("row1", C.c_int32), ("it0", C.c_int32), ("it1", C.c_int32),
        ("shard_index", C.c_int32), ("shard_count", C.c_int32), ("flags", C.c_int32),
        ("num_material_sets", C.c_int32),
    ]


def wrap_i32(x: int) -> int:
    """int32 wrap-around (``seed+1`` at render.py:196 overflows for seed=2**31-1)."""
    return ((int(x) + 2 ** 31) % 2 ** 32) - 2 ** 31


def make_params(camera, res, tex_shape, spp, seed, *, use_tent_filter=True, max_depth=16, rr_depth=2,
                use_mis=True, bilinear=True, prb_mode=ZDR_PRB_CORRECTED, rows=None, its=None, shard=None,
                stats=False, scatter="auto", schedule="auto", light_pdf_counts_env=False, num_material_sets=1, math="exact") -> ZdrParams:
    p = ZdrParams()
    p.cam_fov = float(camera.fov)
    p.cam_origin[:] = list(camera.origin)
    p.cam_target[:] = list(camera.target)
    p.cam_up[:] = list(camera.up)
    p.width, p.height = int(res[0]), int(res[1])                  # res = (W, H), render.py:232
    if len(tex_shape) == 4:                                       # (K, Ht, Wt, 4): K texture sets
        num_material_sets, tex_shape = int(tex_shape[0]), tex_shape[1:]
    p.num_material_sets = int(num_material_sets)
    p.tex_res_x, p.tex_res_y = int(tex_shape[0]), int(tex_shape[1])  # int2(*material.shape[0:2]), render.py:161,170
    p.spp, p.seed = int(spp), wrap_i32(seed)
    p.use_tent_filter = 1 if use_tent_filter else 0
    p.max_depth, p.rr_depth = int(max_depth), int(rr_depth)
    p.use_mis, p.bilinear, p.prb_mode = int(bool(use_mis)), int(bool(bilinear)), int(prb_mode)
    p.row0, p.row1 = (0, p.height) if rows is None else (int(rows[0]), int(rows[1]))
    p.it0, p.it1 = (0, p.spp) if its is None else (int(its[0]), int(its[1]))
    p.shard_index, p.shard_count = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
    p.flags = ((ZDR_FLAG_STATS if stats else 0) | SCATTER_MODES[scatter] | SCHEDULES[schedule]
               | (ZDR_FLAG_LIGHT_PDF_COUNTS_ENV if light_pdf_counts_env else 0) | MATH_MODES[math])
    return p


_HERE = os.path.dirname(os.path.abspath(__file__))
# ZDR_B200_LIB: another build of the same library (kernel-tuning experiments, tools/sweep_*.sh)
LIB_PATH = os.environ.get("ZDR_B200_LIB") or os.path.join(_HERE, "csrc", "libzdr_b200.so")
_lib = None

_f32p, _i32p, _u32p, _u16p, _u64p = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint16), C.POINTER(C.c_uint64))
_vp = C.c_void_p
_pp = C.POINTER(ZdrParams)

# name -> (restype, argtypes): every symbol include/zdr_b200.h declares
SIGNATURES = {
    "zdr_abi_version": (C.c_int, []),
    "zdr_last_error": (C.c_char_p, []),
    "zdr_set_sampler_tables": (C.c_int, [C.c_int, _u32p, C.c_int, C.c_int, _u16p, C.c_int, C.c_int]),
    "zdr_scene_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "zdr_scene_destroy": (None, [_vp]),
    "zdr_scene_add_mesh": (C.c_int, [_vp, _f32p, C.c_int, _i32p, C.c_int, _f32p, _f32p]),
    "zdr_scene_commit": (C.c_int, [_vp]),
    "zdr_scene_clear": (C.c_int, [_vp]),
    "zdr_scene_set_emissions": (C.c_int, [_vp, _f32p, C.c_int]),
    "zdr_scene_set_envmap": (C.c_int, [_vp, _f32p, C.c_int, C.c_int, C.c_int]),
    "zdr_scene_set_material_sets": (C.c_int, [_vp, _i32p, C.c_int]),
    "zdr_scene_set_workspace_limit": (C.c_int, [_vp, C.c_uint64]),
    "zdr_scene_set_workspace": (C.c_int, [_vp, _vp, C.c_uint64]),
    "zdr_scene_trim": (C.c_int, [_vp]),
    "zdr_scene_workspace_bytes": (C.c_uint64, [_vp]),
    "zdr_scene_instance_count": (C.c_int, [_vp]),
    "zdr_scene_light_count": (C.c_int, [_vp]),
    "zdr_scene_triangle_count": (C.c_int, [_vp]),
    "zdr_render_forward": (C.c_int, [_vp, C.c_int, _pp, _vp, _vp, _vp]),
    "zdr_render_backward": (C.c_int, [_vp, C.c_int, _pp, _vp, _vp, _vp, _vp]),
    "zdr_render_duvdxy": (C.c_int, [_vp, _pp, _vp, _vp]),
    "zdr_render_forward_host": (C.c_int, [_vp, C.c_int, _pp, _f32p, _f32p]),
    "zdr_render_backward_host": (C.c_int, [_vp, C.c_int, _pp, _f32p, _f32p, _f32p]),
    "zdr_get_counters": (C.c_int, [_vp, _u64p]),
    "zdr_reset_counters": (C.c_int, [_vp]),
    "zdr_kernel_launches": (C.c_uint64, []),
    "zdr_debug_phase_cycles": (C.c_int, [_vp, _u64p]),
    "zdr_probe_trace_closest": (C.c_int, [_vp, _f32p, C.c_int, _i32p, _f32p]),
    "zdr_probe_trace_any": (C.c_int, [_vp, _f32p, C.c_int, _i32p]),
    "zdr_probe_sampler": (C.c_int, [C.c_int, _i32p, C.c_int, C.c_char_p, _f32p]),
    "zdr_probe_render_samples": (C.c_int, [_vp, C.c_int, _pp, _f32p, _f32p]),
    "zdr_probe_ggx": (C.c_int, [C.c_int, _f32p, C.c_int, _f32p]),
    "zdr_obj_load": (C.c_int, [C.c_char_p, C.POINTER(_f32p), C.POINTER(C.c_int), C.POINTER(_i32p), C.POINTER(C.c_int)]),
    "zdr_obj_free": (None, [_vp]),
    "zdr_obj_last_error": (C.c_char_p, []),
}


class ZdrError(RuntimeError):
    pass


def load_library():
    """Loads libzdr_b200.so (once). Raises if it was not built: no fallback path exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C zdr_b200/csrc`). "
            "zdr_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    if lib.zdr_abi_version() != 1:
        raise ImportError("libzdr_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load_library().zdr_last_error()
        raise ZdrError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def np_ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))
