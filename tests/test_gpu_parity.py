"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): sampler values and primitive ids bit-exact,
images rel-L2 <= 1e-4, gradients rel-L2 <= 1e-3.  The forward pass shares the oracle's numeric
contract, so per-sample radiance is additionally compared bit-for-bit."""
import contextlib
import ctypes as C
import os

import numpy as np
import pytest

from zdr_b200 import _abi
from zdr_b200._abi import ZDR_COLLOCATED, ZDR_DIRECT, ZDR_PATH, make_params, np_ptr

from . import scenes

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4      # relative L2, images
GRAD_TOL = 1e-3     # relative L2, texture gradients


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (d if d > 0 else 1.0)


class GpuScene:
    """Thin ctypes driver of the C ABI with HOST buffers (no torch): what a cgo/JNI-style embedder does."""

    def __init__(self, models, device=0):
        from zdr_b200.load_obj import load_mesh
        from zdr_b200.mathtypes import float4x4
        from zdr_b200.tables import get_tables
        self.lib = _abi.load_library()
        pmj, bn, _ = get_tables()
        _abi.check(self.lib.zdr_set_sampler_tables(device, np_ptr(pmj, C.c_uint32), pmj.shape[0], pmj.shape[1],
                                                    np_ptr(bn, C.c_uint16), bn.shape[0], bn.shape[1]))
        self.h = C.c_void_p()
        _abi.check(self.lib.zdr_scene_create(C.byref(self.h), device))
        cache = {}
        for obj, xf, em in models:
            if obj not in cache:
                cache[obj] = load_mesh(obj)
            verts, tris = cache[obj]
            xfa = np.asarray((xf if xf is not None else float4x4(1.0)).flat(), np.float32)
            ema = np.asarray(list(em), np.float32)
            _abi.check(self.lib.zdr_scene_add_mesh(self.h, np_ptr(verts, C.c_float), verts.shape[0], np_ptr(tris, C.c_int32),
                                                    tris.shape[0] // 3, np_ptr(xfa, C.c_float), np_ptr(ema, C.c_float)))
        _abi.check(self.lib.zdr_scene_commit(self.h))

    def close(self):
        if self.h:
            self.lib.zdr_scene_destroy(self.h)
            self.h = None

    def set_emissions(self, e):
        e = np.ascontiguousarray(e, np.float32)
        _abi.check(self.lib.zdr_scene_set_emissions(self.h, np_ptr(e, C.c_float), e.shape[0]))

    def set_material_sets(self, sets):
        a = np.ascontiguousarray(sets, np.int32)
        _abi.check(self.lib.zdr_scene_set_material_sets(self.h, np_ptr(a, C.c_int32), a.shape[0]))

    def set_envmap(self, img, compensate=True):
        img = np.ascontiguousarray(img, np.float32)
        _abi.check(self.lib.zdr_scene_set_envmap(self.h, np_ptr(img, C.c_float), img.shape[1], img.shape[0], int(compensate)))

    def forward(self, integ, p, mat):
        mat = np.ascontiguousarray(mat, np.float32)
        img = np.zeros((p.height, p.width, 4), np.float32)
        _abi.check(self.lib.zdr_render_forward_host(self.h, integ, C.byref(p), np_ptr(mat, C.c_float), np_ptr(img, C.c_float)))
        return img

    def backward(self, integ, p, mat, dimg):
        mat, dimg = np.ascontiguousarray(mat, np.float32), np.ascontiguousarray(dimg, np.float32)
        dm = np.zeros_like(mat)
        _abi.check(self.lib.zdr_render_backward_host(self.h, integ, C.byref(p), np_ptr(mat, C.c_float), np_ptr(dimg, C.c_float),
                                                      np_ptr(dm, C.c_float)))
        return dm

    def samples(self, integ, p, mat):
        mat = np.ascontiguousarray(mat, np.float32)
        out = np.zeros((p.height, p.width, p.spp, 3), np.float32)
        _abi.check(self.lib.zdr_probe_render_samples(self.h, integ, C.byref(p), np_ptr(mat, C.c_float), np_ptr(out, C.c_float)))
        return out

    def trace_closest(self, rays):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        ip, uvt = np.empty((r.shape[0], 2), np.int32), np.empty((r.shape[0], 3), np.float32)
        _abi.check(self.lib.zdr_probe_trace_closest(self.h, np_ptr(r, C.c_float), r.shape[0], np_ptr(ip, C.c_int32), np_ptr(uvt, C.c_float)))
        return ip, uvt

    def trace_any(self, rays):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        occ = np.empty(r.shape[0], np.int32)
        _abi.check(self.lib.zdr_probe_trace_any(self.h, np_ptr(r, C.c_float), r.shape[0], np_ptr(occ, C.c_int32)))
        return occ

    def counters(self):
        out = (C.c_uint64 * _abi.ZDR_NUM_COUNTERS)()
        _abi.check(self.lib.zdr_get_counters(self.h, out))
        return dict(zip(_abi.COUNTER_NAMES, (int(x) for x in out)))

    def reset_counters(self):
        _abi.check(self.lib.zdr_reset_counters(self.h))


def bits(a):
    """bit pattern of float32 data, with every NaN mapped to one canonical pattern (a NaN sample
    is dropped by integrator.py:27 whatever its payload; x86 and sm_100a differ in NaN sign)."""
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


# ----------------------------------------------------------------------------- sampler (bit-exact)
@pytest.mark.parametrize("spp,bounces", [(1, 12), (4, 12), (64, 12), (256, 12), (1024, 12),
                                         (3, 12), (100, 12),          # not powers of two: cycle walking + signed floor-mod
                                         (100000, 12),                # more samples than a PMJ set holds: table index wraps
                                         (64, 40)])                   # dimension ~290: blue-noise texture index wraps > twice
def test_sampler_bit_exact(oracle, spp, bounces):
    rng = np.random.default_rng(spp)
    n = 4096
    st = np.stack([rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                   rng.choice([0, 1, 12345, 853402567, 19260817, 948377263, 2147483647, -5], n),
                   np.full(n, spp), rng.integers(0, spp, n)], axis=1).astype(np.int32)
    prog = "2" + "112" + "12" + "1" + ("11212" * bounces)     # camera, light, bsdf, rr, ... up to dimension > 64
    ref = oracle.sampler_draws(st, prog)
    out = np.empty_like(ref)
    lib = _abi.load_library()
    from zdr_b200.tables import get_tables
    pmj, bn, _ = get_tables()
    _abi.check(lib.zdr_set_sampler_tables(0, np_ptr(pmj, C.c_uint32), pmj.shape[0], pmj.shape[1], np_ptr(bn, C.c_uint16), bn.shape[0], bn.shape[1]))
    _abi.check(lib.zdr_probe_sampler(0, np_ptr(st, C.c_int32), n, prog.encode(), np_ptr(out, C.c_float)))
    assert np.array_equal(bits(out), bits(ref))
    assert (ref >= 0).all() and (ref < 1).all()


def test_sampler_test_sampler_py_inputs(oracle):
    """test_sampler.py:23-30: pixel (24,345), seed 0, n=32, second next2f()."""
    st = np.array([[24, 345, 0, 32, i] for i in range(32)], np.int32)
    ref = oracle.sampler_draws(st, "22")
    out = np.empty_like(ref)
    lib = _abi.load_library()
    _abi.check(lib.zdr_probe_sampler(0, np_ptr(st, C.c_int32), 32, b"22", np_ptr(out, C.c_float)))
    assert np.array_equal(bits(out), bits(ref))


# --------------------------------------------------------------------------- intersection (exact ids)
def _random_rays(n, seed, lo, hi):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d, np.zeros((n, 1)), np.full((n, 1), 1e30)], axis=1).astype(np.float32)


@pytest.mark.parametrize("name", ["cbox", "spheres"])
def test_trace_prim_ids_exact(oracle, name):
    if name == "cbox":
        models, lo, hi, n = scenes.cbox_path_models(), (-3.5, -0.5, -3.5), (3.5, 6.0, 6.5), 20000
    else:
        models, lo, hi, n = scenes.sphere_grid_models(12, seed=1), (-1, -1, -1), (4, 4, 2), 4000
    g, o = GpuScene(models), oracle.OracleScene(models)
    rays = _random_rays(n, 7, lo, hi)
    # axis-aligned and edge-grazing rays too: shared edges are where tie-breaking matters
    rays[:64, 3:6] = np.eye(3, dtype=np.float32)[np.arange(64) % 3] * np.where(np.arange(64) % 2, 1, -1)[:, None]
    ip_g, uvt_g = g.trace_closest(rays)
    ip_o, uvt_o = o.trace_closest(rays)
    assert np.array_equal(ip_g, ip_o)
    hit = ip_o[:, 0] >= 0
    assert hit.sum() > n // 20
    assert np.array_equal(bits(uvt_g[hit]), bits(uvt_o[hit]))
    # shadow rays: bounded segments
    seg = rays.copy()
    seg[:, 6] = 1e-4
    seg[:, 7] = np.random.default_rng(3).uniform(0.1, 6.0, n)
    assert np.array_equal(g.trace_any(seg), o.trace_any(seg))
    g.close()


def test_trace_shared_edge_rays(oracle):
    """Rays aimed exactly at mesh vertices / edge midpoints of the Cornell box: watertightness and
    the lowest-id tie-break must agree with the oracle's in-order scan."""
    from zdr_b200.load_obj import load_mesh
    models = scenes.cbox_path_models()
    verts, tris = load_mesh(models[0][0])
    P = verts[:, :3][tris.reshape(-1, 3)]                       # (T,3,3)
    targets = np.concatenate([P.reshape(-1, 3), (P[:, 0] + P[:, 1]) / 2, (P[:, 1] + P[:, 2]) / 2, (P[:, 0] + P[:, 2]) / 2])
    origin = np.array([-0.2, 2.6, 6.0], np.float32)
    d = targets - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([np.broadcast_to(origin, d.shape), d, np.zeros((len(d), 1)), np.full((len(d), 1), 1e30)], axis=1).astype(np.float32)
    g, o = GpuScene(models), oracle.OracleScene(models)
    ip_g, uvt_g = g.trace_closest(rays)
    ip_o, uvt_o = o.trace_closest(rays)
    assert np.array_equal(ip_g, ip_o)
    assert np.array_equal(bits(uvt_g[ip_o[:, 0] >= 0]), bits(uvt_o[ip_o[:, 0] >= 0]))
    g.close()


# ------------------------------------------------------------------------------- BRDF + adjoint
def test_ggx_value_and_adjoint(oracle):
    rng = np.random.default_rng(5)
    n = 20000

    def hemi(k):
        v = rng.normal(size=(k, 3))
        v[:, 2] = np.abs(v[:, 2]) + 1e-3
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    inp = np.concatenate([hemi(n), hemi(n), rng.uniform(0, 1, (n, 3)), rng.uniform(0.02, 1, (n, 1)), rng.normal(size=(n, 3))], axis=1).astype(np.float32)
    ref = oracle.ggx(inp)
    out = np.empty_like(ref)
    _abi.check(_abi.load_library().zdr_probe_ggx(0, np_ptr(inp, C.c_float), n, np_ptr(out, C.c_float)))
    assert np.array_equal(bits(out), bits(ref))


# ------------------------------------------------------------------ per-sample radiance + image + grads
CASES = {
    # name: (models, integrator, res, spp, material, kwargs)
    "c1_collocated": (scenes.cbox_collocated_models, ZDR_COLLOCATED, (256, 256), 1, lambda: scenes.random_material((1024, 1024), 0), {}),
    "collocated_spp4_box": (scenes.cbox_collocated_models, ZDR_COLLOCATED, (96, 64), 4, lambda: scenes.random_material((64, 64), 1), {"use_tent_filter": False}),
    "c2_direct_small": (scenes.cbox_direct_models, ZDR_DIRECT, (128, 128), 8, lambda: scenes.random_material((1024, 1024), 0), {}),
    "direct_uv_nomis": (scenes.cbox_path_models, ZDR_DIRECT, (96, 96), 4, lambda: scenes.random_material((256, 256), 2), {"use_mis": False}),
    "direct_uv_nearest": (scenes.cbox_path_models, ZDR_DIRECT, (64, 64), 4, lambda: scenes.random_material((32, 32), 3), {"bilinear": False}),
    "c3_path_small": (scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, lambda: scenes.random_material((1024, 1024), 0), {"max_depth": 8}),
    "path_literal": (scenes.cbox_path_models, ZDR_PATH, (64, 64), 4, lambda: scenes.random_material((128, 128), 4), {"max_depth": 5, "prb_mode": _abi.ZDR_PRB_LITERAL}),
    "path_depth16": (scenes.cbox_path_models, ZDR_PATH, (48, 48), 4, lambda: scenes.random_material((128, 128), 5), {"max_depth": 16}),
    # the wavefront schedule of the same integrator (default from 2^21 camera samples per launch on)
    "c3_path_small_wavefront": (scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, lambda: scenes.random_material((1024, 1024), 0), {"max_depth": 8, "schedule": "wavefront"}),
    "path_literal_wavefront": (scenes.cbox_path_models, ZDR_PATH, (64, 64), 4, lambda: scenes.random_material((128, 128), 4), {"max_depth": 5, "prb_mode": _abi.ZDR_PRB_LITERAL, "schedule": "wavefront"}),
    "path_depth1_wavefront": (scenes.cbox_path_models, ZDR_PATH, (40, 24), 3, lambda: scenes.random_material((64, 64), 6), {"max_depth": 1, "schedule": "wavefront"}),
    # ... with the rays of the traversal stages binned by (origin cell, direction octant) from the second bounce on (k_wf_bin), and
    # on the general traversal kernel instead of the shared-memory / register-stack one small scenes get (k_wf_trace_tiny)
    "c3_path_small_wavefront_binned": (scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, lambda: scenes.random_material((1024, 1024), 0), {"max_depth": 8, "schedule": "wavefront", "env": {"ZDR_WF_BIN": "3"}}),
    "path_literal_wavefront_binned": (scenes.cbox_path_models, ZDR_PATH, (64, 64), 4, lambda: scenes.random_material((128, 128), 4), {"max_depth": 5, "prb_mode": _abi.ZDR_PRB_LITERAL, "schedule": "wavefront", "env": {"ZDR_WF_BIN": "1"}}),
    "path_wavefront_binned_shadow_only": (scenes.cbox_path_models, ZDR_PATH, (64, 64), 4, lambda: scenes.random_material((128, 128), 4), {"max_depth": 5, "schedule": "wavefront", "env": {"ZDR_WF_BIN": "2"}}),
    "c3_path_small_wavefront_general": (scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, lambda: scenes.random_material((1024, 1024), 0), {"max_depth": 8, "schedule": "wavefront", "env": {"ZDR_WF_TINY_OFF": "1", "ZDR_WF_BIN": "0"}}),
    "c3_path_small_wavefront_general_binned": (scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, lambda: scenes.random_material((1024, 1024), 0), {"max_depth": 8, "schedule": "wavefront", "env": {"ZDR_WF_TINY_OFF": "1", "ZDR_WF_BIN": "3"}}),
    # the persistent schedule (path state and ray queues in shared memory; persistent.cu)
    "c3_path_small_persistent": (scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, lambda: scenes.random_material((1024, 1024), 0), {"max_depth": 8, "schedule": "persistent"}),
    "path_literal_persistent": (scenes.cbox_path_models, ZDR_PATH, (64, 64), 4, lambda: scenes.random_material((128, 128), 4), {"max_depth": 5, "prb_mode": _abi.ZDR_PRB_LITERAL, "schedule": "persistent"}),
    "path_depth1_persistent": (scenes.cbox_path_models, ZDR_PATH, (40, 24), 3, lambda: scenes.random_material((64, 64), 6), {"max_depth": 1, "schedule": "persistent"}),
    "path_depth16_persistent": (scenes.cbox_path_models, ZDR_PATH, (48, 48), 4, lambda: scenes.random_material((128, 128), 5), {"max_depth": 16, "schedule": "persistent"}),
    # enough samples that every CTA's pool turns over many times (148 pools x 1536 slots = 227 K paths in flight)
    "path_persistent_many": (scenes.cbox_path_models, ZDR_PATH, (256, 192), 16, lambda: scenes.random_material((256, 256), 7), {"max_depth": 6, "schedule": "persistent"}),
}


@pytest.fixture(scope="module")
def case_cache():
    return {}


@contextlib.contextmanager
def _env(env):
    """Library switches that are read per launch (ZDR_WF_BIN, ZDR_WF_TINY_OFF), set for the duration of a case."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run_case(oracle, name, cache):
    if name in cache:
        return cache[name]
    models_fn, integ, res, spp, mat_fn, kw = CASES[name]
    models, mat = models_fn(), mat_fn()
    g, o = GpuScene(models), oracle.OracleScene(models)
    kw = dict(kw)
    env = kw.pop("env", {})
    out = {"integ": integ, "mat": mat, "g": g, "o": o, "res": res, "spp": spp, "kw": kw, "env": env}
    cache[name] = out
    return out


@pytest.mark.parametrize("name", list(CASES))
def test_forward_parity(oracle, case_cache, name):
    c = _run_case(oracle, name, case_cache)
    p = make_params(scenes.CBOX_CAMERA, c["res"], c["mat"].shape, c["spp"], 0, **c["kw"])
    with _env(c["env"]):
        s_g = c["g"].samples(c["integ"], p, c["mat"])
        img_g = c["g"].forward(c["integ"], p, c["mat"])
    s_o = c["o"].render_samples(c["integ"], p, c["mat"])
    same = bits(s_g) == bits(s_o)
    assert same.all(), f"{(~same).sum()} of {same.size} per-sample radiance values differ; max abs {np.nanmax(np.abs(s_g - s_o))}"
    img_o = c["o"].render_forward(c["integ"], p, c["mat"])
    assert np.array_equal(bits(img_g), bits(img_o))
    assert rel_l2(img_g, img_o) <= IMG_TOL
    assert (img_o[..., 3] == 1).all() and img_o[..., :3].max() > 0


@pytest.mark.parametrize("name", list(CASES))
def test_backward_parity(oracle, case_cache, name):
    c = _run_case(oracle, name, case_cache)
    p = make_params(scenes.CBOX_CAMERA, c["res"], c["mat"].shape, c["spp"], 1, **c["kw"])      # seed+1, render.py:196
    rng = np.random.default_rng(11)
    dimg = rng.normal(size=(c["res"][1], c["res"][0], 4)).astype(np.float32)
    c["g"].reset_counters()
    c["o"].reset_ray_counters()
    with _env(c["env"]):
        dm_g = c["g"].backward(c["integ"], p, c["mat"], dimg)
    dm_o = c["o"].render_backward(c["integ"], p, c["mat"], dimg)
    assert np.abs(dm_o).max() > 0
    assert rel_l2(dm_g, dm_o) <= GRAD_TOL, rel_l2(dm_g, dm_o)
    # the same texels are touched
    assert np.array_equal(dm_g != 0, dm_o != 0) or rel_l2(dm_g, dm_o) < 1e-5
    # ray counts identical (the metric's numerator)
    cg, (oc, oa) = c["g"].counters(), c["o"].ray_counters()
    assert (cg["closest_rays"], cg["any_rays"]) == (oc, oa)


@pytest.mark.parametrize("case", ["c3_path_small", "c3_path_small_wavefront", "c3_path_small_persistent"])
def test_shard_rows_and_samples(oracle, case_cache, case):
    """Row tiles are bit-identical to the full render; sample-range shards sum to it (§8e)."""
    c = _run_case(oracle, case, case_cache)
    W, H = c["res"]
    full = c["g"].forward(c["integ"], make_params(scenes.CBOX_CAMERA, c["res"], c["mat"].shape, c["spp"], 0, **c["kw"]), c["mat"])
    acc = np.zeros_like(full)
    for r0, r1 in [(0, 17), (17, 64), (64, H)]:
        p = make_params(scenes.CBOX_CAMERA, c["res"], c["mat"].shape, c["spp"], 0, rows=(r0, r1), **c["kw"])
        part = c["g"].forward(c["integ"], p, c["mat"])
        assert not part[:r0].any() and not part[r1:].any()
        acc += part
    assert np.array_equal(bits(acc), bits(full))
    acc = np.zeros_like(full)
    for k in range(3):                                   # round-robin 8-row tiles (multi-GPU sharding)
        p = make_params(scenes.CBOX_CAMERA, c["res"], c["mat"].shape, c["spp"], 0, shard=(k, 3), **c["kw"])
        acc += c["g"].forward(c["integ"], p, c["mat"])
    assert np.array_equal(bits(acc), bits(full))
    dimg = np.ones((H, W, 4), np.float32)
    pf = make_params(scenes.CBOX_CAMERA, c["res"], c["mat"].shape, c["spp"], 1, **c["kw"])
    g_full = c["g"].backward(c["integ"], pf, c["mat"], dimg)
    g_sum = np.zeros_like(g_full)
    for r0, r1 in [(0, 40), (40, H)]:
        for i0, i1 in [(0, 3), (3, c["spp"])]:
            p = make_params(scenes.CBOX_CAMERA, c["res"], c["mat"].shape, c["spp"], 1, rows=(r0, r1), its=(i0, i1), **c["kw"])
            g_sum += c["g"].backward(c["integ"], p, c["mat"], dimg)
    assert rel_l2(g_sum, g_full) < 1e-5


def test_scenes_of_one_or_two_leaves_all_schedules(oracle):
    """Scenes whose tree is a single leaf or one node (root link negative / no stack traffic): a turned-over, enlarged copy of the
    light quad as the textured floor under the light, and the floor alone; every schedule (the wavefront one takes the
    shared-memory traversal kernel here), binned and not."""
    from zdr_b200.mathtypes import float4x4
    floor = float4x4(-3, 0, 0, 0, 0, -1, 0, 0, 0, 0, 3, 0, -0.936, 6.319, 6.08, 1)      # half turn about z, x and z enlarged
    light = scenes.asset("cbox-light.obj")
    mat = scenes.random_material((64, 64), 12)
    for models in ([(light, floor, scenes.float3(0.0)), (light, None, scenes.float3(17, 12, 4))], [(light, floor, scenes.float3(0.5, 0.4, 0.3))]):
        g, o = GpuScene(models), oracle.OracleScene(models)
        for sched, env in (("megakernel", {}), ("wavefront", {}), ("wavefront", {"ZDR_WF_BIN": "3"}), ("wavefront", {"ZDR_WF_TINY_OFF": "1"}), ("persistent", {})):
            p = make_params(scenes.CBOX_CAMERA, (64, 48), mat.shape, 4, 0, max_depth=4, schedule=sched)
            with _env(env):
                s_g = g.samples(ZDR_PATH, p, mat)
            s_o = o.render_samples(ZDR_PATH, p, mat)
            assert np.array_equal(bits(s_g), bits(s_o)), (len(models), sched, env)
            pb = make_params(scenes.CBOX_CAMERA, (64, 48), mat.shape, 4, 1, max_depth=4, schedule=sched)
            dimg = np.ones((48, 64, 4), np.float32)
            with _env(env):
                dm_g = g.backward(ZDR_PATH, pb, mat, dimg)
            dm_o = o.render_backward(ZDR_PATH, pb, mat, dimg)
            assert rel_l2(dm_g, dm_o) <= GRAD_TOL or not dm_o.any(), (len(models), sched, env, rel_l2(dm_g, dm_o))
        assert s_o.max() > 0 or len(models) == 1
        g.close()


@pytest.mark.parametrize("name", ["collocated_large", "direct_large", "direct_xform_large", "path_literal_large", "path_literal_large_smooth",
                                  "path_corrected_large", "direct_env_large", "path_literal_env_large"])
def test_cuda_vs_reference_source_fixtures(name):
    """The CUDA path directly against the large committed fixtures that tests/golden/gen_golden.py produced by running the
    REFERENCE'S OWN collocated.py / direct.py / prb.py under the LuisaCompute emulator (48^2-64^2 pixels, 4-8 spp, 128^2 white-noise
    texture, depth 16, roughness down to 0; no oracle in between).  Without an envmap the image is the reference source's BIT FOR BIT
    on every schedule (tests/test_golden.py: so is the oracle's) and the gradient within 1e-5; with one, the north-star bars 1e-4 / 1e-3."""
    from .test_golden import LARGE_FIXTURES, load_large_fixture
    _, models, integ, extra, bit_exact = next(f for f in LARGE_FIXTURES if f[0] == name)
    z, res, spp, seed, mat = load_large_fixture(name)
    g = GpuScene(models())
    if "envmap" in z:
        g.set_envmap(z["envmap"], True)
    kws = [{"schedule": s_} for s_ in ("megakernel", "wavefront", "persistent")] if integ == ZDR_PATH else [{}]
    for kw in kws:
        img = g.forward(integ, make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, seed, **extra, **kw), mat)
        if bit_exact:
            assert np.array_equal(img.view(np.uint32), z["image"].view(np.uint32)), (kw, rel_l2(img, z["image"]))
        # envmap: acosf / atan2f of direction_to_uv are CUDA's here and libm's in the fixture (a pdf cell can flip): the north-star bars
        assert rel_l2(img, z["image"]) <= (1e-6 if bit_exact else IMG_TOL), (kw, rel_l2(img, z["image"]))
        dm = g.backward(integ, make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, seed + 1, **extra, **kw), mat, z["d_image"])
        assert np.abs(z["d_material"]).max() > 0
        # (path_corrected_large: a tenth of the north star's bar; the oracle sits 9e-8 from the fixture and the CUDA path ~4e-8 from
        # the oracle in this mode — profiles/r05f_corrected_fixture.log)
        grad_tol = GRAD_TOL if not bit_exact else 1e-4 if name == "path_corrected_large" else 1e-5
        assert rel_l2(dm, z["d_material"]) <= grad_tol, (kw, rel_l2(dm, z["d_material"]))
    g.close()


def test_cuda_vs_reference_source_material_sets_fixture():
    """... and against the large material-set fixture (the reference's direct.py / interaction.py with the set-index patch, two 64^2
    texture sets, 48 x 48 pixels at 8 spp): image bit for bit, both sets' gradients within 1e-5."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_direct_material_sets_large.npz"))
    res, spp, seed = tuple(int(x) for x in z["res"]), int(z["spp"]), int(z["seed"])
    models, sets = scenes.cbox_material_sets_models()
    mat = z["material"]
    g = GpuScene(models)
    g.set_material_sets(sets)
    img = g.forward(ZDR_DIRECT, make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, seed), mat)
    assert np.array_equal(img.view(np.uint32), z["image"].view(np.uint32)), rel_l2(img, z["image"])
    dm = g.backward(ZDR_DIRECT, make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, seed + 1), mat, z["d_image"])
    for k in range(2):
        assert rel_l2(dm[k], z["d_material"][k]) <= 1e-5, (k, rel_l2(dm[k], z["d_material"][k]))
    g.close()


def test_transformed_instances_and_update_lights(oracle):
    """test_lightstage.py shape: transformed instances, lights toggled with update_lights."""
    models = scenes.sphere_grid_models(6, seed=2)
    models[3] = (models[3][0], models[3][1], scenes.float3(9.0, 7.0, 5.0))
    models[5] = (models[5][0], models[5][1], scenes.float3(0.0))
    g, o = GpuScene(models), oracle.OracleScene(models)
    mat = scenes.random_material((64, 64), 9)
    cam = scenes.Camera(fov=0.9, origin=scenes.float3(-2.0, 2.5, -4.0), target=scenes.float3(1.0, 0.3, 0.3), up=scenes.float3(0, 1, 0))
    for emis in [None, [(0, 0, 0), (0, 0, 0), (3, 3, 3), (0, 0, 0), (0, 0, 0), (5, 1, 1)]]:
        if emis is not None:
            g.set_emissions(np.asarray(emis, np.float32))
            o.update_lights(emis)
        for integ, sched in ((ZDR_DIRECT, "auto"), (ZDR_PATH, "megakernel"), (ZDR_PATH, "wavefront"), (ZDR_PATH, "persistent")):
            p = make_params(cam, (80, 60), mat.shape, 4, 0, max_depth=4, schedule=sched)
            assert np.array_equal(bits(g.samples(integ, p, mat)), bits(o.render_samples(integ, p, mat)))
            dimg = np.ones((60, 80, 4), np.float32)
            assert rel_l2(g.backward(integ, p, mat, dimg), o.render_backward(integ, p, mat, dimg)) <= GRAD_TOL
    g.close()


def test_envmap_path(oracle):
    """C5 shape, small: instanced spheres + procedural envmap, path depth 6."""
    models = scenes.sphere_grid_models(8, seed=4)
    env = scenes.procedural_envmap(128, 64, seed=1)
    g, o = GpuScene(models), oracle.OracleScene(models)
    g.set_envmap(env)
    o.set_envmap(env)
    mat = scenes.random_material((64, 64), 10)
    cam = scenes.Camera(fov=0.9, origin=scenes.float3(-2.0, 2.5, -4.0), target=scenes.float3(1.0, 0.3, 0.3), up=scenes.float3(0, 1, 0))
    for integ, sched in ((ZDR_DIRECT, "auto"), (ZDR_PATH, "megakernel"), (ZDR_PATH, "wavefront"), (ZDR_PATH, "persistent")):
        p = make_params(cam, (64, 48), mat.shape, 8, 0, max_depth=6, schedule=sched)
        img_g, img_o = g.forward(integ, p, mat), o.render_forward(integ, p, mat)
        assert rel_l2(img_g, img_o) <= IMG_TOL, rel_l2(img_g, img_o)     # acosf/atan2f are libm vs CUDA: tolerance, not bits
        dimg = np.ones((48, 64, 4), np.float32)
        assert rel_l2(g.backward(integ, p, mat, dimg), o.render_backward(integ, p, mat, dimg)) <= GRAD_TOL
    g.close()


def test_envmap_with_mesh_lights_and_consistent_light_pdf(oracle):
    """Envmap AND emissive meshes in one scene: the reference's sample_light_pdf leaves the envmap out of its light
    count (light.py:89 vs :26).  Both behaviours — the literal one and ZDR_FLAG_LIGHT_PDF_COUNTS_ENV — must match
    the oracle, and they must differ from each other (the flag is not a no-op)."""
    models = scenes.cbox_path_models()                          # Cornell box (open front) + its area light + an envmap
    env = scenes.procedural_envmap(128, 64, seed=2) * np.float32(0.2)
    g, o = GpuScene(models), oracle.OracleScene(models)
    g.set_envmap(env)
    o.set_envmap(env)
    mat = scenes.random_material((32, 32), 12)
    cam = scenes.CBOX_CAMERA
    dimg = np.ones((48, 64, 4), np.float32)
    for integ, sched in ((ZDR_DIRECT, "auto"), (ZDR_PATH, "megakernel"), (ZDR_PATH, "wavefront"), (ZDR_PATH, "persistent")):
        imgs = []
        for flag in (False, True):
            p = make_params(cam, (64, 48), mat.shape, 8, 3, max_depth=5, schedule=sched, light_pdf_counts_env=flag)
            img_g, img_o = g.forward(integ, p, mat), o.render_forward(integ, p, mat)
            assert rel_l2(img_g, img_o) <= IMG_TOL, (integ, flag, rel_l2(img_g, img_o))
            assert rel_l2(g.backward(integ, p, mat, dimg), o.render_backward(integ, p, mat, dimg)) <= GRAD_TOL
            imgs.append(img_g)
        assert rel_l2(imgs[0], imgs[1]) > 1e-3
    g.close()


@pytest.mark.parametrize("builder", ["csah", "lbvh"])
def test_lbvh_large_scene(oracle, builder, monkeypatch):
    """> 16 K triangles takes the GPU builders (C5 shape: instanced sphere.obj) — clustered SAH (default) and the
    plain LBVH: closest hits, occlusion and a small path render + adjoint must match the oracle's brute-force
    scan exactly."""
    monkeypatch.setenv("ZDR_BVH", builder)
    models = scenes.sphere_grid_models(20, seed=3, emissive_every=7)          # 19 200 triangles, 2 emissive spheres
    g, o = GpuScene(models), oracle.OracleScene(models)
    assert g.lib.zdr_scene_triangle_count(g.h) == 19200
    rays = _random_rays(3000, 11, (-1, -1, -1), (5, 5, 2))
    ip_g, uvt_g = g.trace_closest(rays)
    ip_o, uvt_o = o.trace_closest(rays)
    assert np.array_equal(ip_g, ip_o)
    hit = ip_o[:, 0] >= 0
    assert hit.sum() > 100 and np.array_equal(bits(uvt_g[hit]), bits(uvt_o[hit]))
    seg = rays.copy()
    seg[:, 6], seg[:, 7] = 1e-4, 3.0
    assert np.array_equal(g.trace_any(seg), o.trace_any(seg))
    cam = scenes.Camera(fov=0.9, origin=scenes.float3(-2.0, 3.0, -4.0), target=scenes.float3(2.0, 0.5, 0.3), up=scenes.float3(0, 1, 0))
    mat = scenes.random_material((64, 64), 2)
    # (the wavefront schedule bins the rays of its traversal stages by default in scenes of this size; ZDR_WF_BIN=0: the plain lists)
    for sched, env in (("megakernel", {}), ("wavefront", {}), ("wavefront", {"ZDR_WF_BIN": "0"}), ("persistent", {})):
        p = make_params(cam, (48, 32), mat.shape, 2, 0, max_depth=4, schedule=sched)
        with _env(env):
            s_g = g.samples(ZDR_PATH, p, mat)
        assert np.array_equal(bits(s_g), bits(o.render_samples(ZDR_PATH, p, mat))), (sched, env)
        d = np.ones((32, 48, 4), np.float32)
        with _env(env):
            dm_g = g.backward(ZDR_PATH, p, mat, d)
        assert rel_l2(dm_g, o.render_backward(ZDR_PATH, p, mat, d)) <= GRAD_TOL, (sched, env)
    g.close()


def test_clustered_sah_degenerate_geometry(oracle, tmp_path):
    """Large-scene builder on hostile input: 18 K random triangles of very different sizes, 400 exact copies of one
    triangle (coincident centroids: the binned split has nothing to separate, the builder falls back to splitting
    by position), zero-area triangles, one huge triangle over everything.  Hits must equal the brute-force scan,
    ties going to the lowest scene-order id."""
    rng = np.random.default_rng(5)
    n = 18000
    ctr = rng.uniform(-4, 4, (n, 1, 3))
    size = np.exp(rng.uniform(np.log(1e-3), np.log(0.8), (n, 1, 1)))
    tri = ctr + rng.normal(0, 1, (n, 3, 3)) * size
    dup = np.repeat(np.array([[[0.5, 0.5, 0.0], [1.5, 0.5, 0.2], [0.5, 1.5, 0.1]]]), 400, axis=0)
    flat = np.repeat(rng.uniform(-4, 4, (200, 1, 3)), 3, axis=1)                     # zero area
    huge = np.array([[[-50.0, -50.0, -4.5], [50.0, -50.0, -4.5], [0.0, 80.0, -4.5]]])
    tri = np.concatenate([tri[:9000], dup, flat, tri[9000:], huge]).astype(np.float32)
    path = tmp_path / "soup.obj"
    with open(path, "w") as f:
        for t in tri:
            for v in t:
                f.write(f"v {float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")
        for i in range(len(tri)):
            f.write(f"f {3 * i + 1} {3 * i + 2} {3 * i + 3}\n")
    models = [(str(path), None, scenes.float3(0.0))]
    g, o = GpuScene(models), oracle.OracleScene(models)
    assert g.lib.zdr_scene_triangle_count(g.h) == len(tri)
    rays = _random_rays(4000, 21, (-4, -4, -4), (4, 4, 4))
    rays[:500, 0:3] = (0.8, 0.8, 1.0)                                                # straight at the stack of copies
    rays[:500, 3:6] = rng.normal(0, 0.05, (500, 3)) + (0.0, 0.0, -1.0)
    ip_g, uvt_g = g.trace_closest(rays)
    ip_o, uvt_o = o.trace_closest(rays)
    assert np.array_equal(ip_g, ip_o)
    hit = ip_o[:, 0] >= 0
    assert hit.sum() > 1000 and np.array_equal(bits(uvt_g[hit]), bits(uvt_o[hit]))
    assert not ((ip_o[:500, 1] > 9000) & (ip_o[:500, 1] < 9400)).any()
    assert (ip_o[:500, 1] == 9000).sum() > 30                                        # the first copy wins every tie
    seg = rays.copy()
    seg[:, 6], seg[:, 7] = 1e-4, 2.0
    assert np.array_equal(g.trace_any(seg), o.trace_any(seg))
    g.close()


def test_light_stage_uniform_texture(oracle):
    """C4 shape: sphere.obj lit by toggled emissive quads, 1x1 and constant textures (every scatter of the
    launch lands on one texel), collocated / direct / path, both path schedules."""
    quad = scenes.asset("quad.obj")
    models = [(scenes.asset("sphere.obj"), None, scenes.float3(0.0)),
              (quad, scenes.translate_scale(0.0, 3.0, 0.0, 1.5), scenes.float3(12.0)),
              (quad, scenes.translate_scale(2.5, 2.0, 1.0, 1.0), scenes.float3(0.0))]
    g, o = GpuScene(models), oracle.OracleScene(models)
    cam = scenes.Camera(fov=0.7, origin=scenes.float3(0.0, 0.6, 3.2), target=scenes.float3(0.0, 0.0, 0.0), up=scenes.float3(0, 1, 0))
    for mat in (scenes.constant_material((1, 1)), scenes.constant_material((16, 16), (0.8, 0.4, 0.2, 0.6))):
        for lights in ([(0, 0, 0), (12, 12, 12), (0, 0, 0)], [(0, 0, 0), (0, 0, 0), (9, 6, 3)]):
            g.set_emissions(np.asarray(lights, np.float32))
            o.update_lights(lights)
            for integ, kw in ((ZDR_COLLOCATED, {}), (ZDR_DIRECT, {}), (ZDR_PATH, {"max_depth": 3, "schedule": "megakernel"}),
                              (ZDR_PATH, {"max_depth": 3, "schedule": "wavefront"}), (ZDR_PATH, {"max_depth": 3, "schedule": "persistent"})):
                p = make_params(cam, (64, 64), mat.shape, 4, 2, **kw)
                assert np.array_equal(bits(g.forward(integ, p, mat)), bits(o.render_forward(integ, p, mat)))
                d = np.ones((64, 64, 4), np.float32)
                gg, go = g.backward(integ, p, mat, d), o.render_backward(integ, p, mat, d)
                assert rel_l2(gg, go) <= GRAD_TOL, (integ, rel_l2(gg, go))
    g.close()


@pytest.mark.parametrize("integ,kw", [(ZDR_DIRECT, {}), (ZDR_PATH, {"max_depth": 4, "schedule": "megakernel"}),
                                      (ZDR_PATH, {"max_depth": 4, "schedule": "wavefront"}), (ZDR_PATH, {"max_depth": 4, "schedule": "persistent"})])
def test_ragged_shapes_and_odd_sample_counts(oracle, integ, kw):
    """Non-square image (aspect ratio in integrator.py:23), non-square texture (row stride = shape[0],
    render.py:161 quirk), sample counts that are not powers of two (cycle-walking permutation), a seed that
    wraps at seed+1 (render.py:196), image sizes that are not multiples of the 16x8 block tile."""
    models = scenes.cbox_path_models()
    g, o = GpuScene(models), oracle.OracleScene(models)
    rng = np.random.default_rng(21)
    for (W, H), tex, spp, seed in [((37, 19), (24, 40), 3, 5), ((50, 70), (40, 24), 5, 2147483647), ((16, 8), (7, 7), 12, -3), ((1, 1), (1, 1), 1, 0)]:
        mat = np.clip(rng.random((tex[0], tex[1], 4), dtype=np.float32), 1e-3, 1.0)
        p = make_params(scenes.CBOX_CAMERA, (W, H), mat.shape, spp, seed, **kw)
        assert np.array_equal(bits(g.samples(integ, p, mat)), bits(o.render_samples(integ, p, mat))), (W, H, tex, spp)
        assert np.array_equal(bits(g.forward(integ, p, mat)), bits(o.render_forward(integ, p, mat)))
        pb = make_params(scenes.CBOX_CAMERA, (W, H), mat.shape, spp, _abi.wrap_i32(seed + 1), **kw)
        d = rng.normal(size=(H, W, 4)).astype(np.float32)
        gg, go = g.backward(integ, pb, mat, d), o.render_backward(integ, pb, mat, d)
        assert rel_l2(gg, go) <= GRAD_TOL or np.abs(go).max() == 0
    g.close()


def test_nan_and_degenerate_inputs(oracle):
    """NaN hygiene of integrator.py:27,38-40 and prb.py:100-101,179: NaN cotangents are zeroed, NaN texels
    poison only the samples that read them, degenerate (zero-area) triangles are never hit."""
    models = scenes.cbox_path_models()
    g, o = GpuScene(models), oracle.OracleScene(models)
    mat = scenes.random_material((32, 32), 8)
    mat[5:9, 5:9, :] = np.nan
    p = make_params(scenes.CBOX_CAMERA, (48, 48), mat.shape, 4, 0, max_depth=4)
    img_g, img_o = g.forward(ZDR_PATH, p, mat), o.render_forward(ZDR_PATH, p, mat)
    assert np.array_equal(bits(img_g), bits(img_o)) and np.isfinite(img_o).all()           # NaN samples are dropped
    d = np.ones((48, 48, 4), np.float32)
    d[10:20, 10:20] = np.nan
    gg, go = g.backward(ZDR_PATH, p, scenes.random_material((32, 32), 8), d), o.render_backward(ZDR_PATH, p, scenes.random_material((32, 32), 8), d)
    assert np.isfinite(go).all() and rel_l2(gg, go) <= GRAD_TOL
    g.close()
    # a mesh with a zero-area triangle and a sliver
    verts = np.array([[0, 0, 0, 0, 0, 0, 0, 1], [1, 0, 0, 1, 0, 0, 0, 1], [0, 1, 0, 0, 1, 0, 0, 1], [2, 2, 0, 1, 1, 0, 0, 1], [1e-7, 1, 0, 0, 1, 0, 0, 1]], np.float32)
    tris = np.array([0, 1, 2, 3, 3, 3, 0, 2, 4], np.int32)
    lib = _abi.load_library()
    h = C.c_void_p()
    _abi.check(lib.zdr_scene_create(C.byref(h), 0))
    _abi.check(lib.zdr_scene_add_mesh(h, np_ptr(verts, C.c_float), 5, np_ptr(tris, C.c_int32), 3, None, None))
    _abi.check(lib.zdr_scene_commit(h))
    os_ = oracle.OracleScene([((verts, tris), None, None)])
    rays = _random_rays(2000, 4, (-0.5, -0.5, 1.0), (1.5, 1.5, 2.0))
    rays[:, 3:6] = [0, 0, -1]
    ip, uvt = np.empty((2000, 2), np.int32), np.empty((2000, 3), np.float32)
    _abi.check(lib.zdr_probe_trace_closest(h, np_ptr(rays, C.c_float), 2000, np_ptr(ip, C.c_int32), np_ptr(uvt, C.c_float)))
    ip_o, _ = os_.trace_closest(rays)
    assert np.array_equal(ip, ip_o) and not (ip[:, 1] == 1).any()
    lib.zdr_scene_destroy(h)


def test_error_codes():
    lib = _abi.load_library()
    h = C.c_void_p()
    _abi.check(lib.zdr_scene_create(C.byref(h), 0))
    assert lib.zdr_scene_commit(h) == -2                                                  # ZDR_ERR_STATE: empty scene
    p = make_params(scenes.CBOX_CAMERA, (8, 8), (4, 4), 1, 0)
    assert lib.zdr_render_forward(h, 1, C.byref(p), 1, 1, None) == -2                    # not committed
    verts = np.zeros((3, 8), np.float32)
    bad = np.array([0, 1, 7], np.int32)
    assert lib.zdr_scene_add_mesh(h, np_ptr(verts, C.c_float), 3, np_ptr(bad, C.c_int32), 1, None, None) == -1     # index out of range
    assert b"out of range" in lib.zdr_last_error()
    assert lib.zdr_scene_create(C.byref(C.c_void_p()), 999) == -1                         # bad device
    lib.zdr_scene_destroy(h)
    # non-finite positions are refused at commit by both ingest paths (host flatten / GPU flatten); a failed commit
    # leaves the scene unusable but alive, and a later good commit works
    for n_tris in (4, 20000):
        h = C.c_void_p()
        _abi.check(lib.zdr_scene_create(C.byref(h), 0))
        rng = np.random.default_rng(n_tris)
        verts = np.zeros((3 * n_tris, 8), np.float32)
        verts[:, 0:3] = rng.normal(size=(3 * n_tris, 3))
        verts[:, 5:8] = (0, 0, 1)
        tris = np.arange(3 * n_tris, dtype=np.int32)
        poisoned = verts.copy()
        poisoned[7, 1] = np.inf if n_tris == 4 else np.nan
        _abi.check(lib.zdr_scene_add_mesh(h, np_ptr(poisoned, C.c_float), len(poisoned), np_ptr(tris, C.c_int32), n_tris, None, None))
        assert lib.zdr_scene_commit(h) == -1 and b"non-finite" in lib.zdr_last_error()
        assert lib.zdr_render_forward(h, 1, C.byref(p), 1, 1, None) == -2
        lib.zdr_scene_destroy(h)
        h = C.c_void_p()
        _abi.check(lib.zdr_scene_create(C.byref(h), 0))
        _abi.check(lib.zdr_scene_add_mesh(h, np_ptr(verts, C.c_float), len(verts), np_ptr(tris, C.c_int32), n_tris, None, None))
        _abi.check(lib.zdr_scene_commit(h))
        lib.zdr_scene_destroy(h)
    # sampler tables whose index range does not fit the kernels' 32-bit arithmetic
    tiny16, tiny32 = np.zeros(4, np.uint16), np.zeros(4, np.uint32)
    assert lib.zdr_set_sampler_tables(0, np_ptr(tiny32, C.c_uint32), 5, 65536, np_ptr(tiny16, C.c_uint16), 48, 1 << 14) == -4


@pytest.mark.parametrize("integ,kw", [(ZDR_COLLOCATED, {}), (ZDR_DIRECT, {}), (ZDR_DIRECT, {"scatter": "aggregate"}),
                                      (ZDR_PATH, {"max_depth": 5, "schedule": "megakernel"}), (ZDR_PATH, {"max_depth": 5, "schedule": "wavefront"}),
                                      (ZDR_PATH, {"max_depth": 5, "schedule": "persistent"}),
                                      (ZDR_PATH, {"max_depth": 5, "schedule": "persistent", "prb_mode": _abi.ZDR_PRB_LITERAL})])
def test_material_sets_match_oracle(oracle, integ, kw):
    """Several texture sets ("polymorphic materials", README.md:5; pinned against the patched reference source in
    tests/test_golden.py): Cornell box on set 0, a sphere inside it on set 1, a quad without a set.  Every integrator and
    schedule: image bit-identical to the oracle's, gradients of both sets within the tolerance; the committed fixture too."""
    models, sets = scenes.cbox_material_sets_models()
    g, o = GpuScene(models), oracle.OracleScene(models)
    g.set_material_sets(sets)
    o.set_material_sets(sets)
    mat = np.stack([scenes.random_material((48, 48), 12), scenes.random_material((48, 48), 13)])
    res, spp = (72, 56), 4
    pf = make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, 4, **kw)
    pb = make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, 5, **kw)
    assert pf.num_material_sets == 2 and (pf.tex_res_x, pf.tex_res_y) == (48, 48)
    assert np.array_equal(bits(g.forward(integ, pf, mat)), bits(o.render_forward(integ, pf, mat)))
    dimg = np.random.default_rng(3).uniform(0.5, 1.5, (res[1], res[0], 4)).astype(np.float32)
    dg, do = g.backward(integ, pb, mat, dimg), o.render_backward(integ, pb, mat, dimg)
    assert dg.shape == mat.shape and np.abs(do[0]).max() > 0 and np.abs(do[1]).max() > 0
    for k in range(2):
        assert rel_l2(dg[k], do[k]) <= GRAD_TOL, (k, rel_l2(dg[k], do[k]))
    # a set index beyond the buffer is an error, not an out-of-bounds read
    one = make_params(scenes.CBOX_CAMERA, res, mat[0].shape, spp, 4, **kw)
    with pytest.raises(_abi.ZdrError):
        g.forward(integ, one, mat[0])
    if integ == ZDR_DIRECT and not kw:
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_direct_material_sets.npz"))
        r2 = tuple(int(x) for x in z["res"])
        img = g.forward(ZDR_DIRECT, make_params(scenes.CBOX_CAMERA, r2, z["material"].shape, int(z["spp"]), int(z["seed"])), z["material"])
        assert rel_l2(img, z["image"]) <= 1e-4
        dm = g.backward(ZDR_DIRECT, make_params(scenes.CBOX_CAMERA, r2, z["material"].shape, int(z["spp"]), int(z["seed"]) + 1), z["material"], z["d_image"])
        assert rel_l2(dm, z["d_material"]) <= 1e-3
    g.close()
