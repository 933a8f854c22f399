#!/usr/bin/env python
"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCES (/root/reference/*.py,
imported unmodified) under the LuisaCompute emulator oracle/lc_emu.py.

Run in the build container (needs /root/reference):   python tests/golden/gen_golden.py
The fixtures travel with the repo; the GPU box never reads /root/reference.

Tables: the reference's pbrt-v4 table modules are missing blobs, so the reference runs on the
same surrogate tables the framework ships (zdr_b200/tables.py); the fixture therefore pins the
sampler ALGORITHM (hashing, permutation, index formulae, draw order), not the table contents.

What is recorded
  sampler.npz     pmj02bn.py draws for a grid of (pixel, seed, spp, sample index, draw program)
  functions.npz   camera.py / onb.py / microfacet.py / light.py device functions on seeded inputs,
                  including the LuisaCompute-autodiff gradient of ggx_brdf (emulated tape AD)
  envmap.npz      envmap.py: weight-map kernel cells, MIS compensation + alias tables + pdf (host pipeline), sample_envmap,
                  env_sampled_light_pdf, uv <-> direction, create_alias_table
  render_*.npz    integrator.py kernels built from collocated.py / direct.py on the Cornell box:
                  image, and d_material for a fixed d_image, at fixture sizes
  render_*_large.npz   NOT in the default list (2-15 minutes each): the same kernels and prb.py on 48^2-64^2 pixels at 4-8 spp with
                  lc_emu.CONTRACT_BUILTINS (closed builtins in the numeric contract's fmaf arithmetic) — images bit-identical to
                  the oracle's.  Names: collocated_large direct_large direct_xform_large material_sets_large path_large
                  path_large_smooth direct_env_large path_env_large, and prb.py with the one-line patches that make up
                  prb_mode = corrected: path_corrected_large path_corrected_env_large path_terminal_large
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import lc_emu as E  # noqa: E402
from tests import scenes  # noqa: E402
from zdr_b200.load_obj import read_obj, concat_triangles, load_mesh  # noqa: E402
from zdr_b200.tables import get_tables  # noqa: E402

pmj, bn, prov = get_tables()
assert prov == "surrogate"
ref = E.install(pmj, bn, "/root/reference", "zdr_ref")
import zdr_ref.camera as R_camera  # noqa: E402
import zdr_ref.collocated as R_collocated  # noqa: E402
import zdr_ref.direct as R_direct  # noqa: E402
import zdr_ref.light as R_light  # noqa: E402
import zdr_ref.microfacet as R_mf  # noqa: E402
import zdr_ref.onb as R_onb  # noqa: E402
import zdr_ref.pmj02bn as R_pmj  # noqa: E402
import zdr_ref.render as R_render  # noqa: E402
from zdr_ref.vertex import Vertex  # noqa: E402


def fv(v):
    return [float(c.v) for c in v.c] if isinstance(v, E.Vec) else float(v.v)


# ------------------------------------------------------------------------------------ sampler
def gen_sampler():
    rng = np.random.default_rng(0)
    progs = ["2", "21", "2112", "2112121", "2" + "1121211" * 9]        # up to dimension 65 (past the OOB-wrap region)
    rows, outs, prog_ids = [], [], []
    for pi, prog in enumerate(progs):
        for _ in range(40):
            spp = int(rng.choice([1, 2, 4, 16, 64, 256, 1024]))
            st = [int(rng.integers(0, 2048)), int(rng.integers(0, 2048)),
                  int(rng.choice([0, 1, 12345, 853402567, 19260817, 948377263, 2147483647])), spp, int(rng.integers(0, spp))]
            s = R_pmj.make_pmj02bn_sampler(E.int2(st[0], st[1]), st[2], st[3], st[4])
            vals = []
            for ch in prog:
                if ch == "2":
                    vals.extend(fv(s.next2f()))
                else:
                    vals.append(fv(s.next()))
            rows.append(st); outs.append(np.asarray(vals, np.float32)); prog_ids.append(pi)
    # test_sampler.py:23-30: pixel (24,345), seed 0, n=32, second next2f()
    ts = []
    for i in range(32):
        s = R_pmj.make_pmj02bn_sampler(E.int2(24, 345), 0, 32, i)
        s.next2f()
        ts.append(fv(s.next2f()))
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), progs=np.asarray(progs), state=np.asarray(rows, np.int32),
                        prog_id=np.asarray(prog_ids, np.int32), values=np.asarray(outs, dtype=object),
                        test_sampler_py=np.asarray(ts, np.float32), allow_pickle=True)
    print("sampler.npz:", len(rows), "states")


# ------------------------------------------------------------------------------------ functions
def gen_functions():
    rng = np.random.default_rng(1)
    n = 200

    def hemi():
        v = rng.normal(size=3)
        v[2] = abs(v[2]) + 1e-3
        return (v / np.linalg.norm(v)).astype(np.float32)

    ggx_in, ggx_out, pdf_out = [], [], []
    for _ in range(n):
        wo, wi = hemi(), hemi()
        d, r, g = rng.uniform(0, 1, 3).astype(np.float32), np.float32(rng.uniform(0.03, 1)), rng.normal(size=3).astype(np.float32)
        mat = E.float4(*d, r)
        with E.autodiff():                                   # the reference's AD pattern, collocated.py:44-56
            E.requires_grad(mat)
            f = R_mf.ggx_brdf(E.float3(*wo), E.float3(*wi), mat.xyz, 0.04, mat.w)
            E.backward(f, E.float3(*g))
            dm = E.grad(mat)
        ggx_in.append(np.concatenate([wo, wi, d, [r], g]))
        ggx_out.append(fv(f) + fv(dm))
        pdf_out.append(fv(R_mf.ggx_sample_pdf(E.float3(*wo), E.float3(*wi), mat.xyz, 0.04, mat.w)))
    wm_in, wm_out, cos_in, cos_out = [], [], [], []
    for _ in range(n):
        w, a, u = hemi(), np.float32(rng.uniform(0.001, 1)), rng.uniform(0, 1, 2).astype(np.float32)
        wm_in.append(np.concatenate([w, [a], u]))
        wm_out.append(fv(R_mf.sample_wm(E.float3(*w), E.F(a), E.float2(*u))))
        cos_in.append(u)
        cos_out.append(fv(R_mf.cosine_sample_hemisphere(E.float2(*u))))
    onb_in = [hemi() * rng.choice([-1, 1]) for _ in range(n)]
    onb_out = []
    for nn in onb_in:
        o = R_onb.make_onb(E.float3(*nn))
        onb_out.append(fv(o.tangent) + fv(o.binormal) + fv(o.normal))
    tent_in = rng.uniform(0, 1, n).astype(np.float32)
    tent_out = [fv(R_camera.tent_warp(E.F(u), 1.0)) for u in tent_in]
    tent2 = fv(R_camera.tent_warp(E.float2(0.2, 0.8), 1.0))           # component-wise select on a vector
    tri_in = rng.uniform(0, 1, (n, 2)).astype(np.float32)
    tri_out = [fv(R_light.sample_uniform_triangle(E.float2(*u))) for u in tri_in]
    cam = R_render.Camera(fov=50 / 180 * 3.1415926, origin=E.float3(-0.2, 2.6, 6.0), target=E.float3(-0.2, 2.6, -2.5),
                          up=E.float3(0.0, 1.0, 0.0))
    ndc = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    rays = []
    for p in ndc:
        r = R_camera.generate_ray(cam, E.float2(*p))
        rays.append(fv(r.get_origin()) + fv(r.get_dir()))
    np.savez_compressed(os.path.join(OUT, "functions.npz"), ggx_in=np.asarray(ggx_in, np.float32), ggx_out=np.asarray(ggx_out, np.float32),
                        ggx_pdf=np.asarray(pdf_out, np.float32), wm_in=np.asarray(wm_in, np.float32), wm_out=np.asarray(wm_out, np.float32),
                        cos_in=np.asarray(cos_in, np.float32), cos_out=np.asarray(cos_out, np.float32), onb_in=np.asarray(onb_in, np.float32),
                        onb_out=np.asarray(onb_out, np.float32), tent_in=tent_in, tent_out=np.asarray(tent_out, np.float32),
                        tent_vec=np.asarray(tent2, np.float32), tri_in=tri_in, tri_out=np.asarray(tri_out, np.float32), ndc=ndc,
                        rays=np.asarray(rays, np.float32))
    print("functions.npz")


# ------------------------------------------------------------------------------------ kernels
class RefScene:
    """Scene assembly exactly as render.py:73-128 does it, on the emulated heap / accel."""

    def __init__(self, models):
        self.accel, self.heap = E.Accel(), E.BindlessArray()
        inst_metadata, light_insts, inst_trig_count = [], [], []
        for idx, (obj_file, transform, emission) in enumerate(models):
            em = E.float3(*[float(c) for c in emission]) if emission is not None else E.float3(0.0)
            if em.x > 0 or em.y > 0 or em.z > 0:
                light_insts.append(idx)
            inst_metadata.append(em)
            verts, tris = load_mesh(obj_file)                 # same (v,t,n) rows / fan order as load_obj.py (tests/test_host.py)
            vb = E.Buffer(dtype=Vertex, size=len(verts)); vb.copy_from_array(verts)
            tb = E.buffer([int(t) for t in tris])
            inst_trig_count.append(len(tris) // 3)
            xf = E.float4x4(*transform.flat()) if transform is not None else E.float4x4(1.0)
            self.accel.add(vb, tb, transform=xf)
            self.heap.emplace(idx * 2 + 0, tb)
            self.heap.emplace(idx * 2 + 1, vb)
        self.inst_count = len(models)
        self.light_count = len(light_insts)
        self.heap.emplace(23333, E.buffer(inst_metadata))
        if self.light_count > 0:
            self.heap.emplace(23334, E.buffer(light_insts + [0] * (self.inst_count - self.light_count)))
        self.heap.emplace(23335, E.buffer(inst_trig_count))
        self.accel.update()


def load_envmap_on(heap, img, cells=()):
    """envmap.py:load_envmap on the emulated heap.  Its weight-map kernel (38 M texture fetches at 512 x 256) runs only
    for `cells`; the host pipeline (MIS compensation, alias tables, pdf) then gets the full map from the oracle, whose
    cells gen_envmap() checks against the kernel.  Returns the kernel's values at `cells`."""
    import zdr_ref.envmap as R_env
    from oracle import oracle as o
    o.build()
    osc = o.OracleScene(scenes.cbox_collocated_models())
    osc.set_envmap(img, True)
    full = osc.env_scale_map()
    cellset, captured = set(cells), {}
    orig_numpy = E.Buffer.numpy

    def numpy_hook(self):                      # envmap.py:160: scale_map = scale_map_buffer.numpy()
        captured["cells"] = np.asarray([self.data[y * 512 + x] for x, y in cells], np.float32)
        return full.reshape(-1).copy()
    E.DISPATCH_FILTER = lambda x, y, z: (x, y) in cellset
    E.Buffer.numpy = numpy_hook
    try:
        R_env.load_envmap(heap, img, compensate_mis=True)
    finally:
        E.Buffer.numpy = orig_numpy
        E.DISPATCH_FILTER = None
    return captured["cells"]


def gen_render(name, models, fwd_kernel, bwd_kernel, res, spp, tex, seed, mat_seed, rough_min=0.0, envmap=None):
    sc = RefScene(models)
    env_count = 0
    if envmap is not None:                     # render.py:150-156
        load_envmap_on(sc.heap, envmap)
        env_count = 1
    cam = R_render.Camera(fov=50 / 180 * 3.1415926, origin=E.float3(-0.2, 2.6, 6.0), target=E.float3(-0.2, 2.6, -2.5),
                          up=E.float3(0.0, 1.0, 0.0))
    mat = scenes.random_material((tex, tex), mat_seed)
    # GGX's D term near the specular peak of a very smooth texel is ill-conditioned in binary32 (t = nh^2(a2-1)+1
    # cancels): two correct float32 implementations differ by percents on such a sample, so the path fixture
    # keeps roughness >= rough_min; the GPU-vs-oracle tests (identical arithmetic) cover the full range.
    mat[..., 3] = np.float32(rough_min) + np.float32(1.0 - rough_min) * mat[..., 3]
    W, H = res
    image = E.Buffer(dtype=E.float4, size=W * H, data=np.zeros((W * H, 4), np.float32))
    mbuf = E.Buffer(dtype=float, size=mat.size, data=mat.reshape(-1).copy())
    # render.py:168-171
    fwd_kernel(image, sc.heap, sc.accel, sc.light_count, env_count, mbuf, E.int2(tex, tex), cam, spp, seed, True, dispatch_size=(W, H))
    rng = np.random.default_rng(7)
    d_image = rng.normal(size=(H, W, 4)).astype(np.float32)
    dbuf = E.Buffer(dtype=E.float4, size=W * H, data=d_image.reshape(-1, 4).copy())
    dmat = E.Buffer(dtype=float, size=mat.size, data=np.zeros(mat.size, np.float64))
    # render.py:193-196 (seed+1)
    bwd_kernel(dbuf, sc.heap, sc.accel, sc.light_count, env_count, dmat, mbuf, E.int2(tex, tex), cam, spp, seed + 1, True, dispatch_size=(W, H))
    extra = {} if envmap is None else {"envmap": envmap}
    np.savez_compressed(os.path.join(OUT, f"render_{name}.npz"), res=np.asarray(res), spp=spp, tex=tex, seed=seed, mat_seed=mat_seed, rough_min=rough_min,
                        image=image.data.reshape(H, W, 4).astype(np.float32), d_image=d_image, **extra,
                        d_material=dmat.data.reshape(tex, tex, 4).astype(np.float32))
    print(f"render_{name}.npz: image mean {image.data[:, :3].mean():.5f}, |d_material| sum {np.abs(dmat.data).sum():.4f}")


def gen_loadobj():
    """load_obj.py:1-68 on every asset: vertex de-dup order and fan triangulation fix the primitive ids."""
    import zdr_ref.load_obj as R_obj
    out = {}
    for name in ("cboxuv.obj", "cbox-combined.obj", "cbox-light.obj", "quad.obj", "sphere.obj"):
        vertices, faces = R_obj.read_obj(scenes.asset(name))
        out[name + ":verts"] = np.array([(*v, *t, *n) for v, t, n in vertices], dtype=np.float32)     # render.py:97
        out[name + ":tris"] = np.array(R_obj.concat_triangles(faces), dtype=np.int32)
        if np.isnan(out[name + ":verts"][0, 5]):
            # render.py:95-103: meshes without vn get recompute_normal.py's three kernels (clear, scatter the face normal
            # to the three corners with float atomics, normalise) — run here in program order
            import zdr_ref.recompute_normal as R_rn
            from zdr_ref.vertex import Vertex as V
            vb = E.Buffer(dtype=V, size=len(vertices))
            vb.copy_from_array(out[name + ":verts"].copy())
            tb = E.buffer([int(t) for t in out[name + ":tris"]])
            with np.errstate(invalid="ignore", divide="ignore"):
                R_rn.recompute_normal(vb, tb)
            out[name + ":verts_recomputed"] = np.asarray(vb.data, np.float32)
    np.savez_compressed(os.path.join(OUT, "loadobj.npz"), **out)
    print("loadobj.npz")


def gen_uvgrad():
    """uvgrad.py:19-49 `trace_duvdxy` (ray differentials -> screen-to-texture Jacobian) on camera rays of the
    Cornell box.  The kernel around it (uvgrad.py:75-98) draws its pixel offsets from LuisaCompute's built-in LCG
    sampler, which is not available to pin, so the rays are supplied directly."""
    import zdr_ref.uvgrad as R_uv
    sc = RefScene(scenes.cbox_collocated_models())
    cam = R_render.Camera(fov=50 / 180 * 3.1415926, origin=E.float3(-0.2, 2.6, 6.0), target=E.float3(-0.2, 2.6, -2.5),
                          up=E.float3(0.0, 1.0, 0.0))
    rng = np.random.default_rng(12)
    rays, outs = [], []
    res = 32.0
    for _ in range(150):
        c = rng.uniform(2, 30, 2)
        px = [E.float2(*(2.0 / res * (c + np.asarray(o)) - 1.0)) for o in ((0, 0), (1, 0), (0, 1))]
        rr = [R_camera.generate_ray(cam, q) for q in px]
        g = R_uv.trace_duvdxy(rr[0], rr[1], rr[2], sc.heap, sc.accel)
        rays.append(sum([fv(r.get_origin()) + fv(r.get_dir()) for r in rr], []))
        outs.append(fv(g))
    np.savez_compressed(os.path.join(OUT, "uvgrad.npz"), rays=np.asarray(rays, np.float32), duvdxy=np.asarray(outs, np.float32))
    print("uvgrad.npz: mean |duvdxy|", np.abs(np.asarray(outs)).mean(axis=0))


# prb.py does not compile as shipped: its call sites disagree with the callee signatures in
# interaction.py:9, light.py:23,84 and integrator.py:51-52 (SURVEY.md 8a "PRB defects" item 1).  The
# textual substitutions below are exactly that list — nothing else is changed — and are applied to an
# in-memory copy of the source; the result is executed as module zdr_ref.prb_patched.
PRB_PATCHES = [
    ("surface_interact(hit, heap)", "surface_interact(hit, heap, accel)"),
    ("sample_light_pdf(ray.get_origin(), light_count, env_count, heap, hit.inst, hit.prim, it.p)",
     "sample_light_pdf(ray.get_origin(), light_count, heap, accel, hit.inst, hit.prim, it.p)"),
    ("sample_light_pdf(ray.get_origin(), light_count, heap, hit.inst, hit.prim, it.p)",
     "sample_light_pdf(ray.get_origin(), light_count, heap, accel, hit.inst, hit.prim, it.p)"),
    ("sample_light(it.p, light_count, env_count, heap, sampler)", "sample_light(it.p, light_count, env_count, heap, accel, sampler)"),
    ("sample_light(it.p, light_count, heap, sampler)", "sample_light(it.p, light_count, env_count, heap, accel, sampler)"),
    ("def path_estimator_backward(ray, sampler, heap, accel, light_count,\n", "def path_estimator_backward(ray, sampler, heap, accel, light_count, env_count,\n"),
    ("path_estimator(ray_twin, sampler_twin, heap, accel, light_count, material_buffer, texture_res)",
     "path_estimator(ray_twin, sampler_twin, heap, accel, light_count, env_count, material_buffer, texture_res)"),
]


# The ONE further line that turns prb.py into prb_mode = "corrected" on a scene without envmap (DESIGN.md §2 "PRB adjoint weighting"):
# the cotangent of the BSDF-sample vertex.  prb.py:162 seeds it with beta / pdf * Le; Path Replay Backpropagation needs
# d(Le_remaining) / d(bsdf) = Le_remaining / bsdf per channel (0 where bsdf is 0).  Nothing else of the adjoint walk changes.
PRB_CORRECTED_PATCH = [
    ("backward(bsdf, beta / pdf_bsdf * Le * le_grad)",
     "backward(bsdf, float3(0.0) if bsdf == float3(0.0) else Le / bsdf * le_grad)"),
]


# ... and on a scene WITH an envmap one more: prb.py:28 multiplies the envmap radiance of a miss by the throughput and :31 multiplies
# by it again (SURVEY.md 8a defect 2); corrected mode applies it once.
PRB_CORRECTED_ENV_PATCH = [
    ("emission = beta * heap.texture2d_sample(23332, direction_to_uv(ray.get_dir())).xyz",
     "emission = heap.texture2d_sample(23332, direction_to_uv(ray.get_dir())).xyz"),
]


# ... and on a scene with NON-EMISSIVE instances other than the first: the forward walk stops at them (prb.py:45 "treat everything as
# light other than the first object"), the adjoint walk lacks that line and shades them with the first object's texture
# (SURVEY.md 8a defect 2; literal mode reproduces it); corrected mode stops where the forward walk stopped.
PRB_CORRECTED_TERMINAL_PATCH = [
    ("            Le -= beta * mis_weight * emission\n            break\n",
     "            Le -= beta * mis_weight * emission\n            break\n        if hit.inst > 0:\n            break\n"),
]


def load_patched_prb(extra=(), name="prb_patched"):
    import importlib.util
    import tempfile
    src = open("/root/reference/prb.py").read()
    for a, b in list(PRB_PATCHES) + list(extra):
        assert a in src, a
        src = src.replace(a, b)
    d = tempfile.mkdtemp()
    path = os.path.join(d, f"{name}.py")
    open(path, "w").write(src)
    spec = importlib.util.spec_from_file_location(f"zdr_ref.{name}", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"zdr_ref.{name}"] = mod
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------ material sets
# "Polymorphic materials" (several texture sets) are listed as NOT implemented upstream (README.md:5), so the reference cannot
# pin them as shipped.  The fixture below runs the reference's own direct.py / interaction.py with the smallest textual change
# that adds a set index — nothing else is touched:
#   interaction.py: read_single_bsdf / read_bsdf / write_single_bsdf_grad / write_bsdf_grad take `mat_set` and offset the texel
#                   index by mat_set * texture_res.x * texture_res.y;
#   direct.py:      `mat_set = heap.buffer_read(int, 23336, hit.inst)` next to the emission read, `if hit.inst > 0` becomes
#                   `if mat_set < 0` (direct.py:30,98: "treat everything as light other than the first object" generalised to
#                   "... other than the textured objects"), and every read_bsdf / write_bsdf_grad passes mat_set (the set of
#                   the PRIMARY hit, as mat_grad_uv is the primary hit's uv).
MS_INTERACTION_PATCHES = [
    ("def read_single_bsdf(coord: int2, material_buffer, texture_res):", "def read_single_bsdf(coord: int2, material_buffer, texture_res, mat_set):"),
    ("    idx = coord.x + texture_res.x * coord.y\n    return float4(", "    idx = coord.x + texture_res.x * coord.y + mat_set * texture_res.x * texture_res.y\n    return float4("),
    ("def read_bsdf(uv: float2, material_buffer, texture_res):", "def read_bsdf(uv: float2, material_buffer, texture_res, mat_set):"),
    ("return read_single_bsdf(nearest, material_buffer, texture_res)", "return read_single_bsdf(nearest, material_buffer, texture_res, mat_set)"),
    ("read_single_bsdf(ip + int2(0,0), material_buffer, texture_res)", "read_single_bsdf(ip + int2(0,0), material_buffer, texture_res, mat_set)"),
    ("read_single_bsdf(ip + int2(0,1), material_buffer, texture_res)", "read_single_bsdf(ip + int2(0,1), material_buffer, texture_res, mat_set)"),
    ("read_single_bsdf(ip + int2(1,0), material_buffer, texture_res)", "read_single_bsdf(ip + int2(1,0), material_buffer, texture_res, mat_set)"),
    ("read_single_bsdf(ip + int2(1,1), material_buffer, texture_res)", "read_single_bsdf(ip + int2(1,1), material_buffer, texture_res, mat_set)"),
    ("def write_single_bsdf_grad(coord: int2, dmat, d_material_buffer, texture_res):", "def write_single_bsdf_grad(coord: int2, dmat, d_material_buffer, texture_res, mat_set):"),
    ("    idx = coord.x + texture_res.x * coord.y\n    _ = d_material_buffer", "    idx = coord.x + texture_res.x * coord.y + mat_set * texture_res.x * texture_res.y\n    _ = d_material_buffer"),
    ("def write_bsdf_grad(uv: float2, dmat, d_material_buffer, texture_res):", "def write_bsdf_grad(uv: float2, dmat, d_material_buffer, texture_res, mat_set):"),
    ("write_single_bsdf_grad(nearest, dmat, d_material_buffer, texture_res)", "write_single_bsdf_grad(nearest, dmat, d_material_buffer, texture_res, mat_set)"),
    ("k00 * dmat, d_material_buffer, texture_res)", "k00 * dmat, d_material_buffer, texture_res, mat_set)"),
    ("k01 * dmat, d_material_buffer, texture_res)", "k01 * dmat, d_material_buffer, texture_res, mat_set)"),
    ("k10 * dmat, d_material_buffer, texture_res)", "k10 * dmat, d_material_buffer, texture_res, mat_set)"),
    ("k11 * dmat, d_material_buffer, texture_res)", "k11 * dmat, d_material_buffer, texture_res, mat_set)"),
]
MS_DIRECT_PATCHES = [      # (old, new, occurrences)
    ("from .interaction import read_bsdf, write_bsdf_grad, surface_interact", "from .interaction_ms import read_bsdf, write_bsdf_grad\nfrom .interaction import surface_interact", 1),
    ("    emission = heap.buffer_read(float3, 23333, hit.inst)\n    if hit.inst > 0: # treat everything as light other than the first object",
     "    emission = heap.buffer_read(float3, 23333, hit.inst)\n    mat_set = heap.buffer_read(int, 23336, hit.inst)\n    if mat_set < 0: # treat everything as light other than the textured objects", 2),
    ("read_bsdf(it.uv, material_buffer, texture_res)", "read_bsdf(it.uv, material_buffer, texture_res, mat_set)", 2),
    ("write_bsdf_grad(it.uv, mat_grad, d_material_buffer, texture_res)", "write_bsdf_grad(it.uv, mat_grad, d_material_buffer, texture_res, mat_set)", 2),
    ("write_bsdf_grad(mat_grad_uv, mat_grad, d_material_buffer, texture_res)", "write_bsdf_grad(mat_grad_uv, mat_grad, d_material_buffer, texture_res, mat_set)", 1),
]


def _load_patched(module, patches, new_name):
    import importlib.util
    import tempfile
    src = open(f"/root/reference/{module}.py").read()
    for pt in patches:
        a, b, n = pt if len(pt) == 3 else (pt[0], pt[1], 1)
        assert src.count(a) == n, (module, a, src.count(a))
        src = src.replace(a, b)
    path = os.path.join(tempfile.mkdtemp(), new_name + ".py")
    open(path, "w").write(src)
    spec = importlib.util.spec_from_file_location("zdr_ref." + new_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["zdr_ref." + new_name] = mod
    spec.loader.exec_module(mod)
    return mod


def gen_material_sets(name="direct_material_sets", res=(24, 24), spp=2, tex=32, seed=21):
    """Cornell box (set 0) + light + a textured sphere inside the box (set 1) + a second, untextured quad (set -1): direct.py
    with the material-set patch above, forward and adjoint."""
    _load_patched("interaction", MS_INTERACTION_PATCHES, "interaction_ms")
    dms = _load_patched("direct", MS_DIRECT_PATCHES, "direct_ms")
    models, sets = scenes.cbox_material_sets_models()
    sc = RefScene(models)
    sc.heap.emplace(23336, E.buffer([int(x) for x in sets]))
    cam = R_render.Camera(fov=50 / 180 * 3.1415926, origin=E.float3(-0.2, 2.6, 6.0), target=E.float3(-0.2, 2.6, -2.5),
                          up=E.float3(0.0, 1.0, 0.0))
    W, H = res
    mat = np.stack([scenes.random_material((tex, tex), 12), scenes.random_material((tex, tex), 13)])
    image = E.Buffer(dtype=E.float4, size=W * H, data=np.zeros((W * H, 4), np.float32))
    mbuf = E.Buffer(dtype=float, size=mat.size, data=mat.reshape(-1).copy())
    dms.render_direct_kernel(image, sc.heap, sc.accel, sc.light_count, 0, mbuf, E.int2(tex, tex), cam, spp, seed, True, dispatch_size=(W, H))
    d_image = np.random.default_rng(7).normal(size=(H, W, 4)).astype(np.float32)
    dbuf = E.Buffer(dtype=E.float4, size=W * H, data=d_image.reshape(-1, 4).copy())
    dmat = E.Buffer(dtype=float, size=mat.size, data=np.zeros(mat.size, np.float64))
    dms.render_direct_backward_kernel(dbuf, sc.heap, sc.accel, sc.light_count, 0, dmat, mbuf, E.int2(tex, tex), cam, spp, seed + 1, True, dispatch_size=(W, H))
    dm = dmat.data.reshape(2, tex, tex, 4).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"render_{name}.npz"), res=np.asarray(res), spp=spp, tex=tex, seed=seed, sets=np.asarray(sets, np.int32),
                        material=mat, image=image.data.reshape(H, W, 4).astype(np.float32), d_image=d_image, d_material=dm)
    print(f"render_{name}.npz: image mean {image.data[:, :3].mean():.5f}, |d_material| per set {np.abs(dm).sum(axis=(1, 2, 3))}")


# ------------------------------------------------------------------------------------ envmap.py
ENV_ROWS = (0, 17, 128, 200, 255)            # conditional tables / pdf rows kept in the fixture


def gen_envmap():
    """envmap.py executed from the reference source: the weight-map kernel on a handful of cells, the host pipeline
    (MIS compensation, conditional + marginal alias tables, pdf normalisation) on the FULL weight map, the device
    functions sample_envmap / env_sampled_light_pdf / uv_to_direction / direction_to_uv on the resulting heap, and
    create_alias_table on small vectors.  The 512 x 256 kernel is 38 M emulated texture fetches, so it runs for 28
    cells only; the host pipeline then gets the full map from the oracle (whose cells the fixture checks)."""
    import zdr_ref.envmap as R_env
    img = scenes.procedural_envmap(64, 32, seed=1)
    rng = np.random.default_rng(7)
    cells = sorted({(int(x), int(y)) for x, y in zip(rng.integers(0, 512, 24), rng.integers(0, 256, 24))}
                   | {(0, 0), (511, 255), (0, 255), (511, 0)})
    heap = E.BindlessArray()
    captured = {"cells": load_envmap_on(heap, img, cells)}
    aliases, pdfs = heap.slots[23330].data, np.asarray(heap.slots[23331].data, np.float32)
    prob = np.asarray([float(F_(a.prob)) for a in aliases], np.float32)
    alias = np.asarray([int(a.alias) for a in aliases], np.int32)
    keep = np.concatenate([np.arange(256)] + [256 + r * 512 + np.arange(512) for r in ENV_ROWS])
    pdf_keep = np.concatenate([r * 512 + np.arange(512) for r in ENV_ROWS])
    # device functions on the reference heap
    u = rng.random((96, 2)).astype(np.float32)
    smp = []
    for a, b in u:
        t = R_env.sample_envmap(heap, E.float2(float(a), float(b)))
        smp.append(fv(t.wi) + [float(F_(t.pdf))] + fv(t.eval) + [float(F_(t.dist))])
    d = rng.normal(size=(96, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    pdf_dir = [float(F_(R_env.env_sampled_light_pdf(heap, E.float3(*[float(c) for c in q])))) for q in d]
    uv = rng.random((64, 2)).astype(np.float32)
    dd = rng.normal(size=(64, 3))
    dd = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    mapping = [fv(R_env.uv_to_direction(E.float2(float(a), float(b)))) + fv(R_env.direction_to_uv(E.float3(*[float(c) for c in q])))
               for (a, b), q in zip(uv, dd)]
    # create_alias_table on plain Python floats (binary64 throughout, as the oracle's)
    at_in, at_prob, at_alias, at_pdf = [], [], [], []
    for n in (7, 16, 33):
        v = rng.random(n)
        v[rng.integers(0, n, 2)] = 0.0
        table, pdf = R_env.create_alias_table([float(x) for x in v])
        at_in.append(v); at_prob.append([float(F_(e.prob)) for e in table]); at_alias.append([int(e.alias) for e in table]); at_pdf.append(pdf)
    np.savez_compressed(os.path.join(OUT, "envmap.npz"), img=img, cells=np.asarray(cells, np.int32), cell_weights=captured["cells"],
                        rows=np.asarray(ENV_ROWS), prob=prob[keep], alias=alias[keep], pdf=pdfs[pdf_keep],
                        pdf_sum=np.float64(pdfs.astype(np.float64).sum()), u=u, samples=np.asarray(smp, np.float32), dirs=d,
                        pdf_dir=np.asarray(pdf_dir, np.float32), map_uv=uv, map_dir=dd, mapping=np.asarray(mapping, np.float32),
                        at_in=np.concatenate(at_in), at_prob=np.concatenate(at_prob).astype(np.float32),
                        at_alias=np.concatenate(at_alias).astype(np.int32), at_pdf=np.concatenate(at_pdf), at_len=np.asarray([7, 16, 33]))
    print("envmap: cells", len(cells), "tables", prob.shape, "pdf sum", float(pdfs.astype(np.float64).sum()))


# ------------------------------------------------------------------------------------ render.py: Scene
def gen_scene_api():
    """The reference's own Scene class (render.py:31-241) end to end under the emulator: load_geometry, render through
    Scene.RenderOperator (torch.autograd), update_lights, and backward() restoring the forward call's camera and lights.
    There is no GPU here: the two places render.py names the device (torch.empty(..., device='cuda') and
    torch.cuda.synchronize) are redirected to the CPU for the duration of the run."""
    import torch
    models_flat = scenes.cbox_two_light_models()
    models = [(f, None if m is None else E.float4x4(*m.flat()), E.float3(*[float(c) for c in e])) for f, m, e in models_flat]
    orig_empty, orig_zeros, orig_sync = torch.empty, torch.zeros, torch.cuda.synchronize
    torch.empty = lambda *a, **k: orig_empty(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    torch.zeros = lambda *a, **k: orig_zeros(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        sc = R_render.Scene(models, integrator="direct")
        sc.camera = R_render.Camera(fov=50 / 180 * 3.1415926, origin=E.float3(-0.2, 2.6, 6.0), target=E.float3(-0.2, 2.6, -2.5),
                                    up=E.float3(0.0, 1.0, 0.0))
        res, spp = (20, 20), 2
        mat = torch.from_numpy(scenes.random_material((64, 64), 21)).requires_grad_()
        w = torch.from_numpy(np.random.default_rng(5).uniform(0.5, 1.5, (res[1], res[0], 4)).astype(np.float32))
        img1 = sc.render(mat, res=res, spp=spp, seed=5)
        (img1 * w).sum().backward()
        g1 = mat.grad.clone()
        mat.grad = None
        # render.py:130-148: numbers, None and float3 are all accepted
        sc.update_lights([None, 0, E.float3(9.0, 6.0, 3.0)])
        lc2 = sc.light_count
        img2 = sc.render(mat, res=res, spp=spp, seed=6)
        # the state changes between forward and backward: backward must use the forward call's (render.py:203-218)
        sc.camera = R_render.Camera(fov=0.5, origin=E.float3(1.0, 1.0, 1.0), target=E.float3(0.0, 0.0, 0.0), up=E.float3(0.0, 1.0, 0.0))
        sc.update_lights([E.float3(0.0), E.float3(1.0, 1.0, 1.0), E.float3(0.0)])
        (img2 * w).sum().backward()
        g2 = mat.grad.clone()
    finally:
        torch.empty, torch.zeros, torch.cuda.synchronize = orig_empty, orig_zeros, orig_sync
    np.savez_compressed(os.path.join(OUT, "scene_api.npz"), res=np.asarray(res), spp=spp, w=w.numpy(), img1=img1.detach().numpy(),
                        g1=g1.numpy(), img2=img2.detach().numpy(), g2=g2.numpy(), light_count2=lc2, inst_count=sc.inst_count)
    print("scene_api.npz: means", float(img1[..., :3].mean()), float(img2[..., :3].mean()), "|g|", float(g1.abs().sum()), float(g2.abs().sum()))


def F_(x):
    return x.v if isinstance(x, E.F) else x


if __name__ == "__main__":
    which = sys.argv[1:] or ["sampler", "functions", "collocated", "direct", "path", "loadobj", "uvgrad", "envmap", "direct_env", "path_env", "direct_xform", "scene_api", "material_sets"]
    if "path_env" in which:                    # prb.py (arity-patched) with env_count = 1
        prb = load_patched_prb()
        gen_render("path_literal_env", scenes.cbox_path_models(), prb.render_path_kernel, prb.render_path_backward_kernel, (16, 16), 2, 64, 13, 6,
                   rough_min=0.3, envmap=scenes.procedural_envmap(64, 32, seed=2) * np.float32(0.2))
    if "direct_xform" in which:                # non-rigid instance transforms: p = M v, normals by inverse(transpose(M))
        gen_render("direct_xform", scenes.cbox_sheared_models(), R_direct.render_direct_kernel, R_direct.render_direct_backward_kernel,
                   (20, 20), 2, 64, 17, 10)
    if "direct_env" in which:                  # direct.py with env_count = 1: Cornell box (open front) + light + envmap
        gen_render("direct_env", scenes.cbox_path_models(), R_direct.render_direct_kernel, R_direct.render_direct_backward_kernel,
                   (20, 20), 2, 64, 11, 8, envmap=scenes.procedural_envmap(64, 32, seed=2) * np.float32(0.2))
    if "envmap" in which:
        gen_envmap()
    if "scene_api" in which:
        gen_scene_api()
    if "loadobj" in which:
        gen_loadobj()
    if "uvgrad" in which:
        gen_uvgrad()
    if "sampler" in which:
        gen_sampler()
    if "functions" in which:
        gen_functions()
    if "collocated" in which:
        gen_render("collocated", scenes.cbox_collocated_models(), R_collocated.render_collocated_kernel,
                   R_collocated.render_collocated_backward_kernel, (24, 24), 2, 64, 0, 3)
    if "direct" in which:
        gen_render("direct", scenes.cbox_path_models(), R_direct.render_direct_kernel, R_direct.render_direct_backward_kernel,
                   (20, 20), 2, 64, 5, 4)
    if "material_sets" in which:
        gen_material_sets()
    # larger fixtures (64 x 64 pixels, 8 spp, 128^2 texture: 32 768 samples per pass instead of 512-1 152; minutes under the
    # emulator, so they are not in the default list): every bounce depth, Russian-roulette branch and light-sampling case of the
    # reference's direct.py / prb.py occurs thousands of times in them.  Generated with lc_emu.CONTRACT_BUILTINS: the closed
    # builtins (dot, cross, lerp, reflect, hit.interpolate, matrix * vector) and the three vector sums a X + b Y + c Z evaluate
    # in the numeric contract's fmaf arithmetic.  With the plain-arithmetic builtins the same source gives an image 1.7e-3 away
    # (237 of 4 096 pixels hold a 10+-bounce path that an ulp in hit.interpolate sends elsewhere on the white-noise texture);
    # with them, the reference's source reproduces the oracle to 1e-7, path by path.
    LARGE = ("direct_large", "path_large", "collocated_large", "direct_env_large", "path_env_large", "path_large_smooth",
             "direct_xform_large", "material_sets_large", "path_corrected_large", "path_corrected_env_large", "path_terminal_large")
    if any(w in which for w in LARGE):
        E.CONTRACT_BUILTINS = True
        if "path_corrected_env_large" in which:
            prb = load_patched_prb(PRB_CORRECTED_PATCH + PRB_CORRECTED_ENV_PATCH, "prb_corrected_env")
            gen_render("path_corrected_env_large", scenes.cbox_path_models(), prb.render_path_kernel, prb.render_path_backward_kernel, (48, 48), 4, 128, 39, 23,
                       envmap=scenes.procedural_envmap(64, 32, seed=2) * np.float32(0.2))
        if "path_terminal_large" in which:     # Cornell box + light + two non-emissive instances (sphere, dark quad): both weightings
            models = scenes.cbox_material_sets_models()[0]
            prb = load_patched_prb(PRB_CORRECTED_PATCH + PRB_CORRECTED_TERMINAL_PATCH, "prb_corrected_terminal")
            gen_render("path_corrected_terminal_large", models, prb.render_path_kernel, prb.render_path_backward_kernel, (48, 48), 4, 128, 41, 24)
            prb = load_patched_prb()
            gen_render("path_literal_terminal_large", models, prb.render_path_kernel, prb.render_path_backward_kernel, (48, 48), 4, 128, 41, 24)
        if "path_corrected_large" in which:    # prb.py + the arity substitutions + PRB_CORRECTED_PATCH: pins prb_mode = corrected
            prb = load_patched_prb(PRB_CORRECTED_PATCH, "prb_corrected")
            gen_render("path_corrected_large", scenes.cbox_path_models(), prb.render_path_kernel, prb.render_path_backward_kernel, (48, 48), 4, 128, 37, 22)
        if "direct_xform_large" in which:
            gen_render("direct_xform_large", scenes.cbox_sheared_models(), R_direct.render_direct_kernel, R_direct.render_direct_backward_kernel,
                       (48, 48), 8, 128, 33, 20)
        if "material_sets_large" in which:
            gen_material_sets("direct_material_sets_large", (48, 48), 8, 64, 35)
        if "collocated_large" in which:
            gen_render("collocated_large", scenes.cbox_collocated_models(), R_collocated.render_collocated_kernel,
                       R_collocated.render_collocated_backward_kernel, (64, 64), 8, 128, 25, 16)
        if "direct_env_large" in which:
            gen_render("direct_env_large", scenes.cbox_path_models(), R_direct.render_direct_kernel, R_direct.render_direct_backward_kernel,
                       (48, 48), 8, 128, 27, 17, envmap=scenes.procedural_envmap(64, 32, seed=2) * np.float32(0.2))
        if "path_env_large" in which:
            prb = load_patched_prb()
            gen_render("path_literal_env_large", scenes.cbox_path_models(), prb.render_path_kernel, prb.render_path_backward_kernel, (48, 48), 4, 128, 29, 18,
                       envmap=scenes.procedural_envmap(64, 32, seed=2) * np.float32(0.2))
        if "path_large_smooth" in which:       # the full roughness range [0, 1] (the plain-arithmetic path fixtures keep roughness >= 0.3)
            prb = load_patched_prb()
            gen_render("path_literal_large_smooth", scenes.cbox_path_models(), prb.render_path_kernel, prb.render_path_backward_kernel, (48, 48), 4, 128, 31, 19)
        if "direct_large" in which:
            gen_render("direct_large", scenes.cbox_path_models(), R_direct.render_direct_kernel, R_direct.render_direct_backward_kernel,
                       (64, 64), 8, 128, 23, 15)
        if "path_large" in which:
            prb = load_patched_prb()
            assert prb.max_depth == 16 and prb.rr_depth == 2
            gen_render("path_literal_large", scenes.cbox_path_models(), prb.render_path_kernel, prb.render_path_backward_kernel, (64, 64), 8, 128, 21, 14,
                       rough_min=0.3)
        E.CONTRACT_BUILTINS = False
    if "path" in which:
        prb = load_patched_prb()
        assert prb.max_depth == 16 and prb.rr_depth == 2
        gen_render("path_literal", scenes.cbox_path_models(), prb.render_path_kernel, prb.render_path_backward_kernel, (16, 16), 2, 64, 9, 6, rough_min=0.3)
