"""GPU tests of the "fast" build of the render kernels (ZDR_FLAG_FAST_MATH / Scene.math = "fast"): approximate division, square
root and contraction in the shading arithmetic, sampler and ray / triangle test unchanged.  The exact build is compared with the
oracle bit for bit elsewhere; here the bars are BASELINE.json's tolerances (images rel-L2 <= 1e-4, gradients <= 1e-3) on a
well-conditioned material — roughness >= 0.3, the restriction under which the oracle itself is pinned to the reference's sources
(tests/golden/gen_golden.py), and a texture that is smooth at the scale of a float32 rounding error of the hit point.  On a
1024^2 white-noise texture a 1e-7 change of a bounce direction already changes the fetched material by ~1e-4, so any two
float32 implementations differ there (measured: tests/fast_parity_report.py, profiles/r02v_*)."""
import numpy as np
import pytest

from zdr_b200 import _abi
from zdr_b200._abi import ZDR_COLLOCATED, ZDR_DIRECT, ZDR_PATH, make_params

from . import scenes
from .test_gpu_parity import GpuScene, bits, rel_l2

pytestmark = pytest.mark.gpu

IMG_TOL, GRAD_TOL = 1e-4, 1e-3


def smooth_material(res=(64, 64), seed=0):
    """Low-frequency texture: diffuse in [0.2, 0.8], roughness in [0.3, 0.9]."""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.linspace(0, 1, res[0], dtype=np.float32), np.linspace(0, 1, res[1], dtype=np.float32), indexing="ij")
    m = np.empty((res[0], res[1], 4), np.float32)
    for c in range(4):
        fx, fy, ph = rng.uniform(0.5, 2.0), rng.uniform(0.5, 2.0), rng.uniform(0, 6.28)
        m[..., c] = 0.5 + 0.3 * np.sin(6.2831853 * (fx * x + fy * y) + ph)
    m[..., 3] = 0.6 + (m[..., 3] - 0.5)
    return m


CASES = [
    ("collocated", scenes.cbox_collocated_models, ZDR_COLLOCATED, (128, 128), 2, {}),
    ("direct", scenes.cbox_path_models, ZDR_DIRECT, (96, 96), 8, {}),
    ("path_literal_wavefront", scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, {"max_depth": 6, "prb_mode": _abi.ZDR_PRB_LITERAL, "schedule": "wavefront"}),
    ("path_literal_persistent", scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, {"max_depth": 6, "prb_mode": _abi.ZDR_PRB_LITERAL, "schedule": "persistent"}),
    ("path_literal_megakernel", scenes.cbox_path_models, ZDR_PATH, (64, 64), 4, {"max_depth": 6, "prb_mode": _abi.ZDR_PRB_LITERAL, "schedule": "megakernel"}),
    ("path_corrected_wavefront", scenes.cbox_path_models, ZDR_PATH, (96, 96), 8, {"max_depth": 6, "schedule": "wavefront"}),
]


@pytest.mark.parametrize("name,models_fn,integ,res,spp,kw", CASES, ids=[c[0] for c in CASES])
def test_fast_build_within_tolerance_of_the_oracle(oracle, name, models_fn, integ, res, spp, kw):
    models, mat = models_fn(), smooth_material()
    g, o = GpuScene(models), oracle.OracleScene(models)
    pf = make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, 0, math="fast", **kw)
    pb = make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, 1, math="fast", **kw)
    img_g, img_o = g.forward(integ, pf, mat), o.render_forward(integ, pf, mat)
    dimg = np.random.default_rng(11).normal(size=(res[1], res[0], 4)).astype(np.float32)
    g.reset_counters(); o.reset_ray_counters()
    dm_g, dm_o = g.backward(integ, pb, mat, dimg), o.render_backward(integ, pb, mat, dimg)
    cg, (oc, oa) = g.counters(), o.ray_counters()
    print(name, "image", rel_l2(img_g, img_o), "grad", rel_l2(dm_g, dm_o), "rays", (cg["closest_rays"], cg["any_rays"]), (oc, oa))
    assert rel_l2(img_g, img_o) <= IMG_TOL
    assert np.isfinite(dm_g).all() and np.abs(dm_o).max() > 0
    assert rel_l2(dm_g, dm_o) <= GRAD_TOL
    # the exact build of the same library, same scene handle: bit-identical to the oracle (the flag only selects the kernels)
    px = make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, 0, **kw)
    assert np.array_equal(bits(g.forward(integ, px, mat)), bits(img_o))
    g.close()


def test_fast_build_same_rays_on_direct_lighting(oracle):
    """Sampler draws and ray / triangle tests are the same explicit IEEE operations in both builds: with the direct integrator (rays
    depend on the camera, the sampler and the light samples only) the fast build traces exactly the oracle's rays."""
    models, mat = scenes.cbox_direct_models(), scenes.random_material((256, 256), 3)
    g, o = GpuScene(models), oracle.OracleScene(models)
    p = make_params(scenes.CBOX_CAMERA, (128, 128), mat.shape, 4, 1, math="fast")
    dimg = np.ones((128, 128, 4), np.float32)
    g.reset_counters(); o.reset_ray_counters()
    g.backward(ZDR_DIRECT, p, mat, dimg); o.render_backward(ZDR_DIRECT, p, mat, dimg)
    cg = g.counters()
    assert (cg["closest_rays"], cg["any_rays"]) == o.ray_counters()
    g.close()


def test_scene_math_attribute():
    torch = pytest.importorskip("torch")
    from zdr_b200 import Scene
    scene = Scene(scenes.cbox_path_models(), integrator="path", device=0)
    scene.camera = scenes.CBOX_CAMERA.copy()
    scene.max_depth = 4
    mat = torch.from_numpy(smooth_material()).cuda().requires_grad_()
    out = {}
    for mode in ("exact", "fast"):
        scene.math = mode
        mat.grad = None
        img = scene.render(mat, res=(64, 64), spp=4, seed=0)
        img.sum().backward()
        out[mode] = (img.detach().cpu().numpy(), mat.grad.detach().cpu().numpy().copy())
    assert rel_l2(out["fast"][0], out["exact"][0]) <= IMG_TOL and rel_l2(out["fast"][1], out["exact"][1]) <= GRAD_TOL
    assert not np.array_equal(bits(out["fast"][0]), bits(out["exact"][0]))      # it IS another build
    with pytest.raises(KeyError):
        scene.math = "sloppy"
        scene.render(mat, res=(8, 8), spp=1, seed=0)
