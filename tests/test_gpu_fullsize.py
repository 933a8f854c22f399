"""GPU tests at BASELINE.json's FULL sizes.  The oracle cannot render C3 whole in seconds, so full-size parity
goes through what it can do exactly — whole rows of the full-size job (every pixel's complete 256-sample loop) —
and through size-independent properties: row/tile partitions sum to the whole, the two schedules of the path
integrator agree bit for bit, a masked cotangent equals the launch restricted to the unmasked rows, gradients are
linear in the cotangent.  C2 (direct, 512^2, 64 spp) is small enough to be compared whole."""
import numpy as np
import pytest

from zdr_b200._abi import ZDR_DIRECT, ZDR_PATH, make_params

from . import scenes
from .test_gpu_parity import GRAD_TOL, GpuScene, bits, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3(oracle):
    models = scenes.cbox_path_models()
    g, o = GpuScene(models), oracle.OracleScene(models)
    mat = scenes.random_material((1024, 1024), 0)
    kw = {"max_depth": 8}
    image = g.forward(ZDR_PATH, make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 0, **kw), mat)
    yield {"g": g, "o": o, "mat": mat, "kw": kw, "image": image}
    g.close()


def test_c3_full_size_rows_bit_exact_vs_oracle(c3):
    """Rows of the 1024^2 x 256 spp image (wavefront schedule, the bench's path) against the oracle rendering the
    same rows of the same job: bit-identical."""
    lit = 0.0
    for r0, r1 in [(0, 1), (300, 301), (511, 513), (1023, 1024)]:
        p = make_params(scenes.CBOX_CAMERA, (1024, 1024), c3["mat"].shape, 256, 0, rows=(r0, r1), **c3["kw"])
        ref = c3["o"].render_forward(ZDR_PATH, p, c3["mat"])
        lit = max(lit, float(ref[r0:r1, :, :3].max()))
        assert np.array_equal(bits(c3["image"][r0:r1]), bits(ref[r0:r1]))
    assert lit > 0


def test_c3_full_size_partitions_and_schedules(c3):
    g, mat, kw, full = c3["g"], c3["mat"], c3["kw"], c3["image"]
    assert np.isfinite(full).all() and (full[..., 3] == 1).all()
    # 8 round-robin tile shards (the 8-GPU decomposition) tile the image exactly
    acc = np.zeros_like(full)
    for k in range(8):
        acc += g.forward(ZDR_PATH, make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 0, shard=(k, 8), **kw), mat)
    assert np.array_equal(bits(acc), bits(full))
    # the one-kernel schedule on a band of rows equals the wavefront schedule's rows
    band = g.forward(ZDR_PATH, make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 0, rows=(300, 332), schedule="megakernel", **kw), mat)
    assert np.array_equal(bits(band[300:332]), bits(full[300:332]))
    # every schedule renders the same full-size image: HBM wavefront, persistent shared-memory pools
    for sched in ("wavefront", "persistent"):
        kw2 = dict(kw, schedule=sched)
        img = g.forward(ZDR_PATH, make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 0, **kw2), mat)
        assert np.array_equal(bits(img), bits(full)), sched


def test_c3_full_size_gradient_rows_vs_oracle_and_linearity(c3):
    """Adjoint pass at full size.  (1) the launch restricted to two rows against the oracle's same launch;
    (2) the FULL launch with a cotangent that is zero outside those rows gives the same gradient;
    (3) gradient(2 * cotangent) = 2 * gradient; (4) the halves of the image sum to the whole."""
    g, o, mat, kw = c3["g"], c3["o"], c3["mat"], c3["kw"]
    rows = (640, 642)
    p_rows = make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 1, rows=rows, **kw)
    p_full = make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 1, **kw)
    rng = np.random.default_rng(3)
    dimg = np.zeros((1024, 1024, 4), np.float32)
    dimg[rows[0]:rows[1]] = rng.uniform(0.5, 1.5, (rows[1] - rows[0], 1024, 4)).astype(np.float32)
    ref = o.render_backward(ZDR_PATH, p_rows, mat, dimg)
    got_rows = g.backward(ZDR_PATH, p_rows, mat, dimg)
    assert np.abs(ref).max() > 0 and rel_l2(got_rows, ref) <= GRAD_TOL, rel_l2(got_rows, ref)
    got_masked = g.backward(ZDR_PATH, p_full, mat, dimg)
    assert rel_l2(got_masked, ref) <= GRAD_TOL, rel_l2(got_masked, ref)
    assert rel_l2(g.backward(ZDR_PATH, p_rows, mat, 2.0 * dimg), 2.0 * ref) <= GRAD_TOL
    ones = np.ones((1024, 1024, 4), np.float32)
    whole = g.backward(ZDR_PATH, p_full, mat, ones)
    halves = sum(g.backward(ZDR_PATH, make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 1, rows=r, **kw), mat, ones)
                 for r in [(0, 512), (512, 1024)])
    assert np.isfinite(whole).all() and rel_l2(halves, whole) < 1e-5
    c = g.counters()
    assert c["closest_rays"] > 0
    # the schedules are the same sum in a different order: HBM wavefront (the default at this size) vs persistent pools;
    # and they trace the same rays
    counts = {}
    for sched in ("wavefront", "persistent"):
        g.reset_counters()
        p_s = make_params(scenes.CBOX_CAMERA, (1024, 1024), mat.shape, 256, 1, schedule=sched, **kw)
        assert rel_l2(g.backward(ZDR_PATH, p_s, mat, ones), whole) < 1e-5, sched
        c2 = g.counters()
        counts[sched] = (c2["closest_rays"], c2["any_rays"], c2["camera_samples"], c2["shaded_vertices"], c2["grad_scatters"])
    assert counts["wavefront"] == counts["persistent"] and counts["persistent"][0] > 0, counts


def test_c2_full_size_vs_oracle(oracle):
    """C2 whole: cbox-combined (uv = 0: every scatter on one texel) + light, direct integrator, 512^2, 64 spp."""
    models = scenes.cbox_direct_models()
    g, o = GpuScene(models), oracle.OracleScene(models)
    mat = scenes.random_material((1024, 1024), 0)
    p = make_params(scenes.CBOX_CAMERA, (512, 512), mat.shape, 64, 0)
    assert np.array_equal(bits(g.forward(ZDR_DIRECT, p, mat)), bits(o.render_forward(ZDR_DIRECT, p, mat)))
    p = make_params(scenes.CBOX_CAMERA, (512, 512), mat.shape, 64, 1)
    dimg = np.ones((512, 512, 4), np.float32)
    dg, do = g.backward(ZDR_DIRECT, p, mat, dimg), o.render_backward(ZDR_DIRECT, p, mat, dimg)
    assert np.abs(do).max() > 0 and rel_l2(dg, do) <= GRAD_TOL, rel_l2(dg, do)
    assert np.count_nonzero(do) <= 8                     # all of it lands on the (0, Ht-1) footprint
    g.close()
