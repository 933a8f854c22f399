"""Properties of the CPU oracle itself (the checker must be right before it checks anything) and
the finite-difference arbitration of the two PRB adjoint weightings (fd_validate.py procedure)."""
import numpy as np
import pytest

from zdr_b200._abi import (ZDR_COLLOCATED, ZDR_DIRECT, ZDR_PATH, ZDR_PRB_CORRECTED, ZDR_PRB_LITERAL, make_params)

from . import scenes


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def test_xxhash_matches_independent_python(oracle):
    """pmj02bn.py:59-74 restated with Python ints."""
    M = 0xFFFFFFFF
    P2, P3, P4, P5 = 2246822519, 3266489917, 668265263, 374761393

    def rotl(h):
        return ((h << 17) | (h >> 15)) & M

    def ref(x, y, z, w):
        x, y, z, w = (v & M for v in (x, y, z, w))
        h = (w + P5 + x * P3) & M
        h = (P4 * rotl(h)) & M
        h = (h + y * P3) & M
        h = (P4 * rotl(h)) & M
        h = (h + z * P3) & M
        h = (P4 * rotl(h)) & M
        h = (P2 * (h ^ (h >> 15))) & M
        h = (P3 * (h ^ (h >> 13))) & M
        return h ^ (h >> 16)
    rng = np.random.default_rng(0)
    for _ in range(300):
        a = [int(v) for v in rng.integers(-2 ** 31, 2 ** 31, 4)]
        assert oracle.xxhash32_int4(*a) == ref(*a)


@pytest.mark.parametrize("l", [1, 2, 7, 16, 64, 100, 256, 1000])
def test_permutation_is_a_bijection(oracle, l):
    w = l - 1
    for s in (1, 2, 4, 8, 16):
        w |= w >> s
    # (i + p) % l is a bijection of [0,l) for every p when l is a power of two; for other l only while
    # i + p does not wrap (a property of Kensler's published code, which pmj02bn.py:57 inherits)
    pow2 = l & (l - 1) == 0
    for p in (0, 1, 12345, 0x3FFFFFFF) + ((0xDEADBEEF, 0x7FFFFFFF, 0xFFFFFFFF) if pow2 else ()):
        out = sorted(oracle.permutation_element(i, l, w, p) for i in range(l))
        assert out == list(range(l))


def test_sampler_draws_are_stratified_and_in_range(oracle):
    """next2f() of the first five 2-D dimensions walks a (0,2)-sequence: over a power-of-two number of
    samples of one pixel the points fill every elementary interval exactly once (up to the
    blue-noise toroidal shift, which preserves the per-axis 1-D stratification)."""
    spp = 64
    st = np.array([[17, 33, 5, spp, i] for i in range(spp)], np.int32)
    u = oracle.sampler_draws(st, "2")
    assert (u >= 0).all() and (u < 1).all()
    for axis in range(2):
        shifted = np.sort(u[:, axis])
        gaps = np.diff(np.concatenate([shifted, [shifted[0] + 1]]))
        assert gaps.max() < 2.0 / spp + 1e-6        # one point per 1/spp stratum, shifted cyclically
    u1 = oracle.sampler_draws(st, "21")[:, 2]
    assert sorted(np.floor(u1 * spp).astype(int).tolist()) == list(range(spp))      # generate_1d: one per stratum


def test_ggx_adjoint_matches_central_differences(oracle):
    rng = np.random.default_rng(3)
    n = 400

    def hemi(k):
        v = rng.normal(size=(k, 3))
        v[:, 2] = np.abs(v[:, 2]) + 0.05
        return v / np.linalg.norm(v, axis=1, keepdims=True)
    base = np.concatenate([hemi(n), hemi(n), rng.uniform(0.1, 1, (n, 3)), rng.uniform(0.3, 1, (n, 1)), rng.normal(size=(n, 3))], axis=1).astype(np.float32)
    out = oracle.ggx(base)
    g = base[:, 10:13].astype(np.float64)
    for k, col in enumerate((6, 7, 8, 9)):
        eps = 2e-3
        hi, lo = base.copy(), base.copy()
        hi[:, col] += eps
        lo[:, col] -= eps
        fd = ((oracle.ggx(hi)[:, :3].astype(np.float64) - oracle.ggx(lo)[:, :3]) / (hi[:, col].astype(np.float64) - lo[:, col]).reshape(-1, 1) * g).sum(axis=1)
        assert rel_l2(out[:, 3 + k], fd) < 5e-3, (k, rel_l2(out[:, 3 + k], fd))


def test_ggx_sampling_pdf_integrates_to_one(oracle):
    """furnace-style check of ggx_sample_pdf (microfacet.py:51-58) on a hemisphere quadrature."""
    th, ph = np.meshgrid((np.arange(400) + 0.5) / 400 * np.pi / 2, (np.arange(400) + 0.5) / 400 * 2 * np.pi, indexing="ij")
    wi = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=-1).reshape(-1, 3)
    dw = (np.sin(th) * (np.pi / 2 / 400) * (2 * np.pi / 400)).reshape(-1)
    for wo, r in (((0.0, 0.0, 1.0), 0.6), ((0.6, 0.0, 0.8), 0.8), ((0.3, 0.4, 0.866), 1.0)):
        inp = np.concatenate([np.broadcast_to(np.asarray(wo, np.float32), wi.shape), wi, np.full((len(wi), 1), r)], axis=1).astype(np.float32)
        total = float((oracle.ggx_sample_pdf(inp).astype(np.float64) * dw).sum())
        assert 0.7 < total <= 1.01, total       # the VNDF lobe loses the mass it reflects below the horizon


def test_alias_table_reproduces_distribution(oracle):
    rng = np.random.default_rng(1)
    v = rng.uniform(0, 1, 97) ** 3
    prob, alias, pdf = oracle.create_alias_table(v)
    assert np.isclose(pdf.sum(), 1.0)
    mass = np.zeros_like(v)
    for i in range(len(v)):
        mass[i] += prob[i]
        mass[alias[i]] += 1 - prob[i]
    assert np.allclose(mass / len(v), v / v.sum(), atol=1e-6)


def test_forward_is_deterministic_and_sharding_is_exact(oracle):
    models, mat = scenes.cbox_path_models(), scenes.random_material((64, 64), 2)
    sc = oracle.OracleScene(models)
    p = make_params(scenes.CBOX_CAMERA, (40, 24), mat.shape, 4, 3, max_depth=5)
    a, b = sc.render_forward(ZDR_PATH, p, mat), sc.render_forward(ZDR_PATH, p, mat)
    assert np.array_equal(a, b) and np.isfinite(a).all()
    parts = np.zeros_like(a)
    for k in range(3):       # round-robin 8-row tiles (zdr_params.shard_index)
        parts += sc.render_forward(ZDR_PATH, make_params(scenes.CBOX_CAMERA, (40, 24), mat.shape, 4, 3, max_depth=5, shard=(k, 3)), mat)
    assert np.array_equal(parts, a)


def _directional(oracle, sc, integ, mat, delta, res, spp, seed, eps, **kw):
    """fd_validate.py:57-114: two-sided finite difference of <w, I(mat)> along `delta` (same seed for
    +-eps, :74-81) vs the adjoint estimate <d_material, delta>.  The adjoint renders with seed+1
    (render.py:196), so the forward is differenced at seed+1 to compare like with like."""
    rng = np.random.default_rng(5)
    w = rng.uniform(0.5, 1.5, (res[1], res[0], 4)).astype(np.float32)
    p_fd = make_params(scenes.CBOX_CAMERA, res, mat.shape, spp, seed + 1, **kw)
    hi = sc.render_forward(integ, p_fd, (mat + eps * delta).astype(np.float32)).astype(np.float64)
    lo = sc.render_forward(integ, p_fd, (mat - eps * delta).astype(np.float32)).astype(np.float64)
    fd = ((hi - lo)[..., :3] * w[..., :3]).sum() / (2 * eps)
    dm = sc.render_backward(integ, p_fd, mat, w)           # caller passes seed+1 itself at this level
    ad = (dm.astype(np.float64) * delta).sum()
    return fd, ad


@pytest.mark.parametrize("integ,models", [(ZDR_COLLOCATED, scenes.cbox_collocated_models), (ZDR_DIRECT, scenes.cbox_path_models)])
def test_fd_vs_adjoint_diffuse_exact(oracle, integ, models):
    """The estimators are linear in the diffuse albedo and sampling does not depend on it, so FD with
    the adjoint's own sample set must agree with the adjoint to float accuracy."""
    mat = scenes.random_material((32, 32), 1)
    mat[..., 3] = 0.4 + 0.5 * mat[..., 3]
    delta = np.zeros_like(mat)
    delta[..., :3] = np.random.default_rng(2).uniform(-1, 1, mat[..., :3].shape)
    sc = oracle.OracleScene(models())
    fd, ad = _directional(oracle, sc, integ, mat.astype(np.float64), delta, (48, 48), 4, 11, 1e-2)
    assert abs(fd - ad) <= 2e-3 * abs(fd), (fd, ad)


@pytest.mark.parametrize("integ,models", [(ZDR_COLLOCATED, scenes.cbox_collocated_models), (ZDR_DIRECT, scenes.cbox_path_models)])
def test_fd_vs_adjoint_roughness_statistical(oracle, integ, models):
    """Roughness moves the BSDF samples (detached in the adjoint, direct.py:126-131), so agreement is
    statistical, as fd_validate.py:114 says ('good if last rows similar')."""
    mat = scenes.random_material((16, 16), 1)
    mat[..., 3] = 0.5 + 0.4 * mat[..., 3]
    delta = np.zeros_like(mat)
    delta[..., 3] = 1.0
    sc = oracle.OracleScene(models())
    fd, ad = _directional(oracle, sc, integ, mat.astype(np.float64), delta, (64, 64), 16, 19260817, 1e-2)
    assert abs(fd - ad) <= (1e-3 if integ == ZDR_COLLOCATED else 0.08) * abs(fd), (fd, ad)


def test_fd_arbitrates_prb_weighting(oracle):
    """SURVEY.md 8a defect 3: prb.py:162 seeds the BSDF-sample adjoint with beta/pdf*Le, textbook PRB
    with Le_remaining/f.  Without Russian roulette the path estimator is a polynomial in the diffuse
    albedo on a fixed sample set, so central differences are exact to O(eps^2): they agree with the
    corrected weighting and reject the literal one."""
    mat = scenes.random_material((32, 32), 4)
    mat[..., 3] = 0.5 + 0.4 * mat[..., 3]
    mat[..., :3] = 0.3 + 0.6 * mat[..., :3]
    delta = np.zeros_like(mat)
    delta[..., :3] = np.random.default_rng(8).uniform(0.2, 1, mat[..., :3].shape)
    sc = oracle.OracleScene(scenes.cbox_path_models())
    kw = dict(max_depth=5, rr_depth=99)
    fd, ad_c = _directional(oracle, sc, ZDR_PATH, mat.astype(np.float64), delta, (40, 40), 8, 3, 5e-3, prb_mode=ZDR_PRB_CORRECTED, **kw)
    _, ad_l = _directional(oracle, sc, ZDR_PATH, mat.astype(np.float64), delta, (40, 40), 8, 3, 5e-3, prb_mode=ZDR_PRB_LITERAL, **kw)
    assert abs(fd - ad_c) <= 3e-3 * abs(fd), (fd, ad_c, ad_l)
    assert abs(fd - ad_l) >= 0.05 * abs(fd), (fd, ad_c, ad_l)


def test_envmap_tables_and_render_smoke(oracle):
    env = scenes.procedural_envmap(64, 32, seed=2)
    sc = oracle.OracleScene(scenes.sphere_grid_models(3, seed=1))
    sc.set_envmap(env)
    prob, alias, pdf = sc.env_tables()
    assert np.isfinite(pdf).all() and (prob >= 0).all() and (prob <= 1.0 + 1e-6).all()
    assert np.isclose(pdf.mean(), 1.0, rtol=1e-3)                       # pdf is scaled by pixel_count (envmap.py:197)
    cam = scenes.Camera(fov=0.9, origin=scenes.float3(-2.0, 2.5, -4.0), target=scenes.float3(1.0, 0.3, 0.3), up=scenes.float3(0, 1, 0))
    mat = scenes.random_material((16, 16), 3)
    img = sc.render_forward(ZDR_PATH, make_params(cam, (32, 24), mat.shape, 4, 0, max_depth=4), mat)
    assert np.isfinite(img).all() and img[..., :3].mean() > 0.05


def test_oracle_bvh_equals_brute_force_scan(oracle, monkeypatch):
    """Scenes of >= 4096 triangles are scanned through a median-split BVH inside the oracle (test infrastructure: the
    same triangle test, smallest t, ties to the lowest scene-order triangle).  It must answer exactly as the in-order
    brute-force scan does: hits, barycentrics, occlusion, rendered samples, adjoint and ray counts."""
    rng = np.random.default_rng(2)

    def rays(n, lo, hi):
        o = rng.uniform(lo, hi, (n, 3))
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d[: n // 20, rng.integers(0, 3)] = 0.0                      # axis-parallel directions (slab test with 1/0)
        return np.concatenate([o, d, np.zeros((n, 1)), np.full((n, 1), 1e30)], axis=1).astype(np.float32)
    models = scenes.sphere_grid_models(12, seed=3, emissive_every=5)          # 11 520 triangles
    r = rays(1500, (-1, -1, -1), (4, 4, 2))
    seg = r.copy()
    seg[:, 6], seg[:, 7] = 1e-4, 2.5
    cam = scenes.Camera(fov=0.9, origin=scenes.float3(-2.0, 3.0, -4.0), target=scenes.float3(2.0, 0.5, 0.3), up=scenes.float3(0, 1, 0))
    mat = scenes.random_material((32, 32), 2)
    p = make_params(cam, (24, 16), mat.shape, 2, 0, max_depth=4)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ZDR_ORACLE_BVH", flag)
        sc = oracle.OracleScene(models)
        ip, uvt = sc.trace_closest(r)
        sc.reset_ray_counters()
        smp = sc.render_samples(ZDR_PATH, p, mat)
        grad = sc.render_backward(ZDR_PATH, p, mat, np.ones((16, 24, 4), np.float32))
        out[flag] = (ip, uvt, sc.trace_any(seg), smp, grad, sc.ray_counters())
    a, b = out["0"], out["1"]
    hit = a[0][:, 0] >= 0
    assert hit.sum() > 50 and np.array_equal(a[0], b[0])
    assert np.array_equal(a[1][hit].view(np.uint32), b[1][hit].view(np.uint32))
    assert np.array_equal(a[2], b[2]) and 0 < a[2].sum() < len(a[2])
    assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32)) and a[3].max() > 0
    assert np.allclose(a[4], b[4], rtol=1e-6, atol=1e-9) and a[5] == b[5]


def test_oracle_bvh_equals_scan_on_degenerate_geometry(oracle, monkeypatch, tmp_path):
    """The same equivalence on hostile input: 6 000 triangles of very different sizes, 300 exact copies of one triangle
    (every tie must go to the first copy), zero-area triangles, one huge triangle under everything."""
    rng = np.random.default_rng(5)
    n = 6000
    ctr = rng.uniform(-4, 4, (n, 1, 3))
    size = np.exp(rng.uniform(np.log(1e-3), np.log(0.8), (n, 1, 1)))
    tri = ctr + rng.normal(0, 1, (n, 3, 3)) * size
    dup = np.repeat(np.array([[[0.5, 0.5, 0.0], [1.5, 0.5, 0.2], [0.5, 1.5, 0.1]]]), 300, axis=0)
    flat = np.repeat(rng.uniform(-4, 4, (100, 1, 3)), 3, axis=1)
    huge = np.array([[[-50.0, -50.0, -4.5], [50.0, -50.0, -4.5], [0.0, 80.0, -4.5]]])
    tri = np.concatenate([tri[:3000], dup, flat, tri[3000:], huge]).astype(np.float32)
    path = tmp_path / "soup.obj"
    with open(path, "w") as f:
        for t in tri:
            for v in t:
                f.write(f"v {float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")
        for i in range(len(tri)):
            f.write(f"f {3 * i + 1} {3 * i + 2} {3 * i + 3}\n")
    o = rng.uniform(-4, 4, (3000, 3))
    d = rng.normal(size=(3000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.zeros((3000, 1)), np.full((3000, 1), 1e30)], axis=1).astype(np.float32)
    rays[:400, 0:3] = (0.8, 0.8, 1.0)                                                # straight at the stack of copies
    rays[:400, 3:6] = rng.normal(0, 0.05, (400, 3)) + (0.0, 0.0, -1.0)
    seg = rays.copy()
    seg[:, 6], seg[:, 7] = 1e-4, 2.0
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ZDR_ORACLE_BVH", flag)
        sc = oracle.OracleScene([(str(path), None, scenes.float3(0.0))])
        out[flag] = sc.trace_closest(rays) + (sc.trace_any(seg),)
    a, b = out["0"], out["1"]
    hit = a[0][:, 0] >= 0
    assert hit.sum() > 500 and np.array_equal(a[0], b[0]) and np.array_equal(a[1][hit].view(np.uint32), b[1][hit].view(np.uint32))
    assert np.array_equal(a[2], b[2])
    copies = (a[0][:400, 1] >= 3000) & (a[0][:400, 1] < 3300)
    assert copies.sum() > 20 and (a[0][:400, 1][copies] == 3000).all()


@pytest.mark.parametrize("integ,kw", [(ZDR_COLLOCATED, {}), (ZDR_DIRECT, {}), (ZDR_PATH, {"max_depth": 4}),
                                      (ZDR_PATH, {"max_depth": 4, "prb_mode": ZDR_PRB_LITERAL})])
def test_material_sets_properties(oracle, integ, kw):
    """Material sets ("polymorphic materials", README.md:5): (1) the default assignment is the reference's single-set rule;
    (2) two sets holding the SAME texels render the same image as one set shared by both instances, and the shared set's
    gradient is the sum of the two sets' gradients; (3) a perturbation of set 1 moves only pixels that see instance 2."""
    models, sets = scenes.cbox_material_sets_models()
    res, spp = (24, 24), 2
    tex = scenes.random_material((16, 16), 3)
    two = np.stack([tex, tex])
    d_img = np.random.default_rng(5).uniform(0.5, 1.5, (res[1], res[0], 4)).astype(np.float32)
    sc = oracle.OracleScene(models)
    base = sc.render_forward(integ, make_params(scenes.CBOX_CAMERA, res, tex.shape, spp, 2, **kw), tex)      # default: [0, -1, -1, -1]
    sc.set_material_sets([0, -1, -1, -1])
    assert np.array_equal(base, sc.render_forward(integ, make_params(scenes.CBOX_CAMERA, res, tex.shape, spp, 2, **kw), tex))
    sc.set_material_sets([0, -1, 0, -1])                                  # both textured instances share set 0
    shared = sc.render_forward(integ, make_params(scenes.CBOX_CAMERA, res, tex.shape, spp, 2, **kw), tex)
    g_shared = sc.render_backward(integ, make_params(scenes.CBOX_CAMERA, res, tex.shape, spp, 3, **kw), tex, d_img)
    sc.set_material_sets(sets)                                            # [0, -1, 1, -1]
    split = sc.render_forward(integ, make_params(scenes.CBOX_CAMERA, res, two.shape, spp, 2, **kw), two)
    g_split = sc.render_backward(integ, make_params(scenes.CBOX_CAMERA, res, two.shape, spp, 3, **kw), two, d_img)
    assert np.array_equal(shared, split)
    assert g_split.shape == two.shape and np.abs(g_split[1]).max() > 0
    assert np.allclose(g_split[0] + g_split[1], g_shared, rtol=1e-5, atol=1e-7)
    if integ != ZDR_COLLOCATED:
        assert not np.array_equal(base, split)                            # the sphere is shaded now, it was a black terminal before
    other = two.copy()
    other[1] = scenes.random_material((16, 16), 9)
    moved = sc.render_forward(integ, make_params(scenes.CBOX_CAMERA, res, two.shape, spp, 2, **kw), other)
    assert not np.array_equal(moved, split)
