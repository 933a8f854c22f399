"""The torch surface (zdr_b200.Scene) on the GPU: the drop-in contract of render.py:31-257."""
import numpy as np
import pytest
import torch

from zdr_b200._abi import INTEGRATORS, make_params

from . import scenes

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


@pytest.fixture(scope="module")
def scene():
    from zdr_b200 import Scene
    sc = Scene(scenes.cbox_path_models(), integrator="direct")
    sc.camera = scenes.CBOX_CAMERA.copy()
    return sc


def test_render_shapes_and_alpha(scene):
    mat = torch.from_numpy(scenes.random_material((128, 128), 0)).cuda()
    img = scene.render(mat, res=(96, 64), spp=4)
    assert img.shape == (64, 96, 4) and img.dtype == torch.float32 and img.is_cuda      # (H, W, 4), render.py:163
    assert bool((img[..., 3] == 1).all()) and float(img[..., :3].max()) > 0


@pytest.mark.parametrize("integrator", ["collocated", "direct", "path"])
def test_autograd_matches_oracle(oracle, integrator):
    from zdr_b200 import Scene
    models = scenes.cbox_path_models()
    sc = Scene(models, integrator=integrator)
    sc.camera = scenes.CBOX_CAMERA.copy()
    sc.max_depth = 6
    mat_np = scenes.random_material((96, 96), 3)
    mat = torch.from_numpy(mat_np).cuda().requires_grad_()
    img = sc.render(mat, res=(64, 48), spp=4, seed=5)
    w = torch.from_numpy(np.random.default_rng(1).uniform(0.5, 1.5, (48, 64, 4)).astype(np.float32)).cuda()
    (img * w).sum().backward()
    o = oracle.OracleScene(models)
    ref_img = o.render_forward(INTEGRATORS[integrator], make_params(scenes.CBOX_CAMERA, (64, 48), mat_np.shape, 4, 5, max_depth=6), mat_np)
    ref_grad = o.render_backward(INTEGRATORS[integrator], make_params(scenes.CBOX_CAMERA, (64, 48), mat_np.shape, 4, 6, max_depth=6), mat_np,
                                 w.cpu().numpy())          # seed+1, render.py:196
    assert np.array_equal(img.detach().cpu().numpy(), ref_img)
    assert rel_l2(mat.grad.cpu().numpy(), ref_grad) <= 1e-3


def test_backward_uses_forward_camera_and_lights(scene, oracle):
    """render.py:207-218: camera and emissions are snapshotted at forward time."""
    mat_np = scenes.random_material((64, 64), 2)
    mat = torch.from_numpy(mat_np).cuda().requires_grad_()
    img = scene.render(mat, res=(48, 48), spp=2, seed=1)
    cam_bak = scene.camera
    moved = scenes.CBOX_CAMERA.copy()
    moved.origin = scenes.float3(0.5, 2.0, 5.0)
    scene.camera = moved                                    # changed between forward and backward
    img[..., :3].sum().backward()
    assert scene.camera is moved                            # restored after the adjoint render (render.py:222)
    scene.camera = cam_bak
    o = oracle.OracleScene(scenes.cbox_path_models())
    d = np.ones((48, 48, 4), np.float32)
    d[..., 3] = 0
    ref = o.render_backward(INTEGRATORS["direct"], make_params(scenes.CBOX_CAMERA, (48, 48), mat_np.shape, 2, 2), mat_np, d)
    assert rel_l2(mat.grad.cpu().numpy(), ref) <= 1e-3


def test_non_contiguous_grad_output_and_stream(scene):
    mat = torch.from_numpy(scenes.random_material((64, 64), 2)).cuda().requires_grad_()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):                              # the library enqueues on torch's current stream
        img = scene.render(mat, res=(32, 32), spp=2)
        loss = img.permute(1, 0, 2).flip(0).sum()           # non-contiguous grad_output (render.py:181)
        loss.backward()
    s.synchronize()
    assert torch.isfinite(mat.grad).all() and float(mat.grad.abs().sum()) > 0


def test_uniform_texture_contention_case(oracle):
    """C2/C4 shape: uv == 0 (cbox-combined.obj has no vt) sends every scatter of the image to one texel;
    the warp-aggregated scatter must still sum to the oracle's gradient."""
    from zdr_b200 import Scene
    models = scenes.cbox_direct_models()
    sc = Scene(models, integrator="direct")
    sc.camera = scenes.CBOX_CAMERA.copy()
    mat_np = scenes.constant_material((8, 8))
    mat = torch.from_numpy(mat_np).cuda().requires_grad_()
    sc.reset_counters()
    img = sc.render(mat, res=(128, 128), spp=8, seed=0)
    img[..., :3].sum().backward()
    c = sc.counters()
    o = oracle.OracleScene(models)
    d = np.ones((128, 128, 4), np.float32)
    d[..., 3] = 0
    ref = o.render_backward(INTEGRATORS["direct"], make_params(scenes.CBOX_CAMERA, (128, 128), mat_np.shape, 8, 1), mat_np, d)
    g = mat.grad.cpu().numpy()
    assert (g != 0).sum() == (ref != 0).sum() == 4
    assert rel_l2(g, ref) <= 1e-3
    assert c["red_v4"] * 8 < c["grad_scatters"] * 4         # aggregation: far fewer reductions than texel writes


def test_scatter_modes_agree(oracle):
    """plain, warp-aggregated and block-level (shared-memory duplicate buffer) scatter are schedules of the same sum, on every
    schedule of the path integrator and for the direct integrator on a uv-less mesh (every scatter on one texel)."""
    from zdr_b200 import Scene
    for models, integrator, schedules in ((scenes.cbox_path_models, "path", ("megakernel", "wavefront", "persistent")),
                                          (scenes.cbox_direct_models, "direct", ("auto",))):
        sc = Scene(models(), integrator=integrator)
        sc.camera = scenes.CBOX_CAMERA.copy()
        sc.max_depth = 5
        for schedule in schedules:
            sc.schedule = schedule
            grads, reds = {}, {}
            for mode in ("plain", "aggregate", "block", "auto"):
                sc.scatter_mode = mode
                sc.reset_counters()
                mat = torch.from_numpy(scenes.random_material((64, 64), 7)).cuda().requires_grad_()
                sc.render(mat, res=(64, 64), spp=4, seed=3)[..., :3].sum().backward()
                grads[mode] = mat.grad.cpu().numpy()
                reds[mode] = sc.counters()["red_v4"]
            for mode in ("aggregate", "block", "auto"):
                assert rel_l2(grads[mode], grads["plain"]) < 1e-5, (integrator, schedule, mode)
            assert reds["block"] <= reds["aggregate"] <= reds["plain"], (integrator, schedule, reds)
        sc.close()


def test_render_duvdxy_matches_oracle(scene, oracle):
    """render.py:243-257 / uvgrad.py: screen-to-texture Jacobian; footprint example of example.py:62-64."""
    mat = torch.from_numpy(scenes.random_material((32, 32), 0)).cuda()
    out = scene.render_duvdxy(mat, res=(80, 64), spp=4, seed=3)
    assert out.shape == (64, 80, 4)
    o = oracle.OracleScene(scenes.cbox_path_models())
    ref = o.render_duvdxy(make_params(scenes.CBOX_CAMERA, (80, 64), (32, 32), 4, 3))
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    footprints = torch.det(out.reshape(64, 80, 2, 2)).abs() * 1024 * 1024
    assert float(footprints.mean()) > 1.0


def test_inverse_rendering_recovers_texture():
    """example.py:74-91 in small: Adam on the image loss through Scene.render / backward lowers the image MSE by a lot
    (the gradients point the right way, with the backward seed differing from the forward seed)."""
    import random
    from zdr_b200 import Scene
    sc = Scene(scenes.cbox_path_models(), integrator="path")
    sc.camera = scenes.CBOX_CAMERA.copy()
    sc.max_depth = 4
    torch.manual_seed(0)
    random.seed(0)
    u = torch.linspace(0, 1, 32, device="cuda")
    vv, uu = torch.meshgrid(u, u, indexing="ij")
    gt = torch.stack([0.5 + 0.4 * torch.sin(9 * uu), 0.5 + 0.4 * torch.cos(7 * vv), 0.5 + 0.4 * torch.sin(5 * (uu + vv)),
                      0.6 + 0.2 * torch.cos(6 * uu * vv)], dim=-1).clamp(1e-3, 1).contiguous()
    res = (96, 96)
    target = sc.render(gt, res=res, spp=256)
    mat = torch.rand((32, 32, 4), device="cuda").requires_grad_()
    opt = torch.optim.Adam([mat], lr=0.02)
    with torch.no_grad():
        mse0 = ((sc.render(mat, res=res, spp=256) - target) ** 2).mean().item()
    for _ in range(150):
        img = sc.render(mat, res=res, spp=16, seed=random.randint(0, 2147483647))
        ((img - target) ** 2).sum().backward()
        opt.step()
        opt.zero_grad()
        mat.data.clamp_(min=1e-3, max=1)
    with torch.no_grad():
        mse1 = ((sc.render(mat, res=res, spp=256) - target) ** 2).mean().item()
    assert mse1 < 0.25 * mse0, (mse0, mse1)
    sc.close()


@pytest.mark.parametrize("schedule,res,spp", [("megakernel", (64, 48), 8), ("wavefront", (64, 48), 8),
                                              ("wavefront", (1024, 1024), 8),       # 2^23 samples: two streams inside
                                              ("persistent", (64, 48), 8), ("persistent", (512, 512), 8)])
def test_operator_is_cuda_graph_capturable(schedule, res, spp):
    """The operator launches on the caller's stream with no synchronisation, allocation or host read-back once its
    workspace exists, so forward + adjoint capture into a CUDA graph; replays track the material tensor's contents."""
    import ctypes as C
    from zdr_b200 import Scene
    from zdr_b200._abi import ZDR_PATH, check
    sc = Scene(scenes.cbox_path_models(), integrator="path")
    lib, h = sc._lib, sc._handle
    mats = [torch.from_numpy(scenes.random_material((64, 64), k)).cuda() for k in range(3)]
    material = mats[0].clone()
    image = torch.zeros((res[1], res[0], 4), device="cuda")
    d_image = torch.ones_like(image)
    d_material = torch.zeros_like(material)
    pf = make_params(scenes.CBOX_CAMERA, res, (64, 64), spp, 0, max_depth=4, schedule=schedule)
    pb = make_params(scenes.CBOX_CAMERA, res, (64, 64), spp, 1, max_depth=4, schedule=schedule)

    def run():
        st = torch.cuda.current_stream().cuda_stream
        d_material.zero_()
        check(lib.zdr_render_forward(h, ZDR_PATH, C.byref(pf), material.data_ptr(), image.data_ptr(), st))
        check(lib.zdr_render_backward(h, ZDR_PATH, C.byref(pb), material.data_ptr(), d_image.data_ptr(), d_material.data_ptr(), st))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()                                        # warm-up: workspace allocation, module loading
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        run()
    for m in mats[::-1]:
        material.copy_(m)
        graph.replay()
        torch.cuda.synchronize()
        img_g, grad_g = image.clone(), d_material.clone()
        run()
        torch.cuda.synchronize()
        assert torch.equal(img_g, image)
        assert rel_l2(grad_g.cpu().numpy(), d_material.cpu().numpy()) < 1e-5
        assert float(image[..., :3].max()) > 0 and float(d_material.abs().max()) > 0
    sc.close()


def test_errors_are_python_exceptions(scene):
    with pytest.raises(AssertionError):
        scene.render(torch.zeros((4, 4, 3), device="cuda"), res=(8, 8), spp=1)          # render.py:160
    with pytest.raises(ValueError):
        scene.render(torch.zeros((4, 4, 4)), res=(8, 8), spp=1)                           # host tensor
    with pytest.raises(AssertionError):
        scene.update_lights([None])                                                      # render.py:141


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_equals_single(tmp_path):
    """run with gpurun --gpus 2: Scene.distribute over NCCL (tile axis: bit-identical image; sample axis: equal to
    float accuracy) vs the single-GPU render."""
    import subprocess
    import sys
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ZDR_ROOT"])
from tests import scenes
from zdr_b200 import Scene
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
sc = Scene(scenes.cbox_path_models(), integrator="path", device=rank); sc.camera = scenes.CBOX_CAMERA.copy(); sc.max_depth = 6
mat = torch.from_numpy(scenes.random_material((128, 128), 0)).cuda().requires_grad_()
single = sc.render(mat, res=(96, 80), spp=4, seed=2); single[..., :3].sum().backward(); g1 = mat.grad.clone(); mat.grad = None
sc.distribute()
multi = sc.render(mat, res=(96, 80), spp=4, seed=2); multi[..., :3].sum().backward()
assert torch.equal(single, multi), "image differs"
err = float((mat.grad - g1).norm() / g1.norm()); assert err < 1e-5, err
mat.grad = None
sc.distribute(axis="samples")
multi = sc.render(mat, res=(96, 80), spp=4, seed=2); multi[..., :3].sum().backward()
e_img = float((multi - single).norm() / single.norm()); assert e_img < 1e-6, e_img
assert bool((multi[..., 3] == 1).all())
err2 = float((mat.grad - g1).norm() / g1.norm()); assert err2 < 1e-5, err2
dist.destroy_process_group()
print("rank", rank, "ok", err)
''')
    import os
    env = dict(os.environ, ZDR_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29671", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_load_geometry_twice_replaces_the_scene(oracle):
    """Upstream load_geometry builds a fresh Accel / heap on every call (render.py:74-76): a second call must replace the
    instances (not append), keep Python's and the library's instance counts equal, and render the NEW scene."""
    from zdr_b200 import Scene
    sc = Scene(scenes.cbox_two_light_models(), integrator="direct")
    assert sc.inst_count == 3 and sc._lib.zdr_scene_instance_count(sc._handle) == 3
    sc.load_geometry(scenes.cbox_path_models())
    assert sc.inst_count == 2 and sc._lib.zdr_scene_instance_count(sc._handle) == 2 and sc.light_count == 1
    sc.update_lights([None, scenes.float3(17, 12, 4)])                                 # render.py:141: len == inst_count
    sc.camera = scenes.CBOX_CAMERA.copy()
    mat_np = scenes.random_material((32, 32), 1)
    img = sc.render(torch.from_numpy(mat_np).cuda(), res=(32, 32), spp=2, seed=3)
    ref = oracle.OracleScene(scenes.cbox_path_models()).render_forward(INTEGRATORS["direct"], make_params(scenes.CBOX_CAMERA, (32, 32), mat_np.shape, 2, 3), mat_np)
    assert np.array_equal(img.cpu().numpy(), ref)
    sc.close()


def test_material_view_at_an_odd_offset():
    """The kernels read 16-byte texels; a contiguous view whose storage offset is not a multiple of 4 floats (a slice of a
    flat parameter vector) is copied instead of faulting, and its gradient flows back to the flat vector."""
    from zdr_b200 import Scene
    scene = Scene(scenes.cbox_path_models(), integrator="direct")
    scene.camera = scenes.CBOX_CAMERA.copy()
    base = scenes.random_material((32, 32), 2)
    flat = torch.zeros(1 + base.size, device="cuda")
    flat[1:] = torch.from_numpy(base.ravel()).cuda()
    flat.requires_grad_()
    view = flat[1:].view(32, 32, 4)
    assert view.data_ptr() % 16 != 0 and view.is_contiguous()
    img = scene.render(view, res=(32, 32), spp=2, seed=1)
    img[..., :3].sum().backward()
    m = torch.from_numpy(base).cuda().requires_grad_()
    img2 = scene.render(m, res=(32, 32), spp=2, seed=1)
    img2[..., :3].sum().backward()
    assert torch.equal(img, img2)
    assert rel_l2(flat.grad[1:].cpu().numpy(), m.grad.ravel().cpu().numpy()) < 1e-5 and float(flat.grad[0]) == 0.0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_scene_on_another_device_leaves_the_current_device_alone():
    """Every C entry point restores the caller's current device (torch reads it with cudaGetDevice)."""
    from zdr_b200 import Scene
    torch.cuda.set_device(0)
    sc = Scene(scenes.cbox_path_models(), integrator="direct", device=1)
    assert torch.cuda.current_device() == 0
    sc.camera = scenes.CBOX_CAMERA.copy()
    mat = torch.from_numpy(scenes.random_material((32, 32), 2)).to("cuda:1").requires_grad_()
    with torch.cuda.device(1):
        pass
    img = sc.render(mat, res=(32, 32), spp=2)
    assert torch.cuda.current_device() == 0
    img[..., :3].sum().backward()
    assert torch.cuda.current_device() == 0 and torch.zeros(1, device="cuda").device.index == 0
    assert img.device.index == 1 and float(mat.grad.abs().sum()) > 0
    sc.close()
    assert torch.cuda.current_device() == 0


def test_material_sets_through_the_torch_surface():
    """Scene.set_material_sets + a (K, Ht, Wt, 4) material tensor: image and a gradient of the same shape through autograd;
    too few sets in the tensor is a Python exception."""
    from zdr_b200 import Scene
    models, sets = scenes.cbox_material_sets_models()
    sc = Scene(models, integrator="path")
    sc.camera = scenes.CBOX_CAMERA.copy()
    sc.max_depth = 4
    assert sc.material_sets == [0, -1, -1, -1]
    one = torch.from_numpy(scenes.random_material((32, 32), 1)).cuda().requires_grad_()
    base = sc.render(one, res=(48, 40), spp=4, seed=2)
    sc.set_material_sets(sets)
    with pytest.raises(ValueError):
        sc.render(one, res=(48, 40), spp=4, seed=2)
    two = torch.stack([one.detach(), torch.from_numpy(scenes.random_material((32, 32), 2)).cuda()]).requires_grad_()
    img = sc.render(two, res=(48, 40), spp=4, seed=2)
    assert img.shape == base.shape and not torch.equal(img, base)
    img[..., :3].sum().backward()
    assert two.grad.shape == two.shape and float(two.grad[0].abs().sum()) > 0 and float(two.grad[1].abs().sum()) > 0
    sc.close()


@pytest.mark.parametrize("schedule,res,spp", [("wavefront", (256, 256), 16), ("persistent", (256, 256), 16)])
def test_workspace_limit_external_buffer_and_trim(schedule, res, spp):
    """The path schedules' device workspace can be bounded, provided by the caller (a torch tensor: torch's allocator then
    accounts for it) or released; the image does not depend on how it was batched."""
    from zdr_b200 import Scene
    from zdr_b200._abi import ZdrError
    sc = Scene(scenes.cbox_path_models(), integrator="path")
    sc.camera = scenes.CBOX_CAMERA.copy()
    sc.max_depth = 4
    sc.schedule = schedule
    mat = torch.from_numpy(scenes.random_material((64, 64), 3)).cuda().requires_grad_()
    ref = sc.render(mat, res=res, spp=spp, seed=1)
    ref[..., :3].sum().backward()
    g_ref, mat.grad = mat.grad.clone(), None
    big = sc.workspace_bytes
    assert big > 0
    sc.trim()
    assert sc.workspace_bytes == 0
    sc.set_workspace_limit(24 << 20)                                   # forces several batches
    img = sc.render(mat, res=res, spp=spp, seed=1)
    img[..., :3].sum().backward()
    assert torch.equal(img, ref) and 0 < sc.workspace_bytes <= (24 << 20)
    assert rel_l2(mat.grad.cpu().numpy(), g_ref.cpu().numpy()) < 1e-5
    mat.grad = None
    sc.set_workspace_limit(0)
    buf = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
    sc.set_workspace(buf)
    img = sc.render(mat, res=res, spp=spp, seed=1)
    assert torch.equal(img, ref) and sc.workspace_bytes == buf.numel()
    sc.set_workspace(torch.empty(1 << 16, dtype=torch.uint8, device="cuda"))      # cannot hold one sample per pixel
    with pytest.raises(ZdrError):
        sc.render(mat, res=res, spp=spp, seed=1)
    sc.set_workspace(None)
    assert torch.equal(sc.render(mat, res=res, spp=spp, seed=1), ref)
    sc.close()
