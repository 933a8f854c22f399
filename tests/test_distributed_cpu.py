"""world_size-2 gloo test of the multi-GPU host logic (zdr_b200/distributed.py) on CPU: each rank
renders its round-robin tile shard with the ORACLE as the backend, the collectives sum image and
texture gradient; the result must equal the single-process render — bit-identical for the image
(disjoint rows), to float-sum accuracy for the gradient.  The second axis (sample-index ranges) is
equal to float-sum accuracy in both."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zdr_b200._abi import ZDR_PATH, make_params
from zdr_b200.distributed import rows_of_shard, samples_of_shard, sharded_render

from . import scenes

RES, SPP, SEED, DEPTH = (40, 52), 4, 7, 5


def _backend():
    from oracle import oracle as o
    sc = o.OracleScene(scenes.cbox_path_models())

    def fwd(material, res, spp, seed, shard, its=None):
        p = make_params(scenes.CBOX_CAMERA, res, material.shape, spp, seed, max_depth=DEPTH, shard=shard, its=its)
        return torch.from_numpy(sc.render_forward(ZDR_PATH, p, material.numpy()))

    def bwd(grad_output, d_material, material, res, spp, seed, shard, its=None):
        p = make_params(scenes.CBOX_CAMERA, res, material.shape, spp, seed + 1, max_depth=DEPTH, shard=shard, its=its)   # render.py:196
        d_material += torch.from_numpy(sc.render_backward(ZDR_PATH, p, material.numpy(), grad_output.contiguous().numpy()))
    return fwd, bwd


def _worker(rank, world, port, out_dir, axis="tiles"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fwd, bwd = _backend()
        material = torch.from_numpy(scenes.random_material((64, 64), 1)).requires_grad_()
        img = sharded_render(material, fwd, bwd, res=RES, spp=SPP, seed=SEED, axis=axis)
        w = torch.from_numpy(np.random.default_rng(3).uniform(0.5, 1.5, img.shape).astype(np.float32))
        (img * w).sum().backward()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), img=img.detach().numpy(), grad=material.grad.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_rows_of_shard_partition():
    for h, n in ((52, 2), (1024, 8), (17, 3), (8, 4)):
        rows = [rows_of_shard(h, (i, n)) for i in range(n)]
        assert sorted(sum(rows, [])) == list(range(h))
        assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 8


@pytest.mark.timeout(300)
def test_two_rank_render_equals_single(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    fwd, bwd = _backend()
    material = torch.from_numpy(scenes.random_material((64, 64), 1)).requires_grad_()
    img = sharded_render(material, fwd, bwd, res=RES, spp=SPP, seed=SEED)           # not initialised: shard (0, 1)
    w = torch.from_numpy(np.random.default_rng(3).uniform(0.5, 1.5, img.shape).astype(np.float32))
    (img * w).sum().backward()
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["img"], img.detach().numpy())
        g = material.grad.numpy()
        assert np.linalg.norm(z["grad"] - g) <= 1e-6 * np.linalg.norm(g)


def test_samples_of_shard_partition():
    for spp, n in ((4, 2), (256, 8), (5, 3), (2, 4), (1, 2)):
        ranges = [samples_of_shard(spp, (i, n)) for i in range(n)]
        assert ranges[0][0] == 0 and ranges[-1][1] == spp
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sample_sharding_equals_single(tmp_path, oracle):
    """axis="samples": every rank renders all pixels for half of the sample indices; image and gradient are sums."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), "samples"), nprocs=world, join=True)
    fwd, bwd = _backend()
    material = torch.from_numpy(scenes.random_material((64, 64), 1)).requires_grad_()
    img = sharded_render(material, fwd, bwd, res=RES, spp=SPP, seed=SEED)
    w = torch.from_numpy(np.random.default_rng(3).uniform(0.5, 1.5, img.shape).astype(np.float32))
    (img * w).sum().backward()
    ref_img, ref_grad = img.detach().numpy(), material.grad.numpy()
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.linalg.norm(z["img"] - ref_img) <= 1e-6 * np.linalg.norm(ref_img)
        assert (z["img"][..., 3] == 1).all()
        assert np.linalg.norm(z["grad"] - ref_grad) <= 1e-6 * np.linalg.norm(ref_grad)
