"""C5-shaped job (BASELINE.json configs[4], scaled by flags): N instances of sphere.obj on a jittered lattice +
procedural envmap, path integrator with PRB, depth 16.  Reports BVH build time, Mrays/s, and checks a handful
of closest hits against the oracle's brute-force scan.

    python tools/run_c5.py --inst 10000 --res 960 540 --spp 16
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_c5.py --res 3840 2160 --spp 64          # the literal C5 shape: tile-sharded, NCCL all-reduce of the gradient
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import scenes
from zdr_b200 import Scene

ap = argparse.ArgumentParser()
ap.add_argument("--inst", type=int, default=10000)
ap.add_argument("--res", type=int, nargs=2, default=[960, 540])
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--depth", type=int, default=16)
ap.add_argument("--check-rays", type=int, default=48)
ap.add_argument("--stats", action="store_true", help="also count BVH nodes / triangles per ray (slower)")
ap.add_argument("--cpu-baseline", action="store_true",
                help="also time the CPU oracle (all host threads) on the same scene at 480x270, 2 of --spp samples, fwd + bwd (SURVEY 8d)")
ap.add_argument("--separate", action="store_true", help="keep every sphere a separate instance (10 000 add_mesh calls; only instance 0 is shaded)")
args = ap.parse_args()

from zdr_b200.load_obj import load_mesh

world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

# The reference textures instance 0 only and ends paths on every other instance (direct.py:30, prb.py:45), so a
# scene of 10 000 separate instances would shade almost nothing.  Here the non-emissive spheres are merged into ONE
# textured mesh (instance 0, ~9.5 M triangles); every 97th sphere stays a separate emissive instance.
grid = scenes.sphere_grid_models(args.inst, seed=0, emissive_every=97)
sv, st = load_mesh(grid[0][0])
mv, mt, lights = [], [], []
for i, (obj, xf, em) in enumerate(grid):
    if any(c > 0 for c in em):
        lights.append((obj, xf, em))
        continue
    m = np.asarray(xf.flat(), np.float32).reshape(4, 4).T                       # column-major -> rows
    v = sv.copy()
    v[:, 0:3] = sv[:, 0:3] @ m[:3, :3].T + m[:3, 3]
    mt.append(st + len(mv) * sv.shape[0])
    mv.append(v)
models = [((np.concatenate(mv), np.concatenate(mt)), None, scenes.float3(0.0))] + lights
if args.separate:
    models = grid
t0 = time.perf_counter()
sc = Scene(models, integrator="path", device=local_rank)
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
sc.camera = scenes.sphere_grid_camera(args.inst)
sc.max_depth = args.depth
sc.collect_stats = args.stats
t0 = time.perf_counter()
sc.add_envmap(scenes.procedural_envmap(2048, 1024, seed=0))
t_env = time.perf_counter() - t0
mat = torch.from_numpy(scenes.random_material((2048, 2048), 0)).cuda().requires_grad_()
res = tuple(args.res)
if world > 1:
    sc.distribute()                                  # replicated scene, round-robin 8-row tiles, all-reduce of image + gradient
for rep in range(2):
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sc.reset_counters()
    mat.grad = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    img = sc.render(mat, res=res, spp=args.spp, seed=0)
    ev[1].record()
    img[..., :3].sum().backward()
    ev[2].record()
    torch.cuda.synchronize()
c = sc.counters()
rays = c["closest_rays"] + c["any_rays"]
ms_f, ms_b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
if world > 1:                                        # whole-job rays / max-over-ranks device time
    t = torch.tensor([ms_f, ms_b], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    r = torch.tensor([float(rays), float(c["nodes_visited"]), float(c["tris_tested"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(r)
    ms_f, ms_b, rays = t[0].item(), t[1].item(), int(r[0].item())
    c["nodes_visited"], c["tris_tested"] = r[1].item(), r[2].item()
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
    print(f"{world} GPUs (one process each, NCCL all-reduce of image and gradient inside the timed region)")
print(f"instances {len(models)} triangles {sc.triangle_count:,}  scene+BVH build {t_build:.2f} s (upload + GPU flatten + GPU build, ZDR_BVH={os.environ.get("ZDR_BVH", "default")})  envmap import {t_env:.2f} s")
if args.stats:
    print(f"BVH work per ray: {c['nodes_visited'] / rays:.1f} nodes, {c['tris_tested'] / rays:.1f} triangles; closest {c['closest_rays']:,} any {c['any_rays']:,}")
print(f"{res[0]}x{res[1]} {args.spp} spp depth {args.depth}: fwd {ms_f:.1f} ms, bwd {ms_b:.1f} ms, {rays / 1e6:.1f} Mrays -> {rays / (ms_f + ms_b) / 1e3:.1f} Mrays/s; "
      f"image mean {float(img.detach()[..., :3].mean()):.4f} finite {bool(torch.isfinite(img).all())}, grad |sum| {float(mat.grad.abs().sum()):.3f} finite {bool(torch.isfinite(mat.grad).all())}")
if world > 1:
    dist.destroy_process_group()
if args.cpu_baseline and rank == 0:
    from oracle import oracle as o
    from zdr_b200._abi import ZDR_PATH, make_params
    o.build()
    o.lib().zo_set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    osc = o.OracleScene(models)                       # >= 4096 triangles: the oracle prunes its scan with a BVH (same hits)
    osc.set_envmap(scenes.procedural_envmap(2048, 1024, seed=0))
    t_scene = time.perf_counter() - t0
    mat_np = scenes.random_material((2048, 2048), 0)
    k = min(2, args.spp)
    pf = make_params(sc.camera, (480, 270), mat_np.shape, args.spp, 0, max_depth=args.depth, its=(0, k))
    pb = make_params(sc.camera, (480, 270), mat_np.shape, args.spp, 1, max_depth=args.depth, its=(0, k))
    osc.reset_ray_counters()
    t0 = time.perf_counter()
    osc.render_forward(ZDR_PATH, pf, mat_np)
    osc.render_backward(ZDR_PATH, pb, mat_np, np.ones((270, 480, 4), np.float32))
    dt = time.perf_counter() - t0
    cpu_rays = sum(osc.ray_counters())
    print(f"CPU oracle ({o.lib().zo_num_threads()} threads): scene + BVH + envmap {t_scene:.1f} s; 480x270, samples [0,{k}) of {args.spp}, fwd + bwd: "
          f"{cpu_rays / 1e6:.2f} Mrays in {dt:.1f} s -> {cpu_rays / dt / 1e6:.3f} Mrays/s; GPU / CPU = {rays / (ms_f + ms_b) / 1e3 / (cpu_rays / dt / 1e6):.0f}x")
if args.check_rays:
    import ctypes as C
    from oracle import oracle as o
    from zdr_b200 import _abi
    osc = o.OracleScene(models)
    rng = np.random.default_rng(1)
    cam = sc.camera
    org = np.array(list(cam.origin), np.float32)
    tgt = rng.uniform(0, min(21, args.inst ** (1 / 3)), (args.check_rays, 3)).astype(np.float32)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.concatenate([np.broadcast_to(org, d.shape), d, np.zeros((len(d), 1), np.float32), np.full((len(d), 1), 1e30, np.float32)], axis=1).astype(np.float32)
    ip, uvt = np.empty((len(r), 2), np.int32), np.empty((len(r), 3), np.float32)
    _abi.check(sc._lib.zdr_probe_trace_closest(sc._handle, _abi.np_ptr(r, C.c_float), len(r), _abi.np_ptr(ip, C.c_int32), _abi.np_ptr(uvt, C.c_float)))
    t0 = time.perf_counter()
    ip_o, uvt_o = osc.trace_closest(r)
    print(f"closest hits vs oracle brute force ({len(r)} rays, {time.perf_counter() - t0:.1f} s on the CPU): ids equal {np.array_equal(ip, ip_o)}, "
          f"t/u/v bit-equal {np.array_equal(uvt[ip_o[:, 0] >= 0].view(np.uint32), uvt_o[ip_o[:, 0] >= 0].view(np.uint32))}, hits {(ip_o[:, 0] >= 0).sum()}")
