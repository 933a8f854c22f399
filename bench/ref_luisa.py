#!/usr/bin/env python
"""bench/ref_luisa.py — the TRUE reference arm: upstream zdr on LuisaCompute's CPU backend.

    python bench/ref_luisa.py --zdr /path/to/zdr [--tables DIR] [--config c1|c2|c3] [--steps K] [--warmup W] [--ref-spp S]

Runs the reference's own `Scene` (render.py) — unmodified sources, imported from a scratch copy — through its public
API (`scene.render(...)`, `.backward()`), on the workload and seeds `bench.py` uses, and prints the same JSON line as
`bench.py --impl reference` with `"kind": "reference"`.  It needs what this build container does not have
(SURVEY.md §8c): the `luisa-python` package (LuisaCompute, un-pinned in environment.yml:12) and the two table
modules the checkout is missing (`pmj02tables.py`, `bluenoise.py`: pbrt-v4's PMJ02bn / blue-noise tables; pass their
directory with --tables or ZDR_TABLES_DIR).  Without them it prints `{"impl": "reference", "unavailable": "..."}`
and exits 0, as the bench contract asks.

What is redirected, and nothing else:
  * `luisa.init('cuda')` (render.py:26, pmj02bn.py:8) -> `luisa.init('cpu')`: the north star's comparison target is
    LuisaCompute's CPU backend on the host cores;
  * `tensor.cuda()` / `device='cuda'` / `torch.cuda.synchronize()` (pmj02bn.py:9,14; render.py:163,165,190) -> host
    tensors / no-ops when torch sees no GPU;
  * `prb.py`: the seven call-arity substitutions of SURVEY.md §8a defect 1 (the file does not JIT as shipped) — the
    same list tests/golden/gen_golden.py applies.  The estimator bodies are untouched, so the adjoint is the
    reference's literal weighting (`prb_mode = literal` on our side).
Rays are not counted by the reference; the line reports paths (camera samples) per second and, for the ratio against
`bench.py`, converts with the rays-per-sample figure of the CUDA path's counters (`--rays-per-sample`, default: the C3
figure of profiles/, 7.27 closest + any queries per forward + adjoint sample pair).

This file is the hook BASELINE.md promises; it cannot be exercised in the build container (no luisa), so treat it as
untested there: tests/test_host.py only checks the "unavailable" path.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC, UNIT = "prb_fwd_bwd_mrays_per_s", "Mrays/s"

PRB_PATCHES = [          # SURVEY.md §8a "PRB defects" item 1 — identical to tests/golden/gen_golden.py
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


def unavailable(why: str):
    print(json.dumps({"impl": "reference", "unavailable": why}))
    sys.exit(0)


def stage_package(zdr_dir: str, tables_dir: str | None, max_depth: int) -> str:
    """Copy of the reference package under a scratch directory (the checkout is read-only and misses two modules)."""
    stage = tempfile.mkdtemp(prefix="zdr_ref_")
    pkg = os.path.join(stage, "zdr")
    shutil.copytree(zdr_dir, pkg, ignore=shutil.ignore_patterns(".git", "__pycache__", "results"))
    for name in ("pmj02tables.py", "bluenoise.py"):
        if not os.path.exists(os.path.join(pkg, name)):
            src = os.path.join(tables_dir or "", name)
            if not (tables_dir and os.path.exists(src)):
                unavailable(f"{name} (pbrt-v4 sampler table, a missing blob of the reference checkout) not found; pass --tables DIR")
            shutil.copy(src, os.path.join(pkg, name))
    prb = os.path.join(pkg, "prb.py")
    src = open(prb).read()
    for a, b in PRB_PATCHES:
        if a in src:
            src = src.replace(a, b)
    src = src.replace("max_depth = 16", f"max_depth = {max_depth}")          # prb.py:15: a JIT-time module constant upstream
    open(prb, "w").write(src)
    return stage


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zdr", default=os.environ.get("ZDR_REFERENCE_DIR", "/root/reference"))
    ap.add_argument("--tables", default=os.environ.get("ZDR_TABLES_DIR"))
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3"])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--res", type=int, default=None)
    ap.add_argument("--ref-spp", type=int, default=2, help="spp of the bounded sample one step renders (rays/s does not depend on it)")
    ap.add_argument("--rays-per-sample", type=float, default=None)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    if int(os.environ.get("RANK", "0")) != 0:
        return
    try:
        import luisa
    except ImportError as e:
        unavailable(f"luisa-python (LuisaCompute) is not installed: {e}")
    import torch
    from bench import scenes
    w = scenes.workload(args.config, res=(args.res, args.res) if args.res else None)

    # ---- redirections (see the module docstring)
    _init = luisa.init
    luisa.init = lambda *a, **k: _init("cpu")
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.synchronize = lambda *a, **k: None
        _zeros, _empty = torch.zeros, torch.empty
        strip = lambda f: (lambda *a, **k: f(*a, **{kk: vv for kk, vv in k.items() if not (kk == "device" and str(vv).startswith("cuda"))}))
        torch.zeros, torch.empty = strip(_zeros), strip(_empty)
    stage = stage_package(args.zdr, args.tables, w.depth)
    sys.path.insert(0, stage)
    try:
        import zdr                                            # the reference package itself
    except Exception as e:                                    # noqa: BLE001 - anything LuisaCompute's JIT front-end raises
        unavailable(f"importing the reference failed: {type(e).__name__}: {e}")

    def to_ref(models):
        out = []
        for obj, xf, em in models:
            m = None if xf is None else luisa.float4x4(*xf.flat())
            out.append((obj, m, None if em is None else zdr.float3(*list(em))))
        return out
    scene = zdr.Scene(to_ref(w.models()), integrator=w.integrator)
    cam = w.camera
    scene.camera = zdr.Camera(fov=cam.fov, origin=zdr.float3(*list(cam.origin)), target=zdr.float3(*list(cam.target)), up=zdr.float3(*list(cam.up)))
    material = torch.from_numpy(w.material())
    if torch.cuda.is_available():
        material = material.cuda()
    spp = min(args.ref_spp, w.spp)
    times = []
    for i in range(args.warmup + args.steps):
        m = material.clone().requires_grad_()
        t0 = time.perf_counter()
        img = scene.render(m, res=w.res, spp=spp, seed=0)
        img.sum().backward()
        _ = float(m.grad.abs().sum())
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    samples = 2 * w.res[0] * w.res[1] * spp                   # forward + adjoint camera samples
    rps = args.rays_per_sample if args.rays_per_sample else {"c1": 1.0, "c2": 2.4, "c3": 7.27}[args.config]
    value = samples * rps / (ms * 1e-3) / 1e6
    sample = f"{w.res[0]}x{w.res[1]}, {spp} of {w.spp} spp, fwd+bwd through zdr.Scene.render / backward; rays = camera samples x {rps} (CUDA-path counters)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": w.name(), "sample_per_step": sample, "backend": "LuisaCompute cpu"},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


if __name__ == "__main__":
    main()
