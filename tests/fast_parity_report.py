#!/usr/bin/env python
"""Deviation of the "fast" build of the render kernels (ZDR_FLAG_FAST_MATH) from the CPU oracle on the parity cases, for the
cases' random material as it is (roughness down to 0.02: binary32 GGX is ill conditioned there) and with the roughness clamped
from below: per case image rel-L2 (bar 1e-4), gradient rel-L2 (bar 1e-3), fraction of per-sample radiance values that differ
bitwise / by more than 1e-3, ray counts.
usage (GPU box): python tests/fast_parity_report.py out.json [case ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # repo root
from oracle import oracle  # noqa: E402
from tests import scenes  # noqa: E402
from tests.test_gpu_parity import CASES, GpuScene, bits, rel_l2  # noqa: E402
from zdr_b200._abi import make_params  # noqa: E402


def main(out, names):
    oracle.build(); oracle.lib()
    res = {"cases": {}}
    for name, rmin in [(n, r) for n in names for r in (0.0, 0.15, 0.3)]:
        models_fn, integ, size, spp, mat_fn, kw = CASES[name]
        kw = {k: v for k, v in kw.items() if k != "env"}
        models, mat = models_fn(), mat_fn()
        mat[..., 3] = np.maximum(mat[..., 3], rmin)
        name = f"{name} roughness>={rmin}"
        g, o = GpuScene(models), oracle.OracleScene(models)
        p = make_params(scenes.CBOX_CAMERA, size, mat.shape, spp, 0, math="fast", **kw)
        s_g, s_o = g.samples(integ, p, mat), o.render_samples(integ, p, mat)
        img_g, img_o = g.forward(integ, p, mat), o.render_forward(integ, p, mat)
        pb = make_params(scenes.CBOX_CAMERA, size, mat.shape, spp, 1, math="fast", **kw)
        dimg = np.random.default_rng(11).normal(size=(size[1], size[0], 4)).astype(np.float32)
        g.reset_counters(); o.reset_ray_counters()
        dm_g, dm_o = g.backward(integ, pb, mat, dimg), o.render_backward(integ, pb, mat, dimg)
        cg, (oc, oa) = g.counters(), o.ray_counters()
        fin = np.isfinite(s_o) & np.isfinite(s_g)
        rel = np.abs(s_g - s_o)[fin] / np.maximum(np.abs(s_o[fin]), 1e-6)
        res["cases"][name] = {
            "image_rel_l2": rel_l2(img_g, img_o), "grad_rel_l2": rel_l2(dm_g, dm_o),
            "samples_differing_bitwise": float((bits(s_g) != bits(s_o)).mean()),
            "samples_rel_err_gt_1e-3": float((rel > 1e-3).mean()), "samples_rel_err_median": float(np.median(rel)),
            "rays_gpu": [cg["closest_rays"], cg["any_rays"]], "rays_oracle": [oc, oa]}
        print(name, json.dumps(res["cases"][name]), flush=True)
        g.close()
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:] or ["c1_collocated", "c2_direct_small", "c3_path_small", "c3_path_small_wavefront", "path_literal_wavefront",
                                       "path_depth16_persistent"])
