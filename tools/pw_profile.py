"""Phase profile of the persistent schedule (ZDR_PW_PROFILE=1): share of SM cycles per phase for the forward and the
adjoint pass of a bench config.  usage: ZDR_PW_PROFILE=1 python tools/pw_profile.py [--config c3] [--spp 16]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ZDR_PW_PROFILE", "1")
import torch

from bench import scenes
from zdr_b200 import Scene
from zdr_b200._abi import INTEGRATORS, check, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--res", type=int, default=None)
args = ap.parse_args()
w = scenes.workload(args.config, spp=args.spp, res=(args.res, args.res) if args.res else None)
sc = Scene(w.models(), integrator=w.integrator)
if w.envmap() is not None:
    sc.add_envmap(w.envmap())
mat_np = w.material()
mat = torch.from_numpy(mat_np).cuda()
img = torch.zeros((w.res[1], w.res[0], 4), device="cuda")
dimg, dmat = torch.ones_like(img), torch.zeros_like(mat)
st = torch.cuda.current_stream().cuda_stream
names = ["claim+regen", "extend", "shade A", "shadow", "shade B"]
for rep in range(2):
    for which in ("forward", "adjoint"):
        p = make_params(w.camera, w.res, mat_np.shape, w.spp, 0 if which == "forward" else 1, max_depth=w.depth, schedule="persistent")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if which == "forward":
            check(sc._lib.zdr_render_forward(sc._handle, INTEGRATORS[w.integrator], C.byref(p), mat.data_ptr(), img.data_ptr(), st))
        else:
            check(sc._lib.zdr_render_backward(sc._handle, INTEGRATORS[w.integrator], C.byref(p), mat.data_ptr(), dimg.data_ptr(), dmat.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        out = (C.c_uint64 * 8)()
        check(sc._lib.zdr_debug_phase_cycles(sc._handle, out))
        v = [int(x) for x in out]
        if rep == 1:
            tot = max(sum(v[:5]), 1)
            print(f"{which}: {e0.elapsed_time(e1):.2f} ms; rounds per CTA {v[5] / 148:.0f}, live paths per round {v[6] / max(v[5], 1):.0f}, "
                  f"cycles per round {tot / max(v[5], 1):.0f}; " + ", ".join(f"{n} {100 * c / tot:.1f}%" for n, c in zip(names, v[:5])))
