"""PyTorch surface of the B200 ray-tracer core: ``Scene`` / ``Camera`` / ``RenderOperator``.

Drop-in for the reference's ``render.py`` (``Scene`` at ``render.py:31-257``): same constructor,
same mutable ``camera`` / ``use_tent_filter`` attributes, same ``render`` / ``update_lights`` /
``add_envmap`` methods, gradients through ``Scene.RenderOperator`` (``render.py:201-223``).
Instead of LuisaCompute kernels the operator launches the hand-written sm_100a kernels of
``csrc/`` through the C ABI of ``include/zdr_b200.h``; tensors stay torch-owned, the library
borrows their device pointers and enqueues on torch's current stream (no device-wide barrier,
where the reference does two per call at ``render.py:165,172``).

There is no CPU fallback: constructing a ``Scene`` without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence

import numpy as np
import torch

from . import _abi
from ._abi import INTEGRATORS, PRB_MODES, ZdrError, check, load_library, make_params, np_ptr, wrap_i32
from .load_obj import load_mesh
from .mathtypes import Camera, float3, float4x4
from .tables import get_tables

_tables_on_device = set()


def _ensure_tables(device_index: int) -> str:
    """Uploads the pmj02bn tables once per device (the module-level uploads of pmj02bn.py:9-16).
    Returns their provenance: ``'user:<dir>'`` (real pbrt-v4 tables) or ``'surrogate'`` (tables.py)."""
    pmj, bn, provenance = get_tables()
    if device_index in _tables_on_device:
        return provenance
    lib = load_library()
    check(lib.zdr_set_sampler_tables(device_index, np_ptr(pmj, C.c_uint32), pmj.shape[0], pmj.shape[1],
                                     np_ptr(bn, C.c_uint16), bn.shape[0], bn.shape[1]), "zdr_set_sampler_tables")
    _tables_on_device.add(device_index)
    return provenance


def _aligned16(t: torch.Tensor) -> torch.Tensor:
    """The kernels read texels / cotangents as 16-byte vectors and scatter with red.global.add.v4.f32: a
    contiguous view whose storage offset is not a multiple of 4 floats (a slice of a flat parameter vector) would
    fault.  Such a view is copied once; torch allocations themselves are 512-byte aligned."""
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def _emission3(e) -> float3:
    """render.py:85-88 / :142: None -> 0, scalar -> broadcast."""
    if e is None:
        return float3(0.0)
    if type(e) in {int, float}:
        return float3(e)
    return float3(e)


class Scene:
    """A 3D scene for differentiable rendering (reference: ``render.py:31``).

    Args:
        models: list of ``(obj_file, transform, emission)``; ``transform`` a ``float4x4`` or None,
            ``emission`` a ``float3`` / number / None (``render.py:81-91``).  Instance 0 is the
            textured object, every instance with a positive emission is a light.
            ``obj_file`` may also be a ``(verts Nx8, tris)`` pair of arrays.
        integrator: ``'path'`` | ``'direct'`` | ``'collocated'`` (``render.py:65-69``).
        device: CUDA device index (default: torch's current device).

    Attributes mirroring the reference's JIT-time module constants, settable per scene:
        ``max_depth``/``rr_depth`` (``prb.py:15-16``), ``use_mis`` (``direct.py:14``),
        ``bilinear`` (``interaction.py:33``), ``prb_mode`` (``'corrected'`` | ``'literal'``,
        DESIGN.md "PRB adjoint").
    """

    def __init__(self, models, integrator: str = "direct", device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("zdr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if integrator not in INTEGRATORS:
            raise KeyError(integrator)
        self._lib = load_library()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        #: 'user:<dir>' when ZDR_TABLES_DIR supplied pbrt-v4's tables, 'surrogate' otherwise (tables.py warns once)
        self.sampler_tables_provenance = _ensure_tables(self.device_index)
        self._handle = C.c_void_p()
        check(self._lib.zdr_scene_create(C.byref(self._handle), self.device_index), "zdr_scene_create")
        self.loaded_obj = {}
        self.load_geometry(models)
        self.camera = Camera(fov=40 / 180 * 3.1415926, origin=float3(1.0, 0.5, 0.0), target=float3(0.0, 0.0, 0.0),
                             up=float3(0.0, 1.0, 0.0))                                 # render.py:57-62
        self.integrator = integrator
        self.use_tent_filter = True                                                    # render.py:71
        self.max_depth = 16
        self.rr_depth = 2
        self.use_mis = True
        self.bilinear = True
        self.prb_mode = "corrected"
        self.collect_stats = False
        self.schedule = "auto"          # path integrator: 'auto' | 'wavefront' | 'megakernel' (ZDR_FLAG_*)
        self.scatter_mode = "auto"      # 'auto' | 'aggregate' | 'plain' (ZDR_FLAG_SCATTER_*)
        # 'exact': every operation is the IEEE operation of the numeric contract, images bit-identical to the CPU oracle (default);
        # 'fast': approximate division / sqrt / sin / cos + contraction in the shading arithmetic (ZDR_FLAG_FAST_MATH): 1.1-1.4x
        # faster, sampler and intersections still bit-exact, images / gradients within the tolerances on well-conditioned materials
        self.math = "exact"
        # light.py:89 leaves the envmap out of sample_light_pdf's light count; True uses sample_light's count (light.py:26)
        self.light_pdf_counts_env = False
        # data-parallel shard of the dispatch grid; None = whole image
        self.shard_rows = None
        self.shard = None          # (index, count): round-robin 8-row tiles, see zdr_params.shard_index
        self.shard_samples = None  # (first, last + 1): sample indices of this launch, see zdr_params.it0 / it1
        self._dist_group = None
        self._dist_axis = "tiles"

    # ------------------------------------------------------------------ scene assembly
    def load_geometry(self, models):
        """render.py:73-128.  As upstream, every call starts from an empty accel / heap (render.py:74-76):
        a second call REPLACES the instances."""
        lib = self._lib
        check(lib.zdr_scene_clear(self._handle), "zdr_scene_clear")
        self.loaded_obj = {}                                                           # render.py:76
        emissions = []
        for obj_file, transform, emission in models:
            if transform is None:
                transform = float4x4(1.0)
            emission = _emission3(emission)
            emissions.append(emission)
            if isinstance(obj_file, str):
                if obj_file not in self.loaded_obj:
                    self.loaded_obj[obj_file] = load_mesh(obj_file)
                verts, tris = self.loaded_obj[obj_file]
            else:
                verts, tris = obj_file
            verts = np.ascontiguousarray(verts, dtype=np.float32)
            tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1)
            xf = np.asarray(transform.flat(), dtype=np.float32)
            em = np.asarray(list(emission), dtype=np.float32)
            check(lib.zdr_scene_add_mesh(self._handle, np_ptr(verts, C.c_float), verts.shape[0], np_ptr(tris, C.c_int32),
                                         tris.shape[0] // 3, np_ptr(xf, C.c_float), np_ptr(em, C.c_float)),
                  "zdr_scene_add_mesh")
        self.inst_count = len(emissions)
        if self.inst_count > 10000:
            raise RuntimeError("exceeding maximum number of mesh instances")          # render.py:114-115
        check(lib.zdr_scene_commit(self._handle), "zdr_scene_commit")
        self.env_count = 0
        self.emissions = emissions
        self.light_count = lib.zdr_scene_light_count(self._handle)
        self.material_sets = [0] + [-1] * (self.inst_count - 1)        # direct.py:30: only the first object is textured

    def set_material_sets(self, sets):
        """Several texture sets ("polymorphic materials", listed as not implemented upstream, README.md:5).  ``sets[i]`` is
        the texture set that shades instance ``i`` or ``-1`` for a light / terminal surface; ``render`` then takes a
        ``(K, Ht, Wt, 4)`` material tensor and returns gradients of the same shape.  The default ``[0, -1, -1, ...]`` is the
        reference's rule "treat everything as light other than the first object" (``direct.py:30``, ``prb.py:45``)."""
        sets = [int(x) for x in sets]
        assert len(sets) == self.inst_count
        a = np.asarray(sets, dtype=np.int32)
        check(self._lib.zdr_scene_set_material_sets(self._handle, np_ptr(a, C.c_int32), a.shape[0]), "zdr_scene_set_material_sets")
        self.material_sets = sets

    def update_lights(self, emissions):
        """Rewrite the emission of every instance (``render.py:130-148``)."""
        self.emissions = emissions
        assert len(emissions) == self.inst_count
        e = np.asarray([list(_emission3(x)) for x in emissions], dtype=np.float32).reshape(-1, 3)
        check(self._lib.zdr_scene_set_emissions(self._handle, np_ptr(e, C.c_float), e.shape[0]), "zdr_scene_set_emissions")
        self.light_count = self._lib.zdr_scene_light_count(self._handle)

    def add_envmap(self, filename, compensate_mis: bool = True):
        """``render.py:150-156``.  ``filename``: image path (needs ``imageio``), ``.npy`` path, or
        an ``(H,W,3|4)`` float array."""
        if isinstance(filename, str):
            if filename.endswith(".npy"):
                img = np.load(filename)
            else:
                import imageio                                                      # render.py:151
                img = np.asarray(imageio.imread(filename))
        else:
            img = np.asarray(filename)
        if img.shape[2] == 3:
            img = np.concatenate([img, np.ones_like(img[..., :1])], axis=-1)
        img = np.ascontiguousarray(img, dtype=np.float32)
        check(self._lib.zdr_scene_set_envmap(self._handle, np_ptr(img, C.c_float), img.shape[1], img.shape[0],
                                             int(bool(compensate_mis))), "zdr_scene_set_envmap")
        self.env_count = 1

    # ------------------------------------------------------------------ operator launches
    def _params(self, material, res, spp, seed, rows=None):
        tex_shape = material.shape if material.ndim == 4 else material.shape[0:2]      # (K, Ht, Wt, 4): K texture sets
        p = make_params(self.camera, res, tex_shape, spp, seed, use_tent_filter=self.use_tent_filter,
                        max_depth=self.max_depth, rr_depth=self.rr_depth, use_mis=self.use_mis, bilinear=self.bilinear,
                        prb_mode=PRB_MODES[self.prb_mode] if isinstance(self.prb_mode, str) else int(self.prb_mode),
                        rows=rows if rows is not None else self.shard_rows, its=self.shard_samples, shard=self.shard, stats=self.collect_stats, scatter=self.scatter_mode, schedule=self.schedule,
                        light_pdf_counts_env=self.light_pdf_counts_env, math=self.math)
        return p

    def _check_material(self, material):
        # render.py:160: (Ht, Wt, 4); with material sets (set_material_sets) also (K, Ht, Wt, 4)
        assert (material.ndim == 3 and material.shape[2] == 4) or (material.ndim == 4 and material.shape[3] == 4)
        if (material.shape[0] if material.ndim == 4 else 1) <= max(self.material_sets):
            raise ValueError(f"instance material sets {self.material_sets} need {max(self.material_sets) + 1} texture sets, "
                             f"the material tensor {tuple(material.shape)} holds {material.shape[0] if material.ndim == 4 else 1}")
        if not material.is_cuda or material.device.index != self.device_index:
            raise ValueError(f"material must live on cuda:{self.device_index}")
        if material.dtype != torch.float32:
            raise TypeError("material must be float32")

    def render_forward(self, material, res, spp, seed, integrator: Optional[str] = None, rows=None):
        """``render.py:159-173``: -> ``(H, W, 4)`` float32 cuda tensor, alpha = 1."""
        self._check_material(material)
        material = _aligned16(material.detach().contiguous())
        image = torch.empty((res[1], res[0], 4), dtype=torch.float32, device=self.device)
        p = self._params(material, res, spp, seed, rows)
        if p.row0 > 0 or p.row1 < p.height or p.shard_count > 1 or p.it1 <= p.it0:
            image.zero_()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(self._lib.zdr_render_forward(self._handle, INTEGRATORS[integrator or self.integrator], C.byref(p),
                                           material.data_ptr(), image.data_ptr(), stream), "zdr_render_forward")
        return image

    def render_backward(self, grad_output, d_material, material, res, spp, seed, integrator: Optional[str] = None,
                        rows=None):
        """``render.py:176-199``: accumulates into ``d_material``; the adjoint is estimated with
        ``seed+1`` (``render.py:196``)."""
        self._check_material(material)
        material = _aligned16(material.detach().contiguous())
        d_image = grad_output.reshape((res[0] * res[1], 4)).contiguous()              # render.py:181
        if d_image.dtype != torch.float32:
            d_image = d_image.float()
        d_image = _aligned16(d_image)
        assert d_material.is_contiguous() and d_material.shape == material.shape
        if d_material.data_ptr() % 16 != 0:
            raise ValueError("d_material must be 16-byte aligned (gradients are accumulated with 16-byte vector reductions); "
                             "pass a tensor of its own, not a view at an odd offset")
        p = self._params(material, res, spp, wrap_i32(int(seed) + 1), rows)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(self._lib.zdr_render_backward(self._handle, INTEGRATORS[integrator or self.integrator], C.byref(p),
                                            material.data_ptr(), d_image.data_ptr(), d_material.data_ptr(), stream),
              "zdr_render_backward")
        return d_material, None, None, None, None

    class RenderOperator(torch.autograd.Function):
        """``render.py:201-223``."""

        @staticmethod
        def forward(ctx, material, self, *args):
            ctx.save_for_backward(material)
            ctx.scene = weakref.ref(self)
            ctx.args = args
            ctx.camera = self.camera.copy()
            ctx.emissions = list(self.emissions)
            return self.render_forward(material.detach(), *args)

        @staticmethod
        def backward(ctx, grad_output):
            scene = ctx.scene()
            if scene is None:
                raise RuntimeError("the Scene of this render was garbage-collected before backward()")
            camera_bak = scene.camera                                                 # render.py:216-218
            scene.camera = ctx.camera
            scene.update_lights(ctx.emissions)
            material, = ctx.saved_tensors
            mat_grad = torch.zeros(material.size(), dtype=material.dtype, device=material.device)
            try:
                grads = scene.render_backward(grad_output, mat_grad, material.detach(), *ctx.args)
            finally:
                scene.camera = camera_bak
            return grads

    def render(self, material, *, res, spp, seed=0):
        """Renders the scene; differentiable w.r.t. ``material`` (``render.py:225-241``).

        material: cuda float32 ``(Ht, Wt, 4)`` = (diffuse rgb, roughness); res = (width, height).
        Returns a ``(height, width, 4)`` tensor.  After :meth:`distribute` the call is collective:
        every rank passes the same arguments and gets the full image and the full gradient."""
        if self._dist_group is not None:
            return self._render_distributed(material, res, spp, seed)
        return Scene.RenderOperator.apply(material, self, res, spp, seed)

    def render_duvdxy(self, material, *, res, spp, seed=0):
        """Gradient of the texture coordinates w.r.t. screen-space coordinates (``render.py:243-257``):
        ``(height, width, 4)`` tensor of (dudx, dvdx, dudy, dvdy).  ``material`` only fixes the device, as in
        the reference.  Not differentiable."""
        self._check_material(material)
        image = torch.zeros((res[1], res[0], 4), dtype=torch.float32, device=self.device)
        p = self._params(material, res, spp, seed)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(self._lib.zdr_render_duvdxy(self._handle, C.byref(p), image.data_ptr(), stream), "zdr_render_duvdxy")
        return image

    # ------------------------------------------------------------------ multi-GPU (distributed.py)
    def distribute(self, group=None, axis: str = "tiles"):
        """Shard every later :meth:`render` over the ranks of ``group`` (default: WORLD): replicated
        scene, NCCL all-reduce of image and texture gradient.  ``axis="tiles"``: round-robin 8-row tiles (image
        bit-identical to one GPU); ``axis="samples"``: every rank renders all pixels for its range of sample indices
        (for images too small to tile; image equal to float accuracy)."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if axis not in ("tiles", "samples"):
            raise ValueError(f"unknown sharding axis {axis!r}")
        self._dist_group = group if group is not None else dist.group.WORLD
        self._dist_axis = axis
        return self

    def _render_distributed(self, material, res, spp, seed):
        from .distributed import sharded_render
        camera, emissions = self.camera.copy(), list(self.emissions)
        scene_ref = weakref.ref(self)

        def fwd(mat, res_, spp_, seed_, shard, its=None):
            sc = scene_ref()
            old, sc.shard, sc.shard_samples = (sc.shard, sc.shard_samples), shard, its
            try:
                return sc.render_forward(mat, res_, spp_, seed_)
            finally:
                sc.shard, sc.shard_samples = old

        def bwd(grad_output, d_material, mat, res_, spp_, seed_, shard, its=None):
            sc = scene_ref()
            old, old_cam = (sc.shard, sc.shard_samples), sc.camera
            sc.shard, sc.shard_samples, sc.camera = shard, its, camera       # render.py:216-218: state of the forward call
            sc.update_lights(emissions)
            try:
                sc.render_backward(grad_output, d_material, mat, res_, spp_, seed_)
            finally:
                (sc.shard, sc.shard_samples), sc.camera = old, old_cam
        return sharded_render(material, fwd, bwd, res=res, spp=spp, seed=seed, group=self._dist_group, axis=self._dist_axis)

    # ------------------------------------------------------------------ workspace of the path schedules
    def set_workspace_limit(self, nbytes: int):
        """Upper bound (bytes) of the scene's device workspace (wavefront path state / persistent sample buffer); 0 = the
        library's default policy (up to a quarter of the free memory).  Batches shrink to fit."""
        check(self._lib.zdr_scene_set_workspace_limit(self._handle, int(nbytes)), "zdr_scene_set_workspace_limit")

    def set_workspace(self, buffer: Optional[torch.Tensor]):
        """Use a torch tensor (any dtype, contiguous, on this scene's device) as the workspace instead of memory the library
        allocates with cudaMalloc behind torch's caching allocator; ``None`` goes back to the library's own allocation.  The
        scene keeps a reference to the tensor."""
        if buffer is None:
            check(self._lib.zdr_scene_set_workspace(self._handle, None, 0), "zdr_scene_set_workspace")
        else:
            assert buffer.is_cuda and buffer.device.index == self.device_index and buffer.is_contiguous()
            check(self._lib.zdr_scene_set_workspace(self._handle, buffer.data_ptr(), buffer.numel() * buffer.element_size()), "zdr_scene_set_workspace")
        self._workspace = buffer

    def trim(self):
        """Free the device scratch the library allocated for this scene (re-allocated on demand)."""
        check(self._lib.zdr_scene_trim(self._handle), "zdr_scene_trim")

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.zdr_scene_workspace_bytes(self._handle))

    # ------------------------------------------------------------------ instrumentation
    def counters(self) -> dict:
        out = (C.c_uint64 * _abi.ZDR_NUM_COUNTERS)()
        check(self._lib.zdr_get_counters(self._handle, out), "zdr_get_counters")
        return dict(zip(_abi.COUNTER_NAMES, (int(x) for x in out)))

    def reset_counters(self):
        check(self._lib.zdr_reset_counters(self._handle), "zdr_reset_counters")

    @property
    def triangle_count(self) -> int:
        return self._lib.zdr_scene_triangle_count(self._handle)

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.zdr_scene_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
