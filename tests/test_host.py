"""Host-side logic and the C-ABI surface, no GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from zdr_b200 import _abi
from zdr_b200.load_obj import concat_triangles, load_mesh, read_obj
from zdr_b200.mathtypes import Camera, float3, float4x4
from zdr_b200.tables import get_tables, make_surrogate_tables

from . import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "zdr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zdr_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_ctypes_table():
    names = header_functions()
    assert len(names) >= 20
    assert sorted(_abi.SIGNATURES) == names


def test_library_loads_and_exports_every_symbol():
    lib = _abi.load_library()                     # raises if the in-tree .so is missing: no fallback
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.zdr_abi_version() == 1


def test_header_flags_match_the_python_constants():
    """Every ZDR_FLAG_* of the header has the same value in zdr_b200/_abi.py (the flags select schedules, scatter modes, the build)."""
    src = open(os.path.join(ROOT, "include", "zdr_b200.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(ZDR_FLAG_[A-Z_0-9]+)\s+(\d+)", src)}
    assert len(flags) >= 9 and "ZDR_FLAG_FAST_MATH" in flags
    for name, value in flags.items():
        assert getattr(_abi, name) == value, name
    assert len(set(flags.values())) == len(flags) and all(v & (v - 1) == 0 for v in flags.values())      # distinct single bits
    p = _abi.make_params(scenes.CBOX_CAMERA, (8, 8), (4, 4, 4), 1, 0, math="fast", schedule="wavefront")
    assert p.flags == _abi.ZDR_FLAG_FAST_MATH | _abi.ZDR_FLAG_WAVEFRONT
    with pytest.raises(KeyError):
        _abi.make_params(scenes.CBOX_CAMERA, (8, 8), (4, 4, 4), 1, 0, math="sloppy")


def test_params_struct_layout():
    """zdr_params is ABI: field order / size must match the header (27 x 4 bytes + 10 floats)."""
    src = open(os.path.join(ROOT, "include", "zdr_b200.h")).read()
    body = re.search(r"typedef struct zdr_params \{(.*?)\} zdr_params;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        for n in names.split(","):
            n = n.strip()
            m = re.match(r"(\w+)\[(\d+)\]", n)
            fields.append((m.group(1), int(m.group(2))) if m else (n, 1))
    ct = [(n, (t._length_ if hasattr(t, "_length_") else 1)) for n, t in _abi.ZdrParams._fields_]
    assert fields == ct
    assert C.sizeof(_abi.ZdrParams) == 4 * sum(k for _, k in fields)


def test_compute_entry_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.load_library()
    h = C.c_void_p()
    assert lib.zdr_scene_create(C.byref(h), 0) == -3            # ZDR_ERR_CUDA
    assert b"no CPU fallback" in lib.zdr_last_error()
    with pytest.raises(RuntimeError):
        from zdr_b200 import Scene
        Scene(scenes.cbox_collocated_models(), integrator="collocated")


def test_load_obj_matches_reference_loader():
    z = np.load(os.path.join(ROOT, "tests", "golden", "loadobj.npz"))
    for name in ("cboxuv.obj", "cbox-combined.obj", "cbox-light.obj", "quad.obj", "sphere.obj"):
        vertices, faces = read_obj(scenes.asset(name))
        verts = np.array([(*p, *t, *n) for p, t, n in vertices], dtype=np.float32)
        tris = np.array(concat_triangles(faces), dtype=np.int32)
        assert np.array_equal(tris, z[name + ":tris"]), name
        assert np.array_equal(verts, z[name + ":verts"], equal_nan=True), name
    # quad.obj has no vn: normals are recomputed at load (render.py:101-103); the fixture holds the output of the
    # reference's recompute_normal.py kernels run in program order (float atomics: any order is "the reference's")
    for key in [k for k in z.files if k.endswith(":verts_recomputed")]:
        name = key.split(":")[0]
        verts, _ = load_mesh(scenes.asset(name))
        ref = z[key]
        assert np.array_equal(verts[:, :5], ref[:, :5]) and np.allclose(verts[:, 5:8], ref[:, 5:8], rtol=0, atol=2e-7), name
    assert any(k.endswith(":verts_recomputed") for k in z.files)
    verts, tris = load_mesh(scenes.asset("quad.obj"))
    assert not np.isnan(verts).any() and np.allclose(np.linalg.norm(verts[:, 5:8], axis=1), 1.0)
    # cbox-combined.obj has no vt: uv == 0 (load_obj.py:48) — the scatter-contention case
    verts, _ = load_mesh(scenes.asset("cbox-combined.obj"))
    assert (verts[:, 3:5] == 0).all()


def test_native_obj_loader_equals_python_loader(tmp_path):
    """zdr_obj_load (C++) against load_mesh_python, the line-by-line statement of load_obj.py: identical bytes on the
    shipped assets and on a hostile file (n-gons, a/b/c - a//c - a/b - a corners, negative and zero indices, signed
    zeros, CRLF, comments, extra components, no normals -> recomputed normals incl. a zero-area triangle)."""
    from zdr_b200.load_obj import load_mesh_native, load_mesh_python
    files = [scenes.asset(n) for n in ("cboxuv.obj", "cbox-combined.obj", "cbox-light.obj", "sphere.obj", "quad.obj")]
    rng = np.random.default_rng(0)
    lines = ["# hostile", "", "o thing\r"]
    pts = rng.normal(size=(40, 3)).round(3)
    pts[3] = (0.0, -0.0, 1.0)
    pts[4] = (-0.0, 0.0, 1.0)                       # equal to pts[3] as numbers: merges with it
    for q in pts:
        lines.append(f"v {float(q[0])!r} {float(q[1])!r} {float(q[2])!r} 1.0")
    for q in rng.uniform(0, 1, (12, 2)).round(4):
        lines.append(f"vt {float(q[0])!r} {float(q[1])!r} 0.0\r")
    lines += ["f 1/1 2/2 3/3 4/4 5/5", "f 4/4 5/4 6/1", "f -1 -2 -3", "f 7 8 9 10 11 12 13", "f 0 1 2",
              "f 3 4 5", "f 3/2 4/2 5/2", "f 20/3 21/4 22/5\r", "f 30 30 31", "s off", "usemtl x", "f 33/1 34/1 35/1 36/1"]
    hostile = tmp_path / "hostile.obj"
    hostile.write_text("\n".join(lines) + "\n")
    with_normals = tmp_path / "normals.obj"
    with_normals.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\nvn 0 0 -1\nvt 0.5 0.5\n"
                            "f 1//1 2//1 3//1\nf 2/1/1 4/1/2 3//1\nf 1//1 2//1 4//2\n")
    for f in files + [str(hostile), str(with_normals)]:
        (v0, t0), (v1, t1) = load_mesh_python(f), load_mesh_native(f)
        assert v0.shape == v1.shape and np.array_equal(v0.view(np.uint32), v1.view(np.uint32)), f
        assert np.array_equal(t0, t1), f
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nf 1 2 9\n")
    with pytest.raises(ValueError):
        load_mesh_native(str(bad))
    with pytest.raises(ValueError):
        load_mesh_native(str(tmp_path / "missing.obj"))


def test_mathtypes():
    m = float4x4(1.0)
    assert m.flat() == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    t = float4x4(*range(16))
    assert list(t[1]) == [4.0, 5.0, 6.0, 7.0]                    # column-major, test_lightstage.py:44
    v = float3(1, 2, 3)
    assert list(v * 2 + float3(1.0)) == [3.0, 5.0, 7.0] and v.y == 2.0
    c = Camera()
    assert abs(c.fov - 40 / 180 * 3.1415926) < 1e-12 and list(c.origin) == [1.0, 0.5, 0.0]     # render.py:57-62
    assert _abi.wrap_i32(2147483647 + 1) == -2147483648                                           # seed+1 wrap


def test_sampler_tables_shape_and_provenance():
    pmj, bn, prov = get_tables()
    assert pmj.shape == (5, 65536, 2) and pmj.dtype == np.uint32                                  # pmj02bn.py:10
    assert bn.shape == (48, 128, 128) and bn.dtype == np.uint16                                   # pmj02bn.py:15
    assert prov in ("surrogate",) or prov.startswith("user:")
    p2, b2 = make_surrogate_tables()
    assert np.array_equal(p2, pmj) or prov != "surrogate"
    if prov == "surrogate":
        # the surrogate PMJ sets are (0,2)-sequences: every 16x16 elementary interval of the first 256 points holds one
        pts = pmj[0, :256].astype(np.float64) / 2 ** 32
        cells = (np.floor(pts[:, 0] * 16) * 16 + np.floor(pts[:, 1] * 16)).astype(int)
        assert len(set(cells.tolist())) == 256


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny workload: one JSON line with the
    contract's keys, and all host threads even when the launcher exported OMP_NUM_THREADS=1 (torchrun does)."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--res", "32", "--spp", "4",
                          "--ref-spp", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "prb_fwd_bwd_mrays_per_s" and d["unit"] == "Mrays/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank under torchrun prints nothing and exits 0
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_ref_luisa_hook_reports_unavailable_without_luisa():
    """bench/ref_luisa.py: upstream zdr on LuisaCompute's CPU backend where `luisa-python` and the pbrt tables
    exist; in this container it must say why it cannot run, as one JSON line, and exit 0."""
    import json
    import subprocess
    import sys
    try:
        import luisa  # noqa: F401
        pytest.skip("luisa-python is installed here")
    except ImportError:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench", "ref_luisa.py")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and "luisa" in d["unavailable"]


@pytest.mark.parametrize("script", ["benchmark.py", "lightstage.py", "fd_validate.py", "inverse_cornell.py"])
def test_example_scripts_import_and_parse_arguments(script):
    """examples/ = the reference's caller scripts (benchmark.py, test_lightstage.py, fd_validate.py, example.py) on this package.
    They need a GPU to run (profiles/r05b_examples.log); here: they import the package and get as far as their argument parser."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), "--help"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("usage: " + script), out.stderr[-2000:]
