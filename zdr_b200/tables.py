"""Sampler tables for the pmj02bn sampler (``pmj02bn.py:1-16``).

The reference star-imports ``bluenoise.py`` and ``pmj02tables.py`` (pbrt-v4's
``BlueNoiseTextures[48][128][128]`` u16 and ``PMJ02bnSamples[5][65536][2]`` u32), but both files
are missing from the reference checkout (``.MISSING_LARGE_BLOBS``) and are optimisation outputs
that cannot be regenerated offline.  This module therefore

* loads the real tables when the user supplies them (``ZDR_TABLES_DIR`` pointing at a directory
  with ``pmj02tables.py`` + ``bluenoise.py`` using the pbrt-v4 symbol names, or with
  ``pmj02bn_tables.npz`` holding ``pmj`` and ``bn``), and otherwise
* generates deterministic SURROGATE tables of identical shape and dtype: five independently
  Owen-scrambled base-2 (0,2)-sequences (the stratification property PMJ02 tables have) and
  48 hashed 128x128 u16 offset textures (white, not blue, noise).

Bit-exactness claims of the sampler are "oracle == CUDA on identical tables"; equality with
upstream zdr additionally needs the real pbrt tables.
"""
from __future__ import annotations

import importlib.util
import os
import warnings
from functools import lru_cache
from typing import Tuple

import numpy as np

N_PMJ_SETS = 5
N_PMJ_SAMPLES = 65536
N_BN_TEXTURES = 48
BN_RESOLUTION = 128


def _reverse_bits32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x = ((x >> np.uint32(1)) & np.uint32(0x55555555)) | ((x & np.uint32(0x55555555)) << np.uint32(1))
    x = ((x >> np.uint32(2)) & np.uint32(0x33333333)) | ((x & np.uint32(0x33333333)) << np.uint32(2))
    x = ((x >> np.uint32(4)) & np.uint32(0x0F0F0F0F)) | ((x & np.uint32(0x0F0F0F0F)) << np.uint32(4))
    x = ((x >> np.uint32(8)) & np.uint32(0x00FF00FF)) | ((x & np.uint32(0x00FF00FF)) << np.uint32(8))
    return (x >> np.uint32(16)) | (x << np.uint32(16))


def _owen_scramble(x: np.ndarray, seed: int) -> np.ndarray:
    """Hash-based nested uniform scrambling (Laine-Karras style) of 32-bit fractions."""
    with np.errstate(over="ignore"):
        x = _reverse_bits32(x)
        x = x + np.uint32(seed)
        x ^= x * np.uint32(0x6C50B47C)
        x ^= x * np.uint32(0xB82F1E52)
        x ^= x * np.uint32(0xC7AFE638)
        x ^= x * np.uint32(0x8D22F6E6)
        return _reverse_bits32(x)


def _mix32(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x.astype(np.uint32)
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
        return x


def make_surrogate_tables() -> Tuple[np.ndarray, np.ndarray]:
    i = np.arange(N_PMJ_SAMPLES, dtype=np.uint32)
    vdc = _reverse_bits32(i)                                  # Sobol' dimension 0
    sob1 = np.zeros_like(i)                                   # Sobol' dimension 1
    v = np.uint32(1 << 31)
    for k in range(16):
        sob1 ^= np.where((i >> np.uint32(k)) & np.uint32(1), v, np.uint32(0)).astype(np.uint32)
        v = np.uint32(v ^ (v >> np.uint32(1)))
    pmj = np.empty((N_PMJ_SETS, N_PMJ_SAMPLES, 2), dtype=np.uint32)
    for s in range(N_PMJ_SETS):
        pmj[s, :, 0] = _owen_scramble(vdc, 0x9E3779B9 * (2 * s + 1) & 0xFFFFFFFF)
        pmj[s, :, 1] = _owen_scramble(sob1, 0x85EBCA6B * (2 * s + 2) & 0xFFFFFFFF)
    t, y, x = np.meshgrid(np.arange(N_BN_TEXTURES, dtype=np.uint32), np.arange(BN_RESOLUTION, dtype=np.uint32),
                          np.arange(BN_RESOLUTION, dtype=np.uint32), indexing="ij")
    with np.errstate(over="ignore"):
        h = _mix32(_mix32(t * np.uint32(0x9E3779B1) + np.uint32(0x7F4A7C15)) ^ (y * np.uint32(0x85EBCA77))
                   ^ _mix32(x * np.uint32(0xC2B2AE3D) + np.uint32(1)))
    bn = (h >> np.uint32(16)).astype(np.uint16)
    return pmj, bn


def _load_py_symbols(path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _load_user_tables(directory: str):
    npz = os.path.join(directory, "pmj02bn_tables.npz")
    if os.path.exists(npz):
        z = np.load(npz)
        return np.asarray(z["pmj"], dtype=np.uint32), np.asarray(z["bn"], dtype=np.uint16)
    p_py, b_py = os.path.join(directory, "pmj02tables.py"), os.path.join(directory, "bluenoise.py")
    if os.path.exists(p_py) and os.path.exists(b_py):
        pm = _load_py_symbols(p_py, "_zdr_user_pmj02tables")
        bm = _load_py_symbols(b_py, "_zdr_user_bluenoise")
        pmj = np.asarray(pm.PMJ02bnSamples, dtype=np.uint32)
        bn = np.asarray(bm.BlueNoiseTextures, dtype=np.uint16)
        assert pmj.shape == (pm.nPMJ02bnSets, pm.nPMJ02bnSamples, 2)          # pmj02bn.py:10
        assert bn.shape == (bm.NumBlueNoiseTextures, bm.BlueNoiseResolution, bm.BlueNoiseResolution)  # :15
        return pmj, bn
    return None


@lru_cache(maxsize=1)
def get_tables() -> Tuple[np.ndarray, np.ndarray, str]:
    """-> (pmj uint32 [sets, n, 2], bn uint16 [ntex, res, res], provenance)."""
    d = os.environ.get("ZDR_TABLES_DIR")
    if d:
        got = _load_user_tables(d)
        if got is None:
            raise FileNotFoundError(f"ZDR_TABLES_DIR={d}: no pmj02bn_tables.npz or pmj02tables.py+bluenoise.py")
        return np.ascontiguousarray(got[0]), np.ascontiguousarray(got[1]), f"user:{d}"
    if os.environ.get("ZDR_TABLES_STRICT", "") not in ("", "0"):
        raise FileNotFoundError("ZDR_TABLES_STRICT is set and ZDR_TABLES_DIR is not: refusing the surrogate sampler tables")
    warnings.warn(
        "zdr_b200: pbrt-v4's PMJ02bn / blue-noise tables are not available (the reference's pmj02tables.py and bluenoise.py are "
        "missing blobs); using deterministic SURROGATE tables of the same shape.  Sample positions and noise spectrum differ from "
        "upstream zdr.  Point ZDR_TABLES_DIR at pmj02tables.py + bluenoise.py (or pmj02bn_tables.npz) for the real ones; "
        "ZDR_TABLES_STRICT=1 turns this warning into an error.", RuntimeWarning, stacklevel=2)
    pmj, bn = make_surrogate_tables()
    return np.ascontiguousarray(pmj), np.ascontiguousarray(bn), "surrogate"
