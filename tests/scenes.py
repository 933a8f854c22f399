"""Synthetic workloads: defined in bench/scenes.py (shared by the tests, bench.py and smoke())."""
from bench.scenes import *  # noqa: F401,F403
from bench.scenes import ASSETS, CBOX_CAMERA, SPHERE_CAMERA  # noqa: F401
