#!/bin/bash
# GPU-box experiment: clustered-SAH large-scene builder vs the plain LBVH on the C5-shaped scene.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lbvh_large or clustered_sah or envmap_path" > gpurun_out/csah_pytest.log 2>&1
tail -15 gpurun_out/csah_pytest.log
for b in ${BUILDERS:-lbvh csah}; do
  echo "== ZDR_BVH=$b"
  ZDR_BVH=$b timeout 600 python tools/run_c5.py --inst 10000 --res 1920 1080 --spp 8 --check-rays 48 --stats > gpurun_out/csah_$b.log 2>&1
  tail -4 gpurun_out/csah_$b.log
  ZDR_BVH=$b timeout 600 python tools/run_c5.py --inst 10000 --res 1920 1080 --spp 8 --check-rays 0 2>&1 | tail -2
done
