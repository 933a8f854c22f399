#!/bin/bash
# GPU-box experiment: occupancy cap of k_path (registers per thread) vs speed.
cd "$(dirname "$0")/.."
for mb in 3 4 5 6; do
  sed -i "s/__launch_bounds__(kBlockThreads, [0-9]) k_path/__launch_bounds__(kBlockThreads, $mb) k_path/" zdr_b200/csrc/render.cu
  make -C zdr_b200/csrc -s -j8 > /dev/null 2>&1
  echo "minblocks=$mb $(python bench.py --spp 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])')"
done
