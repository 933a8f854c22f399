#!/bin/bash
# GPU-box experiment: register caps / grid size of the wavefront kernels.
cd "$(dirname "$0")/.."
run() { python bench.py --spp 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --schedule wavefront | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])'; }
for cfg in "8 8 32" "8 8 64" "10 8 32" "12 8 32" "16 8 32" "8 10 32" "10 16 64"; do
  set -- $cfg
  touch zdr_b200/csrc/render.cu
  make -C zdr_b200/csrc -s -j8 EXTRA="-DZDR_WF_SHADE_MINB=$1 -DZDR_WF_EXTEND_MINB=$2 -DZDR_WF_GRID_MULT=$3" > /dev/null 2>&1
  echo "shade_minb=$1 extend_minb=$2 grid_mult=$3 -> $(run)"
done
