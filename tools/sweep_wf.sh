#!/bin/bash
# GPU-box experiment: register caps of the wavefront shading kernels (A, B forward, B adjoint).
cd "$(dirname "$0")/.."
run() { python bench.py --spp 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --schedule wavefront | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])'; }
for cfg in "8 8 8" "8 8 6" "8 8 5" "8 8 4" "8 6 6" "12 8 6" "16 10 6" "8 8 3"; do
  set -- $cfg
  touch zdr_b200/csrc/render.cu
  make -C zdr_b200/csrc -s -j8 EXTRA="-DZDR_WF_SHADE_A_MINB=$1 -DZDR_WF_SHADE_B_MINB=$2 -DZDR_WF_SHADE_BB_MINB=$3" > /dev/null 2>&1
  echo "A=$1 B=$2 Bbwd=$3 -> $(run)"
done
