#!/bin/bash
# GPU-box experiment: refill threshold of the persistent traversal kernel.
cd "$(dirname "$0")/.."
run() { python bench.py --spp 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])'; }
for r in 0 12 16 20 24 28; do
  touch zdr_b200/csrc/render.cu
  make -C zdr_b200/csrc -s -j8 EXTRA="-DZDR_WF_REFILL=$r" > /dev/null 2>&1
  echo "refill=$r -> $(run)"
done
