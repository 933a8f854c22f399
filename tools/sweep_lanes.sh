#!/bin/bash
# GPU-box experiment: number of concurrent state sets / streams of the wavefront schedule (ZDR_WF_LANES).
cd "$(dirname "$0")/.."
b() { python bench.py "$@" --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])'; }
for l in ${LANES:-1 2 3 4}; do
  export ZDR_WF_LANES=$l
  echo "lanes=$l C3@64spp -> $(b --spp 64) | 362^2@256spp -> $(b --res 362 --spp 256) | C3 full -> $(b)"
done
