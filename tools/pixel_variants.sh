#!/bin/bash
# GPU-box: occupancy of the per-pixel kernels (collocated / direct): C1, C2, C4, C4d for the default build and variants.
cd "$(dirname "$0")/.."
R=${1:-r02n}
mkdir -p gpurun_out
b() { timeout 400 python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms")
except Exception as e: print("FAILED", e)'; }
: > gpurun_out/${R}_variants.txt
for lib in default zdr_b200/csrc/variants/lib_*.so default; do
  if [ "$lib" == default ]; then n=default; unset ZDR_B200_LIB; else n=$(basename $lib .so); n=${n#lib_}; export ZDR_B200_LIB=$PWD/$lib; fi
  for c in c1 c2 c4 c4d; do echo "$n $c: $(b --config $c)" | tee -a gpurun_out/${R}_variants.txt; done
done
