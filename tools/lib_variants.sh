#!/bin/bash
# GPU-box: C3 (64 spp and full) on the default build and on every prebuilt variant library.  usage: tools/lib_variants.sh <tag> [bench args]
cd "$(dirname "$0")/.."
R=${1:-r02y}; shift
mkdir -p gpurun_out
b() { timeout 300 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-fast-line 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
O=gpurun_out/${R}_variants.txt; : > $O
for rep in 1 2; do
for lib in default zdr_b200/csrc/variants/lib_*.so; do
  [ $lib == default ] || [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  if [ $lib != default ]; then export ZDR_B200_LIB=$PWD/$lib; else unset ZDR_B200_LIB; fi
  echo "$n C3@64 wavefront: $(b --spp 64 --schedule wavefront "$@")" | tee -a $O
  [ $rep == 1 ] && echo "$n C3 full: $(b "$@")" | tee -a $O
done
done
