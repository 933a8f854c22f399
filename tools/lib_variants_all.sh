#!/bin/bash
# GPU-box: default build and every prebuilt variant on C3 (64 spp wavefront, full size, persistent, one-kernel), C2, C4, C4d.  usage: tools/lib_variants_all.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r03b}; shift
mkdir -p gpurun_out
b() { timeout 300 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-fast-line 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms")
except Exception as e: print("FAILED", e)'; }
O=gpurun_out/${R}_variants.txt; : > $O
for lib in default zdr_b200/csrc/variants/lib_*.so; do
  [ $lib == default ] || [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  if [ $lib != default ]; then export ZDR_B200_LIB=$PWD/$lib; else unset ZDR_B200_LIB; fi
  echo "$n C3@64 wavefront: $(b --spp 64 --schedule wavefront)" | tee -a $O
  echo "$n C3 full: $(b)" | tee -a $O
  echo "$n C3@64 persistent: $(b --spp 64 --schedule persistent)" | tee -a $O
  echo "$n C3 256^2@16 megakernel: $(b --res 256 --spp 16 --schedule megakernel)" | tee -a $O
  echo "$n C2: $(b --config c2)" | tee -a $O
  echo "$n C4: $(b --config c4)" | tee -a $O
  echo "$n C4d: $(b --config c4d)" | tee -a $O
  echo "$n C3@64 wavefront fast: $(b --spp 64 --schedule wavefront --math fast)" | tee -a $O
done
