#!/bin/bash
# GPU-box: staged ray fetch of the wavefront traversal: parity (wavefront cases) + C3 at 64 spp + C5-shaped job, default build and variants.
cd "$(dirname "$0")/.."
R=${1:-r02l}
mkdir -p gpurun_out
b() { timeout 400 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
: > gpurun_out/${R}_variants.txt
for lib in default zdr_b200/csrc/variants/lib_*.so; do
  if [ "$lib" == default ]; then n=default; unset ZDR_B200_LIB; else n=$(basename $lib .so); n=${n#lib_}; export ZDR_B200_LIB=$PWD/$lib; fi
  echo "$n parity: $(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k 'wavefront or c3_full or large or envmap' 2>&1 | tail -1)" | tee -a gpurun_out/${R}_variants.txt
  echo "$n wavefront C3@64: $(b --spp 64 --schedule wavefront)" | tee -a gpurun_out/${R}_variants.txt
  echo "$n wavefront C3 full: $(b --schedule wavefront)" | tee -a gpurun_out/${R}_variants.txt
  echo "$n C5 1920x1080x8: $(b --config c5 --res 1920 1080 --spp 8)" | tee -a gpurun_out/${R}_variants.txt
done
