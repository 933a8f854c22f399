#!/bin/bash
# GPU-box: load-shape variants (vector triangle loads, 256-bit node loads): parity subset + C3 at 64 spp on both big schedules + C5-shaped job.
cd "$(dirname "$0")/.."
R=${1:-r02i}
mkdir -p gpurun_out
b() { timeout 400 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
: > gpurun_out/${R}_variants.txt
for lib in default zdr_b200/csrc/variants/lib_*.so; do
  if [ "$lib" == default ]; then n=default; unset ZDR_B200_LIB; else n=$(basename $lib .so); n=${n#lib_}; export ZDR_B200_LIB=$PWD/$lib; fi
  echo "$n parity: $(timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k 'trace or c3_path_small or large or degenerate' 2>&1 | tail -1)" | tee -a gpurun_out/${R}_variants.txt
  echo "$n wavefront  C3@64: $(b --spp 64 --schedule wavefront)" | tee -a gpurun_out/${R}_variants.txt
  echo "$n persistent C3@64: $(b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_variants.txt
  echo "$n C5 1920x1080x8: $(b --config c5 --res 1920 1080 --spp 8)" | tee -a gpurun_out/${R}_variants.txt
done
