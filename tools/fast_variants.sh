#!/bin/bash
# GPU-box: the "fast" shading builds (variants/lib_fast*.so) against the exact build: what stays bit-exact (sampler, hits), the
# deviation from the oracle (tests/fast_parity_report.py), and the bench lines.  usage: tools/fast_variants.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r02t}
mkdir -p gpurun_out
b() { timeout 300 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
O=gpurun_out/${R}_variants.txt; : > $O
for lib in default zdr_b200/csrc/variants/lib_fast*.so; do
  n=$(basename $lib .so); n=${n#lib_}
  if [ $lib != default ]; then export ZDR_B200_LIB=$PWD/$lib; else unset ZDR_B200_LIB; fi
  if [ $lib != default ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "sampler or trace or large or degenerate" > gpurun_out/${R}_pytest_exact_$n.log 2>&1; tail -1 gpurun_out/${R}_pytest_exact_$n.log | tee -a $O
    timeout 900 python tests/fast_parity_report.py gpurun_out/${R}_parity_$n.json > gpurun_out/${R}_parity_$n.log 2>&1; tail -8 gpurun_out/${R}_parity_$n.log | cut -c1-400
  fi
  echo "$n C3@64 wavefront: $(b --spp 64 --schedule wavefront)" | tee -a $O
  echo "$n C3@64 persistent: $(b --spp 64 --schedule persistent)" | tee -a $O
  echo "$n C3 full: $(b)" | tee -a $O
  echo "$n C2: $(b --config c2)" | tee -a $O
  echo "$n C4: $(b --config c4)" | tee -a $O
  echo "$n C5 960x540x8: $(b --config c5 --res 960 540 --spp 8)" | tee -a $O
done
