#!/bin/bash
# GPU-box session for the persistent shared-memory schedule: parity subset, then C3 at 64 spp under the runtime knobs
# (pool size, refill thresholds, bulk copy) and for every prebuilt variant library.  usage: tools/pw_sweep.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "persistent or fd_arbitrates or graph_capturable" > gpurun_out/${R}_pytest_pw.log 2>&1; tail -5 gpurun_out/${R}_pytest_pw.log
b() { timeout 300 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; last bwd", d["config"]["ms_backward_last_render"], "ms; launches", d["gpu_launches"])
except Exception as e: print("FAILED", e)'; }
echo "wavefront  C3@64: $(b --spp 64 --schedule wavefront)" | tee gpurun_out/${R}_sweep.txt
echo "persistent C3@64 (defaults): $(b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
for pool in 768 1024 1280 1536; do
  echo "persistent pool=$pool: $(ZDR_PW_POOL=$pool b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
done
for rf in 8 14 26 32; do
  echo "persistent refill=$rf: $(ZDR_PW_REFILL=$rf b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
done
for rf in 0 8 14 26 32; do
  echo "persistent refill_any=$rf: $(ZDR_PW_REFILL_ANY=$rf b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
done
echo "persistent no bulk copy: $(ZDR_PW_NO_BULK_COPY=1 b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
echo "persistent scene in global: $(ZDR_PW_NO_SMEM_SCENE=1 b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
for lib in zdr_b200/csrc/variants/lib_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  echo "variant $n: $(ZDR_B200_LIB=$PWD/$lib b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
  for pool in 1024 1280; do echo "variant $n pool=$pool: $(ZDR_PW_POOL=$pool ZDR_B200_LIB=$PWD/$lib b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt; done
done
echo "persistent C3 full: $(b --schedule persistent)" | tee -a gpurun_out/${R}_sweep.txt
