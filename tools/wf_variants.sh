#!/bin/bash
# GPU-box: full GPU test suite on the default build, then the wavefront schedule on C3 at 64 spp for the default build and
# every prebuilt variant.  usage: tools/wf_variants.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r02h}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
b() { timeout 300 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
echo "default wavefront C3@64: $(b --spp 64 --schedule wavefront)" | tee gpurun_out/${R}_variants.txt
echo "default wavefront C3@64 (again): $(b --spp 64 --schedule wavefront)" | tee -a gpurun_out/${R}_variants.txt
for lib in zdr_b200/csrc/variants/lib_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  echo "variant $n wavefront C3@64: $(ZDR_B200_LIB=$PWD/$lib b --spp 64 --schedule wavefront)" | tee -a gpurun_out/${R}_variants.txt
  echo "variant $n persistent C3@64: $(ZDR_B200_LIB=$PWD/$lib b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_variants.txt
done
echo "default persistent C3@64: $(b --spp 64 --schedule persistent)" | tee -a gpurun_out/${R}_variants.txt
echo "default C5 960x540x8: $(b --config c5 --res 960 540 --spp 8)" | tee -a gpurun_out/${R}_variants.txt
for lib in zdr_b200/csrc/variants/lib_*.so; do
  n=$(basename $lib .so); n=${n#lib_}
  echo "variant $n C5 960x540x8: $(ZDR_B200_LIB=$PWD/$lib b --config c5 --res 960 540 --spp 8)" | tee -a gpurun_out/${R}_variants.txt
done
