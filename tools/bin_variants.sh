#!/bin/bash
# GPU-box: GPU test suite, then the HBM wavefront schedule on C3 (64 spp) and on the C5-shaped scene with the traversal
# switches of round 2b: shared-memory / register-stack traversal of small scenes (ZDR_WF_TINY_OFF), ray binning
# (ZDR_WF_BIN bit 0 extension rays, bit 1 shadow rays), and every prebuilt variant.  usage: tools/bin_variants.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r02q}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
b() { timeout 300 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms; launches", d.get("gpu_launches"))
except Exception as e: print("FAILED", e)'; }
O=gpurun_out/${R}_variants.txt; : > $O
for bin in 0 1 2 3; do
  echo "C3@64 wavefront tiny bin=$bin: $(ZDR_WF_BIN=$bin b --spp 64 --schedule wavefront)" | tee -a $O
done
for bin in 0 3; do
  echo "C3@64 wavefront general bin=$bin: $(ZDR_WF_TINY_OFF=1 ZDR_WF_BIN=$bin b --spp 64 --schedule wavefront)" | tee -a $O
done
for lib in zdr_b200/csrc/variants/lib_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  for bin in 0 3; do
    echo "C3@64 wavefront variant $n bin=$bin: $(ZDR_B200_LIB=$PWD/$lib ZDR_WF_BIN=$bin b --spp 64 --schedule wavefront)" | tee -a $O
  done
done
for bin in 0 1 2 3; do
  echo "C5 960x540x8 bin=$bin: $(ZDR_WF_BIN=$bin b --config c5 --res 960 540 --spp 8)" | tee -a $O
done
for bin in ${FULL_BINS:-0 3}; do
  echo "C3 full (256 spp) tiny bin=$bin: $(ZDR_WF_BIN=$bin b)" | tee -a $O
done
