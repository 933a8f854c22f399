#!/bin/bash
# GPU-box experiment: every prebuilt variant of the persistent kernel (tools/build_variant.sh, FILE=persistent) on C3 at 64 spp,
# with its largest pool and with smaller ones (more L1).  usage: tools/pw_variants.sh <tag> [pool fractions...]
cd "$(dirname "$0")/.."
R=${1:-r02d}; shift
mkdir -p gpurun_out
b() { timeout 300 python bench.py "$@" --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --schedule persistent 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
echo "default lib: $(b)" | tee gpurun_out/${R}_variants.txt
for lib in zdr_b200/csrc/variants/lib_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  echo "variant $n (max pool): $(ZDR_B200_LIB=$PWD/$lib b)" | tee -a gpurun_out/${R}_variants.txt
  for pool in ${POOLS:-512 768 1024}; do echo "variant $n pool=$pool: $(ZDR_PW_POOL=$pool ZDR_B200_LIB=$PWD/$lib b)" | tee -a gpurun_out/${R}_variants.txt; done
done
