#!/bin/bash
# GPU-box experiment: run prebuilt library variants (tools/build_variant.sh): C3 at 64 spp, C3 at 1/8 of the pixels
# (one rank's share of an 8-GPU run), the C5-shaped scene.
cd "$(dirname "$0")/.."
b() { python bench.py "$@" --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])'; }
for lib in zdr_b200/csrc/variants/lib_*.so; do
  n=$(basename $lib .so); n=${n#lib_}
  export ZDR_B200_LIB=$PWD/$lib
  echo "$n C3@64spp -> $(b --spp 64) | 362^2@256spp -> $(b --res 362 --spp 256)"
  if [ -z "$NO_C5" ]; then
    echo "$n C5 -> $(python tools/run_c5.py --inst 10000 --res 1920 1080 --spp 16 --check-rays 0 2>&1 | tail -1 | cut -c1-100)"
  fi
done
