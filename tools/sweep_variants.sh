#!/bin/bash
# GPU-box experiment: run prebuilt library variants (tools/build_variant.sh).  x*: C5-shaped scene; s*: C3 at 64 spp.
cd "$(dirname "$0")/.."
for lib in zdr_b200/csrc/variants/lib_*.so; do
  n=$(basename $lib .so); n=${n#lib_}
  if [[ $n == s* || $n == x0_b12 || $n == p* ]]; then
    echo "$n C3@64spp -> $(ZDR_B200_LIB=$PWD/$lib python bench.py --spp 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])')"
  fi
  if [[ $n == x* || $n == p* ]]; then
    echo "$n C5 -> $(ZDR_B200_LIB=$PWD/$lib python tools/run_c5.py --inst 10000 --res 1920 1080 --spp 16 --check-rays 0 2>&1 | tail -1 | cut -c1-100)"
  fi
done
