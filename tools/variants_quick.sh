#!/bin/bash
# GPU box: C3 at 64 spp (wavefront schedule) and at full size on the default build and on the named variant libraries, two rounds.
# usage: tools/variants_quick.sh <tag> name...
cd "$(dirname "$0")/.."
R=$1; shift
mkdir -p gpurun_out
O=gpurun_out/${R}_variants.txt; : > $O
b() { timeout 120 python bench.py "$@" --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-line 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
for rep in 1; do
for n in default "$@"; do
  if [ $n != default ]; then export ZDR_B200_LIB=$PWD/zdr_b200/csrc/variants/lib_$n.so; else unset ZDR_B200_LIB; fi
  echo "$n C3@64: $(b --spp 64 --schedule wavefront)" | tee -a $O
done
done
for n in default "$@"; do
  if [ $n != default ]; then export ZDR_B200_LIB=$PWD/zdr_b200/csrc/variants/lib_$n.so; else unset ZDR_B200_LIB; fi
  echo "$n C3 full: $(b)" | tee -a $O
done
