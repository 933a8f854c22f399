#!/bin/bash
# GPU-box: fast-build variants: deviation from the oracle on two path cases + bench lines.  usage: tools/fast_session2.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r02v}
mkdir -p gpurun_out
for lib in default zdr_b200/csrc/variants/lib_fast*.so; do
  n=$(basename $lib .so); n=${n#lib_}
  if [ $lib != default ]; then export ZDR_B200_LIB=$PWD/$lib; else unset ZDR_B200_LIB; fi
  timeout 900 python tests/fast_parity_report.py gpurun_out/${R}_fast_parity_$n.json c3_path_small_wavefront path_literal_wavefront c1_collocated > gpurun_out/${R}_fast_parity_$n.log 2>&1
  grep -v Warning gpurun_out/${R}_fast_parity_$n.log | cut -c1-230 | tail -9
  for c in c3 c2; do
    python bench.py --config $c --no-cpu-baseline --no-e2e > gpurun_out/${R}_bench_${c}_$n.json 2>> gpurun_out/${R}_bench.err
    python - <<PY
import json
d=json.load(open("gpurun_out/${R}_bench_${c}_$n.json"))
f=d["config"].get("fast_math") or {}
print("$n $c exact", d["value"], "fast", f.get("value"), "image", f.get("image_rel_l2_vs_exact"), "grad", f.get("grad_rel_l2_vs_exact"))
PY
  done
done
