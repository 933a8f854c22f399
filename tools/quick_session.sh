#!/bin/bash
# GPU-box: GPU test suite + the C3 bench line (exact, with its fast_math line) and the persistent schedule.  usage: tools/quick_session.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r02x}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
for c in ${CONFIGS:-c3 c2}; do
  python bench.py --config $c --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${R}_bench_$c.json 2>> gpurun_out/${R}_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${R}_bench_$c.json"))
f=d["config"].get("fast_math") or {}
print("$c", d["value"], d["ms_per_step"], "bwd", d["config"]["ms_backward_last_render"], "e2e", (d.get("e2e") or {}).get("value"), "fast", f.get("value"), f.get("image_rel_l2_vs_exact"), f.get("grad_rel_l2_vs_exact"))
PY
done
