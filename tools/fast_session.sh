#!/bin/bash
# GPU-box: GPU tests, deviation of the fast build from the oracle, bench lines (exact with its fast_math line).  usage: tools/fast_session.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r02u}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
timeout 1200 python tests/fast_parity_report.py gpurun_out/${R}_fast_parity.json > gpurun_out/${R}_fast_parity.log 2>&1; grep -v Warning gpurun_out/${R}_fast_parity.log | cut -c1-330 | tail -20
for c in c3 c2 c4 c4d c1; do
  python bench.py --config $c --no-cpu-baseline > gpurun_out/${R}_bench_$c.json 2>> gpurun_out/${R}_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${R}_bench_$c.json"))
print("$c", d["value"], d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "fast", d["config"].get("fast_math"))
PY
done
python bench.py --schedule persistent --no-cpu-baseline --no-e2e > gpurun_out/${R}_bench_c3_persistent.json 2>> gpurun_out/${R}_bench.err; cut -c1-200 gpurun_out/${R}_bench_c3_persistent.json
