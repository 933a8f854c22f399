#!/bin/bash
# N-GPU box: bench.py under torchrun for c3 (HBM wavefront with 2 / 3 / 4 state sets, persistent schedule, fast build; every line with
# the multi_gpu_check), c5 and c2.  usage: tools/gpu8_session.sh <tag> <N>
cd "$(dirname "$0")/.."
R=${1:-r03f}; N=${2:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$T bench.py --gpus $N > gpurun_out/${R}_bench_c3_${N}gpu.json 2> gpurun_out/${R}_bench_${N}gpu.err; echo "rc=$?"
for L in 3 4; do ZDR_WF_LANES=$L $T bench.py --gpus $N --no-e2e > gpurun_out/${R}_bench_c3_${N}gpu_lanes$L.json 2>> gpurun_out/${R}_bench_${N}gpu.err; done
$T bench.py --gpus $N --math fast --no-e2e > gpurun_out/${R}_bench_c3_${N}gpu_fast.json 2>> gpurun_out/${R}_bench_${N}gpu.err
$T bench.py --gpus $N --schedule persistent --no-e2e > gpurun_out/${R}_bench_c3_${N}gpu_persistent.json 2>> gpurun_out/${R}_bench_${N}gpu.err
if [ -z "$SKIP_C5" ]; then timeout 1200 $T bench.py --gpus $N --config c5 --steps 2 --warmup 2 > gpurun_out/${R}_bench_c5_${N}gpu.json 2>> gpurun_out/${R}_bench_${N}gpu.err; fi
$T bench.py --gpus $N --config c2 > gpurun_out/${R}_bench_c2_${N}gpu.json 2>> gpurun_out/${R}_bench_${N}gpu.err
python - <<PY
import json
for f in ("c3_${N}gpu", "c3_${N}gpu_lanes3", "c3_${N}gpu_lanes4", "c3_${N}gpu_fast", "c3_${N}gpu_persistent", "c5_${N}gpu", "c2_${N}gpu"):
    try:
        d=json.loads(open("gpurun_out/${R}_bench_%s.json" % f).read().strip().splitlines()[-1])
        c=d.get("multi_gpu_check") or {}
        print(f, "value", d["value"], "ms", d["ms_per_step"], "allreduce", d["config"]["ms_allreduce"], "e2e", d.get("e2e",{}).get("value"), "check ok", c.get("ok"), "bits", c.get("image_bits_equal"), "grad", c.get("grad_rel_l2"))
    except Exception as e:
        print(f, "FAILED", e)
PY
