#!/bin/bash
# 2-GPU box: the multi-GPU tests, bench.py under torchrun with its multi_gpu_check, both big schedules.
cd "$(dirname "$0")/.."
R=${1:-r02k}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "two_gpu or another_device" > gpurun_out/${R}_pytest_multi.log 2>&1; tail -3 gpurun_out/${R}_pytest_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$T bench.py --gpus $N > gpurun_out/${R}_bench_c3_${N}gpu.json 2> gpurun_out/${R}_bench_${N}gpu.err; echo "rc=$?"; cut -c1-250 gpurun_out/${R}_bench_c3_${N}gpu.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_bench_c3_${N}gpu.json").read().strip().splitlines()[-1])
print("multi_gpu_check:", d.get("multi_gpu_check"))
print("value", d["value"], "ms", d["ms_per_step"], "allreduce", d["config"]["ms_allreduce"], "e2e", d["e2e"]["value"])
PY
$T bench.py --gpus $N --schedule persistent --no-e2e > gpurun_out/${R}_bench_c3_${N}gpu_persistent.json 2>> gpurun_out/${R}_bench_${N}gpu.err; cut -c1-200 gpurun_out/${R}_bench_c3_${N}gpu_persistent.json
$T bench.py --gpus $N --config c4 > gpurun_out/${R}_bench_c4_${N}gpu.json 2>> gpurun_out/${R}_bench_${N}gpu.err; cut -c1-200 gpurun_out/${R}_bench_c4_${N}gpu.json
