#!/bin/bash
# N-GPU box, short: bench.py under torchrun for c3 and c5 (with the multi_gpu_check).  usage: tools/gpuN_short.sh <tag> <N>
cd "$(dirname "$0")/.."
R=${1:-r04b}; N=${2:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$T bench.py --gpus $N > gpurun_out/${R}_bench_c3_${N}gpu.json 2> gpurun_out/${R}_bench_${N}gpu.err; echo "rc=$?"
timeout 900 $T bench.py --gpus $N --config c5 --steps 2 --warmup 2 > gpurun_out/${R}_bench_c5_${N}gpu.json 2>> gpurun_out/${R}_bench_${N}gpu.err
python - <<PY
import json
for f in ("c3_${N}gpu", "c5_${N}gpu"):
    try:
        d=json.loads(open("gpurun_out/${R}_bench_%s.json" % f).read().strip().splitlines()[-1])
        c=d.get("multi_gpu_check") or {}
        print(f, "value", d["value"], "ms", d["ms_per_step"], "allreduce", d["config"]["ms_allreduce"], "e2e", d.get("e2e",{}).get("value"), "check ok", c.get("ok"), "bits", c.get("image_bits_equal"), "grad", c.get("grad_rel_l2"))
    except Exception as e:
        print(f, "FAILED", e)
PY
