#!/bin/bash
# GPU-box session producing the round's evidence: tests, smoke, bench (both arms), launch lists, --set full capture.
cd "$(dirname "$0")/.."
R=${1:-r01t}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; tail -2 gpurun_out/${R}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; cut -c1-200 gpurun_out/${R}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err; cut -c1-200 gpurun_out/${R}_bench_reference.json
# launch list of the default bench command (per-launch times under ncu are cold-cache and serialised: the SHARES count)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_default.csv \
    python bench.py --no-cpu-baseline > gpurun_out/${R}_ncu_default.log 2>&1
# launch list with DRAM bytes at 16 spp (one batch; feeds profiles/traffic.json)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${R}_launches_spp16.csv python bench.py --spp 16 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_spp16.log 2>&1
# --set full of the stage kernels of the adjoint pass (second bounce of the replay + of the adjoint walk)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_ --launch-skip 160 -c 10 -f -o gpurun_out/prof_${R}_wf \
    python bench.py --spp 16 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_full.log 2>&1
ls -la gpurun_out/prof_${R}_wf.ncu-rep gpurun_out/${R}_launches_*.csv
