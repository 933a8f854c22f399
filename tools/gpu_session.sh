#!/bin/bash
# One GPU-box session: tests, smoke, bench lines of every config (both arms for c3), the ncu launch list with
# instruction-issue + DRAM metrics that feeds bench.py's roofline.  usage: tools/gpu_session.sh <tag> [steps...]
# steps: tests smoke bench ref configs issue full (default: all)
cd "$(dirname "$0")/.."
R=${1:-r02a}; shift
STEPS=${*:-tests smoke bench ref configs issue}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log; fi
if has smoke; then python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log; fi
if has bench; then python bench.py > gpurun_out/${R}_bench_c3.json 2> gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench_c3.json; fi
if has ref; then python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err; cut -c1-200 gpurun_out/${R}_bench_reference.json; fi
if has configs; then
  for c in c1 c2 c4 c4d; do python bench.py --config $c > gpurun_out/${R}_bench_$c.json 2>> gpurun_out/${R}_bench.err; cut -c1-200 gpurun_out/${R}_bench_$c.json; done
  timeout 1500 python bench.py --config c5 --steps 2 --warmup 2 > gpurun_out/${R}_bench_c5.json 2>> gpurun_out/${R}_bench.err; cut -c1-200 gpurun_out/${R}_bench_c5.json
fi
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
if has issue; then
  for c in ${ISSUE_CONFIGS:-c3}; do
    A="--config $c --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"; [[ $c == c3 ]] && A="$A --spp 16"; [[ $c == c5 ]] && A="$A --res 960 540 --spp 8"
    python bench.py $A > gpurun_out/${R}_issue_$c.bench.json 2>> gpurun_out/${R}_bench.err
    timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${R}_issue_$c.csv python bench.py $A > gpurun_out/${R}_issue_$c.ncu.log 2>&1
  done
fi
if has full; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${FULL_KERNELS:-k_} --launch-skip ${FULL_SKIP:-40} -c ${FULL_COUNT:-8} -f -o gpurun_out/prof_${R} \
      python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_full.log 2>&1
fi
ls -la gpurun_out | tail -20
