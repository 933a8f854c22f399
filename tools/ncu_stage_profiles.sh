#!/bin/bash
# GPU-box: ncu --set full captures (with source) of single launches of the wavefront stage kernels at bounce 1 of C3 (16 spp, one
# state set so that the launch order is fixed): shade B forward, shade B adjoint, shade A, closest + shadow traversal.
# usage: tools/ncu_stage_profiles.sh <tag>;  read here with tools/ncu_lines.py / tools/ncu_stalls.py / tools/ncu_summary.py
cd "$(dirname "$0")/.."
R=${1:-r02s}
mkdir -p gpurun_out
export ZDR_WF_LANES=1
A="--spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --schedule wavefront"
cap() {   # name kernel-regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 -c $4 -f -o gpurun_out/prof_${R}_$1 \
      python bench.py $A > gpurun_out/${R}_ncu_$1.log 2>&1
  tail -2 gpurun_out/${R}_ncu_$1.log
}
cap shade_b_fwd k_wf_shade_b 25 1
cap shade_b_adj k_wf_shade_b 41 1
cap shade_a k_wf_shade_a 25 1
cap trace k_wf_trace 50 2
ls -la gpurun_out/prof_${R}_*
