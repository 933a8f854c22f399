#!/bin/bash
# GPU-box experiment: node budget / refill threshold of the traversal kernel on the C5-shaped scene.
cd "$(dirname "$0")/.."
for cfg in "12 20" "6 20" "24 20" "12 26" "6 26" "12 14"; do
  set -- $cfg
  sed -i "s/constexpr int kWfNodeBudget = [0-9]*;/constexpr int kWfNodeBudget = $1;/" zdr_b200/csrc/render.cu
  make -C zdr_b200/csrc -s -j8 EXTRA="-DZDR_WF_REFILL=$2" > /dev/null 2>&1
  echo "budget=$1 refill=$2 -> $(python tools/run_c5.py --inst 10000 --res 1920 1080 --spp 8 --check-rays 0 2>&1 | tail -1 | cut -c1-110)"
done
sed -i "s/constexpr int kWfNodeBudget = [0-9]*;/constexpr int kWfNodeBudget = 12;/" zdr_b200/csrc/render.cu
