#!/bin/bash
# GPU-box session: parity of the new pieces (gradient cache, device alias tables), then the global-state variants of the
# persistent kernel on C3 at 64 spp.
cd "$(dirname "$0")/.."
R=${1:-r02e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -k "envmap or contention or scatter or light_stage or c2_full or uniform or persistent" > gpurun_out/${R}_pytest.log 2>&1; tail -4 gpurun_out/${R}_pytest.log
b() { timeout 300 python bench.py "$@" --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --schedule persistent 2>/dev/null | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d["value"], "Mrays/s", d["ms_per_step"], "ms; bwd", d["config"]["ms_backward_last_render"], "ms")
except Exception as e: print("FAILED", e)'; }
echo "default lib: $(b)" | tee gpurun_out/${R}_variants.txt
for lib in zdr_b200/csrc/variants/lib_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#lib_}
  # parity of the variant on one case
  echo "variant $n parity: $(ZDR_B200_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k 'c3_path_small_persistent or path_persistent_many' 2>&1 | tail -1)" | tee -a gpurun_out/${R}_variants.txt
  for pool in ${POOLS:-1024 2048 4096 8192}; do echo "variant $n pool=$pool: $(ZDR_PW_POOL=$pool ZDR_B200_LIB=$PWD/$lib b)" | tee -a gpurun_out/${R}_variants.txt; done
done
for c in c2 c4 c4d; do python bench.py --config $c --no-cpu-baseline > gpurun_out/${R}_bench_$c.json 2>> gpurun_out/${R}_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/${R}_bench_$c.json').read().strip().splitlines()[-1]); print('$c', d['value'], d['ms_per_step'], 'red/scatter', d['config']['red_v4_per_scatter'], 'e2e', d['e2e']['value'])" | tee -a gpurun_out/${R}_variants.txt; done
