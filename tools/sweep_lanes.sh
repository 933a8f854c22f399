cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
b() { python bench.py "$@" --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["config"]["ms_forward"], d["config"]["ms_backward"])'; }
for l in 1 2 1 2; do
  export ZDR_WF_LANES=$l
  echo "lanes=$l C3@64spp -> $(b --spp 64) | 362^2@256spp -> $(b --res 362 --spp 256) | C3 full -> $(b)"
done
