#!/bin/bash
# GPU box: every script under examples/ at small sizes.  usage: tools/examples_session.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-r05b}; O=gpurun_out/${R}_examples; mkdir -p $O
{
echo "== benchmark.py"; timeout 90 python examples/benchmark.py --iters 10 --res 256 --spp 32 2>&1 | grep -v Warning | tail -6
echo "== benchmark.py (reference sizes, 20 iterations)"; timeout 90 python examples/benchmark.py --iters 20 2>&1 | grep -v Warning | tail -2
echo "== lightstage.py"; timeout 90 python examples/lightstage.py --nlight 8 --res 256 --spp 64 --out $O 2>&1 | grep -v Warning | tail -3
echo "== fd_validate.py cbox path"; timeout 120 python examples/fd_validate.py --res 64 --max-exp 10 --seed 3 2>&1 | grep -v Warning | tail -34
echo "== fd_validate.py sphere + envmap direct"; timeout 120 python examples/fd_validate.py --scene sphere --integrator direct --res 64 --max-exp 8 --seed 5 2>&1 | grep -v Warning | tail -8
echo "== inverse_cornell.py"; timeout 120 python examples/inverse_cornell.py --res 128 --tex 64 --spp 8 --gt-spp 64 --iters 40 --out $O 2>&1 | grep -v Warning | tail -3
} > gpurun_out/${R}_examples.log 2>&1
cat gpurun_out/${R}_examples.log
