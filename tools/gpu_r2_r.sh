#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
SW="--shapes 4096x14336 --efforts 1.0,0.5,0.25 --iters 30 --reps 12 --warm_s 0.3"
for la in 0 1 0 1; do
  export EFFORT_LOOKAHEAD=$la
  timeout 300 python tools/sweep.py $SW --out $O/r2r_sweep_$la.json > $O/r2r_sweep_$la.log 2>&1
  echo "== lookahead $la: $(grep '"us"' $O/r2r_sweep_$la.log | python -c "
import sys,json
print(' '.join(str(json.loads(l)['us']) for l in sys.stdin))")"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "layouts or select" > $O/r2r_parity.log 2>&1; tail -n 2 $O/r2r_parity.log
