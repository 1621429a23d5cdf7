#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
SW="--shapes 4096x14336,4096x4096,14336x4096 --efforts 1.0,0.5,0.25 --iters 30 --reps 12 --warm_s 0.3"
for st in ldgsts3 bulk ldgsts3 bulk; do
  if [ $st = bulk ]; then export EFFORT_STAGE=bulk; else unset EFFORT_STAGE; fi
  timeout 300 python tools/sweep.py $SW --out $O/r2p_sweep_$st.json > $O/r2p_sweep_$st.log 2>&1
  echo "== $st: $(grep '"us"' $O/r2p_sweep_$st.log | python -c "
import sys,json
print(' '.join(str(json.loads(l)['us']) for l in sys.stdin))")"
  timeout 600 python bench.py --quick --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'roofline us', round(d['roofline']['us_per_launch'],2))"
done
