#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
export EFFORT_STAGE=bulk
SW="--shapes 4096x14336 --efforts 1.0,0.25 --iters 30 --reps 12 --warm_s 0.3"
for la in 0 1; do for win in 1 2 4 8; do
  export EFFORT_LOOKAHEAD=$la EFFORT_WINDOW=$win
  timeout 300 python tools/sweep.py $SW --out $O/r2k_sweep_${la}_${win}.json > $O/r2k_sweep_${la}_${win}.log 2>&1
  echo "== lookahead $la window $win"; grep '"us"' $O/r2k_sweep_${la}_${win}.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ', d['shape'], d['effort'], d['us'])"
done; done
export EFFORT_LOOKAHEAD=0 EFFORT_WINDOW=2
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2k_trace_025.txt 2>&1
tail -n 30 $O/r2k_trace_025.txt
