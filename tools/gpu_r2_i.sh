#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $O/r2i_parity.log 2>&1
tail -n 3 $O/r2i_parity.log
for st in ldgsts3 bulk; do
  if [ $st = bulk ]; then export EFFORT_STAGE=bulk; else unset EFFORT_STAGE; fi
  timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2i_trace_025_$st.txt 2>&1
  timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2i_trace_100_$st.txt 2>&1
  SW="--shapes 4096x14336,4096x4096 --efforts 1.0,0.5,0.25 --iters 30 --reps 20 --warm_s 0.3"
  timeout 300 python tools/sweep.py $SW --out $O/r2i_sweep_$st.json > $O/r2i_sweep_$st.log 2>&1
  echo "== $st"; tail -n 8 $O/r2i_sweep_$st.log
done
unset EFFORT_STAGE
tail -n 45 $O/r2i_trace_025_ldgsts3.txt
timeout 300 python -m pytest tests/test_gpu_decode.py -q -k "moe" > $O/r2i_moe.log 2>&1
tail -n 3 $O/r2i_moe.log
