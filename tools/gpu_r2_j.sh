#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
export EFFORT_STAGE=bulk
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "layouts or select or engine" > $O/r2j_parity.log 2>&1
tail -n 3 $O/r2j_parity.log
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2j_trace_025.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2j_trace_100.txt 2>&1
SW="--shapes 4096x14336,4096x4096,14336x4096 --efforts 1.0,0.5,0.25 --iters 30 --reps 20 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out $O/r2j_sweep.json > $O/r2j_sweep.log 2>&1
tail -n 10 $O/r2j_sweep.log
tail -n 45 $O/r2j_trace_025.txt
tail -n 30 $O/r2j_trace_100.txt
