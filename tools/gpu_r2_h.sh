#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2h_trace_025.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2h_trace_100.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x4096 --effort 0.25 > $O/r2h_trace_wq_025.txt 2>&1
SW="--shapes 4096x14336,4096x4096 --efforts 1.0,0.5,0.25 --iters 30 --reps 20 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out $O/r2h_sweep.json > $O/r2h_sweep.log 2>&1
timeout 300 python -m pytest tests/test_gpu_decode.py -q -k "moe" > $O/r2h_moe.log 2>&1
tail -n 40 $O/r2h_trace_025.txt
tail -n 8 $O/r2h_sweep.log
tail -n 3 $O/r2h_moe.log
