#!/bin/bash
# round 2, GPU call B: the TMA pipeline kernel (bucket_mul_v3) -- correctness first, then numbers
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2b_rc.txt
# 1. one operator through v3 with a hard timeout (a hang must not eat the call)
timeout 120 python tools/prof_one.py --shape 4096x14336 --effort 0.25 --n 4 > $O/r2b_first.log 2>&1; echo "first rc=$?" >> $O/r2b_rc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q --maxfail=10 -x > $O/r2b_parity.log 2>&1; echo "parity rc=$?" >> $O/r2b_rc.txt
timeout 600 python -m pytest tests/test_gpu_decode.py -q --maxfail=10 -k "not 32_layers" > $O/r2b_decode.log 2>&1; echo "decode rc=$?" >> $O/r2b_rc.txt
SW="--shapes 4096x14336,4096x4096,14336x4096 --efforts 1.0,0.5,0.25 --iters 30 --reps 20 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out $O/r2b_sweep_v3.json > $O/r2b_sweep_v3.log 2>&1
EFFORT_STAGE=ldgsts timeout 300 python tools/sweep.py $SW --out $O/r2b_sweep_ldgsts.json > $O/r2b_sweep_ldgsts.log 2>&1
EFFORT_STAGE=bulk timeout 300 python tools/sweep.py $SW --out $O/r2b_sweep_bulk.json > $O/r2b_sweep_bulk.log 2>&1
EFFORT_CUTOFF=bisect timeout 300 python tools/sweep.py $SW --out $O/r2b_sweep_bisect.json > $O/r2b_sweep_bisect.log 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2b_trace_025.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2b_trace_100.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x4096 --effort 0.25 > $O/r2b_trace_wq_025.txt 2>&1
timeout 400 python bench.py --steps 32 --warmup 8 --no-cpu > $O/r2b_bench.json 2> $O/r2b_bench.err
timeout 600 python tools/depth_scan.py 0.25 > $O/r2b_depth_025.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_decode.py -q -s -k "32_layers" > $O/r2b_decode32.log 2>&1; echo "decode32 rc=$?" >> $O/r2b_rc.txt
cat $O/r2b_rc.txt
tail -3 $O/r2b_parity.log $O/r2b_decode.log
tail -9 $O/r2b_sweep_v3.log
