#!/bin/bash
# round 2, GPU call E: v4 + speculative L2 prefetch; MoE / Q4 decode tests; bench
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2e_rc.txt
timeout 120 python tools/prof_one.py --shape 4096x14336 --effort 0.25 --n 4 > $O/r2e_first.log 2>&1; echo "first rc=$?" >> $O/r2e_rc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q --maxfail=10 > $O/r2e_parity.log 2>&1; echo "parity rc=$?" >> $O/r2e_rc.txt
timeout 900 python -m pytest tests/test_gpu_decode.py -q --maxfail=10 -k "not 32_layers" > $O/r2e_decode.log 2>&1; echo "decode rc=$?" >> $O/r2e_rc.txt
SW="--shapes 4096x14336,4096x4096,14336x4096,4096x1024 --efforts 1.0,0.5,0.25 --iters 30 --reps 20 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out $O/r2e_sweep_pf.json > $O/r2e_sweep_pf.log 2>&1
EFFORT_PREFETCH=0 timeout 300 python tools/sweep.py $SW --out $O/r2e_sweep_nopf.json > $O/r2e_sweep_nopf.log 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2e_trace_025.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x4096 --effort 0.25 > $O/r2e_trace_wq_025.txt 2>&1
timeout 600 python bench.py --steps 32 --warmup 8 --no-extras --no-cpu > $O/r2e_bench.json 2> $O/r2e_bench.err; echo "bench rc=$?" >> $O/r2e_rc.txt
EFFORT_PREFETCH=0 timeout 600 python bench.py --steps 32 --warmup 8 --quick --no-cpu > $O/r2e_bench_nopf.json 2> $O/r2e_bench_nopf.err
cat $O/r2e_rc.txt
tail -n 3 $O/r2e_parity.log $O/r2e_decode.log
tail -n 12 $O/r2e_sweep_pf.log
