#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2g_rc.txt
timeout 120 python tools/prof_one.py --shape 4096x14336 --effort 0.25 --n 4 > $O/r2g_first.log 2>&1; echo "first rc=$?" >> $O/r2g_rc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q --maxfail=10 > $O/r2g_parity.log 2>&1; echo "parity rc=$?" >> $O/r2g_rc.txt
timeout 900 python -m pytest tests/test_gpu_decode.py -q --maxfail=10 -k "not 32_layers" > $O/r2g_decode.log 2>&1; echo "decode rc=$?" >> $O/r2g_rc.txt
SW="--shapes 4096x14336,4096x4096,14336x4096,4096x1024 --efforts 1.0,0.5,0.25 --iters 30 --reps 20 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out $O/r2g_sweep.json > $O/r2g_sweep.log 2>&1
EFFORT_HINT=0 timeout 300 python tools/sweep.py $SW --out $O/r2g_sweep_nohint.json > $O/r2g_sweep_nohint.log 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2g_trace_025.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2g_trace_100.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x4096 --effort 0.25 > $O/r2g_trace_wq_025.txt 2>&1
timeout 200 python tools/rounds.py > $O/r2g_rounds.txt 2>&1
timeout 600 python bench.py --steps 32 --warmup 8 --no-extras --no-cpu --no-quality > $O/r2g_bench.json 2> $O/r2g_bench.err; echo "bench rc=$?" >> $O/r2g_rc.txt
EFFORT_HINT=0 timeout 600 python bench.py --steps 32 --warmup 8 --quick --no-cpu > $O/r2g_bench_nohint.json 2> $O/r2g_bench_nohint.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:bucket_mul_v4 -s 20 -c 1 -o $O/r2g_v4_e025 python tools/prof_one.py --shape 4096x14336 --effort 0.25 --n 24 > $O/r2g_ncu_a.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:effort -s 300 -c 200 --csv --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu --layers 4 > $O/r2g_launch_bench.log 2>&1
cat $O/r2g_rc.txt
tail -n 3 $O/r2g_parity.log $O/r2g_decode.log
tail -n 12 $O/r2g_sweep.log
cat $O/r2g_rounds.txt
