#!/bin/bash
# round 2, GPU call A: correctness of the new engine + first A/B numbers
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q --maxfail=10 -k "not bulk_stage" > $O/r2a_parity.log 2>&1; echo "parity rc=$?" >> $O/r2a_rc.txt
timeout 300 python -m pytest tests/test_gpu_decode.py -q --maxfail=10 -k "not 32_layers" > $O/r2a_decode.log 2>&1; echo "decode rc=$?" >> $O/r2a_rc.txt
SW="--shapes 4096x14336,4096x4096,14336x4096 --efforts 1.0,0.5,0.25 --iters 30 --reps 20 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out $O/r2a_sweep_v2.json > $O/r2a_sweep_v2.log 2>&1
EFFORT_ENGINE=1 timeout 300 python tools/sweep.py $SW --out $O/r2a_sweep_v1.json > $O/r2a_sweep_v1.log 2>&1
EFFORT_DYN=0 timeout 300 python tools/sweep.py $SW --out $O/r2a_sweep_static.json > $O/r2a_sweep_static.log 2>&1
EFFORT_CUTOFF=bisect timeout 300 python tools/sweep.py $SW --out $O/r2a_sweep_bisect.json > $O/r2a_sweep_bisect.log 2>&1
EFFORT_LAYOUT=slice timeout 300 python tools/sweep.py $SW --out $O/r2a_sweep_slice.json > $O/r2a_sweep_slice.log 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2a_trace_025.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2a_trace_100.txt 2>&1
timeout 400 python bench.py --steps 32 --warmup 8 --no-cpu > $O/r2a_bench.json 2> $O/r2a_bench.err
EFFORT_CHAIN=1 timeout 400 python bench.py --steps 32 --warmup 8 --no-cpu --quick > $O/r2a_bench_chain1.json 2> $O/r2a_bench_chain1.err
# the TMA (cp.async.bulk) ring last: a hang here must not cost the numbers above
timeout 300 python -m pytest tests/test_gpu_parity.py -q --maxfail=10 -k "bulk_stage" > $O/r2a_bulk.log 2>&1; echo "bulk rc=$?" >> $O/r2a_rc.txt
EFFORT_LAYOUT=slice EFFORT_STAGE=bulk timeout 300 python tools/sweep.py $SW --out $O/r2a_sweep_bulk.json > $O/r2a_sweep_bulk.log 2>&1
timeout 600 python -m pytest tests/test_gpu_decode.py -q -s -k "32_layers" > $O/r2a_decode32.log 2>&1; echo "decode32 rc=$?" >> $O/r2a_rc.txt
cat $O/r2a_rc.txt
tail -5 $O/r2a_parity.log $O/r2a_decode.log
cat $O/r2a_sweep_v2.log | tail -12
