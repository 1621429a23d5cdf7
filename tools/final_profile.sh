#!/bin/bash
# Round-end measurement pass on ONE B200 (run through gpurun); everything lands in gpurun_out/r02_* and the
# summaries are copied to profiles/ afterwards (tools/collect_profiles.py).
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
timeout 1200 python bench.py --steps 64 --warmup 8 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; tail -c 2500 $O/r02_bench_n1.json; tail -n 3 $O/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_ref.err; tail -c 900 $O/r02_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:bucket_mul_v|attention_kernel|head_kernel|embed_kernel' -c 500 --csv --log-file $O/r02_launches_decode.csv python bench.py --quick --no-cpu --steps 2 --warmup 1 > $O/r02_ncu_list.log 2>&1
for e in 0.25 1.0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:bucket_mul_v4 -s 20 -c 1 -f -o $O/r02_v4_e$e python tools/prof_one.py --shape 4096x14336 --effort $e --n 24 > $O/r02_ncu_full_$e.log 2>&1
  EFFORT_TRACE=1 timeout 120 python tools/trace_v2.py --shape 4096x14336 --effort $e 2>&1 | tail -40 > $O/r02_timeline_e$e.txt
done
make -s -C tools/ubench all > /dev/null 2>&1
timeout 60 tools/ubench/acc_rate > $O/r02_ubench_acc_rate.txt 2>&1
timeout 60 tools/ubench/stage_cost > $O/r02_ubench_stage_cost.txt 2>&1
timeout 60 tools/ubench/stream_acc > $O/r02_ubench_stream_acc.txt 2>&1
SW="--shapes 4096x14336,4096x4096,4096x1024,14336x4096 --efforts 1.0,0.5,0.25,0.1 --iters 30 --reps 20 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out $O/r02_gemv_sweep.json > $O/r02_gemv_sweep.log 2>&1; tail -n 16 $O/r02_gemv_sweep.log
ls -la $O | grep r02_ | tail -20
