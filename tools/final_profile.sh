#!/bin/bash
# Round-end measurement pass on ONE B200 (run through gpurun); everything lands in gpurun_out/ and the
# summaries are copied to profiles/ by hand afterwards.
set -u
O=gpurun_out
mkdir -p $O
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 64 --warmup 8 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 1500 $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_ref.err; tail -c 700 $O/bench_reference_arm.json
for g in 1 0; do EFFORT_FUSE_GLUE=$g timeout 100 python bench.py --steps 48 --warmup 6 --quick --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fuse_glue=$g', d['value'], d['gpu_launches'])"; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench_4layers.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu --layers 4 > $O/ncu_list.log 2>&1
for e in 0.25 1.0; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:bucket_mul_fused -s 4 -c 1 -f -o $O/fused_e$e python tools/prof_one.py --effort $e > /dev/null 2>&1
  timeout 60 python tools/trace_one.py --effort $e 2>&1 | tail -16 > $O/timeline_e$e.txt
done
ls -la $O | tail -12
