#!/bin/bash
# round 2, GPU call D: ncu captures of the fused kernel (source-level stalls) + launch list of the decode chain
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:bucket_mul_v4 -s 20 -c 2 -o $O/r2d_v4_e025 python tools/prof_one.py --shape 4096x14336 --effort 0.25 --n 24 > $O/r2d_ncu_a.log 2>&1
timeout 600 $NCU -k regex:bucket_mul_v4 -s 20 -c 1 -o $O/r2d_v4_e100 python tools/prof_one.py --shape 4096x14336 --effort 1.0 --n 24 > $O/r2d_ncu_b.log 2>&1
EFFORT_STAGE=ldgsts timeout 600 $NCU -k regex:bucket_mul_v2 -s 20 -c 1 -o $O/r2d_v2_e025 python tools/prof_one.py --shape 4096x14336 --effort 0.25 --n 24 > $O/r2d_ncu_c.log 2>&1
timeout 600 $NCU -k regex:bucket_mul_v4 -s 20 -c 1 -o $O/r2d_v4_wk_e025 python tools/prof_one.py --shape 4096x1024 --effort 0.25 --n 24 > $O/r2d_ncu_d.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $O/r2d_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu --layers 4 > $O/r2d_launch_bench.log 2>&1
ls -la $O/*.ncu-rep
