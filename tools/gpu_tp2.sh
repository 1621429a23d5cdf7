#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_tp.py -x -q > $O/r02_tp_tests_n2.txt 2>&1; tail -n 6 $O/r02_tp_tests_n2.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 64 --warmup 8 --no-cpu > $O/r02_bench_n2_tp.json 2> $O/r02_bench_n2_tp.err
tail -c 1500 $O/r02_bench_n2_tp.json; grep -v "^W\|^\*\|OMP_NUM" $O/r02_bench_n2_tp.err | tail -n 5
