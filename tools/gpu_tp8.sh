#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 64 --warmup 8 --no-cpu > $O/r02_bench_n8_tp.json 2> $O/r02_bench_n8_tp.err
tail -c 1300 $O/r02_bench_n8_tp.json; grep -v "^W\|^\*\|OMP_NUM" $O/r02_bench_n8_tp.err | tail -n 8
