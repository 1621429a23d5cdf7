#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
export EFFORT_STAGE=bulk EFFORT_LOOKAHEAD=0 EFFORT_WINDOW=2
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2l_trace_025.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2l_trace_100.txt 2>&1
export EFFORT_LOOKAHEAD=1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2l_trace_100_la.txt 2>&1
tail -n 28 $O/r2l_trace_025.txt
tail -n 28 $O/r2l_trace_100.txt
tail -n 28 $O/r2l_trace_100_la.txt
