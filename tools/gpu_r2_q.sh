#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for win in 8 2; do
export EFFORT_WINDOW=$win
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2q_trace_025_w$win.txt 2>&1
timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 1.0 > $O/r2q_trace_100_w$win.txt 2>&1
echo "== window $win"
grep -A1 "rep 2" $O/r2q_trace_025_w$win.txt | head -2; grep "streamed\|reduced\|listed" $O/r2q_trace_025_w$win.txt | tail -3; grep "consumers of CTA" $O/r2q_trace_025_w$win.txt
grep "streamed\|reduced\|listed" $O/r2q_trace_100_w$win.txt | tail -3; grep "consumers of CTA" $O/r2q_trace_100_w$win.txt
done
EFFORT_STAGE=pairs-ldgsts timeout 200 python tools/trace_v2.py --shape 4096x14336 --effort 0.25 > $O/r2q_trace_025_ldgsts.txt 2>&1
echo "== ldgsts pairs"; grep "streamed\|reduced\|listed" $O/r2q_trace_025_ldgsts.txt | tail -3; grep "consumers of CTA" $O/r2q_trace_025_ldgsts.txt
