#!/bin/bash
mkdir -p gpurun_out
for e in 0.25 1.0; do EFFORT_TRACE=2 timeout 120 python tools/trace_v2.py --shape 4096x14336 --effort $e 2>&1 | grep "CTA 0 thread 0\|listed\|cutoff\|scored\|streamed\|reduced" | tail -6; done
EFFORT_TRACE=2 timeout 120 python tools/trace_v2.py --shape 4096x4096 --effort 0.25 2>&1 | grep "CTA 0 thread 0\|listed\|cutoff\|scored\|streamed\|reduced" | tail -6
SW="--shapes 4096x14336,4096x4096 --efforts 1.0,0.25 --iters 30 --reps 12 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out gpurun_out/r2u_sweep.json 2>&1 | grep '"us"' | python -c "
import sys,json
print(' '.join(str(json.loads(l)['us']) for l in sys.stdin))"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "layouts or select" 2>&1 | tail -n 2
