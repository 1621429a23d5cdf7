#!/bin/bash
mkdir -p gpurun_out
for e in 0.25 1.0; do EFFORT_TRACE=2 timeout 120 python tools/trace_v2.py --shape 4096x14336 --effort $e 2>&1 | grep "CTA 0 thread 0"; done
SW="--shapes 4096x14336,4096x4096 --efforts 1.0,0.25 --iters 30 --reps 12 --warm_s 0.3"
timeout 300 python tools/sweep.py $SW --out gpurun_out/r2u_sweep.json 2>&1 | grep '"us"' | python -c "
import sys,json
print(' '.join(str(json.loads(l)['us']) for l in sys.stdin))"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decode.py -x -q 2>&1 | tail -n 2
timeout 300 python bench.py --quick --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'roofline us', round(d['roofline']['us_per_launch'],2))"
