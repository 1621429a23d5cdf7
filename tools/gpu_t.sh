#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q > $O/r2t_decode.log 2>&1; tail -n 3 $O/r2t_decode.log
for pf in 0 1 0 1; do
  EFFORT_PREFETCH=$pf timeout 300 python bench.py --quick --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prefetch $pf: tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'roofline us', round(d['roofline']['us_per_launch'],2))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:attention_kernel|head_kernel' -c 40 --csv --log-file $O/r2t_launches.csv python bench.py --quick --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
grep -c attention $O/r2t_launches.csv; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2t_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
import collections
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:30]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in agg.items(): print(k, len(v), sum(v)/len(v)/1000, 'us avg', min(v)/1000, 'min')
PY
