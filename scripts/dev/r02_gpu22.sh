#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02zv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file ${T}_launches_default.csv python -c "
import os
print('INJ', {k:v for k,v in os.environ.items() if 'INJECT' in k or 'NSIGHT' in k or 'NV_COMPUTE' in k or 'NSYS' in k})
import crystalgrowth_b200 as cg
g = cg.Kobayashi(4096, 4096, 1e-4, kernel='fast', seed=1, noise_a=0.01)
for k in range(5):
    g.step(10); g.sync()
print('stats', g.path_stats())
g.close()
" 2>&1 | grep "INJ\|stats"
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/r02zv_launches_default.csv') if not l.startswith('==')))
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except Exception: continue
    ms=v/1e6 if r[iu].startswith('ns') else (v/1e3 if r[iu].startswith('us') else v)
    a=agg.setdefault(r[ik][:40],[0,0.0]); a[0]+=1; a[1]+=ms
for k,(n,ms) in agg.items(): print(k,n,round(ms,3),round(ms/n,4))
PY
timeout 300 python -m pytest tests/test_fast2.py -m gpu -x -q 2>&1 | tail -1
