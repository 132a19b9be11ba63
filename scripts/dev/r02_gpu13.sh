#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02l
timeout 600 python -m pytest tests/test_fast2.py -m gpu -x -q > ${T}_pytest_fast2.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest_fast2.log
tail -15 ${T}_pytest_fast2.log
KOB_TRACE=${T}_trace_pairs.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 --no-invariance > ${T}_trace_bench.json 2> ${T}_trace_bench.err
tail -3 ${T}_trace_bench.err; tail -24 ${T}_trace_pairs.csv
for cm in 0 640; do
KOB_FAST2_CONC=$cm timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance > ${T}_bench_conc$cm.json 2> ${T}_bench_conc$cm.err; tail -2 ${T}_bench_conc$cm.err
python - <<PY
import json
for l in open('${T}_bench_conc$cm.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('conc $cm value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'frac',round(r['frac'],3),'launches',d['gpu_launches'])
PY
done
