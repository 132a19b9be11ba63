#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02k
timeout 900 python -m pytest tests -m gpu -x -q -k "not 16384 and not 2_pow_31 and not 4096" > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -3 ${T}_pytest.log
KOB_TRACE=${T}_trace_pairs.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 --no-invariance > ${T}_trace_bench.json 2> ${T}_trace_bench.err
tail -40 ${T}_trace_pairs.csv
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > ${T}_bench.json 2> ${T}_bench.err; tail -2 ${T}_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r02k_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'frac',round(r['frac'],3),round(r['frac_sec8d_units'],3))
        print('single',round(r['single_step']['value'],1),r['single_step']['launch_ms'],round(r['single_step']['frac'],4))
        print('dense',round(r['dense_field']['value'],1)); print('e2e',d['e2e'],'plugin',d['e2e_plugin']['value'])
PY
