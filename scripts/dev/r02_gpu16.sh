#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02o
timeout 1700 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -4 ${T}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${T}_bench.json 2> ${T}_bench.err; tail -2 ${T}_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r02o_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'frac',round(r['frac'],3),round(r['frac_sec8d_units'],3),'conc',r['pairs_with_concurrent_general_pass'],'launches',d['gpu_launches'])
        print('single',round(r['single_step']['value'],1),r['single_step']['launch_ms'],round(r['single_step']['frac'],4))
        print('dense',round(r['dense_field']['value'],1)); print('e2e',d['e2e']['value'],d['e2e'].get('serial_value'),'plugin',d['e2e_plugin']['value'])
        print('cpu',d['cpu_baseline'])
PY
