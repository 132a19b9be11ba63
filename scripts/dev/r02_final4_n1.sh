#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02zx
timeout 1200 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -3 ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > ${T}_bench.json 2> ${T}_bench.err; tail -2 ${T}_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r02zx_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'frac',round(r['frac'],4),'conc',r['pairs_with_concurrent_general_pass'],'single',round(r['single_step']['value'],1),'dense',round(r['dense_field']['value'],1),'e2e',round(d['e2e']['value'],1),round(d['e2e_plugin']['value'],1),'traffic@',r['traffic_captured_at'])
PY
