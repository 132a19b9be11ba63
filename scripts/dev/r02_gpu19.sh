#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02zd
run() {
  tag=$1; st=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps $st --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 2 > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err; tail -1 ${T}_bench_$tag.err
python - <<PY
import json
for l in open('${T}_bench_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$tag value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'frac',round(r['frac'],3),'launches',d['gpu_launches'],'conc',r['pairs_with_concurrent_general_pass'],'listed',round(r['listed_range_fraction_last_probe'],4))
PY
}
timeout 300 python -m pytest tests/test_fast2.py -m gpu -x -q 2>&1 | tail -1
run s20 20 A=1
run s20seq 20 KOB_FAST2_CONC=0
run s40 40 A=1
run s40seq 40 KOB_FAST2_CONC=0
run s80 80 A=1
run s80seq 80 KOB_FAST2_CONC=0
run s80cap24 80 KOB_FAST2_CONC_SM=24
