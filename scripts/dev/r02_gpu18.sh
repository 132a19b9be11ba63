#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02v
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 2 > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err; tail -1 ${T}_bench_$tag.err
python - <<PY
import json
for l in open('${T}_bench_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$tag value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'frac',round(r['frac'],3),'launches',d['gpu_launches'],'conc',r['pairs_with_concurrent_general_pass'])
PY
}
run m100s24 KOB_FAST2_CONC_SM=24
run m100s28 KOB_FAST2_CONC_SM=28
run m110s24 KOB_FAST2_CONC_SM=24 KOB_FAST2_CONC_MARGIN=110
run m100s24t75 KOB_FAST2_CONC_SM=24 KOB_FAST2_TICKET_US=75
run m100s20 KOB_FAST2_CONC_SM=20 KOB_FAST2_CONC_THR=200
run m90s24 KOB_FAST2_CONC_SM=24 KOB_FAST2_CONC_MARGIN=90
