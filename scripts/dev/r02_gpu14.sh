#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02m
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 2 > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err; tail -1 ${T}_bench_$tag.err
python - <<PY
import json
for l in open('${T}_bench_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$tag value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'frac',round(r['frac'],3),'launches',d['gpu_launches'],'conc',r['pairs_with_concurrent_general_pass'],'listed',r['listed_range_fraction_last_probe'])
PY
}
run seq KOB_FAST2_CONC=0
run d40s37 KOB_FAST2_CONC=100000 KOB_FAST2_CONC_DIV=40 KOB_FAST2_CONC_SM=37
run d20s37 KOB_FAST2_CONC=100000 KOB_FAST2_CONC_DIV=20 KOB_FAST2_CONC_SM=37
run d20s24 KOB_FAST2_CONC=100000 KOB_FAST2_CONC_DIV=20 KOB_FAST2_CONC_SM=24
run d10s48 KOB_FAST2_CONC=100000 KOB_FAST2_CONC_DIV=10 KOB_FAST2_CONC_SM=48
run d60s16 KOB_FAST2_CONC=100000 KOB_FAST2_CONC_DIV=60 KOB_FAST2_CONC_SM=16
