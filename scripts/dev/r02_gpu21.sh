#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02zh
V=crystalgrowth_b200/variants/libkobayashi_cuda_r8.so
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 1 > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err; tail -1 ${T}_bench_$tag.err
python - <<PY
import json
for l in open('${T}_bench_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$tag value',round(d['value'],1),'launch_ms',round(r['launch_ms'],4),'conc',r['pairs_with_concurrent_general_pass'],'listed',round(r['listed_range_fraction_last_probe'],4))
PY
}
run base A=1
run yj128 KOB_FAST2_YJ=128 KOB_FAST2_YJB=32
run yj192 KOB_FAST2_YJ=192 KOB_FAST2_YJB=32
run r8yj192 KOB_LIB_PATH=$V KOB_FAST2_YJ=192 KOB_FAST2_YJB=32
run r8yj128 KOB_LIB_PATH=$V KOB_FAST2_YJ=128 KOB_FAST2_YJB=32
run r8yj96 KOB_LIB_PATH=$V
run yj192seq KOB_FAST2_YJ=192 KOB_FAST2_YJB=32 KOB_FAST2_CONC=0
