#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02q
run() {
  tag=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 3 > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err
python - <<PY
import json
for l in open('${T}_bench_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$tag value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'conc',r['pairs_with_concurrent_general_pass'],'launches',d['gpu_launches'])
PY
}
run seq KOB_FAST2_CONC=0
run conc KOB_FAST2_CONC=1000000
KOB_TRACE=${T}_trace_%p.csv timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 1 > ${T}_trace_bench.json 2> ${T}_trace_bench.err
for f in ${T}_trace_*.csv; do echo $f; tail -30 $f | awk -F, '{printf "%s %s %s | ", substr($1,1,12),$3,$4; if (NR%4==0) print ""}'; echo; done
