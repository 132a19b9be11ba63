#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02s
run() {
  tag=$1; np=$2; shift; shift
  if [ $np = 1 ]; then
  env "$@" timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 3 > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err
  else
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 3 > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err
  fi
python - <<PY
import json
for l in open('${T}_bench_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$tag value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'conc',r['pairs_with_concurrent_general_pass'],'launches',d['gpu_launches'], 'strong', (d.get('strong_65536') or {}).get('value'))
PY
}
timeout 600 python -m pytest tests/test_fast2.py -m gpu -x -q 2>&1 | tail -2

run n1conc 1 KOB_FAST2_CONC=1000000

run n2conc 2 KOB_FAST2_CONC=1000000
