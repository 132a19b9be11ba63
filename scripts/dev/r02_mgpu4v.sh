#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=4
T=gpurun_out/r02za
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {
  tag=$1; shift
  env KOB_FAST2_CONC=0 "$@" > ${T}_bench_$tag.json 2> ${T}_bench_$tag.err
  python - <<PY
import json
for l in open('${T}_bench_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$tag n',d['n_gpus'],'value',round(d['value'],1),'launch_ms',round(r['launch_ms'],4), d.get('seam_waits',{}).get('per_rank_waits'))
PY
}
B="bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --strong-secondary 0 --repeats 1"
run base timeout 600 $TR --master-port 29551 $B
run nonuclei timeout 600 $TR --master-port 29552 $B --nuclei 0
run onenuc timeout 600 $TR --master-port 29553 $B --nuclei 1
run reversed env CUDA_VISIBLE_DEVICES=3,2,1,0 timeout 600 $TR --master-port 29554 $B
run nonoise timeout 600 $TR --master-port 29555 $B --noise 0
