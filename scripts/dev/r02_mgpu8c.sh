#!/bin/bash
# N-GPU long dendrite-growth run (configs[4] in small) + the N>1 bench line with the developed-field leg
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
T=gpurun_out/r02zg${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29571 scripts/long_run_ring.py --edge 8192 --nuclei 64 --steps 40000 --chunk 4000 > ${T}_long_run.md 2> ${T}_long_run.err
grep -v "^\*\|OMP_NUM" ${T}_long_run.md | tail -16
timeout 600 $TR --master-port 29572 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --no-invariance --strong-secondary 0 --repeats 2 > ${T}_bench.json 2> ${T}_bench.err
python - <<PY
import json
for l in open('${T}_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('n',d['n_gpus'],'value',round(d['value'],1),'launch_ms',round(r['launch_ms'],4),'dense',r.get('dense_field'))
PY
