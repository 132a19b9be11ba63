#!/bin/bash
# Are the GPUs of a box equally fast?  N independent single-GPU benches at once (same thermal / power load as the ring), sequential pairs.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-4}
T=gpurun_out/r02y${N}
nvidia-smi --query-gpu=index,name,clocks.max.sm,clocks.max.mem,power.limit,temperature.gpu,ecc.errors.corrected.volatile.total --format=csv > ${T}_gpus.txt; cat ${T}_gpus.txt
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i KOB_FAST2_CONC=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-invariance --repeats 2 > ${T}_solo_$i.json 2> ${T}_solo_$i.err &
done
wait
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
python - <<PY
import json
for i in range($N):
    for l in open('${T}_solo_%d.json' % i):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print('gpu',i,'value',round(d['value'],1),'launch_ms',round(r['launch_ms'],4),'single',round(r['single_step']['value'],1), d['clocks'])
PY
