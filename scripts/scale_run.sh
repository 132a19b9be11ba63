#!/bin/bash
# dev: N-GPU correctness + weak + strong scaling on the box this runs on.  usage: scale_run.sh N
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" -gt 1 ]; then
  timeout 300 $TR --master-port 29601 scripts/mgpu_check.py --kernel fast --ny 1031 --nx 500 --steps 80 2>&1 | grep mgpu_check
  timeout 300 $TR --master-port 29602 scripts/mgpu_check.py --kernel strict --precision f64 --ny 300 --nx 200 --steps 40 2>&1 | grep mgpu_check
  timeout 600 $TR --master-port 29603 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>&1 | grep '^{' > gpurun_out/scale_weak_n$N.json
  timeout 900 $TR --master-port 29604 bench.py --gpus $N --steps 5 --warmup 3 --strong 65536 2>&1 | grep '^{' > gpurun_out/scale_strong_n$N.json
else
  timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu | grep '^{' > gpurun_out/scale_weak_n1.json
  timeout 900 python bench.py --steps 5 --warmup 3 --strong 65536 --no-cpu | grep '^{' > gpurun_out/scale_strong_n1.json
fi
for f in gpurun_out/scale_weak_n$N.json gpurun_out/scale_strong_n$N.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', round(d['value'],1), 'Gcell/s', d['scaling'], 'ms/launch', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],3))"; done
