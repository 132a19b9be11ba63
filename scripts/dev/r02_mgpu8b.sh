#!/bin/bash
# N-GPU session (round 2, concurrent general pass): bitwise shard invariance on a ragged torus and on a dense field (in-library ring
# policy), N=1 bench on the same box (efficiency denominator), weak-scaling bench (default protocol: fresh field per repeat,
# shard invariance pre-check, 65536^2 strong secondary), and the same with the plain far -> general order.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
T=gpurun_out/r02ze${N}
nvidia-smi -L > ${T}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 scripts/mgpu_check.py --nx 3000 --ny 2063 --steps 120 > ${T}_check.log 2>&1
grep "mgpu_check" ${T}_check.log
timeout 600 $TR --master-port 29542 scripts/mgpu_check.py --dense --nx 3000 --ny 4000 --steps 200 > ${T}_check_dense.log 2>&1
grep "mgpu_check" ${T}_check_dense.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance > ${T}_bench_n1.json 2> ${T}_bench_n1.err
timeout 900 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > ${T}_bench.json 2> ${T}_bench.err
tail -3 ${T}_bench.err
KOB_FAST2_CONC=0 timeout 900 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --no-invariance --no-dense --no-single > ${T}_bench_seq.json 2> ${T}_bench_seq.err
python - <<PY
import json
for f in ('${T}_bench_n1.json','${T}_bench.json','${T}_bench_seq.json'):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'n',d['n_gpus'],'value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'conc',r['pairs_with_concurrent_general_pass'],'launches',d['gpu_launches'], 'inv', d.get('shard_invariance'), 'strong', (d.get('strong_65536') or {}).get('value'), (d.get('seam_waits') or {}).get('per_rank_waits'))
PY
