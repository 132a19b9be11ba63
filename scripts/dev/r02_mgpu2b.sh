#!/bin/bash
# 2-GPU session: IPC ring parity tests, in-library ring policy on a dense field, weak-scaling bench with shard invariance;
# N=1 bench on the same box for the efficiency denominator.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02p
nvidia-smi -L > ${T}_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -5 ${T}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/mgpu_check.py --dense --nx 2000 --ny 2000 --steps 300 > ${T}_dense_policy.log 2>&1
grep "mgpu_check" ${T}_dense_policy.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance > ${T}_bench_n1.json 2> ${T}_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > ${T}_bench_n2.json 2> ${T}_bench_n2.err
tail -5 ${T}_bench_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r02p_bench_n1.json','gpurun_out/r02p_bench_n2.json'):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'conc',r['pairs_with_concurrent_general_pass'],'launches',d['gpu_launches'], 'inv', d.get('shard_invariance'), 'strong', d.get('strong_65536'))
PY
