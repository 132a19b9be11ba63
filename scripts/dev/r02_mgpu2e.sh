#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02zf
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or trace or ring" 2>&1 | tail -2
timeout 900 $TR --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > ${T}_bench_n2.json 2> ${T}_bench_n2.err; tail -3 ${T}_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02zf_bench_n2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('n2 value',round(d['value'],1),[round(v,1) for v in d['repeats']['values']],'launch_ms',round(r['launch_ms'],4),'inv',d.get('shard_invariance'),'strong',d['strong_65536']['value'],'dense',r.get('dense_field',{}).get('value'), r.get('dense_field',{}).get('launch_ms'),'e2e',d['e2e']['value'],d['e2e_plugin']['value'], d['seam_waits']['per_rank_waits'])
PY
timeout 600 $TR --master-port 29562 scripts/long_run_ring.py --n 4096 --nuclei 16 --steps 12000 --chunk 3000 2>&1 | grep -v "^\*\|OMP_NUM" | tail -9
