#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-4}
T=gpurun_out/r02zb4_${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in 0 1000000; do
KOB_FAST2_CONC=$mode KOB_TRACE=${T}_trace_c${mode}_%p.csv timeout 600 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --strong-secondary 0 --repeats 1 > ${T}_bench_c$mode.json 2> ${T}_bench_c$mode.err
python - <<PY
import json,glob,csv
for l in open('${T}_bench_c$mode.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('mode $mode n',d['n_gpus'],'value',round(d['value'],1),'launch_ms',round(r['launch_ms'],4),'conc',r['pairs_with_concurrent_general_pass'], d.get('seam_waits'))
for f in sorted(glob.glob('${T}_trace_c${mode}_*.csv')):
    rows=list(csv.DictReader(open(f)))
    # last 100 pair records = timed region
    if '$mode'=='0':
        far=[float(r['duration_us']) for r in rows if r['kernel']=='kob_far2'][-100:]
        gen=[float(r['duration_us']) for r in rows if r['kernel'].startswith('kob_step_fast2')][-100:]
        print(f.split('_')[-1], 'far2 mean',round(sum(far)/len(far),1),'max',max(far),'general mean',round(sum(gen)/len(gen),1),'max',max(gen), 'first10 far', [int(x) for x in far[:10]])
    else:
        pr=[float(r['duration_us']) for r in rows if r['kernel'].startswith('pair')][-100:]
        est=[r['info'] for r in rows if r['kernel'].startswith('pair')][-100:]
        print(f.split('_')[-1], 'pair mean',round(sum(pr)/len(pr),1),'min',min(pr),'max',max(pr), 'first', [int(x) for x in pr[:8]], 'last', [int(x) for x in pr[-8:]], est[0], est[-1])
PY
done
