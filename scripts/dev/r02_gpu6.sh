#!/bin/bash
# Round-2 GPU session 6: 3-stage live ring — parity tests, dense adaptive vs single (per-launch ncu durations), long run.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02f
timeout 1800 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -6 ${T}_pytest.log
one() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['roofline']['launch_ms'],4), d['roofline'].get('paired_substeps'), d['roofline'].get('single_substeps'))"; }
dense() { timeout 200 python bench.py --field dense --steps 5 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 2>&1 | one; }
: > ${T}_variants.txt
echo "== dense KOB_FAST2=0" >> ${T}_variants.txt; KOB_FAST2=0 dense >> ${T}_variants.txt 2>&1
echo "== dense adaptive" >> ${T}_variants.txt; dense >> ${T}_variants.txt 2>&1
cat ${T}_variants.txt
KOB_FAST2=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file ${T}_launches_single.csv python bench.py --field dense --steps 5 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file ${T}_launches_adaptive.csv python bench.py --field dense --steps 5 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 > /dev/null 2>&1
timeout 900 python scripts/long_run.py > ${T}_long_run.md 2>&1
tail -6 ${T}_long_run.md
timeout 600 python bench.py --steps 10 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err
python -c "import sys,json; d=json.load(open('${T}_bench.json')); r=d['roofline']; print('seeded', round(d['value'],1), d['repeats'], 'single', r.get('single_step'), 'dense', r.get('dense_field'), d['cpu_baseline'], d['e2e'], d['e2e_plugin'])"
tail -3 ${T}_bench.err
