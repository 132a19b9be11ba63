#!/bin/bash
# Round-2 GPU session 5: aligned free-mode claims + held-only tier — parity tests, benches, adaptive-vs-single dense, long run.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02e
timeout 1800 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -6 ${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > ${T}_bench.json 2> ${T}_bench.err
python -c "import sys,json; d=json.load(open('${T}_bench.json')); r=d['roofline']; print('seeded', round(d['value'],1), 'dense', r.get('dense_field'))"
KOB_FAST2=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > ${T}_bench_single.json 2>> ${T}_bench.err
python -c "import sys,json; d=json.load(open('${T}_bench_single.json')); r=d['roofline']; print('seeded single', round(d['value'],1), 'dense', r.get('dense_field'))"
one() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['roofline']['launch_ms'],4))"; }
dense() { timeout 200 python bench.py --field dense --steps 5 --warmup 3 --no-cpu --no-e2e --no-dense 2>&1 | one; }
: > ${T}_variants.txt
echo "== dense KOB_FAST2=0" >> ${T}_variants.txt; KOB_FAST2=0 dense >> ${T}_variants.txt 2>&1
echo "== dense adaptive" >> ${T}_variants.txt; dense >> ${T}_variants.txt 2>&1
echo "== dense KOB_FAST_FREE=0" >> ${T}_variants.txt; KOB_FAST2=0 KOB_FAST_FREE=0 dense >> ${T}_variants.txt 2>&1
cat ${T}_variants.txt
timeout 900 python scripts/long_run.py > ${T}_long_run.md 2>&1
tail -8 ${T}_long_run.md
