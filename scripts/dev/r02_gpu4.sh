#!/bin/bash
# Round-2 GPU session 4: TMA input/output rings — parity tests, seeded + dense bench, claim-mode knobs, ncu of dense + seeded launches.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02d
timeout 1500 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -6 ${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > ${T}_bench.json 2> ${T}_bench.err
python -c "import sys,json; d=json.load(open('${T}_bench.json')); r=d['roofline']; print('seeded', round(d['value'],1), 'dense', r.get('dense_field'))"
one() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['roofline']['launch_ms'],4))"; }
dense() { timeout 200 python bench.py --field dense --steps 5 --warmup 3 --no-cpu --no-e2e --no-dense 2>&1 | one; }
: > ${T}_variants.txt
echo "== dense default" >> ${T}_variants.txt; KOB_FAST2=0 dense >> ${T}_variants.txt 2>&1
echo "== dense KOB_FAST_FREE=0" >> ${T}_variants.txt; KOB_FAST2=0 KOB_FAST_FREE=0 dense >> ${T}_variants.txt 2>&1
echo "== dense KOB_FAST_CTA=0" >> ${T}_variants.txt; KOB_FAST2=0 KOB_FAST_CTA=0 dense >> ${T}_variants.txt 2>&1
for e in "KOB_FAST2=0" "KOB_FAST2=0 KOB_FAST_FREE=0" "KOB_FAST2=0 KOB_FAST_CTA=1"; do
  echo "== seeded single-step kernel ($e)" >> ${T}_variants.txt
  env $e timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-dense 2>&1 | one >> ${T}_variants.txt 2>&1
done
cat ${T}_variants.txt
KOB_FAST2=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:kob_step_fast -s 35 -c 1 -o ${T}_dense -f python bench.py --field dense --steps 1 --warmup 3 --no-cpu --no-e2e --no-dense > ${T}_ncu.log 2>&1
KOB_FAST2=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:kob_step_fast -s 35 -c 1 -o ${T}_seeded -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-dense > ${T}_ncu2.log 2>&1
ls -la ${T}_*.ncu-rep
