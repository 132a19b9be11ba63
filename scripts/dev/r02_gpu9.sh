#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02h
timeout 900 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -4 ${T}_pytest.log
one() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['roofline']['launch_ms'],4), d['roofline'].get('paired_substeps'), d['roofline'].get('single_substeps'))"; }
: > ${T}_variants.txt
echo "== dense KOB_FAST2=0" >> ${T}_variants.txt; KOB_FAST2=0 timeout 200 python bench.py --field dense --steps 5 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 2>&1 | one >> ${T}_variants.txt 2>&1
echo "== seeded single-step kernel, 20 steps after 5" >> ${T}_variants.txt
KOB_FAST2=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --repeats 1 2>&1 | one >> ${T}_variants.txt 2>&1
timeout 300 python scripts/dev/dev_dense_real.py make /tmp/dense_real.ckpt 30000 > ${T}_real.txt 2>&1
KOB_FAST2=0 timeout 300 python scripts/dev/dev_dense_real.py run /tmp/dense_real.ckpt 200 >> ${T}_real.txt 2>&1
timeout 300 python scripts/dev/dev_dense_real.py run /tmp/dense_real.ckpt 200 >> ${T}_real.txt 2>&1
cat ${T}_variants.txt ${T}_real.txt
KOB_FAST2=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:kob_step_fast -s 30 -c 1 -o ${T}_real -f python scripts/dev/dev_dense_real.py run /tmp/dense_real.ckpt 40 > ${T}_ncu.log 2>&1
ls -la ${T}_real.ncu-rep
