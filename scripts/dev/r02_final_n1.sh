#!/bin/bash
# Round-2 final single-GPU evidence run: tests, smoke, default bench, ncu launch list + full captures (summarised on the box:
# gpurun brings back at most 64 MiB), long runs, sanitizer.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02z
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > ${T}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -4 ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${T}_smoke.log 2>&1; tail -2 ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench.json 2> ${T}_bench.err; tail -2 ${T}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > ${T}_bench_reference.json 2>> ${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --repeats 1 > /dev/null 2>&1
cap() {  # name, kernel regex, skip, env..., then bench args
  local name=$1 rx=$2 skip=$3; shift 3
  env "$@" timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o /tmp/${name} -f python bench.py $BARGS > ${T}_ncu_${name}.log 2>&1
  python scripts/ncu_summary.py /tmp/${name}.ncu-rep > ${T}_${name}_summary.md 2>&1
  ncu -i /tmp/${name}.ncu-rep --page source --csv 2>/dev/null | gzip > ${T}_${name}_source.csv.gz
  rm -f /tmp/${name}.ncu-rep
}
BARGS="--steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap seeded kob_step_fast 12 KOB_FAST2=0
BARGS="--field dense --steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap dense kob_step_fast 35 KOB_FAST2=0
BARGS="--steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap far2 kob_far2 12 KOB_FAST2=1
BARGS="--steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap general kob_step_fast2 12 KOB_FAST2=1
cuobjdump -xelf all crystalgrowth_b200/libkobayashi_cuda.so > /dev/null 2>&1 && nvdisasm -g -c kob_api.sm_100a.cubin 2>/dev/null | gzip > ${T}_lib_lineinfo.sass.gz; rm -f kob_api.sm_100a.cubin
timeout 900 python scripts/long_run.py > ${T}_long_run_4096.md 2>&1; tail -3 ${T}_long_run_4096.md
timeout 900 python scripts/long_run.py --n 8192 --nuclei 64 --steps 40000 --chunk 4000 > ${T}_long_run_8192.md 2>&1; tail -3 ${T}_long_run_8192.md
bash scripts/sanitize.sh > ${T}_sanitizer.txt 2>&1; cat ${T}_sanitizer.txt
du -sh gpurun_out
