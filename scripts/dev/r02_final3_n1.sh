#!/bin/bash
# Final check of the committed state + fresh ncu captures (plain far -> general order: a profiler serialises kernels).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02zy
timeout 1200 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -3 ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench.json 2> ${T}_bench.err; tail -2 ${T}_bench.err
cap() {  # name, kernel regex, skip, env..., then bench args
  local name=$1 rx=$2 skip=$3; shift 3
  env "$@" timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o /tmp/${name} -f python bench.py $BARGS > ${T}_ncu_${name}.log 2>&1
  python scripts/ncu_summary.py /tmp/${name}.ncu-rep > ${T}_${name}_summary.md 2>&1
  rm -f /tmp/${name}.ncu-rep
}
BARGS="--steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap seeded kob_step_fast 12 KOB_FAST2=0
BARGS="--field dense --steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap dense kob_step_fast 35 KOB_FAST2=0
BARGS="--steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap far2 kob_far2 12 KOB_FAST2=1 KOB_FAST2_CONC=0
BARGS="--steps 1 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1" cap general kob_step_fast2 12 KOB_FAST2=1 KOB_FAST2_CONC=0
head -12 ${T}_far2_summary.md
du -sh gpurun_out
