#!/bin/bash
# Round-2 final single-GPU evidence run (after the early general pass): tests, smoke, default bench, launch trace (CUDA events, no
# profiler), ncu launch list with the plain far -> general order (a profiler serialises kernels: the early general pass would only
# time out), sanitizer.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02zz
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > ${T}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -4 ${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${T}_smoke.log 2>&1; tail -2 ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench.json 2> ${T}_bench.err; tail -2 ${T}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > ${T}_bench_reference.json 2>> ${T}_bench.err
KOB_TRACE=${T}_trace_pairs.csv timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 1 > /dev/null 2>&1
KOB_FAST2_CONC=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --repeats 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${T}_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 > /dev/null 2>&1
timeout 900 python scripts/long_run.py > ${T}_long_run_4096.md 2>&1; tail -3 ${T}_long_run_4096.md
bash scripts/sanitize.sh > ${T}_sanitizer.txt 2>&1; cat ${T}_sanitizer.txt
du -sh gpurun_out
