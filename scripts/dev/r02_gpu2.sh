#!/bin/bash
# ncu captures of the dense-field launch: per-row-vote tier (KOB_FAST_DENSE=0) and straight-line tier (=2)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for dm in 0 2; do
  KOB_FAST2=0 KOB_FAST_DENSE=$dm timeout 600 ncu --set full --import-source on --clock-control none -k regex:kob_step_fast -s 35 -c 1 -o gpurun_out/r02b_dense$dm -f python bench.py --field dense --steps 1 --warmup 3 --no-cpu --no-e2e --no-dense > gpurun_out/r02b_ncu$dm.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
