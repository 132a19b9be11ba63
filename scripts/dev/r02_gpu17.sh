#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02t
KOB_TRACE=${T}_trace.csv timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense --no-single --no-invariance --repeats 1 > ${T}_bench.json 2> ${T}_bench.err
awk -F, 'NR>1{printf "%s %s %s|", substr($1,1,6),int($3),$5; if (NR%6==0) print ""}' ${T}_trace.csv | tail -45
