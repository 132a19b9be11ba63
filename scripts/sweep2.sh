#!/bin/bash
# dev: sweep job-shape knobs "YJ:YJB:FRAC" ...
for cfg in "$@"; do
  IFS=: read -r yj yjb frac <<< "$cfg"
  echo -n "YJ=$yj YJB=$yjb FRAC=$frac : "
  KOB_FAST_YJ=$yj KOB_FAST_YJB=$yjb KOB_FAST_FRAC=$frac python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['launch_ms'],4))"
done
