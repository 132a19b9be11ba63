#!/bin/bash
# dev: sweep library variants x job heights; prints Gcell/s, roofline fraction, ms per launch
for lib in "$@"; do
  IFS=: read -r path yjs <<< "$lib"
  for yj in ${yjs//,/ }; do
    echo -n "$path YJ=$yj : "
    KOB_LIB_PATH=$path KOB_FAST_YJ=$yj python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['launch_ms'],4))"
  done
done
