#!/bin/bash
# dev: two-step kernel sweep, "LIB:LOCK:YJ:YJB"
for cfg in "$@"; do
  IFS=: read -r lib lock yj yjb <<< "$cfg"
  echo -n "F2 LIB=$lib LOCK=$lock YJ=$yj YJB=$yjb : "
  if [ "$lib" != "-" ]; then export KOB_LIB_PATH=$lib; else unset KOB_LIB_PATH; fi
  KOB_FAST2=1 KOB_FAST2_LOCK=$lock KOB_FAST2_YJ=$yj KOB_FAST2_YJB=$yjb python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-dense | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['gpu_launches'])"
done
