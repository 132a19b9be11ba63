#!/bin/bash
# dev: "LIB:NP:CTA:YJ:YJB:FRAC" ... (LIB = path of a library variant or "-"); DENSE=1 also times the dense field
for cfg in "$@"; do
  IFS=: read -r lib np cta yj yjb frac <<< "$cfg"
  echo -n "LIB=$lib NP=$np CTA=$cta YJ=$yj YJB=$yjb FRAC=$frac : "
  if [ "$lib" != "-" ]; then export KOB_LIB_PATH=$lib; else unset KOB_LIB_PATH; fi
  extra="--no-dense"; [ -n "$DENSE" ] && extra=""
  KOB_FAST_NP=$np KOB_FAST_CTA=$cta KOB_FAST_YJ=$yj KOB_FAST_YJB=$yjb KOB_FAST_FRAC=$frac python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e $extra | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value'],1), round(r['frac'],4), round(r['launch_ms'],4), 'dense', round(r.get('dense_field',{}).get('value',0),1))"
done
