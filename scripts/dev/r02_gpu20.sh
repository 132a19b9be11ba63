#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=crystalgrowth_b200/variants/libkobayashi_cuda_loop2.so
python scripts/dev/dev_dense_real.py make /tmp/real30k.kobck 30000 | tail -1
for lib in "" $V; do
  echo "== lib=${lib:-default}"
  KOB_LIB_PATH=$lib python scripts/dev/dev_dense_real.py run /tmp/real30k.kobck 400
  KOB_LIB_PATH=$lib KOB_FAST2=0 python bench.py --field dense --steps 5 --warmup 3 --no-cpu --no-e2e --no-dense --no-single --repeats 1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dense bench', round(d['value'],1), d['roofline']['launch_ms'])"
  KOB_LIB_PATH=$lib python scripts/long_run.py --n 8192 --nuclei 64 --steps 40000 --chunk 8000 | tail -4
done
