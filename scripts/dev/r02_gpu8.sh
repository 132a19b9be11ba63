#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02g
timeout 300 python scripts/dev/dev_dense_real.py make /tmp/dense_real.ckpt 30000 > ${T}_real.txt 2>&1
timeout 300 python scripts/dev/dev_dense_real.py run /tmp/dense_real.ckpt 200 >> ${T}_real.txt 2>&1
cat ${T}_real.txt
KOB_FAST2=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:kob_step_fast -s 30 -c 1 -o ${T}_real -f python scripts/dev/dev_dense_real.py run /tmp/dense_real.ckpt 40 > ${T}_ncu.log 2>&1
ls -la ${T}_real.ncu-rep
timeout 300 python -m pytest tests/test_strips.py tests/test_checkpoint.py tests/test_gpu_parity.py -m gpu -x -q -k "strip or checkpoint or roundtrip" 2>&1 | tail -3
