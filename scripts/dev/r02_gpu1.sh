#!/bin/bash
# Round-2 GPU session 1: parity tests, seeded + dense bench, occupancy variants on the dense field, ncu of the dense launch.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
cat gpurun_out/r02a_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('seeded', round(d['value'],1), 'dense', r.get('dense_field'))"
for v in crystalgrowth_b200/variants/*.so; do
  echo "== $v" >> gpurun_out/r02a_variants.txt
  KOB_LIB_PATH=$PWD/$v timeout 200 python bench.py --field dense --steps 3 --warmup 3 --no-cpu --no-e2e --no-dense 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), d['roofline']['launch_ms'])" >> gpurun_out/r02a_variants.txt 2>&1
done
for dm in 0 2; do
  echo "== KOB_FAST_DENSE=$dm" >> gpurun_out/r02a_variants.txt
  KOB_FAST_DENSE=$dm timeout 200 python bench.py --field dense --steps 3 --warmup 3 --no-cpu --no-e2e --no-dense 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), d['roofline']['launch_ms'])" >> gpurun_out/r02a_variants.txt 2>&1
done
echo "== KOB_FAST_CTA=0 (per-warp jobs)" >> gpurun_out/r02a_variants.txt
KOB_FAST_CTA=0 timeout 200 python bench.py --field dense --steps 3 --warmup 3 --no-cpu --no-e2e --no-dense 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), d['roofline']['launch_ms'])" >> gpurun_out/r02a_variants.txt 2>&1
cat gpurun_out/r02a_variants.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kob_step_fast -s 35 -c 1 -o gpurun_out/r02a_dense -f python bench.py --field dense --steps 1 --warmup 3 --no-cpu --no-e2e --no-dense > gpurun_out/r02a_ncu.log 2>&1
ls -la gpurun_out/r02a_dense.ncu-rep
