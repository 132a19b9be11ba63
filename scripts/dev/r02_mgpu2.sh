#!/bin/bash
# 2-GPU session: IPC ring parity tests, in-library ring policy on a dense field, weak-scaling bench with shard invariance.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=gpurun_out/r02m2
nvidia-smi -L > ${T}_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > ${T}_pytest.log 2>&1; echo "pytest rc=$?" >> ${T}_pytest.log
tail -5 ${T}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/mgpu_check.py --dense --nx 2000 --ny 2000 --steps 300 > ${T}_dense_policy.log 2>&1
grep "mgpu_check" ${T}_dense_policy.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err
tail -c 1500 ${T}_bench.json
tail -5 ${T}_bench.err
