#!/bin/bash
# 8-GPU session: bitwise shard invariance on a ragged torus, weak-scaling bench (with invariance pre-check and the 65536^2 strong secondary).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
T=gpurun_out/r02m${N}
nvidia-smi -L > ${T}_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py --nx 3000 --ny 2063 --steps 120 > ${T}_check.log 2>&1
grep "mgpu_check" ${T}_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/mgpu_check.py --dense --nx 3000 --ny 4000 --steps 200 > ${T}_check_dense.log 2>&1
grep "mgpu_check" ${T}_check_dense.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err
grep -c metric ${T}_bench.json; tail -3 ${T}_bench.err
