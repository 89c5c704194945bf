#!/bin/bash
# config 5 on 1 and 2 GPUs (gpurun --gpus 2) + the 2-rank gather correctness worker
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/gpus.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_worker.py > gpurun_out/mgpu_worker.log 2>&1; echo "mgpu worker exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/mgpu_worker.log | tail -8
for G in 1 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2954$G tools/bench_config5.py > gpurun_out/config5_g$G.json 2> gpurun_out/config5_g$G.err
  echo "config5 g$G exit $?"; grep '^{' gpurun_out/config5_g$G.json; grep -v "^\*\|OMP_NUM" gpurun_out/config5_g$G.err | tail -4
done
