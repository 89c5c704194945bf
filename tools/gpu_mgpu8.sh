#!/bin/bash
# N-GPU validation of the default bench command line the driver uses (gpurun --gpus N).
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
N=${1:-8}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 2000 --warmup 200 > gpurun_out/bench_g${N}.json 2> gpurun_out/bench_g${N}.err
echo "bench g$N exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_g${N}.err | tail -5; cut -c1-1500 gpurun_out/bench_g${N}.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_ref_g${N}.json 2> gpurun_out/bench_ref_g${N}.err
echo "ref g$N exit $?"; cut -c1-700 gpurun_out/bench_ref_g${N}.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 2000 --warmup 200 --gather nccl --no-ant > gpurun_out/bench_g${N}_nccl.json 2> gpurun_out/bench_g${N}_nccl.err
echo "bench g$N nccl exit $?"; cut -c1-400 gpurun_out/bench_g${N}_nccl.json
