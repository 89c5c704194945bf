#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): 2-rank gather test + scaling bench lines.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -x > gpurun_out/pytest_mgpu.log 2>&1; echo "mgpu pytest exit $?"; tail -25 gpurun_out/pytest_mgpu.log
for G in 1 $N; do
  if [ "$G" == "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 2000 --warmup 200 --no-cpu-baseline --no-ant > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err
  else
    for mode in fused nccl; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $G --steps 2000 --warmup 200 --gather $mode > gpurun_out/bench_g${G}_$mode.json 2> gpurun_out/bench_g${G}_$mode.err
      echo "bench g$G $mode exit $?"; tail -3 gpurun_out/bench_g${G}_$mode.err; cat gpurun_out/bench_g${G}_$mode.json | cut -c1-1200
    done
  fi
done
cat gpurun_out/bench_g1.json | cut -c1-600
