#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): 2-rank gather test + scaling bench lines.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
if [ "$2" != "benchonly" ]; then
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_worker.py > gpurun_out/mgpu_worker.log 2>&1; echo "mgpu worker exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/mgpu_worker.log | tail -40
fi
if [ "$2" == "testonly" ]; then exit 0; fi
echo "== single rank under torchrun (e2e check)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 1 --steps 500 --warmup 50 --no-cpu-baseline --no-ant 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('torchrun-1 e2e', d['e2e'])"
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
