#!/bin/bash
# One ncu --set full capture of the fused rollout kernel (bench command line, profiler numbers are not bench values).
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_rollout \
  python bench.py --steps 1000 --warmup 500 --no-cpu-baseline --no-ant > gpurun_out/ncu_rollout.log 2>&1; echo "ncu rollout exit $?"
tail -3 gpurun_out/ncu_rollout.log
ls -la gpurun_out/prof_rollout.ncu-rep
