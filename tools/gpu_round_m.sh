#!/bin/bash
# r01m profile round: launch list of the timed region + full captures (rollout, brax, host-path step).
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2000 --warmup 200 --fused-only > gpurun_out/ncu_launch.log 2>&1; echo "ncu list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 2 -f -o gpurun_out/prof_rollout \
  python bench.py --steps 2000 --warmup 200 --fused-only > gpurun_out/ncu_rollout.log 2>&1; echo "ncu rollout exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:brax_step_kernel -s 3 -c 2 -f -o gpurun_out/prof_brax \
  python bench.py --steps 1000 --warmup 500 --no-cpu-baseline > gpurun_out/ncu_brax.log 2>&1; echo "ncu brax exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_checked_kernel -s 10 -c 2 -f -o gpurun_out/prof_step_checked \
  python bench.py --steps 1000 --warmup 500 --no-cpu-baseline --no-ant > gpurun_out/ncu_stepc.log 2>&1; echo "ncu step_checked exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^step_kernel -s 10 -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 1000 --warmup 500 --no-cpu-baseline --no-ant > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
ls -la gpurun_out | head -40
