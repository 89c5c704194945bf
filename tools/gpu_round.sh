#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list + full captures. Run under gpurun.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 2000 --warmup 200 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2000 --warmup 200 --fused-only > gpurun_out/ncu_launch.log 2>&1; echo "ncu list exit $?"
echo "== ncu full: rollout kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 2 -f -o gpurun_out/prof_rollout \
  python bench.py --steps 2000 --warmup 200 --fused-only > gpurun_out/ncu_rollout.log 2>&1; echo "ncu rollout exit $?"
echo "== ncu full: brax step kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:brax_step_kernel -s 3 -c 2 -f -o gpurun_out/prof_brax \
  python bench.py --steps 1000 --warmup 500 --no-cpu-baseline > gpurun_out/ncu_brax.log 2>&1; echo "ncu brax exit $?"
echo "== ncu full: step kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^step_kernel -s 10 -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 1000 --warmup 500 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
fi
ls -la gpurun_out
