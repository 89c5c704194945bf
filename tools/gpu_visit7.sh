#!/bin/bash
# GPU visit: full parity suite + smoke + bench (default and an awkward K) + e2e probe + extras.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; echo "probe exit $?"; cat gpurun_out/e2e_probe.json; tail -3 gpurun_out/e2e_probe.err
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 7 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k7.json 2> gpurun_out/bench_k7.err; echo "bench k7 exit $?"
cut -c1-400 gpurun_out/bench_k7.json; tail -3 gpurun_out/bench_k7.err
timeout 300 python tools/bench_extras.py > gpurun_out/extras.json 2> gpurun_out/extras.err; echo "extras exit $?"; cat gpurun_out/extras.json; tail -3 gpurun_out/extras.err
