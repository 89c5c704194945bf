#!/bin/bash
# GPU visit: parity suite + bench + reference arm.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 2000 --warmup 200 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cut -c1-1000 gpurun_out/bench_ref.json
timeout 300 python tools/bench_extras.py > gpurun_out/extras.json 2> gpurun_out/extras.err; echo "extras exit $?"; cat gpurun_out/extras.json; tail -3 gpurun_out/extras.err
