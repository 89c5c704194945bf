#!/bin/bash
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
for pack in 3 1; do
CARLB_BRAX_PACK=$pack timeout 300 python bench.py --steps 400 --warmup 40 --no-cpu-baseline 2>gpurun_out/bench_pack$pack.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PACK=$pack value %.4e kernel_ms %.4f | ant fused %.3e api %.3e'%(d['value'], d['roofline']['kernel_ms_avg'], d['ant_8192']['value'], d['ant_8192']['step_api']['value']))"
done
timeout 600 python tools/bench_extras.py > gpurun_out/extras.json 2> gpurun_out/extras.err; echo "extras exit $?"; cat gpurun_out/extras.json
