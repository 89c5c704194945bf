#!/bin/bash
# GPU visit: parity suite, refill-threshold sweep of the rollout kernel, e2e probe, bench (default + K=7).
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
for rf in 16 24 33 12; do
CARLB_ROLLOUT_REFILL=$rf timeout 200 python bench.py --steps 2000 --warmup 500 --no-cpu-baseline --no-ant 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('refill=$rf value %.4e kernel_ms %.4f e2e %.4e'%(d['value'], d['roofline']['kernel_ms_avg'], d['e2e']['value']))"
done | tee gpurun_out/sweep2.txt
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; echo "probe exit $?"; cat gpurun_out/e2e_probe.json; tail -3 gpurun_out/e2e_probe.err
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 7 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k7.json 2> gpurun_out/bench_k7.err; echo "bench k7 exit $?"
cut -c1-300 gpurun_out/bench_k7.json; tail -3 gpurun_out/bench_k7.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cat gpurun_out/bench_ref.json
