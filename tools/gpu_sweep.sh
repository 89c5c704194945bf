#!/bin/bash
# Parameter sweep of the fused rollout kernel (block size x refill threshold); tests first.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for blk in 64 32 128; do for rf in 8 4 16; do
CARLB_ROLLOUT_BLOCK=$blk CARLB_ROLLOUT_REFILL=$rf timeout 200 python bench.py --steps 2000 --warmup 500 --no-cpu-baseline --no-ant 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('block=$blk refill=$rf value %.4e kernel_ms %.4f'%(d['value'], d['roofline']['kernel_ms_avg']))"
done; done | tee gpurun_out/sweep.txt
