#!/bin/bash
# A/B: PDL on/off for the per-step API and the host path.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for v in 1 0; do
  CARLB_ZEROCOPY=$v timeout 300 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-ant 2>gpurun_out/ab_$v.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ZEROCOPY=$v value %.3e e2e_ms %.3f api_us %.2f cold_us %.2f'%(d['value'], d['e2e']['ms_per_step'], d['step_api']['us_per_launch'], d['step_api']['cold_l2_us_per_launch']))"
done
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
