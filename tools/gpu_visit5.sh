#!/bin/bash
# tests + smoke + bench (headline) + extras + Brax FMA A/B; no ncu.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("value %.4e kernel_ms %.4f frac %.3f | api us %.2f | e2e %.3e ms %.3f | ant %.3e"%(d['value'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'], d['step_api']['us_per_launch'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ant_8192']['value']))
PY
timeout 600 python tools/bench_extras.py > gpurun_out/extras.json 2> gpurun_out/extras.err; echo "extras exit $?"; cat gpurun_out/extras.json
echo "== Brax FMA A/B"
CARLB_BRAX_FMAD=1 python -m carl_b200.build --force > /dev/null 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FMAD=1 ant fused %.3e api %.3e'%(d['ant_8192']['value'], d['ant_8192']['step_api']['value']))"
timeout 600 python -m pytest tests/test_brax_parity_gpu.py -q 2>&1 | tail -4
cp gpurun_out/brax_parity_floor.txt gpurun_out/brax_parity_floor_fmad.txt 2>/dev/null
python -m carl_b200.build --force > /dev/null 2>&1
