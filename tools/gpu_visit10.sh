#!/bin/bash
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value %.4e frac %.3f e2e %.4e (%.1f us) cpu %.3e cores %s'%(d['value'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['ms_per_step']*1e3,d['cpu_baseline']['value'],d['cpu_baseline']['cores']))
PY
tail -3 gpurun_out/bench.err
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; echo "probe exit $?"; cat gpurun_out/e2e_probe.json
