#!/bin/bash
# e2e work: in-kernel action check + undo (tests), probe of the host-buffer step, bench line.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_env_api_gpu.py tests/test_classic_parity_gpu.py tests/test_adapters.py tests/test_spaces_and_bench_contract.py -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; echo "probe exit $?"; python -m json.tool gpurun_out/e2e_probe.json; tail -3 gpurun_out/e2e_probe.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value %.4e frac %.3f e2e %.4e (%.1f us) cpu %.3e cores %s ant %.3e step_api %.3e'%(d['value'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['ms_per_step']*1e3,d['cpu_baseline']['value'],d['cpu_baseline']['cores'],d['ant_8192']['value'],d['step_api']['value']))
PY
tail -3 gpurun_out/bench.err
