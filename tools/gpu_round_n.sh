#!/bin/bash
# r01n final round of the session: full GPU suite, smoke, bench, extras, probe, ncu capture of the Brax kernel.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
rm -f gpurun_out/brax_parity_floor.txt
timeout 700 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value %.4e frac %.3f e2e %.4e (%.1f us) cpu %.3e cores %s ant %.3e ant_step %.3e ant_e2e %.3e (%.1f us)'%(d['value'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['ms_per_step']*1e3,d['cpu_baseline']['value'],d['cpu_baseline']['cores'],d['ant_8192']['value'],d['ant_8192']['step_api']['value'],d['ant_8192']['e2e']['value'],d['ant_8192']['e2e']['us_per_step']))
PY
timeout 300 python tools/bench_extras.py > gpurun_out/extras.json 2> gpurun_out/extras.err; echo "extras exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; echo "ref arm exit $?"; cut -c1-300 gpurun_out/bench_reference_arm.json
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:brax_step_kernel -s 3 -c 2 -f -o gpurun_out/prof_brax \
  python bench.py --steps 1000 --warmup 500 --no-cpu-baseline > gpurun_out/ncu_brax.log 2>&1; echo "ncu brax exit $?"
timeout 120 python tools/brax_pack_sweep.py > gpurun_out/pack_sweep_auto.json 2>/dev/null; echo "sweep exit $?"; cat gpurun_out/pack_sweep_auto.json
