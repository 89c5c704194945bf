#!/bin/bash
# r02e (1 GPU): cost of the in-kernel gather machinery without NVLink; full GPU suite incl. the new full-size parity
# tests; e2e bench line; ncu of the deferred-push rollout kernel.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for m in none pipelined sync; do timeout 120 python tools/gather_probe_1gpu.py $m; done
CARLB_GATHER_DEBUG=15 timeout 120 python tools/gather_probe_1gpu.py pipelined
CARLB_ROLLOUT_BLOCK=128 timeout 120 python tools/gather_probe_1gpu.py none
CARLB_ROLLOUT_BLOCK=128 timeout 120 python tools/gather_probe_1gpu.py pipelined
CARLB_ROLLOUT_BLOCK=32 timeout 120 python tools/gather_probe_1gpu.py none
CARLB_ROLLOUT_BLOCK=32 timeout 120 python tools/gather_probe_1gpu.py pipelined
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 10 -c 1 -f -o gpurun_out/prof_rollout_deferred \
  python tools/gather_probe_1gpu.py pipelined --ncu > gpurun_out/ncu_deferred.log 2>&1; echo "ncu deferred exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 10 -c 1 -f -o gpurun_out/prof_rollout_plain \
  python tools/gather_probe_1gpu.py none --ncu > gpurun_out/ncu_plain.log 2>&1; echo "ncu plain exit $?"
rm -f gpurun_out/brax_parity_floor.txt gpurun_out/done_mask_counts.json
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
cat gpurun_out/done_mask_counts.json
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 exit $?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e (%.1f us) async %s ant %.4e ant_fma %s'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']*1e3, d.get('e2e_sync'), d['ant_8192']['value'], d['ant_8192'].get('value_fma')))
PY
