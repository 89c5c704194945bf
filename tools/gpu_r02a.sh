#!/bin/bash
# r02a: (1) can the third-party packages holding the reference arithmetic be installed on the GPU box?
# (2) baseline GPU suite; (3) per-launch cost sweep of the rollout kernel; (4) ncu at the driver's T = 20;
# (5) compute-sanitizer over the checked-step / fused tests.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
{ echo "== pip install gymnasium==0.29.1 jax[cpu] brax==0.12.1 (GPU box, $(date -u))";
  timeout 120 python -m pip install --no-input gymnasium==0.29.1 "jax[cpu]" brax==0.12.1 2>&1 | tail -15;
  echo "exit ${PIPESTATUS[0]}";
  echo "== offline wheelhouse attempt";
  timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse gymnasium==0.29.1 jax brax==0.12.1 2>&1 | tail -8;
  echo "exit ${PIPESTATUS[0]}";
  python -c "import gymnasium" 2>&1 | tail -1; python -c "import jax" 2>&1 | tail -1; python -c "import brax" 2>&1 | tail -1; } > gpurun_out/pip_install_attempt.txt 2>&1
tail -5 gpurun_out/pip_install_attempt.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/rollout_sweep.py > gpurun_out/rollout_sweep.json 2> gpurun_out/rollout_sweep.err; echo "sweep exit $?"; cat gpurun_out/rollout_sweep.json
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; echo "bench k20 exit $?"
rm -f gpurun_out/prof_*.ncu-rep
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 4 -c 2 -f -o gpurun_out/prof_rollout_t20 \
  python bench.py --steps 20 --warmup 100 --fused-only > gpurun_out/ncu_rollout_t20.log 2>&1; echo "ncu t20 exit $?"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_env_api_gpu.py -m gpu -q -x -k "checked or host or step" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/sanitizer_memcheck.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python - <<'PY'
import os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try: print(open('/sys/fs/cgroup/cpu.max').read())
except Exception as e: print(e)
PY
