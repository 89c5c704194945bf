#!/bin/bash
# ONE parameterised GPU visit script (replaces the per-visit scripts of round 1).
#   tools/gpu_visit.sh tests            GPU test-suite + smoke
#   tools/gpu_visit.sh bench            the driver's bench command (+ default, + reference arm)
#   tools/gpu_visit.sh ncu              launch list + ncu --set full of the hot kernels (1 GPU only)
#   tools/gpu_visit.sh sanitizer        compute-sanitizer memcheck / racecheck / synccheck over the fused-gather and host-step paths
#   tools/gpu_visit.sh mgpu             multi-GPU correctness worker on NG ranks (gpurun --gpus NG)
#   tools/gpu_visit.sh scale            bench.py at N in NLIST (default "1 2 4 8", capped at NG)
#   tools/gpu_visit.sh extras           secondary configs and every Brax body on one GPU (tools/bench_extras.py)
# Several stages may be given; outputs go to gpurun_out/ (copy what should be kept into profiles/).
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=${NG:-1}
K=${K:-20}; W=${W:-5}
for stage in "$@"; do
case $stage in
tests)
  rm -f gpurun_out/brax_parity_floor.txt gpurun_out/done_mask_counts.json
  timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
  ;;
bench)
  timeout 600 python bench.py --gpus 1 --steps $K --warmup $W > gpurun_out/bench_k$K.json 2> gpurun_out/bench_k$K.err; echo "bench k$K exit $?"; tail -2 gpurun_out/bench_k$K.err
  timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?"
  timeout 600 python bench.py --impl reference --gpus 1 --steps $K --warmup $W > gpurun_out/bench_reference_arm.json 2>/dev/null; echo "reference arm exit $?"
  python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'unparsable',e); continue
    if d.get('impl')=='reference': print(f,'reference value %.4e cores %s B2 %s'%(d['value'],d['cpu_baseline']['cores'],d.get('python_scalar_all_cores_B2',{}).get('value'))); continue
    print(f,'value %.4e frac %.3f kernel_us %.2f | e2e %.4e (%.1f us) sync %.4e | step_api %.2f us | f64 %.4e | ant %.4e fma %.4e | cpu %s'%(d['value'],d['roofline']['frac'],d['roofline']['kernel_ms_avg']*1e3,d['e2e']['value'],d['e2e']['ms_per_step']*1e3,d['e2e_sync']['value'],d['step_api']['us_per_launch'],d['value_f64']['value'],d['ant_8192']['value'],d['ant_8192']['value_fma'],d.get('cpu_baseline',{}).get('value')))
PY
  ;;
ncu)
  rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps $K --warmup $W --fused-only > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 60 -c 2 -f -o gpurun_out/prof_rollout \
    python bench.py --steps $K --warmup $W --fused-only > gpurun_out/ncu_rollout.log 2>&1; echo "ncu rollout exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:brax_step_kernel -s 4 -c 2 -f -o gpurun_out/prof_brax \
    python tools/ncu_targets.py brax > gpurun_out/ncu_brax.log 2>&1; echo "ncu brax exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:brax_step_kernel -s 4 -c 2 -f -o gpurun_out/prof_brax_fma \
    python tools/ncu_targets.py brax_fma > gpurun_out/ncu_brax_fma.log 2>&1; echo "ncu brax_fma exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 2 -f -o gpurun_out/prof_step \
    python tools/ncu_targets.py step > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_checked_kernel -s 10 -c 2 -f -o gpurun_out/prof_step_checked \
    python tools/ncu_targets.py step_host > gpurun_out/ncu_step_checked.log 2>&1; echo "ncu step_checked exit $?"
  ls -la gpurun_out/*.ncu-rep
  # gpurun copies back at most 64 MiB: summarise every report here, keep only the two that are read at source level
  CARLB_PROF_DIR=gpurun_out/summary python tools/summarize_ncu.py ${TAG:-visit}; echo "summaries exit $?"
  rm -f gpurun_out/prof_step.ncu-rep gpurun_out/prof_step_checked.ncu-rep gpurun_out/prof_brax.ncu-rep
  ;;
sanitizer)
  for tool in memcheck racecheck synccheck; do
    for mode in pipelined sync; do
      timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/gather_probe_1gpu.py $mode --ncu > gpurun_out/sanitizer_${tool}_gather_${mode}.log 2>&1
      echo "sanitizer $tool gather/$mode exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_gather_${mode}.log | tail -1)"
    done
  done
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_env_api_gpu.py -m gpu -q -x -k "split_batch or checked or host" > gpurun_out/sanitizer_memcheck_host_step.log 2>&1
  echo "sanitizer memcheck host-step exit $? : $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_memcheck_host_step.log | tail -1)"
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_brax_parity_gpu.py -m gpu -q -x -k "single_env_step_matches and ant and applied" > gpurun_out/sanitizer_racecheck_brax.log 2>&1
  echo "sanitizer racecheck brax exit $? : $(grep 'RACECHECK SUMMARY' gpurun_out/sanitizer_racecheck_brax.log | tail -1)"
  ;;
mgpu)
  CARLB_MGPU_TIMEOUT=500 timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py > gpurun_out/mgpu_worker_${NG}gpu.log 2>&1
  echo "mgpu worker ($NG ranks) exit $?; MGPU_OK x $(grep -c MGPU_OK gpurun_out/mgpu_worker_${NG}gpu.log)"; tail -3 gpurun_out/mgpu_worker_${NG}gpu.log
  ;;
scale)
  NLIST="${NLIST:-1 2 4 8}" NG=$NG bash tools/gpu_scale.sh
  ;;
extras)
  timeout 400 python tools/bench_extras.py > gpurun_out/extras.json 2> gpurun_out/extras.err; echo "extras exit $?"; tail -2 gpurun_out/extras.err
  ;;
*) echo "unknown stage $stage";;
esac
done
